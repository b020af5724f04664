"""Runs the attention core alone a few times (for ncu): python tools/attn_profile.py B Tq Tk [impl] [bwd] [p_drop] [bits]"""
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from tools import attn_check as A  # noqa: E402

B, Tq, Tk = (int(x) for x in sys.argv[1:4])
impl = int(sys.argv[4]) if len(sys.argv) > 4 else 2
bwd = len(sys.argv) > 5 and sys.argv[5] == "bwd"
p = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
keepbits = len(sys.argv) > 7 and sys.argv[7] == "bits"
g = torch.Generator().manual_seed(1)
qb = torch.randn(B * Tq, A.S3, generator=g).to(A.dev).bfloat16()
kvb = torch.randn(B * Tk, A.S3, generator=g).to(A.dev).bfloat16()
lens = torch.randint(max(1, Tk // 4), Tk + 1, (B,), generator=g)
kmask = (torch.arange(Tk)[None] < lens[:, None]).float().to(A.dev)
dO = torch.randn(B * Tq, 256, generator=g).to(A.dev).bfloat16() if bwd else None
sd = torch.tensor([77], dtype=torch.int64, device=A.dev)
for _ in range(4):
    A.run(impl, B, 8, Tq, Tk, qb, kvb[:, 256:], kvb[:, 512:], kmask, p, sd, 1, dO, keepbits=keepbits)
print("done")
