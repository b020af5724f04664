"""AttentiveStatsPooling at the cfg2 audio shape (B 256, T 250, D 768, bf16): per-kernel-family CUDA-event times from the
library profiler (the scorer GEMM is reported separately from the pooling kernels proper), with algorithmic GB/s."""
import sys

import torch

sys.path.insert(0, ".")
import mmser_b200  # noqa: E402
from mmser_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
B, T, D = 256, 250, 768
pool = mmser_b200.models.AttentiveStatsPooling(D).to(dev)
mask = torch.ones(B, T, device=dev)
mask[:, 200:] = 0
xs = [torch.randn(B, T, D, device=dev).bfloat16().requires_grad_(True) for _ in range(3)]   # rotate: defeat L2 residency
up = torch.randn(B, 2 * D, device=dev).bfloat16()
for it in range(3):
    pool(xs[it % 3], mask).backward(up)
torch.cuda.synchronize()
L.prof_enable(True)
iters = 12
for it in range(iters):
    pool(xs[it % 3], mask).backward(up)
rep = L.prof_report()
L.prof_enable(False)
x_bytes = B * T * D * 2
alg = {"asp_fwd": x_bytes + B * T * 128 * 2, "asp_bwd": 2 * x_bytes + B * T * 128 * 2 * 2}
for k, v in sorted(rep.items()):
    us = v["ms"] / iters * 1e3
    extra = f"  {alg[k] / us / 1e3:7.1f} GB/s algorithmic ({alg[k] / 1e6:.0f} MB)" if k in alg else ""
    print(f"{k:34s} {v['launches'] / iters:4.1f} launches/iter {us:8.1f} us/iter{extra}")
