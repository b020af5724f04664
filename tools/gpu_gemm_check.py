"""GPU check of the two GEMM tiers (SIMT fp32, tcgen05 bf16) against torch.matmul.  Run under gpurun.

Prints one line per case and never stops at the first failure, so a single GPU call surfaces everything.
"""
import sys
import time

import torch

sys.path.insert(0, ".")
import mmser_b200  # noqa: E402
from mmser_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
fails = 0


def ref_gemm(a, b, a_trans, b_trans, bias, act, residual, gate, gate_mode, alpha):
    A = a.float().t() if a_trans else a.float()
    B = b.float() if b_trans else b.float().t()
    v = alpha * (A.double() @ B.double()).float()
    if bias is not None:
        v = v + bias
    if act == L.ACT_RELU:
        v = torch.relu(v)
    elif act == L.ACT_TANH:
        v = torch.tanh(v)
    if gate_mode == L.GATE_RELU:
        v = v * (gate.float() > 0)
    elif gate_mode == L.GATE_TANH:
        v = v * (1 - gate.float() ** 2)
    if residual is not None:
        v = v + residual.float()
    return v


def run_case(name, dtype, M, N, K, a_trans=False, b_trans=False, bias=False, act=0, residual=None, gate_mode=0,
             out_dtype=None, splits=0, alpha=1.0, tol=None):
    global fails
    a = torch.randn((K, M) if a_trans else (M, K), device=dev).to(dtype)
    b = torch.randn((K, N) if b_trans else (N, K), device=dev).to(dtype)
    bias_t = torch.randn(N, device=dev) if bias else None
    out_dtype = out_dtype or dtype
    res = torch.randn(M, N, device=dev).to(residual) if residual is not None else None
    gate = torch.randn(M, N, device=dev).to(dtype).clamp(-0.9, 0.9) if gate_mode else None
    try:
        out = L.gemm(a, b, a_trans=a_trans, b_trans=b_trans, bias=bias_t, act=act, residual=res, gate=gate,
                     gate_mode=gate_mode, out_dtype=out_dtype, splits=splits, alpha=alpha)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"[FAIL] {name}: exception {e}")
        fails += 1
        return
    ref = ref_gemm(a, b, a_trans, b_trans, bias_t, act, res, gate, gate_mode, alpha)
    err = (out.float() - ref).abs()
    scale = ref.abs().max().item() + 1e-6
    rel = err.max().item() / scale
    if tol is None:
        tol = 2e-5 if (dtype == torch.float32) else (1e-2 if out_dtype == torch.bfloat16 else 2e-3)
    ok = rel <= tol and torch.isfinite(out.float()).all().item()
    print(f"[{'ok' if ok else 'FAIL'}] {name}: M={M} N={N} K={K} tA={int(a_trans)} tB={int(b_trans)} "
          f"max_rel={rel:.3e} (tol {tol:.1e})")
    if not ok:
        fails += 1
        # coarse error map: 32-row x 32-col blocks
        Mb, Nb = min(M, 256), min(N, 256)
        e = err[:Mb, :Nb]
        rows = []
        for i in range(0, Mb, 32):
            rows.append(" ".join(f"{e[i:i+32, j:j+32].max().item()/scale:7.1e}" for j in range(0, Nb, 32)))
        print("   block error map (first 256x256):\n   " + "\n   ".join(rows))
        bad = (err / scale > tol).nonzero()[:6]
        for idx in bad:
            i, j = idx.tolist()
            print(f"   [{i},{j}] got {out[i, j].item():.5f} ref {ref[i, j].item():.5f}")


f32, bf = torch.float32, torch.bfloat16
print("== SIMT fp32 ==")
for ta in (False, True):
    for tb in (False, True):
        run_case("simt plain", f32, 300, 192, 100, ta, tb)
run_case("simt bias+relu+res", f32, 257, 130, 77, bias=True, act=L.ACT_RELU, residual=f32)
run_case("simt tanh", f32, 129, 128, 768, bias=True, act=L.ACT_TANH, tol=1e-4)   # tanhf vs torch.tanh: a few ulp
run_case("simt gate relu", f32, 200, 256, 96, gate_mode=L.GATE_RELU)
run_case("simt gate tanh", f32, 200, 128, 96, gate_mode=L.GATE_TANH)
run_case("simt dW splitK", f32, 256, 768, 20000, True, True, tol=1e-4)
run_case("simt small N", f32, 256, 4, 256, bias=True)

print("== tcgen05 bf16 ==")
run_case("tc NN bn256", bf, 1000, 256, 768, bias=True, act=L.ACT_RELU)
run_case("tc NN bn256 f32out", bf, 1000, 256, 768, bias=True, out_dtype=f32)
run_case("tc NN N768 res", bf, 1000, 768, 256, bias=True, residual=bf)
run_case("tc NN bn128 tanh", bf, 640, 128, 768, bias=True, act=L.ACT_TANH)
run_case("tc NN N512 K512 f32res", bf, 256, 512, 512, bias=True, residual=f32, out_dtype=f32)
run_case("tc NN Ktail", bf, 384, 256, 200)
run_case("tc NN long", bf, 20000, 768, 768, bias=True)
run_case("tc dX (B MN-major) bn256", bf, 1000, 768, 256, b_trans=True)
run_case("tc dX (B MN-major) bn128", bf, 1000, 128, 256, b_trans=True)
run_case("tc dX gate relu", bf, 1000, 256, 768, b_trans=True, gate_mode=L.GATE_RELU)
run_case("tc dW (both MN-major) split", bf, 256, 768, 5000, True, True, out_dtype=f32)
run_case("tc dW nosplit", bf, 256, 768, 5000, True, True, out_dtype=f32, splits=1)
run_case("tc dW bn128", bf, 768, 128, 4096, True, True, out_dtype=f32)
run_case("tc A MN-major only", bf, 256, 256, 1024, True, False, out_dtype=f32)


def rowsum_case(name, dtype, M, N, K, splits=0, b_trans=True):
    """dW-type GEMM with the bias gradient (row sums of op(A)) as a by-product."""
    global fails
    a = torch.randn(K, M, device=dev).to(dtype)
    b = (torch.randn(K, N, device=dev) if b_trans else torch.randn(N, K, device=dev)).to(dtype)
    rs = torch.full((M,), 7.0, device=dev)          # must be overwritten, not accumulated
    out = L.gemm(a, b, a_trans=True, b_trans=b_trans, out_dtype=f32, splits=splits, rowsum=rs)
    torch.cuda.synchronize()
    ref = a.double().t() @ (b.double() if b_trans else b.double().t())
    ref_rs = a.double().sum(0)
    e1 = ((out.double() - ref).abs().max() / (ref.abs().max() + 1e-9)).item()
    e2 = ((rs.double() - ref_rs).abs().max() / (ref_rs.abs().max() + 1e-9)).item()
    ok = e1 < 2e-3 and e2 < 1e-4
    print(f"[{'ok' if ok else 'FAIL'}] {name}: M={M} N={N} K={K} splits={splits} dW err={e1:.2e} rowsum err={e2:.2e}")
    if not ok:
        fails += 1
        bad = ((rs.double() - ref_rs).abs() / (ref_rs.abs().max() + 1e-9) > 1e-4).nonzero().flatten()[:8].tolist()
        print("   bad rows:", bad, "got", rs[bad].tolist(), "ref", ref_rs[bad].tolist())


print("== dW GEMM + fused bias gradient ==")
rowsum_case("tc rowsum bn256 split", bf, 768, 768, 20000)
rowsum_case("tc rowsum bn256 nosplit", bf, 256, 768, 5000, splits=1)
rowsum_case("tc rowsum bn128", bf, 768, 128, 4096)
rowsum_case("tc rowsum M tail", bf, 200, 256, 3000)
rowsum_case("tc rowsum tiny", bf, 512, 512, 256)
rowsum_case("tc rowsum B K-major (fallback colsum)", bf, 256, 256, 1024, b_trans=False)
rowsum_case("simt rowsum (fallback colsum)", f32, 192, 128, 700)

print("== timing ==")


def bench(M, N, K, iters=20, **kw):
    a = torch.randn(M, K, device=dev).to(bf)
    b = torch.randn(N, K, device=dev).to(bf)
    out = torch.empty(M, N, device=dev, dtype=bf)
    for _ in range(3):
        L.gemm(a, b, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        L.gemm(a, b, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    t0 = time.perf_counter()
    for _ in range(iters):
        torch.matmul(a, b.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(a, b.t())
    e1.record()
    torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K}: ours {ms*1e3:.1f} us {fl/ms/1e9:.1f} TFLOP/s | cuBLAS {ms_ref*1e3:.1f} us "
          f"{fl/ms_ref/1e9:.1f} TFLOP/s")


try:
    bench(64000, 256, 768)
    bench(64000, 768, 768)
    bench(64000, 768, 256)
    bench(16384, 768, 768)
    bench(256, 512, 512)
    bench(8192, 8192, 8192, iters=5)
except Exception as e:  # noqa: BLE001
    print("bench failed:", e)
    fails += 1

print(f"TOTAL FAILS {fails}")
sys.exit(1 if fails else 0)
