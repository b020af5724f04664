"""Micro-benchmark of the tcgen05 GEMM over the shapes one cfg2 training step launches (run under gpurun).

  python tools/gemm_bench.py                 # all cases, CUDA-event timing, back-to-back launches
  python tools/gemm_bench.py --case 3 --iters 3     # one case (for ncu)

Each line: achieved TFLOP/s and the GB/s of the algorithmic operand + output bytes, next to torch.matmul
(cuBLAS, no epilogue) on the same shape as a sanity reference.
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
import mmser_b200  # noqa: E402,F401
from mmser_b200 import _lib as L  # noqa: E402

bf, f32 = torch.bfloat16, torch.float32
dev = torch.device("cuda:0")

# (name, M, N, K, a_trans, b_trans, bias, act, residual dtype, gate_mode, out dtype)
CASES = [
    ("adapter1 fwd relu",        64000, 256, 768, 0, 0, 1, L.ACT_RELU, None, 0, bf),
    ("adapter2 fwd +res",        64000, 768, 256, 0, 0, 1, 0, bf, 0, bf),
    ("qkv fwd",                  64000, 768, 768, 0, 0, 1, 0, None, 0, bf),
    ("inproj fwd",               64000, 256, 256, 0, 0, 1, 0, None, 0, bf),
    ("scorer fwd tanh",          64000, 128, 768, 0, 0, 1, L.ACT_TANH, None, 0, bf),
    ("dgrad 768<-256 +res",      64000, 768, 256, 0, 1, 0, 0, bf, 0, bf),
    ("dgrad 256<-768 relu gate", 64000, 256, 768, 0, 1, 0, 0, None, L.GATE_RELU, bf),
    ("dgrad 768<-768 +res",      64000, 768, 768, 0, 1, 0, 0, bf, 0, bf),
    ("dgrad 256<-256",           64000, 256, 256, 0, 1, 0, 0, None, 0, bf),
    ("wgrad 768x768 K=64000",    768, 768, 64000, 1, 1, 0, 0, None, 0, f32),
    ("wgrad 256x768 K=64000",    256, 768, 64000, 1, 1, 0, 0, None, 0, f32),
    ("wgrad 768x256 K=64000",    768, 256, 64000, 1, 1, 0, 0, None, 0, f32),
    ("wgrad 256x256 K=64000",    256, 256, 64000, 1, 1, 0, 0, None, 0, f32),
    ("clf block fwd relu",       256, 512, 512, 0, 0, 1, L.ACT_RELU, None, 0, bf),
    ("clf block fwd +res f32",   256, 512, 512, 0, 0, 1, 0, f32, 0, f32),
    ("clf dgrad",                256, 512, 512, 0, 1, 0, 0, None, 0, f32),
    ("text qkv fwd",             16384, 768, 768, 0, 0, 1, 0, None, 0, bf),
    ("square 8192",              8192, 8192, 8192, 0, 0, 0, 0, None, 0, bf),
]


def run(idx, iters, ref=True, cold=False):
    name, M, N, K, ta, tb, bias, act, res, gate, odt = CASES[idx]
    # cold: rotate over enough copies of the token-sized operands that nothing is left in the 126 MB L2 when a copy comes
    # round again -- the situation inside a training step, where every activation is read once per kernel
    per_copy = 2.0 * M * K + M * N * (4 if odt == f32 else 2) * (2 if res is not None else 1) + (2.0 * M * N if gate else 0)
    ncopy = max(1, min(16, int(400e6 / max(per_copy, 1)) + 1)) if cold else 1
    a_ = [torch.randn((K, M) if ta else (M, K), device=dev).to(bf) for _ in range(ncopy)]
    b = torch.randn((K, N) if tb else (N, K), device=dev).to(bf)
    bias_t = torch.randn(N, device=dev) if bias else None
    r_ = [torch.randn(M, N, device=dev).to(res) if res is not None else None for _ in range(ncopy)]
    g_ = [torch.randn(M, N, device=dev).to(bf) if gate else None for _ in range(ncopy)]
    out_ = [torch.empty(M, N, device=dev, dtype=odt) for _ in range(ncopy)]
    a, out = a_[0], out_[0]
    kws = [dict(a_trans=bool(ta), b_trans=bool(tb), bias=bias_t, act=act, residual=r_[i], gate=g_[i], gate_mode=gate, out=out_[i])
           for i in range(ncopy)]
    for i in range(3):
        L.gemm(a_[i % ncopy], b, **kws[i % ncopy])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        L.gemm(a_[i % ncopy], b, **kws[i % ncopy])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    r, g = r_[0], g_[0]
    us_ref = float("nan")
    if ref:
        A = a.t() if ta else a
        Bm = b if tb else b.t()
        for _ in range(3):
            torch.matmul(A, Bm)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            torch.matmul(a_[_ % ncopy].t() if ta else a_[_ % ncopy], Bm)
        e1.record()
        torch.cuda.synchronize()
        us_ref = e0.elapsed_time(e1) / iters * 1e3
    fl = 2.0 * M * N * K
    by = 2.0 * (M * K + N * K) + M * N * (out.element_size() + (r.element_size() if r is not None else 0) + (2 if gate else 0))
    print(f"[{idx:2d}]{' cold' if cold else ''} {name:26s} M={M:6d} N={N:5d} K={K:6d}: {us:8.1f} us {fl/us/1e6:7.1f} TFLOP/s {by/us/1e3:7.1f} GB/s"
          f" | cuBLAS {us_ref:8.1f} us {fl/us_ref/1e6:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=-1)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--cases", default="", help="comma-separated case indices")
    ap.add_argument("--cold", action="store_true", help="rotate operand copies so that every launch reads from HBM")
    args = ap.parse_args()
    L.load()
    sel = [int(x) for x in args.cases.split(",")] if args.cases else (range(len(CASES)) if args.case < 0 else [args.case])
    for i in sel:
        run(i, args.iters if CASES[i][1] * CASES[i][2] * CASES[i][3] < 1e11 else min(args.iters, 5), ref=not args.no_ref,
            cold=args.cold)
