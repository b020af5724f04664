"""Micro-benchmark of the HBM-bound kernels at the cfg2 shapes (run under gpurun): LayerNorm fwd / bwd and the
bias-gradient column sum through the C-ABI, CUDA-event timed, with the algorithmic GB/s next to each.

  python tools/membound_bench.py [--case ln_bwd|ln_fwd|colsum|all] [--iters N]
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
import mmser_b200  # noqa: E402,F401
from mmser_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="all")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    lib = L.load()
    st = L.stream_ptr(dev)
    M, N = 64000, 768
    x = torch.randn(M, N, device=dev).to(bf)
    dy = torch.randn(M, N, device=dev).to(bf)
    y = torch.empty_like(x)
    dx = torch.empty_like(x)
    g = torch.randn(N, device=dev)
    b = torch.randn(N, device=dev)
    stats = torch.empty(M, 2, device=dev)
    dg = torch.zeros(N, device=dev)
    db = torch.zeros(N, device=dev)
    out = torch.zeros(N, device=dev)
    # second, differently-addressed set of buffers so back-to-back iterations do not hit in L2 (126 MB)
    x2, dy2, dx2 = x.clone(), dy.clone(), torch.empty_like(x)

    def ln_fwd():
        L.check(lib.ser_layernorm_fwd(x.data_ptr(), 0, y.data_ptr(), 0, g.data_ptr(), b.data_ptr(), stats.data_ptr(),
                                      M, N, 0, st))

    def ln_bwd(i=[0]):
        i[0] ^= 1
        xa, da, oa = (x, dy, dx) if i[0] else (x2, dy2, dx2)
        L.check(lib.ser_layernorm_bwd(da.data_ptr(), 0, xa.data_ptr(), 0, stats.data_ptr(), g.data_ptr(), b.data_ptr(),
                                      oa.data_ptr(), 0, dg.data_ptr(), db.data_ptr(), M, N, 0, st))

    def colsum(i=[0]):
        i[0] ^= 1
        L.check(lib.ser_colsum((dy if i[0] else dy2).data_ptr(), 0, N, M, N, out.data_ptr(), st))

    ln_fwd()
    cases = {"ln_fwd": (ln_fwd, 2 * M * N * 2), "ln_bwd": (ln_bwd, 3 * M * N * 2), "colsum": (colsum, M * N * 2)}
    for name, (fn, nbytes) in cases.items():
        if args.case not in ("all", name):
            continue
        us = timed(fn, args.iters)
        print(f"{name:8s} {M}x{N} bf16: {us:8.1f} us  {nbytes / us / 1e3:7.1f} GB/s (algorithmic bytes)", flush=True)


if __name__ == "__main__":
    main()
