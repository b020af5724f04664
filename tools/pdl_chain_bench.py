"""Chain of small dependent tcgen05 GEMM launches replayed as one CUDA graph: what a kernel -> kernel boundary costs
with and without programmatic dependent launch (run twice: SER_PDL=0 / SER_PDL=1).

    SER_PDL=0 python tools/pdl_chain_bench.py ; SER_PDL=1 python tools/pdl_chain_bench.py
"""
import os
import sys

import torch

sys.path.insert(0, ".")
import mmser_b200  # noqa: E402,F401
from mmser_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16


def chain(M, N, depth):
    x = [torch.randn(M, N, device=dev).to(bf) * 0.05 for _ in range(2)]
    w = [(torch.randn(N, N, device=dev) * (1.0 / N ** 0.5)).to(bf) for _ in range(depth)]

    def run():
        for i in range(depth):
            L.gemm(x[i & 1], w[i], out=x[(i + 1) & 1])          # out of launch i is the A operand of launch i + 1
    return run, x


def main():
    depth = 40
    for (M, N) in ((256, 512), (256, 256), (4096, 512)):
        run, x = chain(M, N, depth)
        x0 = x[0].clone()
        run(); torch.cuda.synchronize()
        ref = x[depth & 1].clone()
        x[0].copy_(x0)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            run()
            x[0].copy_(x0)
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                run()
        torch.cuda.synchronize()
        x[0].copy_(x0)
        g.replay(); torch.cuda.synchronize()
        same = torch.equal(x[depth & 1], ref)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            e0.record()
            for _ in range(10):
                g.replay()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        # eager (stream order, no graph)
        e0.record()
        for _ in range(10):
            run()
        e1.record(); torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / 10
        print(f"SER_PDL={os.environ.get('SER_PDL', '1')}  M={M} N=K={N} depth={depth}: graph {best * 1000 / depth:6.2f} us/launch, "
              f"eager {eager * 1000 / depth:6.2f} us/launch, graph result identical to eager: {same}")


if __name__ == "__main__":
    main()
