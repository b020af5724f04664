"""AttentiveStatsPooling forward + backward at the cfg2 shapes: kernel durations from CUPTI (torch.profiler), cold
operands (three rotating inputs).  Run with SER_PDL=0 so that a kernel's record does not include its wait on the
previous one.  SER_ASP_SLAB=256|128|64 selects the column slab of the statistics kernels."""
import collections
import os
import re
import sys

import torch

sys.path.insert(0, ".")
import mmser_b200  # noqa: E402

dev = torch.device("cuda:0")
D = 768
from torch.profiler import ProfilerActivity, profile
for (B, T, valid) in ((256, 250, 200), (256, 64, 40), (128, 1500, 1200)):
    pool = mmser_b200.models.AttentiveStatsPooling(D).to(dev)
    mask = torch.ones(B, T, device=dev)
    mask[:, valid:] = 0
    xs = [torch.randn(B, T, D, device=dev).bfloat16().requires_grad_(True) for _ in range(3)]
    up = torch.randn(B, 2 * D, device=dev).bfloat16()
    for it in range(3):
        pool(xs[it % 3], mask).backward(up)
    torch.cuda.synchronize()
    iters = 9
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for it in range(iters):
            pool(xs[it % 3], mask).backward(up)
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and ("asp_" in e.name or "gemm_tc" in e.name):
            nm = re.sub(r"\(.*", "", re.sub(r"void |ser::|\(anonymous namespace\)::", "", e.name))[:60]
            a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
    xb = B * T * D * 2
    print(f"B={B} T={T} D={D} bf16  (x = {xb / 1e6:.0f} MB)  SER_ASP_SLAB={os.environ.get('SER_ASP_SLAB', 'default')}")
    for nm, (n, us) in agg.items():
        per = us / iters
        note = ""
        if "asp_stats" in nm: note = f"  {xb / per / 1e3:7.0f} GB/s (x read once)"
        if "asp_bwd_stats" in nm: note = f"  {2 * xb / per / 1e3:7.0f} GB/s (x read, dx written)"
        print(f"   {nm:60s} {n / iters:4.1f}/iter {per:8.1f} us/iter{note}")
