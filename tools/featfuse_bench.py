"""Per-utterance feature fusion at the cfg2 audio shape (B 256, T 250, hid 768, F 20, bf16, dropout
0.1): per-kernel-family CUDA-event times from the library profiler, and the same op in eager PyTorch (the reference's
expand + cat + Linear(788 -> 768) + ReLU + Dropout, bf16 autocast) on the same GPU for scale."""
import sys

import torch
import torch.nn as nn

sys.path.insert(0, ".")
import mmser_b200  # noqa: E402
from mmser_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
B, T, D, F = 256, 250, 768, 20
fuse = mmser_b200.models.UtteranceFeatureFusion(D, F).to(dev).train()
feats = torch.rand(B, F, device=dev)
xs = [torch.randn(B, T, D, device=dev).bfloat16().requires_grad_(True) for _ in range(3)]   # rotate: defeat L2 residency
up = torch.randn(B, T, D, device=dev).bfloat16()
for it in range(3):
    fuse(xs[it % 3], feats).backward(up)
torch.cuda.synchronize()
L.prof_enable(True)
iters = 12
for it in range(iters):
    fuse(xs[it % 3], feats).backward(up)
rep = L.prof_report()
L.prof_enable(False)
total = 0.0
for k, v in sorted(rep.items()):
    us = v["ms"] / iters * 1e3
    total += us
    print(f"{k:34s} {v['launches'] / iters:4.1f} launches/iter {us:8.1f} us/iter")
print(f"library kernels, forward + backward: {total:.1f} us/iter")

ref = nn.Sequential(nn.Linear(D + F, D), nn.ReLU(), nn.Dropout(0.1)).to(dev).train()


def eager(x):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = ref(torch.cat([x, feats.to(x.dtype).unsqueeze(1).expand(B, T, F)], dim=-1))
    y.backward(up)


for it in range(3):
    eager(xs[it % 3])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for it in range(iters):
    eager(xs[it % 3])
e1.record()
torch.cuda.synchronize()
print(f"eager PyTorch (batched cat + Linear, bf16 autocast), forward + backward: {e0.elapsed_time(e1) / iters * 1e3:.1f} us/iter")
e0.record()
for it in range(iters):
    fuse(xs[it % 3], feats).backward(up)
e1.record()
torch.cuda.synchronize()
print(f"this library through the nn.Module, forward + backward (eager launches): {e0.elapsed_time(e1) / iters * 1e3:.1f} us/iter")
