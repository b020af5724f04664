"""FusedAdamW over the fusion head's parameters (10 groups as in src/train.py:72-83) next to torch.optim.AdamW
(foreach and fused=True), CUDA-event timed.  Algorithmic traffic: 28 bytes per parameter per step."""
import sys

import torch

sys.path.insert(0, ".")
import mmser_b200  # noqa: E402

dev = torch.device("cuda:0")
head = mmser_b200.FusionHead(4).to(dev)
head.features  # noqa: B018
for g in head.GROUPS:
    fp = getattr(getattr(head, g), "_flat", None)
    if fp is not None:
        fp.ensure()
groups = []
for name, lr, wd in (("adapter_a", 1e-4, 0.025), ("adapter_t", 1e-4, 0.025), ("cross", 1e-3, 0.05), ("pool_a", 1e-3, 0.05),
                     ("pool_t", 1e-3, 0.05), ("fusion", 1e-3, 0.05), ("classifier", 1.5e-3, 0.06), ("prototypes", 1e-3, 0.05)):
    groups.append(dict(params=list(getattr(head, name).parameters()), lr=lr, weight_decay=wd))
nparam = sum(p.numel() for g in groups for p in g["params"])
for g in groups:
    for p in g["params"]:
        p.grad = torch.randn_like(p) * 1e-3


def timed(opt, iters=20):
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def copy_groups():
    return [dict(g, params=list(g["params"])) for g in groups]


print(f"{nparam/1e6:.2f} M parameters in {sum(len(g['params']) for g in groups)} tensors, {len(groups)} groups")
for label, make in (("FusedAdamW (this repo)", lambda: mmser_b200.optim.FusedAdamW(copy_groups())),
                    ("torch AdamW foreach", lambda: torch.optim.AdamW(copy_groups(), foreach=True)),
                    ("torch AdamW fused=True", lambda: torch.optim.AdamW(copy_groups(), fused=True))):
    us = timed(make())
    print(f"{label:26s}: {us:8.1f} us/step  {28.0 * nparam / us / 1e3:7.1f} GB/s (28 B/param)")
opt = mmser_b200.optim.FusedAdamW(copy_groups())
opt.step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    opt.clip_grad_norm_(1.0)
e1.record(); torch.cuda.synchronize()
print(f"fused clip_grad_norm_      : {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us  (4 B/param read)")
