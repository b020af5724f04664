"""One training step of the fusion head at a small shape; prints the loss and one checksum per parameter group's
gradient as JSON.  tests/test_gpu_switches.py runs it under different A/B switches (SER_PDL, SER_SIDE_STREAM,
SER_ATTN_KEEPBITS, ...) and compares: the switches change scheduling, never results."""
import json
import sys

import torch

sys.path.insert(0, ".")
import mmser_b200  # noqa: E402
from mmser_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
C = 6
B, Ta, Tt = (2, 610, 250) if "long" in sys.argv else (8, 70, 19)
head = mmser_b200.FusionHead(C, num_layers=4, dropout="reference").to(dev)
head.load_group_state(synth.head_weights(C, 4))
head.train()
torch.manual_seed(3)                                   # dropout seeds are drawn from torch's CPU generator
a, t, am, tm, labels = synth.make_inputs(B, Ta, Tt, C, seed=5)
ins = (a.to(dev).bfloat16(), t.to(dev).bfloat16(), am.to(dev), tm.to(dev), labels.to(dev))
if "persist" in sys.argv:          # one persistent gradient arena, re-zeroed every step (what DataParallelHead.train_step uses)
    head.persistent_grad_arena = True
    head(*ins)["loss"].backward()  # the reported step is the SECOND one: its zero fill meets the first step's gradients
out = head(*ins)
out["loss"].backward()
torch.cuda.synchronize()
res = {"loss": float(out["loss"]), "logits": out["logits"].double().abs().sum().item()}
for g in head.GROUPS:
    gs = [p.grad.double().reshape(-1) for p in getattr(head, g).parameters() if p.grad is not None]
    if gs:
        v = torch.cat(gs)
        res[g] = [v.abs().sum().item(), v.norm().item()]
print(json.dumps(res))
