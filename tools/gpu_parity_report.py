"""Verbose parity report (every tensor with its error, noise floor and limit); run under gpurun.
Usage: python tools/gpu_parity_report.py [f32|bf16|all] [-v] [case ...]"""
import sys
import traceback

import torch

sys.path.insert(0, ".")
from tests import parity_cases as PC  # noqa: E402

dev = torch.device("cuda:0")
torch.set_num_threads(16)
args = [a for a in sys.argv[1:] if not a.startswith("-")]
tier = args[0] if args else "all"
names = args[1:] or list(PC.ALL_CASES)
summary = {}
for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
    if tier not in (tag, "all"):
        continue
    for n in names:
        print(f"== {n}/{tag}")
        try:
            fails, worst = PC.ALL_CASES[n]().check(dtype, dev, verbose="-v" in sys.argv)
            summary[f"{n}/{tag}"] = (len(fails), worst)
        except Exception:  # noqa: BLE001
            traceback.print_exc()
            summary[f"{n}/{tag}"] = (-1, float("inf"))
print("==== SUMMARY (failures, worst err/limit) ====")
for k, (nf, w) in summary.items():
    print(f"{k:40s} fails={nf:4d} worst={w:.3f}")
