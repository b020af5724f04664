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
            case = PC.ALL_CASES[n]()
            fails, worst = case.check(dtype, dev, verbose="-v" in sys.argv)
            # raw numbers against the PLAIN tolerance (no noise term, no outlier set-aside)
            tol = PC.TOL[dtype]
            outs = [r for r in case.records if r[0] == "out"]
            grads = [r for r in case.records if r[0] != "out"]
            wo = max((r[2] for r in outs), default=0.0)
            wg = max((r[2] for r in grads), default=0.0)
            over = sum(1 for r in case.records if r[2] > tol)
            print(f"   plain tolerance {tol:g}: worst forward err = {wo:.2e} ({wo / tol:.2f}x), worst gradient err = {wg:.2e} "
                  f"({wg / tol:.2f}x), tensors above plain tolerance: {over} of {len(case.records)}; "
                  f"worst reference-arithmetic error (R vs E) = {max((r[3] for r in case.records), default=0.0):.2e}")
            summary[f"{n}/{tag}"] = (len(fails), worst, wo / tol, wg / tol, over, len(case.records))
        except Exception:  # noqa: BLE001
            traceback.print_exc()
            summary[f"{n}/{tag}"] = (-1, float("inf"), float("inf"), float("inf"), -1, -1)
print("==== SUMMARY: failures under the noise-aware criterion (tests/parity_cases.py), worst err / limit; then the RAW numbers:")
print("====          worst forward / gradient error as multiples of the PLAIN tolerance (1e-4 fp32 max-norm, 2e-2 bf16 Frobenius)")
for k, (nf, w, wo, wg, over, tot) in summary.items():
    print(f"{k:40s} fails={nf:4d} worst/limit={w:6.3f} | forward {wo:7.2f}x  gradients {wg:7.2f}x  above plain tol: {over}/{tot}")
