"""Which stage's bf16 operand rounding sets the gradient-error floor of the bf16 tier?  (CPU, oracle only.)

E = fp64 oracle.  R(S) = fp32 oracle with the Linear operands rounded to bf16 in the stages S only (what an
implementation that keeps the other stages exact would give at best).  Prints ||R(S) - E|| / ||E|| per parameter group.

    python tools/bf16_floor_experiment.py [B] [Ta] [Tt]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fusion_head_oracle as O          # noqa: E402
from mmser_b200 import synth                        # noqa: E402

STAGES = ("adapter", "cross_attention", "attentive_stats_pooling", "fusion", "classifier")


class _Staged:
    """Wrap one oracle stage so that operand rounding is on (or off) inside it only."""

    def __init__(self, fn, on):
        self.fn, self.on = fn, on

    def __call__(self, *a, **k):
        prev = O._OPERAND_DTYPE
        O._OPERAND_DTYPE = torch.bfloat16 if self.on else None
        try:
            return self.fn(*a, **k)
        finally:
            O._OPERAND_DTYPE = prev


def run(B, Ta, Tt, C, dtype, rounded=()):
    w = synth.head_weights(C)
    a, t, am, tm, labels = synth.make_inputs(B, Ta, Tt, C, seed=1235)
    a, t = a.bfloat16().to(dtype), t.bfloat16().to(dtype)
    ws = {g: {n: v.detach().clone().to(dtype).requires_grad_(v.is_floating_point()) for n, v in d.items()} for g, d in w.items()}
    saved = {s: getattr(O, s) for s in STAGES}
    try:
        for s in STAGES:
            setattr(O, s, _Staged(saved[s], s in rounded))
        out = O.head_forward(a, t, am.to(dtype), tm.to(dtype), labels, ws, C)
        out["loss"].backward()
    finally:
        for s in STAGES:
            setattr(O, s, saved[s])
    grads = {}
    for g, d in ws.items():
        gs = [v.grad.double().reshape(-1) for n, v in d.items() if v.requires_grad and v.grad is not None
              and not n.endswith("anchor_clustering.temperature")]
        if gs:
            grads[g] = torch.cat(gs)
    return grads, out["logits"].detach().double()


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    Ta = int(sys.argv[2]) if len(sys.argv) > 2 else 250
    Tt = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    C = 4
    torch.set_num_threads(os.cpu_count())
    e, el = run(B, Ta, Tt, C, torch.float64)
    arms = {
        "all stages bf16 (autocast everywhere)": STAGES,
        "all but the classifier": STAGES[:-1],
        "classifier only": STAGES[-1:],
        "sequence stages only (adapter+cross)": STAGES[:2],
        "pooling only": ("attentive_stats_pooling",),
        "fusion only": ("fusion",),
        "none (fp32 everywhere)": (),
    }
    print(f"B={B} Ta={Ta} Tt={Tt}: ||grad(R) - grad(E)|| / ||grad(E)|| per parameter group; E = fp64 oracle")
    groups = list(e)
    print(f"{'bf16-rounded stages':42s} " + " ".join(f"{g[:10]:>10s}" for g in groups) + f" {'logits':>10s}")
    for name, st in arms.items():
        r, rl = run(B, Ta, Tt, C, torch.float32, st)
        row = " ".join(f"{((r[g] - e[g]).norm() / e[g].norm()).item():10.4f}" for g in groups)
        print(f"{name:42s} {row} {((rl - el).norm() / el.norm()).item():10.5f}")


if __name__ == "__main__":
    main()
