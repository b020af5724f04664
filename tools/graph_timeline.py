"""Per-kernel timeline of the CUDA-graph-replayed training step (CUPTI through torch.profiler): the durations the
kernels have INSIDE the replay and the gaps between consecutive kernels -- what per-launch CUDA events cannot give.

    python tools/graph_timeline.py [--workload cfg2] [--out gpurun_out/timeline.txt]
"""
import argparse
import collections
import re
import sys

import torch

sys.path.insert(0, ".")
import mmser_b200  # noqa: E402
from mmser_b200 import synth  # noqa: E402
from mmser_b200.parallel import DataParallelHead, GraphedTrainStep  # noqa: E402

WL = {"cfg2": (256, 250, 64, 4), "cfg3": (256, 250, 64, 6), "cfg4": (128, 1500, 256, 4)}


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::|ser::|<unnamed>::", "", name)
    m = re.match(r"([A-Za-z0-9_:]+)(<[^(]*>)?", name)
    return (m.group(1) + (m.group(2) or ""))[:70] if m else name[:70]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--out", default="")
    ap.add_argument("--replays", type=int, default=5)
    ap.add_argument("--dropout", type=float, default=-1.0)
    args = ap.parse_args()
    B, Ta, Tt, C = WL[args.workload]
    import os
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:                                   # under torchrun: the data-parallel step, NCCL kernels included
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    rates = {"cross": 0.1, "fusion": 0.1, "classifier": 0.15} if args.dropout < 0 else args.dropout
    head = mmser_b200.FusionHead(C, dropout=rates).to(dev)
    head.load_group_state(synth.head_weights(C))
    head.train()
    dp = DataParallelHead(head)
    a, t, am, tm, labels = synth.make_inputs(B, Ta, Tt, C, seed=1234 + rank)
    ins = [x.to(dev) for x in (a.bfloat16(), t.bfloat16(), am, tm, labels)]
    g = GraphedTrainStep(dp, *ins, static_inputs=True)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.replays):
            g.replay()
        torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
        if rank != 0:
            torch.distributed.destroy_process_group()
            return
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    evs = sorted(evs, key=lambda e: e.time_range.start)
    evs = [e for e in evs if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    n = len(evs) // args.replays
    lines = []
    if n == 0:
        print("no kernel records (CUPTI unavailable?)")
        return
    # last replay: steady state
    last = evs[-n:]
    t0 = last[0].time_range.start
    dur = collections.OrderedDict()
    gap_after = collections.defaultdict(float)
    total_k = total_gap = 0.0
    lines.append(f"{args.workload}: {n} kernels per replay; last of {args.replays} replays; times in us")
    lines.append(f"{'#':>3s} {'start':>9s} {'dur':>8s} {'gap':>6s}  kernel")
    for i, e in enumerate(last):
        d = e.time_range.end - e.time_range.start
        gap = (last[i + 1].time_range.start - e.time_range.end) if i + 1 < len(last) else 0.0
        nm = short(e.name)
        dur.setdefault(nm, [0, 0.0, 0.0])
        dur[nm][0] += 1; dur[nm][1] += d; dur[nm][2] += gap
        total_k += d; total_gap += gap
        lines.append(f"{i:3d} {e.time_range.start - t0:9.1f} {d:8.1f} {gap:6.1f}  {nm}")
    span = last[-1].time_range.end - t0
    lines.append("")
    lines.append(f"span of the replay {span:.1f} us = kernels {total_k:.1f} us + gaps {total_gap:.1f} us (negative gap = overlap)")
    lines.append(f"{'kernel':70s} {'n':>4s} {'sum dur':>9s} {'avg':>8s} {'sum gap after':>13s}")
    for nm, (c, d, gp) in sorted(dur.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{nm:70s} {c:4d} {d:9.1f} {d / c:8.1f} {gp:13.1f}")
    txt = "\n".join(lines)
    if args.out:
        open(args.out, "w").write(txt + "\n")
    print("\n".join(lines[-(len(dur) + 3):]))


if __name__ == "__main__":
    main()
