#!/usr/bin/env python
"""bench.py -- fusion-head training throughput (forward + backward incl. loss) on N B200s of one node.

  python bench.py --gpus 1 --steps 20 --warmup 5                 # our arm (bf16 tensor-core tier)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # data parallel, one rank per GPU
  python bench.py --impl reference ...                            # the reference's CPU path (oracle port)

One JSON line on stdout (rank 0).  metric = BASELINE.json's "fusion-head train samples/sec"; a step is one
forward+backward of the whole head (adapters -> cross attention -> pooling -> fusion -> 35-block classifier ->
loss) over one synthetic batch of the named shapes.  `value` times the step with inputs resident in HBM;
`e2e` times the public API (FusionHead / DataParallelHead) with HOST (pinned) inputs, host->device copies and a
device->host read of the loss inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

os.environ.setdefault("NCCL_DEBUG", "WARN")     # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1] / configs[2] (per-GPU batch 256) / configs[3]
    "cfg2": dict(B=256, Ta=250, Tt=64, C=4, desc="fusion head fwd+bwd bf16: B=256/GPU, T_audio=250, T_text=64, 4 classes"),
    "cfg3": dict(B=256, Ta=250, Tt=64, C=6, desc="CREMA-D shape DP step: 256/GPU, 6 classes"),
    "cfg4": dict(B=128, Ta=1500, Tt=256, C=4, desc="long utterance: B=128/GPU, T_audio=1500, T_text=256"),
    "cfg1": dict(B=8, Ta=250, Tt=64, C=4, desc="RAVDESS shape B=8"),
}


def fwd_flops_per_sample(Ta, Tt, C):
    """algorithmic forward FLOPs (2*MAC) per sample."""
    tok = (2 * (768 * 256 + 256 * 768) + 3 * 2 * 768 * 256 + 3 * 2 * 256 * 256 + 2 * 256 * 256 + 2 * 256 * 768 +
           2 * (768 * 128 + 128) + 6 * 768)
    return ((Ta + Tt) * tok + 4 * 2 * Ta * Tt * 256 + 2 * 2 * (1536 * 512 + 512 * 512) + 2 * 2 * (512 * 256 + 256) +
            2 * 512 * 512 * 71 + 2 * 512 * 256 + 2 * 256 * C + 2 * 256 * 128 + 2 * 128 * C + 2 * 256 * 64 + 2 * 64)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


def latest_traffic_file():
    d = os.path.join(ROOT, "profiles")
    if not os.path.isdir(d):
        return None
    c = sorted(f for f in os.listdir(d) if f.endswith("_traffic.json"))
    return os.path.join(d, c[-1]) if c else None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            sm.sort()
            out = dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path
# ----------------------------------------------------------------------------------------------------
def parse_dropout(text):
    """--dropout: 'reference' (the rates of the reference's training script) or one rate for every nn.Dropout."""
    from mmser_b200.head import dropout_rates
    return dropout_rates("reference" if text == "reference" else float(text))


def cpu_reference_run(wl, steps, warmup, sample_B, rates=None):
    from oracle import fusion_head_oracle as O
    from oracle import synth
    torch.set_num_threads(os.cpu_count() or 1)
    C = wl["C"]
    w = synth.head_weights(C)
    for grp in w.values():
        for k, v in grp.items():
            if v.is_floating_point() and k not in synth.CLASSIFIER_BUFFERS:
                v.requires_grad_(True)
    a, t, am, tm, labels = synth.make_inputs(sample_B, wl["Ta"], wl["Tt"], C, seed=1234)

    def step():
        for grp in w.values():
            for v in grp.values():
                v.grad = None
        # training-mode dropout like the reference's nn.Dropout layers: fresh Bernoulli masks every step
        src = O.random_dropout({"cross": rates["cross"], "fusion": rates["fusion"], "clf": rates["classifier"]}) \
            if rates and any(v > 0 for v in rates.values()) else None
        with O.dropout_masks(src):
            out = O.head_forward(a, t, am, tm, labels, w, C)
        out["loss"].backward()
        return float(out["loss"].detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=sample_B / dt, ms_per_step=dt * 1e3, cores=torch.get_num_threads(),
                sample=f"{steps} fwd+bwd steps of B={sample_B} at (Ta={wl['Ta']}, Tt={wl['Tt']}, C={C}), fp32, "
                       f"torch {torch.__version__} CPU")


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_B = min(wl["B"], 32)
    steps = max(1, min(args.steps, 8))
    rates = parse_dropout(args.dropout)
    r = cpu_reference_run(wl, steps, max(1, min(args.warmup, 2)), sample_B, rates)
    line = {
        "impl": "reference", "metric": "fusion_head_train_samples_per_sec", "value": r["value"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "shape": wl["desc"], "sample_batch": sample_B, "dropout": rates},
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch.distributed as dist
    import mmser_b200
    from mmser_b200 import _lib as L
    from mmser_b200.parallel import DataParallelHead, GraphedTrainStep
    from oracle import synth          # synthetic weights/inputs only (shared generator); the oracle itself is not used here

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the fusion head has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.load()

    def stage(msg):
        if rank == 0 and os.environ.get("BENCH_VERBOSE"):
            print(f"[bench] {msg}", file=sys.stderr, flush=True)

    peaks = load_peaks()
    B, Ta, Tt, C = wl["B"], wl["Ta"], wl["Tt"], wl["C"]
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32

    rates = parse_dropout(args.dropout)
    head = mmser_b200.FusionHead(C, dropout=rates).to(dev)
    head.load_group_state(synth.head_weights(C))
    head.train()
    dp = DataParallelHead(head)

    # per-rank shard of the synthetic global batch (different seed per rank = different samples)
    a, t, am, tm, labels = synth.make_inputs(B, Ta, Tt, C, seed=1234 + rank)
    host = dict(a=a.to(dtype).pin_memory(), t=t.to(dtype).pin_memory(), am=am.pin_memory(), tm=tm.pin_memory(),
                labels=labels.pin_memory())
    devin = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    torch.cuda.synchronize()

    def eager_step(inp):
        return dp.train_step(inp["a"], inp["t"], inp["am"], inp["tm"], inp["labels"])

    # CUDA graph of the whole step (forward + backward + loss): one cudaGraphLaunch instead of ~160 launches.
    # Data-parallel runs capture the NCCL all-reduces into the same graph (one graph per rank); if the process
    # group cannot be captured the rank falls back to eager launches and says so.
    graphed = None
    if args.graph:
        try:
            # the resident device tensors ARE the graph's static inputs: a replay is the step, no staging copy
            graphed = GraphedTrainStep(dp, devin["a"], devin["t"], devin["am"], devin["tm"], devin["labels"],
                                       static_inputs=True)
        except Exception as e:  # noqa: BLE001
            if world == 1:
                raise
            print(f"[bench rank {rank}] CUDA-graph capture of the data-parallel step failed ({type(e).__name__}: {e}); "
                  "running eager", file=sys.stderr)
            graphed = None
            torch.cuda.synchronize()

    def step(inp):
        if graphed is not None:
            if inp is devin:
                return graphed.replay()
            return graphed(inp["a"], inp["t"], inp["am"], inp["tm"], inp["labels"])
        return eager_step(inp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, tail=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if tail is not None:
            tail()                       # still inside the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    stage(f"setup done (graph={'yes' if graphed is not None else 'no'})")
    # ---------------- device-resident timing (value) ----------------
    for _ in range(max(args.warmup, 3)):
        out = step(devin)
    barrier()
    l0 = L.launch_count()
    eager_step(devin)
    launches_per_step = L.launch_count() - l0          # kernels of ONE step (a graph replay re-issues exactly these)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(lambda: step(devin), args.steps)
    clocks = sampler.stop() if sampler else None
    launches = launches_per_step * args.steps
    loss_val = float(out["loss"].detach())

    stage(f"device-resident timing done: {ms:.3f} ms/step")
    # ---------------- end-to-end through the public API with host buffers ----------------
    # Every step: H2D copy of that step's inputs from pinned host memory (copy stream, double buffered, overlapping the
    # previous step's compute), the step itself (one CUDA graph per input buffer, so the copied tensors ARE the graph's
    # static inputs -- no device-to-device staging), and a D2H read of the step's loss into pinned host memory.  The
    # host consumes the loss one step late (it launches step i+1 first, then waits for loss i), as a training loop
    # that logs its loss does; the last loss is read before the timer stops.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]   # double buffer
    ready = [torch.cuda.Event(), torch.cuda.Event()]        # H2D of buffer i finished (recorded on the copy stream)
    consumed = [torch.cuda.Event(), torch.cuda.Event()]     # compute finished reading buffer i (recorded on the main stream)
    done = [torch.cuda.Event(), torch.cuda.Event()]         # loss of the step on buffer i is in host memory
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    for ev in consumed:
        ev.record()
    e2e_graphs = None
    if graphed is not None:
        try:
            e2e_graphs = [GraphedTrainStep(dp, b["a"], b["t"], b["am"], b["tm"], b["labels"], static_inputs=True)
                          for b in bufs]
        except Exception as e:  # noqa: BLE001
            if world == 1:
                raise
            print(f"[bench rank {rank}] e2e graph capture failed ({type(e).__name__}: {e}); using the staged path",
                  file=sys.stderr)
            e2e_graphs = None

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i])            # never overwrite a buffer the step is still reading
            for k, v in host.items():
                bufs[i][k].copy_(v, non_blocking=True)     # pinned host -> device
            ready[i].record(copy_stream)

    state = {"i": 0, "loss": 0.0, "pending": None}
    prefetch(0)

    def e2e_step():
        i = state["i"]
        torch.cuda.current_stream().wait_event(ready[i])
        o = e2e_graphs[i].replay() if e2e_graphs is not None else step(bufs[i])
        consumed[i].record()
        loss_host[i].copy_(o["loss"].detach().reshape(()), non_blocking=True)     # device -> host read of the step's result
        done[i].record()
        prefetch(1 - i)                                  # next step's host->device copy overlaps this step's compute
        if state["pending"] is not None:                 # consume the previous step's loss (one step late)
            j = state["pending"]
            done[j].synchronize()
            state["loss"] = float(loss_host[j])
        state["pending"] = i
        state["i"] = 1 - i

    def e2e_tail():
        j = state["pending"]
        done[j].synchronize()
        state["loss"] = float(loss_host[j])
        state["pending"] = None

    for _ in range(2):
        e2e_step()
    e2e_tail()
    ms_e2e = timed(e2e_step, max(3, args.steps // 2), tail=e2e_tail)

    stage(f"e2e timing done: {ms_e2e:.3f} ms/step")
    # ---------------- the same step with every dropout rate set to 0 (reported beside the headline) ----------------
    ms_nodrop = None
    if any(v > 0 for v in rates.values()):
        head.set_dropout(0.0)
        g0 = None
        if graphed is not None:
            try:
                g0 = GraphedTrainStep(dp, devin["a"], devin["t"], devin["am"], devin["tm"], devin["labels"])
            except Exception:  # noqa: BLE001
                g0 = None
        step0 = (lambda: g0.replay()) if g0 is not None else (lambda: eager_step(devin))
        for _ in range(3):
            step0()
        ms_nodrop = timed(step0, args.steps)
        head.set_dropout(rates)
        del g0
        stage(f"no-dropout timing done: {ms_nodrop:.3f} ms/step")
    # ---------------- per-kernel-family profile (CUDA events around every launch; separate pass) ----------------
    prof = {}
    nprof = 3
    barrier()
    if rank == 0:
        L.prof_enable(True)
    for _ in range(nprof):             # every rank runs the pass (the steps contain collectives); rank 0 records
        eager_step(devin)              # events cannot be timed inside a graph replay: profile the eager launches
    barrier()
    if rank == 0:
        detail = L.prof_report()
        L.prof_enable(False)
        if args.profile_detail:
            with open(args.profile_detail, "w") as f:
                for k, v in sorted(detail.items(), key=lambda kv: -kv[1]["ms"]):
                    f.write(f"{k:44s} n/step={v['launches']/nprof:6.1f} ms/step={v['ms']/nprof:8.4f} us/launch={1e3*v['ms']/v['launches']:8.1f} "
                            f"TFLOP/s={v['flops']/max(v['ms'],1e-9)/1e9:8.1f} GB/s={v['bytes']/max(v['ms'],1e-9)/1e6:8.1f}\n")
        gemm_detail = {k: v for k, v in detail.items() if k.startswith("gemm_tc")}
        prof = {}
        for k, v in detail.items():            # aggregate shape-tagged records per kernel family
            fam = prof.setdefault(k.split(":")[0], dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            for kk in fam:
                fam[kk] += v[kk]
        for v in prof.values():
            v["ms_per_step"] = v["ms"] / nprof
            v["launches_per_step"] = v["launches"] / nprof

    # ---------------- cpu baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(WORKLOADS["cfg1"] | {"C": C}, steps=6, warmup=1, sample_B=8, rates=rates)
        cpu = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    def finish():
        # Tear-down of a process group whose collectives live inside CUDA graphs can block; everything is
        # measured and printed by now, so synchronise, meet once more and leave without the NCCL destructor.
        sys.stdout.flush(); sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        finish()
        return

    total_B = B * world
    value = total_B / (ms * 1e-3)
    step_flops = 3.0 * fwd_flops_per_sample(Ta, Tt, C) * B             # per GPU, 3x-forward convention
    # dominant kernel family = the tcgen05 GEMM (forward / dgrad / wgrad launches of gemm_tc_kernel)
    fam = {k: v for k, v in prof.items() if k.startswith("gemm_tc")}
    roof = None
    if fam:
        gflops = sum(v["flops"] for v in fam.values())
        gms = sum(v["ms"] for v in fam.values())
        nl = sum(v["launches"] for v in fam.values())
        ach = gflops / (gms * 1e-3) / 1e12
        traffic, tsrc = None, None
        tfile = latest_traffic_file()
        if tfile:                      # DRAM bytes per launch of the same kernel from the committed ncu launch list
            try:
                tk = json.load(open(tfile))["kernels"].get("gemm_tc_kernel")
                if tk:
                    traffic, tsrc = tk["dram_bytes_per_launch"], os.path.relpath(tfile, ROOT)
            except Exception:  # noqa: BLE001
                pass
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 bf16 GEMM, all fwd/dgrad/wgrad launches of a step)",
                "achieved": ach, "peak": peaks["tf_sus"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sus"],
                "traffic": traffic, "traffic_source": tsrc,
                "algorithmic_flops_per_launch": gflops / nl, "algorithmic_bytes_per_launch": sum(v["bytes"] for v in fam.values()) / nl,
                "avg_launch_us": gms * 1e3 / nl, "peak_source": peaks["src"] + " sustained bf16 (kernel timed inside a long step)",
                "launches_per_step": sum(v["launches_per_step"] for v in fam.values()),
                "ms_per_step": sum(v["ms_per_step"] for v in fam.values())}
        # context for `frac`: the family mixes tensor-bound (K = 768), HBM-bound (K = 256) and launch-latency-bound
        # (M = 256) shapes.  (1) time-weighted fraction of each launch's OWN roofline max(flops / tensor peak,
        # algorithmic bytes / HBM peak); (2) the same two numbers over the launches of >= 20 us only.
        def _roof(sel):
            t = sum(v["ms"] for v in sel)
            if t <= 0:
                return None
            t_roof = sum(max(v["flops"] / (peaks["tf_sus"] * 1e12), v["bytes"] / (peaks["hbm"] * 1e9)) * 1e3 for v in sel)
            fl = sum(v["flops"] for v in sel)
            return {"launches_per_step": sum(v["launches"] for v in sel) / nprof, "ms_per_step": t / nprof,
                    "achieved_tflops": fl / (t * 1e-3) / 1e12, "frac_of_tensor_peak": fl / (t * 1e-3) / 1e12 / peaks["tf_sus"],
                    "frac_of_own_roofline": t_roof / t}
        allg = list(gemm_detail.values())
        big = [v for v in allg if v["ms"] / max(v["launches"], 1) >= 0.020]
        roof["all_launches"] = _roof(allg)
        roof["launches_over_20us"] = _roof(big)
    tot_prof_ms = sum(v["ms_per_step"] for v in prof.values()) or 1.0
    families = {k: {"ms_per_step": round(v["ms_per_step"], 4), "launches_per_step": v["launches_per_step"],
                    "share": round(v["ms_per_step"] / tot_prof_ms, 4),
                    "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 2),
                    "gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)} for k, v in sorted(prof.items())}
    line = {
        "metric": "fusion_head_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": args.workload, "shape": wl["desc"], "global_batch": total_B, "per_gpu_batch": B,
                   "parallelism": f"dp{world}", "dropout": rates, "cuda_graph": graphed is not None,
                   "l2": "no explicit flush: one step touches > 2 GB of activations per GPU, far above the 126 MB L2"},
        "step_model_flops_per_gpu": step_flops,
        "model_tflops_per_gpu": step_flops / (ms * 1e-3) / 1e12,
        "model_frac_of_sustained_bf16_peak": step_flops / (ms * 1e-3) / 1e12 / peaks["tf_sus"],
        "no_dropout": None if ms_nodrop is None else {"value": total_B / (ms_nodrop * 1e-3), "unit": "samples/s",
                                                      "ms_per_step": ms_nodrop},
        "roofline": roof, "kernel_families": families,
        "cpu_baseline": cpu,
        "e2e": {"value": total_B / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "api": "mmser_b200.parallel.GraphedTrainStep / DataParallelHead.train_step(FusionHead) with pinned host inputs; "
                       "one graph per input buffer, loss read back every step and consumed one step late"},
        "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
        "clocks": clocks, "loss": loss_val,
    }
    print(json.dumps(line), flush=True)
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dropout", default="reference",
                    help="'reference' = the rates of the reference's training script (cross 0.1, fusion 0.1, classifier 0.15; "
                         "SURVEY.md 8(d)); or one rate for every nn.Dropout of the head, e.g. 0")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--profile-detail", default="", help="write the per-shape kernel table of the profiling pass here")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch > 0:
        wl["B"] = args.batch
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
