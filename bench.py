#!/usr/bin/env python
"""bench.py -- fusion-head throughput on N B200s of one node.

  python bench.py --gpus 1 --steps 20 --warmup 5                 # our arm: cfg2 training step (bf16 tensor-core tier)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # data parallel, one rank per GPU
  python bench.py --workload cfg3|cfg4 ...                        # the other training configs of BASELINE.json
  python bench.py --workload cfg5 ...                             # eval path: 4096 samples x 5 TTA views, OpenMax, sweep
  python bench.py --impl reference ...                            # the reference's CPU path (oracle port) at the same config
  python bench.py --impl eager_gpu ...                            # the reference's arithmetic in stock PyTorch on the same GPU

One JSON line on stdout (rank 0).  metric = BASELINE.json's "fusion-head train samples/sec"; a step is one
forward+backward of the whole head (adapters -> cross attention -> pooling -> fusion -> 35-block classifier ->
loss) over one synthetic batch of the named shapes.  `value` times the step with inputs resident in HBM;
`e2e` times the public API (FusionHead / DataParallelHead) with HOST (pinned) inputs, host->device copies and a
device->host read of the loss inside the timed region.  The line also carries `eager_gpu` (same-box PyTorch eager,
fp32 / tf32 / autocast(bf16), eager and CUDA-graphed: SURVEY.md 8(d)), `with_optimizer_step`, `no_dropout` and, in
data-parallel runs, `no_overlap` (gradient all-reduce after the backward instead of overlapped with it).
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
    # BASELINE.json configs[1] / configs[2] (per-GPU batch 256) / configs[3] / configs[4]
    "cfg2": dict(B=256, Ta=250, Tt=64, C=4, desc="fusion head fwd+bwd bf16: B=256/GPU, T_audio=250, T_text=64, 4 classes"),
    "cfg3": dict(B=256, Ta=250, Tt=64, C=6, desc="CREMA-D shape DP step: 256/GPU, 6 classes"),
    "cfg4": dict(B=128, Ta=1500, Tt=256, C=4, desc="long utterance: B=128/GPU, T_audio=1500, T_text=256"),
    "cfg1": dict(B=8, Ta=250, Tt=64, C=4, desc="RAVDESS shape B=8"),
    "cfg5": dict(B=4096, Ta=250, Tt=64, C=6, V=5, desc="eval path: B=4096 x 5 TTA views, OpenMax (fitted Weibull), view mean, "
                                                        "temperature sweep + scaling, softmax / argmax / energy"),
}
METRIC_TRAIN = "fusion_head_train_samples_per_sec"
METRIC_EVAL = "fusion_head_eval_samples_per_sec"


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


def parse_dropout(text):
    """--dropout: 'reference' (the rates of the reference's training script) or one rate for every nn.Dropout."""
    from mmser_b200.head import dropout_rates
    return dropout_rates("reference" if text == "reference" else float(text))


def train_config(args, wl, world, rates, graphed=None):
    """The `config` object of a training line: identical for our arm, the reference arm and the eager arm."""
    cfg = {"workload": args.workload, "shape": wl["desc"], "global_batch": wl["B"] * world, "per_gpu_batch": wl["B"],
           "parallelism": f"dp{world}", "dropout": rates}
    if graphed is not None:
        cfg["cuda_graph"] = graphed
        cfg["l2"] = "no explicit flush: one step touches > 2 GB of activations per GPU, far above the 126 MB L2"
    return cfg


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path
# ----------------------------------------------------------------------------------------------------
def cpu_reference_run(wl, steps, warmup, sample_B, rates=None, budget_s=None):
    """fwd+bwd steps of the oracle (CPU restatement of the reference, all host threads).  `budget_s`: stop early once
    the timed steps have used this much wall clock (at least one step is always timed)."""
    from oracle import fusion_head_oracle as O
    from mmser_b200 import synth
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
    done, t0 = 0, time.perf_counter()
    for _ in range(steps):
        step()
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    return dict(value=sample_B / dt, ms_per_step=dt * 1e3, cores=torch.get_num_threads(), steps=done,
                sample=f"{done} fwd+bwd steps of B={sample_B} at (Ta={wl['Ta']}, Tt={wl['Tt']}, C={C}), fp32, "
                       f"torch {torch.__version__} CPU")


def cpu_eval_run(wl, sample_B, views, steps=1):
    """cfg5 on the CPU oracle: `views` forwards of the head in eval mode with fitted OpenMax, view mean, temperature
    sweep, scaling, softmax / argmax / energy (src/eval.py:174-206) on a sample of `sample_B` samples."""
    from oracle import fusion_head_oracle as O
    from mmser_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    C = wl["C"]
    w = synth.head_weights(C)
    a, t, am, tm, labels = synth.make_inputs(sample_B, wl["Ta"], wl["Tt"], C, seed=1234)
    view_a = [a] + [a + 0.05 * torch.randn_like(a) * am[..., None] for _ in range(views - 1)]

    def fwd(av):
        x, y = O.adapter(av, w["adapter_a"]), O.adapter(t, w["adapter_t"])
        ea, et = O.cross_attention(x, y, am, tm, w["cross"])
        return O.fusion(O.attentive_stats_pooling(ea, am, w["pool_a"]), O.attentive_stats_pooling(et, tm, w["pool_t"]),
                        w["fusion"])

    with torch.no_grad():
        feats = O.classifier_features(fwd(a), w["classifier"])
        w["classifier"].update(O.fit_weibull(feats, labels, C, w["classifier"]))
        val_logits = O.classifier(fwd(a), w["classifier"], use_openmax=False)
        t0 = time.perf_counter()
        for _ in range(steps):
            lv = torch.stack([O.classifier(fwd(av), w["classifier"], use_openmax=True) for av in view_a])
            T = O.find_optimal_temperature(val_logits, labels)
            lg = O.tta_mean(lv) / T
            _ = torch.softmax(lg, -1), lg.argmax(-1), O.energy_score(lg)
        dt = (time.perf_counter() - t0) / steps
    return dict(value=sample_B / dt, ms_per_step=dt * 1e3, cores=torch.get_num_threads(),
                sample=f"{steps} eval pass(es) of B={sample_B} x {views} views at (Ta={wl['Ta']}, Tt={wl['Tt']}, C={C}), fp32, "
                       f"torch {torch.__version__} CPU")


def run_reference(args, wl):
    """The reference's own (CPU, fp32, PyTorch) implementation of the path on the box's host cores -- the oracle port,
    since the Python reference cannot travel (kind "port").  Same workload, batch, dropout, steps and warm-up as our arm;
    only if the timed steps would exceed ~2.5 minutes are they cut short (`steps` then says how many were timed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rates = parse_dropout(args.dropout)
    if args.workload == "cfg5":
        sB = min(wl["B"], 64)
        r = cpu_eval_run(wl, sB, wl["V"], steps=1)
        line = {"impl": "reference", "metric": METRIC_EVAL, "value": r["value"], "unit": "samples/s", "n_gpus": args.gpus,
                "steps": 1, "warmup": 0, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "shape": wl["desc"], "sample_batch": sB, "views": wl["V"]},
                "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    warm = max(1, min(args.warmup, 3))
    r = cpu_reference_run(wl, args.steps, warm, wl["B"], rates, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC_TRAIN, "value": r["value"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": warm, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": train_config(args, wl, world, rates),
        "steps_requested": args.steps,
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# same-box PyTorch-eager arm (baseline/eager_head.py): the reference's arithmetic through stock torch modules
# ----------------------------------------------------------------------------------------------------
def eager_gpu_measure(wl, rates, dev, steps=5, warmup=3, precisions=("fp32", "tf32", "autocast_bf16"), seed=1234):
    """{precision: {eager_ms, graphed_ms, ...}} for one fwd+bwd step of baseline.eager_head.EagerHead on `dev`.
    fp32 = torch defaults (what the reference runs: no TF32); tf32 = torch.set_float32_matmul_precision('high');
    autocast_bf16 = the whole step under torch.autocast(bfloat16) (the reference autocasts only classifier + loss, and
    in fp16: this is the most favourable reading).  `graphed` = the same step replayed as one CUDA graph, with the three
    host-synchronising constructs of the reference's loss code replaced by sync-free equivalents (graph_safe)."""
    from baseline.eager_head import EagerHead
    from mmser_b200 import synth
    C = wl["C"]
    a, t, am, tm, labels = (x.to(dev) for x in synth.make_inputs(wl["B"], wl["Ta"], wl["Tt"], C, seed=seed))
    weights = synth.head_weights(C)
    out = {}
    prev = torch.get_float32_matmul_precision()
    try:
        for prec in precisions:
            torch.set_float32_matmul_precision("high" if prec == "tf32" else "highest")
            res = {}
            for graphed in (False, True):
                head = EagerHead(C, dropout=rates, graph_safe=graphed).to(dev)
                head.load_group_state(weights)
                head.train()
                params = [p for p in head.parameters()]

                def step():
                    for p in params:
                        p.grad = None
                    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(prec == "autocast_bf16")):
                        o = head(a, t, am, tm, labels)
                    o["loss"].backward()
                    return o["loss"]

                fn = step
                try:
                    if graphed:
                        side = torch.cuda.Stream(device=dev)
                        side.wait_stream(torch.cuda.current_stream(dev))
                        with torch.cuda.stream(side):
                            for _ in range(3):
                                step()
                        torch.cuda.current_stream(dev).wait_stream(side)
                        torch.cuda.synchronize(dev)
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            step()
                        fn = g.replay
                    for _ in range(warmup):
                        fn()
                    torch.cuda.synchronize(dev)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        fn()
                    e1.record()
                    torch.cuda.synchronize(dev)
                    res["graphed_ms" if graphed else "eager_ms"] = e0.elapsed_time(e1) / steps
                except Exception as e:  # noqa: BLE001
                    res["graphed_error" if graphed else "eager_error"] = f"{type(e).__name__}: {e}"[:200]
                    torch.cuda.synchronize(dev)
                del head, params, fn
                torch.cuda.empty_cache()
            best = min([v for k, v in res.items() if k.endswith("_ms")], default=None)
            if best is not None:
                res["best_ms"] = best
                res["samples_per_s"] = wl["B"] / (best * 1e-3)
            out[prec] = res
    finally:
        torch.set_float32_matmul_precision(prev)
    return out


def run_eager(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "eager_gpu", "unavailable": "no CUDA device"}), flush=True)
        return
    if args.workload == "cfg5":
        print(json.dumps({"impl": "eager_gpu", "unavailable": "the eager arm covers the training workloads"}), flush=True)
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    rates = parse_dropout(args.dropout)
    sampler = ClockSampler(dev.index or 0)
    res = eager_gpu_measure(wl, rates, dev, steps=args.steps, warmup=max(args.warmup, 3))
    clocks = sampler.stop()
    best = res.get("autocast_bf16", {}).get("best_ms") or min(v["best_ms"] for v in res.values() if "best_ms" in v)
    line = {"impl": "eager_gpu", "metric": METRIC_TRAIN, "value": wl["B"] / (best * 1e-3), "unit": "samples/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": best, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (torch.autocast) -- fp32 / tf32 variants alongside",
            "data": "synthetic", "config": train_config(args, wl, 1, rates), "variants": res, "clocks": clocks,
            "note": "stock-PyTorch restatement of the reference modules (baseline/eager_head.py, pinned to the oracle by "
                    "tests/test_eager_baseline.py); value = the fastest of eager / CUDA-graphed under autocast(bf16)"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# shared pieces of our arms
# ----------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads to the CPUs of the NUMA node its GPU hangs off (sysfs `local_cpulist` of the GPU's PCI
    function), BEFORE any pinned buffer is allocated: pinned pages are then first-touched on that node, and the
    host-to-device copies of 8 ranks stop sharing one socket's memory controllers (round 1: 44 GB/s per GPU alone,
    21 GB/s with 8 ranks).  Returns a short description for the bench line; never fails the run."""
    try:
        prop = torch.cuda.get_device_properties(local_rank)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        node = open(f"{base}/numa_node").read().strip()
        if os.environ.get("BENCH_NO_NUMA_BIND"):
            return {"gpu_pci": bdf, "numa_node": node, "bound": False, "why": "BENCH_NO_NUMA_BIND"}
        if cpus:
            os.sched_setaffinity(0, cpus)
            torch.set_num_threads(max(1, min(len(cpus), torch.get_num_threads())))
            return {"gpu_pci": bdf, "numa_node": node, "cpus": len(cpus), "bound": True}
        return {"gpu_pci": bdf, "numa_node": node, "bound": False, "why": "no local cpus in this process's affinity mask"}
    except Exception as e:  # noqa: BLE001
        return {"bound": False, "why": f"{type(e).__name__}: {e}"[:120]}


class Ctx:
    """Process-group / device context and the timing helpers every arm of ours uses."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the fusion head has no CPU path (use --impl reference for the CPU baseline)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa = bind_to_gpu_numa_node(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def stage(self, msg):
        if self.rank == 0 and os.environ.get("BENCH_VERBOSE"):
            print(f"[bench] {msg}", file=sys.stderr, flush=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, tail=None):
        """ms per step: CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if tail is not None:
            tail()                       # still inside the timed region
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms) / steps

    def finish(self):
        # Tear-down of a process group whose collectives live inside CUDA graphs can block; everything is
        # measured and printed by now, so synchronise, meet once more and leave without the NCCL destructor.
        sys.stdout.flush(); sys.stderr.flush()
        if self.world > 1:
            torch.cuda.synchronize()
            self.dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)


_EVENT_FLOOR = {"us": None}      # CUDA-event interval of an empty kernel, measured in the profiling pass


def profile_pass(L, ctx, run_once, nprof=3, detail_path=""):
    """Per-kernel-family CUDA-event profile of `nprof` eager passes (rank 0 records): (families, gemm_detail)."""
    ctx.barrier()
    # events cannot be timed inside a graph replay, so the eager launches are profiled -- but issued one by one the
    # device outruns the host and every event interval would contain the host's launch latency (6-8 us on top of the
    # small kernels).  Each profiled pass is therefore enqueued behind a stall kernel that lasts longer than the host
    # needs to enqueue the whole step: the device then runs the step's kernels back to back, as in the replayed graph.
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_once()
    host_us = (time.perf_counter() - t0) * 1e6          # host time to enqueue one eager step
    torch.cuda.synchronize()
    if ctx.rank == 0:
        L.prof_enable(True)
    for _ in range(nprof):             # every rank runs the pass (the steps contain collectives); rank 0 records
        L.prof_stall(1.5 * host_us + 2000.0, torch.cuda.current_device())
        run_once()
        torch.cuda.synchronize()
    if ctx.rank == 0:                  # floor of one event interval: an empty kernel, measured the same way
        L.prof_stall(3000.0, torch.cuda.current_device())
        L.prof_null(32, torch.cuda.current_device())
    ctx.barrier()
    if ctx.rank != 0:
        return {}, {}
    detail = L.prof_report()
    L.prof_enable(False)
    if detail_path:
        with open(detail_path, "w") as f:
            for k, v in sorted(detail.items(), key=lambda kv: -kv[1]["ms"]):
                f.write(f"{k:44s} n/step={v['launches']/nprof:6.1f} ms/step={v['ms']/nprof:8.4f} us/launch={1e3*v['ms']/v['launches']:8.1f} "
                        f"TFLOP/s={v['flops']/max(v['ms'],1e-9)/1e9:8.1f} GB/s={v['bytes']/max(v['ms'],1e-9)/1e6:8.1f}\n")
    null = detail.pop("prof_null", None)
    _EVENT_FLOOR["us"] = (1e3 * null["ms"] / null["launches"]) if null and null["launches"] else None
    gemm_detail = {k: v for k, v in detail.items() if k.startswith("gemm_tc")}
    prof = {}
    for k, v in detail.items():            # aggregate shape-tagged records per kernel family
        fam = prof.setdefault(k.split(":")[0], dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        for kk in fam:
            fam[kk] += v[kk]
    for v in prof.values():
        v["ms_per_step"] = v["ms"] / nprof
        v["launches_per_step"] = v["launches"] / nprof
    return prof, gemm_detail


def roofline_block(prof, gemm_detail, peaks, nprof):
    """`roofline` of the dominant kernel family (all gemm_tc_kernel launches of a step) + per-family table.
    Peak choice (MEASURED_PEAKS.json): a step of a few ms made of ~25 us kernels runs at full clocks and far below the
    power cap -- the BURST bf16 figure is the honest denominator (the sustained one was measured at a 1290 MHz median
    under a 997 W cap); both fractions are reported, `frac` uses burst."""
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
                "achieved": ach, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": ach / peaks["tf_burst"],
                "frac_of_sustained_peak": ach / peaks["tf_sus"],
                "traffic": traffic, "traffic_source": tsrc,
                "algorithmic_flops_per_launch": gflops / nl, "algorithmic_bytes_per_launch": sum(v["bytes"] for v in fam.values()) / nl,
                "avg_launch_us": gms * 1e3 / nl,
                "peak_source": peaks["src"] + " burst bf16 (short kernels at full clocks, no power cap; sustained = "
                               f"{peaks['tf_sus']} TFLOP/s is reported as frac_of_sustained_peak)",
                "launches_per_step": sum(v["launches_per_step"] for v in fam.values()),
                "ms_per_step": sum(v["ms_per_step"] for v in fam.values())}

        # context for `frac`: the family mixes tensor-bound (K = 768), HBM-bound (K = 256) and launch-latency-bound
        # (M = 256) shapes.  (1) time-weighted fraction of each launch's OWN roofline max(flops / tensor peak,
        # algorithmic bytes / HBM peak); (2) the same two numbers over the launches of >= 20 us only.
        def _roof(sel):
            t = sum(v["ms"] for v in sel)
            if t <= 0:
                return None
            t_roof = sum(max(v["flops"] / (peaks["tf_burst"] * 1e12), v["bytes"] / (peaks["hbm"] * 1e9)) * 1e3 for v in sel)
            fl = sum(v["flops"] for v in sel)
            return {"launches_per_step": sum(v["launches"] for v in sel) / nprof, "ms_per_step": t / nprof,
                    "achieved_tflops": fl / (t * 1e-3) / 1e12, "frac_of_tensor_peak": fl / (t * 1e-3) / 1e12 / peaks["tf_burst"],
                    "frac_of_own_roofline": t_roof / t}
        allg = list(gemm_detail.values())
        big = [v for v in allg if v["ms"] / max(v["launches"], 1) >= 0.020]
        roof["all_launches"] = _roof(allg)
        roof["launches_over_20us"] = _roof(big)
        # every record is one event interval around one launch: it contains the launch + drain latency of an isolated
        # kernel and the two event records, which a CUDA-graph replay does not pay per kernel.  The interval of an
        # EMPTY kernel measured the same way is that floor; `frac` above is the raw figure, this one is net of it.
        floor_us = _EVENT_FLOOR["us"]
        if floor_us is not None and gms * 1e3 > nl * floor_us:
            net_ms = gms - nl * floor_us * 1e-3
            roof["event_interval_floor_us"] = floor_us
            roof["frac_net_of_event_floor"] = gflops / (net_ms * 1e-3) / 1e12 / peaks["tf_burst"]
    tot_prof_ms = sum(v["ms_per_step"] for v in prof.values()) or 1.0
    families = {k: {"ms_per_step": round(v["ms_per_step"], 4), "launches_per_step": v["launches_per_step"],
                    "share": round(v["ms_per_step"] / tot_prof_ms, 4),
                    "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 2),
                    "gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1),
                    "frac_of_hbm_peak": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6 / peaks["hbm"], 3)}
                for k, v in sorted(prof.items())}
    return roof, families


# ----------------------------------------------------------------------------------------------------
# our arm: training step (cfg1-cfg4)
# ----------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import mmser_b200
    from mmser_b200 import _lib as L
    from mmser_b200 import synth
    from mmser_b200.parallel import DataParallelHead, GraphedTrainStep

    ctx = Ctx()
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    L.load()
    stage = ctx.stage
    peaks = load_peaks()
    B, Ta, Tt, C = wl["B"], wl["Ta"], wl["Tt"], wl["C"]
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32

    rates = parse_dropout(args.dropout)
    head = mmser_b200.FusionHead(C, dropout=rates).to(dev)
    head.load_group_state(synth.head_weights(C))
    head.train()
    dp = DataParallelHead(head)
    if args.no_overlap:
        dp.reducer.overlap = False

    # per-rank shard of the synthetic global batch (different seed per rank = different samples)
    a, t, am, tm, labels = synth.make_inputs(B, Ta, Tt, C, seed=1234 + rank)
    host = dict(a=a.to(dtype).pin_memory(), t=t.to(dtype).pin_memory(), am=am.pin_memory(), tm=tm.pin_memory(),
                labels=labels.pin_memory())
    devin = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    torch.cuda.synchronize()

    def eager_step(inp):
        return dp.train_step(inp["a"], inp["t"], inp["am"], inp["tm"], inp["labels"])

    def capture(inp, what):
        """CUDA graph of the whole step (forward + backward + loss + NCCL all-reduces in data-parallel runs): one
        cudaGraphLaunch instead of ~100 launches.  If the process group cannot be captured the rank runs eager."""
        if not args.graph:
            return None
        try:
            # the given device tensors ARE the graph's static inputs: a replay is the step, no staging copy
            return GraphedTrainStep(dp, inp["a"], inp["t"], inp["am"], inp["tm"], inp["labels"], static_inputs=True)
        except Exception as e:  # noqa: BLE001
            if world == 1:
                raise
            print(f"[bench rank {rank}] CUDA-graph capture of {what} failed ({type(e).__name__}: {e}); running eager",
                  file=sys.stderr)
            torch.cuda.synchronize()
            return None

    graphed = capture(devin, "the data-parallel step")
    used_graph = graphed is not None

    def step(inp):
        if graphed is not None and inp is devin:
            return graphed.replay()
        return eager_step(inp)

    stage(f"setup done (graph={'yes' if used_graph else 'no'})")
    # ---------------- device-resident timing (value) ----------------
    for _ in range(max(args.warmup, 3)):
        out = step(devin)
    ctx.barrier()
    l0 = L.launch_count()
    eager_step(devin)
    launches_per_step = L.launch_count() - l0          # kernels of ONE step (a graph replay re-issues exactly these)
    ctx.barrier()
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    ms = ctx.timed(lambda: step(devin), args.steps)
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
    # Only the VALID frames cross PCIe: the host holds each modality as [sum(len), 768] packed rows + B + 1 offsets
    # (functional.pack_frames: what a loader that receives per-utterance hidden states has before it pads them), and
    # ser_unpack_frames rebuilds the zero-padded [B, T, 768] tensors and the masks on the device, on the copy stream,
    # straight into the graph's input buffers -- bit-identical to copying the padded tensors (asserted below).
    from mmser_b200.functional import pack_frames, unpack_frames
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]   # double buffer
    packed = os.environ.get("BENCH_E2E_PADDED") is None
    if packed:
        pa, oa = pack_frames(host["a"], host["am"])
        pt, ot = pack_frames(host["t"], host["tm"])
        hostp = dict(pa=pa.pin_memory(), oa=oa.pin_memory(), pt=pt.pin_memory(), ot=ot.pin_memory(), labels=host["labels"])
        stagep = [{k: torch.empty_like(v, device=dev) for k, v in hostp.items() if k != "labels"} for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]        # H2D of buffer i finished (recorded on the copy stream)
    consumed = [torch.cuda.Event(), torch.cuda.Event()]     # compute finished reading buffer i (recorded on the main stream)
    done = [torch.cuda.Event(), torch.cuda.Event()]         # loss of the step on buffer i is in host memory
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    h2d_bytes = sum(v.numel() * v.element_size() for v in (hostp if packed else host).values())
    for ev in consumed:
        ev.record()
    e2e_graphs = None
    if used_graph:
        gs = [capture(b, "the e2e step") for b in bufs]
        e2e_graphs = gs if all(g is not None for g in gs) else None

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i])            # never overwrite a buffer the step is still reading
            if packed:
                for k, v in hostp.items():                 # pinned host -> device: valid frames, offsets, labels
                    (bufs[i] if k == "labels" else stagep[i])[k].copy_(v, non_blocking=True)
                unpack_frames(stagep[i]["pa"], stagep[i]["oa"], Ta, out=bufs[i]["a"], mask_out=bufs[i]["am"])
                unpack_frames(stagep[i]["pt"], stagep[i]["ot"], Tt, out=bufs[i]["t"], mask_out=bufs[i]["tm"])
            else:
                for k, v in host.items():
                    bufs[i][k].copy_(v, non_blocking=True)     # pinned host -> device
            ready[i].record(copy_stream)

    state = {"i": 0, "loss": 0.0, "pending": None}
    prefetch(0)
    if packed:                                             # the rebuilt batch is the padded batch, bit for bit
        copy_stream.synchronize()
        assert all(torch.equal(bufs[0][k], devin[k]) for k in host), "unpack_frames does not reproduce the padded batch"

    def e2e_step():
        i = state["i"]
        torch.cuda.current_stream().wait_event(ready[i])
        o = e2e_graphs[i].replay() if e2e_graphs is not None else eager_step(bufs[i])
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
    ms_e2e = ctx.timed(e2e_step, max(3, args.steps // 2), tail=e2e_tail)
    e2e_graphs = None
    stage(f"e2e timing done: {ms_e2e:.3f} ms/step")

    # ---------------- the step followed by the optimizer step (FusedAdamW, the reference's 10 parameter groups) --------
    ms_opt = None
    if not args.no_optimizer_arm:
        lr = 1e-4
        groups = [dict(params=list(head.adapter_a.parameters()), lr=lr * 0.1, weight_decay=0.025),
                  dict(params=list(head.adapter_t.parameters()), lr=lr * 0.1, weight_decay=0.025),
                  dict(params=list(head.cross.parameters()), lr=lr, weight_decay=0.05),
                  dict(params=list(head.pool_a.parameters()), lr=lr, weight_decay=0.05),
                  dict(params=list(head.pool_t.parameters()), lr=lr, weight_decay=0.05),
                  dict(params=list(head.fusion.parameters()), lr=lr, weight_decay=0.05),
                  dict(params=list(head.classifier.deep_classifier.parameters()), lr=lr * 1.5, weight_decay=0.06),
                  dict(params=list(head.classifier.anchor_clustering.parameters()), lr=lr * 2.0, weight_decay=0.04),
                  dict(params=list(head.classifier.uncertainty_head.parameters()), lr=lr, weight_decay=0.05),
                  dict(params=list(head.prototypes.parameters()), lr=lr, weight_decay=0.05)]     # src/train.py:72-83
        opt = mmser_b200.optim.FusedAdamW(groups, weight_decay=0.05)

        def opt_step():
            step(devin)
            opt.step()

        for _ in range(3):
            opt_step()
        ms_opt = ctx.timed(opt_step, args.steps)
        with torch.no_grad():
            head.load_group_state(synth.head_weights(C))          # back to the benchmark's weights
        del opt
        stage(f"step + optimizer timing done: {ms_opt:.3f} ms/step")

    # ---------------- data-parallel runs: the same step with the all-reduces AFTER the backward (no overlap) ------------
    ms_noov = None
    if world > 1 and not args.no_overlap:
        dp.reducer.overlap = False
        g1 = capture(devin, "the no-overlap step")
        f1 = (lambda: g1.replay()) if g1 is not None else (lambda: eager_step(devin))
        for _ in range(3):
            f1()
        ms_noov = ctx.timed(f1, args.steps)
        dp.reducer.overlap = True
        del g1
        stage(f"no-overlap timing done: {ms_noov:.3f} ms/step")

    # ---------------- the same step with every dropout rate set to 0 (reported beside the headline) ----------------
    ms_nodrop = None
    if any(v > 0 for v in rates.values()):
        head.set_dropout(0.0)
        g0 = capture(devin, "the dropout-free step")
        step0 = (lambda: g0.replay()) if g0 is not None else (lambda: eager_step(devin))
        for _ in range(3):
            step0()
        ms_nodrop = ctx.timed(step0, args.steps)
        head.set_dropout(rates)
        del g0
        stage(f"no-dropout timing done: {ms_nodrop:.3f} ms/step")

    # ---------------- per-kernel-family profile (CUDA events around every launch; separate pass) ----------------
    nprof = 3
    prof, gemm_detail = profile_pass(L, ctx, lambda: eager_step(devin), nprof, args.profile_detail)

    # ---------------- same-box PyTorch eager + cpu baseline (rank 0, N = 1 only) ----------------
    cpu, eager = None, None
    if rank == 0 and world == 1:
        if not args.no_eager_arm:
            graphed = None
            torch.cuda.empty_cache()
            try:
                eager = eager_gpu_measure(wl, rates, dev, steps=5, warmup=3)
            except Exception as e:  # noqa: BLE001
                eager = {"error": f"{type(e).__name__}: {e}"[:300]}
            stage("eager arm done")
        if not args.no_cpu_baseline:
            r = cpu_reference_run(WORKLOADS["cfg1"] | {"C": C}, steps=6, warmup=1, sample_B=8, rates=rates)
            cpu = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank != 0:
        ctx.finish()
        return

    total_B = B * world
    value = total_B / (ms * 1e-3)
    step_flops = 3.0 * fwd_flops_per_sample(Ta, Tt, C) * B             # per GPU, 3x-forward convention
    roof, families = roofline_block(prof, gemm_detail, peaks, nprof)
    if eager and "error" not in eager:
        for v in eager.values():
            if "best_ms" in v:
                v["ours_speedup"] = v["best_ms"] / ms
    line = {
        "metric": METRIC_TRAIN, "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": train_config(args, wl, world, rates, used_graph) | ({"overlap": False} if args.no_overlap else {}),
        "step_model_flops_per_gpu": step_flops,
        "model_tflops_per_gpu": step_flops / (ms * 1e-3) / 1e12,
        "model_frac_of_burst_bf16_peak": step_flops / (ms * 1e-3) / 1e12 / peaks["tf_burst"],
        "model_frac_of_sustained_bf16_peak": step_flops / (ms * 1e-3) / 1e12 / peaks["tf_sus"],
        "no_dropout": None if ms_nodrop is None else {"value": total_B / (ms_nodrop * 1e-3), "unit": "samples/s",
                                                      "ms_per_step": ms_nodrop},
        "no_overlap": None if ms_noov is None else {
            "value": total_B / (ms_noov * 1e-3), "unit": "samples/s", "ms_per_step": ms_noov,
            "what": "gradient all-reduces issued after the backward pass instead of from the per-module backward hooks"},
        "with_optimizer_step": None if ms_opt is None else {
            "value": total_B / (ms_opt * 1e-3), "unit": "samples/s", "ms_per_step": ms_opt,
            "what": "step + FusedAdamW.step() over the reference's 10 parameter groups (src/train.py:72-83)"},
        "roofline": roof, "kernel_families": families,
        "cpu_baseline": cpu,
        "eager_gpu": eager,
        "e2e": {"value": total_B / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "h2d_format": ("packed valid frames + offsets (functional.pack_frames), rebuilt on the device by "
                               "ser_unpack_frames on the copy stream" if packed else "zero-padded tensors"),
                "h2d_bytes_per_step_padded": sum(v.numel() * v.element_size() for v in host.values()),
                "host_numa": ctx.numa,
                "api": "mmser_b200.parallel.GraphedTrainStep / DataParallelHead.train_step(FusionHead) with pinned host inputs; "
                       "one graph per input buffer, loss read back every step and consumed one step late"},
        "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
        "clocks": clocks, "loss": loss_val,
    }
    print(json.dumps(line), flush=True)
    ctx.finish()


# ----------------------------------------------------------------------------------------------------
# our arm: eval path (cfg5)
# ----------------------------------------------------------------------------------------------------
def run_eval(args, wl):
    """BASELINE.json configs[4]: per step, V = 5 test-time-augmentation views of B = 4096 samples go through the head in
    eval mode (OpenMax on, Weibull buffers fitted on a synthetic validation set beforehand), the view logits are
    averaged, the 100-point temperature sweep runs on the validation logits, then /T, softmax, argmax, energy
    (src/eval.py:48-67,174-206; classifier.py:240-275).  The views differ on the audio side only.  Samples shard across
    ranks (replicas only: no collective).  value = samples/s (B per step, each sample = V head forwards)."""
    import mmser_b200
    from mmser_b200 import _lib as L
    from mmser_b200 import functional as SF
    from mmser_b200 import synth

    ctx = Ctx()
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    L.load()
    peaks = load_peaks()
    B, Ta, Tt, C, V = wl["B"], wl["Ta"], wl["Tt"], wl["C"], wl["V"]
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    head = mmser_b200.FusionHead(C, dropout="reference").to(dev)
    head.load_group_state(synth.head_weights(C))
    head.eval()

    a, t, am, tm, labels = synth.make_inputs(B, Ta, Tt, C, seed=1234 + rank)
    g = torch.Generator().manual_seed(99 + rank)
    # TTA views: the reference perturbs the waveform (speed / noise) -> different audio hidden states, same text
    # (one [Ta, 768] perturbation pattern per view, shared by the samples: cheap to generate for 4096 x 250 x 768 inputs)
    views_host = [a.to(dtype).pin_memory()]
    for _ in range(V - 1):
        views_host.append((a + 0.05 * torch.randn(1, Ta, a.shape[2], generator=g) * am[..., None]).to(dtype).pin_memory())
    host = dict(t=t.to(dtype).pin_memory(), am=am.pin_memory(), tm=tm.pin_memory())
    dv = [x.to(dev, non_blocking=True) for x in views_host]
    dd = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    dlabels = labels.to(dev)
    torch.cuda.synchronize()

    with torch.no_grad():
        # validation pass (outside the timed region): fitted Weibull buffers + the logits the sweep is run on
        nfit = head.fit_weibull_on([(dv[0], dd["t"], dd["am"], dd["tm"], dlabels)])
        fused0 = head.features(dv[0], dd["t"], dd["am"], dd["tm"])["fused"]
        val_logits = head.classifier(fused0, use_openmax=False).clone()

    def eval_step(audio_views):
        with torch.no_grad():
            # OpenMax on (eval mode); the text-side work (adapter, q / k / v projections) is done once for the V views
            lv = head.forward_views(audio_views, dd["t"], dd["am"], dd["tm"])
            T = SF.find_optimal_temperature(val_logits, dlabels)      # one launch + one small D2H (the reference: 100 syncs)
            return SF.eval_post(lv, T), T

    for _ in range(max(1, min(args.warmup, 2))):
        post, T = eval_step(dv)
    ctx.barrier()
    l0 = L.launch_count()
    eval_step(dv)
    launches_per_step = L.launch_count() - l0
    steps = max(1, min(args.steps, 10))
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    ms = ctx.timed(lambda: eval_step(dv), steps)
    clocks = sampler.stop() if sampler else None

    # ---- e2e: every step copies the V audio views (+ text, masks) from pinned host memory, view v+1 travelling while
    #      view v is computed, and reads probabilities / predictions / energies back to the host
    from mmser_b200.functional import pack_frames, unpack_frames
    copy_stream = torch.cuda.Stream(device=dev)
    abuf = [torch.empty_like(dv[0]) for _ in range(2)]
    # only the valid frames of every view cross PCIe (functional.pack_frames on the host, ser_unpack_frames on the copy
    # stream rebuilds the zero-padded view in the compute buffer); the text batch and the masks travel as they are
    packed = os.environ.get("BENCH_E2E_PADDED") is None
    if packed:
        offs_host = pack_frames(views_host[0], am)[1].pin_memory()
        views_host = [pack_frames(x, am)[0].pin_memory() for x in views_host]
        pbuf = [torch.empty_like(views_host[0], device=dev) for _ in range(2)]
        offs_dev = offs_host.to(dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    for ev in free:
        ev.record()
    out_host = dict(probs=torch.empty(B, C).pin_memory(), preds=torch.empty(B, dtype=torch.int64).pin_memory(),
                    energy=torch.empty(B).pin_memory())
    h2d = sum(x.numel() * x.element_size() for x in views_host) + sum(v.numel() * v.element_size() for v in host.values())
    d2h = sum(v.numel() * v.element_size() for v in out_host.values())

    def fetch(v, slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[slot])
            if packed:
                pbuf[slot].copy_(views_host[v], non_blocking=True)
                unpack_frames(pbuf[slot], offs_dev, Ta, out=abuf[slot], with_mask=False)
            else:
                abuf[slot].copy_(views_host[v], non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_step():
        with torch.no_grad():
            with torch.cuda.stream(copy_stream):
                for k, v in host.items():
                    dd[k].copy_(v, non_blocking=True)
            fetch(0, 0)
            torch.cuda.current_stream().wait_stream(copy_stream)
            def views():                   # view v + 1 travels while view v is computed
                for v in range(V):
                    slot = v & 1
                    if v + 1 < V:
                        fetch(v + 1, 1 - slot)
                    torch.cuda.current_stream().wait_event(ready[slot])
                    yield abuf[slot]
                    free[slot].record()
            lv = head.forward_views(views(), dd["t"], dd["am"], dd["tm"])
            Tn = SF.find_optimal_temperature(val_logits, dlabels)
            p = SF.eval_post(lv, Tn)
            for k, v in out_host.items():
                v.copy_(p[k], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    e2e_step()
    ms_e2e = ctx.timed(e2e_step, max(1, steps // 2))

    nprof = 1
    prof, gemm_detail = profile_pass(L, ctx, lambda: eval_step(dv), nprof, args.profile_detail)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_eval_run(wl, 32, V, steps=1)
        cpu = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    if rank != 0:
        ctx.finish()
        return
    total_B = B * world
    flops = V * fwd_flops_per_sample(Ta, Tt, C) * B
    roof, families = roofline_block(prof, gemm_detail, peaks, nprof)
    line = {
        "metric": METRIC_EVAL, "value": total_B / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": steps,
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": args.workload, "shape": wl["desc"], "global_batch": total_B, "per_gpu_batch": B, "views": V,
                   "head_forwards_per_step": total_B * V, "parallelism": f"replicas x{world}", "weibull_fit_samples": nfit,
                   "l2": "no explicit flush: one step touches > 50 GB of activations per GPU"},
        "temperature": T, "pred_histogram": torch.bincount(post["preds"].cpu(), minlength=C).tolist(),
        "step_model_flops_per_gpu": flops, "model_tflops_per_gpu": flops / (ms * 1e-3) / 1e12,
        "model_frac_of_burst_bf16_peak": flops / (ms * 1e-3) / 1e12 / peaks["tf_burst"],
        "roofline": roof, "kernel_families": families, "cpu_baseline": cpu,
        "e2e": {"value": total_B / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "api": "FusionHead.eval().forward_views(audio views, t, masks) with pinned host inputs (view v+1 copied while "
                       "view v runs; text-side work once per batch) -> functional.find_optimal_temperature / eval_post -> "
                       "probs, preds, energies read back"},
        "gpu_launches": launches_per_step * steps, "gpu_launches_per_step": launches_per_step, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    ctx.finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dropout", default="reference",
                    help="'reference' = the rates of the reference's training script (cross 0.1, fusion 0.1, classifier 0.15; "
                         "SURVEY.md 8(d)); or one rate for every nn.Dropout of the head, e.g. 0")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager_gpu"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-arm", action="store_true", help="skip the same-box PyTorch-eager measurement of the main line")
    ap.add_argument("--no-optimizer-arm", action="store_true", help="skip the step + FusedAdamW.step() measurement")
    ap.add_argument("--no-overlap", action="store_true",
                    help="data parallel: issue every gradient all-reduce after the backward pass (the main line of a "
                         "default data-parallel run reports this variant as `no_overlap` anyway)")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--profile-detail", default="", help="write the per-shape kernel table of the profiling pass here")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch > 0:
        wl["B"] = args.batch
    if args.impl == "reference":
        run_reference(args, wl)
    elif args.impl == "eager_gpu":
        run_eager(args, wl)
    elif args.workload == "cfg5":
        run_eval(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
