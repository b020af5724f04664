"""Batch data parallelism for the fusion head: one process per GPU, torch.distributed (NCCL over NVLink).

The path shards along the batch only: every op up to the logits is per-sample, so each rank
runs the full head on its slice and the only exchanges are
  1. one asynchronous SUM all-reduce per module of its contiguous fp32 gradient buffer, issued from the module's
     backward hook the moment its gradients are final -- the classifier stack (75.6 MB, first to finish in
     backward) is in flight while fusion / pooling / attention / adapter backward still run;
  2. two tiny forward-time all-reduces that make the loss EXACTLY the global-batch loss of the single-process
     reference: per-class label counts (focal class weights use batch-global counts, losses.py:45) and the raw
     loss sums (so the uncertainty term mean(unc)*mean(correct), every mean's divisor and the non-finite guards
     are evaluated on global values and all ranks take the same branch).
Gradients are pre-divided by the GLOBAL batch inside the loss kernel, so SUM (not AVG) is the right reduction.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.distributed as dist


class GradBucketReducer:
    """Launches one async all-reduce per ready bucket and settles them at the end of the step.
    Backend agnostic (gloo on CPU in the tests, NCCL on the GPUs)."""

    # buckets below this size are latency-bound (RING/LL at 8 GPUs: ~35 us each, 12-31 channels): they are held
    # back and sent as ONE coalesced NCCL group call at the end of the step instead of one launch each
    SMALL_BYTES = 4 << 20

    def __init__(self, group=None):
        self.group = group
        # False: every bucket is held back and all-reduced after the backward pass (one coalesced call) -- the
        # "without overlap" arm of the benchmark
        self.overlap = True
        self.pending: List = []
        self.deferred: List[torch.Tensor] = []
        self.order: List[str] = []          # bucket names in launch order (for tests / logging)

    def _active(self) -> bool:
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def reduce_async(self, name: str, flat: torch.Tensor) -> None:
        self.order.append(name)
        if not self._active():
            return
        if not self.overlap or flat.numel() * flat.element_size() < self.SMALL_BYTES:
            self.deferred.append(flat)
            return
        self.pending.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> None:
        if self.deferred:
            small, self.deferred = self.deferred, []
            cm = getattr(dist, "_coalescing_manager", None)
            done = False
            if cm is not None and small[0].is_cuda:
                try:
                    with cm(group=self.group, device=small[0].device, async_ops=True) as work:
                        for t in small:
                            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                    self.pending.append(work)
                    done = True
                except (TypeError, RuntimeError):
                    done = False
            if not done:
                for t in small:
                    self.pending.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in self.pending:
            w.wait()
        self.pending.clear()

    def reset(self) -> None:
        self.order.clear()


def global_loss_cfg(labels: torch.Tensor, num_classes: int, group=None) -> Dict:
    """loss_cfg entries that turn the per-shard loss kernels into the exact global-batch loss.  Every rank must hold
    the SAME number of samples: B_global = B * world is a host-side divisor baked into the loss descriptor (a
    DistributedSampler with drop_last gives exactly that)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return dict(B_global=int(labels.shape[0]))        # the loss kernel counts the labels itself
    # scatter_add instead of torch.bincount: no device->host sync, CUDA-graph capturable
    # (out-of-range labels -- an error in the reference, losses.py:45 -- are not counted, like the single-process kernel)
    counts = torch.zeros(num_classes, device=labels.device, dtype=torch.float32)
    inside = ((labels >= 0) & (labels < num_classes)).to(torch.float32)
    counts.scatter_add_(0, labels.clamp(0, num_classes - 1), inside)
    # the label counts depend on the labels only: their all-reduce travels during the forward pass and is awaited by
    # the loss (HeadLossFn calls cfg["before_loss"] first)
    work = dist.all_reduce(counts, group=group, async_op=True)

    def reduce_sums(sums: torch.Tensor) -> None:
        if world > 1:
            dist.all_reduce(sums, group=group)

    return dict(counts=counts, B_global=int(labels.shape[0]) * world, all_reduce=reduce_sums, before_loss=work.wait)


class DataParallelHead:
    """Wraps a FusionHead: same call, gradients all-reduced bucket by bucket during backward."""

    def __init__(self, head, group=None, broadcast: bool = True):
        self.head, self.group = head, group
        self.reducer = GradBucketReducer(group)
        self._flats = []
        for name in head.GROUPS:
            mod = getattr(head, name)
            fp = getattr(mod, "_flat", None)
            if fp is None:
                continue
            fp.ensure()
            fp.grad_hook = self._make_hook(name)
            self._flats.append((name, fp))
        self._last: Dict[str, torch.Tensor] = {}
        # train_step() always starts from zero_grad(set_to_none=True), so the gradient buffers of consecutive steps may
        # share one persistent allocation (never true for callers that accumulate gradients over several backwards)
        head.persistent_grad_arena = True
        if dist.is_initialized() and dist.get_world_size(group) > 1 and next(head.parameters()).is_cuda:
            # SMs left to the NCCL kernels that run beside the backward GEMMs (see csrc/gemm_simt.cu set_reserved_sms);
            # SER_SM_RESERVE overrides (0 = none)
            import os
            from . import _lib
            _lib.set_reserved_sms(int(os.environ.get("SER_SM_RESERVE", "0")))
        if broadcast and dist.is_initialized() and dist.get_world_size(group) > 1:
            for _, fp in self._flats:
                dist.broadcast(fp.flat, src=0, group=group)
            dist.broadcast(head.prototypes.prototypes.data, src=0, group=group)

    def _make_hook(self, name: str) -> Callable:
        def hook(fp, gflat):
            self._last[name] = gflat
            self.reducer.reduce_async(name, gflat)
        return hook

    def train_step(self, a_hid, t_hid, a_mask, t_mask, labels, loss_cfg: Optional[dict] = None):
        """forward + backward + gradient all-reduce; returns the head's output dict (loss terms are global)."""
        self.reducer.reset()
        self.head.zero_grad(set_to_none=True)      # gradients of a step are reduced in place; never accumulate across steps
        cfg = global_loss_cfg(labels, self.head.num_labels, self.group)
        if loss_cfg:
            cfg.update(loss_cfg)
        out = self.head(a_hid, t_hid, a_mask, t_mask, labels, loss_cfg=cfg)
        out["loss"].backward()
        pg = self.head.prototypes.prototypes.grad
        if pg is not None:
            self.reducer.reduce_async("prototypes", pg)
        self.reducer.finish()
        # autograd normally adopts our gradient views as .grad (no copy); if it cloned instead, refresh the clone
        for name, fp in self._flats:
            g = self._last.get(name)
            if g is None:
                continue
            base = g.data_ptr()
            for p, o in zip(fp.params, fp.offsets):
                if p.grad is not None and p.grad.data_ptr() != base + 4 * o:
                    p.grad.copy_(g[o:o + p.numel()].view(p.shape))
        return out


class GraphedTrainStep:
    """CUDA-graph replay of DataParallelHead.train_step for fixed shapes (single GPU, or per rank when the
    process group supports capture).  The ~600 small launches of a step are recorded once and replayed with
    one cudaGraphLaunch, which removes the CPU launch bound of the 35-block classifier.

        g = GraphedTrainStep(dp, a, t, a_mask, t_mask, labels)      # warm-up + capture
        out = g(a, t, a_mask, t_mask, labels)                        # copy into static inputs, replay
    Outputs (loss terms, logits, ...) are static tensors overwritten by every replay.  Parameter gradients are written
    into the head's PERSISTENT gradient arena (allocated once, outside capture: DataParallelHead switches it on, see
    FlatParams.shared_grad_arena), so every graph captured over one head (bench.py keeps one per input buffer) writes the
    same `.grad` storage and an optimizer step between replays of any of them works as usual.
    """

    def __init__(self, dp: DataParallelHead, a, t, a_mask, t_mask, labels, warmup: int = 3,
                 static_inputs: bool = False):
        """static_inputs=True adopts the given tensors as the graph's input buffers (no private copies): the caller
        refills them in place -- e.g. with host-to-device copies -- and calls replay()."""
        self.dp = dp
        self.static = [x if (static_inputs or x is None) else x.clone() for x in (a, t, a_mask, t_mask, labels)]
        side = torch.cuda.Stream(device=a.device)
        side.wait_stream(torch.cuda.current_stream(a.device))
        with torch.cuda.stream(side):                       # warm-up on a side stream, as torch.cuda.graphs requires
            for _ in range(warmup):
                dp.train_step(*self.static)
        torch.cuda.current_stream(a.device).wait_stream(side)
        torch.cuda.synchronize(a.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = dp.train_step(*self.static)

    def load_inputs(self, a, t, a_mask, t_mask, labels, non_blocking: bool = True) -> None:
        for dst, src in zip(self.static, (a, t, a_mask, t_mask, labels)):
            if dst is not None and src is not None and dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=non_blocking)

    def replay(self):
        self.graph.replay()
        return self.out

    def __call__(self, a, t, a_mask, t_mask, labels):
        self.load_inputs(a, t, a_mask, t_mask, labels)
        return self.replay()
