"""Optimizer step for the head's parameters.

`FusedAdamW` takes the place of `torch.optim.AdamW` in the reference's training loops (src/train.py:72-83,169-177;
train_crema.py:206-226): same constructor (parameters or parameter groups with per-group lr / weight_decay), same
update rule (decoupled weight decay, bias correction), a `torch.optim.Optimizer` subclass so `LambdaLR`, `state_dict()`
and `zero_grad()` work unchanged.  One multi-tensor CUDA launch per parameter group (24 tensors per launch) instead
of torch's per-op foreach kernels.  `clip_grad_norm_` is the fused counterpart of torch.nn.utils.clip_grad_norm_
(global L2 norm over all groups): it leaves the clip coefficient on the device and the next `step()` multiplies it into
the gradients inside the update kernel -- no pass over the gradients just to scale them, no host synchronisation.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L


def _tables(tensors):
    n = len(tensors)
    return (C.c_void_p * n)(*[t.data_ptr() for t in tensors])


class FusedAdamW(torch.optim.Optimizer):
    # torch.amp.GradScaler.step() hands `grad_scale` / `found_inf` (device tensors, set as attributes) to optimizers
    # that declare this, instead of reading found_inf back to the host and skipping the call (src/train.py:88,169-177)
    _step_supports_amp_scaling = True

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0) or weight_decay < 0.0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._clip_coef: Optional[torch.Tensor] = None      # device scalar set by clip_grad_norm_, consumed by step()
        self._amp_steps: Optional[torch.Tensor] = None      # [n_groups] float32 device step counts (AMP path only)

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """Global gradient-norm clipping over every parameter of every group.  Returns the total norm (device scalar,
        like torch.nn.utils.clip_grad_norm_).  Deviation from torch: `.grad` is left unscaled -- the coefficient stays
        on the device and the NEXT step() multiplies it into the gradients inside the update kernel.  It is bound to
        the gradients it was computed from: zero_grad() (or a skipped step followed by a new backward, which goes
        through zero_grad()) discards it, so it can never leak into a later step that did not clip."""
        grads = [p.grad for g in self.param_groups for p in g["params"] if p.grad is not None]
        if not grads:
            return torch.zeros(())
        for t in grads:
            L.require_cuda(t)
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise L.SerError("FusedAdamW: gradients must be contiguous float32 CUDA tensors")
        dev = grads[0].device
        buf = torch.empty(3, device=dev, dtype=torch.float32)           # scratch | coefficient | norm
        lib = L.load()
        cnt = (C.c_longlong * len(grads))(*[t.numel() for t in grads])
        L.check(lib.ser_grad_clip_coef(len(grads), _tables(grads), cnt, float(max_norm), buf[0:1].data_ptr(),
                                       buf[1:2].data_ptr(), buf[2:3].data_ptr(), L.stream_ptr(dev)), "ser_grad_clip_coef")
        self._clip_coef = buf[1:2]
        return buf[2]

    def zero_grad(self, set_to_none: bool = True) -> None:
        self._clip_coef = None              # a clip coefficient belongs to the gradients it was computed from
        super().zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        coef = self._clip_coef
        self._clip_coef = None
        # set by torch.amp.GradScaler.step() around this call (see _step_supports_amp_scaling)
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        amp = found_inf is not None or grad_scale is not None or self._amp_steps is not None
        work = []
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if params:
                work.append((gi, group, params, self._group_cache(gi, group, params)))
        if not work:
            return loss
        if amp:
            dev = work[0][2][0].device
            if self._amp_steps is None or self._amp_steps.device != dev:
                # from here on the step counts live on the device: a skipped step (found_inf) must not advance them and
                # the host cannot see found_inf without a synchronisation
                host = [0.0] * len(self.param_groups)
                for gi, _, _, cache in work:
                    host[gi] = float(cache["step"])
                self._amp_steps = torch.tensor(host, dtype=torch.float32, device=dev)
            key = tuple(gi for gi, *_ in work)
            masks = self.__dict__.setdefault("_amp_active", {})
            inc = masks.get(key)
            if inc is None or inc.device != self._amp_steps.device:
                active = torch.zeros(len(self.param_groups), dtype=torch.float32)
                active[list(key)] = 1.0
                inc = masks[key] = active.to(self._amp_steps.device)
            if found_inf is not None:
                inc = inc * (1.0 - found_inf.to(torch.float32).reshape(()).clamp(0.0, 1.0))
            self._amp_steps.add_(inc)
        for gi, group, params, cache in work:
            grads = [p.grad for p in params]
            for t in (grads[0], grads[-1]):
                if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                    raise L.SerError("FusedAdamW: gradients must be contiguous float32 CUDA tensors")
            b1, b2 = group["betas"]
            stream = L.stream_ptr(params[0].device)
            if amp:
                L.check(lib.ser_adamw_multi_amp(len(params), cache["p"], _tables(grads), cache["m"], cache["v"], cache["cnt"],
                                                float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                                float(group["weight_decay"]), self._amp_steps[gi:gi + 1].data_ptr(),
                                                None if coef is None else coef.data_ptr(),
                                                None if grad_scale is None else grad_scale.to(torch.float32).reshape(1).data_ptr(),
                                                None if found_inf is None else found_inf.to(torch.float32).reshape(1).data_ptr(),
                                                stream), "ser_adamw_multi_amp")
            else:
                cache["step"] += 1
                step = cache["step"]
                for p in params:                      # torch's state_dict format: the count lives with every parameter
                    self.state[p]["step"] = step
                L.check(lib.ser_adamw_multi(len(params), cache["p"], _tables(grads), cache["m"], cache["v"], cache["cnt"],
                                            float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                            float(group["weight_decay"]), step,
                                            None if coef is None else coef.data_ptr(), stream), "ser_adamw_multi")
            # the kernel wrote the parameters behind autograd's back: bump their version counters like an in-place
            # torch op would (consumers such as FlatParams' cached bf16 operand copies key on them)
            torch.autograd.graph.increment_version(params)
        return loss

    def _group_cache(self, gi, group, params):
        """Pointer tables of a group's parameters and moment buffers: stable across steps (only the gradient table is
        rebuilt per step -- autograd hands out fresh gradient tensors), rebuilt when the parameter set or any
        parameter's storage changes (the key is the full tuple of data pointers).  The bias-correction step count is
        re-read from the per-parameter state, which step() keeps current, so a rebuild never restarts it."""
        caches = self.__dict__.setdefault("_caches", {})
        c = caches.get(gi)
        sig = tuple(p.data_ptr() for p in params)
        if c is not None and c["sig"] == sig:
            return c
        step0 = 0
        for p in params:
            L.require_cuda(p)
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise L.SerError("FusedAdamW: parameters must be contiguous float32 (the head keeps fp32 masters)")
            st = self.state[p]
            if "exp_avg" not in st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            step0 = max(step0, int(st["step"]))
        c = dict(sig=sig, params=params, step=step0, p=_tables(params),
                 m=_tables([self.state[p]["exp_avg"] for p in params]),
                 v=_tables([self.state[p]["exp_avg_sq"] for p in params]),
                 cnt=(C.c_longlong * len(params))(*[p.numel() for p in params]))
        caches[gi] = c
        return c

    def _sync_amp_steps(self) -> None:
        """AMP path: the step counts live on the device; copy them into the per-parameter state (one small D2H copy)."""
        if self._amp_steps is None:
            return
        host = self._amp_steps.cpu().tolist()
        for gi, c in self.__dict__.get("_caches", {}).items():
            c["step"] = int(round(host[gi]))
            for p in c["params"]:
                self.state[p]["step"] = c["step"]

    def state_dict(self):
        self._sync_amp_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for st in self.state.values():                # torch.optim.AdamW checkpoints keep `step` as a tensor
            if torch.is_tensor(st.get("step")):
                st["step"] = int(st["step"].item())
        self.__dict__["_caches"] = {}
        self._amp_steps = None


def clip_grad_norm_(optimizer: FusedAdamW, max_norm: float) -> torch.Tensor:
    """Function form, mirroring torch.nn.utils.clip_grad_norm_(parameters, max_norm) for a FusedAdamW optimizer."""
    return optimizer.clip_grad_norm_(max_norm)
