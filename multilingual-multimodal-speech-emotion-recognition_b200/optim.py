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
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0) or weight_decay < 0.0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._clip_coef: Optional[torch.Tensor] = None      # device scalar set by clip_grad_norm_, consumed by step()

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """Global gradient-norm clipping over every parameter of every group.  Returns the total norm (device scalar,
        like torch.nn.utils.clip_grad_norm_); the scaling itself is applied by the next step()."""
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

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        coef = self._clip_coef
        self._clip_coef = None
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            cache = self._group_cache(gi, group, params)
            cache["step"] += 1
            grads = [p.grad for p in params]
            for t in (grads[0], grads[-1]):
                if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                    raise L.SerError("FusedAdamW: gradients must be contiguous float32 CUDA tensors")
            b1, b2 = group["betas"]
            L.check(lib.ser_adamw_multi(len(params), cache["p"], _tables(grads), cache["m"], cache["v"], cache["cnt"],
                                        float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                        float(group["weight_decay"]), cache["step"],
                                        None if coef is None else coef.data_ptr(), L.stream_ptr(params[0].device)),
                    "ser_adamw_multi")
            # the kernel wrote the parameters behind autograd's back: bump their version counters like an in-place
            # torch op would (consumers such as FlatParams' cached bf16 operand copies key on them)
            torch.autograd.graph.increment_version(params)
        return loss

    def _group_cache(self, gi, group, params):
        """Pointer tables of a group's parameters and moment buffers: stable across steps (only the gradient table is
        rebuilt per step -- autograd hands out fresh gradient tensors), rebuilt when the parameter set or storage moves."""
        caches = self.__dict__.setdefault("_caches", {})
        c = caches.get(gi)
        sig = (len(params), params[0].data_ptr(), params[-1].data_ptr())
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

    def state_dict(self):
        # the per-parameter step counters of torch's format are kept per group while running; write them back first
        for c in self.__dict__.get("_caches", {}).values():
            for p in c["params"]:
                self.state[p]["step"] = c["step"]
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self.__dict__["_caches"] = {}


def clip_grad_norm_(optimizer: FusedAdamW, max_norm: float) -> torch.Tensor:
    """Function form, mirroring torch.nn.utils.clip_grad_norm_(parameters, max_norm) for a FusedAdamW optimizer."""
    return optimizer.clip_grad_norm_(max_norm)
