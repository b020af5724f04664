"""Flat parameter storage for the drop-in modules.

Each module keeps its nn.Parameters exactly as the reference names them (state_dict compatible), but
re-points their storage into ONE contiguous fp32 buffer.  That gives
  * one `ser_cast` launch per forward to produce the bf16 operand copies of every weight,
  * packed weights (q|k|v rows) as plain views, no concatenation,
  * a contiguous gradient buffer for bucketed NCCL all-reduce in data-parallel runs.
Parameters stay ordinary fp32 leaf tensors: optimisers, load_state_dict, .grad all work unchanged.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib

ALIGN = 64  # elements; keeps every tensor 128-byte (bf16) / 256-byte (fp32) aligned


class FlatParams:
    def __init__(self, named: Sequence[Tuple[str, nn.Parameter]]):
        self.names: List[str] = [n for n, _ in named]
        self.params: List[nn.Parameter] = [p for _, p in named]
        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            n = p.numel()
            off += n if n % ALIGN == 0 else n + (ALIGN - n % ALIGN)
        self.total = off
        self.flat: torch.Tensor | None = None
        self._lowp: Dict[torch.dtype, torch.Tensor] = {}
        self._fresh: Dict[torch.dtype, tuple] = {}     # set by precast(): parameter versions the low-precision copy holds
        self._next_grad: torch.Tensor | None = None    # set by shared_grad_arena(): this module's slice of one zero-filled arena
        self.index = {n: i for i, n in enumerate(self.names)}
        # data-parallel hook: called as hook(flat_params, flat_grad_buffer) at the end of the module's backward,
        # i.e. as soon as this module's gradients are final (see parallel.py)
        self.grad_hook = None

    # ---------------------------------------------------------------------------------------------
    def ensure(self) -> None:
        """(Re)build the flat buffer if the parameters moved (e.g. after module.to(device))."""
        p0 = self.params[0]
        if not p0.is_cuda:
            raise _lib.SerError("the B200 fusion head has no CPU path: move the module to a CUDA device")
        if p0.dtype != torch.float32:
            raise _lib.SerError("module parameters must stay float32 (masters); feed bfloat16 inputs to select "
                                "the bf16 tensor-core tier")
        if self.flat is not None and self.flat.device == p0.device:
            base = self.flat.data_ptr()
            if all(p.data_ptr() == base + 4 * o for p, o in zip(self.params, self.offsets)):
                return
        flat = torch.zeros(self.total, device=p0.device, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                v = flat[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
        self.flat = flat
        self._lowp = {}

    def view(self, flat: torch.Tensor, name: str, count: int = 1) -> torch.Tensor:
        """View of `name` inside a flat buffer; count > 1 spans that many consecutive same-shape params
        stacked along dim 0 (packed q|k|v)."""
        i = self.index[name]
        p = self.params[i]
        n = p.numel() * count
        shape = (p.shape[0] * count,) + tuple(p.shape[1:]) if p.dim() > 0 else ()
        if count > 1:
            for j in range(1, count):
                assert self.offsets[i + j] == self.offsets[i] + j * p.numel(), "packed params must be adjacent"
        return flat[self.offsets[i]:self.offsets[i] + n].view(shape)

    def compute_copy(self, dtype: torch.dtype) -> torch.Tensor:
        """Flat buffer in the compute dtype: the fp32 masters themselves, or a fresh bf16 cast (one launch)."""
        self.ensure()
        if dtype == torch.float32:
            return self.flat
        buf = self._lowp_buffer(dtype)
        # already cast for this forward by precast() (one launch for all modules) -- unless a parameter was
        # modified in between (optimizer step, load_state_dict), which bumps its version counter
        stamp = self._fresh.pop(dtype, None)
        if stamp is not None and stamp == self._versions():
            return buf
        lib = _lib.load()
        _lib.check(lib.ser_cast(self.flat.data_ptr(), 1, buf.data_ptr(), 0, self.total,
                                _lib.stream_ptr(self.flat.device)), "ser_cast")
        return buf

    def _versions(self):
        return tuple(p._version for p in self.params)

    def _lowp_buffer(self, dtype: torch.dtype) -> torch.Tensor:
        buf = self._lowp.get(dtype)
        if buf is None or buf.device != self.flat.device:
            buf = torch.empty(self.total, device=self.flat.device, dtype=dtype)
            self._lowp[dtype] = buf
        return buf

    @staticmethod
    def precast(flats: Sequence["FlatParams"], dtype: torch.dtype) -> None:
        """Produce the bf16 operand copies of several modules with ONE launch (ser_cast_multi); each module's next
        compute_copy(dtype) then returns its buffer without casting again.  No-op for float32."""
        if dtype == torch.float32 or not flats:
            return
        for fp in flats:
            fp.ensure()
        lib = _lib.load()
        dev = flats[0].flat.device
        for i in range(0, len(flats), 16):
            chunk = flats[i:i + 16]
            bufs = [fp._lowp_buffer(dtype) for fp in chunk]
            n = len(chunk)
            src = (C.c_void_p * n)(*[fp.flat.data_ptr() for fp in chunk])
            dst = (C.c_void_p * n)(*[b.data_ptr() for b in bufs])
            cnt = (C.c_longlong * n)(*[fp.total for fp in chunk])
            _lib.check(lib.ser_cast_multi(n, src, dst, cnt, _lib.stream_ptr(dev)), "ser_cast_multi")
            for fp in chunk:
                fp._fresh[dtype] = fp._versions()

    def new_grad_buffer(self) -> torch.Tensor:
        """Zero-filled: the backward descriptors are filled with grads_zeroed = 1, so the library neither zeroes its
        split-K / atomically accumulated outputs itself nor touches the alignment padding between parameters."""
        g, self._next_grad = self._next_grad, None
        if g is not None and g.device == self.flat.device:
            return g
        return torch.zeros(self.total, device=self.flat.device, dtype=torch.float32)

    @staticmethod
    def shared_grad_arena(flats: Sequence["FlatParams"], holder=None, zero: bool = True):
        """One zero-filled allocation (one fill kernel) for the gradient buffers of several modules' next backward:
        every module's new_grad_buffer() then returns its slice.  Inside a captured CUDA graph every node costs a few
        microseconds of serialisation, so seven fills become one.
        `holder` (any object): keep ONE arena on it and re-zero it every step instead of allocating -- `.grad` storage is
        then the same for every step and every captured graph.  Only for callers that never accumulate gradients over
        several backward passes (the previous step's `.grad` views alias the new ones).
        `zero=False` (persistent arena only): the caller fills the returned arena with zeros itself before the backward
        pass starts (FusionHead does it on a side stream, beside the latency-bound fusion chain)."""
        if not flats:
            return None
        for fp in flats:
            fp.ensure()
        dev = flats[0].flat.device
        n = sum(fp.total for fp in flats)
        if holder is not None:
            arena = holder.__dict__.get("_grad_arena")
            if arena is None or arena.numel() != n or arena.device != dev:
                arena = torch.empty(n, device=dev, dtype=torch.float32)
                holder.__dict__["_grad_arena"] = arena
            if zero:
                arena.zero_()
        else:
            arena = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        for fp in flats:
            fp._next_grad = arena[off:off + fp.total]
            off += fp.total
        return arena

    def grads_from(self, gflat: torch.Tensor) -> List[torch.Tensor]:
        if self.grad_hook is not None:
            self.grad_hook(self, gflat)
        return [gflat[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]
