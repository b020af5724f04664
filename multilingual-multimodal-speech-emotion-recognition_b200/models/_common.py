"""Shared pieces of the drop-in modules."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib
from ..functional import LayerNormFn, LinearFn


class DropoutSeed:
    """Seed source of one module's in-kernel dropout (csrc/dropout.cuh).

    The kernels read the seed from device memory, so `next()` hands out a fresh 1-element int64 CUDA tensor per
    forward (kept by the autograd node for the backward) and advances the module's counter with an in-stream add:
    a captured CUDA graph therefore draws new masks on every replay.  The starting value is drawn from torch's
    default CPU generator at first use (so torch.manual_seed() makes runs reproducible and no two modules share a
    mask stream) and mixed with the data-parallel rank.  The first seed must be created outside graph capture
    (any eager warm-up step does)."""

    def __init__(self):
        self.counter = None

    def next(self, device) -> torch.Tensor:
        if self.counter is None or self.counter.device != device:
            rank = torch.distributed.get_rank() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 0
            base = int(torch.randint(0, 2 ** 62, (1,)).item()) ^ (rank * 0x0CB92BA72F3D8DD7)
            self.counter = torch.full((1,), base & 0x7FFFFFFFFFFFFFFF, dtype=torch.int64, device=device)
        seed = self.counter.clone()
        self.counter.add_(0x2545F4914F6CDD1D)      # wraps modulo 2^64; the kernels hash all 64 bits
        return seed

    def __deepcopy__(self, memo):
        return DropoutSeed()


def active_dropout(module: nn.Module, p: float) -> float:
    """nn.Dropout semantics: active only in training mode."""
    return float(p) if (module.training and p > 0.0) else 0.0


class Linear(nn.Linear):
    """nn.Linear whose forward runs the library GEMM (so children stay individually callable on CUDA)."""

    def forward(self, x):
        return LinearFn.apply(x, self.weight, self.bias)


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        if not self.elementwise_affine or len(self.normalized_shape) != 1:
            raise _lib.SerError("only affine LayerNorm over the last dimension is supported")
        return LayerNormFn.apply(x, self.weight, self.bias)
