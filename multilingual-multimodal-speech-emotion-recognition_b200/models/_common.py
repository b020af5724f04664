"""Shared pieces of the drop-in modules."""
from __future__ import annotations

import torch.nn as nn

from .. import _lib
from ..functional import LayerNormFn, LinearFn


def check_dropout(module: nn.Module, p: float, where: str) -> None:
    """In-kernel dropout is not implemented yet: the fused path is exact only with dropout inactive
    (p == 0 or module.eval()).  Fail loudly instead of silently training without regularisation."""
    if module.training and p > 0.0:
        raise NotImplementedError(
            f"{where}: dropout p={p} in training mode is not supported by the fused CUDA path yet; "
            "construct the module with dropout=0.0 or call .eval() (parity tests do exactly that, "
            "SURVEY.md section 7 'Dropout')")


class Linear(nn.Linear):
    """nn.Linear whose forward runs the library GEMM (so children stay individually callable on CUDA)."""

    def forward(self, x):
        return LinearFn.apply(x, self.weight, self.bias)


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        if not self.elementwise_affine or len(self.normalized_shape) != 1:
            raise _lib.SerError("only affine LayerNorm over the last dimension is supported")
        return LayerNormFn.apply(x, self.weight, self.bias)
