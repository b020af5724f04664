"""Bottleneck adapter of the two encoders (reference: src/models/audio_encoder.py:19-21,112 and
src/models/text_encoder.py:17-19,57):  seq = seq + Linear(768->256) . ReLU . Linear(256->768)(seq).

`BottleneckAdapter` is an nn.Sequential with the reference's child indices (0, 2), so it can be assigned to
`encoder.adapter` and keeps the checkpoint keys.  Calling it returns the branch only (what the encoders
add to `seq`); `residual_forward` returns seq + branch in one fused pass.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .._params import FlatParams
from ..functional import AdapterFn
from ._common import Linear


class BottleneckAdapter(nn.Sequential):
    def __init__(self, hidden: int = 768, adapter_dim: int = 256):
        super().__init__(Linear(hidden, adapter_dim), nn.ReLU(), Linear(adapter_dim, hidden))
        self._flat = FlatParams([(n, p) for n, p in self.named_parameters()])

    def forward(self, seq: torch.Tensor) -> torch.Tensor:
        return AdapterFn.apply(seq, self._flat, False, *self._flat.params)

    def residual_forward(self, seq: torch.Tensor) -> torch.Tensor:
        return AdapterFn.apply(seq, self._flat, True, *self._flat.params)
