"""Drop-in AttentiveStatsPooling (reference: src/models/pooling.py:6-28) backed by the sm_100a kernels."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .._params import FlatParams
from ..functional import AttentiveStatsPoolingFn
from ._common import Linear


class AttentiveStatsPooling(nn.Module):
    """Masked attentive mean + std over time: [B, T, D] -> [B, 2D].

    Same constructor, forward signature and state_dict keys (attention.0.*, attention.2.*) as the reference.
    """

    def __init__(self, input_dim: int, hidden_dim: int = 128):
        super().__init__()
        self.attention = nn.Sequential(Linear(input_dim, hidden_dim), nn.Tanh(), Linear(hidden_dim, 1))
        self._flat = FlatParams([(n, p) for n, p in self.named_parameters()])

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._flat._out_buffer = self.__dict__.pop("_out_buffer", None)      # see AttentiveStatsPoolingFn.forward
        try:
            return AttentiveStatsPoolingFn.apply(x, mask, self._flat, *self._flat.params)
        finally:
            self._flat._out_buffer = None                                    # never outlives the call it was meant for
