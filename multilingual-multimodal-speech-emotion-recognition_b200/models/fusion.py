"""Drop-in FusionLayer (reference: src/models/fusion.py:5-25): per-modality MLP projection to proj_dim and a
normalised two-way sigmoid gate, executed by the fused fusion kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from .._params import FlatParams
from ..functional import FusionFn
from .. import _lib
from ._common import DropoutSeed, Linear, active_dropout


def _mlp(din: int, dout: int) -> nn.Sequential:
    # indices 0 and 3 hold the Linear layers, as in the reference state_dict (proj_*.0.*, proj_*.3.*)
    return nn.Sequential(Linear(din, dout), nn.ReLU(), nn.Dropout(0.1), Linear(dout, dout))


def _gate(d: int, hidden: int) -> nn.Sequential:
    return nn.Sequential(Linear(d, hidden), nn.ReLU(), Linear(hidden, 1))


class FusionLayer(nn.Module):
    def __init__(self, audio_dim: int, text_dim: int, proj_dim: int):
        super().__init__()
        self.proj_a = _mlp(audio_dim, proj_dim)
        self.proj_t = _mlp(text_dim, proj_dim)
        hidden = max(32, proj_dim // 2)
        self.gate_a = _gate(proj_dim, hidden)
        self.gate_t = _gate(proj_dim, hidden)
        self._flat = FlatParams([(n, p) for n, p in self.named_parameters()])
        self._drop_seed = DropoutSeed()

    def forward(self, audio_vec: torch.Tensor, text_vec: torch.Tensor) -> torch.Tensor:
        # the reference hard-codes Dropout(0.1) in both projections (fusion.py:9,12); the rate is read from the
        # nn.Dropout children so it can be changed the usual way (both branches share one rate in the fused kernel)
        if self.proj_a[2].p != self.proj_t[2].p:
            raise _lib.SerError("FusionLayer: proj_a[2].p and proj_t[2].p must be equal in the fused path")
        p = active_dropout(self, self.proj_a[2].p)
        seed = self._drop_seed.next(audio_vec.device) if p > 0.0 else None
        return FusionFn.apply(audio_vec, text_vec, self._flat, p, seed, *self._flat.params)
