"""Per-utterance feature fusion of the two encoders, the one trainable op between the
adapter and cross attention in the reference's default configuration:

    AudioEncoder.quality_fusion / conditioning_fusion / combined_fusion = Sequential(Linear(hid + {8, 12, 20}, hid),
    ReLU, Dropout(0.1))  (src/models/audio_encoder.py:29-52), applied to cat([seq, features.expand(frames, -1)])
    per utterance (:114-138);  TextEncoder.asr_fusion = the same with 8 ASR features (text_encoder.py:26-30,60-73).

`UtteranceFeatureFusion` is an nn.Sequential with the reference's child layout (0 = Linear, 1 = ReLU, 2 = Dropout), so
it can be assigned to `encoder.combined_fusion` etc. and keeps the checkpoint keys `0.weight [hid, hid+F]`, `0.bias`.
The utterance features are constant over the frames, so the kernels never build the concatenation: the feature part of
the weight becomes a per-utterance bias (csrc/featfuse.cu).

Call forms:
    fuse(seq [B,T,hid], features [B,F])    batched (the B200-native call: all utterances of a batch at once)
    fuse(seq [T,hid],   features [F])      one utterance
    fuse(fused_input [T, hid+F])           the reference's own call (audio_encoder.py:131): the feature columns are read
                                           from the first frame -- they are an expand() of one vector by construction
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import _lib
from .._params import FlatParams
from ..functional import FeatureFusionFn
from ._common import DropoutSeed, Linear, active_dropout


class UtteranceFeatureFusion(nn.Sequential):
    def __init__(self, hidden: int = 768, num_features: int = 20, dropout: float = 0.1):
        super().__init__(Linear(hidden + num_features, hidden), nn.ReLU(), nn.Dropout(dropout))
        self.hidden, self.num_features = hidden, num_features
        self._flat = FlatParams([(n, p) for n, p in self.named_parameters()])
        self._drop_seed = DropoutSeed()

    def forward(self, seq: torch.Tensor, features: Optional[torch.Tensor] = None) -> torch.Tensor:
        D, F = self.hidden, self.num_features
        if features is None:
            if seq.shape[-1] != D + F:
                raise _lib.SerError(f"feature fusion: expected [..., {D + F}] pre-concatenated input, got {tuple(seq.shape)}")
            features = seq[..., 0, D:]
            seq = seq[..., :D]
        single = seq.dim() == 2
        if single:
            seq, features = seq.unsqueeze(0), features.reshape(1, F)
        if seq.dim() != 3 or seq.shape[-1] != D or tuple(features.shape) != (seq.shape[0], F):
            raise _lib.SerError(f"feature fusion: seq {tuple(seq.shape)} / features {tuple(features.shape)} do not match "
                                f"hidden {D}, num_features {F}")
        p = active_dropout(self, self[2].p)
        seed = self._drop_seed.next(seq.device) if p > 0.0 else None
        y = FeatureFusionFn.apply(seq, features, self._flat, p, seed, *self._flat.params)
        return y.squeeze(0) if single else y
