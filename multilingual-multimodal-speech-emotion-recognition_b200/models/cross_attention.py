"""Drop-in CrossModalAttention (reference: src/models/cross_attention.py:6-53).

Bidirectional audio<->text multi-head cross attention with key-padding masks, residual and LayerNorm.
The parameter containers (nn.Linear / nn.MultiheadAttention / nn.LayerNorm) exist only to own the
reference-named parameters; forward() hands everything to one fused C-ABI call per direction pair.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .._params import FlatParams
from ..functional import CrossAttentionFn
from ._common import DropoutSeed, LayerNorm, Linear, active_dropout


class CrossModalAttention(nn.Module):
    def __init__(self, audio_dim: int, text_dim: int, shared_dim: int = 256, num_heads: int = 8, dropout: float = 0.1):
        super().__init__()
        if shared_dim % num_heads != 0:
            raise AssertionError(f"shared_dim {shared_dim} must be divisible by num_heads {num_heads}")
        # audio_dim != text_dim (e.g. wav2vec2-base with xlm-roberta-large) runs the per-layer GEMM path of the library;
        # equal widths with audio_dim == 3 * shared_dim (every reference script) take the folded, pair-batched one
        self.shared_dim, self.num_heads, self.p_drop = shared_dim, num_heads, float(dropout)
        for direction, (q_src, kv_src) in (("a", ("a", "t")), ("t", ("t", "a"))):
            dim_q = audio_dim if q_src == "a" else text_dim
            dim_kv = audio_dim if kv_src == "a" else text_dim
            setattr(self, f"q_{q_src}", Linear(dim_q, shared_dim))
            setattr(self, f"k_{kv_src}", Linear(dim_kv, shared_dim))
            setattr(self, f"v_{kv_src}", Linear(dim_kv, shared_dim))
            setattr(self, f"attn_{direction}", nn.MultiheadAttention(shared_dim, num_heads, dropout=dropout, batch_first=True))
            setattr(self, f"out_{direction}", Linear(shared_dim, dim_q))
        self.dropout = nn.Dropout(dropout)
        self.norm_a = LayerNorm(audio_dim)
        self.norm_t = LayerNorm(text_dim)
        named = dict(self.named_parameters())
        order = []
        for m in ("a", "t"):                       # q|k|v of one modality adjacent -> packed [3S, D] view
            order += [f"q_{m}.weight", f"k_{m}.weight", f"v_{m}.weight", f"q_{m}.bias", f"k_{m}.bias", f"v_{m}.bias"]
        order += [n for n in named if n not in order]
        self._flat = FlatParams([(n, named[n]) for n in order])
        self._drop_seed = DropoutSeed()

    def forward(self, audio_seq: torch.Tensor, text_seq: torch.Tensor, audio_mask: Optional[torch.Tensor] = None,
                text_mask: Optional[torch.Tensor] = None, _text_cache: Optional[dict] = None):
        """Reference signature (cross_attention.py:32).  `_text_cache` (inference only, optional): a dict shared by the
        calls that evaluate several audio views against ONE text batch -- the text-side projections are computed by the
        first call and reused by the others (FusionHead.forward_views)."""
        # attention-weight dropout of both nn.MultiheadAttention modules and self.dropout on the branch outputs
        # (cross_attention.py:18,25,43,51) run inside the kernels; `self.dropout.p` is the single rate, as in the reference
        p = active_dropout(self, self.dropout.p)
        seed = self._drop_seed.next(audio_seq.device) if p > 0.0 else None
        return CrossAttentionFn.apply(audio_seq, text_seq, audio_mask, text_mask, self._flat, self.num_heads, p, seed,
                                      _text_cache, *self._flat.params)
