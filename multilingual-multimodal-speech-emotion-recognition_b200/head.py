"""FusionHead: the reference's hand-wired call sequence (src/train.py:54-69 construction, :145-168 step) as
one nn.Module over the drop-in modules.  This is the public entry point used by bench.py and the tests;
the individual modules remain usable on their own exactly like the reference's."""
from __future__ import annotations

import os

from typing import Dict, Optional

import torch
import torch.nn as nn

from ._params import FlatParams
from .functional import HeadLossFn
from .models import (AdvancedOpenMaxClassifier, AttentiveStatsPooling, BottleneckAdapter, CrossModalAttention,
                     FusionLayer, PrototypeMemory)


# dropout rates of the reference's training script: CrossModalAttention default (cross_attention.py:7), the
# hard-coded nn.Dropout(0.1) of FusionLayer (fusion.py:9,12), AdvancedOpenMaxClassifier(dropout=0.15) (train.py:67)
REFERENCE_DROPOUT = {"cross": 0.1, "fusion": 0.1, "classifier": 0.15}


def dropout_rates(dropout) -> Dict[str, float]:
    """float -> the same rate everywhere; 'reference' -> REFERENCE_DROPOUT; dict -> per-module rates."""
    if isinstance(dropout, str):
        if dropout != "reference":
            raise ValueError("dropout must be a float, a dict or 'reference'")
        return dict(REFERENCE_DROPOUT)
    if isinstance(dropout, dict):
        return {k: float(dropout.get(k, 0.0)) for k in REFERENCE_DROPOUT}
    return {k: float(dropout) for k in REFERENCE_DROPOUT}


class FusionHead(nn.Module):
    def __init__(self, num_labels: int = 4, hidden: int = 768, shared_dim: int = 256, num_heads: int = 8,
                 proj_dim: int = 512, num_layers: int = 35, dropout=0.0):
        super().__init__()
        self.num_labels = num_labels
        rates = dropout_rates(dropout)
        self.adapter_a = BottleneckAdapter(hidden, 256)
        self.adapter_t = BottleneckAdapter(hidden, 256)
        self.cross = CrossModalAttention(hidden, hidden, shared_dim=shared_dim, num_heads=num_heads,
                                         dropout=rates["cross"])
        self.pool_a = AttentiveStatsPooling(hidden)
        self.pool_t = AttentiveStatsPooling(hidden)
        self.fusion = FusionLayer(hidden * 2, hidden * 2, proj_dim)
        self.classifier = AdvancedOpenMaxClassifier(input_dim=proj_dim, num_labels=num_labels, num_layers=num_layers,
                                                    base_dim=proj_dim, dropout=rates["classifier"])
        self.set_dropout(rates)
        self.prototypes = PrototypeMemory(num_labels, proj_dim)
        self.loss_weights = dict(w_ce=1.0, w_focal=0.3, w_unc=0.05, w_proto=0.01)   # train.py:156-168
        self.persistent_grad_arena = False
        # step preparation that nothing before the fusion layer depends on -- the bf16 copy of the classifier's weights
        # (78 % of all parameters) and the zero fill of the persistent gradient arena, both pure HBM traffic -- runs on a
        # side stream beside the fusion MLP / gate chain (a dozen M = B GEMMs on a handful of SMs); SER_PREP_SIDE=0 or
        # this flag keep it at the start of the step on the caller's stream
        self.side_prep = os.environ.get("SER_PREP_SIDE", "1") != "0"

    def _side_stream(self, dev) -> "torch.cuda.Stream":
        s = self.__dict__.get("_side")
        if s is None or s.device != dev:
            s = torch.cuda.Stream(device=dev)
            self.__dict__["_side"] = s
        return s

    def set_dropout(self, dropout) -> Dict[str, float]:
        """Change the dropout rates of the built head (float, dict or 'reference'); returns the rates in effect."""
        rates = dropout_rates(dropout)
        self.cross.dropout.p = self.cross.attn_a.dropout = self.cross.attn_t.dropout = rates["cross"]
        self.cross.p_drop = rates["cross"]
        self.fusion.proj_a[2].p = self.fusion.proj_t[2].p = rates["fusion"]
        self.classifier.p_drop = rates["classifier"]
        for part in (self.classifier.deep_classifier, self.classifier.uncertainty_head):
            for m in part.modules():
                if isinstance(m, nn.Dropout):
                    m.p = rates["classifier"]
        self.dropout_rates = rates
        return rates

    GROUPS = ("adapter_a", "adapter_t", "cross", "pool_a", "pool_t", "fusion", "classifier", "prototypes")

    def load_group_state(self, weights: Dict[str, Dict[str, torch.Tensor]]) -> None:
        for k in self.GROUPS:
            getattr(self, k).load_state_dict(weights[k], strict=True)

    # ------------------------------------------------------------------------------------------------------------
    # checkpoint wire format of the reference (src/train.py:247-262 writes, src/eval.py:109-123 reads): one dict of
    # state_dicts keyed 'audio_encoder', 'text_encoder', 'cross', 'pool_a', 'pool_t', 'fusion', 'classifier',
    # 'prototypes' (+ 'optimizer', 'scheduler', 'epoch', 'f1').  The head owns everything but the encoders; of those it
    # owns the adapters, which live in the encoders' state_dicts under 'adapter.*' (audio_encoder.py:19, text_encoder.py:17).
    # ------------------------------------------------------------------------------------------------------------
    CKPT_GROUPS = ("cross", "pool_a", "pool_t", "fusion", "classifier", "prototypes")
    CKPT_ADAPTERS = {"audio_encoder": "adapter_a", "text_encoder": "adapter_t"}

    def checkpoint_state(self, encoders: Optional[Dict[str, Dict[str, torch.Tensor]]] = None, **extra) -> dict:
        """The reference's checkpoint dict for this head.  `encoders` = {'audio_encoder': state_dict, 'text_encoder':
        state_dict} of the (unchanged) encoder modules, whose 'adapter.*' entries are replaced by this head's adapters;
        without it the two encoder entries hold the adapter keys only.  `extra`: optimizer / scheduler state_dicts,
        epoch, f1 (train.py:258-261)."""
        ckpt = {}
        for enc, grp in self.CKPT_ADAPTERS.items():
            sd = dict(encoders[enc]) if encoders and enc in encoders else {}
            sd.update({f"adapter.{k}": v for k, v in getattr(self, grp).state_dict().items()})
            ckpt[enc] = sd
        for grp in self.CKPT_GROUPS:
            ckpt[grp] = getattr(self, grp).state_dict()
        ckpt.update(extra)
        return ckpt

    def load_checkpoint_state(self, ckpt: dict) -> None:
        """Load a checkpoint written by the reference's train.py (or by checkpoint_state): strict on every group the head
        owns; encoder keys other than 'adapter.*' are the encoders' business and are ignored here."""
        for enc, grp in self.CKPT_ADAPTERS.items():
            sd = {k[len("adapter."):]: v for k, v in ckpt[enc].items() if k.startswith("adapter.")}
            getattr(self, grp).load_state_dict(sd, strict=True)
        for grp in self.CKPT_GROUPS:
            getattr(self, grp).load_state_dict(ckpt[grp], strict=True)

    @torch.no_grad()
    def fit_weibull_on(self, batches) -> int:
        """The Weibull-fitting pass after the last epoch (src/train.py:204-245): run the validation batches
        (a_hid, t_hid, a_mask, t_mask, labels) through the head in eval mode, collect the classifier's 256-d penultimate
        features -- the fused stack hands them out (`classifier.last_features`), so the reference's hand-unrolled walk
        over the classifier's children is not needed -- and call classifier.fit_weibull.  Returns the sample count."""
        was_training = self.training
        self.eval()
        feats, labels = [], []
        try:
            for a_hid, t_hid, a_mask, t_mask, y in batches:
                out = self.features(a_hid, t_hid, a_mask, t_mask)
                self.classifier(out["fused"], use_openmax=False)
                feats.append(self.classifier.last_features.float())
                labels.append(y.to(feats[-1].device))
        finally:
            self.train(was_training)
        if not feats:
            return 0
        f, y = torch.cat(feats), torch.cat(labels)
        self.classifier.fit_weibull(f, y)
        return int(y.numel())

    def features(self, a_hid, t_hid, a_mask=None, t_mask=None):
        # bf16 tier: the operand copies of all modules' weights in one launch instead of one per module
        flats = [getattr(self, g)._flat for g in self.GROUPS if hasattr(getattr(self, g), "_flat")]
        side = self.side_prep and a_hid.is_cuda
        late = [self.classifier._flat] if side else []            # cast beside the fusion chain (below)
        for f in late:
            # everything that may ALLOCATE (the flat master buffer after a .to(), the bf16 copy's storage) happens here, on
            # the caller's stream: the caching allocator ties a block to the stream that was current when it was allocated
            f.ensure()
            if a_hid.dtype != torch.float32:
                f._lowp_buffer(a_hid.dtype)
        FlatParams.precast([f for f in flats if all(f is not l for l in late)], a_hid.dtype)
        arena = None
        if torch.is_grad_enabled():
            # one zero fill for all modules' gradient buffers of this step (a fresh allocation per step, or -- when the
            # caller guarantees zero_grad() before every step, as DataParallelHead.train_step does -- one persistent arena,
            # which is then filled on the side stream)
            defer = side and self.persistent_grad_arena
            arena = FlatParams.shared_grad_arena(flats, holder=self if self.persistent_grad_arena else None, zero=not defer)
            arena = arena if defer else None
        a_seq = self.adapter_a.residual_forward(a_hid)
        t_seq = self.adapter_t.residual_forward(t_hid)
        a_enh, t_enh = self.cross(a_seq, t_seq, a_mask, t_mask)
        # the pooled vectors land in the two halves of one tensor (see AttentiveStatsPoolingFn / FusionFn)
        pooled = torch.empty(2, a_enh.shape[0], 2 * a_enh.shape[2], device=a_enh.device, dtype=a_enh.dtype)
        self.pool_a._out_buffer, self.pool_t._out_buffer = pooled[0], pooled[1]
        a_vec = self.pool_a(a_enh, a_mask)
        t_vec = self.pool_t(t_enh, t_mask)
        if side:
            main, sd = torch.cuda.current_stream(a_hid.device), self._side_stream(a_hid.device)
            sd.wait_stream(main)                                  # fork (a parallel branch of a captured step graph)
            with torch.cuda.stream(sd):
                FlatParams.precast(late, a_hid.dtype)
                if arena is not None:
                    arena.zero_()
        fused = self.fusion(a_vec, t_vec)
        if side:
            main.wait_stream(sd)                                  # join: before the classifier and long before any backward kernel
        return dict(a_enh=a_enh, t_enh=t_enh, a_vec=a_vec, t_vec=t_vec, fused=fused)

    @torch.no_grad()
    def forward_views(self, audio_views, t_hid, a_mask=None, t_mask=None, use_openmax: bool = True) -> torch.Tensor:
        """Test-time augmentation (src/eval.py:174-190): V audio views of the same utterances against ONE text batch ->
        logits [V, B, C] (OpenMax-recalibrated in eval mode, as classifier.forward does).  Equal to calling the head once
        per view, but everything that depends on the text alone is computed once: the text adapter, the text-side
        q / k / v projections (and the folded projection weights) of cross attention -- SURVEY.md section 8(d), cfg5."""
        flats = [getattr(self, g)._flat for g in self.GROUPS if hasattr(getattr(self, g), "_flat")]
        t_seq = self.adapter_t.residual_forward(t_hid)
        cache: dict = {}
        logits = []
        for a_hid in audio_views:
            FlatParams.precast(flats, a_hid.dtype)
            a_seq = self.adapter_a.residual_forward(a_hid)
            a_enh, t_enh = self.cross(a_seq, t_seq, a_mask, t_mask, _text_cache=cache)
            pooled = torch.empty(2, a_enh.shape[0], 2 * a_enh.shape[2], device=a_enh.device, dtype=a_enh.dtype)
            self.pool_a._out_buffer, self.pool_t._out_buffer = pooled[0], pooled[1]
            fused = self.fusion(self.pool_a(a_enh, a_mask), self.pool_t(t_enh, t_mask))
            logits.append(self.classifier(fused, use_openmax=use_openmax))
        return torch.stack(logits)

    def forward(self, a_hid: torch.Tensor, t_hid: torch.Tensor, a_mask: Optional[torch.Tensor] = None,
                t_mask: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                loss_cfg: Optional[dict] = None) -> Dict[str, torch.Tensor]:
        out = self.features(a_hid, t_hid, a_mask, t_mask)
        if labels is None:
            out["logits"] = self.classifier(out["fused"])
            return out
        logits, unc, anchor_loss = self.classifier(out["fused"], use_openmax=False, return_uncertainty=True)
        cfg = dict(self.loss_weights)
        if loss_cfg:
            cfg.update(loss_cfg)
        terms = HeadLossFn.apply(logits, unc, out["fused"], self.prototypes.prototypes, labels, cfg)
        out.update(logits=logits, unc=unc, anchor=anchor_loss, ce=terms[0], focal=terms[1], unc_loss=terms[2],
                   proto=terms[3], loss=terms[4], accuracy=terms[5])      # (+ 0.1 * anchor_loss, train.py:158: identically 0, see classifier)
        return out
