"""Same-box PyTorch-eager baseline of the fusion head -- BASELINE ARM ONLY, not a product path and not the oracle.

SURVEY.md section 2a / 8(d) set the bar for the hot path as "the reference's arithmetic in PyTorch eager (cuBLAS +
unfused ATen kernels) on the same B200, fp32 and autocast(bf16)".  The reference itself cannot travel to the GPU box
(/root/reference does not exist there), so this file rebuilds the head out of STOCK torch modules -- nn.Linear,
nn.MultiheadAttention, nn.LayerNorm, nn.Dropout, F.cross_entropy -- with the reference's module tree and state_dict
key names, so that `load_state_dict(synth.head_weights(C)[group], strict=True)` both loads
the shared synthetic weights and proves the structure.  tests/test_eager_baseline.py holds it to the CPU oracle
(outputs, loss terms and gradients at 1e-5), which in turn is pinned to the reference's own modules.

Only bench.py (`--impl eager_gpu`, and the `eager_gpu` block of the main line) and that test import it.
Reference lines restated: src/models/audio_encoder.py:19-21,112; text_encoder.py:17-19,57; cross_attention.py:6-53;
pooling.py:6-28; fusion.py:5-25; classifier.py:8-238; prototypes.py:5-53; losses.py:7-64; src/train.py:145-168.

`graph_safe=True` swaps the three constructs of the reference's loss code that synchronise with the host
(torch.bincount, `if not torch.isfinite(loss)`) for sync-free equivalents with identical values, so the step can be
captured in a CUDA graph -- the strongest form of the eager baseline (no launch overhead left).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def _adapter(hidden: int = 768, bottleneck: int = 256) -> nn.Sequential:
    return nn.Sequential(nn.Linear(hidden, bottleneck), nn.ReLU(), nn.Linear(bottleneck, hidden))


class EagerCrossModalAttention(nn.Module):
    def __init__(self, dim: int = 768, shared: int = 256, heads: int = 8, dropout: float = 0.1):
        super().__init__()
        self.q_a, self.k_t, self.v_t = nn.Linear(dim, shared), nn.Linear(dim, shared), nn.Linear(dim, shared)
        self.attn_a = nn.MultiheadAttention(shared, heads, dropout=dropout, batch_first=True)
        self.out_a = nn.Linear(shared, dim)
        self.q_t, self.k_a, self.v_a = nn.Linear(dim, shared), nn.Linear(dim, shared), nn.Linear(dim, shared)
        self.attn_t = nn.MultiheadAttention(shared, heads, dropout=dropout, batch_first=True)
        self.out_t = nn.Linear(shared, dim)
        self.dropout = nn.Dropout(dropout)
        self.norm_a, self.norm_t = nn.LayerNorm(dim), nn.LayerNorm(dim)

    def forward(self, a, t, a_mask=None, t_mask=None):
        a_pad = None if a_mask is None else a_mask == 0
        t_pad = None if t_mask is None else t_mask == 0
        ctx_a, _ = self.attn_a(self.q_a(a), self.k_t(t), self.v_t(t), key_padding_mask=t_pad)
        a_enh = self.norm_a(a + self.dropout(self.out_a(ctx_a)))
        ctx_t, _ = self.attn_t(self.q_t(t), self.k_a(a), self.v_a(a), key_padding_mask=a_pad)
        t_enh = self.norm_t(t + self.dropout(self.out_t(ctx_t)))
        return a_enh, t_enh


class EagerAttentiveStatsPooling(nn.Module):
    def __init__(self, dim: int = 768, hidden: int = 128):
        super().__init__()
        self.attention = nn.Sequential(nn.Linear(dim, hidden), nn.Tanh(), nn.Linear(hidden, 1))

    def forward(self, x, mask=None):
        e = self.attention(x).squeeze(-1)
        if mask is not None:
            e = e.masked_fill(mask == 0, float("-inf"))
        w = torch.softmax(e, dim=1).unsqueeze(-1)
        mean = torch.sum(w * x, dim=1)
        var = torch.sum(w * (x - mean.unsqueeze(1)) ** 2, dim=1)
        return torch.cat([mean, torch.sqrt(var + 1e-6)], dim=-1)


class EagerFusionLayer(nn.Module):
    def __init__(self, din: int = 1536, proj: int = 512):
        super().__init__()
        mlp = lambda: nn.Sequential(nn.Linear(din, proj), nn.ReLU(), nn.Dropout(0.1), nn.Linear(proj, proj))   # noqa: E731
        gate = lambda: nn.Sequential(nn.Linear(proj, proj // 2), nn.ReLU(), nn.Linear(proj // 2, 1))           # noqa: E731
        self.proj_a, self.proj_t, self.gate_a, self.gate_t = mlp(), mlp(), gate(), gate()

    def forward(self, av, tv):
        a, t = self.proj_a(av), self.proj_t(tv)
        wa, wt = torch.sigmoid(self.gate_a(a)), torch.sigmoid(self.gate_t(t))
        total = wa + wt + 1e-8
        return (wa / total) * a + (wt / total) * t


class _ResidualBlock(nn.Module):
    def __init__(self, dim: int, p: float):
        super().__init__()
        self.block = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, dim), nn.ReLU(), nn.Dropout(p), nn.Linear(dim, dim),
                                   nn.Dropout(p))

    def forward(self, x):
        return x + self.block(x)


class _DeepClassifier(nn.Module):
    def __init__(self, din: int, classes: int, layers: int, dim: int, p: float):
        super().__init__()
        self.input_projection = nn.Sequential(nn.Linear(din, dim), nn.LayerNorm(dim), nn.ReLU(), nn.Dropout(p))
        self.residual_layers = nn.ModuleList([_ResidualBlock(dim, p) for _ in range(layers)])
        self.layer_norms = nn.ModuleList([nn.LayerNorm(dim) for _ in range(layers)])
        self.output_projection = nn.Sequential(nn.Linear(dim, dim // 2), nn.LayerNorm(dim // 2), nn.ReLU(), nn.Dropout(p),
                                               nn.Linear(dim // 2, classes))


class _AnchorClustering(nn.Module):
    def __init__(self, feat: int, classes: int, anchor_dim: int = 128):
        super().__init__()
        self.class_anchors = nn.Parameter(torch.randn(classes, anchor_dim))
        self.anchor_projection = nn.Sequential(nn.Linear(feat, anchor_dim), nn.LayerNorm(anchor_dim), nn.ReLU(), nn.Dropout(0.1))
        self.temperature = nn.Parameter(torch.tensor(1.0))

    def forward(self, f):
        z = F.normalize(self.anchor_projection(f), p=2, dim=1)
        anchors = F.normalize(self.class_anchors, p=2, dim=1)
        sims = z @ anchors.t()
        # the reference's clustering term: identically zero (similarity minus its own row maximum, clamped at 0)
        loss = torch.clamp(sims - sims.max(dim=1, keepdim=True)[0], min=0.0).mean()
        return sims / self.temperature, loss


class EagerClassifier(nn.Module):
    def __init__(self, classes: int, layers: int = 35, dim: int = 512, dropout: float = 0.15):
        super().__init__()
        self.deep_classifier = _DeepClassifier(dim, classes, layers, dim, dropout)
        self.anchor_clustering = _AnchorClustering(dim // 2, classes)
        self.register_buffer("weibull_alpha", torch.ones(classes))
        self.register_buffer("weibull_beta", torch.ones(classes))
        self.register_buffer("weibull_tau", torch.zeros(classes))
        self.register_buffer("activation_vectors", torch.zeros(classes, dim // 2))
        self.uncertainty_head = nn.Sequential(nn.Linear(dim // 2, 64), nn.ReLU(), nn.Dropout(dropout), nn.Linear(64, 1),
                                              nn.Sigmoid())

    def forward(self, x):
        dc = self.deep_classifier
        h = dc.input_projection(x)
        for norm, block in zip(dc.layer_norms, dc.residual_layers):      # the 35-iteration Python loop of the reference
            h = block(norm(h))
        for i in range(4):
            h = dc.output_projection[i](h)
        _, anchor_loss = self.anchor_clustering(h)
        return dc.output_projection[4](h), self.uncertainty_head(h), anchor_loss


class EagerHead(nn.Module):
    GROUPS = ("adapter_a", "adapter_t", "cross", "pool_a", "pool_t", "fusion", "classifier", "prototypes")

    def __init__(self, classes: int = 4, layers: int = 35, dropout: Optional[Dict[str, float]] = None, graph_safe: bool = False):
        super().__init__()
        r = dropout or {"cross": 0.0, "fusion": 0.0, "classifier": 0.0}
        self.classes, self.graph_safe = classes, graph_safe
        self.adapter_a, self.adapter_t = _adapter(), _adapter()
        self.cross = EagerCrossModalAttention(dropout=r["cross"])
        self.pool_a, self.pool_t = EagerAttentiveStatsPooling(), EagerAttentiveStatsPooling()
        self.fusion = EagerFusionLayer()
        for m in (self.fusion.proj_a[2], self.fusion.proj_t[2]):
            m.p = r["fusion"]
        self.classifier = EagerClassifier(classes, layers, dropout=r["classifier"])
        self.prototypes = nn.ParameterDict({"prototypes": nn.Parameter(torch.randn(classes, 512) * 0.02)})

    def load_group_state(self, weights) -> None:
        for g in self.GROUPS:
            getattr(self, g).load_state_dict(weights[g], strict=True)

    # ---- the three loss modules (losses.py:12-30, 41-64; prototypes.py:13-53) ----
    def _finite_or_zero(self, loss):
        if self.graph_safe:
            return torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss))
        if not torch.isfinite(loss):                      # host sync, as in the reference
            return torch.zeros((), device=loss.device, requires_grad=True)
        return loss

    def _ce_smooth(self, logits, y, eps: float = 0.1):
        C = logits.shape[-1]
        logp = F.log_softmax(logits.clamp(-10.0, 10.0), dim=-1)
        q = torch.full_like(logp, eps / (C - 1)).scatter_(1, y.clamp(0, C - 1).unsqueeze(1), 1.0 - eps)
        return self._finite_or_zero((-q * logp).sum(dim=-1).mean())

    def _cb_focal(self, logits, y, beta: float = 0.9999, gamma: float = 2.0):
        C = self.classes
        with torch.no_grad():
            if self.graph_safe:
                counts = torch.zeros(C, device=y.device).scatter_add_(0, y, torch.ones_like(y, dtype=torch.float32))
            else:
                counts = torch.bincount(y, minlength=C).float()
            base = torch.full((), beta, device=logits.device) if self.graph_safe else torch.tensor(beta, device=logits.device)
            eff = (1.0 - torch.pow(base, counts.clamp(min=1.0))).clamp(min=1e-6)
            w = (1.0 - beta) / eff
            w = w / (w.sum() + 1e-8) * C
        z = logits.clamp(-10.0, 10.0)
        pt = F.softmax(z, dim=-1).gather(1, y.unsqueeze(1)).squeeze(1).clamp(1e-6, 1.0)
        ce = F.cross_entropy(z, y, reduction="none", weight=w.to(z.dtype))
        return self._finite_or_zero((torch.pow(1.0 - pt, gamma) * ce).mean())

    def _proto(self, emb, y, margin: float = 0.5):
        protos = self.prototypes["prototypes"]
        e = emb.clamp(-10.0, 10.0)
        pos = (e - protos[y]).norm(dim=1).mean()
        d = torch.sqrt(((e.unsqueeze(1) - protos.unsqueeze(0)) ** 2).sum(dim=2) + 1e-6)
        own = torch.zeros_like(d, dtype=torch.bool).scatter_(1, y.unsqueeze(1), True)
        d = d.masked_fill(own, float("inf")).clamp(max=10.0)      # the own-class entry stays in the soft-min as 10.0
        neg = (-torch.logsumexp(-d, dim=1)).mean()
        return self._finite_or_zero(pos + margin - neg)

    def forward(self, a, t, a_mask, t_mask, labels):
        a_seq = a + self.adapter_a(a)
        t_seq = t + self.adapter_t(t)
        a_enh, t_enh = self.cross(a_seq, t_seq, a_mask, t_mask)
        a_vec, t_vec = self.pool_a(a_enh, a_mask), self.pool_t(t_enh, t_mask)
        fused = self.fusion(a_vec, t_vec)
        logits, unc, anchor = self.classifier(fused)
        ce, focal = self._ce_smooth(logits, labels), self._cb_focal(logits, labels)
        # [B,1] * [B] broadcasts to [B,B] in the reference (train.py:162): value = mean(unc) * mean(correct)
        unc_loss = torch.mean(unc * (labels == logits.argmax(dim=1)).float())
        proto = self._proto(fused.float(), labels)
        loss = ce + 0.3 * focal + 0.1 * anchor + 0.05 * unc_loss + 0.01 * proto
        return dict(loss=loss, logits=logits, unc=unc, fused=fused, a_enh=a_enh, t_enh=t_enh, a_vec=a_vec, t_vec=t_vec,
                    ce=ce, focal=focal, unc_loss=unc_loss, proto=proto)
