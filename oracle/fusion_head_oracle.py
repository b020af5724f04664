"""CPU ORACLE for the fusion-head hot path -- TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

A plain-PyTorch (CPU, fp32 or fp64) functional restatement of the reference's algorithm, written from
the reference's semantics with every function citing the file:line it follows (paths relative to the
kananmittal/Multilingual-Multimodal-Speech-Emotion-Recognition checkout).  It deliberately avoids
nn.MultiheadAttention / nn.LayerNorm / F.cross_entropy and spells the arithmetic out, so that it is an
independent statement of WHAT the CUDA kernels must compute.  Gradients come from autograd on this
restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker / reported baseline -- never on the path being measured or shipped.

Parity pin: the reference ships no tests or golden vectors.  The oracle is
pinned against outputs of the reference's own modules imported from /root/reference in the build
container: oracle/make_golden.py generates tests/golden/*.pt and tests/test_oracle_golden.py checks this
file against them (<= 2e-6 relative in fp32).

Weights are passed as flat dicts keyed exactly like the reference modules' state_dict()s.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor
W = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------------
# Optional emulation of a reduced-precision-operand implementation (what torch.autocast(bfloat16) does to the
# reference): every Linear rounds its input and weight to `_OPERAND_DTYPE` (straight-through gradient) and
# accumulates in the working precision.  Used by the tests to measure the NOISE FLOOR any bf16-operand
# implementation has against the exact answer (ReLU-mask flips make that floor batch-size dependent).
_OPERAND_DTYPE = None


class operand_rounding:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global _OPERAND_DTYPE
        self.prev, _OPERAND_DTYPE = _OPERAND_DTYPE, self.dtype

    def __exit__(self, *exc):
        global _OPERAND_DTYPE
        _OPERAND_DTYPE = self.prev


# Optional dropout: the reference's nn.Dropout / MultiheadAttention(dropout=p) layers draw from torch's Philox stream,
# which no other implementation can reproduce; what CAN be pinned is the arithmetic given the masks.  Inside
# `with dropout_masks({site: multiplier tensor})` every dropout site multiplies its input by the given tensor
# (0 or 1/(1-p) per element -- torch.nn.functional.dropout's output is exactly x * such a tensor); sites without an
# entry, and all sites outside the context, are the identity (eval mode).  Site names:
#   cross.prob_a / cross.prob_t [B,H,Tq,Tk]   attention weights of attn_a / attn_t (functional.py multi_head_attention_forward)
#   cross.res_a / cross.res_t   [B,T,D]       self.dropout(a_out) / self.dropout(t_out)   (cross_attention.py:43,51)
#   fusion.a / fusion.t         [B,P]         proj_a[2] / proj_t[2]                        (fusion.py:9,12)
#   feat.out                    [B,T,D]       combined_fusion[2] etc.                      (audio_encoder.py:33,42,51)
#   clf.in [B,P]; clf.block{i}.hidden, clf.block{i}.out [B,P]; clf.out [B,F]; clf.unc [B,64]   (classifier.py:83,85,109,127,195)
_DROPOUT_MASKS = None


class dropout_masks:
    def __init__(self, masks):
        self.masks = masks

    def __enter__(self):
        global _DROPOUT_MASKS
        self.prev, _DROPOUT_MASKS = _DROPOUT_MASKS, self.masks

    def __exit__(self, *exc):
        global _DROPOUT_MASKS
        _DROPOUT_MASKS = self.prev


def _drop(x: Tensor, site: str) -> Tensor:
    if _DROPOUT_MASKS is None:
        return x
    if callable(_DROPOUT_MASKS):             # mask source: f(site, x) -> multiplier tensor or None
        m = _DROPOUT_MASKS(site, x)
        return x if m is None else x * m.to(x.dtype).reshape(x.shape)
    if site not in _DROPOUT_MASKS:
        return x
    return x * _DROPOUT_MASKS[site].to(x.dtype).reshape(x.shape)


def random_dropout(rates: Dict[str, float], generator=None):
    """Mask source for `dropout_masks`: fresh Bernoulli keep masks per call, like nn.Dropout in train() mode.
    `rates` maps the site prefix ('cross', 'fusion', 'clf') to p."""
    def f(site: str, x: Tensor):
        p = float(rates.get(site.split(".")[0], 0.0))
        if p <= 0.0:
            return None
        return (torch.rand(x.shape, generator=generator) >= p).to(x.dtype) / (1.0 - p)
    return f


def _ste_round(x: Tensor) -> Tensor:
    return x + (x.to(_OPERAND_DTYPE).to(x.dtype) - x).detach()


def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    if _OPERAND_DTYPE is not None:
        x, w = _ste_round(x), _ste_round(w)
    y = x @ w.t()
    return y if b is None else y + b


def layer_norm(x: Tensor, g: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm over the last dim: biased variance, eps inside the sqrt."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


# --------------------------------------------------------------------------------------------------
# a1  bottleneck adapters  (src/models/audio_encoder.py:19-21,112 ; src/models/text_encoder.py:17-19,57)
# --------------------------------------------------------------------------------------------------
def adapter(x: Tensor, w: W) -> Tensor:
    """seq = seq + Linear(768->256) . ReLU . Linear(256->768)(seq); keys '0.weight','0.bias','2.weight','2.bias'."""
    h = torch.relu(linear(x, w["0.weight"], w["0.bias"]))
    return x + linear(h, w["2.weight"], w["2.bias"])


# --------------------------------------------------------------------------------------------------
# f1  per-utterance feature fusion  (src/models/audio_encoder.py:29-52 applied :114-138,
#     src/models/text_encoder.py:26-30 applied :60-73)
# --------------------------------------------------------------------------------------------------
def utterance_feature_fusion(seq: Tensor, feats: Tensor, w: W) -> Tensor:
    """seq [B,T,hid], feats [B,F]: per utterance the reference expands the feature vector over the frames, concatenates
    it to the hidden states and applies Sequential(Linear(hid+F, hid), ReLU, Dropout): keys '0.weight', '0.bias';
    dropout site 'feat.out' [B,T,hid]."""
    B, T, _ = seq.shape
    fused_input = torch.cat([seq, feats.unsqueeze(1).expand(B, T, feats.shape[-1])], dim=-1)
    return _drop(torch.relu(linear(fused_input, w["0.weight"], w["0.bias"])), "feat.out")


# --------------------------------------------------------------------------------------------------
# a2  CrossModalAttention  (src/models/cross_attention.py:32-53; MHA maths torch/nn/functional.py:6609-6645)
# --------------------------------------------------------------------------------------------------
def _mha(q_in: Tensor, k_in: Tensor, v_in: Tensor, kpm: Optional[Tensor], in_w: Tensor, in_b: Tensor,
         out_w: Tensor, out_b: Tensor, num_heads: int, drop_site: str = "") -> Tensor:
    """nn.MultiheadAttention(batch_first=True) forward, explicit math path; dropout acts on the softmax output.

    in_w [3E,E] is chunked q/k/v; q is scaled by 1/sqrt(dh) *after* its bias; key padding becomes an
    additive -inf on the scores; softmax over keys; out_proj.  A sample whose keys are all padded yields
    NaN for all its queries (softmax over all -inf) -- kept on purpose.
    """
    B, Tq, E = q_in.shape
    Tk = k_in.shape[1]
    dh = E // num_heads
    wq, wk, wv = in_w[:E], in_w[E:2 * E], in_w[2 * E:]
    bq, bk, bv = in_b[:E], in_b[E:2 * E], in_b[2 * E:]
    q = linear(q_in, wq, bq).view(B, Tq, num_heads, dh).transpose(1, 2)  # [B,H,Tq,dh]
    k = linear(k_in, wk, bk).view(B, Tk, num_heads, dh).transpose(1, 2)
    v = linear(v_in, wv, bv).view(B, Tk, num_heads, dh).transpose(1, 2)
    s = (q * (1.0 / math.sqrt(dh))) @ k.transpose(-1, -2)                # [B,H,Tq,Tk]
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    p = _drop(torch.softmax(s, dim=-1), drop_site)
    ctx = (p @ v).transpose(1, 2).reshape(B, Tq, E)
    return linear(ctx, out_w, out_b)


def cross_attention(a: Tensor, t: Tensor, a_mask: Optional[Tensor], t_mask: Optional[Tensor], w: W,
                    num_heads: int = 8) -> Tuple[Tensor, Tensor]:
    """(audio_enh, text_enh); masks are float 1=valid / 0=pad or None (cross_attention.py:34-35)."""
    t_kpm = (t_mask == 0) if t_mask is not None else None
    a_kpm = (a_mask == 0) if a_mask is not None else None
    # A <- T  (cross_attention.py:38-43)
    qa = linear(a, w["q_a.weight"], w["q_a.bias"])
    kt = linear(t, w["k_t.weight"], w["k_t.bias"])
    vt = linear(t, w["v_t.weight"], w["v_t.bias"])
    a_ctx = _mha(qa, kt, vt, t_kpm, w["attn_a.in_proj_weight"], w["attn_a.in_proj_bias"],
                 w["attn_a.out_proj.weight"], w["attn_a.out_proj.bias"], num_heads, "cross.prob_a")
    a_out = linear(a_ctx, w["out_a.weight"], w["out_a.bias"])
    audio_enh = layer_norm(a + _drop(a_out, "cross.res_a"), w["norm_a.weight"], w["norm_a.bias"])
    # T <- A  (cross_attention.py:46-51)
    qt = linear(t, w["q_t.weight"], w["q_t.bias"])
    ka = linear(a, w["k_a.weight"], w["k_a.bias"])
    va = linear(a, w["v_a.weight"], w["v_a.bias"])
    t_ctx = _mha(qt, ka, va, a_kpm, w["attn_t.in_proj_weight"], w["attn_t.in_proj_bias"],
                 w["attn_t.out_proj.weight"], w["attn_t.out_proj.bias"], num_heads, "cross.prob_t")
    t_out = linear(t_ctx, w["out_t.weight"], w["out_t.bias"])
    text_enh = layer_norm(t + _drop(t_out, "cross.res_t"), w["norm_t.weight"], w["norm_t.bias"])
    return audio_enh, text_enh


# --------------------------------------------------------------------------------------------------
# a3  AttentiveStatsPooling  (src/models/pooling.py:15-28)
# --------------------------------------------------------------------------------------------------
def attentive_stats_pooling(x: Tensor, mask: Optional[Tensor], w: W) -> Tensor:
    e = linear(torch.tanh(linear(x, w["attention.0.weight"], w["attention.0.bias"])),
               w["attention.2.weight"], w["attention.2.bias"]).squeeze(-1)          # [B,T]
    if mask is not None:
        e = e.masked_fill(mask == 0, float("-inf"))
    alpha = torch.softmax(e, dim=-1).unsqueeze(-1)                                   # [B,T,1]
    mean = (alpha * x).sum(dim=1)
    var = (alpha * (x - mean.unsqueeze(1)) ** 2).sum(dim=1)                          # two-pass (pooling.py:26)
    std = torch.sqrt(var + 1e-6)
    return torch.cat([mean, std], dim=-1)


# --------------------------------------------------------------------------------------------------
# a4  FusionLayer  (src/models/fusion.py:18-25)
# --------------------------------------------------------------------------------------------------
def fusion(av: Tensor, tv: Tensor, w: W) -> Tensor:
    pa = linear(_drop(torch.relu(linear(av, w["proj_a.0.weight"], w["proj_a.0.bias"])), "fusion.a"),
                w["proj_a.3.weight"], w["proj_a.3.bias"])
    pt = linear(_drop(torch.relu(linear(tv, w["proj_t.0.weight"], w["proj_t.0.bias"])), "fusion.t"),
                w["proj_t.3.weight"], w["proj_t.3.bias"])
    wa = torch.sigmoid(linear(torch.relu(linear(pa, w["gate_a.0.weight"], w["gate_a.0.bias"])),
                              w["gate_a.2.weight"], w["gate_a.2.bias"]))             # [B,1]
    wt = torch.sigmoid(linear(torch.relu(linear(pt, w["gate_t.0.weight"], w["gate_t.0.bias"])),
                              w["gate_t.2.weight"], w["gate_t.2.bias"]))
    wsum = wa + wt + 1e-8
    return (wa / wsum) * pa + (wt / wsum) * pt


# --------------------------------------------------------------------------------------------------
# a5 / a6 / a7  AdvancedOpenMaxClassifier  (src/models/classifier.py:200-305)
# --------------------------------------------------------------------------------------------------
def classifier_features(x: Tensor, w: W, num_layers: int = 35) -> Tensor:
    """[B,512] -> penultimate 256-d features (classifier.py:203-218)."""
    p = "deep_classifier."
    h = torch.relu(layer_norm(linear(x, w[p + "input_projection.0.weight"], w[p + "input_projection.0.bias"]),
                              w[p + "input_projection.1.weight"], w[p + "input_projection.1.bias"]))
    h = _drop(h, "clf.in")
    for i in range(num_layers):
        # outer LayerNorm first; the residual branch starts from the OUTER-LN output (classifier.py:207-212)
        y = layer_norm(h, w[f"{p}layer_norms.{i}.weight"], w[f"{p}layer_norms.{i}.bias"])
        b = f"{p}residual_layers.{i}.block."
        n = layer_norm(y, w[b + "0.weight"], w[b + "0.bias"])
        r = _drop(torch.relu(linear(n, w[b + "1.weight"], w[b + "1.bias"])), f"clf.block{i}.hidden")
        h = y + _drop(linear(r, w[b + "4.weight"], w[b + "4.bias"]), f"clf.block{i}.out")
    f = linear(h, w[p + "output_projection.0.weight"], w[p + "output_projection.0.bias"])
    f = torch.relu(layer_norm(f, w[p + "output_projection.1.weight"], w[p + "output_projection.1.bias"]))
    return _drop(f, "clf.out")


def anchor_clustering(f: Tensor, w: W) -> Tuple[Tensor, Tensor]:
    """ClassAnchorClustering.forward (classifier.py:32-70).  The returned loss is identically 0:
    mean(clamp(sim - max_c sim, min=0))."""
    p = "anchor_clustering."
    z = torch.relu(layer_norm(linear(f, w[p + "anchor_projection.0.weight"], w[p + "anchor_projection.0.bias"]),
                              w[p + "anchor_projection.1.weight"], w[p + "anchor_projection.1.bias"]))
    z = z / z.norm(dim=1, keepdim=True).clamp_min(1e-12)
    an = w[p + "class_anchors"] / w[p + "class_anchors"].norm(dim=1, keepdim=True).clamp_min(1e-12)
    sim_t = (z @ an.t()) / w[p + "temperature"]
    sim = z @ an.t()
    loss = (sim - sim.max(dim=1, keepdim=True)[0]).clamp(min=0).mean()
    return sim_t, loss


def openmax(features: Tensor, logits: Tensor, w: W) -> Tensor:
    """openmax_forward (classifier.py:240-275): Weibull-CDF of distance to class activation vectors."""
    dist = (features.unsqueeze(1) - w["activation_vectors"].unsqueeze(0)).norm(dim=2)     # [B,C]
    beta = w["weibull_beta"].clamp(min=1e-6)
    sx = (dist - w["weibull_tau"]).clamp(min=0)
    cdf = 1 - torch.exp(-torch.pow(sx / beta, w["weibull_alpha"]))
    unknown = torch.maximum(torch.zeros_like(cdf[:, 0]), cdf.max(dim=1)[0])
    scale = torch.where(unknown > 0.3, 1 - unknown * 0.8, torch.ones_like(unknown))
    return logits * scale.unsqueeze(1)


def fit_weibull(features: Tensor, labels: Tensor, num_labels: int, w: W) -> W:
    """fit_weibull (classifier.py:277-305): mean activation, alpha 2.5, beta 1.5*pop-std, tau 0.8*min."""
    out = {k: w[k].clone() for k in ("weibull_alpha", "weibull_beta", "weibull_tau", "activation_vectors")}
    for c in range(num_labels):
        m = labels == c
        if m.sum() == 0:
            continue
        cf = features[m]
        mean = cf.mean(dim=0)
        out["activation_vectors"][c] = mean
        d = (cf - mean).norm(dim=1)
        out["weibull_alpha"][c] = 2.5
        out["weibull_beta"][c] = d.std(unbiased=False) * 1.5     # numpy .std() is the population std
        out["weibull_tau"][c] = d.min() * 0.8
    return out


def classifier(x: Tensor, w: W, num_layers: int = 35, use_openmax: bool = True, training: bool = False,
               return_uncertainty: bool = False):
    """AdvancedOpenMaxClassifier.forward (classifier.py:200-238)."""
    f = classifier_features(x, w, num_layers)
    _, anchor_loss = anchor_clustering(f, w)
    p = "deep_classifier."
    logits = linear(f, w[p + "output_projection.4.weight"], w[p + "output_projection.4.bias"])
    unc = None
    if return_uncertainty:
        u = _drop(torch.relu(linear(f, w["uncertainty_head.0.weight"], w["uncertainty_head.0.bias"])), "clf.unc")
        unc = torch.sigmoid(linear(u, w["uncertainty_head.3.weight"], w["uncertainty_head.3.bias"]))
    if use_openmax and not training:
        logits = openmax(f, logits, w)
    if return_uncertainty:
        return logits, unc, anchor_loss
    return logits


# --------------------------------------------------------------------------------------------------
# a8  PrototypeMemory.prototype_loss  (src/models/prototypes.py:13-53)
# --------------------------------------------------------------------------------------------------
def prototype_loss(emb: Tensor, labels: Tensor, protos: Tensor, margin: float = 0.5) -> Tensor:
    e = emb.clamp(min=-10.0, max=10.0)
    pos = (e - protos[labels]).norm(dim=1).mean()
    d = torch.sqrt(((e.unsqueeze(1) - protos.unsqueeze(0)) ** 2).sum(dim=2) + 1e-6)     # [B,C]
    own = torch.zeros_like(d, dtype=torch.bool)
    own[torch.arange(e.shape[0]), labels] = True
    # own class -> +inf -> clamp(max=10) turns it into the constant 10.0, which STAYS in the soft-min
    nd = d.masked_fill(own, float("inf")).clamp(max=10.0)
    neg = (-torch.logsumexp(-nd, dim=1)).mean()
    loss = pos + margin - neg
    if not torch.isfinite(loss):
        return torch.zeros((), dtype=emb.dtype, requires_grad=True)
    return loss


# --------------------------------------------------------------------------------------------------
# a9 / a10  losses  (src/models/losses.py:12-30, 41-64)
# --------------------------------------------------------------------------------------------------
def label_smoothing_ce(logits: Tensor, target: Tensor, smoothing: float = 0.1) -> Tensor:
    C = logits.shape[-1]
    target = target.long().clamp(min=0, max=max(0, C - 1))
    z = logits.clamp(min=-10.0, max=10.0)
    logp = z - torch.logsumexp(z, dim=-1, keepdim=True)
    logp = torch.nan_to_num(logp, neginf=-1e9)
    q = torch.full_like(logp, smoothing / (C - 1))
    q.scatter_(1, target.unsqueeze(1), 1.0 - smoothing)
    loss = (-q * logp).sum(dim=-1)
    loss = torch.nan_to_num(loss, nan=0.0, posinf=1e6, neginf=1e6).mean()
    if not torch.isfinite(loss):
        return torch.zeros((), dtype=logits.dtype, requires_grad=True)
    return loss


def class_balanced_focal(logits: Tensor, targets: Tensor, beta: float = 0.9999, gamma: float = 2.0,
                         num_classes: Optional[int] = None, counts: Optional[Tensor] = None) -> Tensor:
    """`counts` lets a data-parallel caller pass the GLOBAL per-class counts."""
    weights = None
    if num_classes is not None:
        with torch.no_grad():
            if counts is None:
                counts = torch.bincount(targets, minlength=num_classes)
            cnt = counts.to(torch.float32).clamp(min=1.0)
            eff = (1.0 - torch.pow(torch.tensor(beta, dtype=torch.float32), cnt)).clamp(min=1e-6)
            weights = (1.0 - beta) / eff
            weights = (weights / (weights.sum() + 1e-8) * num_classes).to(logits.dtype)
    z = logits.clamp(min=-10.0, max=10.0)
    logp = z - torch.logsumexp(z, dim=-1, keepdim=True)
    probs = torch.exp(logp)
    pt = probs.gather(1, targets.unsqueeze(1)).squeeze(1).clamp(min=1e-6, max=1.0)
    focal = torch.pow(1.0 - pt, gamma)                       # NOT detached (losses.py:57)
    nll = -logp.gather(1, targets.unsqueeze(1)).squeeze(1)
    ce = nll if weights is None else nll * weights[targets]
    loss = (focal * ce).mean()
    if not torch.isfinite(loss):
        return torch.zeros((), dtype=logits.dtype, requires_grad=True)
    return loss


def supcon_loss(features: Tensor, labels: Tensor, temperature: float = 0.07) -> Tensor:
    """SupConLoss.forward (src/models/losses.py:75-88): the row max INCLUDES the diagonal; positives and the
    denominator exclude it; rows without positives contribute 0 / 1e-12 = 0."""
    f = features / features.norm(dim=-1, keepdim=True).clamp_min(1e-12)            # F.normalize
    logits = f @ f.t() / temperature
    logits = logits - logits.max(dim=1, keepdim=True)[0]
    B = f.shape[0]
    off = 1.0 - torch.eye(B, dtype=f.dtype)
    pos = (labels.unsqueeze(1) == labels.unsqueeze(0)).to(f.dtype) * off
    log_prob = logits - torch.log((torch.exp(logits) * off).sum(dim=1, keepdim=True) + 1e-12)
    return -((pos * log_prob).sum(dim=1) / (pos.sum(dim=1) + 1e-12)).mean()


# --------------------------------------------------------------------------------------------------
# a11  loss composition  (src/train.py:151-168)  and  a12 eval post-processing (src/eval.py, src/utils.py)
# --------------------------------------------------------------------------------------------------
def train_loss(logits: Tensor, unc: Tensor, anchor_loss: Tensor, fused: Tensor, labels: Tensor, protos: Tensor,
               num_classes: int, proto_weight_on: bool = True) -> Dict[str, Tensor]:
    ce = label_smoothing_ce(logits, labels, 0.1)
    focal = class_balanced_focal(logits, labels, 0.9999, 2.0, num_classes)
    correct = (labels == logits.argmax(dim=1)).to(logits.dtype)
    # unc is [B,1], correct is [B]: the product broadcasts to [B,B] (train.py:162)
    unc_loss = (unc * correct).mean()
    out = {"ce": ce, "focal": focal, "anchor": anchor_loss, "unc_loss": unc_loss}
    loss = ce + 0.3 * focal + 0.1 * anchor_loss + 0.05 * unc_loss
    if proto_weight_on:
        proto = prototype_loss(fused, labels, protos)
        out["proto"] = proto
        loss = loss + 0.01 * proto
    out["loss"] = loss
    return out


def energy_score(logits: Tensor) -> Tensor:
    """src/utils.py:12-14."""
    return -torch.logsumexp(logits, dim=-1)


def late_ood_scores(logits: Tensor, features: Tensor, w: W) -> Dict[str, Tensor]:
    """Late-stage OOD scoring: src/models/dual_gate_ood.py:203-220 (energy of logits / T),
    :280-312 (diagonal Mahalanobis distance to each class prototype, eps 1e-8 on the variances), :374-383 (sigmoid(-E),
    exp(-min distance), softmax-weighted mix).  Keys as in LateStageOODDetector.state_dict()."""
    energy = -torch.logsumexp(logits / w["energy_detector.temperature"], dim=-1)
    P, cov = w["prototype_detector.prototypes"], w["prototype_detector.covariances"]
    diff = features.unsqueeze(1) - P.unsqueeze(0)                              # [B,C,D]
    dist = torch.sqrt((diff * diff * (1.0 / (cov + 1e-8)).unsqueeze(0)).sum(-1))
    min_d = dist.min(dim=-1).values
    e_norm, d_norm = torch.sigmoid(-energy), torch.exp(-min_d)
    mix = torch.softmax(w["combination_weights"], dim=0)
    return dict(energy=energy, distances=dist, min_distance=min_d, energy_norm=e_norm, distance_norm=d_norm,
                combined=mix[0] * e_norm + mix[1] * d_norm)


def tta_mean(logits_views: Tensor) -> Tensor:
    """mean of logits over augmentation views [V,B,C] -> [B,C]  (src/eval.py:186-190, README.md:152)."""
    return logits_views.mean(dim=0)


def find_optimal_temperature(logits: Tensor, labels: Tensor) -> float:
    """100-point sweep over logspace(-1, 2) minimising mean|maxprob - correct| (src/eval.py:48-67);
    the first minimum wins (strict <)."""
    best_t, best = 1.0, float("inf")
    for t in torch.logspace(-1, 2, 100):
        p = torch.softmax(logits / t, dim=-1)
        conf, pred = p.max(dim=-1)
        ece = (conf - (pred == labels).float()).abs().mean().item()
        if ece < best:
            best, best_t = ece, float(t)
    return best_t


# --------------------------------------------------------------------------------------------------
# whole head  (src/train.py:145-168)
# --------------------------------------------------------------------------------------------------
def head_forward(a_hid: Tensor, t_hid: Tensor, a_mask: Optional[Tensor], t_mask: Optional[Tensor], labels: Tensor,
                 weights: Dict[str, W], num_classes: int, num_layers: int = 35) -> Dict[str, Tensor]:
    """weights: {'adapter_a','adapter_t','cross','pool_a','pool_t','fusion','classifier','prototypes'}."""
    a = adapter(a_hid, weights["adapter_a"])
    t = adapter(t_hid, weights["adapter_t"])
    a_enh, t_enh = cross_attention(a, t, a_mask, t_mask, weights["cross"])
    av = attentive_stats_pooling(a_enh, a_mask, weights["pool_a"])
    tv = attentive_stats_pooling(t_enh, t_mask, weights["pool_t"])
    fused = fusion(av, tv, weights["fusion"])
    logits, unc, anchor_loss = classifier(fused, weights["classifier"], num_layers, use_openmax=False, training=True,
                                          return_uncertainty=True)
    out = train_loss(logits, unc, anchor_loss, fused, labels, weights["prototypes"]["prototypes"], num_classes)
    out.update({"logits": logits, "unc": unc, "fused": fused, "a_enh": a_enh, "t_enh": t_enh, "a_vec": av, "t_vec": tv})
    return out
