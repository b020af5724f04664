"""Deterministic synthetic weights and inputs of the fusion head: the benchmark's workload generator.

Shared by bench.py (every arm is fed the same tensors), the golden generator and the tests (oracle/synth.py re-exports
this module).  Weights are generated per state-dict key from a name-seeded CPU generator, so the golden fixtures
never need to store the 24.4 M parameters: any process with the same torch build regenerates them bit-for-bit.
Shapes and key names follow the reference modules; inputs follow SURVEY.md section 8(d).
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, Optional, Tuple

import torch

D = 768          # hidden size of wav2vec2-base / xlm-roberta-base
S = 256          # shared attention dim / adapter bottleneck
P = 512          # fusion / classifier width
HID_POOL = 128


def _gen(name: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def _fill(name: str, shape: Tuple[int, ...], kind: str, seed: int) -> torch.Tensor:
    g = _gen(name, seed)
    if kind == "linear_w":            # ~ default nn.Linear scale
        bound = 1.0 / math.sqrt(shape[-1])
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    if kind == "xavier_w":            # classifier._init_weights (classifier.py:134-138)
        bound = math.sqrt(6.0 / (shape[0] + shape[1]))
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    if kind == "bias":                # non-zero on purpose: exercises every bias path
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    if kind == "ln_w":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if kind == "ln_b":
        return 0.05 * torch.randn(shape, generator=g)
    if kind == "randn":
        return torch.randn(shape, generator=g)
    if kind == "proto":
        return 0.5 * torch.randn(shape, generator=g)
    raise ValueError(kind)


def adapter_weights(tag: str, seed: int = 0) -> Dict[str, torch.Tensor]:
    return {
        "0.weight": _fill(f"{tag}.0.weight", (S, D), "linear_w", seed),
        "0.bias": _fill(f"{tag}.0.bias", (S,), "bias", seed),
        "2.weight": _fill(f"{tag}.2.weight", (D, S), "linear_w", seed),
        "2.bias": _fill(f"{tag}.2.bias", (D,), "bias", seed),
    }


def feature_fusion_weights(tag: str, num_features: int, seed: int = 0, hidden: int = D) -> Dict[str, torch.Tensor]:
    """nn.Sequential(nn.Linear(hid + F, hid), nn.ReLU(), nn.Dropout(0.1)): audio_encoder.py:29-52, text_encoder.py:26-30."""
    return {
        "0.weight": _fill(f"{tag}.0.weight", (hidden, hidden + num_features), "linear_w", seed),
        "0.bias": _fill(f"{tag}.0.bias", (hidden,), "bias", seed),
    }


def cross_weights(seed: int = 0, audio_dim: int = D, text_dim: int = D) -> Dict[str, torch.Tensor]:
    w = {}
    dim = {"a": audio_dim, "t": text_dim}
    for n in ("q_a", "k_t", "v_t", "q_t", "k_a", "v_a"):
        w[f"{n}.weight"] = _fill(f"cross.{n}.weight", (S, dim[n[-1]]), "linear_w", seed)
        w[f"{n}.bias"] = _fill(f"cross.{n}.bias", (S,), "bias", seed)
    for n in ("attn_a", "attn_t"):
        w[f"{n}.in_proj_weight"] = _fill(f"cross.{n}.in_proj_weight", (3 * S, S), "linear_w", seed)
        w[f"{n}.in_proj_bias"] = _fill(f"cross.{n}.in_proj_bias", (3 * S,), "bias", seed)
        w[f"{n}.out_proj.weight"] = _fill(f"cross.{n}.out_proj.weight", (S, S), "linear_w", seed)
        w[f"{n}.out_proj.bias"] = _fill(f"cross.{n}.out_proj.bias", (S,), "bias", seed)
    for n in ("out_a", "out_t"):
        w[f"{n}.weight"] = _fill(f"cross.{n}.weight", (dim[n[-1]], S), "linear_w", seed)
        w[f"{n}.bias"] = _fill(f"cross.{n}.bias", (dim[n[-1]],), "bias", seed)
    for n in ("norm_a", "norm_t"):
        w[f"{n}.weight"] = _fill(f"cross.{n}.weight", (dim[n[-1]],), "ln_w", seed)
        w[f"{n}.bias"] = _fill(f"cross.{n}.bias", (dim[n[-1]],), "ln_b", seed)
    return w


def pool_weights(tag: str, seed: int = 0) -> Dict[str, torch.Tensor]:
    return {
        "attention.0.weight": _fill(f"{tag}.attention.0.weight", (HID_POOL, D), "linear_w", seed),
        "attention.0.bias": _fill(f"{tag}.attention.0.bias", (HID_POOL,), "bias", seed),
        "attention.2.weight": _fill(f"{tag}.attention.2.weight", (1, HID_POOL), "linear_w", seed) * 4.0,
        "attention.2.bias": _fill(f"{tag}.attention.2.bias", (1,), "bias", seed),
    }


def fusion_weights(seed: int = 0, audio_dim: int = 2 * D, text_dim: int = 2 * D) -> Dict[str, torch.Tensor]:
    w = {}
    for m in ("a", "t"):
        w[f"proj_{m}.0.weight"] = _fill(f"fusion.proj_{m}.0.weight", (P, audio_dim if m == "a" else text_dim), "linear_w", seed)
        w[f"proj_{m}.0.bias"] = _fill(f"fusion.proj_{m}.0.bias", (P,), "bias", seed)
        w[f"proj_{m}.3.weight"] = _fill(f"fusion.proj_{m}.3.weight", (P, P), "linear_w", seed)
        w[f"proj_{m}.3.bias"] = _fill(f"fusion.proj_{m}.3.bias", (P,), "bias", seed)
        w[f"gate_{m}.0.weight"] = _fill(f"fusion.gate_{m}.0.weight", (P // 2, P), "linear_w", seed)
        w[f"gate_{m}.0.bias"] = _fill(f"fusion.gate_{m}.0.bias", (P // 2,), "bias", seed)
        w[f"gate_{m}.2.weight"] = _fill(f"fusion.gate_{m}.2.weight", (1, P // 2), "linear_w", seed) * 4.0
        w[f"gate_{m}.2.bias"] = _fill(f"fusion.gate_{m}.2.bias", (1,), "bias", seed)
    return w


def classifier_weights(num_labels: int, num_layers: int = 35, seed: int = 0, input_dim: int = P) -> Dict[str, torch.Tensor]:
    w = {}
    p = "deep_classifier."
    w[p + "input_projection.0.weight"] = _fill("clf.in.0.weight", (P, input_dim), "xavier_w", seed)
    w[p + "input_projection.0.bias"] = _fill("clf.in.0.bias", (P,), "bias", seed)
    w[p + "input_projection.1.weight"] = _fill("clf.in.1.weight", (P,), "ln_w", seed)
    w[p + "input_projection.1.bias"] = _fill("clf.in.1.bias", (P,), "ln_b", seed)
    for i in range(num_layers):
        b = f"{p}residual_layers.{i}.block."
        w[b + "0.weight"] = _fill(f"clf.res{i}.0.weight", (P,), "ln_w", seed)
        w[b + "0.bias"] = _fill(f"clf.res{i}.0.bias", (P,), "ln_b", seed)
        w[b + "1.weight"] = _fill(f"clf.res{i}.1.weight", (P, P), "xavier_w", seed)
        w[b + "1.bias"] = _fill(f"clf.res{i}.1.bias", (P,), "bias", seed)
        w[b + "4.weight"] = _fill(f"clf.res{i}.4.weight", (P, P), "xavier_w", seed)
        w[b + "4.bias"] = _fill(f"clf.res{i}.4.bias", (P,), "bias", seed)
    for i in range(num_layers):
        w[f"{p}layer_norms.{i}.weight"] = _fill(f"clf.ln{i}.weight", (P,), "ln_w", seed)
        w[f"{p}layer_norms.{i}.bias"] = _fill(f"clf.ln{i}.bias", (P,), "ln_b", seed)
    w[p + "output_projection.0.weight"] = _fill("clf.out.0.weight", (P // 2, P), "xavier_w", seed)
    w[p + "output_projection.0.bias"] = _fill("clf.out.0.bias", (P // 2,), "bias", seed)
    w[p + "output_projection.1.weight"] = _fill("clf.out.1.weight", (P // 2,), "ln_w", seed)
    w[p + "output_projection.1.bias"] = _fill("clf.out.1.bias", (P // 2,), "ln_b", seed)
    w[p + "output_projection.4.weight"] = _fill("clf.out.4.weight", (num_labels, P // 2), "xavier_w", seed) * 3.0
    w[p + "output_projection.4.bias"] = _fill("clf.out.4.bias", (num_labels,), "bias", seed)
    a = "anchor_clustering."
    w[a + "class_anchors"] = _fill("clf.anchors", (num_labels, 128), "randn", seed)
    w[a + "anchor_projection.0.weight"] = _fill("clf.anchor.0.weight", (128, P // 2), "linear_w", seed)
    w[a + "anchor_projection.0.bias"] = _fill("clf.anchor.0.bias", (128,), "bias", seed)
    w[a + "anchor_projection.1.weight"] = _fill("clf.anchor.1.weight", (128,), "ln_w", seed)
    w[a + "anchor_projection.1.bias"] = _fill("clf.anchor.1.bias", (128,), "ln_b", seed)
    w[a + "temperature"] = torch.tensor(1.0)
    w["weibull_alpha"] = torch.ones(num_labels)
    w["weibull_beta"] = torch.ones(num_labels)
    w["weibull_tau"] = torch.zeros(num_labels)
    w["activation_vectors"] = torch.zeros(num_labels, P // 2)
    w["uncertainty_head.0.weight"] = _fill("clf.unc.0.weight", (64, P // 2), "linear_w", seed)
    w["uncertainty_head.0.bias"] = _fill("clf.unc.0.bias", (64,), "bias", seed)
    w["uncertainty_head.3.weight"] = _fill("clf.unc.3.weight", (1, 64), "linear_w", seed) * 2.0
    w["uncertainty_head.3.bias"] = _fill("clf.unc.3.bias", (1,), "bias", seed)
    return w


CLASSIFIER_BUFFERS = ("weibull_alpha", "weibull_beta", "weibull_tau", "activation_vectors")


def head_weights(num_labels: int, num_layers: int = 35, seed: int = 0) -> Dict[str, Dict[str, torch.Tensor]]:
    return {
        "adapter_a": adapter_weights("adapter_a", seed),
        "adapter_t": adapter_weights("adapter_t", seed),
        "cross": cross_weights(seed),
        "pool_a": pool_weights("pool_a", seed),
        "pool_t": pool_weights("pool_t", seed),
        "fusion": fusion_weights(seed),
        "classifier": classifier_weights(num_labels, num_layers, seed),
        "prototypes": {"prototypes": _fill("prototypes", (num_labels, P), "proto", seed)},
    }


def make_inputs(B: int, Ta: int, Tt: int, C: int, seed: int = 1234, with_masks: bool = True):
    """unit-normal hidden states, right-padded float masks with >= 1 valid token,
    padded positions zero-filled, int64 labels."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    a = torch.randn(B, Ta, D, generator=g)
    t = torch.randn(B, Tt, D, generator=g)
    labels = torch.randint(0, C, (B,), generator=g)
    a_mask: Optional[torch.Tensor] = None
    t_mask: Optional[torch.Tensor] = None
    if with_masks:
        len_a = torch.randint((Ta + 1) // 2, Ta + 1, (B,), generator=g)
        len_t = torch.randint(max(1, (Tt + 3) // 4), Tt + 1, (B,), generator=g)
        a_mask = (torch.arange(Ta)[None, :] < len_a[:, None]).float()
        t_mask = (torch.arange(Tt)[None, :] < len_t[:, None]).float()
        a = a * a_mask[..., None]
        t = t * t_mask[..., None]
    return a, t, a_mask, t_mask, labels
