"""Drop-in AdvancedOpenMaxClassifier (reference: src/models/classifier.py:8-305).

The module tree -- deep_classifier.{input_projection, residual_layers[i].block, layer_norms[i],
output_projection}, anchor_clustering.{class_anchors, anchor_projection, temperature}, uncertainty_head and
the Weibull buffers -- reproduces the reference's names and shapes so checkpoints interchange and so the
training scripts can keep reaching into the children (src/train.py:79-81, 221-236).  forward() bypasses the
children and runs the whole 35-block stack, the heads and (at inference) OpenMax through the C-ABI.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .._params import FlatParams
from .. import functional as SF
from ._common import DropoutSeed, LayerNorm, Linear, active_dropout


class ClassAnchorClustering(nn.Module):
    """Parameter container for the anchor head (classifier.py:8-70).  Its similarity output is discarded by the
    classifier and its loss is identically zero with exactly-zero gradients, so the fused path never evaluates it;
    calling the module directly still works through the library GEMM / LayerNorm."""

    def __init__(self, feature_dim: int, num_classes: int, anchor_dim: int = 128):
        super().__init__()
        self.feature_dim, self.num_classes, self.anchor_dim = feature_dim, num_classes, anchor_dim
        self.class_anchors = nn.Parameter(torch.randn(num_classes, anchor_dim))
        self.anchor_projection = nn.Sequential(Linear(feature_dim, anchor_dim), LayerNorm(anchor_dim), nn.ReLU(),
                                               nn.Dropout(0.1))
        self.temperature = nn.Parameter(torch.tensor(1.0))

    def forward(self, features: torch.Tensor):
        z = nn.functional.normalize(self.anchor_projection(features).float(), p=2, dim=1)
        anchors = nn.functional.normalize(self.class_anchors, p=2, dim=1)
        sims = SF.L.gemm(z.contiguous(), anchors.contiguous())
        loss = (sims - sims.max(dim=1, keepdim=True)[0]).clamp(min=0).mean()      # == 0 by construction
        return sims / self.temperature, loss


class DeepResidualBlock(nn.Module):
    def __init__(self, dim: int, dropout: float = 0.1):
        super().__init__()
        self.block = nn.Sequential(LayerNorm(dim), Linear(dim, dim), nn.ReLU(), nn.Dropout(dropout), Linear(dim, dim),
                                   nn.Dropout(dropout))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x + self.block(x)


class DeepClassifier(nn.Module):
    def __init__(self, input_dim: int, num_classes: int, num_layers: int = 35, base_dim: int = 512, dropout: float = 0.1):
        super().__init__()
        self.input_dim, self.num_classes, self.num_layers, self.base_dim = input_dim, num_classes, num_layers, base_dim
        self.input_projection = nn.Sequential(Linear(input_dim, base_dim), LayerNorm(base_dim), nn.ReLU(), nn.Dropout(dropout))
        self.residual_layers = nn.ModuleList([DeepResidualBlock(base_dim, dropout) for _ in range(num_layers)])
        self.layer_norms = nn.ModuleList([LayerNorm(base_dim) for _ in range(num_layers)])
        self.output_projection = nn.Sequential(Linear(base_dim, base_dim // 2), LayerNorm(base_dim // 2), nn.ReLU(),
                                               nn.Dropout(dropout), Linear(base_dim // 2, num_classes))
        for m in self.modules():                      # xavier-uniform weights, zero biases (classifier.py:134-138)
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.input_projection(x)
        for block, norm in zip(self.residual_layers, self.layer_norms):
            x = block(norm(x))
        return self.output_projection(x)


class AdvancedOpenMaxClassifier(nn.Module):
    def __init__(self, input_dim: int, num_labels: int, num_layers: int = 35, base_dim: int = 512, dropout: float = 0.1,
                 alpha: float = 20.0):
        super().__init__()
        # (input_dim may differ from base_dim, as in the reference: only input_projection[0] sees it; bf16 inputs need a
        #  multiple of 128 -- every reference script uses 512 / 512)
        self.num_labels, self.alpha, self.num_layers, self.p_drop = num_labels, alpha, num_layers, float(dropout)
        self.deep_classifier = DeepClassifier(input_dim, num_labels, num_layers, base_dim, dropout)
        self.anchor_clustering = ClassAnchorClustering(base_dim // 2, num_labels, anchor_dim=128)
        self.register_buffer("weibull_alpha", torch.ones(num_labels))
        self.register_buffer("weibull_beta", torch.ones(num_labels))
        self.register_buffer("weibull_tau", torch.zeros(num_labels))
        self.register_buffer("activation_vectors", torch.zeros(num_labels, base_dim // 2))
        self.uncertainty_head = nn.Sequential(Linear(base_dim // 2, 64), nn.ReLU(), nn.Dropout(dropout), Linear(64, 1),
                                              nn.Sigmoid())
        self._flat = FlatParams([(n, p) for n, p in self.named_parameters()])
        self.last_features: Optional[torch.Tensor] = None     # penultimate 256-d features of the last forward
        self._drop_seed = DropoutSeed()

    def forward(self, x: torch.Tensor, use_openmax: bool = True, return_uncertainty: bool = False):
        # every nn.Dropout of the stack and of the uncertainty head (classifier.py:83,85,109,127,195) runs in-kernel
        p = active_dropout(self, self.p_drop)
        seed = self._drop_seed.next(x.device) if p > 0.0 else None
        logits, unc, feats = SF.ClassifierFn.apply(x, self._flat, self.num_layers, return_uncertainty, p, seed,
                                                   *self._flat.params)
        self.last_features = feats
        # identically 0 in the reference (classifier.py:64-68); one cached device scalar instead of a fill per call
        z = self.__dict__.get("_zero_scalar")
        if z is None or z.device != x.device:
            z = torch.zeros((), device=x.device, dtype=torch.float32)
            self.__dict__["_zero_scalar"] = z
        anchor_loss = z
        if use_openmax and not self.training:
            logits = self.openmax_forward(feats, logits)
        if return_uncertainty:
            return logits, unc, anchor_loss
        return logits

    def openmax_forward(self, features: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
        return SF.openmax(features, logits, self.activation_vectors, self.weibull_alpha, self.weibull_beta,
                          self.weibull_tau)

    @torch.no_grad()
    def fit_weibull(self, features: torch.Tensor, labels: torch.Tensor) -> None:
        """Per-class mean activation vector, alpha = 2.5, beta = 1.5 * population-std of the distances,
        tau = 0.8 * min distance (classifier.py:277-305).  Host-side bookkeeping over a validation set."""
        features = features.float()
        for c in range(self.num_labels):
            sel = labels == c
            if int(sel.sum()) == 0:
                continue
            cf = features[sel]
            mean = cf.mean(dim=0)
            dist = (cf - mean).norm(dim=1)
            self.activation_vectors[c] = mean
            self.weibull_alpha[c] = 2.5
            self.weibull_beta[c] = dist.std(unbiased=False) * 1.5
            self.weibull_tau[c] = dist.min() * 0.8
