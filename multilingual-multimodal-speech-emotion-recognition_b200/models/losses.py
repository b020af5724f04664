"""Drop-in LabelSmoothingCrossEntropy, ClassBalancedFocalLoss and SupConLoss (reference: src/models/losses.py:7-88)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ..functional import HeadLossFn, SupConFn


class LabelSmoothingCrossEntropy(nn.Module):
    def __init__(self, smoothing: float = 0.1):
        super().__init__()
        self.smoothing = smoothing

    def forward(self, logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        terms = HeadLossFn.apply(logits, None, None, None, target,
                                 dict(w_ce=1.0, smoothing=self.smoothing, focal_use_weights=0))
        return terms[4]


class ClassBalancedFocalLoss(nn.Module):
    def __init__(self, beta: float = 0.9999, gamma: float = 2.0, num_classes: Optional[int] = None):
        super().__init__()
        self.beta, self.gamma, self.num_classes = beta, gamma, num_classes
        self.register_buffer("effective_num", torch.tensor(1.0))

    def forward(self, logits: torch.Tensor, targets: torch.Tensor, class_counts: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`class_counts` (optional, [C]) overrides the per-batch bincount -- data-parallel callers pass the global one."""
        cfg = dict(w_focal=1.0, beta=self.beta, gamma=self.gamma, focal_use_weights=int(self.num_classes is not None))
        if class_counts is not None:
            cfg["counts"] = class_counts
        terms = HeadLossFn.apply(logits, None, None, None, targets, cfg)
        return terms[4]


class SupConLoss(nn.Module):
    """Supervised contrastive loss (losses.py:67-88).  The reference's scripts import and construct it
    (train.py:8,86; train_crema.py:31,231) without adding it to the training loss; same constructor and call."""

    def __init__(self, temperature: float = 0.07):
        super().__init__()
        self.temperature = temperature

    def forward(self, features: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        return SupConFn.apply(features, labels, self.temperature)
