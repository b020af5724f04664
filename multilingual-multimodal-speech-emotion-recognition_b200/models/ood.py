"""Late-stage OOD scoring (reference: src/models/dual_gate_ood.py:187-413): free energy of
the temperature-scaled logits, diagonal-Mahalanobis distances of the features to learnable class prototypes, and the
learnable two-way mix of the two normalised scores.  Same constructors, parameter names (state_dict keys
`energy_detector.temperature`, `prototype_detector.prototypes`, `prototype_detector.covariances`, `combination_weights`)
and return types as the reference; the arithmetic of a call is ONE kernel launch (csrc/eval.cu late_ood_kernel) instead
of a Python loop over classes, and `scores()` hands the per-sample tensors out without any host synchronisation
(`forward` reduces them to the reference's LateOODResult of Python scalars, one device->host copy)."""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from .. import functional as SF


class OODReason(Enum):
    """The late-stage members of the reference's OODReason (dual_gate_ood.py:18-31), same values."""
    HIGH_ENERGY = "high_energy"
    HIGH_PROTOTYPE_DISTANCE = "high_prototype_distance"
    COMBINED_THRESHOLD = "combined_threshold"


@dataclass
class LateOODResult:
    """dual_gate_ood.py:43-51"""
    is_ood: bool
    energy_score: float
    prototype_distance: float
    combined_score: float
    confidence_score: float
    reason: Optional[OODReason]


class EnergyBasedOODDetector(nn.Module):
    """dual_gate_ood.py:187-220: (energy [B], logits / temperature)."""

    def __init__(self, temperature: float = 1.0, energy_threshold: float = 0.5):
        super().__init__()
        self.temperature = nn.Parameter(torch.tensor(temperature))
        self.energy_threshold = energy_threshold

    def forward(self, logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        post = SF.eval_post(logits / self.temperature.detach(), 1.0)
        return post["energy"], post["mean_logits"]


class PrototypeDistanceOODDetector(nn.Module):
    """dual_gate_ood.py:246-329: (distances [B,C], min over classes [B])."""

    def __init__(self, num_classes: int, feature_dim: int, distance_threshold: float = 2.0):
        super().__init__()
        self.num_classes, self.feature_dim, self.distance_threshold = num_classes, feature_dim, distance_threshold
        self.prototypes = nn.Parameter(torch.randn(num_classes, feature_dim))
        self.covariances = nn.Parameter(torch.ones(num_classes, feature_dim))
        nn.init.xavier_uniform_(self.prototypes)          # :272-278
        nn.init.ones_(self.covariances)

    def forward(self, features: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        dev = features.device
        out = SF.late_ood(torch.zeros(features.shape[0], self.num_classes, device=dev), features, self.prototypes,
                          self.covariances, torch.ones((), device=dev), torch.zeros(2, device=dev))
        return out["distances"], out["min_distance"]

    @torch.no_grad()
    def update_prototypes(self, features: torch.Tensor, labels: torch.Tensor) -> None:
        """Class means and variances of the given features (:314-328; torch.var = unbiased)."""
        features = features.float()
        for i in range(self.num_classes):
            sel = labels == i
            if int(sel.sum()) > 0:
                cf = features[sel]
                self.prototypes[i].copy_(cf.mean(dim=0))
                self.covariances[i].copy_(cf.var(dim=0) + 1e-8)


class LateStageOODDetector(nn.Module):
    """dual_gate_ood.py:331-413."""

    def __init__(self, num_classes: int, feature_dim: int, energy_weight: float = 0.6, prototype_weight: float = 0.4,
                 combined_threshold: float = 0.5):
        super().__init__()
        self.energy_weight, self.prototype_weight, self.combined_threshold = energy_weight, prototype_weight, combined_threshold
        self.energy_detector = EnergyBasedOODDetector()
        self.prototype_detector = PrototypeDistanceOODDetector(num_classes, feature_dim)
        self.combination_weights = nn.Parameter(torch.tensor([energy_weight, prototype_weight]))

    def scores(self, logits: torch.Tensor, features: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Per-sample tensors (energy, distances, min_distance, energy_norm, distance_norm, combined, is_ood), no sync."""
        out = SF.late_ood(logits, features, self.prototype_detector.prototypes, self.prototype_detector.covariances,
                          self.energy_detector.temperature, self.combination_weights)
        out["is_ood"] = out["combined"] < self.combined_threshold
        return out

    def forward(self, logits: torch.Tensor, features: torch.Tensor) -> LateOODResult:
        s = self.scores(logits, features)
        host = torch.stack([s["energy"].mean(), s["min_distance"].mean(), s["combined"].mean(), s["energy_norm"].mean(),
                            s["distance_norm"].mean(), s["is_ood"].any().float()]).cpu().tolist()
        energy, dist, combined, e_norm, d_norm, any_ood = host
        if e_norm < 0.3:
            reason = OODReason.HIGH_ENERGY
        elif d_norm < 0.3:
            reason = OODReason.HIGH_PROTOTYPE_DISTANCE
        else:
            reason = OODReason.COMBINED_THRESHOLD
        return LateOODResult(is_ood=bool(any_ood), energy_score=energy, prototype_distance=dist, combined_score=combined,
                             confidence_score=combined, reason=reason)
