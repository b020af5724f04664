"""Drop-in PrototypeMemory (reference: src/models/prototypes.py:5-53)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..functional import HeadLossFn


class PrototypeMemory(nn.Module):
    def __init__(self, num_classes: int, dim: int):
        super().__init__()
        self.prototypes = nn.Parameter(torch.randn(num_classes, dim) * 0.02)

    def forward(self) -> torch.Tensor:
        return self.prototypes

    def prototype_loss(self, embeddings: torch.Tensor, labels: torch.Tensor, margin: float = 0.5) -> torch.Tensor:
        """mean ||e - P_y|| + margin - mean softmin_c d(e, P_c), with the reference's clamp / own-class quirks."""
        C = self.prototypes.shape[0]
        dummy_logits = torch.zeros(embeddings.shape[0], C, device=embeddings.device, dtype=torch.float32)
        terms = HeadLossFn.apply(dummy_logits, None, embeddings, self.prototypes, labels,
                                 dict(w_proto=1.0, margin=margin, focal_use_weights=0))
        return terms[4]
