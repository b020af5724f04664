"""`models.losses` of the reference (src/models/losses.py:7-88) served by the B200 drop-in (mmser_b200.models.losses)."""
from mmser_b200.models.losses import LabelSmoothingCrossEntropy, ClassBalancedFocalLoss, SupConLoss  # noqa: F401

__all__ = ['LabelSmoothingCrossEntropy', 'ClassBalancedFocalLoss', 'SupConLoss']
