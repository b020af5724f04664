"""`models.fusion` of the reference (src/models/fusion.py:5-25) served by the B200 drop-in (mmser_b200.models.fusion)."""
from mmser_b200.models.fusion import FusionLayer  # noqa: F401

__all__ = ['FusionLayer']
