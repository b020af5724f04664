"""`models.cross_attention` of the reference (src/models/cross_attention.py:6-53) served by the B200 drop-in (mmser_b200.models.cross_attention)."""
from mmser_b200.models.cross_attention import CrossModalAttention  # noqa: F401

__all__ = ['CrossModalAttention']
