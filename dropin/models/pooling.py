"""`models.pooling` of the reference (src/models/pooling.py:6-28) served by the B200 drop-in (mmser_b200.models.pooling)."""
from mmser_b200.models.pooling import AttentiveStatsPooling  # noqa: F401

__all__ = ['AttentiveStatsPooling']
