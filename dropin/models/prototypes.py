"""`models.prototypes` of the reference (src/models/prototypes.py:5-53) served by the B200 drop-in (mmser_b200.models.prototypes)."""
from mmser_b200.models.prototypes import PrototypeMemory  # noqa: F401

__all__ = ['PrototypeMemory']
