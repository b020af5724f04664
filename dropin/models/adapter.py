"""`models.adapter` of the reference (src/models/audio_encoder.py:19-21, text_encoder.py:17-19) served by the B200 drop-in (mmser_b200.models.adapter)."""
from mmser_b200.models.adapter import BottleneckAdapter  # noqa: F401

__all__ = ['BottleneckAdapter']
