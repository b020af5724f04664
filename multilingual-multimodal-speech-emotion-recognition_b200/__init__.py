"""B200-native fusion head (sm_100a CUDA behind a C-ABI) -- drop-in for the reference's src/models head.

Import as ``mmser_b200`` (see mmser_b200.py at the repo root: the on-disk directory name carries
hyphens, so the alias module gives it an importable name).
"""
from . import _lib  # noqa: F401
from . import functional  # noqa: F401
from . import models  # noqa: F401
from . import optim  # noqa: F401
from .head import FusionHead  # noqa: F401

__all__ = ["_lib", "functional", "models", "optim", "FusionHead"]
