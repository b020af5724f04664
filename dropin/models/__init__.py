"""`models` -- import-path shim that makes the reference's scripts pick up the B200 fusion head UNCHANGED.

The reference's scripts import the hot-path classes by module path (src/train.py:4-9, src/eval.py, train_crema.py):

    from models import AudioEncoder, TextEncoder, FusionLayer, Classifier
    from models.classifier import OpenMaxClassifier, AdvancedOpenMaxClassifier
    from models.cross_attention import CrossModalAttention
    from models.pooling import AttentiveStatsPooling
    from models.losses import LabelSmoothingCrossEntropy, ClassBalancedFocalLoss, SupConLoss
    from models.prototypes import PrototypeMemory

Put THIS directory's parent in front of the reference's `src/` on the import path

    PYTHONPATH=/path/to/repo/dropin:/path/to/repo  python /path/to/reference/src/train.py ...

and those statements resolve to the drop-in modules of `mmser_b200.models` (same constructors, forward signatures and
state_dict keys; sm_100a kernels behind a C-ABI).  Everything the head does not own -- the encoders, the CPU front end,
the legacy `Classifier` / `OpenMaxClassifier` the scripts import but never build -- is still served by the reference's
own `src/models` directory: it is appended to this package's search path, so `models.audio_encoder` etc. resolve there
(module names the shim defines win).  The reference tree is found through $SER_REFERENCE_SRC (its `src/` directory) or
by scanning sys.path for another `models/` directory.  The only edit a script needs is the adapter hand-over shown in
INTEGRATION.md (`audio_encoder.adapter = models.adapter.BottleneckAdapter()`), because the adapters live inside the
reference's encoder classes (src/models/audio_encoder.py:19-21, src/models/text_encoder.py:17-19).
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.abspath(os.path.join(_HERE, "..", ".."))
if _REPO not in sys.path:
    sys.path.append(_REPO)                      # mmser_b200.py (the importable alias of the package) lives there


def _reference_models_dir():
    cand = []
    env = os.environ.get("SER_REFERENCE_SRC")
    if env:
        cand.append(os.path.join(env, "models"))
    for p in sys.path:
        d = os.path.join(p or ".", "models")
        if os.path.abspath(d) != _HERE:
            cand.append(d)
    for d in cand:
        if os.path.isfile(os.path.join(d, "audio_encoder.py")):
            return os.path.abspath(d)
    return None


REFERENCE_MODELS_DIR = _reference_models_dir()
if REFERENCE_MODELS_DIR is not None:
    __path__.append(REFERENCE_MODELS_DIR)       # submodules the shim does not define come from the reference

from .fusion import FusionLayer  # noqa: E402

__all__ = ["AudioEncoder", "TextEncoder", "FusionLayer", "Classifier"]

_LAZY = {"AudioEncoder": "audio_encoder", "TextEncoder": "text_encoder", "Classifier": "classifier"}


def __getattr__(name):
    # The encoders (and the legacy Classifier) belong to the reference: imported on first use, from its own files.
    if name in _LAZY:
        if REFERENCE_MODELS_DIR is None:
            raise ImportError(f"models.{name} is part of the reference (src/models/{_LAZY[name]}.py), which was not found: "
                              "set SER_REFERENCE_SRC to the reference's src/ directory")
        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
