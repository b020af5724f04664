"""`models.classifier` of the reference (src/models/classifier.py:8-305) served by the B200 drop-in.

The hot-path classes come from mmser_b200.models.classifier.  `Classifier` and `OpenMaxClassifier` (classifier.py:308-436)
are legacy 3-layer MLPs the scripts import (src/train.py:4-5) but never construct: they are not part of the head and are
forwarded, on first access, to the reference's own file."""
import importlib.util
import os

from mmser_b200.models.classifier import (AdvancedOpenMaxClassifier, ClassAnchorClustering, DeepClassifier,  # noqa: F401
                                          DeepResidualBlock)

__all__ = ["AdvancedOpenMaxClassifier", "ClassAnchorClustering", "DeepClassifier", "DeepResidualBlock", "Classifier",
           "OpenMaxClassifier"]
_ref = None


def __getattr__(name):
    global _ref
    if name in ("Classifier", "OpenMaxClassifier"):
        if _ref is None:
            from . import REFERENCE_MODELS_DIR
            if REFERENCE_MODELS_DIR is None:
                raise ImportError(f"models.classifier.{name} is a legacy class of the reference that the fusion head does "
                                  "not replace: set SER_REFERENCE_SRC to the reference's src/ directory")
            spec = importlib.util.spec_from_file_location("_reference_models_classifier",
                                                          os.path.join(REFERENCE_MODELS_DIR, "classifier.py"))
            _ref = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(_ref)
        return getattr(_ref, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
