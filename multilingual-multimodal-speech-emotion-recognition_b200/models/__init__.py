"""Drop-in replacements for the hot-path classes of the reference's src/models package."""
from .adapter import BottleneckAdapter
from .classifier import AdvancedOpenMaxClassifier, ClassAnchorClustering, DeepClassifier, DeepResidualBlock
from .cross_attention import CrossModalAttention
from .feature_fusion import UtteranceFeatureFusion
from .fusion import FusionLayer
from .losses import ClassBalancedFocalLoss, LabelSmoothingCrossEntropy, SupConLoss
from .ood import EnergyBasedOODDetector, LateOODResult, LateStageOODDetector, OODReason, PrototypeDistanceOODDetector
from .pooling import AttentiveStatsPooling
from .prototypes import PrototypeMemory

__all__ = [
    "BottleneckAdapter", "AdvancedOpenMaxClassifier", "ClassAnchorClustering", "DeepClassifier", "DeepResidualBlock",
    "CrossModalAttention", "FusionLayer", "ClassBalancedFocalLoss", "LabelSmoothingCrossEntropy", "SupConLoss",
    "AttentiveStatsPooling", "PrototypeMemory", "UtteranceFeatureFusion", "EnergyBasedOODDetector", "PrototypeDistanceOODDetector",
    "LateStageOODDetector", "LateOODResult", "OODReason",
]
