"""The literal drop-in: the reference's import paths resolve to the B200 head through the `dropin/models` shim, and a
training run written like src/train.py (imports, construction, 10-group AdamW, LambdaLR, autocast + GradScaler, the
six-call forward, the hand-unrolled Weibull walk) runs unchanged against it."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT]))


def test_reference_import_paths_resolve_to_the_dropin():
    code = (
        "from models import FusionLayer\n"
        "from models.classifier import AdvancedOpenMaxClassifier\n"
        "from models.cross_attention import CrossModalAttention\n"
        "from models.pooling import AttentiveStatsPooling\n"
        "from models.losses import LabelSmoothingCrossEntropy, ClassBalancedFocalLoss, SupConLoss\n"
        "from models.prototypes import PrototypeMemory\n"
        "import mmser_b200.models as M\n"
        "assert CrossModalAttention is M.CrossModalAttention and FusionLayer is M.FusionLayer\n"
        "assert AdvancedOpenMaxClassifier is M.AdvancedOpenMaxClassifier and PrototypeMemory is M.PrototypeMemory\n"
        "assert AttentiveStatsPooling is M.AttentiveStatsPooling and SupConLoss is M.SupConLoss\n"
        "import models\n"
        "try:\n"
        "    models.AudioEncoder\n"
        "except ImportError as e:\n"
        "    assert 'reference' in str(e) or 'librosa' in str(e) or 'No module' in str(e), e\n"
        "print('OK')\n")
    r = subprocess.run([sys.executable, "-c", code], env=ENV, cwd="/tmp", capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [[], ["--use_amp"], ["--use_amp", "--fused_optimizer"], ["--fused_optimizer"]],
                         ids=["adamw", "adamw_amp", "fused_amp", "fused"])
def test_train_py_style_run_on_the_dropin(flags):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_train_like_reference.py"), *flags], env=ENV,
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "RESULT ok=True" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
