"""Importable alias of the package directory ``multilingual-multimodal-speech-emotion-recognition_b200/``.

The directory name required by the build contract contains hyphens and is therefore not a Python
identifier.  This module turns itself into that package: it sets ``__path__`` to the directory and
executes its ``__init__.py``, so ``import mmser_b200.models.fusion`` etc. resolve inside it.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                         "multilingual-multimodal-speech-emotion-recognition_b200")
__path__ = [_PKG_DIR]
__package__ = "mmser_b200"
if globals().get("__spec__") is not None:
    __spec__.submodule_search_locations = __path__      # makes importlib treat this module as a package
__file__ = _os.path.join(_PKG_DIR, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
