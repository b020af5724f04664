"""Builds tools/probe.cu (hardware probes: debug aids, not product code) into tools/_build/libser_probe.so."""
import ctypes
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mmser_b200 import _lib as L  # noqa: E402


def build_probe():
    """tools/probe.cu is a debug aid, not product code: it is compiled into its own shared object (git-ignored,
    tools/_build/) that links against libser_head.so for the error / launch-count helpers."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.dirname(L.LIB_PATH)
    out = os.path.join(here, "_build", "libser_probe.so")
    src = os.path.join(here, "probe.cu")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                               "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-shared",
                               "-I", os.path.join(pkg, "csrc"), "-I", os.path.join(pkg, "..", "include"), src, "-o", out,
                               "-L", pkg, "-lser_head", f"-Xlinker=-rpath={pkg}", "-lcudart", "-lcuda"])
    L.load()
    return ctypes.CDLL(out)


