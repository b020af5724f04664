"""In-tree build of libser_head.so (CUDA kernels + C-ABI) for sm_100a.

nvcc cross-compiles without a GPU.  Objects are rebuilt only when their source (or a header) is newer.
Usage: python build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libser_head.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", os.path.join(HERE, "..", "include"),
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, verbose):
    out = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", out]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers_mtime()
    # a shipped library that is newer than every source and header is current even when the object files did not
    # travel with it (the build/ directory is not part of a snapshot)
    newest = max([hdr] + [os.path.getmtime(os.path.join(CSRC, s)) for s in _sources()])
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    todo = []
    for src in _sources():
        obj = os.path.join(OBJ, src[:-3] + ".o")
        sm = max(os.path.getmtime(os.path.join(CSRC, src)), hdr)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < sm:
            todo.append(src)
    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for src, rc, log in ex.map(lambda s: _compile(s, verbose), todo):
                if log.strip() and (verbose or rc != 0):
                    print(f"--- {src}\n{log}", file=sys.stderr)
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in _sources()]
    if todo or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout + r.stderr, file=sys.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
