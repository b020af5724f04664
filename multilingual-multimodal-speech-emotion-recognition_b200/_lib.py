"""ctypes binding of libser_head.so (the C-ABI declared in include/ser_head.h).

The library is the only compute path: there is no PyTorch / CPU fallback.  If the shared object is
missing or fails to load, every op raises -- loudly -- instead of silently running something else.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libser_head.so")

SER_F32, SER_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3
GATE_NONE, GATE_RELU, GATE_TANH = 0, 1, 2


class SerError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("A", C.c_void_p), ("lda", C.c_longlong), ("a_trans", C.c_int),
        ("B", C.c_void_p), ("ldb", C.c_longlong), ("b_trans", C.c_int),
        ("C", C.c_void_p), ("ldc", C.c_longlong), ("c_f32", C.c_int),
        ("bias", C.c_void_p),
        ("R", C.c_void_p), ("ldr", C.c_longlong), ("r_f32", C.c_int),
        ("G", C.c_void_p), ("ldg", C.c_longlong), ("g_f32", C.c_int), ("gate_mode", C.c_int),
        ("act", C.c_int), ("accumulate", C.c_int), ("alpha", C.c_float), ("splits", C.c_int),
    ]


_lib = None


def load():
    """Load libser_head.so (building it is `python build.py` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SerError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run __graft_entry__.build()). "
            "There is no fallback path."
        )
    lib = C.CDLL(LIB_PATH)
    lib.ser_version.restype = C.c_int
    lib.ser_last_error.restype = C.c_char_p
    lib.ser_sm_count.restype = C.c_int
    lib.ser_gemm.restype = C.c_int
    lib.ser_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ser_last_error().decode("utf-8", "replace")
        raise SerError(f"{what} failed with code {rc}: {msg}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return SER_F32
    if t == torch.bfloat16:
        return SER_BF16
    raise SerError(f"unsupported dtype {t}: the fusion head runs in float32 or bfloat16")


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SerError(
                "the B200 fusion head has no CPU path: tensors must live on a CUDA device "
                f"(got {t.device})"
            )


def gemm(a, b, *, a_trans=False, b_trans=False, bias=None, act=ACT_NONE, residual=None, gate=None,
         gate_mode=GATE_NONE, out=None, out_dtype=None, accumulate=False, alpha=1.0, splits=0):
    """out[M,N] = epilogue(alpha * op(a) @ op(b)^T).  a: [M,K] (or [K,M] if a_trans); b: [N,K] (or [K,N])."""
    lib = load()
    require_cuda(a, b, bias, residual, gate, out)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_trans else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_trans else (b.shape[0], b.shape[1])
    assert K == Kb, (a.shape, b.shape, a_trans, b_trans)
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype or a.dtype)
    assert out.stride(1) == 1 and tuple(out.shape) == (M, N)
    d = GemmDesc()
    d.dtype = dtype_code(a.dtype)
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_trans = ptr(a), a.stride(0), int(a_trans)
    d.B, d.ldb, d.b_trans = ptr(b), b.stride(0), int(b_trans)
    d.C, d.ldc, d.c_f32 = ptr(out), out.stride(0), int(out.dtype == torch.float32)
    d.bias = ptr(bias)
    if residual is not None:
        d.R, d.ldr, d.r_f32 = ptr(residual), residual.stride(0), int(residual.dtype == torch.float32)
    if gate is not None:
        d.G, d.ldg, d.g_f32 = ptr(gate), gate.stride(0), int(gate.dtype == torch.float32)
    d.gate_mode = gate_mode
    d.act, d.accumulate, d.alpha, d.splits = act, int(accumulate), float(alpha), splits
    check(lib.ser_gemm(C.byref(d), stream_ptr(a.device)), "ser_gemm")
    return out
