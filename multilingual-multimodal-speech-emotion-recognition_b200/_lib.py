"""ctypes binding of libser_head.so (the C-ABI declared in include/ser_head.h).

The library is the only compute path: there is no PyTorch / CPU fallback.  If the shared object is
missing or fails to load, every op raises -- loudly -- instead of silently running something else.

The ctypes structures are generated from include/ser_head.h at import time, so the Python layout can
never drift from the C declaration.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libser_head.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "ser_head.h")

SER_F32, SER_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3
GATE_NONE, GATE_RELU, GATE_TANH = 0, 1, 2


class SerError(RuntimeError):
    pass


# --------------------------------------------------------------------------------------------------
# header -> ctypes
# --------------------------------------------------------------------------------------------------
_SCALARS = {"int": C.c_int, "float": C.c_float, "long long": C.c_longlong, "size_t": C.c_size_t}


def _parse_structs(text: str) -> Dict[str, type]:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    out = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        name, body = m.group(3), m.group(2)
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            if "*" in decl:                       # every pointer flavour is a void*
                for part in decl.split(","):
                    fname = re.findall(r"(\w+)\s*$", part.strip())[0]
                    fields.append((fname, C.c_void_p))
                continue
            for tname, ctype in _SCALARS.items():
                if decl.startswith(tname + " "):
                    for fname in decl[len(tname):].split(","):
                        fields.append((fname.strip(), ctype))
                    break
            else:
                raise SerError(f"cannot parse field declaration '{decl}' in {name}")
        out[name] = type(name, (C.Structure,), {"_fields_": fields})
    return out


def _parse_exports(text: str):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ser_\w+)\s*\(", text)))


with open(HEADER_PATH, "r") as _f:
    _HEADER = _f.read()
STRUCTS = _parse_structs(_HEADER)
EXPORTS = _parse_exports(_HEADER)

GemmDesc = STRUCTS["ser_gemm_desc"]
AdapterDesc = STRUCTS["ser_adapter_desc"]
XattnDesc = STRUCTS["ser_xattn_desc"]
AspDesc = STRUCTS["ser_asp_desc"]
FusionDesc = STRUCTS["ser_fusion_desc"]
ClfDesc = STRUCTS["ser_clf_desc"]
LossDesc = STRUCTS["ser_loss_desc"]
FeatFuseDesc = STRUCTS["ser_featfuse_desc"]
AttnDesc = STRUCTS["ser_attn_desc"]

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
    missing = [n for n in EXPORTS if not hasattr(lib, n)]
    if missing:
        raise SerError(f"libser_head.so does not export {missing}; rebuild it")
    for n in EXPORTS:
        getattr(lib, n).restype = C.c_int
    lib.ser_last_error.restype = C.c_char_p
    for n in ("ser_xattn_bwd_ws_bytes", "ser_fusion_bwd_ws_bytes", "ser_clf_bwd_ws_bytes", "ser_supcon_ws_bytes"):
        getattr(lib, n).restype = C.c_size_t
    P, I, LL, F = C.c_void_p, C.c_int, C.c_longlong, C.c_float
    lib.ser_gemm.argtypes = [C.POINTER(GemmDesc), P]
    lib.ser_cast.argtypes = [P, I, P, I, LL, P]
    lib.ser_cast_multi.argtypes = [I, P, P, P, P]
    lib.ser_set_reserved_sms.argtypes = [I]
    lib.ser_unpack_frames.argtypes = [P, P, P, P, I, I, I, I, P]
    lib.ser_layernorm_fwd.argtypes = [P, I, P, I, P, P, P, I, I, I, P]
    lib.ser_layernorm_bwd.argtypes = [P, I, P, I, P, P, P, P, I, P, P, I, I, I, P]
    lib.ser_colsum.argtypes = [P, I, LL, I, I, P, P]
    for n, S in (("adapter", AdapterDesc), ("xattn", XattnDesc), ("asp", AspDesc), ("fusion", FusionDesc),
                 ("clf", ClfDesc), ("featfuse", FeatFuseDesc)):
        getattr(lib, f"ser_{n}_fwd").argtypes = [C.POINTER(S), P]
        getattr(lib, f"ser_{n}_bwd").argtypes = [C.POINTER(S), P]
    for n in ("ser_attention_fwd", "ser_attention_bwd"):
        getattr(lib, n).argtypes = [C.POINTER(AttnDesc), P]
    for n in ("ser_loss_fwd", "ser_loss_finalize", "ser_loss_bwd"):
        getattr(lib, n).argtypes = [C.POINTER(LossDesc), P]
    lib.ser_xattn_bwd_ws_bytes.argtypes = [I] * 7
    lib.ser_xattn_folded.argtypes = [I] * 3
    lib.ser_fusion_bwd_ws_bytes.argtypes = [I] * 5
    lib.ser_clf_bwd_ws_bytes.argtypes = [I] * 6
    lib.ser_openmax_fwd.argtypes = [P, P, P, P, P, P, P, I, I, I, P]
    lib.ser_late_ood.argtypes = [P, P, I, P, P, P, P, P, P, I, I, I, P]
    lib.ser_eval_post.argtypes = [P, I, I, I, F, P, P, P, P, P]
    lib.ser_temperature_sweep.argtypes = [P, P, I, I, P, I, P, P]
    lib.ser_supcon_ws_bytes.argtypes = [I, I]
    lib.ser_supcon_fwd.argtypes = [P, I, P, I, I, F, P, P, C.c_size_t, P]
    lib.ser_supcon_bwd.argtypes = [P, I, P, I, I, F, P, P, I, P, C.c_size_t, P]
    lib.ser_adamw_multi.argtypes = [I, P, P, P, P, P, F, F, F, F, F, I, P, P]
    lib.ser_adamw_multi_amp.argtypes = [I, P, P, P, P, P, F, F, F, F, F, P, P, P, P, P]
    lib.ser_grad_clip_coef.argtypes = [I, P, P, F, P, P, P, P]
    lib.ser_desc_size.argtypes = [I]
    lib.ser_dropout_mask.argtypes = [P, I, F, LL, I, P, P]
    lib.ser_launch_count.restype = C.c_longlong
    lib.ser_prof_enable.argtypes = [I]
    lib.ser_prof_report.argtypes = [C.c_char_p, I]
    lib.ser_prof_stall.argtypes = [C.c_double, P]
    lib.ser_prof_null.argtypes = [I, P]
    for i, S in enumerate((GemmDesc, AdapterDesc, XattnDesc, AspDesc, FusionDesc, ClfDesc, LossDesc, FeatFuseDesc, AttnDesc)):
        if lib.ser_desc_size(i) != C.sizeof(S):
            raise SerError(f"layout mismatch for {S.__name__}: C {lib.ser_desc_size(i)} vs ctypes {C.sizeof(S)}")
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


def fill(desc, keep: list, **fields):
    """Set descriptor fields.  Tensors become device pointers; lists of tensors become host arrays of
    device pointers (kept alive through `keep`); None becomes NULL."""
    valid = {n for n, _ in desc._fields_}
    for k, v in fields.items():
        if k not in valid:
            raise SerError(f"{type(desc).__name__} has no field '{k}'")
        if isinstance(v, torch.Tensor):
            if not v.is_cuda:
                raise SerError(f"field '{k}': tensor is on {v.device}; the fusion head has no CPU path")
            if not v.is_contiguous():
                raise SerError(f"field '{k}': tensor must be contiguous")
            keep.append(v)
            setattr(desc, k, v.data_ptr())
        elif isinstance(v, (list, tuple)):
            arr = (C.c_void_p * len(v))(*[t.data_ptr() for t in v])
            keep.append(arr)
            keep.extend(v)
            setattr(desc, k, C.cast(arr, C.c_void_p))
        elif v is None:
            setattr(desc, k, None)
        else:
            setattr(desc, k, v)
    return desc


def call(name: str, desc, device) -> None:
    lib = load()
    check(getattr(lib, name)(C.byref(desc), stream_ptr(device)), name)


def gemm(a, b, *, a_trans=False, b_trans=False, bias=None, act=ACT_NONE, residual=None, gate=None,
         gate_mode=GATE_NONE, out=None, out_dtype=None, accumulate=False, alpha=1.0, splits=0, rowsum=None):
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
    if rowsum is not None:                       # [M] fp32: row sums of op(a) (bias gradient of a dW GEMM)
        require_cuda(rowsum)
        assert rowsum.dtype == torch.float32 and rowsum.numel() == M and rowsum.is_contiguous()
        d.rowsum = ptr(rowsum)
    check(lib.ser_gemm(C.byref(d), stream_ptr(a.device)), "ser_gemm")
    return out


def set_reserved_sms(n: int) -> None:
    """Leave n SMs out of the persistent GEMM grids (NCCL's channel CTAs in data-parallel steps)."""
    check(load().ser_set_reserved_sms(int(n)), "ser_set_reserved_sms")


def launch_count() -> int:
    return int(load().ser_launch_count())


def prof_enable(on: bool) -> None:
    load().ser_prof_enable(int(on))


def prof_stall(microseconds: float, device) -> None:
    """Keep the current stream of `device` busy for that long, so that the launches enqueued next run back to back."""
    check(load().ser_prof_stall(float(microseconds), stream_ptr(device)), "ser_prof_stall")


def prof_null(n: int, device) -> None:
    """n profiled launches of an empty kernel (family 'prof_null'): the floor of one event interval."""
    check(load().ser_prof_null(int(n), stream_ptr(device)), "ser_prof_null")


def prof_report():
    """{family: dict(launches, ms, flops, bytes)} of everything recorded since the last report (device-syncs)."""
    buf = C.create_string_buffer(1 << 16)
    load().ser_prof_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms, fl, by = line.split()
        out[name] = dict(launches=int(n), ms=float(ms), flops=float(fl), bytes=float(by))
    return out
