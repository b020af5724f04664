"""Prints which TMEM lane holds accumulator row r for tcgen05.mma cta_group::1 with M = 64 (and M = 128 as a
control).  Run under gpurun: python tools/probe_tmem_layout.py"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from mmser_b200 import _lib as L  # noqa: E402


from tools.probe_build import build_probe  # noqa: E402

lib = build_probe()
lib.ser_debug_probe_tmem_layout.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
for m in (128, 64):
    out = torch.zeros(128, 64, device="cuda")
    rc = lib.ser_debug_probe_tmem_layout(out.data_ptr(), m, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    o = out.cpu()
    print(f"== M={m} rc={rc}")
    rows = {}
    for lane in range(128):
        v0, v1 = float(o[lane, 0]), float(o[lane, 1])
        if v0 > 0:
            rows[lane] = (round(v0) - 1, v1 / v0)          # D[r][0] = r + 1 ; D[r][1] / D[r][0] = 2
    print("lane -> row (col1/col0):", {k: f"{r} ({q:.1f})" for k, (r, q) in rows.items()})
    print("untouched lanes:", [l for l in range(128) if float(o[l, 0]) == -7.0])
    print("row 5 columns 0..7:", [float(x) for x in o[[k for k, (r, _) in rows.items() if r == 5][0], :8]] if rows else None)

# ---- MMA timing: cycles per tcgen05.mma (K = 16) as a function of the tile shape
lib.ser_debug_mma_time.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
for m in (128, 64):
    for n in (32, 64, 128, 256):
        res = []
        for reps in (32, 128):
            buf = torch.zeros(2, device="cuda", dtype=torch.int64)
            for _ in range(2):
                lib.ser_debug_mma_time(buf.data_ptr(), m, n, reps, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
            res.append((reps, int(buf[0]), int(buf[1])))
        (r0, i0, c0), (r1, i1, c1) = res
        print(f"MMA M={m:3d} N={n:3d}: {(c1 - c0) / (r1 - r0):6.1f} cycles per MMA (issue {(i1 - i0) / (r1 - r0):5.1f}); "
              f"{r0} MMAs complete in {c0} cycles")
