"""Tensor-memory load/store throughput (tools/probe.cu ldtm_time_kernel).  Run under gpurun."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from tools.probe_build import build_probe  # noqa: E402

lib = build_probe()
lib.ser_debug_ldtm_time.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
names = {0: "ld 32x32b.x32", 1: "st 32x32b.x32", 3: "ld 16x256b.x8"}
for mode in (0, 3, 1):
    for nw in (1, 2, 4, 8, 16):
        res = []
        for reps in (64, 256):
            buf = torch.zeros(64, device="cuda", dtype=torch.int64)
            for _ in range(2):
                lib.ser_debug_ldtm_time(buf.data_ptr(), nw, reps, mode, 1, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
            res.append((reps, int(buf[:nw].max())))
        (r0, c0), (r1, c1) = res
        per = (c1 - c0) / (r1 - r0)
        print(f"{names[mode]:16s} warps={nw:2d}: {per:7.1f} cycles per instruction per warp  -> {nw * 4096 / per:7.1f} B/clk per SM", flush=True)
