"""Does a copy-engine upload running beside the factorization slow its kernels down?  (GPU)  Times the device-resident
16384^2 LU and a run of 16384 x 64 panels alone and with a 2.1 GB pinned host -> device copy streaming on a second
context's stream; the host-mode pipeline uploads the matrix during the first ~39 ms of the factorization."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rfb200  # noqa: E402
from bench import fill_random  # noqa: E402

n = 16384
ctx, ctx2 = rfb200.Context(0), rfb200.Context(0)
host = ctx.pinned_empty((n, n), np.float64)
fill_random(host)
src = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n); src.upload(host); ctx.sync()
dev = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n)
sink = ctx2.malloc(n * n * 8)
out = {}


def lu_ms(with_copy):
    best = 1e9
    for _ in range(3):
        dev.copy_from(src); ctx.sync(); ctx2.sync()
        if with_copy:
            ctx2.h2d(sink, host)                     # async on ctx2's stream: ~39 ms of copy-engine traffic
        ctx.timer_start(); dev.lu(); best = min(best, ctx.timer_stop())
        ctx2.sync()
    return best


out["lu_alone_ms"] = lu_ms(False)
out["lu_with_h2d_ms"] = lu_ms(True)
print(out, flush=True)

# 16384 x 64 panels (the latency-bound, L2-polling kernel), 100 in a row, restored by a d2d copy of the panel before each
piv = ctx.malloc(64 * 8); info = ctx.malloc(64)
lib, h = ctx._lib, ctx.handle


def panels_ms(with_copy, reps=100):
    best = 1e9
    for _ in range(2):
        ctx.sync(); ctx2.sync()
        if with_copy:
            ctx2.h2d(sink, host)
        ctx.timer_start()
        for _ in range(reps):
            ctx.d2d(dev.ptr, src.ptr, n * 64 * 8)
            ctx._check(lib.rfb_panel_getrf_f64(h, C.c_void_p(dev.ptr), n, 64, n, C.c_void_p(piv), 0, C.c_void_p(info), 0))
        best = min(best, ctx.timer_stop())
        ctx2.sync()
    return best / reps


out["panel_16384x64_alone_us"] = 1e3 * panels_ms(False)
out["panel_16384x64_with_h2d_us"] = 1e3 * panels_ms(True)
print(out, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dma_interference.json", "w"), indent=1)
