"""Per-kernel timings through the C ABI (CUDA events on the library stream).  GPU only."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rfb200  # noqa: E402

ctx = rfb200.Context(0)
lib, h = ctx._lib, ctx.handle
out = {}


def timed(fn, reps=3):
    fn(); ctx.sync()
    best = 1e30
    for _ in range(reps):
        ctx.timer_start(); fn(); best = min(best, ctx.timer_stop())
    return best


rng = np.random.default_rng(0)
# ---- panel: microseconds per column ------------------------------------------------------------
piv = ctx.malloc(64 * 8); info = ctx.malloc(64)
res = {}
for m in (64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
    a = np.asfortranarray(rng.random((m, 64)))
    d = ctx.malloc(a.nbytes); d2 = ctx.malloc(a.nbytes); ctx.h2d(d2, a); ctx.sync()
    def run():
        ctx.d2d(d, d2, a.nbytes)
        ctx._check(lib.rfb_panel_getrf_f64(h, C.c_void_p(d), m, 64, m, C.c_void_p(piv), 0, C.c_void_p(info), 0))
    def copy_only():
        ctx.d2d(d, d2, a.nbytes)
    t = timed(run) - timed(copy_only)
    res[m] = round(t * 1e3 / 64, 3)
    ctx.free(d); ctx.free(d2)
out["panel_us_per_column_n64_T%s" % os.environ.get("RFB_PANEL_THREADS", "auto")] = res
print("panel us/col", res, flush=True)
if os.environ.get("PANEL_ONLY"):
    sys.exit(0)

# ---- gemm -----------------------------------------------------------------------------------------
res = {}
N = 8192
big = ctx.malloc((2 * N) * (2 * N) * 8)
ctx.memset(big, 0, (2 * N) * (2 * N) * 8)
lda = 2 * N
at = lambda r, c: C.c_void_p(big + (r + c * lda) * 8)
for (m, n, k) in [(8192, 8192, 64), (8192, 8192, 128), (8192, 8192, 256), (8192, 8192, 512), (8192, 8192, 1024),
                  (8192, 8192, 2048), (8192, 8192, 8192), (4096, 4096, 4096), (2048, 2048, 2048), (1024, 1024, 1024),
                  (16320, 64, 64), (4096, 64, 64), (8192, 128, 128), (2048, 2048, 64)]:
    for path, name in ((1, "generic"), (2, "tma")):
        ctx.set_default_opts(gemm_path=path)
        f = lambda: ctx._check(lib.rfb_gemm_nn_sub_f64(h, at(k, k) if k < N else at(N, N), at(k, 0) if k < N else at(N, 0),
                                                       at(0, k) if k < N else at(0, N), m, n, k, lda))
        t = timed(f, reps=2 if k >= 2048 else 3)
        res[f"{m}x{n}x{k}:{name}"] = {"ms": round(t, 4), "tflops": round(2.0 * m * n * k / t / 1e9, 2)}
        print(m, n, k, name, res[f"{m}x{n}x{k}:{name}"], flush=True)
ctx.set_default_opts()
out["gemm"] = res

# ---- trsm -----------------------------------------------------------------------------------------
res = {}
for (k, nrhs) in [(64, 8192), (64, 1024), (128, 8192), (256, 8192), (256, 512), (1024, 1024), (2048, 2048), (8192, 8192)]:
    f = lambda: ctx._check(lib.rfb_trsm_llnu_f64(h, at(0, 0), k, at(0, N), nrhs, lda))
    t = timed(f, reps=2)
    res[f"{k}x{nrhs}"] = {"ms": round(t, 4), "tflops": round(1.0 * k * k * nrhs / t / 1e9, 2)}
    print("trsm", k, nrhs, res[f"{k}x{nrhs}"], flush=True)
out["trsm"] = res
out["dmma_peak_tflops"] = ctx.dmma_peak_tflops()
out["copy_gbs"] = ctx.copy_gbs(1 << 30)
json.dump(out, open("gpurun_out/bench_kernels.json", "w"), indent=1)
print(json.dumps(out))
