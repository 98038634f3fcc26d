"""Timings of the widened rows through the C ABI, CUDA events on the library stream:
unpivoted LU, the butterfly transform / solver, batched small LU.  GPU only.  Writes gpurun_out/bench_widened.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rfb200  # noqa: E402

ctx = rfb200.Context(0)
lib, h = ctx._lib, ctx.handle
out = {}
rng = np.random.default_rng(12)


def timed(fn, reps=3, pre=None):
    best = 1e30
    for i in range(reps + 1):
        if pre:
            pre()
        ctx.timer_start(); fn(); t = ctx.timer_stop()
        if i:
            best = min(best, t)
    return best


# ---- LU with and without pivoting, device resident ---------------------------------------------------
for n in (4096, 16384):
    a = np.asfortranarray(rng.random((n, n)))
    a[np.arange(n), np.arange(n)] += n / 4
    src = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n); src.upload(a); ctx.sync()
    dst = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n)
    for nopiv in (0, 1):
        t = timed(lambda: dst.lu(no_pivot=nopiv), reps=3, pre=lambda: dst.copy_from(src))
        out[f"lu_f64_{n}_{'nopivot' if nopiv else 'pivot'}"] = {"ms": round(t, 3), "gflops": round(2 * n ** 3 / 3 / t / 1e6, 1)}
        print(n, "nopiv" if nopiv else "pivot", out[f"lu_f64_{n}_{'nopivot' if nopiv else 'pivot'}"], flush=True)
    ctx.profile_enable(True)
    dst.copy_from(src); dst.lu(no_pivot=1); ctx.sync()
    prof = ctx.profile_read(); ctx.profile_enable(False)
    out[f"lu_f64_{n}_nopivot_classes"] = {k: {"ms": round(v["ms"], 3), "launches": v["launches"]} for k, v in prof.items()}
    print(out[f"lu_f64_{n}_nopivot_classes"], flush=True)
    # ---- butterfly transform: bytes = 2 * 8 * n^2 -----------------------------------------------------
    uv = rfb200.butterfly_generate_random(n)
    duv = ctx.malloc(uv.nbytes); ctx.h2d(duv, uv); ctx.sync()
    t = timed(lambda: ctx._check(lib.rfb_butterfly_mul_f64(h, C.c_void_p(dst.ptr), n, n, C.c_void_p(duv))), reps=5)
    out[f"butterfly_mul_f64_{n}"] = {"ms": round(t, 4), "GBps": round(16.0 * n * n / t / 1e6, 1)}
    print("butterfly_mul", n, out[f"butterfly_mul_f64_{n}"], flush=True)
    # ---- whole butterfly solve, device resident (transform + NoPivot LU + 2 trsm + vec ops) ----------
    b = ctx.malloc(n * 8 * 1)
    # B shares lda with A in device mode: use a separate n x 1 buffer with ldb == lda == n
    info = ctx.malloc(64)
    opts = rfb200._make_opts(rfb200._lib.RFB_MEM_DEVICE)
    def solve():
        ctx._check(lib.rfb_butterfly_solve_f64(h, C.c_void_p(dst.ptr), n, n, C.c_void_p(b), 1, n, C.c_void_p(duv), C.c_void_p(info), C.byref(opts)))
    def pre():
        dst.copy_from(src); ctx.memset(b, 0, n * 8)
    t = timed(solve, reps=3, pre=pre)
    out[f"butterfly_solve_f64_{n}_device"] = {"ms": round(t, 3), "gflops_lu_equiv": round(2 * n ** 3 / 3 / t / 1e6, 1)}
    print("butterfly_solve", n, out[f"butterfly_solve_f64_{n}_device"], flush=True)
    # ---- ldiv!(F, B) on the factors just produced: vector and block right-hand sides ------------------
    for nrhs in (1, 4, 64, 1024):
        bb = ctx.malloc(n * nrhs * 8); ctx.memset(bb, 0, n * nrhs * 8)
        run = lambda: ctx._check(lib.rfb_solve_f64(h, C.c_void_p(dst.ptr), n, n, None, C.c_void_p(bb), nrhs, n, C.byref(opts)))
        t = timed(run, reps=3)
        out[f"ldiv_notipiv_f64_{n}_nrhs{nrhs}"] = {"ms": round(t, 3), "GBps_factors": round(8.0 * n * n / t / 1e6, 1),
                                                   "gflops": round(2.0 * n * n * nrhs / t / 1e6, 1)}
        print("ldiv", n, nrhs, out[f"ldiv_notipiv_f64_{n}_nrhs{nrhs}"], flush=True)
        ctx.free(bb)
    ctx.free(duv); ctx.free(b); ctx.free(info); src.free(); dst.free()

# ---- butterfly solve end to end from host (n = 8192) vs pivoted lu + solve --------------------------------
n = 8192
a = np.asfortranarray(rng.random((n, n))); bvec = rng.random(n)
ws = rfb200.ButterflyWorkspace(a, bvec)
rfb200.butterfly_solve_(ws, ctx=ctx)
t0 = time.perf_counter(); x = rfb200.butterfly_solve_(ws, ctx=ctx); t1 = time.perf_counter()
res = float(np.linalg.norm(a @ x - bvec) / np.linalg.norm(bvec))
t2 = time.perf_counter(); F = rfb200.lu(a, ctx=ctx); y = F.solve(bvec, ctx=ctx); t3 = time.perf_counter()
res2 = float(np.linalg.norm(a @ y - bvec) / np.linalg.norm(bvec))
out["solve_8192_host_e2e"] = {"butterfly_ms": round((t1 - t0) * 1e3, 1), "butterfly_rel_residual": res,
                              "pivoted_lu_solve_ms": round((t3 - t2) * 1e3, 1), "pivoted_rel_residual": res2}
print(out["solve_8192_host_e2e"], flush=True)

# ---- batched small LU, device resident ---------------------------------------------------------------------
for (batch, m, dt) in [(16384, 16, np.float64), (16384, 32, np.float64), (16384, 64, np.float64), (65536, 32, np.float64),
                       (16384, 32, np.float32), (16384, 64, np.float32), (4096, 128, np.float64)]:
    nn = min(m, 64)
    it = np.dtype(dt).itemsize
    a = rng.random((batch, nn, m)).astype(dt)           # each [b] is (n, m) C-order == (m, n) column-major
    d0 = ctx.malloc(a.nbytes); d1 = ctx.malloc(a.nbytes); ctx.h2d(d0, a); ctx.sync()
    piv = ctx.malloc(batch * nn * 8); info = ctx.malloc(batch * 8)
    opts = rfb200._make_opts(rfb200._lib.RFB_MEM_DEVICE)
    f = lib.rfb_lu_batched_f64 if dt == np.float64 else lib.rfb_lu_batched_f32
    run = lambda: ctx._check(f(h, C.c_void_p(d1), m, nn, m, m * nn, batch, C.c_void_p(piv), C.c_void_p(info), C.byref(opts)))
    t = timed(run, reps=3, pre=lambda: ctx.d2d(d1, d0, a.nbytes))
    flops = (m * nn * nn - nn ** 3 / 3.0) * batch
    key = f"batched_{np.dtype(dt).name}_{batch}x{m}x{nn}"
    out[key] = {"ms": round(t, 4), "matrices_per_s": round(batch / t * 1e3), "gflops": round(flops / t / 1e6, 1),
                "GBps_alg": round(2.0 * a.nbytes / t / 1e6, 1)}
    print(key, out[key], flush=True)
    for p in (d0, d1, piv, info):
        ctx.free(p)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_widened.json", "w"), indent=1)
print(json.dumps(out))
