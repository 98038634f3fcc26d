"""Small targets for ncu captures: `python scripts/ncu_target.py lu N` or `gemm M N K path`."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rfb200  # noqa: E402

ctx = rfb200.Context(0)
what = sys.argv[1]
if what == "lu":
    n = int(sys.argv[2])
    a = np.asfortranarray(np.random.default_rng(12).random((n, n)))
    d = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n)
    d.upload(a)
    d.lu()
    ctx.sync()
elif what == "gemm32":
    m, n, k = (int(x) for x in sys.argv[2:5])
    lda = m + k
    big = ctx.malloc(lda * (n + k) * 4)
    ctx.memset(big, 0, lda * (n + k) * 4)
    at = lambda r, c: C.c_void_p(big + (r + c * lda) * 4)
    ctx.set_default_opts(f32_mode=1)
    for _ in range(2):
        ctx._check(ctx._lib.rfb_gemm_nn_sub_f32(ctx.handle, at(k, k), at(k, 0), at(0, k), m, n, k, lda))
    ctx.sync()
elif what == "gemm":
    m, n, k, path = (int(x) for x in sys.argv[2:6])
    lda = m + k
    big = ctx.malloc(lda * (n + k) * 8)
    ctx.memset(big, 0, lda * (n + k) * 8)
    at = lambda r, c: C.c_void_p(big + (r + c * lda) * 8)
    ctx.set_default_opts(gemm_path=path)
    for _ in range(2):
        ctx._check(ctx._lib.rfb_gemm_nn_sub_f64(ctx.handle, at(k, k), at(k, 0), at(0, k), m, n, k, lda))
    ctx.sync()
elif what == "nopiv":
    n = int(sys.argv[2])
    a = np.asfortranarray(np.random.default_rng(12).random((n, n)))
    a[np.arange(n), np.arange(n)] += n / 4
    d = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n)
    d.upload(a)
    uv = rfb200.butterfly_generate_random(n)
    duv = ctx.malloc(uv.nbytes); ctx.h2d(duv, uv)
    ctx._check(ctx._lib.rfb_butterfly_mul_f64(ctx.handle, C.c_void_p(d.ptr), n, n, C.c_void_p(duv)))
    d.lu(no_pivot=1)
    ctx.sync()
elif what == "batched":
    batch, m = int(sys.argv[2]), int(sys.argv[3])
    a = np.random.default_rng(1).random((batch, m, m))
    F = rfb200.lu_batched(a, ctx=ctx)
elif what == "panel":
    m, n = int(sys.argv[2]), int(sys.argv[3])
    a = np.asfortranarray(np.random.default_rng(12).random((m, n)))
    d = ctx.malloc(a.nbytes); ctx.h2d(d, a)
    piv = ctx.malloc(n * 8); info = ctx.malloc(64); ctx.memset(info, 0, 64)
    for _ in range(2):
        ctx.h2d(d, a)
        ctx._check(ctx._lib.rfb_panel_getrf_f64(ctx.handle, C.c_void_p(d), m, n, m, C.c_void_p(piv), 0, C.c_void_p(info), 0))
    ctx.sync()
