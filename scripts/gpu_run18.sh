#!/bin/bash
# run 18: TRSV block kernel v2 (pinned prefetch)
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu.py tests/test_gpu_widened.py tests/test_gpu_kernels.py -q -m gpu -x -k "ldiv or solve or trsm or butterfly or nopivot" 2>&1 | tail -4
timeout 600 python - <<'PY'
import ctypes as C, sys, numpy as np
sys.path.insert(0, '.')
import rfb200
ctx = rfb200.Context(0); lib, h = ctx._lib, ctx.handle
for n in (4096, 16384):
    a = np.asfortranarray(np.random.default_rng(1).random((n, n))); a[np.arange(n), np.arange(n)] += n / 4
    d = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n); d.upload(a); d.lu(no_pivot=1); ctx.sync()
    opts = rfb200._make_opts(rfb200._lib.RFB_MEM_DEVICE)
    for nrhs in (1, 4, 8, 9, 64):
        bb = ctx.malloc(n * nrhs * 8); ctx.memset(bb, 0, n * nrhs * 8)
        run = lambda: ctx._check(lib.rfb_solve_f64(h, C.c_void_p(d.ptr), n, n, None, C.c_void_p(bb), nrhs, n, C.byref(opts)))
        run(); ctx.sync()
        best = 1e9
        for _ in range(3):
            ctx.timer_start(); run(); best = min(best, ctx.timer_stop())
        print(n, nrhs, round(best, 3), "ms", flush=True)
        ctx.free(bb)
    d.free()
PY
