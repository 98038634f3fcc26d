#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tc32 tests first"
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tcgen05" --timeout 300 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_lu.py -m gpu -q -k "tensor_core" --timeout 300 2>&1 | tail -15
echo "== full pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== f32 timing"
timeout 300 python - <<'PY'
import numpy as np, rfb200, time
ctx = rfb200.Context(0)
n = 8192
a = np.asfortranarray(np.random.default_rng(12).random((n, n), dtype=np.float32))
for mode in (0, 1):
    d = rfb200.DeviceMatrix(ctx, n, n, np.float32, lda=n); p = rfb200.DeviceMatrix(ctx, n, n, np.float32, lda=n)
    p.upload(a); ctx.sync()
    for it in range(3):
        d.copy_from(p); ctx.timer_start(); d.lu(f32_mode=mode); ms = ctx.timer_stop()
    print("f32 LU 8192 mode", mode, ms, "ms", 2*n**3/3/ms/1e9, "TFLOP/s")
    d.copy_from(p); ctx.profile_enable(True); d.lu(f32_mode=mode); print(ctx.profile_read()); ctx.profile_enable(False)
    d.free(); p.free()
PY
