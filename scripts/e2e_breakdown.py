"""Where the host-to-host time of rfb200.lu_ goes (GPU): kernel-class times of the factorization inside a host-mode
call for the three early-download modes (2 tiles / 1 row bands: eager interchange order; 0: reference order, one
download at the end), next to the device-resident factorization, plus the unprofiled wall time of each."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rfb200  # noqa: E402
from bench import fill_random  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ctx = rfb200.Context(0)
host = ctx.pinned_empty((n, n), np.float64)
fill_random(host)
hwork = ctx.pinned_empty((n, n), np.float64)
ipiv = np.empty(n, dtype=np.int64)
out = {}
dev = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n)
src = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n)
src.upload(host); ctx.sync()
for i in range(2):
    dev.copy_from(src); ctx.timer_start(); dev.lu(); t = ctx.timer_stop()
dev.copy_from(src); ctx.profile_enable(True); dev.lu(); ctx.sync(); p = ctx.profile_read(); ctx.profile_enable(False)
out["device_resident"] = {"ms": t, "classes_ms": {k: round(v["ms"], 3) for k, v in p.items()}, "launches": {k: v["launches"] for k, v in p.items()}}
print("device", out["device_resident"], flush=True)
for mode in (2, 1, 0):
    ctx.set_early_download(mode)
    ts = []
    for i in range(3):
        np.copyto(hwork, host)
        t0 = time.perf_counter(); rfb200.lu_(hwork, ipiv, ctx=ctx); ts.append(1e3 * (time.perf_counter() - t0))
    np.copyto(hwork, host)
    ctx.profile_enable(True); rfb200.lu_(hwork, ipiv, ctx=ctx); p = ctx.profile_read(); ctx.profile_enable(False)
    out[f"host_mode_{mode}"] = {"ms_runs": [round(x, 2) for x in ts], "classes_ms": {k: round(v["ms"], 3) for k, v in p.items()},
                                "launches": {k: v["launches"] for k, v in p.items()}}
    print("mode", mode, out[f"host_mode_{mode}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/e2e_breakdown.json", "w"), indent=1)
