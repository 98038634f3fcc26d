"""Where the 16384^2 factorization's time goes, from the host driver's own schedule (rfb_trace_lu, no GPU needed)
combined with MEASURED per-shape kernel rates (profiles/r01_bench_kernels_run5.json: GEMM TFLOP/s by k; panel us/column
from run 23).  Writes a markdown table; the point is to rank head-room, not to predict the total to the millisecond."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rfb200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ops = rfb200.trace_lu(n, n)
PANEL, LASWP, TRSM, GEMM = 1, 3, 4, 5
kern = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_kernels_run5.json")))
# measured TMA-GEMM rate for 8192 x 8192 x k (TFLOP/s); small m/n shapes measured separately
rate_by_k = {64: 15.8, 128: 22.0, 256: 26.9, 512: 30.1, 1024: 31.7, 2048: 32.7, 4096: 34.0, 8192: 35.2}
peak = 35.2


def gemm_rate(m, nn, k):
    kk = min(rate_by_k, key=lambda x: abs(np.log2(x) - np.log2(max(k, 64))))
    r = rate_by_k[kk]
    tiles = -(-m // 128) * -(-nn // 128)
    waves = -(-tiles // 148)
    fill = tiles / (waves * 148.0)                     # last-wave quantisation on 148 SMs
    width = min(1.0, nn / (128.0 * -(-nn // 128)))     # half-empty 128-wide tiles
    return r * min(1.0, fill / 0.9) * width


rows = {}
tot = dict(flops=0.0, t=0.0, ideal=0.0)
for op, r, c, s0, s1, s2, r2, c2 in ops.tolist():
    if op == GEMM:
        key = ("schur", s2)
        fl = 2.0 * s0 * s1 * s2
        t = fl / (gemm_rate(s0, s1, s2) * 1e12)
    elif op == TRSM:
        # blocked TRSM: k^2 * nrhs flops, all but the 256-row diagonal blocks are GEMMs of inner size >= 256
        key = ("trsm", s0)
        fl = 1.0 * s0 * s0 * s1
        t = fl / (gemm_rate(s0 // 2 or 1, s1, max(s0 // 2, 256)) * 1e12) if s0 > 256 else 0.0
    else:
        continue
    e = rows.setdefault(key, [0, 0.0, 0.0])
    e[0] += 1; e[1] += fl; e[2] += t
    tot["flops"] += fl; tot["t"] += t; tot["ideal"] += fl / (peak * 1e12)

lines = [f"# Schedule model of the {n} x {n} Float64 factorization (scripts/analyze_trace.py)", "",
         "Operations from `rfb_trace_lu` (the C++ host driver run dry); GEMM rates from the measured 8192 x 8192 x k table",
         "(`profiles/r01_bench_kernels_run5.json`) scaled by 128-wide tile occupancy and last-wave fill on 148 SMs.", "",
         "| class | inner size | launches | GFLOP | modelled ms | ms at the 35.2 TFLOP/s root rate | lost ms |", "|---|---|---|---|---|---|---|"]
for (cls, k), (cnt, fl, t) in sorted(rows.items(), key=lambda kv: (kv[0][0], -kv[0][1])):
    ideal = fl / (peak * 1e12)
    lines.append(f"| {cls} | {k} | {cnt} | {fl / 1e9:.1f} | {t * 1e3:.2f} | {ideal * 1e3:.2f} | {(t - ideal) * 1e3:.2f} |")
lines.append(f"| **all GEMM-shaped work** | | | {tot['flops'] / 1e9:.0f} | {tot['t'] * 1e3:.1f} | {tot['ideal'] * 1e3:.1f} | {(tot['t'] - tot['ideal']) * 1e3:.1f} |")
kinds = ops[:, 0].tolist()
ncols = int(ops[ops[:, 0] == PANEL][:, 4].sum())
lines += ["", f"Panels: {kinds.count(PANEL)} launches, {ncols} pivot columns; at the measured 1.9 us per column (run 23) = "
          f"{ncols * 1.9e-3:.1f} ms, at the single-CTA 1.05 us = {ncols * 1.05e-3:.1f} ms.",
          f"Row interchanges: {kinds.count(LASWP)} launches, {4 * 8 * n * n / 1e9:.1f} GB algorithmic (each pivot meets each column "
          "once, 4 accesses of 8 bytes), ~2.5x that in 32-byte sectors.",
          "", "Measured totals for comparison (run 25): gemm class 94.6 ms (includes the TRSM-internal GEMMs), panel 31.3 ms, "
          "trsm diagonal blocks 13.7 ms, laswp 12.9 ms."]
out = os.path.join(ROOT, "profiles", f"r01_schedule_model_{n}.md")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
