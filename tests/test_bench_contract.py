"""CPU checks of bench.py's contract that need no GPU: the reference arm (times the oracle port on the host cores)
prints exactly one JSON line with the keys the driver reads, and the helper formulas are the documented ones."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--n", "2048"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("LU GFLOP/s (2n^3/3)") and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None                      # BASELINE.md publishes no number for this metric


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_flop_and_residual_helpers():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.lu_flops(16384) == 2.0 * 16384 ** 3 / 3.0
    rng = np.random.default_rng(0)
    a = np.asfortranarray(rng.random((200, 200)))
    from oracle import rf_oracle as O
    f, ipiv, info = O.lu_c(a.copy(order="F"))
    est = bench.hutchinson_residual(a, f, ipiv, nvec=16)
    exact = O.residual_fro_rel(a, f, ipiv)
    # the probe estimator is unbiased for ||.||_F^2 but carries its own O(n eps) rounding (it forms P A x - L (U x) in
    # float64), so near machine precision it over-reports: it must never UNDER-report, and stays far inside the bound
    assert info == 0 and 0.3 * exact < est < 20 * 200 * np.finfo(np.float64).eps
    bad = f.copy(order="F"); bad[150, 100] += 1e-3              # and it sees a real defect
    assert bench.hutchinson_residual(a, bad, ipiv, nvec=16) > 1e-6
