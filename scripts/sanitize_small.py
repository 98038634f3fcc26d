"""Small cases through every kernel family, meant to run under compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
Checks results against the CPU oracle as well, so a sanitizer-clean run is also a parity run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rfb200  # noqa: E402
from oracle import rf_oracle as O  # noqa: E402

ctx = rfb200.Context(0)
rng = np.random.default_rng(3)
# pivoted LU: single CTA, 256-thread CTA, multi-CTA exchange, recursion with laswp lists / trsm / gemm (generic + TMA)
for dtype in (np.float64, np.float32):
    for (m, n) in ((7, 7), (100, 100), (200, 130), (300, 302), (700, 640), (1100, 1100)):
        a = np.asfortranarray(rng.random((m, n)).astype(dtype))
        F = rfb200.lu(a, ctx=ctx)
        _, p, info = O.lu_c(a.copy(order="F"))
        assert F.info == info == 0 and (dtype == np.float32 or np.array_equal(F.ipiv, p)), (m, n, dtype)
# cluster exchange, if enabled through RFB_PANEL_CLUSTER=1
a = np.asfortranarray(rng.random((900, 64)))
F = rfb200.lu(a, ctx=ctx)
assert np.array_equal(F.ipiv, O.panel_c(a.copy(order="F"))[1])
# unpivoted LU + NotIPIV solve (vector -> skinny GEMM, matrix -> tensor GEMM)
for (m, n) in ((64, 64), (300, 300), (520, 400), (1030, 1030)):
    a = np.asfortranarray(rng.random((m, n))); a[np.arange(min(m, n)), np.arange(min(m, n))] += 10
    F = rfb200.lu(a, False, ctx=ctx)
    f, _, info = O.lu_nopiv_c(a.copy(order="F"))
    assert F.info == info == 0 and np.allclose(F.factors, f, atol=1e-10)
    if m == n:
        b = rng.random(n)
        x = rfb200.ldiv_(F, b.copy(), ctx=ctx)
        assert np.linalg.norm(a @ x - b) < 1e-9
        bb = np.asfortranarray(rng.random((n, 40)))
        xx = rfb200.ldiv_(F, bb.copy(order="F"), ctx=ctx)
        assert np.linalg.norm(a @ xx - bb) < 1e-8
# butterfly
for n in (5, 64, 203, 516):
    a = np.asfortranarray(rng.random((n, n))); a[np.arange(n), np.arange(n)] += 10
    b = rng.random(n)
    x = rfb200.butterfly_solve_(rfb200.ButterflyWorkspace(a, b), ctx=ctx)
    assert np.linalg.norm(a @ x - b) < 1e-9 * np.linalg.norm(b) * n
# batched
for (batch, m, n) in ((50, 8, 8), (33, 40, 24), (9, 100, 64), (7, 10, 12), (3, 150, 150)):
    a3 = rng.random((batch, m, n))
    Fs = rfb200.lu_batched(a3, ctx=ctx)
    for b in range(batch):
        if n <= 64 and m <= 128:
            wf, wp, _ = O.panel_c(np.asfortranarray(a3[b]))
            assert np.array_equal(Fs[b].ipiv, wp) and np.array_equal(Fs[b].factors, wf)
# pinned host matrix: eager interchanges + early row downloads
n = 2300
a = np.asfortranarray(rng.random((n, n)))
ref = rfb200.lu(a, ctx=ctx)
pin = ctx.pinned_empty((n, n), np.float64); np.copyto(pin, a)
F = rfb200.lu_(pin, None, ctx=ctx)
assert np.array_equal(F.ipiv, ref.ipiv) and np.array_equal(np.asarray(F.factors), ref.factors)
ctx.close()
print("sanitize_small: all cases OK")
