import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rfb200
from oracle import rf_oracle as O
for (m, n) in [(128, 64), (129, 64), (256, 64), (256, 16), (1024, 64)]:
    ctx = rfb200.Context(0)
    a0 = np.asfortranarray(np.random.default_rng([1, m, n]).random((m, n)))
    want_f, want_p, want_info = O.panel_c(a0.copy(order="F"))
    d = ctx.malloc(a0.nbytes); piv = ctx.malloc(n * 8); info = ctx.malloc(64)
    ctx.h2d(d, a0); ctx.memset(info, 0, 64)
    ctx._check(ctx._lib.rfb_panel_getrf_f64(ctx.handle, C.c_void_p(d), m, n, m, C.c_void_p(piv), 0, C.c_void_p(info), 0))
    got = np.empty_like(a0, order="F"); gp = np.empty(n, dtype=np.int64)
    try:
        ctx.d2h(got, d); ctx.d2h(gp, piv); ctx.sync()
        print(m, n, "OK piv equal", np.array_equal(gp, want_p), "factors equal", np.array_equal(got, want_f), flush=True)
    except Exception as e:
        ctx2 = None
        print(m, n, "FAILED", e, "pivots so far", gp[:8], want_p[:8], flush=True)
    ctx.close()
