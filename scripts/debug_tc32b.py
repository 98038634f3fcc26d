import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rfb200
ctx = rfb200.Context(0)
rng = np.random.default_rng(0)
def run(m, n, k, mode, r0=0):
    k4 = (k + 3) // 4 * 4
    lda = (r0 + k4 + m + 3) // 4 * 4
    big = np.asfortranarray(rng.random((lda, k4 + n), dtype=np.float32))
    ref = big.astype(np.float64)
    want = ref[r0 + k4:r0 + k4 + m, k4:k4 + n] - ref[r0 + k4:r0 + k4 + m, 0:k] @ ref[r0:r0 + k, k4:k4 + n]
    p = ctx.malloc(big.nbytes); ctx.h2d(p, big)
    at = lambda r, c: C.c_void_p(p + (r + c * lda) * 4)
    ctx.set_default_opts(f32_mode=mode)
    ctx._check(ctx._lib.rfb_gemm_nn_sub_f32(ctx.handle, at(r0 + k4, k4), at(r0 + k4, 0), at(r0, k4), m, n, k, lda))
    out = np.empty_like(big, order="F"); ctx.d2h(out, p); ctx.sync(); ctx.free(p)
    got = out[r0 + k4:r0 + k4 + m, k4:k4 + n].astype(np.float64)
    return float(np.abs(got - want).max()), float(np.abs(want).max())
for shape in [(128,128,32),(128,128,64),(128,128,256),(448,64,64),(64,128,64),(64,64,64),(384,128,128),(256,256,256),(64,256,64),(448,448,64),(200,72,40),(4096,4096,512)]:
    e0, s0 = run(*shape, 0); e1, s1 = run(*shape, 1)
    print(shape, "fp32-simt err %.3e" % e0, "tc32 err %.3e" % e1, "scale %.1f" % s0, "k*eps*scale %.3e" % (shape[2]*1.19e-7*s0), flush=True)
