import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rfb200
np.set_printoptions(linewidth=250)
ctx = rfb200.Context(0)
ctx.set_default_opts(f32_mode=1)

def run(Am, Bm):
    m, k = Am.shape; k2, n = Bm.shape
    k4 = (k + 3) // 4 * 4
    lda = k4 + m
    big = np.zeros((lda, k4 + n), dtype=np.float32, order="F")
    big[k4:k4 + m, 0:k] = Am
    big[0:k, k4:k4 + n] = Bm
    p = ctx.malloc(big.nbytes); ctx.h2d(p, big)
    at = lambda r, c: C.c_void_p(p + (r + c * lda) * 4)
    ctx._check(ctx._lib.rfb_gemm_nn_sub_f32(ctx.handle, at(k4, k4), at(k4, 0), at(0, k4), m, n, k, lda))
    out = np.empty_like(big, order="F"); ctx.d2h(out, p); ctx.sync(); ctx.free(p)
    return -out[k4:k4 + m, k4:k4 + n]

M = N = 128
for K in (8, 32):
    print("=== K =", K)
    # Exp 1: decode how A is read
    A = (np.arange(M)[:, None] * 8 + (np.arange(K)[None, :] % 8)).astype(np.float32)  # value = i*8 + kk%8  (< 1024)
    B = np.zeros((K, N), dtype=np.float32); B[np.arange(K), np.arange(K)] = 1
    D = run(A, B)
    want = A @ B
    print("exp1 (A decode) max err", np.abs(D - want).max())
    if np.abs(D - want).max() > 0:
        dec_i = (D[:, :K] // 8).astype(int); dec_k = (D[:, :K] % 8).astype(int)
        print("D[0:12, 0:8]:\n", D[0:12, 0:8]); print("rows 32..36:\n", D[32:36, 0:8])
        print("nonzero columns beyond K:", np.nonzero(np.abs(D[:, K:]).sum(axis=0))[0][:20])
    # Exp 2: decode how B is read
    A = np.zeros((M, K), dtype=np.float32); A[np.arange(K), np.arange(K)] = 1
    B = ((np.arange(K)[:, None] % 8) * 128 + np.arange(N)[None, :]).astype(np.float32)  # value = (kk%8)*128 + j (< 1024)
    D = run(A, B)
    want = A @ B
    print("exp2 (B decode) max err", np.abs(D - want).max())
    if np.abs(D - want).max() > 0:
        print("D[0:8, 0:12]:\n", D[0:8, 0:12]); print("D[0:8, 64:72]:\n", D[0:8, 64:72])
        print("nonzero rows beyond K:", np.nonzero(np.abs(D[K:, :]).sum(axis=1))[0][:20] + K)
    # Exp 3: all ones
    D = run(np.ones((M, K), dtype=np.float32), np.ones((K, N), dtype=np.float32))
    print("exp3 (ones) unique values:", np.unique(D)[:10])
