"""GPU parity tests, kernel level: each sm_100a kernel behind the C ABI against the matching loop of
the CPU oracle on the same seeded inputs (integer / index outputs bit-exact, floating point within
the reference's own n*eps-style bound)."""
import ctypes as C

import numpy as np
import pytest

import rfb200
from oracle import rf_oracle as O
from util import rand_matrix, ref_bound

pytestmark = pytest.mark.gpu

F = {np.float64: "f64", np.float32: "f32"}


class Dev:
    """A column-major host matrix mirrored in device memory; sub-block pointers by (row, col)."""

    def __init__(self, ctx, a):
        self.ctx, self.host = ctx, np.asfortranarray(a)
        self.lda, self.it = a.shape[0], a.itemsize
        self.ptr = ctx.malloc(max(a.nbytes, 16))
        ctx.h2d(self.ptr, self.host)

    def at(self, r, c):
        return C.c_void_p(self.ptr + (r + c * self.lda) * self.it)

    def get(self):
        out = np.empty_like(self.host, order="F")
        self.ctx.d2h(out, self.ptr)
        self.ctx.sync()
        return out

    def free(self):
        self.ctx.free(self.ptr)


def dev_i64(ctx, arr):
    arr = np.ascontiguousarray(arr, dtype=np.int64)
    p = ctx.malloc(max(arr.nbytes, 64))
    ctx.h2d(p, arr)
    return p


def get_i64(ctx, p, n):
    out = np.empty(n, dtype=np.int64)
    ctx.d2h(out, p)
    ctx.sync()
    return out


def fn(ctx, name, dtype):
    return getattr(ctx._lib, f"{name}_{F[dtype]}")


# ---- K1 panel ----------------------------------------------------------------------------------
PANEL_SHAPES = [(1, 1), (5, 3), (8, 8), (64, 64), (100, 16), (128, 64), (129, 64), (300, 40), (1000, 64),
                (4096, 64), (5000, 33), (16384, 64), (20000, 64), (37000, 32)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", PANEL_SHAPES)
def test_panel_matches_oracle(ctx, dtype, shape):
    m, n = shape
    rng = np.random.default_rng([1, m, n])
    a0 = rand_matrix(rng, m, n, dtype)
    want_f, want_p, want_info = O.panel_c(a0.copy(order="F"))
    d = Dev(ctx, a0)
    piv = dev_i64(ctx, np.zeros(n))
    info = dev_i64(ctx, np.zeros(8))
    ctx._check(fn(ctx, "rfb_panel_getrf", dtype)(ctx.handle, d.at(0, 0), m, n, d.lda, C.c_void_p(piv), 0,
                                                C.c_void_p(info), 0))
    got_f, got_p, got_info = d.get(), get_i64(ctx, piv, n), int(get_i64(ctx, info, 1)[0])
    assert np.array_equal(got_p, want_p)                       # bit-exact pivots
    assert got_info == want_info
    assert np.allclose(got_f, want_f, rtol=0, atol=ref_bound(m, dtype))
    d.free(); ctx.free(piv); ctx.free(info)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_panel_ties_zero_column_and_offsets(ctx, dtype):
    rng = np.random.default_rng(5)
    m, n = 700, 48
    a0 = np.asfortranarray(rng.integers(-3, 4, size=(m, n)).astype(dtype))   # many exact ties
    a0[:, 20] = 0                                                            # exactly-zero pivot -> info
    want_f, want_p, want_info = O.panel_c(a0.copy(order="F"))
    # factor it as a sub-block at (row 10, col 6) of a larger allocation; pivots shifted by ipiv_add
    big = np.asfortranarray(rng.random((m + 30, n + 20)).astype(dtype))
    big[10:10 + m, 6:6 + n] = a0
    d = Dev(ctx, big)
    piv = dev_i64(ctx, np.zeros(n)); info = dev_i64(ctx, np.zeros(8))
    ctx._check(fn(ctx, "rfb_panel_getrf", dtype)(ctx.handle, d.at(10, 6), m, n, d.lda, C.c_void_p(piv), 10,
                                                C.c_void_p(info), 6))
    got = d.get()
    assert np.array_equal(get_i64(ctx, piv, n), want_p + 10)
    assert int(get_i64(ctx, info, 1)[0]) == (want_info + 6 if want_info else 0) and want_info > 0
    # same operation order, same FMA, same reciprocal: the panel kernel reproduces the oracle's bits
    assert np.array_equal(got[10:10 + m, 6:6 + n], want_f)
    outside = np.ones_like(big, dtype=bool); outside[10:10 + m, 6:6 + n] = False
    assert np.array_equal(got[outside], big[outside])                        # nothing else touched
    d.free(); ctx.free(piv); ctx.free(info)


def test_panel_nan_and_zero_matrix(ctx):
    a0 = np.asfortranarray(np.random.default_rng(0).random((300, 16)))
    a0[5, 0] = np.nan
    d = Dev(ctx, a0)
    piv = dev_i64(ctx, np.zeros(32)); info = dev_i64(ctx, np.zeros(8))
    ctx._check(ctx._lib.rfb_panel_getrf_f64(ctx.handle, d.at(0, 0), 300, 16, 300, C.c_void_p(piv), 0, C.c_void_p(info), 0))
    _, want_p, _ = O.panel_c(a0.copy(order="F"))
    assert np.array_equal(get_i64(ctx, piv, 16), want_p)      # NaN is never selected (src/lu.jl:301)
    z = Dev(ctx, np.zeros((200, 32), order="F"))
    ctx._check(ctx._lib.rfb_memset(ctx.handle, C.c_void_p(info), 0, 64))
    ctx._check(ctx._lib.rfb_panel_getrf_f64(ctx.handle, z.at(0, 0), 200, 32, 200, C.c_void_p(piv), 0, C.c_void_p(info), 0))
    assert np.array_equal(get_i64(ctx, piv, 32)[:32], np.arange(1, 33))
    assert int(get_i64(ctx, info, 1)[0]) == 1
    d.free(); z.free(); ctx.free(piv); ctx.free(info)


# ---- K2 laswp ----------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(40, 7, 6), (300, 129, 150), (2100, 1000, 1500), (5000, 300, 64)])
def test_laswp_matches_oracle(ctx, dtype, shape):
    m, ncols, npiv = shape
    rng = np.random.default_rng([2, m, ncols])
    a0 = rand_matrix(rng, m, ncols, dtype)
    piv = np.array([rng.integers(i + 1, m + 1) for i in range(npiv)], dtype=np.int64)
    piv[::7] = np.arange(1, npiv + 1)[::7]                    # some i' == i
    if npiv > 3:
        piv[1] = piv[0]                                       # repeated target: sequential semantics matter
    want = O.laswp_c(a0.copy(order="F"), piv)
    d = Dev(ctx, a0); p = dev_i64(ctx, piv + 11)
    ctx._check(fn(ctx, "rfb_laswp", dtype)(ctx.handle, d.at(0, 0), ncols, d.lda, C.c_void_p(p), npiv, 11))
    assert np.array_equal(d.get(), want)                      # pure data movement: bit-exact
    d.free(); ctx.free(p)


# ---- K3 trsm -----------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (64, 64), (65, 130), (128, 64), (150, 140), (200, 700), (256, 256), (257, 100),
                                   (512, 333), (1000, 1000), (2048, 1500)])
def test_trsm_matches_oracle(ctx, dtype, shape):
    k, nrhs = shape
    rng = np.random.default_rng([3, k, nrhs])
    big = rand_matrix(rng, k + 5, k + nrhs + 3, dtype)
    big[:, : k + 1] *= dtype(2.0 / k) if k > 16 else dtype(1)           # keep the solve well scaled
    want = O.trsm_c(big.copy(order="F"), (2, 1), k, (2, k + 2), nrhs)
    d = Dev(ctx, big)
    ctx._check(fn(ctx, "rfb_trsm_llnu", dtype)(ctx.handle, d.at(2, 1), k, d.at(2, k + 2), nrhs, d.lda))
    got = d.get()
    scale = max(1.0, float(np.abs(want).max()))
    assert np.allclose(got, want, rtol=0, atol=ref_bound(k, dtype) * scale)
    d.free()


# ---- K4 gemm -----------------------------------------------------------------------------------
GEMM_SHAPES = [(1, 1, 1), (8, 8, 4), (37, 29, 13), (128, 128, 16), (130, 127, 64), (256, 384, 100), (1000, 513, 300),
               (2048, 2048, 64), (1536, 1024, 1024)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", GEMM_SHAPES)
@pytest.mark.parametrize("offs", [(0, 0), (1, 3)])
def test_gemm_matches_oracle(ctx, dtype, shape, offs):
    m, n, k = shape
    r0, c0 = offs                                            # (1, 3): odd row offset -> not TMA-aligned
    rng = np.random.default_rng([4, m, n, k])
    lda = r0 + k + m + (3 if r0 else 0)
    big = rand_matrix(rng, lda, c0 + k + n, dtype)
    # layout inside `big`: A11-like k x k corner unused; B = rows r0.., cols c0+k..; A = rows r0+k.., cols c0..
    c_off, a_off, b_off = (r0 + k, c0 + k), (r0 + k, c0), (r0, c0 + k)
    want = O.schur_c(big.copy(order="F"), c_off, a_off, b_off, m, n, k, threads=4)
    d = Dev(ctx, big)
    ctx._check(fn(ctx, "rfb_gemm_nn_sub", dtype)(ctx.handle, d.at(*c_off), d.at(*a_off), d.at(*b_off), m, n, k, d.lda))
    got = d.get()
    tol = 4 * k * float(np.finfo(dtype).eps) * max(1.0, float(np.abs(want).max()))
    assert np.allclose(got, want, rtol=0, atol=tol)
    untouched = np.ones_like(big, dtype=bool)
    untouched[c_off[0]:c_off[0] + m, c_off[1]:c_off[1] + n] = False
    assert np.array_equal(got[untouched], big[untouched])
    d.free()


def test_gemm_linearity_large(ctx):
    """Size-independent property at a BASELINE-sized trailing update (8192 x 8192 x 256):
    C - A(B1 + B2) == (C - A B1) - A B2 up to rounding, checked on a random probe."""
    m = n = 8192; k = 256
    rng = np.random.default_rng(9)
    big = rand_matrix(rng, m + k, n + k, np.float64)
    d = Dev(ctx, big)
    ctx._check(ctx._lib.rfb_gemm_nn_sub_f64(ctx.handle, d.at(k, k), d.at(k, 0), d.at(0, k), m, n, k, d.lda))
    got = d.get()
    x = rng.random(n)
    want = big[k:, k:] @ x - big[k:, :k] @ (big[:k, k:] @ x)
    assert np.allclose(got[k:, k:] @ x, want, rtol=1e-10, atol=1e-8)
    d.free()


# ---- K4' Float32 on tcgen05 (3xTF32) -------------------------------------------------------------
@pytest.mark.parametrize("shape", [(128, 128, 32), (128, 128, 8), (256, 384, 96), (1000, 520, 300), (2048, 2048, 64),
                                   (1536, 1024, 1024), (4096, 4096, 512)])
def test_gemm_f32_tcgen05_matches_oracle(ctx, shape):
    """tcgen05.mma kind::tf32 with the 3-term split must be FP32-level accurate: within 0.5*k*eps(Float32)
    of the exact-FP32 oracle loop (measured: 1.5-3x the FFMA kernel's own error, because the tensor core
    accumulates with truncation; plain TF32 would miss this tolerance by ~20x)."""
    m, n, k = shape
    k4 = (k + 3) // 4 * 4
    rng = np.random.default_rng([44, m, n, k])
    lda = (k4 + m + 3) // 4 * 4
    big = rand_matrix(rng, lda, k4 + n, np.float32)
    c_off, a_off, b_off = (k4, k4), (k4, 0), (0, k4)
    want = O.schur_c(big.copy(order="F"), c_off, a_off, b_off, m, n, k, threads=4)
    d = Dev(ctx, big)
    ctx.set_default_opts(f32_mode=1)
    try:
        before = ctx.profile_read()["gemm"]["launches"]
        ctx._check(ctx._lib.rfb_gemm_nn_sub_f32(ctx.handle, d.at(*c_off), d.at(*a_off), d.at(*b_off), m, n, k, d.lda))
        got = d.get()
    finally:
        ctx.set_default_opts()
    tol = 0.5 * max(k, 16) * float(np.finfo(np.float32).eps) * max(1.0, float(np.abs(want).max()))
    err = float(np.abs(got - want).max())
    assert err <= tol, (err, tol)
    untouched = np.ones_like(big, dtype=bool)
    untouched[c_off[0]:c_off[0] + m, c_off[1]:c_off[1] + n] = False
    assert np.array_equal(got[untouched], big[untouched])
    d.free()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (64, 64), (65, 130), (200, 700), (256, 256), (257, 100), (1000, 300), (2048, 1500)])
def test_trsm_upper_matches_numpy(ctx, dtype, shape):
    """Back-substitution kernel (non-unit upper) against a float64 triangular solve."""
    import scipy.linalg as sl
    k, nrhs = shape
    rng = np.random.default_rng([33, k, nrhs])
    big = rand_matrix(rng, k + 5, k + nrhs + 3, dtype)
    u = np.triu(big[2:2 + k, 1:1 + k]).astype(np.float64)
    u[np.arange(k), np.arange(k)] += 2.0                       # well-conditioned diagonal
    big[2:2 + k, 1:1 + k] = np.where(np.triu(np.ones((k, k), dtype=bool)), u, big[2:2 + k, 1:1 + k]).astype(dtype)
    if k > 16:
        big[2:2 + k, 1:1 + k] = np.where(np.triu(np.ones((k, k), dtype=bool), 1), big[2:2 + k, 1:1 + k] * dtype(2.0 / k), big[2:2 + k, 1:1 + k])
    uu = np.triu(big[2:2 + k, 1:1 + k].astype(np.float64))
    want = sl.solve_triangular(uu, big[2:2 + k, k + 2:k + 2 + nrhs].astype(np.float64), lower=False)
    d = Dev(ctx, big)
    ctx._check(fn(ctx, "rfb_trsm_lunn", dtype)(ctx.handle, d.at(2, 1), k, d.at(2, k + 2), nrhs, d.lda))
    got = d.get()
    tol = 50 * max(k, 8) * float(np.finfo(dtype).eps) * max(1.0, float(np.abs(want).max()))
    assert np.abs(got[2:2 + k, k + 2:k + 2 + nrhs] - want).max() <= tol
    mask = np.ones_like(big, dtype=bool); mask[2:2 + k, k + 2:k + 2 + nrhs] = False
    assert np.array_equal(got[mask], big[mask])
    d.free()
