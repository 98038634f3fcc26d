"""GPU parity tests for the rows SURVEY.md section 8f widens into, through the C ABI / host mirror:
pivot = Val(false) + NotIPIV solve (src/lu.jl:27-65, :107-113, :249-254, :323-326), the butterfly solver
(src/butterflylu.jl), Adjoint/Transpose wrappers (src/lu.jl:85-87) and the batched small-matrix LU.
Each test mirrors the reference test it cites and compares with the CPU oracle on the same inputs."""
import ctypes as C

import numpy as np
import pytest

import rfb200
from oracle import rf_oracle as O
from util import hutchinson_residual, rand_matrix, ref_bound
from test_gpu_kernels import Dev, dev_i64, fn, get_i64
from test_oracle_widened import dominant, wilkinson

pytestmark = pytest.mark.gpu

REF_SIZES = list(range(1, 11)) + [50, 130, 300]


# ---- K1' unpivoted panel ----------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(1, 1), (5, 3), (8, 8), (64, 64), (100, 16), (129, 64), (193, 64), (300, 40),
                                   (1000, 64), (5000, 33), (16384, 64), (40000, 32)])
def test_panel_nopiv_matches_oracle_bitwise(ctx, dtype, shape):
    m, n = shape
    rng = np.random.default_rng([21, m, n])
    a0 = dominant(rng, m, n, dtype)
    want_f, want_info = O.panel_nopiv_c(a0.copy(order="F"))
    big = np.asfortranarray(rng.random((m + 7, n + 5)).astype(dtype))      # sub-block of a larger allocation
    big[3:3 + m, 2:2 + n] = a0
    d = Dev(ctx, big)
    info = dev_i64(ctx, np.zeros(8))
    ctx._check(fn(ctx, "rfb_panel_getrf_nopiv", dtype)(ctx.handle, d.at(3, 2), m, n, d.lda, C.c_void_p(info), 11))
    got = d.get()
    assert int(get_i64(ctx, info, 1)[0]) == want_info == 0
    # same operation order, same reciprocal, same FMA as the unblocked loop: identical bits
    assert np.array_equal(got[3:3 + m, 2:2 + n], want_f)
    outside = np.ones_like(big, dtype=bool); outside[3:3 + m, 2:2 + n] = False
    assert np.array_equal(got[outside], big[outside])
    d.free(); ctx.free(info)


def test_panel_nopiv_zero_pivot_and_nan(ctx):
    rng = np.random.default_rng(4)
    m, n = 500, 48
    a0 = np.asfortranarray(rng.integers(-3, 4, size=(m, n)).astype(np.float64))
    a0[np.arange(n), np.arange(n)] = 5
    a0[20, 20] = 0; a0[20, :20] = 0; a0[:20, 20] = 0          # pivot 20 stays exactly zero
    a0[30, 31] = np.nan
    want_f, want_info = O.panel_nopiv_c(a0.copy(order="F"))
    assert want_info == -21
    d = Dev(ctx, a0)
    info = dev_i64(ctx, np.zeros(8))
    ctx._check(ctx._lib.rfb_panel_getrf_nopiv_f64(ctx.handle, d.at(0, 0), m, n, d.lda, C.c_void_p(info), 100))
    got = d.get()
    assert int(get_i64(ctx, info, 1)[0]) == -(100 + 21)
    assert np.array_equal(np.isnan(got), np.isnan(want_f))
    assert np.array_equal(np.nan_to_num(got, nan=7.0), np.nan_to_num(want_f, nan=7.0))
    d.free(); ctx.free(info)


# ---- lu / lu! with pivot = false ---------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("s", REF_SIZES)
def test_nopivot_reference_sweep(ctx, dtype, s):
    """testlu (runtests.jl:14-31) with pivot = false over the reference's shapes (:39-42, :55-58)."""
    rng = np.random.default_rng([13, s, np.dtype(dtype).itemsize])
    for k, (m, n) in enumerate(((s, s), (s, s + 2), (s + 2, s))):
        a0 = dominant(rng, m, n, dtype)
        F = rfb200.lu(a0, rfb200.NoPivot() if k % 2 else False, ctx=ctx)
        want_f, _, want_info = O.lu_nopiv_c(a0.copy(order="F"))
        assert F.info == want_info == 0                                   # runtests.jl:15
        assert isinstance(F.ipiv, rfb200.NotIPIV) and len(F.ipiv) == min(m, n)
        L, U, p = F
        assert np.array_equal(p, np.arange(m))
        e = 10 * np.sqrt(ref_bound(m, dtype))                            # runtests.jl:19 (unpivoted)
        assert np.abs(L.astype(np.float64) @ U.astype(np.float64) - a0).sum(axis=1).max() < e
        # and far tighter than that: the same factorization as the oracle up to summation order
        assert np.allclose(F.factors, want_f, rtol=0, atol=10 * ref_bound(max(m, n), dtype))
        if m == n:                                                        # runtests.jl:21-28
            x = F.solve(a0[:, -1])
            rhs = np.zeros(n); rhs[-1] = 1
            assert np.allclose(x, rhs, rtol=0, atol=100 * e)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_nopivot_user_ipiv(ctx, dtype):
    """runtests.jl:70-84."""
    n = 30
    rng = np.random.default_rng(1)
    a = dominant(rng, n, n, dtype)
    b = rng.random(n).astype(dtype)
    ipiv = np.full(n, np.iinfo(np.int64).max - 7, dtype=np.int64)          # poison
    F = rfb200.lu_(a.copy(order="F"), ipiv, False, False, ctx=ctx)
    assert F.ipiv is ipiv and np.array_equal(ipiv, np.arange(1, n + 1))
    x = rfb200.ldiv_(F, b.copy(), ctx=ctx)                                # the pivoted-LU solve path consumes F.ipiv
    assert np.linalg.norm(a.astype(np.float64) @ x - b) < 1000 * n * np.finfo(dtype).eps


def test_nopivot_negative_info(ctx):
    for n, k in ((300, 100), (300, 299), (130, 64), (40, 7), (1000, 640)):
        a = np.asfortranarray(np.eye(n))
        a[k, k] = 0
        want = O.lu_nopiv_c(a.copy(order="F"))[2]
        F = rfb200.lu(a, False, check=False, ctx=ctx)
        assert F.info == want == -(k + 1)
        with pytest.raises(rfb200.ZeroPivotException):
            rfb200.lu(a, False, ctx=ctx)
    assert rfb200.lu(np.zeros((100, 100)), False, check=False, ctx=ctx).info == -1


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [8, 64, 200, 300])
def test_notipiv_ldiv(ctx, dtype, n):
    """runtests.jl:116-128."""
    rng = np.random.default_rng([2, n])
    a = dominant(rng, n, n, dtype)
    b = rng.random(n).astype(dtype)
    bb = np.asfortranarray(rng.random((n, 3)).astype(dtype))
    F = rfb200.lu(a, False, ctx=ctx)
    x = rfb200.ldiv_(F, b.copy(), ctx=ctx)
    assert x.dtype == dtype and x.shape == (n,)
    assert np.linalg.norm(a.astype(np.float64) @ x - b) < 1000 * n * np.finfo(dtype).eps
    xx = rfb200.ldiv_(F, bb.copy(order="F"), ctx=ctx)
    assert np.linalg.norm(a.astype(np.float64) @ xx - bb) < 1000 * n * np.finfo(dtype).eps
    f, _, _ = O.lu_nopiv_c(a.copy(order="F"))
    assert np.allclose(x, O.ldiv_notipiv_c(f, b.copy()), rtol=0, atol=1000 * n * np.finfo(dtype).eps)


def test_nopivot_4096_properties(ctx):
    n = 4096
    rng = np.random.default_rng([12, 4096])
    a0 = dominant(rng, n, n, np.float64)
    a0[np.arange(n), np.arange(n)] += n / 4                               # strongly dominant: growth-free
    F = rfb200.lu(a0, False, ctx=ctx)
    assert F.info == 0
    r = hutchinson_residual(a0, F.factors, np.arange(1, n + 1))
    assert r < 20 * n * np.finfo(np.float64).eps
    want_f, _, _ = O.lu_nopiv_c(a0.copy(order="F"), threads=8)
    assert np.allclose(F.factors, want_f, rtol=0, atol=1e-9)


# ---- Adjoint / Transpose (src/lu.jl:85-87) -----------------------------------------------------------
def test_adjoint_wrapper(ctx):
    a = rand_matrix(np.random.default_rng(3), 120, 100, np.float64)
    F = rfb200.lu(rfb200.Adjoint(a), ctx=ctx)
    G = rfb200.lu(a, ctx=ctx)
    assert isinstance(F, rfb200.AdjointLU) and np.array_equal(F.parent.factors, G.factors)
    assert np.array_equal(F.parent.ipiv, G.ipiv) and F.info == 0
    assert isinstance(rfb200.lu(rfb200.Transpose(a), False, check=False, ctx=ctx), rfb200.AdjointLU)


# ---- butterfly ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("m", [4, 8, 12, 64, 200, 516, 2048])
def test_butterfly_mul_bitwise(ctx, dtype, m):
    """🦋mul! (src/butterflylu.jl:93-113): the fused one-pass kernel reproduces the two-level loop's bits."""
    rng = np.random.default_rng([3, m])
    a = rand_matrix(rng, m, m, dtype)
    uv = O.butterfly_vals(m, dtype)
    want = O.butterfly_mul_c(a.copy(order="F"), uv)
    got = rfb200.butterfly_mul_(a.copy(order="F"), uv, ctx=ctx)
    assert np.array_equal(got, want)


def test_butterfly_vec_kernels(ctx):
    m, nrhs = 520, 3
    rng = np.random.default_rng(8)
    uv = O.butterfly_vals(m)
    u, v = O.butterfly_materialize(uv, m)
    b = np.asfortranarray(rng.random((m, nrhs)))
    d, duv = Dev(ctx, b), Dev(ctx, uv.reshape(-1, 1))
    ctx._check(ctx._lib.rfb_butterfly_vec_f64(ctx.handle, d.at(0, 0), m, nrhs, m, duv.at(0, 0), 0))
    assert np.allclose(d.get(), u.T @ b, rtol=0, atol=1e-14)              # mul!(tmp, U', b), :50
    ctx.h2d(d.ptr, b)
    ctx._check(ctx._lib.rfb_butterfly_vec_f64(ctx.handle, d.at(0, 0), m, nrhs, m, duv.at(0, 0), 1))
    assert np.allclose(d.get(), v @ b, rtol=0, atol=1e-14)                # mul!(b, V, tmp), :52
    d.free(); duv.free()


@pytest.mark.parametrize("n", list(range(790, 811)))
def test_butterfly_solve_wilkinson(ctx, n):
    """runtests.jl:142-159: ||A x - b|| <= 1e-8 ||b|| on Wilkinson matrices 790..810."""
    rng = np.random.default_rng([1234, n])
    a, b = wilkinson(n), rng.random(n)
    ws = rfb200.ButterflyWorkspace(a.copy(order="F"), b.copy())
    before = ctx.launch_count()
    out = rfb200.butterfly_solve_(ws, True, ctx=ctx)
    assert ctx.launch_count() > before
    assert np.linalg.norm(a @ out - b) <= 1e-8 * np.linalg.norm(b)
    assert np.array_equal(ws.A, a) and np.array_equal(ws.b, b)            # host inputs are not modified
    want, winfo = O.butterfly_solve_oracle(a, b, uv=ws.ws)
    assert winfo == ws.info == 0
    assert np.allclose(out, want, rtol=0, atol=1e-9 * np.abs(want).max())


@pytest.mark.parametrize("dtype,n,nrhs", [(np.float64, 1024, 1), (np.float64, 2050, 4), (np.float32, 512, 2), (np.float64, 5, 1),
                                          (np.float64, 3, 2)])
def test_butterfly_solve_random(ctx, dtype, n, nrhs):
    rng = np.random.default_rng([9, n])
    a = dominant(rng, n, n, dtype)
    b = np.asfortranarray(rng.random((n, nrhs)).astype(dtype))
    ws = rfb200.ButterflyWorkspace(a, b if nrhs > 1 else b[:, 0])
    out = rfb200.butterfly_solve_(ws, ctx=ctx)
    x = out.reshape(n, nrhs)
    assert np.linalg.norm(a.astype(np.float64) @ x - b) < 1000 * n * np.finfo(dtype).eps * np.linalg.norm(b)


# ---- batched small LU --------------------------------------------------------------------------------
def batch_array(rng, batch, m, n, dtype, lda=None):
    lda = lda or m
    buf = np.zeros((batch, n, lda), dtype=dtype)
    a = buf.transpose(0, 2, 1)[:, :m, :]
    a[...] = rng.random((batch, m, n))
    return a


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(100, 8, 8, None), (37, 33, 20, None), (1000, 64, 64, None), (10, 10, 12, None),
                                   (5, 128, 64, None), (64, 40, 40, 48), (3, 1, 1, None), (200, 17, 64, None),
                                   (7, 100, 33, 101)])
def test_batched_small_matches_oracle_bitwise(ctx, dtype, shape):
    batch, m, n, lda = shape
    rng = np.random.default_rng([31, batch, m, n])
    a = batch_array(rng, batch, m, n, dtype, lda)
    a0 = a.copy()
    if batch > 2:
        a[1][:, min(m, n) // 2] = 0                                       # one singular matrix in the batch
        a0 = a.copy()
    before = ctx.launch_count()
    Fs = rfb200.lu_batched_(a, check=False, ctx=ctx)
    assert ctx.launch_count() == before + 1                               # one launch for the whole batch
    for b in range(batch):
        want_f, want_p, want_info = O.panel_c(np.asfortranarray(a0[b]))
        assert Fs[b].info == want_info
        assert np.array_equal(Fs[b].ipiv, want_p)
        assert np.array_equal(Fs[b].factors, want_f)                      # the unblocked loop's own bits
    if batch > 2:
        assert Fs[1].info == min(m, n) // 2 + 1 and Fs[0].info == 0
        with pytest.raises(rfb200.SingularException):
            rfb200.lu_batched(a0, ctx=ctx)


def test_batched_large_falls_back_to_recursive_driver(ctx):
    rng = np.random.default_rng(5)
    a = batch_array(rng, 3, 200, 200, np.float64)
    a0 = a.copy()
    Fs = rfb200.lu_batched_(a, ctx=ctx)
    for b in range(3):
        G = rfb200.lu(a0[b], ctx=ctx)
        assert np.array_equal(Fs[b].factors, G.factors) and np.array_equal(Fs[b].ipiv, G.ipiv)
    Fn = rfb200.lu_batched(np.stack([dominant(rng, 50, 50, np.float64) for _ in range(4)]), False, ctx=ctx)
    assert all(isinstance(f.ipiv, rfb200.NotIPIV) and f.info == 0 for f in Fn)


def test_batched_throughput_shape_16k(ctx):
    """many 32 x 32 Jacobian-sized factorizations in one launch; spot-check against the oracle."""
    rng = np.random.default_rng(6)
    a = batch_array(rng, 16384, 32, 32, np.float64)
    a0 = a.copy()
    Fs = rfb200.lu_batched_(a, ctx=ctx)
    for b in (0, 1, 8191, 16383):
        want_f, want_p, _ = O.panel_c(np.asfortranarray(a0[b]))
        assert np.array_equal(Fs[b].ipiv, want_p) and np.array_equal(Fs[b].factors, want_f)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("pivot", [True, False])
def test_ldiv_with_device_resident_factors(ctx, dtype, pivot):
    """`lu_(A, keep=True)` leaves the factors on the device; `ldiv_` then only moves B (rfb_solve_kept_*) and must give
    exactly what the uploading path gives; once another host-mode call has reused the staging buffer the kept id is
    stale and `ldiv_` silently goes back to uploading F.factors."""
    n = 700
    rng = np.random.default_rng([71, n, int(pivot)])
    a0 = rand_matrix(rng, n, n, dtype)
    if not pivot:
        a0[np.arange(n), np.arange(n)] += dtype(n / 4)
    b = rand_matrix(rng, n, 5, dtype)
    F = rfb200.lu_(a0.copy(order="F"), None, pivot, ctx=ctx, keep=True)
    assert F._kept is not None
    before = ctx.launch_count()
    x_kept = rfb200.ldiv_(F, b.copy(order="F"), ctx=ctx)
    kid = C.c_int64(0)
    ctx._check(ctx._lib.rfb_kept_id(ctx.handle, C.byref(kid)))
    assert kid.value == F._kept[1] and ctx.launch_count() > before           # still resident after the solve
    G = rfb200.LU(F.factors, F.ipiv, F.info)                                  # no kept handle: the uploading path
    x_up = rfb200.ldiv_(G, b.copy(order="F"), ctx=ctx)
    assert np.array_equal(x_kept, x_up)
    ctx._check(ctx._lib.rfb_kept_id(ctx.handle, C.byref(kid)))
    assert kid.value == 0                                                     # the upload reused the staging buffer
    x_again = rfb200.ldiv_(F, b.copy(order="F"), ctx=ctx)                     # stale id -> falls back, same answer
    assert np.array_equal(x_again, x_up)
    v = rfb200.ldiv_(rfb200.lu_(a0.copy(order="F"), None, pivot, ctx=ctx, keep=True), b[:, 0].copy(), ctx=ctx)
    assert np.array_equal(v, x_up[:, 0]) or np.allclose(v, x_up[:, 0], rtol=0, atol=1000 * n * np.finfo(dtype).eps)
    with pytest.raises(rfb200.RfbError):
        fnk = ctx._lib.rfb_solve_kept_f64 if dtype == np.float64 else ctx._lib.rfb_solve_kept_f32
        ctx._check(fnk(ctx.handle, 123456789, C.c_void_p(b.ctypes.data), 5, n))
