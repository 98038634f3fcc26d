"""Property-based shape sweeps (hypothesis), mirroring the reference's strategy of testing against the stdlib over many
shapes (test/runtests.jl:33-68) instead of golden factors: the oracle against LAPACK on the CPU, the CUDA path against
the oracle on the GPU.  Seeds are derived from the drawn shape, so every example is reproducible."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st
from scipy.linalg import lapack

from oracle import rf_oracle as O
from util import assert_pivots_match, assert_testlu, rand_matrix, ref_bound

SHAPES = st.tuples(st.integers(1, 140), st.integers(1, 140))
COMMON = dict(deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@settings(max_examples=60, **COMMON)
@given(shape=SHAPES, f32=st.booleans(), zero_col=st.booleans())
def test_oracle_matches_lapack_on_random_shapes(shape, f32, zero_col):
    m, n = shape
    dtype = np.float32 if f32 else np.float64
    a0 = rand_matrix(np.random.default_rng([5, m, n]), m, n, dtype)
    if zero_col:
        a0[:, (m * 7 + n) % min(m, n)] = 0
    _, piv, linfo = (lapack.sgetrf if f32 else lapack.dgetrf)(a0)
    f, ipiv, info = O.lu_c(a0.copy(order="F"))
    assert info == linfo                                           # runtests.jl:15
    if not zero_col:
        assert np.array_equal(ipiv, piv + 1)
        assert_testlu(a0, f, ipiv, info, linfo, wide=True)
    g, p2, i2 = O.lu_numpy(a0.copy(order="F"))
    assert i2 == info and np.array_equal(p2, ipiv)


@settings(max_examples=40, **COMMON)
@given(shape=SHAPES, f32=st.booleans())
def test_oracle_nopivot_matches_twin_on_random_shapes(shape, f32):
    m, n = shape
    dtype = np.float32 if f32 else np.float64
    a0 = rand_matrix(np.random.default_rng([6, m, n]), m, n, dtype)
    k = min(m, n)
    a0[np.arange(k), np.arange(k)] += 10
    f, _, info = O.lu_nopiv_c(a0.copy(order="F"))
    g, _, i2 = O.lu_numpy(a0.copy(order="F"), pivot=False)
    assert info == i2 == 0
    assert np.allclose(f, g, rtol=0, atol=10 * ref_bound(max(m, n), dtype))


@pytest.mark.gpu
@settings(max_examples=40, **COMMON)
@given(shape=st.tuples(st.integers(1, 400), st.integers(1, 400)), f32=st.booleans(), pivot=st.booleans())
def test_gpu_matches_oracle_on_random_shapes(ctx, shape, f32, pivot):
    import rfb200
    m, n = shape
    dtype = np.float32 if f32 else np.float64
    a0 = rand_matrix(np.random.default_rng([7, m, n]), m, n, dtype)
    if pivot:
        F = rfb200.lu(a0, ctx=ctx)
        f, p, info = O.lu_c(a0.copy(order="F"))
        assert F.info == info == 0
        assert_pivots_match(a0, F.factors, F.ipiv, p, strict=not f32)
        assert_testlu(a0, F.factors, F.ipiv, F.info, 0, wide=True)
    else:
        k = min(m, n)
        a0[np.arange(k), np.arange(k)] += 10
        F = rfb200.lu(a0, False, ctx=ctx)
        f, _, info = O.lu_nopiv_c(a0.copy(order="F"))
        assert F.info == info == 0
        assert np.allclose(F.factors, f, rtol=0, atol=10 * ref_bound(max(m, n), dtype))


@pytest.mark.gpu
@settings(max_examples=25, **COMMON)
@given(batch=st.integers(1, 40), m=st.integers(1, 128), n=st.integers(1, 64), f32=st.booleans())
def test_gpu_batched_matches_oracle_on_random_shapes(ctx, batch, m, n, f32):
    import rfb200
    dtype = np.float32 if f32 else np.float64
    a3 = np.random.default_rng([8, batch, m, n]).random((batch, m, n)).astype(dtype)
    Fs = rfb200.lu_batched(a3, check=False, ctx=ctx)
    for b in range(batch):
        wf, wp, winfo = O.panel_c(np.asfortranarray(a3[b]))
        assert Fs[b].info == winfo and np.array_equal(Fs[b].ipiv, wp) and np.array_equal(Fs[b].factors, wf)
