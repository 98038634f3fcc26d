"""CPU tests that pin the oracle's restatement of the rows SURVEY.md section 8f widens into:
pivot = Val(false) (src/lu.jl:27-65, :107-113, :249-254, :323-326), the NotIPIV solve (:60-64) and the
butterfly solver (src/butterflylu.jl).  Pinned by the reference's own tests for them
(test/runtests.jl:14-31 with the unpivoted tolerance, :70-84, :116-128, :130-159 (wilkinson + 🦋)), by an independent numpy
twin and by the algebraic definition U' A V of the transform."""
import numpy as np
import pytest

from oracle import rf_oracle as O
from util import rand_matrix, ref_bound

REF_SIZES = list(range(1, 11)) + [50, 130, 300]          # runtests.jl:39


def dominant(rng, m, n, dtype):
    """rand(T, n, n) + 10I (runtests.jl:75, :120): safe to factor without pivoting."""
    a = rand_matrix(rng, m, n, dtype)
    k = min(m, n)
    a[np.arange(k), np.arange(k)] += 10
    return a


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("s", REF_SIZES)
def test_nopivot_reference_sweep(dtype, s):
    """testlu with pivot = false (runtests.jl:14-31): info equality and ||L U - A||_inf < 10 sqrt(20 m eps)."""
    rng = np.random.default_rng([13, s, np.dtype(dtype).itemsize])
    for (m, n) in ((s, s), (s, s + 2), (s + 2, s)):
        a0 = dominant(rng, m, n, dtype)
        f, ipiv, info = O.lu_nopiv_c(a0.copy(order="F"))
        assert info == 0 and ipiv is None
        l, u = O.split_lu(np.asarray(f, dtype=np.float64))
        e = 10 * np.sqrt(ref_bound(m, dtype))
        assert np.abs(l @ u - a0).sum(axis=1).max() < e
        f2, p2, i2 = O.lu_numpy(a0.copy(order="F"), pivot=False)
        assert i2 == 0 and np.array_equal(p2, np.arange(1, min(m, n) + 1))
        assert np.allclose(f, f2, rtol=0, atol=ref_bound(max(m, n), dtype) * 10)
        for threads in (3,):
            f3, _, _ = O.lu_nopiv_c(a0.copy(order="F"), threads=threads)
            assert np.allclose(f, f3, rtol=0, atol=ref_bound(max(m, n), dtype) * 10)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_nopivot_user_ipiv_is_identity(dtype):
    """runtests.jl:70-84: a poisoned user ipiv comes back as 1:n, the solve through it is accurate."""
    n = 30
    rng = np.random.default_rng(1)
    a = dominant(rng, n, n, dtype)
    b = rng.random(n).astype(dtype)
    ipiv = np.full(n, np.iinfo(np.int64).max - 7, dtype=np.int64)
    f, p, info = O.lu_nopiv_c(a.copy(order="F"), ipiv)
    assert p is ipiv and np.array_equal(ipiv, np.arange(1, n + 1)) and info == 0
    x = O.ldiv_notipiv_c(f, b.copy())
    assert np.linalg.norm(a.astype(np.float64) @ x - b) < 1000 * n * np.finfo(dtype).eps


def test_nopivot_negative_info():
    """Julia >= 1.11 convention (src/lu.jl:24-25, :323-326, :249-251): first zero pivot k -> info = -k, in
    the leaf and through the recursion's `info -= n1`."""
    assert O.lu_nopiv_c(np.zeros((100, 100), order="F"))[2] == -1
    for n, k in ((300, 100), (300, 299), (130, 64), (40, 7)):
        a = np.asfortranarray(np.eye(n))
        a[k, k] = 0
        assert O.lu_nopiv_c(a.copy(order="F"))[2] == -(k + 1)
        assert O.lu_numpy(a.copy(order="F"), pivot=False)[2] == -(k + 1)
    a = np.asfortranarray(np.eye(20, 64))
    a[5, 5] = 0
    assert O.panel_nopiv_c(a.T.copy(order="F"))[1] == -6


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [8, 64, 200, 300])
def test_notipiv_ldiv(dtype, n):
    """runtests.jl:116-128."""
    rng = np.random.default_rng([2, n])
    a = dominant(rng, n, n, dtype)
    b = rng.random(n).astype(dtype)
    bb = np.asfortranarray(rng.random((n, 3)).astype(dtype))
    f, _, info = O.lu_nopiv_c(a.copy(order="F"))
    assert info == 0
    x = O.ldiv_notipiv_c(f, b.copy())
    assert x.dtype == dtype and np.linalg.norm(a.astype(np.float64) @ x - b) < 1000 * n * np.finfo(dtype).eps
    xx = O.ldiv_notipiv_c(f, bb.copy(order="F"))
    assert np.linalg.norm(a.astype(np.float64) @ xx - bb) < 1000 * n * np.finfo(dtype).eps


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("m", [4, 8, 12, 64, 200])
def test_butterfly_mul_is_ut_a_v(dtype, m):
    """🦋mul! (src/butterflylu.jl:93-113): C loop == numpy twin bit for bit, == U' A V with the materialised
    U, V of materializeUV (:149-178) to rounding."""
    rng = np.random.default_rng([3, m])
    a = rand_matrix(rng, m, m, dtype)
    uv = O.butterfly_vals(m, dtype)
    assert uv.size == 4 * m and np.all(uv > 0.47) and np.all(uv < 0.53)
    c = O.butterfly_mul_c(a.copy(order="F"), uv)
    t = O.butterfly_mul_numpy(a.copy(order="F"), uv)
    assert np.array_equal(c, t)
    u, v = O.butterfly_materialize(uv.astype(np.float64), m)
    want = u.T @ a.astype(np.float64) @ v
    assert np.allclose(c, want, rtol=0, atol=50 * np.finfo(dtype).eps * np.abs(want).max())
    # U and V are (scaled) orthogonal-like butterflies: well conditioned, so the transform is benign
    assert np.linalg.cond(u) < 2 and np.linalg.cond(v) < 2


def wilkinson(n):
    """test/runtests.jl:130-140 (`wilkinson`): the classic growth-factor matrix."""
    a = -np.tril(np.ones((n, n)), -1) + np.eye(n)
    a[:, -1] = 1
    return np.asfortranarray(a)


@pytest.mark.parametrize("n", [790, 797, 803, 808, 810])
def test_butterfly_solve_wilkinson(n):
    """runtests.jl:142-159: ||A x - b|| <= 1e-8 ||b|| on Wilkinson matrices (sizes from 790:810, padded and not)."""
    rng = np.random.default_rng([1234, n])
    a, b = wilkinson(n), rng.random(n)
    x, info = O.butterfly_solve_oracle(a, b)
    assert info == 0
    assert np.linalg.norm(a @ x - b) <= 1e-8 * np.linalg.norm(b)


def test_butterfly_pad():
    a = np.asfortranarray(np.arange(25, dtype=np.float64).reshape(5, 5))
    p = O.butterfly_pad(a)
    assert p.shape == (8, 8) and np.array_equal(p[:5, :5], a)
    assert np.array_equal(p[5:, 5:], np.eye(3)) and not p[:5, 5:].any() and not p[5:, :5].any()
