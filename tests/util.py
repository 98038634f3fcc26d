"""Shared helpers for the parity tests (the assertions of `testlu`, test/runtests.jl:14-31)."""
import numpy as np

from oracle import rf_oracle as O


def rand_matrix(rng, m, n, dtype):
    """rand(T, m, n): uniform [0, 1), column-major (test/runtests.jl:45)."""
    return np.asfortranarray(rng.random((m, n), dtype=np.dtype(dtype).type))


def ref_bound(m, dtype):
    """E = 20 * size(A, 1) * eps(T)   (test/runtests.jl:19)."""
    return 20.0 * m * float(np.finfo(dtype).eps)


def assert_testlu(a0, factors, ipiv, info, info_expected, wide=False):
    """The reference's own acceptance test for a pivoted, serial factorization.

    `wide=True` is for shapes far outside the reference's sweep (its fat cases are only s x (s+2)):
    the inf-norm sums over n columns, so E uses max(m, n); the CPU oracle itself needs that there
    (1000 x 3000: 1.03e-11 against 20*m*eps = 4.4e-12).
    """
    m, n = a0.shape
    assert info == info_expected                                   # runtests.jl:15
    if info != 0:
        return
    e = ref_bound(max(m, n) if wide else m, a0.dtype)
    r = O.residual_inf(a0, factors, ipiv)
    assert r < e, f"||LU - A[p,:]||_inf = {r} >= {e}"              # runtests.jl:20
    if m == n and m > 0:                                           # runtests.jl:21-28
        l, u = O.split_lu(np.asarray(factors, dtype=np.float64))
        p = O.perm_from_ipiv(ipiv, m)
        b = np.asarray(a0, dtype=np.float64)[:, -1][p]
        import scipy.linalg as sl
        y = sl.solve_triangular(l, b, lower=True, unit_diagonal=True)
        x = sl.solve_triangular(u, y, lower=False)
        if np.all(np.isfinite(x)):
            rhs = np.zeros(n)
            rhs[-1] = 1.0
            assert np.allclose(x, rhs, rtol=0, atol=100 * e) or np.allclose(x, rhs, rtol=np.sqrt(np.finfo(a0.dtype).eps), atol=100 * e)


def hutchinson_residual(a0, factors, ipiv, nvec=8, seed=0):
    """Estimate ||P A - L U||_F / ||A||_F with random +-1 probes in O(n^2) per probe."""
    rng = np.random.default_rng(seed)
    m, n = a0.shape
    p = O.perm_from_ipiv(ipiv, m)
    f = np.asarray(factors, dtype=np.float64)
    mn = min(m, n)
    x = rng.integers(0, 2, size=(n, nvec)).astype(np.float64) * 2 - 1
    ux = np.triu(f[:mn, :]) @ x
    lux = np.tril(f[:, :mn], -1) @ ux
    lux[:mn] += ux
    pax = (np.asarray(a0, dtype=np.float64) @ x)[p]
    num = np.linalg.norm(pax - lux) / np.sqrt(nvec)
    return float(num / np.linalg.norm(np.asarray(a0, dtype=np.float64)))


def assert_pivots_match(a0, factors, ipiv, want_ipiv, strict):
    """Pivot vectors must be identical.  With `strict=False` (Float32, where rounding noise of
    different summation orders reaches the gap between the two largest candidates of a column --
    SURVEY.md H4) a difference is accepted only if the FIRST differing step is a documented
    near-tie: the row the oracle picked has |L| >= 1 - c*n*eps in our factorization, i.e. the two
    candidates were equal to within the factorization's own error.  Everything after a legitimate
    near-tie diverges by construction and is covered by the residual test."""
    if np.array_equal(ipiv, want_ipiv):
        return
    assert not strict, f"pivot mismatch at step {int(np.argmax(np.asarray(ipiv) != np.asarray(want_ipiv)))}"
    m, n = a0.shape
    k = int(np.argmax(np.asarray(ipiv) != np.asarray(want_ipiv)))
    perm_o = O.perm_from_ipiv(np.asarray(want_ipiv)[: k + 1], m)   # row the oracle moved to position k
    perm_g = O.perm_from_ipiv(ipiv, m)
    orig = perm_o[k]
    pos = int(np.nonzero(perm_g == orig)[0][0])
    assert pos > k, (k, pos)
    ratio = abs(float(factors[pos, k]))
    tol = 20 * max(m, n) * float(np.finfo(a0.dtype).eps)
    assert ratio >= 1.0 - tol, f"pivot mismatch at step {k} is not a near-tie: |L[{pos},{k}]| = {ratio}"
