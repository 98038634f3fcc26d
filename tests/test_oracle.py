"""CPU tests that PIN the oracle (oracle/rf_oracle.{c,py}).

What pins it (the reference has no golden vectors and cannot run here -- SURVEY.md section 8c):
  * the reference's own acceptance test `testlu` (test/runtests.jl:14-31) over its own shape sweep
    (:39-42) and its singular-column case (:59-64);
  * LAPACK getrf golden fixtures (tests/golden/) -- pivots and info exact, factors to tolerance;
  * an independent numpy restatement of src/lu.jl (oracle.rf_oracle.lu_numpy).
"""
import os
import sys

import numpy as np
import pytest
from scipy.linalg import lapack

from oracle import rf_oracle as O
from util import assert_testlu, rand_matrix, ref_bound

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import CASES, make_input  # noqa: E402

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "getrf_golden.npz"))

REF_SIZES = list(range(1, 11)) + [50, 130, 300]          # [1:10; 50:80:200; 300], runtests.jl:39


def test_nsplit_values():
    # src/lu.jl:158-162, values listed in SURVEY.md section 2
    assert [O.nsplit(np.float64, n) for n in (64, 100, 130, 300, 4096, 9, 15, 16)] == [32, 48, 64, 152, 2048, 4, 7, 8]
    assert [O.nsplit(np.float32, n) for n in (300, 8192, 31, 32)] == [144, 4096, 15, 16]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("s", REF_SIZES)
def test_reference_sweep(dtype, s):
    """square, fat (s, s+2) and tall (s+2, s) like runtests.jl:41-58, plus the singular column."""
    rng = np.random.default_rng([12, s, np.dtype(dtype).itemsize])
    for (m, n) in ((s, s), (s, s + 2), (s + 2, s)):
        a0 = rand_matrix(rng, m, n, dtype)
        getrf = lapack.dgetrf if dtype == np.float64 else lapack.sgetrf
        _, piv, linfo = getrf(a0)
        for threads in (1, 3):                               # JULIA_NUM_THREADS in {1, 3}
            f, ipiv, info = O.lu_c(a0.copy(order="F"), threads=threads)
            assert_testlu(a0, f, ipiv, info, linfo)
            assert np.array_equal(ipiv, piv + 1)
        # singular column: info must equal the stdlib's (runtests.jl:59-64)
        a1 = a0.copy(order="F")
        i = int(rng.integers(0, min(m, n)))
        a1[:, i] = 0
        _, _, linfo = getrf(a1)
        f, ipiv, info = O.lu_c(a1.copy(order="F"))
        assert info == linfo and info > 0


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_golden_lapack(idx):
    name, m, n, dt, special = CASES[idx]
    a0 = make_input(idx)
    assert float(np.asarray(a0, dtype=np.float64).sum()) == float(GOLDEN[name + "__checksum"]), "RNG stream drifted"
    f, ipiv, info = O.lu_c(a0.copy(order="F"))
    assert info == int(GOLDEN[name + "__info"])
    assert np.array_equal(ipiv, GOLDEN[name + "__ipiv"])      # bit-exact pivot indices
    tol = 50 * max(m, n) * np.finfo(a0.dtype).eps
    du = GOLDEN[name + "__diagu"]
    scale = max(1.0, float(np.abs(du).max()))
    assert np.allclose(np.diag(f), du, rtol=0, atol=tol * scale)
    if name + "__lu" in GOLDEN.files:
        assert np.allclose(f, GOLDEN[name + "__lu"], rtol=0, atol=tol * scale)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(5, 5), (48, 48), (49, 49), (64, 64), (100, 103), (131, 97), (300, 300)])
def test_c_matches_numpy_twin(dtype, shape):
    rng = np.random.default_rng([7, shape[0], shape[1]])
    a0 = rand_matrix(rng, shape[0], shape[1], dtype)
    f1, p1, i1 = O.lu_c(a0.copy(order="F"))
    f2, p2, i2 = O.lu_numpy(a0.copy(order="F"))
    assert np.array_equal(p1, p2) and i1 == i2
    assert np.allclose(f1, f2, rtol=0, atol=ref_bound(shape[0], dtype))


def test_kernel_level_entry_points():
    rng = np.random.default_rng(3)
    # panel == unblocked getrf
    a0 = rand_matrix(rng, 200, 16, np.float64)
    f, ipiv, info = O.panel_c(a0.copy(order="F"))
    lu, piv, linfo = lapack.dgetrf(a0)
    assert np.array_equal(ipiv, piv + 1) and info == linfo and np.allclose(f, lu, atol=1e-13)
    # laswp
    a = rand_matrix(rng, 40, 7, np.float64)
    piv = np.array([3, 3, 10, 4, 40, 6], dtype=np.int64)
    want = a.copy()
    for i, ip in enumerate(piv):
        want[[i, ip - 1], :] = want[[ip - 1, i], :]
    assert np.array_equal(O.laswp_c(a.copy(order="F"), piv), want)
    # trsm / schur on sub-blocks of one allocation
    big = rand_matrix(rng, 300, 300, np.float64)
    l = np.tril(big[:150, :150], -1) + np.eye(150)
    want = np.linalg.solve(l, big[:150, 150:290])
    got = O.trsm_c(big.copy(order="F"), (0, 0), 150, (0, 150), 140, threads=2)[:150, 150:290]
    assert np.allclose(got, want, atol=1e-9)
    want = big[150:, 150:] - big[150:, :150] @ big[:150, 150:]
    got = O.schur_c(big.copy(order="F"), (150, 150), (150, 0), (0, 150), 150, 150, 150, threads=2)[150:, 150:]
    assert np.allclose(got, want, atol=1e-11)


def test_edge_cases():
    # empty, zero matrix (info = 1, factorization continues), NaN never selected as pivot
    f, ipiv, info = O.lu_c(np.zeros((0, 0), order="F"))
    assert ipiv.size == 0 and info == 0
    f, ipiv, info = O.lu_c(np.zeros((60, 60), order="F"))
    assert info == 1 and np.array_equal(ipiv, np.arange(1, 61))
    a = np.asfortranarray(np.random.default_rng(0).random((70, 70)))
    a[5, 0] = np.nan
    _, ipiv, _ = O.lu_c(a.copy(order="F"))
    assert ipiv[0] != 6


JULIA_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "julia_golden.npz")


@pytest.mark.skipif(not os.path.exists(JULIA_GOLDEN),
                    reason="tests/golden/julia_golden.npz absent: the Julia reference cannot run in this image "
                           "(generate it with tests/golden/make_golden_julia.jl where Julia exists)")
@pytest.mark.parametrize("idx", range(len(CASES)))
def test_oracle_matches_julia_reference(idx):
    """Pins the oracle against outputs of the real RecursiveFactorization.lu! (src/lu.jl:97-130) on the golden inputs:
    `F.ipiv` / `F.info` exact, `F.factors` within the reference's own bound (test/runtests.jl:19-20)."""
    g = np.load(JULIA_GOLDEN)
    name, m, n, dt, special = CASES[idx]
    a0 = make_input(idx)
    f, ipiv, info = O.lu_c(a0.copy(order="F"))
    for tag in ("serial", "threaded"):
        assert info == int(g[f"{name}.{tag}.info"][0])
        assert np.array_equal(ipiv, g[f"{name}.{tag}.ipiv"])
        if info == 0:
            assert np.abs(f - g[f"{name}.{tag}.factors"]).max() <= ref_bound(max(m, n), a0.dtype)
    fn, _, infon = O.lu_nopiv_c(a0.copy(order="F"))
    assert infon == int(g[f"{name}.nopiv.info"][0])
