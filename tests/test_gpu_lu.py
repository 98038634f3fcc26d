"""GPU parity tests, whole path: rfb200.lu / lu_ (C ABI rfb_lu_f64/f32 with host buffers) against the
CPU oracle and the reference's own acceptance test (test/runtests.jl:14-68)."""
import os
import sys

import numpy as np
import pytest
from scipy.linalg import lapack

import rfb200
from oracle import rf_oracle as O
from util import assert_pivots_match, assert_testlu, hutchinson_residual, rand_matrix

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import CASES, make_input  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "getrf_golden.npz"))
REF_SIZES = list(range(1, 11)) + [50, 130, 300]


def test_native_library_is_the_one_running(ctx):
    info = ctx.device_info()
    assert info["cc"][0] == 10 and info["sm_count"] >= 100
    before = ctx.launch_count()
    rfb200.lu(np.asfortranarray(np.random.default_rng(0).random((200, 200))), ctx=ctx)
    assert ctx.launch_count() > before                      # our kernels launched, not a fallback


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("s", REF_SIZES)
def test_reference_sweep(ctx, dtype, s):
    """test/runtests.jl:39-64: square, fat, tall; info equality; residual bound; singular column."""
    rng = np.random.default_rng([12, s, np.dtype(dtype).itemsize])
    for (m, n) in ((s, s), (s, s + 2), (s + 2, s)):
        a0 = rand_matrix(rng, m, n, dtype)
        _, want_p, want_info = O.lu_c(a0.copy(order="F"))
        F = rfb200.lu(a0, ctx=ctx)
        assert_testlu(a0, F.factors, F.ipiv, F.info, want_info)
        assert np.array_equal(F.ipiv, want_p)               # north_star: pivot indices bit-exact
        a1 = a0.copy(order="F")
        i = int(rng.integers(0, min(m, n)))
        a1[:, i] = 0
        _, want_p, want_info = O.lu_c(a1.copy(order="F"))
        F = rfb200.lu(a1, rfb200.RowMaximum(), check=False, ctx=ctx)
        assert F.info == want_info and F.info > 0
        assert np.array_equal(F.ipiv, want_p)
        with pytest.raises(rfb200.SingularException):
            rfb200.lu(a1, ctx=ctx)                           # check=true -> checknonsingular (src/lu.jl:128)


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_golden_lapack(ctx, idx):
    name, m, n, dt, special = CASES[idx]
    a0 = make_input(idx)
    F = rfb200.lu(a0, check=False, ctx=ctx)
    assert F.info == int(GOLDEN[name + "__info"])
    assert np.array_equal(F.ipiv, GOLDEN[name + "__ipiv"])
    tol = 50 * max(m, n) * np.finfo(a0.dtype).eps
    du = GOLDEN[name + "__diagu"]
    scale = max(1.0, float(np.abs(du).max()))
    assert np.allclose(np.diag(F.factors), du, rtol=0, atol=tol * scale)
    if name + "__lu" in GOLDEN.files:
        assert np.allclose(F.factors, GOLDEN[name + "__lu"], rtol=0, atol=tol * scale)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(64, 64), (65, 65), (127, 200), (200, 127), (500, 500), (1000, 1000), (1111, 1013),
                                   (2048, 2048), (3000, 1000), (1000, 3000)])
def test_shapes_match_oracle(ctx, dtype, shape):
    m, n = shape
    rng = np.random.default_rng([21, m, n])
    a0 = rand_matrix(rng, m, n, dtype)
    want_f, want_p, want_info = O.lu_c(a0.copy(order="F"), threads=8)
    F = rfb200.lu(a0, ctx=ctx)
    assert F.info == want_info == 0
    assert_pivots_match(a0, F.factors, F.ipiv, want_p, strict=(dtype == np.float64))
    assert_testlu(a0, F.factors, F.ipiv, F.info, 0, wide=True)
    if np.array_equal(F.ipiv, want_p):
        scale = max(1.0, float(np.abs(want_f).max()))
        assert np.allclose(F.factors, want_f, rtol=0, atol=200 * max(m, n) * np.finfo(dtype).eps * scale)


@pytest.mark.parametrize("leaf", [16, 32, 64])
def test_leaf_width_option(ctx, leaf):
    a0 = rand_matrix(np.random.default_rng(leaf), 700, 700, np.float64)
    _, want_p, _ = O.lu_c(a0.copy(order="F"))
    F = rfb200.lu(a0, ctx=ctx, leaf_width=leaf)
    assert np.array_equal(F.ipiv, want_p)
    assert_testlu(a0, F.factors, F.ipiv, F.info, 0)


def test_inplace_semantics_and_user_ipiv(ctx):
    """lu!(A, ipiv) (README.md:29): A is overwritten, the caller's ipiv is the one returned."""
    a0 = rand_matrix(np.random.default_rng(1), 333, 333, np.float64)
    a = a0.copy(order="F")
    ipiv = np.full(333, -7, dtype=np.int64)
    F = rfb200.lu_(a, ipiv, True, True, ctx=ctx)             # thread=Val(true) accepted and ignored
    assert F.factors is a and F.ipiv is ipiv
    assert ipiv.min() >= 1 and not np.array_equal(a, a0)
    assert_testlu(a0, a, ipiv, F.info, 0)
    b = a0.copy(order="F")
    rfb200.lu(b, ctx=ctx)
    assert np.array_equal(b, a0)                             # lu() factors a copy


def test_edge_cases(ctx):
    F = rfb200.lu(np.zeros((0, 0), order="F"), ctx=ctx)
    assert F.ipiv.size == 0 and F.info == 0
    F = rfb200.lu(np.zeros((0, 5), order="F"), ctx=ctx)
    assert F.ipiv.size == 0 and F.info == 0
    F = rfb200.lu(np.zeros((130, 130), order="F"), check=False, ctx=ctx)
    assert F.info == 1 and np.array_equal(F.ipiv, np.arange(1, 131))
    eye = np.asfortranarray(np.eye(257))
    F = rfb200.lu(eye, ctx=ctx)
    assert np.array_equal(F.factors, eye) and np.array_equal(F.ipiv, np.arange(1, 258))
    a = rand_matrix(np.random.default_rng(2), 300, 300, np.float64)
    a[17, 0] = np.nan                                        # NaN is never chosen as pivot (src/lu.jl:301)
    _, want_p, _ = O.lu_c(a.copy(order="F"))
    F = rfb200.lu(a, check=False, ctx=ctx)
    assert F.ipiv[0] == want_p[0] != 18


@pytest.mark.parametrize("n", [4096])
def test_baseline_config_4096(ctx, n):
    """BASELINE config "4096x4096 Float64 LU with partial pivoting on 1 B200"."""
    a0 = np.asfortranarray(np.random.default_rng(12).random((n, n)))
    F = rfb200.lu(a0, ctx=ctx)
    _, piv, info = lapack.dgetrf(a0)
    assert F.info == info == 0
    assert np.array_equal(F.ipiv, piv + 1)
    res = O.residual_fro_rel(a0, F.factors, F.ipiv)
    assert res <= 20 * n * np.finfo(np.float64).eps, res     # ||PA-LU||_F/||A||_F <= c n eps, c = 20
    _, want_p, _ = O.lu_c(a0.copy(order="F"), threads=8)
    assert np.array_equal(F.ipiv, want_p)


def test_baseline_config_f32_8192_properties(ctx):
    """BASELINE config "8192x8192 Float32": size-independent checks (oracle too slow at this size)."""
    n = 8192
    a0 = np.asfortranarray(np.random.default_rng(12).random((n, n), dtype=np.float32))
    F = rfb200.lu(a0, ctx=ctx)
    assert F.info == 0
    assert sorted(O.perm_from_ipiv(F.ipiv, n).tolist()) == list(range(n))
    assert np.all(np.abs(np.tril(F.factors, -1)) <= 1.0)     # partial pivoting => |L| <= 1
    res = hutchinson_residual(a0, F.factors, F.ipiv)
    assert res <= 20 * n * np.finfo(np.float32).eps, res
    # pivots against LAPACK sgetrf (same "first largest |a_ik|" rule): identical, or the first difference is a
    # proven near-tie (Float32 rounding noise of a different summation order)
    _, piv, info = lapack.sgetrf(a0.copy(order="F"), overwrite_a=True)
    assert info == 0
    assert_pivots_match(a0, F.factors, F.ipiv, piv + 1, strict=False)


def test_baseline_config_16384_properties(ctx):
    """BASELINE config "16384x16384 Float64": |L| <= 1, valid permutation, residual probe."""
    n = 16384
    a0 = np.asfortranarray(np.random.default_rng(12).random((n, n)))
    F = rfb200.lu(a0, ctx=ctx)
    assert F.info == 0
    assert sorted(O.perm_from_ipiv(F.ipiv, n).tolist()) == list(range(n))
    assert np.all(np.abs(np.tril(F.factors, -1)) <= 1.0)
    res = hutchinson_residual(a0, F.factors, F.ipiv)
    assert res <= 20 * n * np.finfo(np.float64).eps, res
    # north_star "pivot indices bit-exact": at the headline size against LAPACK dgetrf (the oracle, which agrees with
    # LAPACK on every committed fixture, needs minutes here)
    _, piv, info = lapack.dgetrf(a0.copy(order="F"), overwrite_a=True)
    assert info == 0
    assert np.array_equal(F.ipiv, piv + 1), f"first mismatch at step {int(np.argmax(F.ipiv != piv + 1))}"


@pytest.mark.parametrize("n", [300, 512, 2048, 4096])
def test_f32_tensor_core_mode(ctx, n):
    """Float32 LU with the trailing update on tcgen05 (f32_mode = TF32X3; the default from 4096 columns up).  Stated tolerance:
    the north_star bound ||PA-LU||_F/||A||_F <= 20*n*eps(Float32) (met with a ~500x margin), a residual
    within 6x of the exact-FP32 mode's (the tensor core accumulates with truncation), and -- inside the
    reference's own tested range n <= 300 -- the reference's inf-norm bound 20*n*eps as well."""
    a0 = np.asfortranarray(np.random.default_rng([12, n]).random((n, n), dtype=np.float32))
    F1 = rfb200.lu(a0, ctx=ctx, f32_mode=1)
    F0 = rfb200.lu(a0, ctx=ctx, f32_mode=2)
    assert F1.info == 0
    assert sorted(O.perm_from_ipiv(F1.ipiv, n).tolist()) == list(range(n))
    eps = float(np.finfo(np.float32).eps)
    r1, r0 = O.residual_fro_rel(a0, F1.factors, F1.ipiv), O.residual_fro_rel(a0, F0.factors, F0.ipiv)
    assert r1 <= 20 * n * eps, r1
    assert r1 <= 6 * r0, (r1, r0)
    if n <= 300:
        assert_testlu(a0, F1.factors, F1.ipiv, F1.info, 0)


def test_baseline_config_f32_8192_tensor_core(ctx):
    """BASELINE config "8192x8192 Float32 LU, bf16/TF32 tensor-core GEMM with FP32 accumulate"."""
    n = 8192
    a0 = np.asfortranarray(np.random.default_rng(12).random((n, n), dtype=np.float32))
    F = rfb200.lu(a0, ctx=ctx, f32_mode=1)
    assert F.info == 0
    assert np.all(np.abs(np.tril(F.factors, -1)) <= 1.0)
    res = hutchinson_residual(a0, F.factors, F.ipiv)
    assert res <= 20 * n * np.finfo(np.float32).eps, res
    _, piv, info = lapack.sgetrf(a0.copy(order="F"), overwrite_a=True)
    assert info == 0
    assert_pivots_match(a0, F.factors, F.ipiv, piv + 1, strict=False)


@pytest.mark.parametrize("dtype,shape", [(np.float64, (4096, 4096)), (np.float64, (3000, 3000)), (np.float64, (5000, 2500)),
                                         (np.float32, (4100, 4100)), (np.float64, (2048, 2600)), (np.float64, (1500, 1500))])
@pytest.mark.parametrize("mode", [2, 1, 0])
def test_pinned_host_matrix_early_download(ctx, dtype, shape, mode):
    """Page-locked caller matrix: the upload is pipelined, finished subtrees apply their interchanges to all
    columns on their left at once (instead of at the end of each node, src/lu.jl:246), and finished parts of the
    factors travel back to the host while the factorization still runs -- tile by tile (mode 2, the default: unit row
    bands + U12 blocks on a download stream of their own), as row bands of the right-spine nodes (mode 1), or not at
    all (mode 0).  Same swaps in the same per-column order: the result must be IDENTICAL to the pageable-path result
    (reference order), and every element must have arrived (the host matrix is poisoned first)."""
    m, n = shape
    a0 = rand_matrix(np.random.default_rng([77, m, n]), m, n, dtype)
    F_ref = rfb200.lu(a0, ctx=ctx)                       # pageable numpy memory
    a_pin = ctx.pinned_empty((m, n), dtype)
    ctx.set_early_download(mode)
    try:
        np.copyto(a_pin, a0)
        ipiv = np.empty(min(m, n), dtype=np.int64)
        F = rfb200.lu_(a_pin, ipiv, ctx=ctx)
        assert F.info == 0
        assert np.array_equal(F.ipiv, F_ref.ipiv)
        assert np.array_equal(np.asarray(F.factors), F_ref.factors)
        # and without pivoting (no interchanges to move, downloads only)
        G_ref = rfb200.lu(a0 + 10 * np.eye(m, n, dtype=dtype), False, ctx=ctx)
        np.copyto(a_pin, a0 + 10 * np.eye(m, n, dtype=dtype))
        G = rfb200.lu_(a_pin, None, False, ctx=ctx)
        assert np.array_equal(np.asarray(G.factors), G_ref.factors)
        # twice in a row on the same buffers: the second call's upload must not overtake the first call's downloads
        np.copyto(a_pin, a0)
        F2 = rfb200.lu_(a_pin, ipiv, ctx=ctx)
        assert np.array_equal(np.asarray(F2.factors), F_ref.factors)
    finally:
        ctx.set_early_download(2)


@pytest.mark.parametrize("shape", [(50000, 48), (90000, 24), (40000, 200)])
def test_very_tall_matrices_narrow_the_leaf(ctx, shape):
    """More rows than one 64-column cooperative panel grid can hold (148 x 256): the driver narrows the
    leaf to 32 / 16 columns (more co-resident CTAs) instead of failing; pivots still match LAPACK."""
    m, n = shape
    a0 = rand_matrix(np.random.default_rng([31, m, n]), m, n, np.float64)
    F = rfb200.lu(a0, ctx=ctx)
    _, piv, info = lapack.dgetrf(a0)
    assert F.info == info == 0
    assert np.array_equal(F.ipiv, piv + 1)
    assert_testlu(a0, F.factors, F.ipiv, F.info, 0, wide=True)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,nrhs", [(1, 1), (9, 1), (64, 3), (130, 1), (300, 7), (1000, 33), (2500, 512)])
def test_ldiv_solve(ctx, dtype, n, nrhs):
    """`ldiv!(F, B)` on the GPU (forward + back substitution with the K3 kernels): the solve check of
    test/runtests.jl:21-28 (`ldiv!(MF, A[:, end]) == e_n`, atol 100 E) and A X = B residuals."""
    rng = np.random.default_rng([9, n, nrhs])
    a0 = rand_matrix(rng, n, n, dtype)
    F = rfb200.lu(a0, ctx=ctx)
    eps = float(np.finfo(dtype).eps)
    x = F.solve(a0[:, -1].copy(), ctx=ctx)
    rhs = np.zeros(n); rhs[-1] = 1
    assert np.allclose(x, rhs, rtol=0, atol=100 * 20 * n * eps * max(1.0, np.linalg.cond(a0.astype(np.float64)) * eps * 10))
    b = np.asfortranarray(rng.random((n, nrhs), dtype=dtype))
    xs = rfb200.ldiv_(F, b.copy(order="F"), ctx=ctx)
    xw = np.linalg.solve(a0.astype(np.float64), b.astype(np.float64))
    r = np.abs(a0.astype(np.float64) @ xs.astype(np.float64) - b.astype(np.float64)).max()
    assert r <= 1000 * n * eps * max(1.0, np.abs(xw).max()), r        # the reference's solve bound (runtests.jl:82)


@pytest.mark.parametrize("dtype,shape", [(np.float64, (1500, 1500)), (np.float64, (2100, 900)), (np.float32, (1300, 1700)),
                                         (np.float64, (777, 777))])
def test_node_level_laswp_paths_agree(ctx, dtype, shape, monkeypatch):
    """K2 has two list-driven forms: per panel (laswp_list_kernel) and per node (compose the node's panels into the
    net permutation, one pass per column: laswp_compose_kernel + laswp_net_kernel).  The same factorization run with
    the node-level path forced everywhere AND chunked into tiny pivot ranges, with it disabled, and with the defaults
    must give IDENTICAL factors and pivots (row interchanges move bits, they do not compute), equal to the oracle's."""
    m, n = shape
    a0 = rand_matrix(np.random.default_rng([31, m, n]), m, n, dtype)
    F0 = rfb200.lu(a0, ctx=ctx)
    results = []
    for net_min, net_cap in ((64, 192), (1 << 30, 0), (64, 0)):
        monkeypatch.setenv("RFB_LASWP_NET_MIN", str(net_min))
        monkeypatch.setenv("RFB_LASWP_NET_CAP", str(net_cap))
        c2 = rfb200.Context(ctx.device)
        try:
            results.append(rfb200.lu(a0, ctx=c2))
        finally:
            c2.close()
    for Fi in results:
        assert Fi.info == F0.info == 0
        assert np.array_equal(Fi.ipiv, F0.ipiv)
        assert np.array_equal(Fi.factors, F0.factors)
    _, want_p, _ = O.lu_c(a0.copy(order="F"), threads=8)
    if dtype == np.float64:
        assert np.array_equal(F0.ipiv, want_p)
    else:
        assert_pivots_match(a0, F0.factors, F0.ipiv, want_p, strict=False)


@pytest.mark.parametrize("dtype,shape", [(np.float64, (1000, 1000)), (np.float64, (777, 1300)), (np.float64, (1500, 640)),
                                         (np.float32, (900, 900)), (np.float64, (130, 130))])
def test_host_driven_recursion_over_kernel_abi_matches_library_driver(ctx, dtype, shape):
    """north_star: "Julia host code drives the recursion and calls the kernels through a thin ccall shim".  The shim's
    `lu_device!` / `reckernel_device!` (julia/RecursiveFactorizationB200.jl, un-executable here) has an executable twin,
    rfb200.host_recursion.lu_device_, issuing the identical kernel-level calls through ctypes: square, fat (src/lu.jl:148-154)
    and tall shapes, both element types.  Same kernels in the same order as the C++ driver => bit-identical results."""
    from rfb200.host_recursion import lu_device_
    m, n = shape
    a0 = rand_matrix(np.random.default_rng([55, m, n]), m, n, dtype)
    d = rfb200.DeviceMatrix(ctx, m, n, dtype, lda=m + (m & 1))
    d.upload(a0)
    ctx.sync()
    lu_device_(ctx, d.ptr, m, n, d.lda, d.ipiv_ptr, d.info_ptr, dtype)
    f, ipiv, info = d.download()
    d.free()
    F = rfb200.lu(a0, check=False, ctx=ctx, f32_mode=2)      # kernel-level GEMM calls are exact FP32 (RFB_F32_AUTO there)
    assert info == F.info == 0
    assert np.array_equal(ipiv, F.ipiv)
    assert np.array_equal(f, F.factors)
