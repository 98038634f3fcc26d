"""CPU test of the REAL host driver (csrc/rfb_api.cu: lu_device / lu_rec, the twin of src/lu.jl:97-156 and :189-263).

`rfb_trace_lu` runs that C++ recursion without a GPU and returns the operations it would enqueue.  Replaying them with
the oracle's own kernels must reproduce the oracle's LU (same leaf width) BIT FOR BIT -- offsets, sizes and order are
then right -- for both interchange orders (reference order; eager order used for page-locked host matrices), for
pivot = Val(false), and for fat / tall shapes.  Every early row download must see rows that never change afterwards."""
import ctypes as C

import numpy as np
import pytest

import rfb200
from oracle import rf_oracle as O
from util import rand_matrix

PANEL, PANEL_NOPIV, LASWP, TRSM, GEMM, DOWNLOAD, IOTA = 1, 2, 3, 4, 5, 6, 7


def replay(a, ops, pivot=True):
    lib = O._load()
    suf = "f64" if a.dtype == np.float64 else "f32"
    m, n = a.shape
    lda, it, base = m, a.itemsize, a.ctypes.data
    at = lambda r, c: base + (r + c * lda) * it
    mn = min(m, n)
    ipiv = np.zeros(mn, dtype=np.int64)
    info = 0
    snapshots = []
    for op, r, c, s0, s1, s2, r2, c2 in ops.tolist():
        if op == PANEL:
            loc = np.zeros(s1, dtype=np.int64)
            k = getattr(lib, f"rfo_panel_{suf}")(at(r, c), s0, s1, lda, loc.ctypes.data, 0)
            ipiv[c:c + s1] = loc + r
            if k and info == 0:
                info = s2 + k
        elif op == PANEL_NOPIV:
            k = getattr(lib, f"rfo_panel_nopiv_{suf}")(at(r, c), s0, s1, lda, 0)
            if k and info == 0:
                info = -(s2 - k)
        elif op == LASWP:
            assert r == s1                                    # the block starts at the first pivot's row
            for i in range(s1, s2):
                ip = int(ipiv[i]) - 1
                if ip != i:
                    a[[i, ip], c:c + s0] = a[[ip, i], c:c + s0]
        elif op == TRSM:
            O.trsm_c(a, (r, c), s0, (r2, c2), s1)
        elif op == GEMM:
            O.schur_c(a, (r, c), (r2, c2), (c2, c), s0, s1, s2)
        elif op == DOWNLOAD:                                  # tile rows [r, r + s0) x columns [c, c + s1)
            assert 0 <= r and r + s0 <= m and 0 <= c and c + s1 <= n and s0 > 0 and s1 > 0
            snapshots.append((r, c, a[r:r + s0, c:c + s1].copy()))
        elif op == IOTA:
            ipiv[:] = np.arange(1, mn + 1)
        else:
            raise AssertionError(op)
    return a, ipiv, info, snapshots


CASES = [(1, 1), (5, 7), (64, 64), (65, 65), (130, 130), (300, 300), (300, 302), (302, 300), (500, 260), (260, 500),
         (777, 777), (1100, 1100)]


def check_early_downloads(snaps, got, mode):
    """Every early download must have seen elements that never change afterwards, and no element may travel twice.
    Mode 2 (tiles) covers the whole matrix; mode 1 (row bands) covers full-width bands contiguous from row 0."""
    m, n = got.shape
    seen = np.zeros((m, n), dtype=np.int32)
    for r, c, block in snaps:
        assert np.array_equal(block, got[r:r + block.shape[0], c:c + block.shape[1]], equal_nan=True), (r, c, block.shape)
        seen[r:r + block.shape[0], c:c + block.shape[1]] += 1
    assert seen.max(initial=0) <= 1, "an element was downloaded twice"
    if mode == 2 and snaps:
        assert seen.min() == 1, "tile mode must cover the whole matrix (nothing is downloaded at the end)"
    if mode == 1:
        covered = 0
        for r, c, block in snaps:
            assert r == covered and c == 0 and block.shape[1] == n
            covered += block.shape[0]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", CASES)
@pytest.mark.parametrize("pinned", [0, 1, 2])                  # pageable / page-locked with row bands / with tiles
def test_pivoted_schedule_replays_to_the_oracle_lu(dtype, shape, pinned):
    m, n = shape
    a0 = rand_matrix(np.random.default_rng([41, m, n]), m, n, dtype)
    if m > 200:
        a0[:, 150] = 0                                        # a zero pivot somewhere in the middle: info path
    ops = rfb200.trace_lu(m, n, dtype, pinned_host=bool(pinned), early_mode=pinned or None)
    got, ipiv, info, snaps = replay(a0.copy(order="F"), ops)
    want, wp, winfo = O.lu_c(a0.copy(order="F"), blocksize=64, threshold=1)
    assert info == winfo and np.array_equal(ipiv, wp)
    assert np.array_equal(got, want, equal_nan=True)
    if pinned and m >= n and (n >= 1024 or pinned == 2):
        assert snaps, "a page-locked square/tall matrix must send finished elements back early"
    check_early_downloads(snaps, got, pinned)
    if not pinned or m < n:
        assert not snaps


@pytest.mark.parametrize("shape", [(1536, 1536), (2100, 1700), (1024, 1024), (513, 513), (512, 512), (3000, 600)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_early_download_modes_tile_the_matrix(shape, mode):
    """Units of 512 columns: several units, ragged last unit, tall matrices (rows below the square part leave with the
    last unit), a single unit.  Mode 0 keeps the reference's interchange order and downloads nothing early."""
    m, n = shape
    a0 = rand_matrix(np.random.default_rng([44, m, n]), m, n, np.float64)
    ops = rfb200.trace_lu(m, n, np.float64, pinned_host=True, early_mode=mode)
    got, ipiv, info, snaps = replay(a0.copy(order="F"), ops)
    want, wp, winfo = O.lu_c(a0.copy(order="F"), blocksize=64, threshold=1)
    assert info == winfo == 0 and np.array_equal(ipiv, wp) and np.array_equal(got, want)
    check_early_downloads(snaps, got, mode)
    if mode == 0:
        assert not snaps and np.array_equal(ops, rfb200.trace_lu(m, n, np.float64))
    if mode == 2:
        assert sum(b.size for _, _, b in snaps) == m * n
        # what is still to be sent when the last panel finishes is the last unit's band only
        last_panel = int(np.nonzero(ops[:, 0] == PANEL)[0][-1])
        after = ops[last_panel + 1:]
        after = after[after[:, 0] == DOWNLOAD]
        assert len(after) == 1 and int(after[0, 3]) * int(after[0, 4]) <= (m - (n - 512 if n > 512 else 0)) * n


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(64, 64), (130, 130), (300, 302), (500, 260), (777, 777)])
def test_unpivoted_schedule_replays_to_the_oracle_lu(dtype, shape):
    m, n = shape
    a0 = rand_matrix(np.random.default_rng([42, m, n]), m, n, dtype)
    a0[np.arange(min(m, n)), np.arange(min(m, n))] += 10
    if m > 200:
        a0[170, 170] = 0; a0[170, :170] = 0; a0[:170, 170] = 0    # exactly-zero pivot 171 -> info = -171
    ops = rfb200.trace_lu(m, n, dtype, no_pivot=1)
    assert LASWP not in ops[:, 0] and PANEL not in ops[:, 0]
    got, _, info, _ = replay(a0.copy(order="F"), ops, pivot=False)
    want, _, winfo = O.lu_nopiv_c(a0.copy(order="F"), blocksize=64, threshold=1)
    assert info == winfo and (m <= 200 or info == -171)
    assert np.array_equal(got, want, equal_nan=True)


def test_schedule_shape_facts():
    """Launch counts of the 16384^2 factorization quoted in DESIGN.md / bench `launches_by_class`."""
    ops = rfb200.trace_lu(16384, 16384)
    kinds = ops[:, 0].tolist()
    assert kinds.count(PANEL) == 256 and kinds.count(LASWP) == 510 and kinds.count(GEMM) == 255 and kinds.count(TRSM) == 255
    eager = rfb200.trace_lu(16384, 16384, pinned_host=True)
    ek = eager[:, 0].tolist()
    assert ek.count(PANEL) == 256 and ek.count(GEMM) == 255
    down = eager[eager[:, 0] == DOWNLOAD]                      # tiles: 32 unit bands + 31 U12 blocks, the whole matrix once
    assert len(down) == 63 and int((down[:, 3] * down[:, 4]).sum()) == 16384 * 16384
    assert down[-1, 1:5].tolist() == [15872, 0, 512, 16384]   # still to be sent after the last panel: 512 x n
    bands = rfb200.trace_lu(16384, 16384, pinned_host=True, early_mode=1)
    down = bands[bands[:, 0] == DOWNLOAD]
    assert down[:, 1].tolist() == [0, 8192, 12288, 14336, 15360] and int(down[:, 3].sum()) == 15872   # tail: last 512 rows
    # Float32 splits at multiples of 16 (src/lu.jl:158-162)
    f32 = rfb200.trace_lu(8192, 8192, np.float32)
    assert (f32[f32[:, 0] == PANEL][:, 4] <= 64).all()


@pytest.mark.parametrize("opt", [dict(leaf_width=16), dict(leaf_width=32), dict(laswp_path=1)])
def test_driver_options_replay(opt):
    """leaf_width (the GPU analogue of the reference's `blocksize`, src/lu.jl:101) and the ipiv-driven laswp path."""
    m = n = 333
    a0 = rand_matrix(np.random.default_rng([43, m]), m, n, np.float64)
    ops = rfb200.trace_lu(m, n, np.float64, **opt)
    leaf = opt.get("leaf_width", 64)
    assert (ops[ops[:, 0] == PANEL][:, 4] <= leaf).all()
    got, ipiv, info, _ = replay(a0.copy(order="F"), ops)
    want, wp, winfo = O.lu_c(a0.copy(order="F"), blocksize=leaf, threshold=1)
    assert info == winfo == 0 and np.array_equal(ipiv, wp) and np.array_equal(got, want)


def test_very_tall_matrices_narrow_the_leaf_in_the_schedule():
    """One row per thread of a cooperative grid: beyond 148 x 256 rows the driver narrows the leaf."""
    assert (rfb200.trace_lu(30000, 128)[:, 4][rfb200.trace_lu(30000, 128)[:, 0] == PANEL] == 64).all()
    ops = rfb200.trace_lu(50000, 128)
    assert (ops[ops[:, 0] == PANEL][:, 4] <= 32).all()
