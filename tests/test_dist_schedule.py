"""CPU (gloo, world_size 2, 3 and 4) tests of the multi-GPU schedule.

The product's distributed recursion is the C++ scheduler of csrc/rfb_mg.cu.  `rfb_mg_trace` runs exactly that code
dry (no GPU, no NCCL) and returns the operations one rank enqueues; here every rank replays ITS trace with a numpy
backend whose kernels are the CPU oracle's loops and whose broadcast is torch.distributed/gloo, and the gathered
result must equal the single-process oracle factorization (pivots exactly, factors to rounding)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyBackend:
    """Replicated-L 1-D block-cyclic LU on numpy arrays (one per rank), gloo broadcasts."""

    def __init__(self, a_full, nb, rank, world):
        from oracle import rf_oracle as O
        self.O = O
        self.n, self.nb, self.rank, self.world = a_full.shape[0], nb, rank, world
        self.A = np.zeros_like(a_full, order="F")          # only own block columns are filled in
        from rfb200.dist_lu import owner_of
        for j in range((self.n + nb - 1) // nb):
            if owner_of(j, world) == rank:
                self.A[:, j * nb:(j + 1) * nb] = a_full[:, j * nb:(j + 1) * nb]
        self.ipiv = np.zeros(self.n, dtype=np.int64)
        self.info = 0
        self.bcasts = 0

    def factor_block(self, c0, w):
        sub = np.asfortranarray(self.A[c0:, c0:c0 + w])
        f, p, info = self.O.lu_c(sub)
        self.A[c0:, c0:c0 + w] = f
        self.ipiv[c0:c0 + w] = p + c0
        if info and not self.info:
            self.info = info + c0

    def bcast_block(self, c0, w, root):
        panel = torch.from_numpy(np.ascontiguousarray(self.A[c0:, c0:c0 + w]))
        piv = torch.from_numpy(self.ipiv[c0:c0 + w].copy())
        dist.broadcast(panel, src=root)
        dist.broadcast(piv, src=root)
        self.bcasts += 1
        if self.rank != root:
            self.A[c0:, c0:c0 + w] = panel.numpy()
            self.ipiv[c0:c0 + w] = piv.numpy()

    def _swap(self, col0, ncols, k0, k1):
        blk = self.A[:, col0:col0 + ncols]
        for i in range(k0, k1):
            r = int(self.ipiv[i]) - 1
            if r != i:
                blk[[i, r], :] = blk[[r, i], :]

    def update(self, blocks, c0, n1):
        import scipy.linalg as sl
        assert blocks == sorted(blocks)
        for j in blocks:
            col0 = j * self.nb
            ncols = min(self.n, col0 + self.nb) - col0
            self._swap(col0, ncols, c0, c0 + n1)
            l = self.A[c0:c0 + n1, c0:c0 + n1]
            self.A[c0:c0 + n1, col0:col0 + ncols] = sl.solve_triangular(l, self.A[c0:c0 + n1, col0:col0 + ncols], lower=True,
                                                                         unit_diagonal=True)
            self.A[c0 + n1:, col0:col0 + ncols] -= self.A[c0 + n1:, c0:c0 + n1] @ self.A[c0:c0 + n1, col0:col0 + ncols]

    def swap_left(self, c0, n1, k0, k1):
        self._swap(c0, n1, k0, k1)


def _worker(rank, world, port, n, nb, seed, zero_col, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rfb200
    from rfb200.dist_lu import (TRACE_BCAST, TRACE_FACTOR, TRACE_SWAP_LEFT, TRACE_UPDATE, block_range, owner_of,
                                trace_schedule)
    a = np.asfortranarray(np.random.default_rng(seed).random((n, n)))
    if zero_col >= 0:
        a[:, zero_col] = 0
    be = NumpyBackend(a, nb, rank, world)
    for code, x0, x1, x2, x3 in trace_schedule(n, nb, rank, world).tolist():
        if code == TRACE_UPDATE:                      # [1, c0, n1, j, 0]
            assert owner_of(x2, world) == rank
            be.update([x2], x0, x1)
        elif code == TRACE_FACTOR:                    # [2, j, c0, w, 0]
            assert owner_of(x0, world) == rank and (x1, x2) == block_range(x0, n, nb)
            be.factor_block(x1, x2)
        elif code == TRACE_BCAST:                     # [3, j, root, c0, w]
            assert x1 == owner_of(x0, world)
            be.bcast_block(x2, x3, x1)
        elif code == TRACE_SWAP_LEFT:                 # [4, c0, n1, k0, k1]
            be.swap_left(x0, x1, x2, x3)
        else:
            raise AssertionError(code)
    # gather: every rank contributes its own block columns
    full = torch.zeros((n, n), dtype=torch.float64)
    for j in range((n + nb - 1) // nb):
        if owner_of(j, world) == rank:
            c0, w = block_range(j, n, nb)
            full[:, c0:c0 + w] = torch.from_numpy(np.ascontiguousarray(be.A[:, c0:c0 + w]))
    dist.all_reduce(full)
    info = torch.tensor([be.info if be.info else 1 << 60])
    dist.all_reduce(info, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), f=full.numpy(), ipiv=be.ipiv, info=int(info) if int(info) < (1 << 60) else 0,
                 bcasts=be.bcasts)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n,nb,zero_col", [(2, 256, 64, -1), (2, 300, 64, -1), (3, 448, 64, -1), (2, 256, 64, 100),
                                                 (4, 704, 64, -1)])
def test_block_cyclic_schedule_matches_oracle(tmp_path, world, n, nb, zero_col):
    sys.path.insert(0, ROOT)
    from oracle import rf_oracle as O
    mp.spawn(_worker, args=(world, _free_port(), n, nb, 5, zero_col, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "out.npz"))
    a = np.asfortranarray(np.random.default_rng(5).random((n, n)))
    if zero_col >= 0:
        a[:, zero_col] = 0
    want_f, want_p, want_info = O.lu_c(a.copy(order="F"))
    assert int(got["info"]) == want_info
    assert np.array_equal(got["ipiv"], want_p)                   # bit-exact pivots
    assert int(got["bcasts"]) == (n + nb - 1) // nb              # one broadcast per block column
    if want_info == 0:
        assert np.allclose(got["f"], want_f, rtol=0, atol=20 * n * np.finfo(np.float64).eps * max(1.0, np.abs(want_f).max()))
        assert O.residual_inf(a, np.asfortranarray(got["f"]), got["ipiv"]) < 20 * n * np.finfo(np.float64).eps


def test_ownership_helpers():
    sys.path.insert(0, ROOT)
    import rfb200
    from rfb200.dist_lu import block_range, owned_blocks, owner_of
    assert [owner_of(j, 4) for j in range(18)] == [0, 0, 1, 1, 2, 2, 3, 3, 3, 3, 2, 2, 1, 1, 0, 0, 0, 0]
    assert owned_blocks(1, 4, 1000, 128) == [2, 3]
    for world in (1, 2, 3, 8):
        counts = [len(owned_blocks(r, world, 64 * 512, 512)) for r in range(world)]
        assert max(counts) - min(counts) <= 2 and sum(counts) == 64
    assert block_range(7, 1000, 128) == (896, 104)
    assert sum(block_range(j, 1000, 128)[1] for j in range(8)) == 1000


def test_trace_structure():
    """Structure of the C++ schedule (rfb_mg_trace): every block column is broadcast exactly once and in block order on
    every rank; an owned block column receives the contribution of EVERY block column on its left, in order (the last one
    right before it is factored -- the look-ahead step), then is factored, then published; the pivots of every block
    column are applied once to the rank's finished columns on the left (src/lu.jl:246)."""
    sys.path.insert(0, ROOT)
    import rfb200  # noqa: F401
    from rfb200.dist_lu import TRACE_BCAST, TRACE_FACTOR, TRACE_SWAP_LEFT, TRACE_UPDATE, trace_schedule
    n, nb, world = 64 * 13 - 20, 64, 4
    nblk = 13
    for rank in range(world):
        t = trace_schedule(n, nb, rank, world).tolist()
        assert [op[1] for op in t if op[0] == TRACE_BCAST] == list(range(nblk))
        from rfb200.dist_lu import owner_of
        mine = [j for j in range(nblk) if owner_of(j, world) == rank]
        assert [op[1] for op in t if op[0] == TRACE_FACTOR] == mine
        for j in mine:
            srcs = [op[1] // nb for op in t if op[0] == TRACE_UPDATE and op[3] == j]
            assert srcs == list(range(j))                                         # every block column on the left, in order
            i_f = t.index([TRACE_FACTOR, j, j * nb, min(nb, n - j * nb), 0])
            assert all(i < i_f for i, op in enumerate(t) if op[0] == TRACE_UPDATE and op[3] == j)
            for b in range(j):                                                    # a contribution needs its panel
                i_b = next(i for i, op in enumerate(t) if op[0] == TRACE_BCAST and op[1] == b)
                i_u = next(i for i, op in enumerate(t) if op[0] == TRACE_UPDATE and op[3] == j and op[1] == b * nb)
                assert i_b < i_u
        sw = [(op[3], op[4]) for op in t if op[0] == TRACE_SWAP_LEFT]
        assert sw == [(b * nb, min(n, (b + 1) * nb)) for b in range(nblk)]         # pivots of every block column, in order
