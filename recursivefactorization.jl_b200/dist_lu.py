"""Multi-GPU recursive LU: 1-D block-cyclic columns, one process per GPU, NCCL panel broadcast.

BASELINE.json north_star / SURVEY.md section 8e.  The reference has no distributed path at all; this
is the same Toledo recursion (src/lu.jl:189-263) run over *block columns*:

* block column J (width ``block``) is owned by rank ``J % world``;
* a recursion node whose range is one block column is factored by its owner with the single-GPU
  path (``rfb_lu_range``), then the factored panel (rows below its diagonal block included), its
  pivots and its row-exchange lists are broadcast from the owner (``torch.distributed.broadcast``,
  i.e. ``ncclBroadcast`` over NVLink on GPUs, gloo in the CPU tests);
* every rank keeps a full-size column-major buffer in which its own columns and all received L
  panels are valid ("replicated L"), so steps 2-4 of the recursion (row swaps, TRSM, Schur update)
  touch only columns the rank owns and need no communication;
* step 6 (``A21 <- P2 A21``) is applied to the replica on every rank.

The schedule (`run_schedule`) is pure host logic over a small backend interface, so the CPU tests
drive exactly the same code with a numpy/oracle backend over gloo.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from . import _lib


# ----------------------------------------------------------------------------------------------
# host logic shared by every backend
# ----------------------------------------------------------------------------------------------
def owner_of(block: int, world: int) -> int:
    return block % world


def block_range(j: int, n: int, nb: int) -> Tuple[int, int]:
    """(first column, width) of block column j."""
    c0 = j * nb
    return c0, min(n, c0 + nb) - c0


def owned_blocks(rank: int, world: int, n: int, nb: int) -> List[int]:
    return [j for j in range((n + nb - 1) // nb) if owner_of(j, world) == rank]


def run_schedule(be, n: int, nb: int, rank: int, world: int) -> None:
    """Toledo recursion over block columns [0, ceil(n / nb)).  `be` implements:

    factor_block(c0, w)              owner only: LU of columns [c0, c0+w), rows c0.. (global pivots)
    bcast_block(c0, w, root)         everyone: panel rows c0.., its pivots (and exchange lists)
    update(blocks, c0, n1)           src/lu.jl:233-240 on the OWNED block columns `blocks` (ascending, hence
                                     contiguous in the rank's local storage): row swaps with pivots [c0, c0+n1),
                                     A12 <- L11^-1 A12, A22 -= L21 A12, with L11/L21 = columns [c0, c0+n1) of the replica
    swap_left(c0, n1, k0, k1)        src/lu.jl:246: pivots [k0, k1) applied to rows >= k0 of columns [c0, c0+n1)
                                     (the replicated L and the rank's own copy of those columns)
    """
    nblk = (n + nb - 1) // nb

    def leaf(b: int) -> None:
        c0, width = block_range(b, n, nb)
        root = owner_of(b, world)
        if rank == root:
            be.factor_block(c0, width)
        be.bcast_block(c0, width, root)

    def rec(b0: int, nbk: int, first_done: bool) -> None:
        """`first_done`: the leftmost block column of this range was already factored (look-ahead)."""
        if nbk == 1:
            if not first_done:
                leaf(b0)
            return
        c0 = b0 * nb
        width = min(n, (b0 + nbk) * nb) - c0
        nb1 = (nbk + 1) // 2
        n1 = nb1 * nb
        rec(b0, nb1, first_done)                               # src/lu.jl:229
        # Look-ahead: the first block column of the right half is the next one on the critical path.
        # Its owner updates it first and factors + broadcasts it at once, and only then updates its other
        # columns; the other ranks do all their updates of this node while that factorization runs.
        first = b0 + nb1
        mine = [j for j in range(first, b0 + nbk) if owner_of(j, world) == rank]
        if owner_of(first, world) == rank:
            be.update([first], c0, n1)
            leaf(first)
            rest = [j for j in mine if j != first]
            if rest:
                be.update(rest, c0, n1)
        else:
            if mine:
                be.update(mine, c0, n1)
            leaf(first)
        rec(first, nbk - nb1, True)                            # :244
        be.swap_left(c0, n1, c0 + n1, c0 + width)              # :246

    if nblk > 0:
        rec(0, nblk, False)


# ----------------------------------------------------------------------------------------------
# GPU backend
# ----------------------------------------------------------------------------------------------
class DistributedLU:
    """One rank's share of a distributed n x n LU (Float64 / Float32) on its GPU.

    Usage (every rank, under torchrun):
        d = DistributedLU(n, np.float64, block=512)        # uses torch.distributed's default group
        d.set_block(j, host_array_n_by_w)   for j in d.my_blocks
        d.factor(); d.synchronize()                         # enqueued on d.stream
        info = d.info()                                     # global (all-reduced)
        d.get_block(j) / d.gather_to(0)
    """

    def __init__(self, n: int, dtype=np.float64, block: int = 512, ctx=None, group=None, leaf_width: int = 0):
        import torch
        import torch.distributed as dist
        from . import Context, _make_opts
        self.torch, self.dist = torch, dist
        self.n, self.nb, self.dtype = int(n), int(block), np.dtype(dtype)
        if self.dtype not in (np.float64, np.float32):
            raise TypeError("DistributedLU supports float64 / float32")
        if self.nb < 64 or self.nb % 64:
            raise ValueError("block must be a positive multiple of 64")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        # make sure the (eagerly, asynchronously initialised) NCCL communicator is fully up before this
        # process touches the device through a second CUDA runtime (librfb200 links cudart statically)
        if self.world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize()
        self.ctx = ctx or Context(torch.cuda.current_device())
        self._lib, self._h = self.ctx._lib, self.ctx.handle
        # one dedicated stream carries the kernels, the pack/unpack copies AND the NCCL broadcasts, in order
        self.stream = torch.cuda.Stream(device=self.device)
        self.ctx._check(self._lib.rfb_set_stream(self._h, C.c_void_p(self.stream.cuda_stream)))
        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        self.item = self.dtype.itemsize
        self.my_blocks = owned_blocks(self.rank, self.world, self.n, self.nb)
        self.lcol = {}                                  # block -> first local column (own blocks are stored compactly)
        ncl = 0
        for j in self.my_blocks:
            self.lcol[j] = ncl
            ncl += block_range(j, self.n, self.nb)[1]
        self.ncols_local = ncl
        # two column-major buffers with the SAME leading dimension n, so one kernel call can mix them:
        #   L   : full-size replica, valid where L panels were received (rows >= diagonal block)
        #   A   : this rank's own block columns, compact, in block order
        self.L = torch.zeros(self.n * self.n, dtype=tdt, device=self.device)
        self.A = torch.zeros(self.n * max(ncl, 1), dtype=tdt, device=self.device)
        self.ipiv = torch.zeros(self.n, dtype=torch.int64, device=self.device)
        self.info_dev = torch.zeros(8, dtype=torch.int64, device=self.device)
        self.p_dst = torch.empty(2 * self.n + 128, dtype=torch.int32, device=self.device)
        self.p_src = torch.empty(2 * self.n + 128, dtype=torch.int32, device=self.device)
        self.p_width = torch.empty(self.n + 64, dtype=torch.int32, device=self.device)
        self.ctx._check(self._lib.rfb_perm_buffers(self._h, C.c_void_p(self.p_dst.data_ptr()), C.c_void_p(self.p_src.data_ptr()),
                                                   C.c_void_p(self.p_width.data_ptr()), self.n + 64))
        meta = self.nb * (8 + 8 + 8 + 4)
        self.stage = torch.empty(self.n * self.nb * self.item + meta + 256, dtype=torch.uint8, device=self.device)
        self.opts = _make_opts(_lib.RFB_MEM_DEVICE, leaf_width)
        self.bcast_bytes = 0
        suf = "f64" if self.dtype == np.float64 else "f32"
        self._lu_range = getattr(self._lib, f"rfb_lu_range_{suf}")
        self._laswp_range = getattr(self._lib, f"rfb_laswp_range_{suf}")
        self._trsm = getattr(self._lib, f"rfb_trsm_llnu_{suf}")
        self._gemm = getattr(self._lib, f"rfb_gemm_nn_sub_{suf}")
        torch.cuda.synchronize()

    # -- data movement ----------------------------------------------------------------------------
    def _pa(self, r: int, lc: int) -> C.c_void_p:        # own storage, local column lc
        return C.c_void_p(self.A.data_ptr() + (r + lc * self.n) * self.item)

    def _pl(self, r: int, c: int) -> C.c_void_p:         # replica, global column c
        return C.c_void_p(self.L.data_ptr() + (r + c * self.n) * self.item)

    def block_slice(self, j: int):
        lc, w = self.lcol[j], block_range(j, self.n, self.nb)[1]
        return self.A[lc * self.n:(lc + w) * self.n]

    def set_block(self, j: int, host: np.ndarray) -> None:
        c0, w = block_range(j, self.n, self.nb)
        assert host.shape == (self.n, w) and host.dtype == self.dtype
        t = self.torch.from_numpy(np.ascontiguousarray(host.T))           # w x n, rows = columns of A
        with self.torch.cuda.stream(self.stream):
            self.block_slice(j).copy_(t.reshape(-1), non_blocking=False)

    def get_block(self, j: int) -> np.ndarray:
        w = block_range(j, self.n, self.nb)[1]
        with self.torch.cuda.stream(self.stream):
            host = self.block_slice(j).reshape(w, self.n).cpu()
        return np.asfortranarray(host.numpy().T)

    # -- backend interface used by run_schedule -----------------------------------------------------
    def factor_block(self, c0: int, w: int) -> None:
        # the block lives at local column lc: shift the base so that (row c0, col c0) of the "root" view is it
        lc = self.lcol[c0 // self.nb]
        root = self.A.data_ptr() + (lc - c0) * self.n * self.item
        self.ctx._check(self._lu_range(self._h, C.c_void_p(root), self.n, self.n, c0, w,
                                       C.c_void_p(self.ipiv.data_ptr()), C.c_void_p(self.info_dev.data_ptr()),
                                       C.byref(self.opts)))

    def bcast_block(self, c0: int, w: int, root: int) -> None:
        rows = self.n - c0
        pbytes = rows * w * self.item
        st = self.stage.data_ptr()
        off_piv, off_dst, off_src, off_w = pbytes, pbytes + 8 * w, pbytes + 16 * w, pbytes + 24 * w
        total = pbytes + 28 * w
        lib, h = self._lib, self._h
        if self.rank == root:                                             # pack from the own storage
            lc = self.lcol[c0 // self.nb]
            self.ctx._check(lib.rfb_copy2d(h, C.c_void_p(st), rows * self.item, self._pa(c0, lc), self.n * self.item,
                                           rows * self.item, w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(st + off_piv), C.c_void_p(self.ipiv.data_ptr() + 8 * c0), 8 * w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(st + off_dst), C.c_void_p(self.p_dst.data_ptr() + 8 * c0), 8 * w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(st + off_src), C.c_void_p(self.p_src.data_ptr() + 8 * c0), 8 * w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(st + off_w), C.c_void_p(self.p_width.data_ptr() + 4 * c0), 4 * w))
        if self.world > 1:
            self.dist.broadcast(self.stage[:total], src=self.dist.get_global_rank(self.group, root) if self.group else root,
                                group=self.group)
            self.bcast_bytes += total
        # every rank (the owner too) unpacks the panel into its replica
        self.ctx._check(lib.rfb_copy2d(h, self._pl(c0, c0), self.n * self.item, C.c_void_p(st), rows * self.item,
                                       rows * self.item, w))
        if self.rank != root:
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(self.ipiv.data_ptr() + 8 * c0), C.c_void_p(st + off_piv), 8 * w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(self.p_dst.data_ptr() + 8 * c0), C.c_void_p(st + off_dst), 8 * w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(self.p_src.data_ptr() + 8 * c0), C.c_void_p(st + off_src), 8 * w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(self.p_width.data_ptr() + 4 * c0), C.c_void_p(st + off_w), 4 * w))

    def _local_range(self, blocks):
        lc0 = self.lcol[blocks[0]]
        ncols = sum(block_range(j, self.n, self.nb)[1] for j in blocks)
        return lc0, ncols

    def update(self, blocks, c0: int, n1: int) -> None:
        lc0, ncols = self._local_range(blocks)
        base = C.c_void_p(self.A.data_ptr())
        # row swaps on the own columns (a "root" view whose column index is the local one)
        self.ctx._check(self._laswp_range(self._h, base, self.n, lc0, ncols, c0, c0 + n1, C.c_void_p(self.ipiv.data_ptr()), 1))
        self.ctx._check(self._trsm(self._h, self._pl(c0, c0), n1, self._pa(c0, lc0), ncols, self.n))
        m2 = self.n - c0 - n1
        self.ctx._check(self._gemm(self._h, self._pa(c0 + n1, lc0), self._pl(c0 + n1, c0), self._pa(c0, lc0), m2, ncols, n1, self.n))

    def swap_left(self, c0: int, n1: int, k0: int, k1: int) -> None:
        piv = C.c_void_p(self.ipiv.data_ptr())
        self.ctx._check(self._laswp_range(self._h, C.c_void_p(self.L.data_ptr()), self.n, c0, n1, k0, k1, piv, 1))
        mine = [j for j in self.my_blocks if c0 <= j * self.nb < c0 + n1]
        if mine:
            lc0, ncols = self._local_range(mine)
            self.ctx._check(self._laswp_range(self._h, C.c_void_p(self.A.data_ptr()), self.n, lc0, ncols, k0, k1, piv, 1))

    # -- driver -------------------------------------------------------------------------------------
    def factor(self) -> None:
        """Enqueue the whole distributed factorization on `self.stream` (no host synchronisation)."""
        with self.torch.cuda.stream(self.stream):
            self.info_dev.zero_()
            self.p_dst.fill_(-1)
            self.p_src.fill_(-1)
            self.p_width.zero_()
            run_schedule(self, self.n, self.nb, self.rank, self.world)

    def synchronize(self) -> None:
        self.stream.synchronize()
        self.ctx.sync()

    def info(self) -> int:
        """Global info: smallest non-zero per-rank value (first zero-pivot column), else 0."""
        big = self.torch.iinfo(self.torch.int64).max
        with self.torch.cuda.stream(self.stream):
            t = self.info_dev[:1].clone()
            t = self.torch.where(t == 0, self.torch.full_like(t, big), t)
            if self.world > 1:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
            v = int(t.item())
        return 0 if v == big else v

    def pivots(self) -> np.ndarray:
        with self.torch.cuda.stream(self.stream):
            return self.ipiv.cpu().numpy()

    def gather_to(self, dst: int = 0) -> Optional[np.ndarray]:
        """Assemble the factored matrix on rank `dst` (tests / small sizes only)."""
        out = np.empty((self.n, self.n), dtype=self.dtype, order="F") if self.rank == dst else None
        nblk = (self.n + self.nb - 1) // self.nb
        self.synchronize()
        for j in range(nblk):
            c0, w = block_range(j, self.n, self.nb)
            root = owner_of(j, self.world)
            buf = self.block_slice(j) if self.rank == root else None
            if root == dst:
                if self.rank == dst:
                    out[:, c0:c0 + w] = self.get_block(j)
            elif self.rank == root:
                self.dist.send(buf, dst=dst, group=self.group)
            elif self.rank == dst:
                tmp = self.torch.empty(self.n * w, dtype=self.A.dtype, device=self.device)
                self.dist.recv(tmp, src=root, group=self.group)
                out[:, c0:c0 + w] = tmp.reshape(w, self.n).cpu().numpy().T
        return out
