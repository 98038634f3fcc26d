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
    swap(col0, ncols, k0, k1)        pivots [k0, k1) applied to columns [col0, col0+ncols), rows >= k0
    trsm(c0, n1, col0, ncols)        A[c0:c0+n1, cols] <- unitlower(A[c0:c0+n1, c0:c0+n1])^-1 A[c0:c0+n1, cols]
    gemm(c0, n1, col0, ncols)        A[c0+n1:, cols] -= A[c0+n1:, c0:c0+n1] A[c0:c0+n1, cols]
    """
    nblk = (n + nb - 1) // nb

    def leaf(b: int) -> None:
        c0, width = block_range(b, n, nb)
        root = owner_of(b, world)
        if rank == root:
            be.factor_block(c0, width)
        be.bcast_block(c0, width, root)

    def update(j: int, c0: int, n1: int) -> None:              # src/lu.jl:233-240 on one owned block column
        col0, ncols = block_range(j, n, nb)
        be.swap(col0, ncols, c0, c0 + n1)
        be.trsm(c0, n1, col0, ncols)
        be.gemm(c0, n1, col0, ncols)

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
            update(first, c0, n1)
            leaf(first)
            for j in mine:
                if j != first:
                    update(j, c0, n1)
        else:
            for j in mine:
                update(j, c0, n1)
            leaf(first)
        rec(first, nbk - nb1, True)                            # :244
        be.swap(c0, n1, c0 + n1, c0 + width)                   # :246 on the replicated L

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
        self.A = torch.zeros(self.n * self.n, dtype=tdt, device=self.device)          # column-major, lda = n
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
        self.my_blocks = owned_blocks(self.rank, self.world, self.n, self.nb)
        self.bcast_bytes = 0
        suf = "f64" if self.dtype == np.float64 else "f32"
        self._lu_range = getattr(self._lib, f"rfb_lu_range_{suf}")
        self._laswp_range = getattr(self._lib, f"rfb_laswp_range_{suf}")
        self._trsm = getattr(self._lib, f"rfb_trsm_llnu_{suf}")
        self._gemm = getattr(self._lib, f"rfb_gemm_nn_sub_{suf}")
        torch.cuda.synchronize()

    # -- data movement ----------------------------------------------------------------------------
    def _ptr(self, r: int, c: int) -> C.c_void_p:
        return C.c_void_p(self.A.data_ptr() + (r + c * self.n) * self.item)

    def set_block(self, j: int, host: np.ndarray) -> None:
        c0, w = block_range(j, self.n, self.nb)
        assert host.shape == (self.n, w) and host.dtype == self.dtype
        t = self.torch.from_numpy(np.ascontiguousarray(host.T))           # w x n, rows = columns of A
        with self.torch.cuda.stream(self.stream):
            self.A[c0 * self.n:(c0 + w) * self.n].copy_(t.reshape(-1), non_blocking=False)

    def get_block(self, j: int) -> np.ndarray:
        c0, w = block_range(j, self.n, self.nb)
        with self.torch.cuda.stream(self.stream):
            host = self.A[c0 * self.n:(c0 + w) * self.n].reshape(w, self.n).cpu()
        return np.asfortranarray(host.numpy().T)

    # -- backend interface used by run_schedule -----------------------------------------------------
    def factor_block(self, c0: int, w: int) -> None:
        self.ctx._check(self._lu_range(self._h, C.c_void_p(self.A.data_ptr()), self.n, self.n, c0, w,
                                       C.c_void_p(self.ipiv.data_ptr()), C.c_void_p(self.info_dev.data_ptr()),
                                       C.byref(self.opts)))

    def bcast_block(self, c0: int, w: int, root: int) -> None:
        rows = self.n - c0
        pbytes = rows * w * self.item
        st = self.stage.data_ptr()
        off_piv, off_dst, off_src, off_w = pbytes, pbytes + 8 * w, pbytes + 16 * w, pbytes + 24 * w
        total = pbytes + 28 * w
        lib, h = self._lib, self._h
        if self.rank == root:                                             # pack
            self.ctx._check(lib.rfb_copy2d(h, C.c_void_p(st), rows * self.item, self._ptr(c0, c0), self.n * self.item,
                                           rows * self.item, w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(st + off_piv), C.c_void_p(self.ipiv.data_ptr() + 8 * c0), 8 * w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(st + off_dst), C.c_void_p(self.p_dst.data_ptr() + 8 * c0), 8 * w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(st + off_src), C.c_void_p(self.p_src.data_ptr() + 8 * c0), 8 * w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(st + off_w), C.c_void_p(self.p_width.data_ptr() + 4 * c0), 4 * w))
        if self.world > 1:
            self.dist.broadcast(self.stage[:total], src=self.dist.get_global_rank(self.group, root) if self.group else root,
                                group=self.group)
            self.bcast_bytes += total
        if self.rank != root:                                             # unpack
            self.ctx._check(lib.rfb_copy2d(h, self._ptr(c0, c0), self.n * self.item, C.c_void_p(st), rows * self.item,
                                           rows * self.item, w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(self.ipiv.data_ptr() + 8 * c0), C.c_void_p(st + off_piv), 8 * w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(self.p_dst.data_ptr() + 8 * c0), C.c_void_p(st + off_dst), 8 * w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(self.p_src.data_ptr() + 8 * c0), C.c_void_p(st + off_src), 8 * w))
            self.ctx._check(lib.rfb_d2d(h, C.c_void_p(self.p_width.data_ptr() + 4 * c0), C.c_void_p(st + off_w), 4 * w))

    def swap(self, col0: int, ncols: int, k0: int, k1: int) -> None:
        self.ctx._check(self._laswp_range(self._h, C.c_void_p(self.A.data_ptr()), self.n, col0, ncols, k0, k1,
                                          C.c_void_p(self.ipiv.data_ptr()), 1))

    def trsm(self, c0: int, n1: int, col0: int, ncols: int) -> None:
        self.ctx._check(self._trsm(self._h, self._ptr(c0, c0), n1, self._ptr(c0, col0), ncols, self.n))

    def gemm(self, c0: int, n1: int, col0: int, ncols: int) -> None:
        m2 = self.n - c0 - n1
        self.ctx._check(self._gemm(self._h, self._ptr(c0 + n1, col0), self._ptr(c0 + n1, c0), self._ptr(c0, col0),
                                   m2, ncols, n1, self.n))

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
            buf = self.A[c0 * self.n:(c0 + w) * self.n]
            if root == dst:
                if self.rank == dst:
                    out[:, c0:c0 + w] = self.get_block(j)
            elif self.rank == root:
                self.dist.send(buf, dst=dst, group=self.group)
            elif self.rank == dst:
                tmp = self.torch.empty_like(buf)
                self.dist.recv(tmp, src=root, group=self.group)
                out[:, c0:c0 + w] = tmp.reshape(w, self.n).cpu().numpy().T
        return out
