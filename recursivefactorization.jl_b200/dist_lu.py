"""Multi-GPU recursive LU: thin ctypes callers of the C++ driver (csrc/rfb_mg.cu, C ABI `rfb_mg_*`).

BASELINE.json north_star / SURVEY.md section 8e.  The reference has no distributed path at all; the driver runs the
same Toledo recursion (src/lu.jl:189-263) over *block columns* distributed 1-D block-cyclic, broadcasts each factored
block column + its pivots from its owner with ``ncclBroadcast`` and keeps a replica of L on every rank.  Everything on
the data path -- schedule, streams, events, NCCL calls -- is C++ behind the C ABI; nothing here touches torch.

Two handles (see include/rfb200.h):

* ``MultiGpuLU(ngpus)``       one process, G devices: ``lu_(A)`` factors a host matrix in place like ``rfb200.lu_``;
* ``DistributedLU(n, ...)``   one process per GPU (torchrun): the launcher's own transport carries the 128-byte NCCL
  id from rank 0 to the others once (``exchange_id``, e.g. a ``torch.distributed`` / MPI broadcast of a byte string);
  after that every call is librfb200's.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import _lib


# ----------------------------------------------------------------------------------------------
# host logic mirrored from the C++ plan (ownership only; the schedule itself is `rfb_mg_trace`)
# ----------------------------------------------------------------------------------------------
def owner_of(block: int, world: int) -> int:
    """Rank that owns block column `block`: pairs of block columns dealt out cyclically with alternating direction
    (`rfb_mg_owner_of`, the one definition: MgPlan::owner in csrc/rfb_mg.cu)."""
    return int(_lib.load().rfb_mg_owner_of(block, world))


def block_range(j: int, n: int, nb: int) -> Tuple[int, int]:
    """(first column, width) of block column j."""
    c0 = j * nb
    return c0, min(n, c0 + nb) - c0


def owned_blocks(rank: int, world: int, n: int, nb: int) -> List[int]:
    return [j for j in range((n + nb - 1) // nb) if owner_of(j, world) == rank]


TRACE_UPDATE, TRACE_FACTOR, TRACE_BCAST, TRACE_SWAP_LEFT = 1, 2, 3, 4


def trace_schedule(n: int, nb: int, rank: int, world: int) -> np.ndarray:
    """The operations rank `rank` of `world` enqueues for an n x n matrix (``rfb_mg_trace``: the C++ scheduler run dry,
    no GPU, no NCCL): an (nops, 5) int64 array, see csrc/rfb_mg.cu."""
    lib = _lib.load()
    count = C.c_int64(0)
    rc = lib.rfb_mg_trace(n, nb, rank, world, None, 0, C.byref(count))
    if rc != _lib.RFB_OK:
        raise RuntimeError(f"rfb_mg_trace failed ({rc})")
    ops = np.zeros((count.value, 5), dtype=np.int64)
    if count.value:
        rc = lib.rfb_mg_trace(n, nb, rank, world, C.c_void_p(ops.ctypes.data), count.value, C.byref(count))
        if rc != _lib.RFB_OK:
            raise RuntimeError(f"rfb_mg_trace failed ({rc})")
    return ops


class _MgHandle:
    def __init__(self):
        self._lib = _lib.load()
        self._h = C.c_void_p()

    def _check(self, rc: int):
        if rc != _lib.RFB_OK:
            from . import RfbError
            raise RfbError(rc, self._lib.rfb_mg_last_error(self._h).decode() if self._h else "rfb_mg: null handle")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rfb_mg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self) -> float:
        """Wait for the factorization; returns its device time in ms (max over this process's ranks)."""
        ms = C.c_float()
        self._check(self._lib.rfb_mg_sync(self._h, C.byref(ms)))
        return float(ms.value)

    def sched_stats(self, lr: int = 0) -> dict:
        out = (C.c_int64 * 8)()
        self._check(self._lib.rfb_mg_sched_stats(self._h, lr, out))
        return {"bulk_slices": out[0], "critical_enqueues": out[1], "host_idle_ms": out[2] / 1e3, "schedule_loop_ms": out[3] / 1e3,
                "critical_sections_device_ms": out[5] / 1e3, "publications_device_ms": out[6] / 1e3}

    def stats(self) -> dict:
        b, l = C.c_int64(), C.c_int64()
        self._check(self._lib.rfb_mg_stats(self._h, C.byref(b), C.byref(l)))
        return {"bcast_bytes_per_rank": b.value, "launches": l.value}


class MultiGpuLU(_MgHandle):
    """One process driving `ngpus` devices (``rfb_mg_create_all``): the entry a single-process caller of ``lu!`` uses."""

    def __init__(self, ngpus: int, devices=None):
        super().__init__()
        dev = (C.c_int * ngpus)(*devices) if devices is not None else None
        rc = self._lib.rfb_mg_create_all(C.byref(self._h), ngpus, dev)
        if rc != _lib.RFB_OK:
            msg = self._lib.rfb_mg_last_error(self._h).decode() if self._h else "rfb_mg_create_all failed"
            self.close()
            from . import RfbError
            raise RfbError(rc, msg)
        self.ngpus = ngpus

    def lu_(self, A: np.ndarray, ipiv: Optional[np.ndarray] = None, *, check: bool = True, block: int = 0):
        """``lu!(A, ipiv)`` (src/lu.jl:97-130) on all GPUs of this handle: A (square, column-major, float64/float32) is
        overwritten with L\\U; returns the same ``LU`` object as ``rfb200.lu_``."""
        from . import LU, _checknonsingular, _column_major_lda
        if not isinstance(A, np.ndarray) or A.ndim != 2 or A.shape[0] != A.shape[1] or A.dtype not in (np.float64, np.float32):
            raise TypeError("A must be a square float64/float32 numpy matrix")
        lda = _column_major_lda(A)
        if lda is None or not A.flags.writeable:
            raise TypeError("A must be a writeable column-major array")
        n = A.shape[0]
        if ipiv is None:
            ipiv = np.empty(n, dtype=np.int64)
        elif ipiv.dtype != np.int64 or ipiv.shape != (n,) or not ipiv.flags.c_contiguous:
            raise TypeError("ipiv must be a contiguous int64 vector of length n")
        info = C.c_int64(0)
        fn = self._lib.rfb_mg_lu_f64 if A.dtype == np.float64 else self._lib.rfb_mg_lu_f32
        self._check(fn(self._h, C.c_void_p(A.ctypes.data), n, lda, C.c_void_p(ipiv.ctypes.data), C.byref(info), block))
        if check:
            _checknonsingular(info.value)
        return LU(A, ipiv, info.value)


class DistributedLU(_MgHandle):
    """One rank's handle of a distributed n x n LU, one process per GPU.

    `exchange_id(id_bytes_or_None) -> id_bytes`: the launcher's transport; rank 0 passes the 128 bytes made by
    ``rfb_mg_unique_id``, every rank gets them back.  Usage:
        d = DistributedLU(n, np.float64, block=512, rank=r, world=w, device=local_rank, exchange_id=bcast)
        d.set_block(j, host_n_by_w) for j in d.my_blocks;  d.factor();  ms = d.synchronize()
        d.info(); d.pivots(); d.get_block(j)
    """

    def __init__(self, n: int, dtype=np.float64, block: int = 512, *, rank: int, world: int, device: int,
                 exchange_id: Optional[Callable[[Optional[bytes]], bytes]] = None):
        super().__init__()
        self.n, self.nb, self.dtype = int(n), int(block), np.dtype(dtype)
        if self.dtype not in (np.float64, np.float32):
            raise TypeError("DistributedLU supports float64 / float32")
        self.rank, self.world = rank, world
        idbuf = None
        if world > 1:
            if exchange_id is None:
                raise ValueError("world > 1 needs exchange_id")
            mine = None
            if rank == 0:
                raw = (C.c_ubyte * 128)()
                rc = self._lib.rfb_mg_unique_id(raw)
                if rc != _lib.RFB_OK:
                    raise RuntimeError("rfb_mg_unique_id failed (libnccl could not be loaded?)")
                mine = bytes(raw)
            got = exchange_id(mine)
            idbuf = (C.c_ubyte * 128).from_buffer_copy(got)
        rc = self._lib.rfb_mg_create_rank(C.byref(self._h), device, rank, world, idbuf)
        if rc != _lib.RFB_OK:
            self._check(rc)
        self._check(self._lib.rfb_mg_setup(self._h, self.n, self.nb, int(self.dtype == np.float32)))
        self.my_blocks = owned_blocks(rank, world, self.n, self.nb)
        self.item = self.dtype.itemsize
        ctxp = C.c_void_p()
        self._check(self._lib.rfb_mg_rank_ctx(self._h, 0, C.byref(ctxp), None))
        self.ctx_handle = ctxp

    def block_ptr(self, j: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.rfb_mg_block_ptr(self._h, 0, j, C.byref(p)))
        return p.value

    def set_block(self, j: int, host: np.ndarray) -> None:
        """Upload block column j (n x w, column-major) -- asynchronous, the factorization waits for it."""
        w = block_range(j, self.n, self.nb)[1]
        assert host.shape == (self.n, w) and host.dtype == self.dtype and host.flags.f_contiguous
        self._check(self._lib.rfb_mg_load_block(self._h, 0, j, C.c_void_p(host.ctypes.data), self.n, 0))

    def restore_block(self, j: int, dev_ptr: int) -> None:
        """Device-to-device copy of a pristine block column (benchmark restore between steps)."""
        self._check(self._lib.rfb_mg_load_block(self._h, 0, j, C.c_void_p(dev_ptr), self.n, 1))

    def get_block(self, j: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        w = block_range(j, self.n, self.nb)[1]
        if out is None:
            out = np.empty((self.n, w), dtype=self.dtype, order="F")
        self._check(self._lib.rfb_mg_store_block(self._h, 0, j, C.c_void_p(out.ctypes.data), self.n))
        self.synchronize()
        return out

    def store_block_async(self, j: int, out: np.ndarray) -> None:
        self._check(self._lib.rfb_mg_store_block(self._h, 0, j, C.c_void_p(out.ctypes.data), self.n))

    def factor(self) -> None:
        """Run the schedule (returns when the whole factorization is enqueued; `synchronize` waits for it)."""
        self._check(self._lib.rfb_mg_factor(self._h))

    def info(self) -> int:
        """Global info (collective: every rank must call it)."""
        v = C.c_int64()
        self._check(self._lib.rfb_mg_get_info(self._h, C.byref(v)))
        return int(v.value)

    def pivots(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.int64)
        self._check(self._lib.rfb_mg_get_pivots(self._h, C.c_void_p(out.ctypes.data)))
        return out
