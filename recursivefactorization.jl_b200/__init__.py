"""rfb200 -- B200-native recursive LU, host-side mirror of RecursiveFactorization.jl's `lu` / `lu!`.

The product is ``librfb200.so`` (hand-written sm_100a kernels behind the C ABI in
``include/rfb200.h``).  This package is the thin host layer a Python caller uses, shaped like the
reference's Julia API (paths relative to /root/reference):

===========================  =====================================================================
``lu(A, pivot, thread, ...)``   src/lu.jl:19-21   -- factor a copy
``lu_(A, ipiv, pivot, ...)``    src/lu.jl:67-83 and :97-130 (``lu!``) -- factor in place
``LU``                          LinearAlgebra.LU built at src/lu.jl:129 (factors / ipiv / info, L U p)
``SingularException``           raised by ``checknonsingular(info)`` at src/lu.jl:128 when ``check``
``NotIPIV`` / ``pivot=False``   src/lu.jl:27-65: unpivoted factorization with the lazy identity pivot vector
``ldiv_(F, B)``                 ``ldiv!(F, B)`` (LinearAlgebra; NotIPIV overload src/lu.jl:60-64)
``ButterflyWorkspace`` /        src/butterflylu.jl:20-55 (``🦋workspace`` / ``🦋solve!``), ``butterfly_mul_``
``butterfly_solve_``            = ``🦋mul!`` (:93-113)
``Adjoint`` / ``Transpose``     src/lu.jl:85-87
``lu_batched_``                 ``lu!`` over a strided batch of small matrices (README.md:34-35 workload)
===========================  =====================================================================

Arrays follow Julia's layout: column-major (Fortran-ordered) ``float64`` / ``float32``.  Pivots are
int64, 1-based, sequential-swap -- ``LinearAlgebra.LU.ipiv``.  There is no CPU fallback anywhere:
without the CUDA library or without a B200 every call raises.

Import name: the directory is called ``recursivefactorization.jl_b200`` (not importable as is because
of the dot); ``import rfb200`` (repo-root shim) loads it.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional

import numpy as np

from . import _lib
from ._lib import rfb_opts  # noqa: F401  (re-export)

__all__ = [
    "lu", "lu_", "ldiv_", "LU", "SingularException", "ZeroPivotException", "RfbError", "Context", "default_context",
    "DeviceMatrix", "nsplit", "RowMaximum", "NoPivot", "NotIPIV", "Adjoint", "Transpose", "AdjointLU",
    "ButterflyWorkspace", "butterfly_workspace", "butterfly_solve_", "butterfly_mul_", "butterfly_generate_random",
    "lu_batched_", "lu_batched", "trace_lu",
]


# ----------------------------------------------------------------------------------------------
# errors
# ----------------------------------------------------------------------------------------------
class RfbError(RuntimeError):
    """Non-numerical failure reported by librfb200 (bad argument, CUDA error, unsupported shape)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"{_lib.STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class SingularException(ArithmeticError):
    """LinearAlgebra.SingularException(info): U[info, info] is exactly zero (src/lu.jl:128)."""

    def __init__(self, info: int):
        super().__init__(f"matrix is singular to working precision: zero pivot at column {info}")
        self.info = info


class ZeroPivotException(ArithmeticError):
    """LinearAlgebra.ZeroPivotException(k): what ``checknonsingular`` throws for the NEGATIVE info an
    unpivoted factorization reports on Julia >= 1.11 (src/lu.jl:24-25, :323-326)."""

    def __init__(self, info: int):
        super().__init__(f"factorization encountered one or more zero pivots (first at column {info}); "
                         "consider switching to a pivoted LU factorization")
        self.info = info


def _checknonsingular(info: int):
    """LinearAlgebra.checknonsingular (called at src/lu.jl:128)."""
    if info > 0:
        raise SingularException(info)
    if info < 0:
        raise ZeroPivotException(-info)


class NotIPIV:
    """src/lu.jl:27-32: the lazy identity pivot vector of an unpivoted factorization (``ipiv[i] == i``)."""

    def __init__(self, n: int):
        self.len = int(n)

    def __len__(self):
        return self.len

    @property
    def size(self):
        return self.len

    def __getitem__(self, i):
        if isinstance(i, slice):
            return np.arange(1, self.len + 1, dtype=np.int64)[i]
        if not -self.len <= i < self.len:
            raise IndexError(i)
        return (i % self.len) + 1

    def __iter__(self):
        return iter(range(1, self.len + 1))

    def __array__(self, dtype=None, copy=None):
        return np.arange(1, self.len + 1, dtype=dtype or np.int64)

    def __eq__(self, other):
        return np.array_equal(np.asarray(self), np.asarray(other))

    def __repr__(self):
        return f"NotIPIV({self.len})"


class Adjoint:
    """Lazy ``A'`` (LinearAlgebra.Adjoint); ``lu(Adjoint(A)) == AdjointLU(lu(A))`` (src/lu.jl:85-87)."""

    def __init__(self, parent: np.ndarray):
        self.parent = parent


class Transpose(Adjoint):
    """Lazy ``transpose(A)``; identical to Adjoint for the real element types this library handles."""


class RowMaximum:   # LinearAlgebra.RowMaximum(), src/lu.jl:13
    pass


class NoPivot:      # LinearAlgebra.NoPivot(), src/lu.jl:14
    pass


def _normalize_pivot(pivot) -> bool:
    """src/lu.jl:10-17: Val(true)/RowMaximum() -> True, Val(false)/NoPivot() -> False."""
    if isinstance(pivot, (RowMaximum,)) or pivot is RowMaximum:
        return True
    if isinstance(pivot, (NoPivot,)) or pivot is NoPivot:
        return False
    if isinstance(pivot, (bool, np.bool_)):
        return bool(pivot)
    raise TypeError(f"pivot must be True/False, RowMaximum() or NoPivot(); got {pivot!r}")


def nsplit(dtype, n: int) -> int:
    """Split column of the recursion, src/lu.jl:158-162 (host logic shared with the C++ driver)."""
    k = max(2, 128 // np.dtype(dtype).itemsize)
    return ((n + k // 2) // k) * (k // 2) if n >= k else n // 2


# ----------------------------------------------------------------------------------------------
# context
# ----------------------------------------------------------------------------------------------
def _free_pinned(lib, addr: int):
    try:
        lib.rfb_host_free(None, C.c_void_p(addr))      # a NULL context is allowed: cudaFreeHost needs none
    except Exception:
        pass


class Context:
    """Owns one ``rfb_ctx`` (stream + device workspaces) on one GPU.  Not thread-safe."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.rfb_create(C.byref(self._h), device)
        if rc != _lib.RFB_OK:
            msg = self._lib.rfb_last_error(self._h).decode() if self._h else "rfb_create failed"
            if self._h:
                self._lib.rfb_destroy(self._h)
                self._h = C.c_void_p()
            raise RfbError(rc, msg)
        self.device = device

    # -- plumbing -------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != _lib.RFB_OK:
            raise RfbError(rc, self._lib.rfb_last_error(self._h).decode())

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rfb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self) -> dict:
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self._check(self._lib.rfb_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "mem_bytes": mem.value}

    def sync(self):
        self._check(self._lib.rfb_sync(self._h))

    def set_default_opts(self, **kw):
        """Options for the kernel-level ABI calls (gemm_path, trsm_block, ...); no kwargs = defaults."""
        if kw:
            o = _make_opts(**kw)
            self._check(self._lib.rfb_set_default_opts(self._h, C.byref(o)))
        else:
            self._check(self._lib.rfb_set_default_opts(self._h, None))

    def set_early_download(self, mode: int):
        """How `lu_` on a page-locked matrix sends finished factors back while it is still factoring
        (`rfb_set_early_download`): 2 finished tiles (default), 1 row bands at the right spine, 0 off."""
        self._check(self._lib.rfb_set_early_download(self._h, int(mode)))

    def malloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.rfb_malloc(self._h, C.byref(p), nbytes))
        return p.value

    def free(self, ptr: int):
        self._check(self._lib.rfb_free(self._h, C.c_void_p(ptr)))

    def pinned_empty(self, shape, dtype, order="F") -> np.ndarray:
        """numpy array over page-locked host memory (cudaHostAlloc); released when the array is collected."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._check(self._lib.rfb_host_alloc(self._h, C.byref(p), max(n, 16)))
        buf = (C.c_byte * max(n, 16)).from_address(p.value)
        # The allocation lives as long as any array (or view, or LU.factors) built over `buf`: numpy keeps `buf`
        # alive through .base, and the finalizer frees the page-locked block when the last of them is collected --
        # never while a view can still be dereferenced (closing the context does not free it).
        weakref.finalize(buf, _free_pinned, self._lib, p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape, order=order)

    def h2d(self, dst: int, src: np.ndarray):
        self._check(self._lib.rfb_h2d(self._h, C.c_void_p(dst), C.c_void_p(src.ctypes.data), src.nbytes))

    def d2h(self, dst: np.ndarray, src: int):
        self._check(self._lib.rfb_d2h(self._h, C.c_void_p(dst.ctypes.data), C.c_void_p(src), dst.nbytes))

    def d2d(self, dst: int, src: int, nbytes: int):
        self._check(self._lib.rfb_d2d(self._h, C.c_void_p(dst), C.c_void_p(src), nbytes))

    def memset(self, dst: int, value: int, nbytes: int):
        self._check(self._lib.rfb_memset(self._h, C.c_void_p(dst), value, nbytes))

    def timer_start(self):
        self._check(self._lib.rfb_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(self._lib.rfb_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        c = C.c_int64()
        self._check(self._lib.rfb_launch_count(self._h, C.byref(c)))
        return int(c.value)

    def profile_enable(self, on: bool = True):
        self._check(self._lib.rfb_profile_enable(self._h, int(on)))

    def profile_read(self) -> dict:
        ms = (C.c_double * 8)()
        cnt = (C.c_int64 * 8)()
        work = (C.c_double * 8)()
        self._check(self._lib.rfb_profile_read(self._h, ms, cnt, work))
        names = ["panel", "laswp", "trsm_diag", "gemm", "other"]
        return {n: {"ms": ms[i], "launches": cnt[i], "work": work[i]} for i, n in enumerate(names)}

    def dmma_peak_tflops(self, iters: int = 20000) -> float:
        v = C.c_double()
        self._check(self._lib.rfb_bench_dmma_peak(self._h, iters, C.byref(v)))
        return float(v.value)

    def tf32_peak_tflops(self, iters: int = 4000) -> float:
        v = C.c_double()
        self._check(self._lib.rfb_bench_tf32_peak(self._h, iters, C.byref(v)))
        return float(v.value)

    def copy_gbs(self, nbytes: int = 1 << 30, iters: int = 5) -> float:
        v = C.c_double()
        self._check(self._lib.rfb_bench_copy(self._h, nbytes, iters, C.byref(v)))
        return float(v.value)

    # -- whole path -----------------------------------------------------------------------------
    def lu_raw(self, a_ptr: int, m: int, n: int, lda: int, ipiv_ptr: int, info_ptr: int, dtype,
               opts: Optional[rfb_opts] = None) -> None:
        fn = self._lib.rfb_lu_f64 if np.dtype(dtype) == np.float64 else self._lib.rfb_lu_f32
        self._check(fn(self._h, C.c_void_p(a_ptr), m, n, lda, C.c_void_p(ipiv_ptr), C.c_void_p(info_ptr),
                       C.byref(opts) if opts is not None else None))


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    """Process-wide context on the GPU selected by LOCAL_RANK (one process per GPU) or device 0."""
    global _default_ctx
    if _default_ctx is None:
        import os
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


# ----------------------------------------------------------------------------------------------
# LinearAlgebra.LU
# ----------------------------------------------------------------------------------------------
class LU:
    """``LinearAlgebra.LU{T,Matrix{T},Vector{Int64}}`` as returned at src/lu.jl:129.

    ``factors`` holds L strictly below the diagonal (unit diagonal implied) and U on/above it;
    ``ipiv`` is 1-based sequential-swap; ``info`` 0 or the first zero-pivot column.
    """

    def __init__(self, factors: np.ndarray, ipiv: np.ndarray, info: int):
        self.factors, self.ipiv, self.info = factors, ipiv, int(info)
        self._kept = None                    # (context, id) of a device-resident copy (lu_(..., keep=True))

    def __iter__(self):                      # Julia: L, U, p = F
        return iter((self.L, self.U, self.p))

    @property
    def issuccess(self) -> bool:
        return self.info == 0

    @property
    def L(self) -> np.ndarray:
        m, n = self.factors.shape
        mn = min(m, n)
        out = np.tril(self.factors[:, :mn], -1)
        out[np.arange(mn), np.arange(mn)] = 1
        return out

    @property
    def U(self) -> np.ndarray:
        m, n = self.factors.shape
        return np.triu(self.factors[:min(m, n), :])

    @property
    def p(self) -> np.ndarray:
        """0-based row permutation with ``A[p, :] == L @ U`` (Julia's ``F.p`` minus one)."""
        m = self.factors.shape[0]
        p = np.arange(m)
        for i, ip in enumerate(self.ipiv):
            ip = int(ip) - 1
            if ip != i:
                p[i], p[ip] = p[ip], p[i]
        return p

    @property
    def P(self) -> np.ndarray:
        m = self.factors.shape[0]
        out = np.zeros((m, m), dtype=self.factors.dtype)
        out[np.arange(m), self.p] = 1
        return out

    def solve(self, B: np.ndarray, ctx: Optional["Context"] = None) -> np.ndarray:
        """``F \\ B``: returns ``U^-1 L^-1 P B`` computed on the GPU (does not modify ``B``)."""
        return ldiv_(self, np.array(B, dtype=self.factors.dtype, order="F", copy=True), ctx=ctx)

    def __repr__(self):
        return f"LU(factors={self.factors.shape} {self.factors.dtype}, info={self.info})"


class AdjointLU:
    """``adjoint(F)`` / ``transpose(F)`` of a factorization (src/lu.jl:85-87): ``parent`` factors the parent
    matrix, i.e. ``A' == (P' L U)' = U' L' P``."""

    def __init__(self, parent: LU):
        self.parent = parent

    @property
    def info(self):
        return self.parent.info

    @property
    def issuccess(self):
        return self.parent.issuccess

    def __repr__(self):
        return f"AdjointLU({self.parent!r})"


# ----------------------------------------------------------------------------------------------
# lu / lu!
# ----------------------------------------------------------------------------------------------
def _make_opts(mem_space=_lib.RFB_MEM_HOST, leaf_width=0, f32_mode=0, trsm_block=0, gemm_path=0,
               laswp_path=0, no_pivot=0, keep_factors=0) -> rfb_opts:
    o = rfb_opts()
    o.mem_space, o.leaf_width, o.f32_mode = mem_space, leaf_width, f32_mode
    o.trsm_block, o.gemm_path, o.laswp_path, o.no_pivot = trsm_block, gemm_path, laswp_path, int(no_pivot)
    o.keep_factors = int(keep_factors)
    return o


def _column_major_lda(A: np.ndarray) -> Optional[int]:
    """Leading dimension (elements) of a column-major view, or None when `A` is not one: rows must be contiguous
    (stride == itemsize whenever m > 1) and the column stride a multiple of the itemsize with lda >= m (whenever
    n > 1).  numpy reports arbitrary strides for length-1 axes, so only the axes that are walked are constrained --
    but a single-column / single-row view of a larger array is still checked on the axis that is walked."""
    m, n = A.shape
    it = A.itemsize
    if m == 0 or n == 0:
        return max(m, 1)                    # nothing is read or written
    if m > 1 and A.strides[0] != it:
        return None
    if n > 1:
        s1 = A.strides[1]
        if s1 <= 0 or s1 % it or s1 // it < max(m, 1):
            return None
        return s1 // it
    return max(m, 1)


def lu_(A, ipiv: Optional[np.ndarray] = None, pivot=True, thread=False, *, check=True,
        blocksize: Optional[int] = None, threshold: Optional[int] = None, ctx: Optional[Context] = None,
        leaf_width: int = 0, f32_mode: int = 0, trsm_block: int = 0, gemm_path: int = 0,
        laswp_path: int = 0, keep: bool = False):
    """``RecursiveFactorization.lu!`` (src/lu.jl:67-83 and :97-130): factor ``A`` in place.

    ``keep=True`` (square matrices): the factors and pivots also stay resident on the device, and ``ldiv_(F, B)`` /
    ``F.solve(B)`` then only move ``B`` (``rfb_solve_kept_*``) -- until another host-mode call on the same context reuses
    the staging buffer, after which they silently go back to uploading ``F.factors``.  The caller promises not to edit
    ``F.factors`` in between.

    ``A`` must be a column-major float64/float32 matrix (it is overwritten with L\\U and returned
    inside the ``LU``); ``ipiv``, if given, must be an int64 vector of length ``min(m, n)`` and is
    the one returned.  ``thread`` is accepted and ignored (the GPU is always "threaded");
    ``blocksize``/``threshold`` tune the reference's CPU register kernel and are accepted and
    ignored -- the analogous GPU knob is ``leaf_width``.  ``check=True`` raises
    ``SingularException`` when ``info > 0`` like ``checknonsingular`` (src/lu.jl:128).

    ``pivot=False`` / ``NoPivot()`` (src/lu.jl:27-65): no row interchanges; the returned ``ipiv`` is a
    ``NotIPIV`` unless the caller passed a vector, which is then filled with ``1:min(m,n)`` (:107-113); a
    zero pivot gives NEGATIVE ``info`` (Julia >= 1.11, :24-25) and, with ``check``, ``ZeroPivotException``.
    """
    if isinstance(A, Adjoint):                                    # src/lu.jl:85-87
        return AdjointLU(lu_(A.parent, ipiv, pivot, thread, check=check, ctx=ctx, leaf_width=leaf_width,
                             f32_mode=f32_mode, trsm_block=trsm_block, gemm_path=gemm_path, laswp_path=laswp_path))
    piv = _normalize_pivot(pivot)
    if not isinstance(A, np.ndarray) or A.ndim != 2:
        raise TypeError("A must be a 2-D numpy array")
    if A.dtype not in (np.float64, np.float32):
        raise TypeError(f"eltype {A.dtype} is not supported on the B200 path (Float64/Float32 only); no fallback")
    m, n = A.shape
    lda = _column_major_lda(A)
    if lda is None or not A.flags.writeable or not A.flags.aligned:
        raise TypeError("lu_ factors in place and needs a writeable column-major array (unit row stride, "
                        "column stride lda >= m); use lu() to factor a copy of any layout")
    mn = min(m, n)
    if ipiv is None:
        ipiv = np.empty(mn, dtype=np.int64) if piv else NotIPIV(mn)      # init_pivot, src/lu.jl:33-40
    elif isinstance(ipiv, NotIPIV):
        if piv:
            raise TypeError("NotIPIV is only valid with pivot=False")
        if ipiv.len != mn:
            raise ValueError(f"ipiv has length {ipiv.len}, expected min(m, n) = {mn}")
    else:
        if not isinstance(ipiv, np.ndarray) or ipiv.dtype != np.int64 or ipiv.ndim != 1 or \
                not ipiv.flags.c_contiguous:
            raise TypeError("ipiv must be a contiguous int64 vector (Vector{BlasInt})")
        if ipiv.size != mn:
            raise ValueError(f"ipiv has length {ipiv.size}, expected min(m, n) = {mn}")
    ctx = ctx or default_context()
    info = C.c_int64(0)
    opts = _make_opts(_lib.RFB_MEM_HOST, leaf_width, f32_mode, trsm_block, gemm_path, laswp_path, no_pivot=not piv,
                      keep_factors=keep)
    ipiv_ptr = ipiv.ctypes.data if (isinstance(ipiv, np.ndarray) and mn) else 0
    ctx.lu_raw(A.ctypes.data if A.size else 0, m, n, lda, ipiv_ptr, C.addressof(info), A.dtype, opts)
    kept = None
    if keep:
        kid = C.c_int64(0)
        ctx._check(ctx._lib.rfb_kept_id(ctx.handle, C.byref(kid)))
        if kid.value:
            kept = (ctx, kid.value)
    if check:
        _checknonsingular(info.value)
    F = LU(A, ipiv, info.value)
    F._kept = kept
    return F


def ldiv_(F: LU, B: np.ndarray, ctx: Optional[Context] = None) -> np.ndarray:
    """``LinearAlgebra.ldiv!(F::LU, B)`` for a square factorization: overwrite ``B`` (vector or column-major
    matrix of the factors' eltype) with ``U^-1 L^-1 P B`` (forward + back substitution on the GPU)."""
    f = F.factors
    n = f.shape[0]
    if f.ndim != 2 or f.shape[1] != n:
        raise ValueError("ldiv_ needs a square factorization")
    if not isinstance(B, np.ndarray) or B.dtype != f.dtype or B.shape[0] != n or B.ndim not in (1, 2):
        raise TypeError("B must be a numpy vector/matrix with n rows and the factors' eltype")
    if B.ndim == 1:
        ldb = max(n, 1) if (n <= 1 or B.strides[0] == B.itemsize) else None
    else:
        ldb = _column_major_lda(B)
    if ldb is None or not B.flags.writeable:
        raise TypeError("ldiv_ overwrites B and needs a writeable column-major array (unit row stride)")
    ldf = _column_major_lda(f)
    if ldf is None:
        f = np.asfortranarray(f)
        ldf = max(n, 1)
    nrhs = 1 if B.ndim == 1 else B.shape[1]
    kept = getattr(F, "_kept", None)
    if kept is not None and (ctx is None or ctx is kept[0]):       # factors still resident on the device: only B travels
        kctx, kid = kept
        cur = C.c_int64(0)
        kctx._check(kctx._lib.rfb_kept_id(kctx.handle, C.byref(cur)))
        if cur.value == kid:
            fnk = kctx._lib.rfb_solve_kept_f64 if f.dtype == np.float64 else kctx._lib.rfb_solve_kept_f32
            kctx._check(fnk(kctx.handle, kid, C.c_void_p(B.ctypes.data), nrhs, ldb))
            return B
    # NotIPIV: both legs are plain triangular solves, no interchanges (src/lu.jl:60-64)
    ipiv = None if isinstance(F.ipiv, NotIPIV) else np.ascontiguousarray(F.ipiv, dtype=np.int64)
    ctx = ctx or default_context()
    lib = ctx._lib
    fn = lib.rfb_solve_f64 if f.dtype == np.float64 else lib.rfb_solve_f32
    opts = _make_opts(_lib.RFB_MEM_HOST)
    ctx._check(fn(ctx.handle, C.c_void_p(f.ctypes.data), n, ldf,
                  C.c_void_p(ipiv.ctypes.data) if ipiv is not None else None,
                  C.c_void_p(B.ctypes.data), nrhs, ldb, C.byref(opts)))
    return B


def lu(A, pivot=True, thread=False, **kwargs):
    """``RecursiveFactorization.lu`` (src/lu.jl:19-21): ``lu!(copy(A), ...)``."""
    if isinstance(A, Adjoint):                                    # src/lu.jl:85-87
        return AdjointLU(lu(A.parent, pivot, thread, **kwargs))
    A = np.asarray(A)
    if A.ndim != 2:
        raise TypeError("A must be a matrix")
    return lu_(np.array(A, order="F", copy=True), None, pivot, thread, **kwargs)


def trace_lu(m: int, n: int, dtype=np.float64, pinned_host: bool = False, lda: Optional[int] = None,
             early_mode: Optional[int] = None, **opt_kw) -> np.ndarray:
    """The launch sequence `rfb_lu_*` would enqueue for an m x n matrix, without a GPU (`rfb_trace_lu`): an (nops, 8)
    int64 array, see include/rfb200.h.  `pinned_host=True` gives the schedule used for page-locked host matrices;
    `early_mode` (0 off, 1 row bands, 2 tiles) forces one of its early-download schemes instead of the default."""
    if pinned_host and early_mode is not None:
        pinned_host = 10 + int(early_mode)
    lib = _lib.load()
    lda = lda if lda is not None else max(m, 1)
    opts = _make_opts(_lib.RFB_MEM_DEVICE, **opt_kw)
    count = C.c_int64(0)
    is_f32 = int(np.dtype(dtype) == np.float32)
    rc = lib.rfb_trace_lu(is_f32, m, n, lda, C.byref(opts), int(pinned_host), None, 0, C.byref(count))
    if rc != _lib.RFB_OK:
        raise RfbError(rc, "rfb_trace_lu failed")
    ops = np.zeros((count.value, 8), dtype=np.int64)
    if count.value:
        rc = lib.rfb_trace_lu(is_f32, m, n, lda, C.byref(opts), int(pinned_host), C.c_void_p(ops.ctypes.data), count.value,
                              C.byref(count))
        if rc != _lib.RFB_OK:
            raise RfbError(rc, "rfb_trace_lu failed")
    return ops


# ----------------------------------------------------------------------------------------------
# butterfly solver (src/butterflylu.jl)
# ----------------------------------------------------------------------------------------------
def butterfly_generate_random(n: int, dtype=np.float64, seed: int = 888) -> np.ndarray:
    """``🦋generate_random!`` (src/butterflylu.jl:9-19): the 4n butterfly values exp(x)/2, x ~ U(-0.05, 0.05).

    The reference draws them from VectorizedRNG's Xoshift stream, whose output depends on the host's SIMD
    width (test/runtests.jl:143-152) and cannot be reproduced outside Julia; this draws from numpy's
    ``default_rng(seed)`` (default seed 888 like the reference's ``Val(888)``)."""
    rng = np.random.default_rng(seed)
    return (0.5 * np.exp(-0.05 + 0.1 * rng.random(4 * n))).astype(dtype)


class ButterflyWorkspace:
    """``🦋workspace(A, b)`` (src/butterflylu.jl:20-43): the matrix, right-hand side and random butterflies of
    one solve.  Unlike the reference it does not materialise dense ``U``/``V`` (:149-178) -- the device applies
    the butterflies in their factored O(n) form -- and padding to a multiple of 4 (``pad!``, :180-197) happens
    inside the library on the device copy."""

    def __init__(self, A: np.ndarray, b: np.ndarray, seed: int = 888, uv: Optional[np.ndarray] = None):
        if not isinstance(A, np.ndarray) or A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise TypeError("A must be a square numpy matrix")
        if A.dtype not in (np.float64, np.float32):
            raise TypeError(f"eltype {A.dtype} is not supported on the B200 path (Float64/Float32 only); no fallback")
        self.n = A.shape[0]
        self.A = np.asfortranarray(A)
        self.b = np.array(b, dtype=A.dtype, order="F", copy=True)
        if self.b.shape[0] != self.n or self.b.ndim not in (1, 2):
            raise ValueError("b must have n rows")
        npad = self.n if self.n % 4 == 0 else self.n + (4 - self.n % 4)
        self.ws = butterfly_generate_random(npad, A.dtype, seed) if uv is None else np.ascontiguousarray(uv, dtype=A.dtype)
        if self.ws.size != 4 * npad:
            raise ValueError(f"uv must hold 4 * {npad} values")
        self.out = np.empty_like(self.b)
        self.info = 0


butterfly_workspace = ButterflyWorkspace       # src/butterflylu.jl:57


def butterfly_solve_(ws: ButterflyWorkspace, thread=False, ctx: Optional[Context] = None) -> np.ndarray:
    """``🦋solve!(workspace, thread)`` (src/butterflylu.jl:45-55): transform, unpivoted recursive LU, two
    triangular solves, back-transform -- all on the GPU.  Returns ``ws.out`` (the solution of ``A x = b``)."""
    ctx = ctx or default_context()
    lib = ctx._lib
    fn = lib.rfb_butterfly_solve_f64 if ws.A.dtype == np.float64 else lib.rfb_butterfly_solve_f32
    n = ws.n
    ws.out[...] = ws.b
    nrhs = 1 if ws.out.ndim == 1 else ws.out.shape[1]
    info = C.c_int64(0)
    opts = _make_opts(_lib.RFB_MEM_HOST)
    ctx._check(fn(ctx.handle, C.c_void_p(ws.A.ctypes.data), n, max(n, 1), C.c_void_p(ws.out.ctypes.data), nrhs, max(n, 1),
                  C.c_void_p(ws.ws.ctypes.data), C.byref(info), C.byref(opts)))
    ws.info = info.value
    _checknonsingular(info.value)              # lu!(A, Val(false), thread) checks by default (:48)
    return ws.out


def butterfly_mul_(A: np.ndarray, uv: np.ndarray, ctx: Optional[Context] = None) -> np.ndarray:
    """``🦋mul!(A, uv)`` (src/butterflylu.jl:93-113): overwrite the square column-major ``A`` (size % 4 == 0)
    with ``U' A V``."""
    if not isinstance(A, np.ndarray) or A.ndim != 2 or A.shape[0] != A.shape[1] or not A.flags.f_contiguous:
        raise TypeError("A must be a square column-major numpy matrix")
    n = A.shape[0]
    if n % 4:
        raise ValueError("butterfly_mul_ needs a size divisible by 4")
    uv = np.ascontiguousarray(uv, dtype=A.dtype)
    if uv.size != 4 * n:
        raise ValueError(f"uv must hold 4 * {n} values")
    ctx = ctx or default_context()
    lib = ctx._lib
    fn = lib.rfb_butterfly_mul_f64 if A.dtype == np.float64 else lib.rfb_butterfly_mul_f32
    d_a, d_uv = ctx.malloc(max(A.nbytes, 16)), ctx.malloc(max(uv.nbytes, 16))
    try:
        ctx.h2d(d_a, A)
        ctx.h2d(d_uv, uv)
        ctx._check(fn(ctx.handle, C.c_void_p(d_a), n, max(n, 1), C.c_void_p(d_uv)))
        ctx.d2h(A, d_a)
        ctx.sync()
    finally:
        ctx.free(d_a)
        ctx.free(d_uv)
    return A


# ----------------------------------------------------------------------------------------------
# batched small factorizations
# ----------------------------------------------------------------------------------------------
def lu_batched_(A: np.ndarray, ipiv: Optional[np.ndarray] = None, pivot=True, *, check=True,
                ctx: Optional[Context] = None):
    """``lu!`` on every matrix of a batch: ``A`` has shape (batch, m, n) with each ``A[b]`` column-major
    (e.g. ``np.empty((batch, n, m)).transpose(0, 2, 1)``), is overwritten with the packed factors and returned
    as a list of ``LU`` views.  One launch factors all matrices when n <= 64 and m <= 128."""
    piv = _normalize_pivot(pivot)
    if not isinstance(A, np.ndarray) or A.ndim != 3 or A.dtype not in (np.float64, np.float32):
        raise TypeError("A must be a 3-D float64/float32 numpy array (batch, m, n)")
    batch, m, n = A.shape
    it = A.itemsize
    if batch and m and n:
        # (numpy reports arbitrary strides for length-1 axes: only constrain the axes that are actually walked)
        bad = (m > 1 and A.strides[1] != it) or (n > 1 and (A.strides[2] < m * it or A.strides[2] % it)) or \
            (batch > 1 and (A.strides[0] % it or A.strides[0] < (A.strides[2] if n > 1 else m * it) * n)) or not A.flags.writeable
        if bad:
            raise TypeError("each A[b] must be a writeable column-major matrix (strides (s, itemsize, lda*itemsize))")
    mn = min(m, n)
    if piv:
        if ipiv is None:
            ipiv = np.empty((batch, mn), dtype=np.int64)
        elif ipiv.dtype != np.int64 or ipiv.shape != (batch, mn) or not ipiv.flags.c_contiguous:
            raise TypeError("ipiv must be a C-contiguous int64 array of shape (batch, min(m, n))")
    info = np.zeros(batch, dtype=np.int64)
    ctx = ctx or default_context()
    lib = ctx._lib
    fn = lib.rfb_lu_batched_f64 if A.dtype == np.float64 else lib.rfb_lu_batched_f32
    lda = A.strides[2] // it if (m and n > 1) else max(m, 1)
    stride = A.strides[0] // it if (batch > 1 and m and n) else lda * max(n, 1)
    opts = _make_opts(_lib.RFB_MEM_HOST, no_pivot=not piv)
    ctx._check(fn(ctx.handle, C.c_void_p(A.ctypes.data) if A.size else None, m, n, lda, stride, batch,
                  C.c_void_p(ipiv.ctypes.data) if (piv and ipiv.size) else None, C.c_void_p(info.ctypes.data), C.byref(opts)))
    if check:
        for v in info:
            _checknonsingular(int(v))
    return [LU(A[b], ipiv[b] if piv else NotIPIV(mn), int(info[b])) for b in range(batch)]


def lu_batched(A, pivot=True, **kwargs):
    """Factor copies: ``A`` is any (batch, m, n) array-like."""
    A = np.asarray(A)
    if A.ndim != 3:
        raise TypeError("A must be (batch, m, n)")
    buf = np.empty((A.shape[0], A.shape[2], A.shape[1]), dtype=A.dtype).transpose(0, 2, 1)
    buf[...] = A
    return lu_batched_(buf, None, pivot, **kwargs)


# ----------------------------------------------------------------------------------------------
# device-resident use (inputs already in HBM): what the roofline numbers are measured on
# ----------------------------------------------------------------------------------------------
class DeviceMatrix:
    """A column-major matrix + pivot vector + info word living in device memory."""

    def __init__(self, ctx: Context, m: int, n: int, dtype, lda: Optional[int] = None):
        self.ctx, self.m, self.n, self.dtype = ctx, m, n, np.dtype(dtype)
        self.lda = lda if lda is not None else max((m + 1) & ~1, 2)
        self.nbytes = self.lda * max(n, 1) * self.dtype.itemsize
        self.ptr = ctx.malloc(self.nbytes)
        self.ipiv_ptr = ctx.malloc(max(min(m, n), 1) * 8)
        self.info_ptr = ctx.malloc(64)

    def upload(self, a: np.ndarray):
        assert a.shape == (self.m, self.n) and a.dtype == self.dtype
        if self.lda == self.m and a.flags.f_contiguous:
            self.ctx.h2d(self.ptr, a)
        else:
            staged = np.zeros((self.lda, self.n), dtype=self.dtype, order="F")
            staged[: self.m, :] = a
            self.ctx.h2d(self.ptr, staged)
            self.ctx.sync()

    def download(self):
        staged = np.empty((self.lda, self.n), dtype=self.dtype, order="F")
        ipiv = np.empty(min(self.m, self.n), dtype=np.int64)
        info = np.zeros(8, dtype=np.int64)
        self.ctx.d2h(staged, self.ptr)
        if ipiv.size:
            self.ctx.d2h(ipiv, self.ipiv_ptr)
        self.ctx.d2h(info[:1], self.info_ptr)
        self.ctx.sync()
        return np.asfortranarray(staged[: self.m, :]), ipiv, int(info[0])

    def copy_from(self, other: "DeviceMatrix"):
        assert other.nbytes == self.nbytes
        self.ctx.d2d(self.ptr, other.ptr, self.nbytes)

    def lu(self, **opt_kw):
        """Enqueue the factorization where the matrix lies (no copies, no sync)."""
        opts = _make_opts(_lib.RFB_MEM_DEVICE, **opt_kw)
        self.ctx.lu_raw(self.ptr, self.m, self.n, self.lda, self.ipiv_ptr, self.info_ptr, self.dtype, opts)

    def free(self):
        for p in (self.ptr, self.ipiv_ptr, self.info_ptr):
            if p:
                self.ctx.free(p)
        self.ptr = self.ipiv_ptr = self.info_ptr = 0
