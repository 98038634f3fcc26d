"""CPU oracle for the recursive LU hot path -- TEST INFRASTRUCTURE, not product.

Two independent restatements of /root/reference/src/lu.jl (RecursiveFactorization.jl 0.2.30):

* ``lu_c`` & friends: ctypes bindings of ``oracle/librf_oracle.so`` (rf_oracle.c), the fast one,
  also used as the timed CPU baseline (``threads`` = OpenMP threads, the analogue of the
  reference's ``thread=Val(true)``).
* ``lu_numpy``: a small numpy twin written separately, used only to cross-check the C code.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  Bitwise parity with the Julia reference is UNPINNED (see the
header of rf_oracle.c for what is pinned instead).

All matrices are column-major (Fortran order) numpy arrays; pivots are 1-based, sequential-swap
(LAPACK ``ipiv`` semantics), exactly what ``LinearAlgebra.LU.ipiv`` holds (src/lu.jl:129).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librf_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with oracle/Makefile (gcc is part of the image)."""
    src_mtime = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("rf_oracle.c", "rf_oracle_impl.h"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_mtime:
        subprocess.check_call(["make", "-s", "-C", _HERE, "librf_oracle.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(_SO)
        i64, p = C.c_int64, C.c_void_p
        for suf in ("f64", "f32"):
            f = getattr(lib, f"rfo_lu_{suf}")
            f.restype, f.argtypes = i64, [p, i64, i64, i64, p, i64, i64, C.c_int]
            f = getattr(lib, f"rfo_panel_{suf}")
            f.restype, f.argtypes = i64, [p, i64, i64, i64, p, i64]
            f = getattr(lib, f"rfo_laswp_{suf}")
            f.restype, f.argtypes = None, [p, i64, i64, p, i64]
            f = getattr(lib, f"rfo_trsm_{suf}")
            f.restype, f.argtypes = None, [p, i64, p, i64, i64, C.c_int]
            f = getattr(lib, f"rfo_schur_{suf}")
            f.restype, f.argtypes = None, [p, p, p, i64, i64, i64, i64, C.c_int]
            f = getattr(lib, f"rfo_nsplit_pub_{suf}")
            f.restype, f.argtypes = i64, [i64]
            f = getattr(lib, f"rfo_lu_nopiv_{suf}")
            f.restype, f.argtypes = i64, [p, i64, i64, i64, p, i64, i64, C.c_int]
            f = getattr(lib, f"rfo_panel_nopiv_{suf}")
            f.restype, f.argtypes = i64, [p, i64, i64, i64, i64]
            f = getattr(lib, f"rfo_ldiv_notipiv_{suf}")
            f.restype, f.argtypes = None, [p, i64, i64, p, i64, i64]
            f = getattr(lib, f"rfo_butterfly_mul_{suf}")
            f.restype, f.argtypes = None, [p, i64, i64, p]
        _lib = lib
    return _lib


def _suf(a: np.ndarray) -> str:
    if a.dtype == np.float64:
        return "f64"
    if a.dtype == np.float32:
        return "f32"
    raise TypeError(f"oracle handles float64/float32 only, got {a.dtype}")


def _check_f(a: np.ndarray):
    if a.ndim != 2 or not a.flags.f_contiguous:
        raise ValueError("oracle wants a 2-D Fortran-ordered array")


def nsplit(dtype, n: int) -> int:
    """src/lu.jl:158-162."""
    suf = "f64" if np.dtype(dtype) == np.float64 else "f32"
    return int(getattr(_load(), f"rfo_nsplit_pub_{suf}")(n))


def lu_c(a: np.ndarray, blocksize: int = 0, threshold: int = 0, threads: int = 1):
    """In-place LU of Fortran-ordered ``a`` (m x n).  Returns (a, ipiv int64 1-based, info).

    Follows src/lu.jl:97-156 with the reference defaults when blocksize/threshold are 0.
    """
    _check_f(a)
    m, n = a.shape
    ipiv = np.zeros(min(m, n), dtype=np.int64)
    lda = a.strides[1] // a.itemsize if n > 1 else max(m, 1)
    info = getattr(_load(), f"rfo_lu_{_suf(a)}")(a.ctypes.data, m, n, lda, ipiv.ctypes.data,
                                                  blocksize, threshold, threads)
    return a, ipiv, int(info)


def lu_nopiv_c(a: np.ndarray, ipiv=None, blocksize: int = 0, threshold: int = 0, threads: int = 1):
    """``lu!(A, [ipiv,] Val(false))`` (src/lu.jl:97-130 with Pivot = false).  In place; returns
    (a, ipiv or None, info) -- a user ``ipiv`` is filled with 1:min(m,n) (:111-113); a zero pivot
    gives NEGATIVE info (Julia >= 1.11 convention, :24-25, :323-326)."""
    _check_f(a)
    m, n = a.shape
    lda = a.strides[1] // a.itemsize if n > 1 else max(m, 1)
    if ipiv is not None:
        assert ipiv.dtype == np.int64 and ipiv.size == min(m, n)
    info = getattr(_load(), f"rfo_lu_nopiv_{_suf(a)}")(a.ctypes.data, m, n, lda,
                                                        ipiv.ctypes.data if ipiv is not None else None,
                                                        blocksize, threshold, threads)
    return a, ipiv, int(info)


def panel_nopiv_c(a: np.ndarray):
    """src/lu.jl:290-338 with Pivot = false on the whole block (unblocked).  Returns (a, info)."""
    _check_f(a)
    m, n = a.shape
    info = getattr(_load(), f"rfo_panel_nopiv_{_suf(a)}")(a.ctypes.data, m, n, max(m, 1), 0)
    return a, int(info)


def ldiv_notipiv_c(f: np.ndarray, b: np.ndarray):
    """src/lu.jl:60-64: ``ldiv!(F::LU{..,NotIPIV}, B)`` = U^-1 L^-1 B, in place on Fortran-ordered b."""
    _check_f(f)
    n = f.shape[0]
    b2 = b.reshape(n, -1, order="F") if b.ndim == 1 else b
    assert b2.flags.f_contiguous or b2.shape[1] == 1
    getattr(_load(), f"rfo_ldiv_notipiv_{_suf(f)}")(f.ctypes.data, n, max(n, 1), b2.ctypes.data, b2.shape[1], max(n, 1))
    return b


def butterfly_mul_c(a: np.ndarray, uv: np.ndarray):
    """src/butterflylu.jl:93-113 ``🦋mul!(A, uv)``: A <- U' A V in place (square, size % 4 == 0)."""
    _check_f(a)
    m = a.shape[0]
    assert a.shape[1] == m and m % 4 == 0 and uv.dtype == a.dtype and uv.size == 4 * m
    uv = np.ascontiguousarray(uv)
    getattr(_load(), f"rfo_butterfly_mul_{_suf(a)}")(a.ctypes.data, m, max(m, 1), uv.ctypes.data)
    return a


def panel_c(a: np.ndarray):
    """src/lu.jl:290-338 on the whole block (unblocked).  Returns (a, ipiv, info)."""
    _check_f(a)
    m, n = a.shape
    ipiv = np.zeros(min(m, n), dtype=np.int64)
    info = getattr(_load(), f"rfo_panel_{_suf(a)}")(a.ctypes.data, m, n, max(m, 1), ipiv.ctypes.data, 0)
    return a, ipiv, int(info)


def laswp_c(a: np.ndarray, ipiv: np.ndarray):
    """src/lu.jl:164-188 on all columns of ``a``."""
    _check_f(a)
    ipiv = np.ascontiguousarray(ipiv, dtype=np.int64)
    getattr(_load(), f"rfo_laswp_{_suf(a)}")(a.ctypes.data, a.shape[1], max(a.shape[0], 1),
                                             ipiv.ctypes.data, ipiv.size)
    return a


def trsm_c(big: np.ndarray, l_off, k: int, b_off, nrhs: int, threads: int = 1):
    """B <- unitlower(L)^-1 B with L, B sub-blocks (row, col offsets) of one allocation ``big``."""
    _check_f(big)
    lda, it = big.shape[0], big.itemsize
    base = big.ctypes.data
    lp = base + (l_off[0] + l_off[1] * lda) * it
    bp = base + (b_off[0] + b_off[1] * lda) * it
    getattr(_load(), f"rfo_trsm_{_suf(big)}")(lp, k, bp, nrhs, lda, threads)
    return big


def schur_c(big: np.ndarray, c_off, a_off, b_off, m: int, n: int, k: int, threads: int = 1):
    """C -= A B with C, A, B sub-blocks of one allocation ``big`` (src/lu.jl:265-284)."""
    _check_f(big)
    lda, it = big.shape[0], big.itemsize
    base = big.ctypes.data
    ptr = lambda off: base + (off[0] + off[1] * lda) * it
    getattr(_load(), f"rfo_schur_{_suf(big)}")(ptr(c_off), ptr(a_off), ptr(b_off), m, n, k, lda, threads)
    return big


# ----------------------------------------------------------------------------------------------
# numpy twin (independent of the C code; slow, small sizes only)
# ----------------------------------------------------------------------------------------------

def _np_nsplit(itemsize: int, n: int) -> int:
    k = max(2, 128 // itemsize)
    return ((n + k // 2) // k) * (k // 2) if n >= k else n // 2


def _np_leaf(a, ipiv, info, pivot=True):
    """src/lu.jl:290-338."""
    m, n = a.shape
    one = a.dtype.type(1)
    for k in range(len(ipiv)):
        kp = k
        if pivot:
            col = np.abs(a[k:, k])
            amax = a.dtype.type(0)
            for i, v in enumerate(col):      # first strict maximum, NaN never wins
                if v > amax:
                    kp, amax = k + i, v
            ipiv[k] = kp + 1
        if a[kp, k] != 0:
            if kp != k:
                a[[k, kp], :] = a[[kp, k], :]
            a[k + 1:, k] *= one / a[k, k]
        elif info == 0:
            info = k + 1 if pivot else -(k + 1)
        if k == len(ipiv) - 1:
            break
        a[k + 1:, k + 1:] -= np.outer(a[k + 1:, k], a[k, k + 1:])
    return info


def _np_perm(p, a):
    for i, ip in enumerate(p):
        ip = int(ip) - 1
        if ip != i:
            a[[i, ip], :] = a[[ip, i], :]


def _np_trsm(l, b):
    k = l.shape[0]
    for c in range(k):
        b[c + 1:, :] -= np.outer(l[c + 1:, c], b[c, :])


def _np_rec(a, ipiv, info, blocksize, pivot=True):
    m, n = a.shape
    if n <= max(blocksize, 1):
        return _np_leaf(a, ipiv, info, pivot)
    n1 = _np_nsplit(a.itemsize, n)
    p1, p2 = ipiv[:n1], ipiv[n1:]
    info = _np_rec(a[:, :n1], p1, info, blocksize, pivot)
    if pivot:
        _np_perm(p1, a[:, n1:])
    _np_trsm(a[:n1, :n1], a[:n1, n1:])
    a[n1:, n1:] = (-(a[n1:, :n1] @ a[:n1, n1:])) + a[n1:, n1:]
    prev = info
    info = _np_rec(a[n1:, n1:], p2, info, blocksize, pivot)
    if pivot:
        _np_perm(p2, a[n1:, :n1])
    if info != prev:
        info += -n1 if info < 0 else n1
    if pivot:
        p2 += n1
    return info


def lu_numpy(a: np.ndarray, blocksize: int = 0, threshold: int = 0, pivot: bool = True):
    """numpy twin of src/lu.jl:97-156.  In place; returns (a, ipiv, info).  ``pivot=False``: ipiv is
    the identity (what a user vector is filled with, :107-113), negative info."""
    m, n = a.shape
    if blocksize <= 0:
        blocksize = 8 if m * n >= 40000 else 16
    if threshold <= 0:
        threshold = 48
    mn = min(m, n)
    ipiv = np.zeros(mn, dtype=np.int64)
    info = 0
    if mn == 0:
        return a, ipiv, 0
    if not pivot:
        ipiv[:] = np.arange(1, mn + 1)
    if mn > threshold:
        info = _np_rec(a[:, :mn], ipiv, info, blocksize, pivot)
        if m < n:
            if pivot:
                _np_perm(ipiv, a[:, m:])
            _np_trsm(a[:, :m], a[:, m:])
    else:
        info = _np_leaf(a, ipiv, info, pivot)
    return a, ipiv, int(info)


# ----------------------------------------------------------------------------------------------
# butterfly solver (src/butterflylu.jl) -- numpy restatement, used to check the C level loop, the
# materialised U / V and the whole 🦋solve! pipeline
# ----------------------------------------------------------------------------------------------

def butterfly_vals(n: int, dtype=np.float64, seed: int = 888) -> np.ndarray:
    """Stand-in for ``🦋generate_random!`` (src/butterflylu.jl:9-19): 4n values exp(x)/2 with
    x ~ U(-0.05, 0.05).  The reference draws them from VectorizedRNG's Xoshift stream, which is
    SIMD-width dependent (see the comment at test/runtests.jl:143-152) and not reproducible outside
    Julia; any such vector defines a valid transform, and product and oracle are always handed the
    SAME vector."""
    rng = np.random.default_rng(seed)
    return (0.5 * np.exp(-0.05 + 0.1 * rng.random(4 * n))).astype(dtype)


def _np_butterfly_level(a, u, v):
    """src/butterflylu.jl:59-91 on a view, vectorised."""
    mh, nh = a.shape[0] // 2, a.shape[1] // 2
    a11, a21, a12, a22 = a[:mh, :nh].copy(), a[mh:, :nh].copy(), a[:mh, nh:].copy(), a[mh:, nh:].copy()
    t1, t2, t3, t4 = a11 + a12, a21 + a22, a11 - a12, a21 - a22
    u1, u2, v1, v2 = u[:mh, None], u[mh:, None], v[None, :nh], v[None, nh:]
    a[:mh, :nh] = u1 * (t1 + t2) * v1
    a[mh:, :nh] = u2 * (t1 - t2) * v1
    a[:mh, nh:] = u1 * (t3 + t4) * v2
    a[mh:, nh:] = u2 * (t3 - t4) * v2


def butterfly_mul_numpy(a: np.ndarray, uv: np.ndarray):
    """src/butterflylu.jl:93-113."""
    m = a.shape[0]
    h = m // 2
    u1, v1, u2, v2 = uv[:h], uv[h:m], uv[m:m + h], uv[m + h:2 * m]
    _np_butterfly_level(a[:h, :h], u1, v1)
    _np_butterfly_level(a[h:, :h], u2, v1)
    _np_butterfly_level(a[:h, h:], u1, v2)
    _np_butterfly_level(a[h:, h:], u2, v2)
    _np_butterfly_level(a, uv[2 * m:3 * m], uv[3 * m:4 * m])
    return a


def _np_butterfly_block(x):
    """src/butterflylu.jl:134-147 ``🦋!(C, Diagonal(y), Diagonal(z))`` with y, z the halves of x
    (``diagnegbottom``, :115-126): [[D(y), D(z)], [D(y), -D(z)]]."""
    h = x.size // 2
    y, z = np.diag(x[:h]), np.diag(x[h:])
    return np.block([[y, z], [y, -z]])


def butterfly_materialize(uv: np.ndarray, m: int):
    """src/butterflylu.jl:149-178 ``materializeUV``: dense U = Bu2 Bu1, V = Bv2 Bv1."""
    h = m // 2
    z = np.zeros((h, h), dtype=uv.dtype)
    bu2 = np.block([[_np_butterfly_block(uv[:h]), z], [z, _np_butterfly_block(uv[m:m + h])]])
    bv2 = np.block([[_np_butterfly_block(uv[h:m]), z], [z, _np_butterfly_block(uv[m + h:2 * m])]])
    bu1 = _np_butterfly_block(uv[2 * m:3 * m])
    bv1 = _np_butterfly_block(uv[3 * m:4 * m])
    return bu2 @ bu1, bv2 @ bv1


def butterfly_pad(a: np.ndarray):
    """src/butterflylu.jl:180-197 ``pad!``: grow to the next multiple of 4 with an identity corner.
    (Like the reference it adds 4 when the size is already a multiple of 4 -- only called when it
    is not, :34-38.)"""
    m = a.shape[0]
    xn = 4 - m % 4
    out = np.zeros((m + xn, m + xn), dtype=a.dtype, order="F")
    out[:m, :m] = a
    out[np.arange(m, m + xn), np.arange(m, m + xn)] = 1
    return out


def butterfly_solve_oracle(a: np.ndarray, b: np.ndarray, uv=None, threads: int = 1):
    """``🦋solve!(🦋workspace(A, b))`` (src/butterflylu.jl:20-55) restated: pad, transform
    (C loop), NoPivot recursive LU (C), tmp = U'b, NotIPIV ldiv!, x = V tmp.  The padding rows of
    b are zeros here (the reference appends rand(xn), :37 -- they only touch the discarded tail
    of the solution because the padded system is block diagonal)."""
    n = a.shape[0]
    a = np.asfortranarray(a.copy())
    b = b.astype(a.dtype).copy()
    if n % 4:
        a = butterfly_pad(a)
        b = np.concatenate([b, np.zeros(a.shape[0] - n, dtype=a.dtype)])
    m = a.shape[0]
    if uv is None:
        uv = butterfly_vals(m, a.dtype)
    butterfly_mul_c(a, uv)
    _, _, info = lu_nopiv_c(a, threads=threads)
    u, v = butterfly_materialize(uv, m)
    tmp = np.asfortranarray((u.T @ b).reshape(m, 1))
    ldiv_notipiv_c(a, tmp)
    x = v @ tmp[:, 0]
    return x[:n], info


# ----------------------------------------------------------------------------------------------
# helpers shared by the tests / bench (what `testlu` in test/runtests.jl:14-31 computes)
# ----------------------------------------------------------------------------------------------

def perm_from_ipiv(ipiv: np.ndarray, m: int) -> np.ndarray:
    """Row permutation p (0-based) such that (P A) = A[p, :]; LinearAlgebra's ``F.p``."""
    p = np.arange(m)
    for i, ip in enumerate(np.asarray(ipiv)):
        ip = int(ip) - 1
        if ip != i:
            p[i], p[ip] = p[ip], p[i]
    return p


def split_lu(f: np.ndarray):
    """(L, U) from packed factors: L m x min(m,n) unit lower, U min(m,n) x n upper."""
    m, n = f.shape
    mn = min(m, n)
    l = np.tril(f[:, :mn], -1) + np.eye(m, mn, dtype=f.dtype)
    u = np.triu(f[:mn, :])
    return l, u


def residual_inf(a0: np.ndarray, f: np.ndarray, ipiv: np.ndarray) -> float:
    """||L U - A[p,:]||_inf, computed in float64 (test/runtests.jl:20)."""
    l, u = split_lu(np.asarray(f, dtype=np.float64))
    p = perm_from_ipiv(ipiv, a0.shape[0])
    r = l @ u - np.asarray(a0, dtype=np.float64)[p, :]
    return float(np.abs(r).sum(axis=1).max()) if r.size else 0.0


def residual_fro_rel(a0: np.ndarray, f: np.ndarray, ipiv: np.ndarray) -> float:
    """||P A - L U||_F / ||A||_F in float64 (BASELINE.json metric)."""
    l, u = split_lu(np.asarray(f, dtype=np.float64))
    p = perm_from_ipiv(ipiv, a0.shape[0])
    a64 = np.asarray(a0, dtype=np.float64)
    den = np.linalg.norm(a64)
    return float(np.linalg.norm(l @ u - a64[p, :]) / (den if den > 0 else 1.0))
