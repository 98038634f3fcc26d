"""Host-driven recursion over the KERNEL-LEVEL C ABI: the executable twin of `reckernel_device!` / `lu_device!` in
julia/RecursiveFactorizationB200.jl (the north_star's "Julia host code drives the recursion and calls the kernels
through a thin ccall shim").  No Julia runtime exists here, so the identical call sequence is issued through ctypes and
checked on the GPU bit for bit against the library's own C++ driver (`rfb_lu_*`, csrc/rfb_api.cu:lu_rec).

Restates lu! / _recurse! / reckernel! (src/lu.jl:97-130, :145-156, :189-263):
  leaf (n <= leaf columns)                         -> rfb_lu_range_*   (one K1 launch, global pivots / info, exchange list)
  apply_permutation! (:233, :246, :151)            -> rfb_laswp_range_* (list-driven K2)
  ldiv!(UnitLowerTriangular(A11), A12) (:235,:153) -> rfb_trsm_llnu_*
  schur_complement! (:240)                         -> rfb_gemm_nn_sub_*
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, nsplit


def lu_device_(ctx, d_a: int, m: int, n: int, lda: int, d_ipiv: int, d_info: int, dtype, leaf: int = 64) -> None:
    """Pivoted lu!(A, ipiv; check=false) of the m x n device matrix at `d_a` (leading dimension `lda`), enqueued on the
    context's stream.  `d_ipiv`: min(m, n) int64 on the device, `d_info`: one int64 on the device."""
    dtype = np.dtype(dtype)
    suf = "f64" if dtype == np.float64 else "f32"
    lib, h, es = ctx._lib, ctx.handle, dtype.itemsize
    lu_range = getattr(lib, f"rfb_lu_range_{suf}")
    laswp_range = getattr(lib, f"rfb_laswp_range_{suf}")
    trsm = getattr(lib, f"rfb_trsm_llnu_{suf}")
    gemm = getattr(lib, f"rfb_gemm_nn_sub_{suf}")
    mn = min(m, n)
    cap = mn + 64
    perm = [ctx.malloc(4 * 2 * cap), ctx.malloc(4 * 2 * cap), ctx.malloc(4 * cap)]
    opts = _lib.rfb_opts()
    opts.mem_space, opts.leaf_width = _lib.RFB_MEM_DEVICE, leaf
    ipiv_p, info_p, root = C.c_void_p(d_ipiv), C.c_void_p(d_info), C.c_void_p(d_a)

    def at(r, c):
        return C.c_void_p(d_a + es * (r + c * lda))

    def rec(c0, nn):                                                   # reckernel!, src/lu.jl:189-263
        if nn <= leaf:                                                 # :192-195
            ctx._check(lu_range(h, root, m, lda, c0, nn, ipiv_p, info_p, C.byref(opts)))
            return
        n1 = nsplit(dtype, nn)                                         # :196-198
        n2 = nn - n1
        rec(c0, n1)                                                    # :229
        ctx._check(laswp_range(h, root, lda, c0 + n1, n2, c0, c0 + n1, ipiv_p, 1))            # :233
        ctx._check(trsm(h, at(c0, c0), n1, at(c0, c0 + n1), n2, lda))                         # :235
        ctx._check(gemm(h, at(c0 + n1, c0 + n1), at(c0 + n1, c0), at(c0, c0 + n1), m - c0 - n1, n2, n1, lda))   # :240
        rec(c0 + n1, n2)                                               # :244
        ctx._check(laswp_range(h, root, lda, c0, n1, c0 + n1, c0 + nn, ipiv_p, 1))            # :246

    try:
        ctx._check(lib.rfb_perm_buffers(h, C.c_void_p(perm[0]), C.c_void_p(perm[1]), C.c_void_p(perm[2]), cap))
        ctx.memset(d_info, 0, 8)
        if mn:
            rec(0, mn)                                                 # :147
            if m < n:                                                  # fat tail, :148-154
                ctx._check(laswp_range(h, root, lda, m, n - m, 0, mn, ipiv_p, 1))
                ctx._check(trsm(h, root, m, at(0, m), n - m, lda))
        ctx.sync()
    finally:
        # hand the context's exchange-list slots back (the whole-path driver allocates its own on demand)
        ctx._check(lib.rfb_perm_buffers_release(h))
        for p in perm:
            ctx.free(p)
