"""ctypes binding of librfb200.so (the C ABI declared in include/rfb200.h).

The shared library is the product; this module only loads it and declares signatures.  If the
library is missing the import fails loudly -- there is no Python/NumPy fallback for any entry point.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librfb200.so")

RFB_OK, RFB_ERR_ARG, RFB_ERR_CUDA, RFB_ERR_NCCL, RFB_ERR_UNSUPPORTED, RFB_ERR_NOMEM, RFB_ERR_INTERNAL = range(7)
RFB_MEM_HOST, RFB_MEM_DEVICE = 0, 1
RFB_F32_AUTO, RFB_F32_TF32X3, RFB_F32_FP32 = 0, 1, 2

STATUS_NAMES = {
    RFB_OK: "RFB_OK", RFB_ERR_ARG: "RFB_ERR_ARG", RFB_ERR_CUDA: "RFB_ERR_CUDA", RFB_ERR_NCCL: "RFB_ERR_NCCL",
    RFB_ERR_UNSUPPORTED: "RFB_ERR_UNSUPPORTED", RFB_ERR_NOMEM: "RFB_ERR_NOMEM", RFB_ERR_INTERNAL: "RFB_ERR_INTERNAL",
}


class rfb_opts(C.Structure):
    """Mirror of `struct rfb_opts` (include/rfb200.h)."""
    _fields_ = [
        ("mem_space", C.c_int32),
        ("leaf_width", C.c_int32),
        ("f32_mode", C.c_int32),
        ("trsm_block", C.c_int32),
        ("gemm_path", C.c_int32),
        ("laswp_path", C.c_int32),
        ("no_pivot", C.c_int32),
        ("keep_factors", C.c_int32),
        ("reserved", C.c_int32 * 8),
    ]


_i64, _p, _int = C.c_int64, C.c_void_p, C.c_int

# name -> (restype, argtypes); one entry per function declared in include/rfb200.h
SIGNATURES = {
    "rfb_version": (_int, []),
    "rfb_create": (_int, [C.POINTER(_p), _int]),
    "rfb_destroy": (_int, [_p]),
    "rfb_last_error": (C.c_char_p, [_p]),
    "rfb_device_info": (_int, [_p, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int), C.POINTER(C.c_size_t)]),
    "rfb_set_default_opts": (_int, [_p, C.POINTER(rfb_opts)]),
    "rfb_set_early_download": (_int, [_p, _int]),
    "rfb_lu_f64": (_int, [_p, _p, _i64, _i64, _i64, _p, _p, C.POINTER(rfb_opts)]),
    "rfb_lu_f32": (_int, [_p, _p, _i64, _i64, _i64, _p, _p, C.POINTER(rfb_opts)]),
    "rfb_panel_getrf_f64": (_int, [_p, _p, _i64, _i64, _i64, _p, _i64, _p, _i64]),
    "rfb_panel_getrf_f32": (_int, [_p, _p, _i64, _i64, _i64, _p, _i64, _p, _i64]),
    "rfb_laswp_f64": (_int, [_p, _p, _i64, _i64, _p, _i64, _i64]),
    "rfb_laswp_f32": (_int, [_p, _p, _i64, _i64, _p, _i64, _i64]),
    "rfb_trsm_llnu_f64": (_int, [_p, _p, _i64, _p, _i64, _i64]),
    "rfb_trsm_llnu_f32": (_int, [_p, _p, _i64, _p, _i64, _i64]),
    "rfb_gemm_nn_sub_f64": (_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i64]),
    "rfb_gemm_nn_sub_f32": (_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i64]),
    "rfb_ipiv_shift": (_int, [_p, _p, _i64, _i64]),
    "rfb_trsm_lunn_f64": (_int, [_p, _p, _i64, _p, _i64, _i64]),
    "rfb_trsm_lunn_f32": (_int, [_p, _p, _i64, _p, _i64, _i64]),
    "rfb_solve_f64": (_int, [_p, _p, _i64, _i64, _p, _p, _i64, _i64, C.POINTER(rfb_opts)]),
    "rfb_solve_f32": (_int, [_p, _p, _i64, _i64, _p, _p, _i64, _i64, C.POINTER(rfb_opts)]),
    "rfb_kept_id": (_int, [_p, C.POINTER(_i64)]),
    "rfb_solve_kept_f64": (_int, [_p, _i64, _p, _i64, _i64]),
    "rfb_solve_kept_f32": (_int, [_p, _i64, _p, _i64, _i64]),
    "rfb_panel_getrf_nopiv_f64": (_int, [_p, _p, _i64, _i64, _i64, _p, _i64]),
    "rfb_panel_getrf_nopiv_f32": (_int, [_p, _p, _i64, _i64, _i64, _p, _i64]),
    "rfb_butterfly_mul_f64": (_int, [_p, _p, _i64, _i64, _p]),
    "rfb_butterfly_mul_f32": (_int, [_p, _p, _i64, _i64, _p]),
    "rfb_butterfly_vec_f64": (_int, [_p, _p, _i64, _i64, _i64, _p, _int]),
    "rfb_butterfly_vec_f32": (_int, [_p, _p, _i64, _i64, _i64, _p, _int]),
    "rfb_butterfly_solve_f64": (_int, [_p, _p, _i64, _i64, _p, _i64, _i64, _p, _p, C.POINTER(rfb_opts)]),
    "rfb_butterfly_solve_f32": (_int, [_p, _p, _i64, _i64, _p, _i64, _i64, _p, _p, C.POINTER(rfb_opts)]),
    "rfb_lu_batched_f64": (_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, C.POINTER(rfb_opts)]),
    "rfb_lu_batched_f32": (_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, C.POINTER(rfb_opts)]),
    "rfb_trace_lu": (_int, [_int, _i64, _i64, _i64, C.POINTER(rfb_opts), _int, _p, _i64, C.POINTER(_i64)]),
    "rfb_lu_range_f64": (_int, [_p, _p, _i64, _i64, _i64, _i64, _p, _p, C.POINTER(rfb_opts)]),
    "rfb_lu_range_f32": (_int, [_p, _p, _i64, _i64, _i64, _i64, _p, _p, C.POINTER(rfb_opts)]),
    "rfb_laswp_range_f64": (_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _int]),
    "rfb_laswp_range_f32": (_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _int]),
    "rfb_perm_buffers": (_int, [_p, _p, _p, _p, _i64]),
    "rfb_perm_buffers_release": (_int, [_p]),
    "rfb_copy2d": (_int, [_p, _p, C.c_size_t, _p, C.c_size_t, C.c_size_t, C.c_size_t]),
    "rfb_set_stream": (_int, [_p, _p]),
    "rfb_mg_unique_id": (_int, [_p]),
    "rfb_mg_create_rank": (_int, [C.POINTER(_p), _int, _int, _int, _p]),
    "rfb_mg_create_all": (_int, [C.POINTER(_p), _int, C.POINTER(_int)]),
    "rfb_mg_destroy": (_int, [_p]),
    "rfb_mg_last_error": (C.c_char_p, [_p]),
    "rfb_mg_setup": (_int, [_p, _i64, _i64, _int]),
    "rfb_mg_owner_of": (_int, [_i64, _int]),
    "rfb_mg_local_ranks": (_int, [_p, C.POINTER(_int), C.POINTER(_int)]),
    "rfb_mg_rank_ctx": (_int, [_p, _int, C.POINTER(_p), C.POINTER(_int)]),
    "rfb_mg_block_ptr": (_int, [_p, _int, _i64, C.POINTER(_p)]),
    "rfb_mg_load_block": (_int, [_p, _int, _i64, _p, _i64, _int]),
    "rfb_mg_store_block": (_int, [_p, _int, _i64, _p, _i64]),
    "rfb_mg_factor": (_int, [_p]),
    "rfb_mg_sync": (_int, [_p, C.POINTER(C.c_float)]),
    "rfb_mg_get_pivots": (_int, [_p, _p]),
    "rfb_mg_get_info": (_int, [_p, C.POINTER(_i64)]),
    "rfb_mg_stats": (_int, [_p, C.POINTER(_i64), C.POINTER(_i64)]),
    "rfb_mg_sched_stats": (_int, [_p, _int, C.POINTER(_i64)]),
    "rfb_mg_lu_f64": (_int, [_p, _p, _i64, _i64, _p, C.POINTER(_i64), _i64]),
    "rfb_mg_lu_f32": (_int, [_p, _p, _i64, _i64, _p, C.POINTER(_i64), _i64]),
    "rfb_lu_f64_mg": (_int, [C.POINTER(_int), _int, _p, _i64, _i64, _p, C.POINTER(_i64), _i64]),
    "rfb_lu_f32_mg": (_int, [C.POINTER(_int), _int, _p, _i64, _i64, _p, C.POINTER(_i64), _i64]),
    "rfb_mg_trace": (_int, [_i64, _i64, _int, _int, _p, _i64, C.POINTER(_i64)]),
    "rfb_malloc": (_int, [_p, C.POINTER(_p), C.c_size_t]),
    "rfb_free": (_int, [_p, _p]),
    "rfb_host_alloc": (_int, [_p, C.POINTER(_p), C.c_size_t]),
    "rfb_host_free": (_int, [_p, _p]),
    "rfb_h2d": (_int, [_p, _p, _p, C.c_size_t]),
    "rfb_d2h": (_int, [_p, _p, _p, C.c_size_t]),
    "rfb_d2d": (_int, [_p, _p, _p, C.c_size_t]),
    "rfb_memset": (_int, [_p, _p, _int, C.c_size_t]),
    "rfb_sync": (_int, [_p]),
    "rfb_timer_start": (_int, [_p]),
    "rfb_timer_stop": (_int, [_p, C.POINTER(C.c_float)]),
    "rfb_launch_count": (_int, [_p, C.POINTER(_i64)]),
    "rfb_profile_enable": (_int, [_p, _int]),
    "rfb_profile_read": (_int, [_p, C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(C.c_double)]),
    "rfb_bench_dmma_peak": (_int, [_p, _int, C.POINTER(C.c_double)]),
    "rfb_bench_copy": (_int, [_p, C.c_size_t, _int, C.POINTER(C.c_double)]),
    "rfb_bench_tf32_peak": (_int, [_p, _int, C.POINTER(C.c_double)]),
}

_lib = None


def load() -> C.CDLL:
    """Load librfb200.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == ABI drift, let it surface
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib
