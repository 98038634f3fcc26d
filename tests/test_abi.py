"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/rfb200.h declares, the ctypes table matches the header, host-side logic agrees with the
oracle, and argument errors / the missing-GPU case fail loudly (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import rfb200
from oracle import rf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "rfb200.h")).read()


def declared_functions():
    code = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(rfb_[a-z0-9_]+)\s*\(", code)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(rfb200._lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/rfb200.h but not exported"


def test_ctypes_table_matches_header():
    assert sorted(rfb200._lib.SIGNATURES) == declared_functions()
    assert ctypes.sizeof(rfb200.rfb_opts) == 64
    assert rfb200._lib.load().rfb_version() >= 100


def test_option_struct_is_mirrored_field_for_field():
    """`struct rfb_opts` in the header, the ctypes Structure and the Julia shim's RfbOpts must list the same fields in
    the same order (a silent mismatch would shift every option by four bytes)."""
    body = re.search(r"typedef struct rfb_opts \{(.*?)\} rfb_opts;", HEADER, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = re.findall(r"int32_t\s+([a-z_0-9]+)(?:\[(\d+)\])?;", body)
    py_fields = [(n, str(t._length_) if hasattr(t, "_length_") else "") for n, t in rfb200.rfb_opts._fields_]
    assert [(n, l) for n, l in c_fields] == py_fields
    assert 4 * sum(int(l or 1) for _, l in c_fields) == 64
    jl = open(os.path.join(ROOT, "recursivefactorization.jl_b200", "julia", "RecursiveFactorizationB200.jl")).read()
    jbody = re.search(r"struct RfbOpts\n(.*?)\nend", jl, flags=re.S).group(1)
    j_fields = re.findall(r"^\s*([a-z_0-9]+)::(Int32|NTuple\{(\d+), Int32\})", jbody, flags=re.M)
    assert [(n, l) for n, _, l in j_fields] == [(n, l) for n, l in c_fields]


def test_julia_shim_only_calls_exported_symbols():
    """Every C symbol the (un-executable) Julia shim names must be declared in include/rfb200.h, for both element types,
    and it must bind the whole reference surface: lu!, ldiv!, the butterfly solver, the batched call, the multi-GPU handle
    and the kernel-level entry points its own recursion drives."""
    jl = open(os.path.join(ROOT, "recursivefactorization.jl_b200", "julia", "RecursiveFactorizationB200.jl")).read()
    code = "\n".join(line.split("#", 1)[0] for line in jl.splitlines())           # comments name types like rfb_opts
    called = set()
    for tok in re.findall(r"(?<![A-Za-z0-9_])rfb_[a-z0-9_]+", code):
        called |= {tok + "f64", tok + "f32"} if tok.endswith("_") else {tok}     # Symbol("rfb_lu_range_", suf)
    want = {"rfb_create", "rfb_destroy", "rfb_perm_buffers", "rfb_memset", "rfb_mg_create_all", "rfb_mg_destroy"}
    for base in ("rfb_lu", "rfb_solve", "rfb_butterfly_solve", "rfb_lu_batched", "rfb_mg_lu", "rfb_lu_range", "rfb_laswp_range",
                 "rfb_trsm_llnu", "rfb_gemm_nn_sub"):
        want |= {base + "_f64", base + "_f32"}
    assert want <= called, want - called
    assert called <= set(declared_functions()), called - set(declared_functions())


def test_no_oracle_or_cpu_fallback_in_product():
    pkg = os.path.join(ROOT, "recursivefactorization.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f)).read()
                assert "rf_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f
                # no library LU/BLAS on the product path either (north_star: no cuBLAS/cuSOLVER, no Triton)
                banned_all = ["import scipy", "from scipy", "getrf(", "cublas", "cusolver", "import triton"]
                banned_all.append("import torch")   # (round 2: the multi-GPU driver is C++ too; no torch anywhere in the product)
                for banned in banned_all:
                    assert banned not in text.lower(), (f, banned)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_nsplit_matches_oracle(dtype):
    for n in list(range(0, 200)) + [300, 4096, 8192, 16384, 32768]:
        assert rfb200.nsplit(dtype, n) == O.nsplit(dtype, n)


def test_argument_errors_are_loud():
    a = np.zeros((4, 4), order="F")
    with pytest.raises(TypeError):
        rfb200.lu(a.astype(np.complex128))
    with pytest.raises(TypeError):
        rfb200.lu_(a, rfb200.NotIPIV(4), True)             # NotIPIV only with pivot = Val(false), src/lu.jl:33-40
    with pytest.raises(ValueError):
        rfb200.lu_(a, rfb200.NotIPIV(3), False)
    with pytest.raises(TypeError):
        rfb200.lu_batched_(np.zeros((2, 4, 4)))            # each matrix must be column-major
    with pytest.raises(TypeError):
        rfb200.ButterflyWorkspace(np.zeros((4, 5)), np.zeros(4))
    with pytest.raises(TypeError):
        rfb200.lu_(np.zeros((4, 4), order="C"))            # in-place needs column-major
    with pytest.raises(ValueError):
        rfb200.lu_(a, np.zeros(3, dtype=np.int64))         # ipiv length must be min(m, n)
    with pytest.raises(TypeError):
        rfb200.lu_(a, np.zeros(4, dtype=np.int32))         # Vector{BlasInt} is int64
    with pytest.raises(TypeError):
        rfb200.lu(a, "yes")


def test_notipiv_is_the_lazy_identity():
    """src/lu.jl:27-32."""
    p = rfb200.NotIPIV(5)
    assert len(p) == 5 and p[0] == 1 and p[4] == 5 and list(p) == [1, 2, 3, 4, 5] and p == np.arange(1, 6)
    F = rfb200.LU(np.asfortranarray(np.eye(5)), p, 0)
    assert np.array_equal(F.p, np.arange(5))
    with pytest.raises(rfb200.ZeroPivotException):
        rfb200._checknonsingular(-3)
    with pytest.raises(rfb200.SingularException):
        rfb200._checknonsingular(3)


def test_lu_object_properties():
    f = np.asfortranarray(np.array([[4.0, 3.0], [0.5, 2.0]]))
    F = rfb200.LU(f, np.array([2, 2], dtype=np.int64), 0)
    assert np.array_equal(F.L, [[1, 0], [0.5, 1]]) and np.array_equal(F.U, [[4, 3], [0, 2]])
    assert np.array_equal(F.p, [1, 0]) and F.issuccess
    l, u, p = F
    assert np.array_equal(F.P @ np.array([[1.0, 2], [3, 4]]), np.array([[3.0, 4], [1, 2]]))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    with pytest.raises(rfb200.RfbError) as ei:
        rfb200.Context(0)
    assert "no CPU fallback" in str(ei.value)


def test_lu_rejects_views_that_are_not_column_major():
    """a single-column / single-row view with non-unit strides must not reach the library (it would read
    and overwrite the parent's neighbouring elements).  The layout check runs before any context is created."""
    import rfb200
    bad = [np.zeros((5, 3))[:, 0:1],                      # m x 1 view of a C-ordered array: row stride 24
           np.zeros((4, 6), order="F")[0:1, :].T,         # n x 1 with row stride 32
           np.zeros((6, 6))[:, :],                        # C-ordered square
           np.zeros((8, 8), order="F")[::2, :]]           # row stride 16
    for a in bad:
        with pytest.raises(TypeError):
            rfb200.lu_(a)
    ok = rfb200._column_major_lda
    assert ok(np.zeros((5, 3), order="F")) == 5
    assert ok(np.zeros((10, 6), order="F")[:8, :]) == 10       # lda > m view
    assert ok(np.zeros((10, 6), order="F")[:8, 2:3]) == 8      # single column of an F array: contiguous
    assert ok(np.zeros((4, 6), order="F")[0:1, :]) == 4        # 1 x n view: columns 4 elements apart
    assert ok(np.zeros((5, 3))[:, 0:1]) is None
    assert ok(np.zeros((0, 0), order="F")) == 1
    f = rfb200.LU(np.zeros((3, 3), order="F"), np.arange(1, 4), 0)
    with pytest.raises(TypeError):
        rfb200.ldiv_(f, np.zeros((3, 4))[:, 1:2])              # B: m x 1 view with row stride 32
    with pytest.raises(TypeError):
        rfb200.ldiv_(f, np.zeros(6)[::2])                      # strided vector
