"""The shipped librfb200.so must contain the Blackwell instructions DESIGN.md says its hot kernels use (checked on the
SASS of the built library with cuobjdump, no GPU needed): TMA tensor loads + FP64 tensor MMAs + TMA bulk f64 reduce-adds
in the trailing GEMM, tcgen05 MMAs with TMEM loads in the Float32 GEMM, DMMA in the triangular solve, redux.sync (SASS CREDUX) in
the pivot search.  A rebuild that silently falls back to SIMT code (wrong arch flags, a refactor that drops the inline PTX)
fails here instead of showing up as a slower bench."""
import collections
import os
import re
import shutil
import subprocess

import pytest

import rfb200

LIB = os.path.join(os.path.dirname(os.path.abspath(rfb200.__file__)), "librfb200.so")
if not os.path.exists(LIB):      # rfb200.py is a shim at the repo root; the library lives in the package directory
    LIB = os.path.join(os.path.dirname(os.path.abspath(rfb200.__file__)), "recursivefactorization.jl_b200", "librfb200.so")


@pytest.fixture(scope="module")
def sass():
    cuobjdump, cxxfilt = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump", shutil.which("c++filt")
    if not os.path.exists(cuobjdump) or not cxxfilt:
        pytest.skip("cuobjdump / c++filt not available")
    out = subprocess.run([cuobjdump, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, cur = collections.defaultdict(collections.Counter), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_x]+)*)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
    names = subprocess.run([cxxfilt] + list(per), capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", LIB], capture_output=True, text=True).stdout
    return {name: ops for name, ops in zip(names, per.values())}


def kernels(sass, needle):
    hit = {k: v for k, v in sass.items() if needle in k}
    assert hit, f"no kernel named *{needle}* in the library"
    return hit


def count(ops, prefix):
    return sum(n for op, n in ops.items() if op.startswith(prefix))


def test_f64_trailing_gemm_is_tma_fed_dmma_with_bulk_reduce_epilogue(sass):
    for name, ops in kernels(sass, "gemm_f64_tma_kernel<true>").items():
        assert count(ops, "UTMALDG.2D") >= 17 and count(ops, "DMMA.8x8x4") == 128, name
        assert count(ops, "UBLKRED.G.S.ADD.F64.RN") == 1, name          # the shipped epilogue (include/...)
    for name, ops in kernels(sass, "gemm_f64_tma_kernel<false>").items():
        assert count(ops, "UTMALDG.2D") >= 17 and count(ops, "DMMA.8x8x4") == 128 and count(ops, "UBLKRED") == 0, name


def test_f32_trailing_gemm_is_tcgen05_with_tmem(sass):
    for name, ops in kernels(sass, "gemm_f32_tc_kernel").items():
        assert count(ops, "UTCHMMA") >= 12 and count(ops, "LDTM") >= 1 and count(ops, "UTMALDG.2D") >= 5, name
        assert count(ops, "UTCBAR") >= 1, name


def test_triangular_solve_and_panel(sass):
    for name, ops in kernels(sass, "trsm_dmma_kernel").items():
        assert count(ops, "DMMA.8x8x4") >= 100, name
    pivoted = {k: v for k, v in kernels(sass, "panel_kernel<").items() if "nopiv" not in k}
    # warp argmax = three redux.sync (max, max, min) per reduction; sm_100a lowers them to CREDUX.MAX / CREDUX.MIN
    assert pivoted and all(count(ops, "CREDUX.MAX") >= 2 and count(ops, "CREDUX.MIN") >= 1 for ops in pivoted.values())


def test_no_local_memory_spills_in_the_gemm_kernels(sass):
    for needle in ("gemm_f64_tma_kernel", "gemm_f32_tc_kernel"):
        for name, ops in kernels(sass, needle).items():
            assert count(ops, "LDL") == 0 and count(ops, "STL") == 0, name
