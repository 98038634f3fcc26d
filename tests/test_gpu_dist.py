"""GPU tests of the multi-GPU path (C++ driver csrc/rfb_mg.cu behind rfb_mg_*).

* one process, G devices (rfb_mg_create_all / rfb_mg_lu_f64 -- the entry a Julia caller of lu! binds): runs with
  G = 1 on every box (the whole scheduler, replica, publish and node-end swap code is exercised; only the NCCL call
  is skipped) and with G = 2 / 4 where the box has them;
* one process per GPU under torchrun (rfb_mg_create_rank + ncclBroadcast): needs >= 2 GPUs."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import rfb200
from oracle import rf_oracle as O
from rfb200.dist_lu import MultiGpuLU
from util import assert_pivots_match, hutchinson_residual, rand_matrix

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
    return sum(1 for l in out.splitlines() if l.startswith("GPU "))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("g", [1, 2, 4])
@pytest.mark.parametrize("n,nb,dtype,zero_col", [(1024, 128, np.float64, -1), (1100, 192, np.float64, -1), (2048, 256, np.float64, 700),
                                                 (1536, 128, np.float32, -1), (640, 64, np.float64, -1), (1920, 64, np.float64, -1)])
def test_one_process_many_devices_matches_oracle(ctx, g, n, nb, dtype, zero_col):
    """rfb_mg_lu_* on a host matrix: same LU object as rfb200.lu_ -- pivots equal to the oracle's (bit-exact for
    Float64), info equal, ||PA - LU||_F / ||A||_F within 20 n eps; pivots also equal to the single-GPU path's."""
    if _ngpu() < g:
        pytest.skip(f"needs {g} GPUs")
    a0 = rand_matrix(np.random.default_rng([41, n, nb]), n, n, dtype)
    if zero_col >= 0:
        a0[:, zero_col] = 0
    mg = MultiGpuLU(g)
    try:
        a = a0.copy(order="F")
        F = mg.lu_(a, check=False, block=nb)
        assert F.factors is a
        _, want_p, want_info = O.lu_c(a0.copy(order="F"), threads=8)
        F1 = rfb200.lu(a0, check=False, ctx=ctx)
        assert F.info == want_info == F1.info
        if dtype == np.float64:
            assert np.array_equal(F.ipiv, want_p)
            assert np.array_equal(F.ipiv, F1.ipiv)
        else:
            assert_pivots_match(a0, F.factors, F.ipiv, want_p, strict=False)
        if want_info == 0:
            assert np.all(np.abs(np.tril(F.factors, -1)) <= 1.0)
            assert O.residual_fro_rel(a0, F.factors, F.ipiv) <= 20 * n * np.finfo(dtype).eps
        # a second factorization on the same handle reuses the buffers
        b = a0.copy(order="F")
        F2 = mg.lu_(b, check=False, block=nb)
        assert np.array_equal(F2.ipiv, F.ipiv) and np.array_equal(b, a)
    finally:
        mg.close()


def test_one_process_larger_matrix_properties(ctx):
    """8192^2 through the multi-GPU driver on all GPUs of the box (1 on the driver's boxes): pivots equal the
    single-GPU path's, residual probe within the bound."""
    g = max(1, min(_ngpu(), 8))
    n = 8192
    a0 = np.asfortranarray(np.random.default_rng(12).random((n, n)))
    mg = MultiGpuLU(g)
    try:
        a = a0.copy(order="F")
        F = mg.lu_(a, block=512)
        F1 = rfb200.lu(a0, ctx=ctx)
        assert np.array_equal(F.ipiv, F1.ipiv)
        assert hutchinson_residual(a0, F.factors, F.ipiv) <= 20 * n * np.finfo(np.float64).eps
    finally:
        mg.close()


@pytest.mark.parametrize("n,nb,dt,zero_col", [(2048, 256, "f64", -1), (3000, 192, "f64", -1), (2048, 128, "f32", -1), (1024, 128, "f64", 700)])
def test_distributed_matches_single_gpu(n, nb, dt, zero_col):
    g = _ngpu()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if g < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "dist_check.py"), str(n), str(nb), dt, str(zero_col)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"], out
