"""GPU test of the multi-GPU path (needs >= 2 GPUs; the single-GPU boxes skip it):
torchrun with one process per GPU, distributed LU vs the single-GPU path on the same matrix."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n,nb,dt,zero_col", [(2048, 256, "f64", -1), (3000, 192, "f64", -1), (2048, 128, "f32", -1), (1024, 128, "f64", 700)])
def test_distributed_matches_single_gpu(n, nb, dt, zero_col):
    g = _ngpu()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if g < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "scripts", "dist_check.py"), str(n), str(nb), dt, str(zero_col)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"], out
