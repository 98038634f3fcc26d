"""Multi-GPU parity check (run under torchrun, one process per GPU):
distributed 1-D block-cyclic LU vs the single-GPU path on the same seeded matrix."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rfb200  # noqa: E402
from rfb200.dist_lu import DistributedLU, block_range  # noqa: E402


def block_data(n, nb, j, dtype, zero_col=-1):
    c0, w = block_range(j, n, nb)
    a = np.random.default_rng([12, j]).random((n, w), dtype=np.dtype(dtype).type)
    if c0 <= zero_col < c0 + w:
        a[:, zero_col - c0] = 0
    return np.asfortranarray(a)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    dtype = np.float64 if (len(sys.argv) <= 3 or sys.argv[3] == "f64") else np.float32
    zero_col = int(sys.argv[4]) if len(sys.argv) > 4 else -1
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    d = DistributedLU(n, dtype, block=nb)
    for j in d.my_blocks:
        d.set_block(j, block_data(n, nb, j, dtype, zero_col))
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(d.stream)
    d.factor()
    e1.record(d.stream)
    d.synchronize()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    info = d.info()
    full = d.gather_to(0)
    ipiv = d.pivots()
    out = {"n": n, "nb": nb, "world": world, "ms": ms, "info": info, "bcast_MB_per_rank": d.bcast_bytes / 1e6}
    if rank == 0:
        nblk = (n + nb - 1) // nb
        a0 = np.empty((n, n), dtype=dtype, order="F")
        for j in range(nblk):
            c0, w = block_range(j, n, nb)
            a0[:, c0:c0 + w] = block_data(n, nb, j, dtype, zero_col)
        F = rfb200.lu(a0, check=False, ctx=d.ctx)               # single-GPU path, same library
        out["info_single"] = F.info
        out["pivots_equal_single_gpu"] = bool(np.array_equal(ipiv, F.ipiv))
        out["max_abs_diff_vs_single_gpu"] = float(np.abs(full - F.factors).max())
        if info == 0:
            p = F.p
            eps = float(np.finfo(dtype).eps)
            l = np.tril(full.astype(np.float64), -1) + np.eye(n)
            u = np.triu(full.astype(np.float64))
            x = np.random.default_rng(0).integers(0, 2, size=(n, 4)).astype(np.float64) * 2 - 1
            pd = np.arange(n)
            for i, ip in enumerate(ipiv):
                ip = int(ip) - 1
                if ip != i:
                    pd[i], pd[ip] = pd[ip], pd[i]
            r = a0.astype(np.float64)[pd, :] @ x - l @ (u @ x)
            out["residual_fro_rel_est"] = float(np.linalg.norm(r) / 2.0 / np.linalg.norm(a0.astype(np.float64)))
            out["bound_20_n_eps"] = 20 * n * eps
        ok = out["pivots_equal_single_gpu"] and out["info"] == out["info_single"] and \
            (info != 0 or out["residual_fro_rel_est"] <= out["bound_20_n_eps"])
        out["ok"] = bool(ok)
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
