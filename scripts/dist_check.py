"""Multi-GPU parity check (run under torchrun, one process per GPU): the C++ distributed LU (rfb_mg_*, NCCL broadcast)
vs the single-GPU path on the same seeded matrix.  torch.distributed (gloo, CPU) is only the launcher's transport: it
carries the 128-byte NCCL id once and gathers the results for the comparison; the data path is librfb200's."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rfb200  # noqa: E402
from rfb200.dist_lu import DistributedLU, block_range, owner_of  # noqa: E402


def block_data(n, nb, j, dtype, zero_col=-1):
    c0, w = block_range(j, n, nb)
    a = np.random.default_rng([12, j]).random((n, w), dtype=np.dtype(dtype).type)
    if c0 <= zero_col < c0 + w:
        a[:, zero_col - c0] = 0
    return np.asfortranarray(a)


def exchange(mine):
    box = [mine]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    dtype = np.float64 if (len(sys.argv) <= 3 or sys.argv[3] == "f64") else np.float32
    zero_col = int(sys.argv[4]) if len(sys.argv) > 4 else -1
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    d = DistributedLU(n, dtype, block=nb, rank=rank, world=world, device=local_rank, exchange_id=exchange)
    for j in d.my_blocks:
        d.set_block(j, block_data(n, nb, j, dtype, zero_col))
    d.synchronize()
    dist.barrier()
    d.factor()
    ms = d.synchronize()
    info = d.info()
    ipiv = d.pivots()
    # gather the factored block columns on rank 0 (CPU transport; test sizes only)
    full = torch.zeros((n, n), dtype=torch.float64)
    for j in d.my_blocks:
        c0, w = block_range(j, n, nb)
        full[:, c0:c0 + w] = torch.from_numpy(d.get_block(j).astype(np.float64))
    dist.all_reduce(full)
    full = np.asfortranarray(full.numpy().astype(dtype))
    allp = [None] * world
    dist.all_gather_object(allp, ipiv.tolist())
    out = {"n": n, "nb": nb, "world": world, "ms": ms, "info": info, "bcast_MB_per_rank": d.stats()["bcast_bytes_per_rank"] / 1e6,
           "pivots_same_on_every_rank": all(p == allp[0] for p in allp)}
    if rank == 0:
        nblk = (n + nb - 1) // nb
        a0 = np.empty((n, n), dtype=dtype, order="F")
        for j in range(nblk):
            c0, w = block_range(j, n, nb)
            a0[:, c0:c0 + w] = block_data(n, nb, j, dtype, zero_col)
        ctx = rfb200.Context(local_rank)
        F = rfb200.lu(a0, check=False, ctx=ctx)               # single-GPU path, same library
        out["info_single"] = F.info
        out["pivots_equal_single_gpu"] = bool(np.array_equal(ipiv, F.ipiv))
        out["max_abs_diff_vs_single_gpu"] = float(np.abs(full - F.factors).max())
        if info == 0:
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
        piv_ok = out["pivots_equal_single_gpu"] or dtype == np.float32     # Float32: near-ties may differ
        ok = piv_ok and out["pivots_same_on_every_rank"] and out["info"] == out["info_single"] and \
            (info != 0 or out["residual_fro_rel_est"] <= out["bound_20_n_eps"])
        out["ok"] = bool(ok)
        print(json.dumps(out), flush=True)
        ctx.close()
    d.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
