import os, sys, ctypes as C
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
def step(msg):
    torch.cuda.synchronize(); print(f"[{lr}] ok: {msg}", flush=True)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr)); step("init pg")
t = torch.ones(1024, device="cuda"); dist.broadcast(t, src=0); step("plain broadcast")
import rfb200
ctx = rfb200.Context(lr); step("ctx created")
z = torch.zeros(1 << 20, dtype=torch.float64, device="cuda"); step("zeros")
from rfb200.dist_lu import DistributedLU, block_range
d = DistributedLU(2048, np.float64, block=256, ctx=ctx); step("DistributedLU ctor")
for j in d.my_blocks:
    d.set_block(j, np.asfortranarray(np.random.default_rng([12, j]).random((2048, 256))))
step("set_block")
d.factor_block(0, 256) if dist.get_rank() == 0 else None; step("factor_block")
d.bcast_block(0, 256, 0); step("bcast_block")
d.factor(); step("factor")
print(lr, "info", d.info(), flush=True)
dist.destroy_process_group()
