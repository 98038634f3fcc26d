#!/bin/bash
set -u
cd /root/repo
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for g in 8 4; do
echo "== bench dist 32768 ($g GPUs)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus $g --steps 3 --warmup 1 > gpurun_out/bench_dist${g}_32768.json 2> gpurun_out/bench_dist$g.err; echo rc=$?; python - <<PY
import json
for line in open('gpurun_out/bench_dist${g}_32768.json'):
    if line.startswith('{'):
        d=json.loads(line); print('$g GPUs', d['value'], d['ms_per_step'], d['e2e'], d['checks'], d['gpu_launches'], d['clocks'])
PY
tail -3 gpurun_out/bench_dist$g.err | cut -c1-300
done
echo "== dist_check 8 GPUs 4096/256"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 scripts/dist_check.py 4096 256 f64 2>&1 | grep "^{" | cut -c1-500
