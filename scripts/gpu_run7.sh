#!/bin/bash
set -u
cd /root/repo
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
mkdir -p gpurun_out
echo "== full pytest gpu (incl. 2-GPU dist tests)"
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench dist 32768 (2 GPUs)"
for blk in 512 1024; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 2 --warmup 1 --skip-e2e --block $blk > gpurun_out/bench_dist2_32768_b$blk.json 2> gpurun_out/bench_dist2.err; echo rc=$?; python - <<PY
import json
for line in open('gpurun_out/bench_dist2_32768_b$blk.json'):
    if line.startswith('{'):
        d=json.loads(line); print('block $blk', d['value'], d['ms_per_step'], d['checks'], d['gpu_launches'])
PY
done
echo "== 1-GPU 32768 for the strong-scaling baseline"
timeout 600 python bench.py --size 32768 --steps 2 --warmup 1 --skip-e2e --skip-cpu-baseline > gpurun_out/bench_1gpu_32768.json 2> gpurun_out/bench_1gpu_32768.err; python - <<PY
import json
for line in open('gpurun_out/bench_1gpu_32768.json'):
    if line.startswith('{'):
        d=json.loads(line); print('1gpu 32768', d['value'], d['ms_per_step'], d['roofline']['share_of_step_ms'])
PY
