#!/bin/bash
set -u
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_check.py 3000 192 f64 2>&1 | grep "^{" | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 2 --warmup 1 --skip-e2e > gpurun_out/bench_dist2_32768.json 2> gpurun_out/bench_dist2_32768.err; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench_dist2_32768.json')); print(d['value'], d['ms_per_step'], d['checks'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 --skip-e2e --block 256 > gpurun_out/bench_dist2_32768_b256.json 2>> gpurun_out/bench_dist2_32768.err; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/bench_dist2_32768_b256.json')); print('block256', d['value'], d['ms_per_step'], d['checks'])"
