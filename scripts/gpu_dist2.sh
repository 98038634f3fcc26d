#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== dist_check 2048/256 f64"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2048 256 f64 2>&1 | tail -5
echo "== dist_check 3000/192 f64"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_check.py 3000 192 f64 2>&1 | tail -3
echo "== dist_check 8192/512 f64"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dist_check.py 8192 512 f64 2>&1 | tail -3
echo "== dist_check 1024/128 singular col"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/dist_check.py 1024 128 f64 700 2>&1 | tail -3
echo "== bench dist 16384 (2 GPUs)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 2 --warmup 1 --size 16384 > gpurun_out/bench_dist2_16384.json 2> gpurun_out/bench_dist2_16384.err; echo rc=$?; cat gpurun_out/bench_dist2_16384.json | cut -c1-1500; tail -5 gpurun_out/bench_dist2_16384.err
echo "== bench dist 32768 (2 GPUs)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_dist2_32768.json 2> gpurun_out/bench_dist2_32768.err; echo rc=$?; cat gpurun_out/bench_dist2_32768.json | cut -c1-1500; tail -5 gpurun_out/bench_dist2_32768.err
