#!/bin/bash
# First GPU pass: sanity + sanitizer on tiny cases, parity tests, short bench, launch list.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/host.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== sanitizer (memcheck) on small kernel tests"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "shape0 or shape1 or shape2 or shape3 or ties" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -8 gpurun_out/sanitizer.log
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== bench 4096"
timeout 600 python bench.py --n 4096 --steps 3 --warmup 2 --cpu-sample-n 4096 > gpurun_out/bench_4096.json 2> gpurun_out/bench_4096.err; echo "bench rc=$?"; cat gpurun_out/bench_4096.json; tail -5 gpurun_out/bench_4096.err
echo "== bench 16384"
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_16384.json 2> gpurun_out/bench_16384.err; echo "bench rc=$?"; cat gpurun_out/bench_16384.json; tail -5 gpurun_out/bench_16384.err
