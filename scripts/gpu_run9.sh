#!/bin/bash
set -u
cd /root/repo
mkdir -p gpurun_out
echo "== trsm tests"
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "trsm" --timeout 300 2>&1 | tail -5
echo "== pytest gpu (all)"
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== kernel microbench"
timeout 900 python scripts/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1; echo "rc=$?"; grep -E "^trsm|^panel" gpurun_out/bench_kernels.log
echo "== bench 4096 / 16384"
timeout 600 python bench.py --size 4096 --steps 3 --warmup 2 --cpu-sample-n 4096 > gpurun_out/bench_4096.json 2> gpurun_out/bench_4096.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_4096.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['share_of_step_ms'], d['roofline']['achieved'])"
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_16384.json 2> gpurun_out/bench_16384.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_16384.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['share_of_step_ms'], d['roofline']['achieved'], d['clocks'])"
