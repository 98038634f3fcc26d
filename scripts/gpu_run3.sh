#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== kernel microbench"
timeout 900 python scripts/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1; echo "rc=$?"; grep -v "^{" gpurun_out/bench_kernels.log | tail -45
echo "== panel T=256"
PANEL_ONLY=1 RFB_PANEL_THREADS=256 timeout 300 python scripts/bench_kernels.py 2>&1 | tail -2
echo "== bench 4096 / 16384"
timeout 600 python bench.py --n 4096 --steps 3 --warmup 2 --cpu-sample-n 4096 > gpurun_out/bench_4096.json 2> gpurun_out/bench_4096.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_4096.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['share_of_step_ms'], d['roofline']['achieved'])"; tail -5 gpurun_out/bench_4096.err
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_16384.json 2> gpurun_out/bench_16384.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_16384.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['share_of_step_ms'], d['roofline']['achieved'], d['clocks'], d['checks'])"; tail -5 gpurun_out/bench_16384.err
echo "== ncu panel (full set) m=8192"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 2 -c 1 -f -o gpurun_out/prof_panel python scripts/ncu_target.py lu 8192 > gpurun_out/ncu_panel.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_panel.log
