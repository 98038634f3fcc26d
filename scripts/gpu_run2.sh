#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== kernel microbench"
timeout 900 python scripts/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1; echo "rc=$?"; tail -45 gpurun_out/bench_kernels.log
echo "== bench 4096 / 16384"
timeout 600 python bench.py --n 4096 --steps 3 --warmup 2 --cpu-sample-n 4096 > gpurun_out/bench_4096.json 2> gpurun_out/bench_4096.err; echo "bench rc=$?"; cat gpurun_out/bench_4096.json; tail -5 gpurun_out/bench_4096.err
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_16384.json 2> gpurun_out/bench_16384.err; echo "bench rc=$?"; cat gpurun_out/bench_16384.json; tail -5 gpurun_out/bench_16384.err
echo "== ncu panel (full set)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 8 -c 2 -f -o gpurun_out/prof_panel python scripts/ncu_target.py lu 4096 > gpurun_out/ncu_panel.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_panel.log
echo "== ncu gemm tma (full set)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_tma -s 1 -c 1 -f -o gpurun_out/prof_gemm_tma python scripts/ncu_target.py gemm 8192 8192 2048 2 > gpurun_out/ncu_gemm.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_gemm.log
echo "== ncu launch list (LU 4096)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lu4096.csv python scripts/ncu_target.py lu 4096 > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches_lu4096.csv
