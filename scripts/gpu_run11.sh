#!/bin/bash
set -u
cd /root/repo
mkdir -p gpurun_out
echo "== ncu full: root GEMM 8192^3 (TMA + DMMA kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_tma -s 1 -c 1 -f -o gpurun_out/prof_gemm_root python scripts/ncu_target.py gemm 8192 8192 8192 2 > gpurun_out/ncu_gemm_root.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_gemm_root.log
echo "== ncu full: tcgen05 f32 GEMM 4096^3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f32_tc -s 1 -c 1 -f -o gpurun_out/prof_gemm_tc32 python scripts/ncu_target.py gemm32 4096 4096 4096 > gpurun_out/ncu_gemm_tc32.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_gemm_tc32.log
echo "== ncu full: panel kernel (m=16384 first panel) + trsm block + laswp list"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 0 -c 1 -f -o gpurun_out/prof_panel16k python scripts/ncu_target.py lu 16384 > gpurun_out/ncu_panel16k.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none -k regex:trsm_block -s 200 -c 1 -f -o gpurun_out/prof_trsm_block python scripts/ncu_target.py lu 16384 > gpurun_out/ncu_trsm.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none -k regex:laswp_list -s 506 -c 1 -f -o gpurun_out/prof_laswp python scripts/ncu_target.py lu 16384 > gpurun_out/ncu_laswp.log 2>&1; echo rc=$?
echo "== ncu launch list of the bench command"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench16384.csv python bench.py --steps 1 --warmup 1 --skip-e2e --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo rc=$?; wc -l gpurun_out/launches_bench16384.csv
echo "== final bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_16384.json 2> gpurun_out/bench_16384.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_16384.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_reference.json
