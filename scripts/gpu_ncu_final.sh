#!/bin/bash
# ncu evidence for the shipped K4 kernel (reduce-add epilogue): full capture of the root 8192^3 launch + launch list of one LU
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_tma_kernel -s 1 -c 1 -f -o gpurun_out/r2_gemm_reduce_root \
    python scripts/ncu_target.py gemm 8192 8192 8192 2 > gpurun_out/r2_ncu11.log 2>&1
echo "ncu full rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_lu16384_run11.csv \
    python scripts/ncu_target.py lu 16384 > gpurun_out/r2_ncu11b.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/r2_launches_lu16384_run11.csv
