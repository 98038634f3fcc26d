#!/bin/bash
# run 12: first GPU pass over the widened rows (NoPivot, butterfly, batched) + full parity suite
set -u
cd /root/repo
mkdir -p gpurun_out
echo "== new parity tests"
timeout 900 python -m pytest tests/test_gpu_widened.py -q -m gpu -x 2>&1 | tail -15
echo "== full gpu suite"
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5
echo "== widened bench"
timeout 900 python scripts/bench_widened.py > gpurun_out/bench_widened.log 2>&1; echo rc=$?; tail -25 gpurun_out/bench_widened.log
echo "== ncu: nopivot panel, butterfly, batched"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_nopiv -s 0 -c 1 -f -o gpurun_out/prof_panel_nopiv python scripts/ncu_target.py nopiv 16384 > gpurun_out/ncu_nopiv.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:butterfly_mul -s 0 -c 1 -f -o gpurun_out/prof_butterfly_mul python scripts/ncu_target.py nopiv 16384 > gpurun_out/ncu_bfly.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 0 -c 1 -f -o gpurun_out/prof_batched python scripts/ncu_target.py batched 16384 32 > gpurun_out/ncu_batched.log 2>&1; echo rc=$?
echo "== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_16384_run12.json 2> gpurun_out/bench_16384_run12.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_16384_run12.json
