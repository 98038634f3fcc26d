#!/bin/bash
# run 13: NoPivot panel v2, skinny GEMM (vector solves), eager interchanges + early row downloads (e2e)
set -u
cd /root/repo
mkdir -p gpurun_out
echo "== widened + lu parity tests"
timeout 900 python -m pytest tests/test_gpu_widened.py tests/test_gpu_lu.py tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -15
echo "== widened bench"
timeout 900 python scripts/bench_widened.py > gpurun_out/bench_widened.log 2>&1; echo rc=$?; grep -v "^{" gpurun_out/bench_widened.log | tail -40
echo "== ncu: nopivot panel v2, skinny gemm"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_nopiv -s 0 -c 1 -f -o gpurun_out/prof_panel_nopiv_v2 python scripts/ncu_target.py nopiv 16384 > gpurun_out/ncu_nopiv.log 2>&1; echo rc=$?
echo "== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_16384_run13.json 2> gpurun_out/bench_16384_run13.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_16384_run13.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','checks','gpu_launches')})
print(d['roofline']['share_of_step_ms'], d['roofline']['achieved'], d['clocks'])
PY
