#!/bin/bash
# run 14: speculative reciprocal in the pivoted / unpivoted / batched panel kernels
set -u
cd /root/repo
mkdir -p gpurun_out
echo "== parity tests"
timeout 900 python -m pytest tests/test_gpu_widened.py tests/test_gpu_lu.py tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -15
echo "== panel us/col"
PANEL_ONLY=1 timeout 300 python scripts/bench_kernels.py 2>&1 | tail -3
echo "== widened bench"
timeout 900 python scripts/bench_widened.py > gpurun_out/bench_widened.log 2>&1; echo rc=$?; grep -v "^{\"" gpurun_out/bench_widened.log | tail -40
echo "== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_16384_run14.json 2> gpurun_out/bench_16384_run14.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_16384_run14.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','checks','gpu_launches')})
print(d['roofline']['share_of_step_ms'], d['roofline']['achieved'], d['clocks'])
PY
