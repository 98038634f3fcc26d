#!/bin/bash
# run 26: geometric upload chunks + shorter download tail; hypothesis GPU sweeps; final bench
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_property_sweep.py tests/test_gpu_lu.py -q -m gpu -x 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 --skip-others --skip-cpu-baseline > gpurun_out/bench_16384_run26.json 2> gpurun_out/bench_16384_run26.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_16384_run26.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e')}, d['checks'])
PY
