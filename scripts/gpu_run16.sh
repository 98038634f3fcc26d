#!/bin/bash
# run 16: live-width chunked window loops in the panel kernel
set -u
cd /root/repo
mkdir -p gpurun_out
echo "== parity tests"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lu.py tests/test_gpu_widened.py -q -m gpu -x 2>&1 | tail -5
RFB_PANEL_CLUSTER=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k panel 2>&1 | tail -3
echo "== panel us/col (cluster off / on)"
PANEL_ONLY=1 timeout 300 python scripts/bench_kernels.py 2>&1 | tail -1
RFB_PANEL_CLUSTER=1 PANEL_ONLY=1 timeout 300 python scripts/bench_kernels.py 2>&1 | tail -1
echo "== batched"
timeout 900 python scripts/bench_widened.py 2>&1 | grep "^batched\|pivot\|nopiv "
echo "== bench 4096 and 16384"
timeout 600 python bench.py --n 4096 --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_4096_run16.json 2> gpurun_out/bench_4096_run16.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_16384_run16.json 2> gpurun_out/bench_16384_run16.err; echo "bench rc=$?"; python - <<'PY'
import json
for f in ('gpurun_out/bench_4096_run16.json','gpurun_out/bench_16384_run16.json'):
    d=json.load(open(f))
    print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')})
    print(d['roofline']['share_of_step_ms'], d['roofline']['achieved'], d['checks'])
PY
