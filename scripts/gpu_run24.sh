#!/bin/bash
# run 24: pipelined panel exchange (forwarded header values, pivot row fetched one step late)
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k panel 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_lu.py tests/test_gpu_widened.py -q -m gpu -x 2>&1 | tail -4
PANEL_ONLY=1 timeout 300 python scripts/bench_kernels.py 2>&1 | tail -1
RFB_PANEL_PIPE=0 PANEL_ONLY=1 timeout 300 python scripts/bench_kernels.py 2>&1 | tail -1
for n in 4096 16384; do
timeout 600 python bench.py --n $n --steps 5 --warmup 3 --skip-cpu-baseline --skip-others --skip-e2e > gpurun_out/bench_${n}_run24.json 2> gpurun_out/bench_${n}_run24.err; echo "bench rc=$?"
done
python - <<'PY'
import json
for f in ('gpurun_out/bench_4096_run24.json','gpurun_out/bench_16384_run24.json'):
    try:
        d=json.load(open(f))
        print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['share_of_step_ms'], d['checks'])
    except Exception as e: print(f, 'failed', e)
PY
