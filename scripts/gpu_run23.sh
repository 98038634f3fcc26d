#!/bin/bash
# run 23: deferred window update (rank-1 update in the shadow of the header round trip)
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lu.py tests/test_gpu_widened.py -q -m gpu -x 2>&1 | tail -3
RFB_PANEL_CLUSTER=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k panel 2>&1 | tail -2
PANEL_ONLY=1 timeout 300 python scripts/bench_kernels.py 2>&1 | tail -1
RFB_PANEL_CLUSTER=1 PANEL_ONLY=1 timeout 300 python scripts/bench_kernels.py 2>&1 | tail -1
for n in 4096 16384; do
timeout 600 python bench.py --n $n --steps 5 --warmup 3 --skip-cpu-baseline --skip-others --skip-e2e > gpurun_out/bench_${n}_run23.json 2> gpurun_out/bench_${n}_run23.err; echo "bench rc=$?"
done
python - <<'PY'
import json
for f in ('gpurun_out/bench_4096_run23.json','gpurun_out/bench_16384_run23.json'):
    d=json.load(open(f))
    print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['share_of_step_ms'], d['checks'])
PY
