#!/bin/bash
# run 15: cluster/DSMEM panel exchange for <= 4096 rows, unpivoted panel v3
set -u
cd /root/repo
mkdir -p gpurun_out
echo "== parity tests"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lu.py tests/test_gpu_widened.py -q -m gpu -x 2>&1 | tail -15
echo "== panel us/col (cluster on / off)"
PANEL_ONLY=1 timeout 300 python scripts/bench_kernels.py 2>&1 | tail -2
RFB_PANEL_CLUSTER=0 PANEL_ONLY=1 timeout 300 python scripts/bench_kernels.py 2>&1 | tail -2
echo "== widened bench"
timeout 900 python scripts/bench_widened.py > gpurun_out/bench_widened.log 2>&1; echo rc=$?; grep -v "^{\"" gpurun_out/bench_widened.log | head -12
echo "== ncu: nopivot panel v3 + cluster panel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_nopiv -s 0 -c 1 -f -o gpurun_out/prof_panel_nopiv_v3 python scripts/ncu_target.py nopiv 16384 > gpurun_out/ncu_nopiv.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 0 -c 1 -f -o gpurun_out/prof_panel_cluster python scripts/ncu_target.py lu 4096 > gpurun_out/ncu_cluster.log 2>&1; echo rc=$?
echo "== bench 4096 and 16384"
timeout 600 python bench.py --n 4096 --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_4096_run15.json 2> gpurun_out/bench_4096_run15.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_16384_run15.json 2> gpurun_out/bench_16384_run15.err; echo "bench rc=$?"; python - <<'PY'
import json
for f in ('gpurun_out/bench_4096_run15.json','gpurun_out/bench_16384_run15.json'):
    d=json.load(open(f))
    print({k:d[k] for k in ('value','ms_per_step','e2e','checks','gpu_launches')})
    print(d['roofline']['share_of_step_ms'], d['roofline']['achieved'], d['clocks'])
PY
