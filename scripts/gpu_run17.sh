#!/bin/bash
# run 17: row-parallel triangular block kernel for vector solves; bench with other_configs; launch list
set -u
cd /root/repo
mkdir -p gpurun_out
echo "== full gpu suite"
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -4
echo "== widened bench"
timeout 900 python scripts/bench_widened.py > gpurun_out/bench_widened.log 2>&1; echo rc=$?; grep "^ldiv\|butterfly_solve\|butterfly_ms" gpurun_out/bench_widened.log
echo "== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_16384_run17.json 2> gpurun_out/bench_16384_run17.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_16384_run17.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')})
print(d['roofline']['share_of_step_ms'], d['roofline']['achieved'], d['checks'])
for k,v in d['other_configs'].items(): print(k, v)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_run17.json 2> gpurun_out/bench_reference_run17.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_reference_run17.json
echo "== ncu launch list of the bench command"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench16384_run17.csv python bench.py --steps 1 --warmup 1 --skip-e2e --skip-cpu-baseline --skip-others > gpurun_out/bench_under_ncu.log 2>&1; echo rc=$?; wc -l gpurun_out/launches_bench16384_run17.csv
