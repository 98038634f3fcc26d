#!/bin/bash
# run 25: final verification of the committed state (deferred window update): full suite, smoke, bench both arms, panel ncu
set -u
cd /root/repo
mkdir -p gpurun_out
echo "== full gpu suite"
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3
echo "== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_16384_run25.json 2> gpurun_out/bench_16384_run25.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_run25.json 2> gpurun_out/bench_reference_run25.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_16384_run25.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')})
print(d['roofline']['share_of_step_ms'], d['roofline']['achieved'], d['roofline']['frac'], d['checks'])
for k,v in d['other_configs'].items(): print(k, v)
print(json.load(open('gpurun_out/bench_reference_run25.json'))['value'])
PY
echo "== ncu panel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 1 -c 1 -f -o gpurun_out/prof_panel16k_run25 python scripts/ncu_target.py panel 16384 64 > gpurun_out/ncu_panel16k.log 2>&1; echo rc=$?
