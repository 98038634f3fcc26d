#!/bin/bash
# One GPU call: the GPU test suite, then the 1-GPU bench (no 32768^2 side entry, no CPU baseline: both unchanged).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2_v8}
( time timeout 240 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
( time timeout 240 python bench.py --steps 3 --warmup 3 --skip-big --skip-cpu-baseline ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("ms", d["ms_per_step"], "e2e", d["e2e"], "roof", {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "peak_measured")}, d["checks"])
PY
