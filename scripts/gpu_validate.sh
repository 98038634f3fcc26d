#!/bin/bash
# One GPU call: the GPU test suite, then the 1-GPU bench (no 32768^2 side entry, no CPU baseline).
#   gpurun --timeout 500 -- 'bash scripts/gpu_validate.sh r2_vNN'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2_v}
( time timeout 240 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
( time timeout 240 python bench.py --steps 3 --warmup 3 --skip-big --skip-cpu-baseline ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_row_bands"], "roof",
              {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "peak_measured", "share_of_step_ms")}, d["checks"])
PY
