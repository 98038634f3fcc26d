#!/bin/bash
# One GPU call: A/B of the K4 variants, then the GPU test suite and the 1-GPU bench under the fastest bit-identical variant.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2_v10}
( time timeout 150 python scripts/ab_gemm_persist.py ) > gpurun_out/${TAG}_ab.log 2>&1
echo "ab rc=$?"; tail -4 gpurun_out/${TAG}_ab.log
eval "$(python - <<'PY'
import json
env = {"rmw": "", "reduce": "RFB_GEMM_EPILOGUE=1", "persist": "RFB_GEMM_EPILOGUE=1 RFB_GEMM_PERSIST=1",
       "persist_k1024": "RFB_GEMM_EPILOGUE=1 RFB_GEMM_PERSIST=1 RFB_GEMM_PERSIST_MAXK=1024",
       "persist_k2048_t149": "RFB_GEMM_EPILOGUE=1 RFB_GEMM_PERSIST=1 RFB_GEMM_PERSIST_MAXK=2048 RFB_GEMM_PERSIST_MINTILES=149"}
best = "rmw"
try:
    d = json.load(open("gpurun_out/ab_gemm_persist.json"))
    ok = all(d["bitwise"].values()) and all(v["bitwise_equal"] for v in d["lu"].values())
    if ok:
        lu = {k: v["ms"] for k, v in d["lu"]["16384"].items() if isinstance(v, dict)}
        best = min(lu, key=lu.get)
except Exception as e:
    print("echo 'no A/B result: %s'" % str(e).replace("'", ""))
print("echo 'chosen variant: %s'" % best)
for kv in env[best].split():
    print("export " + kv)
PY
)"
env | grep RFB_ || true
( time timeout 240 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
( time timeout 240 python bench.py --steps 3 --warmup 3 --skip-big --skip-cpu-baseline ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_row_bands"], "roof", {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "peak_measured", "share_of_step_ms")}, d["checks"])
PY
