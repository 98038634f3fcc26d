#!/bin/bash
# run 19: compute-sanitizer (memcheck, racecheck, synccheck) over small cases of every kernel family
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_small.py 2>&1 | tail -3
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitizer_$tool.log 2>&1; echo rc=$?
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small|Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
echo "== cluster path under memcheck"
RFB_PANEL_CLUSTER=1 timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 10 python scripts/sanitize_small.py > gpurun_out/sanitizer_memcheck_cluster.log 2>&1; echo rc=$?
grep -E "ERROR SUMMARY|sanitize_small" gpurun_out/sanitizer_memcheck_cluster.log | head
