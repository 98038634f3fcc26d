#!/bin/bash
# run 21: source-level ncu of the panel kernel, single CTA (128 x 64) and 8 CTAs (1024 x 64)
set -u
cd /root/repo
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 1 -c 1 -f -o gpurun_out/prof_panel_g1 python scripts/ncu_target.py panel 128 64 > gpurun_out/ncu_panel_g1.log 2>&1; echo rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 1 -c 1 -f -o gpurun_out/prof_panel_g8 python scripts/ncu_target.py panel 1024 64 > gpurun_out/ncu_panel_g8.log 2>&1; echo rc=$?
