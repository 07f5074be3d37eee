#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run4.log
: > $L
echo "== e2e profile cfg5 full" | tee -a $L
timeout 900 python -W ignore scripts/e2e_profile.py cfg5 1.0 --report 2>&1 | grep -v "^$" | head -150 | tee -a $L
echo "== ncu stream cfg4 0.05" | tee -a $L
python -W ignore scripts/profile_cfg.py cfg4 0.05 1 > gpurun_out/r2_prof2_plain_cfg4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:em_stream_kernel -c 1 -o gpurun_out/r2b_stream_cfg4 -f python -W ignore scripts/profile_cfg.py cfg4 0.05 1 > gpurun_out/r2_prof2_ncu_cfg4.log 2>&1
tail -n 3 gpurun_out/r2_prof2_plain_cfg4.log | tee -a $L
