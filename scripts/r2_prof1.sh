#!/bin/bash
mkdir -p gpurun_out
python -W ignore scripts/profile_cfg.py cfg4 0.05 1 > gpurun_out/r2_prof1_plain_cfg4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:em_stream_kernel -c 1 -o gpurun_out/r2_stream_cfg4 -f python -W ignore scripts/profile_cfg.py cfg4 0.05 1 > gpurun_out/r2_prof1_ncu_cfg4.log 2>&1
python -W ignore scripts/profile_cfg.py cfg5 0.03 1 > gpurun_out/r2_prof1_plain_cfg5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"em_clane_kernel|em_lane_kernel<512|em_stream_kernel" -c 9 -o gpurun_out/r2_cfg5_003 -f python -W ignore scripts/profile_cfg.py cfg5 0.03 1 > gpurun_out/r2_prof1_ncu_cfg5.log 2>&1
tail -3 gpurun_out/r2_prof1_plain_cfg4.log gpurun_out/r2_prof1_plain_cfg5.log
tail -3 gpurun_out/r2_prof1_ncu_cfg4.log gpurun_out/r2_prof1_ncu_cfg5.log
