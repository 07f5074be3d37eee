#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_p3_launches.csv python -W ignore scripts/profile_cfg.py cfg5 0.1 1 > gpurun_out/r2_p3_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tile_build_kernel" -c 1 -o /tmp/tb -f python -W ignore scripts/profile_cfg.py cfg5 0.1 1 > gpurun_out/r2_p3_b.log 2>&1
python scripts/ncu_summary.py /tmp/tb.ncu-rep gpurun_out/r2l_tile_build_ncu > /dev/null 2>&1
timeout 100 python scripts/ncu_hotlines.py /tmp/tb.ncu-rep tile_build 40 > gpurun_out/r2l_tile_build_hotlines.txt 2>&1
ls -la gpurun_out
