#!/bin/bash
mkdir -p gpurun_out
python -W ignore scripts/profile_cfg.py cfg5 0.1 2 > gpurun_out/r2_p1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"em_" -c 40 -o gpurun_out/r2l_cfg5_x0.1_em -f python -W ignore scripts/profile_cfg.py cfg5 0.1 1 > gpurun_out/r2_p1_ncu.log 2>&1
tail -3 gpurun_out/r2_p1_plain.log; tail -3 gpurun_out/r2_p1_ncu.log; ls -la gpurun_out/r2l_cfg5_x0.1_em.ncu-rep
