#!/bin/bash
# ncu --set full of the EM kernels of one cfg5 fit at x0.05; the report stays on the box (too large), the summaries come home
mkdir -p gpurun_out
python -W ignore scripts/profile_cfg.py cfg5 0.05 2 > gpurun_out/r2_p2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"em_" -c 40 -o /tmp/r2l_em -f python -W ignore scripts/profile_cfg.py cfg5 0.05 1 > gpurun_out/r2_p2_ncu.log 2>&1
tail -2 gpurun_out/r2_p2_plain.log
python scripts/ncu_summary.py /tmp/r2l_em.ncu-rep gpurun_out/r2l_cfg5_x0.05_em_ncu > /dev/null 2>&1
python scripts/ncu_hotlines.py /tmp/r2l_em.ncu-rep em_clane 45 > gpurun_out/r2l_clane_hotlines_all.txt 2>&1
ls -la /tmp/r2l_em.ncu-rep gpurun_out/
