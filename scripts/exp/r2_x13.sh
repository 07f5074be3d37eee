#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_x13.log
: > $L
timeout 200 python -m pytest tests/test_gpu_em.py tests/test_gpu_model.py tests/test_gpu_scale.py -x -q -m gpu 2>&1 | tail -2 | tee -a $L
echo "== compact push loop, cfg5 x0.2" | tee -a $L
timeout 100 python -W ignore scripts/profile_cfg.py cfg5 0.2 3 2>&1 | grep -E "^[12] |rate" | sed 's/, .lane0.*resident/, resident/' | cut -c1-220 | tee -a $L
