#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run7.log
: > $L
timeout 900 python -W ignore -m pytest tests/test_gpu_count.py tests/test_gpu_model.py tests/test_mtx.py -m gpu -x -q 2>&1 | tail -15 | tee -a $L
echo "== e2e profile cfg2 report" | tee -a $L
timeout 600 python -W ignore scripts/e2e_profile.py cfg2 1.0 --report 2>&1 | grep -E "^call|^pass" | tee -a $L
echo "== e2e profile cfg5 report" | tee -a $L
timeout 900 python -W ignore scripts/e2e_profile.py cfg5 1.0 --report 2>&1 | grep -v "^$" | head -120 | tee -a $L
