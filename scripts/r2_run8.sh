#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run8.log
: > $L
timeout 1200 python -W ignore -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee -a $L
echo "== e2e profile cfg2 report" | tee -a $L
timeout 600 python -W ignore scripts/e2e_profile.py cfg2 1.0 --report 2>&1 | grep -v "^$" | tail -45 | cut -c1-180 | tee -a $L
echo "== e2e profile cfg5 report" | tee -a $L
timeout 900 python -W ignore scripts/e2e_profile.py cfg5 1.0 --report 2>&1 | grep -v "^$" | cut -c1-180 | tee -a $L
