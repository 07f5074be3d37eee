#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run11.log
: > $L
timeout 1200 python -W ignore -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee -a $L
for lib in default es1408 es1536 es2816; do
  if [ $lib = default ]; then unset STELLARSCOPE_B200_LIB; else export STELLARSCOPE_B200_LIB=$PWD/stellarscope_b200/variants/libss_$lib.so; fi
  for cfg in "cfg4 0.05" "cfg3 0.2"; do
    echo "== $lib $cfg" | tee -a $L
    timeout 600 python -W ignore tests/tools/config_bench.py $cfg 2>&1 | grep -E "fit:|rror|assert|Trace" | tee -a $L
  done
done
unset STELLARSCOPE_B200_LIB
echo "== e2e profile cfg5 report" | tee -a $L
timeout 900 python -W ignore scripts/e2e_profile.py cfg5 1.0 --report 2>&1 | grep -E "^call|^pass" | tee -a $L
