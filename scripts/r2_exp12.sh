#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_exp12.log
: > $L
for lib in default es1408 es2816; do
  if [ $lib = default ]; then unset STELLARSCOPE_B200_LIB; else export STELLARSCOPE_B200_LIB=$PWD/stellarscope_b200/variants/libss_$lib.so; fi
  for cfg in "cfg4 0.05" "cfg3 0.2"; do
    echo "== $lib $cfg" | tee -a $L
    timeout 600 python -W ignore tests/tools/config_bench.py $cfg 2>&1 | grep -E "layout|fit:|rror|assert|Trace" | cut -c1-250 | tee -a $L
  done
done
