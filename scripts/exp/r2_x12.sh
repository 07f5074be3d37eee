#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_x12.log
: > $L
for rep in 1 2; do
for lib in pack nopack; do
  if [ $lib = pack ]; then unset STELLARSCOPE_B200_LIB; else export STELLARSCOPE_B200_LIB=$PWD/stellarscope_b200/variants/libss_nopack.so; fi
  echo "== $lib cfg5 x0.2" | tee -a $L
  timeout 300 python -W ignore scripts/profile_cfg.py cfg5 0.2 4 2>&1 | grep -E "^[123] " | sed 's/, .lane0.*resident/, resident/' | cut -c1-200 | tee -a $L
done
done
