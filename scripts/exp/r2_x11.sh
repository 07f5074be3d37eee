#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_x11.log
: > $L
for occ in 0 1; do
  for sc in 0.1 0.3; do
    echo "== SS_STREAM_OCC=$occ cfg5 x$sc" | tee -a $L
    SS_STREAM_OCC=$occ timeout 300 python -W ignore scripts/profile_cfg.py cfg5 $sc 3 2>&1 | grep -E "^2 |rate" | cut -c1-420 | tee -a $L
  done
done
