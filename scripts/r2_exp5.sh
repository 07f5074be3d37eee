#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_exp5.log
: > $L
for cfg in "8 8" "1 1" "8 1" "1 2" "1 4"; do
  set -- $cfg
  export SS_SMALL_TILE_THR=$1 SS_MAX_CLUSTER_TILES=$2
  echo "== cfg5 0.1 SS_SMALL_TILE_THR=$1 SS_MAX_CLUSTER_TILES=$2" | tee -a $L
  timeout 600 python -W ignore tests/tools/config_bench.py cfg5 0.1 2>&1 | grep -E "layout|fit:|parity|OK|rror|assert" | tee -a $L
done
