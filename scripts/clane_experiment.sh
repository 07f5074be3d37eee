#!/bin/bash
# Round-2 experiment (cluster lane kernel): 2048-alignment tiles for every multi-tile cell and
# 256-thread CTAs (two clusters' CTAs per SM) against the default (4096-alignment tiles, 512 threads).
# Run on the GPU box:  bash scripts/clane_experiment.sh   (about 3 minutes; logs under gpurun_out/)
mkdir -p gpurun_out
for cfg in "8 512" "1 512" "1 256"; do
  set -- $cfg
  export SS_SMALL_TILE_THR=$1 SS_CLANE_THREADS=$2
  echo "== SS_SMALL_TILE_THR=$1 SS_CLANE_THREADS=$2" | tee -a gpurun_out/clane_experiment.log
  python -W ignore -m pytest tests/test_gpu_em.py tests/test_gpu_layout.py -m gpu -x -q 2>&1 | tail -1 | tee -a gpurun_out/clane_experiment.log
  python -W ignore tests/tools/config_bench.py cfg5 0.1 2>&1 | grep -E "layout|fit:|parity|OK|Error|error" | tee -a gpurun_out/clane_experiment.log
done
