#!/bin/bash
# round-2 experiment 1: merged stream pass (header staged) + cluster-kernel knobs
mkdir -p gpurun_out
L=gpurun_out/r2_exp1.log
: > $L
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | tee -a $L
nproc | tee -a $L; free -g | head -2 | tee -a $L
timeout 900 python -W ignore -m pytest tests/test_gpu_em.py tests/test_gpu_layout.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3 | tee -a $L
for cfg in "cfg4 0.05" "cfg3 0.2"; do
  echo "== $cfg" | tee -a $L
  timeout 600 python -W ignore tests/tools/config_bench.py $cfg 2>&1 | grep -E "layout|fit:|parity|OK|rror|assert" | tee -a $L
done
for cfg in "8 512" "1 512" "1 256"; do
  set -- $cfg
  export SS_SMALL_TILE_THR=$1 SS_CLANE_THREADS=$2
  echo "== cfg5 0.1 SS_SMALL_TILE_THR=$1 SS_CLANE_THREADS=$2" | tee -a $L
  timeout 600 python -W ignore tests/tools/config_bench.py cfg5 0.1 2>&1 | grep -E "layout|fit:|parity|OK|rror|assert" | tee -a $L
done
