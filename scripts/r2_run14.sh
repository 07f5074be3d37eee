#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run14.log
: > $L
timeout 1200 python -W ignore -m pytest tests/test_gpu_em.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -4 | tee -a $L
for thr in 512 1024; do
  export SS_CLANE_THREADS=$thr
  echo "== cfg5 0.1 SS_CLANE_THREADS=$thr" | tee -a $L
  timeout 600 python -W ignore tests/tools/config_bench.py cfg5 0.1 2>&1 | grep -E "fit:|rror|assert|Trace|parity" | cut -c1-250 | tee -a $L
done
unset SS_CLANE_THREADS
for cfg in "cfg4 0.05" "cfg3 0.2" "cfg3 1.0"; do
  echo "== $cfg" | tee -a $L
  timeout 600 python -W ignore tests/tools/config_bench.py $cfg 2>&1 | grep -E "fit:|rror|assert|Trace|parity" | cut -c1-250 | tee -a $L
done
