#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run6.log
: > $L
timeout 900 python -W ignore -m pytest tests/test_gpu_em.py tests/test_gpu_layout.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -15 | tee -a $L
for mode in 1 0; do
  export SS_CELL_MODE=$mode
  echo "== cfg5 0.1 SS_CELL_MODE=$mode" | tee -a $L
  timeout 600 python -W ignore tests/tools/config_bench.py cfg5 0.1 2>&1 | grep -E "layout|fit:|parity|OK|rror|assert|Trace" | tee -a $L
done
unset SS_CELL_MODE
for cfg in "cfg4 0.05" "cfg3 0.2" "cfg2 1.0"; do
  echo "== $cfg" | tee -a $L
  timeout 600 python -W ignore tests/tools/config_bench.py $cfg 2>&1 | grep -E "layout|fit:|parity|OK|rror|assert|Trace" | tee -a $L
done
