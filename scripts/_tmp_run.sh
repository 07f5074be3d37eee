#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run18.log
: > $L
timeout 1200 python -W ignore -m pytest tests/test_gpu_em.py tests/test_gpu_model.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -4 | tee -a $L
for cfg in "cfg4 0.05" "cfg3 0.2" "cfg3 1.0" "cfg5 0.1"; do
  echo "== $cfg" | tee -a $L
  timeout 600 python -W ignore tests/tools/config_bench.py $cfg 2>&1 | grep -E "fit:|rror|assert|Trace" | cut -c1-250 | tee -a $L
done
