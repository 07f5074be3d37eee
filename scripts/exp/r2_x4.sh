#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_x4.log
: > $L
run() {
  echo "== $*" | tee -a $L
  env "${@:2}" timeout 300 python -W ignore tests/tools/config_bench.py cfg5 0.1 2>&1 | grep -E "layout|fit:|oracle|OK|rror|assert|Trace" | cut -c1-330 | tee -a $L
}
run split+t3072 SS_CLUSTER_TILE=3072
run split+t4096 SS_CLUSTER_TILE=4096
run split+t3072-nolight SS_CLUSTER_TILE=3072 SS_CLANE_LIGHT=0
timeout 600 python -m pytest tests/test_gpu_em.py tests/test_gpu_model.py tests/test_gpu_layout.py -x -q -m gpu 2>&1 | tail -5 | tee -a $L
echo "== clane phase profile (cycles per iteration), cfg5 x0.02" | tee -a $L
STELLARSCOPE_B200_LIB=$PWD/stellarscope_b200/variants/libss_prof.so SS_CLUSTER_TILE=3072 timeout 300 python -W ignore scripts/profile_cfg.py cfg5 0.02 1 2>&1 | grep -E "clane|rate" | sort | uniq | head -40 | tee -a $L
