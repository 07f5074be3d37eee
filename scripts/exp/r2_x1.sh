#!/bin/bash
# balanced cluster tiles + light cluster variant vs the committed library, cfg5 x0.1 (one GPU)
mkdir -p gpurun_out
L=gpurun_out/r2_x1.log
: > $L
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | tee -a $L
run() {  # name, env...
  echo "== $*" | tee -a $L
  env "${@:2}" timeout 300 python -W ignore tests/tools/config_bench.py cfg5 0.1 2>&1 | grep -E "layout|fit:|oracle|OK|rror|assert|Trace" | cut -c1-330 | tee -a $L
}
run base STELLARSCOPE_B200_LIB=$PWD/stellarscope_b200/variants/libss_base.so
run balanced SS_CLANE_LIGHT=0 SS_CLUSTER_TILE=4096
run balanced+light SS_CLANE_LIGHT=1 SS_CLUSTER_TILE=4096
run t3072+light SS_CLANE_LIGHT=1 SS_CLUSTER_TILE=3072
run t2560+light SS_CLANE_LIGHT=1 SS_CLUSTER_TILE=2560
run t3072 SS_CLANE_LIGHT=0 SS_CLUSTER_TILE=3072
python tests/tools/empty_rank_check.py 2>&1 | tail -3 | tee -a $L
