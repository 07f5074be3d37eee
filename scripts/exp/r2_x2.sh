#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_x2.log
: > $L
run() {
  echo "== $*" | tee -a $L
  env "${@:2}" timeout 300 python -W ignore tests/tools/config_bench.py cfg5 0.1 2>&1 | grep -E "layout|fit:|oracle|OK|rror|assert|Trace" | cut -c1-330 | tee -a $L
}
run carve100+t3072 SS_CLUSTER_TILE=3072
run carve0+t3072 SS_CARVEOUT=0 SS_CLUSTER_TILE=3072
run carve100+t3584 SS_CLUSTER_TILE=3584
echo "== clane phase profile (cycles per iteration), cfg5 x0.02" | tee -a $L
STELLARSCOPE_B200_LIB=$PWD/stellarscope_b200/variants/libss_prof.so SS_CLUSTER_TILE=3072 timeout 300 python -W ignore scripts/profile_cfg.py cfg5 0.02 1 2>&1 | grep -E "clane|rate" | sort | uniq | head -80 | tee -a $L
python tests/tools/empty_rank_check.py 2>&1 | tail -3 | tee -a $L
