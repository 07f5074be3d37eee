#!/bin/bash
# round 2, GPU call: full gpu tests + new bench at reduced and full scale (1 GPU)
mkdir -p gpurun_out
L=gpurun_out/r2_run2.log
: > $L
echo skip-pytest
echo "== bench cfg5 x0.05" | tee -a $L
timeout 900 python -W ignore bench.py --steps 3 --warmup 3 --scale 0.05 --parity 40 > gpurun_out/r2_bench_cfg5_005.json 2> gpurun_out/r2_bench_cfg5_005.err; echo rc=$? | tee -a $L
tail -c 3000 gpurun_out/r2_bench_cfg5_005.err | tee -a $L
echo "== bench cfg5 full" | tee -a $L
timeout 1500 python -W ignore bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg5_full.json 2> gpurun_out/r2_bench_cfg5_full.err; echo rc=$? | tee -a $L
tail -c 3000 gpurun_out/r2_bench_cfg5_full.err | tee -a $L
