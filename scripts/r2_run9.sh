#!/bin/bash
# 2 GPUs: sharded path against the goldens + strong-scaling bench at reduced scale
mkdir -p gpurun_out
L=gpurun_out/r2_run9.log
: > $L
timeout 900 python -W ignore -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/dist_check.py 2>&1 | grep -E "rank|OK|MISMATCH|rror|Trace|assert" | tail -60 | tee -a $L
echo "== bench cfg5 x0.2 N=2" | tee -a $L
timeout 900 python -W ignore -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --scale 0.2 --parity 40 > gpurun_out/r2_bench_cfg5_02_n2.json 2> gpurun_out/r2_bench_cfg5_02_n2.err; echo rc=$? | tee -a $L
tail -c 2500 gpurun_out/r2_bench_cfg5_02_n2.err | tee -a $L
echo "== bench cfg4 x0.1 N=2" | tee -a $L
timeout 900 python -W ignore -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --scale 0.1 --config cfg4 --no-full-flow > gpurun_out/r2_bench_cfg4_01_n2.json 2> gpurun_out/r2_bench_cfg4_01_n2.err; echo rc=$? | tee -a $L
tail -c 2500 gpurun_out/r2_bench_cfg4_01_n2.err | tee -a $L
