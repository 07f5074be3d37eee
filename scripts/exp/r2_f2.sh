#!/bin/bash
# two GPUs: parity of the sharded paths against the reference's goldens, strong-scaling bench of cfg5 at x0.5
mkdir -p gpurun_out
timeout 600 python -W ignore -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/tools/dist_check.py > gpurun_out/r2m_dist_check_2gpu.log 2>&1; echo dist_check rc=$?
tail -4 gpurun_out/r2m_dist_check_2gpu.log
timeout 600 python -W ignore -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tests/tools/empty_rank_check.py > gpurun_out/r2m_empty_rank_2gpu.log 2>&1; echo empty rc=$?
bash scripts/bench_multi.sh 2 --scale 0.5 --steps 4 --warmup 2 --parity 40 --cpu-budget-aln 200000 --no-ref-as-written
cp gpurun_out/bench_n2.json gpurun_out/r2m_bench_cfg5_x0.5_2gpu.json
