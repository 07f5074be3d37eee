#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -W ignore -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/tools/dist_check.py > gpurun_out/r2n_dist_check_2gpu.log 2>&1; echo dist_check rc=$?
grep -c OK gpurun_out/r2n_dist_check_2gpu.log; grep -i -E "error|Traceback|FAIL" gpurun_out/r2n_dist_check_2gpu.log | head -5
bash scripts/bench_multi.sh 2 --scale 0.2 --steps 3 --warmup 2 --parity 20 --no-full-flow --cpu-budget-aln 100000 --no-ref-as-written
