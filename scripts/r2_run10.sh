#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run10.log
: > $L
timeout 1200 python -W ignore -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee -a $L
for cfg in cfg5 cfg2 cfg3 cfg4; do
  echo "== bench $cfg" | tee -a $L
  timeout 1500 python -W ignore bench.py --steps 5 --warmup 3 --config $cfg > gpurun_out/r2_bench_${cfg}_n1.json 2> gpurun_out/r2_bench_${cfg}_n1.err; echo rc=$? | tee -a $L
  tail -c 1500 gpurun_out/r2_bench_${cfg}_n1.err | tee -a $L
done
echo "== reference arm" | tee -a $L
timeout 900 python -W ignore bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo rc=$? | tee -a $L
