#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run15.log
: > $L
for iv in 0.02 0.25 5; do
  echo "== bench cfg3 clock interval $iv" | tee -a $L
  timeout 900 python -W ignore bench.py --steps 6 --warmup 3 --config cfg3 --no-full-flow --parity 0 --cpu-budget-aln 200000 --no-ref-as-written --clock-interval $iv > gpurun_out/r2_b15.json 2> gpurun_out/r2_b15.err; echo rc=$? | tee -a $L
  python -c "
import json; l=json.loads(open('gpurun_out/r2_b15.json').read().strip().splitlines()[-1]); print('cfg3 value %.1f G ms/step %.2f em(last) %.2f samples %s'%(l['value']/1e9,l['ms_per_step'],l['config']['kernel_ms_rank0']['em'], l['clocks']))" | tee -a $L
done
