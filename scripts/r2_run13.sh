#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run13.log
: > $L
timeout 1200 python -W ignore -m pytest tests/test_gpu_em.py tests/test_gpu_model.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -8 | tee -a $L
for cfg in "cfg4 0.05" "cfg3 0.2" "cfg5 0.1"; do
  echo "== $cfg" | tee -a $L
  timeout 600 python -W ignore tests/tools/config_bench.py $cfg 2>&1 | grep -E "fit:|rror|assert|Trace|parity" | cut -c1-250 | tee -a $L
done
echo "== bench cfg3 full" | tee -a $L
timeout 900 python -W ignore bench.py --steps 5 --warmup 3 --config cfg3 --no-full-flow > gpurun_out/r2_bench_cfg3_n1b.json 2> gpurun_out/r2_bench_cfg3_n1b.err; echo rc=$? | tee -a $L
python -c "
import json; l=json.loads(open('gpurun_out/r2_bench_cfg3_n1b.json').read().strip().splitlines()[-1]); print('cfg3 value %.1f G ms/step %.2f frac %.3f'%(l['value']/1e9,l['ms_per_step'],l['roofline']['frac']))" | tee -a $L
echo "== launch list bench cfg5 x0.1" | tee -a $L
python -W ignore bench.py --steps 2 --warmup 3 --scale 0.1 --parity 0 --no-full-flow > gpurun_out/r2_launch_plain.json 2> gpurun_out/r2_launch_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches_bench_cfg5_x0.1.csv python -W ignore bench.py --steps 2 --warmup 3 --scale 0.1 --parity 0 --no-full-flow > gpurun_out/r2_launch_ncu.log 2>&1
tail -2 gpurun_out/r2_launch_ncu.log | cut -c1-300 | tee -a $L
