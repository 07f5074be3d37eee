#!/bin/bash
# round-2 final measurements, one GPU: full gpu tests, default bench (cfg5), reference arm, launch list at x0.1
mkdir -p gpurun_out
L=gpurun_out/r2_f1.log
: > $L
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee -a $L
echo "== bench cfg5 full (default flags)" | tee -a $L
timeout 1200 python -W ignore bench.py > gpurun_out/r2m_bench_cfg5_1gpu.json 2> gpurun_out/r2m_bench_cfg5_1gpu.err; echo rc=$? | tee -a $L
tail -c 600 gpurun_out/r2m_bench_cfg5_1gpu.err | tee -a $L
echo "== reference arm" | tee -a $L
timeout 600 python -W ignore bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_reference_arm.json 2> gpurun_out/r2m_bench_reference_arm.err; echo rc=$? | tee -a $L
echo "== launch list x0.1" | tee -a $L
timeout 600 python -W ignore bench.py --scale 0.1 --steps 2 --warmup 1 --parity 0 --no-full-flow --cpu-budget-aln 100000 --no-ref-as-written > gpurun_out/r2m_launch_plain.json 2>/dev/null &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2m_launches_bench_cfg5_x0.1.csv python -W ignore bench.py --scale 0.1 --steps 2 --warmup 1 --parity 0 --no-full-flow --cpu-budget-aln 100000 --no-ref-as-written > /dev/null 2>&1; echo rc=$? | tee -a $L
python - <<'PY' | tee -a $L
import json
l=json.loads(open('gpurun_out/r2m_bench_cfg5_1gpu.json').read().strip().splitlines()[-1])
print('value %.1f G  ms/step %.1f  frac %.3f  e2e %.3f s  build %.1f ms  launches %s' % (l['value']/1e9, l['ms_per_step'], l['roofline']['frac'], l['e2e']['seconds_per_step'], l['config']['layout_build_ms'], l['gpu_launches']))
print(l['parity_sample']['rank0'])
print(l.get('e2e_full'))
print(l['clocks'])
PY
