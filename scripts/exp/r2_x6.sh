#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_x6.log
: > $L
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee -a $L
echo "== bench cfg5 full" | tee -a $L
timeout 900 python -W ignore bench.py --steps 5 --warmup 3 > gpurun_out/r2_x6_bench_cfg5.json 2> gpurun_out/r2_x6_bench_cfg5.err; echo rc=$? | tee -a $L
tail -c 1500 gpurun_out/r2_x6_bench_cfg5.err | tee -a $L
python - <<'PY' | tee -a $L
import json
l=json.loads(open('gpurun_out/r2_x6_bench_cfg5.json').read().strip().splitlines()[-1])
print('value %.1f G  ms/step %.1f  frac %.3f  e2e %.3f s  build %.1f ms  parity %s' % (l['value']/1e9, l['ms_per_step'], l['roofline']['frac'], l['e2e']['seconds_per_step'], l['config']['layout_build_ms'], {k: l['parity_sample'][k] for k in ('cells','counts_equal','iters_equal','lnl_1e10','pi_1e6')}))
print({k: round(v,1) for k,v in l['config']['kernel_ms_rank0'].items() if not k.startswith('lane')})
print(l.get('e2e_full'))
PY
