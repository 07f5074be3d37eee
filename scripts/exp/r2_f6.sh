#!/bin/bash
# end of round 2: full gpu tests, smoke, default bench
mkdir -p gpurun_out
L=gpurun_out/r2_f6.log
: > $L
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee -a $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a $L
echo "== bench (defaults)" | tee -a $L
timeout 1200 python -W ignore bench.py > gpurun_out/r2n_bench_cfg5_1gpu.json 2> gpurun_out/r2n_bench_cfg5_1gpu.err; echo rc=$? | tee -a $L
python - <<'PY' | tee -a $L
import json
l=json.loads(open('gpurun_out/r2n_bench_cfg5_1gpu.json').read().strip().splitlines()[-1])
print('value %.1f G  ms/step %.1f  frac %.3f  on_chip %.3f  e2e %.3f s = %.1f G  build %.1f ms  launches %s' % (l['value']/1e9, l['ms_per_step'], l['roofline']['frac'], l['roofline']['on_chip']['frac'], l['e2e']['seconds_per_step'], l['e2e']['value']/1e9, l['config']['layout_build_ms'], l['gpu_launches']))
print({k:v for k,v in l['parity_sample'].items() if k not in ('what','rank0')})
print(l.get('e2e_full'))
print(l['clocks'])
PY
