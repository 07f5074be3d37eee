#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_f3.log
: > $L
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee -a $L
for cfg in cfg2 cfg3 cfg4; do
  echo "== bench $cfg" | tee -a $L
  timeout 900 python -W ignore bench.py --config $cfg --steps 6 --warmup 3 --cpu-budget-aln 500000 --no-ref-as-written > gpurun_out/r2m_bench_${cfg}_1gpu.json 2> gpurun_out/r2m_bench_${cfg}_1gpu.err; echo rc=$? | tee -a $L
  python - <<PY | tee -a $L
import json
l=json.loads(open('gpurun_out/r2m_bench_${cfg}_1gpu.json').read().strip().splitlines()[-1])
print('value %.1f G  ms/step %.2f  frac %.3f  e2e %.4f s = %.1f G  build %.1f ms  e2e_full %s  parity %s' % (l['value']/1e9, l['ms_per_step'], l['roofline']['frac'], l['e2e']['seconds_per_step'], l['e2e']['value']/1e9, l['config']['layout_build_ms'], (l.get('e2e_full') or {}).get('seconds'), {k:v for k,v in (l['parity_sample'] or {}).items() if k not in ('what','rank0')}))
PY
done
