#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -W ignore bench.py --steps 3 --warmup 3 --parity 40 --no-full-flow --cpu-budget-aln 50000 --no-ref-as-written > gpurun_out/r2n_bench_cfg5_1gpu_final_short.json 2> gpurun_out/r2n_bench_cfg5_1gpu_final_short.err; echo rc=$?
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2n_bench_cfg5_1gpu_final_short.json').read().strip().splitlines()[-1])
print('value %.1f G  ms/step %.2f  frac %.3f  e2e %.4f s  build %.1f ms  parity %s' % (l['value']/1e9, l['ms_per_step'], l['roofline']['frac'], l['e2e']['seconds_per_step'], l['config']['layout_build_ms'], {k:v for k,v in (l['parity_sample'] or {}).items() if k not in ('what','rank0')}))
PY
