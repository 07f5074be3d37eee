#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_f5.log
: > $L
timeout 600 python -m pytest tests/test_gpu_layout.py tests/test_gpu_em.py tests/test_gpu_model.py tests/test_gpu_scale.py -x -q -m gpu 2>&1 | tail -2 | tee -a $L
echo "== bench cfg5 full, short" | tee -a $L
timeout 900 python -W ignore bench.py --steps 4 --warmup 3 --parity 40 --no-full-flow --cpu-budget-aln 100000 --no-ref-as-written > gpurun_out/r2_f5_cfg5.json 2> gpurun_out/r2_f5_cfg5.err; echo rc=$? | tee -a $L
python - <<PY | tee -a $L
import json
l=json.loads(open('gpurun_out/r2_f5_cfg5.json').read().strip().splitlines()[-1])
print('value %.1f G  ms/step %.2f  e2e %.4f s  build %.1f ms  parity %s' % (l['value']/1e9, l['ms_per_step'], l['e2e']['seconds_per_step'], l['config']['layout_build_ms'], {k:v for k,v in (l['parity_sample'] or {}).items() if k not in ('what','rank0')}))
PY
