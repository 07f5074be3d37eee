#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_f4.log
: > $L
timeout 600 python -m pytest tests/test_gpu_layout.py tests/test_gpu_em.py tests/test_gpu_model.py tests/test_gpu_scale.py -x -q -m gpu 2>&1 | tail -2 | tee -a $L
for cfg in "cfg3 1.0" "cfg5 0.25" "cfg2 1.0"; do
  set -- $cfg
  echo "== bench $1 x$2" | tee -a $L
  timeout 600 python -W ignore bench.py --config $1 --scale $2 --steps 4 --warmup 3 --parity 10 --no-full-flow --cpu-budget-aln 100000 --no-ref-as-written > gpurun_out/r2_f4_$1.json 2> gpurun_out/r2_f4_$1.err; echo rc=$? | tee -a $L
  python - <<PY | tee -a $L
import json
l=json.loads(open('gpurun_out/r2_f4_$1.json').read().strip().splitlines()[-1])
print('value %.1f G  ms/step %.2f  e2e %.4f s  build %.1f ms' % (l['value']/1e9, l['ms_per_step'], l['e2e']['seconds_per_step'], l['config']['layout_build_ms']))
PY
done
