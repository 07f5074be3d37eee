#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_x10.log
: > $L
timeout 600 python -m pytest tests/test_gpu_layout.py tests/test_gpu_em.py tests/test_gpu_model.py tests/test_gpu_scale.py -x -q -m gpu 2>&1 | tail -3 | tee -a $L
for lib in new; do
  if [ $lib = new ]; then unset STELLARSCOPE_B200_LIB; else export STELLARSCOPE_B200_LIB=$PWD/stellarscope_b200/variants/libss_base.so; fi
  echo "== $lib: bench cfg5 x0.25" | tee -a $L
  timeout 600 python -W ignore bench.py --steps 3 --warmup 2 --scale 0.25 --parity 20 --no-full-flow --cpu-budget-aln 200000 --no-ref-as-written > gpurun_out/r2_x10_$lib.json 2> gpurun_out/r2_x10_$lib.err; echo rc=$? | tee -a $L
  python - <<PY | tee -a $L
import json
l=json.loads(open('gpurun_out/r2_x10_$lib.json').read().strip().splitlines()[-1])
print('value %.1f G  ms/step %.1f  e2e %.3f s  build %.1f ms  breakdown %s' % (l['value']/1e9, l['ms_per_step'], l['e2e']['seconds_per_step'], l['config']['layout_build_ms'], l['e2e'].get('host_breakdown_rank0_s')))
PY
done
