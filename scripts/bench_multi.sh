#!/bin/bash
# strong-scaling bench under torchrun on N GPUs of one box:  gpurun --gpus N -- 'bash scripts/bench_multi.sh N [extra bench args]'
N=${1:-2}; shift
mkdir -p gpurun_out
python -W ignore -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo rc=$?
tail -c 1500 gpurun_out/bench_n$N.err
python - <<PY
import json
l=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N=%d value %.1f G ms/step %.1f e2e %.2f G (%.3f s/step) h2d/step %.2f GB'%(l['n_gpus'],l['value']/1e9,l['ms_per_step'],l['e2e']['value']/1e9,l['e2e']['seconds_per_step'],l['e2e']['h2d_bytes_per_step']/1e9))
print('host breakdown rank0', l['e2e'].get('host_breakdown_rank0_s'))
print('parity', {k:v for k,v in (l['parity_sample'] or {}).items() if k not in ('what','rank0')})
print('build ms', l['config']['layout_build_ms'], 'gen s', l['config']['generate_s'])
PY
