"""Single-GPU timing of the row-sharded (dense) pseudobulk path (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stellarscope_b200 import synth
from stellarscope_b200.engine import Engine
A = int(float(sys.argv[1])) if len(sys.argv) > 1 else 25_000_000
d = synth.generate(n_cells=5000, n_alignments=A, n_loci=28000, seed=0)
eng = Engine(0)
eng.set_sharding(0, 1, None, peers=True, K=d.K)
eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
eng.build_layout(np.zeros(d.N, dtype=np.int32), 1)
eng.enable_timing(True)
for i in range(3):
    eng.fit(); torch.cuda.synchronize()
    t = eng.last_timings()
res = eng.results(traces=False)
it = int(res['iters'][0])
print(f'dense pseudobulk {d.A} alignments: em {t["em"]:.3f} ms, {it} iterations, {d.A * it / t["em"] / 1e6:.1f} G aln*iter/s, '
      f'{12.0 * d.A * it / t["em"] / 1e6:.0f} GB/s algorithmic (12 B/aln)')
