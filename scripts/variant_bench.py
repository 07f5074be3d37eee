"""Kernel-variant comparison (development aid): configs[1] workload, device time of the EM kernels.
   STELLARSCOPE_B200_LIB=<variant .so> python scripts/variant_bench.py [n_fits]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from stellarscope_b200 import synth
from stellarscope_b200.engine import Engine

n_fits = int(sys.argv[1]) if len(sys.argv) > 1 else 6
d = synth.generate(n_cells=10000, n_alignments=5_000_000, n_loci=28000, seed=0)
eng = Engine(0)
eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
eng.build_layout(d.row_cell, d.n_cells)
eng.enable_timing(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
best = None
for i in range(n_fits):
    flush.fill_(1)
    eng.fit()
    torch.cuda.synchronize()
    t = eng.last_timings()
    if i > 0 and (best is None or t['em'] < best['em']):
        best = t
res = eng.results(traces=False)
segA = eng.export_layout_array('seg_A', np.int64)
units = float((segA * res['iters']).sum())
name = os.path.basename(os.environ.get('STELLARSCOPE_B200_LIB', 'default'))
print(f'{name}: em {best["em"]:.3f} ms  rate {units / best["em"] / 1e6:.1f} G aln*iter/s  resident {best["resident"]:.3f} cluster {best["cluster"]:.3f} '
      + ' '.join(f'{k[4:]}:{v:.2f}' for k, v in best.items() if k.startswith('lane')), flush=True)
