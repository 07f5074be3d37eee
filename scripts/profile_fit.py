"""Profiling driver (ncu): configs[1] workload, a few EM fits.  python scripts/profile_fit.py [n_fits] [scale]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from stellarscope_b200 import synth
from stellarscope_b200.engine import Engine

n_fits = int(sys.argv[1]) if len(sys.argv) > 1 else 3
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
mode = sys.argv[3] if len(sys.argv) > 3 else 'individual'
d = synth.generate(n_cells=int(10000 * scale), n_alignments=int(5_000_000 * scale), n_loci=28000, seed=0)
eng = Engine(0)
eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
if mode == 'individual':
    eng.build_layout(d.row_cell, d.n_cells)
else:
    eng.build_layout(np.zeros(d.N, dtype=np.int32), 1)
eng.enable_timing(True)
print(eng.layout_info())
for i in range(n_fits):
    eng.fit()
    torch.cuda.synchronize()
    print(i, eng.last_timings())
res = eng.results(traces=False)
print('iters mean', res['iters'].mean(), 'max', res['iters'].max(), 'launches', eng.launch_count())
