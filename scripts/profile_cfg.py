"""Profiling driver (ncu): one of the BASELINE.json configurations at a scale, a few EM fits on ONE GPU.

    python scripts/profile_cfg.py cfg2|cfg3|cfg4|cfg5 [scale] [n_fits]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from stellarscope_b200 import synth
from stellarscope_b200.engine import Engine

CFG = {
    'cfg2': dict(n_cells=10_000, n_alignments=5_000_000, mode='individual'),
    'cfg3': dict(n_cells=50_000, n_alignments=100_000_000, mode='celltype', n_celltypes=20, umi_dup_rate=0.15),
    'cfg4': dict(n_cells=100_000, n_alignments=500_000_000, mode='pseudobulk'),
    'cfg5': dict(n_cells=100_000, n_alignments=500_000_000, mode='individual'),
}
name = sys.argv[1]
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
n_fits = int(sys.argv[3]) if len(sys.argv) > 3 else 1
c = dict(CFG[name])
mode = c.pop('mode')
c['n_cells'] = max(1, int(c['n_cells'] * scale))
c['n_alignments'] = int(c['n_alignments'] * scale)
d = synth.generate_device(n_loci=28_000, seed=0, **c)
if c.get('umi_dup_rate'):
    d = synth.apply_row_mask(d, synth.umi_dedup_mask(d))
if mode == 'individual':
    row_seg, S = d.row_cell, d.n_cells
elif mode == 'celltype':
    row_seg, S = d.cell_type[d.row_cell].astype(np.int32), d.n_celltypes
else:
    row_seg, S = np.zeros(d.N, dtype=np.int32), 1
eng = Engine(0)
eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
eng.build_layout(row_seg, S)
eng.enable_timing(True)
info = eng.layout_info()
print({k: info[k] for k in ('n_tiles', 'n_resident_tiles', 'n_cluster_tiles', 'n_stream_tiles', 'nnz')})
for i in range(n_fits):
    eng.fit()
    torch.cuda.synchronize()
    t = eng.last_timings()
    print(i, {k: round(v, 3) for k, v in t.items() if v > 0})
res = eng.results(traces=False)
seg_A = eng.export_layout_array('seg_A', np.int64)
units = float((seg_A * res['iters'].astype(np.int64)).sum())
print('iters mean', res['iters'].mean(), 'max', res['iters'].max(), 'units', units, 'rate G/s', units / t['em'] / 1e6)
