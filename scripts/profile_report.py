"""Profiling driver (ncu) for everything around the EM loop: layout build, emit, the reassignment kernel, the
sort-free count kernels, the CSR compaction.   python scripts/profile_report.py [cfg5] [scale]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from stellarscope_b200.engine import Engine

cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg5'
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
d = bench.make_workload(cfg, scale)
row_seg, S = bench.row_models(d, bench.CONFIGS[cfg]['mode'])
eng = Engine(0)
eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
eng.build_layout(row_seg, S)
eng.fit()
z = eng.emit_z()
rc = torch.from_numpy(d.row_cell).to(z.device)
torch.cuda.synchronize()
for mode in ('best_exclude', 'best_conf', 'best_average', 'total_hits'):
    t0 = time.perf_counter()
    flag, nbest, info = eng.reassign(mode, z, conf_thresh=0.9, tie_tol=1e-9)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ptr, idx, frac = eng.reassign_csr(flag, nbest, fractions=(mode == 'best_average'))
    t2 = time.perf_counter()
    if mode == 'best_average':
        out = eng.count(flag, rc)
    else:
        out = eng.count_cells(flag, rc, d.n_cells)
    t3 = time.perf_counter()
    print(f'{mode:13s} reassign {1e3 * (t1 - t0):7.2f} ms  csr (+D2H) {1e3 * (t2 - t1):7.2f} ms  count (+D2H) {1e3 * (t3 - t2):7.2f} ms  '
          f'entries {len(out[1])}', flush=True)
print('alignments', d.A, 'rows', d.N)
