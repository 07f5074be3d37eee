"""Where does the reassign + count time go?  (development aid)   python scripts/reassign_profile.py [scale]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stellarscope_b200 import synth
from stellarscope_b200.engine import Engine

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
d = synth.generate(n_cells=int(100_000 * scale), n_alignments=int(500_000_000 * scale), n_loci=28000, seed=5)
eng = Engine(0)
eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
eng.build_layout(d.row_cell, d.n_cells)
eng.fit(); z = eng.emit_z(); torch.cuda.synchronize()
rc = torch.from_numpy(d.row_cell).cuda()
def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f'{label:40s} {1e3 * (time.perf_counter() - t0):8.2f} ms', flush=True); return r
for rep in range(2):
    for m in ('best_exclude', 'total_hits'):
        flag, nbest, info = T(f'{m}: eng.reassign', lambda: eng.reassign(m, z, conf_thresh=0.9, tie_tol=1e-9))
        out = T(f'{m}: eng.count', lambda: eng.count(flag, rc))
        print('   entries', len(out[2]), 'selected', int(info[0]))
# device part of ss_count alone (outputs preallocated, no D2H of the COO)
import ctypes as C
cap = eng.A
oc = torch.empty(cap, dtype=torch.int32, device='cuda'); ol = torch.empty(cap, dtype=torch.int32, device='cuda')
ov = torch.empty(cap, dtype=torch.float64, device='cuda')
k = eng._keep
for m in ('best_exclude', 'total_hits'):
    flag, nbest, info = eng.reassign(m, z, conf_thresh=0.9, tie_tol=1e-9)
    n = C.c_int64(0)
    for rep in range(2):
        T(f'{m}: ss_count only (device + syncs)', lambda: eng.lib.ss_count(eng.h, eng._stream(), eng.N, eng.K, eng._p(k['row_ptr']),
          eng._p(k['col_idx']), eng._p(flag), None, eng._p(rc), None, cap, eng._p(oc), eng._p(ol), eng._p(ov), C.byref(n)))
    T(f'{m}: D2H of {n.value} COO entries', lambda: (oc[:n.value].cpu(), ol[:n.value].cpu(), ov[:n.value].cpu()))
