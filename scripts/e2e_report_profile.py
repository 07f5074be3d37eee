"""cProfile of the second half of the drop-in flow on config 2: Stellarscope.reassign + the aggregation of output_report
(development aid).   python scripts/e2e_report_profile.py"""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from stellarscope_b200 import model
d = bench.make_workload(0)
opts = bench.BenchOpts()
opts.reassign_mode = ['best_exclude', 'best_conf', 'best_random', 'best_average', 'total_hits']
st = bench.make_st(d, opts)
names = list(st.bcode_ridx_map)
st.read_index = {f'r{i:08d}': i for i in range(d.N)}
st.read_bcode_map = {f'r{i:08d}': names[c] for i, c in enumerate(d.row_cell.tolist())}
st.read_umi_map = {}
st.feat_index = {('__no_feature' if j == 0 else f'L{j:05d}'): j for j in range(d.K)}
st.filtlist = None
tl, pool = model._fit_pooling_model(st, opts)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model.reassign(st, tl)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    counts, bc, ft = model.count_matrices(st, tl)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'reassign (5 modes) {1e3 * (t1 - t0):.1f} ms, count_matrices {1e3 * (t2 - t1):.1f} ms', flush=True)
# the file output: device formatter of output_report against scipy.io.mmwrite on the same matrices
import tempfile, scipy.io
from stellarscope_b200 import mtx
tmpd = tempfile.mkdtemp(prefix='ss_report_')
opts.outfile_path = lambda suffix: os.path.join(tmpd, suffix)
opts.version = 'bench'
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i, (mode, cm) in enumerate(counts.items()):
        mtx.mmwrite(os.path.join(tmpd, f'dev_{mode}.mtx'), cm, comment=model._mtx_meta(st, mode), engine=tl._engine,
                    device_coo=getattr(cm, '_ss_dev_coo', None))
    t1 = time.perf_counter()
    for i, (mode, cm) in enumerate(counts.items()):
        scipy.io.mmwrite(os.path.join(tmpd, f'ref_{mode}.mtx'), cm, comment=model._mtx_meta(st, mode))
    t2 = time.perf_counter()
    nent = sum(cm.nnz for cm in counts.values())
    print(f'MTX files of 5 modes ({nent} entries): device formatter {1e3 * (t1 - t0):.1f} ms, scipy.io.mmwrite {1e3 * (t2 - t1):.1f} ms',
          flush=True)
for mode, cm in counts.items():
    a, b = (open(os.path.join(tmpd, f'{k}_{mode}.mtx'), 'rb').read() for k in ('ref', 'dev'))
    same = a == b if np.issubdtype(cm.dtype, np.integer) else (scipy.io.mmread(os.path.join(tmpd, f'dev_{mode}.mtx')).tocsc() != cm).nnz == 0
    print(f'  {mode}: {"identical bytes" if a == b else ("equal matrices" if same else "MISMATCH")}')
if '--no-cprofile' in sys.argv:
    sys.exit(0)
pr = cProfile.Profile(); pr.enable()
model.reassign(st, tl); counts, bc, ft = model.count_matrices(st, tl)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
