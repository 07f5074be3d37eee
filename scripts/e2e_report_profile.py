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
pr = cProfile.Profile(); pr.enable()
model.reassign(st, tl); counts, bc, ft = model.count_matrices(st, tl)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
