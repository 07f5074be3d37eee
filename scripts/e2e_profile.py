"""Where the time of one reference-facing call goes (host side): cProfile of model._fit_pooling_model on a bench workload.

    python scripts/e2e_profile.py [cfg5] [scale] [--report]     (--report: + Stellarscope.reassign sweep and count matrices)
"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from stellarscope_b200 import model

cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg5'
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
d = bench.make_workload(cfg, scale)
opts = bench.BenchOpts(bench.CONFIGS[cfg]['mode'])
st = bench.make_st(d, opts)
for i in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tl, pool = model._fit_pooling_model(st, opts)
    torch.cuda.synchronize()
    print(f'call {i}: {time.perf_counter() - t0:.3f} s (device fit {tl.fit_seconds:.3f} s)', flush=True)
    tl.close()
pr = cProfile.Profile()
pr.enable()
tl, pool = model._fit_pooling_model(st, opts)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
if '--report' in sys.argv:
    opts.reassign_mode = ['best_exclude', 'best_conf', 'best_random', 'best_average', 'total_hits']
    st.opts = opts
    for i in range(2):
        t0 = time.perf_counter()
        model.reassign(st, tl)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        counts, bc, ft = model.count_matrices(st, tl)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f'pass {i}: reassign x5 {t1 - t0:.3f} s, count_matrices {t2 - t1:.3f} s', flush=True)
    pr = cProfile.Profile()
    pr.enable()
    model.reassign(st, tl)
    counts, bc, ft = model.count_matrices(st, tl)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(30)
