"""Where does the e2e time go?  (host prep / H2D / build / fit / results)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from stellarscope_b200 import model
from stellarscope_b200.engine import Engine

d = bench.make_workload(0)
opts = bench.BenchOpts()
st = bench.make_st(d, opts)
def T(label, fn, n=3):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f'{label:28s} {1e3*min(ts):8.2f} ms'); return r
T('pooling_row_segments', lambda: model.pooling_row_segments(st, opts))
seg, keys = model.pooling_row_segments(st, opts)
eng = Engine(0)
T('set_matrix (pin+H2D)', lambda: eng.set_matrix(d.indptr, d.indices, d.scores, d.K))
T('build_layout', lambda: eng.build_layout(seg, len(keys)))
T('fit', lambda: eng.fit())
T('results(traces)', lambda: eng.results(traces=True))
tl = model.TelescopeLikelihood(st.raw_scores, opts, row_segment=seg, n_segments=len(keys))
T('fit_segments (FitInfo objs)', lambda: tl.fit_segments())
T('TelescopeLikelihood ctor', lambda: model.TelescopeLikelihood(st.raw_scores, opts, row_segment=seg, n_segments=len(keys)), n=2)
T('_fit_pooling_model', lambda: model._fit_pooling_model(st, opts), n=2)
