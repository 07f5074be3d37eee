"""cProfile of the e2e call (development aid)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from stellarscope_b200 import model
d = bench.make_workload(0)
opts = bench.BenchOpts()
st = bench.make_st(d, opts)
for _ in range(3):
    tl, pi = model._fit_pooling_model(st, opts); torch.cuda.synchronize(); tl.close()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    tl, pi = model._fit_pooling_model(st, opts); torch.cuda.synchronize(); tl.close()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
