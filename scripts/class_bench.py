"""Per-size-class throughput of the resident kernel: uniform cell sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stellarscope_b200 import synth
from stellarscope_b200.engine import Engine

sizes = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [100, 200, 400, 800, 1600, 3200]
total = int(float(sys.argv[2])) if len(sys.argv) > 2 else 3_000_000
for sz in sizes:
    n_cells = max(1, total // sz)
    d = synth.generate(n_cells=n_cells, n_alignments=total, n_loci=28000, seed=1, sigma=0.05)
    eng = Engine(0)
    eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
    eng.build_layout(d.row_cell, d.n_cells)
    eng.enable_timing(True)
    info = eng.layout_info()
    best = None
    for i in range(4):
        eng.fit()
        torch.cuda.synchronize()
        t = eng.last_timings()
        best = t if best is None or t['em'] < best['em'] else best
    res = eng.results(traces=False)
    segA = eng.export_layout_array('seg_A', np.int64)
    units = float((segA * res['iters']).sum())
    cls = {k: round(v, 3) for k, v in best.items() if k.startswith('lane') and v > 0 or k == 'stream' and v > 0}
    print(f'size {sz:5d} cells {n_cells:6d} tiles {info["n_tiles"]} stream_tiles {info["n_stream_tiles"]} '
          f'mean_iters {res["iters"].mean():.1f} em_ms {best["em"]:.3f} rate {units / best["em"] / 1e6:.1f} G/s  {cls}', flush=True)
    eng.close()
