"""Celltype pooling over N GPUs (torchrun): BASELINE.json configs[2] shape, scaled -- whole-model packing (LPT)
against packing + wide models fitted by all ranks together.

    python -m torch.distributed.run --nproc-per-node N scripts/celltype_dist_bench.py [scale] [dominant_share]

``dominant_share`` > 0 moves that fraction of the cells into cell type 0 (tissues usually have a dominant type).

Every rank generates the same synthetic data set (50k cells x scale, 100M alignments x scale, 20 cell types with
Dirichlet(0.5) sizes) and calls the drop-in ``_fit_pooling_model``; reported: wall time of the call from a barrier
to the last rank's return (max over ranks), the device time of the fits, the models each rank got.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
from stellarscope_b200 import model, parallel, synth

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
d = synth.generate(n_cells=int(50_000 * scale), n_alignments=int(100_000_000 * scale), n_loci=28000, seed=3, n_celltypes=20)
dominant = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
if dominant > 0:
    d.cell_type = d.cell_type.copy()
    d.cell_type[np.random.default_rng(11).random(d.n_cells) < dominant] = 0
opts = bench.BenchOpts()
opts.pooling_mode, opts.shard_models = 'celltype', True
st = bench.make_st(d, opts)
names = list(st.bcode_ridx_map)
st.celltypes = [f'type{t}' for t in range(d.n_celltypes)]
st.ctype_bcode_map = {f'type{t}': [names[c] for c in np.flatnonzero(d.cell_type == t)] for t in range(d.n_celltypes)}
lens = np.diff(d.indptr)
seg_A = np.bincount(d.cell_type[d.row_cell], weights=lens, minlength=d.n_celltypes)
if rank == 0:
    print(f'{world} GPU(s): {d.n_cells} cells, {d.A} alignments, {d.n_celltypes} cell types; shares of the largest: '
          f'{np.round(np.sort(seg_A / seg_A.sum())[::-1][:6], 3).tolist()}', flush=True)
ref = None
for label, wide in (('whole-model packing (LPT)', None), ('packing + wide models (automatic choice)', 'auto'),
                    ('packing + the four largest models wide', 'top4')):
    opts.wide_models = list(np.argsort(-seg_A)[:4]) if wide == 'top4' else wide
    best = None
    for i in range(3):
        torch.cuda.synchronize(dev)
        dist.barrier()
        t0 = time.perf_counter()
        tl, pool = model._fit_pooling_model(st, opts)
        torch.cuda.synchronize(dev)
        t = torch.tensor([time.perf_counter() - t0, float(tl.fit_seconds or 0.0), float(getattr(tl, 'fit_seconds_wide', 0.0))],
                         dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i > 0 and (best is None or float(t[0]) < best[0]):
            best = (float(t[0]), float(t[1]), float(t[2]))
        keys = list(pool.models_info.keys())
        its = np.array([len(pool.models_info[k].iterations) for k in keys])
        lnl = float(pool.total_lnl)
        nloc, wide = tl.n_segments, list(map(int, getattr(tl, 'wide_model_ids', [])))
        tl.close()
    units = float((seg_A * its).sum())
    loc = [None] * world
    dist.all_gather_object(loc, nloc)
    if ref is None:
        ref = (its.copy(), lnl)
    same = bool(np.array_equal(its, ref[0])) and abs(lnl - ref[1]) <= 1e-9 * abs(ref[1])
    if rank == 0:
        dev_ms = (best[1] + best[2]) * 1e3
        print(f'{label}: EM fit {dev_ms:.1f} ms device (packed models {best[1] * 1e3:.1f} ms, max over ranks, + wide models '
              f'{best[2] * 1e3:.1f} ms) = {units / dev_ms / 1e6:.1f} G aln*iter/s; whole drop-in call {best[0] * 1e3:.0f} ms wall '
              f'(host preparation is not sharded); models per rank {loc}, wide {wide}; '
              f'iterations / lnL equal to the first variant: {same}', flush=True)
dist.barrier()
dist.destroy_process_group()
