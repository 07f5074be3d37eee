"""BASELINE.json configs[4] ("atlas": 100k cells, ~500M alignments, individual mode) over N GPUs -- STRONG scaling:
the total work is fixed, every rank takes 1/N of the cells (individual mode shards by cell with no data-path
collective).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/cfg5_dist_bench.py [scale=1.0]
    python scripts/cfg5_dist_bench.py [scale]                      # one GPU

Rank r generates cells r, r + N, ... of the configuration as its own synthetic data set (same generator and size
distribution, seed 100 + r: drawing 100k/N cells from LogNormal(sigma = 1) is what a random 1/N of the atlas looks
like; the shards differ by a few percent in alignments, which the report shows), builds its layout and fits.
Timed: layout build and EM fit per rank on the device (CUDA events), max over ranks after a barrier.  Reported:
alignments x iterations / s of the whole job, the per-rank spread, and the sampled-oracle parity of a few cells per
rank (CPU, test infrastructure).  Written without a GPU at hand (round 1); the calls are the ones
tests/tools/config_bench.py uses.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from stellarscope_b200 import synth
from stellarscope_b200.engine import Engine

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n_oracle = int(sys.argv[2]) if len(sys.argv) > 2 else 4
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
distributed = 'RANK' in os.environ
if distributed:
    dist.init_process_group('nccl', device_id=dev)
rank = dist.get_rank() if distributed else 0
world = dist.get_world_size() if distributed else 1

n_cells = int(100_000 * scale) // world
n_aln = int(500_000_000 * scale) // world
t0 = time.perf_counter()
d = synth.generate(n_cells=n_cells, n_alignments=n_aln, n_loci=28_000, seed=100 + rank)
gen_s = time.perf_counter() - t0
eng = Engine(local)
eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
torch.cuda.synchronize(dev)
if distributed:
    dist.barrier()
t0 = time.perf_counter()
eng.build_layout(d.row_cell, d.n_cells)
torch.cuda.synchronize(dev)
build_s = time.perf_counter() - t0
eng.enable_timing(True)
best = None
for i in range(3):
    if distributed:
        dist.barrier()
    eng.fit()
    torch.cuda.synchronize(dev)
    t = eng.last_timings()
    if best is None or t['em'] < best['em']:
        best = t
res = eng.results(traces=False)
seg_A = eng.export_layout_array('seg_A', np.int64)
it = res['iters'].astype(np.int64)
units = float((seg_A * it).sum())
info = eng.layout_info()
stat = torch.tensor([best['em'], build_s * 1e3, units, float(d.A), float(info['n_cluster_tiles']),
                     float(info['n_stream_tiles']), float(info['n_tiles'])], dtype=torch.float64, device=dev)
if distributed:
    allst = [torch.zeros_like(stat) for _ in range(world)]
    dist.all_gather(allst, stat)
    allst = torch.stack(allst).cpu().numpy()
else:
    allst = stat.cpu().numpy()[None]

# sampled oracle parity (CPU): a few random cells and the largest one of this rank
from oracle.em_oracle import OracleLikelihood, Opts, submatrix_rows
rng = np.random.default_rng(rank)
bounds = np.searchsorted(d.row_cell, np.arange(d.n_cells + 1))
pick = np.unique(np.concatenate([rng.choice(d.n_cells, size=min(n_oracle, d.n_cells), replace=False),
                                 [int(np.argmax(seg_A))]]))
ptr64, idx64 = d.indptr.astype(np.int64), d.indices.astype(np.int64)
worst = 0.0
for s in pick:
    rows = np.arange(bounds[s], bounds[s + 1])
    ptr, idx, sc, take = submatrix_rows(ptr64, idx64, d.scores, rows)
    om = OracleLikelihood(ptr, idx, sc, d.K, Opts()).em()
    assert res['iters'][s] == len(om.fitinfo.iterations), (rank, s)
    assert abs(res['lnl'][s] - float(om.lnl)) <= 1e-10 * abs(float(om.lnl)), (rank, s)
    pi, th = eng.dense_params(int(s))
    po = np.asarray(om.pi, dtype=np.float64)
    ok = po > 1e-290
    worst = max(worst, float(np.max(np.abs(pi[ok] - po[ok]) / po[ok])))
assert worst < 1e-6
w = torch.tensor([worst], dtype=torch.float64, device=dev)
if distributed:
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
if rank == 0:
    em_ms, build_ms = allst[:, 0], allst[:, 1]
    tot_units, tot_A = allst[:, 2].sum(), allst[:, 3].sum()
    print(f'cfg5 x{scale} over {world} GPU(s): {n_cells * world} cells, {int(tot_A)} alignments '
          f'(per rank {allst[:, 3].min() / 1e6:.1f}-{allst[:, 3].max() / 1e6:.1f} M; generated in {gen_s:.0f} s)')
    print(f'layout build per rank {build_ms.min():.0f}-{build_ms.max():.0f} ms; tiles {int(allst[:, 6].sum())} '
          f'(cluster {int(allst[:, 4].sum())}, stream {int(allst[:, 5].sum())})')
    print(f'EM fit per rank {em_ms.min():.1f}-{em_ms.max():.1f} ms (device); whole job '
          f'{tot_units / em_ms.max() / 1e6:.1f} G aln*iter/s at the slowest rank\'s time '
          f'({tot_units / world / em_ms.mean() / 1e6:.1f} G per GPU on average)')
    print(f'oracle parity on {len(pick)} sampled cells per rank incl. its largest: iterations equal, lnL 1e-10, '
          f'max rel pi err {float(w):.2e}')
    print('OK')
eng.close()
if distributed:
    dist.barrier()
    dist.destroy_process_group()
