"""Row-sharded pseudobulk fit over N GPUs (torchrun): BASELINE.json configs[3] shape, scaled.

    python -m torch.distributed.run --nproc-per-node N scripts/pseudobulk_dist_bench.py [alignments_per_rank]

Every rank generates its own block of rows (seed = rank; the union is one pseudobulk model over all
loci), builds its layout and fits with (a) the persistent kernel exchanging over NVLink peer memory and
(b) the host loop with one NCCL all-reduce per iteration.  Device time (CUDA events), max over ranks.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from stellarscope_b200 import synth
from stellarscope_b200.engine import Engine

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
A = int(float(sys.argv[1])) if len(sys.argv) > 1 else 25_000_000
d = synth.generate(n_cells=2000, n_alignments=A, n_loci=28000, seed=rank)
seg = np.zeros(d.N, dtype=np.int32)
out = {}
for peers in (True, False):
    eng = Engine(local)
    eng.set_sharding(rank, world, None, peers=peers, K=d.K)
    eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
    eng.build_layout(seg, 1)
    best = None
    for i in range(4):
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eng.fit()
        e.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if i > 0:
            best = float(ms) if best is None else min(best, float(ms))
    res = eng.results(traces=False)
    it = int(res['iters'][0])
    tot = torch.tensor([float(d.A)], dtype=torch.float64, device=dev)
    dist.all_reduce(tot)
    pi, th = eng.dense_params_sharded()
    chk = torch.tensor([float(it), float(res['lnl'][0]), float(pi.sum())], dtype=torch.float64, device=dev)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(bool(torch.equal(c, allc[0])) for c in allc)
    out[peers] = (best, it, float(tot))
    if rank == 0:
        mode = ('persistent kernel, exchange over peer memory (peer_mode=%s)' % eng.peer_mode) if peers else \
            'host loop, NCCL all-reduce per iteration (%d all-reduces)' % getattr(eng, 'n_allreduce', 0)
        print(f'{world} GPU(s), {int(float(tot))} alignments, {it} iterations, {mode}: fit {best:.3f} ms = '
              f'{best / it * 1e3:.1f} us/iteration, {float(tot) * it / best / 1e6:.1f} G aln*iter/s, '
              f'identical on all ranks: {same}, sum(pi) = {float(pi.sum()):.12f}', flush=True)
    eng.close()
dist.barrier()
dist.destroy_process_group()
