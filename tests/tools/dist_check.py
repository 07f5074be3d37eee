"""2+ GPU check (torchrun): model sharding of individual / celltype pooling against the golden fixtures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch, torch.distributed as dist
from tests import helpers as H
from stellarscope_b200 import model

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
for name in ('individual', 'celltype', 'individual_umicounts'):
    g = H.load_golden(name)
    opts = H.golden_opts(g, cls=H.StOpts)
    opts.reassign_mode = ['best_exclude', 'best_average', 'total_hits']
    st = H.make_st(g, opts)
    tl, pool = model._fit_pooling_model(st, opts)
    keys = list(pool.models_info.keys())
    its = [len(pool.models_info[k].iterations) for k in keys]
    lnl = [pool.models_info[k].final_lnl for k in keys]
    good = its == g['iters'].tolist() and np.allclose(lnl, g['lnl'], rtol=1e-10)
    model.reassign(st, tl)
    counts, bc, ft = model.count_matrices(st, tl)
    for m in opts.reassign_mode:
        good &= bool(np.allclose(np.asarray(counts[m].todense()), g[f'c_{m}'], rtol=1e-12, atol=0))
    print(f'rank {rank}/{world} {name}: local models {tl.n_segments} of {len(keys)} -> {"OK" if good else "MISMATCH"}', flush=True)
    ok &= good
# pseudobulk: ONE model, rows sharded over the ranks; exchange over peer memory (persistent kernel)
# and through the NCCL hook (host loop)
for name in ('pseudobulk', 'pseudobulk_lnl', 'pseudobulk_priors'):
    for peers in (True, False):
        g = H.load_golden(name)
        opts = H.golden_opts(g, cls=H.StOpts)
        opts.peer_exchange = peers
        opts.reassign_mode = ['best_exclude', 'best_average', 'total_hits']
        st = H.make_st(g, opts)
        tl, pool = model._fit_pooling_model(st, opts)
        fi = pool.models_info['pseudobulk']
        good = len(fi.iterations) == int(g['iters'][0]) and abs(fi.final_lnl - g['lnl'][0]) <= 1e-10 * abs(g['lnl'][0])
        good &= fi.nobs == int(g['nobs'][0]) and fi.nfeats == int(g['nfeats'][0])
        d = np.array([x[2] for x in fi.iterations]); dg = g['diffs'][0, :len(d)]
        good &= bool(np.all(np.abs(d - dg) <= 1e-12 + 1e-6 * dg))
        r, okp = H.rel_err(np.asarray(tl.pi.toarray()).ravel(), g['pi'][0])
        good &= bool(r < 1e-6 and okp)
        # identical decision and parameters on every rank
        v = torch.tensor([float(len(fi.iterations)), float(fi.final_lnl)], dtype=torch.float64, device='cuda')
        vs = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(vs, v)
        good &= all(bool(torch.equal(x, vs[0])) for x in vs)
        model.reassign(st, tl)
        counts, bc, ft = model.count_matrices(st, tl)
        for m in opts.reassign_mode:
            good &= bool(np.allclose(np.asarray(counts[m].todense()), g[f'c_{m}'], rtol=1e-12, atol=0))
        print(f'rank {rank}/{world} {name} peers={peers} (peer_mode={tl._engine.peer_mode}): iters {len(fi.iterations)} '
              f'-> {"OK" if good else "MISMATCH"}', flush=True)
        ok &= good
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
