"""2+ GPU check (torchrun): model sharding of individual / celltype pooling against the golden fixtures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
