"""2+ GPU check (torchrun): model sharding of individual / celltype pooling (incl. wide models fitted by all ranks
together), the row-sharded pseudobulk fit and --umi_counts under sharding, against the golden fixtures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.sparse as sp
import torch, torch.distributed as dist
from tests import helpers as H
from stellarscope_b200 import model

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, world = dist.get_rank(), dist.get_world_size()
ok = True


def check_sharded_reassign(g, st, tl, infos, modes):
    """Rank-local reassignment matrices against the golden ones (own rows equal, other rows empty) and the
    all-reduced ReassignInfo counters against the reference's."""
    good = True
    rows = tl._row_ids
    mask = np.zeros(tl.N, dtype=bool)
    mask[rows] = True
    for mode in modes:
        r = sp.csr_matrix(st.reassignments[mode]); r.eliminate_zeros(); r.sort_indices()
        ref = sp.csr_matrix((g[f'r_{mode}_data'], g[f'r_{mode}_indices'], g[f'r_{mode}_indptr']), shape=r.shape)
        keep = sp.diags(mask.astype(ref.dtype)) @ ref
        keep = sp.csr_matrix(keep); keep.eliminate_zeros(); keep.sort_indices()
        same = np.array_equal(r.indptr, keep.indptr) and np.array_equal(r.indices, keep.indices) and \
            np.allclose(r.data, keep.data, rtol=1e-15, atol=0)
        ri, gi = infos[mode], g[f'r_{mode}_info']
        cnt = ri.assigned == gi[0] and ri.ambiguous == gi[1] and (ri.unaligned if ri.unaligned is not None else -1) == gi[2]
        if f'r_{mode}_dist' in g.files:
            cnt &= sorted(ri.ambigous_dist.items()) == [tuple(x) for x in g[f'r_{mode}_dist'].tolist()]
        if not (same and cnt):
            print(f'  rank {rank} mode {mode}: matrix {"ok" if same else "DIFF"} counters {"ok" if cnt else "DIFF"}', flush=True)
        good &= bool(same and cnt)
    return good


# (golden, wide models): the largest cell type of the `celltype` fixture holds 74 % of the alignments: 'auto' (with the
# lockstep cost set to 0 for these tiny fixtures) fits it with all ranks together (rows in blocks); [1, 2] makes two
# of the three models wide (on 2 ranks one rank then has no model of its own), [0, 1, 2] all three (no rank has a
# model of its own); None never splits a model
for name, wide in (('individual', 'auto'), ('celltype', 'auto'), ('celltype', [1, 2]), ('celltype', [0, 1, 2]), ('celltype', None),
                   ('celltype_umidedup', 'auto'), ('individual_umicounts', 'auto')):
    g = H.load_golden(name)
    opts = H.golden_opts(g, cls=H.StOpts)
    opts.wide_models, opts.wide_model_cost = wide, 0
    opts.reassign_mode = list(H.MODES)                       # all seven, in the order the goldens drew opts.rng
    opts.rng = np.random.default_rng(2024)
    st = H.make_st(g, opts)
    tl, pool = model._fit_pooling_model(st, opts)
    keys = list(pool.models_info.keys())
    its = [pool.models_info[k].n_iterations for k in keys]
    lnl = [pool.models_info[k].final_lnl for k in keys]
    good = its == g['iters'].tolist() and np.allclose(lnl, g['lnl'], rtol=1e-10)
    infos = model.reassign(st, tl)
    good &= check_sharded_reassign(g, st, tl, infos, opts.reassign_mode)
    counts, bc, ft = model.count_matrices(st, tl)
    for m in opts.reassign_mode:
        good &= bool(np.allclose(np.asarray(counts[m].todense()), g[f'c_{m}'], rtol=1e-12, atol=0))
    n_up = tl._engine.N
    tot = torch.tensor([float(n_up)], dtype=torch.float64, device='cuda'); dist.all_reduce(tot)
    good &= int(tot.item()) == tl.N                           # the ranks' uploads partition the rows
    pi_ok = True
    print(f'rank {rank}/{world} {name} wide={wide}: local models {tl.n_segments} of {len(keys)}, wide '
          f'{list(getattr(tl, "wide_model_ids", []))} -> {"OK" if good else "MISMATCH"}', flush=True)
    ok &= good
# pseudobulk: ONE model, rows sharded over the ranks; exchange over peer memory (persistent kernel)
# and through the NCCL hook (host loop)
for name in ('pseudobulk', 'pseudobulk_lnl', 'pseudobulk_priors'):
    for peers in (True, False):
        g = H.load_golden(name)
        opts = H.golden_opts(g, cls=H.StOpts)
        opts.peer_exchange = peers
        opts.reassign_mode = list(H.MODES)
        opts.rng = np.random.default_rng(2024)
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
        infos = model.reassign(st, tl)
        good &= check_sharded_reassign(g, st, tl, infos, opts.reassign_mode)
        counts, bc, ft = model.count_matrices(st, tl)
        for m in opts.reassign_mode:
            good &= bool(np.allclose(np.asarray(counts[m].todense()), g[f'c_{m}'], rtol=1e-12, atol=0))
        print(f'rank {rank}/{world} {name} peers={peers} (peer_mode={tl._engine.peer_mode}): iters {len(fi.iterations)} '
              f'-> {"OK" if good else "MISMATCH"}', flush=True)
        ok &= good
# --umi_counts with models / rows split over the ranks: the distinct UMIs of a cell are counted where all its reads
# are (whole cells per rank); checked against the same call without sharding (every rank alone, all rows)
for name, wide in (('celltype', [1, 2]), ('pseudobulk', None), ('individual_umicounts', 'auto')):
    g = H.load_golden(name)
    res = []
    for shard in (True, False):
        opts = H.golden_opts(g, cls=H.StOpts)
        opts.umi_counts, opts.ignore_umi = True, True
        opts.wide_models, opts.wide_model_cost = wide, 0
        opts.shard_models = shard
        opts.reassign_mode = ['best_exclude', 'best_average', 'total_hits']
        st = H.make_st(g, opts)
        tl, pool = model._fit_pooling_model(st, opts)
        model.reassign(st, tl)
        counts, bc, ft = model.count_matrices(st, tl)
        res.append({m: np.asarray(counts[m].todense()) for m in opts.reassign_mode})
        tl.close()
    good = all(np.allclose(res[0][m], res[1][m], rtol=1e-12, atol=0) for m in res[0])
    good &= bool(res[0]['best_exclude'].sum() > 0)
    print(f'rank {rank}/{world} {name} umi_counts sharded vs alone: {"OK" if good else "MISMATCH"}', flush=True)
    ok &= good
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
