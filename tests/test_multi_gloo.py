"""Multi-rank host logic on CPU: world_size 2, gloo backend.  The per-rank fit is the oracle
(test infrastructure) -- what is checked here is the sharding / gather plumbing of
stellarscope_b200.parallel."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from stellarscope_b200 import parallel
from tests import helpers as H


def test_lpt_shards_balanced_and_deterministic():
    rng = np.random.default_rng(0)
    sizes = np.round(rng.lognormal(6, 1, 500)).astype(int)
    for world in (1, 2, 4, 8):
        sh = parallel.lpt_shards(sizes, world)
        assert sorted(np.concatenate(sh).tolist()) == list(range(500))
        loads = np.array([sizes[b].sum() for b in sh])
        assert loads.max() <= loads.mean() + sizes.max()
        sh2 = parallel.lpt_shards(sizes, world)
        assert all(np.array_equal(a, b) for a, b in zip(sh, sh2))


def test_row_blocks_respect_cells():
    d = H.small_synth(seed=9, n_cells=40, n_alignments=9000)
    b = parallel.row_blocks(d.indptr, d.row_cell, 4)
    assert b[0] == 0 and b[-1] == d.N and np.all(np.diff(b) >= 0)
    for x in b[1:-1]:
        assert d.row_cell[x] != d.row_cell[x - 1]
    per = np.diff(d.indptr.astype(np.int64)[b])
    assert per.max() < 0.5 * d.A


def _oracle_fit_local(ptr, idx, dat, K, seg_local, n_local):
    from oracle.em_oracle import OracleLikelihood, Opts, submatrix_rows
    out = []
    for s in range(n_local):
        rows = np.flatnonzero(seg_local == s)
        p, i, sc, _ = submatrix_rows(ptr, idx, dat, rows)
        m = OracleLikelihood(p, i, sc, K, Opts()).em()
        out.append((len(m.fitinfo.iterations), float(m.lnl)))
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    d = H.small_synth(seed=31, n_cells=12, n_alignments=3000, n_loci=150)
    recs, rows, take = parallel.fit_models_sharded(d.indptr.astype(np.int64), d.indices.astype(np.int64), d.scores,
                                                  d.K, d.row_cell.astype(np.int64), d.n_cells, _oracle_fit_local)
    q.put((rank, recs, rows.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_fit_models_sharded_world2_gloo():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = H.small_synth(seed=31, n_cells=12, n_alignments=3000, n_loci=150)
    single = _oracle_fit_local(d.indptr.astype(np.int64), d.indices.astype(np.int64), d.scores, d.K,
                               d.row_cell.astype(np.int64), d.n_cells)
    rows_all = []
    for rank, recs, rows in got:
        assert recs == single                    # every rank sees every model's record, same as one process
        rows_all += rows
    assert sorted(rows_all) == list(range(d.N))   # rows are partitioned across the two ranks


def test_row_block_equal_alignments():
    d = H.small_synth(seed=10, n_cells=30, n_alignments=8000)
    ptr = d.indptr.astype(np.int64)
    for world in (1, 2, 3, 8):
        blocks = [parallel.row_block(ptr, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == d.N
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        per = np.array([ptr[b] - ptr[a] for a, b in blocks])
        assert per.max() - per.min() <= 2 * np.diff(ptr).max()
    assert [parallel.row_block(np.array([0, 3]), r, 4) for r in range(4)] == [(0, 0), (0, 1), (1, 1), (1, 1)]


def _sharded_pseudobulk_worker(rank, world, port, q):
    """The row-sharded pseudobulk protocol of csrc/ss_dense.cu restated in numpy over gloo: s_max by
    all-reduce MAX, [pisum0, weights] by all-reduce SUM once, the K-vector of partial theta sums by
    all-reduce SUM per iteration, identical update and decision on every rank."""
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    d = H.small_synth(seed=41, n_cells=8, n_alignments=4000, n_loci=120, unique_frac=0.1)
    ptr = d.indptr.astype(np.int64)
    r0, r1 = parallel.row_block(ptr, rank, world)
    a, b = ptr[r0], ptr[r1]
    lens = np.diff(ptr[r0:r1 + 1])
    cols, sc = d.indices[a:b].astype(np.int64), d.scores[a:b].astype(np.float64)
    K = d.K

    def allreduce(x, op):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        dist.all_reduce(t, op=op)
        return t.numpy()
    smax = allreduce(np.array([sc.max() if len(sc) else 0.0]), dist.ReduceOp.MAX)[0]
    Q = np.expm1((sc * (1.0 / smax)) * 100.0)
    row = np.repeat(np.arange(r1 - r0), lens)
    w = np.zeros(r1 - r0)
    np.maximum.at(w, row, Q)
    amb = lens > 1
    stat = np.zeros(K + 2)
    np.add.at(stat, cols[~amb[row]], Q[~amb[row]])                  # pisum0
    stat[K], stat[K + 1] = w[amb].sum(), w.sum()
    stat = allreduce(stat, dist.ReduceOp.SUM)
    pisum0, wamb, wtot = stat[:K], stat[K], stat[K + 1]
    wmax = np.expm1((smax * (1.0 / smax)) * 100.0)
    thpw, pipw = 200000 * wmax, 0.0
    thd, pid = wamb + thpw * K, wtot + pipw * K
    pi = np.full(K, 1.0 / K)
    th = np.full(K, 1.0 / K)
    am = amb[row]
    it = 0
    while True:
        it += 1
        u = Q * pi[cols] * np.where(am, th[cols], 1.0)
        rs = np.zeros(r1 - r0)
        np.add.at(rs, row, u)
        z = u / rs[row]
        T = np.zeros(K)
        np.add.at(T, cols[am], (z * w[row])[am])
        T = allreduce(T, dist.ReduceOp.SUM)                         # the one exchange per iteration
        th_n, pi_n = (T + thpw) / thd, (pisum0 + T + pipw) / pid
        diff = np.abs(pi_n - pi).sum()
        pi, th = pi_n, th_n
        if diff < 1e-7 or it >= 100:
            break
    q.put((rank, it, pi.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_pseudobulk_protocol_world2_gloo():
    from oracle.em_oracle import OracleLikelihood, Opts
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_pseudobulk_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = H.small_synth(seed=41, n_cells=8, n_alignments=4000, n_loci=120, unique_frac=0.1)
    m = OracleLikelihood(d.indptr.astype(np.int64), d.indices.astype(np.int64), d.scores, d.K, Opts()).em()
    po = np.asarray(m.pi, dtype=np.float64)
    assert got[0][1] == got[1][1] == len(m.fitinfo.iterations)      # same decision on both ranks = the single-process one
    assert got[0][2] == got[1][2]                                   # bit-identical parameters on both ranks
    ok = po > 1e-290
    assert np.max(np.abs(np.array(got[0][2])[ok] - po[ok]) / po[ok]) < 1e-9


def test_wide_models_are_split_by_rows_and_left_out_of_the_packing():
    """Celltype sharding: a model that whole-model packing (LPT) cannot balance is fitted by all
    ranks (rows cut into blocks of equal alignment count), the others are packed whole."""
    from stellarscope_b200 import parallel
    rng = np.random.default_rng(3)
    n_rows, n_models, world = 4000, 6, 4
    lens = rng.integers(2, 9, n_rows)
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    seg = rng.choice(n_models, size=n_rows, p=[0.55, 0.2, 0.1, 0.07, 0.05, 0.03]).astype(np.int32)
    seg[rng.random(n_rows) < 0.02] = -1                                  # rows of no model
    shards, seg_A, wide = parallel.shard_models(indptr, seg, n_models, world, wide='auto', wide_cost=0, wide_gain=0.1)
    assert wide.tolist() == [0, 1]                                       # 55 % and 20 % against a fair share of 25 %
    assert sorted(np.concatenate(shards).tolist()) == [2, 3, 4, 5]       # every packed model on exactly one rank
    assert int(seg_A.sum()) == int(lens[seg >= 0].sum())
    # a wide model: the ranks' blocks partition its rows, in order, with balanced alignment counts
    rows = np.flatnonzero(seg == 0)
    blocks = [parallel.model_row_block(lens, rows, r, world) for r in range(world)]
    assert np.array_equal(np.concatenate(blocks), rows)
    loads = np.array([lens[b].sum() for b in blocks])
    assert loads.max() - loads.min() <= 2 * lens.max()
    # the fullest rank carries much less than with packing alone
    plain = parallel.makespan(seg_A, world, [], 0)
    assert plain == max(seg_A[b].sum() for b in parallel.lpt_shards(seg_A, world))
    assert parallel.makespan(seg_A, world, wide, 0) < 0.6 * plain
    # the lockstep cost of a wide model decides: expensive exchange -> fewer / no wide models
    assert parallel.wide_models(seg_A, world, cost=0.05 * seg_A.sum(), gain=0.03).tolist() == [0]
    assert parallel.wide_models(seg_A, world, cost=seg_A.sum(), gain=0.03).size == 0
    # BASELINE config 3 shape: 20 Dirichlet(0.5) cell types on 8 GPUs, largest 16.8 % (fair share 12.5 %) -- splitting
    # would cost more than the 1.34 x imbalance; a dominant type of 40 % is split
    c3 = (np.array([.168, .14, .13, .118, .08, .063, .048, .046, .045, .043, .033, .021, .019, .014, .012, .01, .005, .003,
                    .002, .002]) * 100e6).astype(np.int64)
    assert parallel.wide_models(c3, 8).size == 0
    dom = np.concatenate([[int(0.4 * 100e6)], (c3[1:] * 0.6 / c3[1:].sum() * 100e6).astype(np.int64)])
    assert parallel.wide_models(dom, 8).tolist() == [0]
    # switches: None / one rank split nothing; an explicit list is taken as it is; the two-value form is unchanged
    assert parallel.wide_models(seg_A, world, None).size == 0 and parallel.wide_models(seg_A, 1, 0).size == 0
    assert parallel.shard_models(indptr, seg, n_models, world, wide=[5, 2])[2].tolist() == [2, 5]
    assert parallel.shard_models(indptr, seg, n_models, world, wide=[])[2].size == 0
    assert len(parallel.shard_models(indptr, seg, n_models, world)) == 2


def test_cell_blocks_keep_the_rows_of_a_cell_on_one_rank():
    from stellarscope_b200 import parallel
    rng = np.random.default_rng(4)
    n_rows, n_cells, world = 6000, 150, 4
    lens = rng.integers(2, 9, n_rows)
    row_cell = rng.integers(0, n_cells, n_rows).astype(np.int32)         # rows of a cell are scattered (BAM order)
    rows = np.flatnonzero(rng.random(n_rows) < 0.7)                      # the rows of one model
    blocks = [parallel.model_cell_block(lens, rows, row_cell, r, world) for r in range(world)]
    assert np.array_equal(np.sort(np.concatenate(blocks)), rows)         # a partition of the model's rows
    owner = {}
    for r, b in enumerate(blocks):
        for c in np.unique(row_cell[b]):
            assert owner.setdefault(int(c), r) == r                      # no cell on two ranks
    loads = np.array([lens[b].sum() for b in blocks], dtype=float)
    assert loads.max() / loads.mean() < 1.1
    assert parallel.model_cell_block(lens, np.zeros(0, dtype=np.int64), row_cell, 0, world).size == 0


def _report_worker(rank, world, port, q, tmpdir):
    """output_report under model sharding: rank 0 alone writes, every rank returns after the files exist."""
    import types
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from stellarscope_b200 import model
    calls = []
    model.count_matrices = lambda st, tl: ({'best_exclude': None}, ['BC0'], ['__no_feature'])

    def fake_write(st, tl, counts, bc, ft):
        import time
        time.sleep(0.5)                              # the other rank must wait for this
        calls.append(rank)
        open(os.path.join(tmpdir, 'TE_counts.mtx'), 'w').write('x')
    model._write_report = fake_write
    st = types.SimpleNamespace(opts=types.SimpleNamespace(shard_models=True))
    model.output_report(st, None)
    q.put((rank, calls, os.path.exists(os.path.join(tmpdir, 'TE_counts.mtx'))))
    dist.barrier()
    dist.destroy_process_group()


def test_output_report_single_writer_world2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_report_worker, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == (0, [0], True)                  # rank 0 wrote
    assert got[1] == (1, [], True)                   # rank 1 did not write and returned after the file existed


def test_plan_local_rows_partitions_every_row():
    """Packed models, a wide model cut into blocks and rows of no model: every row on exactly one rank."""
    d = H.small_synth(seed=12, n_cells=40, n_alignments=9000, n_loci=200)
    seg = (d.row_cell % 5).astype(np.int32)
    seg[d.row_cell % 11 == 0] = -1                                # cells without a model
    for world in (1, 2, 3, 8):
        for wide in ([], [2], [0, 1, 2, 3, 4]):
            got = []
            for rank in range(world):
                p = parallel.plan_local_rows(d.indptr, seg, 5, world, rank, wide=wide, row_cell=d.row_cell)
                rows = p['rows']
                assert np.all(np.diff(rows) > 0)
                assert len(p['local_seg']) == len(rows)
                # local model ids index `mine`; rows of wide / no model carry -1
                fitted = p['local_seg'] >= 0
                assert np.array_equal(p['mine'][p['local_seg'][fitted]], seg[rows][fitted])
                for w, m in zip(p['wide'], p['wide_local']):
                    assert np.all(seg[rows][m] == w)
                got.append(rows)
            allr = np.concatenate(got)
            assert np.array_equal(np.sort(allr), np.arange(d.N)), (world, wide)
            if world > 1 and wide:
                # a wide model's rows are spread over the ranks
                w = wide[0]
                per_rank = [int((seg[r] == w).sum()) for r in got]
                assert min(per_rank) > 0


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n = 3 + 2 * rank                                           # ragged contributions; rank 0 of the second call: none
    g = parallel.gather_concat({'a': np.arange(n, dtype=np.int32) + 100 * rank,
                                'b': np.full((n, 2), rank + 0.5),
                                'c': (np.arange(n) + rank).astype(np.uint16)})
    g2 = parallel.gather_concat({'x': np.arange(0 if rank == 0 else 4, dtype=np.int64)})
    s = parallel.allreduce_sum_host(np.array([rank + 1, 10], dtype=np.int64))
    q.put((rank, {k: v.tolist() for k, v in g.items()}, g2['x'].tolist(), g2['counts'].tolist(), s.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_concat_world2_gloo():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g, x, xc, s in got:
        assert g['a'] == [0, 1, 2, 100, 101, 102, 103, 104]
        assert g['b'] == [[0.5, 0.5]] * 3 + [[1.5, 1.5]] * 5
        assert g['c'] == [0, 1, 2, 1, 2, 3, 4, 5]
        assert g['counts'] == [3, 5]
        assert x == [0, 1, 2, 3] and xc == [0, 4]
        assert s == [3, 20]


def _plan_inputs():
    d = H.small_synth(seed=21, n_cells=60, n_alignments=15000, n_loci=300)
    seg_sorted = d.row_cell.astype(np.int32)                     # rows of a cell are consecutive, cells ascending
    perm = np.random.default_rng(5).permutation(d.n_cells).astype(np.int32)
    seg_mixed = perm[d.row_cell]                                 # consecutive but not ascending: the generic route
    return d, seg_sorted, seg_mixed


def test_plan_of_a_sorted_row_map_equals_the_generic_plan(monkeypatch):
    """Row -> model map sorted by model (barcode-sorted input): the plan comes from binary searches and row ranges and
    is the one the passes over all rows give."""
    d, seg_sorted, seg_mixed = _plan_inputs()
    for world in (1, 2, 3, 8):
        for rank in range(world):
            fast = parallel.plan_local_rows(d.indptr, seg_sorted, d.n_cells, world, rank)
            with monkeypatch.context() as m:
                m.setattr(parallel, '_sorted_by_segment', lambda *a, **k: False)
                slow = parallel.plan_local_rows(d.indptr, seg_sorted, d.n_cells, world, rank)
            for k in ('rows', 'local_seg', 'mine', 'seg_A'):
                assert np.array_equal(fast[k], slow[k]), (world, rank, k)
    assert not parallel._sorted_by_segment(seg_mixed, 1, 0, False)
    assert parallel._sorted_by_segment(seg_sorted, 1, 0, False)


def _plan_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    d, seg_sorted, seg_mixed = _plan_inputs()
    out = {}
    for name, seg in (('sorted', seg_sorted), ('mixed', seg_mixed)):
        verdict = parallel._sorted_by_segment(seg, world, rank, True)
        p = parallel.plan_local_rows(d.indptr, seg, d.n_cells, world, rank, collective=True)
        ref = parallel.plan_local_rows(d.indptr, seg, d.n_cells, world, rank, collective=False)
        out[name] = (bool(verdict), all(np.array_equal(p[k], ref[k]) for k in ('rows', 'local_seg', 'mine', 'seg_A')),
                     p['rows'].tolist())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_collective_plan_world2_gloo():
    """The collective form of the plan (1/world slices + all-reduce; sortedness verdict combined over the ranks) gives
    every rank the plan of the stand-alone call, and the ranks' rows partition the matrix."""
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_plan_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d, _, _ = _plan_inputs()
    for name, want_sorted in (('sorted', True), ('mixed', False)):
        assert got[0][name][0] == got[1][name][0] == want_sorted
        assert got[0][name][1] and got[1][name][1]
        rows = np.sort(np.concatenate([got[0][name][2], got[1][name][2]]))
        assert np.array_equal(rows, np.arange(d.N))
