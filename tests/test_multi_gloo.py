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
