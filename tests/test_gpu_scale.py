"""Oracle parity at the BASELINE.json shapes (K = 28,001 loci), through the public model API.

The small-size parity tests (test_gpu_model.py, test_gpu_em.py) run on <= 1500 loci and <= 48k alignments; here
the same bar -- iteration counts equal, lnL 1e-10, pi 1e-6, best_exclude counts bit-exact -- is applied
 * to configs[1] at full size (10k cells, 5M alignments, individual pooling): a stratified sample of cells that
   covers every kernel path (register-resident lane kernel classes, the multi-tile cell-stream kernel, the
   cooperative streaming kernel) incl. the largest cells;
 * to celltype(20) pooling and to pseudobulk on 2M alignments: EVERY model against the oracle (streaming kernels).
The oracle (numpy, oracle/em_oracle.py) runs one model per host core.
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
TIE_TOL = 1e-9


def _oracle_model(job):
    """(iters, lnl, pi, keep_exact, keep_tol) of one model; keep_*: best_exclude selection per stored alignment of
    the model's rows (exact ties as in the reference / ties within the stated relative tolerance)."""
    ptr, idx, sc, K, okw = job
    from oracle.em_oracle import OracleLikelihood, Opts
    m = OracleLikelihood(ptr, idx, sc, K, Opts(**okw)).em()
    z = np.asarray(m.z, dtype=np.float64)
    n = len(ptr) - 1
    rix = np.repeat(np.arange(n), np.diff(ptr))
    zmax = np.zeros(n)
    np.maximum.at(zmax, rix, z)
    keeps = []
    for tol in (0.0, TIE_TOL):
        best = (z >= zmax[rix] * (1.0 - tol)) & (z > 0)
        nb = np.bincount(rix[best], minlength=n)
        keeps.append(best & (nb[rix] == 1))
    return len(m.fitinfo.iterations), float(m.lnl), np.asarray(m.pi, dtype=np.float64).ravel(), keeps[0], keeps[1]


def _oracle_models(d, row_lists, okw):
    from oracle.em_oracle import submatrix_rows
    ptr64 = d.indptr.astype(np.int64)
    jobs, takes = [], []
    for rows in row_lists:
        ptr, idx, sc, take = submatrix_rows(ptr64, d.indices.astype(np.int64), d.scores, rows)
        jobs.append((ptr, idx, sc, d.K, okw))
        takes.append(take)
    cores = min(len(jobs), os.cpu_count() or 1)
    if cores > 1:
        with mp.get_context('fork').Pool(cores) as pool:
            out = pool.map(_oracle_model, jobs, 1)
    else:
        out = [_oracle_model(j) for j in jobs]
    return out, takes


def _counts_from_keep(d, takes, keeps, cells=None):
    """(K x n_cells) dense-able count matrix of the selected alignments (scipy sparse)."""
    import scipy.sparse as sp
    pos = np.concatenate([t[k] for t, k in zip(takes, keeps)]) if takes else np.zeros(0, dtype=np.int64)
    row_of = np.repeat(np.arange(d.N), np.diff(d.indptr))
    r = row_of[pos]
    return sp.coo_matrix((np.ones(len(pos), dtype=np.int64), (d.indices[pos], d.row_cell[r])),
                         shape=(d.K, d.n_cells)).tocsc()


def _st_opts(d, mode):
    import bench
    opts = bench.BenchOpts(mode)
    opts.shard_models = False
    return bench.make_st(d, opts), opts


def _check_models(pool, tl, oracle, model_ids):
    worst = 0.0
    for s, (it, lnl, pi_o, _, _) in zip(model_ids, oracle):
        fi = pool.models_info[s]
        assert fi.n_iterations == it, (s, fi.n_iterations, it)
        assert abs(fi.final_lnl - lnl) <= 1e-10 * abs(lnl), (s, fi.final_lnl, lnl)
        pi = tl.segment_params(s)[0]
        big = pi_o > H.PI_FLOOR
        worst = max(worst, float(np.max(np.abs(pi[big] - pi_o[big]) / pi_o[big])))
    assert worst < 1e-6, worst
    return worst


def _assert_counts(gpu, exact, tol, cols=None):
    if cols is not None:
        gpu, exact, tol = gpu[:, cols], exact[:, cols], tol[:, cols]
    a = (gpu != exact).nnz == 0
    b = (gpu != tol).nnz == 0
    assert a or b, 'best_exclude counts differ from the oracle under both tie rules'


def test_individual_configs1_full_size_vs_oracle():
    from stellarscope_b200 import model, synth
    d = synth.generate(n_cells=10_000, n_alignments=5_000_000, n_loci=28_000, seed=0)
    assert d.K == 28_001
    st, opts = _st_opts(d, 'individual')
    tl, pool = model._fit_pooling_model(st, opts)
    bounds = np.searchsorted(d.row_cell, np.arange(d.n_cells + 1))
    seg_A = np.diff(d.indptr.astype(np.int64)[bounds])
    rng = np.random.default_rng(5)
    # strata: size classes of the single-tile cells (lane kernel CTA sizes follow the size), multi-tile cells, the largest
    pick = set(np.argsort(-seg_A)[:3].tolist())
    cls = np.minimum(13, np.log2(np.maximum(seg_A, 1)).astype(int))
    for c in np.unique(cls):
        members = np.flatnonzero(cls == c)
        pick.update(rng.choice(members, size=min(10, len(members)), replace=False).tolist())
    pick = sorted(pick)
    info = tl._engine.layout_info()
    assert info['n_cluster_segments'] + info['n_stream_segments'] > 0        # multi-tile paths are exercised
    ntiles = tl._engine.export_layout_array('seg_ntiles', np.int32)
    assert (ntiles[pick] > 1).sum() >= 3
    oracle, takes = _oracle_models(d, [np.arange(bounds[s], bounds[s + 1]) for s in pick], bench_opts())
    _check_models(pool, tl, oracle, pick)
    model.reassign(st, tl)
    counts, bc, ft = model.count_matrices(st, tl)
    gpu = counts['best_exclude'].tocsc().astype(np.int64)
    exact = _counts_from_keep(d, takes, [o[3] for o in oracle])
    tol = _counts_from_keep(d, takes, [o[4] for o in oracle])
    _assert_counts(gpu, exact, tol, cols=np.array(pick))
    tl.close()


def bench_opts():
    import bench
    return dict(bench.OPTS)


@pytest.mark.parametrize('mode', ['celltype', 'pseudobulk'])
def test_streamed_models_2M_alignments_vs_oracle(mode):
    """celltype(20) and pseudobulk at 2M alignments, 28k loci, 10 % unique rows (pisum0 path): every model against
    the oracle, best_exclude counts of every cell bit-exact."""
    from stellarscope_b200 import model, synth
    d = synth.generate(n_cells=2000, n_alignments=2_000_000, n_loci=28_000, seed=3, n_celltypes=20, unique_frac=0.1)
    st, opts = _st_opts(d, mode)
    tl, pool = model._fit_pooling_model(st, opts)
    if mode == 'celltype':
        ct_row = d.cell_type[d.row_cell]
        row_lists = [np.flatnonzero(ct_row == t) for t in range(d.n_celltypes)]
        ids = list(range(d.n_celltypes))
        assert tl._engine.layout_info()['n_stream_tiles'] + tl._engine.layout_info()['n_cluster_tiles'] > 100
    else:
        row_lists = [np.arange(d.N)]
        ids = ['pseudobulk']
    oracle, takes = _oracle_models(d, row_lists, bench_opts())
    if mode == 'celltype':
        _check_models(pool, tl, oracle, ids)
    else:
        it, lnl, pi_o, _, _ = oracle[0]
        fi = pool.models_info['pseudobulk']
        assert fi.n_iterations == it and abs(fi.final_lnl - lnl) <= 1e-10 * abs(lnl)
        r, ok = H.rel_err(np.asarray(tl.pi.toarray()).ravel(), pi_o)
        assert r < 1e-6 and ok
    model.reassign(st, tl)
    counts, bc, ft = model.count_matrices(st, tl)
    gpu = counts['best_exclude'].tocsc().astype(np.int64)
    exact = _counts_from_keep(d, takes, [o[3] for o in oracle])
    tol = _counts_from_keep(d, takes, [o[4] for o in oracle])
    _assert_counts(gpu, exact, tol)
    assert gpu.sum() > 0.5 * d.N
    tl.close()


def test_rank_local_rows_equal_the_full_fit():
    """A rank of a sharded run holds only its rows (``row_ids``): the rows are cut out of the host matrix while they
    are staged for the upload (``Engine.set_matrix_rows`` -- at this size through the pinned ring with the chunked
    gather).  The models fitted from the cut equal the same models of the full fit, the device matrix equals the
    host-side cut, and the rank-local reassignment matrix has the global shape with the other rows empty."""
    import scipy.sparse as sp
    from stellarscope_b200 import hostutil, model, synth
    d = synth.generate(n_cells=3000, n_alignments=20_000_000, n_loci=28_000, seed=11)
    st, opts = _st_opts(d, 'individual')
    full, pool = model._fit_pooling_model(st, opts)
    it_full = np.array([pool.models_info[s].n_iterations for s in range(d.n_cells)])
    lnl_full = np.array([pool.models_info[s].final_lnl for s in range(d.n_cells)])
    full.close()
    mine = np.arange(1, d.n_cells, 2)                                   # "this rank's" cells: every second one
    local_id = np.full(d.n_cells, -1, dtype=np.int32)
    local_id[mine] = np.arange(len(mine), dtype=np.int32)
    rows = np.flatnonzero(local_id[d.row_cell] >= 0).astype(np.int64)
    m = sp.csr_matrix((d.scores, d.indices, d.indptr), shape=(d.N, d.K))
    tl = model.TelescopeLikelihood(m, opts, row_segment=local_id[d.row_cell[rows]], n_segments=len(mine), row_ids=rows)
    assert tl._engine.N == len(rows) and tl.N == d.N
    ptr, idx, dat = hostutil.gather_rows(d.indptr, d.indices, d.scores, rows)
    assert idx.nbytes > (32 << 20)                                       # the chunked ring route was taken
    k = tl._engine._keep
    assert np.array_equal(k['row_ptr'].cpu().numpy(), ptr.astype(np.int32))
    assert np.array_equal(k['col_idx'].cpu().numpy(), idx)
    assert np.array_equal(k['score'].cpu().numpy().view(np.uint16), dat)
    infos = tl.fit_segments()
    assert np.array_equal(np.array([infos[i].n_iterations for i in range(len(mine))]), it_full[mine])
    assert np.allclose(np.array([infos[i].final_lnl for i in range(len(mine))]), lnl_full[mine], rtol=1e-12, atol=0)
    mat, rinfo = tl.reassign('best_exclude')
    assert mat.shape == (d.N, d.K)
    lens = np.diff(mat.indptr)
    other = np.ones(d.N, dtype=bool)
    other[rows] = False
    assert lens[other].sum() == 0 and lens[rows].sum() == rinfo.assigned
    tl.close()
