"""CUDA EM path (through the C ABI) against the oracle on seeded inputs."""
import numpy as np
import pytest

from oracle.em_oracle import OracleLikelihood, Opts, submatrix_rows
from stellarscope_b200 import layout_host, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

PI_RTOL = 1e-6        # north_star: final pi/theta within 1e-6 relative in fp64
LNL_RTOL = 1e-10


def _fit_and_compare(d, row_seg, seg_rows, opts, use_import=False):
    from stellarscope_b200.engine import Engine
    S = len(seg_rows)
    eng = Engine()
    eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
    if use_import:
        eng.import_layout(layout_host.build(d.indptr, d.indices, d.scores, row_seg, S, d.K))
    else:
        eng.build_layout(row_seg, S)
    eng.fit(opts.em_epsilon, opts.max_iter, opts.pi_prior, opts.theta_prior, opts.use_likelihood)
    res = eng.results()
    z = eng.emit_z().cpu().numpy()
    ptr64, idx64 = d.indptr.astype(np.int64), d.indices.astype(np.int64)
    for s, rows in enumerate(seg_rows):
        ptr, idx, sc, take = submatrix_rows(ptr64, idx64, d.scores, rows)
        om = OracleLikelihood(ptr, idx, sc, d.K, opts).em()
        n_it = len(om.fitinfo.iterations)
        assert res['iters'][s] == n_it, f'segment {s}'
        assert bool(res['flags'][s] & 1) == om.fitinfo.converged
        assert bool(res['flags'][s] & 2) == om.fitinfo.reached_max
        assert res['nobs'][s] == om.fitinfo.nobs and res['nfeats'][s] == om.fitinfo.nfeats
        pi, th = eng.dense_params(s)
        r, ok = H.rel_err(pi, np.asarray(om.pi, dtype=np.float64))
        assert r < PI_RTOL and ok, f'pi segment {s}: {r}'
        r, ok = H.rel_err(th, np.asarray(om.theta, dtype=np.float64))
        assert r < PI_RTOL, f'theta segment {s}: {r}'
        lo = float(om.lnl)
        assert abs(res['lnl'][s] - lo) <= LNL_RTOL * abs(lo), f'lnl segment {s}'
        do = np.array([x[2] for x in om.fitinfo.iterations])
        dg = res['diffs'][s, :n_it]
        assert np.all(np.abs(dg - do) <= 1e-12 + 1e-6 * do), f'diff trace segment {s}'
        if len(take):
            assert np.max(np.abs(z[take] - np.asarray(om.z, dtype=np.float64))) < 1e-9
    eng.close()
    return res


@pytest.mark.parametrize('use_import', [False, True])
def test_individual_resident(use_import):
    d = H.small_synth(seed=21)
    _fit_and_compare(d, d.row_cell, d.cell_rows(), Opts(), use_import)


@pytest.mark.parametrize('use_import', [False, True])
def test_pseudobulk_streaming(use_import):
    d = H.small_synth(seed=22, n_alignments=25000)
    res = _fit_and_compare(d, np.zeros(d.N, dtype=np.int32), [np.arange(d.N)], Opts(), use_import)
    assert res['iters'][0] > 1


def test_pseudobulk_many_tiles_streaming_kernel():
    # > 8 tiles: the cooperative streaming kernel (segments of 2..8 tiles run as clusters)
    d = H.small_synth(seed=30, n_alignments=48000, n_loci=1500)
    from stellarscope_b200.engine import Engine
    _fit_and_compare(d, np.zeros(d.N, dtype=np.int32), [np.arange(d.N)], Opts())


def test_cluster_segments_of_every_size():
    # cells of 2..8 tiles next to single-tile cells
    sizes = [3000, 6000, 9000, 13000, 17000, 21000, 25000, 29000, 500, 800]
    parts = [H.small_synth(seed=40 + i, n_cells=1, n_alignments=a, n_loci=900, unique_frac=0.05) for i, a in enumerate(sizes)]
    indptr = [np.zeros(1, dtype=np.int64)]
    idx, sc, cell = [], [], []
    off = 0
    for c, p in enumerate(parts):
        indptr.append(p.indptr[1:].astype(np.int64) + off)
        off += p.A
        idx.append(p.indices)
        sc.append(p.scores)
        cell.append(np.full(p.N, c, dtype=np.int32))
    import types
    d = types.SimpleNamespace(indptr=np.concatenate(indptr).astype(np.int32), indices=np.concatenate(idx),
                              scores=np.concatenate(sc), K=parts[0].K, row_cell=np.concatenate(cell),
                              n_cells=len(parts))
    d.N = len(d.indptr) - 1
    rows = [np.flatnonzero(d.row_cell == c) for c in range(d.n_cells)]
    from stellarscope_b200.engine import Engine
    eng = Engine()
    eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
    eng.build_layout(d.row_cell, d.n_cells)
    info = eng.layout_info()
    assert info['n_cluster_segments'] >= 6 and info['n_cluster_tiles'] >= 20
    eng.close()
    _fit_and_compare(d, d.row_cell, rows, Opts())
    _fit_and_compare(d, d.row_cell, rows, Opts(use_likelihood=True, max_iter=40))


def test_celltype_mixed_sizes():
    d = H.small_synth(seed=23, n_cells=40, n_alignments=40000, n_celltypes=5, n_loci=800)
    _fit_and_compare(d, d.cell_type[d.row_cell].astype(np.int32), d.celltype_rows(), Opts())


def test_use_likelihood_resident_and_streaming():
    d = H.small_synth(seed=24, n_alignments=12000)
    _fit_and_compare(d, d.row_cell, d.cell_rows(), Opts(use_likelihood=True))
    _fit_and_compare(d, np.zeros(d.N, dtype=np.int32), [np.arange(d.N)], Opts(use_likelihood=True))


def test_priors_and_max_iter():
    d = H.small_synth(seed=25)
    _fit_and_compare(d, d.row_cell, d.cell_rows(), Opts(pi_prior=2, theta_prior=1000, max_iter=9))
    _fit_and_compare(d, d.row_cell, d.cell_rows(), Opts(theta_prior=0, pi_prior=0, em_epsilon=1e-4))


def test_all_unique_cells_and_tiny_cells():
    d = H.small_synth(seed=26, unique_frac=1.0, n_cells=6, n_alignments=200)
    _fit_and_compare(d, d.row_cell, d.cell_rows(), Opts())
    d = H.small_synth(seed=27, n_cells=60, n_alignments=200)
    _fit_and_compare(d, d.row_cell, d.cell_rows(), Opts())


def test_zero_denominator_is_model_error():
    """theta_prior = 0 and no ambiguous row -> 0/0 -> StellarscopeError class (model.py:267-272)."""
    from stellarscope_b200.engine import Engine
    indptr = np.array([0, 1, 2, 3, 5], dtype=np.int32)          # rows 0-2 unique, row 3 ambiguous
    indices = np.array([1, 2, 1, 0, 3], dtype=np.int32)
    scores = np.array([200, 180, 150, 170, 190], dtype=np.uint16)
    seg = np.array([0, 0, 1, 2], dtype=np.int32)
    eng = Engine()
    eng.set_matrix(indptr, indices, scores, 5)
    eng.build_layout(seg, 3)
    eng.fit(1e-7, 100, 0.0, 0.0, False)
    flags = eng.results()['flags']
    assert (flags[:2] & 4).all() and not (flags[2] & 4)
    eng.close()


def test_rows_outside_models_get_no_z():
    from stellarscope_b200.engine import Engine
    d = H.small_synth(seed=29)
    seg = d.row_cell.copy()
    seg[seg == 0] = -1
    eng = Engine()
    eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
    eng.build_layout(seg, d.n_cells)
    eng.fit()
    z = eng.emit_z().cpu().numpy()
    rows = np.flatnonzero(d.row_cell == 0)
    for r in rows:
        assert np.all(z[d.indptr[r]:d.indptr[r + 1]] == 0)
    # every fitted non-empty row is a probability vector
    sums = np.add.reduceat(z, d.indptr[:-1].astype(np.int64))
    fitted = (seg >= 0) & (np.diff(d.indptr) > 0)
    assert np.allclose(sums[fitted], 1.0, atol=1e-12)
    eng.close()


def test_full_size_properties_config2_shape():
    """Size-independent checks at a larger size (rows sum to 1, pi sums to 1, idempotent refit)."""
    from stellarscope_b200.engine import Engine
    d = synth.generate(n_cells=2000, n_alignments=1_000_000, n_loci=28000, seed=5)
    eng = Engine()
    eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
    eng.build_layout(d.row_cell, d.n_cells)
    eng.fit()
    r1 = eng.results()
    z = eng.emit_z().cpu().numpy()
    sums = np.add.reduceat(z, d.indptr[:-1].astype(np.int64))
    assert np.allclose(sums, 1.0, atol=1e-12)
    p = eng.params()
    for s in (0, 7, 1999):
        a, n = p['seg_col0'][s], p['seg_ncol'][s]
        tot = p['pi'][a:a + n].sum() + (d.K - n) * p['pi_rest'][s]
        assert abs(tot - 1.0) < 1e-9
    eng.fit()
    r2 = eng.results()
    assert np.array_equal(r1['iters'], r2['iters'])
    assert np.allclose(r1['lnl'], r2['lnl'], rtol=1e-12, atol=0)
    eng.close()


def test_long_rows_take_the_generic_paths():
    """Reads with up to 60 alignments (rows longer than 32 leave the lane groups) in small cells, one
    multi-tile cell and the pseudobulk model -- lane, cluster and streaming kernels against the oracle."""
    d = H.small_synth(seed=31, n_cells=8, n_alignments=30000, n_loci=400)
    # every 37th read becomes a read with 33..60 alignments on random loci
    rng = np.random.default_rng(31)
    idx, sc, ptr = [], [], [0]
    for i in range(d.N):
        a, b = d.indptr[i], d.indptr[i + 1]
        if i % 37 == 0:
            k = int(rng.integers(33, 61))
            cols = np.sort(rng.choice(np.arange(1, d.K), size=k, replace=False))
            idx.append(cols.astype(np.int32))
            sc.append((140 + rng.binomial(120, 0.65, k)).astype(np.uint16))
        else:
            idx.append(d.indices[a:b])
            sc.append(d.scores[a:b])
        ptr.append(ptr[-1] + len(idx[-1]))
    d.indptr, d.indices, d.scores = np.array(ptr, dtype=np.int32), np.concatenate(idx), np.concatenate(sc)
    assert np.diff(d.indptr).max() > 32
    _fit_and_compare(d, d.row_cell, d.cell_rows(), Opts())
    _fit_and_compare(d, np.zeros(d.N, dtype=np.int32), [np.arange(d.N)], Opts())


def test_non_canonical_input_is_canonicalised_like_the_reference():
    """Unsorted column indices, explicit zeros and duplicate entries: the device builder reports
    SS_ERR_NOT_CANONICAL, the host canonicalises (csr_matrix + eliminate_zeros, model.py:113-114) and the
    fit equals the fit of the canonical matrix."""
    import scipy.sparse as sp
    from stellarscope_b200 import model
    from stellarscope_b200.engine import Engine, NotCanonicalError
    d = H.small_synth(seed=32, n_cells=5, n_alignments=3000, n_loci=200)
    m = sp.csr_matrix((d.scores.astype(np.int64), d.indices, d.indptr), shape=(d.N, d.K))
    # scramble: reverse the order inside every row, add an explicit zero and split one entry in two
    idx, dat, ptr = [], [], [0]
    for i in range(d.N):
        a, b = d.indptr[i], d.indptr[i + 1]
        ci, cv = list(d.indices[a:b][::-1]), list(d.scores[a:b][::-1].astype(np.int64))
        if i % 7 == 0 and b > a:
            ci.append(ci[0]); cv.append(3); cv[0] -= 3          # duplicate column, same total
        if i % 5 == 0:
            free = [c for c in range(d.K) if c not in ci][:1]
            ci += free; cv += [0] * len(free)                    # explicit zero
        idx += ci; dat += cv; ptr.append(len(idx))
    bad = sp.csr_matrix((np.array(dat), np.array(idx), np.array(ptr)), shape=(d.N, d.K))
    assert not bad.has_sorted_indices
    eng = Engine()
    eng.set_matrix(bad.indptr, bad.indices, bad.data, d.K)
    with pytest.raises(NotCanonicalError):
        eng.build_layout(np.zeros(d.N, dtype=np.int32), 1)
    eng.close()
    o = H.StOpts()
    t_bad = model.TelescopeLikelihood(bad, o)
    t_ok = model.TelescopeLikelihood(m, o)
    t_bad.em()
    t_ok.em()
    assert len(t_bad.fitinfo.iterations) == len(t_ok.fitinfo.iterations)
    assert t_bad.lnl == t_ok.lnl
    assert np.array_equal(t_bad.raw_scores.indices, t_ok.raw_scores.indices)
    assert np.array_equal(np.asarray(t_bad.pi.todense()), np.asarray(t_ok.pi.todense()))
    assert np.array_equal(t_bad.z.data, t_ok.z.data)


def test_capacity_and_empty_inputs():
    from stellarscope_b200 import model, _lib
    from stellarscope_b200.engine import Engine, EngineError
    import scipy.sparse as sp
    # a read with more alignments than half a tile: explicit capacity error, not a wrong answer
    K = 3000
    indptr = np.array([0, 2500, 2502], dtype=np.int32)
    indices = np.concatenate([np.arange(2500), [1, 2]]).astype(np.int32)
    scores = np.full(2502, 150, dtype=np.uint16)
    eng = Engine()
    eng.set_matrix(indptr, indices, scores, K)
    with pytest.raises(EngineError) as ei:
        eng.build_layout(np.zeros(2, dtype=np.int32), 1)
    assert ei.value.status == _lib.SS_ERR_CAPACITY
    eng.close()
    # no alignments at all: the reference's M-step divides by zero -> StellarscopeError (model.py:267-272)
    empty = sp.csr_matrix((4, 10), dtype=np.uint16)
    with pytest.raises(model.StellarscopeError):
        model.TelescopeLikelihood(empty, H.StOpts()).em()
