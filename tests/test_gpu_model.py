"""The reference-facing Python API (stellarscope_b200.model) against golden
outputs of the real reference (tests/golden, written by oracle/make_golden.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

from tests import helpers as H

pytestmark = pytest.mark.gpu

TIE_TOL = 1e-9


def _fit(name):
    from stellarscope_b200 import model
    g = H.load_golden(name)
    opts = H.golden_opts(g, cls=H.StOpts)
    st = H.make_st(g, opts)
    tl, poolinfo = model._fit_pooling_model(st, opts)
    return g, opts, st, tl, poolinfo


@pytest.mark.parametrize('name', H.golden_names())
def test_fit_pooling_model_matches_reference(name):
    g, opts, st, tl, poolinfo = _fit(name)
    keys = list(poolinfo.models_info.keys())
    assert len(keys) == int(g['n_models'])
    assert keys == (['pseudobulk'] if opts.pooling_mode == 'pseudobulk' else list(range(len(keys))))
    for i, k in enumerate(keys):
        fi = poolinfo.models_info[k]
        assert len(fi.iterations) == g['iters'][i], (name, i)
        assert fi.converged == bool(g['converged'][i]) and fi.reached_max == bool(g['reached_max'][i])
        assert fi.nobs == g['nobs'][i] and fi.nfeats == g['nfeats'][i] and fi.nparams == 2
        assert abs(fi.final_lnl - g['lnl'][i]) <= 1e-10 * abs(g['lnl'][i])
        d = np.array([x[2] for x in fi.iterations])
        dg = g['diffs'][i, :len(d)]
        assert np.all(np.abs(d - dg) <= 1e-12 + 1e-6 * dg)
        pi, theta = tl.segment_params(i)
        r, ok = H.rel_err(pi, g['pi'][i])
        assert r < 1e-6 and ok, (name, i, r)
        r, ok = H.rel_err(theta, g['theta'][i])
        assert r < 1e-6
    assert abs(float(poolinfo.total_lnl) - float(g['total_lnl'])) <= 1e-10 * abs(float(g['total_lnl']))
    assert poolinfo.total_obs == int(g['total_obs']) and poolinfo.total_params == int(g['total_params'])
    assert abs(float(poolinfo.AIC) - float(g['AIC'])) <= 1e-9 * abs(float(g['AIC']))
    assert abs(float(poolinfo.BIC) - float(g['BIC'])) <= 1e-9 * abs(float(g['BIC']))
    # z, in the CSR order of the fitted matrix
    z = tl.z
    assert z.shape == st.raw_scores.shape and z.dtype == np.float64
    assert np.max(np.abs(z.data - g['z'])) < 1e-9
    if opts.pooling_mode == 'pseudobulk':
        assert abs(tl.lnl - g['lnl'][0]) <= 1e-10 * abs(g['lnl'][0])
        r, ok = H.rel_err(np.asarray(tl.pi.toarray()).ravel(), g['pi'][0])
        assert r < 1e-6 and ok


@pytest.mark.parametrize('name', H.golden_names())
def test_reassign_and_counts_match_reference(name):
    from stellarscope_b200 import model
    g, opts, st, tl, poolinfo = _fit(name)
    ptr, idx, sc = H.golden_matrix(g)
    margin = H.top2_margin(ptr, g['z'])
    safe_rows = (margin > TIE_TOL) | (margin == 0)          # outside the tie tolerance, or exact ties
    assert safe_rows.all(), 'fixture has a near-tie inside the tolerance; parity is not claimed there'
    opts.rng = np.random.default_rng(2024)
    infos = model.reassign(st, tl)
    for mode in H.MODES:
        r = sp.csr_matrix(st.reassignments[mode])
        r.eliminate_zeros()
        r.sort_indices()
        assert r.shape == st.raw_scores.shape
        exp_dtype = np.float64 if mode == 'best_average' else np.int8
        assert st.reassignments[mode].dtype == exp_dtype
        assert H.csr_equal(r.indptr, r.indices, r.data, g[f'r_{mode}_indptr'], g[f'r_{mode}_indices'],
                           g[f'r_{mode}_data'], rtol=1e-15 if mode == 'best_average' else 0), (name, mode)
        ri, gi = infos[mode], g[f'r_{mode}_info']
        assert ri.assigned == gi[0] and ri.ambiguous == gi[1], (name, mode)
        assert (ri.unaligned if ri.unaligned is not None else -1) == gi[2]
        if f'r_{mode}_dist' in g.files:
            assert sorted(ri.ambigous_dist.items()) == [tuple(x) for x in g[f'r_{mode}_dist'].tolist()]
    counts, bc_list, ft_list = model.count_matrices(st, tl)
    assert bc_list == g['barcodes'].tolist()
    assert ft_list[0] == '__no_feature'
    for mode in H.MODES:
        c = counts[mode]
        exp = g[f'c_{mode}']
        assert c.shape == exp.shape
        if mode == 'best_average' or bool(g['opt_umi_counts']):
            assert c.dtype == np.float64
            assert np.allclose(np.asarray(c.todense()), exp, rtol=1e-12, atol=0), (name, mode)
        else:
            assert c.dtype == np.int32
            assert np.array_equal(np.asarray(c.todense()), exp), (name, mode)


def test_output_report_writes_mtx(tmp_path):
    import scipy.io
    from stellarscope_b200 import model
    g, opts, st, tl, poolinfo = _fit('individual')
    opts.reassign_mode = ['best_exclude', 'best_average']
    opts.outfile_path = lambda suffix: str(tmp_path / suffix)
    model.reassign(st, tl)
    model.output_report(st, tl)
    assert (tmp_path / 'barcodes.tsv').read_text().split() == g['barcodes'].tolist()
    m0 = scipy.io.mmread(str(tmp_path / 'TE_counts.mtx'))
    assert np.array_equal(np.asarray(m0.todense()), g['c_best_exclude'])
    head = open(tmp_path / 'TE_counts.mtx').read(300)
    assert head.startswith('%%MatrixMarket matrix coordinate integer general') and 'PN: stellarscope' in head
    m1 = scipy.io.mmread(str(tmp_path / 'TE_counts.best_average.mtx'))
    assert np.allclose(np.asarray(m1.todense()), g['c_best_average'])


def test_error_behaviour():
    from stellarscope_b200 import model
    g, opts, st, tl, poolinfo = _fit('pseudobulk')
    with pytest.raises(ValueError):
        tl.reassign('best_guess')
    with pytest.raises(model.StellarscopeError):
        tl.reassign('best_conf', 0.5)
    opts2 = H.golden_opts(g, cls=H.StOpts, ignore_umi=False)
    st2 = H.make_st(g, opts2)
    with pytest.raises(model.StellarscopeError):
        model._fit_pooling_model(st2, opts2)                 # "UMI corrected matrix not found"


def test_binmax_known_answers_through_ss_reassign():
    """reference tests/test_sparse_plus.py:73-81 through the reassign kernel."""
    import torch
    from stellarscope_b200.engine import Engine
    eng = Engine()
    m = sp.csr_matrix(np.array([[6, 0, 2], [0, 0, 3], [4, 5, 6], [5, 5, 1], [0, 0, 0]], dtype=np.float64))
    eng.set_matrix(m.indptr, m.indices, np.ones(m.nnz, dtype=np.uint16), 3)
    z = torch.from_numpy(m.data).to(eng.device)
    flag, nbest, info = eng.reassign('best_average', z, tie_tol=0.0)
    assert flag.cpu().numpy().tolist() == [1, 0, 1, 0, 0, 1, 1, 1, 0]
    assert nbest.cpu().numpy().tolist() == [1, 1, 1, 2, 0]
    assert info[:3].tolist() == [3, 1, 1]
    flag, nbest, info = eng.reassign('best_exclude', z, tie_tol=0.0)
    assert flag.cpu().numpy().tolist() == [1, 0, 1, 0, 0, 1, 0, 0, 0]
    eng.close()


@pytest.mark.parametrize('name', ['pseudobulk', 'pseudobulk_lnl', 'pseudobulk_priors'])
@pytest.mark.parametrize('peers', [True, False])
def test_row_sharded_dense_path_single_rank(name, peers):
    """The row-sharded (dense K-vector) fit of csrc/ss_dense.cu on ONE rank -- persistent kernel
    (peers=True) and host-loop variant (peers=False) -- against the golden pseudobulk fits of the
    real reference.  The multi-rank exchange is covered by tests/tools/dist_check.py (2+ GPUs)."""
    from stellarscope_b200 import model
    g = H.load_golden(name)
    opts = H.golden_opts(g, cls=H.StOpts)
    st = H.make_st(g, opts)
    tl = model.TelescopeLikelihood(st.raw_scores, opts, shard=(0, 1, None, peers))
    tl.em()
    fi = tl.fitinfo
    assert len(fi.iterations) == g['iters'][0]
    assert fi.converged == bool(g['converged'][0]) and fi.reached_max == bool(g['reached_max'][0])
    assert fi.nobs == g['nobs'][0] and fi.nfeats == g['nfeats'][0]
    assert abs(fi.final_lnl - g['lnl'][0]) <= 1e-10 * abs(g['lnl'][0])
    d = np.array([x[2] for x in fi.iterations])
    dg = g['diffs'][0, :len(d)]
    assert np.all(np.abs(d - dg) <= 1e-12 + 1e-6 * dg)
    r, ok = H.rel_err(np.asarray(tl.pi.toarray()).ravel(), g['pi'][0])
    assert r < 1e-6 and ok
    r, ok = H.rel_err(tl.theta, g['theta'][0])
    assert r < 1e-6
    assert np.max(np.abs(tl.z.data - g['z'])) < 1e-9


def test_config1_reproduces_telescope_report():
    """BASELINE.json configs[0] through the CUDA path: pooling_mode=pseudobulk, reassign_mode=best_exclude
    (+ best_conf / total_hits / initial_unique) on the matrix of the bundled BAM + GTF, against the
    reference's own golden ``data/telescope_report.tsv`` (independent of the oracle)."""
    from stellarscope_b200 import model
    from tests.test_oracle import config1_report_check
    g = H.load_golden('config1')
    opts = H.golden_opts(g, cls=H.StOpts)
    opts.reassign_mode = ['best_exclude', 'best_conf', 'total_hits', 'initial_unique']
    st = H.make_st(g, opts)
    tl, poolinfo = model._fit_pooling_model(st, opts)
    assert len(poolinfo.models_info['pseudobulk'].iterations) == 16
    assert abs(tl.lnl - 94724.99236416568) <= 1e-10 * 94724.99236416568       # SURVEY.md section 6 probe
    model.reassign(st, tl)
    counts, bc, ft = model.count_matrices(st, tl)
    pi = np.asarray(tl.pi.toarray()).ravel()
    config1_report_check(g, pi, {m: np.asarray(counts[m].todense())[:, 0] for m in opts.reassign_mode})


def test_tie_rule_inside_and_outside_the_tolerance():
    """The device calls two alignments tied when z >= max * (1 - tie_tol) (relative, 1e-9 by default; the
    reference compares for exact equality, sparse_plus.py:171-174, and produces exact ties only from bit-identical
    parameters -- SURVEY.md 0.10).  A row whose runner-up lies INSIDE the margin without being equal is a tie here
    (ambiguous, dropped by best_exclude, split by best_average); just outside the margin it is assigned; with
    tie_tol = 0 the rule is the reference's."""
    import torch
    from stellarscope_b200.engine import Engine
    eng = Engine()
    rows = [[0.5, 0.5 * (1 - 1e-12), 1e-3],          # inside the tolerance, not equal
            [0.5, 0.5 * (1 - 1e-6), 1e-3],           # outside
            [0.4, 0.4, 0.2],                          # exact tie
            [0.7, 0.2, 0.1]]
    m = sp.csr_matrix(np.array(rows))
    eng.set_matrix(m.indptr, m.indices, np.ones(m.nnz, dtype=np.uint16), 3)
    z = torch.from_numpy(m.data.copy()).to(eng.device)
    flag, nbest, info = eng.reassign('best_exclude', z, tie_tol=1e-9)
    assert nbest.cpu().numpy().tolist() == [2, 1, 2, 1]
    assert flag.cpu().numpy().reshape(4, 3).tolist() == [[0, 0, 0], [1, 0, 0], [0, 0, 0], [1, 0, 0]]
    assert (int(info[0]), int(info[1]), int(info[2])) == (2, 2, 0)
    flag, nbest, info = eng.reassign('best_average', z, tie_tol=1e-9)
    assert flag.cpu().numpy().reshape(4, 3).tolist() == [[1, 1, 0], [1, 0, 0], [1, 1, 0], [1, 0, 0]]
    flag, nbest, info = eng.reassign('best_exclude', z, tie_tol=0.0)       # the reference's exact rule
    assert nbest.cpu().numpy().tolist() == [1, 1, 2, 1]
    assert (int(info[0]), int(info[1])) == (3, 1)
    eng.close()


def test_single_steps_and_init_parameters_match_the_oracle():
    """estep / mstep / calculate_lnl (reference model.py:202-323) and pi_init / theta_init (:336-337) of the drop-in
    class against the oracle's, on a model with unique and ambiguous rows; one hand-driven EM iteration equals the
    first iteration of em()."""
    from oracle.em_oracle import OracleLikelihood, Opts
    from stellarscope_b200 import model
    d = H.small_synth(seed=21, n_cells=6, n_alignments=4000, n_loci=150, unique_frac=0.15)
    opts = H.StOpts()
    m = sp.csr_matrix((d.scores, d.indices, d.indptr), shape=(d.N, d.K))
    tl = model.TelescopeLikelihood(m, opts)
    om = OracleLikelihood(d.indptr.astype(np.int64), d.indices.astype(np.int64), d.scores, d.K, Opts())
    pi0, th0 = np.repeat(1.0 / d.K, d.K), np.repeat(1.0 / d.K, d.K)
    z = tl.estep(sp.csr_matrix(pi0), th0)
    zo = np.asarray(om.estep(pi0, th0), dtype=np.float64)
    assert z.shape == (d.N, d.K) and np.max(np.abs(z.data - zo)) < 1e-12
    pi1, th1 = tl.mstep(z)
    pio, tho = om.mstep(zo)
    pio, tho = np.asarray(pio, dtype=np.float64).ravel(), np.asarray(tho, dtype=np.float64).ravel()
    assert sp.issparse(pi1) and pi1.shape == (1, d.K) and th1.shape == (d.K,)
    assert np.allclose(np.asarray(pi1.todense()).ravel(), pio, rtol=1e-12, atol=0)
    assert np.allclose(th1, tho, rtol=1e-12, atol=0)
    lnl = tl.calculate_lnl(z, pi1, th1)
    lo = float(om.calculate_lnl(zo, pio, tho))
    assert abs(lnl - lo) <= 1e-10 * abs(lo)
    tl.em()
    om.em()
    assert np.allclose(np.asarray(tl.pi_init.todense()).ravel(), np.asarray(om.pi_init, dtype=np.float64).ravel(), rtol=1e-9, atol=0)
    assert np.allclose(tl.theta_init, np.asarray(om.theta_init, dtype=np.float64).ravel(), rtol=1e-9, atol=0)
    assert np.allclose(np.asarray(tl.pi_init.todense()).ravel(), pio, rtol=1e-9, atol=0)     # the hand-driven iteration
    assert len(tl.fitinfo.iterations) == len(om.fitinfo.iterations)
    tl.close()
