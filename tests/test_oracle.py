"""Pin the CPU restatement (oracle/em_oracle.py): reference known answers,
golden fixtures written by the real reference, and -- when the checkout is
present (build container) -- the reference itself on fresh inputs."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import em_oracle as eo
from oracle import refshim
from tests import helpers as H


# ---- known answers of the reference's own tests (stellarscope/tests/test_sparse_plus.py) ----
def _csr(dense):
    m = sp.csr_matrix(np.array(dense))
    return m.indptr.astype(np.int64), m.indices, m.data


def test_norm_rows_known_answer():
    # test_sparse_plus.py:47-53: [[1,0,2],[0,0,3],[4,5,6]] row-normalised
    ptr, idx, dat = _csr([[1, 0, 2], [0, 0, 3], [4, 5, 6]])
    got = eo.norm_rows(ptr, dat.astype(np.float64))
    exp = np.array([1 / 3., 2 / 3., 1., 4 / 15., 5 / 15., 6 / 15.])
    assert np.allclose(got, exp, rtol=1e-15)


def test_norm_rows_zero_row():
    # test_sparse_plus.py:55-63: a row of zeros stays zero
    ptr = np.array([0, 2, 2, 5])
    dat = np.array([1., 2., 4., 5., 6.])
    got = eo.norm_rows(ptr, dat)
    assert np.allclose(got, [1 / 3., 2 / 3., 4 / 15., 5 / 15., 6 / 15.])
    got0 = eo.norm_rows(np.array([0, 2]), np.array([0., 0.]))
    assert np.array_equal(got0, [0., 0.])


def test_binmax_rows_known_answer():
    # test_sparse_plus.py:73-81: [[6,0,2],[0,0,3],[4,5,6]] -> [[1,0,0],[0,0,1],[0,0,1]]
    ptr, idx, dat = _csr([[6, 0, 2], [0, 0, 3], [4, 5, 6]])
    flag = eo.binmax_rows(ptr, dat.astype(np.float64))
    assert flag.tolist() == [True, False, True, False, False, True]
    # ties are all kept
    ptr, idx, dat = _csr([[5, 5, 1], [0, 2, 2], [0, 0, 0]])
    assert eo.binmax_rows(ptr, dat.astype(np.float64)).tolist() == [True, True, False, True, True]


def test_rng_replay_equivalence():
    """rng.choice(range(a,b)) == a + rng.integers(0,b-a), scalar or vectorised, same final state
    (sparse_plus.py:235)."""
    n = np.random.default_rng(0).integers(2, 9, 500)
    r1 = np.random.default_rng(42)
    a = np.array([int(r1.choice(range(5, 5 + k))) - 5 for k in n])
    r2 = np.random.default_rng(42)
    b = r2.integers(0, n)
    assert np.array_equal(a, b)
    assert r1.bit_generator.state == r2.bit_generator.state


# ---- golden fixtures from the real reference ------------------------------------------------
@pytest.mark.parametrize('name', H.golden_names())
def test_oracle_matches_reference_golden(name):
    g = H.load_golden(name)
    opts = H.golden_opts(g)
    ptr, idx, sc = H.golden_matrix(g)
    res = eo.fit_pooling_model(ptr, idx, sc, int(g['K']), opts, segments=H.golden_segments(g))
    keys = list(res.fits.keys())
    assert len(keys) == int(g['n_models'])
    for i, k in enumerate(keys):
        fi = res.fits[k]
        m = res.models[k]
        assert len(fi.iterations) == g['iters'][i], (name, i)
        assert fi.converged == bool(g['converged'][i]) and fi.reached_max == bool(g['reached_max'][i])
        assert fi.nobs == g['nobs'][i] and fi.nfeats == g['nfeats'][i]
        assert abs(float(fi.final_lnl) - g['lnl'][i]) <= 1e-12 * abs(g['lnl'][i])
        r, ok = H.rel_err(np.asarray(m.pi, dtype=np.float64), g['pi'][i])
        assert r < 1e-9 and ok
        r, ok = H.rel_err(np.asarray(m.theta, dtype=np.float64), g['theta'][i])
        assert r < 1e-12
        d = np.array([x[2] for x in fi.iterations])
        dg = g['diffs'][i, :len(d)]
        assert np.all(np.abs(d - dg) <= 1e-12 + 1e-6 * dg)
    # z is stored in the pattern of the fitted matrix
    assert np.max(np.abs(res.z - g['z'])) < 1e-12
    assert abs(res.total_lnl - float(g['total_lnl'])) <= 1e-12 * abs(float(g['total_lnl']))
    assert res.total_obs == int(g['total_obs']) and res.total_params == int(g['total_params'])
    assert abs(res.AIC - float(g['AIC'])) <= 1e-9 * abs(float(g['AIC']))
    assert abs(res.BIC - float(g['BIC'])) <= 1e-9 * abs(float(g['BIC']))


@pytest.mark.parametrize('name', H.golden_names())
def test_oracle_reassign_and_counts_match_golden(name):
    g = H.load_golden(name)
    opts = H.golden_opts(g)
    ptr, idx, sc = H.golden_matrix(g)
    K = int(g['K'])
    # the reference reassigns on ret_model (full matrix) with the summed z
    full = eo.OracleLikelihood(ptr, idx, sc, K, opts)
    rng = np.random.default_rng(2024)
    names = [f'BC{c:04d}' for c in range(int(g['n_cells']))]
    row_bc = [names[c] for c in g['row_cell']]
    row_umi = [f'U{u:08d}' for u in g['row_umi']] if bool(g['opt_umi_counts']) else None
    bc_list = sorted(names)
    for mode in H.MODES:
        rp, ri, rd, info = full.reassign(mode, float(g['opt_conf_prob']), rng=rng, z=g['z'])
        assert H.csr_equal(rp, ri, rd, g[f'r_{mode}_indptr'], g[f'r_{mode}_indices'], g[f'r_{mode}_data'],
                           rtol=1e-15 if mode == 'best_average' else 0), (name, mode)
        ginfo = g[f'r_{mode}_info']
        assert info['assigned'] == ginfo[0] and info['ambiguous'] == ginfo[1]
        assert (info['unaligned'] if info['unaligned'] is not None else -1) == ginfo[2]
        if f'r_{mode}_dist' in g.files:
            assert sorted(info['ambigous_dist'].items()) == [tuple(x) for x in g[f'r_{mode}_dist'].tolist()]
        counts = eo.aggregate_counts(rp, ri, rd, K, row_bc, bc_list, row_umi=row_umi)
        assert np.allclose(counts, g[f'c_{mode}'], rtol=1e-12, atol=0), (name, mode)
    assert g['barcodes'].tolist() == bc_list


# ---- live reference (build container only) ----------------------------------------------------
@pytest.mark.skipif(not refshim.available(), reason='reference checkout not present')
@pytest.mark.parametrize('seed,uf', [(101, 0.0), (102, 0.3)])
def test_oracle_matches_live_reference(seed, uf):
    ref = refshim.load()
    d = H.small_synth(seed=seed, unique_frac=uf, n_cells=5, n_alignments=3000, n_loci=200)
    M = sp.csr_matrix((d.scores, d.indices, d.indptr), shape=(d.N, d.K))
    ro = refshim.RefOpts()
    with refshim.cli_fp_policy():
        tl = ref.TelescopeLikelihood(M, ro)
        tl.em()
    om = eo.OracleLikelihood(d.indptr, d.indices, d.scores, d.K, eo.Opts()).em()
    assert len(tl.fitinfo.iterations) == len(om.fitinfo.iterations)
    assert abs(float(tl.lnl) - float(om.lnl)) <= 1e-13 * abs(float(tl.lnl))
    r, ok = H.rel_err(np.asarray(om.pi, dtype=np.float64), np.asarray(tl.pi.toarray(), dtype=np.float64).ravel())
    assert r < 1e-10 and ok
    assert np.max(np.abs(np.asarray(om.z, dtype=np.float64) - tl.z.data.astype(np.float64))) < 1e-13


def test_empty_model_raises():
    with pytest.raises(eo.OracleError):
        eo.OracleLikelihood(np.array([0, 0, 0]), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.uint16), 5,
                            eo.Opts()).em()


# ---- config 1 (bundled BAM + GTF) against the reference's own golden report -------------------------
def config1_report_check(g, pi, counts):
    """``stellarscope/data/telescope_report.tsv`` (Telescope 1.0.2 on the bundled files) against a
    fit of the config-1 matrix: final_prop <-> pi (3 printed digits), final_count <-> best_exclude,
    final_conf <-> best_conf at 0.9, init_aligned <-> total_hits, unique_count <-> initial_unique
    (SURVEY.md 8c).  ``counts[mode]`` = per-locus counts of the single cell."""
    col = {n: j for j, n in enumerate(g['feat_names'].tolist())}
    checked = 0
    for i, name in enumerate(g['report_names'].tolist()):
        j = col[name]
        fp = float(g['report_final_prop'][i])
        assert abs(pi[j] - fp) <= 0.006 * fp + 1e-300, (name, pi[j], fp)
        assert counts['best_exclude'][j] == g['report_final_count'][i], name
        assert counts['best_conf'][j] == g['report_final_conf'][i], name
        assert counts['total_hits'][j] == g['report_init_aligned'][i], name
        assert counts['initial_unique'][j] == g['report_unique_count'][i], name
        checked += 1
    assert checked == 59                                    # telescope_report.tsv:3-61
    return checked


def test_oracle_config1_reproduces_telescope_report():
    g = H.load_golden('config1')
    assert g['indptr'][-1] == 18471 and len(g['indptr']) - 1 == 1000 and int(g['K']) == 100
    assert int(g['minAS']) == 240 and int(g['maxAS']) == 300
    opts = H.golden_opts(g)
    ptr, idx, sc = H.golden_matrix(g)
    K = int(g['K'])
    m = eo.OracleLikelihood(ptr, idx, sc, K, opts).em()
    assert len(m.fitinfo.iterations) == 16
    names = ['sample_01']
    counts = {}
    for mode in ('best_exclude', 'best_conf', 'total_hits', 'initial_unique'):
        rp, ri, rd, _ = m.reassign(mode, 0.9, rng=np.random.default_rng(0))
        counts[mode] = eo.aggregate_counts(rp, ri, rd, K, ['sample_01'] * (len(ptr) - 1), names)[:, 0]
    config1_report_check(g, np.asarray(m.pi, dtype=np.float64), counts)


@pytest.mark.skipif(not refshim.available(), reason='reference checkout not present')
def test_loader_restatement_reproduces_config1_fixture():
    """oracle/loader_oracle.py on the bundled BAM + GTF gives the committed fixture matrix."""
    import os
    from oracle.loader_oracle import load_alignment
    data = os.path.join(refshim.REFERENCE_ROOT, 'stellarscope', 'data')
    r = load_alignment(os.path.join(data, 'alignment.bam'), os.path.join(data, 'annotation.gtf'),
                       barcode_tag='RG', ignore_umi=True)
    g = H.load_golden('config1')
    m = r['matrix']
    assert np.array_equal(m.indptr, g['indptr']) and np.array_equal(m.indices, g['indices'])
    assert np.array_equal(m.data, g['scores'])
    assert sorted(r['feat_index'], key=r['feat_index'].get) == g['feat_names'].tolist()
    # with default tags every fragment is skipped as "Missing CB" (model.py:1099-1101): nothing to fit
    r0 = load_alignment(os.path.join(data, 'alignment.bam'), os.path.join(data, 'annotation.gtf'))
    assert r0['matrix'].shape[0] == 0
