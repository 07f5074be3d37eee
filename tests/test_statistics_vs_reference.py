"""The result records of the path (FitInfo / PoolInfo / ReassignInfo / UMIInfo, reference
stellarscope/utils/statistics.py:269-625) against the reference's own classes filled with the same content: what
goes into the log and into stats.final.tsv (``output_stats`` :628-644 concatenates the ``to_dataframe()`` frames) must
be the same text.  Needs the reference checkout (build container); the GPU box has the goldens instead."""
import io
import logging
import os

import numpy as np
import pytest

from stellarscope_b200 import statistics as ours

pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/stellarscope'),
                                reason='reference checkout not present')


@pytest.fixture(scope='module')
def theirs():
    from oracle import refshim
    refshim.load()
    import stellarscope.utils.statistics as s
    return s


def _log_text(fn):
    buf = io.StringIO()
    h = logging.StreamHandler(buf)
    root = logging.getLogger()
    old = root.level
    root.addHandler(h)
    root.setLevel(logging.DEBUG)
    try:
        fn()
    finally:
        root.removeHandler(h)
        root.setLevel(old)
    return buf.getvalue()


def _frames_equal(a, b):
    assert list(a.columns) == list(b.columns)
    assert a.astype(str).values.tolist() == b.astype(str).values.tolist()


def _fill_fit(f, nobs, nfeats, lnl, iters, converged):
    f.nobs, f.nfeats, f.epsilon, f.max_iter = nobs, nfeats, 1e-7, 100
    f.final_lnl = lnl
    f.converged, f.reached_max = converged, not converged
    f.iterations = [(i + 1, 0.01 * i, 10.0 ** -(i + 1)) for i in range(iters)]


def test_fitinfo_and_poolinfo(theirs):
    pa, pb = ours.PoolInfo('individual'), theirs.PoolInfo('individual')
    for i, (nobs, nf, lnl, it, cv) in enumerate([(40, 7, 95123.25, 12, True), (3, 2, 17.5, 100, False),
                                                 (1200, 88, 4.2e6, 31, True)]):
        fa, fb = ours.FitInfo(), theirs.FitInfo()
        _fill_fit(fa, nobs, nf, lnl, it, cv)
        _fill_fit(fb, nobs, nf, lnl, it, cv)
        assert fa.fitted == fb.fitted and fa.nparams == fb.nparams
        _frames_equal(fa.to_dataframe(), fb.to_dataframe())
        pa.models_info[i], pb.models_info[i] = fa, fb
    pa.nmodels = pb.nmodels = 3
    assert pa.fitted_models == pb.fitted_models
    assert pa.total_obs == pb.total_obs and pa.total_params == pb.total_params
    assert float(pa.total_lnl) == float(pb.total_lnl)
    assert float(pa.AIC) == float(pb.AIC) and float(pa.BIC) == float(pb.BIC)
    _frames_equal(pa.to_dataframe(), pb.to_dataframe())
    assert _log_text(pa.log) == _log_text(pb.log)


@pytest.mark.parametrize('mode', ['best_exclude', 'best_conf', 'best_random', 'best_average', 'initial_unique',
                                  'initial_random', 'total_hits'])
def test_reassigninfo(theirs, mode):
    ra, rb = ours.ReassignInfo(mode), theirs.ReassignInfo(mode)
    for r in (ra, rb):
        r.assigned, r.ambiguous, r.unaligned = 5400, 321, 17
        if hasattr(rb, 'ambiguous_dist') or mode in ('best_exclude', 'best_random', 'best_average', 'initial_random'):
            r.ambiguous_dist = {0: 17, 1: 5400, 2: 300, 3: 21}
    assert ra.format() == rb.format()
    assert ra.format(include_mode=False, show_unaligned=True) == rb.format(include_mode=False, show_unaligned=True)
    _frames_equal(ra.to_dataframe(), rb.to_dataframe())
    assert _log_text(ra.log) == _log_text(rb.log)


def test_umiinfo(theirs):
    rng = np.random.default_rng(4)
    sizes = np.concatenate([np.ones(500, dtype=int), rng.integers(2, 9, 80), [14, 40]])
    groups = {(f'BC{i % 7}', f'U{i}'): {f'r{i}_{k}': None for k in range(int(n))} for i, n in enumerate(sizes)}
    ua, ub = ours.UMIInfo(sizes), theirs.UMIInfo(groups)
    for u in (ua, ub):
        u.ncomps_umi.update({1: 60, 2: 19, 3: 3})
        u.nexclude = 211
    assert ua.nexclude == ub.nexclude
    _frames_equal(ua.to_dataframe(), ub.to_dataframe())
    _frames_equal(ua.to_dataframe(binned=False), ub.to_dataframe(binned=False))
    assert _log_text(ua.loginit) == _log_text(ub.loginit)
    assert _log_text(ua.log) == _log_text(ub.log)


def test_output_stats_accepts_our_records(theirs, tmp_path):
    """``output_stats`` (statistics.py:628-644) of the reference on this package's records."""
    p = ours.PoolInfo('pseudobulk')
    f = ours.FitInfo()
    _fill_fit(f, 10, 4, 12.5, 3, True)
    p.models_info['pseudobulk'] = f
    p.nmodels = 1
    r = ours.ReassignInfo('best_exclude')
    r.assigned, r.ambiguous, r.unaligned = 7, 2, 1
    out = str(tmp_path / 'stats.final.tsv')
    theirs.output_stats([p, r], out)
    text = open(out).read()
    assert text.startswith('stage\tvar') or 'PoolInfo' in text
    assert 'total_lnl' in text and 'ReassignInfo' in text


def test_sparse_plus_file_format_and_equality(theirs, tmp_path):
    """csr_matrix_plus.save / load / check_equal: files written by either implementation are read by the other
    (reference sparse_plus.py:242-245, 398-405)."""
    import scipy.sparse as sp
    import stellarscope.utils.sparse_plus as ref_sp
    from stellarscope_b200 import sparse_plus as our_sp
    rng = np.random.default_rng(0)
    m = sp.random(40, 30, density=0.1, random_state=1, format='csr', dtype=np.float64)
    a, b = our_sp.csr_matrix_plus(m), ref_sp.csr_matrix_plus(m)
    fa, fb = str(tmp_path / 'a.npz'), str(tmp_path / 'b.npz')
    a.save(fa)
    b.save(fb)
    assert ref_sp.csr_matrix_plus.load(fa).check_equal(b)
    assert our_sp.csr_matrix_plus.load(fb).check_equal(a)
    assert a.check_equal(b) and b.check_equal(a)
    c = our_sp.csr_matrix_plus(m * 2)
    assert not a.check_equal(c) and not a.check_equal(our_sp.csr_matrix_plus(m[:, :29]))
