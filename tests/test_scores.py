"""Score-matrix construction (reference ``mapping_to_matrix``, model.py:958-978): the CPU
restatement and the CUDA path against golden outputs of the real reference (tests/golden/scores_*.npz,
oracle/make_golden_scores.py)."""
import glob
import os
import types

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import refshim
from oracle import scores_oracle as so

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NAMES = sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(GOLDEN, 'scores_*.npz')))


def _load(name):
    return np.load(os.path.join(GOLDEN, f'scores_{name}.npz'), allow_pickle=False)


def _same_csr(m, g):
    return (np.array_equal(m.indptr, g['indptr']) and np.array_equal(m.indices, g['indices'])
            and np.array_equal(m.data, g['data']) and m.data.dtype == np.uint16)


def test_golden_fixtures_present():
    assert NAMES == ['many_dups', 'mixed', 'wide']


@pytest.mark.parametrize('name', NAMES)
def test_scores_oracle_matches_reference_golden(name):
    g = _load(name)
    args = (g['read'], g['feat'], g['alnscore'], g['alnlen'], int(g['min_as']), int(g['n_reads']), int(g['n_feats']))
    m, nofeat = so.mapping_to_matrix(*args)
    assert _same_csr(m, g) and nofeat == 0
    m2 = so.mapping_to_matrix_loop(*args)
    assert _same_csr(m2, g)


def test_scores_oracle_counts_reads_without_feature():
    # read 1 only maps to column 0, read 2 has no record at all
    m, nofeat = so.mapping_to_matrix([0, 0, 1], [3, 0, 0], [10, 12, 11], [50, 50, 50], 10, 3, 5)
    assert nofeat == 2
    assert m.toarray().tolist() == [[53, 0, 0, 51, 0], [52, 0, 0, 0, 0], [0, 0, 0, 0, 0]]
    with pytest.raises(OverflowError):
        so.mapping_to_matrix([0], [1], [70000], [50], 10, 1, 2)


@pytest.mark.skipif(not refshim.available(), reason='reference checkout not present')
def test_scores_oracle_matches_live_reference():
    from oracle import make_golden_scores as mg
    ref = refshim.load()
    c = mg.make_case(seed=99, n_reads=150, n_loci=40, dup_frac=0.5)
    m_ref, _, _, _ = mg.run_reference(ref, c)
    m, nofeat = so.mapping_to_matrix(c['read'], c['feat'], c['alnscore'], c['alnlen'], c['min_as'], c['n_reads'], c['n_feats'])
    assert nofeat == 0
    assert np.array_equal(m.indptr, m_ref.indptr) and np.array_equal(m.indices, m_ref.indices)
    assert np.array_equal(m.data, m_ref.data)


def _tuples(g):
    fnames = ['__no_feature'] + [f'L{j:05d}' for j in range(1, int(g['n_feats']))]
    read_index, maps = {}, []
    for r, j, s, ln in zip(g['read'].tolist(), g['feat'].tolist(), g['alnscore'].tolist(), g['alnlen'].tolist()):
        q = f'read{r:07d}'
        read_index.setdefault(q, len(read_index))
        maps.append(('PM', q, fnames[j], s, ln))
    return maps, read_index, {f: i for i, f in enumerate(fnames)}


def test_code_mappings_host_helper():
    """The C walk of the loader's tuple list gives the dict look-ups of model.py:967-968 (host logic, no GPU)."""
    from stellarscope_b200 import loader
    g = _load('mixed')
    maps, read_index, feat_index = _tuples(g)
    read, feat, alnscore, alnlen = loader.code_mappings(maps, read_index, feat_index)
    assert np.array_equal(read, g['read']) and np.array_equal(feat, g['feat'])
    assert np.array_equal(alnscore, g['alnscore']) and np.array_equal(alnlen, g['alnlen'])
    # generic containers take the Python route with the same result
    r2 = loader.code_mappings(tuple(maps), read_index, feat_index)
    assert all(np.array_equal(a, b) for a, b in zip(r2, (read, feat, alnscore, alnlen)))
    with pytest.raises(KeyError):
        loader.code_mappings([('PM', 'nobody', '__no_feature', 1, 1)], read_index, feat_index)
    with pytest.raises(ValueError):
        loader.code_mappings([('PM', 'read0000000', '__no_feature', 1)], read_index, feat_index)


def _make_st(g):
    from tests import helpers as H
    maps, read_index, feat_index = _tuples(g)
    st = types.SimpleNamespace()
    st.opts = H.StOpts()
    st.opts.no_feature_key = '__no_feature'
    st.run_info = {}
    st.read_index = read_index
    st._load_sequential = lambda annot: (maps, types.SimpleNamespace(minAS=int(g['min_as'])))
    loci = {f: None for f in list(feat_index)[1:]}
    annot = types.SimpleNamespace(loci=loci, feature_length=lambda: {k: 1000 for k in loci})
    return st, annot


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_cuda_load_alignment_matches_reference_golden(name):
    """``stellarscope_b200.loader.load_alignment`` -- the drop-in for ``Stellarscope.load_alignment`` -- through
    the C ABI: bit-exact CSR."""
    from stellarscope_b200 import loader
    g = _load(name)
    st, annot = _make_st(g)
    alninfo = loader.load_alignment(st, annot)
    assert alninfo.minAS == int(g['min_as'])
    assert st.shape == (int(g['n_reads']), int(g['n_feats']))
    assert st.run_info['annotated_features'] == int(g['n_feats']) - 1
    m = st.raw_scores
    assert m.shape == st.shape and m.has_canonical_format
    assert _same_csr(m, g)


@pytest.mark.gpu
def test_cuda_scores_build_large_against_oracle_and_feeds_the_layout():
    """2M records, heavy repetition, shuffled order: bit-exact against the CPU restatement; the device-resident
    result is the engine's matrix and the layout builder accepts it as canonical."""
    from stellarscope_b200.engine import Engine
    rng = np.random.default_rng(5)
    n, n_reads, n_feats = 2_000_000, 300_000, 20_000
    read = rng.integers(0, n_reads, n).astype(np.int32)
    feat = (rng.integers(0, 40, n) + (read.astype(np.int64) * 7919) % (n_feats - 40)).astype(np.int32)
    alnscore = rng.integers(-30, 150, n).astype(np.int32)
    alnlen = rng.integers(20, 300, n).astype(np.int32)
    ref, nofeat = so.mapping_to_matrix(read, feat, alnscore, alnlen, -30, n_reads, n_feats)
    eng = Engine()
    indptr, indices, data, info = eng.scores_build(read, feat, alnscore, alnlen, -30, n_reads, n_feats)
    assert int(info[0]) == ref.nnz and int(info[3]) == nofeat
    assert np.array_equal(indptr, ref.indptr) and np.array_equal(indices, ref.indices) and np.array_equal(data, ref.data)
    assert eng.A == ref.nnz and eng.N == n_reads
    seg = (np.arange(n_reads) % 500).astype(np.int32)
    eng.build_layout(seg, 500)                       # canonical check of the builder passes on the device copy
    assert eng.layout_info()['nnz'] == ref.nnz
    eng.close()


@pytest.mark.gpu
def test_cuda_scores_build_edge_cases():
    from stellarscope_b200.engine import Engine, EngineError
    eng = Engine()
    z = np.zeros(0, dtype=np.int32)
    indptr, indices, data, info = eng.scores_build(z, z, z, z, 0, 5, 3, keep=False)       # no record at all
    assert indptr.tolist() == [0] * 6 and len(indices) == 0 and info.tolist() == [0, 0, 0, 5]
    one = lambda *v: np.array(v, dtype=np.int32)
    indptr, indices, data, info = eng.scores_build(one(2, 2, 0), one(1, 1, 0), one(5, 9, 5), one(10, 10, 10), 5, 4, 2,
                                                   keep=False)
    assert indptr.tolist() == [0, 1, 1, 2, 2] and indices.tolist() == [0, 1] and data.tolist() == [11, 15]
    assert info.tolist() == [2, 0, 0, 3]                                                # reads 0, 1, 3 lack a feature column
    with pytest.raises(OverflowError):
        eng.scores_build(one(0), one(1), one(70000), one(10), 0, 1, 2, keep=False)      # > uint16
    with pytest.raises(OverflowError):
        eng.scores_build(one(0), one(1), one(0), one(0), 1, 1, 2, keep=False)           # value 0 would not be stored
    with pytest.raises(EngineError):
        eng.scores_build(one(3), one(1), one(5), one(10), 0, 2, 2, keep=False)          # read id outside the matrix
    eng.close()
