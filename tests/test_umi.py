"""UMI de-duplication (reference model.py:561-668, 1148-1219): the CPU restatement and
the CUDA path against golden outputs of the real reference (tests/golden/umi_*.npz, oracle/make_golden_umi.py)."""
import glob
import os
from collections import Counter

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import umi_oracle as uo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NAMES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, 'umi_*.npz')))


def _load(name):
    g = np.load(os.path.join(GOLDEN, f'umi_{name}.npz'), allow_pickle=False)
    key = g['row_cell'].astype(np.int64) * (int(g['row_umi'].max()) + 1) + g['row_umi']
    _, row_group = np.unique(key, return_inverse=True)
    return g, row_group.astype(np.int32)


@pytest.mark.parametrize('name', NAMES)
def test_umi_oracle_matches_reference_golden(name):
    g, row_group = _load(name)
    r = uo.dedup(g['indptr'], g['indices'], g['scores'], row_group)
    assert np.array_equal(r['excluded'], g['excluded'])
    assert r['nexclude'] == int(g['nexclude'])
    assert sorted(r['ncomps_umi'].items()) == [tuple(x) for x in g['ncomps'].tolist()]
    assert sorted(r['rpu_counter'].items()) == [tuple(x) for x in g['rpu'].tolist()]
    ptr, idx, sc = uo.corrected_matrix(g['indptr'], g['indices'], g['scores'], r['excluded'])
    assert np.array_equal(ptr, g['c_indptr']) and np.array_equal(idx, g['c_indices']) and np.array_equal(sc, g['c_scores'])


def _make_st(g, tmp_path, opts_cls):
    import types
    N, K = len(g['indptr']) - 1, int(g['K'])
    st = types.SimpleNamespace()
    st.opts = opts_cls()
    st.opts.outfile_path = lambda suffix: str(tmp_path / suffix)
    st.raw_scores = sp.csr_matrix((g['scores'], g['indices'], g['indptr']), shape=(N, K))
    st.shape = (N, K)
    st.read_index = {f'read{i:06d}': i for i in range(N)}
    st.read_bcode_map = {f'read{i:06d}': f'BC{g["row_cell"][i]:03d}' for i in range(N)}
    st.read_umi_map = {f'read{i:06d}': f'UMI{g["row_umi"][i]:05d}' for i in range(N)}
    return st


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_cuda_dedup_umi_matches_reference_golden(name, tmp_path):
    """``stellarscope_b200.umi.dedup_umi`` -- the drop-in for ``Stellarscope.dedup_umi`` -- through the C ABI."""
    from stellarscope_b200 import umi
    from tests import helpers as H
    g, _ = _load(name)
    st = _make_st(g, tmp_path, H.StOpts)
    info = umi.dedup_umi(st, output_report=True, summary=False)
    dups = np.asarray(st.umi_dups.todense()).ravel().astype(bool)
    assert np.array_equal(dups, g['excluded'])
    cor = sp.csr_matrix(st.corrected)
    cor.eliminate_zeros()
    cor.sort_indices()
    assert st.corrected.shape == st.raw_scores.shape
    assert np.array_equal(cor.indptr, g['c_indptr']) and np.array_equal(cor.indices, g['c_indices'])
    assert np.array_equal(cor.data, g['c_scores'])
    assert info.nexclude == int(g['nexclude'])
    assert sorted(info.ncomps_umi.items()) == [tuple(x) for x in g['ncomps'].tolist()]
    assert sorted(info.rpu_counter.items()) == [tuple(x) for x in g['rpu'].tolist()]
    assert info.num_umi == sum(v for _, v in g['rpu'].tolist())
    assert (tmp_path / 'umi_tracking.txt').read_text() == str(g['report'])


@pytest.mark.gpu
def test_cuda_dedup_umi_large_against_oracle():
    """50k reads with heavy duplication through the engine call, row for row against the CPU restatement."""
    from stellarscope_b200.engine import Engine
    rng = np.random.default_rng(9)
    N, K = 50_000, 500
    lens = rng.integers(1, 7, N)
    indptr = np.zeros(N + 1, dtype=np.int32)
    np.cumsum(lens, out=indptr[1:])
    indices = np.concatenate([np.sort(rng.choice(K, size=l, replace=False)) for l in lens]).astype(np.int32)
    for i in np.flatnonzero(lens == 1):                    # every read overlaps a feature (model.py:969-971)
        if indices[indptr[i]] == 0:
            indices[indptr[i]] = 1
    scores = rng.choice(np.array([150, 160, 170, 180, 200], dtype=np.uint16), size=int(indptr[-1]))
    row_group = rng.integers(0, 12_000, N).astype(np.int32)
    ref = uo.dedup(indptr, indices, scores, row_group)
    eng = Engine()
    eng.set_matrix(indptr, indices, scores, K)
    out = eng.umi_dedup(row_group, 12_000)
    assert np.array_equal(out['excluded'], ref['excluded'])
    assert Counter(out['group_ncomp'][out['group_size'] > 1].tolist()) == ref['ncomps_umi']
    eng.close()
