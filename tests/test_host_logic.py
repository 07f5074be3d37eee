"""Host-side logic that needs no GPU: layout specification invariants, pooling row maps,
random tie replay, result records, synthetic generator."""
import types

import numpy as np
import pytest

from oracle import em_oracle as eo
from stellarscope_b200 import layout_host, statistics, synth
from tests import helpers as H


def test_synth_is_deterministic_and_well_formed():
    a = H.small_synth(seed=3)
    b = H.small_synth(seed=3)
    assert np.array_equal(a.indices, b.indices) and np.array_equal(a.scores, b.scores)
    lens = np.diff(a.indptr)
    assert lens.min() >= 1 and a.scores.min() >= 140 and a.scores.max() <= 260
    for r in range(0, a.N, 97):                       # strictly ascending columns inside a row
        c = a.indices[a.indptr[r]:a.indptr[r + 1]]
        assert np.all(np.diff(c) > 0)
    assert np.all(np.diff(a.row_cell) >= 0)           # rows of a cell are contiguous


def test_layout_spec_roundtrip():
    """The tiles hold every alignment of every ambiguous row exactly once, JDS and column-major
    indices are consistent, unique rows are folded into pisum0."""
    d = H.small_synth(seed=4, n_alignments=12000)
    S = d.n_cells
    hl = layout_host.build(d.indptr, d.indices, d.scores, d.row_cell, S, d.K)
    lens = np.diff(d.indptr)
    seen = np.zeros(d.A, dtype=int)
    for t in range(hl.n_tiles):
        seg, nrows, nent, ncol, maxlen, nlong, eo_, ro, co, jo, rfrag, cfrag = hl.tile_hdr[t][:12]
        rgt, cgt = hl.tile_hdr[t][12:16], hl.tile_hdr[t][16:22]
        assert nent <= layout_host.E_HARD
        rl = hl.r_len[ro:ro + nrows].astype(int)
        assert np.all(np.diff(rl) <= 0) and rl[0] == maxlen and rl.sum() == nent
        jp = hl.j_ptr[jo:jo + maxlen + 1].astype(int)
        cp = hl.c_ptr[co:co + ncol + 1].astype(int)
        assert jp[0] == 0 and jp[-1] == nent and cp[0] == 0 and cp[-1] == nent
        cnt = np.diff(cp)
        assert np.all(np.diff(cnt) <= 0) and (cnt > 32).sum() == nlong
        assert rgt.tolist() == [(rl > x).sum() for x in (4, 8, 16, 32)]
        assert cgt.tolist() == [(cnt > x).sum() for x in (4, 8, 16, 32, 64, 128)]
        opos = hl.e_opos[eo_:eo_ + nent]
        seen[opos] += 1
        lcol = hl.e_lcol[eo_:eo_ + nent].astype(int)
        cpos = hl.e_cpos[eo_:eo_ + nent].astype(int)
        assert sorted(cpos.tolist()) == list(range(nent))
        assert np.all((cpos >= cp[lcol]) & (cpos < cp[lcol + 1]))
        scol = hl.c_scol[co:co + ncol]
        assert np.array_equal(hl.sc_gcol[scol][lcol], d.indices[opos])
        for r in range(nrows):
            row = hl.r_orow[ro + r]
            assert d.row_cell[row] == seg and lens[row] == rl[r]
            for p in range(rl[r]):
                assert opos[jp[p] + r] == d.indptr[row] + p
    amb_e = np.repeat(lens > 1, lens)
    assert np.array_equal(seen, amb_e.astype(int))
    # unique rows
    uni = np.flatnonzero(lens == 1)
    assert np.array_equal(hl.u_row, uni) and hl.sc_nuniq.sum() == len(uni)
    assert hl.seg_nobs.sum() == d.N and hl.seg_A.sum() == d.A


def test_layout_spec_matches_oracle_constants():
    d = H.small_synth(seed=5, n_cells=3, n_alignments=900)
    hl = layout_host.build(d.indptr, d.indices, d.scores, np.zeros(d.N, dtype=np.int32), 1, d.K)
    m = eo.OracleLikelihood(d.indptr, d.indices, d.scores, d.K, eo.Opts())
    assert hl.seg_smax[0] == m.max_score
    assert np.isclose(hl.seg_wtot[0], m.total_wt, rtol=1e-13) and np.isclose(hl.seg_wamb[0], m.ambig_wt, rtol=1e-13)
    assert np.isclose(hl.seg_wmax[0], m.w.max(), rtol=0)
    ps0 = np.zeros(d.K)
    ps0[hl.sc_gcol] = hl.sc_pisum0
    assert np.allclose(ps0, m.pisum0, rtol=1e-13)
    assert hl.seg_nobs[0] == m.fitinfo.nobs and hl.seg_ncol[0] == m.fitinfo.nfeats


def test_pooling_row_segments():
    from stellarscope_b200.model import pooling_row_segments, StellarscopeError
    st = types.SimpleNamespace(shape=(7, 3))
    st.bcode_ridx_map = {'B': {4, 5}, 'A': {0, 2}, 'E': set(), 'C': {6}}
    st.celltypes = ['t1', 't0', 't2']
    st.ctype_bcode_map = {'t0': ['A', 'Z'], 't1': ['C', 'B'], 't2': ['E']}
    o = types.SimpleNamespace(pooling_mode='pseudobulk')
    seg, keys = pooling_row_segments(st, o)
    assert seg.tolist() == [0] * 7 and keys == ['pseudobulk']
    o.pooling_mode = 'individual'                     # dict order, empty barcodes skipped (model.py:734-740)
    seg, keys = pooling_row_segments(st, o)
    assert seg.tolist() == [1, -1, 1, -1, 0, 0, 2] and keys == [0, 1, 2]
    o.pooling_mode = 'celltype'                       # st.celltypes order (model.py:710-720)
    seg, keys = pooling_row_segments(st, o)
    assert seg.tolist() == [1, -1, 1, -1, 0, 0, 0] and keys == [0, 1]
    st.ctype_bcode_map['t2'] = ['A']
    with pytest.raises(StellarscopeError):
        pooling_row_segments(st, o)


def test_random_replay_matches_oracle_choose_random():
    from stellarscope_b200.model import _replay_random_choice
    rng = np.random.default_rng(1)
    nb = rng.integers(0, 4, 300)
    lens = nb + rng.integers(0, 3, 300)
    ptr = np.concatenate([[0], np.cumsum(lens)])
    row = np.repeat(np.arange(300), lens)
    flag = np.zeros(ptr[-1], dtype=bool)
    for r in range(300):
        flag[ptr[r] + rng.choice(lens[r], nb[r], replace=False)] = True if nb[r] else False
    a = _replay_random_choice(flag, row, nb, np.random.default_rng(9))
    b = eo._choose_random(flag, row, nb, np.random.default_rng(9))
    assert np.array_equal(a, b)
    assert np.array_equal(np.bincount(row[a], minlength=300), (nb > 0).astype(int))


def test_info_records():
    p = statistics.PoolInfo('individual')
    for i, (lnl, nobs, nf) in enumerate([(10.5, 4, 3), (20.25, 6, 5)]):
        f = statistics.FitInfo(nobs=nobs, nfeats=nf, epsilon=1e-7, max_iter=100)
        f.final_lnl = lnl
        p.models_info[i] = f
    assert p.fitted_models == 2 and p.total_obs == 10 and p.total_params == 16
    assert float(p.total_lnl) == 30.75
    assert float(p.AIC) == 2 * 16 - 2 * 30.75
    assert np.isclose(float(p.BIC), 16 * np.log(10) - 2 * 30.75)
    assert set(p.to_dataframe()['var']) >= {'AIC', 'BIC', 'total_lnl', 'total_obs', 'total_params'}
    r = statistics.ReassignInfo('total_hits')
    r.assigned, r.ambiguous, r.unaligned = 7, 2, 1
    assert r.format() == ['total_hits - 7 total alignments.',
                          'total_hits - 2 reads initially had multiple alignments.']
    assert statistics.counter_from_hist([3, 0, 5]) == {0: 3, 2: 5}


def test_pooling_row_segments_walks_the_reference_containers():
    """row -> model map from ``bcode_ridx_map`` (dict of row sets) / celltype maps through the C helper
    (csrc/ss_hostutil.c): model ids in dict order, barcodes without reads get no model
    (reference model.py:735-739), a row claimed twice is an error."""
    import types
    from stellarscope_b200 import model
    st = types.SimpleNamespace(shape=(10, 4))
    st.bcode_ridx_map = {'B1': {0, 3}, 'EMPTY': set(), 'B2': {1, 2, 9}, 'B3': {5}}
    o = types.SimpleNamespace(pooling_mode='individual')
    seg, keys = model.pooling_row_segments(st, o)
    assert keys == [0, 1, 2]
    assert seg.tolist() == [0, 1, 1, 0, -1, 2, -1, -1, -1, 1]
    st.celltypes = ['t0', 't1', 't2']
    st.ctype_bcode_map = {'t0': ['B3', 'B1'], 't1': ['EMPTY', 'NOPE'], 't2': ['B2']}
    o.pooling_mode = 'celltype'
    seg, keys = model.pooling_row_segments(st, o)
    assert keys == [0, 1]                                  # t1 has no reads: not fitted
    assert seg.tolist() == [0, 1, 1, 0, -1, 0, -1, -1, -1, 1]
    st.bcode_ridx_map['B4'] = {9}
    o.pooling_mode = 'individual'
    with pytest.raises(model.StellarscopeError):
        model.pooling_row_segments(st, o)
    o.pooling_mode = 'pseudobulk'
    seg, keys = model.pooling_row_segments(st, o)
    assert keys == ['pseudobulk'] and not seg.any()
    # large input: the threaded walk gives the same map
    rng = np.random.default_rng(0)
    perm = rng.permutation(300000)
    cuts = np.sort(rng.choice(300000, 999, replace=False))
    sets = [set(x.tolist()) for x in np.split(perm, cuts)]
    st2 = types.SimpleNamespace(shape=(300000, 4), bcode_ridx_map={f'b{i}': s for i, s in enumerate(sets)})
    seg, keys = model.pooling_row_segments(st2, o.__class__(pooling_mode='individual'))
    exp = np.empty(300000, dtype=np.int32)
    for i, x in enumerate(np.split(perm, cuts)):
        exp[x] = i
    assert np.array_equal(seg, exp) and len(keys) == 1000


def test_segment_fit_infos_lazy_records():
    from stellarscope_b200.statistics import SegmentFitInfos, PoolInfo
    res = dict(iters=np.array([3, 5]), flags=np.array([1, 2]), lnl=np.array([1.5, 2.25]), nobs=np.array([10, 20]),
               nfeats=np.array([4, 6]), diffs=np.arange(20, dtype=float).reshape(2, 10))
    infos = SegmentFitInfos(res, 1e-7, 10, 0.5).keyed(['a', 'b'])
    assert list(infos.keys()) == ['a', 'b'] and len(infos) == 2
    fb = infos['b']
    assert fb.reached_max and not fb.converged and fb.nobs == 20 and fb.nfeats == 6 and fb.final_lnl == 2.25
    assert fb.iterations == [(i + 1, 0.5, 10.0 + i) for i in range(5)]
    p = PoolInfo('individual')
    p.models_info = infos
    assert float(p.total_lnl) == 3.75 and p.total_obs == 30 and p.total_params == 20 and p.fitted_models == 2
    assert abs(float(p.AIC) - (2 * 20 - 2 * 3.75)) < 1e-12


def test_row_outputs_c_walk_equals_the_python_route():
    """Row -> output column / UMI id maps of output_report (model.py:1389-1394): the C walk of the reference's dicts
    gives what the generator expressions over ``sorted(read_index)`` give."""
    import types
    from collections import OrderedDict
    from stellarscope_b200 import model
    rng = np.random.default_rng(8)
    N = 5000
    order = rng.permutation(N)                                   # dict order != row order
    read_index = {f'q{r:05d}': int(r) for r in order}
    bcs = [f'BC{c:03d}' for c in range(40)]
    read_bcode_map = {q: bcs[int(rng.integers(0, 40))] for q in read_index}
    read_umi_map = {q: f'U{int(rng.integers(0, 300)):04d}' for q in read_index}
    bc_pos = {b: i for i, b in enumerate(sorted(bcs[:35]))}      # 5 barcodes are filtered out -> -1
    for umi_counts in (False, True):
        st = types.SimpleNamespace(read_index=read_index, read_bcode_map=read_bcode_map, read_umi_map=read_umi_map,
                                   opts=types.SimpleNamespace(umi_counts=umi_counts))
        fast = model._row_outputs(st, bc_pos)
        st2 = types.SimpleNamespace(read_index=OrderedDict(read_index), read_bcode_map=read_bcode_map, read_umi_map=read_umi_map,
                                    opts=st.opts)                # not a plain dict -> Python route
        slow = model._row_outputs(st2, bc_pos)
        assert np.array_equal(fast[0], slow[0]) and (fast[0] == -1).sum() > 0
        if umi_counts:
            assert np.array_equal(fast[1], slow[1])
        else:
            assert fast[1] is None and slow[1] is None
    bad = types.SimpleNamespace(read_index={'nobody': 0}, read_bcode_map={}, read_umi_map={}, opts=types.SimpleNamespace(umi_counts=False))
    with pytest.raises(KeyError):
        model._row_outputs(bad, bc_pos)


def test_cluster_segments_are_cut_into_equal_tiles():
    """A segment of 2..8 large tiles (it runs as a thread-block cluster whose CTAs wait for the slowest one every
    iteration) is cut into EQUAL tiles of about E_CLUSTER alignments; single-tile and streamed segments keep their cut
    (csrc/ss_build.cu seg_chunk, mirrored by layout_host.build)."""
    rng = np.random.default_rng(3)
    sizes = [900, 4000, 5000, 9000, 20000, 30000, 40000]       # ambiguous alignments per cell (3 per row)
    indptr, indices, seg = [0], [], []
    for s, a in enumerate(sizes):
        for _ in range(a // 3):
            c = np.sort(rng.choice(400, size=3, replace=False)) + 1
            indices.extend(c.tolist())
            indptr.append(len(indices))
            seg.append(s)
    indptr = np.asarray(indptr, dtype=np.int32)
    indices = np.asarray(indices, dtype=np.int32)
    scores = rng.integers(140, 260, len(indices)).astype(np.uint16)
    hl = layout_host.build(indptr, indices, scores, np.asarray(seg, dtype=np.int32), len(sizes), 401)
    nent = {s: [] for s in range(len(sizes))}
    for t in range(hl.n_tiles):
        nent[int(hl.tile_hdr[t][0])].append(int(hl.tile_hdr[t][2]))
    chunk = hl.chunk
    for s, a in enumerate(sizes):
        a = a // 3 * 3
        assert sum(nent[s]) == a and max(nent[s]) <= layout_host.E_HARD
        if a <= chunk:
            assert len(nent[s]) == 1
        elif a <= layout_host.STREAM_SEG_TILES * chunk:
            want = min(-(-a // min(chunk, layout_host.E_CLUSTER)), layout_host.STREAM_SEG_TILES)
            assert len(nent[s]) == want, (a, nent[s])
            assert max(nent[s]) - min(nent[s]) <= 6, (a, nent[s])          # equal up to a row
        else:
            assert max(nent[s]) <= layout_host.E_STREAM                     # streamed: small tiles
