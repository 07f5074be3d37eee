"""Checkpoint side-car: drop-in for Stellarscope.save / load (reference model.py:1259-1289).
The oracle is the reference's own format -- a pickle of the object: what ``load(save(st))`` gives must equal what
``pickle.loads(pickle.dumps(st))`` gives, attribute by attribute, and the integer arrays the hot path takes from a
loaded checkpoint must equal the ones it derives from the reference's dicts."""
import os
import pickle
import types
from collections import defaultdict

import numpy as np
import pytest
import scipy.sparse as sp

from stellarscope_b200 import checkpoint, model
from tests import helpers as H

ROW_MAPS = checkpoint.ROW_MAPS


def make_st(seed=3, n_cells=25, n_alignments=6000, umi=True, empty_barcode=False):
    """A Stellarscope-like object filled the way load_alignment fills it (store_read_info, model.py:1030-1061)."""
    d = H.small_synth(seed=seed, n_cells=n_cells, n_alignments=n_alignments, n_loci=120)
    rng = np.random.default_rng(seed)
    st = types.SimpleNamespace()
    st.opts = types.SimpleNamespace(ignore_umi=not umi, umi_counts=umi, pooling_mode='individual',
                                    outfile_path=lambda s: s)
    st.opts.outfile_path = None            # lambdas do not pickle; the reference's opts hold none
    st.run_info = {'version': 't', 'annotated_features': d.K - 1}
    st.feature_length = {f'L{j}': 100 + j for j in range(1, d.K)}
    st.feat_index = {('__no_feature' if j == 0 else f'L{j}'): j for j in range(d.K)}
    st.shape = (d.N, d.K)
    st.raw_scores = sp.csr_matrix((d.scores, d.indices, d.indptr), shape=(d.N, d.K))
    st.read_index, st.read_bcode_map, st.read_umi_map = {}, {}, {}
    st.bcode_ridx_map, st.bcode_umi_map = defaultdict(set), defaultdict(list)
    # reads arrive interleaved over the cells, a read may be stored more than once (several fragments)
    order = rng.permutation(d.N)
    cell_of = d.row_cell
    rows_new = np.empty(d.N, dtype=np.int64)
    for new_row, old in enumerate(order.tolist()):
        q = f'read:{old:07d}/x'
        bc = f'BC{int(cell_of[old]):04d}-1'
        u = 'ACGT'[int(rng.integers(4))] * 3 + f'{int(rng.integers(40)):02d}'
        for rep in range(1 + int(rng.integers(2))):
            row = st.read_index.setdefault(q, len(st.read_index))
            st.read_bcode_map.setdefault(q, bc)
            st.bcode_ridx_map[bc].add(row)
            if umi:
                st.read_umi_map.setdefault(q, u)
                st.bcode_umi_map[bc].append(u)
        rows_new[new_row] = old
    st.raw_scores = st.raw_scores[rows_new]          # row r of the matrix = r-th stored read
    if empty_barcode:
        st.bcode_ridx_map['BCEMPTY-1']                # a key without rows (defaultdict access)
    st.filtlist = {}
    st.bcode_ctype_map = {}
    st.ctype_bcode_map = defaultdict(set)
    st.celltypes = []
    st.corrected = None
    st.umi_dups = None
    st.reassignments = {}
    return st


def assert_same_object(a, b):
    assert list(a.__dict__) == list(b.__dict__)                # same attributes, nothing added
    for k, va in a.__dict__.items():
        vb = getattr(b, k)
        if sp.issparse(va):
            assert type(va) is type(vb) and va.shape == vb.shape and va.dtype == vb.dtype, k
            assert (va != vb).nnz == 0, k
        elif k == 'reassignments':
            assert list(va) == list(vb)
            for m in va:
                assert (va[m] != vb[m]).nnz == 0 and va[m].dtype == vb[m].dtype
        elif k == 'opts':
            assert va.__dict__ == vb.__dict__
        elif k in ROW_MAPS:
            da = va._mat() if isinstance(va, checkpoint.LazyMap) else va
            db = vb._mat() if isinstance(vb, checkpoint.LazyMap) else vb
            assert type(da) is type(db), k
            assert getattr(da, 'default_factory', None) is getattr(db, 'default_factory', None), k
            assert dict(va) == dict(vb), k
            assert list(va) == list(vb), k                     # same key order
        else:
            assert va == vb, k


@pytest.mark.parametrize('umi', [True, False])
def test_roundtrip_equals_pickle_roundtrip(tmp_path, umi):
    st = make_st(umi=umi, empty_barcode=True)
    st.corrected = st.raw_scores.multiply(sp.csr_matrix(np.ones((st.shape[0], 1), dtype=np.uint16))).tocsr()
    st.reassignments = {'best_exclude': sp.csr_matrix(st.raw_scores > 200).astype(np.int8),
                        'best_average': sp.csr_matrix(st.raw_scores, dtype=np.float64) / 7}
    f = str(tmp_path / 'checkpoint.load_alignment.pickle')
    assert checkpoint.save(st, f) is True
    assert os.path.exists(checkpoint.sidecar_path(f))
    assert os.path.getsize(f) < 20000                          # the pickle itself holds no per-read container
    ours = checkpoint.load(f)
    ref = pickle.loads(pickle.dumps(st))                       # the reference's save + load
    for k in ROW_MAPS:
        assert isinstance(getattr(ours, k), checkpoint.LazyMap) and not getattr(ours, k).materialised
        assert len(getattr(ours, k)) == len(getattr(ref, k)) or k == 'bcode_umi_map'
    assert_same_object(ours, ref)
    assert isinstance(ours.bcode_ridx_map['BC0001-1'], set)
    assert ours.bcode_ridx_map['never seen'] == set()          # defaultdict behaviour survives
    # a loaded object saves again (stages.py:159, stellarscope_assign.py:178) and pickles with the reference's save
    g = str(tmp_path / 'checkpoint.final.pickle')
    again = checkpoint.load(f)
    checkpoint.save(again, g)
    assert not any(getattr(again, k).materialised for k in ROW_MAPS)     # straight from the row table
    assert_same_object(checkpoint.load(g), ref)
    assert_same_object(pickle.loads(pickle.dumps(checkpoint.load(f))), ref)


def test_hot_path_arrays_from_table_equal_dict_route(tmp_path):
    st = make_st(seed=8, umi=True, empty_barcode=True)
    st.celltypes = ['A', 'B', 'none']
    names = list(st.bcode_ridx_map)
    st.ctype_bcode_map = defaultdict(set, {'A': set(names[:7]), 'B': set(names[7:15]) | {'absent-1'},
                                          'none': {'BCEMPTY-1'}})
    f = str(tmp_path / 'c.pickle')
    checkpoint.save(st, f)
    for mode in ('individual', 'celltype', 'pseudobulk'):
        opts = types.SimpleNamespace(pooling_mode=mode)
        ld = checkpoint.load(f)
        info = {}
        seg_t, keys_t = model.pooling_row_segments(ld, opts, info=info)
        seg_d, keys_d = model.pooling_row_segments(st, opts)
        assert np.array_equal(seg_t, seg_d) and keys_t == keys_d, mode
        if 'all_assigned' in info:             # what the sharded fit takes instead of a pass over the rows
            assert info['all_assigned'] == bool(len(seg_t) == 0 or seg_t.min() >= 0), mode
        assert checkpoint.row_table(ld) is not None and not ld.read_index.materialised
    ld = checkpoint.load(f)
    assert np.array_equal(model._row_cells(ld), model._row_cells(st))
    for umi_counts in (False, True):
        for filt in (None, {b: i for i, b in enumerate(reversed(names[:10]))}):
            for o in (st, ld):
                o.opts.umi_counts = umi_counts
                o.filtlist = filt or {}
            bc_list = sorted(filt, key=filt.get) if filt else sorted(st.bcode_ridx_map.keys())
            bc_pos = {b: i for i, b in enumerate(bc_list)}
            cell_t, umi_t = model._row_outputs(ld, bc_pos)
            cell_d, umi_d = model._row_outputs(st, bc_pos)
            assert np.array_equal(cell_t, cell_d)
            if umi_counts:                                     # ids are labels: same partition of the rows
                assert len(set(zip(umi_t.tolist(), umi_d.tolist()))) == len(set(umi_t.tolist())) == len(set(umi_d.tolist()))
    from stellarscope_b200 import umi as umimod
    g_t, pairs_t = umimod._pair_ids(ld)
    g_d, pairs_d = umimod._pair_ids(st)
    assert np.array_equal(g_t, g_d) and [tuple(p) for p in pairs_t] == [tuple(p) for p in pairs_d]
    assert not ld.read_index.materialised and not ld.read_umi_map.materialised


def test_dirty_table_falls_back_to_dicts(tmp_path):
    st = make_st(seed=5, umi=False)
    f = str(tmp_path / 'c.pickle')
    checkpoint.save(st, f)
    ld = checkpoint.load(f)
    ld.read_index['late read'] = len(ld.read_index)            # a write: the arrays no longer describe the object
    assert checkpoint.row_table(ld) is None
    ld = checkpoint.load(f)
    ld.bcode_ridx_map['BC0000-1']                              # a set was handed out: it may be changed
    assert checkpoint.row_table(ld) is None
    opts = types.SimpleNamespace(pooling_mode='individual')
    seg, keys = model.pooling_row_segments(ld, opts)           # container route on the materialised maps
    seg0, keys0 = model.pooling_row_segments(st, opts)
    assert np.array_equal(seg, seg0) and keys == keys0


def test_reference_pickles_and_odd_objects(tmp_path):
    st = make_st(seed=6, umi=True)
    f = str(tmp_path / 'legacy.pickle')
    with open(f, 'wb') as fh:
        pickle.dump(st, fh)                                    # what the reference's save writes
    ld = checkpoint.load(f)
    assert type(ld.read_index) is dict and checkpoint.row_table(ld) is None
    assert_same_object(ld, st)
    # containers of another shape (rows not numbered in insertion order): saved as a plain pickle, nothing lost
    odd = make_st(seed=6, umi=False)
    k0, k1 = list(odd.read_index)[:2]
    odd.read_index[k0], odd.read_index[k1] = odd.read_index[k1], odd.read_index[k0]
    g = str(tmp_path / 'odd.pickle')
    checkpoint.save(odd, g)
    assert not os.path.exists(checkpoint.sidecar_path(g))
    assert_same_object(checkpoint.load(g), odd)
    # a side-car of another save is refused
    h1, h2 = str(tmp_path / 'h1.pickle'), str(tmp_path / 'h2.pickle')
    checkpoint.save(st, h1)
    checkpoint.save(st, h2)
    os.replace(checkpoint.sidecar_path(h2), checkpoint.sidecar_path(h1))
    with pytest.raises(ValueError):
        checkpoint.load(h1)
    # a side-car that went missing is an error, not an empty object
    h = str(tmp_path / 'h.pickle')
    checkpoint.save(st, h)
    os.remove(checkpoint.sidecar_path(h))
    with pytest.raises(FileNotFoundError):
        checkpoint.load(h)


@pytest.mark.skipif(not os.path.isdir('/root/reference/stellarscope'), reason='reference checkout not present')
def test_real_reference_save_load(tmp_path):
    """The real ``Stellarscope.save`` / ``load`` (unmodified reference) against the side-car on the same object."""
    from oracle import refshim
    ref = refshim.load()
    st = make_st(seed=11, umi=True)
    obj = ref.Stellarscope.__new__(ref.Stellarscope)           # __init__ opens the BAM with pysam; the state is what counts
    obj.__dict__.update(st.__dict__)
    obj.raw_scores = ref.csr_matrix(obj.raw_scores)
    a, b = str(tmp_path / 'ref.pickle'), str(tmp_path / 'ours.pickle')
    assert ref.Stellarscope.save(obj, a) is True
    theirs = ref.Stellarscope.load(a)
    assert checkpoint.save(obj, b) is True
    ours = checkpoint.load(b)
    assert type(ours) is type(theirs)
    assert_same_object(ours, theirs)


def test_lazy_map_behaves_like_the_dict(tmp_path):
    st = make_st(seed=4, umi=True)
    f = str(tmp_path / 'c.pickle')
    checkpoint.save(st, f)
    ld = checkpoint.load(f)
    ri, ref = ld.read_index, st.read_index
    q = next(iter(ref))
    assert len(ri) == len(ref) and not ri.materialised          # len needs no dict
    assert q in ri and ri[q] == ref[q] and ri.get('nope') is None and ri.get(q) == ref[q]
    assert sorted(ri, key=ri.get) == sorted(ref, key=ref.get)   # the idiom of output_report (model.py:1389)
    assert list(ri.items()) == list(ref.items()) and list(ri.values()) == list(ref.values())
    assert ri == ref and dict(ri) == ref
    assert ri.setdefault(q, -5) == ref[q]
    with pytest.raises(KeyError):
        ri['nope']
    import copy
    assert copy.deepcopy(ld.read_bcode_map) == st.read_bcode_map
    assert ld.read_umi_map.pop(q) == st.read_umi_map[q] and q not in ld.read_umi_map
    assert checkpoint.row_table(ld) is None                     # that was a write
    assert 'read_index' in repr(ri)


def test_plain_dict_containers_keep_their_type(tmp_path):
    st = make_st(seed=2, umi=True)
    st.bcode_ridx_map = dict(st.bcode_ridx_map)                # e.g. an object somebody rebuilt by hand
    st.bcode_umi_map = dict(st.bcode_umi_map)
    f = str(tmp_path / 'c.pickle')
    checkpoint.save(st, f)
    ld = checkpoint.load(f)
    assert_same_object(ld, pickle.loads(pickle.dumps(st)))
    with pytest.raises(KeyError):
        ld.bcode_ridx_map['never seen']


@pytest.mark.skipif(not os.path.isdir('/root/reference/stellarscope'), reason='reference checkout not present')
def test_install_rebinds_the_reference_names(tmp_path):
    """The one-import binding of INTEGRATION.md against the real reference module, in a subprocess (install() rebinds
    names of the imported reference for good): every hot-path name points at this package afterwards, and the
    reference's own ``Stellarscope.save`` / ``load`` entry points go through the side-car."""
    import subprocess
    import sys
    import textwrap
    st = make_st(seed=13, umi=True)
    src = str(tmp_path / 'src.pickle')
    with open(src, 'wb') as fh:
        pickle.dump(st.__dict__, fh)
    code = textwrap.dedent(f'''
        import pickle, sys
        sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
        from oracle import refshim
        ref = refshim.load()
        import stellarscope_b200.model as b200
        from stellarscope_b200 import checkpoint
        assert b200.install() is ref
        assert ref.TelescopeLikelihood is b200.TelescopeLikelihood
        assert ref._fit_pooling_model is b200._fit_pooling_model and ref._em_wrapper is b200._em_wrapper
        import stellarscope
        assert b200.StellarscopeError is stellarscope.StellarscopeError
        import stellarscope.utils.statistics as ref_stats
        assert b200.ReassignInfo is ref_stats.ReassignInfo         # the reference's own record classes
        assert issubclass(b200.PoolInfo, ref_stats.PoolInfo) and b200.PoolInfo.log is ref_stats.PoolInfo.log
        import numpy as np
        from stellarscope_b200.statistics import SegmentFitInfos, PoolInfo as LocalPool
        res = dict(iters=np.array([3, 5]), flags=np.array([1, 2]), lnl=np.array([10.5, -2.25]), nobs=np.array([7, 9]),
                   nfeats=np.array([4, 6]), diffs=np.zeros((2, 5)))
        a, b = b200.PoolInfo('individual'), LocalPool('individual')
        a.models_info = SegmentFitInfos(res, 1e-7, 5, 0.0).keyed([0, 1])
        b.models_info = SegmentFitInfos(res, 1e-7, 5, 0.0).keyed([0, 1])
        assert (a.total_lnl, a.total_obs, a.total_params, a.AIC, a.BIC) == (b.total_lnl, b.total_obs, b.total_params, b.AIC, b.BIC)
        assert a.to_dataframe().equals(b.to_dataframe())
        obj = ref.Stellarscope.__new__(ref.Stellarscope)
        obj.__dict__.update(pickle.load(open({src!r}, 'rb')))
        for name in ('reassign', 'output_report', 'dedup_umi', 'load_alignment', 'save'):
            assert getattr(ref.Stellarscope, name).__module__ == 'stellarscope_b200.model', name
        f = {str(tmp_path / 'c.pickle')!r}
        assert obj.save(f) is True                              # stages.py:97 calls exactly this
        back = ref.Stellarscope.load(f)                         # stages.py:108
        assert type(back) is ref.Stellarscope and isinstance(back.read_index, checkpoint.LazyMap)
        assert dict(back.read_bcode_map) == obj.read_bcode_map
        seg_a, keys_a = b200.pooling_row_segments(back, back.opts)
        seg_b, keys_b = b200.pooling_row_segments(obj, obj.opts)
        assert (seg_a == seg_b).all() and keys_a == keys_b and not back.read_index.materialised
        print('install ok')
    ''')
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'install ok' in r.stdout, r.stdout + r.stderr
