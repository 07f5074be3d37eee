"""Checkpoint side-car -- drop-in for ``Stellarscope.save`` / ``Stellarscope.load`` (reference
stellarscope/utils/model.py:1259-1289; written after load_alignment, after dedup_umi and at the end of a
run, stages.py:97,159, stellarscope_assign.py:178; read by ``stellarscope resume``, stages.py:108).
SURVEY.md section 8f, rank 4.

The reference pickles the whole object.  At atlas scale almost all of that is five per-read containers of
Python strings (``read_index``, ``read_bcode_map``, ``read_umi_map``, ``bcode_ridx_map``, ``bcode_umi_map``),
which pickle walks object by object, and ``resume`` rebuilds them only to have the hot path turn them back
into integer arrays (row -> model, row -> output column, row -> (barcode, UMI) pair).

Here the checkpoint is a small pickle (everything else, unchanged) plus a side-car ``<name>.ssrows.npz``
holding the per-read containers as a *row table* -- the read names as one blob in row order, an integer
barcode / UMI code per row, the barcode / UMI strings once -- and the sparse matrices as their three arrays.
``load`` returns the object with the containers as lazy mappings over the row table: the hot path
(``pooling_row_segments``, ``_row_cells``, ``_row_outputs``, ``umi._pair_ids``) reads the integer arrays and
never builds a dict; any other access (``update_sam`` looks reads up by name, model.py:1455) materialises the
real dict on first use.  Checkpoints written by the reference (plain pickles) load as before.

Host-only: no device work, nothing here computes the model.
"""
from __future__ import annotations

import copy
import os
import pickle
from collections import defaultdict
from collections.abc import MutableMapping
from itertools import chain

import numpy as np
import scipy.sparse

MAGIC = 'stellarscope_b200.checkpoint/1'
ROW_MAPS = ('read_index', 'read_bcode_map', 'read_umi_map', 'bcode_ridx_map', 'bcode_umi_map')
MATRICES = ('raw_scores', 'corrected', 'umi_dups')


def sidecar_path(filename):
    return os.fspath(filename) + '.ssrows.npz'


class RowTable:
    """The per-read containers of ``Stellarscope`` as arrays (rows = ``read_index`` values 0..N-1).

    names_blob   read names in row order, newline-joined (BAM query names hold no newline)
    bcodes       barcodes in the key order of ``bcode_ridx_map``; row_bcode[r] = index into it
    umis         distinct UMIs in first-row order; row_umi[r] = index into it (None when UMIs are ignored)
    bu_ptr/bu_code  ``bcode_umi_map`` (barcode -> list of UMIs, one per stored fragment, model.py:1061) as CSR
    dirty        set when a lazy mapping was written to (or a container with mutable values was handed out):
                 the arrays may no longer describe the object and the hot path goes back to the dicts
    """

    def __init__(self, n_rows, names_blob, bcodes, row_bcode, umis=None, row_umi=None, bu_ptr=None, bu_code=None):
        self.N = int(n_rows)
        self.names_blob = names_blob
        self.bcodes = list(bcodes)
        self.row_bcode = np.ascontiguousarray(row_bcode, dtype=np.int32)
        self.umis = None if umis is None else list(umis)
        self.row_umi = None if row_umi is None else np.ascontiguousarray(row_umi, dtype=np.int32)
        self.bu_ptr, self.bu_code = bu_ptr, bu_code
        self.dirty = False
        self._names = None

    def names(self):
        if self._names is None:
            blob = self.names_blob
            if isinstance(blob, np.ndarray):
                blob = blob.tobytes()
            self._names = blob.decode().split('\n') if self.N else []
            if len(self._names) != self.N:
                raise ValueError('checkpoint side-car: read-name blob does not hold one name per row')
        return self._names

    # -- the containers, materialised (what pickle.load of the reference's checkpoint would have given) --
    def build(self, kind):
        if kind == 'read_index':
            return dict(zip(self.names(), range(self.N)))
        if kind == 'read_bcode_map':
            lab = np.array(self.bcodes + [None], dtype=object)[self.row_bcode].tolist()
            return dict(zip(self.names(), lab))
        if kind == 'read_umi_map':
            if self.row_umi is None:
                return {}
            lab = np.array(self.umis + [None], dtype=object)[self.row_umi].tolist()
            return dict(zip(self.names(), lab))
        if kind == 'bcode_ridx_map':
            out = defaultdict(set) if self.default_maps.get(kind, True) else {}
            order = np.argsort(self.row_bcode, kind='stable')
            bounds = np.searchsorted(self.row_bcode[order], np.arange(len(self.bcodes) + 1))
            for i, b in enumerate(self.bcodes):
                out[b] = set(order[bounds[i]:bounds[i + 1]].tolist())
            return out
        if kind == 'bcode_umi_map':
            out = defaultdict(list) if self.default_maps.get(kind, True) else {}
            if self.bu_ptr is not None:
                u = np.array(self.umis + [None], dtype=object)
                for i, b in enumerate(self.bcodes):
                    a, e = int(self.bu_ptr[i]), int(self.bu_ptr[i + 1])
                    if (e > a) if self.bu_keys is None else (b in self.bu_keys):
                        out[b] = u[self.bu_code[a:e]].tolist()
            return out
        raise KeyError(kind)

    bu_keys = None          # barcodes that are keys of bcode_umi_map (None: all with entries)
    default_maps = {}       # container kind -> it was a defaultdict (the reference's are, model.py:875-876)

    def n_keys(self, kind):
        if kind in ('read_index', 'read_bcode_map'):
            return self.N
        if kind == 'read_umi_map':
            return self.N if self.row_umi is not None else 0
        if kind == 'bcode_ridx_map':
            return len(self.bcodes)
        return None


class LazyMap(MutableMapping):
    """One of the per-read containers of a loaded checkpoint.  Behaves like the dict it stands for; the dict is
    built from the row table on first access.  Pickles as the plain dict (so the reference's own
    ``Stellarscope.save`` keeps working on a loaded object)."""
    __slots__ = ('_table', '_kind', '_d')
    _MUTABLE_VALUES = ('bcode_ridx_map', 'bcode_umi_map')

    def __init__(self, table, kind):
        self._table, self._kind, self._d = table, kind, None

    @property
    def materialised(self):
        return self._d is not None

    def _mat(self):
        if self._d is None:
            self._d = self._table.build(self._kind)
            if self._kind in self._MUTABLE_VALUES:
                self._table.dirty = True       # the caller may change the sets / lists it gets
        return self._d

    def __getitem__(self, k):
        return self._mat()[k]

    def __setitem__(self, k, v):
        self._table.dirty = True
        self._mat()[k] = v

    def __delitem__(self, k):
        self._table.dirty = True
        del self._mat()[k]

    def __iter__(self):
        return iter(self._mat())

    def __len__(self):
        if self._d is None:
            n = self._table.n_keys(self._kind)
            if n is not None:
                return n
        return len(self._mat())

    def __contains__(self, k):
        return k in self._mat()

    def __repr__(self):
        return f'LazyMap({self._kind}, {"materialised" if self._d is not None else "on the row table"})'

    def __reduce__(self):
        d = self._mat()
        if isinstance(d, defaultdict):
            return (defaultdict, (d.default_factory, dict(d)))
        return (dict, (d,))


def attach_row_table(st, table, kinds=None):
    """Give ``st`` the per-read containers of ``table`` as lazy mappings -- what ``load`` does for a side-car
    checkpoint.  For callers that hold the row arrays already (``stellarscope resume`` from a side-car, synthetic
    benchmark input): the hot path then reads the arrays and never builds a dict of 10^8 Python objects."""
    for k in (kinds or ROW_MAPS):
        setattr(st, k, LazyMap(table, k))
    return st


def row_table(st):
    """The row table of an object loaded from a side-car checkpoint, or None when the object's containers are
    (or may have become) something else."""
    m = getattr(st, 'read_index', None)
    t = m._table if isinstance(m, LazyMap) else None       # the object carries no attribute the reference's does not
    if t is None or t.dirty:
        return None
    for k in ROW_MAPS[:4]:
        m = getattr(st, k, None)
        if not isinstance(m, LazyMap) or m._table is not t:
            return None
    return t


# ---------------------------------------------------------------------------
# save
# ---------------------------------------------------------------------------
def _table_from_dicts(st):
    """Row table of the reference's containers, or None when they are not of the shape load_alignment builds
    (then the checkpoint is a plain pickle)."""
    from . import hostutil
    ri, rb, ru = st.read_index, st.read_bcode_map, st.read_umi_map
    groups = st.bcode_ridx_map
    if not (type(ri) is dict and type(rb) is dict and type(ru) is dict and isinstance(groups, dict)):
        return None
    N = len(ri)
    if len(rb) != N or (ru and len(ru) != N):
        return None
    if N and not np.array_equal(np.fromiter(ri.values(), dtype=np.int64, count=N), np.arange(N)):
        return None                                        # rows are numbered in insertion order (model.py:1045)
    bcodes = list(groups)
    if any(type(b) is not str or '\n' in b for b in bcodes):
        return None
    code = {b: i for i, b in enumerate(bcodes)}
    row_bcode = np.full(N, -1, dtype=np.int32)
    try:
        hostutil.row_labels(ri, rb, code, row_bcode)
        by_set = np.full(N, -1, dtype=np.int32)
        hostutil.fill_row_segments(list(groups.values()), np.arange(len(bcodes), dtype=np.int32), by_set)
    except (KeyError, ValueError, IndexError, TypeError):
        return None
    if not np.array_equal(by_set, row_bcode):              # the row sets say the same as the per-read map
        return None
    umis = row_umi = bu_ptr = bu_code = None
    bu_keys = None
    if ru:
        ids = {}
        row_umi = np.full(N, -1, dtype=np.int32)
        try:
            hostutil.row_labels(ri, ru, None, row_umi, ids)
        except (KeyError, TypeError):
            return None
        umis = list(ids)
        if any(type(u) is not str for u in umis):
            return None
        bum = getattr(st, 'bcode_umi_map', None)
        if bum:
            if not set(bum) <= set(code):
                return None
            lens = np.fromiter((len(bum.get(b, ())) for b in bcodes), dtype=np.int64, count=len(bcodes))
            bu_ptr = np.zeros(len(bcodes) + 1, dtype=np.int64)
            np.cumsum(lens, out=bu_ptr[1:])
            try:
                bu_code = np.fromiter(map(ids.__getitem__, chain.from_iterable(bum.get(b, ()) for b in bcodes)),
                                      dtype=np.int32, count=int(bu_ptr[-1]))
            except KeyError:
                return None
            bu_keys = set(bum)
    elif getattr(st, 'bcode_umi_map', None):
        return None
    names = list(ri)
    if any('\n' in q for q in names[:1000]):               # spot check; the join / split round trip is verified below
        return None
    blob = '\n'.join(names).encode()
    if blob.count(b'\n') != max(N - 1, 0):
        return None
    t = RowTable(N, blob, bcodes, row_bcode, umis, row_umi, bu_ptr, bu_code)
    t.bu_keys = bu_keys
    t._names = names
    t.default_maps = {k: isinstance(getattr(st, k, None), defaultdict) for k in ('bcode_ridx_map', 'bcode_umi_map')}
    for k, factory in (('bcode_ridx_map', set), ('bcode_umi_map', list)):
        m = getattr(st, k, None)
        if isinstance(m, defaultdict) and m.default_factory is not factory:
            return None                                    # not the reference's container
    return t


def _mat_arrays(prefix, m, out, meta):
    m = scipy.sparse.csr_matrix(m) if not scipy.sparse.isspmatrix_csr(m) else m
    out[prefix + '.data'], out[prefix + '.indices'], out[prefix + '.indptr'] = m.data, m.indices, m.indptr
    meta[prefix] = (tuple(int(x) for x in m.shape), type(m).__module__, type(m).__qualname__)


def _mat_from(prefix, z, meta):
    shape, mod, qual = meta[prefix]
    cls = scipy.sparse.csr_matrix
    try:
        import importlib
        c = importlib.import_module(mod)
        for part in qual.split('.'):
            c = getattr(c, part)
        if isinstance(c, type) and issubclass(c, scipy.sparse.csr_matrix):
            cls = c
    except Exception:
        pass
    return cls((z[prefix + '.data'], z[prefix + '.indices'], z[prefix + '.indptr']), shape=shape)


def save(st, filename):
    """``Stellarscope.save(self, filename)`` (model.py:1259-1274).  Returns True."""
    t = row_table(st)
    if t is None and not any(isinstance(getattr(st, k, None), LazyMap) for k in ROW_MAPS):
        t = _table_from_dicts(st) if all(hasattr(st, k) for k in ROW_MAPS[:4]) else None
    if t is None:
        with open(filename, 'wb') as fh:                   # containers of another shape: the reference's format
            pickle.dump(st, fh)
        return True
    arrays, meta = {}, {}
    light = copy.copy(st)
    for k in ROW_MAPS:
        if hasattr(st, k):
            setattr(light, k, None)
    for k in MATRICES:
        m = getattr(st, k, None)
        if m is not None and scipy.sparse.issparse(m):
            _mat_arrays('m.' + k, m, arrays, meta)
            setattr(light, k, None)
    rm = getattr(st, 'reassignments', None)
    if isinstance(rm, dict) and rm and all(scipy.sparse.issparse(v) for v in rm.values()):
        for mode, m in rm.items():
            _mat_arrays('r.' + mode, m, arrays, meta)
        light.reassignments = {mode: None for mode in rm}
    arrays['names'] = np.frombuffer(t.names_blob, dtype=np.uint8) if not isinstance(t.names_blob, np.ndarray) else t.names_blob
    arrays['row_bcode'] = t.row_bcode
    arrays['bcodes'] = np.frombuffer('\n'.join(t.bcodes).encode(), dtype=np.uint8)
    if t.row_umi is not None:
        arrays['row_umi'] = t.row_umi
        arrays['umis'] = np.frombuffer('\n'.join(t.umis).encode(), dtype=np.uint8)
    if t.bu_ptr is not None:
        arrays['bu_ptr'], arrays['bu_code'] = t.bu_ptr, t.bu_code
    token = os.urandom(8).hex()                            # ties the pickle to its side-car (checked by load)
    arrays['token'] = np.frombuffer(token.encode(), dtype=np.uint8)
    head = {'magic': MAGIC, 'token': token, 'n_rows': t.N, 'n_bcodes': len(t.bcodes), 'n_umis': None if t.umis is None else len(t.umis),
            'matrices': meta, 'maps': [k for k in ROW_MAPS if hasattr(st, k)],
            'bu_keys': None if t.bu_keys is None else sorted(t.bu_keys), 'default_maps': dict(t.default_maps)}
    side = sidecar_path(filename)
    # both files are written completely under temporary names first (a full disk or a crash while writing leaves
    # the previous checkpoint pair untouched), then put in place by two renames back to back
    tmp_side, tmp_pkl = side + '.tmp', str(filename) + '.tmp'
    try:
        with open(tmp_side, 'wb') as fh:
            np.savez(fh, **arrays)                         # stored, not deflated: the arrays go out at disk speed
        with open(tmp_pkl, 'wb') as fh:
            pickle.dump((MAGIC, head, light), fh)
    except BaseException:
        for f in (tmp_side, tmp_pkl):
            if os.path.exists(f):
                os.remove(f)
        raise
    os.replace(tmp_side, side)
    os.replace(tmp_pkl, filename)
    return True


# ---------------------------------------------------------------------------
# load
# ---------------------------------------------------------------------------
def _split(blob_arr, n):
    if n == 0:
        return []
    out = blob_arr.tobytes().decode().split('\n')
    if len(out) != n:
        raise ValueError('checkpoint side-car: string table of the wrong length')
    return out


def load(filename):
    """``Stellarscope.load(filename)`` (model.py:1276-1289): the object of a side-car checkpoint, or whatever a
    plain pickle (a checkpoint of the reference) holds."""
    with open(filename, 'rb') as fh:
        obj = pickle.load(fh)
    if not (isinstance(obj, tuple) and len(obj) == 3 and obj[0] == MAGIC):
        return obj
    _, head, st = obj
    side = sidecar_path(filename)
    if not os.path.exists(side):
        raise FileNotFoundError(f'{side}: the side-car of checkpoint {filename} is missing')
    with np.load(side) as z:
        if 'token' in head and ('token' not in z.files or z['token'].tobytes().decode() != head['token']):
            raise ValueError(f'{side} does not belong to checkpoint {filename} (written by another save)')
        bcodes = _split(z['bcodes'], head['n_bcodes'])
        umis = _split(z['umis'], head['n_umis']) if head['n_umis'] is not None else None
        t = RowTable(head['n_rows'], z['names'], bcodes, z['row_bcode'], umis,
                     z['row_umi'] if 'row_umi' in z.files else None,
                     z['bu_ptr'] if 'bu_ptr' in z.files else None, z['bu_code'] if 'bu_code' in z.files else None)
        t.bu_keys = None if head.get('bu_keys') is None else set(head['bu_keys'])
        t.default_maps = dict(head.get('default_maps') or {})
        for key in head['matrices']:
            m = _mat_from(key, z, head['matrices'])
            if key.startswith('m.'):
                setattr(st, key[2:], m)
            else:
                st.reassignments[key[2:]] = m
    for k in head['maps']:
        setattr(st, k, LazyMap(t, k))
    return st
