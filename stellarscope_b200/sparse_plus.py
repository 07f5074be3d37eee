"""``csr_matrix_plus`` -- scipy CSR with the per-cell aggregation the report needs.

Mirror of the part of the reference's ``stellarscope/utils/sparse_plus.py``
class that the hot path's callers use on reassignment matrices
(``groupby_sum`` :367-396 via ``output_report`` model.py:1311, 1336;
``check_equal`` :242-245; ``save``/``load`` :398-405).  The aggregation runs on
the device (``ss_count``).  A matrix that ``TelescopeLikelihood.reassign`` just produced remembers the device
flags it was compacted from (``_ss_dev``), so that the aggregation that follows does not upload it again; the
attribute is dropped on pickling and is not inherited by derived matrices, so instances still pickle like plain
scipy matrices (``Stellarscope.save``, model.py:1259-1274).
"""
from __future__ import annotations

from collections.abc import Mapping

import numpy as np
import scipy.sparse


class StellarscopeError(Exception):
    pass


_engine_cache = {}


def _default_engine():
    from .engine import Engine
    if 'e' not in _engine_cache:
        _engine_cache['e'] = Engine()
    return _engine_cache['e']


def _rebuild(cls, data, indices, indptr, shape):
    return cls((data, indices, indptr), shape=shape)


class csr_matrix_plus(scipy.sparse.csr_matrix):

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop('_ss_dev', None)               # (engine, device flags): never pickled
        return state

    def __reduce_ex__(self, protocol):
        return (_rebuild, (type(self), self.data, self.indices, self.indptr, self.shape))

    def check_equal(self, other):
        """Same shape and no differing element (the reference's method of this name, sparse_plus.py:242-245)."""
        same_shape = tuple(self.shape) == tuple(other.shape)
        return bool(same_shape and (self != other).nnz == 0)

    _NPZ_FIELDS = ('data', 'indices', 'indptr', 'shape')       # file format of the reference's save (sparse_plus.py:398-405)

    def save(self, filename):
        np.savez(filename, **{k: getattr(self, k) for k in self._NPZ_FIELDS})

    @classmethod
    def load(cls, filename):
        with np.load(filename) as z:
            data, indices, indptr, shape = (z[k] for k in cls._NPZ_FIELDS)
        return cls((data, indices, indptr), shape=tuple(int(x) for x in shape))

    # -- device aggregation ----------------------------------------------------
    def _device_coo(self, row_group, row_umi, engine, keep_device=False):
        """COO (group, col, value) of sum over rows of each group, on the device."""
        dev = getattr(self, '_ss_dev', None)
        if (dev is not None and engine is not None and dev[0] is engine and getattr(engine, 'h', None)
                and engine._keep.get('row_ptr') is dev[2] and dev[1].numel() == engine.A):
            # produced by this engine's ss_reassign a moment ago: its flags are still on the device
            rows = dev[3] if len(dev) > 3 else None
            if rows is not None:                 # sharded run: the engine holds this rank's rows only
                row_group = np.ascontiguousarray(row_group[rows])
                row_umi = None if row_umi is None else np.ascontiguousarray(row_umi[rows])
            return engine.count(dev[1], row_group, row_umi=row_umi, keep_device=keep_device)
        eng = _default_engine()          # a separate handle: the model's engine keeps its score matrix
        m = self
        if not m.has_sorted_indices:
            m = m.copy()
            m.sort_indices()
        eng.set_matrix(m.indptr, m.indices, np.zeros(m.nnz, dtype=np.uint16), m.shape[1])
        flag = eng._to_dev((m.data != 0).astype(np.uint8) if m.nnz else np.zeros(1, np.uint8), np.uint8)
        weight = None
        if not np.issubdtype(m.dtype, np.integer) or (m.nnz and (m.data != 1).any()):
            weight = eng._to_dev(m.data.astype(np.float64), np.float64)
        return eng.count(flag, row_group, weight=weight, row_umi=row_umi, keep_device=keep_device)

    def _count_cells_fast(self, row_cell, n_cells, engine, cache=None):
        """Integer counts of a matrix whose selection flags are still on the engine's device (``_ss_dev``): the
        sort-free per-cell kernel (``ss_count_cells``) hands back the CSC arrays of the (K x n_cells) matrix
        directly -- int32 counts, no (cell, locus) key per entry to sort or to bring home."""
        dev = getattr(self, '_ss_dev', None)
        if not (dev is not None and engine is not None and dev[0] is engine and getattr(engine, 'h', None)
                and engine._keep.get('row_ptr') is dev[2] and dev[1].numel() == engine.A and engine.A > 0):
            return None
        rows = dev[3] if len(dev) > 3 else None
        rc = cache.get('row_cell_dev') if cache is not None else None
        if rc is None:
            if rows is not None:
                row_cell = np.ascontiguousarray(row_cell[rows])
            rc = engine._to_dev(row_cell, np.int32)
            if cache is not None:
                cache['row_cell_dev'] = rc
        nbest = dev[4] if len(dev) > 4 else None          # best_average: fractions 1 / nbest
        if nbest is None and not np.issubdtype(self.dtype, np.integer):
            return None
        res = engine.count_cells(dev[1], rc, n_cells, keep_device=True, nbest=nbest)
        if res is None:
            return None
        indptr, col, cnt, kept = res
        out = scipy.sparse.csc_matrix((cnt, col, indptr), shape=(self.shape[1], n_cells))
        if kept is not None:
            out._ss_dev_csc = kept           # (column pointers, loci, counts) on the device: mtx.mmwrite expands them
        return out

    def count_by_cell(self, row_cell, n_cells, row_umi=None, engine=None, cache=None):
        """(K x n_cells) matrix of per-cell column sums -- ``agg_bc`` /
        ``agg_bc_umi`` (model.py:1302-1345).  ``row_cell[i]`` = output column
        of row i (-1: dropped).  int32 for integer matrices, float64 otherwise
        and always float64 in UMI mode (model.py:1341)."""
        integer = np.issubdtype(self.dtype, np.integer) and row_umi is None
        if row_umi is None:
            out = self._count_cells_fast(np.ascontiguousarray(row_cell, dtype=np.int32), n_cells, engine, cache)
            if out is not None:
                return out
        cell, col, val, kept = self._device_coo(np.ascontiguousarray(row_cell, dtype=np.int32),
                                                None if row_umi is None else np.ascontiguousarray(row_umi, np.int32),
                                                engine, keep_device=True)
        dtype = np.int32 if integer else np.float64
        # the device returns the triples sorted by (cell, locus) without duplicates: that is the CSC form of the
        # (K x n_cells) matrix -- no host sort
        indptr = np.zeros(n_cells + 1, dtype=np.int64)
        np.cumsum(np.bincount(cell, minlength=n_cells), out=indptr[1:])
        out = scipy.sparse.csc_matrix((val.astype(dtype), col, indptr), shape=(self.shape[1], n_cells))
        if kept is not None:
            # the same triples, still on the device, in the storage order of `out` (cell-major): mtx.mmwrite formats
            # them without a second upload (MTX row = locus, MTX column = cell)
            out._ss_dev_coo = (kept[1], kept[0], kept[2])
        return out

    def groupby_sum(self, by, group_index: bool = True, dtype=np.float64, proc: int = 1):
        """Reference signature (sparse_plus.py:367-396): ``by`` maps row index ->
        group label; returns the (n_groups x K) matrix of column sums per group
        and the labels in first-seen order.  ``proc`` is ignored."""
        if not isinstance(by, Mapping):
            raise NotImplementedError('only Mapping `by` (like the reference)')
        labels = {}
        row_group = np.full(self.shape[0], -1, dtype=np.int32)
        for k, v in by.items():
            row_group[k] = labels.setdefault(v, len(labels))
        cell, col, val = self._device_coo(row_group, None, None)
        indptr = np.zeros(len(labels) + 1, dtype=np.int64)                  # sorted by (group, column): already CSR
        np.cumsum(np.bincount(cell, minlength=len(labels)), out=indptr[1:])
        mat = scipy.sparse.csr_matrix((val.astype(dtype), col, indptr), shape=(len(labels), self.shape[1]))
        if group_index:
            return type(self)(mat), list(labels.keys())
        return type(self)(mat)
