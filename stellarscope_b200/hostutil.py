"""Loader of the CPython host helper ``_ss_hostutil`` (csrc/ss_hostutil.c).

Built in-tree with gcc on first use; host-side container walking only -- no
model arithmetic lives here.
"""
from __future__ import annotations

import importlib.util

from . import build as _build

_mod = None


def load():
    global _mod
    if _mod is None:
        if _build.hostutil_needs_build():
            _build.build_hostutil()
        spec = importlib.util.spec_from_file_location('_ss_hostutil', _build.HOSTUTIL)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mod = mod
    return _mod


def fill_row_segments(groups, ids, seg):
    """seg[row] = ids[i] for every row of the i-th container of ``groups``."""
    return load().fill_row_segments(groups, ids, seg)


def number_groups(groups, ids):
    """ids[i] = running index of groups[i] among the non-empty containers (-1: empty); returns their number."""
    return load().number_groups(groups, ids)


def pair_ids(read_index, bcode_map, umi_map, row_group):
    """row_group[row] = index of the read's (barcode, UMI) pair in the returned first-seen list."""
    return load().pair_ids(read_index, bcode_map, umi_map, row_group)


def code_mappings(mappings, read_index, feat_index, read, feat, alnscore, alnlen):
    """Integer-code the loader's (code, query_name, feat_name, alnscore, alnlen) tuples into int32 arrays."""
    return load().code_mappings(mappings, read_index, feat_index, read, feat, alnscore, alnlen)


def row_labels(read_index, label_of_read, position_of_label, out, ids=None):
    """out[read_index[q]] = position_of_label.get(label_of_read[q], -1); with ``ids`` (a dict) running ids in row order."""
    return load().row_labels(read_index, label_of_read, position_of_label, out, ids)


def gather_rows(indptr, indices, data, rows):
    """Row-compacted CSR ``(ptr int64[n+1], idx, dat)`` of the given rows (any order) of a host CSR matrix:
    one memcpy per run of consecutive rows, a few threads, no index array."""
    import numpy as np
    indptr, indices, data = (np.ascontiguousarray(a) for a in (indptr, indices, data))
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    ptr = np.empty(len(rows) + 1, dtype=np.int64)
    total = load().row_extents(indptr, rows, ptr)
    idx = np.empty(total, dtype=indices.dtype)
    dat = np.empty(total, dtype=data.dtype)
    load().gather_rows(indptr, indices, data, rows, ptr, idx, dat)
    return ptr, idx, dat
