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


def segment_alignments(indptr, row_segment, n_segments):
    """int64[n_segments]: alignments of the rows of every model (rows with a negative model id count nowhere)."""
    import numpy as np
    out = np.zeros(int(n_segments), dtype=np.int64)
    load().segment_alignments(np.ascontiguousarray(indptr), np.ascontiguousarray(row_segment, dtype=np.int32), out)
    return out


def lpt_assign(seg_A, world):
    """int32[S]: rank of every model under longest-processing-time packing (larger first, ties by index; the
    least loaded rank, ties by rank) -- the assignment ``parallel.lpt_shards`` makes."""
    import numpy as np
    seg_A = np.ascontiguousarray(seg_A, dtype=np.int64)
    out = np.empty(len(seg_A), dtype=np.int32)
    load().lpt_assign(seg_A, int(world), out)
    return out


def rows_of_rank(row_segment, model_rank, local_id, rank):
    """(rows int64[n], local_seg int32[n]): the rows whose model is packed onto ``rank`` (ascending) and the
    local id of their model."""
    import numpy as np
    row_segment = np.ascontiguousarray(row_segment, dtype=np.int32)
    model_rank = np.ascontiguousarray(model_rank, dtype=np.int32)
    local_id = np.ascontiguousarray(local_id, dtype=np.int32)
    counts = np.zeros(64, dtype=np.int64)
    load().count_rows(row_segment, model_rank, int(rank), counts)
    off = np.zeros(64, dtype=np.int64)
    np.cumsum(counts[:-1], out=off[1:])
    n = int(counts.sum())
    rows = np.empty(n, dtype=np.int64)
    lseg = np.empty(n, dtype=np.int32)
    load().fill_rows(row_segment, model_rank, local_id, int(rank), off, rows, lseg)
    return rows, lseg


def rows_of_ranges(starts, ends):
    """(rows int64[n], local_seg int32[n]) for models whose rows are the ranges ``[starts[i], ends[i])`` (ascending,
    disjoint): the rows in order and the index of their model."""
    import numpy as np
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    ends = np.ascontiguousarray(ends, dtype=np.int64)
    off = np.zeros(len(starts) + 1, dtype=np.int64)
    np.cumsum(ends - starts, out=off[1:])
    n = int(off[-1])
    rows = np.empty(n, dtype=np.int64)
    lseg = np.empty(n, dtype=np.int32)
    load().fill_ranges(starts, ends, off, rows, lseg)
    return rows, lseg


def group_sizes(codes, n_groups):
    """int64[n_groups]: number of rows with every code (negative codes count nowhere) -- ``np.bincount`` without
    the int64 copy of 10^8 codes, on a few threads."""
    import numpy as np
    out = np.zeros(int(n_groups), dtype=np.int64)
    load().segment_alignments(np.zeros(0, dtype=np.int32), np.ascontiguousarray(codes, dtype=np.int32), out)
    return out


def row_extents(indptr, rows):
    """int64[n + 1]: prefix sums of the lengths of the given rows (the row pointers of the row-compacted matrix)."""
    import numpy as np
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    ptr = np.empty(len(rows) + 1, dtype=np.int64)
    load().row_extents(np.ascontiguousarray(indptr), rows, ptr)
    return ptr
