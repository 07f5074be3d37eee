"""UMI de-duplication -- drop-in for ``Stellarscope.dedup_umi`` (reference stellarscope/utils/model.py:1148-1219;
``select_umi_representatives`` :561-668), the step right before the EM path.

The reference loops over every barcode+UMI pair in Python, builds an O(n^2) adjacency matrix per duplicate group and
calls scipy's connected components.  Here the host only integer-codes the (barcode, UMI) pairs; grouping,
components and the choice of the representatives happen in ``ss_umi_dedup`` (csrc/ss_umi.cu).  Nothing is
computed on the CPU; without a CUDA device the call fails.
"""
from __future__ import annotations

import logging as lg
from collections import Counter

import numpy as np
import scipy.sparse

from .engine import Engine
from .sparse_plus import csr_matrix_plus
from .statistics import UMIInfo


def _pair_ids(st):
    """Row -> id of its (barcode, UMI) pair, ids in first-seen order (the order of the reference's
    ``bcumi_read`` dict, model.py:1161-1164) and the pairs themselves."""
    from . import checkpoint, hostutil
    N = st.shape[0]
    table = checkpoint.row_table(st)
    if table is not None and table.row_umi is not None and N == table.N:
        # loaded from a side-car checkpoint: (barcode code, UMI code) per row -> pair ids in first-row order
        nu = max(len(table.umis), 1)
        key = table.row_bcode.astype(np.int64) * nu + table.row_umi
        uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
        order = np.argsort(first, kind='stable')
        rank = np.empty(len(uniq), dtype=np.int32)
        rank[order] = np.arange(len(uniq), dtype=np.int32)
        pairs = [(table.bcodes[k // nu], table.umis[k % nu]) for k in uniq[order].tolist()]
        return rank[inv.reshape(-1)].astype(np.int32), pairs
    row_group = np.full(N, -1, dtype=np.int32)
    if all(type(m) is dict for m in (st.read_index, st.read_bcode_map, st.read_umi_map)):
        pairs = hostutil.pair_ids(st.read_index, st.read_bcode_map, st.read_umi_map, row_group)   # C walk of the dicts
    else:
        ids = {}
        bmap, umap = st.read_bcode_map, st.read_umi_map
        for qname, row in st.read_index.items():
            row_group[row] = ids.setdefault((bmap[qname], umap[qname]), len(ids))
        pairs = list(ids)
    return row_group, pairs


def dedup_umi(st, output_report=True, summary=True):
    """Sets ``st.umi_dups`` (N x 1 indicator of the excluded reads), ``st.corrected`` (the score matrix with
    the excluded rows emptied, model.py:1216) and returns the ``UMIInfo`` record."""
    raw = scipy.sparse.csr_matrix(st.raw_scores)
    if not raw.has_sorted_indices:
        raw = raw.copy()
        raw.sort_indices()
    N, K = raw.shape
    row_group, pairs = _pair_ids(st)
    eng = Engine.acquire()
    try:
        eng.set_matrix(raw.indptr, raw.indices, raw.data, K)
        out = eng.umi_dedup(row_group, len(pairs))
    finally:
        eng.release()
    excluded, gsize, gcomp = out['excluded'], out['group_size'], out['group_ncomp']
    umiinfo = UMIInfo(gsize)
    umiinfo.loginit(lg.INFO if summary else lg.DEBUG)
    dupg = gsize > 1
    vals, freq = np.unique(gcomp[dupg], return_counts=True)
    umiinfo.ncomps_umi = Counter({int(v): int(f) for v, f in zip(vals, freq)})
    umiinfo.nexclude = int(excluded.sum())

    if output_report:
        _write_tracking(st, raw, row_group, pairs, out)

    ex_rows = np.flatnonzero(excluded)
    st.umi_dups = csr_matrix_plus((np.ones(len(ex_rows), dtype=np.ubyte), (ex_rows, np.zeros(len(ex_rows), dtype=np.int64))),
                                  shape=(N, 1))                                     # rowid(), sparse_plus.py:408-441
    lens = np.diff(raw.indptr).astype(np.int64)
    keep = np.repeat(~excluded, lens)
    ptr = np.zeros(N + 1, dtype=raw.indptr.dtype)
    np.cumsum(np.where(excluded, 0, lens), out=ptr[1:])
    st.corrected = csr_matrix_plus((raw.data[keep], raw.indices[keep], ptr), shape=raw.shape)
    return umiinfo


def _write_tracking(st, raw, row_group, pairs, out):
    """``umi_tracking.txt`` (model.py:1157-1158, 1200-1206): for every duplicated pair its reads with the
    component label (scipy numbers components by their first read), EX / REP and the feature -> score dict."""
    from . import checkpoint
    table = checkpoint.row_table(st)
    if table is not None and table.N == raw.shape[0]:
        qnames = table.names()
    else:
        qnames = [None] * raw.shape[0]
        for q, r in st.read_index.items():
            qnames[r] = q
    order = np.argsort(row_group, kind='stable')                 # reads of a pair in row order
    bounds = np.searchsorted(row_group[order], np.arange(len(pairs) + 1))
    root, excluded = out['comp_root'], out['excluded']
    with open(st.opts.outfile_path('umi_tracking.txt'), 'w') as fh:
        for g, (bc, umi) in enumerate(pairs):
            rows = order[bounds[g]:bounds[g + 1]]
            if len(rows) < 2:
                continue
            print(f'{bc}\t{umi}', file=fh)
            label = {}
            for r in rows:
                comp = label.setdefault(int(root[r]), len(label))
                a, b = raw.indptr[r], raw.indptr[r + 1]
                vec = {ft: sc for ft, sc in zip(raw.indices[a:b], raw.data[a:b])}
                vec.pop(0, None)
                print(f'\t{qnames[r]}\t{comp}\t{"EX" if excluded[r] else "REP"}\t{str(vec)}', file=fh)
