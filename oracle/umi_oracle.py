"""TEST INFRASTRUCTURE -- CPU restatement of the reference's UMI de-duplication, the step right before
the EM path.  Follows ``Stellarscope.dedup_umi``
(stellarscope/utils/model.py:1148-1219) and ``select_umi_representatives`` (:561-668) on integer-coded
inputs.  Only tests, ``__graft_entry__.smoke()`` and the CPU-baseline legs may import this module.

Reads (rows of the score matrix) with the same barcode + UMI are duplicates of one molecule.  Inside
such a group, reads that share a feature (column != 0, ``__no_feature`` is dropped, model.py:1186) are
connected; every connected component keeps ONE representative -- the read with the largest
``(max/sum, -n_features, max, sum, position in the group)`` (model.py:578-601, sorted descending) -- and
all other reads are excluded (their rows are zeroed, model.py:1216).
"""
from __future__ import annotations

from collections import Counter

import numpy as np


def read_stats(scores):
    """model.py:578-601"""
    s = [int(x) for x in scores]
    return (max(s) / sum(s), -len(s), max(s), sum(s))


def dedup(indptr, indices, scores, row_group):
    """``row_group[i]`` = id of the (barcode, UMI) pair of row i (-1: row takes no part).
    Returns dict(excluded bool[N], comp int32[N] (scipy-style component label inside the group, -1 for
    rows of singleton groups), ncomps_umi Counter, nexclude, rpu_counter Counter)."""
    N = len(indptr) - 1
    groups = {}
    for i in range(N):
        g = int(row_group[i])
        if g >= 0:
            groups.setdefault(g, []).append(i)            # insertion order = row order (read_index order)
    excluded = np.zeros(N, dtype=bool)
    comp = np.full(N, -1, dtype=np.int32)
    ncomps_umi = Counter()
    nexclude = 0
    for g, rows in groups.items():
        if len(rows) == 1:
            continue                                       # model.py:1175-1176
        vecs = []
        for r in rows:
            a, b = indptr[r], indptr[r + 1]
            vecs.append({int(c): int(s) for c, s in zip(indices[a:b], scores[a:b]) if c != 0})
        n = len(rows)
        # connected components of "share a feature" (model.py:625-636; the shortcut :613-620 is the
        # one-component special case of the same rule)
        label = [-1] * n
        nc = 0
        for i in range(n):                                 # scipy labels components in order of their first node
            if label[i] >= 0:
                continue
            label[i] = nc
            stack = [i]
            while stack:
                u = stack.pop()
                for v in range(n):
                    if label[v] < 0 and vecs[u].keys() & vecs[v].keys():
                        label[v] = nc
                        stack.append(v)
            nc += 1
        ex = [True] * n
        for c in range(nc):
            ranks = sorted(((*read_stats(vecs[r].values()), r) for r in range(n) if label[r] == c), reverse=True)
            ex[ranks[0][-1]] = False                       # model.py:648-655
        for k, r in enumerate(rows):
            excluded[r] = ex[k]
            comp[r] = label[k]
        ncomps_umi[nc] += 1
        nexclude += sum(ex)
    rpu = Counter(len(v) for v in groups.values())
    return dict(excluded=excluded, comp=comp, ncomps_umi=ncomps_umi, nexclude=nexclude, rpu_counter=rpu)


def corrected_matrix(indptr, indices, scores, excluded):
    """``raw_scores.multiply(bool_inv(umi_dups))`` (model.py:1216): the excluded rows lose their entries."""
    lens = np.diff(indptr).astype(np.int64)
    keep = np.repeat(~excluded, lens)
    ptr = np.zeros(len(indptr), dtype=np.int64)
    np.cumsum(np.where(excluded, 0, lens), out=ptr[1:])
    return ptr, np.asarray(indices)[keep], np.asarray(scores)[keep]
