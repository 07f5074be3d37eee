"""TEST INFRASTRUCTURE -- CPU restatement of the reference's score-matrix construction, the producer of the
EM path's input.  Follows ``mapping_to_matrix`` inside
``Stellarscope.load_alignment`` (stellarscope/utils/model.py:958-978) on integer-coded records.
Only tests, ``__graft_entry__.smoke()`` and the CPU-baseline legs may import this module.

Pinned against outputs of the real reference (``oracle/make_golden_scores.py`` ->
``tests/golden/scores_*.npz``).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def mapping_to_matrix(read, feat, alnscore, alnlen, min_as, n_reads, n_feats):
    """Returns (csr uint16 canonical, n_reads_without_feature).

    model.py:962     rescale = s - minAS + 1
    model.py:966-969 m[i, j] = max(m[i, j], rescale(alnscore) + alnlen)   (dok, uint16)
    model.py:971-974 every read must have a non-zero value in a column >= 1
    model.py:977     dok -> CSR (coo.tocsr + sum_duplicates: columns ascending)
    """
    read = np.asarray(read, dtype=np.int64)
    feat = np.asarray(feat, dtype=np.int64)
    val = (np.asarray(alnscore, dtype=np.int64) - int(min_as) + 1) + np.asarray(alnlen, dtype=np.int64)
    if len(val) and (val.min() <= 0 or val.max() > 65535):
        raise OverflowError('score outside (0, 65535]')
    key = read * n_feats + feat
    order = np.lexsort((val, key))                      # by key, then value: the last of a run is its maximum
    key, val = key[order], val[order]
    last = np.ones(len(key), dtype=bool)
    last[:-1] = key[1:] != key[:-1]
    key, val = key[last], val[last]
    rows, cols = key // max(n_feats, 1), key % max(n_feats, 1)
    m = sp.csr_matrix((val.astype(np.uint16), (rows, cols)), shape=(n_reads, n_feats))
    m.sort_indices()
    has_feat = np.zeros(n_reads, dtype=bool)
    has_feat[rows[cols >= 1]] = True
    return m, int((~has_feat).sum())


def mapping_to_matrix_loop(read, feat, alnscore, alnlen, min_as, n_reads, n_feats):
    """The reference's loop, literally (small cases)."""
    d = {}
    for i, j, s, ln in zip(read, feat, alnscore, alnlen):
        k = (int(i), int(j))
        d[k] = max(d.get(k, 0), (int(s) - int(min_as) + 1) + int(ln))
    rows = np.array([k[0] for k in d], dtype=np.int64)
    cols = np.array([k[1] for k in d], dtype=np.int64)
    vals = np.array(list(d.values()), dtype=np.int64)
    m = sp.csr_matrix((vals.astype(np.uint16), (rows, cols)), shape=(n_reads, n_feats))
    m.sort_indices()
    return m
