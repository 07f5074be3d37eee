"""Score-matrix construction -- drop-in for ``mapping_to_matrix`` inside ``Stellarscope.load_alignment``
(reference stellarscope/utils/model.py:946-992).

The reference inserts every (fragment, locus) mapping into a ``dok_matrix`` from a Python loop,
``m[i, j] = max(m[i, j], (alnscore - minAS + 1) + alnlen)``, converts it to CSR and asserts that every read has a
stored value in a feature column.  Here the host only does the two dict look-ups per mapping (strings -> row /
column ids, a C walk of the tuple list); the max-reduction, the CSR construction and the check run in
``ss_scores_build`` (csrc/ss_scores.cu).  The BAM iteration itself (``_load_sequential``, pysam) stays the
reference's.  Nothing is computed on the CPU; without a CUDA device the call fails.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine
from .sparse_plus import csr_matrix_plus


def code_mappings(mappings, read_index, feat_index):
    """(code, query_name, feat_name, alnscore, alnlen) tuples -> int32 arrays read, feat, alnscore, alnlen."""
    from . import hostutil
    n = len(mappings)
    out = np.empty((4, n), dtype=np.int32)
    if type(mappings) is list and type(read_index) is dict and type(feat_index) is dict:
        hostutil.code_mappings(mappings, read_index, feat_index, out[0], out[1], out[2], out[3])
    else:
        for i, (_code, qname, fname, alnscore, alnlen) in enumerate(mappings):
            out[0, i], out[1, i], out[2, i], out[3, i] = read_index[qname], feat_index[fname], alnscore, alnlen
    return out[0], out[1], out[2], out[3]


def mapping_to_matrix(st, mappings, min_as, engine=None):
    """Sets ``st.shape`` and ``st.raw_scores`` from the loader's mapping tuples (model.py:958-978)."""
    st.shape = (len(st.read_index), len(st.feat_index))
    assert st.feat_index[st.opts.no_feature_key] == 0                       # model.py:972
    read, feat, alnscore, alnlen = code_mappings(mappings, st.read_index, st.feat_index)
    eng = engine or Engine.acquire()
    try:
        indptr, indices, data, info = eng.scores_build(read, feat, alnscore, alnlen, min_as, st.shape[0], st.shape[1],
                                                       keep=engine is not None)
    finally:
        if engine is None:
            eng.release()
    # every read must have a non-zero value in a feature column (model.py:971-974)
    assert int(info[3]) == 0
    st.raw_scores = csr_matrix_plus((data, indices, indptr), shape=st.shape)
    return


def load_alignment(st, annotation):
    """``Stellarscope.load_alignment`` (model.py:946-992) with the matrix built on the device."""
    st.run_info['annotated_features'] = len(annotation.loci)
    st.feature_length = annotation.feature_length().copy()
    st.feat_index = {st.opts.no_feature_key: 0, }
    for locus in annotation.loci.keys():
        _ = st.feat_index.setdefault(locus, len(st.feat_index))
    maps, alninfo = st._load_sequential(annotation)                         # the reference's BAM iteration (pysam)
    mapping_to_matrix(st, maps, alninfo.minAS)
    return alninfo
