"""A rank that holds no rows / no models (fewer models than ranks): the calls of the sharded path on an empty local
matrix.  python tests/tools/empty_rank_check.py   (one GPU)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.sparse as sp
from tests import helpers as H
from stellarscope_b200 import model, synth

d = H.small_synth(seed=5, n_cells=6, n_alignments=3000, n_loci=120)
m = sp.csr_matrix((d.scores, d.indices, d.indptr), shape=(d.N, d.K))
opts = H.StOpts(pooling_mode='individual')
for rows, seg, nseg in ((np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int32), 0),
                        (np.arange(10, dtype=np.int64), np.full(10, -1, dtype=np.int32), 0)):
    tl = model.TelescopeLikelihood(m, opts, row_segment=seg, n_segments=nseg, row_ids=rows)
    infos = tl.fit_segments()
    assert len(infos) == 0
    for mode in H.MODES:
        mat, ri = tl.reassign(mode, 0.9 if mode == 'best_conf' else None)
        assert mat.shape == m.shape
        if mode in ('best_exclude', 'best_conf', 'best_random', 'best_average'):
            assert mat.nnz == 0
        cm = mat.count_by_cell(d.row_cell, d.n_cells, engine=tl._engine)
        assert cm.shape == (d.K, d.n_cells)
        print(len(rows), mode, 'nnz', mat.nnz, 'counts', cm.sum(), 'info', ri.assigned, ri.ambiguous, ri.unaligned)
    z = tl.z
    assert z.shape == m.shape and z.nnz == m.nnz and not z.data.any()
    tl.close()
print('OK')
