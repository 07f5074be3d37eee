"""Golden MTX files for tests/test_mtx.py: written by scipy.io.mmwrite -- the writer the reference calls
(stellarscope/utils/model.py:1418) -- from small seeded count matrices, with a metadata comment of the shape
mtx_meta produces (:1347-1368).  Run from the repo root:  python tests/tools/make_golden_mtx.py"""
import os

import numpy as np
import scipy.io
import scipy.sparse as sp

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'golden')
rng = np.random.default_rng(20261020)
K, Cn, n = 60, 23, 320
comment = '\n '.join(['PN: stellarscope', 'VN: 1.5', 'pooling_mode: individual', 'reassign_mode: best_exclude',
                      'command: stellarscope assign aln.bam genes.gtf'])
r, c = rng.integers(0, K, n), rng.integers(0, Cn, n)
mi = sp.coo_matrix((rng.integers(1, 3000, n).astype(np.int32), (r, c)), shape=(K, Cn)).tocsc()
mi.sum_duplicates(); mi.sort_indices(); mi = mi.astype(np.int32)
mr = sp.coo_matrix((rng.integers(1, 9, n) / rng.integers(1, 13, n), (r, c)), shape=(K, Cn)).tocsc()
mr.sum_duplicates(); mr.sort_indices()
scipy.io.mmwrite(os.path.join(OUT, 'mtx_int.mtx'), mi, comment=comment)
scipy.io.mmwrite(os.path.join(OUT, 'mtx_real.mtx'), mr, comment=comment)
np.savez_compressed(os.path.join(OUT, 'mtx_cases.npz'), shape=np.array([K, Cn]), comment=np.array(comment),
                    int_data=mi.data, int_indices=mi.indices, int_indptr=mi.indptr,
                    real_data=mr.data, real_indices=mr.indices, real_indptr=mr.indptr)
print('written', mi.nnz, mr.nnz)
