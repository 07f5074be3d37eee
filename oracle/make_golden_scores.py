"""Golden vectors of the score-matrix construction from the REAL reference (build container only).

    python -m oracle.make_golden_scores

Runs the unmodified ``Stellarscope.load_alignment`` (reference model.py:946-992; its nested
``mapping_to_matrix`` :958-978) with the BAM iteration (``_load_sequential``, pysam) replaced by seeded synthetic
mapping tuples, and stores inputs + the resulting ``raw_scores`` CSR in tests/golden/scores_*.npz.
"""
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def make_case(seed, n_reads=600, n_loci=80, max_loci_per_read=6, dup_frac=0.35, min_as=57, nofeat_frac=0.2):
    """One record per (fragment, locus) mapping as the loader emits them (model.py:1131-1132), in BAM order
    (reads consecutive, loci of a read in any order); some (read, locus) pairs occur several times with
    different scores (the max wins), some reads also map outside the annotation (column 0)."""
    rng = np.random.default_rng(seed)
    read, feat, alnscore, alnlen = [], [], [], []
    for r in range(n_reads):
        k = int(rng.integers(1, max_loci_per_read + 1))
        loci = rng.choice(np.arange(1, n_loci + 1), size=k, replace=False).tolist()
        if rng.random() < nofeat_frac:
            loci.append(0)
        recs = []
        for j in loci:
            reps = 1 if rng.random() > dup_frac else int(rng.integers(2, 5))
            for _ in range(reps):
                recs.append((j, int(rng.integers(min_as, min_as + 120)), int(rng.integers(30, 152))))
        for i in rng.permutation(len(recs)):
            j, s, ln = recs[i]
            read.append(r); feat.append(j); alnscore.append(s); alnlen.append(ln)
    return dict(read=np.array(read, dtype=np.int32), feat=np.array(feat, dtype=np.int32),
                alnscore=np.array(alnscore, dtype=np.int32), alnlen=np.array(alnlen, dtype=np.int32),
                min_as=min_as, n_reads=n_reads, n_feats=n_loci + 1)


def run_reference(ref, c):
    """The real ``load_alignment`` on the records of ``c``."""
    nofeat = '__no_feature'
    loci = {f'L{j:05d}': None for j in range(1, c['n_feats'])}
    fnames = [nofeat] + list(loci)
    st = types.SimpleNamespace()
    st.opts = refshim.RefOpts(no_feature_key=nofeat)
    st.run_info = {}
    st.read_index = {}
    maps = []
    for r, j, s, ln in zip(c['read'].tolist(), c['feat'].tolist(), c['alnscore'].tolist(), c['alnlen'].tolist()):
        q = f'read{r:07d}'
        st.read_index.setdefault(q, len(st.read_index))          # store_read_info, model.py:1044
        maps.append(('PM', q, fnames[j], s, ln))
    alninfo = types.SimpleNamespace(minAS=c['min_as'])
    st._load_sequential = lambda annot: (maps, alninfo)
    annot = types.SimpleNamespace(loci=loci, feature_length=lambda: {k: 1000 for k in loci})
    ref.Stellarscope.load_alignment(st, annot)
    m = sp.csr_matrix(st.raw_scores)
    assert m.dtype == np.uint16
    # dok -> CSR keeps the dok's insertion order inside a row with the scipy of this container (1.18; older
    # versions sorted through coo.sum_duplicates): the stored fixture is the same matrix with ascending columns
    nnz = m.nnz
    m.sum_duplicates()
    assert m.nnz == nnz and m.has_canonical_format
    return m, st, maps, annot


def main():
    ref = refshim.load()
    for name, kw in (('mixed', dict(seed=11)),
                     ('many_dups', dict(seed=12, n_reads=300, n_loci=12, dup_frac=0.8, max_loci_per_read=8)),
                     ('wide', dict(seed=13, n_reads=400, n_loci=3000, dup_frac=0.1, nofeat_frac=0.5, min_as=-20))):
        c = make_case(**kw)
        m, _, _, _ = run_reference(ref, c)
        out = dict(c)
        out.update(indptr=m.indptr, indices=m.indices, data=m.data)
        np.savez_compressed(os.path.join(GOLDEN, f'scores_{name}.npz'), **out)
        print(f'{name}: records={len(c["read"])} shape={m.shape} nnz={m.nnz} score range {m.data.min()}..{m.data.max()}')


if __name__ == '__main__':
    main()
