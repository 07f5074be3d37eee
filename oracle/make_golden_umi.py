"""Golden vectors of the UMI de-duplication from the REAL reference (build container only).

    python -m oracle.make_golden_umi

Runs the unmodified ``Stellarscope.dedup_umi`` (reference model.py:1148-1219) on seeded synthetic
reads whose duplicate groups have every overlap structure (all reads share a locus, chains, disjoint
sets, ties in the ranking statistics) and stores inputs + outputs in tests/golden/umi_*.npz.
"""
import os
import sys
import tempfile
import types

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def make_case(seed, n_cells=6, n_umi=400, K=60, dup_frac=0.45):
    """rows grouped by (cell, umi); duplicate groups of 2..7 reads built from a small locus pool so that
    reads overlap in many ways; some reads carry column 0 (__no_feature) too."""
    rng = np.random.default_rng(seed)
    rows = []           # (cell, umi, {col: score})
    for u in range(n_umi):
        cell = int(rng.integers(0, n_cells))
        n = 1 if rng.random() > dup_frac else int(rng.integers(2, 8))
        pool = rng.choice(np.arange(1, K), size=int(rng.integers(2, 9)), replace=False)
        for _ in range(n):
            k = int(rng.integers(1, min(5, len(pool)) + 1))
            cols = rng.choice(pool, size=k, replace=False)
            # few distinct score values -> ties in max / sum / ratio
            vec = {int(c): int(rng.choice([150, 160, 170, 180, 200])) for c in cols}
            if rng.random() < 0.2:
                vec[0] = int(rng.integers(140, 200))
            rows.append((cell, u, vec))
    order = rng.permutation(len(rows))                      # reads of a group are not adjacent in the BAM
    rows = [rows[i] for i in order]
    N = len(rows)
    indptr = np.zeros(N + 1, dtype=np.int32)
    idx, sc = [], []
    for i, (_, _, vec) in enumerate(rows):
        cs = sorted(vec)
        idx += cs
        sc += [vec[c] for c in cs]
        indptr[i + 1] = len(idx)
    return dict(indptr=indptr, indices=np.array(idx, dtype=np.int32), scores=np.array(sc, dtype=np.uint16), K=K,
                row_cell=np.array([r[0] for r in rows], dtype=np.int32),
                row_umi=np.array([r[1] for r in rows], dtype=np.int32))


def run_case(name, ref, **kw):
    c = make_case(**kw)
    N, K = len(c['indptr']) - 1, c['K']
    tmp = tempfile.mkdtemp()
    st = types.SimpleNamespace()
    st.opts = refshim.RefOpts()
    st.opts.outfile_path = lambda suffix: os.path.join(tmp, suffix)
    st.raw_scores = ref.csr_matrix(sp.csr_matrix((c['scores'], c['indices'], c['indptr']), shape=(N, K)))
    st.shape = (N, K)
    st.read_index = {f'read{i:06d}': i for i in range(N)}
    st.read_bcode_map = {f'read{i:06d}': f'BC{c["row_cell"][i]:03d}' for i in range(N)}
    st.read_umi_map = {f'read{i:06d}': f'UMI{c["row_umi"][i]:05d}' for i in range(N)}
    with refshim.cli_fp_policy():
        info = ref.Stellarscope.dedup_umi(st, output_report=True, summary=False)
    dups = np.asarray(st.umi_dups.todense()).ravel().astype(bool)
    cor = sp.csr_matrix(st.corrected)
    cor.eliminate_zeros()
    cor.sort_indices()
    out = dict(c)
    out.update(excluded=dups, c_indptr=cor.indptr, c_indices=cor.indices, c_scores=cor.data,
               nexclude=info.nexclude,
               ncomps=np.array(sorted(info.ncomps_umi.items())), rpu=np.array(sorted(info.rpu_counter.items())),
               report=np.array(open(os.path.join(tmp, 'umi_tracking.txt')).read()))
    np.savez_compressed(os.path.join(GOLDEN, f'umi_{name}.npz'), **out)
    print(f'{name}: N={N} groups={len(set(zip(c["row_cell"].tolist(), c["row_umi"].tolist())))} excluded={int(dups.sum())} '
          f'ncomps={dict(info.ncomps_umi)}')


def main():
    ref = refshim.load()
    run_case('mixed', ref, seed=1)
    run_case('dense_overlap', ref, seed=2, K=12, n_umi=300, dup_frac=0.7)
    run_case('sparse_overlap', ref, seed=3, K=400, n_umi=300, dup_frac=0.6)


if __name__ == '__main__':
    main()
