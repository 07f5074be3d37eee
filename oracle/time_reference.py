"""Timing of the REAL reference's fit ("the reference exactly as written on a <=200-cell
subsample to document its O(cells x N_total) behaviour").  TEST / BENCH INFRASTRUCTURE.

    python -m oracle.time_reference [n_cells ...]                       # build container: /root/reference
    python oracle/time_reference.py --npz sample.npz --ref baseline/_ref --nproc 16 --json
                                                                        # bench.py's cpu_baseline.reference_as_written leg:
                                                                        # the unmodified package under baseline/_ref (it
                                                                        # travels to the GPU box) on a sample bench.py wrote

Workload: the first n cells of BASELINE.json configs[1] (synthetic 10x, seed 0, ~500 alignments per cell),
``_fit_pooling_model`` in individual mode, serial (``nproc=1``) and with ``nproc = os.cpu_count()``; the unit is
the bench's: sum over cells of alignments x EM iterations, per second of wall time.  The numpy port
(oracle/em_oracle.py, what bench.py times on the GPU box) runs on the same cells for comparison.
"""
import os
import sys
import time
import types

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if '--ref' in sys.argv:                                    # before oracle.refshim reads the location
    os.environ['STELLARSCOPE_REFERENCE'] = os.path.abspath(sys.argv[sys.argv.index('--ref') + 1])
from oracle import em_oracle, refshim  # noqa: E402
from oracle.make_golden import make_st  # noqa: E402
from stellarscope_b200 import synth  # noqa: E402


def subsample(d, n_cells):
    """The rows of the first ``n_cells`` cells (rows of a cell are contiguous) as their own data set."""
    n_rows = int(np.searchsorted(d.row_cell, n_cells))
    nnz = int(d.indptr[n_rows])
    s = types.SimpleNamespace(indptr=d.indptr[:n_rows + 1].copy(), indices=d.indices[:nnz].copy(), scores=d.scores[:nnz].copy(),
                              N=n_rows, K=d.K, A=nnz, n_cells=n_cells, row_cell=d.row_cell[:n_rows].copy(),
                              cell_type=d.cell_type[:n_cells].copy(), n_celltypes=d.n_celltypes, row_umi=d.row_umi[:n_rows].copy())
    s.cell_rows = lambda: [np.flatnonzero(s.row_cell == c) for c in range(n_cells)]
    return s


def main_sample():
    """One timed call of the reference's ``_fit_pooling_model`` (individual mode, ``nproc`` workers) on the sample
    bench.py wrote; prints one JSON line."""
    import json
    a = sys.argv
    z = np.load(a[a.index('--npz') + 1])
    nproc = int(a[a.index('--nproc') + 1]) if '--nproc' in a else (os.cpu_count() or 1)
    n = int(z['n_cells'])
    d = types.SimpleNamespace(indptr=z['indptr'], indices=z['indices'], scores=z['scores'], N=len(z['indptr']) - 1,
                              K=int(z['K']), A=int(z['indptr'][-1]), n_cells=n, row_cell=z['row_cell'],
                              cell_type=np.zeros(n, dtype=np.int32), n_celltypes=1,
                              row_umi=np.arange(len(z['indptr']) - 1, dtype=np.int32))
    d.cell_rows = lambda: [np.flatnonzero(d.row_cell == c) for c in range(n)]
    ref = refshim.load()
    opts = refshim.RefOpts(pooling_mode='individual', nproc=nproc)
    st, _ = make_st(d, opts, False)
    a_cell = np.bincount(d.row_cell, weights=np.diff(d.indptr), minlength=n)
    with refshim.cli_fp_policy():
        t0 = time.perf_counter()
        tl, pool = ref._fit_pooling_model(st, opts, processes=nproc)
        dt = time.perf_counter() - t0
    iters = np.array([len(pool.models_info[k].iterations) for k in pool.models_info])
    units = float((a_cell * iters).sum())
    print(json.dumps(dict(value=units / dt, unit='aln*iter/s', cores=nproc, kind='_ref', seconds=dt, cells=n,
                          alignments=d.A, rows=d.N, mean_iters=float(iters.mean()),
                          what='unmodified reference package (baseline/_ref), stellarscope.utils.model.'
                               '_fit_pooling_model(st, opts) individual mode as written (one full-height masked '
                               'matrix per cell, multiprocessing.Pool(nproc))')))


def main():
    if '--npz' in sys.argv:
        return main_sample()
    sizes = [int(a) for a in sys.argv[1:]] or [50, 100, 200]
    ref = refshim.load()
    full = synth.generate(n_cells=10_000, n_alignments=5_000_000, n_loci=28_000, seed=0)
    ncpu = os.cpu_count()
    print(f'host: {ncpu} cores; reference = {refshim.REFERENCE_ROOT}', flush=True)
    for n in sizes:
        d = subsample(full, n)
        lens = np.diff(d.indptr)
        a_cell = np.bincount(d.row_cell, weights=lens, minlength=n)
        for nproc in (1, ncpu):
            opts = refshim.RefOpts(pooling_mode='individual', nproc=nproc)
            st, _ = make_st(d, opts, False)
            with refshim.cli_fp_policy():
                t0 = time.perf_counter()
                tl, pool = ref._fit_pooling_model(st, opts, processes=nproc)
                dt = time.perf_counter() - t0
            iters = np.array([len(pool.models_info[k].iterations) for k in pool.models_info])
            units = float((a_cell * iters).sum())
            print(f'reference as written: {n:4d} cells, {d.N} rows, {d.A} alignments, nproc={nproc}: {dt:7.2f} s '
                  f'({dt / n * 1e3:.0f} ms per cell), {units / dt / 1e6:.3f} M aln*iter/s, mean iterations {iters.mean():.1f}',
                  flush=True)
        # the numpy port on row-compacted per-cell matrices (what bench.py --impl reference runs), one core
        m = sp.csr_matrix((d.scores, d.indices, d.indptr), shape=(d.N, d.K))
        t0 = time.perf_counter()
        tot = 0.0
        for c in range(n):
            rows = np.flatnonzero(d.row_cell == c)
            sub = m[rows]
            fit = em_oracle.OracleLikelihood(sub.indptr, sub.indices, sub.data, d.K, em_oracle.Opts()).em()
            tot += sub.nnz * len(fit.fitinfo.iterations)
        dt = time.perf_counter() - t0
        print(f'numpy port, 1 core, row-compacted cells: {n:4d} cells: {dt:7.2f} s, {tot / dt / 1e6:.3f} M aln*iter/s', flush=True)


if __name__ == '__main__':
    main()
