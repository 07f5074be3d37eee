"""Generate tests/golden/*.npz by running the REAL reference (build container only).

    python -m oracle.make_golden

Runs ``_fit_pooling_model``, ``Stellarscope.reassign`` and
``Stellarscope.output_report`` of the unmodified reference checkout
(/root/reference, imported through oracle/refshim.py) on small seeded
synthetic inputs and stores inputs + outputs.  The GPU tests compare the CUDA
path with these files; the CPU tests compare the numpy restatement
(oracle/em_oracle.py) with them.  The reference cannot travel to the GPU box,
the fixtures can.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import scipy.io
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from stellarscope_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
MODES = ['best_exclude', 'best_conf', 'best_random', 'best_average', 'initial_unique',
         'initial_random', 'total_hits']

CASES = {
    # name: (synth kwargs, opts kwargs, use umi-corrected matrix)
    'pseudobulk': (dict(n_cells=12, n_alignments=2500, n_loci=150, seed=11, unique_frac=0.15),
                   dict(pooling_mode='pseudobulk'), False),
    'individual': (dict(n_cells=12, n_alignments=2500, n_loci=150, seed=12, unique_frac=0.15),
                   dict(pooling_mode='individual'), False),
    'celltype': (dict(n_cells=15, n_alignments=3000, n_loci=150, seed=13, unique_frac=0.1, n_celltypes=3),
                 dict(pooling_mode='celltype'), False),
    'pseudobulk_lnl': (dict(n_cells=8, n_alignments=2000, n_loci=120, seed=14, unique_frac=0.1),
                       dict(pooling_mode='pseudobulk', use_likelihood=True), False),
    'individual_lnl': (dict(n_cells=8, n_alignments=2000, n_loci=120, seed=15, unique_frac=0.1),
                       dict(pooling_mode='individual', use_likelihood=True), False),
    'individual_umicounts': (dict(n_cells=10, n_alignments=2000, n_loci=120, seed=16, unique_frac=0.1,
                                  umi_dup_rate=0.2),
                             dict(pooling_mode='individual', umi_counts=True), False),
    'celltype_umidedup': (dict(n_cells=12, n_alignments=2500, n_loci=150, seed=17, unique_frac=0.1,
                               n_celltypes=3, umi_dup_rate=0.15),
                          dict(pooling_mode='celltype', ignore_umi=False), True),
    'pseudobulk_priors': (dict(n_cells=6, n_alignments=1500, n_loci=100, seed=18, unique_frac=0.2),
                          dict(pooling_mode='pseudobulk', pi_prior=3, theta_prior=500, max_iter=30), False),
    'individual_maxiter': (dict(n_cells=6, n_alignments=1500, n_loci=100, seed=19, unique_frac=0.0),
                           dict(pooling_mode='individual', max_iter=7), False),
    # second batch: aggregation by UMI under celltype pooling, priors and tiny / mostly-unique cells under individual
    # pooling (cells without ambiguous rows), a single iteration, another confidence threshold
    'celltype_umicounts': (dict(n_cells=14, n_alignments=2600, n_loci=140, seed=21, unique_frac=0.1, n_celltypes=3,
                                umi_dup_rate=0.2),
                           dict(pooling_mode='celltype', umi_counts=True), False),
    'individual_priors': (dict(n_cells=9, n_alignments=1800, n_loci=110, seed=22, unique_frac=0.15),
                          dict(pooling_mode='individual', pi_prior=2, theta_prior=1000), False),
    'individual_sparsecells': (dict(n_cells=24, n_alignments=700, n_loci=90, seed=23, unique_frac=0.85),
                               dict(pooling_mode='individual'), False),
    'pseudobulk_maxiter1': (dict(n_cells=5, n_alignments=1200, n_loci=100, seed=24, unique_frac=0.1),
                            dict(pooling_mode='pseudobulk', max_iter=1), False),
    'individual_conf06': (dict(n_cells=8, n_alignments=1600, n_loci=100, seed=25, unique_frac=0.1),
                          dict(pooling_mode='individual', conf_prob=0.6), False),
    # BASELINE.json configs[0]: bundled BAM + GTF, pseudobulk, defaults (golden: data/telescope_report.tsv)
    'config1': ('config1', dict(pooling_mode='pseudobulk'), False),
}


def make_st(d, opts, corrected):
    """A namespace that carries what the hot path reads from ``Stellarscope``
    (SURVEY.md 8b/8c)."""
    N = d.N
    M = sp.csr_matrix((d.scores, d.indices, d.indptr), shape=(N, d.K))
    st = types.SimpleNamespace()
    st.opts = opts
    st.raw_scores = M
    st.shape = M.shape
    st.corrected = None
    if corrected:
        keep = synth.umi_dedup_mask(d)
        dm = synth.apply_row_mask(d, keep)
        st.corrected = sp.csr_matrix((dm.scores, dm.indices, dm.indptr), shape=(N, d.K))
    names = [f'BC{c:04d}' for c in range(d.n_cells)]
    st.bcode_ridx_map = {}
    for c, rows in enumerate(d.cell_rows()):
        st.bcode_ridx_map[names[c]] = set(int(r) for r in rows)
    st.celltypes = [f'type{t}' for t in range(d.n_celltypes)]
    st.ctype_bcode_map = {f'type{t}': [names[c] for c in np.flatnonzero(d.cell_type == t)]
                          for t in range(d.n_celltypes)}
    st.read_index = {f'read{i:07d}': i for i in range(N)}
    st.read_bcode_map = {f'read{i:07d}': names[d.row_cell[i]] for i in range(N)}
    st.read_umi_map = {f'read{i:07d}': f'U{d.row_umi[i]:08d}' for i in range(N)}
    st.feat_index = {('__no_feature' if j == 0 else f'locus{j:05d}'): j for j in range(d.K)}
    st.filtlist = None
    st.reassignments = {}
    return st, names


def config1_dataset():
    """BASELINE.json configs[0]: the BAM + GTF the reference bundles, through the loader restatement
    (oracle/loader_oracle.py) with ``--barcode_tag RG --ignore_umi`` (the bundled BAM has no CB/UB tags,
    SURVEY.md 0.5).  Returns the data set plus the golden columns of ``data/telescope_report.tsv``."""
    from oracle.loader_oracle import load_alignment
    data = os.path.join(refshim.REFERENCE_ROOT, 'stellarscope', 'data')
    r = load_alignment(os.path.join(data, 'alignment.bam'), os.path.join(data, 'annotation.gtf'),
                       barcode_tag='RG', ignore_umi=True)
    m = r['matrix']
    N, K = m.shape
    d = types.SimpleNamespace(indptr=m.indptr.astype(np.int32), indices=m.indices.astype(np.int32),
                              scores=m.data.astype(np.uint16), N=N, K=K, A=int(m.nnz), n_cells=1,
                              row_cell=np.zeros(N, dtype=np.int32), cell_type=np.zeros(1, dtype=np.int32),
                              n_celltypes=1, row_umi=np.arange(N, dtype=np.int32))
    d.cell_rows = lambda: [np.arange(N)]
    feats = sorted(r['feat_index'], key=r['feat_index'].get)
    rep = {}
    lines = open(os.path.join(data, 'telescope_report.tsv')).read().splitlines()
    cols = lines[1].split('\t')
    for ln in lines[2:]:
        f = ln.split('\t')
        rep[f[0]] = dict(zip(cols[1:], f[1:]))
    extra = dict(feat_names=np.array(feats), minAS=r['minAS'], maxAS=r['maxAS'],
                 report_names=np.array(list(rep.keys())))
    for c in ('final_count', 'final_conf', 'final_prop', 'init_aligned', 'unique_count'):
        extra['report_' + c] = np.array([float(rep[k][c]) for k in rep])
    return d, extra


def run_case(name, skw, okw, corrected, ref):
    extra = {}
    if skw == 'config1':
        d, extra = config1_dataset()
    else:
        d = synth.generate(**skw)
    tmpdir = tempfile.mkdtemp()
    opts = refshim.RefOpts(reassign_mode=MODES, rng=np.random.default_rng(2024), **okw)
    opts.outfile_path = lambda suffix: os.path.join(tmpdir, suffix)
    st, names = make_st(d, opts, corrected)
    out = dict(indptr=d.indptr, indices=d.indices, scores=d.scores, K=d.K, row_cell=d.row_cell,
               n_cells=d.n_cells, cell_type=d.cell_type, n_celltypes=d.n_celltypes, row_umi=d.row_umi)
    for k in ('pooling_mode', 'use_likelihood', 'em_epsilon', 'max_iter', 'pi_prior', 'theta_prior',
              'conf_prob', 'ignore_umi', 'umi_counts'):
        out['opt_' + k] = getattr(opts, k)
    out['corrected'] = corrected
    if corrected:
        out['c_indptr'], out['c_indices'], out['c_scores'] = (st.corrected.indptr, st.corrected.indices,
                                                              st.corrected.data)
    with refshim.cli_fp_policy():
        # capture per-model parameters: the reference discards them (model.py:671-673)
        models = []
        orig_em = ref.TelescopeLikelihood.em

        def em_spy(self, *a, **k):
            r = orig_em(self, *a, **k)
            models.append(self)
            return r
        ref.TelescopeLikelihood.em = em_spy
        try:
            tl, poolinfo = ref._fit_pooling_model(st, opts)
        finally:
            ref.TelescopeLikelihood.em = orig_em
        rinfo = ref.Stellarscope.reassign(st, tl)
        ref.Stellarscope.output_report(st, tl)
    keys = list(poolinfo.models_info.keys())
    fis = [poolinfo.models_info[k] for k in keys]
    MI = opts.max_iter
    out['n_models'] = len(keys)
    out['iters'] = np.array([len(f.iterations) for f in fis])
    out['converged'] = np.array([f.converged for f in fis])
    out['reached_max'] = np.array([f.reached_max for f in fis])
    out['lnl'] = np.array([float(f.final_lnl) for f in fis])
    out['nobs'] = np.array([f.nobs for f in fis])
    out['nfeats'] = np.array([f.nfeats for f in fis])
    diffs = np.zeros((len(fis), MI))
    for i, f in enumerate(fis):
        diffs[i, :len(f.iterations)] = [float(x[2]) for x in f.iterations]
    out['diffs'] = diffs
    out['pi'] = np.array([np.asarray(m.pi.toarray(), dtype=np.float64).ravel() for m in models])
    out['theta'] = np.array([np.asarray(m.theta, dtype=np.float64).ravel() for m in models])
    out['promoted'] = np.array([m.pi.dtype != np.float64 for m in models])
    z = sp.csr_matrix(tl.z, dtype=np.float64)
    z.sort_indices()
    full = sp.csr_matrix(st.corrected if corrected else st.raw_scores)
    full.eliminate_zeros()
    full.sort_indices()
    zd = np.zeros(full.nnz)
    # z has the pattern of the fitted rows; place into the pattern of the full matrix
    K = d.K
    frow = np.repeat(np.arange(full.shape[0]), np.diff(full.indptr))
    zrow = np.repeat(np.arange(z.shape[0]), np.diff(z.indptr))
    pos = np.searchsorted(frow.astype(np.int64) * K + full.indices, zrow.astype(np.int64) * K + z.indices)
    zd[pos] = z.data
    out['z'] = zd
    out['total_lnl'] = float(poolinfo.total_lnl)
    out['total_obs'] = poolinfo.total_obs
    out['total_params'] = poolinfo.total_params
    out['AIC'] = float(poolinfo.AIC)
    out['BIC'] = float(poolinfo.BIC)
    for mi, mode in enumerate(MODES):
        r = sp.csr_matrix(st.reassignments[mode])
        r.eliminate_zeros()
        r.sort_indices()
        out[f'r_{mode}_indptr'] = r.indptr
        out[f'r_{mode}_indices'] = r.indices
        out[f'r_{mode}_data'] = r.data
        ri = rinfo[mode]
        out[f'r_{mode}_info'] = np.array([ri.assigned, ri.ambiguous,
                                          -1 if ri.unaligned is None else ri.unaligned])
        dist = getattr(ri, 'ambigous_dist', None)
        if dist is not None:
            ks = sorted(dist)
            out[f'r_{mode}_dist'] = np.array([[k, dist[k]] for k in ks])
        fn = 'TE_counts.mtx' if mi == 0 else f'TE_counts.{mode}.mtx'
        out[f'c_{mode}'] = np.asarray(scipy.io.mmread(os.path.join(tmpdir, fn)).todense())
    out['barcodes'] = np.array(open(os.path.join(tmpdir, 'barcodes.tsv')).read().split())
    out.update(extra)
    np.savez_compressed(os.path.join(GOLDEN, f'ref_{name}.npz'), **out)
    print(f'{name}: N={d.N} A={d.A} K={d.K} models={len(keys)} iters={out["iters"].tolist()[:6]} '
          f'lnl={out["total_lnl"]:.6f} promoted={bool(out["promoted"].any())}')


def main():
    ref = refshim.load()
    os.makedirs(GOLDEN, exist_ok=True)
    only = sys.argv[1:]
    for name, (skw, okw, corrected) in CASES.items():
        if only and name not in only:
            continue
        run_case(name, skw, okw, corrected, ref)


if __name__ == '__main__':
    main()
