"""Ad-hoc GPU parity check used during development (gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from stellarscope_b200 import synth, layout_host
from stellarscope_b200.engine import Engine
from oracle.em_oracle import OracleLikelihood, Opts, submatrix_rows

def check(name, d, row_seg, S, seg_rows, opts_kw=None):
    opts_kw = opts_kw or {}
    t = time.time()
    hl = layout_host.build(d.indptr, d.indices, d.scores, row_seg, S, d.K)
    print(f'[{name}] layout: tiles={hl.n_tiles} chunk={hl.chunk} maxlen={hl.maxlen} segcols={len(hl.sc_gcol)} uniq={len(hl.u_row)} ({time.time()-t:.2f}s)')
    eng = Engine()
    eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
    eng.import_layout(hl)
    print('   info', eng.layout_info())
    o = Opts(**opts_kw)
    eng.fit(o.em_epsilon, o.max_iter, o.pi_prior, o.theta_prior, o.use_likelihood)
    res = eng.results()
    z = eng.emit_z().cpu().numpy()
    worst = dict(pi=0, th=0, lnl=0, diff=0, z=0)
    bad_iters = 0
    for s, rows in enumerate(seg_rows):
        ptr, idx, sc, take = submatrix_rows(d.indptr.astype(np.int64), d.indices.astype(np.int64), d.scores, rows)
        om = OracleLikelihood(ptr, idx, sc, d.K, o).em()
        it_o = len(om.fitinfo.iterations)
        if it_o != res['iters'][s]:
            bad_iters += 1
            print('   ITER MISMATCH seg', s, it_o, res['iters'][s], 'flags', res['flags'][s])
            continue
        pi, th = eng.dense_params(s)
        po = np.asarray(om.pi, dtype=np.float64); to = np.asarray(om.theta, dtype=np.float64)
        floor = 1e-290
        m = po > floor
        worst['pi'] = max(worst['pi'], np.max(np.abs(pi[m] - po[m]) / po[m]))
        worst['th'] = max(worst['th'], np.max(np.abs(th - to) / to))
        worst['lnl'] = max(worst['lnl'], abs(res['lnl'][s] - float(om.lnl)) / abs(float(om.lnl)))
        do = np.array([x[2] for x in om.fitinfo.iterations])
        dg = res['diffs'][s, :it_o]
        worst['diff'] = max(worst['diff'], np.max(np.abs(dg - do) / (1e-6 * do + 1e-12)))
        worst['z'] = max(worst['z'], np.max(np.abs(z[take] - np.asarray(om.z, dtype=np.float64))) if len(take) else 0)
    print(f'[{name}] segs={S} iter mismatches={bad_iters} worst rel: ' + ' '.join(f'{k}={v:.3g}' for k, v in worst.items()),
          'launches', eng.launch_count())
    eng.close()

d = synth.generate(n_cells=40, n_alignments=30000, n_loci=600, seed=3, unique_frac=0.1)
print('data', d.N, d.A, d.K)
check('individual', d, d.row_cell, d.n_cells, d.cell_rows())
check('pseudobulk', d, np.zeros(d.N, dtype=np.int32), 1, [np.arange(d.N)])
check('individual-lnl', d, d.row_cell, d.n_cells, d.cell_rows(), dict(use_likelihood=True))
check('pseudobulk-lnl', d, np.zeros(d.N, dtype=np.int32), 1, [np.arange(d.N)], dict(use_likelihood=True))
d2 = synth.generate(n_cells=6, n_alignments=60000, n_loci=2000, seed=4, unique_frac=0.0, n_celltypes=3)
ct = d2.cell_type[d2.row_cell]
check('celltype', d2, ct, d2.n_celltypes, d2.celltype_rows())
