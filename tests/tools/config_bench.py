"""Device-time measurement + validation of the BASELINE.json configurations on ONE GPU.

    python tests/tools/config_bench.py cfg2|cfg3|cfg4|cfg5 [scale] [n_oracle_cells]

cfg2: 10k cells, 5M alignments, individual          cfg3: 50k cells, 100M alignments, celltype(20), UMI dedup
cfg4: one model, 500M alignments (pseudobulk)       cfg5: 100k cells, 500M alignments, individual, reassign sweep
``scale`` multiplies cells and alignments (cfg4/5 at 1.0 need ~35 GB of host memory to generate).

Checks (size independent):
  * every model converged or reached max_iter, iteration counts in [1, max_iter]
  * sum_j pi_j = 1 for every model (M-step invariant, model.py:243-246), rows of z sum to 1
  * total_hits counts = number of alignments; best_exclude <= number of fitted rows
  * a random sample of models re-fitted by the CPU oracle: same iteration count, pi within 1e-6, lnL 1e-10
and reports alignments x iterations / s, the algorithmic bandwidth and the layout build time.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from oracle.em_oracle import OracleLikelihood, Opts, submatrix_rows
from stellarscope_b200 import synth
from stellarscope_b200.engine import Engine

CFG = {
    'cfg2': dict(n_cells=10_000, n_alignments=5_000_000, mode='individual'),
    'cfg3': dict(n_cells=50_000, n_alignments=100_000_000, mode='celltype', n_celltypes=20, umi_dup_rate=0.15),
    'cfg4': dict(n_cells=100_000, n_alignments=500_000_000, mode='pseudobulk'),
    'cfg5': dict(n_cells=100_000, n_alignments=500_000_000, mode='individual'),
}


def main():
    name = sys.argv[1]
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    n_oracle = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    c = dict(CFG[name])
    mode = c.pop('mode')
    c['n_cells'] = max(1, int(c['n_cells'] * scale))
    c['n_alignments'] = int(c['n_alignments'] * scale)
    t0 = time.perf_counter()
    d = (synth.generate_device if torch.cuda.is_available() else synth.generate)(n_loci=28_000, seed=0, **c)
    if c.get('umi_dup_rate'):
        d = synth.apply_row_mask(d, synth.umi_dedup_mask(d))          # rows zeroed like model.py:1216
    print(f'{name} x{scale}: {d.n_cells} cells, {d.N} rows, {d.A} alignments, mode {mode} '
          f'(generated in {time.perf_counter() - t0:.1f} s)', flush=True)
    if mode == 'individual':
        row_seg, S = d.row_cell, d.n_cells
        seg_rows = None
    elif mode == 'celltype':
        row_seg, S = d.cell_type[d.row_cell].astype(np.int32), d.n_celltypes
    else:
        row_seg, S = np.zeros(d.N, dtype=np.int32), 1
    eng = Engine(0)
    eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.build_layout(row_seg, S)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    info = eng.layout_info()
    print('layout:', {k: info[k] for k in ('n_tiles', 'n_resident_tiles', 'n_cluster_tiles', 'n_stream_tiles',
                                           'n_cluster_segments', 'n_stream_segments', 'n_segcols')},
          f'build {1e3 * build_s:.1f} ms', flush=True)
    eng.enable_timing(True)
    best = None
    for i in range(3):
        eng.fit()
        torch.cuda.synchronize()
        t = eng.last_timings()
        if best is None or t['em'] < best['em']:
            best = t
    res = eng.results(traces=False)
    p = eng.params()
    seg_A = eng.export_layout_array('seg_A', np.int64)
    seg_Aamb = eng.export_layout_array('seg_Aamb', np.int64)
    seg_namb = eng.export_layout_array('seg_namb', np.int32).astype(np.int64)
    it = res['iters'].astype(np.int64)
    fitted = seg_A > 0
    units = float((seg_A * it).sum())
    alg = float(((12 * seg_Aamb + 4 * seg_namb + 16 * p['seg_ncol'].astype(np.int64)) * it).sum())
    em_s = best['em'] / 1e3
    print(f'fit: em {best["em"]:.3f} ms (lane {best["resident"]:.3f} cluster {best["cluster"]:.3f} [heavy {best["cluster_heavy"]:.3f} '
          f'light {best["cluster_light"]:.3f}] stream {best["stream"]:.3f}) '
          f'mean iters {it[fitted].mean():.1f} max {it.max()}  rate {units / em_s / 1e9:.1f} G aln*iter/s  '
          f'algorithmic {alg / em_s / 1e9:.0f} GB/s', flush=True)
    # ---- properties ----
    assert np.all((it[fitted] >= 1) & (it[fitted] <= 100))
    assert np.all((res['flags'][fitted] & 3) != 0) and not np.any(res['flags'] & 4 & np.where(fitted, 4, 0))
    # sum of pi over the K loci of each model = 1
    col0, ncol = p['seg_col0'].astype(np.int64), p['seg_ncol'].astype(np.int64)
    seg_sum = np.zeros(len(col0))
    nz = ncol > 0
    seg_sum[nz] = np.add.reduceat(p['pi'], col0[nz])          # per-model sums (a global prefix sum would lose 1e-9)
    pisum = seg_sum + (d.K - ncol) * p['pi_rest']
    assert np.max(np.abs(pisum[fitted] - 1.0)) < 1e-9, np.max(np.abs(pisum[fitted] - 1.0))
    z = eng.emit_z()
    rowsum = torch.zeros(d.N, dtype=torch.float64, device=z.device)
    rows_t = torch.repeat_interleave(torch.arange(d.N, device=z.device),
                                     torch.from_numpy(np.diff(d.indptr).astype(np.int64)).to(z.device))
    rowsum.index_add_(0, rows_t, z)
    lens = np.diff(d.indptr)
    fit_rows = torch.from_numpy((lens > 0) & (row_seg >= 0)).to(z.device)
    assert float((rowsum[fit_rows] - 1.0).abs().max()) < 1e-9
    del rows_t, rowsum
    # ---- reassign sweep + counts ----
    row_cell_t = torch.from_numpy(d.row_cell).to(z.device)
    for m in ('best_exclude', 'best_conf', 'best_random', 'best_average', 'total_hits'):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        flag, nbest, rinfo = eng.reassign(m, z, conf_thresh=0.9, tie_tol=1e-9)
        cell, col, val = eng.count(flag, row_cell_t)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if m == 'total_hits':
            assert int(val.sum()) == d.A and int(rinfo[0]) == d.A
        if m == 'best_exclude':
            assert int(val.sum()) == int(rinfo[0]) <= int((lens > 0).sum())
        print(f'reassign {m:13s} + count: {1e3 * dt:8.1f} ms  assigned {int(rinfo[0])} ambiguous {int(rinfo[1])} '
              f'count entries {len(val)}', flush=True)
    # ---- sampled oracle parity ----
    rng = np.random.default_rng(1)
    o = Opts()
    ptr64, idx64 = d.indptr.astype(np.int64), d.indices.astype(np.int64)
    if mode == 'individual':
        bounds = np.searchsorted(d.row_cell, np.arange(d.n_cells + 1))
        big = np.argsort(-seg_A)[:3]                                   # the largest cells (multi-tile paths)
        pick = np.unique(np.concatenate([rng.choice(d.n_cells, size=min(n_oracle, d.n_cells), replace=False), big]))
        worst = 0.0
        t0 = time.perf_counter()
        for s in pick:
            rows = np.arange(bounds[s], bounds[s + 1])
            ptr, idx, sc, take = submatrix_rows(ptr64, idx64, d.scores, rows)
            om = OracleLikelihood(ptr, idx, sc, d.K, o).em()
            assert res['iters'][s] == len(om.fitinfo.iterations), (s, res['iters'][s], len(om.fitinfo.iterations))
            lo = float(om.lnl)
            assert abs(res['lnl'][s] - lo) <= 1e-10 * abs(lo), (s, res['lnl'][s], lo)
            pi, th = eng.dense_params(int(s))
            po = np.asarray(om.pi, dtype=np.float64)
            ok = po > 1e-290
            worst = max(worst, float(np.max(np.abs(pi[ok] - po[ok]) / po[ok])))
        assert worst < 1e-6
        print(f'oracle parity on {len(pick)} sampled cells (incl. the 3 largest, {int(seg_A[big].max())} alignments): '
              f'iterations equal, lnL 1e-10, max rel pi err {worst:.2e}  ({time.perf_counter() - t0:.1f} s CPU)', flush=True)
    print('OK', flush=True)
    eng.close()


if __name__ == '__main__':
    main()
