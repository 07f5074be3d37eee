#!/usr/bin/env python
"""Benchmark of the EM read-reassignment hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg5|cfg2|cfg3|cfg4]
                    [--scale S]

One "step" = one EM fit (``ss_em_fit``: every model of the pooling mode to convergence) over ONE synthetic data
set.  Default workload = BASELINE.json ``configs[4]`` ("atlas": 100k cells, ~500M alignments, 28k loci,
pooling_mode=individual): the configuration the north-star target is quoted on; it fits one B200.  Under
``torchrun`` with N ranks the SAME data set is sharded (STRONG scaling): individual / celltype pooling packs
whole models onto ranks by longest-processing-time on their alignment counts (no data-path collective,
SURVEY.md 8e); pseudobulk (``--config cfg4``) cuts the rows of the one model into N blocks and exchanges the
partial theta sums once per iteration inside the kernel (NVLink peer memory).

``value``  alignments x EM-iterations / s of the whole job, inputs resident in HBM, device time of the fit (CUDA
           events on the launching stream), max over ranks.
``e2e``    the same metric through the reference-facing Python API
           (``stellarscope_b200.model._fit_pooling_model(st, opts)``) starting from HOST scipy matrices on every
           rank: model packing, cutting the rank's rows out of the host matrix, H2D, device layout build, fit, D2H
           and the gather of the per-model records over the ranks -- wall clock between barriers, max over ranks.
``parity_sample``  after the timed region: sampled models re-fitted by the CPU oracle (iteration counts equal,
           pi 1e-6, lnL 1e-10) and their best_exclude counts compared bit-exactly; a mismatch exits non-zero.
``--impl reference``  the reference's CPU implementation of the path on the host cores (``baseline/_ref`` --
           the unmodified reference package -- when it is present, else the numpy oracle port), on a bounded
           sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'alignments_x_em_iterations_per_sec'
UNIT = 'aln*iter/s'
N_LOCI = 28_000
CONFIGS = {
    'cfg2': dict(n_cells=10_000, n_alignments=5_000_000, mode='individual',
                 label='configs[1]: synthetic 10x, 10k cells, 5M multimapped alignments, 28k loci, pooling_mode=individual'),
    'cfg3': dict(n_cells=50_000, n_alignments=100_000_000, mode='celltype', n_celltypes=20, umi_dup_rate=0.15,
                 label='configs[2]: synthetic 10x, 50k cells, 100M alignments, pooling_mode=celltype (20 celltypes), '
                       'UMI de-duplication applied as a row mask'),
    'cfg4': dict(n_cells=100_000, n_alignments=500_000_000, mode='pseudobulk',
                 label='configs[3]: synthetic pseudobulk, one model over 500M alignments, rows sharded over the ranks'),
    'cfg5': dict(n_cells=100_000, n_alignments=500_000_000, mode='individual',
                 label='configs[4]: synthetic atlas, 100k cells, ~500M alignments, 28k loci, pooling_mode=individual'),
}
OPTS = dict(em_epsilon=1e-7, max_iter=100, pi_prior=0, theta_prior=200000)   # stellarscope_assign.yaml:231-250
TIE_TOL = 1e-9
# dram__bytes_read.sum + dram__bytes_write.sum of the EM kernels of ONE fit, from ncu --set full captures
# (profiles/): config -> (bytes per fit at scale 1.0 or None, source)
NCU_TRAFFIC = {
    ('cfg2', 1.0): (111.0e6, 'profiles/r1e_em_full_ncu.csv (one fit, sum over the 16 EM kernels)'),
    ('cfg5', 0.1): (8.762e9, 'profiles/r2i_cfg5_x0.1_em_ncu.csv (one fit at scale 0.1, sum over the 21 EM kernels: '
                             '7.5 GB of it is em_stream_kernel re-reading the tiles of the 92 largest cells every '
                             'iteration; lane + cluster kernels read their tiles once: 1.2 GB)'),
    ('cfg5', 0.05): (0.621e9, 'profiles/r2l_cfg5_x0.05_em_ncu.csv (one fit at scale 0.05, sum over the 28 EM kernels; '
                              'the tiles of the 46 streamed cells stay in L2 at this scale, every tile is read from '
                              'DRAM about once)'),
}


def measured_hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index, interval=0.25):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.interval = interval
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                      '-i', str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(',')])
            except Exception:
                pass
            self._halt.wait(self.interval)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                   r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                continue
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------
def make_workload(cfg, scale, seed=0, keep_device=False):
    """The synthetic data set of a configuration (stellarscope_b200.synth; the large draws on the GPU when
    there is one: numpy needs minutes for 5x10^8 alignments)."""
    from stellarscope_b200 import synth
    c = {k: v for k, v in CONFIGS[cfg].items() if k not in ('mode', 'label')}
    c['n_cells'] = max(1, int(c['n_cells'] * scale))
    c['n_alignments'] = max(100, int(c['n_alignments'] * scale))
    try:
        import torch
        on_gpu = torch.cuda.is_available()
    except Exception:
        on_gpu = False
    if on_gpu:
        d = synth.generate_device(n_loci=N_LOCI, seed=seed, keep_device=keep_device and not c.get('umi_dup_rate'), **c)
    else:
        d = synth.generate(n_loci=N_LOCI, seed=seed, **c)
    if c.get('umi_dup_rate'):
        d = synth.apply_row_mask(d, synth.umi_dedup_mask(d))          # rows zeroed like model.py:1216
    return d


def row_models(d, mode):
    """Row -> model map of the pooling mode and the number of models."""
    if mode == 'individual':
        return d.row_cell, d.n_cells
    if mode == 'celltype':
        return d.cell_type[d.row_cell].astype(np.int32), d.n_celltypes
    return np.zeros(d.N, dtype=np.int32), 1


class BenchOpts:
    def __init__(self, mode):
        self.__dict__.update(OPTS)
        self.use_likelihood = False
        self.pooling_mode = mode
        self.ignore_umi = True
        self.umi_counts = False
        self.nproc = 1
        self.reassign_mode = ['best_exclude']
        self.conf_prob = 0.9
        self.rng = np.random.default_rng(0)
        self.shard_models = True       # strong scaling: the models of the ONE data set are packed onto the ranks
        self.version = 'bench'


def make_st(d, opts):
    """The attributes of ``Stellarscope`` the hot path reads.  The per-read containers are handed
    over the way ``stellarscope resume`` gets them from a side-car checkpoint (``checkpoint.RowTable``: a barcode
    code per row): at 1.4x10^8 reads the reference's dict of Python sets would not fit a benchmark step."""
    import types
    import scipy.sparse as sp
    from stellarscope_b200 import checkpoint
    st = types.SimpleNamespace()
    st.opts = opts
    st.raw_scores = sp.csr_matrix((d.scores, d.indices, d.indptr), shape=(d.N, d.K), copy=False)
    st.shape = st.raw_scores.shape
    st.corrected = None
    bcodes = [f'BC{c:06d}' for c in range(d.n_cells)]
    table = checkpoint.RowTable(d.N, b'', bcodes, d.row_cell)
    checkpoint.attach_row_table(st, table)
    st.celltypes = [f'type{t}' for t in range(d.n_celltypes)] if opts.pooling_mode == 'celltype' else []
    st.ctype_bcode_map = ({f'type{t}': [bcodes[c] for c in np.flatnonzero(d.cell_type == t)]
                           for t in range(d.n_celltypes)} if opts.pooling_mode == 'celltype' else {})
    st.feat_index = {('__no_feature' if j == 0 else f'locus{j:05d}'): j for j in range(d.K)}
    st.filtlist = None
    st.reassignments = {}
    return st


# ---------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------
def _cpu_fit_cells(cells):
    """cells: list of (ptr, idx, scores, K) row-compacted per-model matrices."""
    from oracle.em_oracle import OracleLikelihood, Opts
    o = Opts(**OPTS)
    units = 0
    for ptr, idx, sc, K in cells:
        m = OracleLikelihood(ptr, idx, sc, K, o).em()
        units += int(ptr[-1]) * len(m.fitinfo.iterations)
    return units


def model_rows(d, row_seg, s):
    """Row ids of model ``s`` (contiguous for the synthetic individual mode, scattered otherwise)."""
    if row_seg is d.row_cell:
        a, b = np.searchsorted(d.row_cell, [s, s + 1])
        return np.arange(a, b)
    return np.flatnonzero(row_seg == s)


def cpu_sample_rate(d, mode, cores, budget_aln, seed=0):
    """alignments x iterations / s of the oracle port (numpy restatement of the reference EM) on ``cores``
    processes over a random sample of whole models of the workload -- about ``budget_aln`` alignments in total
    (row-compacted per-model submatrices, bit-identical to the reference's masking).  The
    matrices are cut out before the clock starts and the worker pool is up before the clock starts."""
    import multiprocessing as mp
    from oracle.em_oracle import submatrix_rows
    rng = np.random.default_rng(seed)
    row_seg, S = row_models(d, mode)
    indptr = d.indptr.astype(np.int64)
    if mode == 'individual':
        seg_A = np.diff(indptr[np.searchsorted(d.row_cell, np.arange(S + 1))])
    else:
        seg_A = np.bincount(row_seg, weights=np.diff(indptr), minlength=S).astype(np.int64)
    order = rng.permutation(S)
    pick, tot = [], 0
    for s in order:
        if seg_A[s] == 0 or (pick and tot + seg_A[s] > budget_aln):
            continue
        pick.append(int(s))
        tot += int(seg_A[s])
        if tot >= 0.9 * budget_aln:
            break
    if mode == 'pseudobulk':
        # one model over everything: a contiguous block of rows of the budget's size is the bounded sample
        r1 = int(np.searchsorted(indptr, min(budget_aln, int(indptr[-1]))))
        rows_of = {0: np.arange(0, max(1, r1))}
        pick = [0]
    else:
        rows_of = {s: model_rows(d, row_seg, s) for s in pick}
    idx64 = d.indices
    cells = []
    for s in pick:
        ptr, idx, sc, _ = submatrix_rows(indptr, idx64, d.scores, rows_of[s])
        cells.append((ptr, idx.astype(np.int64), sc, d.K))
    cells.sort(key=lambda c: -int(c[0][-1]))                 # longest first, round-robin: balanced chunks
    use = max(1, min(cores, len(cells)))
    chunks = [c for c in (cells[i::use] for i in range(use)) if c]
    if use > 1:
        with mp.get_context('fork').Pool(use) as pool:
            pool.map(abs, range(use))                        # workers are up
            t0 = time.perf_counter()
            units = sum(pool.map(_cpu_fit_cells, chunks, 1))
            dt = time.perf_counter() - t0
    else:
        t0 = time.perf_counter()
        units = sum(_cpu_fit_cells(c) for c in chunks)
        dt = time.perf_counter() - t0
    return units / dt, units, dt, len(cells), use, sum(int(c[0][-1]) for c in cells)


def reference_as_written(d, mode, cores, n_cells=25, first_cell=0, time_limit=300.0):
    """The UNMODIFIED reference (``baseline/_ref``: the package copied / installed by ``__graft_entry__.build``)
    on ``n_cells`` consecutive cells of the workload through its own ``_fit_pooling_model(st, opts)`` with
    ``nproc=cores`` -- exactly as written, i.e. one full-height masked copy of the matrix per cell (SURVEY.md
    0.7: its cost per cell grows with the number of rows handed in, so a 25-cell sample FLATTERS it); run in a
    subprocess (``oracle/time_reference.py``) so that its ``multiprocessing.Pool`` forks from a process that
    holds no CUDA context.  Returns a dict or None (not installed / failed / over the time limit)."""
    ref_dir = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_dir, 'stellarscope')) or mode != 'individual':
        return None
    import tempfile
    first_cell = first_cell % max(1, d.n_cells - n_cells)
    c0, c1 = first_cell, min(first_cell + n_cells, d.n_cells)
    r0, r1 = (int(x) for x in np.searchsorted(d.row_cell, [c0, c1]))
    a0, a1 = int(d.indptr[r0]), int(d.indptr[r1])
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, 'sample.npz')
        np.savez(f, indptr=(d.indptr[r0:r1 + 1].astype(np.int64) - a0).astype(np.int32), indices=d.indices[a0:a1],
                 scores=d.scores[a0:a1], K=d.K, row_cell=d.row_cell[r0:r1] - c0, n_cells=c1 - c0)
        try:
            out = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'time_reference.py'), '--npz', f,
                                  '--ref', ref_dir, '--nproc', str(cores), '--json'],
                                 capture_output=True, text=True, timeout=time_limit)
            for line in out.stdout.splitlines()[::-1]:
                if line.startswith('{'):
                    return json.loads(line)
        except Exception:
            return None
    return None


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cfg = args.config
    mode = CONFIGS[cfg]['mode']
    cores = os.cpu_count() or 1
    # the sample is drawn from the workload itself when a GPU can generate it in seconds; on a box without one from
    # a 1/50 replica (same generator, same per-model size distribution)
    try:
        import torch
        on_gpu = torch.cuda.is_available()
    except Exception:
        on_gpu = False
    scale = args.scale if on_gpu else min(args.scale, 0.02)
    d = make_workload(cfg, scale, seed=0)
    if on_gpu:
        torch.cuda.empty_cache()
    replica = '' if scale == args.scale else ' (1/50 replica: no GPU to generate the full data set)'
    budget = int(args.cpu_budget_aln)
    have_ref = mode == 'individual' and os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'stellarscope'))
    line_cpu = None
    if have_ref:
        # the reference itself, through its own public API, 25 cells per step
        for k in range(args.warmup):
            reference_as_written(d, mode, cores, first_cell=25 * (1000 + k))
        tot_units, tot_dt, last = 0.0, 0.0, None
        for k in range(args.steps):
            r = reference_as_written(d, mode, cores, first_cell=25 * k)
            if r is None:
                have_ref = False
                break
            tot_units += r['value'] * r['seconds']
            tot_dt += r['seconds']
            last = r
        if have_ref:
            value = tot_units / tot_dt
            prate, _, pdt, pn, pused, paln = cpu_sample_rate(d, mode, cores, budget, seed=0)
            line_cpu = dict(value=value, unit=UNIT, cores=cores, kind='reference',
                            sample=f'{last["cells"]} consecutive cells (~{last["alignments"]} alignments) per step of this '
                                   f'workload{replica}: the UNMODIFIED reference (baseline/_ref) through '
                                   f'stellarscope.utils.model._fit_pooling_model(st, opts), individual mode, '
                                   f'nproc={cores}.  Its cost per cell grows with the rows handed in (a full-height '
                                   f'masked matrix per cell), so this small sample flatters it',
                            port=dict(value=prate, unit=UNIT, cores=pused, kind='port',
                                      sample=f'{pn} random models ({paln} alignments), numpy restatement on row-compacted '
                                             f'per-model matrices, one process per core, {pdt:.1f} s'))
    if line_cpu is None:
        for _ in range(args.warmup):
            cpu_sample_rate(d, mode, cores, budget // 8, seed=99)
        tot_units, tot_dt, n_models, used, aln = 0, 0.0, 0, cores, 0
        for k in range(args.steps):
            r, units, dt, n, used, a = cpu_sample_rate(d, mode, cores, budget, seed=k)
            tot_units += units
            tot_dt += dt
            n_models, aln = n, a
        value = tot_units / tot_dt
        line_cpu = dict(value=value, unit=UNIT, cores=used, kind='port',
                        sample=f'{n_models} randomly drawn models ({aln} alignments) per step of this workload{replica}, '
                               f'numpy restatement of the reference EM (oracle/em_oracle.py), one process per core, '
                               f'row-compacted per-model matrices')
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * tot_dt / max(1, args.steps), higher_is_better=True, scaling='strong',
                vs_baseline=None, dtype='f64', data='synthetic', impl='reference',
                config=dict(workload=CONFIGS[cfg]['label'] + ' (bounded sample)', config=cfg, scale=args.scale, **OPTS),
                cpu_baseline=line_cpu,
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# parity sample (oracle = test infrastructure, after the timed region)
# ---------------------------------------------------------------------------
def parity_sample(d, mode, eng, rows_global, local_seg, mine, res, n_sample, seed, cpu_budget_aln=6_000_000):
    """Re-fit sampled models of THIS rank with the CPU oracle and compare: iteration counts equal, lnL 1e-10, pi
    1e-6 (values >= 1e-290), best_exclude counts per locus bit-exact (rows whose top-2 margin of z lies inside the
    relative tie tolerance 1e-9 are compared under both tie rules).  The sample is stratified over the kernel
    paths: models by number of tiles (1 = lane kernel, 2..8 = cluster kernel, more = streaming kernel) and,
    among single-tile models, by size class; the largest models of the rank are always included."""
    import torch
    from oracle.em_oracle import OracleLikelihood, Opts, submatrix_rows
    rng = np.random.default_rng(seed)
    seg_A = eng.export_layout_array('seg_A', np.int64)
    seg_nt = eng.export_layout_array('seg_ntiles', np.int32)
    S = len(seg_A)
    fitted = np.flatnonzero(seg_A > 0)
    if len(fitted) == 0:
        return dict(cells=0, max_rel_pi=0.0, counts_equal=True, iters_equal=True, lnl_ok=True, near_tie_rows=0,
                    by_path={})
    # strata: (tiles, size class)
    cls = np.where(seg_nt == 1, np.minimum(12, np.log2(np.maximum(seg_A, 1)).astype(int)), 100 + np.minimum(seg_nt, 9))
    pick = set(int(s) for s in fitted[np.argsort(-seg_A[fitted])[:3]])
    strata = {}
    for s in rng.permutation(fitted):
        strata.setdefault(int(cls[s]), []).append(int(s))
    per = max(2, n_sample // max(1, len(strata)))
    budget = cpu_budget_aln
    for k in sorted(strata):
        for s in strata[k][:per]:
            if s not in pick and seg_A[s] <= budget:
                pick.add(s)
                budget -= int(seg_A[s])
    pick = sorted(pick)
    # device results of the sampled models
    p = eng.params_device()
    z = eng.emit_z()
    local_seg_t = torch.from_numpy(np.ascontiguousarray(local_seg)).to(z.device)
    flag, nbest, rinfo = eng.reassign('best_exclude', z, tie_tol=TIE_TOL)
    cell, col, val = eng.count(flag, local_seg_t)            # counts per (model, locus), sorted by model
    ptr64, idx64 = d.indptr.astype(np.int64), d.indices
    o = Opts(**OPTS)
    worst_pi, near, ok_it, ok_lnl, ok_cnt = 0.0, 0, True, True, True
    by_path = {}
    bad = []
    for s in pick:
        rows_l = np.flatnonzero(local_seg == s)
        rows = rows_global[rows_l]
        ptr, idx, sc, _ = submatrix_rows(ptr64, idx64, d.scores, rows)
        om = OracleLikelihood(ptr, idx.astype(np.int64), sc, d.K, o).em()
        it_ok = int(res['iters'][s]) == len(om.fitinfo.iterations)
        lo = float(om.lnl)
        l_ok = abs(float(res['lnl'][s]) - lo) <= 1e-10 * abs(lo)
        a, n = int(p['seg_col0'][s]), int(p['seg_ncol'][s])
        pi = np.full(d.K, float(p['pi_rest'][s]))
        pi[p['gcol'][a:a + n].cpu().numpy()] = p['pi'][a:a + n].cpu().numpy()
        po = np.asarray(om.pi, dtype=np.float64).ravel()
        big = po > 1e-290
        rel = float(np.max(np.abs(pi[big] - po[big]) / po[big])) if big.any() else 0.0
        worst_pi = max(worst_pi, rel)
        # best_exclude counts per locus of this model: exact-tie rule and tolerance rule of the oracle's z
        zo = np.asarray(om.z, dtype=np.float64)
        rix = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
        zmax = np.zeros(len(ptr) - 1)
        np.maximum.at(zmax, rix, zo)

        def counts(tol):
            best = (zo >= zmax[rix] * (1.0 - tol)) & (zo > 0)
            nb = np.bincount(rix[best], minlength=len(ptr) - 1)
            keep = best & (nb[rix] == 1)
            return np.bincount(idx[keep], minlength=d.K)
        c_exact, c_tol = counts(0.0), counts(TIE_TOL)
        near += int(np.abs(c_exact - c_tol).sum() > 0)
        lo_, hi_ = np.searchsorted(cell, [s, s + 1])
        c_gpu = np.zeros(d.K, dtype=np.int64)
        c_gpu[col[lo_:hi_]] = val[lo_:hi_].astype(np.int64)
        c_ok = bool(np.array_equal(c_gpu, c_exact) or np.array_equal(c_gpu, c_tol))
        ok_it &= it_ok
        ok_lnl &= l_ok
        ok_cnt &= c_ok
        path = 'lane' if seg_nt[s] == 1 else ('cluster' if seg_nt[s] <= 8 else 'stream')
        by_path[path] = by_path.get(path, 0) + 1
        if not (it_ok and l_ok and c_ok and rel < 1e-6):
            bad.append((int(s), int(seg_A[s]), int(seg_nt[s]), it_ok, l_ok, c_ok, rel))
    out = dict(cells=len(pick), max_rel_pi=worst_pi, counts_equal=bool(ok_cnt), iters_equal=bool(ok_it),
               lnl_ok=bool(ok_lnl), near_tie_models=near, by_path=by_path,
               largest_alignments=int(seg_A[pick].max()) if pick else 0)
    if bad:
        out['mismatches'] = bad[:10]
    return out


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def device_local_matrix(d, rows, dev):
    """Device CSR (int32 row_ptr, int32 col_idx, uint16 score) of the given rows of the data set, cut out of the
    device-resident full matrix when the generator left one, else uploaded from the host (untimed set-up)."""
    import torch
    full = getattr(d, 'dev', None)
    if full is None:
        from stellarscope_b200 import hostutil
        ptr, idx, dat = hostutil.gather_rows(d.indptr, d.indices, d.scores, rows)
        return (torch.from_numpy(ptr.astype(np.int32)).to(dev), torch.from_numpy(idx).to(dev),
                torch.from_numpy(dat.view(np.int16)).to(dev).view(torch.uint16))
    indptr, indices, scores = full
    if len(rows) == d.N:
        return indptr, indices, scores
    r = torch.from_numpy(rows).to(dev)
    p64 = indptr.to(torch.int64)
    lens = p64[r + 1] - p64[r]
    lptr = torch.zeros(len(rows) + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=lptr[1:])
    tot = int(lptr[-1])
    take = torch.repeat_interleave(p64[r] - lptr[:-1], lens, output_size=tot) + torch.arange(tot, device=dev)
    return lptr.to(torch.int32), indices[take], scores.view(torch.int16)[take].view(torch.uint16)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from stellarscope_b200.engine import Engine
    from stellarscope_b200 import model, parallel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    cfg = args.config
    mode = CONFIGS[cfg]['mode']
    t_gen = time.perf_counter()
    d = make_workload(cfg, args.scale, seed=0, keep_device=True)     # every rank generates the SAME data set
    gen_s = time.perf_counter() - t_gen
    if world > 1:
        chk = torch.tensor([float(d.A), float(d.N), float(d.scores[::97].astype(np.int64).sum()),
                            float(d.indices[::89].astype(np.int64).sum())], dtype=torch.float64, device=dev)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not torch.equal(lo, hi):
            raise RuntimeError('the ranks generated different data sets')
    opts = BenchOpts(mode)
    row_seg, S = row_models(d, mode)

    # ---- resident-input path: this rank's rows in HBM, layout built on device, timed region = the fit ----
    if mode == 'pseudobulk':
        r0, r1 = parallel.row_block(d.indptr, rank, world)
        rows = np.arange(r0, r1, dtype=np.int64)
        local_seg = np.zeros(len(rows), dtype=np.int32)
        mine = np.array([0])
    else:
        plan = parallel.plan_local_rows(d.indptr, row_seg, S, world, rank, collective=world > 1)
        rows, local_seg, mine = plan['rows'], plan['local_seg'], plan['mine']
    eng = Engine(local)
    if mode == 'pseudobulk' and world > 1:
        eng.set_sharding(rank, world, None, peers=True, K=d.K)
    eng.set_matrix_device(*device_local_matrix(d, rows, dev), d.K)
    if hasattr(d, 'dev'):
        d.dev = None                                                  # the full device copy may go
        d.dev_row_cell = None
    build_ms = None
    for _ in range(2):                         # the first build pays for the memory pool; report the second
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        eng.build_layout(local_seg, len(mine))
        torch.cuda.synchronize(dev)
        build_ms = 1e3 * (time.perf_counter() - t0)
    info = eng.layout_info()
    eng.enable_timing(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def one_fit():
        flush.fill_(1)                                                  # evict L2 between timed steps
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eng.fit(**OPTS)
        e.record()
        return s, e

    sampler = ClockSampler(local, args.clock_interval)          # samples clocks / throttle reasons from the warm-up on
    sampler.start()
    for _ in range(args.warmup):
        one_fit()
    barrier()
    l0 = eng.launch_count()
    evs = []
    for _ in range(args.steps):
        evs.append(one_fit())
    barrier()
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    fit_ms = [s.elapsed_time(e) for s, e in evs]
    groups = eng.last_timings() if not getattr(eng, 'sharded', False) else {}
    res = eng.results(traces=False)
    seg_A = eng.export_layout_array('seg_A', np.int64)
    seg_Aamb = eng.export_layout_array('seg_Aamb', np.int64)
    seg_namb = eng.export_layout_array('seg_namb', np.int32).astype(np.int64)
    seg_ntiles = eng.export_layout_array('seg_ntiles', np.int32)
    iters = res['iters'].astype(np.int64)
    units_step = int((seg_A * iters).sum())
    total_ms = float(sum(fit_ms))
    # algorithmic bytes of the fused E+M kernels, SURVEY.md 8d: per iteration
    # 12 B per ambiguous alignment + 4 B per ambiguous row + 16 B per active column
    if getattr(eng, 'sharded', False):
        seg_ncol = np.array([d.K], dtype=np.int64)
    else:
        seg_ncol = eng.params_device()['seg_ncol'].astype(np.int64)
    per_iter = 12 * seg_Aamb + 4 * seg_namb + 16 * seg_ncol
    alg_bytes = float((per_iter * iters)[seg_ntiles > 0].sum())
    lane_b = float((per_iter * iters)[seg_ntiles == 1].sum())
    clu_b = float((per_iter * iters)[(seg_ntiles > 1) & (seg_ntiles <= 8)].sum())
    str_b = float((per_iter * iters)[seg_ntiles > 8].sum())
    em_ms = groups.get('em', total_ms / max(1, args.steps))

    # ---- parity sample (this rank's models; CPU oracle) ----
    par = None
    if mode != 'pseudobulk' and args.parity > 0:
        par = parity_sample(d, mode, eng, rows, local_seg, mine, res, max(4, args.parity // world), seed=rank)
    eng.close()
    del eng, flush
    torch.cuda.empty_cache()

    # ---- e2e: reference-facing API from host scipy matrices ----
    st = make_st(d, opts)
    e2e_times, e2e_units = [], 0
    h2d = d2h = 0
    n_e2e = max(1, min(args.steps, 3))
    for i in range(1 + n_e2e):                      # first pass = warm-up
        barrier()
        t0 = time.perf_counter()
        tl, poolinfo = model._fit_pooling_model(st, opts)
        barrier()
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_times.append(dt)
        iters_e = tl._results['iters'].astype(np.int64)
        e2e_units = int((tl._engine.export_layout_array('seg_A', np.int64) * iters_e).sum())
        h2d = tl._engine.h2d_bytes
        d2h = (4 * 4 + 8) * tl.n_segments      # iterations, flags, nobs, nfeats, lnL per model (the diff traces stay
                                               # on the device until a FitInfo.iterations is looked at)
        nmod = len(poolinfo.models_info)
        host_t = getattr(tl, 'host_timings', None)
        if i < n_e2e:
            tl.close()
    e2e_dt = float(np.mean(e2e_times))

    # ---- the whole drop-in flow once: fit (above) + Stellarscope.reassign over the mode sweep + per-cell count
    #      matrices + the MatrixMarket files (BASELINE.json configs[4]: reassign_mode sweep) ----
    full = None
    if args.full_flow and (world == 1 or args.full_flow_multi):
        import shutil
        import tempfile
        sweep = ['best_exclude', 'best_conf', 'best_random', 'best_average', 'total_hits']
        opts.reassign_mode = sweep
        st.opts = opts
        tmpdir = tempfile.mkdtemp(prefix='ss_bench_')
        opts.outfile_path = lambda suffix: os.path.join(tmpdir, suffix)
        try:
            barrier()
            t0 = time.perf_counter()
            model.reassign(st, tl)
            barrier()
            t1 = time.perf_counter()
            write = shutil.disk_usage(tmpdir).free > 40 * d.A           # ~14 bytes of text per count entry, 5 files
            if write:
                model.output_report(st, tl)
            else:
                model.count_matrices(st, tl)
            barrier()
            t2 = time.perf_counter()
            mtx_bytes = sum(os.path.getsize(os.path.join(tmpdir, f)) for f in os.listdir(tmpdir)) if rank == 0 else 0
            full = dict(reassign_s=t1 - t0, report_s=t2 - t1, mtx_written=bool(write), mtx_bytes=int(mtx_bytes),
                        modes=sweep)
        finally:
            shutil.rmtree(tmpdir, ignore_errors=True)
            st.reassignments = {}
    tl.close()

    # ---- reduce over ranks ----
    tvec = torch.tensor([total_ms, e2e_dt, em_ms, build_ms, full['reassign_s'] if full else 0.0,
                         full['report_s'] if full else 0.0], dtype=torch.float64, device=dev)
    pv = [1.0, 1.0, 1.0, 1.0, 0.0, 0.0] if par is None else [
        float(par['counts_equal']), float(par['iters_equal']), float(par['lnl_ok']), float(par['max_rel_pi'] < 1e-6),
        float(par['max_rel_pi']), float(par['cells'])]
    uvec = torch.tensor([float(units_step), float(e2e_units), float(launches), float(h2d), float(d2h), alg_bytes,
                         lane_b, clu_b, str_b, float(info['n_tiles']), float(info['n_cluster_tiles']),
                         float(info['n_stream_tiles']), pv[5], float(len(mine))],
                        dtype=torch.float64, device=dev)
    pmin = torch.tensor(pv[:4], dtype=torch.float64, device=dev)
    pmax = torch.tensor([pv[4]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
        dist.all_reduce(uvec, op=dist.ReduceOp.SUM)
        dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
    tmax_ms, e2e_max, em_max, build_max, reassign_max, report_max = tvec.tolist()
    (units_all, e2e_units_all, launches_all, h2d_all, d2h_all, alg_all, lane_all, clu_all, str_all, tiles_all,
     ctiles_all, stiles_all, par_cells, models_all) = uvec.tolist()
    if mode == 'pseudobulk':
        # ONE model: every rank reports the same iteration count for its block; units = all alignments x iterations
        e2e_units_all = e2e_units_all
    value = units_all * args.steps / (tmax_ms / 1e3)
    e2e_value = e2e_units_all / e2e_max
    parity_ok = bool(pmin.min().item() >= 1.0)

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        achieved = alg_all / (em_max / 1e3) / 1e9 if em_max > 0 else 0.0
        cores = os.cpu_count() or 1
        cpu_rate, cpu_units, cpu_dt, cpu_n, cpu_used, cpu_aln = cpu_sample_rate(d, mode, cores, args.cpu_budget_aln, seed=1)
        asw = reference_as_written(d, mode, cores) if args.ref_as_written else None
        traffic, traffic_src = NCU_TRAFFIC.get((cfg, float(args.scale)), (None, None))
        if world != 1:
            traffic = None
        if traffic is None:
            traffic_src = {f'{k[0]} x{k[1]}': dict(bytes_per_fit=v[0], source=v[1]) for k, v in NCU_TRAFFIC.items()}
        fitted = seg_A > 0
        # the resource that binds the on-chip kernels (lane / cluster): issue slots.  Warp instructions per unit from the
        # ncu capture of one cfg5 fit (sum of smsp__inst_executed over the EM kernels / units of the fit)
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_hz = 1e6 * float((clocks or {}).get('sm_mhz') or 1965.0)
        winst = 3.59 if mode == 'individual' else None
        on_chip = None if winst is None else dict(
            bound='issue_slots', warp_instructions_per_unit=winst,
            achieved=winst * value / world / 1e9, peak=sm_count * 4 * sm_hz / 1e9, unit='G warp-instructions/s per GPU',
            frac=winst * value / world / (sm_count * 4 * sm_hz),
            source='profiles/r2l_cfg5_x0.05_em_ncu.csv: 8.72 G warp instructions for 2.43 G alignments x iterations (captured '
                   'before the pushes of the cluster kernel moved into a compact loop: a few percent fewer since); peak = '
                   'SMs x 4 schedulers x SM clock (the kernels also wait on shared memory and barriers: 45-62 % of the '
                   'issue slots is what the best of them reach)')
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=tmax_ms / args.steps, higher_is_better=True, scaling='strong', vs_baseline=None,
            dtype='f64', data='synthetic',
            config=dict(workload=CONFIGS[cfg]['label'], config=cfg, scale=args.scale,
                        l2='inputs larger than L2 for cfg3-cfg5; 256 MB written between timed steps as well',
                        n_cells=d.n_cells, alignments=d.A, rows=d.N, K=d.K,
                        models=1 if mode == 'pseudobulk' else int(models_all), **OPTS,
                        mean_iters=float(iters[fitted].mean()) if fitted.any() else 0.0,
                        units_per_step=int(units_all),
                        per_model_fit_us=1e3 * (tmax_ms / args.steps) / max(1.0, models_all),
                        layout_build_ms=build_max, generate_s=gen_s, tiles=int(tiles_all),
                        cluster_tiles=int(ctiles_all), stream_tiles=int(stiles_all),
                        kernel_ms_rank0=groups,
                        parallelism=(f'rows of the one model in {world} block(s), partial theta sums exchanged per '
                                     f'iteration inside the kernel (NVLink peer memory)' if mode == 'pseudobulk' else
                                     f'models packed onto {world} GPU(s) by LPT on alignment counts, no data-path '
                                     f'collective; per-model records gathered as tensors')),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d_all), d2h_bytes_per_step=int(d2h_all),
                     seconds_per_step=e2e_max, models_returned=int(nmod), host_breakdown_rank0_s=host_t,
                     api='stellarscope_b200.model._fit_pooling_model(st, opts), st from host scipy CSR + row table'),
            e2e_full=(dict(seconds=e2e_max + reassign_max + report_max, fit_s=e2e_max, reassign_s=reassign_max,
                           count_and_mtx_s=report_max, value=e2e_units_all / (e2e_max + reassign_max + report_max),
                           unit=UNIT, modes=full['modes'], mtx_written=full['mtx_written'], mtx_bytes=full['mtx_bytes'],
                           what='_fit_pooling_model + Stellarscope.reassign over the five modes + output_report '
                                '(per-cell count matrices, barcodes / features, one MatrixMarket file per mode), '
                                'host matrices in, files out; max over ranks per stage') if full else None),
            gpu_launches=int(launches_all),
            roofline=dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak,
                          traffic=traffic, traffic_source=traffic_src,
                          kernel='fused E+M kernels of one fit, launched concurrently: em_lane_kernel (single-tile '
                                 'models, register-resident), em_clane_kernel (thread-block clusters: models of 2..8 '
                                 'tiles), em_stream_kernel (larger models, tiles streamed by TMA); '
                                 'em_dense_kernel for the row-sharded model',
                          note='achieved = ALGORITHMIC bytes / span of the EM kernels: iterations x (12 B per ambiguous '
                               'alignment + 4 B per row + 16 B per active column), SURVEY.md 8d.  Models that fit on '
                               'chip are read from HBM ONCE and iterated out of registers / shared memory, so for '
                               'them this is an EFFECTIVE bandwidth, not DRAM traffic (ncu: DRAM traffic of a cfg2 '
                               'fit is 1.3 % of these bytes; the binding resources are the shared-memory pipe and '
                               'issue slots, profiles/); only the streamed models (stream_bytes) consume HBM '
                               'bandwidth per iteration',
                          binding_resource='shared-memory pipe / issue slots for lane + cluster kernels (on-chip '
                                           'resident), HBM + L2 for the streaming kernel',
                          on_chip=on_chip,
                          ncu_utilisation=dict(
                              source='profiles/r2l_cfg5_x0.05_em_ncu.csv, profiles/r2h_stream_final_cfg4_ncu.csv (ncu --set full; '
                                     'percent of peak sustained; shares = serialised kernel times of one cfg5 fit)',
                              em_clane_kernel_light=dict(share_of_em_time_cfg5=0.52, issue_slots='32-39', shared_memory_pipe='24-41',
                                                         registers=64, warps_per_sm='28-31 (two CTAs per SM)'),
                              em_clane_kernel_heavy=dict(share_of_em_time_cfg5=0.20, issue_slots='26-33', shared_memory_pipe='15-22',
                                                         registers=128, warps_per_sm=16),
                              em_lane_kernel=dict(share_of_em_time_cfg5=0.12, issue_slots='31-45', shared_memory_pipe='39-56',
                                                  registers=128, warps_per_sm='12-16 (32 for the 1024-thread class)'),
                              em_stream_kernel=dict(share_of_em_time_cfg5=0.16, issue_slots='21 (cfg5 x0.05) / 55 (cfg4)',
                                                    shared_memory_pipe='17 / 51', dram_pct_of_measured_peak=38.8,
                                                    registers=64, warps_per_sm=32)),
                          peak_source=peak_src, kernel_ms=em_max, algorithmic_bytes=alg_all,
                          lane_bytes=lane_all, cluster_bytes=clu_all, stream_bytes=str_all),
            parity_sample=dict(cells=int(par_cells), max_rel_pi=float(pmax.item()), counts_equal=bool(pmin[0].item() >= 1),
                               iters_equal=bool(pmin[1].item() >= 1), lnl_1e10=bool(pmin[2].item() >= 1),
                               pi_1e6=bool(pmin[3].item() >= 1), rank0=par,
                               what='models re-fitted by the CPU oracle after the timed region (all ranks, stratified '
                                    'over lane / cluster / streaming paths and size classes, largest models '
                                    'included); best_exclude counts per locus compared bit-exactly') if par is not None
            else None,
            cpu_baseline=dict(value=cpu_rate, unit=UNIT, cores=cpu_used, kind='port',
                              sample=f'{cpu_n} random models of this workload ({cpu_aln} alignments), numpy restatement of '
                                     f'the reference EM (oracle/em_oracle.py), one process per core, {cpu_dt:.1f} s',
                              **({'reference_as_written': asw} if asw else {})),
            clocks=clocks)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if parity_ok else 3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='cfg5', choices=sorted(CONFIGS))
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--parity', type=int, default=120, help='models re-fitted by the CPU oracle (0: skip)')
    ap.add_argument('--cpu-budget-aln', dest='cpu_budget_aln', type=float, default=4e6,
                    help='alignments per CPU-baseline sample (about 10-30 s of CPU work)')
    ap.add_argument('--clock-interval', dest='clock_interval', type=float, default=0.25,
                    help='seconds between nvidia-smi clock samples during the timed region')
    ap.add_argument('--no-ref-as-written', dest='ref_as_written', action='store_false')
    ap.add_argument('--full-flow-multi', dest='full_flow_multi', action='store_true',
                    help='run the e2e_full pass under torchrun as well (default: single GPU only)')
    ap.add_argument('--no-full-flow', dest='full_flow', action='store_false',
                    help='skip the one pass of reassign sweep + count matrices + MatrixMarket files (e2e_full)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
