#!/usr/bin/env python
"""Benchmark of the EM read-reassignment hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one EM fit (``ss_em_fit``: every cell's model to convergence) over
one batch of synthetic 10x-style input.  Workload at any N: BASELINE.json
``configs[1]`` per GPU -- 10k cells, 5M multimapped alignments over 28k loci,
pooling_mode=individual (weak scaling: every rank fits its own 10k cells;
individual mode shards cells across GPUs with no data-path collective,
SURVEY.md 8e).

``value``  alignments x EM-iterations / s, inputs resident in HBM, device time
           of the fit (CUDA events), max over ranks.
``e2e``    same metric through the reference-facing Python API
           (``stellarscope_b200.model._fit_pooling_model``) starting from HOST
           scipy matrices: host prep + H2D + device layout build + fit + D2H of
           the per-model results, wall clock with a device synchronise.
``--impl reference``  the CPU restatement of the reference's algorithm
           (oracle/em_oracle.py, numpy, one process per host core) on a bounded
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
WORKLOAD = dict(n_cells=10_000, n_alignments=5_000_000, n_loci=28_000)
# dram__bytes_read.sum + dram__bytes_write.sum of the 16 fused E+M kernels of ONE fit of this workload, from the
# ncu --set full capture profiles/r1e_em_full_ncu.csv (110.8 MB read + 0.2 MB written): every cell is read from
# HBM once; the iterations run out of registers / shared memory
NCU_DRAM_BYTES_PER_FIT = 111.0e6
NCU_DRAM_SOURCE = 'profiles/r1e_em_full_ncu.csv (ncu --set full, one fit, sum over the 16 EM kernels)'
CPU_SAMPLE_CELLS = 10000     # the CPU baseline fits every cell of the workload (about 10-15 s on 16 cores)
OPTS = dict(em_epsilon=1e-7, max_iter=100, pi_prior=0, theta_prior=200000)   # stellarscope_assign.yaml:231-250


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

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
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
            self._halt.wait(0.02)

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


def make_workload(seed, scale=1.0):
    from stellarscope_b200 import synth
    return synth.generate(n_cells=max(1, int(WORKLOAD['n_cells'] * scale)),
                          n_alignments=max(100, int(WORKLOAD['n_alignments'] * scale)),
                          n_loci=WORKLOAD['n_loci'], seed=seed)


def make_st(d, opts):
    """The attributes of ``Stellarscope`` the hot path reads."""
    import types
    import scipy.sparse as sp
    st = types.SimpleNamespace()
    st.opts = opts
    st.raw_scores = sp.csr_matrix((d.scores, d.indices, d.indptr), shape=(d.N, d.K))
    st.shape = st.raw_scores.shape
    st.corrected = None
    st.bcode_ridx_map = {f'BC{c:06d}': set(rows.tolist()) for c, rows in enumerate(d.cell_rows())}
    st.celltypes = []
    st.ctype_bcode_map = {}
    return st


class BenchOpts:
    def __init__(self):
        self.__dict__.update(OPTS)
        self.use_likelihood = False
        self.pooling_mode = 'individual'
        self.ignore_umi = True
        self.umi_counts = False
        self.nproc = 1
        self.reassign_mode = ['best_exclude']
        self.conf_prob = 0.9
        self.rng = np.random.default_rng(0)
        self.shard_models = False      # weak scaling: every rank fits its own 10k cells


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ---------------------------------------------------------------------------
def _cpu_fit_cells(cells):
    """cells: list of (ptr, idx, scores, K) row-compacted per-cell matrices."""
    from oracle.em_oracle import OracleLikelihood, Opts
    o = Opts(**OPTS)
    units = 0
    for ptr, idx, sc, K in cells:
        m = OracleLikelihood(ptr, idx, sc, K, o).em()
        units += int(ptr[-1]) * len(m.fitinfo.iterations)
    return units


def cpu_sample_rate(d, n_cells_sample, cores, seed=0):
    """alignments x iterations / s of the oracle port on ``cores`` processes over a
    random sample of cells (row-compacted per-cell submatrices, bit-identical to
    the reference's masking).  The per-cell matrices are cut out
    before the clock starts (the reference's generator does that inside
    ``_fit_pooling_model``; it is not part of the EM); the worker pool is up
    before the clock starts as well."""
    import multiprocessing as mp
    from oracle.em_oracle import submatrix_rows
    rng = np.random.default_rng(seed)
    cell_rows = d.cell_rows()
    pick = rng.choice(len(cell_rows), size=min(n_cells_sample, len(cell_rows)), replace=False)
    indptr, indices = d.indptr.astype(np.int64), d.indices.astype(np.int64)
    cells = []
    for i in pick:
        ptr, idx, sc, _ = submatrix_rows(indptr, indices, d.scores, cell_rows[i])
        cells.append((ptr, idx, sc, d.K))
    # longest cells first, round-robin: balanced chunks
    cells.sort(key=lambda c: -int(c[0][-1]))
    chunks = [cells[i::cores] for i in range(cores)]
    chunks = [c for c in chunks if c]
    if cores > 1:
        with mp.get_context('fork').Pool(cores) as pool:
            pool.map(abs, range(cores))                      # workers are up
            t0 = time.perf_counter()
            units = sum(pool.map(_cpu_fit_cells, chunks, 1))
            dt = time.perf_counter() - t0
    else:
        t0 = time.perf_counter()
        units = sum(_cpu_fit_cells(c) for c in chunks)
        dt = time.perf_counter() - t0
    return units / dt, units, dt, len(cells)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    d = make_workload(seed=0, scale=0.1)            # the sample is drawn from a 1/10 replica (same per-cell sizes)
    n_sample = 1000
    for _ in range(args.warmup):
        cpu_sample_rate(d, cores, cores, seed=99)
    rates, tot_units, tot_dt = [], 0, 0.0
    for k in range(args.steps):
        r, units, dt, n = cpu_sample_rate(d, n_sample, cores, seed=k)
        rates.append(r)
        tot_units += units
        tot_dt += dt
    value = tot_units / tot_dt
    sample = (f'{n_sample} randomly drawn cells per step of the configs[1] workload shape '
              f'(10k cells / 5M alignments, 1/10 replica), numpy restatement of the reference EM, '
              f'one process per core, row-compacted per-cell matrices')
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * tot_dt / max(1, args.steps), higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f64', data='synthetic', impl='reference',
                config=dict(workload='configs[1]: synthetic 10x, 10k cells, 5M alignments, 28k loci, '
                                     'pooling_mode=individual (bounded sample)', **OPTS),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind='port', sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from stellarscope_b200.engine import Engine
    from stellarscope_b200 import model

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

    d = make_workload(seed=rank)
    opts = BenchOpts()

    # ---- resident-input path: matrix in HBM, layout built on device, timed region = the fit ----
    eng = Engine(local)
    eng.set_matrix(d.indptr, d.indices, d.scores, d.K)
    build_ms = None
    for _ in range(2):                         # the first build pays for the memory pool; report the second
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        eng.build_layout(d.row_cell, d.n_cells)
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

    sampler = ClockSampler(local)          # samples clocks / throttle reasons from the warm-up on
    sampler.start()
    for _ in range(args.warmup):
        one_fit()
    barrier()
    l0 = eng.launch_count()
    evs = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        evs.append(one_fit())
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    fit_ms = [s.elapsed_time(e) for s, e in evs]
    groups = eng.last_timings()
    res = eng.results(traces=False)
    p = eng.params()
    # units of one step: sum over models of (alignments of the model) x (its iterations)
    seg_A = eng.export_layout_array('seg_A', np.int64)
    seg_Aamb = eng.export_layout_array('seg_Aamb', np.int64)
    seg_namb = eng.export_layout_array('seg_namb', np.int32)
    seg_ntiles = eng.export_layout_array('seg_ntiles', np.int32)
    iters = res['iters'].astype(np.int64)
    units_step = int((seg_A * iters).sum())
    total_ms = float(sum(fit_ms))
    # algorithmic bytes of the fused E+M kernels, SURVEY.md 8d: per iteration
    # 12 B per ambiguous alignment + 4 B per ambiguous row + 16 B per active column
    per_iter = 12 * seg_Aamb + 4 * seg_namb.astype(np.int64) + 16 * p['seg_ncol'].astype(np.int64)
    resident = seg_ntiles == 1
    alg_bytes_all = float((per_iter * iters)[seg_ntiles > 0].sum())
    alg_bytes = float((per_iter * iters)[resident].sum())
    multi_bytes = alg_bytes_all - alg_bytes

    # ---- e2e: reference-facing API from host scipy matrices ----
    st = make_st(d, opts)
    e2e_times, e2e_units = [], 0
    h2d = d2h = 0
    n_e2e = max(1, min(args.steps, 3))
    for i in range(1 + n_e2e):                      # first pass = warm-up
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        tl, poolinfo = model._fit_pooling_model(st, opts)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_times.append(dt)
        iters_e = tl._results['iters'].astype(np.int64)
        e2e_units = int((tl._engine.export_layout_array('seg_A', np.int64) * iters_e).sum())
        h2d = tl._engine.h2d_bytes
        d2h = (4 * 4 + 8) * tl.n_segments      # iterations, flags, nobs, nfeats, lnL per model (the diff traces stay
                                               # on the device until a FitInfo.iterations is looked at)
        tl.close()
    e2e_dt = float(np.mean(e2e_times))

    # ---- reduce over ranks ----
    tvec = torch.tensor([total_ms, e2e_dt], dtype=torch.float64, device=dev)
    uvec = torch.tensor([float(units_step), float(e2e_units), float(launches), float(h2d), float(d2h)],
                        dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
        dist.all_reduce(uvec, op=dist.ReduceOp.SUM)
    tmax_ms, e2e_max = tvec.tolist()
    units_all, e2e_units_all, launches_all, h2d_all, d2h_all = uvec.tolist()
    value = units_all * args.steps / (tmax_ms / 1e3)
    e2e_value = e2e_units_all / e2e_max

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        em_ms = groups['em']                      # all fused E+M kernels of one fit (they run concurrently)
        achieved = alg_bytes_all / (em_ms / 1e3) / 1e9 if em_ms > 0 else 0.0
        cores = os.cpu_count() or 1
        cpu_rate, cpu_units, cpu_dt, cpu_n = cpu_sample_rate(d, CPU_SAMPLE_CELLS, cores, seed=1)
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=tmax_ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
            dtype='f64', data='synthetic',
            config=dict(workload='configs[1] per GPU: synthetic 10x, 10k cells, 5M multimapped alignments, 28k loci, '
                                 'pooling_mode=individual', l2='flushed (256 MB write) between timed steps',
                        n_cells=d.n_cells, alignments=d.A, rows=d.N, K=d.K, **OPTS,
                        mean_iters=float(iters[seg_A > 0].mean()), units_per_step=units_step,
                        per_cell_fit_us=1e3 * (tmax_ms / args.steps) / max(1, d.n_cells),
                        layout_build_ms=build_ms, tiles=info['n_tiles'], stream_tiles=info['n_stream_tiles'],
                        cluster_tiles=info['n_cluster_tiles'],
                        kernel_ms=groups, parallelism=f'cells sharded over {world} GPU(s), no collective'),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d_all), d2h_bytes_per_step=int(d2h_all),
                     seconds_per_step=e2e_max, api='stellarscope_b200.model._fit_pooling_model(st, opts)'),
            gpu_launches=int(launches_all),
            roofline=dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak,
                          traffic=NCU_DRAM_BYTES_PER_FIT if world == 1 else None,
                          traffic_source=NCU_DRAM_SOURCE,
                          kernel='fused E+M kernels of one fit: em_lane_kernel (one launch per CTA-size class) + '
                                 'em_clane_kernel (thread-block clusters: cells of 2..8 tiles), launched concurrently',
                          note='EFFECTIVE bandwidth: algorithmic bytes = iterations x (12 B/ambiguous alignment + '
                               '4 B/row + 16 B/active column), SURVEY.md 8d; every model is read from HBM once and '
                               'iterated out of shared memory / registers, so this is not DRAM traffic (see profiles/ '
                               'for dram__bytes); the binding resource is the shared-memory data path',
                          peak_source=peak_src, kernel_ms=em_ms, algorithmic_bytes=alg_bytes_all,
                          resident_bytes=alg_bytes, multi_tile_bytes=multi_bytes,
                          span_ms=dict(resident=groups['resident'], cluster=groups.get('cluster', 0.0),
                                       stream=groups['stream'])),
            cpu_baseline=dict(value=cpu_rate, unit=UNIT, cores=cores, kind='port',
                              sample=f'{cpu_n} random cells of this workload, numpy restatement of the reference EM '
                                     f'(oracle/em_oracle.py), one process per core, {cpu_dt:.1f} s'),
            clocks=clocks)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
