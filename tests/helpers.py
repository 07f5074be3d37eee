"""Shared test helpers: golden fixtures, the `st` namespace the hot path reads,
comparison utilities."""
import glob
import os
import types

import numpy as np
import scipy.sparse as sp

from oracle.em_oracle import Opts
from stellarscope_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MODES = ['best_exclude', 'best_conf', 'best_random', 'best_average', 'initial_unique',
         'initial_random', 'total_hits']
PI_FLOOR = 1e-290      # below this the reference's values only exist in float128


def golden_names():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, 'ref_*.npz')))


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, f'ref_{name}.npz'), allow_pickle=False)


def golden_opts(g, cls=Opts, **extra):
    kw = dict(em_epsilon=float(g['opt_em_epsilon']), max_iter=int(g['opt_max_iter']),
              pi_prior=float(g['opt_pi_prior']), theta_prior=float(g['opt_theta_prior']),
              use_likelihood=bool(g['opt_use_likelihood']), pooling_mode=str(g['opt_pooling_mode']),
              conf_prob=float(g['opt_conf_prob']), ignore_umi=bool(g['opt_ignore_umi']),
              umi_counts=bool(g['opt_umi_counts']))
    kw.update(extra)
    return cls(**kw)


def golden_matrix(g):
    """(indptr, indices, scores) of the matrix the fit uses (corrected if UMI-deduplicated)."""
    if bool(g['corrected']):
        return g['c_indptr'].astype(np.int64), g['c_indices'].astype(np.int64), g['c_scores']
    return g['indptr'].astype(np.int64), g['indices'].astype(np.int64), g['scores']


def golden_segments(g):
    """Row lists per model, in the reference's model order."""
    mode = str(g['opt_pooling_mode'])
    N = len(g['indptr']) - 1
    rc = g['row_cell']
    if mode == 'pseudobulk':
        return [np.arange(N)]
    if mode == 'individual':
        return [np.flatnonzero(rc == c) for c in range(int(g['n_cells']))]
    ct = g['cell_type'][rc]
    return [np.flatnonzero(ct == t) for t in range(int(g['n_celltypes']))]


class StOpts:
    def __init__(self, **kw):
        self.em_epsilon = 1e-7
        self.max_iter = 100
        self.pi_prior = 0
        self.theta_prior = 200000
        self.use_likelihood = False
        self.pooling_mode = 'pseudobulk'
        self.reassign_mode = list(MODES)
        self.conf_prob = 0.9
        self.ignore_umi = True
        self.umi_counts = False
        self.nproc = 1
        self.rng = np.random.default_rng(2024)
        self.version = 'test'
        self.__dict__.update(kw)


def make_st(g, opts):
    """Namespace carrying what the hot path reads from ``Stellarscope``
    (same construction as oracle/make_golden.py)."""
    N = len(g['indptr']) - 1
    K = int(g['K'])
    st = types.SimpleNamespace()
    st.opts = opts
    st.raw_scores = sp.csr_matrix((g['scores'], g['indices'], g['indptr']), shape=(N, K))
    st.shape = st.raw_scores.shape
    st.corrected = None
    if bool(g['corrected']):
        st.corrected = sp.csr_matrix((g['c_scores'], g['c_indices'], g['c_indptr']), shape=(N, K))
    n_cells = int(g['n_cells'])
    names = [f'BC{c:04d}' for c in range(n_cells)]
    rc = g['row_cell']
    st.bcode_ridx_map = {names[c]: set(int(r) for r in np.flatnonzero(rc == c)) for c in range(n_cells)}
    nct = int(g['n_celltypes'])
    st.celltypes = [f'type{t}' for t in range(nct)]
    st.ctype_bcode_map = {f'type{t}': [names[c] for c in np.flatnonzero(g['cell_type'] == t)] for t in range(nct)}
    st.read_index = {f'read{i:07d}': i for i in range(N)}
    st.read_bcode_map = {f'read{i:07d}': names[rc[i]] for i in range(N)}
    st.read_umi_map = {f'read{i:07d}': f'U{g["row_umi"][i]:08d}' for i in range(N)}
    st.feat_index = {('__no_feature' if j == 0 else f'locus{j:05d}'): j for j in range(K)}
    st.filtlist = None
    st.reassignments = {}
    return st


def rel_err(a, b, floor=PI_FLOOR):
    """max |a-b|/|b| over entries with |b| > floor; entries below must be < floor-ish absolute."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    big = np.abs(b) > floor
    r = np.max(np.abs(a[big] - b[big]) / np.abs(b[big])) if big.any() else 0.0
    small_ok = np.all(np.abs(a[~big]) <= 1e-280) if (~big).any() else True
    return r, small_ok


def top2_margin(indptr, z):
    """Relative gap between the best and the second best z of every row (inf for rows with < 2)."""
    n = len(indptr) - 1
    out = np.full(n, np.inf)
    for i in range(n):
        v = np.sort(z[indptr[i]:indptr[i + 1]])[::-1]
        if len(v) >= 2 and v[0] > 0:
            out[i] = (v[0] - v[1]) / v[0]
    return out


def csr_equal(ptr_a, idx_a, dat_a, ptr_b, idx_b, dat_b, rtol=0):
    if not (np.array_equal(ptr_a, ptr_b) and np.array_equal(idx_a, idx_b)):
        return False
    if rtol == 0:
        return np.array_equal(np.asarray(dat_a, dtype=np.float64), np.asarray(dat_b, dtype=np.float64))
    return np.allclose(dat_a, dat_b, rtol=rtol, atol=0)


def small_synth(seed=0, **kw):
    args = dict(n_cells=20, n_alignments=6000, n_loci=300, seed=seed, unique_frac=0.1)
    args.update(kw)
    return synth.generate(**args)
