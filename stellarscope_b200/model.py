"""Host-side mirror of the reference's model API for the EM path.

Same names, argument meaning and error behaviour as
``stellarscope/utils/model.py`` (``TelescopeLikelihood`` :87-507,
``_fit_pooling_model`` :675-804, ``Stellarscope.reassign`` :1239-1256,
``Stellarscope.output_report`` :1291-1420) -- but every model of a pooling
mode lives in ONE device layout and is fitted by one call into
libstellarscope_b200.so.  Nothing here evaluates the model on the CPU; without
a CUDA device construction fails.
"""
from __future__ import annotations

import logging as lg
import sys
import time
from collections import OrderedDict
from typing import Optional

import numpy as np
import scipy.sparse

from .engine import Engine, ModelError, NotCanonicalError, to_host
from .sparse_plus import csr_matrix_plus
from .statistics import FitInfo, PoolInfo, ReassignInfo, SegmentFitInfos, counter_from_hist

TIE_TOL_DEFAULT = 1e-9


class StellarscopeError(Exception):
    """Stands in for ``stellarscope.StellarscopeError`` when the reference
    package is not importable; ``install()`` rebinds it to the real class."""


def _canonical_csr(m):
    m = scipy.sparse.csr_matrix(m)
    if not m.has_canonical_format:
        m = m.copy()
        m.sum_duplicates()
    if (m.data == 0).any():
        m = m.copy()
        m.eliminate_zeros()                      # model.py:114
    if not m.has_sorted_indices:
        m.sort_indices()
    return m


class TelescopeLikelihood(object):
    """Drop-in for the reference class of the same name.

    ``row_segment`` / ``n_segments`` are an extension used by
    ``_fit_pooling_model``: row i belongs to model ``row_segment[i]`` (-1: to
    none).  Without them the object is one model over the whole matrix, exactly
    like the reference's ``TelescopeLikelihood(score_matrix, opts)``.
    """

    REASSIGN_MODES = [          # model.py:92-100
        'best_exclude', 'best_conf', 'best_random', 'best_average',
        'initial_unique', 'initial_random', 'total_hits',
    ]

    def __init__(self, score_matrix, opts, row_segment=None, n_segments=None, device=None,
                 tie_tol=TIE_TOL_DEFAULT, shard=None, device_matrix=None, row_ids=None, group=None):
        self.opts = opts
        self.epsilon = opts.em_epsilon
        self.max_iter = opts.max_iter
        self.pi_prior = opts.pi_prior
        self.theta_prior = opts.theta_prior
        self.tie_tol = tie_tol
        # The reference canonicalises on construction (csr_matrix(...), eliminate_zeros, model.py:113-114).
        # Checking that on the host costs a pass over the matrix; ss_layout_build checks it on the device
        # in the pass it makes anyway and the host only canonicalises when that check fails.
        m = score_matrix if scipy.sparse.isspmatrix_csr(score_matrix) else scipy.sparse.csr_matrix(score_matrix)
        self.N, self.K = m.shape
        self.scale_factor = 100.
        # ``row_ids`` (sharded runs): ascending global ids of the rows THIS rank holds.  The object
        # keeps the reference's global shape (N, K, raw_scores); the device only ever sees the rank's rows --
        # they are cut out of the host matrix before the upload and ``row_segment`` is indexed by local row.
        self._row_ids = None if row_ids is None else np.ascontiguousarray(row_ids, dtype=np.int64)
        self._group = group
        if m.nnz > 2 ** 31 - 1:
            raise OverflowError(f'{m.nnz} alignments: the device layout uses int32 row pointers (at most 2^31 - 1)')
        self._engine = Engine(device) if shard is not None else Engine.acquire(device)
        if shard is not None:
            # ONE model whose rows are spread over ranks: (rank, world, process group[, peers])
            rank, world, group = shard[:3]
            self._engine.set_sharding(rank, world, group, peers=(shard[3] if len(shard) > 3 else True), K=self.K)
        for attempt in (0, 1):
            if m.dtype != np.uint16 and m.nnz and (m.data.max() > 65535 or m.data.min() < 0):
                raise StellarscopeError('alignment scores outside [0, 65535] are not supported (uint16, model.py:964)')
            if device_matrix is not None and attempt == 0:
                self._engine.set_matrix_device(*device_matrix, self.K)         # canonical copy of another model object
            elif self._row_ids is not None:
                self._engine.set_matrix_rows(m.indptr, m.indices, m.data, self._row_ids, self.K)
            else:
                self._engine.set_matrix(m.indptr, m.indices, m.data, self.K)    # H2D in flight ...
            if attempt == 0:
                if callable(row_segment):                                      # ... while the host walks the row sets
                    row_segment, self.segment_keys = row_segment()
                    n_segments = len(self.segment_keys)
                elif row_segment is None:
                    row_segment = np.zeros(self._engine.N, dtype=np.int32)
                    n_segments = 1
                self.row_segment = np.ascontiguousarray(row_segment, dtype=np.int32)
                self.n_segments = int(n_segments)
            try:
                self._engine.build_layout(self.row_segment, self.n_segments)
                break
            except NotCanonicalError:
                if attempt == 1:
                    raise
                m = _canonical_csr(m)
        self.raw_scores = m if isinstance(m, csr_matrix_plus) else csr_matrix_plus(m)
        self._max_score = None

        self._z = None
        self._z_dev = None
        self.pi = csr_matrix_plus(np.repeat(1. / self.K, self.K))     # model.py:164
        self.pi_init = None
        self.theta = np.repeat(1. / self.K, self.K)                   # model.py:172
        self.theta_init = None
        self.lnl = float('inf')
        self.fitinfo = FitInfo(epsilon=self.epsilon, max_iter=self.max_iter)
        self.segment_fitinfo = None
        self.fit_seconds = None

    def close(self):
        """Give the device resources back (the engine handle returns to a pool)."""
        eng, self._engine = getattr(self, '_engine', None), None
        self._steps_obj = None
        if eng is not None:
            eng.release()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- lazily materialised views the reference exposes as attributes ----------
    @property
    def max_score(self):
        """model.py:115 (the fit takes its per-model maxima on the device)."""
        if self._max_score is None:
            self._max_score = int(self.raw_scores.data.max()) if self.raw_scores.nnz else 0
        return self._max_score

    @property
    def Q(self):
        """model.py:129-130 for a single-model object (lazy, host; not used by the fit)."""
        if self.n_segments != 1:
            raise AttributeError('Q is per model; this object holds several models')
        q = self.raw_scores.astype(np.float64)
        if q.nnz:
            q.data = np.expm1((q.data * (1.0 / self.max_score)) * self.scale_factor)
        return csr_matrix_plus(q)

    def _y(self, pred):
        n = np.diff(self.raw_scores.indptr)
        return csr_matrix_plus(np.asmatrix(pred(n)).T, dtype=np.ubyte)

    @property
    def Y_amb(self):
        return self._y(lambda n: n > 1)           # model.py:141

    @property
    def Y_uni(self):
        return self._y(lambda n: n == 1)

    @property
    def Y_0(self):
        return self._y(lambda n: n == 0)

    @property
    def z(self):
        """N x K CSR float64 posterior (model.py:356; summed over models, :790).
        Copied from the device on first access."""
        if self._z is None:
            if self.segment_fitinfo is None:
                return None
            zd = to_host(self._z_device())
            if self._row_ids is not None:
                # this rank's rows only (the other ranks' rows are zero here, like the rows of no model)
                zd = self._scatter_entries(zd)
            m = csr_matrix_plus((zd, self.raw_scores.indices.copy(), self.raw_scores.indptr.copy()),
                                shape=self.raw_scores.shape)
            self._z = m
        return self._z

    @z.setter
    def z(self, value):
        self._z = value

    def _scatter_entries(self, local_vals):
        """Per-alignment values of the local rows -> array over all alignments of the global matrix (0 elsewhere)."""
        indptr = self.raw_scores.indptr
        rows = self._row_ids
        lens = (indptr[rows + 1] - indptr[rows]).astype(np.int64)
        lptr = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum(lens, out=lptr[1:])
        pos = np.repeat(indptr[rows].astype(np.int64) - lptr[:-1], lens) + np.arange(lptr[-1])
        out = np.zeros(self.raw_scores.nnz, dtype=local_vals.dtype)
        out[pos] = local_vals[:lptr[-1]]
        return out

    def _global_indptr(self, local_ptr, dtype):
        """Row pointers over all N rows from the row pointers over this rank's rows (the other rows are empty)."""
        if self._row_ids is None:
            return local_ptr.astype(dtype, copy=False)
        lens = np.zeros(self.N, dtype=np.int64)
        lens[self._row_ids] = np.diff(local_ptr)
        out = np.zeros(self.N + 1, dtype=np.int64)
        np.cumsum(lens, out=out[1:])
        return out.astype(dtype, copy=False)

    def _sharded(self):
        if self._row_ids is None:
            return False
        world, _ = _dist_info()
        return world > 1

    def _z_device(self):
        if self._z_dev is None:
            self._z_dev = self._engine.emit_z()
        return self._z_dev

    # -- fit -------------------------------------------------------------------
    def fit_segments(self):
        """Fit every model of the layout; returns the list of FitInfo in segment order."""
        import torch
        o = self.opts
        eng = self._engine
        t0 = time.perf_counter()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(torch.cuda.current_stream(eng.device))
        eng.fit(self.epsilon, self.max_iter, self.pi_prior, self.theta_prior,
                bool(getattr(o, 'use_likelihood', False)))
        end.record(torch.cuda.current_stream(eng.device))
        res = eng.results(traces=True)
        self.fit_seconds = start.elapsed_time(end) / 1e3
        self.fit_wall_seconds = time.perf_counter() - t0
        if (res['flags'] & 4).any():
            bad = int(np.flatnonzero(res['flags'] & 4)[0])
            # 0/0 in the M-step: FloatingPointError -> StellarscopeError, model.py:267-272
            raise StellarscopeError(('invalid value encountered in divide',
                                     f'model {bad} has no alignments or a zero denominator'))
        tot_it = max(1, int(res['iters'].sum()))
        itime = self.fit_seconds / tot_it
        infos = SegmentFitInfos(res, self.epsilon, self.max_iter, itime)
        self.segment_fitinfo = infos
        self._results = res
        self._z = None
        self._z_dev = None
        return infos

    def em(self, loglev=lg.DEBUG):
        """model.py:325-373 for a single-model object."""
        if self.n_segments != 1:
            raise StellarscopeError('em() is for a single model; use fit_segments()')
        # pi_init / theta_init: the parameters after the FIRST M-step (model.py:336-337) = the result of a
        # one-iteration fit; it costs one iteration of ~50 and keeps the attribute contract of the reference
        eng = self._engine
        eng.fit(self.epsilon, 1, self.pi_prior, self.theta_prior, False)
        pi1, theta1 = eng.dense_params(0)
        self.pi_init, self.theta_init = csr_matrix_plus(pi1), theta1
        fi = self.fit_segments()[0]
        self.fitinfo = fi
        pi, theta = self._engine.dense_params(0)
        self.pi = csr_matrix_plus(pi)
        self.theta = theta
        self.lnl = fi.final_lnl
        for inum, _t, diff in fi.iterations:
            lg.log(loglev, 'Iteration {:d}, diff={:.5g}'.format(inum, diff))
        if fi.converged:
            lg.log(loglev, f'EM converged after {len(fi.iterations):d} iterations.')
        elif fi.reached_max:
            lg.log(loglev, f'EM terminated after {len(fi.iterations):d} iterations.')
        lg.log(loglev, f'Final log-likelihood: {self.lnl:f}.')
        return

    # -- single steps (model.py:202-323): for callers that drive the iteration themselves; em() does not use them --
    def _steps(self):
        if getattr(self, '_steps_obj', None) is None:
            from .steps import ModelSteps
            self._steps_obj = ModelSteps(self)
        return self._steps_obj

    def estep(self, pi, theta):
        """E(z[i,j]) for the given parameters: N x K CSR (model.py:202-226)."""
        return self._steps().estep(pi, theta)

    def mstep(self, z):
        """MAP estimates ``(pi 1 x K sparse, theta K array)`` from z (model.py:229-272)."""
        try:
            return self._steps().mstep(z)
        except FloatingPointError as e:
            raise StellarscopeError(e.args)                    # model.py:267-272

    def calculate_lnl(self, z, pi, theta):
        """sum_ij z_ij log(Q_ij pi_j theta_j^Y_i) (model.py:276-323)."""
        return self._steps().calculate_lnl(z, pi, theta)

    def segment_params(self, seg):
        """(pi, theta) dense K-vectors of one model of a pooled fit."""
        return self._engine.dense_params(seg)

    # -- reassign: model.py:375-507 --------------------------------------------
    def reassign(self, mode: str, thresh: Optional[float] = None):
        if mode not in self.REASSIGN_MODES:
            msg = f'Argument "method" should be one of {self.REASSIGN_MODES}'
            raise ValueError(msg)
        if mode == 'best_conf' and (thresh is None or thresh <= 0.5):
            raise StellarscopeError(f"Confidence threshold ({thresh}) must be > 0.5")
        needs_z = mode in ('best_exclude', 'best_conf', 'best_random', 'best_average')
        z = None
        if needs_z:
            if self.segment_fitinfo is None:
                raise StellarscopeError('reassign_mat was not set.')     # model.py:504-505 (no z)
            z = self._z_device()
        flag, nbest, info = self._engine.reassign(mode, z, thresh if thresh is not None else 0.0, self.tie_tol)
        if self._sharded():
            # every row lives on exactly one rank (parallel.plan_local_rows): the counters and the tie histogram of
            # the whole matrix are the sums over the ranks
            from . import parallel
            info = parallel.allreduce_sum_host(np.asarray(info, dtype=np.int64), self._group)
        rinfo = ReassignInfo(mode)
        rinfo.assigned, rinfo.ambiguous, rinfo.unaligned = int(info[0]), int(info[1]), int(info[2])
        if mode in ('best_exclude', 'best_random', 'best_average', 'initial_random'):
            rinfo.ambigous_dist = counter_from_hist(info[4:])
        indptr_dtype = self.raw_scores.indptr.dtype
        if mode not in ('best_random', 'initial_random'):
            # the selection comes back as CSR straight from the device (no pass over all alignments on the host)
            ptr, idx, frac = self._engine.reassign_csr(flag, nbest, fractions=(mode == 'best_average'))
            data = frac if mode == 'best_average' else np.ones(len(idx), dtype=np.int8)
            mat = csr_matrix_plus((data, idx, self._global_indptr(ptr, indptr_dtype)), shape=self.raw_scores.shape)
            # lets count_by_cell aggregate without uploading the matrix again (while the engine still holds this
            # matrix); best_average also needs the tie-set sizes (its entries are 1 / nbest)
            mat._ss_dev = (self._engine, flag, self._engine._keep.get('row_ptr'), self._row_ids) + \
                ((nbest,) if mode == 'best_average' else ())
            return mat, rinfo
        # random modes: the tie sets come from the device, the draws from the reference's numpy generator in row
        # order (one vectorised rng.integers call gives the stream of the reference's per-row rng.choice calls,
        # tests/test_oracle.py), the picked alignment is kept on the device and the matrix compacted there
        tied_rows, nb_tied = self._engine.tied_rows(nbest)
        if self._sharded():
            # one stream of draws for the whole matrix, in global row order, on every rank (the reference draws
            # row by row, sparse_plus.py:233-238): gather the tie sets, draw them all, keep the own rows' picks
            from . import parallel
            g = parallel.gather_concat({'row': self._row_ids[tied_rows], 'n': nb_tied.astype(np.int64)}, self._group)
            order = np.argsort(g['row'], kind='stable')
            picks_all = self.opts.rng.integers(0, g['n'][order]) if len(order) else np.zeros(0, dtype=np.int64)
            picks = picks_all[np.searchsorted(g['row'][order], self._row_ids[tied_rows])]
        else:
            picks = self.opts.rng.integers(0, nb_tied) if len(tied_rows) else np.zeros(0, dtype=np.int64)
        self._engine.reassign_pick(flag, tied_rows, picks)
        ptr, idx, _ = self._engine.reassign_csr(flag)
        mat = csr_matrix_plus((np.ones(len(idx), dtype=np.int8), idx, self._global_indptr(ptr, indptr_dtype)),
                              shape=self.raw_scores.shape)
        mat._ss_dev = (self._engine, flag, self._engine._keep.get('row_ptr'), self._row_ids)
        return mat, rinfo


def _apply_random_choice(flag, row, nbest, tied_rows, picks):
    """``choose_random(1, rng)`` (sparse_plus.py:231-240) on the tie sets the device
    found: for every row with more than one tied-best alignment, in row order,
    the reference draws ``rng.choice(range(start, end))``; numpy produces the
    same stream for one vectorised ``rng.integers(0, n)`` call (checked in
    tests/test_oracle.py), so the draws (``picks``, one per row of ``tied_rows``) are made on the host and
    applied as a mask."""
    if len(tied_rows) == 0:
        return flag
    pos = np.flatnonzero(flag)                 # tied-best alignments, CSR order
    prow = row[pos]
    start = np.zeros(len(nbest) + 1, dtype=np.int64)
    np.cumsum(nbest, out=start[1:])            # start of each row inside `pos`
    in_tied = nbest[prow] > 1
    out = flag.copy()
    out[pos[in_tied]] = False
    out[pos[start[tied_rows] + picks]] = True
    return out


def _replay_random_choice(flag, row, nbest, rng):
    """Single-process form: draws and applies (kept for the tests of the draw order)."""
    tied_rows = np.flatnonzero(nbest > 1)
    if len(tied_rows) == 0:
        return flag
    return _apply_random_choice(flag, row, nbest, tied_rows, rng.integers(0, nbest[tied_rows]))


def _em_wrapper(tl: TelescopeLikelihood):
    tl.em()
    return tl.z, tl.fitinfo


def pooling_row_segments(st, opts, collective=False, info=None):
    """Row -> model map of a pooling mode (model.py:701-743).  Returns
    ``(row_segment int32[N], keys)`` where ``keys[i]`` is the key of model i in
    ``PoolInfo.models_info`` (model.py:766, 791).  The reference's containers
    (``bcode_ridx_map``: barcode -> set of row ids) are walked by the C helper
    ``csrc/ss_hostutil.c``.  ``collective``: all ranks of the default process group make this call together (passes
    over per-row arrays may then be done in 1/world slices and combined).  ``info`` (a dict) receives
    ``all_assigned`` when the call knows without another pass whether every row belongs to a model."""
    from . import hostutil
    N = st.shape[0]
    if opts.pooling_mode == 'pseudobulk':
        return np.zeros(N, dtype=np.int32), ['pseudobulk']
    if opts.pooling_mode in ('individual', 'celltype'):
        from . import checkpoint
        table = checkpoint.row_table(st)
        if table is not None:
            return _row_segments_from_table(st, opts, table, collective, info)
    seg = np.full(N, -1, dtype=np.int32)
    groups, ids = [], []
    i = 0
    if opts.pooling_mode == 'individual':
        groups = list(st.bcode_ridx_map.values())
        ids = np.empty(len(groups), dtype=np.int32)
        i = hostutil.number_groups(groups, ids)               # barcodes without reads get no model (model.py:735-739)
        if i < len(groups):
            for _bcode, rowset in st.bcode_ridx_map.items():
                if not rowset:
                    lg.info(f'        ...not fitting "{_bcode}" -> no reads found')
        what = 'a read belongs to more than one barcode'
    elif opts.pooling_mode == 'celltype':
        for _ctype in st.celltypes:
            n0 = len(groups)
            for _bcode in st.ctype_bcode_map[_ctype]:
                rowset = st.bcode_ridx_map.get(_bcode)
                if rowset:
                    groups.append(rowset)
            if len(groups) == n0:
                lg.info(f'        ...not fitting "{_ctype}" -> no reads found')
                continue
            ids.extend([i] * (len(groups) - n0))
            i += 1
        ids = np.asarray(ids, dtype=np.int32)
        what = 'a barcode belongs to more than one celltype (not supported)'
    else:
        raise StellarscopeError(f'unknown pooling mode {opts.pooling_mode}')
    try:
        hostutil.fill_row_segments(groups, ids, seg)
    except ValueError as e:
        raise StellarscopeError(f'{what}: {e}')
    return seg, list(range(i))


def _row_segments_from_table(st, opts, table, collective=False, info=None):
    """``pooling_row_segments`` for an object loaded from a side-car checkpoint (``checkpoint.RowTable``): the
    barcode code of every row is already an array, no container is walked."""
    code = table.row_bcode
    nb = len(table.bcodes)
    from . import hostutil
    world, rank = _dist_info() if collective else (1, 0)
    if world > 1:
        # every rank looks at 1/world of the rows; the sums (and the "a row without a barcode exists" flag) are combined
        from . import parallel
        n = len(code)
        part = code[n * rank // world: n * (rank + 1) // world]
        loc = np.concatenate([hostutil.group_sizes(part, nb), [int(len(part) > 0 and int(part.min()) < 0)]])
        tot = parallel.allreduce_sum_host(loc.astype(np.int64))
        counts, all_coded = tot[:nb], int(tot[nb]) == 0
    else:
        all_coded = len(code) == 0 or int(code.min()) >= 0        # every row belongs to a barcode (the usual case)
        counts = hostutil.group_sizes(code, nb)
    if opts.pooling_mode == 'individual':
        has = counts > 0
        if all_coded and has.all():
            if info is not None:
                info['all_assigned'] = True
            return code, list(range(nb))                          # model of a row = its barcode code: no pass at all
        model_of = np.where(has, np.cumsum(has) - 1, -1).astype(np.int32)   # barcodes without reads get no model
        for b in np.flatnonzero(~has):
            lg.info(f'        ...not fitting "{table.bcodes[b]}" -> no reads found')
        n_models = int(has.sum())
    else:
        code_of = {b: i for i, b in enumerate(table.bcodes)}
        model_of = np.full(nb, -1, dtype=np.int32)
        n_models = 0
        for _ctype in st.celltypes:
            found = False
            for _bcode in st.ctype_bcode_map[_ctype]:
                c = code_of.get(_bcode)
                if c is None or counts[c] == 0:
                    continue
                if model_of[c] not in (-1, n_models):
                    raise StellarscopeError('a barcode belongs to more than one celltype (not supported): '
                                            f'{_bcode}')
                model_of[c] = n_models
                found = True
            if not found:
                lg.info(f'        ...not fitting "{_ctype}" -> no reads found')
                continue
            n_models += 1
    if all_coded:
        seg = model_of[code]
    else:
        seg = np.where(code >= 0, model_of[np.maximum(code, 0)], -1).astype(np.int32)
    if info is not None:                                          # rows of a barcode without a model are loose
        info['all_assigned'] = bool(all_coded and not np.any((counts > 0) & (model_of < 0)))
    return seg, list(range(n_models))


def _row_cells(st):
    """Cell (running index of the barcode in ``bcode_ridx_map``) of every row, -1 for rows of no barcode."""
    from . import checkpoint, hostutil
    table = checkpoint.row_table(st)
    if table is not None:
        return table.row_bcode.copy()
    groups = list(st.bcode_ridx_map.values())
    cell = np.full(st.shape[0], -1, dtype=np.int32)
    hostutil.fill_row_segments(groups, np.arange(len(groups), dtype=np.int32), cell)
    return cell


def _fit_pooling_model(st, opts, processes: int = 1, progress: int = 100):
    """Fit model using different pooling modes (model.py:675-804).

    ``processes`` / ``opts.nproc`` are ignored for the fit: all models run
    concurrently on the GPU.  Returns ``(TelescopeLikelihood, PoolInfo)``."""
    poolinfo = PoolInfo(opts.pooling_mode)
    if opts.ignore_umi:                                     # model.py:749-758
        if getattr(st, 'corrected', None) is not None:
            lg.warning('Ignoring UMI corrected matrix')
        fullmat = st.raw_scores
    else:
        if getattr(st, 'corrected', None) is None:
            raise StellarscopeError("UMI corrected matrix not found")
        if st.corrected.shape != st.shape:
            raise StellarscopeError("UMI corrected matrix shape mismatch")
        fullmat = st.corrected

    if opts.pooling_mode == 'pseudobulk':
        lg.info('Models to fit: 1')
        world, rank = _dist_info()
        if world > 1 and getattr(opts, 'shard_models', True):
            # one model: shard its ROWS over the ranks (blocks of equal alignment count); the theta sums
            # are exchanged once per iteration.  Rows of other ranks belong to no local
            # model, so z -- and the counts derived from it -- are rank-local and add up over ranks.
            from . import parallel
            # no canonicalising pass over the whole matrix on the host: the device checks the rank's rows while it
            # builds the layout and only a failed check makes this rank canonicalise (TelescopeLikelihood.__init__)
            m = fullmat if scipy.sparse.isspmatrix_csr(fullmat) else scipy.sparse.csr_matrix(fullmat)
            if getattr(opts, 'umi_counts', False):
                # distinct UMIs are counted per cell: all reads of a cell on one rank
                rows = parallel.model_cell_block(np.diff(m.indptr), np.arange(m.shape[0]), _row_cells(st), rank, world)
            else:
                r0, r1 = parallel.row_block(m.indptr, rank, world)
                rows = np.arange(r0, r1, dtype=np.int64)
            # the rank uploads ITS block of rows only
            ret_model = TelescopeLikelihood(m, opts, row_segment=np.zeros(len(rows), dtype=np.int32), n_segments=1,
                                            row_ids=rows,
                                            shard=(rank, world, None, getattr(opts, 'peer_exchange', True)))
        else:
            ret_model = TelescopeLikelihood(fullmat, opts)
        poolinfo.nmodels = 1
        ret_model.em()
        poolinfo.models_info['pseudobulk'] = ret_model.fitinfo
        return ret_model, poolinfo

    poolinfo.nmodels = (len(st.bcode_ridx_map) if opts.pooling_mode == 'individual'
                        else len(st.celltypes))             # model.py:772, 778
    lg.info(f'Models to fit: {poolinfo.nmodels}')
    world, rank = _dist_info()
    if not getattr(opts, 'shard_models', True):
        world = 1                    # every rank fits its own, independent data set
    if world == 1:
        # the row -> model map is computed while the matrix is on its way to the device
        ret_model = TelescopeLikelihood(fullmat, opts, row_segment=lambda: pooling_row_segments(st, opts))
        keys = ret_model.segment_keys
        infos = ret_model.fit_segments().keyed(keys)
    else:
        # models are independent: shard them over the ranks (one process per GPU), no collective on
        # the data path; every rank keeps the global row numbering, rows of other ranks' models are
        # simply not fitted here
        from . import parallel
        _t = [time.perf_counter()]
        seg_info = {}
        row_segment, keys = pooling_row_segments(st, opts, collective=True, info=seg_info)
        _t.append(time.perf_counter())
        m = fullmat if scipy.sparse.isspmatrix_csr(fullmat) else scipy.sparse.csr_matrix(fullmat)
        # a model much larger than a rank's fair share (a dominant cell type on 8 GPUs) cannot be balanced by packing
        # whole models: such WIDE models are fitted by all ranks together, rows cut into blocks like pseudobulk
        # opts.wide_models: None (default: pack whole models), 'auto' (parallel.wide_models) or a list of model indices;
        # opts.wide_model_cost: lockstep overhead of one wide model in alignments of single-GPU work
        umi_counts = bool(getattr(opts, 'umi_counts', False))
        # (a pass over 10^8 rows on every rank when the row -> model map did not already tell)
        loose = (not seg_info['all_assigned']) if 'all_assigned' in seg_info else \
            (len(row_segment) > 0 and int(row_segment.min()) < 0)
        row_cell = _row_cells(st) if (loose or (umi_counts and getattr(opts, 'wide_models', None))) else None
        plan = parallel.plan_local_rows(m.indptr, row_segment, len(keys), world, rank,
                                        wide=getattr(opts, 'wide_models', None) or [],
                                        wide_cost=getattr(opts, 'wide_model_cost', 2_000_000),
                                        row_cell=row_cell, cells_whole=umi_counts, all_assigned=not loose,
                                        collective=True)
        mine, wide, rows = plan['mine'], plan['wide'], plan['rows']
        _t.append(time.perf_counter())
        # every rank cuts ITS rows out of the host matrix and uploads only them (1/world of the H2D traffic)
        ret_model = TelescopeLikelihood(m, opts, row_segment=plan['local_seg'], n_segments=len(mine), row_ids=rows)
        ret_model.model_ids = mine
        _t.append(time.perf_counter())
        local = ret_model.fit_segments()
        _t.append(time.perf_counter())
        # per-model records of all ranks: fixed-width arrays through all_gather_into_tensor (no pickling; the
        # per-iteration traces stay on the rank that fitted the model unless opts.gather_traces is set)
        lr = local._res
        rec = np.stack([np.asarray(mine, dtype=np.float64)] +
                       [np.asarray(lr[k], dtype=np.float64) for k in ('iters', 'flags', 'nobs', 'nfeats', 'lnl')], axis=1) \
            if len(mine) else np.zeros((0, 6))
        payload = {'rec': rec}
        want_traces = bool(getattr(opts, 'gather_traces', False))
        if want_traces:
            payload['diffs'] = (np.asarray(lr['diffs']).reshape(len(mine), -1) if len(mine)
                                else np.zeros((0, ret_model.max_iter)))
        g = parallel.gather_concat(payload)
        ids = g['rec'][:, 0].astype(np.int64)
        res = dict(iters=g['rec'][:, 1].astype(np.int32), flags=g['rec'][:, 2].astype(np.int32),
                   nobs=g['rec'][:, 3].astype(np.int32), nfeats=g['rec'][:, 4].astype(np.int32), lnl=g['rec'][:, 5])
        my0 = int(g['counts'][:rank].sum())
        res['diffs'] = g['diffs'] if want_traces else _ShardedTrace(lr.get('diffs'), my0, len(mine), len(ids),
                                                                   ret_model.max_iter)
        itime = float(parallel.allreduce_sum_host(np.array([local._itime]))[0]) / world
        _t.append(time.perf_counter())
        # where this rank's wall time went (seconds): row -> model map, packing plan + row selection, cutting the rows
        # out + upload + layout build, fit (incl. the wait for the device), gather of the records
        ret_model.host_timings = dict(zip(('row_segments', 'plan', 'cut_upload_build', 'fit', 'gather_records'),
                                          np.diff(_t).tolist()))
        parts_ids, parts_res = [ids], [res]
        if len(wide):
            eng = ret_model._engine
            dev_matrix = tuple(eng._keep[k] for k in ('row_ptr', 'col_idx', 'score'))
            z_all = ret_model._z_device()
            for w, mask in zip(wide, plan['wide_local']):
                seg_w = np.where(mask, 0, -1).astype(np.int32)
                tw = TelescopeLikelihood(ret_model.raw_scores, opts, row_segment=seg_w, n_segments=1, device_matrix=dev_matrix,
                                         row_ids=rows, shard=(rank, world, None, getattr(opts, 'peer_exchange', True)))
                info_w = tw.fit_segments()                  # the same record on every rank
                z_all += tw._z_device()                     # disjoint rows: the rank-local z of the wide model
                rw = dict(info_w._res)
                rw['diffs'] = np.asarray(rw['diffs']).reshape(1, -1)
                parts_ids.append(np.array([w], dtype=np.int64))
                parts_res.append(rw)
                itime = max(itime, info_w._itime)
                ret_model.fit_seconds_wide = getattr(ret_model, 'fit_seconds_wide', 0.0) + (tw.fit_seconds or 0.0)
                tw._engine.sync_ranks()
                tw.close()
            ret_model.wide_model_ids = wide
        # concatenate the per-rank result arrays; order[i] = row of model i in the concatenation
        ids = np.concatenate(parts_ids)
        fields = ('iters', 'flags', 'lnl', 'nobs', 'nfeats')
        res = {k: np.concatenate([np.asarray(p[k]) for p in parts_res]) for k in fields}
        res['diffs'] = parts_res[0]['diffs'] if len(parts_res) == 1 else _StackedTrace([p['diffs'] for p in parts_res])
        order = np.empty(len(keys), dtype=np.int64)
        order[ids] = np.arange(len(ids))
        infos = SegmentFitInfos(res, ret_model.epsilon, ret_model.max_iter, itime, keys=keys, order=order)
    poolinfo.models_info = infos
    lg.info(f'        ...{len(infos)} models fitted')
    ret_model.lnl = poolinfo.total_lnl                      # model.py:803
    return ret_model, poolinfo


class _ShardedTrace:
    """Row j of the gathered per-model traces: the |delta pi| trace of a model fitted on THIS rank (rows
    ``my0 .. my0 + n_mine`` of the gathered order, read lazily from the device), NaN for the models of other
    ranks (their traces stay where they were computed; ``opts.gather_traces = True`` ships them)."""

    def __init__(self, local, my0, n_mine, n_all, max_iter):
        self._local, self._my0, self._n, self._all, self._mi = local, my0, n_mine, n_all, max_iter

    def __len__(self):
        return self._all

    def __getitem__(self, j):
        j = int(j)
        if self._my0 <= j < self._my0 + self._n and self._local is not None:
            return self._local[j - self._my0]
        return np.full(self._mi, np.nan)


class _StackedTrace:
    """Row-wise concatenation of trace blocks (each indexable by row)."""

    def __init__(self, blocks):
        self._blocks = blocks
        self._starts = np.cumsum([0] + [len(b) for b in blocks])

    def __len__(self):
        return int(self._starts[-1])

    def __getitem__(self, j):
        b = int(np.searchsorted(self._starts, int(j), side='right')) - 1
        return self._blocks[b][int(j) - int(self._starts[b])]


def _dist_info():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
    except Exception:
        pass
    return 1, 0


# ---------------------------------------------------------------------------
# Stellarscope.reassign / output_report  (model.py:1239-1256, 1291-1420)
# ---------------------------------------------------------------------------
def reassign(st, tl: TelescopeLikelihood):
    """``Stellarscope.reassign``: every mode of ``opts.reassign_mode`` in order
    (the modes share ``opts.rng``)."""
    if not hasattr(st, 'reassignments'):
        st.reassignments = {}
    rmode_info = {}
    for _rmode in st.opts.reassign_mode:
        _thresh = st.opts.conf_prob if _rmode == 'best_conf' else None
        _rmat, _rinfo = tl.reassign(_rmode, _thresh)
        st.reassignments[_rmode] = _rmat
        rmode_info[_rmode] = _rinfo
    return rmode_info


def _row_outputs(st, bc_pos):
    """Output column (barcode position, -1: dropped) and -- with ``--umi_counts`` -- UMI id of every row, from the
    reference's per-read dicts (model.py:1389-1394: rows in ``read_index`` order).  The dicts are walked by the C
    helper; other mapping types take the Python route."""
    N = len(st.read_index)
    want_umi = bool(st.opts.umi_counts)
    from . import checkpoint
    table = checkpoint.row_table(st)
    if table is not None and (not want_umi or table.row_umi is not None):
        # loaded from a side-car checkpoint: barcode / UMI codes per row are arrays already
        pos = np.fromiter((bc_pos.get(b, -1) for b in table.bcodes), dtype=np.int32, count=len(table.bcodes))
        code = table.row_bcode
        if len(code) and int(code.min()) >= 0 and np.array_equal(pos, np.arange(len(pos), dtype=np.int32)):
            row_cell = code                       # barcodes are in output order already and every row has one
        elif len(code) and int(code.min()) >= 0:
            row_cell = pos[code]
        else:
            row_cell = np.where(code >= 0, pos[np.maximum(code, 0)] if len(pos) else -1, -1).astype(np.int32)
        return row_cell, (table.row_umi.copy() if want_umi else None)
    if type(st.read_index) is dict and type(st.read_bcode_map) is dict and (not want_umi or type(st.read_umi_map) is dict):
        from . import hostutil
        row_cell = None
        groups = getattr(st, 'bcode_ridx_map', None)
        if type(groups) is dict and groups:
            # barcode -> row set is filled together with read -> barcode (store_read_info, model.py:1044-1053): walking
            # the ~10^4 sets is 100 x cheaper than one dict look-up per read; used when the sets cover every row once
            cand = np.full(N, -1, dtype=np.int32)            # the helper's own "unclaimed" value (threaded path)
            ids = np.fromiter((bc_pos.get(b, -1) for b in groups), dtype=np.int32, count=len(groups))
            try:
                # number of rows the sets claimed: equal to N only when they cover every row (each exactly once --
                # a row claimed by two barcodes raises); otherwise read_bcode_map decides, like the reference
                if hostutil.fill_row_segments(list(groups.values()), ids, cand) == N:
                    row_cell = cand
            except (ValueError, IndexError):
                row_cell = None
        if row_cell is None:
            row_cell = np.full(N, -1, dtype=np.int32)
            hostutil.row_labels(st.read_index, st.read_bcode_map, bc_pos, row_cell)
        row_umi = None
        if want_umi:
            row_umi = np.full(N, -1, dtype=np.int32)
            hostutil.row_labels(st.read_index, st.read_umi_map, None, row_umi, {})
        return row_cell, row_umi
    qnames = sorted(st.read_index, key=st.read_index.get)           # model.py:1389
    row_cell = np.fromiter((bc_pos.get(st.read_bcode_map[q], -1) for q in qnames), dtype=np.int32,
                           count=len(qnames))
    row_umi = None
    if want_umi:
        umi_ids = {}
        row_umi = np.fromiter((umi_ids.setdefault(st.read_umi_map[q], len(umi_ids)) for q in qnames),
                              dtype=np.int32, count=len(qnames))
    return row_cell, row_umi


def count_matrices(st, tl: TelescopeLikelihood):
    """The aggregation half of ``output_report`` (agg_bc / agg_bc_umi,
    model.py:1302-1345): ``{mode: (numft x numbc) scipy sparse matrix}`` plus the barcode
    and feature lists in output order (:1371-1378)."""
    if getattr(st, 'filtlist', None):
        bc_list = sorted(st.filtlist, key=st.filtlist.get)
    else:
        from . import checkpoint
        table = checkpoint.row_table(st)
        bc_list = sorted(table.bcodes) if table is not None else sorted(st.bcode_ridx_map.keys())
    ft_list = sorted(st.feat_index, key=st.feat_index.get)
    numft, numbc = len(ft_list), len(bc_list)
    bc_pos = {b: i for i, b in enumerate(bc_list)}
    row_cell, row_umi = _row_outputs(st, bc_pos)
    out = OrderedDict()
    world, rank = _dist_info()
    if not getattr(st.opts, 'shard_models', True):
        world = 1
    dev_cache = {}                       # the row -> output column map goes to the device once for all modes
    for _rmode in st.opts.reassign_mode:
        remat = st.reassignments[_rmode]
        cm = remat.count_by_cell(row_cell, numbc, row_umi=row_umi, engine=tl._engine, cache=dev_cache)
        if world > 1:
            cm.__dict__.pop('_ss_dev_coo', None)        # device triples of this rank's part only
            cm.__dict__.pop('_ss_dev_csc', None)
            # every rank holds the counts of its own rows (a partition of all rows, parallel.plan_local_rows): the
            # COO triples of all ranks are concatenated -- tensors through all_gather_into_tensor, no pickled
            # matrices -- and entries of the same (locus, cell) add up (a cell cut across ranks: pseudobulk blocks)
            from . import parallel
            cm = _merge_rank_counts(cm, world, parallel)
        out[_rmode] = cm
        if out[_rmode].shape != (numft, numbc):
            raise StellarscopeError(f'Incompatible shapes: {out[_rmode].shape} != {(numft, numbc)}')
    return out, bc_list, ft_list


def _merge_rank_counts(cm, world, parallel):
    """The count matrix of the whole run from the ranks' parts (each rank counted its own rows).  The parts travel as
    arrays (``parallel.gather_concat``).  Usually no cell is cut across ranks (models are packed whole; wide /
    pseudobulk models with ``--umi_counts`` keep cells whole): the columns of the result are then the ranks' columns
    side by side and are put in place with one scatter per rank -- no sort, no sum.  Otherwise (a cell's rows on
    several ranks) the triples are concatenated and entries of the same (locus, cell) add up."""
    cm = scipy.sparse.csc_matrix(cm)
    n_cells = cm.shape[1]
    lens = np.diff(cm.indptr).astype(np.int64)
    L = parallel.gather_concat({'lens': lens})['lens'].reshape(world, n_cells)
    g = parallel.gather_concat({'c': cm.indices.astype(np.int32), 'v': np.ascontiguousarray(cm.data)})
    if ((L > 0).sum(axis=0) > 1).any():
        cells = np.concatenate([np.repeat(np.arange(n_cells, dtype=np.int32), L[r]) for r in range(world)])
        merged = scipy.sparse.coo_matrix((g['v'], (g['c'], cells)), shape=cm.shape).tocsc()    # sums duplicates
        return merged.astype(cm.dtype)
    tot = L.sum(axis=0)
    indptr = np.zeros(n_cells + 1, dtype=np.int64)
    np.cumsum(tot, out=indptr[1:])
    out_idx = np.empty(int(indptr[-1]), dtype=np.int32)
    out_val = np.empty(int(indptr[-1]), dtype=cm.dtype)
    off = 0
    for r in range(world):
        lr = L[r]
        nr = int(lr.sum())
        if nr:
            src0 = np.cumsum(lr) - lr                                   # first entry of the cell in rank r's block
            dest = np.repeat(indptr[:-1] - src0, lr) + np.arange(nr, dtype=np.int64)
            out_idx[dest] = g['c'][off:off + nr]
            out_val[dest] = g['v'][off:off + nr]
        off += nr
    return scipy.sparse.csc_matrix((out_val, out_idx, indptr), shape=cm.shape)


def output_report(st, tl: TelescopeLikelihood):
    """``Stellarscope.output_report``: barcodes.tsv, features.tsv, TE_counts[.mode].mtx."""
    counts, bc_list, ft_list = count_matrices(st, tl)
    world, rank = _dist_info()
    if world > 1 and getattr(st.opts, 'shard_models', True):
        # the ranks of a sharded run hold the same (summed) matrices and the same output paths: one of them writes
        import torch.distributed as dist
        try:
            if rank == 0:
                _write_report(st, tl, counts, bc_list, ft_list)
        finally:
            dist.barrier()                       # the files exist when any rank returns
        return
    _write_report(st, tl, counts, bc_list, ft_list)
    return


def _write_report(st, tl, counts, bc_list, ft_list):
    """barcodes.tsv, features.tsv (model.py:1380-1384) and one MTX file per reassignment mode (:1411-1418)."""
    from . import mtx
    with open(st.opts.outfile_path('barcodes.tsv'), 'w') as outh:
        print('\n'.join(bc_list), file=outh)
    with open(st.opts.outfile_path('features.tsv'), 'w') as outh:
        print('\n'.join(ft_list), file=outh)
    for rmode_i, (_rmode, _counts) in enumerate(counts.items()):
        if rmode_i == 0:
            out_mtx = st.opts.outfile_path('TE_counts.mtx')
        else:
            out_mtx = st.opts.outfile_path(f'TE_counts.{_rmode}.mtx')
        mtx.mmwrite(out_mtx, _counts, comment=_mtx_meta(st, _rmode), engine=tl._engine,
                    device_coo=getattr(_counts, '_ss_dev_coo', None))                      # model.py:1418
    return


def _mtx_meta(st, reassign_mode):
    """MTX comment block (model.py:1347-1368)."""
    o = st.opts
    meta = OrderedDict({'PN': 'stellarscope', 'VN': getattr(o, 'version', '')})
    for k in ('samfile', 'checkpoint', 'gtffile', 'filtered_bc'):
        if hasattr(o, k):
            meta[k] = getattr(o, k)
    meta['pooling_mode'] = o.pooling_mode
    meta['reassign_mode'] = reassign_mode
    if o.ignore_umi:
        meta['ignore_umi'] = 'Yes'
        meta['umi_counts'] = 'Yes' if o.umi_counts else 'No'
    meta['command'] = ' '.join(sys.argv)
    return '\n '.join(f'{k}: {v}' for k, v in meta.items())


def install():
    """Make ``stellarscope assign`` / ``resume`` use this implementation:
    rebinds the hot-path names of the reference's ``stellarscope.utils.model``
    (must be importable) -- see INTEGRATION.md."""
    global StellarscopeError, PoolInfo, ReassignInfo
    import stellarscope
    import stellarscope.utils.model as ref
    import stellarscope.utils.statistics as ref_stats
    from . import statistics
    StellarscopeError = stellarscope.StellarscopeError
    # the result records are the reference's own classes from here on (PoolInfo only adds vectorised totals)
    PoolInfo, ReassignInfo = statistics.bind_reference_records(ref_stats)
    from . import sparse_plus
    sparse_plus.StellarscopeError = StellarscopeError
    ref.TelescopeLikelihood = TelescopeLikelihood
    ref._fit_pooling_model = _fit_pooling_model
    ref._em_wrapper = _em_wrapper
    ref.Stellarscope.reassign = lambda self, tl: reassign(self, tl)
    ref.Stellarscope.output_report = lambda self, tl: output_report(self, tl)
    from . import umi
    ref.Stellarscope.dedup_umi = lambda self, output_report=True, summary=True: umi.dedup_umi(self, output_report, summary)
    from . import loader
    ref.Stellarscope.load_alignment = lambda self, annotation: loader.load_alignment(self, annotation)
    from . import checkpoint
    ref.Stellarscope.save = lambda self, filename: checkpoint.save(self, filename)
    ref.Stellarscope.load = classmethod(lambda cls, filename: checkpoint.load(filename))
    return ref
