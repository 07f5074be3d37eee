"""CPU restatement (numpy) of Stellarscope's EM read-reassignment path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``stellarscope_b200/`` may import
this module; it is the checker for ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

What it restates (all citations relative to the reference checkout,
``stellarscope/utils/``):

* ``TelescopeLikelihood.__init__``      model.py:102-200
* ``estep`` / ``norm(1)``               model.py:202-226, sparse_plus.py:31-64
* ``mstep``                             model.py:229-272
* ``calculate_lnl``                     model.py:276-323
* ``em``                                model.py:325-373
* ``reassign`` (7 modes)                model.py:375-507,
  ``binmax(1)`` sparse_plus.py:134-178, ``choose_random`` :199-240
* ``_fit_pooling_model``                model.py:675-804
* ``output_report.agg_bc/agg_bc_umi``   model.py:1302-1345,
  ``groupby_sum`` sparse_plus.py:367-396
* ``FitInfo`` / ``PoolInfo`` numbers    statistics.py:269-393

The arithmetic of the reference executes in scipy.sparse ``_sparsetools`` and
numpy (third party, unpinned: ``scipy>=1.2.1``, ``numpy>=1.16.3`` in the
reference's setup.py:48-49; 1.18.1 / 2.3.5 in this image).  This file works on
raw CSR arrays with numpy only and keeps the reference's *operation order*
where it is observable (reciprocal-then-multiply in ``scale`` and ``norm``,
row-major accumulation order of column sums).

Extended precision: the CLI sets ``np.seterr(all='raise')``
(``__main__.py:30``) and on an underflow the reference recomputes in
``np.float128`` and stays promoted (model.py:223-226, sparse_plus.py:282-296).
``promote=True`` (default) restates that: work in float64 with underflow
trapped, switch the model to ``np.longdouble`` at the first trap.  Which exact
scipy call traps first is an implementation detail of scipy; the effect of the
promotion point on pi/theta/lnL is < 1e-15 relative (checked against the
imported reference by ``oracle/make_golden.py``).

Parity pinning: checked against (i) the known-answer matrices in the
reference's ``tests/test_sparse_plus.py:39-81`` (tests/test_oracle.py),
(ii) ``stellarscope/data/telescope_report.tsv`` through the config-1 fixture,
(iii) outputs of the imported reference itself on seeded inputs
(``tests/golden/*.npz`` written by ``oracle/make_golden.py``).
"""
from __future__ import annotations

from collections import Counter
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

REASSIGN_MODES = [  # model.py:92-100
    'best_exclude', 'best_conf', 'best_random', 'best_average',
    'initial_unique', 'initial_random', 'total_hits',
]


class OracleError(Exception):
    """Stands in for StellarscopeError (stellarscope/__init__.py)."""


@dataclass
class Opts:
    """The option attributes the hot path reads."""
    em_epsilon: float = 1e-7
    max_iter: int = 100
    pi_prior: float = 0
    theta_prior: float = 200000
    use_likelihood: bool = False
    pooling_mode: str = 'pseudobulk'
    reassign_mode: tuple = ('best_exclude',)
    conf_prob: float = 0.9
    ignore_umi: bool = True
    umi_counts: bool = False
    nproc: int = 1
    rng: Optional[np.random.Generator] = None


# --------------------------------------------------------------------------
# small CSR helpers (sparse_plus.py)
# --------------------------------------------------------------------------
def csr_eliminate_zeros(indptr, indices, data):
    """scipy ``eliminate_zeros`` on raw arrays (model.py:114)."""
    keep = data != 0
    if keep.all():
        return indptr, indices, data
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    cnt = np.bincount(rows[keep], minlength=len(indptr) - 1)
    new_ptr = np.zeros(len(indptr), dtype=indptr.dtype)
    np.cumsum(cnt, out=new_ptr[1:])
    return new_ptr, indices[keep], data[keep]


def row_of_entry(indptr):
    return np.repeat(np.arange(len(indptr) - 1, dtype=np.int64), np.diff(indptr))


def row_reduce(op, vals, indptr, empty=0):
    """Reduce ``vals`` over each CSR row; rows without entries get ``empty``."""
    n = len(indptr) - 1
    out = np.full(n, empty, dtype=vals.dtype)
    nz = np.flatnonzero(np.diff(indptr) > 0)
    if len(nz):
        out[nz] = op.reduceat(vals, indptr[nz])
    return out


def recip0(v):
    """sparse_plus.py:31-37 -- reciprocal with 1/0 -> 0."""
    with np.errstate(divide='ignore'):
        r = 1.0 / v
    r[np.isinf(r)] = 0
    return r


def norm_rows(indptr, data):
    """``csr_matrix_plus.norm(1)`` sparse_plus.py:63-64 (reciprocal, multiply)."""
    rs = row_reduce(np.add, data, indptr)
    return data * recip0(rs)[row_of_entry(indptr)]


def binmax_rows(indptr, data):
    """``binmax(1)`` sparse_plus.py:170-178: 1 where data == row max.

    scipy's sparse ``max(1)`` includes the implicit zeros of a row that is not
    completely filled, so a stored value only wins if it is >= 0 ... in the
    reference's use every stored value is > 0 or the row max is an explicit
    value; rows whose max is <= 0 produce nothing after ``eliminate_zeros``
    (an all-zero z row, model.py:433).  ``ncols`` is not needed for that."""
    rmax = row_reduce(np.maximum, data, indptr, empty=0)
    # implicit zeros: a row shorter than the matrix width has max >= 0
    rmax = np.maximum(rmax, 0)
    flag = (data == rmax[row_of_entry(indptr)]) & (data != 0)
    return flag


class _ColSummer:
    """Column sums in row-major accumulation order, any float dtype.

    scipy's ``csr.sum(0)`` is ``ones @ M`` -> accumulates row by row, i.e. for
    each column the addends are taken in ascending row order (model.py:191,
    239).  ``np.bincount`` does the same but only in float64, so the
    longdouble path uses a stable column sort + ``reduceat``."""

    def __init__(self, indices, K):
        self.K = K
        self.indices = indices
        self.perm = np.argsort(indices, kind='stable')
        sc = indices[self.perm]
        if len(sc):
            head = np.ones(len(sc), dtype=bool)
            head[1:] = sc[1:] != sc[:-1]
            self.starts = np.flatnonzero(head)
            self.cols = sc[self.starts]
        else:
            self.starts = np.zeros(0, dtype=np.int64)
            self.cols = np.zeros(0, dtype=np.int64)

    def __call__(self, vals):
        out = np.zeros(self.K, dtype=vals.dtype)
        if vals.dtype == np.float64:
            # bincount: sequential, ascending entry (= row-major) order
            with np.errstate(under='ignore'):
                return np.bincount(self.indices, weights=vals, minlength=self.K)
        if len(self.starts):
            out[self.cols] = np.add.reduceat(vals[self.perm], self.starts)
        return out


# --------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------
@dataclass
class FitInfo:  # statistics.py:269-307
    nobs: int = 0
    nfeats: int = 0
    nparams: int = 2
    epsilon: float = 0.0
    max_iter: int = 0
    converged: bool = False
    reached_max: bool = False
    iterations: list = field(default_factory=list)  # (inum, itime, diff)
    final_lnl: Optional[float] = None


class OracleLikelihood:
    """Restatement of ``TelescopeLikelihood`` on raw CSR arrays.

    ``indptr, indices, scores`` describe the N x K uint16 score matrix
    (model.py:113-118); explicit zeros are dropped like ``eliminate_zeros``.
    """

    def __init__(self, indptr, indices, scores, K, opts, promote=True):
        indptr = np.asarray(indptr, dtype=np.int64)
        indices = np.asarray(indices, dtype=np.int64)
        scores = np.asarray(scores)
        indptr, indices, scores = csr_eliminate_zeros(indptr, indices, scores)
        self.indptr, self.indices, self.raw = indptr, indices, scores
        self.N, self.K = len(indptr) - 1, int(K)
        self.opts = opts
        self.epsilon = opts.em_epsilon
        self.max_iter = opts.max_iter
        self.promote = promote
        self.promoted = False
        self.rows = row_of_entry(indptr)

        # model.py:115, 129-130 and sparse_plus.py:126-128:
        # Q = expm1( (raw * (1/max)) * 100 )
        self.max_score = int(scores.max()) if len(scores) else 0
        if len(scores):
            scaled = scores.astype(np.float64) * (1.0 / self.max_score)
            self.Q = np.expm1(scaled * 100.0)
        else:
            self.Q = np.zeros(0, dtype=np.float64)

        # model.py:140-143
        nnz_row = np.diff(indptr)
        self.y_amb = nnz_row > 1
        self.y_uni = nnz_row == 1
        self.y_0 = nnz_row == 0
        self.amb_e = self.y_amb[self.rows]          # entry belongs to an ambiguous row
        self.uni_e = self.y_uni[self.rows]

        # model.py:164, 172
        self.pi = np.repeat(1.0 / self.K, self.K)
        self.theta = np.repeat(1.0 / self.K, self.K)
        self.pi_init = self.theta_init = None
        self.z = None
        self.lnl = float('inf')

        # model.py:183-194
        self.w = row_reduce(np.maximum, self.Q, indptr, empty=0.0)   # Q.max(1)
        self.total_wt = self.w.sum()
        self.ambig_wt = self.w[self.y_amb].sum()
        self.unique_wt = self.w[self.y_uni].sum()
        wmax = self.w.max() if self.N else 0.0
        self.pi_prior_wt = opts.pi_prior * wmax
        self.theta_prior_wt = opts.theta_prior * wmax
        self._colsum = _ColSummer(indices, self.K)
        self.pisum0 = self._colsum(np.where(self.uni_e, self.Q, 0.0))
        self.theta_denom = self.ambig_wt + self.theta_prior_wt * self.K
        self.pi_denom = self.total_wt + self.pi_prior_wt * self.K

        # statistics.py:289-294
        self.fitinfo = FitInfo(
            nobs=int((nnz_row > 0).sum()),
            nfeats=int(len(np.unique(indices))),
            epsilon=self.epsilon, max_iter=self.max_iter)

    # -- helpers ----------------------------------------------------------
    def _promote(self):
        xd = np.longdouble
        self.promoted = True
        self.Q = self.Q.astype(xd)
        self.pi = self.pi.astype(xd)
        self.theta = self.theta.astype(xd)

    def _guard(self, fn, *args):
        """Run ``fn`` with underflow trapped; on a trap promote and redo.

        model.py:223-226 (estep), :267-272 (mstep), :290-320 (lnl)."""
        if not self.promote:
            with np.errstate(under='ignore'):
                return fn(*args)
        try:
            with np.errstate(divide='raise', over='raise', under='raise',
                             invalid='raise'):
                return fn(*args)
        except FloatingPointError as e:
            if 'underflow' not in str(e):
                raise OracleError(str(e))
            self._promote()
            args = tuple(a.astype(np.longdouble) if isinstance(a, np.ndarray)
                         and a.dtype.kind == 'f' else a for a in args)
            with np.errstate(under='ignore'):
                return fn(*args)

    # -- E step: model.py:207-210 ------------------------------------------
    def estep(self, pi, theta):
        def _e(pi, theta):
            pij = pi[self.indices]
            # (_ambQ * pi) * theta  +  _uniQ * pi
            u = np.where(self.amb_e, (self.Q * pij) * theta[self.indices],
                         self.Q * pij)
            return norm_rows(self.indptr, u)
        return self._guard(_e, pi, theta)

    # -- M step: model.py:234-246 ------------------------------------------
    def mstep(self, z):
        def _m(z):
            weighted = z * self.w[self.rows]
            thetasum = self._colsum(np.where(self.amb_e, weighted,
                                             np.zeros(1, dtype=weighted.dtype)))
            if self.theta_denom == 0 or self.pi_denom == 0:
                raise OracleError('invalid value encountered in divide')
            theta_hat = (thetasum + self.theta_prior_wt) / self.theta_denom
            pisum = self.pisum0 + thetasum
            pi_hat = (pisum + self.pi_prior_wt) / self.pi_denom
            return pi_hat, theta_hat
        return self._guard(_m, z)

    # -- lnL: model.py:291-317 ----------------------------------------------
    def calculate_lnl(self, z, pi, theta):
        def _l(z, pi, theta):
            pitheta = pi * theta
            inner = np.where(self.amb_e, self.Q * pitheta[self.indices],
                             self.Q * pi[self.indices])
            # the reference takes log over *stored* entries of (_amb + _uni);
            # an entry that became exactly 0 is dropped by the sparse add
            ok = inner > 0
            li = np.zeros(len(inner), dtype=inner.dtype)
            li[ok] = np.log(inner[ok])
            return (z * li).sum()
        return self._guard(_l, z, pi, theta)

    # -- EM loop: model.py:325-373 -------------------------------------------
    def em(self):
        inum = 0
        use_l = self.opts.use_likelihood
        fi = self.fitinfo
        if len(self.Q) == 0:
            # theta_denom == 0 -> FloatingPointError -> StellarscopeError
            raise OracleError('empty model (no alignments)')
        while not (fi.converged or fi.reached_max):
            z = self.estep(self.pi, self.theta)
            pi, theta = self.mstep(z)
            inum += 1
            if inum == 1:
                self.pi_init, self.theta_init = pi, theta
            with np.errstate(under='ignore'):
                diff_est = np.abs(pi - self.pi).sum()
            if use_l:
                lnl = self.calculate_lnl(z, pi, theta)
                fi.converged = bool(abs(lnl - self.lnl) < self.epsilon)
                self.lnl = lnl
            else:
                fi.converged = bool(diff_est < self.epsilon)
            fi.reached_max = inum >= self.max_iter
            self.z, self.pi, self.theta = z, pi, theta
            fi.iterations.append((inum, 0.0, float(diff_est)))
        if not use_l:
            self.lnl = self.calculate_lnl(self.z, self.pi, self.theta)
        fi.final_lnl = self.lnl
        return self

    # -- reassign: model.py:375-507 -------------------------------------------
    def reassign(self, mode, thresh=None, rng=None, z=None):
        """Returns ``(indptr, indices, data, info)`` -- a CSR with zeros
        eliminated -- and ``info = dict(assigned, ambiguous, unaligned,
        ambigous_dist)``."""
        if mode not in REASSIGN_MODES:
            raise ValueError(f'Argument "method" should be one of {REASSIGN_MODES}')
        z = self.z if z is None else z
        rng = rng if rng is not None else self.opts.rng
        info = dict(assigned=0, ambiguous=0, unaligned=None, ambigous_dist=None)
        N = self.N
        if mode in ('best_exclude', 'best_random', 'best_average'):
            z64 = np.asarray(z)
            flag = binmax_rows(self.indptr, z64)
            nbest = np.bincount(self.rows[flag], minlength=N)
            info.update(assigned=int((nbest == 1).sum()),
                        ambiguous=int((nbest > 1).sum()),
                        unaligned=int((nbest == 0).sum()),
                        ambigous_dist=Counter(nbest.tolist()))
            if mode == 'best_exclude':
                keep = flag & (nbest[self.rows] == 1)
                vals = np.ones(int(keep.sum()), dtype=np.int8)
            elif mode == 'best_random':
                keep = _choose_random(flag, self.rows, nbest, rng)
                vals = np.ones(int(keep.sum()), dtype=np.int8)
            else:
                keep = flag
                vals = norm_rows(_ptr(self.rows[keep], N),
                                 np.ones(int(keep.sum()), dtype=np.float64))
        elif mode == 'best_conf':
            if thresh is None or thresh <= 0.5:
                raise OracleError(f'Confidence threshold ({thresh}) must be > 0.5')
            z64 = np.asarray(z)
            keep = z64 > thresh
            vals = np.ones(int(keep.sum()), dtype=np.int8)
            # rowmax = z.max(1): sparse N x 1; ``.data`` holds the nonzero maxima
            rmax = np.maximum(row_reduce(np.maximum, z64, self.indptr, empty=0), 0)
            stored = rmax != 0
            info.update(assigned=int((rmax[stored] > thresh).sum()),
                        ambiguous=int((rmax[stored] <= thresh).sum()),
                        unaligned=int(N - stored.sum()))
        elif mode == 'initial_unique':
            keep = self.uni_e & (self.Q > 0)
            vals = np.ones(int(keep.sum()), dtype=np.int8)
            info.update(assigned=int(keep.sum()), ambiguous=int(self.y_amb.sum()),
                        unaligned=int(self.y_0.sum()))
        elif mode == 'initial_random':
            flag = binmax_rows(self.indptr, self.raw.astype(np.int64))
            nbest = np.bincount(self.rows[flag], minlength=N)
            info.update(assigned=int((nbest == 1).sum()),
                        ambiguous=int((nbest > 1).sum()),
                        unaligned=int((nbest == 0).sum()),
                        ambigous_dist=Counter(nbest.tolist()))
            keep = _choose_random(flag, self.rows, nbest, rng)
            vals = np.ones(int(keep.sum()), dtype=np.int8)
        else:  # total_hits
            keep = self.raw > 0
            vals = np.ones(int(keep.sum()), dtype=np.int8)
            per_row = np.diff(self.indptr)
            info.update(assigned=int(keep.sum()),
                        ambiguous=int((per_row > 1).sum()),
                        unaligned=int((per_row == 0).sum()))
        rows = self.rows[keep]
        return _ptr(rows, N), self.indices[keep], vals, info


def _ptr(sorted_rows, N):
    ptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(np.bincount(sorted_rows, minlength=N), out=ptr[1:])
    return ptr


def _choose_random(flag, rows, nbest, rng):
    """``choose_random(1, rng)`` sparse_plus.py:231-240 on the binmax matrix
    (zeros already eliminated): for every row with more than one stored entry,
    in row order, ``rng.choice(range(start, end))`` keeps one."""
    keep = flag.copy()
    pos = np.flatnonzero(flag)                 # entries of the compacted matrix
    prow = rows[pos]
    start = np.zeros(len(nbest) + 1, dtype=np.int64)
    np.cumsum(nbest, out=start[1:])
    for r in np.flatnonzero(nbest > 1):
        a, b = start[r], start[r + 1]
        chosen = rng.choice(range(a, b))
        drop = pos[a:b]
        keep[drop] = False
        keep[pos[chosen]] = True
    return keep


# --------------------------------------------------------------------------
# pooling driver: model.py:675-804
# --------------------------------------------------------------------------
def submatrix_rows(indptr, indices, data, rows):
    """Row-compacted submatrix.  The reference masks the full-height matrix
    (``_fullmat.multiply(rowid(rows, N))`` model.py:722-723, 742-743); dropping
    the all-zero rows instead is bit-identical for pi/theta/z/lnL
    because empty rows contribute nothing to any sum."""
    rows = np.asarray(rows, dtype=np.int64)
    lens = indptr[rows + 1] - indptr[rows]
    ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    take = np.repeat(indptr[rows] - ptr[:-1], lens) + np.arange(ptr[-1])
    return ptr, indices[take], data[take], take


@dataclass
class PoolResult:
    z: np.ndarray                      # data of the N x K z matrix, CSR order of the full matrix
    fits: dict                         # models_info: key -> FitInfo  (model.py:766, 791)
    models: dict                       # key -> OracleLikelihood (pi/theta per model; not kept by the reference)
    total_lnl: float = 0.0
    total_obs: int = 0
    total_params: int = 0
    AIC: float = 0.0
    BIC: float = 0.0


def fit_pooling_model(indptr, indices, scores, K, opts, segments=None, promote=True):
    """``_fit_pooling_model`` model.py:675-804.

    ``segments``: for 'individual' the list of row-id lists in
    ``st.bcode_ridx_map`` insertion order (model.py:734-735, rows are sorted);
    for 'celltype' one list per ``st.celltypes`` entry (:710-716, concatenated
    in barcode order -- ``rowid`` sorts and de-duplicates them,
    sparse_plus.py:434).  Empty lists are skipped but still consume no index
    (model.py:718-720, 738-740: ``continue`` before ``yield``; ``enumerate``
    numbers only the yielded models)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    indptr, indices, scores = csr_eliminate_zeros(indptr, indices, np.asarray(scores))
    if opts.pooling_mode == 'pseudobulk':
        m = OracleLikelihood(indptr, indices, scores, K, opts, promote).em()
        res = PoolResult(z=np.asarray(m.z, dtype=np.float64),
                         fits={'pseudobulk': m.fitinfo}, models={'pseudobulk': m})
    else:
        z = np.zeros(len(indices), dtype=np.float64)
        fits, models = {}, {}
        i = 0
        for rows in segments:
            rows = sorted(set(int(r) for r in rows))
            if not rows:
                continue
            ptr, idx, sc, take = submatrix_rows(indptr, indices, scores, rows)
            m = OracleLikelihood(ptr, idx, sc, K, opts, promote).em()
            with np.errstate(under='ignore'):
                z[take] += np.asarray(m.z, dtype=np.float64)   # ret_model.z += _z (:790)
            fits[i] = m.fitinfo
            models[i] = m
            i += 1
        res = PoolResult(z=z, fits=fits, models=models)
    # statistics.py:337-393
    res.total_lnl = float(sum(np.longdouble(f.final_lnl) for f in res.fits.values()))
    res.total_obs = sum(f.nobs for f in res.fits.values())
    res.total_params = sum(f.nparams * f.nfeats for f in res.fits.values())
    res.AIC = 2 * res.total_params - 2 * res.total_lnl
    res.BIC = (res.total_params * np.log(res.total_obs) - 2 * res.total_lnl
               if res.total_obs else float('nan'))
    return res


# --------------------------------------------------------------------------
# count aggregation: model.py:1302-1345, sparse_plus.py:367-396
# --------------------------------------------------------------------------
def aggregate_counts(rindptr, rindices, rdata, K, row_bc, bc_list, row_umi=None):
    """``agg_bc`` / ``agg_bc_umi``: dense (K x len(bc_list)) count matrix.

    ``row_bc[i]`` is the barcode label of row i (model.py:1389-1390),
    ``bc_list`` the output barcode order (:1371-1375).  With ``row_umi`` the
    rows are first summed per (barcode, UMI), every nonzero set to 1, then
    summed per barcode (:1334-1345; the result is float64 because of
    ``np.ones``, :1341)."""
    rows = row_of_entry(np.asarray(rindptr, dtype=np.int64))
    integer = np.issubdtype(np.asarray(rdata).dtype, np.integer)
    dtype = np.int32 if integer else np.float64
    bc_pos = {b: i for i, b in enumerate(bc_list)}
    out = np.zeros((K, len(bc_list)), dtype=dtype)
    if row_umi is None:
        for e in range(len(rindices)):
            b = bc_pos.get(row_bc[rows[e]])
            if b is not None:
                out[rindices[e], b] += rdata[e]
        return out
    seen = {}
    for e in range(len(rindices)):
        r = rows[e]
        key = (row_bc[r], row_umi[r], int(rindices[e]))
        seen[key] = seen.get(key, 0) + rdata[e]
    out = np.zeros((K, len(bc_list)), dtype=np.float64)
    for (bc, _umi, col), v in seen.items():
        if v != 0 and bc in bc_pos:
            out[col, bc_pos[bc]] += 1.0
    return out
