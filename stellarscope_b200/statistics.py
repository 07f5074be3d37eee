"""Result records of the hot path: FitInfo, PoolInfo, ReassignInfo.

Same attribute / property names and derived quantities as the reference's
``stellarscope/utils/statistics.py:269-495`` so that ``stages.py`` logging and
``output_stats`` (statistics.py:628-644) keep working; the values are filled
from device results instead of from a fitted scipy model.
"""
from __future__ import annotations

import logging as lg
from collections import Counter
from typing import Optional

import numpy as np


def _property_names(obj):
    """Properties of the record, as the reference lists them (statistics.py:18-20) -- minus the ones that are plain
    attributes there and properties only here (``_not_in_reference``), so ``to_dataframe`` / stats.final.tsv show
    the reference's rows."""
    cls = obj.__class__
    skip = getattr(cls, '_not_in_reference', ())
    return [p for p in dir(cls) if isinstance(getattr(cls, p), property) and p not in skip]


class GenericInfo:
    """statistics.py:21-44"""

    def __init__(self):
        self.infotype = self.__class__.__name__

    def __str__(self):
        return '\n'.join(f'{k}\t{v}' for k, v in vars(self).items())

    def to_dataframe(self):
        import pandas as pd
        names = _property_names(self)
        return pd.DataFrame({'stage': self.infotype, 'mode': '', 'var': names,
                             'value': [getattr(self, v) for v in names]}, dtype=object)


class FitInfo(GenericInfo):
    """statistics.py:269-307.  ``iterations`` holds ``(inum, itime, diff_est)``
    (model.py:359); ``itime`` is the device time of the whole fit divided by the
    number of iterations (the reference measures each iteration on the host)."""
    _not_in_reference = ('iterations', 'n_iterations')     # lazy here, a plain list attribute in the reference

    def __init__(self, nobs=None, nfeats=None, epsilon=None, max_iter=None):
        super().__init__()
        self.converged = False
        self.reached_max = False
        self.nparams = 2
        self.nobs = nobs
        self.nfeats = nfeats
        self.epsilon = epsilon
        self.max_iter = max_iter
        self._iterations = []
        self._trace = None            # (n, itime, diffs ndarray): materialised on first access
        self.final_lnl = None

    @property
    def iterations(self):
        """list of ``(inum, itime, diff_est)`` (model.py:359).  With thousands of models the
        tuples are only built when somebody looks at them."""
        if self._trace is not None:
            n, itime, diffs = self._trace
            self._iterations = [(i + 1, itime, float(diffs[i])) for i in range(n)]
            self._trace = None
        return self._iterations

    @iterations.setter
    def iterations(self, value):
        self._trace = None
        self._iterations = value

    @property
    def n_iterations(self) -> int:
        return self._trace[0] if self._trace is not None else len(self._iterations)

    def set_trace(self, n, itime, diffs):
        self._trace = (int(n), float(itime), diffs)

    @property
    def fitted(self) -> bool:
        return self.converged | self.reached_max


class SegmentFitInfos:
    """``FitInfo`` records of all models of one fit, backed by the arrays that came back from
    the device.  Behaves like a read-only sequence (and, with ``keys``, like the dict
    ``PoolInfo.models_info`` of the reference, model.py:766, 791): a record is only built when
    somebody asks for it -- with 10^4..10^5 models, building them all costs more than the fit."""

    def __init__(self, res, epsilon, max_iter, itime, keys=None, order=None):
        self._res = res                      # dict of arrays: iters, flags, lnl, nobs, nfeats, diffs
        self._eps, self._mi, self._itime = epsilon, max_iter, itime
        self._keys = list(keys) if keys is not None else None
        self._pos = None if self._keys is None else {k: i for i, k in enumerate(self._keys)}
        self._order = order                  # optional: position -> row of the arrays
        self._cache = {}

    # -- sequence protocol --
    def __len__(self):
        return len(self._keys) if self._keys is not None else len(self._res['iters'])

    def _make(self, i):
        fi = self._cache.get(i)
        if fi is None:
            r = self._res
            j = i if self._order is None else int(self._order[i])
            fi = FitInfo(nobs=int(r['nobs'][j]), nfeats=int(r['nfeats'][j]), epsilon=self._eps, max_iter=self._mi)
            fl = int(r['flags'][j])
            fi.converged = bool(fl & 1)
            fi.reached_max = bool(fl & 2)
            fi.set_trace(int(r['iters'][j]), self._itime, r['diffs'][j])
            fi.final_lnl = float(r['lnl'][j])
            self._cache[i] = fi
        return fi

    def __getitem__(self, k):
        if self._keys is not None:
            return self._make(self._pos[k])
        if isinstance(k, slice):
            return [self._make(i) for i in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        if not 0 <= k < len(self):
            raise IndexError(k)
        return self._make(k)

    def __iter__(self):
        if self._keys is not None:
            return iter(self._keys)
        return (self._make(i) for i in range(len(self)))

    # -- mapping protocol (when keyed) --
    def keys(self):
        return list(self._keys)

    def values(self):
        return (self._make(i) for i in range(len(self)))

    def items(self):
        return ((k, self._make(i)) for i, k in enumerate(self._keys))

    def __contains__(self, k):
        return k in self._pos if self._pos is not None else False

    def keyed(self, keys):
        out = SegmentFitInfos(self._res, self._eps, self._mi, self._itime, keys=keys, order=self._order)
        return out

    # -- vectorised totals (PoolInfo) --
    def _rows(self):
        return slice(None) if self._order is None else np.asarray(self._order)

    def total_lnl(self):
        """float128 accumulation in model order (statistics.py:338-343)."""
        v = np.asarray(self._res['lnl'])[self._rows()].astype(np.longdouble)
        return np.cumsum(v)[-1] if len(v) else np.longdouble(0.0)

    def total_obs(self):
        return int(np.asarray(self._res['nobs'])[self._rows()].astype(np.int64).sum())

    def total_params(self):
        return int(2 * np.asarray(self._res['nfeats'])[self._rows()].astype(np.int64).sum())


class PoolInfo(GenericInfo):
    """statistics.py:309-407."""

    def __init__(self, pooling_mode: Optional[str] = None):
        super().__init__()
        self.pooling_mode = pooling_mode
        self.nmodels = 0
        self.models_info = {}
        self._total_params = None
        self._total_obs = None
        self._total_lnl = None
        self._AIC = None
        self._BIC = None

    @property
    def fitted_models(self) -> int:
        return len(self.models_info)

    @property
    def total_lnl(self):
        if self._total_lnl is None:
            if isinstance(self.models_info, SegmentFitInfos):
                self._total_lnl = self.models_info.total_lnl()
            else:
                tot = np.longdouble(0.0)               # float128 accumulation, statistics.py:340
                for fi in self.models_info.values():
                    tot += fi.final_lnl
                self._total_lnl = tot
        return self._total_lnl

    @property
    def total_obs(self) -> int:
        if self._total_obs is None:
            if isinstance(self.models_info, SegmentFitInfos):
                self._total_obs = self.models_info.total_obs()
            else:
                self._total_obs = sum(fi.nobs for fi in self.models_info.values())
        return self._total_obs

    @property
    def total_params(self) -> int:
        if self._total_params is None:
            if isinstance(self.models_info, SegmentFitInfos):
                self._total_params = self.models_info.total_params()
            else:
                self._total_params = sum(fi.nparams * fi.nfeats for fi in self.models_info.values())
        return self._total_params

    @property
    def AIC(self):
        if self._AIC is None:
            self._AIC = 2 * self.total_params - (2 * self.total_lnl)
        return self._AIC

    @property
    def BIC(self):
        if self._BIC is None:
            self._BIC = self.total_params * np.log(self.total_obs) - (2 * self.total_lnl)
        return self._BIC

    def log(self, loglev=lg.INFO):
        lg.log(loglev, f'Complete data log-likelihood (lnL): {self.total_lnl}')
        lg.log(loglev, f'  Number of models estimated: {self.fitted_models}')
        lg.log(loglev, f'  Total observations: {self.total_obs}')
        lg.log(loglev, f'  Total parameters estimated: {self.total_params}')
        lg.log(loglev, f'    AIC: {self.AIC}')
        lg.log(loglev, f'    BIC: {self.BIC}')

    def to_dataframe(self):
        ret = super().to_dataframe()
        ret['mode'] = self.pooling_mode
        return ret


class ReassignInfo(GenericInfo):
    """statistics.py:410-495 (the attribute is spelled ``ambigous_dist`` by the
    code that fills it, model.py:440)."""

    explanations = {
        'best_exclude': 'remain ambiguous -> excluded',
        'best_conf': 'are low confidence -> excluded',
        'best_random': 'remain ambiguous -> randomly assigned',
        'best_average': 'remain ambiguous -> divided evenly',
        'initial_unique': 'were initially ambiguous -> discarded',
        'initial_random': 'were initially ambiguous -> randomly assigned',
        'total_hits': 'initially had multiple alignments',
    }

    def __init__(self, reassign_mode: str):
        super().__init__()
        self.reassign_mode = reassign_mode
        self.explanation = self.explanations[reassign_mode]
        self._assigned = 0
        self._ambiguous = 0
        self._unaligned = None
        self.ambiguous_dist = None

    @property
    def assigned(self):
        return self._assigned

    @assigned.setter
    def assigned(self, val):
        self._assigned = val

    @property
    def ambiguous(self):
        return self._ambiguous

    @ambiguous.setter
    def ambiguous(self, val):
        self._ambiguous = val

    @property
    def unaligned(self):
        return self._unaligned

    @unaligned.setter
    def unaligned(self, val):
        self._unaligned = val

    def format(self, include_mode=True, show_unaligned=False):
        m = self.reassign_mode
        prefix = f'{m} - ' if include_mode else ''
        ret = []
        if m == 'total_hits':
            ret.append(f'{prefix}{self.assigned} total alignments.')
        elif m == 'initial_unique':
            ret.append(f'{prefix}{self.assigned} uniquely mapped.')
        else:
            ret.append(f'{prefix}{self.assigned} assigned.')
        ret.append(f'{prefix}{self.ambiguous} reads {self.explanation}.')
        if show_unaligned and self.unaligned is not None:
            ret.append(f'{prefix}{self.unaligned} had no alignments.')
        return ret

    def log(self, loglev=lg.INFO):
        lg.log(loglev, f'Reassignment using {self.reassign_mode}')
        for line in self.format(False):
            lg.log(loglev, f'  {line}')

    def to_dataframe(self):
        ret = super().to_dataframe()
        ret['mode'] = self.reassign_mode
        return ret


def counter_from_hist(hist):
    """Device histogram of min(nbest, 63) -> ``Counter(nbest)`` (model.py:440)."""
    return Counter({int(k): int(v) for k, v in enumerate(hist) if v})


class UMIInfo(GenericInfo):
    """statistics.py:498-626 -- filled from the per-pair read / component counts the device returns
    instead of from the reference's dict of dicts."""

    def __init__(self, group_sizes=None):
        super().__init__()
        self.rpu_counter = Counter()
        self.rpu_bins = []
        self.rpu_hist = None
        if group_sizes is not None and len(group_sizes):
            self.init_rpu(group_sizes)
        self.ncomps_umi = Counter()
        self._nexclude = 0

    num_umi = property(lambda self: sum(self.rpu_counter.values()))
    uni_umi = property(lambda self: self.rpu_counter[1])
    dup_umi = property(lambda self: sum(f for n, f in self.rpu_counter.items() if n > 1))
    max_rpu = property(lambda self: max(self.rpu_counter))
    possible_dups = property(lambda self: sum((n - 1) * f for n, f in self.rpu_counter.items()))
    max_comps = property(lambda self: max(self.ncomps_umi) if self.ncomps_umi else 0)

    @property
    def nexclude(self):
        return self._nexclude

    @nexclude.setter
    def nexclude(self, val):
        self._nexclude = val

    def init_rpu(self, group_sizes):
        """statistics.py:552-570 with the reads-per-pair counts as an array."""
        rpu = np.asarray(group_sizes)
        vals, freq = np.unique(rpu, return_counts=True)
        self.rpu_counter = Counter({int(v): int(f) for v, f in zip(vals, freq)})
        if self.max_rpu <= 5:
            self.rpu_bins = list(range(1, self.max_rpu + 2))
        else:
            self.rpu_bins = [1, 2, 3, 4, 5, 6, 11, 21]
            if self.max_rpu > 20:
                self.rpu_bins.append(self.max_rpu + 1)
        self.rpu_hist, _bins = np.histogram(rpu, self.rpu_bins)

    def _bin_label(self, b_i):
        bs, be = self.rpu_bins[b_i], self.rpu_bins[b_i + 1]
        if be == self.rpu_bins[-1]:
            return f'>{bs - 1}'
        return f'{bs}' if be - bs == 1 else f'{bs}-{be - 1}'

    def loginit(self, loglev=lg.INFO):
        lg.log(loglev, f'Number of BC+UMI pairs: {self.num_umi}')
        lg.log(loglev, f'  unique UMIs: {self.uni_umi}')
        lg.log(loglev, f'  duplicated UMIs: {self.dup_umi}')
        lg.log(loglev, f'  max reads per UMI: {self.max_rpu}')
        for b_i, v in enumerate(self.rpu_hist):
            lg.log(loglev, f'    UMIs with {self._bin_label(b_i)} reads: {v}')
        lg.log(loglev, f'  Possible duplicate reads: {self.possible_dups}')
        lg.log(loglev, 'Finding duplicates and selecting representatives...')

    def log(self, loglev=lg.INFO):
        lg.log(loglev, f'    UMIs with 1 component: {self.ncomps_umi[1]}')
        lg.log(loglev, f'    UMIs with 2 components: {self.ncomps_umi[2]}')
        lg.log(loglev, f'    UMIs with 3 components: {self.ncomps_umi[3]}')
        _gt3 = sum(v for k, v in self.ncomps_umi.items() if k > 3)
        if _gt3:
            lg.log(loglev, f'    UMIs with >3 components: {_gt3}')
        lg.log(loglev, f'Total UMI duplicate reads excluded: {self.nexclude}')

    def to_dataframe(self, binned=True):
        ret = super().to_dataframe()
        if binned:
            for b_i, v in enumerate(self.rpu_hist):
                ret.loc[len(ret.index)] = ['UMIInfo', 'umidist', self._bin_label(b_i), v]
        else:
            for nread in range(1, self.max_rpu + 1):
                ret.loc[len(ret.index)] = ['UMIInfo', 'umidist', nread, self.rpu_counter[nread]]
        for ncomp in range(1, self.max_comps + 1):
            ret.loc[len(ret.index)] = ['UMIInfo', 'compdist', ncomp, self.ncomps_umi[ncomp]]
        return ret


def bind_reference_records(ref_stats):
    """``model.install()``: use the reference's OWN record classes where the reference package is importable
    (``stellarscope.utils.statistics``): ``ReassignInfo`` as it is, ``PoolInfo`` subclassed only to keep the
    vectorised totals over the device result arrays (its loop over ``models_info.items()`` would build 10^5
    ``FitInfo`` objects).  The classes above remain for the reference-less GPU box; tests compare the two
    (tests/test_statistics_vs_reference.py).  Returns ``(PoolInfo, ReassignInfo)``."""
    class PoolInfo(ref_stats.PoolInfo):
        __doc__ = ref_stats.PoolInfo.__doc__

        def __init__(self, pooling_mode=None):
            super().__init__(pooling_mode)
            self.infotype = 'PoolInfo'

        @property
        def total_lnl(self):
            if self._total_lnl is None and isinstance(self.models_info, SegmentFitInfos):
                self._total_lnl = self.models_info.total_lnl()
            return ref_stats.PoolInfo.total_lnl.fget(self)

        @property
        def total_obs(self):
            if self._total_obs is None and isinstance(self.models_info, SegmentFitInfos):
                self._total_obs = self.models_info.total_obs()
            return ref_stats.PoolInfo.total_obs.fget(self)

        @property
        def total_params(self):
            if self._total_params is None and isinstance(self.models_info, SegmentFitInfos):
                self._total_params = self.models_info.total_params()
            return ref_stats.PoolInfo.total_params.fget(self)

    PoolInfo.__name__ = PoolInfo.__qualname__ = 'PoolInfo'
    return PoolInfo, ref_stats.ReassignInfo
