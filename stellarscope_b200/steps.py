"""Single steps of the model -- ``TelescopeLikelihood.estep / mstep / calculate_lnl`` of the reference
(stellarscope/utils/model.py:202-323) for callers that drive the iteration themselves.

``em()`` / ``fit_segments()`` never come here: the fit runs in the fused kernels of libstellarscope_b200.so with the
loop on the device.  These three methods exist because the reference class has them (SURVEY.md section 1, L2
interface); they evaluate one step for ONE model on the device with torch segment operations over the matrix the
engine already holds (no CPU evaluation, no second upload) and hand back the reference's types: ``estep`` an N x K
CSR matrix, ``mstep`` ``(pi as 1 x K sparse, theta as K array)``, ``calculate_lnl`` a float.  fp64; no float128
promotion (values below 1e-290 flush).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse
import torch

from .sparse_plus import csr_matrix_plus


def _dense(v, K):
    if scipy.sparse.issparse(v):
        v = np.asarray(v.todense())
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1)[:K])


class ModelSteps:
    def __init__(self, tl):
        eng = tl._engine
        if tl.n_segments != 1 or getattr(tl, '_row_ids', None) is not None:
            raise NotImplementedError('estep / mstep / calculate_lnl are for a single model held by one GPU')
        self.tl, self.dev = tl, eng.device
        self.N, self.K = eng.N, eng.K
        rp = eng._keep['row_ptr'].to(torch.int64)
        self.ci = eng._keep['col_idx'].to(torch.int64)
        sc = eng._keep['score'].to(torch.int32)
        lens = rp[1:] - rp[:-1]
        self.rows = torch.repeat_interleave(torch.arange(self.N, device=self.dev), lens, output_size=int(self.ci.numel()))
        smax = int(sc.max()) if sc.numel() else 1
        self.Q = torch.expm1((sc.to(torch.float64) * (1.0 / smax)) * 100.0)          # model.py:129-130
        self.w = torch.zeros(self.N, dtype=torch.float64, device=self.dev)
        if self.Q.numel():
            self.w.scatter_reduce_(0, self.rows, self.Q, 'amax', include_self=True)   # model.py:183
        self.amb_row = lens > 1
        self.amb_e, uni_e = self.amb_row[self.rows], (lens == 1)[self.rows]
        wmax = float(self.w.max()) if self.N else 0.0
        self.pi_prior_wt = tl.pi_prior * wmax                                          # model.py:189-190
        self.theta_prior_wt = tl.theta_prior * wmax
        self.pisum0 = torch.zeros(self.K, dtype=torch.float64, device=self.dev)
        self.pisum0.index_add_(0, self.ci[uni_e], self.Q[uni_e])                       # model.py:191
        self.theta_denom = float(self.w[self.amb_row].sum()) + self.theta_prior_wt * self.K
        self.pi_denom = float(self.w.sum()) + self.pi_prior_wt * self.K
        self.keys = self.rows * self.K + self.ci                                       # ascending (canonical CSR)

    def _inner(self, pi, theta):
        """Q_ij pi_j theta_j on ambiguous rows, Q_ij pi_j on unique rows (model.py:207-209)."""
        p = torch.from_numpy(_dense(pi, self.K)).to(self.dev)
        t = torch.from_numpy(_dense(theta, self.K)).to(self.dev)
        qp = self.Q * p[self.ci]
        return torch.where(self.amb_e, qp * t[self.ci], qp)

    def estep(self, pi, theta):
        v = self._inner(pi, theta)
        rs = torch.zeros(self.N, dtype=torch.float64, device=self.dev).index_add_(0, self.rows, v)
        inv = torch.where(rs > 0, 1.0 / torch.where(rs > 0, rs, torch.ones_like(rs)), torch.zeros_like(rs))
        z = (v * inv[self.rows]).cpu().numpy()                                         # norm(1), sparse_plus.py:42-64
        m = self.tl.raw_scores
        return csr_matrix_plus((z, m.indices.copy(), m.indptr.copy()), shape=m.shape)

    def _z_entries(self, z):
        z = scipy.sparse.csr_matrix(z)
        rows = np.repeat(np.arange(z.shape[0], dtype=np.int64), np.diff(z.indptr))
        return (torch.from_numpy(rows).to(self.dev), torch.from_numpy(z.indices.astype(np.int64)).to(self.dev),
                torch.from_numpy(np.asarray(z.data, dtype=np.float64)).to(self.dev))

    def mstep(self, z):
        zr, zc, zv = self._z_entries(z)
        weighted = zv * self.w[zr] * self.amb_row[zr].to(torch.float64)                # model.py:236-239
        thetasum = torch.zeros(self.K, dtype=torch.float64, device=self.dev).index_add_(0, zc, weighted)
        if not (self.theta_denom > 0 and self.pi_denom > 0):
            raise FloatingPointError('invalid value encountered in divide')
        theta_hat = (thetasum + self.theta_prior_wt) / self.theta_denom
        pi_hat = (self.pisum0 + thetasum + self.pi_prior_wt) / self.pi_denom
        return csr_matrix_plus(pi_hat.cpu().numpy()), theta_hat.cpu().numpy()

    def calculate_lnl(self, z, pi, theta):
        inner = self._inner(pi, theta)
        zr, zc, zv = self._z_entries(z)
        zk = zr * self.K + zc
        pos = torch.searchsorted(self.keys, zk).clamp_(max=max(int(self.keys.numel()) - 1, 0))
        hit = (self.keys[pos] == zk) & (inner[pos] > 0) if self.keys.numel() else torch.zeros_like(zk, dtype=torch.bool)
        return float((zv[hit] * torch.log(inner[pos][hit])).sum())                     # model.py:303-317
