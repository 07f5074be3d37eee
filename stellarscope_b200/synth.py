"""Seeded synthetic 10x-style score matrices.

The reference has no benchmark input of its own for the EM path; BASELINE.json
names synthetic shapes ("10k cells, 5M multimapped alignments over 28k HERV/L1
loci", ...).  This generator produces the N x K uint16 score matrix that
``Stellarscope.load_alignment`` would hand to ``_fit_pooling_model``
(reference ``utils/model.py:958-978`` builds it from a BAM; values are
``(AS - minAS + 1) + alnlen``), together with the row -> cell map that the
reference keeps in ``bcode_ridx_map`` (model.py:875, 1053).

Model of the data (all draws from ``numpy.random.default_rng(seed)``):

* K = ``n_loci`` + 1 columns, column 0 is ``__no_feature``.
* loci are grouped in families of 4..64 members (HERV / L1 like); a read picks
  a family by Zipf(1.1) popularity and k distinct members,
  k = 2 + Poisson(1.5) clipped to [2, 32] (and to the family size);
  ``unique_frac`` of the rows get k = 1 instead (exercises pisum0).
* score = 140 + Binomial(120, p), p = 0.75 for the read's true locus and 0.65
  for the other members -> integers in [140, 260].
* with probability ``nofeat_frac`` a row carries an extra alignment to col 0.
* structural ties: in every family with >= 6 members the last member is a
  *mirror* of member 0 -- it is never drawn on its own and is added, with the
  same score, to every row that contains member 0 (duplicated loci that always
  co-occur; the reference then produces exact ties).
  ``tie_frac`` of the rows are forced to have member 0 as their true locus.
* cell sizes ~ LogNormal(sigma=1) normalised to the target number of
  alignments; the rows of a cell are contiguous.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_PRIMES = np.array([67, 71, 73, 79, 83, 89, 97, 101], dtype=np.int64)


@dataclass
class SynthData:
    indptr: np.ndarray      # int32 [N+1]
    indices: np.ndarray     # int32 [A], ascending within a row
    scores: np.ndarray      # uint16 [A]
    K: int
    row_cell: np.ndarray    # int32 [N], cell id of each row (non-decreasing)
    n_cells: int
    cell_type: np.ndarray   # int32 [n_cells], celltype id of each cell
    n_celltypes: int
    row_umi: np.ndarray     # int32 [N], UMI id (unique within a cell unless duplicated)

    @property
    def N(self):
        return len(self.indptr) - 1

    @property
    def A(self):
        return int(self.indptr[-1])

    def cell_rows(self):
        """list of row-id arrays, one per cell (bcode_ridx_map order)."""
        bounds = np.searchsorted(self.row_cell, np.arange(self.n_cells + 1))
        return [np.arange(bounds[c], bounds[c + 1]) for c in range(self.n_cells)]

    def celltype_rows(self):
        ct_of_row = self.cell_type[self.row_cell]
        return [np.flatnonzero(ct_of_row == t) for t in range(self.n_celltypes)]


def make_families(n_loci, rng):
    """Partition loci 1..n_loci into families of 4..64 members."""
    sizes = []
    left = n_loci
    while left > 0:
        s = int(min(left, rng.integers(4, 65)))
        if left - s < 4 and left - s > 0:
            s = left
        sizes.append(s)
        left -= s
    sizes = np.array(sizes, dtype=np.int64)
    start = np.concatenate([[1], 1 + np.cumsum(sizes)[:-1]])
    return start, sizes


def generate(n_cells=100, n_alignments=50_000, n_loci=28_000, seed=0,
             unique_frac=0.0, tie_frac=0.05, nofeat_frac=0.03, n_celltypes=1,
             umi_dup_rate=0.0, sigma=1.0, max_k=32):
    rng = np.random.default_rng(seed)
    K = n_loci + 1
    fstart, fsize = make_families(n_loci, rng)
    F = len(fsize)
    has_mirror = fsize >= 6
    draw_size = np.where(has_mirror, fsize - 1, fsize)      # members drawable on their own
    pop = 1.0 / np.arange(1, F + 1) ** 1.1
    pop = pop[rng.permutation(F)]
    pop /= pop.sum()

    # expected alignments per row -> number of rows
    mean_k = (1 - unique_frac) * 3.5 + unique_frac * 1.0 + nofeat_frac + 0.08
    N = max(n_cells, int(round(n_alignments / mean_k)))

    # cell sizes (rows per cell), every cell gets at least one row
    wts = rng.lognormal(0.0, sigma, n_cells)
    rows_per_cell = np.maximum(1, np.floor(wts / wts.sum() * N).astype(np.int64))
    N = int(rows_per_cell.sum())
    row_cell = np.repeat(np.arange(n_cells, dtype=np.int32), rows_per_cell)

    fam = rng.choice(F, size=N, p=pop)
    m = draw_size[fam]
    k = 2 + rng.poisson(1.5, N)
    k = np.minimum(np.minimum(k, max_k), m)
    if unique_frac > 0:
        k = np.where(rng.random(N) < unique_frac, 1, k)
    start = rng.integers(0, 1 << 30, N) % m
    stride = _PRIMES[rng.integers(0, len(_PRIMES), N)]
    forced = has_mirror[fam] & (rng.random(N) < tie_frac)
    start = np.where(forced, 0, start)                       # true locus := member 0

    ptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(k, out=ptr[1:])
    A0 = int(ptr[-1])
    erow = np.repeat(np.arange(N, dtype=np.int64), k)
    t = np.arange(A0, dtype=np.int64) - ptr[erow]            # slot within the row
    member = (start[erow] + t * stride[erow]) % m[erow]
    col = fstart[fam[erow]] + member
    p = np.where(t == 0, 0.75, 0.65)
    score = 140 + rng.binomial(120, p)

    rows_l, cols_l, sc_l = [erow], [col], [score]
    # mirrors of member 0
    mir = (member == 0) & has_mirror[fam[erow]]
    rows_l.append(erow[mir])
    cols_l.append(fstart[fam[erow[mir]]] + fsize[fam[erow[mir]]] - 1)
    sc_l.append(score[mir])
    # __no_feature
    nf = np.flatnonzero(rng.random(N) < nofeat_frac)
    rows_l.append(nf)
    cols_l.append(np.zeros(len(nf), dtype=np.int64))
    sc_l.append(140 + rng.binomial(120, 0.65, len(nf)))

    rows = np.concatenate(rows_l)
    cols = np.concatenate(cols_l)
    sc = np.concatenate(sc_l)
    order = np.lexsort((cols, rows))
    rows, cols, sc = rows[order], cols[order], sc[order]
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=N), out=indptr[1:])

    # celltypes: Dirichlet(0.5) proportions over cells
    if n_celltypes > 1:
        props = rng.dirichlet(np.full(n_celltypes, 0.5))
        props = np.maximum(props, 1.0 / (10 * n_cells))
        cell_type = rng.choice(n_celltypes, size=n_cells, p=props / props.sum()).astype(np.int32)
        cell_type[:min(n_celltypes, n_cells)] = np.arange(min(n_celltypes, n_cells))
    else:
        cell_type = np.zeros(n_cells, dtype=np.int32)

    # UMIs: unique per row, then a fraction of rows copies the UMI of the
    # previous row of the same cell (PCR duplicate)
    umi = np.arange(N, dtype=np.int32)
    if umi_dup_rate > 0:
        dup = rng.random(N) < umi_dup_rate
        dup[0] = False
        dup[1:] &= row_cell[1:] == row_cell[:-1]
        idx = np.where(dup, 0, np.arange(N))
        np.maximum.accumulate(idx, out=idx)
        umi = umi[idx]

    return SynthData(indptr=indptr.astype(np.int32), indices=cols.astype(np.int32),
                     scores=sc.astype(np.uint16), K=K, row_cell=row_cell,
                     n_cells=n_cells, cell_type=cell_type, n_celltypes=max(1, n_celltypes),
                     row_umi=umi)


def umi_dedup_mask(data: SynthData):
    """Row mask of a simple UMI de-duplication (keep the first row of each
    (cell, UMI) group).  The reference picks a representative per connected
    component (model.py:561-668, 1148-1219; out of scope here) and
    zeroes the other rows (``model.py:1216``); for the synthetic data "first
    row" plays that role.  Returns bool[N], True = row kept."""
    keep = np.ones(data.N, dtype=bool)
    keep[1:] = ~((data.row_umi[1:] == data.row_umi[:-1]) &
                 (data.row_cell[1:] == data.row_cell[:-1]))
    return keep


def apply_row_mask(data: SynthData, keep):
    """Zero the masked rows (their entries disappear, the row stays)."""
    lens = np.diff(data.indptr).astype(np.int64)
    ek = np.repeat(keep, lens)
    new_len = np.where(keep, lens, 0)
    ptr = np.zeros(data.N + 1, dtype=np.int64)
    np.cumsum(new_len, out=ptr[1:])
    return SynthData(indptr=ptr.astype(np.int32), indices=data.indices[ek],
                     scores=data.scores[ek], K=data.K, row_cell=data.row_cell,
                     n_cells=data.n_cells, cell_type=data.cell_type,
                     n_celltypes=data.n_celltypes, row_umi=data.row_umi)


def generate_device(n_cells=100, n_alignments=50_000, n_loci=28_000, seed=0, unique_frac=0.0, tie_frac=0.05,
                    nofeat_frac=0.03, n_celltypes=1, umi_dup_rate=0.0, sigma=1.0, max_k=32, device=None,
                    keep_device=False):
    """``generate`` for the large configurations (BASELINE.json configs[2..4]: 10^8 .. 5x10^8 alignments), with
    the per-row / per-alignment draws made by torch on the GPU: numpy needs ~3 minutes for 5x10^8 alignments, the
    device a few seconds.  Same model of the data, same small host-side draws (families, popularity, cell sizes,
    celltypes -- from ``numpy.random.default_rng(seed)``); the large draws come from a ``torch.Generator`` seeded
    with ``seed``, so the matrix is a different sample of the same distribution than ``generate`` gives, and it is
    reproducible on the same GPU model (all ranks of a bench run generate identical data; ``bench.py`` checks a
    checksum).  Benchmark input only: nothing of the product depends on it.

    Returns a ``SynthData`` of host arrays; with ``keep_device`` also ``.dev = (indptr, indices, scores)`` device
    tensors (int32, int32, uint16)."""
    import torch
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    rng = np.random.default_rng(seed)
    K = n_loci + 1
    fstart, fsize = make_families(n_loci, rng)
    F = len(fsize)
    has_mirror = fsize >= 6
    draw_size = np.where(has_mirror, fsize - 1, fsize)
    pop = 1.0 / np.arange(1, F + 1) ** 1.1
    pop = pop[rng.permutation(F)]
    pop /= pop.sum()
    mean_k = (1 - unique_frac) * 3.5 + unique_frac * 1.0 + nofeat_frac + 0.08
    N = max(n_cells, int(round(n_alignments / mean_k)))
    wts = rng.lognormal(0.0, sigma, n_cells)
    rows_per_cell = np.maximum(1, np.floor(wts / wts.sum() * N).astype(np.int64))
    N = int(rows_per_cell.sum())
    if n_celltypes > 1:
        props = rng.dirichlet(np.full(n_celltypes, 0.5))
        props = np.maximum(props, 1.0 / (10 * n_cells))
        cell_type = rng.choice(n_celltypes, size=n_cells, p=props / props.sum()).astype(np.int32)
        cell_type[:min(n_celltypes, n_cells)] = np.arange(min(n_celltypes, n_cells))
    else:
        cell_type = np.zeros(n_cells, dtype=np.int32)

    g = torch.Generator(device=dev)
    g.manual_seed(int(seed) + 0x5eed)
    i64 = torch.int64
    import contextlib
    with (torch.cuda.device(dev) if dev.type == 'cuda' else contextlib.nullcontext()):
        t_fstart = torch.from_numpy(fstart).to(dev)
        t_fsize = torch.from_numpy(fsize).to(dev)
        t_mirror = torch.from_numpy(has_mirror).to(dev)
        t_draw = torch.from_numpy(draw_size).to(dev)
        cdf = torch.from_numpy(np.cumsum(pop)).to(dev)
        fam = torch.searchsorted(cdf, torch.rand(N, generator=g, device=dev, dtype=torch.float64)).clamp_(max=F - 1)
        m = t_draw[fam]
        k = 2 + torch.poisson(torch.full((N,), 1.5, device=dev), generator=g).to(i64)
        k = torch.minimum(k.clamp_(max=max_k), m)
        if unique_frac > 0:
            k = torch.where(torch.rand(N, generator=g, device=dev) < unique_frac, torch.ones_like(k), k)
        start = torch.randint(0, 1 << 30, (N,), generator=g, device=dev) % m
        stride = torch.from_numpy(_PRIMES).to(dev)[torch.randint(0, len(_PRIMES), (N,), generator=g, device=dev)]
        forced = t_mirror[fam] & (torch.rand(N, generator=g, device=dev) < tie_frac)
        start = torch.where(forced, torch.zeros_like(start), start)
        ptr = torch.zeros(N + 1, dtype=i64, device=dev)
        torch.cumsum(k, 0, out=ptr[1:])
        A0 = int(ptr[-1])
        erow = torch.repeat_interleave(torch.arange(N, device=dev), k, output_size=A0)
        t = torch.arange(A0, device=dev) - ptr[erow]
        member = (start[erow] + t * stride[erow]) % m[erow]
        efam = fam[erow]
        col = t_fstart[efam] + member
        p = torch.where(t == 0, 0.75, 0.65).to(torch.float32)
        del t
        score = 140 + torch.binomial(torch.full((A0,), 120.0, device=dev), p, generator=g).to(i64)
        del p
        # one sortable key per alignment: row | column | score (51 bits)
        keys = [(erow << 23) | (col << 8) | (score - 140)]
        mir = (member == 0) & t_mirror[efam]
        mr = erow[mir]
        mf = efam[mir]
        keys.append((mr << 23) | ((t_fstart[mf] + t_fsize[mf] - 1) << 8) | (score[mir] - 140))
        del member, efam, col, mir, mr, mf, erow, score
        nf = torch.nonzero(torch.rand(N, generator=g, device=dev) < nofeat_frac).reshape(-1)
        nfs = torch.binomial(torch.full((len(nf),), 120.0, device=dev), torch.full((len(nf),), 0.65, device=dev),
                             generator=g).to(i64)
        keys.append((nf << 23) | nfs)
        key = torch.cat(keys)
        del keys
        key, _ = torch.sort(key)
        rows = key >> 23
        indices = ((key >> 8) & 0x7fff).to(torch.int32)
        scores = ((key & 0xff) + 140).to(torch.int32)
        del key
        indptr = torch.zeros(N + 1, dtype=i64, device=dev)
        torch.cumsum(torch.bincount(rows, minlength=N), 0, out=indptr[1:])
        del rows
        if int(indptr[-1]) >= 2 ** 31:
            raise OverflowError('more than 2^31 - 1 alignments: the score matrix uses int32 row pointers')
        indptr = indptr.to(torch.int32)
        row_cell_t = torch.repeat_interleave(torch.arange(n_cells, device=dev, dtype=torch.int32),
                                             torch.from_numpy(rows_per_cell).to(dev), output_size=N)
        umi = torch.arange(N, dtype=torch.int32, device=dev)
        if umi_dup_rate > 0:
            dup = torch.rand(N, generator=g, device=dev) < umi_dup_rate
            dup[0] = False
            dup[1:] &= row_cell_t[1:] == row_cell_t[:-1]
            idx = torch.where(dup, torch.zeros_like(umi), umi)
            umi = torch.cummax(idx, 0).values.to(torch.int32)
        scores16 = scores.to(torch.uint16)
        del scores
        out = SynthData(indptr=indptr.cpu().numpy(), indices=indices.cpu().numpy(), scores=scores16.cpu().numpy(),
                        K=K, row_cell=row_cell_t.cpu().numpy(), n_cells=n_cells, cell_type=cell_type,
                        n_celltypes=max(1, n_celltypes), row_umi=umi.cpu().numpy())
        if keep_device:
            out.dev = (indptr, indices, scores16)
            out.dev_row_cell = row_cell_t
    return out
