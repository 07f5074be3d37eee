"""Multi-GPU drivers: one process per GPU (torchrun), torch.distributed for the plumbing.

* individual / celltype pooling: the models are independent (reference
  ``_fit_pooling_model`` model.py:701-743 fits them one after another or in a
  ``multiprocessing.Pool``, :793-801) -> shard MODELS across ranks by
  longest-processing-time bin packing on their alignment counts; every rank
  builds the layout of its own rows and fits; no collective on the data path,
  only a final gather of the per-model records.
* pseudobulk: one model -> ROWS are sharded and the K-vector of theta sums is
  all-reduced once per iteration (see ``fit_pseudobulk_sharded``).
"""
from __future__ import annotations

import heapq

import numpy as np


def lpt_shards(sizes, world):
    """Longest-processing-time assignment of items (models) to ``world`` bins.
    Deterministic (ties by index).  Returns a list of sorted index arrays."""
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.lexsort((np.arange(len(sizes)), -sizes))
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    bins = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        bins[r].append(int(i))
        heapq.heappush(heap, (load + int(sizes[i]), r))
    return [np.array(sorted(b), dtype=np.int64) for b in bins]


def model_cell_block(lens, rows, row_cell, rank, world):
    """Like ``model_row_block`` but whole CELLS go to a rank (``row_cell[i]`` = cell of row i; the rows of a cell
    need not be contiguous): the cells of the model, in id order, are cut into ``world`` consecutive groups of
    about equal alignment count.  Needed when the counting is per (cell, UMI) -- ``--umi_counts`` -- because the
    distinct UMIs of a cell can only be counted where all its reads are ("keep a cell's rows
    together so counts stay rank-local")."""
    rows = np.asarray(rows, dtype=np.int64)
    if len(rows) == 0:
        return rows
    cells, inv = np.unique(np.asarray(row_cell)[rows], return_inverse=True)
    a_cell = np.bincount(inv, weights=np.asarray(lens, dtype=np.int64)[rows], minlength=len(cells))
    mid = np.cumsum(a_cell) - 0.5 * a_cell                        # midpoint rule: monotone -> consecutive groups
    cell_rank = np.minimum(world - 1, (mid * world / max(float(a_cell.sum()), 1.0)).astype(np.int64))
    return rows[cell_rank[inv] == rank]


def extract_rows(indptr, indices, data, rows):
    """Row-compacted CSR of ``rows`` (ascending)."""
    rows = np.asarray(rows, dtype=np.int64)
    lens = (indptr[rows + 1] - indptr[rows]).astype(np.int64)
    ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    take = np.repeat(indptr[rows].astype(np.int64) - ptr[:-1], lens) + np.arange(ptr[-1])
    return ptr, indices[take], data[take], take


def shard_models(indptr, row_segment, n_segments, world, wide=None, wide_cost=2_000_000, wide_gain=0.03):
    """(per-rank model index arrays, alignments per model) for a row -> model map.

    With ``wide`` a third value is returned, the WIDE models: ``'auto'`` chooses them (``wide_models``), a list
    names them.  They are left out of the per-rank lists: their ROWS are cut into ``world`` blocks and fitted
    by all ranks together like a pseudobulk model (celltype row)."""
    lens = np.diff(indptr).astype(np.int64)
    fitted = row_segment >= 0
    seg_A = np.bincount(row_segment[fitted], weights=lens[fitted], minlength=n_segments).astype(np.int64)
    if wide is None:
        return lpt_shards(seg_A, world), seg_A
    if isinstance(wide, str):
        wide_ids = wide_models(seg_A, world, wide_cost, wide_gain)
    else:
        wide_ids = np.array(sorted(int(w) for w in wide), dtype=np.int64)
        if world < 2:
            wide_ids = wide_ids[:0]
    narrow = np.setdiff1d(np.arange(n_segments, dtype=np.int64), wide_ids)
    shards = [narrow[b] for b in lpt_shards(seg_A[narrow], world)]
    return shards, seg_A, wide_ids


def makespan(seg_A, world, wide, cost):
    """Load of the fullest rank in alignments: its packed models (LPT) plus, for every wide model, 1/world of
    its alignments and ``cost`` -- the price of fitting in lockstep with the other ranks (one exchange and two
    grid-wide barriers per iteration, ~20 us x ~80 iterations, expressed in alignments of single-GPU work)."""
    seg_A = np.asarray(seg_A, dtype=np.int64)
    wide = np.asarray(wide, dtype=np.int64)
    narrow = np.setdiff1d(np.arange(len(seg_A)), wide)
    packed = max((float(seg_A[narrow[b]].sum()) for b in lpt_shards(seg_A[narrow], world)), default=0.0)
    return packed + float(seg_A[wide].sum()) / world + float(cost) * len(wide)


def wide_models(seg_A, world, cost=2_000_000, gain=0.03):
    """Models that are fitted by all ranks together instead of being packed onto one rank (ascending indices).

    Packing whole models (LPT) cannot balance a model that is much larger than a rank's fair share -- a
    dominant cell type on 8 GPUs.  Greedy on ``makespan``: the largest packed model becomes wide as long as
    that shortens the fullest rank's load by more than ``gain`` (relative).  ``cost=None`` never splits."""
    seg_A = np.asarray(seg_A, dtype=np.int64)
    if world < 2 or cost is None or len(seg_A) == 0:
        return np.zeros(0, dtype=np.int64)
    order = np.lexsort((np.arange(len(seg_A)), -seg_A))           # candidates, largest first
    wide = []
    now = makespan(seg_A, world, wide, cost)
    for cand in order:
        new = makespan(seg_A, world, wide + [int(cand)], cost)
        if not new < (1.0 - gain) * now:
            break
        wide.append(int(cand))
        now = new
    return np.array(sorted(wide), dtype=np.int64)


def model_row_block(lens, rows, rank, world):
    """The block of ``rows`` (ascending row ids of ONE model) that ``rank`` fits when the model's rows are cut
    into ``world`` consecutive blocks of about equal alignment count."""
    rows = np.asarray(rows, dtype=np.int64)
    ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(np.asarray(lens, dtype=np.int64)[rows], out=ptr[1:])
    a, b = row_block(ptr, rank, world)
    return rows[a:b]


def local_problem(indptr, indices, data, row_segment, my_models):
    """Rows and local model ids of one rank.  Returns (rows, ptr, idx, dat, take, local_seg)."""
    remap = np.full(int(row_segment.max()) + 2 if len(row_segment) else 1, -1, dtype=np.int64)
    remap[my_models] = np.arange(len(my_models))
    seg_local_all = np.where(row_segment >= 0, remap[np.maximum(row_segment, 0)], -1)
    rows = np.flatnonzero(seg_local_all >= 0)
    ptr, idx, dat, take = extract_rows(indptr, indices, data, rows)
    return rows, ptr, idx, dat, take, seg_local_all[rows].astype(np.int32)


def fit_models_sharded(indptr, indices, data, K, row_segment, n_segments, fit_local, group=None):
    """Fit independent models across the ranks of ``group``.

    ``fit_local(ptr, idx, dat, K, local_row_segment, n_local)`` fits this rank's
    models and returns a list of per-model records (any picklable object) in
    local model order.  Returns, on every rank, the list of records of ALL
    models in global model order, plus this rank's (rows, take) so that the
    caller can place row-wise results (z) back.
    """
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    shards, _seg_A = shard_models(indptr, row_segment, n_segments, world)
    mine = shards[rank]
    rows, ptr, idx, dat, take, seg_local = local_problem(indptr, indices, data, row_segment, mine)
    local_records = fit_local(ptr, idx, dat, K, seg_local, len(mine))
    if len(local_records) != len(mine):
        raise RuntimeError('fit_local returned a wrong number of records')
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine.tolist(), local_records), group=group)
    else:
        gathered = [(mine.tolist(), local_records)]
    out = [None] * n_segments
    for models, recs in gathered:
        for m, r in zip(models, recs):
            out[m] = r
    return out, rows, take


def row_blocks(indptr, row_cell, world):
    """Pseudobulk sharding: contiguous row blocks of about equal alignment count
    whose boundaries fall between cells (a cell's rows stay on one rank, so the
    per-cell counts stay rank-local).  Returns ``world + 1`` row boundaries."""
    N = len(indptr) - 1
    A = int(indptr[-1])
    cell_start = np.flatnonzero(np.concatenate([[True], row_cell[1:] != row_cell[:-1]]))
    cand = np.concatenate([cell_start, [N]])
    bounds = [0]
    for r in range(1, world):
        target = A * r // world
        j = int(np.argmin(np.abs(indptr[cand].astype(np.int64) - target)))
        bounds.append(max(bounds[-1], int(cand[j])))
    bounds.append(N)
    return np.array(bounds, dtype=np.int64)


def row_block(indptr, rank, world):
    """Row range [r0, r1) of ``rank`` when the rows of one model are cut into ``world`` contiguous
    blocks of about equal alignment count (pseudobulk sharding without cell information: per-cell
    counts of the rank-local rows are added up over the ranks afterwards)."""
    N = len(indptr) - 1
    A = int(indptr[-1])
    targets = (A * np.arange(world + 1, dtype=np.int64)) // world
    bounds = np.searchsorted(np.asarray(indptr, dtype=np.int64), targets, side='left')
    bounds[0], bounds[-1] = 0, N
    bounds = np.maximum.accumulate(np.minimum(bounds, N))
    return int(bounds[rank]), int(bounds[rank + 1])


# ---------------------------------------------------------------------------
# Rank-local rows of a sharded run and the gathers of its results
# ---------------------------------------------------------------------------
def plan_local_rows(indptr, row_segment, n_segments, world, rank, wide=None, wide_cost=2_000_000, row_cell=None,
                    cells_whole=False, all_assigned=None, collective=False):
    """Which rows this rank holds when the models of a pooling mode are sharded over ``world`` ranks.

    Every row of the matrix goes to exactly ONE rank, so that whatever is computed per row (z, the reassignment
    matrices, the per-cell counts, the ``ReassignInfo`` counters) adds up over the ranks to the single-GPU result:

    * rows of a packed model -> the rank the model was packed onto (LPT, ``shard_models``);
    * rows of a WIDE model -> cut into ``world`` blocks of about equal alignment count (whole cells per block
      with ``cells_whole``, i.e. ``--umi_counts``);
    * rows of no model (barcodes without a cell type, UMI-dropped rows: never fitted, but still counted by the
      modes that do not use z -- model.py:468-502) -> spread by cell (``row_cell``) or, without it, in
      ``world`` contiguous blocks.

    Returns a dict: ``rows`` (ascending global row ids of this rank, int64), ``local_seg`` (int32 per local row:
    index into ``mine`` or -1), ``mine`` (global ids of the models packed onto this rank), ``wide`` (global ids of
    the wide models), ``wide_local`` (per wide model: bool mask over the LOCAL rows = this rank's block),
    ``seg_A`` (alignments per model).  ``collective``: the call is made by all ranks of the default process group
    at the same time (the per-model alignment counts are then computed in 1/world slices and all-reduced)."""
    indptr = np.asarray(indptr)
    row_segment = np.asarray(row_segment)
    N = len(row_segment)
    no_wide = wide is None or (not isinstance(wide, str) and len(wide) == 0)
    if all_assigned is None:
        all_assigned = N == 0 or int(row_segment.min()) >= 0
    if no_wide and row_segment.dtype == np.int32 and all_assigned:
        # the common case -- every row belongs to a packed model: three threaded passes in C instead of ~10 numpy
        # passes over per-row arrays (1.4x10^8 rows at atlas scale)
        from . import hostutil
        if N > 0 and _sorted_by_segment(row_segment, world, rank, collective):
            # the rows of every model are consecutive (reads of a barcode-sorted BAM, `stellarscope cellsort`; the row ->
            # model map is then non-decreasing): the plan needs no pass over the rows at all -- the first row of every
            # model by binary search, its alignments from the row pointers there, and the rank's rows as ranges
            bounds = np.searchsorted(row_segment, np.arange(n_segments + 1, dtype=row_segment.dtype)).astype(np.int64)
            seg_A = (indptr[bounds[1:]].astype(np.int64) - indptr[bounds[:-1]].astype(np.int64))
            model_rank = hostutil.lpt_assign(seg_A, world)
            mine = np.flatnonzero(model_rank == rank).astype(np.int64)
            rows, local_seg = hostutil.rows_of_ranges(bounds[mine], bounds[mine + 1])
            return dict(rows=rows, local_seg=local_seg, mine=mine, wide=np.zeros(0, dtype=np.int64), wide_local=[],
                        seg_A=seg_A)
        if collective and world > 1:
            # every rank counts the alignments per model over 1/world of the rows; the sums are all-reduced
            r0, r1 = N * rank // world, N * (rank + 1) // world
            part = hostutil.segment_alignments(indptr[r0:r1 + 1], row_segment[r0:r1], n_segments)
            seg_A = allreduce_sum_host(part).astype(np.int64)
        else:
            seg_A = hostutil.segment_alignments(indptr, row_segment, n_segments)
        model_rank = hostutil.lpt_assign(seg_A, world)
        mine = np.flatnonzero(model_rank == rank).astype(np.int64)
        local_id = np.full(n_segments, -1, dtype=np.int32)
        local_id[mine] = np.arange(len(mine), dtype=np.int32)
        rows, local_seg = hostutil.rows_of_rank(row_segment, model_rank, local_id, rank)
        return dict(rows=rows, local_seg=local_seg, mine=mine, wide=np.zeros(0, dtype=np.int64), wide_local=[],
                    seg_A=seg_A)
    shards, seg_A, wide_ids = shard_models(indptr, row_segment, n_segments, world,
                                           wide=wide if wide is not None else [], wide_cost=wide_cost)
    mine = np.asarray(shards[rank], dtype=np.int64)
    model_rank = np.full(n_segments + 1, -1, dtype=np.int32)
    for r, b in enumerate(shards):
        model_rank[b] = r
    owner = model_rank[np.where(row_segment >= 0, row_segment, n_segments)]      # -1: wide model or no model
    lens = None
    wide_rows = []
    for w in wide_ids:
        rows_w = np.flatnonzero(row_segment == w)
        if lens is None:
            lens = np.diff(indptr).astype(np.int64)
        if cells_whole and row_cell is not None:
            blk = model_cell_block(lens, rows_w, row_cell, rank, world)
        else:
            blk = model_row_block(lens, rows_w, rank, world)
        wide_rows.append(blk)
        owner[rows_w] = -2                                   # not "no model"
        owner[blk] = rank
    loose = np.flatnonzero(owner == -1)
    if len(loose):
        if row_cell is not None:
            c = np.asarray(row_cell)[loose]
            owner[loose] = np.where(c >= 0, c % world, 0).astype(np.int32)
        else:
            owner[loose] = np.minimum(world - 1, (np.arange(len(loose), dtype=np.int64) * world) // len(loose))
    rows = np.flatnonzero(owner == rank).astype(np.int64)
    remap = np.full(n_segments + 1, -1, dtype=np.int32)
    remap[mine] = np.arange(len(mine), dtype=np.int32)
    seg_rows = row_segment[rows]
    local_seg = remap[np.where(seg_rows >= 0, seg_rows, n_segments)].astype(np.int32)
    wide_local = []
    for blk in wide_rows:
        m = np.zeros(len(rows), dtype=bool)
        m[np.searchsorted(rows, blk)] = True
        wide_local.append(m)
    return dict(rows=rows, local_seg=local_seg, mine=mine, wide=np.asarray(wide_ids, dtype=np.int64),
                wide_local=wide_local, seg_A=seg_A)


def _sorted_by_segment(row_segment, world, rank, collective):
    """True when the row -> model map is non-decreasing.  ``collective``: every rank looks at 1/world of the rows (and
    the first row of the next slice) and the verdicts are combined -- all ranks get the same answer."""
    N = len(row_segment)
    if collective and world > 1:
        r0, r1 = N * rank // world, min(N, N * (rank + 1) // world + 1)
        part = row_segment[r0:r1]
        bad = int(len(part) > 1 and bool(np.any(part[1:] < part[:-1])))
        return int(allreduce_sum_host(np.array([bad], dtype=np.int64))[0]) == 0
    step = 1 << 24                                            # in pieces: no temporary of the size of the map
    for a in range(0, max(N - 1, 0), step):
        b = min(N, a + step + 1)
        if np.any(row_segment[a + 1:b] < row_segment[a:b - 1]):
            return False
    return True


def _comm_device(group=None):
    """Device the collectives of ``group`` take their tensors on (NCCL: the current CUDA device; gloo: the host)."""
    import torch
    import torch.distributed as dist
    if dist.get_backend(group) == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def gather_concat(arrays, group=None):
    """All-gather of variable-length 1-D (or row-wise 2-D) numpy arrays: every rank gets, per name, the
    concatenation over the ranks in rank order, plus ``counts`` (rows contributed by every rank).  All arrays of
    one call have the same number of rows on a rank.  The payload travels as tensors (``all_gather_into_tensor``
    of buffers padded to the largest contribution) -- no pickling (results of 10^5 models / 10^8 count entries)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    names = list(arrays)
    n = len(arrays[names[0]]) if names else 0
    dev = _comm_device(group)
    cnt = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnt, mine, group=group)
    counts = cnt.cpu().numpy()
    nmax = int(counts.max()) if world else 0
    out = {}
    for k in names:
        a = np.ascontiguousarray(arrays[k])
        if len(a) != n:
            raise ValueError('gather_concat: arrays of one call must have the same length')
        tail = tuple(a.shape[1:])
        width = int(np.prod(tail, dtype=np.int64)) * a.dtype.itemsize       # bytes per row: travels as uint8
        buf = torch.zeros(max(nmax * width, 1), dtype=torch.uint8, device=dev)
        if n:
            buf[:n * width].copy_(torch.from_numpy(a.reshape(-1).view(np.uint8)))
        allb = torch.empty(world * max(nmax * width, 1), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allb, buf, group=group)
        h = allb.cpu().numpy().reshape(world, max(nmax * width, 1))
        parts = [h[r, :int(counts[r]) * width] for r in range(world)]
        out[k] = np.concatenate(parts).view(a.dtype).reshape((-1,) + tail)
    out['counts'] = counts
    return out


def allreduce_sum_host(a, group=None):
    """Sum of a small host array over the ranks (int64 / float64), returned as numpy."""
    import torch
    import torch.distributed as dist
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a.copy()).to(_comm_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()
