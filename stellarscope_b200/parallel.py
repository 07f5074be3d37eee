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


def extract_rows(indptr, indices, data, rows):
    """Row-compacted CSR of ``rows`` (ascending)."""
    rows = np.asarray(rows, dtype=np.int64)
    lens = (indptr[rows + 1] - indptr[rows]).astype(np.int64)
    ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    take = np.repeat(indptr[rows].astype(np.int64) - ptr[:-1], lens) + np.arange(ptr[-1])
    return ptr, indices[take], data[take], take


def shard_models(indptr, row_segment, n_segments, world):
    """(per-rank model index arrays, alignments per model) for a row -> model map."""
    lens = np.diff(indptr).astype(np.int64)
    fitted = row_segment >= 0
    seg_A = np.bincount(row_segment[fitted], weights=lens[fitted], minlength=n_segments).astype(np.int64)
    return lpt_shards(seg_A, world), seg_A


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
