"""Host (numpy) builder of the tiled segment layout -- the *format specification*.

The product builds this layout on the device (``csrc/ss_layout.cu``,
``ss_layout_build``).  This module restates the same construction with numpy so
that (a) the device builder can be checked array-for-array and (b) the EM
kernels can be exercised against a layout whose provenance is independent of
the build kernels (``ss_layout_import``, debug/test entry point).  It is NOT
used by ``fit_pooling_model``.

Layout (DESIGN.md "Data layout in HBM"):

segment  = one EM model (a cell, a celltype, or everything: reference
           ``_fit_pooling_model`` model.py:675-804)
tile     = <= E_CAP alignments of consecutive *ambiguous* rows (>= 2
           alignments, model.py:141) of one segment.  Rows inside a tile are
           in jagged-diagonal (JDS) order: sorted by length descending, the
           p-th alignment of JDS row r sits at ``jptr[p] + r``.
tile col = the distinct loci a tile touches, ordered by (count desc, locus
           asc); ``cpos`` gives every alignment its slot in column-major order
           (``cptr[c] + rank of its row among the rows of column c``).
seg col  = the distinct loci a segment touches (ambiguous + unique rows),
           ascending; ``scol`` maps a tile column to its absolute slot.

Unique rows (exactly one alignment) never enter a tile: they are folded into
``pisum0`` (model.py:191), a per-column count and a per-segment sum of log Q.
"""
from __future__ import annotations

from dataclasses import dataclass

import os

import numpy as np

E_HARD = 4096        # hard cap on alignments per tile (shared-memory budget, csrc/ss_common.cuh)
E_STREAM = 2048      # tile size of streamed segments (more than STREAM_SEG_TILES large tiles)
STREAM_SEG_TILES = 8
E_CLUSTER = 3072     # target tile size of the segments that run as thread-block clusters (csrc/ss_common.cuh)
ALIGN = 8            # every per-tile slice starts at a multiple of 8 elements


FRAG, FRAG_MAXG_ROW, FRAG_MAXG_COL = 4, 8, 32     # csrc/ss_common.cuh
HDR_WORDS = 24       # int64 words per tile header (csrc/ss_common.cuh)


def frag_group(n):
    f = (n + FRAG - 1) // FRAG
    g = 1
    while g < f:
        g <<= 1
    return g


def _pad(n, a=ALIGN):
    return (n + a - 1) // a * a


@dataclass
class HostLayout:
    # scalars
    N: int
    K: int
    A: int
    S: int
    n_tiles: int
    chunk: int
    maxlen: int
    # tile headers (int64 [n_tiles, 24]): seg, nrows, nent, ncol, maxlen, nlong, ent_off, row_off, col_off, jds_off,
    # rfrag, cfrag (lanes the rows / columns occupy in the lane kernel; longer ones excluded),
    # rows longer than 4/8/16/32, columns with more than 4/8/16/32/64/128 rows, 2 spare words
    tile_hdr: np.ndarray
    # entry arrays
    e_lcol: np.ndarray   # u16
    e_cpos: np.ndarray   # u16
    e_q: np.ndarray      # f64
    e_opos: np.ndarray   # i32  original CSR position
    # row arrays
    r_len: np.ndarray    # u16
    r_w: np.ndarray      # f64
    r_orow: np.ndarray   # i32  original row id
    # tile column arrays
    c_ptr: np.ndarray    # u16  (ncol+1 per tile)
    c_scol: np.ndarray   # i32  absolute segment-column slot
    c_pisum0: np.ndarray  # f64
    c_nuniq: np.ndarray  # i32
    j_ptr: np.ndarray    # u16  (maxlen+1 per tile)
    # segment arrays
    seg_tile0: np.ndarray   # i32
    seg_ntiles: np.ndarray  # i32
    seg_col0: np.ndarray    # i32
    seg_ncol: np.ndarray    # i32
    seg_smax: np.ndarray    # i32
    seg_wmax: np.ndarray    # f64
    seg_wtot: np.ndarray    # f64
    seg_wamb: np.ndarray    # f64
    seg_logq_uniq: np.ndarray  # f64
    seg_nobs: np.ndarray    # i32
    seg_namb: np.ndarray    # i32 ambiguous rows
    seg_nuniq: np.ndarray   # i32 unique rows
    seg_A: np.ndarray       # i64 all alignments
    seg_Aamb: np.ndarray    # i64 alignments in ambiguous rows
    # segment-column arrays
    sc_gcol: np.ndarray     # i32
    sc_pisum0: np.ndarray   # f64
    sc_nuniq: np.ndarray    # i32
    sc_inamb: np.ndarray    # i8   1 if some tile touches the column
    # unique rows
    u_row: np.ndarray       # i32 original row ids of unique rows of fitted segments
    u_pos: np.ndarray       # i32 CSR position of their single alignment

    LONG_COL = 32


def build(indptr, indices, scores, row_seg, S, K, e_hard=E_HARD, long_col=32):
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    scores = np.asarray(scores)
    row_seg = np.asarray(row_seg, dtype=np.int64)
    N = len(indptr) - 1
    A = int(indptr[-1])
    lens = np.diff(indptr)
    fitted = row_seg >= 0
    amb = fitted & (lens > 1)
    uni = fitted & (lens == 1)
    erow = np.repeat(np.arange(N, dtype=np.int64), lens)
    eseg = row_seg[erow]

    # ---- per segment score max and Q (model.py:115, 129-130) ----
    seg_smax = np.zeros(S, dtype=np.int64)
    ef = eseg >= 0
    np.maximum.at(seg_smax, eseg[ef], scores[ef].astype(np.int64))
    with np.errstate(divide='ignore'):
        inv = 1.0 / seg_smax.astype(np.float64)
    Q = np.zeros(A, dtype=np.float64)
    Q[ef] = np.expm1((scores[ef].astype(np.float64) * inv[eseg[ef]]) * 100.0)
    with np.errstate(invalid='ignore'):
        seg_wmax = np.where(seg_smax > 0,
                            np.expm1((seg_smax.astype(np.float64) * inv) * 100.0), 0.0)
    # row weights w_i = max_j Q_ij (model.py:183)
    w = np.zeros(N, dtype=np.float64)
    nz = np.flatnonzero(lens > 0)
    w[nz] = np.maximum.reduceat(Q, indptr[nz])
    w[~fitted] = 0.0
    seg_wtot = np.zeros(S)
    seg_wamb = np.zeros(S)
    for s_arr, mask in ((seg_wtot, amb | uni), (seg_wamb, amb)):
        rr = np.flatnonzero(mask)
        order = np.argsort(row_seg[rr], kind='stable')
        rr = rr[order]
        # sequential (row order) sum per segment
        np.add.at(s_arr, row_seg[rr], w[rr])
    seg_nobs = np.bincount(row_seg[amb | uni], minlength=S).astype(np.int32)
    seg_namb = np.bincount(row_seg[amb], minlength=S).astype(np.int32)
    seg_nuniq = np.bincount(row_seg[uni], minlength=S).astype(np.int32)
    seg_A = np.bincount(row_seg[fitted], weights=lens[fitted], minlength=S).astype(np.int64)
    seg_Aamb = np.bincount(row_seg[amb], weights=lens[amb], minlength=S).astype(np.int64)

    # ---- segment columns: distinct (seg, col) over all fitted entries ----
    key = eseg[ef] * K + indices[ef]
    ukey = np.unique(key)
    sc_seg = ukey // K
    sc_gcol = (ukey % K).astype(np.int32)
    seg_ncol = np.bincount(sc_seg, minlength=S).astype(np.int32)
    seg_col0 = np.zeros(S, dtype=np.int32)
    seg_col0[1:] = np.cumsum(seg_ncol)[:-1]
    # unique rows -> pisum0, nuniq, sum log Q   (row order accumulation)
    u_row = np.flatnonzero(uni)
    u_pos = indptr[u_row]
    sc_pisum0 = np.zeros(len(ukey))
    sc_nuniq = np.zeros(len(ukey), dtype=np.int32)
    seg_logq_uniq = np.zeros(S)
    if len(u_row):
        uslot = np.searchsorted(ukey, row_seg[u_row] * K + indices[u_pos])
        np.add.at(sc_pisum0, uslot, Q[u_pos])
        np.add.at(sc_nuniq, uslot, 1)
        np.add.at(seg_logq_uniq, row_seg[u_row], np.log(Q[u_pos]))

    # ---- tiles over ambiguous rows ----
    ar = np.flatnonzero(amb)
    # (segment, first locus, row) order: consecutive rows -- and the tiles of a multi-tile segment -- touch
    # a narrow band of loci (csrc/ss_build.cu row_sortkey_kernel); column 0 does not count as first locus
    first = indices[indptr[ar]].astype(np.int64)
    second = indices[np.minimum(indptr[ar] + 1, len(indices) - 1)].astype(np.int64) if len(ar) else first
    first = np.where(first == 0, second, first)
    ar = ar[np.lexsort((first, row_seg[ar]))]                # lexsort is stable
    alen = lens[ar]
    maxlen = int(alen.max()) if len(ar) else 0
    if maxlen > e_hard // 2:
        raise ValueError(f'a read with {maxlen} alignments exceeds the tile capacity')
    chunk = e_hard - maxlen + 1 if maxlen else e_hard
    aseg = row_seg[ar]
    pre = np.zeros(len(ar) + 1, dtype=np.int64)
    np.cumsum(alen, out=pre[1:])
    seg_row0 = np.searchsorted(aseg, np.arange(S))
    seg_ent0 = pre[np.minimum(seg_row0, len(ar))]
    o = pre[:-1] - seg_ent0[aseg] if len(ar) else np.zeros(0, dtype=np.int64)
    # segments that will be streamed are cut into smaller tiles (csrc/ss_build.cu seg_chunk)
    chunk_small = E_STREAM - maxlen + 1 if (e_hard == E_HARD and maxlen <= E_STREAM // 2) else chunk
    seg_aamb = np.bincount(aseg, weights=alen, minlength=S).astype(np.int64) if len(ar) else np.zeros(S, dtype=np.int64)
    small_thr = int(os.environ.get('SS_SMALL_TILE_THR', STREAM_SEG_TILES))      # experiment knob, like the device builder
    # segments of 2..small_thr large tiles (thread-block clusters) are cut into EQUAL tiles of about `target` alignments
    target = min(chunk, int(os.environ.get('SS_CLUSTER_TILE', E_CLUSTER)))
    nt_bal = np.minimum(-(-seg_aamb // target), small_thr)
    seg_chunk = np.where(seg_aamb > small_thr * chunk, chunk_small,
                         np.where(seg_aamb <= chunk, chunk, -(-seg_aamb // np.maximum(nt_bal, 1))))
    tin = o // seg_chunk[aseg] if len(ar) else o
    seg_ntiles = np.zeros(S, dtype=np.int32)
    if len(ar):
        np.maximum.at(seg_ntiles, aseg, (tin + 1).astype(np.int32))
    seg_tile0 = np.zeros(S, dtype=np.int32)
    seg_tile0[1:] = np.cumsum(seg_ntiles)[:-1]
    n_tiles = int(seg_ntiles.sum())
    rtile = seg_tile0[aseg] + tin if len(ar) else np.zeros(0, dtype=np.int64)

    hdr = np.zeros((n_tiles, HDR_WORDS), dtype=np.int64)
    sc_inamb = np.zeros(len(ukey), dtype=np.int8)
    ent_off = row_off = col_off = jds_off = 0
    E_l, P_l, Q_l, O_l = [], [], [], []
    RL_l, RW_l, RO_l = [], [], []
    CP_l, CS_l, CPS_l, CNU_l, J_l = [], [], [], [], []
    tile_row0 = np.searchsorted(rtile, np.arange(n_tiles + 1))
    for t in range(n_tiles):
        a, b = tile_row0[t], tile_row0[t + 1]
        rows = ar[a:b]
        rl = alen[a:b]
        s = int(aseg[a])
        order = np.argsort(-rl, kind='stable')               # (len desc, rank asc)
        rows, rl = rows[order], rl[order]
        nrows, nent, ml = len(rows), int(rl.sum()), int(rl[0])
        cnt = np.array([(rl > p).sum() for p in range(ml)], dtype=np.int64)
        jp = np.zeros(ml + 1, dtype=np.int64)
        np.cumsum(cnt, out=jp[1:])
        # entries in JDS order
        er = np.repeat(np.arange(nrows), rl)                 # JDS row of each CSR-order entry
        ep = np.arange(nent) - np.repeat(np.cumsum(rl) - rl, rl)   # slot in row
        opos = np.repeat(indptr[rows], rl) + ep
        e_idx = jp[ep] + er                                  # JDS position
        g = np.empty(nent, dtype=np.int64); g[e_idx] = indices[opos]
        qq = np.empty(nent); qq[e_idx] = Q[opos]
        oo = np.empty(nent, dtype=np.int64); oo[e_idx] = opos
        rr = np.empty(nent, dtype=np.int64); rr[e_idx] = er
        # tile columns
        ug, inv_c, cc = np.unique(g, return_inverse=True, return_counts=True)
        corder = np.lexsort((ug, -cc))                       # (count desc, gcol asc)
        rank_of = np.empty(len(ug), dtype=np.int64); rank_of[corder] = np.arange(len(ug))
        lcol = rank_of[inv_c]
        ccs = cc[corder]
        cptr = np.zeros(len(ug) + 1, dtype=np.int64)
        np.cumsum(ccs, out=cptr[1:])
        # slot inside the column: ascending JDS row
        o2 = np.lexsort((rr, lcol))
        cpos = np.empty(nent, dtype=np.int64)
        cpos[o2] = np.arange(nent)                           # == cptr[lcol] + rank
        slot = seg_col0[s] + np.searchsorted(sc_gcol[seg_col0[s]:seg_col0[s] + seg_ncol[s]],
                                             ug[corder])
        nlong = int((ccs > long_col).sum())
        sc_inamb[slot] = 1
        rfrag = int(sum(frag_group(int(x)) for x in rl if x <= FRAG * FRAG_MAXG_ROW))
        cfrag = int(sum(frag_group(int(x)) for x in ccs if x <= FRAG * FRAG_MAXG_COL))
        # number of rows longer than 4/8/16/32 alignments, of columns with more than 4..128 rows
        # (lane-group boundaries of the lane kernel, csrc/ss_em.cu)
        rgt = [int((rl > x).sum()) for x in (4, 8, 16, 32)]
        cgt = [int((ccs > x).sum()) for x in (4, 8, 16, 32, 64, 128)]
        hdr[t, :12] = (s, nrows, nent, len(ug), ml, nlong, ent_off, row_off, col_off, jds_off, rfrag, cfrag)
        hdr[t, 12:16] = rgt
        hdr[t, 16:22] = cgt

        def padded(x, n, dtype):
            out = np.zeros(n, dtype=dtype); out[:len(x)] = x; return out
        pe, pr = _pad(nent), _pad(nrows)
        pc, pj = _pad(len(ug) + 1), _pad(ml + 1)
        E_l.append(padded(lcol, pe, np.uint16)); P_l.append(padded(cpos, pe, np.uint16))
        Q_l.append(padded(qq, pe, np.float64)); O_l.append(padded(oo, pe, np.int32))
        RL_l.append(padded(rl, pr, np.uint16)); RW_l.append(padded(w[rows], pr, np.float64))
        RO_l.append(padded(rows, pr, np.int32))
        CP_l.append(padded(cptr, pc, np.uint16)); CS_l.append(padded(slot, pc, np.int32))
        CPS_l.append(padded(sc_pisum0[slot], pc, np.float64))
        CNU_l.append(padded(sc_nuniq[slot], pc, np.int32))
        J_l.append(padded(jp, pj, np.uint16))
        ent_off += pe; row_off += pr; col_off += pc; jds_off += pj

    def cat(lst, dtype):
        return np.concatenate(lst) if lst else np.zeros(0, dtype=dtype)

    return HostLayout(
        N=N, K=K, A=A, S=S, n_tiles=n_tiles, chunk=chunk, maxlen=maxlen, tile_hdr=hdr,
        e_lcol=cat(E_l, np.uint16), e_cpos=cat(P_l, np.uint16), e_q=cat(Q_l, np.float64),
        e_opos=cat(O_l, np.int32), r_len=cat(RL_l, np.uint16), r_w=cat(RW_l, np.float64),
        r_orow=cat(RO_l, np.int32), c_ptr=cat(CP_l, np.uint16), c_scol=cat(CS_l, np.int32),
        c_pisum0=cat(CPS_l, np.float64), c_nuniq=cat(CNU_l, np.int32), j_ptr=cat(J_l, np.uint16),
        seg_tile0=seg_tile0, seg_ntiles=seg_ntiles, seg_col0=seg_col0, seg_ncol=seg_ncol,
        seg_smax=seg_smax.astype(np.int32), seg_wmax=seg_wmax, seg_wtot=seg_wtot,
        seg_wamb=seg_wamb, seg_logq_uniq=seg_logq_uniq, seg_nobs=seg_nobs, seg_namb=seg_namb,
        seg_nuniq=seg_nuniq, seg_A=seg_A, seg_Aamb=seg_Aamb, sc_gcol=sc_gcol,
        sc_pisum0=sc_pisum0, sc_nuniq=sc_nuniq, sc_inamb=sc_inamb,
        u_row=u_row.astype(np.int32), u_pos=u_pos.astype(np.int32))
