"""Fragment-major ("FM") form of a tile -- FORMAT SPECIFICATION for the next streaming kernel,
written down and checked on the CPU before the device builder / kernel exist.  Like ``layout_host.py`` this is
not product code: nothing on the product path imports it.

Why: the streaming tile pass (celltype / pseudobulk / very large cells; ``lane_stream_pass``, csrc/ss_em.cuh)
re-derives, for every tile and every iteration, which alignments a lane owns from the JDS arrays -- 176
thread-instructions per alignment and iteration against 83 for the register-resident lane kernel.  The FM form
stores exactly what the lane kernels derive at start-up (``fq / fo / fw / fg`` of ``em_clane_kernel``,
csrc/ss_em.cu) as one fixed-size record per *row-fragment lane* and one descriptor per *column-fragment lane*,
in lane order, so that a streamed tile costs a few coalesced vector loads per lane and no index arithmetic.

Row-fragment record (64 bytes, slot v = lane + j * threads, j < fr):
    q[4]    f64   Q of the lane's <= 4 alignments (0 for padding)
    off[4]  u32   (lcol * 8) | (cpos * 8) << 16 : byte offsets into the x vector and the column-major scratch;
                  padding points at the dummy slots (ncol, pe)
    w       f64   row weight (max_j Q_ij, model.py:183), same in all lanes of the row's group
    meta    u32   n (alignments in this lane, 0..4) | group << 8  (group = lanes of the row: 1, 2, 4 or 8)
    pad     u32
Rows longer than 32 alignments and rows beyond the ``fr * threads`` lanes stay in the JDS arrays (generic path),
as in the lane kernels; ``row_generic`` lists them.

Column-fragment descriptor (8 bytes, slot v = lane + j * threads, j < fc):
    cb      u32   byte offset of the lane's first slot in the column-major scratch
    cm      u32   n (slots of this lane, 0..4) | group << 4 | (column + 1 if the lane leads the group else 0) << 12
Columns with more than 128 rows (one warp each) and columns beyond the lanes are listed in ``col_generic``.

Byte budget (celltype-like synthetic segment, 2048-alignment tiles, records of the used lanes rounded up to whole
warps): 2.96 alignments per 4-slot row lane -> 22.2 B per alignment for the row records + 2.8 B for the column
descriptors = 24.9 B against ~15 B for the JDS arrays the present kernel streams.  If that turns the pass
HBM-bound, the compact variant stores the uint16 score instead of Q (40-byte record, ~16.4 B per alignment) and
looks Q up in a per-segment table in shared memory (scores span ~120 values).

The lane-to-fragment mapping is the one of ``em_lane_kernel`` / ``em_clane_kernel``: rows sorted by length
(descending) are cut into groups of 8 / 4 / 2 / 1 lanes by the header counts H_RGT4..H_RGT32, columns (sorted by
count) into groups of 32..1 lanes by H_CGT4..H_CGT128; a group never straddles a warp because the class
boundaries are multiples of the group size.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .layout_host import FRAG, HostLayout, _pad

ROW_REC = np.dtype([('q', '<f8', (FRAG,)), ('off', '<u4', (FRAG,)), ('w', '<f8'), ('meta', '<u4'), ('pad', '<u4')])
COL_REC = np.dtype([('cb', '<u4'), ('cm', '<u4')])
assert ROW_REC.itemsize == 64 and COL_REC.itemsize == 8


@dataclass
class FMTile:
    rows: np.ndarray          # ROW_REC [fr * threads]
    cols: np.ndarray          # COL_REC [fc * threads]
    row_generic: np.ndarray   # JDS rows that are not covered by the records
    col_generic: np.ndarray   # tile columns that are not covered by the descriptors
    ncol: int
    pe: int                   # padded number of alignments: index of the dummy scratch slot


def tile_slices(L: HostLayout, t: int):
    h = L.tile_hdr[t]
    nrows, nent, ncol, ml = int(h[1]), int(h[2]), int(h[3]), int(h[4])
    eo, ro, co, jo = int(h[6]), int(h[7]), int(h[8]), int(h[9])
    return dict(nrows=nrows, nent=nent, ncol=ncol, maxlen=ml, hdr=h,
                lcol=L.e_lcol[eo:eo + nent].astype(np.int64), cpos=L.e_cpos[eo:eo + nent].astype(np.int64),
                q=L.e_q[eo:eo + nent], rlen=L.r_len[ro:ro + nrows].astype(np.int64), w=L.r_w[ro:ro + nrows],
                cptr=L.c_ptr[co:co + ncol + 1].astype(np.int64), jptr=L.j_ptr[jo:jo + ml + 1].astype(np.int64))


def build_fm(L: HostLayout, t: int, threads: int = 512, fr: int = 2, fc: int = 4) -> FMTile:
    s = tile_slices(L, t)
    h, nrows, ncol, nent = s['hdr'], s['nrows'], s['ncol'], s['nent']
    pe = _pad(nent)
    dummy = (ncol << 3) | (pe << 19)
    rows = np.zeros(fr * threads, dtype=ROW_REC)
    rows['off'][:] = dummy
    rows['meta'][:] = 1 << 8
    # ---- rows: classes by length (H_RGT4, H_RGT8, H_RGT16, H_RGT32 = header words 12..15) ----
    rgt4, rgt8, rgt16, rgt32 = (int(x) for x in h[12:16])
    n_huge = rgt32
    n8, n4, n2, n1 = rgt16 - rgt32, rgt8 - rgt16, rgt4 - rgt8, nrows - rgt4
    V8 = 8 * n8
    V4 = V8 + 4 * n4
    V2 = V4 + 2 * n2
    V = V2 + n1
    Vc = min(V, fr * threads)
    # the lane budget fr * threads and the class boundaries are multiples of the group sizes, so a group is never
    # cut: rows [n_huge, r_ovf) are covered, the others take the generic path (as in the kernels)
    if Vc >= V:
        r_ovf = nrows
    elif Vc < V8:
        r_ovf = n_huge + Vc // 8
    elif Vc < V4:
        r_ovf = n_huge + n8 + (Vc - V8) // 4
    elif Vc < V2:
        r_ovf = n_huge + n8 + n4 + (Vc - V4) // 2
    else:
        r_ovf = n_huge + n8 + n4 + n2 + (Vc - V2)
    for vl in range(Vc):
        if vl < V8:
            g, r, k = 8, n_huge + vl // 8, vl % 8
        elif vl < V4:
            g, r, k = 4, n_huge + n8 + (vl - V8) // 4, (vl - V8) % 4
        elif vl < V2:
            g, r, k = 2, n_huge + n8 + n4 + (vl - V4) // 2, (vl - V4) % 2
        else:
            g, r, k = 1, n_huge + n8 + n4 + n2 + (vl - V2), 0
        assert r < r_ovf
        n = max(0, min(FRAG, int(s['rlen'][r]) - FRAG * k))
        rec = rows[vl]
        for p in range(n):
            e = int(s['jptr'][FRAG * k + p]) + r
            rec['q'][p] = s['q'][e]
            rec['off'][p] = (int(s['lcol'][e]) << 3) | (int(s['cpos'][e]) << 19)
        rec['w'] = s['w'][r]
        rec['meta'] = n | (g << 8)
    row_generic = np.concatenate([np.arange(n_huge), np.arange(r_ovf, nrows)]).astype(np.int64)
    # ---- columns: classes by count (H_CGT4 .. H_CGT128 = header words 16..21) ----
    b4, b8, b16, b32, b64, b128 = (int(x) for x in h[16:22])
    W32 = 32 * (b64 - b128)
    W16 = W32 + 16 * (b32 - b64)
    W8 = W16 + 8 * (b16 - b32)
    W4 = W8 + 4 * (b8 - b16)
    W2 = W4 + 2 * (b4 - b8)
    W = W2 + (ncol - b4)
    Wc = min(W, fc * threads)
    cols = np.zeros(fc * threads, dtype=COL_REC)
    cols['cm'][:] = 1 << 4
    if Wc >= W:
        c_ovf = ncol
    elif Wc < W32:
        c_ovf = b128 + Wc // 32
    elif Wc < W16:
        c_ovf = b64 + (Wc - W32) // 16
    elif Wc < W8:
        c_ovf = b32 + (Wc - W16) // 8
    elif Wc < W4:
        c_ovf = b16 + (Wc - W8) // 4
    elif Wc < W2:
        c_ovf = b8 + (Wc - W4) // 2
    else:
        c_ovf = b4 + (Wc - W2)
    for vl in range(Wc):
        if vl < W32:
            g, c, k = 32, b128 + vl // 32, vl % 32
        elif vl < W16:
            g, c, k = 16, b64 + (vl - W32) // 16, (vl - W32) % 16
        elif vl < W8:
            g, c, k = 8, b32 + (vl - W16) // 8, (vl - W16) % 8
        elif vl < W4:
            g, c, k = 4, b16 + (vl - W8) // 4, (vl - W8) % 4
        elif vl < W2:
            g, c, k = 2, b8 + (vl - W4) // 2, (vl - W4) % 2
        else:
            g, c, k = 1, b4 + (vl - W2), 0
        assert c < c_ovf
        b = int(s['cptr'][c])
        n = int(s['cptr'][c + 1]) - b
        cn = max(0, min(FRAG, n - FRAG * k))
        cols[vl]['cb'] = 8 * (b + FRAG * k)
        cols[vl]['cm'] = cn | (g << 4) | ((c + 1 if k == 0 else 0) << 12)
    col_generic = np.concatenate([np.arange(b128), np.arange(c_ovf, ncol)]).astype(np.int64)
    return FMTile(rows=rows, cols=cols, row_generic=row_generic, col_generic=col_generic, ncol=ncol, pe=pe)


def _group_sum(v, groups):
    """Sum over aligned groups of ``groups[i]`` consecutive lanes (what the xor-shuffle butterflies compute)."""
    out = v.copy()
    i = 0
    n = len(v)
    while i < n:
        g = int(groups[i])
        if g > 1:
            out[i:i + g] = v[i:i + g].sum()
        i += g
    return out


def iterate_fm(fm: FMTile, s: dict, x: np.ndarray) -> np.ndarray:
    """One fused E+M pass of a tile from its FM form: returns T_c = sum_i w_i z_ic per tile column (model.py:236)
    given x_c = pi_c * theta_c -- the arithmetic of phases A and B of the lane kernels, lane by lane."""
    ncol, pe = fm.ncol, fm.pe
    xs = np.zeros(ncol + 1)
    xs[:ncol] = x
    val = np.zeros(pe + 8)
    R = fm.rows
    lo = (R['off'] & 0xffff) >> 3
    hi = R['off'] >> 19
    n = R['meta'] & 0xff
    g = np.maximum(R['meta'] >> 8, 1)
    acc = (R['q'] * xs[lo]).sum(axis=1)
    acc = _group_sum(acc, g)
    with np.errstate(divide='ignore', invalid='ignore'):
        sc = np.where(acc > 0, R['w'] / acc, 0.0)
    for p in range(FRAG):
        m = n > p
        val[hi[m, p]] = R['q'][m, p] * sc[m]
    for r in fm.row_generic:                                 # generic path: straight from the JDS arrays
        ln = int(s['rlen'][r])
        e = s['jptr'][:ln] + r
        a = float((s['q'][e] * xs[s['lcol'][e]]).sum())
        val[s['cpos'][e]] = s['q'][e] * (s['w'][r] / a if a > 0 else 0.0)
    T = np.zeros(ncol)
    C = fm.cols
    cn = C['cm'] & 15
    cg = np.maximum((C['cm'] >> 4) & 63, 1)
    first = C['cb'] >> 3
    part = np.zeros(len(C))
    for k in range(FRAG):
        m = cn > k
        part[m] += val[first[m] + k]
    part = _group_sum(part, cg)
    lead = C['cm'] >> 12
    T[lead[lead > 0] - 1] = part[lead > 0]
    for c in fm.col_generic:
        T[c] = val[s['cptr'][c]:s['cptr'][c + 1]].sum()
    return T * x            # the scratch holds Q_ic w_i / r_i; x_c is applied once per column (as in the kernels)


def iterate_direct(s: dict, x: np.ndarray) -> np.ndarray:
    """The same pass straight from the definition (E-step model.py:207-210, M-step :236) on the tile's entries."""
    nrows = s['nrows']
    T = np.zeros(s['ncol'])
    for r in range(nrows):
        ln = int(s['rlen'][r])
        e = s['jptr'][:ln] + r
        u = s['q'][e] * x[s['lcol'][e]]
        tot = u.sum()
        if tot > 0:
            np.add.at(T, s['lcol'][e], s['w'][r] * u / tot)
    return T
