// EM kernels: per-fit preparation, the resident kernel (one CTA owns a whole
// segment and iterates to convergence out of shared memory), the streaming
// kernel (cooperative grid, segments that span several tiles, iterates in
// lockstep with an on-device convergence test), emit (final z + lnL) and the
// per-segment finalisation.
//
// Reference being replaced: TelescopeLikelihood.em / estep / mstep /
// calculate_lnl, stellarscope/utils/model.py:202-373.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>

#include "ss_em.cuh"
#include "ss_handle.h"

namespace cg = cooperative_groups;

// streaming kernel: CTA size and fragment slots per lane (rows / columns) of the tile pass
#ifndef SS_STREAM_THREADS
#define SS_STREAM_THREADS 512
#define SS_STREAM_FR 2
#define SS_STREAM_FC 3
#endif

namespace ss {

struct StreamArgs {
  const int* tiles;
  int n_tiles;
  const int* segs;
  int n_segs;
  const int* col_seg;    // [n_cols] index into segs
  const int* col_slot;   // [n_cols] absolute slot
  long long n_cols;
  int* ctl;              // [0] = number of active segments, [1 + i] = segment i done (iters)
  const int* chunk_seg;  // [n_chunks] column chunks (SS_COL_CHUNK slots of one segment): segment ...
  const int* chunk_j0;   // ... and first column of the chunk (segment-relative)
  int n_chunks;
  StreamGeom geom;       // shared-memory geometry of the streaming tile pass
};

// ---------------------------------------------------------------------------
// prepare: per segment constants (model.py:189-194) and default parameters of the columns no tile
// touches.  Two kernels: one warp per column chunk (SS_COL_CHUNK slots of one segment -- a segment
// with many columns does not serialise on one warp), then one warp per segment that adds the
// chunks' partial sums in chunk order.
// ---------------------------------------------------------------------------
struct ColChunks {
  int n;
  const int* seg;        // [n] segment of the chunk
  const int* j0;         // [n] first column of the chunk (segment-relative)
  const int* seg_chunk0; // [S] first chunk of the segment
  double* part;          // [2][n]
};

struct SegConst {
  double invK, thpw, pipw, inv_thd, inv_pid;
  bool bad;
};
__device__ __forceinline__ SegConst seg_constants(const Layout& L, const FitParams& P, int s) {
  SegConst c;
  const double K = (double)P.K;
  c.invK = 1.0 / K;
  const double wmax = L.seg_wmax[s];
  c.thpw = P.theta_prior * wmax;
  c.pipw = P.pi_prior * wmax;
  const double thd = L.seg_wamb[s] + c.thpw * K;
  const double pid = L.seg_wtot[s] + c.pipw * K;
  c.bad = !(thd > 0.0) || !(pid > 0.0);
  c.inv_thd = c.bad ? 0.0 : 1.0 / thd;
  c.inv_pid = c.bad ? 0.0 : 1.0 / pid;
  return c;
}

__global__ void fit_prepare_cols_kernel(Layout L, FitState F, FitParams P, const int8_t* __restrict__ sc_inamb, ColChunks cc) {
  const int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ch >= cc.n) return;
  const int s = cc.seg[ch];
  const SegConst k = seg_constants(L, P, s);
  const int c0 = L.seg_col0[s], nc = L.seg_ncol[s];
  const int jb = cc.j0[ch], je = min(nc, jb + SS_COL_CHUNK);
  const double th_rest = k.thpw * k.inv_thd;
  const double invK = k.invK;
  double d1 = 0.0, lx = 0.0;
  // Single-tile segments (lane kernel) keep their state on chip and publish pi / theta / x of the tile's columns
  // themselves: only the columns no tile touches need their (constant) parameters here.  Multi-tile segments
  // (cluster / streaming kernels) start from the per-slot state below.
  const bool full = L.seg_ntiles[s] > 1;
  for (int j0 = jb; j0 < je; j0 += 128) {
    double ps0[4];
    int nu[4];
    bool out[4];                  // column outside the tiles
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u + lane;
      const bool ok = j < je;
      out[u] = ok && !sc_inamb[c0 + j];
      ps0[u] = ok ? L.sc_pisum0[c0 + j] : 0.0;
      nu[u] = out[u] ? L.sc_nuniq[c0 + j] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u + lane;
      if (j >= je) continue;
      const int slot = c0 + j;
      const double p1 = (ps0[u] + k.pipw) * k.inv_pid;
      if (full || out[u]) {
        F.sc_pi[slot] = p1;
        F.sc_theta[slot] = th_rest;
      }
      if (full) {
        F.sc_ptprev[slot] = invK * invK;
        F.sc_ptfin[slot] = p1 * th_rest;
        F.sc_pt[0][slot] = invK * invK;
        F.sc_pt[1][slot] = invK * invK;
        F.sc_T[slot] = 0.0;
      }
      if (out[u]) {
        d1 += fabs(p1 - invK);
        if (nu[u] > 0 && p1 > 0.0) lx += (double)nu[u] * log(p1);
      }
    }
  }
  d1 = warp_sum(d1);
  lx = warp_sum(lx);
  if (lane == 0) {
    cc.part[ch] = d1;
    cc.part[cc.n + ch] = lx;
  }
}

__global__ void fit_prepare_kernel(Layout L, FitState F, FitParams P, ColChunks cc) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= L.S) return;
  const int s = warp;
  const SegConst k = seg_constants(L, P, s);
  const int nc = L.seg_ncol[s];
  const int ch0 = cc.seg_chunk0[s], nch = (nc + SS_COL_CHUNK - 1) / SS_COL_CHUNK;
  double d1 = 0.0, lx = 0.0;
  for (int i = lane; i < nch; i += 32) {
    d1 += cc.part[ch0 + i];
    lx += cc.part[cc.n + ch0 + i];
  }
  d1 = warp_sum(d1);
  lx = warp_sum(lx);
  if (lane == 0) {
    const double pi_rest = k.pipw * k.inv_pid;
    d1 += ((double)P.K - (double)nc) * fabs(pi_rest - k.invK);
    F.seg_thpw[s] = k.thpw;
    F.seg_pipw[s] = k.pipw;
    F.seg_inv_thd[s] = k.inv_thd;
    F.seg_inv_pid[s] = k.inv_pid;
    F.seg_diff1_extra[s] = d1;
    F.seg_lnl_extra[s] = lx + L.seg_logq_uniq[s];
    F.seg_diffacc[s] = 0.0;
    F.seg_lnl[s] = 0.0;
    F.seg_iters[s] = 0;
    F.seg_flags[s] = k.bad ? 4 : 0;
  }
}

// ---------------------------------------------------------------------------
// lane kernel: one CTA per single-tile segment (a cell in individual mode), iterated to
// convergence on chip.
//
// The tile's static data (Q, index words, row weights, the per-column state pi / x) lives in
// REGISTERS for the whole fit: every lane owns FR row fragments and FC column fragments of <= 4
// entries.  A row (column) with more than 4 entries is spread over an aligned group of 2/4/8
// (..32) lanes whose partial sums are combined by xor shuffles, so every lane does the same small
// amount of work whatever the row / column lengths are (no intra-CTA imbalance at the barriers).
// Shared memory only carries what rows and columns exchange: the column-major scratch `val`
// (8 B per alignment) and the vector x = pi*theta (ping-pong).  Per alignment and iteration that
// is one gather (x_j), one scatter (val) and one column read -- the shared-memory data path is the
// binding resource of this kernel, so nothing else goes through it.
//
// Lane layout (from the tile header: rows are sorted by length, columns by count, descending):
// virtual lane vl = tid + j * THREADS for fragment slot j; the 8-lane groups come first, then the
// 4-, 2- and 1-lane ones, so that groups are aligned and the lanes of a warp have (almost always)
// the same fragment length.  Per slot the warp-uniform maximum fragment length / group size is
// computed once; inside the loop only uniform branches remain, lanes with fewer entries work on
// padding (q = 0, a dummy column whose x is 0, a dummy scratch slot).
// Rows longer than 32 / columns longer than 128 and whatever does not fit the FR*THREADS
// (FC*THREADS) lanes take generic paths that read the tile from global memory.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double group_xor_sum(double a, int g, int gmax) {
  // butterfly inside aligned lane groups of size g (1,2,4,..,32); executed by all 32 lanes, gmax is
  // the largest group of the warp (warp-uniform)
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    if (o < gmax) {
      const double t = __shfl_xor_sync(full, a, o);
      if (o < g) a += t;
    }
  }
  return a;
}

// deterministic block sum for a run-time number of warps (every thread returns the same value;
// one barrier, which also publishes the caller's preceding shared-memory writes).  `red` holds
// 2 x 32 doubles; callers alternate `par`.
__device__ __forceinline__ double lane_block_sum(double v, int nw, double* red, int par) {
  v = warp_sum(v);
  if (nw == 1) {
    __syncwarp();
    return v;
  }
  const int lane = threadIdx.x & 31;
  double* r = red + 32 * (par & 1);
  if (lane == 0) r[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = lane < nw ? r[lane] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    if (o < nw) t += __shfl_xor_sync(0xffffffffu, t, o);        // nw is block-uniform; lanes >= nw hold 0
  return __shfl_sync(0xffffffffu, t, 0);
}

// THREADS = blockDim.x (a multiple of 32, chosen per tile by the launch plan so that all warps
// carry the same number of fragment slots)
template <int TT, bool LNL, int FR, int FC, int MAXREG>
__global__ void __maxnreg__(MAXREG) em_lane_kernel(EmArgs a) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double s_red[64];
  const int THREADS = TT ? TT : blockDim.x;
  const int NW = THREADS >> 5;
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = a.tile_list[blockIdx.x];
  const long long* __restrict__ hd = a.L.tile_hdr + (long long)HDR_WORDS * t;
  const TileView v = load_tile_view(a.L.tile_hdr, t);
  const int seg = v.seg;
  double* val = reinterpret_cast<double*>(dsm);      // pe + 8 (slot pe: scratch of the padding entries)
  double* pt = val + v.pe + 8;                       // pc  (slot ncol: x of the dummy column = 0)
  double* pt2 = pt + v.pc;                           // pc
  double* aux = pt2 + v.pc;                          // pc  pi of the columns on the generic path; all columns after the fit
  // global views of the tile
  const double* __restrict__ gq = a.L.e_q + v.eo;
  const uint32_t* __restrict__ gidx = a.L.e_idx + v.eo;
  const uint16_t* __restrict__ grlen = a.L.r_len + v.ro;
  const double* __restrict__ gw = a.L.r_w + v.ro;
  const uint16_t* __restrict__ gjptr = a.L.j_ptr + v.jo;
  const uint16_t* __restrict__ gcptr = a.L.c_ptr + v.co;
  const double* __restrict__ ps0 = a.L.c_pisum0 + v.co;
  const int* __restrict__ nuq = a.L.c_nuniq + v.co;
  const int* __restrict__ scol = a.L.c_scol + v.co;

  const double invK = 1.0 / (double)a.P.K;       // model.py:164,172
  for (int c = tid; c < v.ncol; c += THREADS) {
    aux[c] = invK;
    pt[c] = invK * invK;
  }
  if (tid == 0) {
    pt[v.ncol] = 0.0;
    pt2[v.ncol] = 0.0;
  }
  const double thpw = a.F.seg_thpw[seg], pipw = a.F.seg_pipw[seg];
  const double inv_thd = a.F.seg_inv_thd[seg], inv_pid = a.F.seg_inv_pid[seg];
  const double diff1_extra = a.F.seg_diff1_extra[seg];
  const double lnl_extra = a.F.seg_lnl_extra[seg];
  const double eps = a.P.eps;
  const int max_iter = a.P.max_iter;
  const bool has_uniq = a.L.seg_nuniq[seg] > 0;
  double* diffs = a.F.seg_diffs + (long long)seg * max_iter;
  double* lnls = a.F.seg_lnls ? a.F.seg_lnls + (long long)seg * max_iter : nullptr;

  // ---- row fragments ----
  const int rgt4 = (int)hd[H_RGT4], rgt8 = (int)hd[H_RGT8], rgt16 = (int)hd[H_RGT16], rgt32 = (int)hd[H_RGT32];
  const int n_huge = rgt32;                                     // generic path
  const int n8 = rgt16 - rgt32, n4 = rgt8 - rgt16, n2 = rgt4 - rgt8, n1 = v.nrows - rgt4;
  const int V8 = 8 * n8, V4 = V8 + 4 * n4, V2 = V4 + 2 * n2, V = V2 + n1;   // virtual lanes
  const int Vc = min(V, FR * THREADS);                          // lanes held in registers
  // first row that is NOT covered by the register fragments (group sizes divide 32)
  int r_ovf;
  if (Vc >= V) r_ovf = v.nrows;
  else if (Vc < V8) r_ovf = n_huge + Vc / 8;
  else if (Vc < V4) r_ovf = n_huge + n8 + (Vc - V8) / 4;
  else if (Vc < V2) r_ovf = n_huge + n8 + n4 + (Vc - V4) / 2;
  else r_ovf = n_huge + n8 + n4 + n2 + (Vc - V2);
  const bool generic_rows = n_huge > 0 || r_ovf < v.nrows;

  const uint32_t dummy_off = ((uint32_t)v.ncol << 3) | ((uint32_t)v.pe << 19);
  double fq[FR][FRAG], fw[FR];
  uint32_t fo[FR][FRAG];        // byte offset of x_j | byte offset of the scratch slot << 16
  int fg[FR], nA[FR], gA[FR];
#pragma unroll
  for (int j = 0; j < FR; ++j) {
    const int vl = tid + j * THREADS;
    int n = 0;
    fg[j] = 1; fw[j] = 0.0;
#pragma unroll
    for (int p = 0; p < FRAG; ++p) { fq[j][p] = 0.0; fo[j][p] = dummy_off; }
    if (vl < Vc) {
      int r, k, g;
      if (vl < V8) { g = 8; r = n_huge + vl / 8; k = vl % 8; }
      else if (vl < V4) { g = 4; r = n_huge + n8 + (vl - V8) / 4; k = (vl - V8) % 4; }
      else if (vl < V2) { g = 2; r = n_huge + n8 + n4 + (vl - V4) / 2; k = (vl - V4) % 2; }
      else { g = 1; r = n_huge + n8 + n4 + n2 + (vl - V2); k = 0; }
      const int len = grlen[r];
      n = max(0, min(FRAG, len - FRAG * k));
      fg[j] = g; fw[j] = gw[r];
#pragma unroll
      for (int p = 0; p < FRAG; ++p)
        if (p < n) {
          const int e = (int)gjptr[FRAG * k + p] + r;
          const uint32_t ix = gidx[e];
          fq[j][p] = gq[e];
          fo[j][p] = ((ix & 0xffffu) << 3) | ((ix >> 16) << 19);
        }
    }
    nA[j] = __reduce_max_sync(full, n);
    gA[j] = __reduce_max_sync(full, fg[j]);
  }
  // ---- column fragments ----
  const int b128 = (int)hd[H_CGT128];                           // generic (warp-strided) path
  const int b64 = (int)hd[H_CGT64], b32 = (int)hd[H_CGT32], b16 = (int)hd[H_CGT16], b8 = (int)hd[H_CGT8],
            b4 = (int)hd[H_CGT4];
  const int W32 = 32 * (b64 - b128), W16 = W32 + 16 * (b32 - b64), W8 = W16 + 8 * (b16 - b32);
  const int W4 = W8 + 4 * (b8 - b16), W2 = W4 + 2 * (b4 - b8), W = W2 + (v.ncol - b4);
  const int Wc = min(W, FC * THREADS);
  int c_ovf;
  if (Wc >= W) c_ovf = v.ncol;
  else if (Wc < W32) c_ovf = b128 + Wc / 32;
  else if (Wc < W16) c_ovf = b64 + (Wc - W32) / 16;
  else if (Wc < W8) c_ovf = b32 + (Wc - W16) / 8;
  else if (Wc < W4) c_ovf = b16 + (Wc - W8) / 4;
  else if (Wc < W2) c_ovf = b8 + (Wc - W4) / 2;
  else c_ovf = b4 + (Wc - W2);
  const bool generic_cols = b128 > 0 || c_ovf < v.ncol;

  int cb[FC], cn[FC], cg[FC], cc[FC];       // byte offset of the first slot, #slots (<= FRAG), group size, column (-1: not leader)
  int nB[FC], gB[FC];
  double cpi[FC], cps0[FC], cx[FC];
#pragma unroll
  for (int j = 0; j < FC; ++j) {
    const int vl = tid + j * THREADS;
    cb[j] = 0; cn[j] = 0; cg[j] = 1; cc[j] = -1;
    cpi[j] = invK; cps0[j] = 0.0; cx[j] = invK * invK;
    if (vl < Wc) {
      int c, k, g;
      if (vl < W32) { g = 32; c = b128 + vl / 32; k = vl % 32; }
      else if (vl < W16) { g = 16; c = b64 + (vl - W32) / 16; k = (vl - W32) % 16; }
      else if (vl < W8) { g = 8; c = b32 + (vl - W16) / 8; k = (vl - W16) % 8; }
      else if (vl < W4) { g = 4; c = b16 + (vl - W8) / 4; k = (vl - W8) % 4; }
      else if (vl < W2) { g = 2; c = b8 + (vl - W4) / 2; k = (vl - W4) % 2; }
      else { g = 1; c = b4 + (vl - W2); k = 0; }
      const int b = gcptr[c], n = (int)gcptr[c + 1] - b;
      cb[j] = 8 * (b + FRAG * k);
      cn[j] = max(0, min(FRAG, n - FRAG * k));
      cg[j] = g;
      cc[j] = k == 0 ? c : -1;
      cps0[j] = has_uniq ? ps0[c] : 0.0;
    }
    nB[j] = __reduce_max_sync(full, cn[j]);
    gB[j] = __reduce_max_sync(full, cg[j]);
  }
  __syncthreads();

  double* pt_cur = pt;
  double* pt_nxt = pt2;
  unsigned char* const vb = reinterpret_cast<unsigned char*>(val);
  double lnl_prev = INFINITY, lnl_amb = 0.0;
  double finv[FR];                           // 1 / row sum of the last E-step (likelihood mode)
  int it = 0, par = 0;
  bool done = false, conv = false;
  while (!done) {
    ++it;
    // ---- phase A: rows.  r_i = sum_j Q_ij x_j ; val[i,j] = Q_ij w_i / r_i ----
    const unsigned char* const xb = reinterpret_cast<const unsigned char*>(pt_cur);
#pragma unroll
    for (int j = 0; j < FR; ++j) {
      if (nA[j] > 0) {                                            // warp-uniform
        double acc = 0.0;
#pragma unroll
        for (int p = 0; p < FRAG; ++p)
          if (p < nA[j]) acc = fma(fq[j][p], *reinterpret_cast<const double*>(xb + (fo[j][p] & 0xffffu)), acc);
        if (gA[j] > 1) acc = group_xor_sum(acc, fg[j], gA[j]);
        const bool pos = acc > 0.0;
        double inv = 1.0 / (pos ? acc : 1.0);                     // padding lanes: no 1/0 slow path
        inv = pos ? inv : 0.0;
        if (LNL) finv[j] = inv;
        const double sc = fw[j] * inv;
#pragma unroll
        for (int p = 0; p < FRAG; ++p)
          if (p < nA[j]) *reinterpret_cast<double*>(vb + (fo[j][p] >> 16)) = fq[j][p] * sc;
      } else if (LNL) {
        finv[j] = 0.0;
      }
    }
    if (generic_rows) {
      for (int r = tid; r < v.nrows; r += THREADS) {               // rows outside the register lanes (rare)
        const int rr = r < n_huge ? r : r_ovf + (r - n_huge);
        if (rr >= v.nrows) break;
        const int len = grlen[rr];
        double acc = 0.0;
        for (int p = 0; p < len; ++p) {
          const int e = (int)gjptr[p] + rr;
          acc = fma(gq[e], pt_cur[gidx[e] & 0xffffu], acc);
        }
        const double sc = acc > 0.0 ? gw[rr] * (1.0 / acc) : 0.0;
        for (int p = 0; p < len; ++p) {
          const int e = (int)gjptr[p] + rr;
          val[gidx[e] >> 16] = gq[e] * sc;
        }
      }
    }
    if (NW == 1) __syncwarp(); else __syncthreads();
    // ---- phase B: thetasum per tile column -> theta', pi', x' = pi' theta', |delta pi| ----
    double dsum = 0.0;
#pragma unroll
    for (int j = 0; j < FC; ++j) {
      if (nB[j] > 0) {                                            // warp-uniform
        double asum = 0.0;
        const unsigned char* const cp = vb + cb[j];
#pragma unroll
        for (int k = 0; k < FRAG; ++k)
          if (k < nB[j]) {
            if (k < cn[j]) asum += *reinterpret_cast<const double*>(cp + 8 * k);
          }
        if (gB[j] > 1) asum = group_xor_sum(asum, cg[j], gB[j]);
        const double T = cx[j] * asum;
        const double th = (T + thpw) * inv_thd;                 // model.py:240
        const double pn = (cps0[j] + T + pipw) * inv_pid;       // model.py:243-244
        const double xn = pn * th;
        if (cc[j] >= 0) {
          dsum += fabs(pn - cpi[j]);
          pt_nxt[cc[j]] = xn;
        }
        cpi[j] = pn;
        cx[j] = xn;
      }
    }
    if (generic_cols) {
      for (int c = warp; c < b128; c += NW) {                     // very long columns: one warp each
        const int b = gcptr[c], n = (int)gcptr[c + 1] - b;
        const double asum = col_sum_warp(val, b, n, lane);
        if (lane == 0) {
          const double T = pt_cur[c] * asum;
          const double th = (T + thpw) * inv_thd;
          const double pn = ((has_uniq ? ps0[c] : 0.0) + T + pipw) * inv_pid;
          dsum += fabs(pn - aux[c]);
          aux[c] = pn;
          pt_nxt[c] = pn * th;
        }
      }
      for (int c = c_ovf + tid; c < v.ncol; c += THREADS) {       // columns beyond the register lanes
        const int b = gcptr[c], n = (int)gcptr[c + 1] - b;
        const double T = pt_cur[c] * col_sum_thread(val, b, n);
        const double th = (T + thpw) * inv_thd;
        const double pn = ((has_uniq ? ps0[c] : 0.0) + T + pipw) * inv_pid;
        dsum += fabs(pn - aux[c]);
        aux[c] = pn;
        pt_nxt[c] = pn * th;
      }
    }
    double diff = lane_block_sum(dsum, NW, s_red, par++);     // barrier: publishes pt_nxt / aux
    if (it == 1) diff += diff1_extra;
    const bool reached = it >= max_iter;
    if (!LNL) {
      conv = diff < eps;                                       // model.py:351
    } else {
      // lnL with the E-step's z (old x) and the new parameters: model.py:344
      double l = 0.0;
#pragma unroll
      for (int j = 0; j < FR; ++j)
#pragma unroll
        for (int p = 0; p < FRAG; ++p)
          if (p < nA[j]) {
            const int c = (fo[j][p] & 0xffffu) >> 3;
            const double z = (fq[j][p] * pt_cur[c]) * finv[j];
            const double pn = pt_nxt[c];
            if (z > 0.0 && pn > 0.0) l = fma(z, log(fq[j][p] * pn), l);
          }
      if (generic_rows) {
        for (int r = tid; r < v.nrows; r += THREADS) {
          const int rr = r < n_huge ? r : r_ovf + (r - n_huge);
          if (rr >= v.nrows) break;
          const int len = grlen[rr];
          double acc = 0.0;
          for (int p = 0; p < len; ++p) {
            const int e = (int)gjptr[p] + rr;
            acc = fma(gq[e], pt_cur[gidx[e] & 0xffffu], acc);
          }
          const double inv = acc > 0.0 ? 1.0 / acc : 0.0;
          for (int p = 0; p < len; ++p) {
            const int e = (int)gjptr[p] + rr;
            const int c = gidx[e] & 0xffffu;
            const double z = (gq[e] * pt_cur[c]) * inv;
            const double pn = pt_nxt[c];
            if (z > 0.0 && pn > 0.0) l = fma(z, log(gq[e] * pn), l);
          }
        }
      }
      if (has_uniq) {
#pragma unroll
        for (int j = 0; j < FC; ++j)
          if (cc[j] >= 0) {
            const int nu = nuq[cc[j]];
            if (nu > 0 && cpi[j] > 0.0) l += (double)nu * log(cpi[j]);
          }
        for (int c = warp; c < b128; c += NW)
          if (lane == 0) { const int nu = nuq[c]; if (nu > 0 && aux[c] > 0.0) l += (double)nu * log(aux[c]); }
        for (int c = c_ovf + tid; c < v.ncol; c += THREADS) {
          const int nu = nuq[c];
          if (nu > 0 && aux[c] > 0.0) l += (double)nu * log(aux[c]);
        }
      }
      lnl_amb = lane_block_sum(l, NW, s_red, par++);
      const double lnl = lnl_amb + lnl_extra;
      conv = fabs(lnl - lnl_prev) < eps;                       // model.py:345-347
      lnl_prev = lnl;
      if (tid == 0 && lnls) lnls[it - 1] = lnl;
    }
    done = conv || reached;
    if (tid == 0) diffs[it - 1] = diff;
    if (!done) { double* tmp = pt_cur; pt_cur = pt_nxt; pt_nxt = tmp; }
  }
  // ---- converged / reached max: publish parameters, final lnL ----
  // pt_cur = x of the last E-step (tl.z belongs to it), pt_nxt = final x; `val` still holds the last E-step.
  // theta' = x'/pi' would lose the last bit: redo the column sum exactly like the last iteration did
#pragma unroll
  for (int j = 0; j < FC; ++j) {
    if (nB[j] > 0) {
      double asum = 0.0;
      const unsigned char* const cp = vb + cb[j];
#pragma unroll
      for (int k = 0; k < FRAG; ++k)
        if (k < nB[j]) {
          if (k < cn[j]) asum += *reinterpret_cast<const double*>(cp + 8 * k);
        }
      if (gB[j] > 1) asum = group_xor_sum(asum, cg[j], gB[j]);
      if (cc[j] >= 0) {
        const int slot = scol[cc[j]];
        a.F.sc_pi[slot] = cpi[j];
        a.F.sc_theta[slot] = (pt_cur[cc[j]] * asum + thpw) * inv_thd;
        aux[cc[j]] = cpi[j];
      }
    }
  }
  if (generic_cols) {
    for (int c = warp; c < b128; c += NW) {
      const int b = gcptr[c], n = (int)gcptr[c + 1] - b;
      const double asum = col_sum_warp(val, b, n, lane);      // same tree as the iteration
      if (lane == 0) {
        a.F.sc_pi[scol[c]] = aux[c];
        a.F.sc_theta[scol[c]] = (pt_cur[c] * asum + thpw) * inv_thd;
      }
    }
    for (int c = c_ovf + tid; c < v.ncol; c += THREADS) {
      const int b = gcptr[c], n = (int)gcptr[c + 1] - b;
      a.F.sc_pi[scol[c]] = aux[c];
      a.F.sc_theta[scol[c]] = (pt_cur[c] * col_sum_thread(val, b, n) + thpw) * inv_thd;
    }
  }
  for (int c = tid; c < v.ncol; c += THREADS) {
    a.F.sc_ptprev[scol[c]] = pt_cur[c];
    a.F.sc_ptfin[scol[c]] = pt_nxt[c];
  }
  double l = 0.0;
  if (!LNL) {
    // final lnL: z of the last E-step, final parameters (model.py:368-370); generic row pass from global
    for (int r = tid; r < v.nrows; r += THREADS) {
      const int len = grlen[r];
      double acc = 0.0;
      for (int p = 0; p < len; ++p) {
        const int e = (int)gjptr[p] + r;
        acc = fma(gq[e], pt_cur[gidx[e] & 0xffffu], acc);
      }
      const double inv = acc > 0.0 ? 1.0 / acc : 0.0;
      for (int p = 0; p < len; ++p) {
        const int e = (int)gjptr[p] + r;
        const int c = gidx[e] & 0xffffu;
        const double z = (gq[e] * pt_cur[c]) * inv;
        const double pn = pt_nxt[c];
        if (z > 0.0 && pn > 0.0) l = fma(z, log(gq[e] * pn), l);
      }
    }
    // + the unique rows of the tile's columns, n_uniq_j * log(pi_j) (the columns no tile touches are in
    // seg_lnl_extra): the tile of a single-tile segment carries the segment's whole ambiguous-column part
    if (has_uniq) {
#pragma unroll
      for (int j = 0; j < FC; ++j)
        if (cc[j] >= 0) {
          const int nu = nuq[cc[j]];
          if (nu > 0 && cpi[j] > 0.0) l += (double)nu * log(cpi[j]);
        }
      for (int c = warp; c < b128; c += NW)
        if (lane == 0) { const int nu = nuq[c]; if (nu > 0 && aux[c] > 0.0) l += (double)nu * log(aux[c]); }
      for (int c = c_ovf + tid; c < v.ncol; c += THREADS) {
        const int nu = nuq[c];
        if (nu > 0 && aux[c] > 0.0) l += (double)nu * log(aux[c]);
      }
    }
    l = lane_block_sum(l, NW, s_red, par++);
  } else {
    l = lnl_amb;                 // lnL of the last iteration: ambiguous rows + unique rows of the tile's columns
  }
  if (tid == 0) {
    a.F.tile_lnl[t] = l;
    a.F.seg_iters[seg] = it;
    a.F.seg_flags[seg] = (conv ? 1 : 0) | (it >= max_iter ? 2 : 0);
  }
}

// ---------------------------------------------------------------------------
// cluster lane kernel: one thread-block CLUSTER per segment of 2..8 tiles (a cell with more than
// one tile of alignments).  Every CTA keeps the rows of its tile in registers like the lane kernel
// (FR row fragments per lane).  There is NO cluster-wide barrier inside the iteration: what the CTAs
// exchange -- the partial sums of the columns several tiles hold, and one |delta pi| partial per
// CTA -- is PUSHED into the consumers' shared memory with st.async, which signals the consumer's
// mbarrier (complete_tx); a CTA waits only for the bytes it expects.  (A cluster barrier per
// iteration costs MEMBAR.ALL.GPU + the barrier + CCTL.IVALL -- the L1 invalidation turns every
// register spill into an L2 round trip -- and an exchange through L2 one more round trip; ncu /
// clock64 profile: half of an iteration.)  Per iteration
//   phase A   local rows: r_i = sum_j Q_ij x_j, scatter Q_ij w_i / r_i into the column-major scratch
//   phase B   column sums (lane fragments, xor-shuffle groups).  A column held by this tile alone
//             (the common case: the builder orders the rows by locus, so tiles overlap in few
//             columns) is finished right there by its leader lane: theta', pi', x' = pi' theta',
//             |delta pi|.  The partial sum of a shared column is pushed to the other holders.
//   wait      for the mbarrier of this pass: the partial sums of my shared columns from their other
//             holders and the |delta pi| partials of the previous iteration have arrived
//   shared    every holder of a shared column adds the partial sums in tile order (bit-identical
//             theta', pi', x' in all holders -- no owner, no second exchange); |delta pi| is summed
//             by the first holder.
// The cluster-wide |delta pi| of iteration k is therefore complete one pass later: convergence of
// iteration k is detected in pass k+1 (whose phases A / B ran speculatively and are discarded: x is
// triple-buffered so that x_{k-1} and x_k survive); one more pass repeats iteration k from x_{k-1}
// -- the same sums in the same order -- and publishes pi and theta.  No global memory traffic and
// no atomics inside the loop.  Buffers and mbarriers alternate with the pass parity; a CTA cannot
// run two passes ahead of a peer because it needs the peer's |delta pi| partial of the pass before.
// (The likelihood-convergence mode keeps the three-barrier kernel below.)
// ---------------------------------------------------------------------------
struct ClusterArgs {
  const int* segs;
  const uint16_t* xref;
  const long long* seg_xref_off;
  const int* seg_xmax;             // receive slots per tile (cluster lane kernel)
};

constexpr int SS_MAX_CLUSTER_TILES = 8;

__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
  return r;
}
// remote store of 8 bytes into a peer CTA's shared memory; completes 8 bytes on the peer's mbarrier
__device__ __forceinline__ void st_async_f64(uint32_t raddr, double v, uint32_t rmbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr),
               "l"(__double_as_longlong(v)), "r"(rmbar)
               : "memory");
}
__device__ __forceinline__ uint32_t ld_dsmem_u16(uint32_t raddr) {
  uint16_t v;
  asm volatile("ld.shared::cluster.u16 %0, [%1];" : "=h"(v) : "r"(raddr));
  return v;
}
// wait for a phase of an mbarrier that peers complete with remote stores (st.async lands in THIS CTA's shared memory
// before its complete_tx, so the default cta-scope acquire is what CUTLASS' cluster barriers use too; a cluster-scope
// acquire makes ptxas add CCTL.IVALL -- an L1 invalidation per pass); a lost peer turns into a trap instead of a
// hung GPU
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (unsigned long long spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1ull << 24)) __trap();
  }
}

// Two instantiations (ss_handle.h): heavy <512, 4, 6, 128> -- one CTA owns its SM -- and light <512, 2, 3, 64> for
// clusters whose tiles need at most 2 row-fragment slots per lane and at most half an SM's shared memory: two CTAs
// (of different clusters) share an SM
template <int THREADS, int FR_, int FC_, int MAXREG>
__global__ void __maxnreg__(MAXREG) em_clane_kernel(EmArgs a, ClusterArgs ca) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double s_red[64];
  constexpr int FR = FR_, FC = FC_;
  constexpr int NW = THREADS / 32;
  const unsigned full = 0xffffffffu;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int seg = ca.segs[blockIdx.x / nt];
  const int t0 = a.L.seg_tile0[seg];
  const int t = t0 + rank;
  const long long* __restrict__ hd = a.L.tile_hdr + (long long)HDR_WORDS * t;
  const TileView v = load_tile_view(a.L.tile_hdr, t);
  // the same carve in every CTA of the cluster: sizes of the segment's largest tile; xmax (launch plan) bounds the
  // partial sums a tile can receive per pass: sum over the segment's tile columns of (holders - 1)
  int pemax = 0, pcmax = 0;
  for (int k = 0; k < nt; ++k) {
    const long long* hk = a.L.tile_hdr + (long long)HDR_WORDS * (t0 + k);
    pemax = max(pemax, pad8((int)hk[H_NENT]));
    pcmax = max(pcmax, pad8((int)hk[H_NCOL] + 1));
  }
  const int xmax = ca.seg_xmax[seg];
  // x of three consecutive iterations (rotating: the convergence of iteration k is known one pass late, when the
  // column lanes have already written x_{k+1}; x_{k-1} and x_k must survive), pi, the prior pseudo counts of the
  // columns, the column-major scratch, the receive buffers of the two pass parities
  double* xb0 = reinterpret_cast<double*>(dsm);
  double* xb1 = xb0 + pcmax;
  double* xb2 = xb1 + pcmax;
  double* spi = xb2 + pcmax;                           // pi
  double* sps0 = spi + pcmax;                          // pisum0 of the column (unique rows, model.py:191)
  double* val = sps0 + pcmax;                          // pemax + 8 (slot pe: scratch of the padding entries)
  double* rpart = val + pemax + 8;                     // [2][xmax] partial sums of my shared columns from their other holders
  uint16_t* xr = reinterpret_cast<uint16_t*>(rpart + 2 * xmax);   // [ncol][nt]: see below
  uint16_t* shl = xr + (size_t)nt * pcmax;             // the shared columns, ascending (bit 15: an earlier tile holds it too)
  uint16_t* rb = shl + pcmax;                          // [ncol] first receive slot of a shared column
  uint16_t* shp = rb + pcmax;                          // per list entry: the scratch slot its column sum is parked in
  uint8_t* shr = reinterpret_cast<uint8_t*>(shp + pcmax);   // [ncol] another tile holds the column too
  __shared__ __align__(8) uint64_t s_mbar[2];
  __shared__ double s_dall[2][SS_MAX_CLUSTER_TILES];   // [parity][rank]: |delta pi| partial of every CTA of the cluster (pushed)

  const double* __restrict__ gq = a.L.e_q + v.eo;
  const uint32_t* __restrict__ gidx = a.L.e_idx + v.eo;
  const uint16_t* __restrict__ grlen = a.L.r_len + v.ro;
  const double* __restrict__ gw = a.L.r_w + v.ro;
  const uint16_t* __restrict__ gjptr = a.L.j_ptr + v.jo;
  const uint16_t* __restrict__ gcptr = a.L.c_ptr + v.co;
  const double* __restrict__ ps0 = a.L.c_pisum0 + v.co;
  const int* __restrict__ scol = a.L.c_scol + v.co;
  const int col0 = a.L.seg_col0[seg];
  const uint16_t* __restrict__ gxref = ca.xref + ca.seg_xref_off[seg];
  const bool has_uniq = a.L.seg_nuniq[seg] > 0;

  const double invK = 1.0 / (double)a.P.K;
  for (int c = tid; c < v.ncol; c += THREADS) {
    spi[c] = invK;
    xb0[c] = invK * invK;
    sps0[c] = has_uniq ? ps0[c] : 0.0;
  }
  // The shared columns as a list in ascending order, and for each of them the first of its receive slots -- the other
  // holders' partial sums arrive in consecutive slots, in tile order (ordered compaction chunk by chunk: a packed
  // warp scan of (shared ? 1 : 0) | (other holders << 16)).  xr[c][k] = local index of column c in tile k (0xffff:
  // tile k does not hold it) for now.
  __shared__ int s_wcnt[NW + 1];
  int n_sh = 0, n_recv = 0;
  {
    int run = 0;                                           // packed totals of the chunks before
    for (int c0 = 0; c0 < v.ncol; c0 += THREADS) {
      const int c = c0 + tid;
      int others = 0;
      bool not_first = false;
      if (c < v.ncol) {
        const long long j = scol[c] - col0;
        for (int k = 0; k < nt; ++k) {
          const uint16_t x = gxref[j * nt + k];
          xr[c * nt + k] = x;
          const bool held = x != 0xffffu;
          others += (k != rank && held) ? 1 : 0;
          not_first |= k < rank && held;
        }
        shr[c] = others > 0 ? 1 : 0;
      }
      const int mine = others > 0 ? (1 | (others << 16)) : 0;
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(full, incl, o);
        if (lane >= o) incl += up;
      }
      if (lane == 31) s_wcnt[warp] = incl;
      __syncthreads();
      int base = run, tot = 0;
      for (int w = 0; w < NW; ++w) {
        const int cw = s_wcnt[w];
        if (w < warp) base += cw;
        tot += cw;
      }
      if (others > 0) {
        const int at = base + incl - mine;                 // exclusive prefix: list position | first receive slot << 16
        shl[at & 0xffff] = (uint16_t)(c | (not_first ? 0x8000 : 0));
        shp[at & 0xffff] = gcptr[c];
        rb[c] = (uint16_t)(at >> 16);
      }
      run += tot;
      __syncthreads();
    }
    n_sh = run & 0xffff;
    n_recv = run >> 16;
  }
  if (tid == 0) {
    xb0[v.ncol] = 0.0;
    xb1[v.ncol] = 0.0;
    xb2[v.ncol] = 0.0;
    mbar_init(&s_mbar[0], 1);
    mbar_init(&s_mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 2 * SS_MAX_CLUSTER_TILES) (&s_dall[0][0])[tid] = 0.0;
  __syncthreads();
  cluster.sync();                              // every CTA's lists and mbarriers are in place
  // xr[c][k] := the slot of tile k's receive buffer that takes MY partial sum of column c (0xfffe: k is this tile,
  // 0xffff: tile k does not hold the column): tile k's first slot of the column (read from its shared memory) + the
  // number of holders before me other than k
  for (int i = tid; i < n_sh; i += THREADS) {
    const int c = shl[i] & 0x7fff;
    uint16_t* xc_ = xr + c * nt;
    int before = 0;                                        // holders j < rank
    for (int k = 0; k < rank; ++k) before += xc_[k] != 0xffffu ? 1 : 0;
    for (int k = 0; k < nt; ++k) {
      const unsigned ck = xc_[k];
      if (k == rank) { xc_[k] = 0xfffeu; continue; }
      if (ck == 0xffffu) continue;
      const uint32_t rbk = ld_dsmem_u16(mapa_u32(smem_u32(rb + ck), (uint32_t)k));
      xc_[k] = (uint16_t)(rbk + before - (k < rank ? 1 : 0));
    }
  }
  const double thpw = a.F.seg_thpw[seg], pipw = a.F.seg_pipw[seg];
  const double inv_thd = a.F.seg_inv_thd[seg], inv_pid = a.F.seg_inv_pid[seg];
  const double diff1_extra = a.F.seg_diff1_extra[seg];
  const double eps = a.P.eps;
  const int max_iter = a.P.max_iter;
  double* diffs = a.F.seg_diffs + (long long)seg * max_iter;

  // ---- row fragments (as in the lane kernel) ----
  const int rgt4 = (int)hd[H_RGT4], rgt8 = (int)hd[H_RGT8], rgt16 = (int)hd[H_RGT16], rgt32 = (int)hd[H_RGT32];
  const int n_huge = rgt32;
  const int n8 = rgt16 - rgt32, n4 = rgt8 - rgt16, n2 = rgt4 - rgt8, n1 = v.nrows - rgt4;
  const int V8 = 8 * n8, V4 = V8 + 4 * n4, V2 = V4 + 2 * n2, V = V2 + n1;
  const int Vc = min(V, FR * THREADS);
  int r_ovf;
  if (Vc >= V) r_ovf = v.nrows;
  else if (Vc < V8) r_ovf = n_huge + Vc / 8;
  else if (Vc < V4) r_ovf = n_huge + n8 + (Vc - V8) / 4;
  else if (Vc < V2) r_ovf = n_huge + n8 + n4 + (Vc - V4) / 2;
  else r_ovf = n_huge + n8 + n4 + n2 + (Vc - V2);
  const bool generic_rows = n_huge > 0 || r_ovf < v.nrows;
  const uint32_t dummy_off = ((uint32_t)v.ncol << 3) | ((uint32_t)v.pe << 19);
  double fq[FR][FRAG], fw[FR];
  uint32_t fo[FR][FRAG];
  int fg[FR], nA[FR], gA[FR];
#pragma unroll
  for (int j = 0; j < FR; ++j) {
    const int vl = tid + j * THREADS;
    int n = 0;
    fg[j] = 1; fw[j] = 0.0;
#pragma unroll
    for (int p = 0; p < FRAG; ++p) { fq[j][p] = 0.0; fo[j][p] = dummy_off; }
    if (vl < Vc) {
      int r, k, g;
      if (vl < V8) { g = 8; r = n_huge + vl / 8; k = vl % 8; }
      else if (vl < V4) { g = 4; r = n_huge + n8 + (vl - V8) / 4; k = (vl - V8) % 4; }
      else if (vl < V2) { g = 2; r = n_huge + n8 + n4 + (vl - V4) / 2; k = (vl - V4) % 2; }
      else { g = 1; r = n_huge + n8 + n4 + n2 + (vl - V2); k = 0; }
      const int len = grlen[r];
      n = max(0, min(FRAG, len - FRAG * k));
      fg[j] = g; fw[j] = gw[r];
#pragma unroll
      for (int p = 0; p < FRAG; ++p)
        if (p < n) {
          const int e = (int)gjptr[FRAG * k + p] + r;
          const uint32_t ix = gidx[e];
          fq[j][p] = gq[e];
          fo[j][p] = ((ix & 0xffffu) << 3) | ((ix >> 16) << 19);
        }
    }
    nA[j] = __reduce_max_sync(full, n);
    gA[j] = __reduce_max_sync(full, fg[j]);
  }
  // ---- column fragments: descriptors only (the column state lives in shared memory) ----
  const int b128 = (int)hd[H_CGT128];
  const int b64 = (int)hd[H_CGT64], b32 = (int)hd[H_CGT32], b16 = (int)hd[H_CGT16], b8 = (int)hd[H_CGT8],
            b4 = (int)hd[H_CGT4];
  const int W32 = 32 * (b64 - b128), W16 = W32 + 16 * (b32 - b64), W8 = W16 + 8 * (b16 - b32);
  const int W4 = W8 + 4 * (b8 - b16), W2 = W4 + 2 * (b4 - b8), W = W2 + (v.ncol - b4);
  const int Wc = min(W, FC * THREADS);
  int c_ovf;
  if (Wc >= W) c_ovf = v.ncol;
  else if (Wc < W32) c_ovf = b128 + Wc / 32;
  else if (Wc < W16) c_ovf = b64 + (Wc - W32) / 16;
  else if (Wc < W8) c_ovf = b32 + (Wc - W16) / 8;
  else if (Wc < W4) c_ovf = b16 + (Wc - W8) / 4;
  else if (Wc < W2) c_ovf = b8 + (Wc - W4) / 2;
  else c_ovf = b4 + (Wc - W2);
  // byte offset of the first slot; #slots | group << 4 | (leader ? column + 1 : 0) << 12.  (Packing the warp maxima
  // nB / gB into cm as well, the row descriptors fg / nA / gA into one word and parking the per-segment constants in
  // shared memory was measured: the spill count of the 64-register variant did not move and the extra field
  // extraction cost 3.5 % -- config 5 x0.2 90.2 against 87.1 ms, profiles/r2l_descriptor_packing_experiment.log)
  int cb[FC], cm[FC], nB[FC], gB[FC];
#pragma unroll
  for (int j = 0; j < FC; ++j) {
    const int vl = tid + j * THREADS;
    int cnj = 0, cgj = 1, ccj = -1;
    cb[j] = 0;
    if (vl < Wc) {
      int c, k, g;
      if (vl < W32) { g = 32; c = b128 + vl / 32; k = vl % 32; }
      else if (vl < W16) { g = 16; c = b64 + (vl - W32) / 16; k = (vl - W32) % 16; }
      else if (vl < W8) { g = 8; c = b32 + (vl - W16) / 8; k = (vl - W16) % 8; }
      else if (vl < W4) { g = 4; c = b16 + (vl - W8) / 4; k = (vl - W8) % 4; }
      else if (vl < W2) { g = 2; c = b8 + (vl - W4) / 2; k = (vl - W4) % 2; }
      else { g = 1; c = b4 + (vl - W2); k = 0; }
      const int b = gcptr[c], n = (int)gcptr[c + 1] - b;
      cb[j] = 8 * (b + FRAG * k);
      cnj = max(0, min(FRAG, n - FRAG * k));
      cgj = g;
      ccj = k == 0 ? c : -1;
    }
    cm[j] = cnj | (cgj << 4) | ((ccj + 1) << 12);
    nB[j] = __reduce_max_sync(full, cnj);
    gB[j] = __reduce_max_sync(full, cgj);
  }
  __syncthreads();

  double* x_prev = xb2;                        // x_{it-2}
  double* x_cur = xb0;                         // x_{it-1}: what the E-step of iteration `it` reads
  double* x_nxt = xb1;                         // x_it
  unsigned char* const vb = reinterpret_cast<unsigned char*>(val);
  const uint32_t rpart_s = smem_u32(rpart), mbar_s = smem_u32(&s_mbar[0]), dall_s = smem_u32(&s_dall[0][0]);
  int it = 0, fin = 0, par = 0, xc = 0;
  bool conv = false;
  bool final_pass = false;                     // the pass that repeats iteration `fin` to publish its parameters
  bool dsum_sent = false;                      // the previous pass pushed |delta pi| partials (they complete on this pass's mbarrier)
  double dsum = 0.0;
#ifdef SS_CLANE_PROFILE
  long long tA = 0, tB1 = 0, tS = 0, tB2 = 0, tR = 0, c0, c1;
#define SS_TICK(acc) c1 = clock64(); acc += c1 - c0; c0 = c1;
  c0 = clock64();
#else
#define SS_TICK(acc)
#endif
  // new parameters of tile column c from its column sum `tot` (model.py:240-246).  Iterating: pi' and x' = pi' theta'
  // go to shared memory, |delta pi| is added up (`count`: this CTA is the first holder of the column).  Final pass:
  // pi' and theta' of iteration `fin` go to the segment's column table.
  auto update_col = [&](int c, double tot, const double* x_in, bool count) {
    const double T = x_in[c] * tot;
    const double th = (T + thpw) * inv_thd;
    const double pn = (sps0[c] + T + pipw) * inv_pid;
    if (!final_pass) {
      if (count) dsum += fabs(pn - spi[c]);
      spi[c] = pn;
      x_nxt[c] = pn * th;
    } else if (count) {
      const int slot = scol[c];
      a.F.sc_pi[slot] = pn;
      a.F.sc_theta[slot] = th;
    }
  };
  while (true) {
    if (!final_pass) ++it;
    ++xc;
    const int pp = xc & 1;                                        // parity of this pass: receive buffer and mbarrier
    const bool work = final_pass || it <= max_iter;               // phases A / B run (not in the pass after max_iter)
    if (tid == 0) {
      // arm this pass's mbarrier: the partial sums of my shared columns + the |delta pi| partials of the pass before
      const uint32_t bytes = (work ? 8u * (uint32_t)n_recv : 0u) + (dsum_sent ? 8u * (uint32_t)(nt - 1) : 0u);
      if (bytes) mbar_expect_tx(&s_mbar[pp], bytes);
      else mbar_arrive(&s_mbar[pp]);
    }
    // my partial sum of shared column c goes to the other holders
    auto push_col = [&](int c, double asum) {
      const uint16_t* xc_ = xr + c * nt;
#pragma unroll
      for (int k = 0; k < SS_MAX_CLUSTER_TILES; ++k)
        if (k < nt) {
          const unsigned sl = xc_[k];
          if (sl < 0xfffeu)
            st_async_f64(mapa_u32(rpart_s + 8u * (uint32_t)(pp * xmax + (int)sl), (uint32_t)k), asum,
                         mapa_u32(mbar_s + 8u * (uint32_t)pp, (uint32_t)k));
        }
    };
    const double* const x_in = final_pass ? x_prev : x_cur;
    if (work) {
      // ---- phase A ----
      const unsigned char* const xb = reinterpret_cast<const unsigned char*>(x_in);
#pragma unroll
      for (int j = 0; j < FR; ++j) {
        if (nA[j] > 0) {
          double acc = 0.0;
#pragma unroll
          for (int p = 0; p < FRAG; ++p)
            if (p < nA[j]) acc = fma(fq[j][p], *reinterpret_cast<const double*>(xb + (fo[j][p] & 0xffffu)), acc);
          if (gA[j] > 1) acc = group_xor_sum(acc, fg[j], gA[j]);
          const bool pos = acc > 0.0;
          double inv = 1.0 / (pos ? acc : 1.0);
          inv = pos ? inv : 0.0;
          const double sc = fw[j] * inv;
#pragma unroll
          for (int p = 0; p < FRAG; ++p)
            if (p < nA[j]) *reinterpret_cast<double*>(vb + (fo[j][p] >> 16)) = fq[j][p] * sc;
        }
      }
      if (generic_rows) {
        for (int r = tid; r < v.nrows; r += THREADS) {
          const int rr = r < n_huge ? r : r_ovf + (r - n_huge);
          if (rr >= v.nrows) break;
          const int len = grlen[rr];
          double acc = 0.0;
          for (int p = 0; p < len; ++p) {
            const int e = (int)gjptr[p] + rr;
            acc = fma(gq[e], x_in[gidx[e] & 0xffffu], acc);
          }
          const double sc = acc > 0.0 ? gw[rr] * (1.0 / acc) : 0.0;
          for (int p = 0; p < len; ++p) {
            const int e = (int)gjptr[p] + rr;
            val[gidx[e] >> 16] = gq[e] * sc;
          }
        }
      }
      __syncthreads();
      SS_TICK(tA)
      // ---- phase B: column sums.  The sum is parked in the column's first scratch slot; a column other tiles hold
      // too is pushed to them first, the others are finished by the thread that parked them ----
#pragma unroll
      for (int j = 0; j < FC; ++j) {
        if (nB[j] > 0) {
          double asum = 0.0;
          unsigned char* const cp = vb + cb[j];
          const int cnj = cm[j] & 15;
#pragma unroll
          for (int k = 0; k < FRAG; ++k)
            if (k < nB[j]) {
              if (k < cnj) asum += *reinterpret_cast<const double*>(cp + 8 * k);
            }
          if (gB[j] > 1) asum = group_xor_sum(asum, (cm[j] >> 4) & 63, gB[j]);
          const int c1 = cm[j] >> 12;
          if (c1 > 0) {
            if (shr[c1 - 1]) *reinterpret_cast<double*>(cp) = asum;
            else update_col(c1 - 1, asum, x_in, true);
          }
        }
      }
      for (int c = warp; c < b128; c += NW) {                     // very long columns: one warp each
        const int b = gcptr[c], n = (int)gcptr[c + 1] - b;
        const double asum = col_sum_warp(val, b, n, lane);
        if (lane == 0) {
          if (shr[c]) val[b] = asum;
          else update_col(c, asum, x_in, true);
        }
      }
      for (int c = c_ovf + tid; c < v.ncol; c += THREADS) {       // columns beyond the register lanes
        const int b = gcptr[c], n = (int)gcptr[c + 1] - b;
        const double asum = col_sum_thread(val, b, n);
        if (shr[c]) val[b] = asum;
        else update_col(c, asum, x_in, true);
      }
    }
    SS_TICK(tB1)
    __syncthreads();                                              // the parked sums of the shared columns are visible
    // push the sums of the shared columns to their other holders: one lane per list entry.  (Pushing from the column
    // lanes as soon as a sum is known put ~20 instructions per holder and slot into EVERY warp -- nearly all warps have
    // a shared column among their 96 -- about a quarter of the kernel's instructions.)
    if (work)
      for (int i = tid; i < n_sh; i += THREADS) push_col(shl[i] & 0x7fff, val[shp[i]]);
    mbar_wait_cluster(&s_mbar[pp], (uint32_t)((xc - 1) >> 1) & 1u);
    SS_TICK(tS)
    if (!final_pass && it > 1) {
      // ---- decision for iteration it-1 (all CTAs add the partials in rank order) ----
      double diff = 0.0;
      for (int r = 0; r < nt; ++r) diff += s_dall[pp][r];
      if (it - 1 == 1) diff += diff1_extra;
      if (rank == 0 && tid == 0) diffs[it - 2] = diff;
      conv = diff < eps;                                          // model.py:351
      if (conv || it - 1 >= max_iter) {
        // iteration fin = it-1 is final: x_fin = x_cur, the E-step that counts read x_prev.  What this pass wrote
        // (x_nxt, pi of the unshared columns) is discarded; one more pass repeats iteration fin from x_prev -- the
        // same sums in the same order -- and publishes pi and theta
        fin = it - 1;
        final_pass = true;
        dsum_sent = false;
        continue;
      }
    }
    // ---- shared columns: my partial sum and the ones the other holders pushed, added in tile order ----
    if (work) {
      const double* const rp = rpart + pp * xmax;
      for (int i = tid; i < n_sh; i += THREADS) {
        const int c = shl[i] & 0x7fff;
        const bool first = (shl[i] & 0x8000) == 0;                // no earlier tile holds this column
        const uint16_t* xc_ = xr + c * nt;
        int sl = rb[c];
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < SS_MAX_CLUSTER_TILES; ++k)
          if (k < nt) {
            const unsigned e = xc_[k];
            if (e == 0xfffeu) tot += val[shp[i]];
            else if (e != 0xffffu) tot += rp[sl++];
          }
        update_col(c, tot, x_in, first);
      }
    }
    SS_TICK(tB2)
    if (final_pass) break;
    dsum = block_sum<THREADS>(dsum, s_red, par++);                // barrier: publishes x_nxt
    if (tid < nt) {                                               // |delta pi| partial to every CTA; it completes on the NEXT pass's mbarrier
      const uint32_t off = 8u * (uint32_t)((pp ^ 1) * SS_MAX_CLUSTER_TILES + rank);
      if (tid == rank) s_dall[pp ^ 1][rank] = dsum;
      else st_async_f64(mapa_u32(dall_s + off, (uint32_t)tid), dsum, mapa_u32(mbar_s + 8u * (uint32_t)(pp ^ 1), (uint32_t)tid));
    }
    dsum_sent = true;
    dsum = 0.0;
    double* tmp = x_prev; x_prev = x_cur; x_cur = x_nxt; x_nxt = tmp;
    SS_TICK(tR)
  }
#ifdef SS_CLANE_PROFILE
  if (tid == 0 && blockIdx.x < 4)
    printf("clane seg %d rank %d/%d nent %d ncol %d iters %d: cycles/iter A %lld B1 %lld sync %lld B2 %lld red %lld\n", seg, rank,
           nt, v.nent, v.ncol, it, tA / it, tB1 / it, tS / it, tB2 / it, tR / it);
#endif
#undef SS_TICK
  // ---- publish: iteration `fin` is final.  x_cur = x_fin, x_prev = x_{fin-1} (the last E-step that counts) ----
  for (int c = tid; c < v.ncol; c += THREADS) {
    const int slot = scol[c];
    a.F.sc_ptprev[slot] = x_prev[c];
    a.F.sc_ptfin[slot] = x_cur[c];
  }
  // final lnL: z of the last E-step (x_{fin-1}), final parameters (model.py:368-370)
  double l = 0.0;
  for (int r = tid; r < v.nrows; r += THREADS) {
    const int len = grlen[r];
    double acc = 0.0;
    for (int p = 0; p < len; ++p) {
      const int e = (int)gjptr[p] + r;
      acc = fma(gq[e], x_prev[gidx[e] & 0xffffu], acc);
    }
    const double inv = acc > 0.0 ? 1.0 / acc : 0.0;
    for (int p = 0; p < len; ++p) {
      const int e = (int)gjptr[p] + r;
      const int c = gidx[e] & 0xffffu;
      const double z = (gq[e] * x_prev[c]) * inv;
      const double pn = x_cur[c];
      if (z > 0.0 && pn > 0.0) l = fma(z, log(gq[e] * pn), l);
    }
  }
  l = block_sum<THREADS>(l, s_red, par++);
  if (tid == 0) {
    a.F.tile_lnl[t] = l;
    if (rank == 0) {
      a.F.seg_iters[seg] = fin;
      a.F.seg_flags[seg] = (conv ? 1 : 0) | (fin >= max_iter ? 2 : 0);
    }
  }
  cluster.sync();                                                 // nobody leaves while its shared memory may be read
}

// ---------------------------------------------------------------------------
// three-barrier cluster kernel (likelihood-convergence mode only): one thread-block CLUSTER per segment of 2..8 tiles.  Every CTA keeps its
// tile resident in shared memory (like the resident kernel); the columns of the segment are
// owned round-robin by the CTAs (owner of segment column j = j mod nt).  Per iteration:
//   phase A (local rows) | barrier | phase B (local partial column sums -> `part`)
//   cluster.sync | owners pull the partial sums of their columns from every tile over DSMEM
//   in fixed tile order (deterministic), compute theta', pi', x', |delta pi|
//   cluster.sync | every CTA pulls x' of its tile columns from the owners and the nt partial
//   |delta pi| sums, takes the (identical) convergence decision.
// No global memory traffic and no atomics inside the loop.
// ---------------------------------------------------------------------------

template <int THREADS, bool LNL, int NT>
__global__ void __launch_bounds__(THREADS) em_cluster_kernel(EmArgs a, ClusterArgs ca) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double s_red[64];
  __shared__ double s_dsum[2];          // [0] |delta pi| of the owned columns, [1] lnL part
  __shared__ uint64_t s_bar;
  constexpr int NW = THREADS / 32;
  constexpr int nt = NT;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster.block_rank();
  const int seg = ca.segs[blockIdx.x / nt];
  const int t = a.L.seg_tile0[seg] + rank;
  const TileView v = load_tile_view(a.L.tile_hdr, t);
  TileSmem s = carve_tile(dsm, v);
  const int Kc = a.L.seg_ncol[seg], col0 = a.L.seg_col0[seg];
  const int no_max = (Kc + nt - 1) / nt;                    // owned columns, upper bound (same on every rank)
  const int no_pad = pad8(no_max + 1);
  const int n_owned = rank < Kc ? (Kc - rank + nt - 1) / nt : 0;
  // owner arrays live at the same offset in every CTA of the cluster (largest tile of the segment)
  size_t tile_bytes = 0;
  int aux_off[NT];                                           // offset (doubles) of `part` in every tile's image
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    const TileView vk = load_tile_view(a.L.tile_hdr, a.L.seg_tile0[seg] + k);
    const size_t bts = tile_smem_bytes(vk.nent, vk.nrows, vk.ncol, vk.maxlen);
    tile_bytes = bts > tile_bytes ? bts : tile_bytes;
    aux_off[k] = 2 * vk.pe + vk.pr + vk.pc;                  // q, val, w, pt precede aux (carve_tile)
  }
  tile_bytes = (tile_bytes + 127) & ~(size_t)127;
  double* own_x0 = reinterpret_cast<double*>(dsm + tile_bytes);
  double* own_x1 = own_x0 + no_pad;
  double* own_pi = own_x1 + no_pad;
  double* own_ps0 = own_pi + no_pad;
  uint16_t* own_xref = reinterpret_cast<uint16_t*>(own_ps0 + no_pad);   // [n_owned][NT]
  double* part = s.aux;                                      // partial column sums of this tile

  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) tile_issue_loads(a.L, v, s, &s_bar);
  const double invK = 1.0 / (double)a.P.K;
  const bool has_uniq = a.L.seg_nuniq[seg] > 0;
  const uint16_t* __restrict__ xref = ca.xref + ca.seg_xref_off[seg];
  for (int c = tid; c < v.ncol; c += THREADS) s.pt[c] = invK * invK;
  for (int i = tid; i < n_owned; i += THREADS) {
    const int j = rank + i * nt;
    own_x0[i] = invK * invK;
    own_pi[i] = invK;
    own_ps0[i] = has_uniq ? a.L.sc_pisum0[col0 + j] : 0.0;
#pragma unroll
    for (int k = 0; k < NT; ++k) own_xref[i * NT + k] = xref[(long long)j * nt + k];
  }
  const double thpw = a.F.seg_thpw[seg], pipw = a.F.seg_pipw[seg];
  const double inv_thd = a.F.seg_inv_thd[seg], inv_pid = a.F.seg_inv_pid[seg];
  const double diff1_extra = a.F.seg_diff1_extra[seg];
  const double eps = a.P.eps;
  const int max_iter = a.P.max_iter;
  const int* __restrict__ scol = a.L.c_scol + v.co;
  double* diffs = a.F.seg_diffs + (long long)seg * max_iter;
  double* lnls = a.F.seg_lnls ? a.F.seg_lnls + (long long)seg * max_iter : nullptr;
  double* dsm_d = reinterpret_cast<double*>(dsm);
  mbar_wait(&s_bar, 0);
  __syncthreads();
  cluster.sync();                                            // everybody's shared memory is initialised

  double* pt_cur = s.pt;
  double* pt_nxt = s.pt2;
  int cur = 0;                                               // own_x buffer holding the current x
  double lnl_prev = INFINITY;
  int it = 0, par = 0;
  bool done = false, conv = false;
  while (!done) {
    ++it;
    tile_phase_a<THREADS>(v, s, pt_cur);
    __syncthreads();
    for (int c = warp; c < v.nlong; c += NW) {
      const int b = s.cptr[c], n = (int)s.cptr[c + 1] - b;
      const double asum = col_sum_warp(s.val, b, n, lane);
      if (lane == 0) part[c] = asum;
    }
    for (int c = v.nlong + tid; c < v.ncol; c += THREADS) {
      const int b = s.cptr[c], n = (int)s.cptr[c + 1] - b;
      part[c] = col_sum_thread(s.val, b, n);
    }
    cluster.sync();
    // ---- owners: thetasum of the owned segment columns, new parameters ----
    double* ox_cur = cur ? own_x1 : own_x0;
    double* ox_nxt = cur ? own_x0 : own_x1;
    double dsum = 0.0;
    for (int i = tid; i < n_owned; i += THREADS) {
      double pk[NT];
      bool in_tile = false;
#pragma unroll
      for (int k = 0; k < NT; ++k) {                         // all remote loads in flight together
        const unsigned c = own_xref[i * NT + k];
        pk[k] = 0.0;
        if (c != 0xffffu) {
          pk[k] = cluster.map_shared_rank(dsm_d + aux_off[k], k)[c];
          in_tile = true;
        }
      }
      if (!in_tile) continue;       // column of unique rows only: constant, handled by prepare / finalize
      double asum = 0.0;
#pragma unroll
      for (int k = 0; k < NT; ++k) asum += pk[k];             // fixed tile order
      const double T = ox_cur[i] * asum;
      const double th = (T + thpw) * inv_thd;                           // model.py:240
      const double pn = (own_ps0[i] + T + pipw) * inv_pid;              // model.py:243-244
      dsum += fabs(pn - own_pi[i]);
      own_pi[i] = pn;
      ox_nxt[i] = pn * th;
    }
    dsum = block_sum<THREADS>(dsum, s_red, par++);
    if (tid == 0) s_dsum[0] = dsum;
    cluster.sync();
    // ---- everybody: x' of the tile's columns, convergence ----
    double diff = 0.0;
#pragma unroll
    for (int k = 0; k < NT; ++k) diff += cluster.map_shared_rank(s_dsum, k)[0];
    if (it == 1) diff += diff1_extra;
    {
      const double* ox = cur ? own_x0 : own_x1;              // buffer written in this iteration
      for (int c = tid; c < v.ncol; c += THREADS) {
        const int j = scol[c] - col0;
        pt_nxt[c] = cluster.map_shared_rank(ox, j % nt)[j / nt];
      }
    }
    const bool reached = it >= max_iter;
    if (!LNL) {
      conv = diff < eps;
      __syncthreads();
    } else {
      __syncthreads();
      double l = tile_lnl_pass<THREADS>(v, s, pt_cur, pt_nxt, true, nullptr, nullptr);
      if (has_uniq)
        for (int i = tid; i < n_owned; i += THREADS) {
          bool in_tile = false;
#pragma unroll
          for (int k = 0; k < NT; ++k) in_tile |= own_xref[i * NT + k] != 0xffffu;
          const int nu = a.L.sc_nuniq[col0 + rank + i * nt];
          if (in_tile && nu > 0 && own_pi[i] > 0.0) l += (double)nu * log(own_pi[i]);
        }
      l = block_sum<THREADS>(l, s_red, par++);
      if (tid == 0) s_dsum[1] = l;
      cluster.sync();
      double lnl = a.F.seg_lnl_extra[seg];
#pragma unroll
      for (int k = 0; k < NT; ++k) lnl += cluster.map_shared_rank(s_dsum, k)[1];
      conv = fabs(lnl - lnl_prev) < eps;
      lnl_prev = lnl;
      if (rank == 0 && tid == 0 && lnls) lnls[it - 1] = lnl;
    }
    done = conv || reached;
    if (rank == 0 && tid == 0) diffs[it - 1] = diff;
    if (!done) {
      double* tmp = pt_cur; pt_cur = pt_nxt; pt_nxt = tmp;
      cur ^= 1;
    }
  }
  // ---- publish ----
  {
    const double* ox_cur = cur ? own_x1 : own_x0;
    const double* ox_nxt = cur ? own_x0 : own_x1;
    for (int i = tid; i < n_owned; i += THREADS) {
      const int j = rank + i * nt;
      bool in_tile = false;
#pragma unroll
      for (int k = 0; k < NT; ++k) in_tile |= own_xref[i * NT + k] != 0xffffu;
      if (!in_tile) continue;
      a.F.sc_pi[col0 + j] = own_pi[i];
      a.F.sc_ptprev[col0 + j] = ox_cur[i];
      a.F.sc_ptfin[col0 + j] = ox_nxt[i];
      // theta' = x' / pi' would lose the last bit: recompute it like the iteration did
      const double xo = ox_cur[i];
      double asum = 0.0;
#pragma unroll
      for (int k = 0; k < NT; ++k) {
        const unsigned c = own_xref[i * NT + k];
        if (c != 0xffffu) asum += cluster.map_shared_rank(dsm_d + aux_off[k], k)[c];
      }
      a.F.sc_theta[col0 + j] = (xo * asum + thpw) * inv_thd;
    }
  }
  double l = tile_lnl_pass<THREADS>(v, s, pt_cur, pt_nxt, true, nullptr, nullptr);
  l = block_sum<THREADS>(l, s_red, par++);
  if (tid == 0) {
    a.F.tile_lnl[t] = l;
    if (rank == 0) {
      a.F.seg_iters[seg] = it;
      a.F.seg_flags[seg] = (conv ? 1 : 0) | (it >= max_iter ? 2 : 0);
    }
  }
  cluster.sync();                                            // nobody leaves while its shared memory may be read
}

// ---------------------------------------------------------------------------
// streaming kernel: cooperative grid over the tiles of multi-tile segments
// ---------------------------------------------------------------------------
template <int THREADS, bool LNL>
__global__ void __launch_bounds__(THREADS, LNL ? 1 : 2) em_stream_kernel(EmArgs a, StreamArgs sa) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double s_red[64];
  __shared__ uint64_t s_bar;
  constexpr int NW = THREADS / 32;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long gtid = (long long)blockIdx.x * THREADS + tid;
  const long long gthreads = (long long)gridDim.x * THREADS;
  const int max_iter = a.P.max_iter;
  const double eps = a.P.eps;
  __shared__ StreamPipe s_pipe;
  uint32_t phase = 0;
  uint32_t n_use = 0u;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    stream_pipe_init(&s_pipe, THREADS / 32 - STREAM_NPROD);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  for (int it = 1; it <= max_iter; ++it) {
    const int cur = (it - 1) & 1;
    const double* __restrict__ ptg = a.F.sc_pt[cur];
    // ---- tile pass: phase A + B, partial thetasum -> global accumulators (double-buffered TMA, lane fragments) ----
    lane_stream_pass<THREADS, SS_STREAM_FR, SS_STREAM_FC>(a.L, a.L.c_scol, sa.tiles, sa.n_tiles, a.F.seg_iters, ptg, a.F.sc_T, sa.geom, dsm,
                                    &s_pipe, n_use);
    grid.sync();
    // ---- update pass over the columns of the active multi-tile segments: one warp per column chunk of one segment
    //      (per-segment constants loaded once, coalesced column loads, four columns per lane in flight; a converged
    //      segment costs one load per chunk) ----
    double* __restrict__ ptn = a.F.sc_pt[cur ^ 1];
    {
      const int gwarp = (int)(gtid >> 5), nwarps = (int)(gthreads >> 5);
      for (int ci = gwarp; ci < sa.n_chunks; ci += nwarps) {
        const int seg = sa.chunk_seg[ci];
        if (a.F.seg_iters[seg] != 0) continue;
        const int c0 = a.L.seg_col0[seg], nc = a.L.seg_ncol[seg];
        const int jb = sa.chunk_j0[ci], je = min(nc, jb + SS_COL_CHUNK);
        const double thpw = a.F.seg_thpw[seg], inv_thd = a.F.seg_inv_thd[seg];
        const double pipw = a.F.seg_pipw[seg], inv_pid = a.F.seg_inv_pid[seg];
        const double pi0 = 1.0 / (double)a.P.K;                  // pi_0 = 1/K for every locus (model.py:164)
        double d = 0.0;
        for (int j0 = jb; j0 < je; j0 += 128) {
          double T[4], p0[4], po[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + 32 * u + lane;
            const bool ok = j < je;
            T[u] = ok ? a.F.sc_T[c0 + j] : 0.0;
            p0[u] = ok ? a.L.sc_pisum0[c0 + j] : 0.0;
            po[u] = (ok && it > 1) ? a.F.sc_pi[c0 + j] : pi0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + 32 * u + lane;
            if (j < je) {
              const int slot = c0 + j;
              const double th = (T[u] + thpw) * inv_thd;
              const double pn = (p0[u] + T[u] + pipw) * inv_pid;
              d += fabs(pn - po[u]);
              a.F.sc_pi[slot] = pn;
              a.F.sc_theta[slot] = th;
              ptn[slot] = pn * th;
              a.F.sc_T[slot] = 0.0;
            }
          }
        }
        d = warp_sum(d);
        if (lane == 0) atomicAdd(&a.F.seg_diffacc[seg], d);
      }
    }
    grid.sync();
    if (LNL) {
      // second tile pass: lnL with z from x_cur and the parameters x_next
      for (int i = blockIdx.x; i < sa.n_tiles; i += gridDim.x) {
        const int t = sa.tiles[i];
        const TileView v = load_tile_view(a.L.tile_hdr, t);
        if (a.F.seg_iters[v.seg] != 0) continue;
        TileSmem s = carve_tile(dsm, v);
        if (tid == 0) {
          fence_proxy_async();
          tile_issue_loads(a.L, v, s, &s_bar);
        }
        const int* __restrict__ scol = a.L.c_scol + v.co;
        for (int c = tid; c < v.ncol; c += THREADS) {
          s.pt[c] = ptg[scol[c]];
          s.pt2[c] = ptn[scol[c]];
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        __syncthreads();
        double l = tile_lnl_pass<THREADS>(v, s, s.pt, s.pt2, true, nullptr, nullptr);
        l = block_sum<THREADS>(l, s_red, 0);
        if (tid == 0) atomicAdd(&a.F.seg_lnl[v.seg], l);
        __syncthreads();
      }
      // unique-row part over the columns
      for (long long x = gtid; x < sa.n_cols; x += gthreads) {
        const int seg = sa.segs[sa.col_seg[x]];
        if (a.F.seg_iters[seg] != 0) continue;
        const int slot = sa.col_slot[x];
        const int nu = a.L.sc_nuniq[slot];
        const double pn = a.F.sc_pi[slot];
        if (nu > 0 && pn > 0.0) atomicAdd(&a.F.seg_lnl[seg], (double)nu * log(pn));
      }
      grid.sync();
    }
    // ---- convergence decision, one thread per segment (block 0 only) ----
    if (blockIdx.x == 0) {
      int still = 0;
      for (int i = tid; i < sa.n_segs; i += THREADS) {
        const int seg = sa.segs[i];
        if (a.F.seg_iters[seg] != 0) continue;
        double diff = a.F.seg_diffacc[seg];
        a.F.seg_diffacc[seg] = 0.0;
        if (it == 1) diff += a.F.seg_diff1_extra[seg];
        a.F.seg_diffs[(long long)seg * max_iter + it - 1] = diff;
        bool conv;
        if (LNL) {
          const double lnl = a.F.seg_lnl[seg] + a.L.seg_logq_uniq[seg];
          const double prev = it == 1 ? INFINITY : a.F.seg_lnls[(long long)seg * max_iter + it - 2];
          a.F.seg_lnls[(long long)seg * max_iter + it - 1] = lnl;
          conv = fabs(lnl - prev) < eps;
          a.F.seg_lnl[seg] = 0.0;
        } else {
          conv = diff < eps;
        }
        const bool reached = it >= max_iter;
        if (conv || reached) {
          a.F.seg_flags[seg] = (conv ? 1 : 0) | (reached ? 2 : 0);
          a.F.seg_iters[seg] = it;
        } else {
          ++still;
        }
      }
      still = __syncthreads_count(still);
      if (tid == 0) sa.ctl[0] = still;
    }
    grid.sync();
    if (*((volatile int*)sa.ctl) == 0) break;
  }
}

// freeze x_prev / x_final of the multi-tile segments (buffer parity = iteration count)
__global__ void stream_freeze_kernel(FitState F, StreamArgs sa) {
  const long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= sa.n_cols) return;
  const int seg = sa.segs[sa.col_seg[x]];
  const int slot = sa.col_slot[x];
  const int it = F.seg_iters[seg];
  F.sc_ptprev[slot] = F.sc_pt[(it - 1) & 1][slot];
  F.sc_ptfin[slot] = F.sc_pt[it & 1][slot];
}

// emit for streamed tiles: final lnL part (+ z) from x_prev / x_final
template <int THREADS>
__global__ void __launch_bounds__(THREADS) emit_tiles_kernel(EmArgs a, int want_lnl) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double s_red[64];
  __shared__ uint64_t s_bar;
  const int tid = threadIdx.x;
  const int t = a.tile_list[blockIdx.x];
  const TileView v = load_tile_view(a.L.tile_hdr, t);
  TileSmem s = carve_tile(dsm, v);
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) tile_issue_loads(a.L, v, s, &s_bar);
  const int* __restrict__ scol = a.L.c_scol + v.co;
  for (int c = tid; c < v.ncol; c += THREADS) {
    s.pt[c] = a.F.sc_ptprev[scol[c]];
    s.pt2[c] = a.F.sc_ptfin[scol[c]];
  }
  mbar_wait(&s_bar, 0);
  __syncthreads();
  double l = tile_lnl_pass<THREADS>(v, s, s.pt, s.pt2, want_lnl != 0, a.z_out, a.L.e_opos + v.eo);
  if (want_lnl) {
    l = block_sum<THREADS>(l, s_red, 0);
    if (tid == 0) a.F.tile_lnl[t] = l;
  }
}

// z of unique rows: u * (1/u) (norm(1), sparse_plus.py:64) -- 1 up to rounding; 0 if pi_j == 0
__global__ void emit_unique_kernel(Layout L, double* __restrict__ z_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.n_uniq) return;
  z_out[L.u_pos[i]] = 1.0;
}

// unique rows on the tile columns of multi-tile segments: sum of n_uniq_j * log(pi_j), one warp per column
// chunk (a single tile adds this part itself -- lane kernel; the columns no tile touches are in seg_lnl_extra)
__global__ void fit_finalize_cols_kernel(Layout L, FitState F, const int8_t* __restrict__ sc_inamb, ColChunks cc) {
  const int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ch >= cc.n) return;
  const int s = cc.seg[ch];
  if (L.seg_ntiles[s] < 2) return;
  const int c0 = L.seg_col0[s], nc = L.seg_ncol[s];
  const int jb = cc.j0[ch], je = min(nc, jb + SS_COL_CHUNK);
  double l = 0.0;
  for (int j0 = jb; j0 < je; j0 += 128) {
    int nu[4];
    double p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u + lane;
      nu[u] = j < je && sc_inamb[c0 + j] ? L.sc_nuniq[c0 + j] : 0;
      p[u] = j < je ? F.sc_pi[c0 + j] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (nu[u] > 0 && p[u] > 0.0) l += (double)nu[u] * log(p[u]);
  }
  l = warp_sum(l);
  if (lane == 0) cc.part[ch] = l;
}

// per segment: final lnL = sum of tile parts + unique-row part; segments without tiles
// (no ambiguous rows) are solved in closed form.  One warp per segment.
__global__ void fit_finalize_kernel(Layout L, FitState F, FitParams P, ColChunks cc) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= L.S) return;
  const int s = warp;
  if (F.seg_flags[s] & 4) return;
  const int nt = L.seg_ntiles[s], t0 = L.seg_tile0[s];
  double l = 0.0;
  if (nt > 1) {
    const int ch0 = cc.seg_chunk0[s], nch = (L.seg_ncol[s] + SS_COL_CHUNK - 1) / SS_COL_CHUNK;
    for (int i = lane; i < nch; i += 32) l += cc.part[ch0 + i];
  }
  l = warp_sum(l);
  double lt = 0.0;
  for (int t = lane; t < nt; t += 32) lt += F.tile_lnl[t0 + t];
  lt = warp_sum(lt);
  if (lane == 0) {
    F.seg_lnl[s] = lt + l + F.seg_lnl_extra[s];
    if (nt == 0) {
      // T = 0 in every iteration: pi is constant after iteration 1
      const double d1 = F.seg_diff1_extra[s];
      double* diffs = F.seg_diffs + (long long)s * P.max_iter;
      int it = 1;
      bool conv;
      if (P.use_likelihood) {
        // lnL is the same in iterations 1 and 2 -> converged at 2
        double* lnls = F.seg_lnls + (long long)s * P.max_iter;
        diffs[0] = d1;
        lnls[0] = F.seg_lnl[s];
        conv = false;
        while (!conv && it < P.max_iter) {
          ++it;
          diffs[it - 1] = 0.0;
          lnls[it - 1] = F.seg_lnl[s];
          conv = 0.0 < P.eps;
        }
      } else {
        diffs[0] = d1;
        conv = d1 < P.eps;
        while (!conv && it < P.max_iter) {
          ++it;
          diffs[it - 1] = 0.0;
          conv = 0.0 < P.eps;
        }
      }
      F.seg_iters[s] = it;
      F.seg_flags[s] = (conv ? 1 : 0) | (it >= P.max_iter ? 2 : 0);
    }
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
void free_fit(ss_handle_s* h) {
  dev_free_all(h->fit_allocs);
  h->F = FitState();
  h->stream_ctl = nullptr;
  h->has_fit = false;
}


// (a preferred shared-memory carveout of 100 % for all EM kernels was measured: no difference, the driver's choice
// already lets two light cluster CTAs / two streaming CTAs share an SM)
template <class F>
static cudaError_t set_smem(F fn, size_t smem) {
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

template <int TT, int FR, int FC, int MAXREG>
static cudaError_t launch_lane(ss_handle_s* h, cudaStream_t st, const EmArgs& a, int n, int threads, size_t smem,
                               bool lnl) {
  if (n == 0) return cudaSuccess;
  cudaError_t e;
  if (lnl) {
    auto fn = em_lane_kernel<TT, true, FR, FC, MAXREG>;
    e = set_smem(fn, smem);
    if (e != cudaSuccess) return e;
    fn<<<n, threads, smem, st>>>(a);
  } else {
    auto fn = em_lane_kernel<TT, false, FR, FC, MAXREG>;
    e = set_smem(fn, smem);
    if (e != cudaSuccess) return e;
    fn<<<n, threads, smem, st>>>(a);
  }
  h->launches++;
  return cudaGetLastError();
}

cudaError_t ensure_aux_streams(ss_handle_s* h) {
  if (h->ev_fork) return cudaSuccess;
  cudaError_t e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  if (e != cudaSuccess) return e;
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  for (int i = 0; i < SS_NUM_AUX_STREAMS; ++i) {
    // the cluster / streaming kernels have few, long-running CTAs that each need a whole SM: the
    // block scheduler must place them before the many short CTAs of the lane kernels
    e = cudaStreamCreateWithPriority(&h->aux[i], cudaStreamNonBlocking, i >= SS_NUM_CLASSES ? prio_hi : prio_lo);
    if (e != cudaSuccess) return e;
    e = cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

constexpr int STREAM_THREADS = SS_STREAM_THREADS;
constexpr int CLUSTER_THREADS = 512;

int em_fit(ss_handle_s* h, cudaStream_t st, const FitParams& P) {
  if (!h->has_layout) {
    h->set_error("ss_em_fit: no layout (call ss_layout_build first)");
    return SS_ERR_STATE;
  }
  if (P.max_iter < 1 || !(P.eps == P.eps)) {
    h->set_error("ss_em_fit: max_iter must be >= 1 and em_epsilon a number");
    return SS_ERR_ARG;
  }
  Layout& L = h->L;
  FitState& F = h->F;
  const size_t S = (size_t)L.S, NC = (size_t)L.n_segcols, MI = (size_t)P.max_iter;
  // the fit state is allocated once per (layout, max_iter, likelihood mode) and reused by later fits
  const bool reuse = h->fit_allocs.size() > 0 && h->fit_S == S && h->fit_NC == NC && h->fit_MI == MI &&
                     h->fit_T == (size_t)L.n_tiles && h->fit_lnl == (P.use_likelihood != 0);
  if (!reuse) {
    free_fit(h);
    auto& pool = h->fit_allocs;
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.seg_thpw, S));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.seg_pipw, S));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.seg_inv_thd, S));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.seg_inv_pid, S));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.seg_diff1_extra, S));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.seg_lnl_extra, S));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.sc_pi, NC));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.sc_pt[0], NC));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.sc_pt[1], NC));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.sc_ptprev, NC));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.sc_ptfin, NC));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.sc_theta, NC));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.sc_T, NC));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.seg_iters, S));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.seg_flags, S));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.seg_lnl, S));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.seg_diffs, S * MI));
    if (P.use_likelihood) SS_CUDA_CHECK(h, dev_alloc(pool, &F.seg_lnls, S * MI));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.seg_diffacc, S));
    SS_CUDA_CHECK(h, dev_alloc(pool, &F.tile_lnl, (size_t)L.n_tiles));
    SS_CUDA_CHECK(h, dev_alloc(pool, &h->stream_ctl, 4));
    h->fit_S = S; h->fit_NC = NC; h->fit_MI = MI; h->fit_T = (size_t)L.n_tiles;
    h->fit_lnl = P.use_likelihood != 0;
  }
  h->has_fit = false;
  SS_CUDA_CHECK(h, cudaMemsetAsync(F.seg_diffs, 0, std::max<size_t>(1, S * MI) * sizeof(double), st));
  if (P.use_likelihood)
    SS_CUDA_CHECK(h, cudaMemsetAsync(F.seg_lnls, 0, std::max<size_t>(1, S * MI) * sizeof(double), st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(F.tile_lnl, 0, (size_t)std::max(1, L.n_tiles) * sizeof(double), st));
  h->P = P;
  h->P.K = L.K;

#define SS_MARK(i) if (h->timing) SS_CUDA_CHECK(h, cudaEventRecord(h->ev[i], st))
  SS_MARK(0);
  if (h->dense_mode()) {
    // one model, rows sharded over ranks (or forced on one rank): dense K-long parameter vectors
    for (int i = 0; i < SS_NUM_AUX_STREAMS; ++i) h->aux_used[i] = false;
    const int rc = em_fit_dense(h, st, h->P);
    if (rc != SS_OK) return rc;
    if (h->timing) {
      SS_MARK(3); SS_MARK(4); SS_MARK(5);
      h->ev_valid = true;
    }
    return SS_OK;
  }
  ColChunks chunks;
  chunks.n = h->n_col_chunks;
  chunks.seg = h->chunk_seg;
  chunks.j0 = h->chunk_j0;
  chunks.seg_chunk0 = h->seg_chunk0;
  chunks.part = h->chunk_part;
  if (L.S > 0) {
    const int warps_per_block = 8;
    if (h->n_col_chunks > 0) {
      fit_prepare_cols_kernel<<<(h->n_col_chunks + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(
          L, F, h->P, h->sc_inamb, chunks);
      h->launches++;
    }
    fit_prepare_kernel<<<(L.S + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(L, F, h->P, chunks);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
  }

  EmArgs a;
  a.L = L;
  a.F = F;
  a.P = h->P;
  a.z_out = nullptr;
  const bool lnl = P.use_likelihood != 0;

  // All EM kernels of the fit run concurrently: the cooperative streaming kernel (segments
  // spanning several tiles) and one resident kernel per size class, each on its own stream
  // forked from / joined back into the caller's stream.
  SS_MARK(1);
  SS_CUDA_CHECK(h, ensure_aux_streams(h));
  SS_CUDA_CHECK(h, cudaEventRecord(h->ev_fork, st));
  for (int i = 0; i < SS_NUM_AUX_STREAMS; ++i) h->aux_used[i] = false;
  if (h->n_stream_tiles > 0) {
    cudaStream_t sx = h->aux[SS_NUM_CLASSES];
    h->aux_used[SS_NUM_CLASSES] = true;
    SS_CUDA_CHECK(h, cudaStreamWaitEvent(sx, h->ev_fork, 0));
    if (h->timing) SS_CUDA_CHECK(h, cudaEventRecord(h->tev_s[SS_NUM_CLASSES], sx));
    StreamArgs sa;
    sa.tiles = h->stream_tiles;
    sa.n_tiles = h->n_stream_tiles;
    sa.segs = h->stream_segs;
    sa.n_segs = h->n_stream_segs;
    sa.col_seg = h->stream_col_seg;
    sa.col_slot = h->stream_col_slot;
    sa.n_cols = h->n_stream_cols;
    sa.ctl = h->stream_ctl;
    sa.chunk_seg = h->stream_chunk_seg;
    sa.chunk_j0 = h->stream_chunk_j0;
    sa.n_chunks = h->n_stream_chunks;
    SS_CUDA_CHECK(h, cudaMemsetAsync(h->stream_ctl, 0, 4 * sizeof(int), sx));
    void* fn = lnl ? (void*)em_stream_kernel<STREAM_THREADS, true> : (void*)em_stream_kernel<STREAM_THREADS, false>;
    int occ = 0;
    const size_t smem = pick_stream_geom(h->geom_stream, h->max_smem_optin - 2048, lnl ? h->stream_smem : 0, fn, STREAM_THREADS,
                                         &sa.geom, &occ);
    if (occ < 1) {
      h->set_error("streaming kernel does not fit on an SM");
      return SS_ERR_CAPACITY;
    }
    SS_CUDA_CHECK(h, set_smem(fn, smem));
    // (capping the cooperative grid at one CTA per SM while cluster / lane kernels run beside it was measured: config 5
    // x0.1 45.5 -> 43.1 ms, x0.3 130.7 -> 140.2 ms -- dropped)
    int grid = std::min(h->n_stream_tiles, occ * h->num_sms);
    a.tile_list = h->stream_tiles;
    a.n_list = h->n_stream_tiles;
    void* args[] = {(void*)&a, (void*)&sa};
    SS_CUDA_CHECK(h, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(STREAM_THREADS), args, smem, sx));
    h->launches++;
    stream_freeze_kernel<<<(unsigned)((h->n_stream_cols + 255) / 256), 256, 0, sx>>>(F, sa);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
    if (h->timing) SS_CUDA_CHECK(h, cudaEventRecord(h->tev_e[SS_NUM_CLASSES], sx));
    SS_CUDA_CHECK(h, cudaEventRecord(h->ev_join[SS_NUM_CLASSES], sx));
  }
  // clusters: segments of 2..8 tiles
  for (int pass = 0; pass < 2 * (ss_handle_s::MAX_CLUSTER - 1); ++pass) {
    // heavy clusters first (their CTAs each need a whole SM), large ones first
    const bool light = pass >= ss_handle_s::MAX_CLUSTER - 1;
    const int nt = ss_handle_s::MAX_CLUSTER - pass % (ss_handle_s::MAX_CLUSTER - 1);
    const int n_clu = light ? (lnl ? 0 : h->clu_count_light[nt]) : h->clu_count[nt] + (lnl ? h->clu_count_light[nt] : 0);
    if (n_clu == 0) continue;
    const int si = SS_NUM_CLASSES + 1 + (nt - 2) + (light ? ss_handle_s::MAX_CLUSTER - 1 : 0);
    cudaStream_t sx = h->aux[si];
    h->aux_used[si] = true;
    SS_CUDA_CHECK(h, cudaStreamWaitEvent(sx, h->ev_fork, 0));
    if (h->timing) SS_CUDA_CHECK(h, cudaEventRecord(h->tev_s[si], sx));
    ClusterArgs ca;
    ca.segs = light ? h->clu_segs_light[nt] : h->clu_segs[nt];
    ca.xref = h->clu_xref;
    ca.seg_xref_off = h->seg_xref_off;
    ca.seg_xmax = h->seg_xmax;
    a.tile_list = nullptr;
    a.n_list = 0;
    void (*kfn)(EmArgs, ClusterArgs) = nullptr;
    size_t smem = 0;
    int threads = CLUSTER_THREADS;
    if (!lnl) {
      // one cluster barrier per iteration.  512 threads x 128 registers (4 row / 6 column fragment slots per lane).
      // Measured and dropped: 1024 threads x 64 registers with half the slots (32 warps per SM) -- 53.2 ms against
      // 52.3 ms on config 5 x0.1; 256-thread CTAs on 2048-alignment tiles -- 55.8 ms (twice the cluster members)
      if (light) kfn = em_clane_kernel<CLANE_LIGHT_THREADS, CLANE_LIGHT_FR, CLANE_LIGHT_FC, SS_CLANE_LIGHT_MAXREG>;
      else kfn = em_clane_kernel<CLANE_THREADS, CLANE_FR, CLANE_FC, SS_CLANE_MAXREG>;
      threads = light ? CLANE_LIGHT_THREADS : CLANE_THREADS;
      smem = (light ? h->clu_smem_light[nt] : h->clu_smem_lane[nt]) + 256;
    } else {
      switch (nt) {
#define SS_CLU_CASE(N)                               \
  case N:                                            \
    kfn = em_cluster_kernel<CLUSTER_THREADS, true, N>; \
    break;
        SS_CLU_CASE(2) SS_CLU_CASE(3) SS_CLU_CASE(4) SS_CLU_CASE(5) SS_CLU_CASE(6) SS_CLU_CASE(7) SS_CLU_CASE(8)
#undef SS_CLU_CASE
      }
      smem = h->clu_smem[nt] + 256;
    }
    SS_CUDA_CHECK(h, set_smem((const void*)kfn, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_clu * nt));
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = sx;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)nt;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SS_CUDA_CHECK(h, cudaLaunchKernelEx(&cfg, kfn, a, ca));
    h->launches++;
    if (h->timing) SS_CUDA_CHECK(h, cudaEventRecord(h->tev_e[si], sx));
    SS_CUDA_CHECK(h, cudaEventRecord(h->ev_join[si], sx));
  }
  for (int c = SS_NUM_CLASSES - 1; c >= 0; --c) {
    if (h->cls_count[c] == 0) continue;
    cudaStream_t sx = h->aux[c];
    h->aux_used[c] = true;
    SS_CUDA_CHECK(h, cudaStreamWaitEvent(sx, h->ev_fork, 0));
    if (h->timing) SS_CUDA_CHECK(h, cudaEventRecord(h->tev_s[c], sx));
    a.tile_list = h->cls_tiles[c];
    a.n_list = h->cls_count[c];
    const size_t smem = h->cls_smem[c];
    cudaError_t e = cudaSuccess;
    // one instantiation per CTA size: a compile-time thread count is worth ~7% (index arithmetic, reduction tree)
#define SS_LANE_CASE(T)                                                                              \
  case T:                                                                                            \
    e = launch_lane<T, SS_FR, SS_FC, SS_MAXREG>(h, sx, a, h->cls_count[c], T, smem, lnl);             \
    break;
    if (c < SS_NUM_CLASSES - 1) {
      switch (SS_CLS_THREADS[c]) {
        SS_LANE_CASE(32) SS_LANE_CASE(64) SS_LANE_CASE(96) SS_LANE_CASE(128) SS_LANE_CASE(160) SS_LANE_CASE(192)
        SS_LANE_CASE(224) SS_LANE_CASE(256) SS_LANE_CASE(320) SS_LANE_CASE(384) SS_LANE_CASE(448) SS_LANE_CASE(512)
        default: e = cudaErrorInvalidValue; break;
      }
    } else {
      e = launch_lane<SS_BIG_THREADS_V, SS_FR_BIG, SS_FC_BIG, SS_MAXREG_BIG>(h, sx, a, h->cls_count[c],
                                                                              SS_BIG_THREADS_V, smem, lnl);
    }
#undef SS_LANE_CASE
    SS_CUDA_CHECK(h, e);
    if (h->timing) SS_CUDA_CHECK(h, cudaEventRecord(h->tev_e[c], sx));
    SS_CUDA_CHECK(h, cudaEventRecord(h->ev_join[c], sx));
  }
  for (int i = 0; i < SS_NUM_AUX_STREAMS; ++i)
    if (h->aux_used[i]) SS_CUDA_CHECK(h, cudaStreamWaitEvent(st, h->ev_join[i], 0));
  SS_MARK(2);
  SS_MARK(3);
  // final lnL of the streamed tiles
  if (h->n_stream_tiles > 0) {
    a.tile_list = h->stream_tiles;
    a.n_list = h->n_stream_tiles;
    const size_t smem = h->stream_smem;
    SS_CUDA_CHECK(h, cudaFuncSetAttribute(emit_tiles_kernel<STREAM_THREADS>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    emit_tiles_kernel<STREAM_THREADS><<<h->n_stream_tiles, STREAM_THREADS, smem, st>>>(a, 1);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
  }
  SS_MARK(4);
  if (L.S > 0) {
    const int warps_per_block = 8;
    if (h->n_col_chunks > 0 && h->n_stream_tiles + h->n_cluster_tiles > 0) {
      fit_finalize_cols_kernel<<<(h->n_col_chunks + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(
          L, F, h->sc_inamb, chunks);
      h->launches++;
    }
    fit_finalize_kernel<<<(L.S + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(L, F, h->P, chunks);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
  }
  SS_MARK(5);
  h->ev_valid = h->timing;
  h->has_fit = true;
  return SS_OK;
}

// z in the CSR order of the input matrix
int em_emit_z(ss_handle_s* h, cudaStream_t st, double* z) {
  if (!h->has_fit) {
    h->set_error("ss_em_emit_z: no fit");
    return SS_ERR_STATE;
  }
  Layout& L = h->L;
  SS_CUDA_CHECK(h, cudaMemsetAsync(z, 0, (size_t)L.A * sizeof(double), st));
  if (L.n_tiles > 0) {
    EmArgs a;
    a.L = L;
    a.F = h->F;
    a.P = h->P;
    a.z_out = z;
    a.tile_list = h->all_tiles;
    a.n_list = L.n_tiles;
    const size_t smem = h->all_smem;
    SS_CUDA_CHECK(h, cudaFuncSetAttribute(emit_tiles_kernel<STREAM_THREADS>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    emit_tiles_kernel<STREAM_THREADS><<<L.n_tiles, STREAM_THREADS, smem, st>>>(a, 0);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
  }
  if (L.n_uniq > 0) {
    emit_unique_kernel<<<(L.n_uniq + 255) / 256, 256, 0, st>>>(L, z);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
  }
  return SS_OK;
}

}  // namespace ss
