// Tile-level E+M primitives shared by the resident and the streaming kernels.
//
// One EM iteration of a tile (reference: TelescopeLikelihood.estep model.py:207-210,
// mstep :236-246), with x_j = pi_j * theta_j taken from `pt`:
//   phase A (row pass, one thread per JDS row):
//       r_i  = sum_j Q_ij x_j                      (row sum of the un-normalised E-step)
//       s_i  = w_i / r_i                           (w_i = max_j Q_ij, model.py:183)
//       val[cpos(i,j)] = Q_ij s_i                  (= w_i z_ij / x_j, scattered to column-major order)
//   phase B (column pass, one thread -- or one warp for long columns -- per tile column):
//       T_j  = x_j * sum_i val[i,j]                (= sum_i w_i z_ij = thetasum_j, model.py:239)
// Every sum runs in a fixed order that depends only on the sparsity pattern, so two
// loci that always co-occur with equal scores get bit-identical T.
//
// Per alignment the tile holds Q (f64) and one packed index word (tile-local column | slot in
// the column-major scratch << 16).
#pragma once
#include "ss_common.cuh"

namespace ss {

// kernel argument block shared by the EM kernels
struct EmArgs {
  Layout L;
  FitState F;
  FitParams P;
  const int* tile_list;
  int n_list;
  double* z_out;
};

struct TileView {
  int seg, nrows, nent, ncol, maxlen, nlong;
  int pe, pr, pc, pj;
  long long eo, ro, co, jo;
};

__device__ __forceinline__ TileView load_tile_view(const long long* __restrict__ tile_hdr, int t) {
  const long long* h = tile_hdr + (long long)HDR_WORDS * t;
  TileView v;
  v.seg = (int)h[H_SEG];
  v.nrows = (int)h[H_NROWS];
  v.nent = (int)h[H_NENT];
  v.ncol = (int)h[H_NCOL];
  v.maxlen = (int)h[H_MAXLEN];
  v.nlong = (int)h[H_NLONG];
  v.eo = h[H_ENT_OFF];
  v.ro = h[H_ROW_OFF];
  v.co = h[H_COL_OFF];
  v.jo = h[H_JDS_OFF];
  v.pe = pad8(v.nent);
  v.pr = pad8(v.nrows);
  v.pc = pad8(v.ncol + 1);
  v.pj = pad8(v.maxlen + 1);
  return v;
}

// shared-memory image of a tile
struct TileSmem {
  double* q;       // pe  alignment weights Q_ij, JDS order
  uint32_t* idx;   // pe  tile-local column | column-major slot << 16
  double* val;     // pe  scratch, column-major order
  double* w;       // pr  row weights
  double* pt;      // pc  x_j of the current E-step
  double* aux;     // pc  resident: pi_j ; streaming: unused
  double* pt2;     // pc  x_j after the M-step (ping-pong partner of pt)
  uint16_t* rlen;  // pr
  uint16_t* cptr;  // pc  column-major start of each tile column (ncol+1 used)
  uint16_t* jptr;  // pj  JDS start of each in-row position (maxlen+1 used)
};

__host__ __device__ inline size_t tile_smem_bytes(int nent, int nrows, int ncol, int maxlen) {
  const size_t pe = pad8(nent), pr = pad8(nrows), pc = pad8(ncol + 1), pj = pad8(maxlen + 1);
  return 8 * (2 * pe + pr + 3 * pc) + 4 * pe + 2 * (pr + pc + pj);
}

__device__ __forceinline__ TileSmem carve_tile(unsigned char* base, const TileView& v) {
  TileSmem s;
  double* d = reinterpret_cast<double*>(base);
  s.q = d;            d += v.pe;
  s.val = d;          d += v.pe;
  s.w = d;            d += v.pr;
  s.pt = d;           d += v.pc;
  s.aux = d;          d += v.pc;
  s.pt2 = d;          d += v.pc;
  s.idx = reinterpret_cast<uint32_t*>(d);
  uint16_t* u = reinterpret_cast<uint16_t*>(s.idx + v.pe);
  s.rlen = u;         u += v.pr;
  s.cptr = u;         u += v.pc;
  s.jptr = u;
  return s;
}

// issue the TMA bulk copies of a tile's read-only arrays (one thread)
__device__ __forceinline__ void tile_issue_loads(const Layout& L, const TileView& v, const TileSmem& s,
                                                 uint64_t* bar) {
  const uint32_t bytes = 12u * v.pe + 8u * v.pr + 2u * (v.pr + v.pc + v.pj);
  mbar_expect_tx(bar, bytes);
  bulk_g2s(s.q, L.e_q + v.eo, 8u * v.pe, bar);
  bulk_g2s(s.idx, L.e_idx + v.eo, 4u * v.pe, bar);
  bulk_g2s(s.w, L.r_w + v.ro, 8u * v.pr, bar);
  bulk_g2s(s.rlen, L.r_len + v.ro, 2u * v.pr, bar);
  bulk_g2s(s.cptr, L.c_ptr + v.co, 2u * v.pc, bar);
  bulk_g2s(s.jptr, L.j_ptr + v.jo, 2u * v.pj, bar);
}

// phase A of a streamed tile: one thread per JDS row
template <int THREADS>
__device__ __forceinline__ void tile_phase_a(const TileView& v, const TileSmem& s, const double* __restrict__ pt) {
  for (int r = threadIdx.x; r < v.nrows; r += THREADS) {
    const int len = s.rlen[r];
    double acc = 0.0;
    for (int p = 0; p < len; ++p) {
      const int e = s.jptr[p] + r;
      acc = fma(s.q[e], pt[s.idx[e] & 0xffffu], acc);
    }
    const double sc = acc > 0.0 ? s.w[r] * (1.0 / acc) : 0.0;
    for (int p = 0; p < len; ++p) {
      const int e = s.jptr[p] + r;
      s.val[s.idx[e] >> 16] = s.q[e] * sc;
    }
  }
}

// shared memory of the lane kernel: scratch (+ one padding line), x ping-pong, pi
__host__ __device__ inline size_t lane_smem_bytes(int nent, int ncol) {
  return 8 * ((size_t)pad8(nent) + 8 + 3 * (size_t)pad8(ncol + 1));
}

// shared memory of the cluster lane kernel (ss_em.cu): x ping-pong, pi, theta, partial sums: 5 pcmax
// doubles; scratch: pemax + 8 doubles; cross reference: nt * pcmax u16; shared-column flags: pcmax bytes
__host__ __device__ inline size_t clane_smem_bytes(int pemax, int pcmax, int nt) {
  return 8 * (5 * (size_t)pcmax + (size_t)pemax + 8) + 2 * (size_t)nt * (size_t)pcmax + (size_t)pcmax + 64;
}

// phase B helper: sum of the column-major slots [b, b+n) of a tile column, by one thread
__device__ __forceinline__ double col_sum_thread(const double* __restrict__ val, int b, int n) {
  const double* p = val + b;
  double a;
  switch (n) {
    case 1: a = p[0]; break;
    case 2: a = p[0] + p[1]; break;
    case 3: a = (p[0] + p[1]) + p[2]; break;
    case 4: a = ((p[0] + p[1]) + p[2]) + p[3]; break;
    case 5: a = (((p[0] + p[1]) + p[2]) + p[3]) + p[4]; break;
    case 6: a = ((((p[0] + p[1]) + p[2]) + p[3]) + p[4]) + p[5]; break;
    case 7: a = (((((p[0] + p[1]) + p[2]) + p[3]) + p[4]) + p[5]) + p[6]; break;
    case 8: a = ((((((p[0] + p[1]) + p[2]) + p[3]) + p[4]) + p[5]) + p[6]) + p[7]; break;
    default:
      a = 0.0;
      for (int k = 0; k < n; ++k) a += p[k];
      break;
  }
  return a;
}
__device__ __forceinline__ double col_sum_warp(const double* __restrict__ val, int b, int n, int lane) {
  double a = 0.0;
  for (int k = lane; k < n; k += 32) a += val[b + k];
  return warp_sum(a);
}

// lnL contribution of the tile's (ambiguous) rows and optional z output
// (calculate_lnl model.py:291-317 with z from `ptold`, parameters `ptnew`).
template <int THREADS>
__device__ __forceinline__ double tile_lnl_pass(const TileView& v, const TileSmem& s,
                                                const double* __restrict__ ptold,
                                                const double* __restrict__ ptnew, bool want_lnl,
                                                double* __restrict__ z_out, const int* __restrict__ opos) {
  double l = 0.0;
  for (int r = threadIdx.x; r < v.nrows; r += THREADS) {
    const int len = s.rlen[r];
    double acc = 0.0;
    for (int p = 0; p < len; ++p) {
      const int e = s.jptr[p] + r;
      acc = fma(s.q[e], ptold[s.idx[e] & 0xffffu], acc);
    }
    const double inv = acc > 0.0 ? 1.0 / acc : 0.0;
    for (int p = 0; p < len; ++p) {
      const int e = s.jptr[p] + r;
      const int c = s.idx[e] & 0xffffu;
      const double q = s.q[e];
      const double z = (q * ptold[c]) * inv;
      if (want_lnl) {
        const double pn = ptnew[c];
        if (z > 0.0 && pn > 0.0) l = fma(z, log(q * pn), l);
      }
      if (z_out) z_out[opos[e]] = z;
    }
  }
  return l;
}

// deterministic block sum: every thread returns the same value.  `red` holds 2 x 32 doubles;
// callers alternate `par` between consecutive uses so one barrier per call is enough (the
// barrier also publishes the caller's preceding shared-memory writes).
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* red, int par) {
  constexpr int NW = THREADS / 32;
  v = warp_sum(v);
  if (THREADS == 32) {
    __syncwarp();
    return v;
  }
  double* r = red + 32 * (par & 1);
  const int lane = threadIdx.x & 31;
  if (lane == 0) r[threadIdx.x >> 5] = v;
  __syncthreads();
  // every warp reduces the NW partials the same way (one lane per partial, xor tree, broadcast):
  // cheaper on the shared-memory data path than every lane reading all partials
  double t = lane < NW ? r[lane] : 0.0;
#pragma unroll
  for (int o = NW / 2; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return __shfl_sync(0xffffffffu, t, 0);
}

template <int THREADS>
__device__ __forceinline__ void group_sync() {
  if (THREADS == 32) __syncwarp(); else __syncthreads();
}

}  // namespace ss

namespace ss {

// ---------------------------------------------------------------------------------------------------
// Streaming tile pass with lane fragments (segments too large to stay on chip: celltype / pseudobulk
// models and cells beyond the cluster size).  One E+M half-iteration over the tiles assigned to this
// CTA: r_i = sum_j Q_ij x_j per row, partial thetasum per tile column -> fp64 RED into T[colmap[c]].
//   * the tile arrays (Q, index words, row weights, lengths, column pointers, JDS pointers, column map)
//     arrive by TMA bulk copies into one of TWO shared-memory buffers: the next tile is in flight while
//     the current one is computed (HBM is the bound of this pass: 12 B per alignment + ~6 B of row /
//     column metadata);
//   * the compute is the lane kernel's: rows and columns are cut into fragments of <= 4 alignments,
//     long rows / columns span aligned lane groups combined by xor shuffles, so that a tile costs the
//     same few hundred cycles in every warp instead of "the longest row of warp 0".
// ---------------------------------------------------------------------------------------------------
struct StreamGeom {
  int pe, pr, pc, pj;   // padded maxima over the streamed tiles
  int nbuf;             // 2: double buffered, 1: the largest tile leaves no room for a second buffer
};
constexpr int STREAM_HDR_BYTES = 256;   // the tile header (HDR_WORDS * 8 = 192 B) travels with the tile
constexpr int STREAM_DC_INTS = 32;      // class boundaries derived once per tile (lane_tile_compute)
__host__ __device__ inline size_t stream_buf_bytes(const StreamGeom& g) {
  return STREAM_HDR_BYTES + 12 * (size_t)g.pe + 10 * (size_t)g.pr + 6 * (size_t)g.pc + 2 * (size_t)g.pj;
}
__host__ __device__ inline size_t stream_smem_bytes(const StreamGeom& g) {
  return (size_t)g.nbuf * stream_buf_bytes(g) + 8 * ((size_t)g.pe + 8) + 8 * (size_t)g.pc + 4 * STREAM_DC_INTS + 128;
}

// double buffered when two buffers + scratch fit `limit` bytes of shared memory
inline StreamGeom make_stream_geom(const int* gm, size_t limit) {
  StreamGeom g;
  g.pe = gm[0]; g.pr = gm[1]; g.pc = gm[2]; g.pj = gm[3];
  g.nbuf = 2;
  if (stream_smem_bytes(g) > limit) g.nbuf = 1;
  return g;
}

struct StreamBuf {
  const long long* hdr;
  const double* q; const double* w; const uint32_t* idx; const int* gcol;
  const uint16_t* rlen; const uint16_t* cptr; const uint16_t* jptr;
};
__device__ __forceinline__ StreamBuf carve_stream(unsigned char* base, const StreamGeom& g) {
  StreamBuf b;
  unsigned char* p = base;
  b.hdr = reinterpret_cast<const long long*>(p);  p += STREAM_HDR_BYTES;
  b.q = reinterpret_cast<const double*>(p);       p += 8 * (size_t)g.pe;
  b.w = reinterpret_cast<const double*>(p);       p += 8 * (size_t)g.pr;
  b.idx = reinterpret_cast<const uint32_t*>(p);   p += 4 * (size_t)g.pe;
  b.gcol = reinterpret_cast<const int*>(p);       p += 4 * (size_t)g.pc;
  b.rlen = reinterpret_cast<const uint16_t*>(p);  p += 2 * (size_t)g.pr;
  b.cptr = reinterpret_cast<const uint16_t*>(p);  p += 2 * (size_t)g.pc;
  b.jptr = reinterpret_cast<const uint16_t*>(p);
  return b;
}
__device__ __forceinline__ void stream_issue_loads(const Layout& L, const int* colmap, int t, const TileView& v,
                                                   const StreamBuf& b, uint64_t* bar) {
  const uint32_t bytes = 8u * HDR_WORDS + 12u * v.pe + 10u * v.pr + 6u * v.pc + 2u * v.pj;
  mbar_expect_tx(bar, bytes);
  bulk_g2s((void*)b.hdr, L.tile_hdr + (long long)HDR_WORDS * t, 8u * HDR_WORDS, bar);
  bulk_g2s((void*)b.q, L.e_q + v.eo, 8u * v.pe, bar);
  bulk_g2s((void*)b.idx, L.e_idx + v.eo, 4u * v.pe, bar);
  bulk_g2s((void*)b.w, L.r_w + v.ro, 8u * v.pr, bar);
  bulk_g2s((void*)b.rlen, L.r_len + v.ro, 2u * v.pr, bar);
  bulk_g2s((void*)b.cptr, L.c_ptr + v.co, 2u * v.pc, bar);
  bulk_g2s((void*)b.jptr, L.j_ptr + v.jo, 2u * v.pj, bar);
  bulk_g2s((void*)b.gcol, colmap + v.co, 4u * v.pc, bar);
}

__device__ __forceinline__ double group_xor_sum_s(double a, int g, int gmax) {
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

#ifdef SS_STREAM_PROFILE
__device__ long long g_st_prof[8];
#define ST_TICK(i) { const long long c1_ = clock64(); if (blockIdx.x == 0 && threadIdx.x == 0) g_st_prof[i] += c1_ - c0_; c0_ = c1_; }
#define ST_T0 long long c0_ = clock64();
#else
#define ST_TICK(i)
#define ST_T0
#endif
// compute of one tile whose arrays are in `b`; xs = tile-local x (ncol + 1 slots), val = scratch (pe + 8)
template <int THREADS, int FR, int FC>
__device__ __forceinline__ void lane_tile_compute(const long long* __restrict__ hd, const TileView& v, const StreamBuf& b,
                                                  const double* __restrict__ x, double* __restrict__ T, double* xs,
                                                  double* val, int* dc) {
  constexpr int NW = THREADS / 32;
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ST_T0
  // ---- tile-local x ----
  for (int c = tid; c < v.ncol; c += THREADS) xs[c] = __ldcg(x + b.gcol[c]);
  if (tid == 0) xs[v.ncol] = 0.0;
  // ---- lane-group boundaries of the tile: derived once (one thread), read by everybody after the barrier ----
  if (tid == THREADS - 1) {
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
    dc[0] = n_huge; dc[1] = n8; dc[2] = n4; dc[3] = n2; dc[4] = V8; dc[5] = V4; dc[6] = V2; dc[7] = Vc; dc[8] = r_ovf;
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
    dc[9] = b128; dc[10] = b64; dc[11] = b32; dc[12] = b16; dc[13] = b8; dc[14] = b4;
    dc[15] = W32; dc[16] = W16; dc[17] = W8; dc[18] = W4; dc[19] = W2; dc[20] = Wc; dc[21] = c_ovf;
  }
  __syncthreads();
  ST_TICK(1)
  // ---- phase A: row fragments ----
  {
    const int n_huge = dc[0], n8 = dc[1], n4 = dc[2], n2 = dc[3];
    const int V8 = dc[4], V4 = dc[5], V2 = dc[6], Vc = dc[7], r_ovf = dc[8];
#pragma unroll
    for (int j = 0; j < FR; ++j) {
      const int vl = tid + j * THREADS;
      if (j * THREADS >= Vc) break;                                  // block-uniform
      int n = 0, g = 1, r = 0, k = 0;
      if (vl < Vc) {
        if (vl < V8) { g = 8; r = n_huge + vl / 8; k = vl % 8; }
        else if (vl < V4) { g = 4; r = n_huge + n8 + (vl - V8) / 4; k = (vl - V8) % 4; }
        else if (vl < V2) { g = 2; r = n_huge + n8 + n4 + (vl - V4) / 2; k = (vl - V4) % 2; }
        else { g = 1; r = n_huge + n8 + n4 + n2 + (vl - V2); k = 0; }
        n = max(0, min(FRAG, (int)b.rlen[r] - FRAG * k));
      }
      const int nmax = __reduce_max_sync(full, n);
      if (nmax == 0) continue;                                        // warp-uniform
      const int gmax = __reduce_max_sync(full, g);
      double fq[FRAG];
      uint32_t fi[FRAG];
      double acc = 0.0;
#pragma unroll
      for (int p = 0; p < FRAG; ++p) {
        fq[p] = 0.0;
        fi[p] = 0u;
        if (p < n) {
          const int e = (int)b.jptr[FRAG * k + p] + r;
          fi[p] = b.idx[e];
          fq[p] = b.q[e];
          acc = fma(fq[p], xs[fi[p] & 0xffffu], acc);
        }
      }
      if (gmax > 1) acc = group_xor_sum_s(acc, g, gmax);
      const bool pos = acc > 0.0;
      double inv = 1.0 / (pos ? acc : 1.0);
      inv = pos ? inv : 0.0;
      const double sc = (n > 0 ? b.w[r] : 0.0) * inv;
#pragma unroll
      for (int p = 0; p < FRAG; ++p)
        if (p < n) val[fi[p] >> 16] = fq[p] * sc;
    }
    if (n_huge > 0 || r_ovf < v.nrows) {
      for (int r = tid; r < v.nrows; r += THREADS) {                 // rows outside the lanes (rare)
        const int rr = r < n_huge ? r : r_ovf + (r - n_huge);
        if (rr >= v.nrows) break;
        const int len = b.rlen[rr];
        double acc = 0.0;
        for (int p = 0; p < len; ++p) {
          const int e = (int)b.jptr[p] + rr;
          acc = fma(b.q[e], xs[b.idx[e] & 0xffffu], acc);
        }
        const double sc = acc > 0.0 ? b.w[rr] * (1.0 / acc) : 0.0;
        for (int p = 0; p < len; ++p) {
          const int e = (int)b.jptr[p] + rr;
          val[b.idx[e] >> 16] = b.q[e] * sc;
        }
      }
    }
  }
  __syncthreads();
  ST_TICK(2)
  // ---- phase B: column fragments -> RED into T ----
  {
    const int b128 = dc[9], b64 = dc[10], b32 = dc[11], b16 = dc[12], b8 = dc[13], b4 = dc[14];
    const int W32 = dc[15], W16 = dc[16], W8 = dc[17], W4 = dc[18], W2 = dc[19], Wc = dc[20], c_ovf = dc[21];
#pragma unroll
    for (int j = 0; j < FC; ++j) {
      const int vl = tid + j * THREADS;
      if (j * THREADS >= Wc) break;                                  // block-uniform
      int n = 0, g = 1, c = 0, k = 0, base = 0;
      if (vl < Wc) {
        if (vl < W32) { g = 32; c = b128 + vl / 32; k = vl % 32; }
        else if (vl < W16) { g = 16; c = b64 + (vl - W32) / 16; k = (vl - W32) % 16; }
        else if (vl < W8) { g = 8; c = b32 + (vl - W16) / 8; k = (vl - W16) % 8; }
        else if (vl < W4) { g = 4; c = b16 + (vl - W8) / 4; k = (vl - W8) % 4; }
        else if (vl < W2) { g = 2; c = b8 + (vl - W4) / 2; k = (vl - W4) % 2; }
        else { g = 1; c = b4 + (vl - W2); k = 0; }
        const int cb = b.cptr[c];
        base = cb + FRAG * k;
        n = max(0, min(FRAG, (int)b.cptr[c + 1] - base));
      }
      const int gmax = __reduce_max_sync(full, g);
      double asum = 0.0;
#pragma unroll
      for (int i = 0; i < FRAG; ++i)
        if (i < n) asum += val[base + i];
      if (gmax > 1) asum = group_xor_sum_s(asum, g, gmax);
      if (vl < Wc && k == 0) atomicAdd(&T[b.gcol[c]], xs[c] * asum);
    }
    for (int c = warp; c < b128; c += NW) {                          // very long columns: one warp each
      const int cb = b.cptr[c], n = (int)b.cptr[c + 1] - cb;
      const double asum = col_sum_warp(val, cb, n, lane);
      if (lane == 0) atomicAdd(&T[b.gcol[c]], xs[c] * asum);
    }
    for (int c = c_ovf + tid; c < v.ncol; c += THREADS) {            // columns beyond the lanes
      const int cb = b.cptr[c], n = (int)b.cptr[c + 1] - cb;
      atomicAdd(&T[b.gcol[c]], xs[c] * col_sum_thread(val, cb, n));
    }
  }
  ST_TICK(3)
}

// the tiles tiles[i] (or i itself when tiles == nullptr), i = blockIdx.x, blockIdx.x + gridDim.x, ...;
// tiles of segments with seg_done[seg] != 0 are skipped.  `uses[b]` counts the fills of buffer b (mbarrier parity).
template <int THREADS, int FR, int FC>
__device__ __forceinline__ void lane_stream_pass(const Layout& L, const int* __restrict__ colmap, const int* __restrict__ tiles,
                                                 int n_items, const int* seg_done, const double* __restrict__ x,
                                                 double* __restrict__ T, const StreamGeom& g, unsigned char* dsm,
                                                 uint64_t* bars, uint32_t* uses) {
  const int tid = threadIdx.x;
  const size_t bb = stream_buf_bytes(g);
  double* val = reinterpret_cast<double*>(dsm + (size_t)g.nbuf * bb);
  double* xs = val + g.pe + 8;
  int* dc = reinterpret_cast<int*>(xs + g.pc);
  auto next_item = [&](int i) {
    while (i < n_items) {
      const int t = tiles ? tiles[i] : i;
      if (!seg_done || ((volatile const int*)seg_done)[(int)L.tile_hdr[(long long)HDR_WORDS * t + H_SEG]] == 0) break;
      i += gridDim.x;
    }
    return i;
  };
  int cur = next_item(blockIdx.x);
  int nb = 0;
  if (cur < n_items && tid == 0) {
    const int t = tiles ? tiles[cur] : cur;
    fence_proxy_async();
    stream_issue_loads(L, colmap, t, load_tile_view(L.tile_hdr, t), carve_stream(dsm, g), &bars[0]);
  }
  while (cur < n_items) {
    const int b_cur = g.nbuf == 2 ? (nb & 1) : 0;
    const int nxt = next_item(cur + gridDim.x);
    if (g.nbuf == 2 && nxt < n_items && tid == 0) {
      const int t = tiles ? tiles[nxt] : nxt;
      fence_proxy_async();
      stream_issue_loads(L, colmap, t, load_tile_view(L.tile_hdr, t), carve_stream(dsm + (size_t)(b_cur ^ 1) * bb, g),
                         &bars[b_cur ^ 1]);
    }
    const StreamBuf b = carve_stream(dsm + (size_t)b_cur * bb, g);
#ifdef SS_STREAM_PROFILE
    long long c0_ = clock64();
#endif
    mbar_wait(&bars[b_cur], uses[b_cur] & 1);
    uses[b_cur]++;
    ST_TICK(0)
    const TileView v = load_tile_view(b.hdr, 0);                  // the header arrived with the tile (shared memory)
    lane_tile_compute<THREADS, FR, FC>(b.hdr, v, b, x, T, xs, val, dc);
    __syncthreads();                                              // buffers, val and xs are free again
#ifdef SS_STREAM_PROFILE
    c0_ = clock64() - 0;   // (B tail + barrier counted with the next wait)
    if (blockIdx.x == 0 && tid == 0) g_st_prof[7] += 1;
#endif
    if (g.nbuf == 1 && nxt < n_items && tid == 0) {
      const int t2 = tiles ? tiles[nxt] : nxt;
      fence_proxy_async();
      stream_issue_loads(L, colmap, t2, load_tile_view(L.tile_hdr, t2), carve_stream(dsm, g), &bars[0]);
    }
    cur = nxt;
    ++nb;
  }
}

}  // namespace ss
