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

// shared memory of the cluster lane kernel (ss_em.cu): x ping-pong, pi, theta: 4 pcmax doubles;
// scratch: pemax + 8 doubles; cross reference: nt * pcmax u16
__host__ __device__ inline size_t clane_smem_bytes(int pemax, int pcmax, int nt) {
  return 8 * (4 * (size_t)pcmax + (size_t)pemax + 8) + 2 * (size_t)nt * (size_t)pcmax + 64;
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
