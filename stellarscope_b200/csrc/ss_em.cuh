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

// shared memory of the cluster lane kernel (ss_em.cu): x of three iterations, pi, pisum0: 5 pcmax doubles; scratch:
// pemax + 8 doubles; receive buffers of the two pass parities: 2 xmax doubles (xmax = padded sum over the segment's
// columns of (holders - 1): what a tile can receive per pass); cross reference: nt * pcmax u16; list of the shared
// columns, their first receive slots and their park slots: 3 pcmax u16; shared-column flags: pcmax bytes
__host__ __device__ inline size_t clane_smem_bytes(int pemax, int pcmax, int nt, int xmax) {
  return 8 * (5 * (size_t)pcmax + (size_t)pemax + 8 + 2 * (size_t)xmax) + 2 * (size_t)(nt + 3) * (size_t)pcmax +
         (size_t)pcmax + 64;
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
//
// The CTA is WARP-SPECIALISED (round 2: ncu showed the pass waiting, not computing -- every tile paid an L2
// round trip for its x values, two dependent global loads for "is this segment still active", and a
// one-thread derivation of the lane boundaries, all in front of a CTA barrier):
//   * the last warp is the PRODUCER: it picks the next active tile of this CTA, issues the TMA bulk copies
//     of the tile arrays (Q, index words, row weights, lengths, column pointers, JDS pointers, column map,
//     tile header) into one of two shared-memory buffers, gathers the tile-local x values from global
//     memory (many loads in flight per lane), derives the lane-group boundaries, and hands the buffer over
//     through an mbarrier -- one tile ahead of the consumers, so none of that latency is on their path;
//   * the other warps are CONSUMERS: wait for a complete buffer, phase A (row fragments), one barrier,
//     phase B (column fragments -> RED), one barrier, release the buffer.
// The compute is the lane kernel's: rows and columns are cut into fragments of <= 4 alignments, long rows
// / columns span aligned lane groups combined by xor shuffles.
// ---------------------------------------------------------------------------------------------------
struct StreamGeom {
  int pe, pr, pc, pj;   // padded maxima over the streamed tiles
  int nbuf;             // 2: double buffered, 1: the largest tile leaves no room for a second buffer
  int nval;             // column-major scratch arrays: one per buffer when two CTAs still share an SM with that
                        // (no end-of-tile barrier among the consumers), else one (and the barrier)
  int nxs;              // tile-local x arrays: one per buffer, or ONE when only that keeps two CTAs on an SM (the
                        // producer then gathers the next tile's x after the consumers have left the current tile)
};
constexpr int STREAM_HDR_BYTES = 256;   // the tile header (HDR_WORDS * 8 = 192 B) travels with the tile
constexpr int STREAM_DC_INTS = 24;      // lane-group boundaries of a tile (derived by the producer)
constexpr int STREAM_MAX_ITEMS = 1024;
constexpr int STREAM_NPROD = 1;         // producer warps of a streaming CTA  // items of a CTA that the per-pass compaction handles (more: checked one by one)
__host__ __device__ inline size_t stream_buf_bytes(const StreamGeom& g) {
  return STREAM_HDR_BYTES + 12 * (size_t)g.pe + 10 * (size_t)g.pr + 6 * (size_t)g.pc + 2 * (size_t)g.pj;
}
// per buffer: tile arrays + tile-local x + the column-major scratch (a scratch per buffer lets a warp start the next
// tile's row phase while slower warps are still in this tile's column phase: ONE consumer barrier per tile)
__host__ __device__ inline size_t stream_smem_bytes(const StreamGeom& g) {
  return (size_t)g.nbuf * stream_buf_bytes(g) + (size_t)g.nxs * 8 * (size_t)g.pc + (size_t)g.nval * 8 * ((size_t)g.pe + 8) + 128;
}

// Shared-memory configuration of a streaming kernel `fn` (CTA of `threads`) for tiles of geometry gm: the richest
// one that still lets TWO CTAs share an SM (asked of the occupancy calculator, not estimated: the kernels' static
// shared memory and the per-CTA reservation count) -- two scratch arrays + two x arrays, then one scratch, then one x
// array as well; if no configuration reaches two CTAs, the richest that fits one.  `floor_smem`: other users of the
// dynamic shared memory in the same kernel (likelihood pass).  Returns the dynamic shared memory to launch with.
inline size_t pick_stream_geom(const int* gm, size_t limit, size_t floor_smem, const void* fn, int threads, StreamGeom* out,
                               int* occ_out) {
  static const int cand[4][3] = {{2, 2, 2}, {2, 1, 2}, {2, 1, 1}, {1, 1, 1}};      // nbuf, nval, nxs
  StreamGeom g;
  g.pe = gm[0]; g.pr = gm[1]; g.pc = gm[2]; g.pj = gm[3];
  StreamGeom best = g;
  size_t best_smem = 0;
  int best_occ = 0;
  for (int pass = 0; pass < 2 && best_occ == 0; ++pass)          // pass 0: two CTAs per SM; pass 1: anything that fits
    for (int c = 0; c < 4; ++c) {
      g.nbuf = cand[c][0]; g.nval = cand[c][1]; g.nxs = cand[c][2];
      const size_t smem = stream_smem_bytes(g) > floor_smem ? stream_smem_bytes(g) : floor_smem;
      if (smem > limit) continue;
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) continue;
      int occ = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem) != cudaSuccess) continue;
      if (occ >= (pass == 0 ? 2 : 1)) { best = g; best_smem = smem; best_occ = occ; break; }
    }
  *out = best;
  *occ_out = best_occ;
  return best_smem;
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
// the same with the column map on a barrier of its own (issued first: it is small and the producer needs it
// before anything else -- the gather of the tile-local x then overlaps the flight of the big arrays)
__device__ __forceinline__ void stream_issue_loads2(const Layout& L, const int* colmap, int t, const TileView& v,
                                                    const StreamBuf& b, uint64_t* bar, uint64_t* bar_map) {
  mbar_expect_tx(bar_map, 4u * v.pc);
  bulk_g2s((void*)b.gcol, colmap + v.co, 4u * v.pc, bar_map);
  const uint32_t bytes = 8u * HDR_WORDS + 12u * v.pe + 10u * v.pr + 2u * v.pc + 2u * v.pj;
  mbar_expect_tx(bar, bytes);
  bulk_g2s((void*)b.hdr, L.tile_hdr + (long long)HDR_WORDS * t, 8u * HDR_WORDS, bar);
  bulk_g2s((void*)b.q, L.e_q + v.eo, 8u * v.pe, bar);
  bulk_g2s((void*)b.idx, L.e_idx + v.eo, 4u * v.pe, bar);
  bulk_g2s((void*)b.w, L.r_w + v.ro, 8u * v.pr, bar);
  bulk_g2s((void*)b.rlen, L.r_len + v.ro, 2u * v.pr, bar);
  bulk_g2s((void*)b.cptr, L.c_ptr + v.co, 2u * v.pc, bar);
  bulk_g2s((void*)b.jptr, L.j_ptr + v.jo, 2u * v.pj, bar);
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

// control block of the pipeline (static shared memory of the kernel that runs the pass)
struct StreamPipe {
  uint64_t full[2];     // the TMA bytes of buffer b have landed
  uint64_t mapf[2];     // ... its column map has (a barrier of its own: the producer gathers x from it first)
  uint64_t ready[2];    // buffer b is complete: tile arrays, tile-local x, boundaries (producer -> consumers)
  uint64_t empty[2];    // the consumers are done with buffer b (consumers -> producer)
  int tile[2];          // tile in buffer b, -1: end of the pass
  int dc[2][STREAM_DC_INTS];
  // this CTA's ACTIVE items of the current pass (segments that have converged are left out), compacted once per
  // pass by all threads: the producer then walks a list instead of chasing tile -> segment -> done flag through
  // global memory for every item it skips (late iterations of a celltype fit skip most of them)
  int n_active;
  int warp_cnt[32];
  uint16_t items[STREAM_MAX_ITEMS];   // ordinal j of the active item blockIdx + j * gridDim
};
__device__ __forceinline__ void stream_pipe_init(StreamPipe* p, int consumer_warps) {   // one thread, before a CTA barrier
  for (int b = 0; b < 2; ++b) {
    mbar_init(&p->full[b], 1);
    mbar_init(&p->mapf[b], 1);
    mbar_init(&p->ready[b], 1);
    mbar_init(&p->empty[b], consumer_warps);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int COUNT>
__device__ __forceinline__ void consumer_bar() {      // named barrier 1: the consumer warps only
  asm volatile("bar.sync 1, %0;" ::"n"(COUNT) : "memory");
}

// lane-group boundaries of a tile from its header, for a pass with FR * CONS row lanes / FC * CONS column lanes
__device__ __forceinline__ void stream_boundaries(const long long* hd, int nrows, int ncol, int row_lanes, int col_lanes,
                                                  int* dc) {
  const int rgt4 = (int)hd[H_RGT4], rgt8 = (int)hd[H_RGT8], rgt16 = (int)hd[H_RGT16], rgt32 = (int)hd[H_RGT32];
  const int n_huge = rgt32;
  const int n8 = rgt16 - rgt32, n4 = rgt8 - rgt16, n2 = rgt4 - rgt8, n1 = nrows - rgt4;
  const int V8 = 8 * n8, V4 = V8 + 4 * n4, V2 = V4 + 2 * n2, V = V2 + n1;
  const int Vc = min(V, row_lanes);
  int r_ovf;
  if (Vc >= V) r_ovf = nrows;
  else if (Vc < V8) r_ovf = n_huge + Vc / 8;
  else if (Vc < V4) r_ovf = n_huge + n8 + (Vc - V8) / 4;
  else if (Vc < V2) r_ovf = n_huge + n8 + n4 + (Vc - V4) / 2;
  else r_ovf = n_huge + n8 + n4 + n2 + (Vc - V2);
  dc[0] = n_huge; dc[1] = n8; dc[2] = n4; dc[3] = n2; dc[4] = V8; dc[5] = V4; dc[6] = V2; dc[7] = Vc; dc[8] = r_ovf;
  const int b128 = (int)hd[H_CGT128];
  const int b64 = (int)hd[H_CGT64], b32 = (int)hd[H_CGT32], b16 = (int)hd[H_CGT16], b8 = (int)hd[H_CGT8],
            b4 = (int)hd[H_CGT4];
  const int W32 = 32 * (b64 - b128), W16 = W32 + 16 * (b32 - b64), W8 = W16 + 8 * (b16 - b32);
  const int W4 = W8 + 4 * (b8 - b16), W2 = W4 + 2 * (b4 - b8), W = W2 + (ncol - b4);
  const int Wc = min(W, col_lanes);
  int c_ovf;
  if (Wc >= W) c_ovf = ncol;
  else if (Wc < W32) c_ovf = b128 + Wc / 32;
  else if (Wc < W16) c_ovf = b64 + (Wc - W32) / 16;
  else if (Wc < W8) c_ovf = b32 + (Wc - W16) / 8;
  else if (Wc < W4) c_ovf = b16 + (Wc - W8) / 4;
  else if (Wc < W2) c_ovf = b8 + (Wc - W4) / 2;
  else c_ovf = b4 + (Wc - W2);
  dc[9] = b128; dc[10] = b64; dc[11] = b32; dc[12] = b16; dc[13] = b8; dc[14] = b4;
  dc[15] = W32; dc[16] = W16; dc[17] = W8; dc[18] = W4; dc[19] = W2; dc[20] = Wc; dc[21] = c_ovf;
}

// compute of one tile whose arrays are in `b` (consumer warps: CONS threads, tid < CONS);
// xs = tile-local x (ncol + 1 slots, filled by the producer), val = scratch (pe + 8)
template <int CONS, int FR, int FC>
__device__ __forceinline__ void lane_tile_compute(const TileView& v, const StreamBuf& b, double* __restrict__ T,
                                                  const double* __restrict__ xs, double* __restrict__ val,
                                                  const int* __restrict__ dc) {
  constexpr int NW = CONS / 32;
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- phase A: row fragments ----
  {
    const int n_huge = dc[0], n8 = dc[1], n4 = dc[2], n2 = dc[3];
    const int V8 = dc[4], V4 = dc[5], V2 = dc[6], Vc = dc[7], r_ovf = dc[8];
#pragma unroll
    for (int j = 0; j < FR; ++j) {
      const int vl = tid + j * CONS;
      if (j * CONS >= Vc) break;                                     // block-uniform
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
      for (int r = tid; r < v.nrows; r += CONS) {                    // rows outside the lanes (rare)
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
  consumer_bar<CONS>();
  // ---- phase B: column fragments -> RED into T ----
  {
    const int b128 = dc[9], b64 = dc[10], b32 = dc[11], b16 = dc[12], b8 = dc[13], b4 = dc[14];
    const int W32 = dc[15], W16 = dc[16], W8 = dc[17], W4 = dc[18], W2 = dc[19], Wc = dc[20], c_ovf = dc[21];
#pragma unroll
    for (int j = 0; j < FC; ++j) {
      const int vl = tid + j * CONS;
      if (j * CONS >= Wc) break;                                     // block-uniform
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
    for (int c = c_ovf + tid; c < v.ncol; c += CONS) {               // columns beyond the lanes
      const int cb = b.cptr[c], n = (int)b.cptr[c + 1] - cb;
      atomicAdd(&T[b.gcol[c]], xs[c] * col_sum_thread(val, cb, n));
    }
  }
}

// The tiles tiles[i] (or i itself when tiles == nullptr), i = blockIdx.x, blockIdx.x + gridDim.x, ...; tiles of
// segments with seg_done[seg] != 0 are skipped.  `n_use` counts the buffer fills of this CTA over the life of the
// kernel (mbarrier parities); every thread carries the same value.  Called by all THREADS threads of the CTA; ends
// with a CTA barrier.
template <int THREADS, int FR, int FC>
__device__ __forceinline__ void lane_stream_pass(const Layout& L, const int* __restrict__ colmap, const int* __restrict__ tiles,
                                                 int n_items, const int* seg_done, const double* __restrict__ x,
                                                 double* __restrict__ T, const StreamGeom& g, unsigned char* dsm,
                                                 StreamPipe* pipe, uint32_t& n_use) {
  // producer warps (the last warps of the CTA).  Two producers, each filling every other use, were measured: the
  // consumers' wait for a complete buffer did not shrink and the lost consumer warp cost 9 % on config 4 (2.87 against
  // 2.63 ms): the wait is for the data, not for the producer's instructions
  constexpr int NPROD = STREAM_NPROD;
  constexpr int CONS = THREADS - 32 * NPROD;             // consumer threads
  const int tid = threadIdx.x, lane = tid & 31;
  const size_t bb = stream_buf_bytes(g);
  const int nbuf = g.nbuf;
  const int nval = g.nbuf == 2 ? g.nval : 1;
  double* val0 = reinterpret_cast<double*>(dsm + (size_t)nbuf * bb);           // nval scratch arrays of pe + 8
  double* xs0 = val0 + (size_t)g.nval * (g.pe + 8);                            // nxs tile-local x arrays of pc
  const int nxs = g.nxs;
  uint32_t n = n_use;
  // ---- active items of this CTA (all threads): items[k] = ordinal j of the k-th active item blockIdx + j * gridDim ----
  const int n_mine = (int)blockIdx.x < n_items ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const bool listed = seg_done != nullptr && n_mine <= STREAM_MAX_ITEMS;
  if (listed) {
    for (int j0 = 0; j0 < n_mine; j0 += THREADS) {                   // one round for up to THREADS items
      const int j = j0 + tid;
      bool act = false;
      if (j < n_mine) {
        const int i = (int)blockIdx.x + j * (int)gridDim.x;
        const int t = tiles ? tiles[i] : i;
        act = ((volatile const int*)seg_done)[(int)L.tile_hdr[(long long)HDR_WORDS * t + H_SEG]] == 0;
      }
      const unsigned m = __ballot_sync(0xffffffffu, act);
      if (lane == 0) pipe->warp_cnt[tid >> 5] = __popc(m);
      if (j0 == 0 && tid == 0) pipe->n_active = 0;
      __syncthreads();
      int before = pipe->n_active;
      for (int w = 0; w < (tid >> 5); ++w) before += pipe->warp_cnt[w];
      if (act) pipe->items[before + __popc(m & ((1u << lane) - 1u))] = (uint16_t)j;
      __syncthreads();
      if (tid == 0) {
        int tot = pipe->n_active;
        for (int w = 0; w < THREADS / 32; ++w) tot += pipe->warp_cnt[w];
        pipe->n_active = tot;
      }
      __syncthreads();
    }
  }
  // number of tiles of this pass when it is known up front (list, or no segment can drop out); -1: the single
  // producer finds out by walking (more items than the list holds)
  const int M = listed ? pipe->n_active : (seg_done == nullptr ? n_mine : -1);
  if (tid >= CONS) {
    // ------------------------------ producer warps ------------------------------
    // producer p fills the uses k = p, p + 2, ... of this pass (with two buffers: always the same buffer), so a
    // producer has two consumer tile times for one tile's copies + x gather; use M is the end marker
    const int p = (tid - CONS) >> 5;
    const int kstep = M >= 0 ? NPROD : 1;
    int i = (int)blockIdx.x;                                          // walking producer only
    auto skip_done = [&](int ii) {
      while (ii < n_items) {
        const int t = tiles ? tiles[ii] : ii;
        if (!seg_done || ((volatile const int*)seg_done)[(int)L.tile_hdr[(long long)HDR_WORDS * t + H_SEG]] == 0) break;
        ii += gridDim.x;
      }
      return ii;
    };
    if (M < 0) i = skip_done(i);
    uint32_t n_end = n;                                               // first use after this pass (walking producer)
    for (int k = p; M >= 0 ? k <= M : p == 0; k += kstep) {
      const uint32_t nk = n + (uint32_t)k;
      const int slot = nbuf == 2 ? (int)(nk & 1u) : 0;
      const uint32_t kk = nbuf == 2 ? (nk >> 1) : nk;
      if (kk >= 1) mbar_wait(&pipe->empty[slot], (kk - 1) & 1u);      // the consumers have released this buffer
      const bool last = M >= 0 ? k == M : i >= n_items;
      if (last) {
        if (lane == 0) {
          pipe->tile[slot] = -1;
          mbar_arrive(&pipe->full[slot]);
          mbar_arrive(&pipe->mapf[slot]);
          mbar_arrive(&pipe->ready[slot]);
        }
        n_end = nk + 1;
        break;
      }
      const int item = M >= 0 ? (int)blockIdx.x + (listed ? (int)pipe->items[k] : k) * (int)gridDim.x : i;
      const int t = tiles ? tiles[item] : item;
      const TileView v = load_tile_view(L.tile_hdr, t);
      const StreamBuf b = carve_stream(dsm + (size_t)slot * bb, g);
      if (lane == 0) {
        fence_proxy_async();
        stream_issue_loads2(L, colmap, t, v, b, &pipe->full[slot], &pipe->mapf[slot]);
      }
      if (M < 0) i = skip_done(i + gridDim.x);                        // overlaps the flight of the copies
      mbar_wait(&pipe->mapf[slot], kk & 1u);                          // the column map is in; the big arrays still fly
      if (nxs == 1 && nbuf == 2 && nk >= 1) {
        // one x array for both buffers: the consumers must have left the PREVIOUS tile (use nk - 1, other buffer)
        const uint32_t np_ = nk - 1;
        mbar_wait(&pipe->empty[(int)(np_ & 1u)], (np_ >> 1) & 1u);
      }
      // tile-local x: 8 loads in flight per lane
      double* xs = xs0 + (size_t)(nxs == 2 ? slot : 0) * g.pc;
      for (int c0 = 0; c0 < v.ncol; c0 += 256) {
        double xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = c0 + 32 * u + lane;
          xv[u] = c < v.ncol ? __ldcg(x + b.gcol[c]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = c0 + 32 * u + lane;
          if (c < v.ncol) xs[c] = xv[u];
        }
      }
      mbar_wait(&pipe->full[slot], kk & 1u);
      if (lane == 0) {
        xs[v.ncol] = 0.0;
        stream_boundaries(b.hdr, v.nrows, v.ncol, FR * CONS, FC * CONS, pipe->dc[slot]);
        pipe->tile[slot] = t;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&pipe->ready[slot]);                // release: x, boundaries, tile id
    }
    // every thread leaves with the same use count: the consumers count theirs, the producers take it from the pass
    if (M >= 0) n = n + (uint32_t)M + 1u;
    else if (p == 0) n = n_end;
    if (M < 0) {                                                      // the idle second producer learns it from the first
      if (p == 0 && lane == 0) pipe->n_active = (int)(n_end - n_use);
    }
  } else {
    // ------------------------------ consumer warps ------------------------------
    while (true) {
      const int slot = nbuf == 2 ? (int)(n & 1u) : 0;
      const uint32_t k = nbuf == 2 ? (n >> 1) : n;
      mbar_wait(&pipe->ready[slot], k & 1u);
      const int t = pipe->tile[slot];
      mbar_wait(&pipe->full[slot], k & 1u);                          // complete already; orders the reads after the copies
      ++n;
      if (t < 0) {
        __syncwarp();                                                // the warp has read the marker
        if (lane == 0) mbar_arrive(&pipe->empty[slot]);
        break;
      }
      const StreamBuf b = carve_stream(dsm + (size_t)slot * bb, g);
      const TileView v = load_tile_view(b.hdr, 0);                   // the header arrived with the tile
      lane_tile_compute<CONS, FR, FC>(v, b, T, xs0 + (size_t)(nxs == 2 ? slot : 0) * g.pc, val0 + (size_t)(nval == 2 ? slot : 0) * (g.pe + 8),
                                      pipe->dc[slot]);
      if (nval != 2) consumer_bar<CONS>();                           // one scratch: everybody has left the column phase
      // this warp is done with the buffer (tile arrays, x, scratch); the producer refills it when all warps are.
      // No CTA-wide barrier here: the row phase of the next tile writes the OTHER scratch, and nobody reaches the
      // tile after that before everybody has passed the next tile's row | column barrier, i.e. left this column phase
      __syncwarp();
      if (lane == 0) mbar_arrive(&pipe->empty[slot]);
    }
  }
  __syncthreads();
  if (M < 0 && tid >= CONS) n = n_use + (uint32_t)pipe->n_active;     // walking mode: uses counted by producer 0
  n_use = n;
  __syncthreads();
}

}  // namespace ss
