// Device builder of the tiled segment layout (ss_layout_build).
//
// Replaces, for all models of a pooling mode at once, TelescopeLikelihood.__init__
// (reference utils/model.py:102-200) and the row selection of _fit_pooling_model
// (:701-743).  The construction is specified by stellarscope_b200/layout_host.py;
// this file produces the same arrays on the device.  Device-wide sorts / scans use
// CUB (library code, one-time build step); the per-tile work (JDS ordering, tile
// column discovery, column-major slots) is a hand-written CTA-per-tile kernel (its three
// block-wide sorts are CUB's BlockRadixSort).
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "ss_em.cuh"
#include "ss_handle.h"

namespace ss {

namespace {

// temporaries come from the stream-ordered pool (cudaMallocAsync): no device-wide
// synchronisation per allocation, memory is reused across builds
struct TmpPool {
  cudaStream_t st;
  std::vector<void*> p;
  explicit TmpPool(cudaStream_t s) : st(s) {}
  ~TmpPool() {
    for (void* q : p) cudaFreeAsync(q, st);
  }
  template <class T>
  cudaError_t get(T** out, size_t n) {
    *out = nullptr;
    if (n == 0) n = 1;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(out), n * sizeof(T), st);
    if (e == cudaSuccess) p.push_back(*out);
    return e;
  }
};

#define CUB_CALL(call_with_tmp)                \
  do {                                         \
    void* d_tmp = nullptr;                     \
    size_t tmp_bytes = 0;                      \
    SS_CUDA_CHECK(h, (call_with_tmp));         \
    unsigned char* _t;                         \
    SS_CUDA_CHECK(h, tmp.get(&_t, tmp_bytes)); \
    d_tmp = _t;                                \
    SS_CUDA_CHECK(h, (call_with_tmp));         \
    h->launches++;                             \
  } while (0)

// ---- step 1: classify rows, per-segment integer statistics --------------------
// counters: [0] = max row length among ambiguous rows, [1] = 1 when the matrix is not canonical
// (a zero score, column indices not strictly ascending inside a row, or outside [0, K))
__global__ void row_class_kernel(int n_rows, int n_cols, const int* __restrict__ row_ptr,
                                 const int* __restrict__ col_idx, const uint16_t* __restrict__ score,
                                 const int* __restrict__ row_seg, int n_seg, uint8_t* __restrict__ is_amb,
                                 uint8_t* __restrict__ is_uni, int* __restrict__ seg_smax, int* __restrict__ seg_nobs,
                                 int* __restrict__ seg_namb, int* __restrict__ seg_nuniq,
                                 unsigned long long* __restrict__ seg_A, unsigned long long* __restrict__ seg_Aamb,
                                 int* __restrict__ counters) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  int seg = -1, len = 0, rmax = 0;
  if (row < n_rows) {
    seg = row_seg[row];
    if (seg >= n_seg) seg = -1;
    const int b = row_ptr[row], e = row_ptr[row + 1];
    len = e - b;
    bool bad = len < 0;
    int prev = -1;
    for (int k = b; k < e; ++k) {
      const int sc = (int)score[k], c = col_idx[k];
      rmax = max(rmax, sc);
      bad |= sc == 0 || c <= prev || c >= n_cols;
      prev = c;
    }
    if (bad) counters[1] = 1;
    is_amb[row] = seg >= 0 && len > 1;
    is_uni[row] = seg >= 0 && len == 1;
  }
  const bool fitted = seg >= 0 && len > 0;
  // warp aggregation when the whole warp works on one segment (rows of a cell are usually adjacent)
  const unsigned full = 0xffffffffu;
  const int s0 = __shfl_sync(full, seg, 0);
  if (__all_sync(full, seg == s0)) {
    if (s0 >= 0) {
      const int wmax = __reduce_max_sync(full, rmax);
      const int nobs = __reduce_add_sync(full, fitted ? 1 : 0);
      const int namb = __reduce_add_sync(full, (fitted && len > 1) ? 1 : 0);
      const int nuni = __reduce_add_sync(full, (fitted && len == 1) ? 1 : 0);
      const int a_all = __reduce_add_sync(full, fitted ? len : 0);
      const int a_amb = __reduce_add_sync(full, (fitted && len > 1) ? len : 0);
      const int lmax = __reduce_max_sync(full, (fitted && len > 1) ? len : 0);
      if ((threadIdx.x & 31) == 0) {
        if (wmax) atomicMax(&seg_smax[s0], wmax);
        if (nobs) atomicAdd(&seg_nobs[s0], nobs);
        if (namb) atomicAdd(&seg_namb[s0], namb);
        if (nuni) atomicAdd(&seg_nuniq[s0], nuni);
        if (a_all) atomicAdd(&seg_A[s0], (unsigned long long)a_all);
        if (a_amb) atomicAdd(&seg_Aamb[s0], (unsigned long long)a_amb);
        if (lmax) atomicMax(&counters[0], lmax);
      }
    }
  } else if (fitted) {
    atomicMax(&seg_smax[seg], rmax);
    atomicAdd(&seg_nobs[seg], 1);
    atomicAdd(&seg_A[seg], (unsigned long long)len);
    if (len > 1) {
      atomicAdd(&seg_namb[seg], 1);
      atomicAdd(&seg_Aamb[seg], (unsigned long long)len);
      atomicMax(&counters[0], len);
    } else {
      atomicAdd(&seg_nuniq[seg], 1);
    }
  }
}

__global__ void iota_kernel(int n, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

__global__ void gather_seg_len_kernel(int n, const int* __restrict__ rows, const int* __restrict__ row_seg,
                                      const int* __restrict__ row_ptr, int* __restrict__ seg_out,
                                      long long* __restrict__ len_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rows[i];
  if (seg_out) seg_out[i] = row_seg[r];
  if (len_out) len_out[i] = row_ptr[r + 1] - row_ptr[r];
}

// sort key of an ambiguous row: (segment, first locus).  Inside a segment the rows are ordered by the
// first locus they align to (the second one when the first is column 0, __no_feature): reads multimap
// inside a repeat family, so consecutive rows -- and therefore the tiles of a multi-tile segment -- touch
// a narrow band of loci.  Tiles then share few columns: fewer partial sums to exchange between the CTAs
// of a cluster, fewer RED operations and a smaller gather of x per streamed tile.
__global__ void row_sortkey_kernel(int n, const int* __restrict__ rows, const int* __restrict__ row_seg,
                                   const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                   unsigned long long* __restrict__ key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rows[i];
  const int b = row_ptr[r];
  int c = col_idx[b];
  if (c == 0) c = col_idx[b + 1];                 // ambiguous rows have at least two alignments
  key[i] = ((unsigned long long)(unsigned)row_seg[r] << 32) | (unsigned)c;
}
__global__ void key_to_seg_kernel(int n, const unsigned long long* __restrict__ key, int* __restrict__ seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) seg[i] = (int)(key[i] >> 32);
}

// rows of a segment are cut into tiles every `chunk` alignments.  Segments that will be streamed (more than
// STREAM_SEG_TILES large tiles) use the small tile size; a segment that fits ONE large tile is one tile; the
// segments in between (2..8 tiles: one thread-block cluster per segment) are cut into EQUAL tiles of about `target`
// alignments -- all CTAs of a cluster wait for the slowest one at the cluster barrier of every iteration, so a
// 5000-alignment cell runs as 2 x 2500 rather than 4080 + 920 (layout_host.py: seg_chunk)
__host__ __device__ inline int seg_chunk(long long a_amb, int chunk_big, int chunk_small, int thr, int target) {
  if (a_amb > (long long)thr * chunk_big) return chunk_small;
  if (a_amb <= chunk_big) return chunk_big;
  long long nt = (a_amb + target - 1) / target;
  if (nt > thr) nt = thr;                      // a_amb <= thr * chunk_big: the tiles still fit
  return (int)((a_amb + nt - 1) / nt);
}

// per segment: first sorted ambiguous row, first entry offset, number of tiles
__global__ void seg_tiles_kernel(int n_seg, int n_amb, const long long* __restrict__ seg_row0 /* exclusive scan of namb */,
                                 const int* __restrict__ seg_namb, const long long* __restrict__ pre /* n_amb+1 */,
                                 int chunk_big, int chunk_small, int small_thr, int target, const long long* __restrict__ seg_Aamb,
                                 long long* __restrict__ seg_ntiles64, long long* __restrict__ seg_ent0) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  const int chunk = seg_chunk(seg_Aamb[s], chunk_big, chunk_small, small_thr, target);
  const long long r0 = seg_row0[s];
  const int n = seg_namb[s];
  const long long e0 = pre[min(r0, (long long)n_amb)];
  seg_ent0[s] = e0;
  if (n == 0) {
    seg_ntiles64[s] = 0;
  } else {
    const long long last_start = pre[r0 + n - 1] - e0;
    seg_ntiles64[s] = last_start / chunk + 1;
  }
}

// tile of every sorted ambiguous row and the first row of every tile
__global__ void row_tile_kernel(int n_amb, const int* __restrict__ aseg, const long long* __restrict__ pre,
                                const long long* __restrict__ seg_ent0, const long long* __restrict__ seg_tile0,
                                int chunk_big, int chunk_small, int small_thr, int target, const long long* __restrict__ seg_Aamb,
                                int* __restrict__ rtile) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_amb) return;
  const int s = aseg[k];
  const int chunk = seg_chunk(seg_Aamb[s], chunk_big, chunk_small, small_thr, target);
  rtile[k] = (int)(seg_tile0[s] + (pre[k] - seg_ent0[s]) / chunk);
}
__global__ void tile_heads_kernel(int n_amb, int n_tiles, const int* __restrict__ rtile, int* __restrict__ tile_row0) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 0) tile_row0[n_tiles] = n_amb;
  if (k >= n_amb) return;
  if (k == 0 || rtile[k] != rtile[k - 1]) tile_row0[rtile[k]] = k;
}

// padded sizes known before the tile kernel
__global__ void tile_sizes_kernel(int n_tiles, const int* __restrict__ tile_row0, const long long* __restrict__ pre,
                                  long long* __restrict__ pad_ent, long long* __restrict__ pad_row) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const int a = tile_row0[t], b = tile_row0[t + 1];
  pad_ent[t] = pad8((int)(pre[b] - pre[a]));
  pad_row[t] = pad8(b - a);
}

// ---- block-wide helpers ----------------------------------------------------------
// Stable sort of keys[0, n) (shared memory, capacity BT * ITEMS) by the bits [begin_bit, end_bit): the block-wide
// radix sort of CUB (library code, like the device-wide sorts of this file) on a blocked arrangement in registers.
// 15 four-bit passes per tile in all; the bitonic network this replaces took 210 compare-exchange steps and was 70 %
// of the builder's instructions (ncu, profiles/r2l_tile_build_*).  Slots >= n take part as all-ones keys and, the
// sort being stable, stay behind the keys proper.
template <int BT, int ITEMS>
struct TileSort {
  using BRS = cub::BlockRadixSort<unsigned long long, BT, ITEMS>;
  using Temp = typename BRS::TempStorage;
  static __device__ __forceinline__ void sort(unsigned long long* keys, int n, int begin_bit, int end_bit, Temp& tmp) {
    unsigned long long k[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int idx = (int)threadIdx.x * ITEMS + i;
      k[i] = idx < n ? keys[idx] : ~0ull;
    }
    __syncthreads();                                   // everything is read before the striped write-back below
    BRS(tmp).SortBlockedToStriped(k, begin_bit, end_bit);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) keys[i * BT + (int)threadIdx.x] = k[i];
    __syncthreads();
  }
};
__host__ __device__ constexpr int ceil_log2(int n) {
  int b = 0;
  while ((1 << b) < n) ++b;
  return b;
}
// exclusive scan of n ints in shared memory, in place; returns the total. scratch: blockDim.x + 1 ints.  Every thread
// sums a contiguous piece; the per-thread sums are scanned by warp shuffles (a thread-0 loop over 1024 partial sums
// cost more than the sorts of a tile)
__device__ int block_excl_scan(int* data, int n, int* scratch) {
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
  const int per = (n + T - 1) / T;
  const int b = min(n, tid * per), e = min(n, b + per);
  int s = 0;
  for (int i = b; i < e; ++i) s += data[i];
  int incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nw ? scratch[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += up;
    }
    if (lane < nw) scratch[lane] = wi - w;                  // exclusive prefix of the warp totals
    if (lane == 31) scratch[T] = wi;                        // total (nw <= 32)
  }
  __syncthreads();
  int run = scratch[warp] + incl - s;
  const int total = scratch[T];
  for (int i = b; i < e; ++i) {
    const int v = data[i];
    data[i] = run;
    run += v;
  }
  __syncthreads();
  return total;
}

constexpr int BUILD_THREADS = 256;

struct TileBuildArgs {
  const int* row_ptr;
  const int* col_idx;
  const uint16_t* score;
  const int* ar;            // sorted ambiguous rows
  const int* aseg;
  const int* tile_row0;
  const long long* pre;
  const long long* ent_off; // exclusive scans of the padded sizes
  const long long* row_off;
  const int* seg_smax;
  int n_tiles;
  // outputs at their final offsets
  double* e_q;
  uint32_t* e_idx;
  int* e_opos;
  uint16_t* r_len;
  double* r_w;
  int* r_orow;
  // outputs at temporary (upper bound) offsets: column arrays at ent_off[t] + 8 t, jds at (maxlen_pad) * t
  uint16_t* tmp_cptr;
  int* tmp_gcol;
  uint16_t* tmp_jptr;
  int jds_stride;
  int gbits;                // bits of a column index (n_cols - 1)
  // per tile scalars
  int *t_ncol, *t_maxlen, *t_nlong, *t_seg, *t_nrows, *t_nent, *t_rfrag, *t_cfrag;
  int* t_gt;               // [n_tiles][10]: rows longer than 4/8/16/32, columns with more than 4..128 rows
  double* t_wsum;
};

// ---- step 5: one CTA per tile ---------------------------------------------------
// Instantiated for three tile-size classes (ECAP = 256 / 1024 / E_HARD alignments) so that the
// many small tiles of individual-mode cells do not pay for the shared memory of the largest
// one; every instance is launched over all tiles and a CTA returns when the tile is not of its
// class (ELOW < nent <= ECAP).
template <int ECAP, int BT>
constexpr size_t tile_build_smem() {
  return (size_t)ECAP * 8 * 2 + (size_t)(ECAP * 4 + 2) * 4 + (size_t)(ECAP / 2 * 2 + ECAP + 2) * 4 +
         (size_t)(BT + 1) * 4 + 64;
}

template <int ECAP, int ELOW, int BT>
__global__ void __launch_bounds__(BT) tile_build_kernel(TileBuildArgs a) {
  extern __shared__ __align__(16) unsigned char dsm[];
  constexpr int RCAP = ECAP / 2;
  const int tid = threadIdx.x;
  const int t = blockIdx.x;
  const int row0 = a.tile_row0[t], nrows = a.tile_row0[t + 1] - row0;
  const int nent = (int)(a.pre[row0 + nrows] - a.pre[row0]);
  if (nent <= ELOW || nent > ECAP) return;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(dsm);   // ECAP
  unsigned long long* ckeys = keys + ECAP;                                  // ECAP
  int* colidx = reinterpret_cast<int*>(ckeys + ECAP);                       // ECAP: column (gcol order) of each sorted entry
  int* colstart = colidx + ECAP;                                            // ECAP + 1
  int* rank_of = colstart + ECAP + 1;                                       // ECAP
  int* cptr = rank_of + ECAP;                                               // ECAP + 1
  int* rrow = cptr + ECAP + 1;                                              // RCAP  original row of JDS row r
  int* rlen = rrow + RCAP;                                                  // RCAP
  int* jptr = rlen + RCAP;                                                  // ECAP + 2 (a row may be as long as the tile)
  int* scratch = jptr + ECAP + 2;                                           // BT + 1
  __shared__ double s_red[32];
  using SortE = TileSort<BT, ECAP / BT>;                                    // entries / columns: up to ECAP keys
  using SortR = TileSort<BT, (RCAP + BT - 1) / BT>;                         // rows: up to RCAP keys; entries / columns of a half-full tile
  __shared__ union { typename SortE::Temp e; typename SortR::Temp r; } s_sort;
  constexpr int HALF = BT * ((RCAP + BT - 1) / BT);                         // keys the small sorter takes (streamed tiles: 2048 of 4096)
  constexpr int LEN_BITS = ceil_log2(E_HARD) + 1;                           // E_HARD - len / - count; all-ones = padding sorts last
  const int seg = a.aseg[row0];
  const long long eo = a.ent_off[t], ro = a.row_off[t];

  // (a) JDS row order: (len desc, rank asc)
  int lmax = 0;
  for (int r = tid; r < nrows; r += BT) {
    const int row = a.ar[row0 + r];
    const int len = a.row_ptr[row + 1] - a.row_ptr[row];
    keys[r] = ((unsigned long long)(unsigned)(E_HARD - len) << 32) | (unsigned)r;
  }
  __syncthreads();
  SortR::sort(keys, nrows, 32, 32 + LEN_BITS, s_sort.r);         // by length; stable: rank order inside a length
  for (int r = tid; r < nrows; r += BT) {
    const int rank = (int)(keys[r] & 0xffffffffu);
    const int len = E_HARD - (int)(keys[r] >> 32);
    rrow[r] = a.ar[row0 + rank];
    rlen[r] = len;
  }
  __syncthreads();
  lmax = rlen[0];
  // (b) jptr[p] = sum_{p' < p} #(rows with len > p'); rows are sorted, so #(len > p) = first r with len <= p
  for (int p = tid; p <= lmax; p += BT) jptr[p] = 0;
  __syncthreads();
  for (int p = tid; p < lmax; p += BT) {
    int lo = 0, hi = nrows;                  // first r with rlen[r] <= p
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (rlen[mid] > p) lo = mid + 1; else hi = mid;
    }
    jptr[p] = lo;
  }
  __syncthreads();
  block_excl_scan(jptr, lmax + 1, scratch);   // jptr[lmax] = nent
  // (c) entries
  const double inv = 1.0 / (double)a.seg_smax[seg];
  double wsum = 0.0;
  int rfrag_local = 0;
  for (int r = tid; r < nrows; r += BT) {
    const int row = rrow[r], len = rlen[r];
    if (len <= FRAG * FRAG_MAXG_ROW) rfrag_local += frag_group(len);
    const int b = a.row_ptr[row];
    double w = 0.0;
    for (int p = 0; p < len; ++p) {
      const int e = jptr[p] + r;
      const int g = a.col_idx[b + p];
      const double q = expm1(((double)a.score[b + p] * inv) * 100.0);   // model.py:129-130, sparse_plus.py:128
      w = fmax(w, q);
      a.e_q[eo + e] = q;
      a.e_opos[eo + e] = b + p;
      keys[e] = ((unsigned long long)(unsigned)g << 32) | ((unsigned long long)(unsigned)r << 16) | (unsigned)e;
    }
    a.r_len[ro + r] = (uint16_t)len;
    a.r_w[ro + r] = w;
    a.r_orow[ro + r] = row;
    wsum += w;
  }
  {
    // deterministic block sum of the row weights
    double v = warp_sum(wsum);
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
  }
  // zero the padding of the entry / row arrays
  for (int e = nent + tid; e < pad8(nent); e += BT) {
    a.e_q[eo + e] = 0.0; a.e_idx[eo + e] = 0u; a.e_opos[eo + e] = 0;
  }
  for (int r = nrows + tid; r < pad8(nrows); r += BT) {
    a.r_len[ro + r] = 0; a.r_w[ro + r] = 0.0; a.r_orow[ro + r] = 0;
  }
  __syncthreads();
  if (tid == 0) {
    double tsum = 0.0;
    for (int i = 0; i < BT / 32; ++i) tsum += s_red[i];
    a.t_wsum[t] = tsum;
  }
  // (d) tile columns: sort by (locus, JDS row) -- least significant field first
  if (nent <= HALF) {
    SortR::sort(keys, nent, 16, 16 + ceil_log2(RCAP), s_sort.r);
    SortR::sort(keys, nent, 32, 32 + a.gbits, s_sort.r);
  } else {
    SortE::sort(keys, nent, 16, 16 + ceil_log2(RCAP), s_sort.e);
    SortE::sort(keys, nent, 32, 32 + a.gbits, s_sort.e);
  }
  for (int i = tid; i < nent; i += BT)
    colidx[i] = (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32)) ? 1 : 0;
  __syncthreads();
  const int ncol = block_excl_scan(colidx, nent, scratch);   // colidx[i] = (#heads before i); head i has the column index
  // the scan is exclusive: column of entry i = colidx[i] + head(i) - 1 -> recompute head
  for (int i = tid; i < nent; i += BT) {
    const bool head = (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32));
    const int ci = colidx[i] + (head ? 1 : 0) - 1;
    colidx[i] = ci;
    if (head) colstart[ci] = i;
  }
  if (tid == 0) colstart[ncol] = nent;
  __syncthreads();
  // order columns by (count desc, locus asc)
  for (int c = tid; c < ncol; c += BT) {
    const int cnt = colstart[c + 1] - colstart[c];
    const unsigned long long g = keys[colstart[c]] >> 32;
    ckeys[c] = ((unsigned long long)(unsigned)(E_HARD - cnt) << 48) | (g << 16) | (unsigned)c;
  }
  __syncthreads();
  if (ncol <= HALF) SortR::sort(ckeys, ncol, 48, 48 + LEN_BITS, s_sort.r);    // by count; stable: the columns are in locus order
  else SortE::sort(ckeys, ncol, 48, 48 + LEN_BITS, s_sort.e);
  int nlong_local = 0;
  const long long co_tmp = eo + 8LL * t;
  for (int lc = tid; lc < ncol; lc += BT) {
    const int c = (int)(ckeys[lc] & 0xffffu);
    const int cnt = E_HARD - (int)(ckeys[lc] >> 48);
    rank_of[c] = lc;
    cptr[lc] = cnt;
    a.tmp_gcol[co_tmp + lc] = (int)((ckeys[lc] >> 16) & 0xffffffffu);
    nlong_local += cnt > LONG_COL;
    if (cnt <= FRAG * FRAG_MAXG_COL) nlong_local += frag_group(cnt) << 16;   // packed: cfrag in the high half
  }
  if (tid == 0) cptr[ncol] = 0;
  __syncthreads();
  block_excl_scan(cptr, ncol + 1, scratch);     // cptr[ncol] = nent
  for (int i = tid; i < nent; i += BT) {
    const int ci = colidx[i];
    const int lc = rank_of[ci];
    const int e = (int)(keys[i] & 0xffffu);
    a.e_idx[eo + e] = pack_idx((unsigned)lc, (unsigned)(cptr[lc] + (i - colstart[ci])));
  }
  // lane-group boundaries: rows are sorted by length, columns by count (both descending), so
  // "number of rows longer than x" is the first position whose length is <= x
  if (tid < 10) {
    const int x = tid < 4 ? (4 << tid) : (4 << (tid - 4));    // 4 8 16 32 | 4 8 16 32 64 128
    int lo = 0, hi = tid < 4 ? nrows : ncol;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const int v = tid < 4 ? rlen[mid] : cptr[mid + 1] - cptr[mid];
      if (v > x) lo = mid + 1; else hi = mid;
    }
    a.t_gt[10 * t + tid] = lo;
  }
  for (int lc = tid; lc <= ncol; lc += BT) a.tmp_cptr[co_tmp + lc] = (uint16_t)cptr[lc];
  for (int p = tid; p <= lmax; p += BT) a.tmp_jptr[(long long)a.jds_stride * t + p] = (uint16_t)jptr[p];
  // number of long columns / column fragments / row fragments: block reduce
  {
    int vr = rfrag_local;
    for (int o = 16; o > 0; o >>= 1) vr += __shfl_xor_sync(0xffffffffu, vr, o);
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = vr;
    __syncthreads();
    int rtot = 0;
    for (int i = 0; i < BT / 32; ++i) rtot += scratch[i];
    int v = nlong_local;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int i = 0; i < BT / 32; ++i) tot += scratch[i];
      a.t_nlong[t] = tot & 0xffff;
      a.t_cfrag[t] = tot >> 16;
      a.t_rfrag[t] = rtot;
      a.t_ncol[t] = ncol;
      a.t_maxlen[t] = lmax;
      a.t_seg[t] = seg;
      a.t_nrows[t] = nrows;
      a.t_nent[t] = nent;
    }
  }
}

__global__ void tile_pad2_kernel(int n_tiles, const int* __restrict__ t_ncol, const int* __restrict__ t_maxlen,
                                 long long* __restrict__ pad_col, long long* __restrict__ pad_jds) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  pad_col[t] = pad8(t_ncol[t] + 1);
  pad_jds[t] = pad8(t_maxlen[t] + 1);
}

__global__ void tile_hdr_kernel(int n_tiles, const int* t_seg, const int* t_nrows, const int* t_nent, const int* t_ncol,
                                const int* t_maxlen, const int* t_nlong, const int* t_rfrag, const int* t_cfrag,
                                const int* t_gt, const long long* ent_off,
                                const long long* row_off, const long long* col_off, const long long* jds_off,
                                long long* __restrict__ hdr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  long long* h = hdr + (long long)HDR_WORDS * t;
  h[H_SEG] = t_seg[t]; h[H_NROWS] = t_nrows[t]; h[H_NENT] = t_nent[t]; h[H_NCOL] = t_ncol[t];
  h[H_MAXLEN] = t_maxlen[t]; h[H_NLONG] = t_nlong[t];
  h[H_ENT_OFF] = ent_off[t]; h[H_ROW_OFF] = row_off[t]; h[H_COL_OFF] = col_off[t]; h[H_JDS_OFF] = jds_off[t];
  h[H_RFRAG] = t_rfrag[t]; h[H_CFRAG] = t_cfrag[t];
  for (int i = 0; i < 10; ++i) h[H_RGT4 + i] = t_gt[10 * t + i];
  h[H_CGT128 + 1] = 0; h[H_CGT128 + 2] = 0;
}

// copy cptr / jptr to their final offsets, emit (seg, locus) keys of the tile columns
__global__ void tile_compact_kernel(int n_tiles, const long long* __restrict__ hdr, const uint16_t* __restrict__ tmp_cptr,
                                    const int* __restrict__ tmp_gcol, const uint16_t* __restrict__ tmp_jptr,
                                    int jds_stride, const long long* __restrict__ ent_off,
                                    const long long* __restrict__ key_off, uint16_t* __restrict__ c_ptr,
                                    uint16_t* __restrict__ j_ptr, unsigned long long* __restrict__ keys_out) {
  const int t = blockIdx.x;
  const long long* h = hdr + (long long)HDR_WORDS * t;
  const int ncol = (int)h[H_NCOL], maxlen = (int)h[H_MAXLEN], seg = (int)h[H_SEG];
  const long long co = h[H_COL_OFF], jo = h[H_JDS_OFF], src = ent_off[t] + 8LL * t, ko = key_off[t];
  for (int c = threadIdx.x; c < pad8(ncol + 1); c += blockDim.x) c_ptr[co + c] = c <= ncol ? tmp_cptr[src + c] : 0;
  for (int p = threadIdx.x; p < pad8(maxlen + 1); p += blockDim.x)
    j_ptr[jo + p] = p <= maxlen ? tmp_jptr[(long long)jds_stride * t + p] : 0;
  for (int c = threadIdx.x; c < ncol; c += blockDim.x)
    keys_out[ko + c] = ((unsigned long long)(unsigned)seg << 32) | (unsigned)tmp_gcol[src + c];
}

// (seg, locus) keys of the unique rows
__global__ void uniq_keys_kernel(int n, const int* __restrict__ u_row, const int* __restrict__ row_ptr,
                                 const int* __restrict__ col_idx, const int* __restrict__ row_seg,
                                 int* __restrict__ u_pos, unsigned long long* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = u_row[i];
  const int p = row_ptr[r];
  u_pos[i] = p;
  keys[i] = ((unsigned long long)(unsigned)row_seg[r] << 32) | (unsigned)col_idx[p];
}

__device__ __forceinline__ long long lower_bound_u64(const unsigned long long* a, long long n, unsigned long long v) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void seg_cols_kernel(int n_seg, const unsigned long long* __restrict__ sc_key, long long n_sc,
                                int* __restrict__ seg_col0, int* __restrict__ seg_ncol) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  const long long a = lower_bound_u64(sc_key, n_sc, (unsigned long long)(unsigned)s << 32);
  const long long b = lower_bound_u64(sc_key, n_sc, (unsigned long long)(unsigned)(s + 1) << 32);
  seg_col0[s] = (int)a;
  seg_ncol[s] = (int)(b - a);
}

__global__ void sc_gcol_kernel(long long n, const unsigned long long* __restrict__ sc_key, int* __restrict__ gcol) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gcol[i] = (int)(sc_key[i] & 0xffffffffu);
}

// unique rows sorted by (seg, locus): one thread per run accumulates pisum0, count, sum log Q in row order
__global__ void uniq_runs_kernel(int n, const unsigned long long* __restrict__ ukeys_sorted,
                                 const int* __restrict__ uidx_sorted, const int* __restrict__ u_pos,
                                 const uint16_t* __restrict__ score, const int* __restrict__ seg_smax,
                                 const unsigned long long* __restrict__ sc_key, long long n_sc,
                                 double* __restrict__ sc_pisum0, int* __restrict__ sc_nuniq,
                                 double* __restrict__ sc_logq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = ukeys_sorted[i];
  if (i > 0 && ukeys_sorted[i - 1] == key) return;      // not a run head
  const int seg = (int)(key >> 32);
  const double inv = 1.0 / (double)seg_smax[seg];
  double ps = 0.0, lq = 0.0;
  int cnt = 0;
  for (int k = i; k < n && ukeys_sorted[k] == key; ++k) {
    const double q = expm1(((double)score[u_pos[uidx_sorted[k]]] * inv) * 100.0);
    ps += q;
    lq += log(q);
    ++cnt;
  }
  const long long slot = lower_bound_u64(sc_key, n_sc, key);
  sc_pisum0[slot] = ps;
  sc_nuniq[slot] = cnt;
  sc_logq[slot] = lq;
}

// tile column -> segment column slot, copies of pisum0 / nuniq in tile order
__global__ void tile_scol_kernel(const long long* __restrict__ hdr, const long long* __restrict__ ent_off,
                                 const int* __restrict__ tmp_gcol, const int* __restrict__ seg_col0,
                                 const int* __restrict__ seg_ncol, const int* __restrict__ sc_gcol,
                                 const double* __restrict__ sc_pisum0, const int* __restrict__ sc_nuniq,
                                 int8_t* __restrict__ sc_inamb, int* __restrict__ c_scol,
                                 double* __restrict__ c_pisum0, int* __restrict__ c_nuniq) {
  const int t = blockIdx.x;
  const long long* h = hdr + (long long)HDR_WORDS * t;
  const int ncol = (int)h[H_NCOL], seg = (int)h[H_SEG];
  const long long co = h[H_COL_OFF], src = ent_off[t] + 8LL * t;
  const int c0 = seg_col0[seg], nc = seg_ncol[seg];
  for (int c = threadIdx.x; c < pad8(ncol + 1); c += blockDim.x) {
    if (c < ncol) {
      const int g = tmp_gcol[src + c];
      int lo = 0, hi = nc;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sc_gcol[c0 + mid] < g) lo = mid + 1; else hi = mid;
      }
      const int slot = c0 + lo;
      c_scol[co + c] = slot;
      c_pisum0[co + c] = sc_pisum0[slot];
      c_nuniq[co + c] = sc_nuniq[slot];
      sc_inamb[slot] = 1;
    } else {
      c_scol[co + c] = 0;
      c_pisum0[co + c] = 0.0;
      c_nuniq[co + c] = 0;
    }
  }
}

// per segment totals: W_amb = sum of tile weight sums, W_uni = sum of pisum0, log Q of unique rows, w_max
__global__ void seg_totals_kernel(int n_seg, const int* __restrict__ seg_tile0, const int* __restrict__ seg_ntiles,
                                  const double* __restrict__ t_wsum, const int* __restrict__ seg_col0,
                                  const int* __restrict__ seg_ncol, const double* __restrict__ sc_pisum0,
                                  const double* __restrict__ sc_logq, const int* __restrict__ seg_smax,
                                  double* __restrict__ seg_wamb, double* __restrict__ seg_wtot,
                                  double* __restrict__ seg_logq, double* __restrict__ seg_wmax) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_seg) return;
  const int s = warp;
  double wa = 0.0, wu = 0.0, lq = 0.0;
  for (int t = lane; t < seg_ntiles[s]; t += 32) wa += t_wsum[seg_tile0[s] + t];
  for (int j = lane; j < seg_ncol[s]; j += 32) {
    wu += sc_pisum0[seg_col0[s] + j];
    lq += sc_logq[seg_col0[s] + j];
  }
  wa = warp_sum(wa);
  wu = warp_sum(wu);
  lq = warp_sum(lq);
  if (lane == 0) {
    seg_wamb[s] = wa;
    seg_wtot[s] = wa + wu;
    seg_logq[s] = lq;
    const int sm = seg_smax[s];
    seg_wmax[s] = sm > 0 ? expm1(((double)sm * (1.0 / (double)sm)) * 100.0) : 0.0;   // model.py:189-190
  }
}

__global__ void narrow_kernel(int n, const long long* __restrict__ in, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)in[i];
}
__global__ void widen_int_kernel(int n, const int* __restrict__ in, long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

inline unsigned grid_for(long long n, int b = 256) { return (unsigned)std::max<long long>(1, (n + b - 1) / b); }

}  // namespace

int layout_build(ss_handle_s* h, cudaStream_t st, int n_rows, int n_cols, long long nnz, const int* row_ptr,
                 const int* col_idx, const uint16_t* score, const int* row_segment, int n_segments) {
  free_fit(h);
  free_dense(h);
  free_layout(h);
  Layout& L = h->L;
  auto& pool = h->layout_allocs;
  TmpPool tmp(st);
  const int N = n_rows, S = n_segments;
  L.N = N; L.K = n_cols; L.S = S; L.A = nnz;
  const size_t Sz = (size_t)std::max(S, 1);

  // ---- step 1 ----
  uint8_t *is_amb, *is_uni;
  int* counters;
  unsigned long long *segA_u, *segAamb_u;
  SS_CUDA_CHECK(h, tmp.get(&is_amb, (size_t)N));
  SS_CUDA_CHECK(h, tmp.get(&is_uni, (size_t)N));
  SS_CUDA_CHECK(h, tmp.get(&counters, 8));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_smax, Sz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_nobs, Sz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_namb, Sz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_nuniq, Sz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_A, Sz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_Aamb, Sz));
  segA_u = reinterpret_cast<unsigned long long*>(L.seg_A);
  segAamb_u = reinterpret_cast<unsigned long long*>(L.seg_Aamb);
  SS_CUDA_CHECK(h, cudaMemsetAsync(counters, 0, 8 * sizeof(int), st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(L.seg_smax, 0, Sz * 4, st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(L.seg_nobs, 0, Sz * 4, st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(L.seg_namb, 0, Sz * 4, st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(L.seg_nuniq, 0, Sz * 4, st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(L.seg_A, 0, Sz * 8, st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(L.seg_Aamb, 0, Sz * 8, st));
  if (N > 0) {
    row_class_kernel<<<grid_for(N), 256, 0, st>>>(N, n_cols, row_ptr, col_idx, score, row_segment, S, is_amb, is_uni, L.seg_smax,
                                                  L.seg_nobs, L.seg_namb, L.seg_nuniq, segA_u, segAamb_u, counters);
    h->launches++;
  }
  // rows sharded over ranks: s_max is the maximum of the WHOLE matrix (model.py:115)
  if (int rc = dense_allreduce_smax(h, st)) return rc;
  // ---- step 2: ambiguous / unique row lists ----
  int *iota, *ar_unsorted, *d_nsel;
  SS_CUDA_CHECK(h, tmp.get(&iota, (size_t)N));
  SS_CUDA_CHECK(h, tmp.get(&ar_unsorted, (size_t)N));
  SS_CUDA_CHECK(h, tmp.get(&d_nsel, 2));
  int* u_row_tmp;
  SS_CUDA_CHECK(h, tmp.get(&u_row_tmp, (size_t)N));
  int n_amb = 0, n_uniq = 0, maxlen = 0;
  if (N > 0) {
    iota_kernel<<<grid_for(N), 256, 0, st>>>(N, iota);
    h->launches++;
    CUB_CALL(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, iota, is_amb, ar_unsorted, d_nsel, N, st));
    CUB_CALL(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, iota, is_uni, u_row_tmp, d_nsel + 1, N, st));
    int hsel[2];
    SS_CUDA_CHECK(h, cudaMemcpyAsync(hsel, d_nsel, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    int hcnt[2] = {0, 0};
    SS_CUDA_CHECK(h, cudaMemcpyAsync(hcnt, counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
    n_amb = hsel[0];
    n_uniq = hsel[1];
    maxlen = hcnt[0];
    if (hcnt[1]) {
      h->set_error("score matrix is not canonical CSR (needs non-zero scores and strictly ascending column "
                   "indices in [0, K) inside every row)");
      return SS_ERR_NOT_CANONICAL;
    }
  }
  if (maxlen > E_HARD / 2) {
    h->set_error("a read has " + std::to_string(maxlen) + " alignments; the tile capacity is " +
                 std::to_string(E_HARD / 2));
    return SS_ERR_CAPACITY;
  }
  const int chunk = maxlen ? E_HARD - maxlen + 1 : E_HARD;
  const int chunk_small = maxlen <= E_STREAM / 2 ? E_STREAM - maxlen + 1 : chunk;
  L.maxlen = maxlen;
  L.chunk = chunk;
  L.n_uniq = n_uniq;

  // sort ambiguous rows by segment (stable: row order inside a segment)
  int *aseg_in, *aseg, *ar;
  SS_CUDA_CHECK(h, tmp.get(&aseg_in, (size_t)n_amb));
  SS_CUDA_CHECK(h, tmp.get(&aseg, (size_t)n_amb));
  SS_CUDA_CHECK(h, tmp.get(&ar, (size_t)n_amb));
  long long *alen, *pre;
  SS_CUDA_CHECK(h, tmp.get(&alen, (size_t)n_amb + 1));
  SS_CUDA_CHECK(h, tmp.get(&pre, (size_t)n_amb + 1));
  int n_tiles = 0;
  long long *seg_row0, *seg_namb64, *seg_ntiles64, *seg_tile0_64, *seg_ent0;
  SS_CUDA_CHECK(h, tmp.get(&seg_row0, Sz + 1));
  SS_CUDA_CHECK(h, tmp.get(&seg_namb64, Sz + 1));
  SS_CUDA_CHECK(h, tmp.get(&seg_ntiles64, Sz + 1));
  SS_CUDA_CHECK(h, tmp.get(&seg_tile0_64, Sz + 1));
  SS_CUDA_CHECK(h, tmp.get(&seg_ent0, Sz + 1));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_tile0, Sz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_ntiles, Sz));
  SS_CUDA_CHECK(h, cudaMemsetAsync(L.seg_tile0, 0, Sz * 4, st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(L.seg_ntiles, 0, Sz * 4, st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(alen, 0, ((size_t)n_amb + 1) * 8, st));
  int* rtile = nullptr;
  int* tile_row0 = nullptr;
  if (n_amb > 0) {
    // stable sort by (segment, first locus): row order inside equal keys is kept
    unsigned long long *rkey_in, *rkey;
    SS_CUDA_CHECK(h, tmp.get(&rkey_in, (size_t)n_amb));
    SS_CUDA_CHECK(h, tmp.get(&rkey, (size_t)n_amb));
    row_sortkey_kernel<<<grid_for(n_amb), 256, 0, st>>>(n_amb, ar_unsorted, row_segment, row_ptr, col_idx, rkey_in);
    h->launches++;
    int bits = 33;
    while ((1LL << (bits - 32)) < S && bits < 64) ++bits;
    CUB_CALL(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, rkey_in, rkey, ar_unsorted, ar, n_amb, 0, bits, st));
    key_to_seg_kernel<<<grid_for(n_amb), 256, 0, st>>>(n_amb, rkey, aseg);
    h->launches++;
    gather_seg_len_kernel<<<grid_for(n_amb), 256, 0, st>>>(n_amb, ar, row_segment, row_ptr, nullptr, alen);
    h->launches++;
    CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, alen, pre, n_amb + 1, st));
    // ---- steps 3/4: tiles ----
    widen_int_kernel<<<grid_for(S), 256, 0, st>>>(S, L.seg_namb, seg_namb64);
    h->launches++;
    CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, seg_namb64, seg_row0, S, st));
    seg_tiles_kernel<<<grid_for(S), 256, 0, st>>>(S, n_amb, seg_row0, L.seg_namb, pre, chunk, chunk_small, h->small_tile_thr, std::min(chunk, h->cluster_tile), L.seg_Aamb,
                                                  seg_ntiles64, seg_ent0);
    h->launches++;
    CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, seg_ntiles64, seg_tile0_64, S, st));
    narrow_kernel<<<grid_for(S), 256, 0, st>>>(S, seg_ntiles64, L.seg_ntiles);
    narrow_kernel<<<grid_for(S), 256, 0, st>>>(S, seg_tile0_64, L.seg_tile0);
    h->launches += 2;
    long long last_t0 = 0, last_nt = 0;
    SS_CUDA_CHECK(h, cudaMemcpyAsync(&last_t0, seg_tile0_64 + (S - 1), 8, cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaMemcpyAsync(&last_nt, seg_ntiles64 + (S - 1), 8, cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
    n_tiles = (int)(last_t0 + last_nt);
    SS_CUDA_CHECK(h, tmp.get(&rtile, (size_t)n_amb));
    SS_CUDA_CHECK(h, tmp.get(&tile_row0, (size_t)n_tiles + 1));
    row_tile_kernel<<<grid_for(n_amb), 256, 0, st>>>(n_amb, aseg, pre, seg_ent0, seg_tile0_64, chunk, chunk_small, h->small_tile_thr,
                                                     std::min(chunk, h->cluster_tile), L.seg_Aamb, rtile);
    tile_heads_kernel<<<grid_for(n_amb), 256, 0, st>>>(n_amb, n_tiles, rtile, tile_row0);
    h->launches += 2;
  }
  L.n_tiles = n_tiles;
  const size_t Tz = (size_t)std::max(n_tiles, 1);

  // offsets of entry / row arrays
  long long *pad_ent, *pad_row, *ent_off, *row_off, *pad_col, *pad_jds, *col_off, *jds_off, *key_cnt, *key_off;
  SS_CUDA_CHECK(h, tmp.get(&pad_ent, Tz + 1));
  SS_CUDA_CHECK(h, tmp.get(&pad_row, Tz + 1));
  SS_CUDA_CHECK(h, tmp.get(&ent_off, Tz + 1));
  SS_CUDA_CHECK(h, tmp.get(&row_off, Tz + 1));
  SS_CUDA_CHECK(h, tmp.get(&pad_col, Tz + 1));
  SS_CUDA_CHECK(h, tmp.get(&pad_jds, Tz + 1));
  SS_CUDA_CHECK(h, tmp.get(&col_off, Tz + 1));
  SS_CUDA_CHECK(h, tmp.get(&jds_off, Tz + 1));
  SS_CUDA_CHECK(h, tmp.get(&key_cnt, Tz + 1));
  SS_CUDA_CHECK(h, tmp.get(&key_off, Tz + 1));
  int *t_ncol, *t_maxlen, *t_nlong, *t_seg, *t_nrows, *t_nent, *t_rfrag, *t_cfrag;
  double* t_wsum;
  SS_CUDA_CHECK(h, tmp.get(&t_rfrag, Tz));
  SS_CUDA_CHECK(h, tmp.get(&t_cfrag, Tz));
  SS_CUDA_CHECK(h, tmp.get(&t_ncol, Tz));
  SS_CUDA_CHECK(h, tmp.get(&t_maxlen, Tz));
  SS_CUDA_CHECK(h, tmp.get(&t_nlong, Tz));
  SS_CUDA_CHECK(h, tmp.get(&t_seg, Tz));
  SS_CUDA_CHECK(h, tmp.get(&t_nrows, Tz));
  SS_CUDA_CHECK(h, tmp.get(&t_nent, Tz));
  SS_CUDA_CHECK(h, tmp.get(&t_wsum, Tz));
  int* t_gt;
  SS_CUDA_CHECK(h, tmp.get(&t_gt, Tz * 10));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.tile_hdr, Tz * HDR_WORDS));
  long long tot[4] = {0, 0, 0, 0};
  uint16_t* tmp_cptr = nullptr;
  int* tmp_gcol = nullptr;
  uint16_t* tmp_jptr = nullptr;
  const int jds_stride = pad8(maxlen + 1);
  long long n_tilekeys = 0;
  if (n_tiles > 0) {
    SS_CUDA_CHECK(h, cudaMemsetAsync(pad_ent + n_tiles, 0, 8, st));
    SS_CUDA_CHECK(h, cudaMemsetAsync(pad_row + n_tiles, 0, 8, st));
    tile_sizes_kernel<<<grid_for(n_tiles), 256, 0, st>>>(n_tiles, tile_row0, pre, pad_ent, pad_row);
    h->launches++;
    CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, pad_ent, ent_off, n_tiles + 1, st));
    CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, pad_row, row_off, n_tiles + 1, st));
    SS_CUDA_CHECK(h, cudaMemcpyAsync(&tot[0], ent_off + n_tiles, 8, cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaMemcpyAsync(&tot[1], row_off + n_tiles, 8, cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  }
  L.tot_ent = tot[0];
  L.tot_row = tot[1];
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.e_q, (size_t)L.tot_ent));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.e_idx, (size_t)L.tot_ent));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.e_opos, (size_t)L.tot_ent));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.r_len, (size_t)L.tot_row));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.r_w, (size_t)L.tot_row));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.r_orow, (size_t)L.tot_row));
  if (n_tiles > 0) {
    const size_t tmpc = (size_t)L.tot_ent + 8 * (size_t)n_tiles + 8;
    SS_CUDA_CHECK(h, tmp.get(&tmp_cptr, tmpc));
    SS_CUDA_CHECK(h, tmp.get(&tmp_gcol, tmpc));
    SS_CUDA_CHECK(h, tmp.get(&tmp_jptr, (size_t)jds_stride * n_tiles));
    TileBuildArgs a;
    a.row_ptr = row_ptr; a.col_idx = col_idx; a.score = score; a.ar = ar; a.aseg = aseg; a.tile_row0 = tile_row0;
    a.pre = pre; a.ent_off = ent_off; a.row_off = row_off; a.seg_smax = L.seg_smax; a.n_tiles = n_tiles;
    a.e_q = L.e_q; a.e_idx = L.e_idx; a.e_opos = L.e_opos;
    a.r_len = L.r_len; a.r_w = L.r_w; a.r_orow = L.r_orow;
    a.tmp_cptr = tmp_cptr; a.tmp_gcol = tmp_gcol; a.tmp_jptr = tmp_jptr; a.jds_stride = jds_stride;
    a.gbits = 1;
    while ((1LL << a.gbits) < (long long)n_cols) ++a.gbits;
    a.t_ncol = t_ncol; a.t_maxlen = t_maxlen; a.t_nlong = t_nlong; a.t_seg = t_seg; a.t_nrows = t_nrows;
    a.t_nent = t_nent; a.t_wsum = t_wsum; a.t_rfrag = t_rfrag; a.t_cfrag = t_cfrag; a.t_gt = t_gt;
    // the three size classes work on disjoint tiles: run them side by side on internal streams
    SS_CUDA_CHECK(h, ensure_aux_streams(h));
    SS_CUDA_CHECK(h, cudaEventRecord(h->ev_fork, st));
#define SS_TILE_BUILD(ECAP, ELOW, BT, SIDX)                                                               \
  do {                                                                                                    \
    constexpr size_t smem = tile_build_smem<ECAP, BT>();                                                  \
    cudaStream_t sx = h->aux[SIDX];                                                                       \
    SS_CUDA_CHECK(h, cudaStreamWaitEvent(sx, h->ev_fork, 0));                                             \
    SS_CUDA_CHECK(h, cudaFuncSetAttribute(tile_build_kernel<ECAP, ELOW, BT>,                              \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    tile_build_kernel<ECAP, ELOW, BT><<<n_tiles, BT, smem, sx>>>(a);                                      \
    h->launches++;                                                                                        \
    SS_CUDA_CHECK(h, cudaGetLastError());                                                                 \
    SS_CUDA_CHECK(h, cudaEventRecord(h->ev_join[SIDX], sx));                                              \
    SS_CUDA_CHECK(h, cudaStreamWaitEvent(st, h->ev_join[SIDX], 0));                                       \
  } while (0)
    SS_TILE_BUILD(E_HARD, 1024, 1024, 0);      // largest tiles: one CTA per SM
    SS_TILE_BUILD(1024, 256, 256, 1);
    SS_TILE_BUILD(256, 0, 128, 2);
#undef SS_TILE_BUILD
    SS_CUDA_CHECK(h, cudaMemsetAsync(pad_col + n_tiles, 0, 8, st));
    SS_CUDA_CHECK(h, cudaMemsetAsync(pad_jds + n_tiles, 0, 8, st));
    SS_CUDA_CHECK(h, cudaMemsetAsync(key_cnt + n_tiles, 0, 8, st));
    tile_pad2_kernel<<<grid_for(n_tiles), 256, 0, st>>>(n_tiles, t_ncol, t_maxlen, pad_col, pad_jds);
    widen_int_kernel<<<grid_for(n_tiles), 256, 0, st>>>(n_tiles, t_ncol, key_cnt);
    h->launches += 2;
    CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, pad_col, col_off, n_tiles + 1, st));
    CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, pad_jds, jds_off, n_tiles + 1, st));
    CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, key_cnt, key_off, n_tiles + 1, st));
    SS_CUDA_CHECK(h, cudaMemcpyAsync(&tot[2], col_off + n_tiles, 8, cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaMemcpyAsync(&tot[3], jds_off + n_tiles, 8, cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaMemcpyAsync(&n_tilekeys, key_off + n_tiles, 8, cudaMemcpyDeviceToHost, st));
    tile_hdr_kernel<<<grid_for(n_tiles), 256, 0, st>>>(n_tiles, t_seg, t_nrows, t_nent, t_ncol, t_maxlen, t_nlong,
                                                       t_rfrag, t_cfrag, t_gt, ent_off, row_off, col_off, jds_off,
                                                       L.tile_hdr);
    h->launches++;
    h->hdr_host.resize((size_t)n_tiles * HDR_WORDS);
    SS_CUDA_CHECK(h, cudaMemcpyAsync(h->hdr_host.data(), L.tile_hdr, (size_t)n_tiles * HDR_WORDS * 8,
                                     cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  }
  L.tot_col = tot[2];
  L.tot_jds = tot[3];
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.c_ptr, (size_t)L.tot_col));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.c_scol, (size_t)L.tot_col));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.c_pisum0, (size_t)L.tot_col));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.c_nuniq, (size_t)L.tot_col));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.j_ptr, (size_t)L.tot_jds));

  // ---- step 6: segment columns ----
  const long long n_keys = n_tilekeys + n_uniq;
  unsigned long long *keys_all, *keys_sorted, *sc_key;
  SS_CUDA_CHECK(h, tmp.get(&keys_all, (size_t)n_keys));
  SS_CUDA_CHECK(h, tmp.get(&keys_sorted, (size_t)n_keys));
  SS_CUDA_CHECK(h, tmp.get(&sc_key, (size_t)n_keys));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.u_row, (size_t)n_uniq));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.u_pos, (size_t)n_uniq));
  if (n_tiles > 0) {
    tile_compact_kernel<<<n_tiles, 128, 0, st>>>(n_tiles, L.tile_hdr, tmp_cptr, tmp_gcol, tmp_jptr, jds_stride, ent_off,
                                                 key_off, L.c_ptr, L.j_ptr, keys_all);
    h->launches++;
  }
  if (n_uniq > 0) {
    SS_CUDA_CHECK(h, cudaMemcpyAsync(L.u_row, u_row_tmp, (size_t)n_uniq * 4, cudaMemcpyDeviceToDevice, st));
    uniq_keys_kernel<<<grid_for(n_uniq), 256, 0, st>>>(n_uniq, L.u_row, row_ptr, col_idx, row_segment, L.u_pos,
                                                       keys_all + n_tilekeys);
    h->launches++;
  }
  long long n_sc = 0;
  if (n_keys > 0) {
    if (n_keys > 0x7fffffffLL) {
      h->set_error("too many (segment, locus) pairs");
      return SS_ERR_CAPACITY;
    }
    int kbits = 33;
    while ((1LL << (kbits - 32)) < S && kbits < 64) ++kbits;
    CUB_CALL(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, keys_all, keys_sorted, (int)n_keys, 0, kbits, st));
    CUB_CALL(cub::DeviceSelect::Unique(d_tmp, tmp_bytes, keys_sorted, sc_key, d_nsel, (int)n_keys, st));
    int nsel = 0;
    SS_CUDA_CHECK(h, cudaMemcpyAsync(&nsel, d_nsel, 4, cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
    n_sc = nsel;
  }
  L.n_segcols = n_sc;
  const size_t NCz = (size_t)std::max<long long>(n_sc, 1);
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_col0, Sz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_ncol, Sz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.sc_gcol, NCz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.sc_pisum0, NCz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.sc_nuniq, NCz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->sc_inamb, NCz));
  double* sc_logq;
  SS_CUDA_CHECK(h, tmp.get(&sc_logq, NCz));
  SS_CUDA_CHECK(h, cudaMemsetAsync(L.sc_pisum0, 0, NCz * 8, st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(L.sc_nuniq, 0, NCz * 4, st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(h->sc_inamb, 0, NCz, st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(sc_logq, 0, NCz * 8, st));
  if (S > 0) {
    seg_cols_kernel<<<grid_for(S), 256, 0, st>>>(S, sc_key, n_sc, L.seg_col0, L.seg_ncol);
    h->launches++;
  }
  if (n_sc > 0) {
    sc_gcol_kernel<<<grid_for(n_sc), 256, 0, st>>>(n_sc, sc_key, L.sc_gcol);
    h->launches++;
  }
  if (n_uniq > 0) {
    unsigned long long* uk_sorted;
    int *uidx, *uidx_sorted;
    SS_CUDA_CHECK(h, tmp.get(&uk_sorted, (size_t)n_uniq));
    SS_CUDA_CHECK(h, tmp.get(&uidx, (size_t)n_uniq));
    SS_CUDA_CHECK(h, tmp.get(&uidx_sorted, (size_t)n_uniq));
    iota_kernel<<<grid_for(n_uniq), 256, 0, st>>>(n_uniq, uidx);
    h->launches++;
    int kbits = 33;
    while ((1LL << (kbits - 32)) < S && kbits < 64) ++kbits;
    CUB_CALL(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, keys_all + n_tilekeys, uk_sorted, uidx, uidx_sorted,
                                             n_uniq, 0, kbits, st));
    uniq_runs_kernel<<<grid_for(n_uniq), 256, 0, st>>>(n_uniq, uk_sorted, uidx_sorted, L.u_pos, score, L.seg_smax,
                                                       sc_key, n_sc, L.sc_pisum0, L.sc_nuniq, sc_logq);
    h->launches++;
  }
  // ---- step 7: tile columns -> segment columns; segment totals ----
  if (n_tiles > 0) {
    tile_scol_kernel<<<n_tiles, 128, 0, st>>>(L.tile_hdr, ent_off, tmp_gcol, L.seg_col0, L.seg_ncol, L.sc_gcol,
                                              L.sc_pisum0, L.sc_nuniq, h->sc_inamb, L.c_scol, L.c_pisum0, L.c_nuniq);
    h->launches++;
  }
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_wmax, Sz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_wtot, Sz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_wamb, Sz));
  SS_CUDA_CHECK(h, dev_alloc(pool, &L.seg_logq_uniq, Sz));
  if (S > 0) {
    seg_totals_kernel<<<grid_for((long long)S * 32), 256, 0, st>>>(S, L.seg_tile0, L.seg_ntiles, t_wsum, L.seg_col0,
                                                                    L.seg_ncol, L.sc_pisum0, sc_logq, L.seg_smax,
                                                                    L.seg_wamb, L.seg_wtot, L.seg_logq_uniq,
                                                                    L.seg_wmax);
    h->launches++;
  }
  SS_CUDA_CHECK(h, cudaGetLastError());
  // ---- launch plan ----
  h->seg_ntiles_host.assign(Sz, 0);
  h->seg_tile0_host.assign(Sz, 0);
  std::vector<int> ncol_host(Sz, 0), col0_host(Sz, 0);
  if (S > 0) {
    SS_CUDA_CHECK(h, cudaMemcpyAsync(h->seg_ntiles_host.data(), L.seg_ntiles, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaMemcpyAsync(h->seg_tile0_host.data(), L.seg_tile0, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaMemcpyAsync(ncol_host.data(), L.seg_ncol, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaMemcpyAsync(col0_host.data(), L.seg_col0, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
  }
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  int rc = layout_plan(h, st, ncol_host, col0_host);
  if (rc != SS_OK) return rc;
  h->has_layout = true;
  return SS_OK;
}

}  // namespace ss
