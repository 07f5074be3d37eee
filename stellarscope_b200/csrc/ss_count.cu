// Sort-free per-cell count aggregation (integer counts).
//
// Reference being replaced: Stellarscope.output_report agg_bc (utils/model.py:1302-1320) with
// csr_matrix_plus.groupby_sum (utils/sparse_plus.py:367-396): counts[locus, cell] = number of selected
// alignments of the cell's rows at the locus.
//
// ss_count (ss_reassign.cu) sorts one 64-bit (cell, locus) key per selected alignment -- 5x10^8 keys for
// total_hits on the atlas configuration (0.94 s, round 1).  Nothing needs sorting: a cell's rows are grouped
// with one counting pass over the rows, then ONE CTA per cell accumulates the cell's selected alignments into
// shared-memory bins over the loci (shared-memory atomics only), remembers which bins it touched, sorts that
// short list in shared memory and writes (locus, count) in ascending locus order.  The result is the CSC form of
// the (loci x cells) count matrix: column pointers, loci, int32 counts -- 8 bytes per entry to bring home instead
// of 16.
#include <cub/cub.cuh>

#include "ss_handle.h"

namespace ss {

namespace {

constexpr int CC_THREADS = 1024;
constexpr int CC_LIST = 8192;          // touched-bin list kept in shared memory; beyond it the bins are scanned
constexpr int CC_LIST_FRAC = 1024;     // ... of the fractional variant (its bins are doubles: 224 KB for 28k loci)

// rows and selected alignments per cell; sortedness of row_cell
__global__ void cc_hist_kernel(int n_rows, int n_cells, const int* __restrict__ row_ptr, const uint8_t* __restrict__ flag,
                               const int* __restrict__ row_cell, int* __restrict__ cell_rows,
                               unsigned long long* __restrict__ cell_sel, int* __restrict__ unsorted) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  int c = -1, sel = 0;
  if (row < n_rows) {
    c = row_cell[row];
    if (c >= n_cells) c = -1;
    if (c >= 0) {
      for (int k = row_ptr[row]; k < row_ptr[row + 1]; ++k) sel += flag[k] != 0;
    }
    if (row + 1 < n_rows) {
      const int cn = row_cell[row + 1];
      if (c >= 0 && cn >= 0 && cn < c) *unsorted = 1;
      if ((c < 0) != (cn < 0 || cn >= n_cells) ) *unsorted = 1;     // dropped rows inside: take the general path
    }
  }
  // warp aggregation: the rows of a cell are usually neighbours
  const unsigned full = 0xffffffffu;
  const int c0 = __shfl_sync(full, c, 0);
  if (__all_sync(full, c == c0)) {
    int one = c >= 0 ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sel += __shfl_xor_sync(full, sel, o);
      one += __shfl_xor_sync(full, one, o);
    }
    if ((threadIdx.x & 31) == 0 && c0 >= 0) {
      atomicAdd(&cell_rows[c0], one);
      atomicAdd(&cell_sel[c0], (unsigned long long)sel);
    }
  } else if (c >= 0) {
    atomicAdd(&cell_rows[c], 1);
    atomicAdd(&cell_sel[c], (unsigned long long)sel);
  }
}

__global__ void cc_bounds_kernel(int n_cells, int n_cols, const unsigned long long* __restrict__ cell_sel,
                                 long long* __restrict__ ub) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_cells) return;
  ub[c] = c < n_cells ? (long long)min(cell_sel[c], (unsigned long long)n_cols) : 0;
}
__global__ void cc_widen_kernel(int n, const int* __restrict__ in, long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) out[i] = i < n ? in[i] : 0;
}

// rows grouped by cell (general case: the rows of a cell are scattered over the matrix)
__global__ void cc_group_rows_kernel(int n_rows, int n_cells, const int* __restrict__ row_cell,
                                     const long long* __restrict__ cell_row0, int* __restrict__ cursor,
                                     int* __restrict__ rows_by_cell) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const int c = row_cell[row];
  if (c < 0 || c >= n_cells) return;
  rows_by_cell[cell_row0[c] + atomicAdd(&cursor[c], 1)] = row;
}

// one CTA per cell (persistent: cells are handed out through an atomic counter).  VT = unsigned int: integer counts;
// VT = double: every selected alignment of row r adds 1 / nbest[r] (best_average, model.py:448) -- the bins are
// twice as wide, the touched-bin list shorter (LIST entries), the additions are floating-point atomics (their order,
// and with it the last bit of a sum, is not fixed)
template <class VT, int LIST>
__global__ void __launch_bounds__(CC_THREADS) cc_count_kernel(
    int n_cells, int n_cols, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
    const uint8_t* __restrict__ flag, const int* __restrict__ nbest, const long long* __restrict__ cell_row0,
    const int* __restrict__ rows_by_cell, const int* __restrict__ first_row, const unsigned long long* __restrict__ cell_sel,
    const long long* __restrict__ ub_off, int* __restrict__ next_cell, unsigned int* __restrict__ tmp_col,
    VT* __restrict__ tmp_cnt, long long* __restrict__ cell_nnz) {
  constexpr int CC_LIST = LIST;
  extern __shared__ __align__(16) unsigned char dsm[];
  VT* bins = reinterpret_cast<VT*>(dsm);                                // n_cols
  unsigned int* list = reinterpret_cast<unsigned int*>(bins + ((n_cols + 3) & ~3));   // CC_LIST
  __shared__ int s_cell, s_nlist;
  __shared__ unsigned int s_scan[CC_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n_cols; i += CC_THREADS) bins[i] = VT(0);
  __syncthreads();
  while (true) {
    if (tid == 0) {
      s_cell = atomicAdd(next_cell, 1);
      s_nlist = 0;
    }
    __syncthreads();
    const int c = s_cell;
    if (c >= n_cells) break;
    const long long r0 = cell_row0[c], r1 = cell_row0[c + 1];
    if (cell_sel[c] == 0ull) {
      if (tid == 0) cell_nnz[c] = 0;
      __syncthreads();
      continue;
    }
    const long long base_row = rows_by_cell ? 0 : (long long)first_row[c];
    for (long long i = r0 + tid; i < r1; i += CC_THREADS) {
      const int row = rows_by_cell ? rows_by_cell[i] : (int)(base_row + (i - r0));
      const int b = row_ptr[row], e = row_ptr[row + 1];
      VT one = VT(1);
      if (sizeof(VT) == 8 && nbest) one = VT(1.0 / (double)max(1, nbest[row]));
      for (int k = b; k < e; ++k) {
        if (!flag[k]) continue;
        const unsigned col = (unsigned)col_idx[k];
        if (col >= (unsigned)n_cols) continue;
        if (atomicAdd(&bins[col], one) == VT(0)) {
          const int p = atomicAdd(&s_nlist, 1);
          if (p < CC_LIST) list[p] = col;
        }
      }
    }
    __syncthreads();
    const int nd = s_nlist;
    const long long o = ub_off[c];
    if (nd <= CC_LIST) {
      // bitonic sort of the touched loci (padded to a power of two with 0xffffffff)
      int np2 = 1;
      while (np2 < nd) np2 <<= 1;
      for (int i = nd + tid; i < np2; i += CC_THREADS) list[i] = 0xffffffffu;
      __syncthreads();
      for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < np2; i += CC_THREADS) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const unsigned a = list[i], b2 = list[ixj];
              const bool up = (i & k) == 0;
              if ((a > b2) == up) { list[i] = b2; list[ixj] = a; }
            }
          }
          __syncthreads();
        }
      }
      for (int i = tid; i < nd; i += CC_THREADS) {
        const unsigned col = list[i];
        tmp_col[o + i] = col;
        tmp_cnt[o + i] = bins[col];
        bins[col] = VT(0);
      }
    } else {
      // many distinct loci: scan all bins in locus order (every thread owns a contiguous chunk)
      const int chunk = (n_cols + CC_THREADS - 1) / CC_THREADS;
      const int b0 = min(n_cols, tid * chunk), b1 = min(n_cols, b0 + chunk);
      unsigned int mine = 0;
      for (int i = b0; i < b1; ++i) mine += bins[i] != VT(0);
      unsigned int incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      if (lane == 31) s_scan[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        unsigned int v = s_scan[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned int t = __shfl_up_sync(0xffffffffu, v, d);
          if (lane >= d) v += t;
        }
        s_scan[lane] = v;
      }
      __syncthreads();
      long long w = o + (long long)(incl - mine) + (warp > 0 ? s_scan[warp - 1] : 0u);
      for (int i = b0; i < b1; ++i) {
        const VT v = bins[i];
        if (v != VT(0)) {
          tmp_col[w] = (unsigned)i;
          tmp_cnt[w] = v;
          ++w;
          bins[i] = VT(0);
        }
      }
    }
    if (tid == 0) cell_nnz[c] = nd;
    __syncthreads();
  }
}

// first row of every cell when row_cell is sorted: the scan of the rows per cell counts only rows WITH a cell, so
// it is the first row only if no dropped rows precede; computed explicitly instead
__global__ void cc_first_row_kernel(int n_rows, int n_cells, const int* __restrict__ row_cell, int* __restrict__ first_row) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const int c = row_cell[row];
  if (c < 0 || c >= n_cells) return;
  if (row == 0 || row_cell[row - 1] != c) first_row[c] = row;
}

// tmp (upper-bound offsets) -> output (exact offsets): one warp per cell
template <class VT, class OT>
__global__ void cc_compact_kernel(int n_cells, const long long* __restrict__ ub_off, const long long* __restrict__ out_ptr,
                                  const unsigned int* __restrict__ tmp_col, const VT* __restrict__ tmp_cnt,
                                  int* __restrict__ out_col, OT* __restrict__ out_cnt) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= n_cells) return;
  const long long src = ub_off[c], dst = out_ptr[c];
  const long long n = out_ptr[c + 1] - dst;
  for (long long i = lane; i < n; i += 32) {
    out_col[dst + i] = (int)tmp_col[src + i];
    out_cnt[dst + i] = (OT)tmp_cnt[src + i];
  }
}

struct TmpPool {
  std::vector<void*> p;
  ~TmpPool() {
    for (void* q : p) cudaFreeAsync(q, tl_alloc_stream);
  }
  template <class T>
  cudaError_t get(T** out, size_t n) {
    return dev_alloc(p, out, n);
  }
};

}  // namespace

#define CUB_CALL(call_with_tmp)                                                      \
  do {                                                                               \
    void* d_tmp = nullptr;                                                           \
    size_t tmp_bytes = 0;                                                            \
    SS_CUDA_CHECK(h, (call_with_tmp));                                               \
    unsigned char* _t;                                                               \
    SS_CUDA_CHECK(h, tmp.get(&_t, tmp_bytes));                                       \
    d_tmp = _t;                                                                      \
    SS_CUDA_CHECK(h, (call_with_tmp));                                               \
    h->launches++;                                                                   \
  } while (0)

// returns SS_OK, or SS_ERR_CAPACITY with *n_out = required entries when cap is too small / the loci do not fit the bins
int count_cells(ss_handle_s* h, cudaStream_t st, int n_rows, int n_cols, int n_cells, const int* row_ptr, const int* col_idx,
                const uint8_t* flag, const int* nbest, const int* row_cell, long long cap, long long* out_ptr, int* out_col,
                int* out_cnt, double* out_val, long long* n_out) {
  if (!out_ptr || !n_out || n_cells < 0 || n_cols <= 0 || (out_val && out_cnt) ||
      (n_rows > 0 && (!row_ptr || !col_idx || !flag || !row_cell))) {
    h->set_error("ss_count_cells: bad argument");
    return SS_ERR_ARG;
  }
  *n_out = 0;
  const bool frac = out_val != nullptr;            // fractional counts: 1 / nbest per selected alignment
  const size_t smem = (frac ? 8 : 4) * (size_t)((n_cols + 3) & ~3) + 4 * (size_t)(frac ? CC_LIST_FRAC : CC_LIST);
  if (smem + 1024 > h->max_smem_optin) {
    h->set_error("ss_count_cells: the loci do not fit the shared-memory bins (use ss_count)");
    return SS_ERR_CAPACITY;
  }
  SS_CUDA_CHECK(h, cudaMemsetAsync(out_ptr, 0, ((size_t)n_cells + 1) * sizeof(long long), st));
  if (n_rows <= 0 || n_cells == 0) {
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
    return SS_OK;
  }
  TmpPool tmp;
  const int B = 256, GR = (n_rows + B - 1) / B, GC = (n_cells + 1 + B - 1) / B;
  int *cell_rows, *unsorted, *cursor = nullptr, *rows_by_cell = nullptr, *first_row = nullptr, *next_cell;
  unsigned long long* cell_sel;
  long long *cell_rows64, *cell_row0, *ub, *ub_off, *cell_nnz;
  SS_CUDA_CHECK(h, tmp.get(&cell_rows, (size_t)n_cells + 1));
  SS_CUDA_CHECK(h, tmp.get(&cell_sel, (size_t)n_cells + 1));
  SS_CUDA_CHECK(h, tmp.get(&unsorted, 2));
  next_cell = unsorted + 1;
  SS_CUDA_CHECK(h, tmp.get(&cell_rows64, (size_t)n_cells + 1));
  SS_CUDA_CHECK(h, tmp.get(&cell_row0, (size_t)n_cells + 1));
  SS_CUDA_CHECK(h, tmp.get(&ub, (size_t)n_cells + 1));
  SS_CUDA_CHECK(h, tmp.get(&ub_off, (size_t)n_cells + 1));
  SS_CUDA_CHECK(h, tmp.get(&cell_nnz, (size_t)n_cells + 1));
  SS_CUDA_CHECK(h, cudaMemsetAsync(cell_rows, 0, ((size_t)n_cells + 1) * sizeof(int), st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(cell_sel, 0, ((size_t)n_cells + 1) * sizeof(unsigned long long), st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(unsorted, 0, 2 * sizeof(int), st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(cell_nnz, 0, ((size_t)n_cells + 1) * sizeof(long long), st));
  cc_hist_kernel<<<GR, B, 0, st>>>(n_rows, n_cells, row_ptr, flag, row_cell, cell_rows, cell_sel, unsorted);
  cc_widen_kernel<<<GC, B, 0, st>>>(n_cells, cell_rows, cell_rows64);
  cc_bounds_kernel<<<GC, B, 0, st>>>(n_cells, n_cols, cell_sel, ub);
  h->launches += 3;
  CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, cell_rows64, cell_row0, n_cells + 1, st));
  CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, ub, ub_off, n_cells + 1, st));
  int uns = 0;
  long long ub_total = 0;
  SS_CUDA_CHECK(h, cudaMemcpyAsync(&uns, unsorted, sizeof(int), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaMemcpyAsync(&ub_total, ub_off + n_cells, sizeof(long long), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  if (ub_total == 0) return SS_OK;
  if (uns) {
    SS_CUDA_CHECK(h, tmp.get(&cursor, (size_t)n_cells));
    SS_CUDA_CHECK(h, tmp.get(&rows_by_cell, (size_t)n_rows));
    SS_CUDA_CHECK(h, cudaMemsetAsync(cursor, 0, (size_t)n_cells * sizeof(int), st));
    cc_group_rows_kernel<<<GR, B, 0, st>>>(n_rows, n_cells, row_cell, cell_row0, cursor, rows_by_cell);
  } else {
    SS_CUDA_CHECK(h, tmp.get(&first_row, (size_t)n_cells));
    SS_CUDA_CHECK(h, cudaMemsetAsync(first_row, 0, (size_t)n_cells * sizeof(int), st));
    cc_first_row_kernel<<<GR, B, 0, st>>>(n_rows, n_cells, row_cell, first_row);
  }
  h->launches++;
  unsigned int *tmp_col, *tmp_cnt = nullptr;
  double* tmp_val = nullptr;
  SS_CUDA_CHECK(h, tmp.get(&tmp_col, (size_t)ub_total));
  if (frac) SS_CUDA_CHECK(h, tmp.get(&tmp_val, (size_t)ub_total));
  else SS_CUDA_CHECK(h, tmp.get(&tmp_cnt, (size_t)ub_total));
  const void* kfn = frac ? (const void*)cc_count_kernel<double, CC_LIST_FRAC> : (const void*)cc_count_kernel<unsigned int, CC_LIST>;
  SS_CUDA_CHECK(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  SS_CUDA_CHECK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, CC_THREADS, smem));
  const int grid = std::min(n_cells, std::max(1, occ) * h->num_sms);
  if (frac)
    cc_count_kernel<double, CC_LIST_FRAC><<<grid, CC_THREADS, smem, st>>>(n_cells, n_cols, row_ptr, col_idx, flag, nbest, cell_row0,
                                                                       rows_by_cell, first_row, cell_sel, ub_off, next_cell,
                                                                       tmp_col, tmp_val, cell_nnz);
  else
    cc_count_kernel<unsigned int, CC_LIST><<<grid, CC_THREADS, smem, st>>>(n_cells, n_cols, row_ptr, col_idx, flag, nullptr, cell_row0,
                                                                        rows_by_cell, first_row, cell_sel, ub_off, next_cell,
                                                                        tmp_col, tmp_cnt, cell_nnz);
  h->launches++;
  SS_CUDA_CHECK(h, cudaGetLastError());
  CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, cell_nnz, out_ptr, n_cells + 1, st));
  long long total = 0;
  SS_CUDA_CHECK(h, cudaMemcpyAsync(&total, out_ptr + n_cells, sizeof(long long), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  *n_out = total;
  if (total > cap) {
    h->set_error("ss_count_cells: output capacity too small");
    return SS_ERR_CAPACITY;
  }
  if (total > 0) {
    const unsigned gcmp = (unsigned)(((long long)n_cells * 32 + B - 1) / B);
    if (frac) cc_compact_kernel<double, double><<<gcmp, B, 0, st>>>(n_cells, ub_off, out_ptr, tmp_col, tmp_val, out_col, out_val);
    else cc_compact_kernel<unsigned int, int><<<gcmp, B, 0, st>>>(n_cells, ub_off, out_ptr, tmp_col, tmp_cnt, out_col, out_cnt);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
  }
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));      // temporaries are freed on return
  return SS_OK;
}

}  // namespace ss
