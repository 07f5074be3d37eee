// Reassignment pass and per-cell count aggregation.
//
// Reference being replaced: TelescopeLikelihood.reassign (utils/model.py:375-507),
// csr_matrix_plus.binmax(1) (utils/sparse_plus.py:134-178) and
// Stellarscope.output_report agg_bc / agg_bc_umi (utils/model.py:1302-1345) with
// csr_matrix_plus.groupby_sum (utils/sparse_plus.py:367-396).
#include <cub/cub.cuh>

#include "ss_handle.h"

namespace ss {

constexpr int HIST_BINS = 64;

// One thread per row of the input CSR.  flag[] gets the reassignment matrix of the
// mode (as 0/1 per alignment), nbest[] the size of the row's tie set.
__global__ void reassign_kernel(int mode, double conf, double tol, int n_rows, const int* __restrict__ row_ptr,
                                const uint16_t* __restrict__ score, const double* __restrict__ z,
                                uint8_t* __restrict__ flag, int* __restrict__ nbest_out,
                                unsigned long long* __restrict__ info) {
  __shared__ unsigned int s_hist[HIST_BINS];
  __shared__ unsigned long long s_cnt[3];
  for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x) s_hist[i] = 0;
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long assigned = 0, ambiguous = 0, unaligned = 0;
  if (row < n_rows) {
    const int b = row_ptr[row], e = row_ptr[row + 1];
    const int len = e - b;
    int nb = 0;
    if (mode == SS_BEST_EXCLUDE || mode == SS_BEST_RANDOM || mode == SS_BEST_AVERAGE) {
      // binmax(1): entries equal to the row maximum (implicit zeros included, so the
      // maximum of a row is >= 0 and an all-zero row selects nothing)
      double m = 0.0;
      for (int k = b; k < e; ++k) m = fmax(m, z[k]);
      const double thr = m * (1.0 - tol);
      if (m > 0.0)
        for (int k = b; k < e; ++k) nb += (z[k] >= thr);
      const bool keep_all = mode != SS_BEST_EXCLUDE || nb == 1;
      for (int k = b; k < e; ++k) flag[k] = (m > 0.0 && z[k] >= thr && keep_all) ? 1 : 0;
      assigned = nb == 1;
      ambiguous = nb > 1;
      unaligned = nb == 0;
      atomicAdd(&s_hist[min(nb, HIST_BINS - 1)], 1u);
    } else if (mode == SS_BEST_CONF) {
      double m = 0.0;
      for (int k = b; k < e; ++k) {
        const double v = z[k];
        m = fmax(m, v);
        const bool sel = v > conf;          // strict, model.py:459
        flag[k] = sel ? 1 : 0;
        nb += sel;
      }
      assigned = m > conf;                   // model.py:463-465 on the stored row maxima
      ambiguous = (m != 0.0) && !(m > conf);
      unaligned = m == 0.0;
    } else if (mode == SS_INITIAL_UNIQUE) {
      for (int k = b; k < e; ++k) flag[k] = len == 1 ? 1 : 0;
      nb = len == 1;
      assigned = len == 1;
      ambiguous = len > 1;
      unaligned = len == 0;
    } else if (mode == SS_INITIAL_RANDOM) {
      int m = 0;
      for (int k = b; k < e; ++k) m = max(m, (int)score[k]);
      if (m > 0)
        for (int k = b; k < e; ++k) nb += ((int)score[k] == m);
      for (int k = b; k < e; ++k) flag[k] = (m > 0 && (int)score[k] == m) ? 1 : 0;
      assigned = nb == 1;
      ambiguous = nb > 1;
      unaligned = nb == 0;
      atomicAdd(&s_hist[min(nb, HIST_BINS - 1)], 1u);
    } else {  // SS_TOTAL_HITS
      for (int k = b; k < e; ++k) flag[k] = 1;
      nb = len;
      assigned = (unsigned long long)len;    // total number of alignments, model.py:498
      ambiguous = len > 1;
      unaligned = len == 0;
    }
    nbest_out[row] = nb;
  }
  // block-level aggregation of the counters
  for (int o = 16; o > 0; o >>= 1) {
    assigned += __shfl_xor_sync(0xffffffffu, assigned, o);
    ambiguous += __shfl_xor_sync(0xffffffffu, ambiguous, o);
    unaligned += __shfl_xor_sync(0xffffffffu, unaligned, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_cnt[0], assigned);
    atomicAdd(&s_cnt[1], ambiguous);
    atomicAdd(&s_cnt[2], unaligned);
  }
  __syncthreads();
  if (threadIdx.x < 3 && s_cnt[threadIdx.x]) atomicAdd(&info[threadIdx.x], s_cnt[threadIdx.x]);
  for (int i = threadIdx.x; i < HIST_BINS; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&info[4 + i], (unsigned long long)s_hist[i]);
}

int reassign(ss_handle_s* h, cudaStream_t st, int mode, double conf, double tol, int n_rows, const int* row_ptr,
             const uint16_t* score, const double* z, uint8_t* flag, int* nbest, long long* info) {
  if (mode < 0 || mode > SS_TOTAL_HITS) {
    h->set_error("ss_reassign: unknown mode");   // ValueError, model.py:424-426
    return SS_ERR_ARG;
  }
  if (mode == SS_BEST_CONF && !(conf > 0.5)) {
    h->set_error("ss_reassign: confidence threshold must be > 0.5");  // model.py:454-457
    return SS_ERR_MODEL;
  }
  const bool needs_z = mode <= SS_BEST_AVERAGE;
  if (!info || (n_rows > 0 && ((needs_z && !z) || ((mode == SS_INITIAL_RANDOM) && !score) || !flag || !nbest ||
                               !row_ptr))) {       // n_rows == 0: a rank of a sharded run without rows of its own
    h->set_error("ss_reassign: missing buffer");
    return SS_ERR_ARG;
  }
  SS_CUDA_CHECK(h, cudaMemsetAsync(info, 0, (4 + HIST_BINS) * sizeof(long long), st));
  if (n_rows > 0) {
    reassign_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(mode, conf, tol, n_rows, row_ptr, score, z, flag, nbest,
                                                          reinterpret_cast<unsigned long long*>(info));
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
  }
  return SS_OK;
}

// ---------------------------------------------------------------------------
// count aggregation
// ---------------------------------------------------------------------------
__global__ void count_rows_kernel(int n_rows, const int* __restrict__ row_ptr, const uint8_t* __restrict__ flag,
                                  const int* __restrict__ row_cell, int* __restrict__ per_row) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  int c = 0;
  if (row_cell[row] >= 0)
    for (int k = row_ptr[row]; k < row_ptr[row + 1]; ++k) c += flag[k] != 0;
  per_row[row] = c;
}

__global__ void count_fill_kernel(int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                  const uint8_t* __restrict__ flag, const double* __restrict__ weight,
                                  const int* __restrict__ row_cell, const int* __restrict__ row_umi,
                                  const long long* __restrict__ offs, unsigned long long* __restrict__ key_a,
                                  unsigned int* __restrict__ key_col, double* __restrict__ val) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const int cell = row_cell[row];
  if (cell < 0) return;
  long long o = offs[row];
  for (int k = row_ptr[row]; k < row_ptr[row + 1]; ++k) {
    if (!flag[k]) continue;
    if (row_umi) {
      key_a[o] = ((unsigned long long)(unsigned)cell << 32) | (unsigned)row_umi[row];
      key_col[o] = (unsigned)col_idx[k];
    } else {
      key_a[o] = ((unsigned long long)(unsigned)cell << 32) | (unsigned)col_idx[k];
    }
    if (val) val[o] = weight ? weight[k] : 1.0;
    ++o;
  }
}

__global__ void widen_kernel(long long n, const int* __restrict__ in, long long* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// (cell|umi, col) sorted pairs -> head flag of each distinct triple, re-keyed as (cell|col)
__global__ void umi_heads_kernel(long long n, const unsigned long long* __restrict__ key_a,
                                 const unsigned int* __restrict__ key_col, unsigned long long* __restrict__ out_key,
                                 uint8_t* __restrict__ head) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool hd = i == 0 || key_a[i] != key_a[i - 1] || key_col[i] != key_col[i - 1];
  head[i] = hd;
  out_key[i] = (key_a[i] & 0xffffffff00000000ull) | key_col[i];
}

__global__ void unpack_kernel(long long n, const unsigned long long* __restrict__ keys, const int* __restrict__ cnt,
                              const double* __restrict__ sums, int* __restrict__ out_cell, int* __restrict__ out_col,
                              double* __restrict__ out_val) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out_cell[i] = (int)(keys[i] >> 32);
  out_col[i] = (int)(keys[i] & 0xffffffffu);
  out_val[i] = sums ? sums[i] : (double)cnt[i];
}

struct TmpPool {       // temporaries of one call: allocated (dev_alloc) and released in the order of the call's stream
  std::vector<void*> p;
  ~TmpPool() {
    for (void* q : p) cudaFreeAsync(q, tl_alloc_stream);
  }
  template <class T>
  cudaError_t get(T** out, size_t n) {
    return dev_alloc(p, out, n);
  }
};

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

int count(ss_handle_s* h, cudaStream_t st, int n_rows, int n_cols, const int* row_ptr, const int* col_idx,
          const uint8_t* flag, const double* weight, const int* row_cell, const int* row_umi, long long cap,
          int* out_cell, int* out_col, double* out_val, long long* n_out) {
  (void)n_cols;
  if (!row_ptr || !col_idx || !flag || !row_cell || !n_out) {
    h->set_error("ss_count: missing buffer");
    return SS_ERR_ARG;
  }
  *n_out = 0;
  if (n_rows <= 0) return SS_OK;
  TmpPool tmp;
  const int B = 256, G = (n_rows + B - 1) / B;
  int* per_row;
  long long *per_row64, *offs;
  SS_CUDA_CHECK(h, tmp.get(&per_row, (size_t)n_rows));
  SS_CUDA_CHECK(h, tmp.get(&per_row64, (size_t)n_rows));
  SS_CUDA_CHECK(h, tmp.get(&offs, (size_t)n_rows + 1));
  count_rows_kernel<<<G, B, 0, st>>>(n_rows, row_ptr, flag, row_cell, per_row);
  widen_kernel<<<G, B, 0, st>>>(n_rows, per_row, per_row64);
  h->launches += 2;
  CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, per_row64, offs, n_rows, st));
  long long last_off = 0, last_cnt = 0;
  SS_CUDA_CHECK(h, cudaMemcpyAsync(&last_off, offs + (n_rows - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaMemcpyAsync(&last_cnt, per_row64 + (n_rows - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  const long long M = last_off + last_cnt;
  if (M == 0) return SS_OK;
  if (M > 0x7fffffffLL) {
    h->set_error("ss_count: more than 2^31 selected alignments in one call");
    return SS_ERR_CAPACITY;
  }
  const int m = (int)M;
  const int GM = (m + B - 1) / B;
  unsigned long long *key_a, *key_b;
  unsigned int *kc_a = nullptr, *kc_b = nullptr;
  double *val_a = nullptr, *val_b = nullptr;
  SS_CUDA_CHECK(h, tmp.get(&key_a, (size_t)m));
  SS_CUDA_CHECK(h, tmp.get(&key_b, (size_t)m));
  const bool umi = row_umi != nullptr;
  const bool weighted = weight != nullptr && !umi;
  if (umi) {
    SS_CUDA_CHECK(h, tmp.get(&kc_a, (size_t)m));
    SS_CUDA_CHECK(h, tmp.get(&kc_b, (size_t)m));
  }
  if (weighted) {
    SS_CUDA_CHECK(h, tmp.get(&val_a, (size_t)m));
    SS_CUDA_CHECK(h, tmp.get(&val_b, (size_t)m));
  }
  count_fill_kernel<<<G, B, 0, st>>>(n_rows, row_ptr, col_idx, flag, weight, row_cell, row_umi, offs, key_a, kc_a,
                                     val_a);
  h->launches++;
  SS_CUDA_CHECK(h, cudaGetLastError());

  unsigned long long* keys_sorted = key_b;
  long long m2 = m;
  if (umi) {
    // sort by (cell|umi, col): LSD -- first the column, then (stable) the 64-bit key
    CUB_CALL(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, kc_a, kc_b, key_a, key_b, m, 0, 32, st));
    CUB_CALL(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, key_b, key_a, kc_b, kc_a, m, 0, 64, st));
    // distinct (cell, umi, col) triples -> one count each (model.py:1340-1341), re-keyed (cell|col)
    uint8_t* head;
    int* d_nsel;
    SS_CUDA_CHECK(h, tmp.get(&head, (size_t)m));
    SS_CUDA_CHECK(h, tmp.get(&d_nsel, 1));
    umi_heads_kernel<<<GM, B, 0, st>>>(m, key_a, kc_a, key_b, head);
    h->launches++;
    unsigned long long* sel;
    SS_CUDA_CHECK(h, tmp.get(&sel, (size_t)m));
    CUB_CALL(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, key_b, head, sel, d_nsel, m, st));
    int nsel = 0;
    SS_CUDA_CHECK(h, cudaMemcpyAsync(&nsel, d_nsel, sizeof(int), cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
    m2 = nsel;
    CUB_CALL(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, sel, key_a, (int)m2, 0, 64, st));
    keys_sorted = key_a;
  } else if (weighted) {
    CUB_CALL(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, key_a, key_b, val_a, val_b, m, 0, 64, st));
  } else {
    CUB_CALL(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, key_a, key_b, m, 0, 64, st));
  }
  // run-length encode / reduce by key
  unsigned long long* uniq;
  int* cnts = nullptr;
  double* sums = nullptr;
  int* d_nruns;
  SS_CUDA_CHECK(h, tmp.get(&uniq, (size_t)m2));
  SS_CUDA_CHECK(h, tmp.get(&d_nruns, 1));
  if (weighted) {
    SS_CUDA_CHECK(h, tmp.get(&sums, (size_t)m2));
    CUB_CALL(cub::DeviceReduce::ReduceByKey(d_tmp, tmp_bytes, keys_sorted, uniq, val_b, sums, d_nruns, cub::Sum(),
                                            (int)m2, st));
  } else {
    SS_CUDA_CHECK(h, tmp.get(&cnts, (size_t)m2));
    CUB_CALL(cub::DeviceRunLengthEncode::Encode(d_tmp, tmp_bytes, keys_sorted, uniq, cnts, d_nruns, (int)m2, st));
  }
  int nruns = 0;
  SS_CUDA_CHECK(h, cudaMemcpyAsync(&nruns, d_nruns, sizeof(int), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  *n_out = nruns;
  if (nruns > cap) {
    h->set_error("ss_count: output capacity too small");
    return SS_ERR_CAPACITY;
  }
  if (nruns > 0) {
    unpack_kernel<<<(nruns + B - 1) / B, B, 0, st>>>(nruns, uniq, cnts, sums, out_cell, out_col, out_val);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));  // temporaries are freed on return
  }
  return SS_OK;
}

// ---------------------------------------------------------------------------
// choose_random(1, rng) (sparse_plus.py:199-240): in every row with several tied-best alignments keep the one the
// host drew.  The draws stay on the host (numpy's PCG64 stream, in row order -- INTEGRATION.md section 3); only the
// tied rows and their picks travel.  One thread per tied row.
// ---------------------------------------------------------------------------
__global__ void pick_kernel(int n_tied, const int* __restrict__ tied_rows, const int* __restrict__ picks,
                            const int* __restrict__ row_ptr, uint8_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tied) return;
  const int row = tied_rows[i], want = picks[i];
  int seen = 0;
  for (int k = row_ptr[row]; k < row_ptr[row + 1]; ++k) {
    if (!flag[k]) continue;
    if (seen != want) flag[k] = 0;
    ++seen;
  }
}

int reassign_pick(ss_handle_s* h, cudaStream_t st, int n_tied, const int* tied_rows, const int* picks, const int* row_ptr,
                  uint8_t* flag) {
  if (n_tied < 0 || (n_tied > 0 && (!tied_rows || !picks || !row_ptr || !flag))) {
    h->set_error("ss_reassign_pick: bad argument");
    return SS_ERR_ARG;
  }
  if (n_tied > 0) {
    pick_kernel<<<(n_tied + 255) / 256, 256, 0, st>>>(n_tied, tied_rows, picks, row_ptr, flag);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
  }
  return SS_OK;
}

// ---------------------------------------------------------------------------
// the reassignment matrix as CSR: compaction of the flagged alignments
// ---------------------------------------------------------------------------
__global__ void select_count_kernel(int n_rows, const int* __restrict__ row_ptr, const uint8_t* __restrict__ flag,
                                    long long* __restrict__ per_row) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row > n_rows) return;
  int c = 0;
  if (row < n_rows && flag)
    for (int k = row_ptr[row]; k < row_ptr[row + 1]; ++k) c += flag[k] != 0;
  per_row[row] = c;                                  // per_row[n_rows] = 0: the scan's last element is the total
}

__global__ void select_fill_kernel(int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                   const uint8_t* __restrict__ flag, const int* __restrict__ nbest,
                                   const long long* __restrict__ out_ptr, int* __restrict__ out_idx,
                                   double* __restrict__ out_val) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  long long o = out_ptr[row];
  const double v = (out_val && nbest && nbest[row] > 0) ? 1.0 / (double)nbest[row] : 1.0;   // best / nbest, model.py:448
  for (int k = row_ptr[row]; k < row_ptr[row + 1]; ++k) {
    if (!flag[k]) continue;
    out_idx[o] = col_idx[k];
    if (out_val) out_val[o] = v;
    ++o;
  }
}

int reassign_csr(ss_handle_s* h, cudaStream_t st, int n_rows, const int* row_ptr, const int* col_idx, const uint8_t* flag,
                 const int* nbest, long long cap, long long* out_ptr, int* out_idx, double* out_val, long long* n_out) {
  *n_out = 0;
  // a matrix without alignments (a rank of a sharded run without rows of its own) has empty col_idx / flag arrays
  if (n_rows < 0 || !row_ptr || !out_ptr || (cap > 0 && !out_idx)) {
    h->set_error("ss_reassign_csr: bad argument");
    return SS_ERR_ARG;
  }
  TmpPool tmp;
  const int B = 256;
  long long* per_row;
  SS_CUDA_CHECK(h, tmp.get(&per_row, (size_t)n_rows + 1));
  select_count_kernel<<<(n_rows + 1 + B - 1) / B, B, 0, st>>>(n_rows, row_ptr, flag, per_row);
  h->launches++;
  CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, per_row, out_ptr, n_rows + 1, st));
  long long total = 0;
  SS_CUDA_CHECK(h, cudaMemcpyAsync(&total, out_ptr + n_rows, sizeof(long long), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  *n_out = total;
  if (total > cap) {
    h->set_error("ss_reassign_csr: output capacity too small");
    return SS_ERR_CAPACITY;
  }
  if (total > 0 && !col_idx) {
    h->set_error("ss_reassign_csr: bad argument");
    return SS_ERR_ARG;
  }
  if (total > 0 && n_rows > 0) {
    select_fill_kernel<<<(n_rows + B - 1) / B, B, 0, st>>>(n_rows, row_ptr, col_idx, flag, nbest, out_ptr, out_idx, out_val);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
  }
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));      // temporaries are freed on return
  return SS_OK;
}

}  // namespace ss
