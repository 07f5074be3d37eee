// Score-matrix construction on the device -- the producer of the path's input.
//
// Reference being replaced: mapping_to_matrix inside Stellarscope.load_alignment
// (stellarscope/utils/model.py:958-978): one Python dok_matrix insert per (fragment, locus) mapping,
//     m[i, j] = max(m[i, j], (alnscore - minAS + 1) + alnlen)          (uint16)
// then dok -> CSR and the check that every read has a stored value in a feature column >= 1.
//
// Device formulation: one 64-bit key (read << 32 | feature) per mapping -> radix sort of (key, score)
// over the bits in use -> run heads carry the maximum of their run -> compaction of the heads into
// col_idx / score; row_ptr by binary search of the row boundaries in the sorted keys.  The output is
// canonical CSR (columns ascending, no duplicates) -- what coo.tocsr() + sum_duplicates leaves in the
// reference and what ss_layout_build wants.
#include <cub/cub.cuh>

#include "ss_handle.h"

namespace ss {

namespace {

struct ScTmp {
  cudaStream_t st;
  std::vector<void*> p;
  explicit ScTmp(cudaStream_t s) : st(s) {}
  ~ScTmp() {
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

#define SC_CUB(call_with_tmp)                  \
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

inline unsigned sgrid(long long n, int b = 256) { return (unsigned)std::max<long long>(1, (n + b - 1) / b); }

// info[1]: mappings whose value leaves (0, 65535]; info[2]: mappings with a read / feature id out of range
__global__ void sc_keys_kernel(long long n, const int* __restrict__ read, const int* __restrict__ feat,
                               const int* __restrict__ alnscore, const int* __restrict__ alnlen, int min_as, int n_reads,
                               int n_feats, unsigned long long* __restrict__ keys, uint16_t* __restrict__ vals,
                               unsigned long long* __restrict__ info) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = read[i], f = feat[i];
  const long long v = ((long long)alnscore[i] - min_as + 1) + alnlen[i];      // rescale(alnscore) + alnlen, model.py:962,969
  const bool bad_id = r < 0 || r >= n_reads || f < 0 || f >= n_feats;
  const bool bad_v = v <= 0 || v > 65535;
  if (bad_id) atomicAdd(&info[2], 1ull);
  if (bad_v) atomicAdd(&info[1], 1ull);
  keys[i] = bad_id ? ~0ull : ((unsigned long long)(unsigned)r << 32 | (unsigned)f);
  vals[i] = bad_v ? (uint16_t)0 : (uint16_t)v;
}

// sorted keys: a run of equal keys is one stored element; its head looks ahead for the maximum
// (runs are the alignments of one fragment on one locus -- a handful)
__global__ void sc_heads_kernel(long long n, const unsigned long long* __restrict__ keys, const uint16_t* __restrict__ vals,
                                uint8_t* __restrict__ head, uint16_t* __restrict__ vmax) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i];
  const bool hd = i == 0 || keys[i - 1] != k;
  head[i] = hd;
  if (!hd) return;
  uint16_t m = vals[i];
  for (long long j = i + 1; j < n && keys[j] == k; ++j) m = max(m, vals[j]);
  vmax[i] = m;
}

struct ScItem {
  unsigned long long key;
  uint16_t v;
};

__global__ void sc_pack_kernel(long long n, const unsigned long long* __restrict__ keys, const uint16_t* __restrict__ vmax,
                               ScItem* __restrict__ items) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ScItem it;
  it.key = keys[i];
  it.v = vmax[i];
  items[i] = it;
}

__global__ void sc_emit_kernel(const long long* __restrict__ d_nnz, const ScItem* __restrict__ items, int* __restrict__ col_idx,
                               uint16_t* __restrict__ score, unsigned long long* __restrict__ ukeys) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *d_nnz) return;
  const ScItem it = items[i];
  col_idx[i] = (int)(unsigned)(it.key & 0xffffffffull);
  score[i] = it.v;
  ukeys[i] = it.key;
}

// row_ptr[r] = first stored element with read >= r; rows whose last stored column is 0 (or that are empty)
// have no feature column -- the reference's assertion, model.py:971-974
__global__ void sc_rowptr_kernel(int n_reads, const long long* __restrict__ d_nnz, const unsigned long long* __restrict__ ukeys,
                                 int* __restrict__ row_ptr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_reads) return;
  const long long nnz = *d_nnz;
  const unsigned long long probe = (unsigned long long)(unsigned)r << 32;
  long long lo = 0, hi = nnz;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (ukeys[mid] < probe) lo = mid + 1; else hi = mid;
  }
  row_ptr[r] = (int)lo;
}

__global__ void sc_check_kernel(int n_reads, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                unsigned long long* __restrict__ info) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const int a = row_ptr[r], b = row_ptr[r + 1];
  if (b == a || col_idx[b - 1] < 1) atomicAdd(&info[3], 1ull);
}

}  // namespace

int scores_build(ss_handle_s* h, cudaStream_t st, long long n, const int* read, const int* feat, const int* alnscore,
                 const int* alnlen, int min_as, int n_reads, int n_feats, int* row_ptr, int* col_idx, uint16_t* score,
                 long long* info_host) {
  ScTmp tmp(st);
  for (int i = 0; i < 4; ++i) info_host[i] = 0;
  if (n_reads < 0 || n_feats < 0 || n < 0) {
    h->set_error("ss_scores_build: negative size");
    return SS_ERR_ARG;
  }
  if (n > 0x7fffffffLL) {
    h->set_error("ss_scores_build: more than 2^31-1 mappings");
    return SS_ERR_CAPACITY;
  }
  SS_CUDA_CHECK(h, cudaMemsetAsync(row_ptr, 0, ((size_t)n_reads + 1) * sizeof(int), st));
  unsigned long long* info;
  SS_CUDA_CHECK(h, tmp.get(&info, 4));
  SS_CUDA_CHECK(h, cudaMemsetAsync(info, 0, 4 * sizeof(unsigned long long), st));
  if (n > 0) {
    unsigned long long *keys, *keys_s, *ukeys;
    uint16_t *vals, *vals_s, *vmax;
    uint8_t* head;
    ScItem *items, *items_sel;
    long long* d_nnz;
    SS_CUDA_CHECK(h, tmp.get(&keys, (size_t)n));
    SS_CUDA_CHECK(h, tmp.get(&keys_s, (size_t)n));
    SS_CUDA_CHECK(h, tmp.get(&ukeys, (size_t)n));
    SS_CUDA_CHECK(h, tmp.get(&vals, (size_t)n));
    SS_CUDA_CHECK(h, tmp.get(&vals_s, (size_t)n));
    SS_CUDA_CHECK(h, tmp.get(&vmax, (size_t)n));
    SS_CUDA_CHECK(h, tmp.get(&head, (size_t)n));
    SS_CUDA_CHECK(h, tmp.get(&items, (size_t)n));
    SS_CUDA_CHECK(h, tmp.get(&items_sel, (size_t)n));
    SS_CUDA_CHECK(h, tmp.get(&d_nnz, 1));
    sc_keys_kernel<<<sgrid(n), 256, 0, st>>>(n, read, feat, alnscore, alnlen, min_as, n_reads, n_feats, keys, vals, info);
    h->launches++;
    // stop on bad input before anything is written to the outputs
    unsigned long long hinfo[4];
    SS_CUDA_CHECK(h, cudaMemcpyAsync(hinfo, info, sizeof(hinfo), cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
    info_host[1] = (long long)hinfo[1];
    info_host[2] = (long long)hinfo[2];
    if (hinfo[2]) {
      h->set_error("ss_scores_build: read / feature index outside the matrix");
      return SS_ERR_ARG;
    }
    if (hinfo[1]) {
      h->set_error("ss_scores_build: score outside (0, 65535] (uint16 score matrix, model.py:964)");
      return SS_ERR_CAPACITY;
    }
    int bits = 32;
    while (bits < 64 && (1LL << (bits - 32)) < (long long)n_reads) ++bits;
    SC_CUB(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, keys, keys_s, vals, vals_s, (int)n, 0, bits, st));
    sc_heads_kernel<<<sgrid(n), 256, 0, st>>>(n, keys_s, vals_s, head, vmax);
    sc_pack_kernel<<<sgrid(n), 256, 0, st>>>(n, keys_s, vmax, items);
    h->launches += 2;
    SC_CUB(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, items, head, items_sel, d_nnz, (int)n, st));
    sc_emit_kernel<<<sgrid(n), 256, 0, st>>>(d_nnz, items_sel, col_idx, score, ukeys);
    sc_rowptr_kernel<<<sgrid((long long)n_reads + 1), 256, 0, st>>>(n_reads, d_nnz, ukeys, row_ptr);
    h->launches += 2;
    long long nnz = 0;
    SS_CUDA_CHECK(h, cudaMemcpyAsync(&nnz, d_nnz, sizeof(long long), cudaMemcpyDeviceToHost, st));
    if (n_reads > 0) {
      sc_check_kernel<<<sgrid(n_reads), 256, 0, st>>>(n_reads, row_ptr, col_idx, info);
      h->launches++;
    }
    SS_CUDA_CHECK(h, cudaMemcpyAsync(hinfo, info, sizeof(hinfo), cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaGetLastError());
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
    info_host[0] = nnz;
    info_host[3] = (long long)hinfo[3];
  } else {
    info_host[3] = n_reads;        // no mapping at all: every read fails the feature-column check
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  }
  return SS_OK;
}

}  // namespace ss
