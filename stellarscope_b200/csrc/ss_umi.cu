// UMI de-duplication on the device -- the step right before the EM path.
//
// Reference being replaced: Stellarscope.dedup_umi (stellarscope/utils/model.py:1148-1219) and
// select_umi_representatives (:561-668): Python dict loops over every barcode+UMI pair, an O(n^2)
// dok_matrix adjacency per duplicate group and scipy connected_components.
//
// Reads (rows) with the same barcode+UMI are duplicates of one molecule.  Inside such a group reads
// that share a feature (column != 0) are connected; every connected component keeps the read with the
// largest (max/sum, -n_features, max, sum, row) and the other reads are excluded.
// Device formulation: one (group, feature) key per alignment of the duplicate groups -> radix sort ->
// equal neighbours are unions in a lock-free union-find over rows (smaller row = root, so the result
// does not depend on the order of the unions) -> merge sort of the duplicate rows by
// (root, ranking tuple) -> the last row of every root segment is the representative.
#include <cub/cub.cuh>

#include "ss_handle.h"

namespace ss {

namespace {

struct UmiTmp {
  cudaStream_t st;
  std::vector<void*> p;
  explicit UmiTmp(cudaStream_t s) : st(s) {}
  ~UmiTmp() {
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

#define UMI_CUB(call_with_tmp)                 \
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

inline unsigned ugrid(long long n, int b = 256) { return (unsigned)std::max<long long>(1, (n + b - 1) / b); }

__global__ void umi_group_size_kernel(int n_rows, const int* __restrict__ row_group, int n_groups, int* __restrict__ gsize) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int g = row_group[r];
  if (g >= 0 && g < n_groups) atomicAdd(&gsize[g], 1);
}

// rows of duplicate groups: number of (group, feature) keys = alignments on features != 0; ranking statistics
struct UmiRec {
  int root;       // smallest row of the connected component
  int row;
  double ratio;   // max / sum            model.py:596
  int nlen;       // -n_features          :597
  int mx, sm;     // max, sum             :598-599
};

__global__ void umi_row_stats_kernel(int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                     const uint16_t* __restrict__ score, const int* __restrict__ row_group, int n_groups,
                                     const int* __restrict__ gsize, long long* __restrict__ nkeys, uint8_t* __restrict__ is_dup,
                                     int* __restrict__ parent) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int g = row_group[r];
  const bool dup = g >= 0 && g < n_groups && gsize[g] > 1;
  int n = 0;
  if (dup)
    for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) n += col_idx[k] != 0;
  nkeys[r] = n;
  is_dup[r] = dup;
  parent[r] = r;
}

__global__ void umi_emit_keys_kernel(int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                     const int* __restrict__ row_group, const uint8_t* __restrict__ is_dup,
                                     const long long* __restrict__ key_off, unsigned long long* __restrict__ keys,
                                     int* __restrict__ vals) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows || !is_dup[r]) return;
  long long o = key_off[r];
  const unsigned long long g = (unsigned long long)(unsigned)row_group[r] << 32;
  for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
    const int c = col_idx[k];
    if (c != 0) {
      keys[o] = g | (unsigned)c;
      vals[o] = r;
      ++o;
    }
  }
}

__device__ __forceinline__ int uf_find(int* parent, int x) {
  int p = ((volatile int*)parent)[x];
  while (p != x) {
    const int gp = ((volatile int*)parent)[p];
    if (gp != p) atomicCAS(&parent[x], p, gp);      // path halving (only ever lowers a parent)
    x = p;
    p = gp;
  }
  return x;
}

// equal neighbouring keys = two reads of one group on the same feature -> one component
__global__ void umi_union_kernel(long long n_keys, const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                 int* __restrict__ parent) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (i >= n_keys || keys[i] != keys[i - 1]) return;
  int a = vals[i], b = vals[i - 1];
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) break;
    if (a < b) { const int t = a; a = b; b = t; }   // a > b: hang the larger root under the smaller
    if (atomicCAS(&parent[a], a, b) == a) break;
  }
}

__global__ void umi_records_kernel(int n_dup, const int* __restrict__ dup_rows, const int* __restrict__ row_ptr,
                                   const int* __restrict__ col_idx, const uint16_t* __restrict__ score, int* __restrict__ parent,
                                   UmiRec* __restrict__ rec, int* __restrict__ comp_root) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_dup) return;
  const int r = dup_rows[i];
  int mx = 0, sm = 0, n = 0;
  for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k)
    if (col_idx[k] != 0) {
      const int s = score[k];
      mx = max(mx, s);
      sm += s;
      ++n;
    }
  UmiRec q;
  q.root = uf_find(parent, r);
  q.row = r;
  q.ratio = sm > 0 ? (double)mx / (double)sm : 0.0;
  q.nlen = -n;
  q.mx = mx;
  q.sm = sm;
  rec[i] = q;
  comp_root[r] = q.root;
}

struct UmiLess {   // ascending (root, ranking tuple): the representative of a component is its LAST record
  __device__ bool operator()(const UmiRec& a, const UmiRec& b) const {
    if (a.root != b.root) return a.root < b.root;
    if (a.ratio != b.ratio) return a.ratio < b.ratio;
    if (a.nlen != b.nlen) return a.nlen < b.nlen;
    if (a.mx != b.mx) return a.mx < b.mx;
    if (a.sm != b.sm) return a.sm < b.sm;
    return a.row < b.row;
  }
};

__global__ void umi_mark_kernel(int n_dup, const UmiRec* __restrict__ rec, const int* __restrict__ row_group,
                                uint8_t* __restrict__ excluded, int* __restrict__ group_ncomp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_dup) return;
  const bool rep = i + 1 == n_dup || rec[i + 1].root != rec[i].root;
  excluded[rec[i].row] = rep ? 0 : 1;
  if (rep) atomicAdd(&group_ncomp[row_group[rec[i].row]], 1);
}

__global__ void umi_iota_kernel(int n, int* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

}  // namespace

int umi_dedup(ss_handle_s* h, cudaStream_t st, int n_rows, const int* row_ptr, const int* col_idx, const uint16_t* score,
              const int* row_group, int n_groups, uint8_t* excluded, int* comp_root, int* group_size, int* group_ncomp) {
  UmiTmp tmp(st);
  const size_t N = (size_t)std::max(n_rows, 1), G = (size_t)std::max(n_groups, 1);
  SS_CUDA_CHECK(h, cudaMemsetAsync(excluded, 0, N, st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(comp_root, 0xff, N * sizeof(int), st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(group_size, 0, G * sizeof(int), st));
  SS_CUDA_CHECK(h, cudaMemsetAsync(group_ncomp, 0, G * sizeof(int), st));
  if (n_rows == 0 || n_groups == 0) return SS_OK;
  int *parent, *iota, *dup_rows, *d_nsel;
  uint8_t* is_dup;
  long long *nkeys, *key_off;
  SS_CUDA_CHECK(h, tmp.get(&nkeys, N + 1));
  SS_CUDA_CHECK(h, tmp.get(&parent, N));
  SS_CUDA_CHECK(h, tmp.get(&iota, N));
  SS_CUDA_CHECK(h, tmp.get(&dup_rows, N));
  SS_CUDA_CHECK(h, tmp.get(&d_nsel, 2));
  SS_CUDA_CHECK(h, tmp.get(&is_dup, N));
  SS_CUDA_CHECK(h, tmp.get(&key_off, N + 1));
  umi_group_size_kernel<<<ugrid(n_rows), 256, 0, st>>>(n_rows, row_group, n_groups, group_size);
  umi_row_stats_kernel<<<ugrid(n_rows), 256, 0, st>>>(n_rows, row_ptr, col_idx, score, row_group, n_groups, group_size,
                                                      nkeys, is_dup, parent);
  umi_iota_kernel<<<ugrid(n_rows), 256, 0, st>>>(n_rows, iota);
  h->launches += 3;
  // offsets of the keys of every row; duplicate rows
  SS_CUDA_CHECK(h, cudaMemsetAsync(nkeys + n_rows, 0, sizeof(long long), st));
  UMI_CUB(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, nkeys, key_off, n_rows + 1, st));
  UMI_CUB(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, iota, is_dup, dup_rows, d_nsel, n_rows, st));
  long long n_keys = 0;
  int n_dup = 0;
  SS_CUDA_CHECK(h, cudaMemcpyAsync(&n_keys, key_off + n_rows, sizeof(long long), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaMemcpyAsync(&n_dup, d_nsel, sizeof(int), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  if (n_dup == 0) return SS_OK;
  if (n_keys > 0x7fffffffLL) {
    h->set_error("ss_umi_dedup: too many alignments in duplicate groups");
    return SS_ERR_CAPACITY;
  }
  if (n_keys > 0) {
    unsigned long long *keys, *keys_s;
    int *vals, *vals_s;
    SS_CUDA_CHECK(h, tmp.get(&keys, (size_t)n_keys));
    SS_CUDA_CHECK(h, tmp.get(&keys_s, (size_t)n_keys));
    SS_CUDA_CHECK(h, tmp.get(&vals, (size_t)n_keys));
    SS_CUDA_CHECK(h, tmp.get(&vals_s, (size_t)n_keys));
    umi_emit_keys_kernel<<<ugrid(n_rows), 256, 0, st>>>(n_rows, row_ptr, col_idx, row_group, is_dup, key_off, keys, vals);
    h->launches++;
    int bits = 33;
    while ((1LL << (bits - 32)) < n_groups && bits < 64) ++bits;
    UMI_CUB(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, keys, keys_s, vals, vals_s, (int)n_keys, 0, bits, st));
    umi_union_kernel<<<ugrid(n_keys), 256, 0, st>>>(n_keys, keys_s, vals_s, parent);
    h->launches++;
  }
  UmiRec* rec;
  SS_CUDA_CHECK(h, tmp.get(&rec, (size_t)n_dup));
  umi_records_kernel<<<ugrid(n_dup), 256, 0, st>>>(n_dup, dup_rows, row_ptr, col_idx, score, parent, rec, comp_root);
  h->launches++;
  UMI_CUB(cub::DeviceMergeSort::SortKeys(d_tmp, tmp_bytes, rec, n_dup, UmiLess(), st));
  umi_mark_kernel<<<ugrid(n_dup), 256, 0, st>>>(n_dup, rec, row_group, excluded, group_ncomp);
  h->launches++;
  SS_CUDA_CHECK(h, cudaGetLastError());
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));      // the temporaries are released in stream order after this
  return SS_OK;
}

}  // namespace ss
