// MatrixMarket text of a count matrix, formatted on the device.
//
// Reference being replaced: the body written by scipy.io.mmwrite at the end of
// Stellarscope.output_report (utils/model.py:1418): one line "row col value" (1-based) per stored
// entry, in storage order.  At atlas scale that writer is the slowest step left of the drop-in flow
// (8-20 M entries/s on the host: 10-25 s per reassignment mode for configs[4]); here the COO triples
// are turned into text by two HBM-bound passes (lengths -> exclusive scan -> bytes), and the host
// only streams the finished bytes into the file.
//   pass 1  mtx_len_kernel    16 B read + 8 B written per entry
//   scan    cub::DeviceScan   16 B per entry
//   pass 2  mtx_write_kernel  24 B read per entry, text written once: every CTA formats its 256 lines
//                             into shared memory at the byte offsets of the scan and copies the block
//                             out with 16-byte stores (the lines have odd lengths; byte-wise global
//                             stores would touch every 32-byte sector ~3 times)
#include <cub/cub.cuh>

#include "ss_handle.h"
#include "ss_mtx_fmt.cuh"

namespace ss {

namespace {

constexpr int MTX_B = 256;   // entries per CTA

struct MtxTmp {              // temporaries of one call, stream-ordered
  std::vector<void*> p;
  ~MtxTmp() {
    for (void* q : p) cudaFreeAsync(q, tl_alloc_stream);
  }
  template <class T>
  cudaError_t get(T** out, size_t n) { return dev_alloc(p, out, n); }
};

__global__ void mtx_len_kernel(long long n, const int* __restrict__ row, const int* __restrict__ col,
                               const double* __restrict__ val, int integer, long long* __restrict__ len,
                               int* __restrict__ bad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { len[n] = 0; return; }
  const double v = val[i];
  const int r = row[i], c = col[i];
  bool ok = r >= 0 && c >= 0;
  if (integer) ok = ok && fabs(v) < 9007199254740992.0 && v == (double)(long long)v;
  else ok = ok && ssmtx::real_supported(v);
  if (!ok) {
    atomicExch(bad, 1);
    len[i] = 0;
    return;
  }
  len[i] = ssmtx::line_len(r, c, v, integer != 0);
}

__global__ void __launch_bounds__(MTX_B) mtx_write_kernel(long long n, const int* __restrict__ row,
                                                          const int* __restrict__ col, const double* __restrict__ val,
                                                          int integer, const long long* __restrict__ off,
                                                          unsigned char* __restrict__ out) {
  __shared__ __align__(16) char buf[MTX_B * ssmtx::MAX_LINE_CHARS + 32];
  const long long i0 = (long long)blockIdx.x * MTX_B;
  const long long i1 = min(i0 + MTX_B, n);
  const long long base = off[i0], end = off[i1];
  const int pad = (int)(base & 15);                 // shared and global offsets agree modulo 16
  const long long i = i0 + threadIdx.x;
  if (i < i1) ssmtx::put_line(buf + pad + (int)(off[i] - base), row[i], col[i], val[i], integer != 0);
  __syncthreads();
  const int span = pad + (int)(end - base);         // bytes [pad, span) of buf are text
  unsigned char* const g = out + (base - pad);      // 16-byte aligned (out is)
  for (int w = threadIdx.x; w * 16 < span; w += MTX_B) {
    const int b0 = w * 16;
    if (b0 >= pad && b0 + 16 <= span) {
      *reinterpret_cast<uint4*>(g + b0) = *reinterpret_cast<const uint4*>(buf + b0);
    } else {
      const int lo = max(b0, pad), hi = min(b0 + 16, span);
      for (int b = lo; b < hi; ++b) g[b] = (unsigned char)buf[b];
    }
  }
}

}  // namespace

int mtx_measure(ss_handle_s* h, cudaStream_t st, long long n, const int* row, const int* col, const double* val,
                int integer, long long* off, long long* n_bytes) {
  *n_bytes = 0;
  if (n < 0 || !off || (n > 0 && (!row || !col || !val))) {
    h->set_error("ss_mtx_measure: bad argument");
    return SS_ERR_ARG;
  }
  MtxTmp tmp;
  int* bad;
  SS_CUDA_CHECK(h, tmp.get(&bad, 1));
  SS_CUDA_CHECK(h, cudaMemsetAsync(bad, 0, sizeof(int), st));
  const unsigned grid = (unsigned)((n + 1 + MTX_B - 1) / MTX_B);
  mtx_len_kernel<<<grid, MTX_B, 0, st>>>(n, row, col, val, integer, off, bad);
  h->launches++;
  SS_CUDA_CHECK(h, cudaGetLastError());
  {
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    SS_CUDA_CHECK(h, cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, off, off, n + 1, st));
    unsigned char* t;
    SS_CUDA_CHECK(h, tmp.get(&t, tmp_bytes));
    d_tmp = t;
    SS_CUDA_CHECK(h, cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, off, off, n + 1, st));   // in place
    h->launches++;
  }
  long long total = 0;
  int hbad = 0;
  SS_CUDA_CHECK(h, cudaMemcpyAsync(&total, off + n, sizeof(long long), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  if (hbad) {
    h->set_error(integer ? "ss_mtx_measure: negative index or a value that is not an integer below 2^53"
                         : "ss_mtx_measure: negative index or a value outside 0, 2^-64 <= |x| < 2^63");
    return SS_ERR_ARG;
  }
  *n_bytes = total;
  return SS_OK;
}

int mtx_write(ss_handle_s* h, cudaStream_t st, long long n, const int* row, const int* col, const double* val,
              int integer, const long long* off, unsigned char* out) {
  if (n < 0 || !off || (n > 0 && (!row || !col || !val || !out))) {
    h->set_error("ss_mtx_write: bad argument");
    return SS_ERR_ARG;
  }
  if (reinterpret_cast<uintptr_t>(out) & 15) {
    h->set_error("ss_mtx_write: the text buffer must be 16-byte aligned");
    return SS_ERR_ARG;
  }
  if (n == 0) return SS_OK;
  const unsigned grid = (unsigned)((n + MTX_B - 1) / MTX_B);
  mtx_write_kernel<<<grid, MTX_B, 0, st>>>(n, row, col, val, integer, off, out);
  h->launches++;
  SS_CUDA_CHECK(h, cudaGetLastError());
  return SS_OK;
}

}  // namespace ss
