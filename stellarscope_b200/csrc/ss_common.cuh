// stellarscope_b200 -- shared declarations for the sm_100a kernels.
//
// Segment  = one EM model (reference: one TelescopeLikelihood, utils/model.py:87).
// Tile     = <= E_HARD alignments of consecutive ambiguous rows of one segment,
//            rows in jagged-diagonal order, plus the column-major slot (cpos) of
//            every alignment.  See DESIGN.md "Data layout in HBM" and
//            stellarscope_b200/layout_host.py (format specification).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace ss {

constexpr int E_HARD = 4096;     // hard cap on alignments per tile (worst case fits 227 KB of shared memory)
#ifndef SS_E_STREAM
#define SS_E_STREAM 2048
#endif
constexpr int E_STREAM = SS_E_STREAM;   // tile size of segments that are streamed (more than STREAM_SEG_TILES large tiles):
                                 // small tiles let 2-3 CTAs share an SM, so one CTA's loads / atomics overlap another's compute
constexpr int STREAM_SEG_TILES = 8;
#ifndef SS_E_CLUSTER
#define SS_E_CLUSTER 3072
#endif
constexpr int E_CLUSTER = SS_E_CLUSTER; // target tile size of segments of 2..STREAM_SEG_TILES tiles (thread-block clusters): equal tiles
constexpr int ALIGN_ELEMS = 8;   // per-tile slices start at multiples of 8 elements (16 B for u16)
constexpr int LONG_COL = 32;     // tile columns with more rows than this are reduced by a warp
constexpr int HDR_WORDS = 24;    // int64 words per tile header

// tile header word indices.  H_RGT*: rows of the tile longer than 4/8/16/32 alignments, H_CGT*: tile
// columns with more than 4..128 rows -- the lane-group boundaries of the lane kernel (rows are sorted
// by length, columns by count, both descending).
enum { H_SEG = 0, H_NROWS, H_NENT, H_NCOL, H_MAXLEN, H_NLONG, H_ENT_OFF, H_ROW_OFF, H_COL_OFF, H_JDS_OFF,
       H_RFRAG, H_CFRAG, H_RGT4, H_RGT8, H_RGT16, H_RGT32, H_CGT4, H_CGT8, H_CGT16, H_CGT32, H_CGT64, H_CGT128 };

// Fragments (resident kernel): a row / column is cut into pieces of FRAG entries handled by
// one lane each; a row with f pieces occupies a lane group of size pow2ceil(f) (<= FRAG_MAXG_ROW),
// a column one of size <= FRAG_MAXG_COL; longer ones take the generic path.
constexpr int FRAG = 4;
constexpr int FRAG_MAXG_ROW = 8;    // rows up to 32 alignments
constexpr int FRAG_MAXG_COL = 32;   // columns up to 128 rows
__host__ __device__ inline int frag_group(int len) {      // lanes a row/column of `len` entries occupies
  const int f = (len + FRAG - 1) / FRAG;
  int g = 1;
  while (g < f) g <<= 1;
  return g;
}

__host__ __device__ inline int pad8(int n) { return (n + 7) & ~7; }

// per alignment of a tile: Q_ij (e_q) and a packed index word (e_idx) =
// tile-local column | column-major slot << 16
__host__ __device__ inline uint32_t pack_idx(uint32_t lcol, uint32_t cpos) { return lcol | (cpos << 16); }

// Device-resident layout (all pointers are device pointers owned by the handle).
struct Layout {
  int N = 0, K = 0, S = 0, n_tiles = 0, maxlen = 0, chunk = 0, n_uniq = 0;
  long long A = 0, n_segcols = 0;
  long long tot_ent = 0, tot_row = 0, tot_col = 0, tot_jds = 0;
  long long* tile_hdr = nullptr;
  // per alignment (tile order)
  double* e_q = nullptr;
  uint32_t* e_idx = nullptr;
  int* e_opos = nullptr;
  // per ambiguous row (tile order)
  uint16_t* r_len = nullptr;
  double* r_w = nullptr;
  int* r_orow = nullptr;
  // per tile column
  uint16_t* c_ptr = nullptr;
  int* c_scol = nullptr;
  double* c_pisum0 = nullptr;
  int* c_nuniq = nullptr;
  uint16_t* j_ptr = nullptr;
  // per segment
  int *seg_tile0 = nullptr, *seg_ntiles = nullptr, *seg_col0 = nullptr, *seg_ncol = nullptr;
  int *seg_smax = nullptr, *seg_nobs = nullptr, *seg_namb = nullptr, *seg_nuniq = nullptr;
  double *seg_wmax = nullptr, *seg_wtot = nullptr, *seg_wamb = nullptr, *seg_logq_uniq = nullptr;
  long long *seg_A = nullptr, *seg_Aamb = nullptr;
  // per segment column
  int* sc_gcol = nullptr;
  double* sc_pisum0 = nullptr;
  int* sc_nuniq = nullptr;
  // unique rows
  int *u_row = nullptr, *u_pos = nullptr;
};

// Per-fit parameters and state (device pointers owned by the handle).
struct FitParams {
  double eps;
  int max_iter;
  double pi_prior, theta_prior;
  int use_likelihood;
  int K;
};

struct FitState {
  // per segment constants derived from the priors (model.py:189-194)
  double *seg_thpw = nullptr, *seg_pipw = nullptr, *seg_inv_thd = nullptr, *seg_inv_pid = nullptr;
  double *seg_diff1_extra = nullptr;   // contribution of columns outside the tiles to diff in iteration 1
  double *seg_lnl_extra = nullptr;     // lnL of unique rows on columns outside the tiles + sum log Q of unique rows
  // per segment column state
  double *sc_pi = nullptr;             // current pi
  double *sc_pt[2] = {nullptr, nullptr};  // pi*theta, double buffered by iteration parity (streaming)
  double *sc_ptprev = nullptr;         // pi*theta used by the last E-step (for emit)
  double *sc_ptfin = nullptr;          // final pi*theta
  double *sc_theta = nullptr;          // final theta
  double *sc_T = nullptr;              // streaming accumulators
  // per segment outputs
  int *seg_iters = nullptr, *seg_flags = nullptr;  // flags: bit0 converged, bit1 reached_max, bit2 error
  double *seg_lnl = nullptr;
  double *seg_diffs = nullptr;         // [S * max_iter]
  double *seg_lnls = nullptr;          // [S * max_iter] (use_likelihood only)
  double *seg_diffacc = nullptr;       // streaming: per segment diff accumulator
  double *tile_lnl = nullptr;          // [n_tiles]
};

#define SS_CUDA_CHECK(h, expr)                                                        \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      (h)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));             \
      return SS_ERR_CUDA;                                                             \
    }                                                                                 \
  } while (0)

// ---- device helpers ----------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared (cp.async.bulk; SASS: UBLKCP). bytes % 16 == 0,
// both addresses 16 B aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace ss
