// Handle behind the C ABI (include/stellarscope_b200.h).
#pragma once
#include <string>
#include <vector>

#include "../../include/stellarscope_b200.h"
#include "ss_common.cuh"

// lane kernel (csrc/ss_em.cu): row / column fragments (of <= 4 alignments) held in registers per
// thread.  Single-tile segments are binned by the smallest CTA size whose lanes hold all their
// fragments (launch plan, ss_layout.cu); the last class (1024 threads, half the registers per
// thread) takes the tiles that need more lanes than 512 threads offer.  The -D knobs below exist
// for tuning runs (measured on B200, configs[1]: 128 registers beat 96 / 80 / 64, FR=2 FC=4 beat
// 1/2, 2/2 and 4/4; the class granularity is worth ~2%).
#ifndef SS_FR_V
#define SS_FR_V 2
#define SS_FC_V 4
#define SS_MAXREG_V 128
#endif
#ifndef SS_FR_BIG_V
#define SS_FR_BIG_V 2
#define SS_FC_BIG_V 2
#define SS_MAXREG_BIG_V 64
#define SS_BIG_THREADS_V 1024
#endif
#ifndef SS_CLS_SET
#define SS_CLS_SET 0
#endif
#if SS_CLS_SET == 0
#define SS_CLS_LIST 32, 64, 96, 128, 160, 192, 224, 256, 320, 384, 448, 512
#elif SS_CLS_SET == 1
#define SS_CLS_LIST 32, 64, 128, 256, 512
#else
#define SS_CLS_LIST 32, 64, 96, 128, 192, 256, 384, 512
#endif
constexpr int SS_FR = SS_FR_V, SS_FC = SS_FC_V, SS_MAXREG = SS_MAXREG_V;
constexpr int SS_FR_BIG = SS_FR_BIG_V, SS_FC_BIG = SS_FC_BIG_V, SS_MAXREG_BIG = SS_MAXREG_BIG_V;
constexpr int SS_CLS_THREADS[] = {SS_CLS_LIST, SS_BIG_THREADS_V};
constexpr int SS_NUM_CLASSES = (int)(sizeof(SS_CLS_THREADS) / sizeof(int));
constexpr int SS_NUM_AUX_STREAMS = SS_NUM_CLASSES + 1 + 7 + 7;   // size classes, streaming kernel, cluster sizes 2..8 (heavy, light)
// cluster lane kernel (csrc/ss_em.cu): CTA size, row / column fragment slots per lane, register cap.  Heavy: one CTA
// owns its SM; light: two CTAs per SM (tiles whose row fragments fit 2 slots and whose shared memory is at most half
// an SM's -- launch plan, ss_layout.cu)
#ifndef SS_CLANE_THREADS
#define SS_CLANE_THREADS 512
#define SS_CLANE_FR 4
#define SS_CLANE_FC 6
#define SS_CLANE_MAXREG 128
#endif
#ifndef SS_CLANE_LIGHT_FR
#define SS_CLANE_LIGHT_FR 2
#define SS_CLANE_LIGHT_FC 3
#define SS_CLANE_LIGHT_MAXREG 64
#endif
constexpr int CLANE_THREADS = SS_CLANE_THREADS, CLANE_FR = SS_CLANE_FR, CLANE_FC = SS_CLANE_FC;
constexpr int CLANE_LIGHT_THREADS = 512, CLANE_LIGHT_FR = SS_CLANE_LIGHT_FR, CLANE_LIGHT_FC = SS_CLANE_LIGHT_FC;
constexpr int SS_COL_CHUNK = 512;
constexpr int SS_DN_MAXW = 8;   // ranks of a row-sharded (dense) fit: the GPUs of one NVSwitch domain
constexpr int SS_DN_THREADS = 512;    // CTA size of the row-sharded kernels = columns per update chunk
constexpr int SS_DN_TAIL = 8;   // scalar slots behind the K columns of an exchange buffer ([0] = lnL part)

struct ss_handle_s {
  int device = 0;
  int num_sms = 148;
  size_t max_smem_optin = 0;
  std::string err;
  long long launches = 0;
  // experiment knobs (environment, read by ss_create): SS_SMALL_TILE_THR = segments with more than this many
  // E_HARD tiles are cut into E_STREAM tiles (default STREAM_SEG_TILES: only the streamed ones; 1: every multi-tile
  // segment); SS_CLUSTER_TILE = target tile size of the segments that run as thread-block clusters (cut into equal
  // tiles of about this many alignments, ss_build.cu seg_chunk); SS_CLANE_LIGHT=0 switches the two-CTAs-per-SM
  // variant of the cluster lane kernel off
  int small_tile_thr = ss::STREAM_SEG_TILES;
  int cluster_tile = ss::E_CLUSTER;
  bool clane_light = true;

  // ---- layout ----
  bool has_layout = false;
  ss::Layout L;
  std::vector<void*> layout_allocs;
  std::vector<long long> hdr_host;            // tile headers, HDR_WORDS per tile
  std::vector<int> seg_ntiles_host, seg_tile0_host;
  // launch plan
  int* cls_tiles[SS_NUM_CLASSES] = {};   // device tile lists (single-tile segments), by size class
  int cls_count[SS_NUM_CLASSES] = {};
  size_t cls_smem[SS_NUM_CLASSES] = {};  // dynamic shared memory of the class (largest tile)
  int* stream_tiles = nullptr;    // device: tiles of multi-tile segments
  int n_stream_tiles = 0;
  int* stream_segs = nullptr;     // device: multi-tile segment ids
  int n_stream_segs = 0;
  int* stream_col_seg = nullptr;  // device: for every column of a multi-tile segment, index into stream_segs
  int* stream_col_slot = nullptr; // device: its absolute slot
  long long n_stream_cols = 0;
  int* stream_chunk_seg = nullptr;   // device: column chunks (SS_COL_CHUNK slots) of the streamed segments: segment ...
  int* stream_chunk_j0 = nullptr;    // ... and first column (segment-relative)
  int n_stream_chunks = 0;
  // column chunks (SS_COL_CHUNK slots of one segment each): the per-column work of prepare / finalize runs one
  // warp per chunk, so a segment with many columns does not serialise on one warp; partial sums are combined
  // per segment in chunk order (deterministic)
  int n_col_chunks = 0;
  int* chunk_seg = nullptr;       // device: segment of the chunk
  int* chunk_j0 = nullptr;        // device: first column (segment-relative) of the chunk
  int* seg_chunk0 = nullptr;      // device: first chunk of the segment
  double* chunk_part = nullptr;   // device: [2][n_col_chunks] partial sums
  size_t stream_smem = 0;
  // segments of 2..8 tiles: one thread-block cluster per segment (DSMEM exchange)
  static constexpr int MAX_CLUSTER = 8;
  int* clu_segs[MAX_CLUSTER + 1] = {};     // device: segment ids by number of tiles
  int clu_count[MAX_CLUSTER + 1] = {};
  size_t clu_smem[MAX_CLUSTER + 1] = {};       // three-barrier kernel (likelihood mode)
  size_t clu_smem_lane[MAX_CLUSTER + 1] = {};  // cluster lane kernel
  // "light" clusters: every tile's row fragments fit 2 slots of a 512-thread CTA and the shared memory of a CTA is at
  // most half an SM's, so two CTAs (of different clusters) share an SM -- one cluster's barrier / L2 exchange
  // overlaps the other's row and column phases (em_clane_kernel<512, 2, 3, 64>)
  int* clu_segs_light[MAX_CLUSTER + 1] = {};
  int clu_count_light[MAX_CLUSTER + 1] = {};
  size_t clu_smem_light[MAX_CLUSTER + 1] = {};
  uint16_t* clu_xref = nullptr;            // device: [segment column][tile of the segment] -> tile column or 0xffff
  long long* seg_xref_off = nullptr;       // device: offset of a segment's block in clu_xref (-1: none)
  int* seg_xmax = nullptr;                 // device: receive slots per tile of the cluster lane kernel (padded sum over the segment's tile columns of holders - 1)
  int n_cluster_tiles = 0, n_cluster_segs = 0;
  int* all_tiles = nullptr;       // device: 0 .. n_tiles-1
  size_t all_smem = 0;            // smem of the largest tile (likelihood-mode carve)
  int geom_all[4] = {8, 8, 8, 8};     // padded maxima (pe, pr, pc, pj) over all tiles / over the streamed tiles
  int geom_stream[4] = {8, 8, 8, 8};
  int8_t* sc_inamb = nullptr;     // device: segment column is touched by a tile
  long long nnz_amb = 0, rows_amb = 0;

  // ---- fit ----
  bool has_fit = false;
  ss::FitState F;
  ss::FitParams P;
  std::vector<void*> fit_allocs;
  size_t fit_S = 0, fit_NC = 0, fit_MI = 0, fit_T = 0;
  bool fit_lnl = false;
  int* stream_ctl = nullptr;      // device control words of the streaming kernel

  // measurement aid: events around the kernel groups of the last fit
  bool timing = false;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ev_valid = false;
  // the EM kernels of one fit run concurrently on internal streams forked from the caller's stream
  cudaStream_t aux[SS_NUM_AUX_STREAMS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[SS_NUM_AUX_STREAMS] = {};
  bool aux_used[SS_NUM_AUX_STREAMS] = {};
  cudaEvent_t tev_s[SS_NUM_AUX_STREAMS] = {}, tev_e[SS_NUM_AUX_STREAMS] = {};  // timing (when enabled)

  // multi-GPU hook
  ss_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;

  // ---- row-sharded single-model fit (ss_dense.cu) ----
  bool dn_force = false;          // dense path on a single rank (tests)
  bool dn_peers = false;          // exchange buffers in peer-mapped symmetric memory
  bool dn_persistent = false;     // whole fit in one cooperative kernel (peers, or a single rank)
  int dn_world = 1, dn_rank = 0;
  void* dn_peer_base[SS_DN_MAXW] = {};
  size_t dn_peer_bytes = 0;
  unsigned long long dn_epoch = 0;
  bool dn_ready = false;
  std::vector<void*> dense_allocs;
  double* dn_stat = nullptr;      // [3K + tail]: pisum0 | nuniq | active | scalars, summed over ranks
  double dn_tail_host[SS_DN_TAIL] = {};
  int* dn_tgcol = nullptr;
  double *dn_pi = nullptr, *dn_theta = nullptr, *dn_x[2] = {nullptr, nullptr}, *dn_T[2] = {nullptr, nullptr};
  double *dn_part = nullptr, *dn_lnl = nullptr;
  int *dn_ctl = nullptr, *dn_counts = nullptr;
  int dn_nchunks = 0, dn_Kp = 0;
  bool dense_mode() const { return dn_force || dn_peers || (allreduce != nullptr && dn_world > 1); }

  void set_error(const std::string& m) { err = m; }
};

namespace ss {
// ss_em.cu
int em_fit(ss_handle_s* h, cudaStream_t st, const FitParams& P);
int em_emit_z(ss_handle_s* h, cudaStream_t st, double* z);
void free_fit(ss_handle_s* h);
cudaError_t ensure_aux_streams(ss_handle_s* h);   // internal streams + fork / join events of the handle
// ss_umi.cu
int reassign_csr(ss_handle_s* h, cudaStream_t st, int n_rows, const int* row_ptr, const int* col_idx, const uint8_t* flag,
                 const int* nbest, long long cap, long long* out_ptr, int* out_idx, double* out_val, long long* n_out);
int umi_dedup(ss_handle_s* h, cudaStream_t st, int n_rows, const int* row_ptr, const int* col_idx, const uint16_t* score,
              const int* row_group, int n_groups, uint8_t* excluded, int* comp_root, int* group_size, int* group_ncomp);
int scores_build(ss_handle_s* h, cudaStream_t st, long long n, const int* read, const int* feat, const int* alnscore,
                 const int* alnlen, int min_as, int n_reads, int n_feats, int* row_ptr, int* col_idx, uint16_t* score,
                 long long* info_host);
// ss_mtx.cu
int mtx_measure(ss_handle_s* h, cudaStream_t st, long long n, const int* row, const int* col, const double* val,
                int integer, long long* off, long long* n_bytes);
int mtx_write(ss_handle_s* h, cudaStream_t st, long long n, const int* row, const int* col, const double* val,
              int integer, const long long* off, unsigned char* out);
// ss_dense.cu
int em_fit_dense(ss_handle_s* h, cudaStream_t st, const FitParams& P);
int dense_allreduce_smax(ss_handle_s* h, cudaStream_t st);
void free_dense(ss_handle_s* h);
// ss_layout.cu
int layout_build(ss_handle_s* h, cudaStream_t st, int n_rows, int n_cols, long long nnz, const int* row_ptr,
                 const int* col_idx, const uint16_t* score, const int* row_segment, int n_segments);
int layout_import(ss_handle_s* h, cudaStream_t st, const ss_host_layout* HL);
int layout_plan(ss_handle_s* h, cudaStream_t st, const std::vector<int>& seg_ncol, const std::vector<int>& seg_col0);
void free_layout(ss_handle_s* h);
// ss_reassign.cu
int reassign(ss_handle_s* h, cudaStream_t st, int mode, double conf, double tol, int n_rows, const int* row_ptr,
             const uint16_t* score, const double* z, uint8_t* flag, int* nbest, long long* info);
int count(ss_handle_s* h, cudaStream_t st, int n_rows, int n_cols, const int* row_ptr, const int* col_idx,
          const uint8_t* flag, const double* weight, const int* row_cell, const int* row_umi, long long cap,
          int* out_cell, int* out_col, double* out_val, long long* n_out);

int reassign_pick(ss_handle_s* h, cudaStream_t st, int n_tied, const int* tied_rows, const int* picks, const int* row_ptr,
                  uint8_t* flag);
// ss_count.cu
int count_cells(ss_handle_s* h, cudaStream_t st, int n_rows, int n_cols, int n_cells, const int* row_ptr, const int* col_idx,
                const uint8_t* flag, const int* nbest, const int* row_cell, long long cap, long long* out_ptr, int* out_col,
                int* out_cnt, double* out_val, long long* n_out);

// Device memory of a handle comes from the device's stream-ordered pool (cudaMallocAsync; the
// pool keeps its memory, ss_create): a layout build makes ~50 allocations and cudaMalloc would
// cost more than the kernels.  Entry points set the stream the allocations are ordered on.
extern thread_local cudaStream_t tl_alloc_stream;
template <class T>
inline cudaError_t dev_alloc(std::vector<void*>& pool, T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(p), n * sizeof(T), tl_alloc_stream);
  if (e == cudaSuccess) pool.push_back(*p);
  return e;
}
// free everything in `pool`; the device is synchronised first (kernels of internal streams may
// still use the memory)
inline void dev_free_all(std::vector<void*>& pool) {
  if (pool.empty()) return;
  cudaDeviceSynchronize();
  for (void* p : pool) cudaFreeAsync(p, tl_alloc_stream);
  pool.clear();
}
}  // namespace ss
