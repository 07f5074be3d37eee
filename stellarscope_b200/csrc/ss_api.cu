// extern "C" boundary (include/stellarscope_b200.h).
#include <cstdlib>
#include <cstring>
#include <new>

#include "ss_handle.h"

using namespace ss;

#define SS_REQUIRE_HANDLE(h) \
  if (!(h)) return SS_ERR_ARG

// unpack one half of the packed index words (export for tests)
__global__ void unpack_idx_kernel(long long n, const uint32_t* __restrict__ idx, int field, uint16_t* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint16_t)(field == 0 ? (idx[i] & 0xffffu) : (idx[i] >> 16));
}

extern "C" {

int ss_abi_version(void) { return SS_ABI_VERSION; }

int ss_create(int device, ss_handle_t* out) {
  if (!out) return SS_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev) return SS_ERR_CUDA;
  ss_handle_s* h = new (std::nothrow) ss_handle_s();
  if (!h) return SS_ERR_CUDA;
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess) {
    delete h;
    return SS_ERR_CUDA;
  }
  int v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
  h->num_sms = v > 0 ? v : 148;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  h->max_smem_optin = (size_t)v;
  // keep the stream-ordered pool's memory across synchronisations (temporaries of the builder)
  cudaMemPool_t mp;
  if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  if (const char* e = getenv("SS_SMALL_TILE_THR")) {
    const int v = atoi(e);
    if (v >= 1 && v <= STREAM_SEG_TILES) h->small_tile_thr = v;
  }
  if (const char* e = getenv("SS_CLUSTER_TILE")) {
    const int v = atoi(e);
    if (v >= 512 && v <= E_HARD) h->cluster_tile = v;
  }
  if (const char* e = getenv("SS_CLANE_LIGHT")) h->clane_light = atoi(e) != 0;
  *out = h;
  return SS_OK;
}

int ss_destroy(ss_handle_t h) {
  SS_REQUIRE_HANDLE(h);
  cudaSetDevice(h->device);
  tl_alloc_stream = nullptr;
  free_fit(h);
  free_dense(h);
  free_layout(h);
  for (int i = 0; i < 6; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (int i = 0; i < SS_NUM_AUX_STREAMS; ++i) {
    if (h->tev_s[i]) cudaEventDestroy(h->tev_s[i]);
    if (h->tev_e[i]) cudaEventDestroy(h->tev_e[i]);
  }
  for (int i = 0; i < SS_NUM_AUX_STREAMS; ++i) {
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    if (h->aux[i]) cudaStreamDestroy(h->aux[i]);
  }
  delete h;
  return SS_OK;
}

int ss_release(ss_handle_t h) {
  SS_REQUIRE_HANDLE(h);
  cudaSetDevice(h->device);
  tl_alloc_stream = nullptr;          // the stream of the last call may be gone
  free_fit(h);
  free_dense(h);
  free_layout(h);
  return SS_OK;
}

const char* ss_last_error(ss_handle_t h) { return h ? h->err.c_str() : "null handle"; }

int ss_layout_build(ss_handle_t h, void* stream, int32_t n_rows, int32_t n_cols, int64_t nnz, const int32_t* row_ptr,
                    const int32_t* col_idx, const uint16_t* score, const int32_t* row_segment, int32_t n_segments) {
  SS_REQUIRE_HANDLE(h);
  if (n_rows < 0 || n_cols <= 0 || nnz < 0 || n_segments < 0 || !row_ptr || (n_rows > 0 && !row_segment) ||
      (nnz > 0 && (!col_idx || !score))) {          // n_rows == 0: a rank of a sharded run without rows of its own
    h->set_error("ss_layout_build: bad argument");
    return SS_ERR_ARG;
  }
  if (n_cols > 65535 * 16) {
    h->set_error("ss_layout_build: more than 1M loci not supported");
    return SS_ERR_ARG;
  }
  if (nnz > 2147483647LL) {
    // row_ptr is int32: a matrix with 2^31 or more alignments cannot be described (scipy switches to int64
    // index arrays there); the caller shards the rows over more GPUs
    h->set_error("ss_layout_build: more than 2^31 - 1 alignments (int32 row pointers)");
    return SS_ERR_CAPACITY;
  }
  cudaSetDevice(h->device);
  tl_alloc_stream = (cudaStream_t)stream;
  return layout_build(h, (cudaStream_t)stream, n_rows, n_cols, nnz, row_ptr, col_idx, score, row_segment,
                      n_segments);
}

int ss_layout_import(ss_handle_t h, void* stream, const ss_host_layout* L) {
  SS_REQUIRE_HANDLE(h);
  if (!L) {
    h->set_error("ss_layout_import: null layout");
    return SS_ERR_ARG;
  }
  cudaSetDevice(h->device);
  tl_alloc_stream = (cudaStream_t)stream;
  return layout_import(h, (cudaStream_t)stream, L);
}

int ss_layout_get_info(ss_handle_t h, ss_layout_info* o) {
  SS_REQUIRE_HANDLE(h);
  if (!o || !h->has_layout) {
    h->set_error("ss_layout_get_info: no layout");
    return SS_ERR_STATE;
  }
  const Layout& L = h->L;
  memset(o, 0, sizeof(*o));
  o->n_rows = L.N; o->n_cols = L.K; o->n_segments = L.S; o->n_tiles = L.n_tiles;
  o->max_row_len = L.maxlen; o->chunk = L.chunk; o->n_unique_rows = L.n_uniq;
  o->n_stream_tiles = h->n_stream_tiles; o->n_stream_segments = h->n_stream_segs;
  o->n_resident_tiles = L.n_tiles - h->n_stream_tiles - h->n_cluster_tiles;
  o->n_cluster_tiles = h->n_cluster_tiles; o->n_cluster_segments = h->n_cluster_segs;
  o->nnz = L.A; o->n_segcols = L.n_segcols;
  o->tot_ent = L.tot_ent; o->tot_row = L.tot_row; o->tot_col = L.tot_col; o->tot_jds = L.tot_jds;
  o->nnz_ambiguous = h->nnz_amb; o->rows_ambiguous = h->rows_amb;
  return SS_OK;
}

int ss_layout_export(ss_handle_t h, const char* name, void* dst, int64_t* bytes) {
  SS_REQUIRE_HANDLE(h);
  if (!h->has_layout || !name || !bytes) {
    h->set_error("ss_layout_export: no layout / bad argument");
    return SS_ERR_STATE;
  }
  const Layout& L = h->L;
  struct Ent { const char* n; const void* p; size_t b; };
  const size_t T = (size_t)L.n_tiles, S = (size_t)L.S, NC = (size_t)L.n_segcols;
  const Ent table[] = {
      {"tile_hdr", L.tile_hdr, T * HDR_WORDS * 8},
      {"e_lcol", nullptr, (size_t)L.tot_ent * 2}, {"e_cpos", nullptr, (size_t)L.tot_ent * 2},
      {"e_q", L.e_q, (size_t)L.tot_ent * 8}, {"e_opos", L.e_opos, (size_t)L.tot_ent * 4},
      {"e_idx", L.e_idx, (size_t)L.tot_ent * 4},
      {"r_len", L.r_len, (size_t)L.tot_row * 2}, {"r_w", L.r_w, (size_t)L.tot_row * 8},
      {"r_orow", L.r_orow, (size_t)L.tot_row * 4},
      {"c_ptr", L.c_ptr, (size_t)L.tot_col * 2}, {"c_scol", L.c_scol, (size_t)L.tot_col * 4},
      {"c_pisum0", L.c_pisum0, (size_t)L.tot_col * 8}, {"c_nuniq", L.c_nuniq, (size_t)L.tot_col * 4},
      {"j_ptr", L.j_ptr, (size_t)L.tot_jds * 2},
      {"seg_tile0", L.seg_tile0, S * 4}, {"seg_ntiles", L.seg_ntiles, S * 4}, {"seg_col0", L.seg_col0, S * 4},
      {"seg_ncol", L.seg_ncol, S * 4}, {"seg_smax", L.seg_smax, S * 4}, {"seg_nobs", L.seg_nobs, S * 4},
      {"seg_namb", L.seg_namb, S * 4}, {"seg_nuniq", L.seg_nuniq, S * 4}, {"seg_wmax", L.seg_wmax, S * 8},
      {"seg_wtot", L.seg_wtot, S * 8}, {"seg_wamb", L.seg_wamb, S * 8}, {"seg_logq_uniq", L.seg_logq_uniq, S * 8},
      {"seg_A", L.seg_A, S * 8}, {"seg_Aamb", L.seg_Aamb, S * 8},
      {"sc_gcol", L.sc_gcol, NC * 4}, {"sc_pisum0", L.sc_pisum0, NC * 8}, {"sc_nuniq", L.sc_nuniq, NC * 4},
      {"sc_inamb", h->sc_inamb, NC}, {"u_row", L.u_row, (size_t)L.n_uniq * 4}, {"u_pos", L.u_pos, (size_t)L.n_uniq * 4},
  };
  for (const Ent& e : table) {
    if (strcmp(e.n, name) == 0) {
      if (!dst) {
        *bytes = (int64_t)e.b;
        return SS_OK;
      }
      if ((size_t)*bytes < e.b) {
        h->set_error("ss_layout_export: destination too small");
        return SS_ERR_ARG;
      }
      *bytes = (int64_t)e.b;
      cudaSetDevice(h->device);
      if (e.b && !e.p) {
        // one half of the packed index words
        const int field = strcmp(name, "e_lcol") == 0 ? 0 : 1;
        uint16_t* tmp = nullptr;
        SS_CUDA_CHECK(h, cudaMalloc(&tmp, e.b));
        unpack_idx_kernel<<<(unsigned)((L.tot_ent + 255) / 256), 256>>>(L.tot_ent, L.e_idx, field, tmp);
        cudaError_t ce = cudaMemcpy(dst, tmp, e.b, cudaMemcpyDeviceToHost);
        cudaFree(tmp);
        SS_CUDA_CHECK(h, ce);
        return SS_OK;
      }
      if (e.b) SS_CUDA_CHECK(h, cudaMemcpy(dst, e.p, e.b, cudaMemcpyDeviceToHost));
      return SS_OK;
    }
  }
  h->set_error(std::string("ss_layout_export: unknown array ") + name);
  return SS_ERR_ARG;
}

int ss_em_fit(ss_handle_t h, void* stream, double em_epsilon, int32_t max_iter, double pi_prior, double theta_prior,
              int32_t use_likelihood) {
  SS_REQUIRE_HANDLE(h);
  cudaSetDevice(h->device);
  tl_alloc_stream = (cudaStream_t)stream;
  FitParams P;
  P.eps = em_epsilon;
  P.max_iter = max_iter;
  P.pi_prior = pi_prior;
  P.theta_prior = theta_prior;
  P.use_likelihood = use_likelihood ? 1 : 0;
  P.K = h->L.K;
  return em_fit(h, (cudaStream_t)stream, P);
}

#define COPY_OUT(dst, src, n, T)                                                                          \
  if (dst) SS_CUDA_CHECK(h, cudaMemcpyAsync(dst, src, (size_t)(n) * sizeof(T), cudaMemcpyDeviceToDevice, st))

int ss_em_get_results(ss_handle_t h, void* stream, int32_t* iters, int32_t* flags, double* lnl, double* diffs,
                      double* lnls, int32_t* nobs, int32_t* nfeats) {
  SS_REQUIRE_HANDLE(h);
  if (!h->has_fit) {
    h->set_error("ss_em_get_results: no fit");
    return SS_ERR_STATE;
  }
  cudaSetDevice(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t S = (size_t)h->L.S;
  COPY_OUT(iters, h->F.seg_iters, S, int);
  COPY_OUT(flags, h->F.seg_flags, S, int);
  COPY_OUT(lnl, h->F.seg_lnl, S, double);
  COPY_OUT(diffs, h->F.seg_diffs, S * h->P.max_iter, double);
  if (lnls) {
    if (!h->F.seg_lnls) {
      h->set_error("ss_em_get_results: lnL trace only exists with use_likelihood");
      return SS_ERR_STATE;
    }
    COPY_OUT(lnls, h->F.seg_lnls, S * h->P.max_iter, double);
  }
  if (h->dense_mode() && h->dn_ready) {
    // rows sharded over ranks: observations / loci of the whole model, not of this rank's block
    COPY_OUT(nobs, h->dn_counts, 1, int);
    COPY_OUT(nfeats, h->dn_counts + 1, 1, int);
  } else {
    COPY_OUT(nobs, h->L.seg_nobs, S, int);
    COPY_OUT(nfeats, h->L.seg_ncol, S, int);
  }
  return SS_OK;
}

__global__ void rest_params_kernel(int S, const double* thpw, const double* pipw, const double* inv_thd,
                                   const double* inv_pid, double* pi_rest, double* theta_rest) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  if (pi_rest) pi_rest[s] = pipw[s] * inv_pid[s];
  if (theta_rest) theta_rest[s] = thpw[s] * inv_thd[s];
}

int ss_em_get_params(ss_handle_t h, void* stream, int32_t* seg_col0, int32_t* seg_ncol, int32_t* gcol, double* pi,
                     double* theta, double* pi_rest, double* theta_rest) {
  SS_REQUIRE_HANDLE(h);
  if (!h->has_fit) {
    h->set_error("ss_em_get_params: no fit");
    return SS_ERR_STATE;
  }
  cudaSetDevice(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t S = (size_t)h->L.S, NC = (size_t)h->L.n_segcols;
  COPY_OUT(seg_col0, h->L.seg_col0, S, int);
  COPY_OUT(seg_ncol, h->L.seg_ncol, S, int);
  COPY_OUT(gcol, h->L.sc_gcol, NC, int);
  COPY_OUT(pi, h->F.sc_pi, NC, double);
  COPY_OUT(theta, h->F.sc_theta, NC, double);
  if ((pi_rest || theta_rest) && S > 0) {
    rest_params_kernel<<<(unsigned)((S + 255) / 256), 256, 0, st>>>((int)S, h->F.seg_thpw, h->F.seg_pipw,
                                                                    h->F.seg_inv_thd, h->F.seg_inv_pid, pi_rest,
                                                                    theta_rest);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
  }
  return SS_OK;
}

int ss_em_emit_z(ss_handle_t h, void* stream, double* z) {
  SS_REQUIRE_HANDLE(h);
  if (!z) {
    h->set_error("ss_em_emit_z: null buffer");
    return SS_ERR_ARG;
  }
  cudaSetDevice(h->device);
  return em_emit_z(h, (cudaStream_t)stream, z);
}

int ss_reassign(ss_handle_t h, void* stream, int32_t mode, double conf_thresh, double tie_tol, int32_t n_rows,
                const int32_t* row_ptr, const uint16_t* score, const double* z, uint8_t* flag, int32_t* nbest,
                int64_t* info) {
  SS_REQUIRE_HANDLE(h);
  cudaSetDevice(h->device);
  return reassign(h, (cudaStream_t)stream, mode, conf_thresh, tie_tol, n_rows, row_ptr, score, z, flag, nbest,
                  (long long*)info);
}

int ss_count(ss_handle_t h, void* stream, int32_t n_rows, int32_t n_cols, const int32_t* row_ptr,
             const int32_t* col_idx, const uint8_t* flag, const double* weight, const int32_t* row_cell,
             const int32_t* row_umi, int64_t cap, int32_t* out_cell, int32_t* out_col, double* out_val,
             int64_t* n_out) {
  SS_REQUIRE_HANDLE(h);
  cudaSetDevice(h->device);
  tl_alloc_stream = (cudaStream_t)stream;          // temporaries are allocated and freed in this stream's order
  long long n = 0;
  int rc = count(h, (cudaStream_t)stream, n_rows, n_cols, row_ptr, col_idx, flag, weight, row_cell, row_umi, cap,
                 out_cell, out_col, out_val, &n);
  if (n_out) *n_out = n;
  return rc;
}

int ss_reassign_pick(ss_handle_t h, void* stream, int32_t n_tied, const int32_t* tied_rows, const int32_t* picks,
                     const int32_t* row_ptr, uint8_t* flag) {
  SS_REQUIRE_HANDLE(h);
  cudaSetDevice(h->device);
  return reassign_pick(h, (cudaStream_t)stream, n_tied, tied_rows, picks, row_ptr, flag);
}

int ss_count_cells(ss_handle_t h, void* stream, int32_t n_rows, int32_t n_cols, int32_t n_cells, const int32_t* row_ptr,
                   const int32_t* col_idx, const uint8_t* flag, const int32_t* nbest, const int32_t* row_cell, int64_t cap,
                   int64_t* out_ptr, int32_t* out_col, int32_t* out_cnt, double* out_val, int64_t* n_out) {
  SS_REQUIRE_HANDLE(h);
  cudaSetDevice(h->device);
  tl_alloc_stream = (cudaStream_t)stream;
  long long n = 0;
  int rc = count_cells(h, (cudaStream_t)stream, n_rows, n_cols, n_cells, row_ptr, col_idx, flag, nbest, row_cell, cap,
                       reinterpret_cast<long long*>(out_ptr), out_col, out_cnt, out_val, &n);
  if (n_out) *n_out = n;
  return rc;
}

int ss_reassign_csr(ss_handle_t h, void* stream, int32_t n_rows, const int32_t* row_ptr, const int32_t* col_idx,
                    const uint8_t* flag, const int32_t* nbest, int64_t cap, int64_t* out_ptr, int32_t* out_idx,
                    double* out_val, int64_t* n_out) {
  SS_REQUIRE_HANDLE(h);
  if (!n_out) {
    h->set_error("ss_reassign_csr: bad argument");
    return SS_ERR_ARG;
  }
  cudaSetDevice(h->device);
  tl_alloc_stream = (cudaStream_t)stream;
  long long n = 0;
  const int rc = reassign_csr(h, (cudaStream_t)stream, n_rows, row_ptr, col_idx, flag, nbest, cap,
                              reinterpret_cast<long long*>(out_ptr), out_idx, out_val, &n);
  *n_out = n;
  return rc;
}

int ss_umi_dedup(ss_handle_t h, void* stream, int32_t n_rows, const int32_t* row_ptr, const int32_t* col_idx,
                 const uint16_t* score, const int32_t* row_group, int32_t n_groups, uint8_t* excluded, int32_t* comp_root,
                 int32_t* group_size, int32_t* group_ncomp) {
  SS_REQUIRE_HANDLE(h);
  if (n_rows < 0 || n_groups < 0 || !row_ptr || !row_group || !excluded || !comp_root || !group_size || !group_ncomp) {
    h->set_error("ss_umi_dedup: bad argument");
    return SS_ERR_ARG;
  }
  cudaSetDevice(h->device);
  return umi_dedup(h, (cudaStream_t)stream, n_rows, row_ptr, col_idx, score, row_group, n_groups, excluded, comp_root,
                   group_size, group_ncomp);
}

int ss_scores_build(ss_handle_t h, void* stream, int64_t n, const int32_t* read, const int32_t* feat,
                    const int32_t* alnscore, const int32_t* alnlen, int32_t min_as, int32_t n_reads, int32_t n_feats,
                    int32_t* row_ptr, int32_t* col_idx, uint16_t* score, int64_t* info) {
  SS_REQUIRE_HANDLE(h);
  if (!row_ptr || !info || (n > 0 && (!read || !feat || !alnscore || !alnlen || !col_idx || !score))) {
    h->set_error("ss_scores_build: bad argument");
    return SS_ERR_ARG;
  }
  cudaSetDevice(h->device);
  return scores_build(h, (cudaStream_t)stream, n, read, feat, alnscore, alnlen, min_as, n_reads, n_feats, row_ptr, col_idx,
                      score, reinterpret_cast<long long*>(info));
}

int ss_mtx_measure(ss_handle_t h, void* stream, int64_t n, const int32_t* row, const int32_t* col, const double* val,
                   int32_t integer, int64_t* off, int64_t* n_bytes) {
  SS_REQUIRE_HANDLE(h);
  if (!n_bytes) {
    h->set_error("ss_mtx_measure: bad argument");
    return SS_ERR_ARG;
  }
  cudaSetDevice(h->device);
  tl_alloc_stream = (cudaStream_t)stream;
  long long nb = 0;
  const int rc = mtx_measure(h, (cudaStream_t)stream, n, row, col, val, integer, reinterpret_cast<long long*>(off), &nb);
  *n_bytes = nb;
  return rc;
}

int ss_mtx_write(ss_handle_t h, void* stream, int64_t n, const int32_t* row, const int32_t* col, const double* val,
                 int32_t integer, const int64_t* off, uint8_t* out) {
  SS_REQUIRE_HANDLE(h);
  cudaSetDevice(h->device);
  return mtx_write(h, (cudaStream_t)stream, n, row, col, val, integer, reinterpret_cast<const long long*>(off), out);
}

int ss_set_allreduce(ss_handle_t h, ss_allreduce_fn fn, void* user) {
  SS_REQUIRE_HANDLE(h);
  h->allreduce = fn;
  h->allreduce_user = user;
  return SS_OK;
}

int64_t ss_peer_buffer_bytes(int32_t n_cols) {
  const int nchunks = (n_cols + SS_DN_TAIL + 64 + SS_DN_THREADS - 1) / SS_DN_THREADS;
  return (int64_t)(2 * (int64_t)nchunks * SS_DN_THREADS + SS_DN_MAXW) * 8;
}

int ss_set_peers(ss_handle_t h, int32_t rank, int32_t world, void* const* peer_bufs, int64_t bytes) {
  SS_REQUIRE_HANDLE(h);
  if (world < 1 || world > SS_DN_MAXW || rank < 0 || rank >= world) {
    h->set_error("ss_set_peers: need 0 <= rank < world <= 8");
    return SS_ERR_ARG;
  }
  cudaSetDevice(h->device);
  free_dense(h);
  h->dn_world = world;
  h->dn_rank = rank;
  h->dn_epoch = 0;
  h->dn_peers = peer_bufs != nullptr;
  // a single rank has nothing to exchange: the persistent kernel runs on plain device buffers
  // (bytes < 0 selects the host-loop variant, for tests)
  h->dn_persistent = h->dn_peers || (world == 1 && bytes >= 0);
  h->dn_force = true;
  h->dn_peer_bytes = peer_bufs ? (size_t)bytes : 0;
  for (int r = 0; r < SS_DN_MAXW; ++r) h->dn_peer_base[r] = (peer_bufs && r < world) ? peer_bufs[r] : nullptr;
  return SS_OK;
}

int ss_em_get_dense_params(ss_handle_t h, void* stream, double* pi, double* theta) {
  SS_REQUIRE_HANDLE(h);
  if (!h->has_fit || !h->dense_mode() || !h->dn_ready) {
    h->set_error("ss_em_get_dense_params: no row-sharded fit");
    return SS_ERR_STATE;
  }
  cudaSetDevice(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  COPY_OUT(pi, h->dn_pi, h->L.K, double);
  COPY_OUT(theta, h->dn_theta, h->L.K, double);
  return SS_OK;
}

int64_t ss_launch_count(ss_handle_t h) { return h ? h->launches : 0; }

int ss_lane_classes(int32_t* threads, int32_t cap) {
  if (!threads) return SS_NUM_CLASSES;
  for (int i = 0; i < SS_NUM_CLASSES && i < cap; ++i) threads[i] = SS_CLS_THREADS[i];
  return SS_NUM_CLASSES;
}

int ss_set_timing(ss_handle_t h, int32_t enabled) {
  SS_REQUIRE_HANDLE(h);
  cudaSetDevice(h->device);
  h->timing = enabled != 0;
  if (h->timing && !h->ev[0]) {
    for (int i = 0; i < 6; ++i) SS_CUDA_CHECK(h, cudaEventCreate(&h->ev[i]));
    for (int i = 0; i < SS_NUM_AUX_STREAMS; ++i) {
      SS_CUDA_CHECK(h, cudaEventCreate(&h->tev_s[i]));
      SS_CUDA_CHECK(h, cudaEventCreate(&h->tev_e[i]));
    }
  }
  return SS_OK;
}

int ss_get_timings(ss_handle_t h, double* ms, int32_t n) {
  SS_REQUIRE_HANDLE(h);
  if (!ms || n < 5 + SS_NUM_AUX_STREAMS || !h->ev_valid) {
    h->set_error("ss_get_timings: timing not enabled / no fit / buffer too small");
    return SS_ERR_STATE;
  }
  cudaSetDevice(h->device);
  SS_CUDA_CHECK(h, cudaEventSynchronize(h->ev[5]));
  for (int i = 0; i < 5; ++i) {
    float t = 0.f;
    SS_CUDA_CHECK(h, cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]));
    ms[i] = t;
  }
  // spans of the concurrent EM kernels: size classes 0..5, then the streaming kernel
  for (int i = 0; i < SS_NUM_AUX_STREAMS; ++i) {
    float t = 0.f;
    if (h->aux_used[i]) SS_CUDA_CHECK(h, cudaEventElapsedTime(&t, h->tev_s[i], h->tev_e[i]));
    if (5 + i < n) ms[5 + i] = t;
  }
  return SS_OK;
}

}  // extern "C"
