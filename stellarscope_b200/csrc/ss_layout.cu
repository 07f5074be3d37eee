// Layout management: import of a host-built layout (debug/test), the launch plan
// (which tile goes to which kernel) and -- below -- the device builder.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "ss_em.cuh"
#include "ss_handle.h"

#ifndef SS_PLAN_RULE
#define SS_PLAN_RULE 1     // 1: smallest CTA that holds all fragments; 0: cost model (same speed on B200)
#endif

namespace ss {

thread_local cudaStream_t tl_alloc_stream = nullptr;

void free_layout(ss_handle_s* h) {
  dev_free_all(h->layout_allocs);
  h->L = Layout();
  h->hdr_host.clear();
  h->seg_ntiles_host.clear();
  h->seg_tile0_host.clear();
  for (int c = 0; c < SS_NUM_CLASSES; ++c) {
    h->cls_tiles[c] = nullptr;
    h->cls_count[c] = 0;
    h->cls_smem[c] = 0;
  }
  h->stream_tiles = h->stream_segs = h->stream_col_seg = h->stream_col_slot = nullptr;
  h->stream_chunk_seg = h->stream_chunk_j0 = nullptr; h->n_stream_chunks = 0;
  h->n_stream_tiles = h->n_stream_segs = 0;
  h->n_stream_cols = 0;
  h->all_tiles = nullptr;
  h->n_col_chunks = 0;
  h->chunk_seg = h->chunk_j0 = h->seg_chunk0 = nullptr;
  h->chunk_part = nullptr;
  h->sc_inamb = nullptr;
  for (int k = 0; k <= ss_handle_s::MAX_CLUSTER; ++k) {
    h->clu_segs[k] = h->clu_segs_light[k] = nullptr;
    h->clu_count[k] = h->clu_count_light[k] = 0;
    h->clu_smem[k] = h->clu_smem_lane[k] = h->clu_smem_light[k] = 0;
  }
  h->clu_xref = nullptr; h->seg_xref_off = nullptr; h->seg_xmax = nullptr; h->n_cluster_tiles = h->n_cluster_segs = 0;
  h->has_layout = false;
}

__global__ void pack_idx_kernel(long long n, const uint16_t* __restrict__ lcol, const uint16_t* __restrict__ cpos,
                                uint32_t* __restrict__ idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = pack_idx(lcol[i], cpos[i]);
}

// cross reference of the cluster kernel: for every column of a clustered segment, its local
// column index inside each tile of the segment
__global__ void cluster_xref_kernel(const int* __restrict__ tiles, int n, Layout L, const long long* __restrict__ seg_xref_off,
                                    uint16_t* __restrict__ xref) {
  const int t = tiles[blockIdx.x];
  const long long* h = L.tile_hdr + (long long)HDR_WORDS * t;
  const int seg = (int)h[H_SEG], ncol = (int)h[H_NCOL];
  const long long co = h[H_COL_OFF];
  const int nt = L.seg_ntiles[seg], k = t - L.seg_tile0[seg], col0 = L.seg_col0[seg];
  const long long off = seg_xref_off[seg];
  for (int c = threadIdx.x; c < ncol; c += blockDim.x)
    xref[off + (long long)(L.c_scol[co + c] - col0) * nt + k] = (uint16_t)c;
}

// per-column lists of the streamed segments: index of the segment in the list of streamed segments, absolute slot of
// the column.  One CTA per streamed segment.
__global__ void stream_cols_kernel(const int* __restrict__ segs, const long long* __restrict__ first,
                                   const int* __restrict__ seg_col0, const int* __restrict__ seg_ncol,
                                   int* __restrict__ col_seg, int* __restrict__ col_slot) {
  const int si = blockIdx.x, s = segs[si];
  const long long o = first[si];
  const int c0 = seg_col0[s], n = seg_ncol[s];
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    col_seg[o + j] = si;
    col_slot[o + j] = c0 + j;
  }
}

// number of a segment's columns that a tile touches (the others belong to unique rows only); one warp per segment
__global__ void seg_tile_cols_kernel(int S, const int* __restrict__ seg_col0, const int* __restrict__ seg_ncol,
                                     const int8_t* __restrict__ sc_inamb, int* __restrict__ out) {
  const int s = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (s >= S) return;
  const int c0 = seg_col0[s], n = seg_ncol[s];
  int cnt = 0;
  for (int j = lane; j < n; j += 32) cnt += sc_inamb[c0 + j] != 0;
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) out[s] = cnt;
}

template <class T>
static cudaError_t upload(std::vector<void*>& pool, T** dst, const T* src, size_t n, cudaStream_t st) {
  cudaError_t e = dev_alloc(pool, dst, n);
  if (e != cudaSuccess || n == 0) return e;
  return cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
}

int layout_import(ss_handle_s* h, cudaStream_t st, const ss_host_layout* HL) {
  free_fit(h);
  free_layout(h);
  Layout& L = h->L;
  auto& pool = h->layout_allocs;
  L.N = HL->N; L.K = HL->K; L.S = HL->S; L.n_tiles = HL->n_tiles; L.maxlen = HL->maxlen; L.chunk = HL->chunk;
  L.n_uniq = HL->n_uniq; L.A = HL->A; L.n_segcols = HL->n_segcols;
  L.tot_ent = HL->tot_ent; L.tot_row = HL->tot_row; L.tot_col = HL->tot_col; L.tot_jds = HL->tot_jds;
  const size_t T = (size_t)L.n_tiles, S = (size_t)L.S, NC = (size_t)L.n_segcols;
  SS_CUDA_CHECK(h, upload(pool, &L.tile_hdr, (const long long*)HL->tile_hdr, T * HDR_WORDS, st));
  {
    std::vector<void*> tmp;
    uint16_t *d_lcol, *d_cpos;
    SS_CUDA_CHECK(h, upload(tmp, &d_lcol, HL->e_lcol, (size_t)L.tot_ent, st));
    SS_CUDA_CHECK(h, upload(tmp, &d_cpos, HL->e_cpos, (size_t)L.tot_ent, st));
    SS_CUDA_CHECK(h, upload(pool, &L.e_q, HL->e_q, (size_t)L.tot_ent, st));
    SS_CUDA_CHECK(h, dev_alloc(pool, &L.e_idx, (size_t)L.tot_ent));
    if (L.tot_ent > 0) {
      pack_idx_kernel<<<(unsigned)((L.tot_ent + 255) / 256), 256, 0, st>>>(L.tot_ent, d_lcol, d_cpos, L.e_idx);
      h->launches++;
    }
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
    for (void* q : tmp) cudaFreeAsync(q, st);
  }
  SS_CUDA_CHECK(h, upload(pool, &L.e_opos, HL->e_opos, (size_t)L.tot_ent, st));
  SS_CUDA_CHECK(h, upload(pool, &L.r_len, HL->r_len, (size_t)L.tot_row, st));
  SS_CUDA_CHECK(h, upload(pool, &L.r_w, HL->r_w, (size_t)L.tot_row, st));
  SS_CUDA_CHECK(h, upload(pool, &L.r_orow, HL->r_orow, (size_t)L.tot_row, st));
  SS_CUDA_CHECK(h, upload(pool, &L.c_ptr, HL->c_ptr, (size_t)L.tot_col, st));
  SS_CUDA_CHECK(h, upload(pool, &L.c_scol, HL->c_scol, (size_t)L.tot_col, st));
  SS_CUDA_CHECK(h, upload(pool, &L.c_pisum0, HL->c_pisum0, (size_t)L.tot_col, st));
  SS_CUDA_CHECK(h, upload(pool, &L.c_nuniq, HL->c_nuniq, (size_t)L.tot_col, st));
  SS_CUDA_CHECK(h, upload(pool, &L.j_ptr, HL->j_ptr, (size_t)L.tot_jds, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_tile0, HL->seg_tile0, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_ntiles, HL->seg_ntiles, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_col0, HL->seg_col0, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_ncol, HL->seg_ncol, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_smax, HL->seg_smax, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_nobs, HL->seg_nobs, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_namb, HL->seg_namb, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_nuniq, HL->seg_nuniq, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_wmax, HL->seg_wmax, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_wtot, HL->seg_wtot, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_wamb, HL->seg_wamb, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_logq_uniq, HL->seg_logq_uniq, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_A, (const long long*)HL->seg_A, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.seg_Aamb, (const long long*)HL->seg_Aamb, S, st));
  SS_CUDA_CHECK(h, upload(pool, &L.sc_gcol, HL->sc_gcol, NC, st));
  SS_CUDA_CHECK(h, upload(pool, &L.sc_pisum0, HL->sc_pisum0, NC, st));
  SS_CUDA_CHECK(h, upload(pool, &L.sc_nuniq, HL->sc_nuniq, NC, st));
  SS_CUDA_CHECK(h, upload(pool, &h->sc_inamb, HL->sc_inamb, NC, st));
  SS_CUDA_CHECK(h, upload(pool, &L.u_row, HL->u_row, (size_t)L.n_uniq, st));
  SS_CUDA_CHECK(h, upload(pool, &L.u_pos, HL->u_pos, (size_t)L.n_uniq, st));
  h->hdr_host.assign((const long long*)HL->tile_hdr, (const long long*)HL->tile_hdr + T * HDR_WORDS);
  h->seg_ntiles_host.assign(HL->seg_ntiles, HL->seg_ntiles + S);
  h->seg_tile0_host.assign(HL->seg_tile0, HL->seg_tile0 + S);
  std::vector<int> ncol_host(HL->seg_ncol, HL->seg_ncol + S), col0_host(HL->seg_col0, HL->seg_col0 + S);
  int rc = layout_plan(h, st, ncol_host, col0_host);
  if (rc != SS_OK) return rc;
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  h->has_layout = true;
  return SS_OK;
}

// Decide which kernel runs each tile.  Tiles of single-tile segments go to the
// resident kernel, binned by shared-memory need (CTAs per SM); tiles of larger
// segments go to the streaming kernel.
int layout_plan(ss_handle_s* h, cudaStream_t st, const std::vector<int>& seg_ncol,
                const std::vector<int>& seg_col0) {
  Layout& L = h->L;
  auto& pool = h->layout_allocs;
  const int T = L.n_tiles;
  std::vector<int> cls[SS_NUM_CLASSES], stream, all(T);
  std::iota(all.begin(), all.end(), 0);
  size_t all_smem = 0;
  h->nnz_amb = 0;
  h->rows_amb = 0;
  for (int c = 0; c < SS_NUM_CLASSES; ++c) h->cls_smem[c] = 0;
  h->stream_smem = 0;
  for (int i = 0; i < 4; ++i) h->geom_all[i] = h->geom_stream[i] = 8;
  // segments of 2..8 tiles whose tiles (plus the owner arrays) fit shared memory run as clusters
  const int MAXC = ss_handle_s::MAX_CLUSTER;
  std::vector<char> seg_clustered((size_t)std::max(L.S, 1), 0);
  std::vector<int> clu[ss_handle_s::MAX_CLUSTER + 1], clu_light[ss_handle_s::MAX_CLUSTER + 1];
  // shared memory a light CTA may use: half of an SM's (the opt-in maximum per CTA is the SM's minus 1 KB; every
  // resident CTA reserves 1 KB), minus the kernel's static shared memory
  const size_t light_cap = (h->max_smem_optin + 1024) / 2 - 1024 - 1536;
  std::vector<long long> xoff((size_t)std::max(L.S, 1), -1);
  long long xref_total = 0;
  for (int k = 0; k <= MAXC; ++k) h->clu_smem[k] = h->clu_smem_lane[k] = h->clu_smem_light[k] = 0;
  // columns of every segment that its tiles touch: sum over the tiles' columns minus this = sum over the columns of
  // (holders - 1) = the partial sums a tile of the cluster lane kernel can receive per pass (its receive buffer)
  std::vector<int> tile_cols((size_t)std::max(L.S, 1), 0), seg_xmax((size_t)std::max(L.S, 1), 0);
  if (L.S > 0) {
    int* d_cnt;
    SS_CUDA_CHECK(h, dev_alloc(pool, &d_cnt, (size_t)L.S));
    seg_tile_cols_kernel<<<(unsigned)(((long long)L.S * 32 + 255) / 256), 256, 0, st>>>(L.S, L.seg_col0, L.seg_ncol, h->sc_inamb, d_cnt);
    h->launches++;
    SS_CUDA_CHECK(h, cudaMemcpyAsync(tile_cols.data(), d_cnt, (size_t)L.S * 4, cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  }
  for (int sgm = 0; sgm < L.S; ++sgm) {
    const int nt = h->seg_ntiles_host[sgm];
    if (nt < 2 || nt > MAXC) continue;
    const size_t no_pad = (size_t)pad8((seg_ncol[sgm] + nt - 1) / nt + 1);
    const size_t owner_bytes = 8 * 4 * no_pad + 2 * no_pad * nt;
    size_t need = 0;
    int pemax = 0, pcmax = 0, xsum = 0;
    long long rfmax = 0;
    for (int k = 0; k < nt; ++k) {
      const long long* hd = &h->hdr_host[(size_t)(h->seg_tile0_host[sgm] + k) * HDR_WORDS];
      rfmax = std::max(rfmax, hd[H_RGT32] > 0 ? (long long)1 << 40 : hd[H_RFRAG]);   // no generic rows in a light tile
      need = std::max(need, tile_smem_bytes((int)hd[H_NENT], (int)hd[H_NROWS], (int)hd[H_NCOL], (int)hd[H_MAXLEN]) +
                                owner_bytes);
      pemax = std::max(pemax, pad8((int)hd[H_NENT]));
      pcmax = std::max(pcmax, pad8((int)hd[H_NCOL] + 1));
      xsum += (int)hd[H_NCOL];
    }
    seg_xmax[sgm] = pad8(std::max(xsum - tile_cols[sgm], 0) + 1);
    const size_t need_lane = clane_smem_bytes(pemax, pcmax, nt, seg_xmax[sgm]);
    if (need + 2048 > h->max_smem_optin || need_lane + 2048 > h->max_smem_optin) continue;
    seg_clustered[sgm] = 1;
    h->clu_smem[nt] = std::max(h->clu_smem[nt], need);
    if (h->clane_light && rfmax <= CLANE_LIGHT_FR * CLANE_LIGHT_THREADS && need_lane + 256 <= light_cap) {
      clu_light[nt].push_back(sgm);
      h->clu_smem_light[nt] = std::max(h->clu_smem_light[nt], need_lane);
    } else {
      clu[nt].push_back(sgm);
      h->clu_smem_lane[nt] = std::max(h->clu_smem_lane[nt], need_lane);
    }
    xoff[sgm] = xref_total;
    xref_total += (long long)seg_ncol[sgm] * nt;
  }
  std::vector<int> cluster_tiles;
  for (int t = 0; t < T; ++t) {
    const long long* hd = &h->hdr_host[(size_t)t * HDR_WORDS];
    const int seg = (int)hd[H_SEG], nrows = (int)hd[H_NROWS], nent = (int)hd[H_NENT], ncol = (int)hd[H_NCOL],
              maxlen = (int)hd[H_MAXLEN];
    h->nnz_amb += nent;
    h->rows_amb += nrows;
    const size_t b0 = tile_smem_bytes(nent, nrows, ncol, maxlen);
    all_smem = std::max(all_smem, b0);
    const int gm[4] = {pad8(nent), pad8(nrows), pad8(ncol + 1), pad8(maxlen + 1)};
    for (int i = 0; i < 4; ++i) h->geom_all[i] = std::max(h->geom_all[i], gm[i]);
    if (b0 + 1024 > h->max_smem_optin) {
      h->set_error("tile exceeds the shared-memory capacity");
      return SS_ERR_CAPACITY;
    }
    if (h->seg_ntiles_host[seg] == 1) {
      // lane kernel: among the CTA sizes whose lanes hold all row / column fragments in registers, the
      // one with the smallest (warps x time per iteration): a warp's iteration costs about 50
      // instructions per row slot, 36 per column slot and 55 of fixed overhead, and the CTA waits for
      // its slowest warp at two barriers, so a partly filled extra slot is paid by every warp
      const long long rfrag = hd[H_RFRAG], cfrag = hd[H_CFRAG];
      const long long vr = (rfrag + 31) / 32, vc = (cfrag + 31) / 32;
      int c = SS_NUM_CLASSES - 1;
      long long best = -1;
      for (int k = 0; k < SS_NUM_CLASSES - 1; ++k) {
        const long long w = SS_CLS_THREADS[k] / 32;
        const long long sr = (vr + w - 1) / w, sc = (vc + w - 1) / w;
        if (sr > SS_FR || sc > SS_FC) continue;
#if SS_PLAN_RULE == 0
        const long long cost = w * (50 * sr + 36 * sc + 55);
        if (best < 0 || cost < best) { best = cost; c = k; }
#else
        (void)best;
        c = k;                       // smallest CTA whose lanes hold all fragments
        break;
#endif
      }
      cls[c].push_back(t);
      h->cls_smem[c] = std::max(h->cls_smem[c], lane_smem_bytes(nent, ncol));
    } else if (seg_clustered[seg]) {
      cluster_tiles.push_back(t);
    } else {
      stream.push_back(t);
      h->stream_smem = std::max(h->stream_smem, b0);
      for (int i = 0; i < 4; ++i) h->geom_stream[i] = std::max(h->geom_stream[i], gm[i]);
    }
  }
  h->all_smem = all_smem;
  // largest first, ties by id: sorted as packed integer keys (a comparator that looks the sizes up in the 40 MB header
  // table misses the cache on every comparison)
  auto sort_by_size = [&](std::vector<int>& v, auto size_of) {
    std::vector<unsigned long long> keys(v.size());
    for (size_t i = 0; i < v.size(); ++i)
      keys[i] = ((unsigned long long)(unsigned)(0x7fffffff - (int)size_of(v[i])) << 32) | (unsigned)v[i];
    std::sort(keys.begin(), keys.end());
    for (size_t i = 0; i < v.size(); ++i) v[i] = (int)(keys[i] & 0xffffffffu);
  };
  auto tile_nent = [&](int t) { return h->hdr_host[(size_t)t * HDR_WORDS + H_NENT]; };
  for (int c = 0; c < SS_NUM_CLASSES; ++c) {
    sort_by_size(cls[c], tile_nent);
    h->cls_count[c] = (int)cls[c].size();
    SS_CUDA_CHECK(h, upload(pool, &h->cls_tiles[c], (const int*)cls[c].data(), cls[c].size(), st));
  }
  sort_by_size(stream, tile_nent);
  h->n_stream_tiles = (int)stream.size();
  SS_CUDA_CHECK(h, upload(pool, &h->stream_tiles, (const int*)stream.data(), stream.size(), st));
  SS_CUDA_CHECK(h, upload(pool, &h->all_tiles, (const int*)all.data(), all.size(), st));
  // column chunks of all segments (prepare / finalize)
  {
    std::vector<int> cseg, cj0, sc0((size_t)std::max(L.S, 1), 0);
    for (int s = 0; s < L.S; ++s) {
      sc0[s] = (int)cseg.size();
      for (int j = 0; j < seg_ncol[s]; j += SS_COL_CHUNK) {
        cseg.push_back(s);
        cj0.push_back(j);
      }
    }
    h->n_col_chunks = (int)cseg.size();
    SS_CUDA_CHECK(h, upload(pool, &h->chunk_seg, (const int*)cseg.data(), cseg.size(), st));
    SS_CUDA_CHECK(h, upload(pool, &h->chunk_j0, (const int*)cj0.data(), cj0.size(), st));
    SS_CUDA_CHECK(h, upload(pool, &h->seg_chunk0, (const int*)sc0.data(), sc0.size(), st));
    SS_CUDA_CHECK(h, dev_alloc(pool, &h->chunk_part, (size_t)std::max(2 * h->n_col_chunks, 1)));
  }
  // multi-tile segments and their columns
  // (the per-column lists -- 10^7 entries for the ~900 streamed cells of the atlas configuration -- are filled on
  // the device: stream_cols_kernel)
  std::vector<int> segs;
  std::vector<long long> seg_first;                      // first entry of a streamed segment in the column lists
  long long n_stream_cols = 0;
  for (int s = 0; s < L.S; ++s) {
    if (h->seg_ntiles_host[s] > 1 && !seg_clustered[s]) {
      segs.push_back(s);
      seg_first.push_back(n_stream_cols);
      n_stream_cols += seg_ncol[s];
    }
  }
  // column chunks of the streamed segments: the update pass of the streaming kernel runs one warp per chunk
  {
    std::vector<int> sch_seg, sch_j0;
    for (int s2 : segs)
      for (int j = 0; j < seg_ncol[s2]; j += SS_COL_CHUNK) {
        sch_seg.push_back(s2);
        sch_j0.push_back(j);
      }
    h->n_stream_chunks = (int)sch_seg.size();
    SS_CUDA_CHECK(h, upload(pool, &h->stream_chunk_seg, (const int*)sch_seg.data(), sch_seg.size(), st));
    SS_CUDA_CHECK(h, upload(pool, &h->stream_chunk_j0, (const int*)sch_j0.data(), sch_j0.size(), st));
  }
  h->n_stream_segs = (int)segs.size();
  h->n_stream_cols = n_stream_cols;
  SS_CUDA_CHECK(h, upload(pool, &h->stream_segs, (const int*)segs.data(), segs.size(), st));
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->stream_col_seg, (size_t)n_stream_cols));
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->stream_col_slot, (size_t)n_stream_cols));
  if (!segs.empty()) {
    long long* d_first;
    SS_CUDA_CHECK(h, upload(pool, &d_first, (const long long*)seg_first.data(), seg_first.size(), st));
    stream_cols_kernel<<<(unsigned)segs.size(), 256, 0, st>>>(h->stream_segs, d_first, L.seg_col0, L.seg_ncol, h->stream_col_seg,
                                                              h->stream_col_slot);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
  }
  // cluster segments
  h->n_cluster_tiles = (int)cluster_tiles.size();
  h->n_cluster_segs = 0;
  for (int k = 2; k <= MAXC; ++k) {
    // big segments first
    auto last_tile_nent = [&](int sg) { return h->hdr_host[(size_t)(h->seg_tile0_host[sg] + k - 1) * HDR_WORDS + H_NENT]; };
    sort_by_size(clu[k], last_tile_nent);
    sort_by_size(clu_light[k], last_tile_nent);
    h->clu_count[k] = (int)clu[k].size();
    h->clu_count_light[k] = (int)clu_light[k].size();
    h->n_cluster_segs += h->clu_count[k] + h->clu_count_light[k];
    // one list per cluster size, heavy segments first: the likelihood-mode kernel takes them all
    clu[k].insert(clu[k].end(), clu_light[k].begin(), clu_light[k].end());
    SS_CUDA_CHECK(h, upload(pool, &h->clu_segs[k], (const int*)clu[k].data(), clu[k].size(), st));
    h->clu_segs_light[k] = h->clu_segs[k] + h->clu_count[k];
  }
  SS_CUDA_CHECK(h, upload(pool, &h->seg_xref_off, (const long long*)xoff.data(), xoff.size(), st));
  SS_CUDA_CHECK(h, upload(pool, &h->seg_xmax, (const int*)seg_xmax.data(), seg_xmax.size(), st));
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->clu_xref, (size_t)std::max<long long>(xref_total, 1)));
  if (xref_total > 0) {
    int* d_ct;
    SS_CUDA_CHECK(h, upload(pool, &d_ct, (const int*)cluster_tiles.data(), cluster_tiles.size(), st));
    SS_CUDA_CHECK(h, cudaMemsetAsync(h->clu_xref, 0xff, (size_t)xref_total * sizeof(uint16_t), st));
    cluster_xref_kernel<<<(unsigned)cluster_tiles.size(), 128, 0, st>>>(d_ct, (int)cluster_tiles.size(), L,
                                                                         h->seg_xref_off, h->clu_xref);
    h->launches++;
    SS_CUDA_CHECK(h, cudaGetLastError());
  }
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));  // host vectors go out of scope
  return SS_OK;
}

}  // namespace ss
