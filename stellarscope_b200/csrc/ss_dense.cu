// Row-sharded fit of ONE model over several GPUs (pooling_mode=pseudobulk).
//
// Reference being replaced: TelescopeLikelihood.em (stellarscope/utils/model.py:325-373) on the
// whole matrix -- single-threaded scipy in the reference.  Every rank holds a contiguous block
// of the rows (its own tiles); the parameters live in dense K-long vectors that are identical on
// every rank.  One EM iteration =
//     tile pass      local rows: r_i = sum_j Q_ij x_j, partial thetasum T_j (model.py:207-210, 236-239)
//     exchange       T = sum over ranks of the partial T   (the ONE exchange step of the path)
//     column update  theta', pi', x' = pi' theta', |delta pi| (model.py:240-246, 340), identical on all ranks
//
// Two ways to run the exchange:
//   peers   (ss_set_peers)     one persistent cooperative kernel per rank runs the whole fit; the
//           partial sums sit in NVLink-mapped symmetric memory, ranks signal each other with
//           release/acquire flags and every rank adds the W partial vectors itself, in rank order,
//           straight out of its peers' memory (P2P loads).  No host round trip, no NCCL call inside
//           the loop, and because every rank adds in the same order the convergence decision is
//           bit-identical everywhere (no broadcast needed).
//   hook    (ss_set_allreduce) the host loop launches pass / all-reduce (caller's NCCL) / update
//           per iteration -- the baseline the fused kernel is measured against, and the fallback
//           when peer memory is unavailable.
// With world == 1 the peers path is a plain single-GPU dense fit (used by the tests).
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "ss_em.cuh"
#include "ss_handle.h"

namespace cg = cooperative_groups;

namespace ss {

constexpr int DN_THREADS = SS_DN_THREADS;
constexpr int SS_ST_FR = 2, SS_ST_FC = 3;      // fragments per lane of the streaming tile pass (1024 lanes)
constexpr long long DN_SPIN_LIMIT = 20000000000LL;   // ~10 s of SM clocks: a lost peer must not hang the GPU

struct DenseArgs {
  int K, Kp, world, rank, n_tiles, nchunks, max_iter, summed;
  double eps;
  double* T[2];                         // this rank's exchange buffers, Kp + SS_DN_TAIL doubles each
  const double* peerT[2][SS_DN_MAXW];   // every rank's buffers (peers mode), by rank
  unsigned long long* flags;            // [SS_DN_MAXW], this rank's memory: last epoch signalled by rank r
  unsigned long long* peer_flags[SS_DN_MAXW];
  unsigned long long epoch0;
  const double* pisum0;                 // K, summed over ranks
  const double* nuniq;                  // K, summed over ranks (double)
  double* pi;
  double* theta;
  double* x[2];
  double* part;                         // [2 * nchunks]: |delta pi| and unique-row lnL partials per chunk
  const int* tgcol;                     // locus of every tile column
  double thpw, pipw, inv_thd, inv_pid, logq_uniq;
  StreamGeom geom;                      // shared-memory geometry of the streaming tile pass
  int* ctl;                             // [0] iterations (0 = running), [1] flags, [2] communication error
  double* diffs;
  double* lnls;
  double* lnl_out;
};

__device__ __forceinline__ double ld_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- tile pass: partial thetasum of the local rows -> T (atomics; order-free, absorbed by the tie tolerance)
// ---- lnL of the local ambiguous rows: z from xold, parameters xnew (model.py:291-317) -> *acc
template <int THREADS>
__device__ __forceinline__ void dense_lnl_pass(const EmArgs& a, const DenseArgs& d, const double* xold,
                                               const double* xnew, double* acc, unsigned char* dsm, uint64_t* bar,
                                               uint32_t& phase, double* s_red) {
  const int tid = threadIdx.x;
  for (int t = blockIdx.x; t < d.n_tiles; t += gridDim.x) {
    const TileView v = load_tile_view(a.L.tile_hdr, t);
    TileSmem s = carve_tile(dsm, v);
    if (tid == 0) {
      fence_proxy_async();
      tile_issue_loads(a.L, v, s, bar);
    }
    const int* __restrict__ gc = d.tgcol + v.co;
    for (int c = tid; c < v.ncol; c += THREADS) {
      s.pt[c] = __ldcg(xold + gc[c]);
      s.pt2[c] = __ldcg(xnew + gc[c]);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    __syncthreads();
    double l = tile_lnl_pass<THREADS>(v, s, s.pt, s.pt2, true, nullptr, nullptr);
    l = block_sum<THREADS>(l, s_red, 0);
    if (tid == 0) atomicAdd(acc, l);
    __syncthreads();
  }
}

// ---- column update over fixed chunks of THREADS columns (identical arithmetic on every rank)
template <int THREADS>
__device__ __forceinline__ void dense_update(const DenseArgs& d, int cur, int it, double* s_red) {
  const int tid = threadIdx.x;
  const double invK = 1.0 / (double)d.K;
  const bool remote = d.world > 1 && !d.summed;
  for (int chunk = blockIdx.x; chunk < d.nchunks; chunk += gridDim.x) {
    const int j = chunk * THREADS + tid;
    double dj = 0.0, lu = 0.0;
    if (j < d.K) {
      double T = 0.0;
      if (remote) {
        double pk[SS_DN_MAXW];
#pragma unroll
        for (int r = 0; r < SS_DN_MAXW; ++r) pk[r] = r < d.world ? ld_sys_f64(d.peerT[cur][r] + j) : 0.0;
#pragma unroll
        for (int r = 0; r < SS_DN_MAXW; ++r) T += pk[r];          // fixed rank order
      } else {
        T = __ldcg(d.T[cur] + j);
      }
      const double th = (T + d.thpw) * d.inv_thd;                 // model.py:240
      const double pn = (d.pisum0[j] + T + d.pipw) * d.inv_pid;   // model.py:243-244
      const double pold = it == 1 ? invK : d.pi[j];               // pi_0 = 1/K, model.py:164
      dj = fabs(pn - pold);                                       // model.py:340
      d.pi[j] = pn;
      d.theta[j] = th;
      d.x[cur ^ 1][j] = pn * th;
      const double nu = d.nuniq[j];
      if (nu > 0.0 && pn > 0.0) lu = nu * log(pn);                // unique rows: z = 1, model.py:291-317
    }
    // the other buffer is free (every peer has signalled this iteration): clear it for the next pass
    if (j < d.Kp + SS_DN_TAIL) d.T[cur ^ 1][j] = 0.0;
    const double dsum = block_sum<THREADS>(dj, s_red, 0);
    const double lsum = block_sum<THREADS>(lu, s_red, 1);
    if (tid == 0) {
      d.part[chunk] = dsum;
      d.part[d.nchunks + chunk] = lsum;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ double dense_part_sum(const DenseArgs& d, int which) {
  // lane-strided loads (in flight together) + xor tree: a fixed order, identical in every warp, CTA and rank
  double v = 0.0;
  for (int c = threadIdx.x & 31; c < d.nchunks; c += 32) v += __ldcg(d.part + which * d.nchunks + c);
  return warp_sum(v);
}

// cross-rank barrier + visibility of the exchange buffers; called by every thread of the grid after a grid.sync
__device__ __forceinline__ void dense_exchange(const DenseArgs& d, unsigned long long epoch, cg::grid_group& grid) {
  if (d.world <= 1 || d.summed) return;
  if (blockIdx.x == 0 && threadIdx.x < d.world) {
    __threadfence_system();
    st_release_sys_u64(d.peer_flags[threadIdx.x] + d.rank, epoch);
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(d.flags + threadIdx.x) < epoch) {
      if (clock64() - t0 > DN_SPIN_LIMIT) {
        d.ctl[2] = 1;
        break;
      }
    }
    __threadfence_system();
  }
  grid.sync();
}

// ---- peers mode: the whole fit in one cooperative kernel ---------------------------------------
template <int THREADS, bool LNL>
__global__ void __launch_bounds__(THREADS, LNL ? 1 : 2) em_dense_kernel(EmArgs a, DenseArgs d) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double s_red[64];
  __shared__ uint64_t s_bar;
  __shared__ StreamPipe s_pipe;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x;
  uint32_t phase = 0;
  uint32_t n_use = 0u;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    stream_pipe_init(&s_pipe, THREADS / 32 - STREAM_NPROD);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned long long epoch = d.epoch0;
  double lnl_prev = INFINITY, lnl = 0.0;
  int it = 0;
  bool conv = false, comm_err = false;
#define K_TICK(i)
  while (true) {
    ++it;
    const int cur = (it - 1) & 1;
    lane_stream_pass<THREADS, SS_ST_FR, SS_ST_FC>(a.L, d.tgcol, nullptr, d.n_tiles, nullptr, d.x[cur], d.T[cur], d.geom,
                                                    dsm, &s_pipe, n_use);
    K_TICK(0)
    grid.sync();
    K_TICK(1)
    dense_exchange(d, ++epoch, grid);
    dense_update<THREADS>(d, cur, it, s_red);
    K_TICK(2)
    grid.sync();
    K_TICK(3)
    const double diff = dense_part_sum(d, 0);
    K_TICK(4)
    if (LNL) {
      // lnL with the E-step's z (old x) and the new parameters: model.py:344
      dense_lnl_pass<THREADS>(a, d, d.x[cur], d.x[cur ^ 1], d.T[cur] + d.Kp, dsm, &s_bar, phase, s_red);
      grid.sync();
      dense_exchange(d, ++epoch, grid);
      double la = 0.0;
      if (d.world > 1 && !d.summed) {
        for (int r = 0; r < d.world; ++r) la += ld_sys_f64(d.peerT[cur][r] + d.Kp);
      } else {
        la = __ldcg(d.T[cur] + d.Kp);
      }
      lnl = la + dense_part_sum(d, 1) + d.logq_uniq;
      conv = fabs(lnl - lnl_prev) < d.eps;                     // model.py:345-347
      lnl_prev = lnl;
      if (blockIdx.x == 0 && tid == 0) d.lnls[it - 1] = lnl;
    } else {
      conv = diff < d.eps;                                      // model.py:351
    }
    if (blockIdx.x == 0 && tid == 0) d.diffs[it - 1] = diff;
    comm_err = *((volatile int*)(d.ctl + 2)) != 0;
    if (conv || it >= d.max_iter || comm_err) break;
  }
  const int cur = (it - 1) & 1;      // x[cur]: last E-step, x[cur ^ 1]: final parameters
  if (!LNL && !comm_err) {
    // final lnL (model.py:368-370); the buffer of the other parity was cleared by the last update
    dense_lnl_pass<THREADS>(a, d, d.x[cur], d.x[cur ^ 1], d.T[cur ^ 1] + d.Kp, dsm, &s_bar, phase, s_red);
    grid.sync();
    dense_exchange(d, ++epoch, grid);
    double la = 0.0;
    if (d.world > 1 && !d.summed) {
      for (int r = 0; r < d.world; ++r) la += ld_sys_f64(d.peerT[cur ^ 1][r] + d.Kp);
    } else {
      la = __ldcg(d.T[cur ^ 1] + d.Kp);
    }
    lnl = la + dense_part_sum(d, 1) + d.logq_uniq;
    comm_err = *((volatile int*)(d.ctl + 2)) != 0;
  }
  // drain: nobody leaves (and lets the host reuse its exchange buffers) while a peer may still read them
  if (!comm_err) dense_exchange(d, ++epoch, grid);
  if (blockIdx.x == 0 && tid == 0) {
    d.ctl[0] = it;
    d.ctl[1] = (conv ? 1 : 0) | (it >= d.max_iter ? 2 : 0) | (comm_err ? 8 : 0);
    *d.lnl_out = lnl;
  }
}

// ---- hook mode: the same steps as separate launches --------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2) dense_pass_kernel(EmArgs a, DenseArgs d, int cur, int lnl_pass, int tail_buf) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double s_red[64];
  __shared__ uint64_t s_bar;
  __shared__ StreamPipe s_pipe;
  if (*((volatile int*)d.ctl) != 0 && !lnl_pass) return;       // the fit has finished: no-op
  uint32_t phase = 0;
  uint32_t n_use = 0u;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    stream_pipe_init(&s_pipe, THREADS / 32 - STREAM_NPROD);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (lnl_pass)
    dense_lnl_pass<THREADS>(a, d, d.x[cur], d.x[cur ^ 1], d.T[tail_buf] + d.Kp, dsm, &s_bar, phase, s_red);
  else
    lane_stream_pass<THREADS, SS_ST_FR, SS_ST_FC>(a.L, d.tgcol, nullptr, d.n_tiles, nullptr, d.x[cur], d.T[cur], d.geom,
                                                    dsm, &s_pipe, n_use);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) dense_update_kernel(DenseArgs d, int cur, int it) {
  __shared__ double s_red[64];
  if (*((volatile int*)d.ctl) != 0) return;
  dense_update<THREADS>(d, cur, it, s_red);
}

// one thread: convergence decision of iteration `it` (lnl_mode: T[tail_buf] tail holds the summed lnL part)
__global__ void dense_decide_kernel(DenseArgs d, int it, int lnl_mode, int tail_buf, double* lnl_prev) {
  // launched with one warp: the partial sums are reduced by all 32 lanes, lane 0 decides
  const double diff = dense_part_sum(d, 0);
  const double lu = dense_part_sum(d, 1);
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (d.ctl[0] != 0) return;
  bool conv;
  if (lnl_mode) {
    const double lnl = d.T[tail_buf][d.Kp] + lu + d.logq_uniq;
    conv = fabs(lnl - *lnl_prev) < d.eps;
    *lnl_prev = lnl;
    d.lnls[it - 1] = lnl;
    *d.lnl_out = lnl;
  } else {
    conv = diff < d.eps;
  }
  d.diffs[it - 1] = diff;
  if (conv || it >= d.max_iter) {
    d.ctl[0] = it;
    d.ctl[1] = (conv ? 1 : 0) | (it >= d.max_iter ? 2 : 0);
  }
}
__global__ void dense_final_lnl_kernel(DenseArgs d, int tail_buf) {
  const double lu = dense_part_sum(d, 1);
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  *d.lnl_out = d.T[tail_buf][d.Kp] + lu + d.logq_uniq;
}

// ---- preparation -------------------------------------------------------------------------------
// scatter the per-slot statistics of the local layout into dense K-vectors: buf = [pisum0 | nuniq | active | tail]
__global__ void dense_scatter_kernel(Layout L, int K, double* __restrict__ buf) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L.n_segcols) {
    const int g = L.sc_gcol[i];
    buf[g] = L.sc_pisum0[i];
    buf[K + g] = (double)L.sc_nuniq[i];
    buf[2 * K + g] = 1.0;
  }
  if (i == 0) {
    double* tail = buf + 3LL * K;
    tail[0] = L.S > 0 ? L.seg_wamb[0] : 0.0;
    tail[1] = L.S > 0 ? L.seg_wtot[0] : 0.0;
    tail[2] = L.S > 0 ? L.seg_logq_uniq[0] : 0.0;
    tail[3] = L.S > 0 ? (double)L.seg_nobs[0] : 0.0;
    tail[4] = L.S > 0 ? (double)L.seg_A[0] : 0.0;
    tail[5] = L.S > 0 ? (double)L.seg_Aamb[0] : 0.0;
  }
}
__global__ void dense_tgcol_kernel(Layout L, int* __restrict__ tgcol) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L.tot_col) tgcol[i] = L.sc_gcol[L.c_scol[i]];      // padding slots point at slot 0: harmless
}
// nfeats = number of loci with an alignment on any rank; nobs (statistics.py:289-292)
__global__ void dense_counts_kernel(int K, const double* __restrict__ buf, int* __restrict__ out) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int c = 0;
  for (int j = threadIdx.x; j < K; j += blockDim.x) c += buf[2LL * K + j] > 0.0;
  atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) {
    out[0] = (int)(buf[3LL * K + 3] + 0.5);   // nobs
    out[1] = s_cnt;                           // nfeats
  }
}
__global__ void dense_init_kernel(DenseArgs d) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const double invK = 1.0 / (double)d.K;
  if (j < d.K) {
    d.pi[j] = invK;
    d.theta[j] = invK;
    d.x[0][j] = invK * invK;
    d.x[1][j] = invK * invK;
  }
  if (j < d.Kp + SS_DN_TAIL) {
    d.T[0][j] = 0.0;
    d.T[1][j] = 0.0;
  }
  if (j < 4) d.ctl[j] = 0;
}
// results back into the per-slot arrays the rest of the library reads (emit z, params, results)
__global__ void dense_publish_kernel(Layout L, FitState F, DenseArgs d, int max_iter) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int it = d.ctl[0];
  const int cur = (it - 1) & 1;
  if (i < L.n_segcols) {
    const int g = L.sc_gcol[i];
    F.sc_pi[i] = d.pi[g];
    F.sc_theta[i] = d.theta[g];
    F.sc_ptprev[i] = d.x[cur][g];
    F.sc_ptfin[i] = d.x[cur ^ 1][g];
  }
  if (i == 0 && L.S > 0) {
    F.seg_iters[0] = it;
    F.seg_flags[0] = d.ctl[1];
    F.seg_lnl[0] = *d.lnl_out;
    F.seg_thpw[0] = d.thpw;
    F.seg_pipw[0] = d.pipw;
    F.seg_inv_thd[0] = d.inv_thd;
    F.seg_inv_pid[0] = d.inv_pid;
  }
  for (long long k = i; k < max_iter; k += (long long)gridDim.x * blockDim.x) {
    F.seg_diffs[k] = k < it ? d.diffs[k] : 0.0;
    if (F.seg_lnls) F.seg_lnls[k] = k < it ? d.lnls[k] : 0.0;
  }
}

// int <-> double for the MAX all-reduce of s_max at build time (model.py:115: max over the WHOLE matrix)
__global__ void smax_to_double_kernel(const int* in, double* out) { out[0] = (double)in[0]; }
__global__ void smax_from_double_kernel(const double* in, int* out) { out[0] = (int)(in[0] + 0.5); }

int dense_allreduce_smax(ss_handle_s* h, cudaStream_t st) {
  if (!h->allreduce || h->L.S != 1) return SS_OK;
  double* tmp = nullptr;
  SS_CUDA_CHECK(h, cudaMallocAsync(reinterpret_cast<void**>(&tmp), sizeof(double), st));
  smax_to_double_kernel<<<1, 1, 0, st>>>(h->L.seg_smax, tmp);
  h->launches++;
  const int rc = h->allreduce(h->allreduce_user, (void*)st, tmp, 1, 1);
  if (rc != 0) {
    h->set_error("all-reduce hook failed (s_max)");
    return SS_ERR_CUDA;
  }
  smax_from_double_kernel<<<1, 1, 0, st>>>(tmp, h->L.seg_smax);
  h->launches++;
  SS_CUDA_CHECK(h, cudaFreeAsync(tmp, st));
  return SS_OK;
}

void free_dense(ss_handle_s* h) {
  dev_free_all(h->dense_allocs);
  h->dn_ready = false;
}

// once per layout: dense statistics summed over the ranks, tile column -> locus map
static int dense_prepare(ss_handle_s* h, cudaStream_t st) {
  Layout& L = h->L;
  const int K = L.K;
  auto& pool = h->dense_allocs;
  free_dense(h);
  const size_t nstat = 3 * (size_t)K + SS_DN_TAIL;
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->dn_stat, nstat));
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->dn_tgcol, (size_t)std::max<long long>(L.tot_col, 1)));
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->dn_pi, (size_t)K));
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->dn_theta, (size_t)K));
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->dn_x[0], (size_t)K));
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->dn_x[1], (size_t)K));
  const int nchunks = (K + SS_DN_TAIL + 64 + DN_THREADS - 1) / DN_THREADS;
  h->dn_nchunks = nchunks;
  h->dn_Kp = nchunks * DN_THREADS - SS_DN_TAIL;
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->dn_part, 2 * (size_t)nchunks));
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->dn_ctl, 4));
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->dn_counts, 2));
  SS_CUDA_CHECK(h, dev_alloc(pool, &h->dn_lnl, 2));
  if (!h->dn_peers) {
    // exchange buffers in ordinary device memory (hook mode / single rank)
    SS_CUDA_CHECK(h, dev_alloc(pool, &h->dn_T[0], (size_t)h->dn_Kp + SS_DN_TAIL));
    SS_CUDA_CHECK(h, dev_alloc(pool, &h->dn_T[1], (size_t)h->dn_Kp + SS_DN_TAIL));
  } else {
    if ((size_t)(2 * (h->dn_Kp + SS_DN_TAIL) + SS_DN_MAXW) * 8 > h->dn_peer_bytes) {
      h->set_error("ss_set_peers: exchange buffer too small for this number of loci");
      return SS_ERR_ARG;
    }
    double* base = reinterpret_cast<double*>(h->dn_peer_base[h->dn_rank]);
    h->dn_T[0] = base;
    h->dn_T[1] = base + h->dn_Kp + SS_DN_TAIL;
  }
  SS_CUDA_CHECK(h, cudaMemsetAsync(h->dn_stat, 0, nstat * sizeof(double), st));
  const long long n = std::max<long long>(L.n_segcols, 1);
  dense_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L, K, h->dn_stat);
  h->launches++;
  if (L.tot_col > 0) {
    dense_tgcol_kernel<<<(unsigned)((L.tot_col + 255) / 256), 256, 0, st>>>(L, h->dn_tgcol);
    h->launches++;
  }
  if (h->allreduce && h->dn_world > 1) {
    const int rc = h->allreduce(h->allreduce_user, (void*)st, h->dn_stat, (long long)nstat, 0);
    if (rc != 0) {
      h->set_error("all-reduce hook failed (model statistics)");
      return SS_ERR_CUDA;
    }
  } else if (h->dn_world > 1) {
    h->set_error("row-sharded fit over several ranks needs the all-reduce hook (ss_set_allreduce) for the build-time sums");
    return SS_ERR_STATE;
  }
  dense_counts_kernel<<<1, 256, 0, st>>>(K, h->dn_stat, h->dn_counts);
  h->launches++;
  SS_CUDA_CHECK(h, cudaMemcpyAsync(h->dn_tail_host, h->dn_stat + 3LL * K, SS_DN_TAIL * sizeof(double),
                                   cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  SS_CUDA_CHECK(h, cudaGetLastError());
  h->dn_ready = true;
  return SS_OK;
}

int em_fit_dense(ss_handle_s* h, cudaStream_t st, const FitParams& P) {
  Layout& L = h->L;
  FitState& F = h->F;
  if (L.S != 1) {
    h->set_error("row-sharded (dense) fit needs exactly one model; shard models over ranks instead");
    return SS_ERR_STATE;
  }
  if (L.K > 1 << 24) {
    h->set_error("row-sharded fit: too many loci");
    return SS_ERR_CAPACITY;
  }
  if (!h->dn_ready) {
    const int rc = dense_prepare(h, st);
    if (rc != SS_OK) return rc;
  }
  const int K = L.K;
  // model constants from the GLOBAL weight sums (model.py:189-194)
  const double wamb = h->dn_tail_host[0], wtot = h->dn_tail_host[1];
  double wmax_h = 0.0;
  SS_CUDA_CHECK(h, cudaMemcpyAsync(&wmax_h, L.seg_wmax, sizeof(double), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
  DenseArgs d;
  memset(&d, 0, sizeof(d));
  d.K = K; d.Kp = h->dn_Kp; d.world = h->dn_world; d.rank = h->dn_rank; d.n_tiles = L.n_tiles;
  d.nchunks = h->dn_nchunks; d.max_iter = P.max_iter; d.eps = P.eps;
  d.summed = h->dn_peers ? 0 : 1;       // no peer buffers: T already is the sum over ranks when the update reads it
  d.T[0] = h->dn_T[0]; d.T[1] = h->dn_T[1];
  for (int r = 0; r < SS_DN_MAXW; ++r) {
    if (h->dn_peers && r < h->dn_world) {
      double* base = reinterpret_cast<double*>(h->dn_peer_base[r]);
      d.peerT[0][r] = base;
      d.peerT[1][r] = base + h->dn_Kp + SS_DN_TAIL;
      d.peer_flags[r] = reinterpret_cast<unsigned long long*>(base + 2 * (h->dn_Kp + SS_DN_TAIL));
    } else {
      d.peerT[0][r] = d.peerT[1][r] = nullptr;
      d.peer_flags[r] = nullptr;
    }
  }
  d.flags = h->dn_peers ? d.peer_flags[h->dn_rank] : nullptr;
  d.epoch0 = h->dn_epoch;
  d.pisum0 = h->dn_stat; d.nuniq = h->dn_stat + K;
  d.pi = h->dn_pi; d.theta = h->dn_theta; d.x[0] = h->dn_x[0]; d.x[1] = h->dn_x[1];
  d.part = h->dn_part; d.tgcol = h->dn_tgcol;
  d.thpw = P.theta_prior * wmax_h;
  d.pipw = P.pi_prior * wmax_h;
  const double thd = wamb + d.thpw * (double)K, pid = wtot + d.pipw * (double)K;
  d.logq_uniq = h->dn_tail_host[2];
  d.ctl = h->dn_ctl; d.diffs = F.seg_diffs; d.lnls = F.seg_lnls; d.lnl_out = h->dn_lnl;
  if (!(thd > 0.0) || !(pid > 0.0)) {
    // 0/0 in the M-step: FloatingPointError -> StellarscopeError in the reference (model.py:267-272)
    int flag = 4;
    SS_CUDA_CHECK(h, cudaMemcpyAsync(F.seg_flags, &flag, sizeof(int), cudaMemcpyHostToDevice, st));
    SS_CUDA_CHECK(h, cudaMemsetAsync(F.seg_iters, 0, sizeof(int), st));
    SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
    h->has_fit = true;
    return SS_OK;
  }
  d.inv_thd = 1.0 / thd;
  d.inv_pid = 1.0 / pid;

  EmArgs a;
  a.L = L; a.F = F; a.P = h->P; a.z_out = nullptr; a.tile_list = nullptr; a.n_list = 0;
  const bool lnl = P.use_likelihood != 0;
  int occ_pick = 0;
  const void* pick_fn = h->dn_persistent ? (lnl ? (const void*)em_dense_kernel<DN_THREADS, true> : (const void*)em_dense_kernel<DN_THREADS, false>)
                                         : (const void*)dense_pass_kernel<DN_THREADS>;
  // the final lnL pass (every mode) carves a whole tile out of the same dynamic shared memory (h->all_smem)
  const size_t smem = std::max<size_t>(pick_stream_geom(h->geom_all, h->max_smem_optin - 2048, h->all_smem,
                                                        pick_fn, DN_THREADS, &d.geom, &occ_pick), 1024);
  if (occ_pick < 1) {
    h->set_error("dense kernel does not fit on an SM");
    return SS_ERR_CAPACITY;
  }
  const int nK = d.Kp + SS_DN_TAIL;
  dense_init_kernel<<<(nK + 255) / 256, 256, 0, st>>>(d);
  h->launches++;
  if (h->timing) SS_CUDA_CHECK(h, cudaEventRecord(h->ev[1], st));
  if (h->dn_persistent) {
    // ---- one persistent cooperative kernel ----
    void* fn = lnl ? (void*)em_dense_kernel<DN_THREADS, true> : (void*)em_dense_kernel<DN_THREADS, false>;
    SS_CUDA_CHECK(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SS_CUDA_CHECK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, DN_THREADS, smem));
    if (occ < 1) {
      h->set_error("dense kernel does not fit on an SM");
      return SS_ERR_CAPACITY;
    }
    const int grid = std::max(1, std::min(std::max(L.n_tiles, d.nchunks), occ * h->num_sms));
    void* args[] = {(void*)&a, (void*)&d};
    SS_CUDA_CHECK(h, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(DN_THREADS), args, smem, st));
    h->launches++;
    h->dn_epoch += 2ull * (unsigned long long)P.max_iter + 4ull;
  } else {
    // ---- hook mode: host loop, one all-reduce (caller's NCCL) per iteration ----
    SS_CUDA_CHECK(h, cudaFuncSetAttribute(dense_pass_kernel<DN_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
    const bool multi = h->allreduce && h->dn_world > 1;
    const int grid_t = std::max(1, std::min(L.n_tiles, h->num_sms));   // persistent CTAs: the next tile is prefetched
    double* lnl_prev = h->dn_lnl + 1;
    const double inf = INFINITY;
    SS_CUDA_CHECK(h, cudaMemcpyAsync(lnl_prev, &inf, sizeof(double), cudaMemcpyHostToDevice, st));
    int it_done = 0, it = 0;
    while (it_done == 0) {
      ++it;
      const int cur = (it - 1) & 1;
      if (L.n_tiles > 0) {
        dense_pass_kernel<DN_THREADS><<<grid_t, DN_THREADS, smem, st>>>(a, d, cur, 0, 0);
        h->launches++;
      }
      if (multi && h->allreduce(h->allreduce_user, (void*)st, d.T[cur], (long long)nK, 0) != 0) {
        h->set_error("all-reduce hook failed (theta sums)");
        return SS_ERR_CUDA;
      }
      dense_update_kernel<DN_THREADS><<<d.nchunks, DN_THREADS, 0, st>>>(d, cur, it);
      h->launches++;
      if (lnl) {
        if (L.n_tiles > 0) {
          dense_pass_kernel<DN_THREADS><<<grid_t, DN_THREADS, smem, st>>>(a, d, cur, 1, cur);
          h->launches++;
        }
        if (multi && h->allreduce(h->allreduce_user, (void*)st, d.T[cur] + d.Kp, SS_DN_TAIL, 0) != 0) {
          h->set_error("all-reduce hook failed (lnL)");
          return SS_ERR_CUDA;
        }
      }
      dense_decide_kernel<<<1, 32, 0, st>>>(d, it, lnl ? 1 : 0, cur, lnl_prev);
      h->launches++;
      SS_CUDA_CHECK(h, cudaMemcpyAsync(&it_done, d.ctl, sizeof(int), cudaMemcpyDeviceToHost, st));
      SS_CUDA_CHECK(h, cudaStreamSynchronize(st));
      if (it >= P.max_iter && it_done == 0) {
        h->set_error("dense fit: iteration bookkeeping out of step");
        return SS_ERR_STATE;
      }
    }
    if (!lnl) {
      const int cur = (it_done - 1) & 1;
      if (L.n_tiles > 0) {
        dense_pass_kernel<DN_THREADS><<<grid_t, DN_THREADS, smem, st>>>(a, d, cur, 1, cur ^ 1);
        h->launches++;
      }
      if (multi && h->allreduce(h->allreduce_user, (void*)st, d.T[cur ^ 1] + d.Kp, SS_DN_TAIL, 0) != 0) {
        h->set_error("all-reduce hook failed (final lnL)");
        return SS_ERR_CUDA;
      }
      dense_final_lnl_kernel<<<1, 32, 0, st>>>(d, cur ^ 1);
      h->launches++;
    }
  }
  if (h->timing) SS_CUDA_CHECK(h, cudaEventRecord(h->ev[2], st));
  const long long n = std::max<long long>(std::max<long long>(L.n_segcols, P.max_iter), 1);
  dense_publish_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 65535), 256, 0, st>>>(L, F, d, P.max_iter);
  h->launches++;
  SS_CUDA_CHECK(h, cudaGetLastError());
  h->has_fit = true;
  return SS_OK;
}

}  // namespace ss
