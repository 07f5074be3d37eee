/* stellarscope_b200 -- C ABI of the B200-native EM read-reassignment path.
 *
 * The reference (nixonlab/stellarscope) is pure Python and has no FFI on this
 * path; its boundary is a Python class contract.  This
 * header is the C-ABI a binding for that contract calls.  Each entry point
 * names the reference code it replaces (paths relative to the reference
 * checkout, stellarscope/utils/).
 *
 * Conventions
 *  - every function returns an ss_status (0 = ok); the message of the last
 *    failure on a handle is returned by ss_last_error().  No C++ exception
 *    crosses the boundary.
 *  - the CALLER owns every input and output buffer (device pointers unless the
 *    name ends in _host); the handle owns the tiled layout and its scratch and
 *    frees them in ss_destroy().
 *  - all work is enqueued on the caller's cudaStream_t (passed as void*) and is
 *    asynchronous unless stated; one handle per (process, GPU); a handle is
 *    not re-entrant.
 */
#ifndef STELLARSCOPE_B200_H
#define STELLARSCOPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS_ABI_VERSION 1

typedef struct ss_handle_s* ss_handle_t;

enum ss_status {
  SS_OK = 0,
  SS_ERR_ARG = 1,      /* bad argument (ValueError in the reference, model.py:424-426) */
  SS_ERR_CUDA = 2,     /* CUDA runtime failure */
  SS_ERR_STATE = 3,    /* call order (no layout / no fit yet) */
  SS_ERR_MODEL = 4,    /* StellarscopeError class: empty model, zero denominator (model.py:267-272) */
  SS_ERR_CAPACITY = 5, /* a read has more alignments than a tile can hold */
  SS_ERR_NOT_CANONICAL = 6 /* ss_layout_build: zero scores / unsorted or duplicate column indices; the
                              reference canonicalises on construction (csr_matrix + eliminate_zeros,
                              model.py:113-114) -- the caller does that on the host and retries */
};

/* reassignment modes, order of TelescopeLikelihood.REASSIGN_MODES (model.py:92-100) */
enum ss_reassign_mode {
  SS_BEST_EXCLUDE = 0,
  SS_BEST_CONF = 1,
  SS_BEST_RANDOM = 2,
  SS_BEST_AVERAGE = 3,
  SS_INITIAL_UNIQUE = 4,
  SS_INITIAL_RANDOM = 5,
  SS_TOTAL_HITS = 6
};

/* sizes of the layout held by a handle */
typedef struct ss_layout_info {
  int32_t n_rows, n_cols, n_segments, n_tiles, max_row_len, chunk;
  int32_t n_unique_rows, n_resident_tiles, n_stream_tiles, n_stream_segments;
  int32_t n_cluster_tiles, n_cluster_segments;
  int64_t nnz, n_segcols, tot_ent, tot_row, tot_col, tot_jds;
  int64_t nnz_ambiguous, rows_ambiguous;
} ss_layout_info;

/* host-side description of a pre-built layout (debug / test import only;
 * format specification: stellarscope_b200/layout_host.py) */
typedef struct ss_host_layout {
  int32_t N, K, S, n_tiles, maxlen, chunk, n_uniq;
  int64_t A, n_segcols, tot_ent, tot_row, tot_col, tot_jds;
  const int64_t* tile_hdr;
  const uint16_t* e_lcol; const uint16_t* e_cpos; const double* e_q; const int32_t* e_opos;
  const uint16_t* r_len; const double* r_w; const int32_t* r_orow;
  const uint16_t* c_ptr; const int32_t* c_scol; const double* c_pisum0; const int32_t* c_nuniq;
  const uint16_t* j_ptr;
  const int32_t* seg_tile0; const int32_t* seg_ntiles; const int32_t* seg_col0; const int32_t* seg_ncol;
  const int32_t* seg_smax; const int32_t* seg_nobs; const int32_t* seg_namb; const int32_t* seg_nuniq;
  const double* seg_wmax; const double* seg_wtot; const double* seg_wamb; const double* seg_logq_uniq;
  const int64_t* seg_A; const int64_t* seg_Aamb;
  const int32_t* sc_gcol; const double* sc_pisum0; const int32_t* sc_nuniq; const int8_t* sc_inamb;
  const int32_t* u_row; const int32_t* u_pos;
} ss_host_layout;

int ss_abi_version(void);

/* one handle per (process, GPU) */
int ss_create(int device, ss_handle_t* out);
int ss_destroy(ss_handle_t h);
/* drop the layout, the fit state and their device memory; the handle (streams, events) stays usable */
int ss_release(ss_handle_t h);
const char* ss_last_error(ss_handle_t h);

/* Build the tiled segment layout on the device.
 * Replaces TelescopeLikelihood.__init__ (model.py:102-200: Q = expm1(100*s/s_max),
 * Y_amb/Y_uni/Y_0, row weights, weight sums, pisum0) for EVERY model of a pooling
 * mode at once, and the row selection of _fit_pooling_model (model.py:701-743).
 *   row_ptr[n_rows+1], col_idx[nnz] (ascending within a row), score[nnz] (> 0):
 *       the N x K uint16 CSR score matrix (st.corrected / st.raw_scores, model.py:749-758)
 *   row_segment[n_rows]: model index of each row (cell, celltype, 0 for
 *       pseudobulk), -1 = row belongs to no model.
 * The matrix must be canonical (checked on the device in the same pass: SS_ERR_NOT_CANONICAL).
 * Synchronises the stream (sizes of the layout are data dependent). */
int ss_layout_build(ss_handle_t h, void* stream, int32_t n_rows, int32_t n_cols, int64_t nnz,
                    const int32_t* row_ptr, const int32_t* col_idx, const uint16_t* score,
                    const int32_t* row_segment, int32_t n_segments);

/* Debug / test: install a layout that was built on the host. Synchronous. */
int ss_layout_import(ss_handle_t h, void* stream, const ss_host_layout* L);

int ss_layout_get_info(ss_handle_t h, ss_layout_info* out);

/* Copy one named layout array to a caller host buffer (tests: device builder
 * vs format specification).  Returns the byte size in *bytes when dst_host is
 * NULL. Synchronous. */
int ss_layout_export(ss_handle_t h, const char* name, void* dst_host, int64_t* bytes);

/* Run EM to convergence for every segment.
 * Replaces TelescopeLikelihood.em (model.py:325-373) = estep (:202-226) + mstep
 * (:229-272) + convergence (:340-355) + calculate_lnl (:276-323), for all models
 * of _fit_pooling_model (:785-801) in one call. */
int ss_em_fit(ss_handle_t h, void* stream, double em_epsilon, int32_t max_iter, double pi_prior,
              double theta_prior, int32_t use_likelihood);

/* Per-segment results of the last fit (any pointer may be NULL):
 *   iters[S]  number of iterations (len(FitInfo.iterations), statistics.py:280)
 *   flags[S]  bit0 converged, bit1 reached_max, bit2 model error (SS_ERR_MODEL class),
 *             bit3 a peer did not answer (row-sharded fit)
 *   lnl[S]    FitInfo.final_lnl (model.py:368-370)
 *   diffs[S*max_iter] diff_est trace (model.py:340, 359), unused slots = 0
 *   lnls[S*max_iter]  lnL trace (only with use_likelihood)
 *   nobs[S], nfeats[S]  FitInfo.nobs / nfeats (statistics.py:289-292) */
int ss_em_get_results(ss_handle_t h, void* stream, int32_t* iters, int32_t* flags, double* lnl,
                      double* diffs, double* lnls, int32_t* nobs, int32_t* nfeats);

/* Per segment-column parameters of the last fit: for segment s the slots
 * seg_col0[s] .. seg_col0[s]+seg_ncol[s]-1 hold the loci gcol[] (ascending)
 * the model touches, their pi and theta; every other locus of that model has
 * pi = pi_rest[s], theta = theta_rest[s].  (tl.pi / tl.theta, model.py:357) */
int ss_em_get_params(ss_handle_t h, void* stream, int32_t* seg_col0, int32_t* seg_ncol, int32_t* gcol,
                     double* pi, double* theta, double* pi_rest, double* theta_rest);

/* Posterior matrix z of the last E-step in the CSR order of the input matrix
 * (tl.z, model.py:356; ret_model.z += _z, :790).  Rows outside every model and
 * empty rows get nothing (their entries do not exist). */
int ss_em_emit_z(ss_handle_t h, void* stream, double* z /* nnz */);

/* Reassignment pass (TelescopeLikelihood.reassign, model.py:375-507; binmax
 * sparse_plus.py:134-178).  Works on the CSR order of the input matrix.
 *   z        posterior (nnz) -- from ss_em_emit_z; ignored by the initial_* / total_hits modes
 *   flag[nnz] out: 1 where the alignment is selected by the mode (for the two
 *            *_random modes: 1 on every tied-best alignment; the host replays
 *            numpy's PCG64 draws, sparse_plus.py:233-238)
 *   nbest[n_rows] out: number of tied-best alignments of the row (0 for rows without)
 *   info[4 + 64] out: assigned, ambiguous, unaligned, (reserved), histogram of
 *            min(nbest, 63)  (ReassignInfo, statistics.py:410-462; model.py:437-440)
 *   tie_tol  relative tolerance of "equal to the row maximum"; 0 = exact */
int ss_reassign(ss_handle_t h, void* stream, int32_t mode, double conf_thresh, double tie_tol,
                int32_t n_rows, const int32_t* row_ptr, const uint16_t* score, const double* z,
                uint8_t* flag, int32_t* nbest, int64_t* info);

/* The reassignment matrix itself as CSR (the first return value of TelescopeLikelihood.reassign,
 * model.py:375-507; stored in Stellarscope.reassignments[mode] :1254): compaction of the flagged alignments.
 *   flag[nnz], nbest[n_rows]: outputs of ss_reassign (nbest may be NULL)
 *   out_ptr[n_rows + 1] (int64), out_idx[cap]: row pointers / column of every selected alignment, row order kept
 *   out_val[cap] or NULL: 1 / nbest of the row (the fractions of best_average, model.py:448)
 * *n_out = number of selected alignments (host value; SS_ERR_CAPACITY if it exceeds cap).  Synchronises. */
int ss_reassign_csr(ss_handle_t h, void* stream, int32_t n_rows, const int32_t* row_ptr, const int32_t* col_idx,
                    const uint8_t* flag, const int32_t* nbest, int64_t cap, int64_t* out_ptr, int32_t* out_idx,
                    double* out_val, int64_t* n_out);

/* Count aggregation (Stellarscope.output_report agg_bc / agg_bc_umi,
 * model.py:1302-1345; groupby_sum sparse_plus.py:367-396).
 *   flag/weight: the reassignment matrix in CSR order (weight NULL = 1 per flagged alignment)
 *   row_cell[n_rows]: output barcode index of the row, -1 = drop
 *   row_umi[n_rows] or NULL: UMI id of the row (umi_counts mode)
 * Writes up to cap COO triples (cell, locus, value) sorted by (cell, locus);
 * *n_out receives the number produced (host value, call synchronises). */
int ss_count(ss_handle_t h, void* stream, int32_t n_rows, int32_t n_cols, const int32_t* row_ptr,
             const int32_t* col_idx, const uint8_t* flag, const double* weight, const int32_t* row_cell,
             const int32_t* row_umi, int64_t cap, int32_t* out_cell, int32_t* out_col, double* out_val,
             int64_t* n_out);

/* The random modes (best_random model.py:446, initial_random :482-491; csr_matrix_plus.choose_random
 * sparse_plus.py:199-240): in row tied_rows[i] -- a row in which ss_reassign flagged several tied-best alignments --
 * keep only the picks[i]-th flagged alignment (0-based, in CSR order).  The picks are drawn by the caller (the
 * reference's numpy generator, in row order); flag is updated in place.  Asynchronous. */
int ss_reassign_pick(ss_handle_t h, void* stream, int32_t n_tied, const int32_t* tied_rows, const int32_t* picks,
                     const int32_t* row_ptr, uint8_t* flag);

/* Per-cell counts without a sort (agg_bc, model.py:1302-1320, for every reassignment mode without --umi_counts):
 * the (loci x cells) count matrix in CSC form.  One CTA per cell accumulates the cell's selected alignments in
 * shared-memory bins over the loci.
 *   row_cell[n_rows]: output column of the row (0 .. n_cells-1; anything else: row dropped)
 *   out_ptr[n_cells + 1] out: column pointers; out_col[cap] out: loci, ascending per cell
 *   integer modes: out_cnt[cap] (int32 counts), out_val = NULL, nbest ignored
 *   best_average:  out_val[cap] (double), out_cnt = NULL, nbest[n_rows] from ss_reassign: every selected alignment
 *                  of row r adds 1 / nbest[r] (model.py:448)
 *   *n_out: entries produced (host value, call synchronises).  SS_ERR_CAPACITY: cap too small (*n_out = needed)
 *   or n_cols too large for the shared-memory bins (use ss_count). */
int ss_count_cells(ss_handle_t h, void* stream, int32_t n_rows, int32_t n_cols, int32_t n_cells, const int32_t* row_ptr,
                   const int32_t* col_idx, const uint8_t* flag, const int32_t* nbest, const int32_t* row_cell, int64_t cap,
                   int64_t* out_ptr, int32_t* out_col, int32_t* out_cnt, double* out_val, int64_t* n_out);

/* UMI de-duplication, the step before the path (Stellarscope.dedup_umi model.py:1148-1219,
 * select_umi_representatives :561-668).  Works on the CSR score matrix (raw_scores).
 *   row_group[n_rows]: id of the barcode+UMI pair of the row (0 .. n_groups-1), -1 = row takes no part
 *   excluded[n_rows] out: 1 for the duplicate reads that are NOT the representative of their connected
 *            component (reads of one pair that share a feature != column 0 are connected; the
 *            representative has the largest (max/sum, -n_features, max, sum, row), model.py:578-601)
 *   comp_root[n_rows] out: smallest row of the read's component, -1 for reads of single-read pairs
 *   group_size[n_groups], group_ncomp[n_groups] out: reads / connected components of every pair
 *            (UMIInfo.rpu_counter, UMIInfo.ncomps_umi, statistics.py:498-570)
 * Synchronises the stream. */
int ss_umi_dedup(ss_handle_t h, void* stream, int32_t n_rows, const int32_t* row_ptr, const int32_t* col_idx,
                 const uint16_t* score, const int32_t* row_group, int32_t n_groups, uint8_t* excluded,
                 int32_t* comp_root, int32_t* group_size, int32_t* group_ncomp);

/* Score-matrix construction, the producer of the path's input (mapping_to_matrix inside
 * Stellarscope.load_alignment, model.py:958-978): one (read, feature, alnscore, alnlen) record per
 * fragment x locus mapping (read = read_index[query_name], feat = feat_index[feat_name], both coded by
 * the caller); stores  max over the records of one (read, feat) of  (alnscore - min_as + 1) + alnlen
 * as uint16 and returns canonical CSR (columns ascending, no duplicates).
 *   read/feat/alnscore/alnlen[n]: device arrays
 *   row_ptr[n_reads + 1], col_idx[cap >= n], score[cap >= n] out: device arrays
 *   info[4] out (HOST): stored elements, records whose value leaves (0, 65535], records with an index
 *            outside the matrix, reads without a stored value in a feature column >= 1 (the reference
 *            asserts there are none, model.py:971-974 -- the caller raises)
 * SS_ERR_CAPACITY / SS_ERR_ARG (nothing written) when info[1] / info[2] is not 0.  Synchronises the stream. */
int ss_scores_build(ss_handle_t h, void* stream, int64_t n, const int32_t* read, const int32_t* feat,
                    const int32_t* alnscore, const int32_t* alnlen, int32_t min_as, int32_t n_reads, int32_t n_feats,
                    int32_t* row_ptr, int32_t* col_idx, uint16_t* score, int64_t* info);

/* MatrixMarket body of a count matrix: the text scipy.io.mmwrite writes for the
 * (features x barcodes) matrices at the end of Stellarscope.output_report (model.py:1418, header and
 * metadata comment :1347-1368 stay with the caller): one line "row+1 col+1 value\n" per entry, in the
 * order of the arrays.
 *   row/col[n] (int32), val[n] (double): COO triples, device arrays (the order is kept)
 *   integer != 0: values are printed as integers ("coordinate integer", int32 count matrices);
 *            0: as the correctly rounded 17-significant-digit decimal without trailing zeros ("coordinate
 *            real": best_average and --umi_counts), which parses back to the same double
 *   off[n + 1] (int64, device, caller-owned): ss_mtx_measure fills it with the byte offset of every line
 *            (off[n] = total) and returns the total in *n_bytes (HOST; synchronises the stream);
 *   out[*n_bytes] (device, 16-byte aligned): ss_mtx_write writes the text there (asynchronous).
 * SS_ERR_ARG (nothing written) for a negative index, a non-integer value with integer != 0, or a real
 * value outside 0 / 2^-64 <= |x| < 2^63 (count matrices hold integers and sums of 1/nbest). */
int ss_mtx_measure(ss_handle_t h, void* stream, int64_t n, const int32_t* row, const int32_t* col, const double* val,
                   int32_t integer, int64_t* off, int64_t* n_bytes);
int ss_mtx_write(ss_handle_t h, void* stream, int64_t n, const int32_t* row, const int32_t* col, const double* val,
                 int32_t integer, const int64_t* off, uint8_t* out);

/* ---- one model, rows sharded over the GPUs of a box (pooling_mode=pseudobulk) ----
 * The reference fits the pseudobulk model single-threaded (model.py:762-767).  Here every rank
 * (one process per GPU) holds a block of rows of the SAME model; ss_layout_build / ss_em_fit are
 * called on every rank with n_segments = 1 and the library runs the exchange step of the path:
 * the sum over ranks of the theta sums (model.py:239) once per iteration, plus build-time sums
 * (s_max MAX, model.py:115; pisum0, weights, model.py:183-194).
 *
 * ss_set_allreduce: hook that all-reduces `count` doubles in place on `stream` (op 0 = sum,
 *   1 = max) -- the caller's NCCL communicator.  Used for the build-time sums and, when no peer
 *   buffers are installed, once per iteration from a host loop.  NULL removes the hook.
 * ss_set_peers: rank / world of the sharded fit and (optionally) peer-mapped exchange buffers:
 *   peer_bufs[r] = address, valid on THIS device, of rank r's buffer of `bytes` bytes
 *   (>= ss_peer_buffer_bytes(K), zero-filled, e.g. torch symmetric memory over NVLink).  With
 *   peer buffers the whole fit runs in one persistent kernel per rank that exchanges the partial
 *   sums over peer memory itself; every rank adds them in rank order, so all ranks take the
 *   identical convergence decision.  world = 1 with peer_bufs = NULL selects the same persistent
 *   kernel on a single GPU (bytes < 0: the host-loop variant).  The caller must keep the ranks in step (barrier) before each fit. */
typedef int (*ss_allreduce_fn)(void* user, void* stream, double* buf, int64_t count, int32_t op /*0 sum, 1 max*/);
int ss_set_allreduce(ss_handle_t h, ss_allreduce_fn fn, void* user);
int64_t ss_peer_buffer_bytes(int32_t n_cols);
int ss_set_peers(ss_handle_t h, int32_t rank, int32_t world, void* const* peer_bufs, int64_t bytes);
/* pi / theta of the row-sharded model over all K loci (tl.pi, tl.theta, model.py:357) */
int ss_em_get_dense_params(ss_handle_t h, void* stream, double* pi, double* theta);

/* number of kernel launches issued by this handle since creation */
int64_t ss_launch_count(ss_handle_t h);

/* Measurement aid (bench.py): when enabled, ss_em_fit brackets its kernel groups with CUDA
 * events; ss_get_timings synchronises on them and returns the device time in ms of the last
 * fit: [0] prepare, [1] all EM kernels (they run concurrently), [2] unused, [3] emit,
 * [4] finalize, then the span of each concurrent EM kernel: [5 .. 5+C-1] the lane kernel of
 * size class 0..C-1 (C = ss_lane_classes), [5+C] the streaming kernel, [5+C+1 .. 5+C+7] the
 * cluster kernels for segments of 2..8 tiles.  n >= 5 + C + 8.
 * ss_lane_classes: CTA sizes (threads) of the lane-kernel size classes; returns C. */
int ss_lane_classes(int32_t* threads, int32_t cap);
int ss_set_timing(ss_handle_t h, int32_t enabled);
int ss_get_timings(ss_handle_t h, double* ms, int32_t n);

#ifdef __cplusplus
}
#endif
#endif
