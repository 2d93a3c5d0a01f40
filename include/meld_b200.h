/*
 * meld_b200.h -- C-ABI of libmeld_b200.so, the sm_100a engine behind
 * meld.MELD().fit_transform / meld.utils.normalize_densities.
 *
 * The reference (KrishnaswamyLab/MELD) has no FFI of its own: its seam is the
 * Python API plus the duck-typed graph protocol used by meld/filter.py
 * (graph.estimate_lmax(), graph.lmax, graph.L, graph.N).  Each entry point below
 * cites the reference call it stands in for; INTEGRATION.md shows the ctypes stub
 * a maintainer would add to the reference.
 *
 * Conventions
 *  - Every function returns 0 on success or a negative meld_b200_status; the text
 *    of the last failure on the calling thread is meld_b200_last_error().
 *  - All array arguments are DEVICE pointers unless the name ends in `_host`.
 *    The caller (PyTorch, as the device-memory container) owns every buffer it
 *    passes in; the library owns only the opaque handles it creates.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    Calls are asynchronous with respect to the host unless stated otherwise.
 *  - A handle is not thread-safe.  No C++ exception crosses this boundary.
 *  - There is no CPU fallback: without a CUDA device every compute call fails
 *    with MELD_B200_ERR_CUDA.
 */
#ifndef MELD_B200_H_
#define MELD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum meld_b200_status {
  MELD_B200_OK = 0,
  MELD_B200_ERR_INVALID = -1, /* bad argument (NULL, shape, range)            */
  MELD_B200_ERR_CUDA = -2,    /* a CUDA runtime / driver call failed          */
  MELD_B200_ERR_NOMEM = -3,   /* device or host allocation failed             */
  MELD_B200_ERR_UNSUPPORTED = -4, /* outside the engine-supported subset      */
  MELD_B200_ERR_INTERNAL = -5
} meld_b200_status;

/* Opaque device graph: CSR of the combinatorial Laplacian L = D - W (float64
 * values, int32 columns) for rows [row0, row0 + n_rows) of an n_cols-wide
 * operator, plus the row-block partition the Chebyshev kernel stages by TMA. */
typedef struct meld_b200_graph meld_b200_graph_t;

/* ---- library ---------------------------------------------------------------- */
int meld_b200_version(void);              /* major*10000 + minor*100 + patch     */
const char *meld_b200_last_error(void);   /* thread-local, never NULL            */
/* Kernels this library has launched so far in this process (its own, not CUB's).  */
int64_t meld_b200_launch_count(void);
/* Times this library has made the host wait for a stream so far (cudaStreamSynchronize inside its calls).       */
int64_t meld_b200_sync_count(void);
/* sm_count / cc_major / cc_minor of the current device (host pointers).        */
int meld_b200_device_info(int *sm_count_host, int *cc_major_host, int *cc_minor_host);

/* Launch-configuration knobs for bench sweeps and tests only (the defaults are the shipped configuration; the key
 * list is in INTEGRATION.md: Chebyshev kernels blk_chunk, group, flat_*, pad_width, ctas_per_sm; graph build reorder,
 * clusters, kmeans_iters, km_var_pct, prune*, tl_*, tc_multicast, ...).  Unknown keys are an error.  Takes effect
 * for graphs created afterwards.                                                                                  */
int meld_b200_set_tuning(const char *key, int value);

/* ---- graph construction ------------------------------------------------------ */

/* Stands in for graphtools.Graph(X, knn, decay, thresh, anisotropy, use_pygsp=True)
 * as reached from MELD.fit (reference meld/meld.py:117-118, :273): exact
 * Euclidean kNN, alpha-decay kernel K_ij = exp(-(d_ij/eps_i)^decay) for every j
 * with K_ij >= thresh, (K+K^T)/2, anisotropy K_ij/(q_i q_j)^a, W = K - diag,
 * L = diag(W 1) - W.  X is (n, d) row-major float64 (the post-PCA data_nu).
 * decay = 0 stands for the reference's decay=None: the unweighted kernel K_ij = 1 for the knn nearest
 * cells of i (self included), thresh and bandwidth_scale ignored (graphtools kNNGraph, binary branch).
 * flags: bit0 keep the un-symmetrised kernel for meld_b200_graph_export_knn_kernel,
 *        bit1 use the SIMT fp32 candidate search instead of the tcgen05 one
 *        (test cross-check only).
 * Synchronises the stream (row counts come back to size the CSR).               */
#define MELD_B200_FLAG_KEEP_KNN_KERNEL 1
#define MELD_B200_FLAG_SIMT_SEARCH 2
int meld_b200_knn_graph_build(const double *X, int64_t n, int64_t d, int knn, double decay,
                              double thresh, double anisotropy, double bandwidth_scale,
                              int flags, void *stream, meld_b200_graph_t **graph_out);

/* Stands in for graphtools.Graph(..., thresh=0) -> TraditionalGraph ("exact" dense graph; what the reference's
 * own known-answer test builds, test/test_meld.py:58-67): the same kernel for EVERY pair whose value has not
 * underflowed to zero, eps_i = distance to the knn-th non-self neighbour.  n <= 16384.  flags: bit0 as above. */
int meld_b200_dense_graph_build(const double *X, int64_t n, int64_t d, int knn, double decay, double anisotropy,
                                double bandwidth_scale, int flags, void *stream, meld_b200_graph_t **graph_out);

/* ---- sharded build: the two stages of meld_b200_knn_graph_build, for one-process-per-GPU drivers ------
 * Stage 1 is row-local: candidate search (pass 1 + pass 2 against ALL n cells), exact float64 distances
 * and eps_i for query rows [row_begin, row_end) only -- it shards over ranks with no exchange.  Rows are
 * in the INTERNAL cell order (every rank derives the same Morton permutation from the same X);
 * row_begin must be a multiple of 512, row_end a multiple of 512 or n.
 * Stage 2 needs every row: the caller all-gathers (NCCL) the per-row counts, candidate columns, squared
 * distances and eps of all ranks, in row order, and every rank builds the same Laplacian.               */
typedef struct meld_b200_cands meld_b200_cands_t;
int meld_b200_knn_candidates(const double *X, int64_t n, int64_t d, int knn, double decay, double thresh,
                             double bandwidth_scale, int64_t row_begin, int64_t row_end, int flags, void *stream,
                             meld_b200_cands_t **cands_out);
int meld_b200_cands_info(const meld_b200_cands_t *c, int64_t *n_rows_host, int64_t *total_host, int *has_perm_host,
                         int64_t *max_per_row_host);
/* counts: n_rows int64; cand: total int32; d2: total f64; eps: n_rows f64; perm: n int32 (any may be NULL) */
int meld_b200_cands_export(const meld_b200_cands_t *c, int64_t *counts, int32_t *cand, double *d2, double *eps,
                           int32_t *perm, void *stream);
int meld_b200_cands_destroy(meld_b200_cands_t *c);
/* counts: n int64 (candidates per row, row order), cand / d2: total entries, eps: n, perm: n or NULL.    */
int meld_b200_graph_from_candidates(int64_t n, const int64_t *counts, const int32_t *cand, const double *d2,
                                    int64_t total, const double *eps, const int32_t *perm, int knn, double decay,
                                    double thresh, double anisotropy, double bandwidth_scale, int flags, void *stream,
                                    meld_b200_graph_t **graph_out);

/* ---- sharded stage 2: every rank assembles ONLY its rows of L (nothing O(nnz) is replicated or exchanged) -------
 * Given a rank's stage-1 result `c` (rows [row_begin, row_end)), eps of ALL rows (the caller all-gathers the ranks'
 * eps, 8 n bytes) and the ranks' row bounds (world + 1 values, host):
 *   stage2_begin     kernel values of the local rows; entries whose reverse edge is below threshold must also appear
 *                    in row j: local j are appended here, remote j become 16-byte records {int32 j, int32 i, f64 v}.
 *                    Returns the number of records per owner rank (host, `world` values).
 *   stage2_records   writes the records, grouped by owner rank in rank order, into the caller's device buffer
 *   (caller)         all-to-all-v of the records (rank w receives the records of its rows)
 *   stage2_assemble  counts, row pointers, fill, rows sorted by column; writes the local row sums q (row_end -
 *                    row_begin doubles) into the caller's device buffer
 *   (caller)         all-gather of q (8 n bytes)
 *   stage2_finish    anisotropy + Laplacian of the local rows with the global q -> a graph handle holding rows
 *                    [row_begin, row_end) of the n-column operator (like meld_b200_graph_row_slice), with the cell
 *                    order of `c`.
 * The handle borrows `c` (destroy the stage-2 handle first); `c`'s candidate arrays are consumed.                  */
typedef struct meld_b200_stage2 meld_b200_stage2_t;
int meld_b200_stage2_begin(meld_b200_cands_t *c, const double *eps_full, const int64_t *bounds_host, int world,
                           int knn, double decay, double thresh, double anisotropy, double bandwidth_scale,
                           void *stream, meld_b200_stage2_t **out, int64_t *send_counts_host);
int meld_b200_stage2_records(meld_b200_stage2_t *st, void *records_out, void *stream);
int meld_b200_stage2_assemble(meld_b200_stage2_t *st, const void *recv_records, int64_t n_recv, void *stream,
                              double *q_local_out);
int meld_b200_stage2_finish(meld_b200_stage2_t *st, const double *q_full, void *stream,
                            meld_b200_graph_t **graph_out);
int meld_b200_stage2_destroy(meld_b200_stage2_t *st);

/* Test hook: run only the reduced-precision candidate search of knn_graph_build and return the
 * per-row pass-2 key (float, s-space) and candidate count; used to cross-check the tcgen05 search
 * against the SIMT one.  Synchronous.                                            */
int meld_b200_debug_candidate_search(const double *X, int64_t n, int64_t d, int knn, double decay,
                                     double thresh, double bandwidth_scale, int flags, void *stream,
                                     float *key2_out, int32_t *cnt_out, int64_t *cap_host);

/* Parity / prebuilt-graph entry (reference: MELD.fit(graph) short-circuit,
 * meld/benchmark.py:194-195): adopt a CSR Laplacian built elsewhere.  Rows
 * [row0, row0+n_rows) of an operator with n_cols columns; indptr has n_rows+1
 * int64 entries starting at 0.  The arrays are copied; the caller keeps its own. */
int meld_b200_graph_from_csr(int64_t n_rows, int64_t n_cols, int64_t row0, int64_t nnz,
                             const int64_t *indptr, const int32_t *indices, const double *data,
                             void *stream, meld_b200_graph_t **graph_out);

int meld_b200_graph_info(const meld_b200_graph_t *g, int64_t *n_rows_host, int64_t *n_cols_host,
                         int64_t *row0_host, int64_t *nnz_host);
/* Copies L out (caller-allocated: indptr n_rows+1 int64, indices nnz int32, data nnz f64). */
int meld_b200_graph_export_csr(const meld_b200_graph_t *g, int64_t *indptr, int32_t *indices,
                               double *data, void *stream);
/* Graphs built by meld_b200_knn_graph_build keep their rows in an internal cell order (Morton curve
 * of the leading features; row a of the exported CSR = caller's cell perm[a], columns likewise).
 * cheby_filter takes and returns signals in the CALLER's order.  perm_out: n_rows int32 (device),
 * may be NULL; *is_identity_host = 1 when no reordering was applied.                       */
int meld_b200_graph_permutation(const meld_b200_graph_t *g, int32_t *perm_out, int *is_identity_host,
                                void *stream);
/* Un-symmetrised alpha-decay kernel (graphtools kNNGraph.build_kernel_to_data);
 * only when built with MELD_B200_FLAG_KEEP_KNN_KERNEL.                           */
int meld_b200_graph_knn_kernel_nnz(const meld_b200_graph_t *g, int64_t *nnz_host);
int meld_b200_graph_export_knn_kernel(const meld_b200_graph_t *g, int64_t *indptr, int32_t *indices,
                                      double *data, void *stream);
/* Build statistics (host): [0] candidate-search passes run, [1] max candidates/row,
 * [2] candidate capacity used, [3] rows that overflowed on the first emit pass,
 * [4] search implementation (0 tcgen05, 1 SIMT), [5], [6] reserved (0), [7] row blocks of the
 * nonzero-balanced partition.                                        */
int meld_b200_graph_build_stats(const meld_b200_graph_t *g, int64_t *stats8_host);
/* CUDA-event timings of the last build (host): [0] ms of search pass 1, [1] ms of pass 2,
 * [2] flops issued by pass 2 and [3] by pass 1 (2 x 128 x 256 x K' per (row tile, column tile)
 * product that survived the bounding-ball pruning), [4] flops of an unpruned pass
 * (2 x rows x padded columns x K'), [5..7] reserved.                                   */
int meld_b200_graph_build_times(const meld_b200_graph_t *g, double *times8_host);
int meld_b200_graph_destroy(meld_b200_graph_t *g);

/* ---- filter ------------------------------------------------------------------- */

/* Stands in for pygsp Graph.estimate_lmax() (reference meld/filter.py:39):
 * Lanczos on L, returns 1.01 * largest Ritz value (the reference's 1.01 factor).
 * Stops when the bound min(r, r^2 / gap) of the Ritz value's error (r = Ritz residual, gap = distance to the
 * second Ritz value; Kato-Temple) is below rel_tol * theta (default 1e-5), or at max_iters.  Synchronous.
 * Only for full (row0 == 0, n_rows == n_cols) graphs.                            */
int meld_b200_estimate_lmax(meld_b200_graph_t *g, int max_iters, double rel_tol, void *stream,
                            double *lmax_host, int *iters_host);

/* Stands in for pygsp cheby_op(G, c, signal) as called by
 * pygsp.filters.Filter.filter(method="chebyshev") from meld/filter.py:56-59.
 * coeffs_host: m+1 Chebyshev coefficients (host).  S, R: (n, p) row-major float64
 * device arrays, p <= 8 per call.  R may not alias S.                            */
int meld_b200_cheby_filter(meld_b200_graph_t *g, double lmax, const double *coeffs_host, int n_coeffs,
                           const double *S, int p, double *R, void *stream);

/* Shared-basis parameter sweep (reference pattern meld/benchmark.py:186-200 and
 * notebooks/MELD_Quickstart.ipynb:729-747: MELD(beta=b).fit(graph).transform(labels) for hundreds of b on
 * one graph).  Filters that differ only in their Chebyshev coefficients share the basis T_k(L) S, so the
 * recurrence runs ONCE: the m+1 terms are kept in library workspace and every filter f gets
 *   R[f] = c[f][0]/2 T_0 + sum_{k>=1} c[f][k] T_k .
 * coeffs_host: n_filters x n_coeffs row-major (host).  S: (n, p) f64, p <= 8.  R: n_filters x n x p f64.    */
int meld_b200_cheby_sweep(meld_b200_graph_t *g, double lmax, const double *coeffs_host, int n_filters,
                          int n_coeffs, const double *S, int p, double *R, void *stream);

/* One term of the three-term recurrence on this graph's rows, for callers that
 * exchange T between ranks after each step (row-partitioned multi-GPU).  All signal
 * arrays are in the graph's INTERNAL cell order (meld_b200_graph_permute_signal):
 *   y      = L[rows,:] @ T_cur                       (T_cur: n_cols x p, global rows)
 *   T_new  = alpha * (y - shift * T_cur[row0+i]) - gamma * T_old[i]
 *   R[i]   = (r_scale ? R[i] : 0) + c * T_new[i] + c_cur * T_cur[row0+i]
 * T_old / T_new / R are (n_rows, p) local-row arrays; T_new may alias T_old;
 * T_old may be NULL when gamma == 0, R may be NULL to skip the accumulation.    */
int meld_b200_cheby_step(meld_b200_graph_t *g, const double *T_cur, const double *T_old, double *T_new,
                         double *R, int p, double alpha, double shift, double gamma, double c,
                         double c_cur, int r_accumulate, void *stream);

/* Signals between the caller's cell order and the graph's internal one (identity for graphs adopted from
 * CSR): to_internal = 1: out[a] = in[perm[a]]; 0: out[perm[a]] = in[a].  in / out: (n_cols, p) f64.        */
int meld_b200_graph_permute_signal(const meld_b200_graph_t *g, const double *in, int p, int to_internal,
                                   double *out, void *stream);

/* ---- row-partitioned filter over the GPUs of one NVSwitch box ---------------------------------------------
 * The reference has no distributed code (reference setup.py:45); what is sharded is the recurrence behind
 * meld/filter.py:59.  One process per GPU.  Rank r owns rows [row_begin_r, row_end_r) of L in the internal
 * cell order (meld_b200_graph_row_slice of the full graph) and, per term, computes its rows of T_k and stores
 * them straight into every peer's HBM over NVLink from inside the SpMM kernel (no NCCL call and no host on the
 * data path); ranks meet through flag words in each other's memory.
 *   meld_b200_dist_create    allocates this rank's block (flags + two full-length signal buffers)
 *   meld_b200_dist_export    writes its CUDA IPC handle (meld_b200_dist_handle_bytes() bytes, host)
 *   meld_b200_dist_connect   maps the blocks of all ranks (world x handle bytes, rank order, host); the
 *                            caller all-gathers the handles with whatever it has (torch.distributed)
 * Every rank must issue the same sequence of meld_b200_cheby_filter_dist calls.                             */
typedef struct meld_b200_dist meld_b200_dist_t;
int meld_b200_graph_row_slice(const meld_b200_graph_t *g, int64_t row_begin, int64_t row_end, void *stream,
                              meld_b200_graph_t **slice_out);
int meld_b200_dist_create(int rank, int world, int64_t n_rows_total, int p_max, void *stream,
                          meld_b200_dist_t **dist_out);
int meld_b200_dist_handle_bytes(void);
int meld_b200_dist_export(const meld_b200_dist_t *d, void *blob_host);
int meld_b200_dist_connect(meld_b200_dist_t *d, const void *all_blobs_host);
/* Test hook: contexts of all "ranks" created in ONE process on one device (CUDA IPC cannot map a block into
 * the process that exported it) are connected by pointer; the ranks' call sequences then run on different
 * streams of that device and meet through the same flag protocol.                                          */
int meld_b200_dist_connect_local(meld_b200_dist_t *d, meld_b200_dist_t *const *all, int count);
/* Halo of a row slice.  Per term a rank's rows of T_k only have to reach the peers whose own rows reference them
 * as columns.  meld_b200_graph_mark_columns sets ref[c] = 1 (device bytes, n_cols long, caller-zeroed) for every
 * column the slice references; the caller exchanges the row ranges (all-to-all of bytes: rank w's marks of rank
 * r's rows go to rank r), and meld_b200_graph_set_halo turns recv[w * chunk + i] (w = 0..world-1) into the
 * per-row peer mask the step kernel tests before each peer store.  Without a halo every row goes to every peer. */
int meld_b200_graph_mark_columns(const meld_b200_graph_t *slice, uint8_t *ref, void *stream);
int meld_b200_graph_set_halo(meld_b200_graph_t *slice, const uint8_t *recv, int64_t chunk, int world, int rank,
                             void *stream);
/* 1 when a flag wait timed out (a peer died or left the call sequence); synchronises the device.           */
int meld_b200_dist_error(const meld_b200_dist_t *d, int *err_host);
int meld_b200_dist_destroy(meld_b200_dist_t *d);
/* meld_b200_cheby_filter on the rows of `slice`: S and R are the FULL (n, p) signals in the caller's order
 * (every rank passes the same S and receives the same R).                                                   */
int meld_b200_cheby_filter_dist(meld_b200_graph_t *slice, meld_b200_dist_t *d, double lmax,
                                const double *coeffs_host, int n_coeffs, const double *S, int p, double *R,
                                void *stream);

/* meld_b200_estimate_lmax on a row-partitioned operator: every rank holds its rows (`slice`); per Lanczos step the
 * ranks exchange their rows of the new vector by peer stores and two scalars (w.Lw, |w|^2) through rank-ordered
 * slots in peer memory, so all ranks compute bit-identical coefficients and stop at the same step.           */
int meld_b200_estimate_lmax_dist(meld_b200_graph_t *slice, meld_b200_dist_t *d, int max_iters, double rel_tol,
                                 void *stream, double *lmax_host, int *iters_host);

/* Frees the library-owned build arena (several GB after a large build; it is re-grown on the next build).  */
int meld_b200_release_workspace(void);

/* ---- small fused host-side helpers -------------------------------------------- */

/* meld/meld.py:143-191 + :229-232: one-hot indicators from integer label codes
 * (codes[i] in [0,p)), optionally column-normalised to sum 1.  S: (n, p) f64.   */
int meld_b200_indicator_matrix(const int32_t *codes, int64_t n, int p, int sample_normalize, double *S,
                               void *stream);
/* meld/utils.py:35-47: row-wise L1 normalisation, zero rows stay zero.          */
int meld_b200_l1_normalize_rows(const double *in, int64_t n, int p, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MELD_B200_H_ */
