/* scl_b200.h -- C ABI of libscl_b200.so: the B200 (sm_100a) descriptor-space hot path of
 * janinethoma/soft_contrastive_learning.
 *
 * The reference is pure Python/TF-1.10 and has no FFI of its own; its boundary for this path is a set of
 * Python call sites.  Each entry point below names the reference call site it replaces (file:line relative
 * to the reference root).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host; tensors are
 *    row-major, contiguous, 16-byte aligned; outputs are pre-allocated by the caller;
 *  - scalar results (losses) are written to device memory: no entry point synchronises the stream
 *    except where stated (scl_knn_query);
 *  - workspace and stream are per call, so concurrent calls from several host threads are safe when they use
 *    different workspaces (train.py runs up to three threads per session).  The ONLY process-wide mutable state is
 *    the explicitly documented set of knobs below -- scl_set_tuning (initialised once from SCL_* environment
 *    variables when the library is loaded; no entry point calls getenv), scl_set_gemm_precision and the
 *    scl_knn_timing measurement hook (mutex-guarded) -- set them before concurrent use.  Per-device attributes
 *    (dynamic shared-memory limits) are cached per device ordinal, so one process may drive several GPUs;
 *  - return value 0 on success, a negative scl_status otherwise; never throws, never exits;
 *  - there is no CPU fallback: on a device that is not compute capability 10.x every compute entry point
 *    returns SCL_ERR_ARCH.
 */
#ifndef SCL_B200_H
#define SCL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* scl_stream_t; /* cudaStream_t */

typedef enum {
  SCL_OK = 0,
  SCL_ERR_BAD_ARG = -1,     /* null pointer / unknown enum */
  SCL_ERR_BAD_SHAPE = -2,   /* unsupported or inconsistent sizes */
  SCL_ERR_ALIGN = -3,       /* pointer not 16-byte aligned */
  SCL_ERR_WORKSPACE = -4,   /* workspace too small */
  SCL_ERR_CUDA = -5,        /* a CUDA runtime/driver call failed: see scl_last_error() */
  SCL_ERR_ARCH = -6,        /* device is not sm_100 */
  SCL_ERR_UNSUPPORTED = -7
} scl_status;

int scl_version(void);
const char* scl_strerror(int status);
const char* scl_last_error(void); /* thread-local detail string of the last SCL_ERR_CUDA */
int scl_device_ok(void);          /* 0 when the current device is compute capability 10.x */

/* Process-wide tuning / test knobs, by the name of the environment variable that initialises them at load time:
 *   SCL_WMS_STREAM, SCL_WMS_STREAM_CFG, SCL_WMS_CLUSTER, SCL_WMS_CHUNKED, SCL_TUPLE_CLUSTER, SCL_KNN_TC_VARIANT,
 *   SCL_KNN_SYNC, SCL_KNN_SYNC_WINDOW, SCL_KNN_SYNC_SUBS, SCL_KNN_RANGES, SCL_KNN_GROUP_M, SCL_KNN_CHUNK_Q,
 *   SCL_KNN_STAGE2, SCL_GEMM_SIMT, SCL_NV_FUSED.
 * value = INT32_MIN restores "unset" (the built-in choice).  They select between kernels that compute the same result
 * (the parity tests force each one); unknown names return SCL_ERR_BAD_ARG. */
int scl_set_tuning(const char* name, int value);
int scl_get_tuning(const char* name, int* value);

/* ------------------------------------------------------------------------------------------------
 * Multi-similarity family parameters.
 * wms_loss(distances, embeddings, d_alpha, d_beta, alpha=2, beta=50, lamb=1, eps=0.1, ms_mining=True,
 *          wfunction='exp', sumfunction='ms')                         model/losses.py:5
 * ms_loss(labels, embeddings, alpha=2, beta=50, lamb=1, eps=0.1, ms_mining=True)   model/losses.py:76
 * ---------------------------------------------------------------------------------------------- */
enum { SCL_WF_EXP = 0, SCL_WF_LIN = 1, SCL_WF_TANH = 2 };   /* wfunction, losses.py:11-19 */
enum { SCL_SUM_MS = 0, SCL_SUM_PLAIN = 1 };                 /* sumfunction, losses.py:39-58 */

typedef struct {
  float d_alpha, d_beta;          /* GPS sigmoid (train.py:852 ALPHA, BETA) */
  float alpha, beta, lamb, eps;   /* multi-similarity constants */
  int32_t ms_mining;
  int32_t wfunction;
  int32_t sumfunction;
} scl_ms_params;

/* W1, tuple mode.  Replaces wms_loss at train/train.py:852 (distances placeholder train.py:684-686,
 * built by train.py:557-563) together with its TF-autodiff backward (train.py:874-878).
 *   emb   [T,S,D] f32   descriptors, tuple row order [anchor, positives, negatives]   (train.py:503)
 *   dist  [T,S,S] f32   pairwise Euclidean metres (not squared)
 *   loss  [1]     f32   mean over tuples of the per-tuple losses.py:5-60 value
 *   per_tuple [T] f32   optional (may be NULL)
 *   demb  [T,S,D] f32   d loss / d emb  (may be NULL: forward only)
 *   kept  [T,S,2] u32   optional: bit j of kept[t,i,0] / kept[t,i,1] = pair (i,j) survived positive /
 *                       negative mining (losses.py:36-37, tested with >0 at :50,:53)
 * Requires 2 <= S <= 32, D % 4 == 0. */
int scl_wms_tuple_workspace_bytes(int T, int S, int D, size_t* bytes);
int scl_wms_tuple_fwd_bwd(const float* emb, const float* dist, int T, int S, int D, const scl_ms_params* p,
                          float* loss, float* per_tuple, float* demb, uint32_t* kept,
                          void* workspace, size_t workspace_bytes, scl_stream_t stream);

/* W1 / W2, flat mode: one Gram matrix over the whole batch (model/losses.py:25, :94).
 *   wms: dist [B,B] f32; ms: labels [B] i32 (train.py:822-826 classes, any integer coding).
 *   kept [2,B,B] u8 optional.  Requires D % 4 == 0. */
int scl_ms_flat_workspace_bytes(int B, int D, size_t* bytes);
int scl_wms_flat_fwd_bwd(const float* emb, const float* dist, int B, int D, const scl_ms_params* p,
                         float* loss, float* demb, uint8_t* kept,
                         void* workspace, size_t workspace_bytes, scl_stream_t stream);
int scl_ms_flat_fwd_bwd(const float* emb, const int32_t* labels, int B, int D, const scl_ms_params* p,
                        float* loss, float* demb, uint8_t* kept,
                        void* workspace, size_t workspace_bytes, scl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * L1-L5: triplet family on tuples.  Replaces, by `kind`:
 *   SCL_TRIPLET          pointnetvlad_cls.triplet_loss          train/train.py:701
 *   SCL_LAZY_TRIPLET     pointnetvlad_cls.lazy_triplet_loss     train/train.py:703
 *   SCL_QUADRUPLET       pointnetvlad_cls.quadruplet_loss       train/train.py:707-708
 *   SCL_LAZY_QUADRUPLET  pointnetvlad_cls.lazy_quadruplet_loss  train/train.py:710-712
 *   SCL_EVIL_TRIPLET     evil_triplet_loss                      model/losses.py:63-73
 *   SCL_EVIL_QUADRUPLET  evil_quadruplet_loss                   model/losses.py:197-214
 * and, with dist_term != NONE, distance_triplet_loss (model/losses.py:239-264; calls train.py:719-747):
 *   loss = triplet(kind) + lam * {distance_loss | huber_distance_loss}(a, pos, sq_d_dists, d_max, f_max).
 *   emb [T,S,D] f32 with S = 1+P+N(+1 if the kind has an `other` negative, last row)   (train.py:589-592,654)
 *   sq_d_dists [T,P] f32 squared metres anchor->positive (train.py:529-534), NULL when dist_term == NONE
 * ---------------------------------------------------------------------------------------------- */
enum { SCL_TRIPLET = 0, SCL_LAZY_TRIPLET = 1, SCL_QUADRUPLET = 2, SCL_LAZY_QUADRUPLET = 3,
       SCL_EVIL_TRIPLET = 4, SCL_EVIL_QUADRUPLET = 5,
       /* distance_quadruplet_loss, model/losses.py:267-307 (calls train/train.py:729-763): distance_triplet_loss with
        * triplet_loss / lazy_triplet_loss + the distance-term second hinge; needs dist_term != NONE and the `other` row */
       SCL_DISTANCE_QUADRUPLET = 6, SCL_DISTANCE_LAZY_QUADRUPLET = 7 };
enum { SCL_DIST_NONE = 0, SCL_DIST_SQUARED = 1, SCL_DIST_HUBER = 2 };

typedef struct {
  int32_t kind;
  int32_t dist_term;
  float m1, m2;                       /* MARGIN_1, MARGIN_2 (train.py:1254-1256) */
  float lam;                          /* LAM (train.py:1257) */
  float d_max_squared, f_max_squared; /* train.py:695-696 */
} scl_tuple_params;

int scl_tuple_loss_workspace_bytes(int T, int P, int N, int D, size_t* bytes);
int scl_tuple_loss_fwd_bwd(const float* emb, int T, int P, int N, int D, const float* sq_d_dists,
                           const scl_tuple_params* p, float* loss, float* demb,
                           void* workspace, size_t workspace_bytes, scl_stream_t stream);

/* L6: logratio_loss, model/losses.py:125-135 (call train/train.py:854-855, distances train.py:564-571).
 * Tuple mode = mean over tuples of the reference's T=1 formula.  strict_reference=1 reproduces the
 * reference's broadcast (feature ratio over all (n,p) pairs, GPS ratio element-wise (k,k); needs P==N);
 * strict_reference=0 uses the all-pairs GPS ratio of Kim et al.
 *   emb [T,1+P+N,D], sq_pos [T,P], sq_neg [T,N] (squared metres). */
int scl_logratio_fwd_bwd(const float* emb, int T, int P, int N, int D, const float* sq_pos, const float* sq_neg,
                         int strict_reference, float* loss, float* demb,
                         void* workspace, size_t workspace_bytes, scl_stream_t stream);

/* pairwise_distance_loss, model/losses.py:627-646: all-pairs squared feature distances of
 * [anchor, positives] (formula of :656-661) against all-pairs squared metres, squared (huber = 0) or Huber (huber = 1,
 * labels = scaled_f, predictions = scaled_d, delta = 1), mean over T*n*n.
 *   emb [T,n,D] f32 (n = 1+P <= 32), pairwise_sq_d [T,n,n] f32 (train.py:535-537), demb may be NULL.
 * Workspace: scl_wms_tuple_workspace_bytes(T, n, D). */
int scl_pairwise_distance_loss_fwd_bwd(const float* emb, const float* pairwise_sq_d, int T, int n, int D,
                                       float d_max_squared, float f_max_squared, int huber, float* loss, float* demb,
                                       void* workspace, size_t workspace_bytes, scl_stream_t stream);

/* D1: _pairwise_squared_distances, model/losses.py:656-661:  out[t,i,j] = r_i - 2 x_i.x_j + r_j
 * (no clamp, diagonal not forced to zero).  x [T,n,D] -> out [T,n,n]. */
int scl_pairwise_sqdist(const float* x, int T, int n, int D, float* out, scl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * N1: NetVLAD head.  Replaces  x = tf.nn.l2_normalize(x, axis=-1); x = layers.netVLAD(x, 64)
 * at model/nets.py:66-67 (and model/grad_nets.py:66-67) and its autodiff backward.
 *   x        [B,HW,C] f32   conv5_3 map, NHWC flattened over space (C = 512 in the reference)
 *   assign_w [C,K]    f32   'assignment/kernel' [1,1,C,K]
 *   centers  [C,K]    f32   'cluster_centers'  [1,1,1,C,K] (stored negated upstream, hence added)
 *   out      [B,C*K]  f32   index c*K + k
 * Workspace keeps what backward needs (soft assignments, raw VLAD, norms); pass the SAME workspace to bwd.
 * Requires C % 4 == 0, K == 64. */
int scl_netvlad_workspace_bytes(int B, int HW, int C, int K, size_t* bytes);
int scl_netvlad_fwd(const float* x, const float* assign_w, const float* centers, int B, int HW, int C, int K,
                    float* out, void* workspace, size_t workspace_bytes, scl_stream_t stream);
int scl_netvlad_bwd(const float* x, const float* assign_w, const float* centers, const float* dout,
                    int B, int HW, int C, int K, float* dx, float* dassign_w, float* dcenters,
                    void* workspace, size_t workspace_bytes, scl_stream_t stream);

/* P1: PCA-whitening projection.  Replaces train/train.py:646-652
 *   y = matmul(x - m, v, adjoint_b=True) / sqrt(var)           (eval twin: evaluation/top-n.py:74-77)
 *   x [B,Din], v [Dout,Din], m [Din], var [Dout] -> y [B,Dout];  backward: dx = (dy / sqrt(var)) v. */
int scl_pca_workspace_bytes(int B, int Din, int Dout, size_t* bytes);
int scl_pca_fwd(const float* x, const float* v, const float* m, const float* var, int B, int Din, int Dout,
                float* y, void* workspace, size_t workspace_bytes, scl_stream_t stream);
int scl_pca_bwd(const float* dy, const float* v, const float* var, int B, int Din, int Dout,
                float* dx, void* workspace, size_t workspace_bytes, scl_stream_t stream);

/* P1 with a prepared projection matrix.  v is a fed constant of the training loop (train/train.py:281-283, 647-649) and of
 * top-n.py's sweep: scl_pca_prepare splits it ONCE into fp16 hi / lo halves with one power-of-two scale (the "shadow",
 * 256-byte aligned, scl_pca_shadow_bytes); scl_pca_fwd_prepared / _bwd_prepared then run the same projection on the
 * pre-split f16 tensor-core engine (csrc/tc_gemm_h3.cu): fp32-grade (22 significant bits per operand, three products,
 * fp32 accumulation flushed every 128 k) like scl_pca_fwd, 2.5-3x faster.  Requires Din, Dout multiples of 8 and
 * <= 32768 (SCL_ERR_UNSUPPORTED otherwise: use scl_pca_fwd / scl_pca_bwd).  Workspace: scl_pca_prepared_workspace_bytes. */
int scl_pca_shadow_bytes(int Din, int Dout, size_t* bytes);
int scl_pca_prepare(const float* v, int Din, int Dout, void* shadow, size_t shadow_bytes, scl_stream_t stream);
int scl_pca_prepared_workspace_bytes(int B, int Din, int Dout, size_t* bytes);
int scl_pca_fwd_prepared(const float* x, const void* shadow, const float* m, const float* var, int B, int Din, int Dout,
                         float* y, void* workspace, size_t workspace_bytes, scl_stream_t stream);
int scl_pca_bwd_prepared(const float* dy, const void* shadow, const float* var, int B, int Din, int Dout,
                         float* dx, void* workspace, size_t workspace_bytes, scl_stream_t stream);

/* P0: the data-sized steps of the PCA *fit*, `PCA(whiten=True, n_components=d).fit(pca_f)` at
 * evaluation/top-n.py:74-75 (training twin: the incremental PCA of train/train.py:1039-1053, source absent upstream).
 *   mean[D] = column means of x [n,D] (float64 accumulation, fixed order), xc [n,D] = x - mean (may be NULL).
 * The fit's contractions (Gram xc xc^T when n <= D, covariance xc^T xc otherwise; back-projection xc^T U) are
 * scl_gemm_tf32 calls; the caller solves the small symmetric eigenproblem in between (see netvlad.pca_fit). */
int scl_pca_center_workspace_bytes(int n, int D, size_t* bytes);
int scl_pca_center(const float* x, int n, int D, float* mean, float* xc, void* workspace, size_t workspace_bytes,
                   scl_stream_t stream);

/* Precision of the tensor-core contractions (PCA, flat-mode Gram and its backward): 0 = fp32-grade 3xTF32 (default:
 * meets the 1e-5 tolerance of the reference's fp32 graph), 1 = one TF32 pass (relative error ~1e-3, three times the
 * throughput).  PROCESS-WIDE setting (one of the documented knobs, see the conventions at the top): set it before
 * concurrent use; scl_gemm_tf32 takes the precision per call. */
int scl_set_gemm_precision(int mode);
int scl_get_gemm_precision(void);
/* The contraction engine itself: C[M,N] = A . B^T (* colscale[n]) on tcgen05 kind::tf32.
 *   a_mn = 0: A(m,k) = A[m*lda + k];  a_mn = 1: A(m,k) = A[k*lda + m]  (same for B with n);  colscale may be NULL.
 * Requires lda, ldb, ldc multiples of 4 and 16-byte aligned pointers. */
int scl_gemm_tf32(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc,
                  int a_mn, int b_mn, const float* colscale, int precision, scl_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * R1: exact brute-force kNN.  Replaces
 *   KDTree(ref_f).query(query_f, k=N, return_distance=True, sort_results=True)
 * at evaluation/top-n.py:103-106 (and train/train.py:1181-1182 k=5, :451 k=MINING_CACHE_SIZE).
 *
 * An index is the fp32 database shard itself plus a device-side "shadow": fp16 copy (row pitch Dp =
 * round_up(D,64)) for the tensor-core candidate pass and exact fp32 squared norms.
 *   scl_knn_shadow_bytes   size of the shadow for R rows
 *   scl_knn_build          fills the shadow from db [R,D] f32 (one pass over the database)
 *   scl_knn_query          dist [Q,k] f64 Euclidean ascending, idx [Q,k] i64 = local row + idx_offset,
 *                          ties ordered by index; exact: every query is either certified against the fp16 rounding
 *                          bound, or resolved by the second tensor stage (all rows inside the bound collected and
 *                          rescored), or recomputed by the exact fp32->fp64 scan.
 *                          stats (optional, device i32[8]): {n_queries, n_certified, n_refused by the first pass,
 *                          path, n_resolved_by_stage2, n_exact_scan, pipeline_chunks, n_settled_by_the_bound
 *                          (scl_knn_query_end only)}
 *   force_path: 0 auto, 1 exact scan only, 2 tensor pass (+ stage 2 / scan as needed), 3 tensor pass with every query
 *               forced through the exact scan as well, 4 ... through stage 2 as well (test hooks).
 * Requires D % 4 == 0, k <= 1024.  The tensor pass keeps k' = 64 candidates per query and needs slack above k: it is
 * used for k <= 32 (top-n.py N = 25, train.py k = 5) when Q*R >= 2^22 and R >= 4096; larger k -- the hard-negative
 * mining cache of train.py:451, k = MINING_CACHE_SIZE = 1000 -- always takes the exact float64 scan.  k > R pads the
 * tail with (inf, -1) where sklearn raises.
 * Synchronisation: a fully certified call synchronises the stream once (to learn that nothing was refused); a call
 * with refused queries synchronises a second time.  Queries are processed in chunks; the merge / rescore / certificate
 * of one chunk run on an internal helper stream under the tensor pass of the next (joined before the call returns). */
int scl_knn_shadow_bytes(int64_t R, int D, size_t* bytes);
int scl_knn_build(const float* db, int64_t R, int D, void* shadow, size_t shadow_bytes, scl_stream_t stream);
int scl_knn_query_workspace_bytes(int64_t R, int D, int Q, int k, size_t* bytes);
int scl_knn_query(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q, int k,
                  int64_t idx_offset, int force_path, double* dist, int64_t* idx, int32_t* stats,
                  void* workspace, size_t workspace_bytes, scl_stream_t stream);

/* Sharded retrieval in two phases (the all-gather + merge of evaluation/top-n.py:106's one query
 * call when the database rows are split over the GPUs of a box).  A rank's exact rescore of k..64 candidates per query
 * does not shrink with its shard, and of the G*k rows the ranks return only k survive the merge.  Split in two:
 *   scl_knn_query_begin   tensor pass + candidate selection; ub (device f32 [Q,k], ascending per query) = upper bounds,
 *                         in the units of |r|^2 - 2 q.r, on the exact distances of this shard's k best candidates
 *                         (+inf where it has fewer).  Asynchronous on `stream`.
 *   -- the caller all-gathers `ub` over the ranks ([G,Q,k], 4*Q*k bytes per rank) --
 *   scl_knn_bound_reduce  bound[q] = k-th smallest of the G*k gathered values: k rows of the whole database are at or
 *                         below it.
 *   scl_knn_query_end     given that bound, rescores only the candidates of this shard that can still be among the
 *                         GLOBAL k nearest and returns them sorted by (distance, index); the rest of the k slots is
 *                         padded with (inf, -1).  Merged over the ranks (scl_topk_merge) the lists give the exact global
 *                         top-k, ties by index: every row left out is STRICTLY farther than k other rows.  Queries whose
 *                         shard holds more than 64 rows inside the bound fall back to the exact local top-k of
 *                         scl_knn_query (certificate / second tensor stage / exact scan).  Same workspace (and the SAME
 *                         workspace contents: no other call on it in between), shadow, queries and sizes as `begin`.
 * SCL_ERR_UNSUPPORTED when this shard's sizes do not take the tensor pass (see scl_knn_query): such a rank contributes
 * +inf to the gather and calls scl_knn_query instead.  stats as in scl_knn_query.
 *
 * Pipelined form.  The first phase is ONE persistent tensor launch that works through the queries in groups
 * (scl_knn_query_groups: n_groups groups of group_queries queries, the last possibly shorter; a function of D and Q
 * only, so every rank sees the same groups) and signals the completion of each group in device memory:
 *   scl_knn_query_launch       query preparation + the tensor launch, on `stream`
 *   scl_knn_query_begin_group  makes ITS stream wait for the signal of `group` (cuStreamWaitValue32 -- no SM is held
 *                              while waiting), then selects the candidates of the group's queries; ub = the group's rows
 *   scl_knn_query_end_group    as scl_knn_query_end for the group's queries; bound / dist / idx = the group's rows
 * Called on a second stream, the selection, the exchange, the rescore and the shard merge of group g overlap the tensor
 * kernel's work on group g+1.  group = -1 addresses all queries (begin = launch + begin_group(-1), end = end_group(-1)).
 * launch, begin_group and end_group of one query belong to ONE host thread; the next launch on the same workspace must
 * be ordered after the last end_group (e.g. the launch stream waits for the second stream). */
int scl_knn_query_begin(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q, int k,
                        float* ub, void* workspace, size_t workspace_bytes, scl_stream_t stream);
int scl_knn_bound_reduce(const float* ub_all, int G, int Q, int k, float* bound, scl_stream_t stream);
int scl_knn_query_end(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q, int k,
                      int64_t idx_offset, const float* bound, double* dist, int64_t* idx, int32_t* stats,
                      void* workspace, size_t workspace_bytes, scl_stream_t stream);
int scl_knn_query_groups(int D, int Q, int* n_groups, int* group_queries);
int scl_knn_query_launch(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q, int k,
                         void* workspace, size_t workspace_bytes, scl_stream_t stream);
int scl_knn_query_begin_group(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q, int k,
                              int group, float* ub, void* workspace, size_t workspace_bytes, scl_stream_t stream);
int scl_knn_query_end_group(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q, int k,
                            int64_t idx_offset, int group, const float* bound, double* dist, int64_t* idx, int32_t* stats,
                            void* workspace, size_t workspace_bytes, scl_stream_t stream);

/* Test hook: when set (per host thread) and capacity_floats >= Q*R, the tensor pass of the following scl_knn_query
 * calls on this thread also writes its raw fp16-pass scores [Q,R] there.  NULL switches it off. */
int scl_knn_set_debug_scores(float* scores, size_t capacity_floats);

/* Measurement hook (bench.py), process-wide and mutex-guarded: returns the accumulated device time (CUDA events on the launching stream) and the
 * number of launches of the tensor-pass kernel since the last reset, then, if enable >= 0, resets the counters and
 * switches the timing on (1) or off (0).  Pass enable = -1 to read without resetting. */
int scl_knn_timing(int enable, double* tensor_pass_ms_sum, int* tensor_pass_calls);

/* Merge of G per-shard sorted top-k lists (after an all-gather) -> d [Q,k], i [Q,k], ordered by (distance, index).
 * Shard g's lists are d_all + g*shard_stride and i_all + g*shard_stride (elements; 0 = Q*k, i.e. dense [G,Q,k]
 * arrays): a packed per-rank message [dist Q*k | idx Q*k] gathered ONCE is merged in place with
 * d_all = buf, i_all = buf + Q*k, shard_stride = 2*Q*k.  SURVEY.md section 8(e). */
int scl_topk_merge(const double* d_all, const int64_t* i_all, int G, int Q, int k, int64_t shard_stride,
                   double* d, int64_t* i, scl_stream_t stream);

/* R2: geographic bookkeeping of evaluation/top-n.py:69,110-113 without materialising the [Q,R] matrix:
 *   top_g_dists[q,j] = |query_xy[q] - ref_xy[top_i[q,j]]|,  gt_i[q] = argmin_r |query_xy[q]-ref_xy[r]|,
 *   gt_g_dist[q] = that minimum.  xy arrays are f64 [.,2]; top_i indexes ref_xy directly. */
int scl_geo_topn(const double* query_xy, const double* ref_xy, int Q, int64_t R, const int64_t* top_i, int k,
                 double* top_g_dists, int64_t* gt_i, double* gt_g_dist, scl_stream_t stream);

/* R3: recall curves.  evaluation/roc.py:213-216 / train/train.py:368-375:
 *   curve[n,x] = 100 * |{q : min_{j<=n} top_g_dists[q,j] < thresholds[x]}| / Q   for n < k. */
int scl_recall_curves(const double* top_g_dists, int Q, int k, const double* thresholds, int n_thresholds,
                      double* curves, scl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SCL_B200_H */
