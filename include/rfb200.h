/*
 * rfb200.h -- C ABI of librfb200.so: B200-native (sm_100a) recursive left-looking LU with partial
 * pivoting, the drop-in for the hot path of RecursiveFactorization.jl (`lu` / `lu!`).
 *
 * Reference citations are relative to /root/reference (RecursiveFactorization.jl 0.2.30).
 *
 * Conventions
 *   - every function returns an int status (RFB_OK == 0); numerical singularity is NOT an error:
 *     it is reported through `info` exactly like src/lu.jl:321-327 / :248-255 (0 = ok, k > 0 = first
 *     exactly-zero pivot at global column k, factorization continues).  Throwing
 *     SingularException (src/lu.jl:128 `checknonsingular`) is the caller's (Julia / Python) job.
 *   - matrices are column-major with leading dimension `lda` (elements); `ipiv` is int64
 *     (= Julia BlasInt), 1-based, sequential-swap (LAPACK) semantics, length min(m, n):
 *     byte-for-byte what `LinearAlgebra.LU.ipiv` holds (src/lu.jl:129).
 *   - the library never keeps a host pointer past the call and never frees caller memory.
 *   - there is no CPU fallback: without a usable sm_100 device rfb_create fails.
 *   - a context is not thread-safe; use one per host thread / Julia task.
 */
#ifndef RFB200_H
#define RFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rfb_ctx rfb_ctx;

enum {
    RFB_OK = 0,
    RFB_ERR_ARG = 1,         /* bad argument (null pointer, lda < m, ...)                     */
    RFB_ERR_CUDA = 2,        /* CUDA runtime / driver failure; see rfb_last_error             */
    RFB_ERR_NCCL = 3,        /* NCCL failure (multi-GPU path)                                 */
    RFB_ERR_UNSUPPORTED = 4, /* shape / option outside what the kernels implement             */
    RFB_ERR_NOMEM = 5,       /* device allocation failed                                      */
    RFB_ERR_INTERNAL = 6     /* device-side protocol error (e.g. panel exchange timed out)    */
};

enum { RFB_MEM_HOST = 0, RFB_MEM_DEVICE = 1 };

/* Float32 trailing-update arithmetic (config "8192x8192 Float32 LU, bf16/TF32 tensor-core GEMM with FP32 accumulate"). */
enum {
    RFB_F32_AUTO = 0,   /* default.  Whole-path calls (rfb_lu_f32, rfb_mg_lu_f32): TF32X3 when min(m, n) >= 4096, exact
                           FP32 below.  Measured reason for the threshold: the 3xTF32 residual is ~3.6x the FP32 one, which
                           keeps north_star's ||PA-LU||_F/||A||_F <= 20 n eps with a ~500x margin at every size but
                           crosses the reference's own ABSOLUTE inf-norm bound 20*m*eps (test/runtests.jl:19-20, which
                           the reference only applies at n <= 300) near n = 1000; below 4096 the trailing update is a
                           small share of the time anyway (the pivot chain dominates).
                           Kernel-level calls (rfb_gemm_nn_sub_f32 / rfb_trsm_*_f32): exact FP32.                        */
    RFB_F32_TF32X3 = 1, /* tcgen05 kind::tf32, 3-term split, FP32 accumulate in TMEM: ~2.4x faster trailing update,
                           residual ~3.6x the FP32 mode's (the tensor core adds into its accumulator with truncation),
                           still ~500x inside 20*n*eps in the Frobenius metric; views that are not 16-byte aligned
                           with lda % 4 == 0 fall through to the FP32 tiles                                              */
    RFB_F32_FP32 = 2    /* exact FP32 FFMA register tiles everywhere                                                     */
};

/* Options of the whole-path calls.  Zero-initialise for defaults.
 * The reference's `blocksize` / `threshold` keywords (src/lu.jl:101-102) tune a CPU register
 * kernel; here the analogous knob is `leaf_width` (columns factored by one panel-kernel launch). */
typedef struct rfb_opts {
    int32_t mem_space;   /* RFB_MEM_HOST: A/ipiv are host pointers (copied in and out);
                            RFB_MEM_DEVICE: A/ipiv are device pointers, nothing is copied     */
    int32_t leaf_width;  /* 0 = default (64); one of 16, 32, 64                               */
    int32_t f32_mode;    /* RFB_F32_*                                                         */
    int32_t trsm_block;  /* 0 = default; diagonal block of the blocked TRSM                   */
    int32_t gemm_path;   /* 0 = auto, 1 = force generic (cp.async) tiles, 2 = force TMA tiles */
    int32_t laswp_path;  /* 0 = auto, 1 = force the ipiv-driven kernel                        */
    int32_t no_pivot;    /* 1 = pivot = Val(false) / NoPivot() (src/lu.jl:27-65): no row interchanges,
                            `ipiv` may be NULL (NotIPIV) or is filled with 1:min(m,n) (:107-113), a
                            zero pivot is reported as NEGATIVE info (Julia >= 1.11, :24-25, :323-326) */
    int32_t keep_factors; /* host-mode rfb_lu_*: 1 = keep the factors + pivots resident on the device after the call so that
                            rfb_solve_kept_* can solve with them without a second upload (see rfb_kept_id)            */
    int32_t reserved[8];
} rfb_opts;

/* ---- context ------------------------------------------------------------------------------ */
int rfb_version(void);
int rfb_create(rfb_ctx **out, int device);            /* one stream + workspace on `device`   */
int rfb_destroy(rfb_ctx *ctx);
const char *rfb_last_error(rfb_ctx *ctx);             /* valid until the next call on ctx     */
int rfb_device_info(rfb_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *mem_bytes);
/* options used by the kernel-level calls below (which take no rfb_opts); NULL restores defaults */
int rfb_set_default_opts(rfb_ctx *ctx, const rfb_opts *opts);
/* Host-mode rfb_lu_* on a PAGE-LOCKED caller matrix send finished parts of the factors back while the factorization
 * is still running (results are identical; only the time at which host memory is written changes):
 *   2 = finished tiles (default): every 512-column unit's row band and every node's U12 block as soon as it is final,
 *   1 = row bands at the right spine of the recursion, 0 = off (one download at the end, reference interchange order).
 * Environment RFB_EARLY_DOWNLOAD sets the initial value. */
int rfb_set_early_download(rfb_ctx *ctx, int mode);

/* ---- whole path: twin of lu!(A, ipiv, Val(true), thread; check=false)  src/lu.jl:97-156 ----
 * Factors the m x n matrix in place (L strictly below the diagonal, U on/above: the layout of
 * `LU.factors`), writes min(m,n) pivots and *info.  With RFB_MEM_HOST the matrix is copied to
 * the device, factored there and copied back; with RFB_MEM_DEVICE it is factored where it lies
 * and the call returns after enqueueing (use rfb_sync). */
int rfb_lu_f64(rfb_ctx *ctx, double *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv,
               int64_t *info, const rfb_opts *opts);
int rfb_lu_f32(rfb_ctx *ctx, float *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv,
               int64_t *info, const rfb_opts *opts);

/* ---- kernel level: what a host-side restatement of reckernel! (src/lu.jl:189-263) calls. ----
 * All pointers are DEVICE pointers into one column-major allocation sharing `lda`; calls are
 * enqueued on the context stream and do not synchronise.
 *
 * rfb_panel_getrf: src/lu.jl:290-338 (_generic_lufact!) on an m x n panel, n <= 64, m >= n.
 *   ipiv_dev[k] = (block-local 1-based pivot row) + ipiv_add;  if the k-th pivot is exactly zero
 *   and *info_dev == 0 then *info_dev = col_offset + k + 1.
 * rfb_laswp: src/lu.jl:164-188 (apply_permutation!): for i in 0..npiv-1 swap rows i and
 *   ipiv_dev[i] - 1 - ipiv_sub of the (rows x ncols) block at A.
 * rfb_trsm_llnu: TriangularSolve.ldiv!(UnitLowerTriangular(L), B) (call sites src/lu.jl:235,:153):
 *   B (k x nrhs) <- L^-1 B, reading only the strict lower triangle of the k x k block at L.
 * rfb_gemm_nn_sub: src/lu.jl:265-284 (schur_complement!): C (m x n) <- C - A (m x k) * B (k x n),
 *   product accumulated from zero and added to C once.
 * rfb_ipiv_shift: src/lu.jl:256-260 (P2 .+= n1).
 */
int rfb_panel_getrf_f64(rfb_ctx *ctx, double *A, int64_t m, int64_t n, int64_t lda,
                        int64_t *ipiv_dev, int64_t ipiv_add, int64_t *info_dev, int64_t col_offset);
int rfb_panel_getrf_f32(rfb_ctx *ctx, float *A, int64_t m, int64_t n, int64_t lda,
                        int64_t *ipiv_dev, int64_t ipiv_add, int64_t *info_dev, int64_t col_offset);
int rfb_laswp_f64(rfb_ctx *ctx, double *A, int64_t ncols, int64_t lda, const int64_t *ipiv_dev,
                  int64_t npiv, int64_t ipiv_sub);
int rfb_laswp_f32(rfb_ctx *ctx, float *A, int64_t ncols, int64_t lda, const int64_t *ipiv_dev,
                  int64_t npiv, int64_t ipiv_sub);
int rfb_trsm_llnu_f64(rfb_ctx *ctx, const double *L, int64_t k, double *B, int64_t nrhs, int64_t lda);
int rfb_trsm_llnu_f32(rfb_ctx *ctx, const float *L, int64_t k, float *B, int64_t nrhs, int64_t lda);
int rfb_gemm_nn_sub_f64(rfb_ctx *ctx, double *C, const double *A, const double *B, int64_t m,
                        int64_t n, int64_t k, int64_t lda);
int rfb_gemm_nn_sub_f32(rfb_ctx *ctx, float *C, const float *A, const float *B, int64_t m,
                        int64_t n, int64_t k, int64_t lda);
int rfb_ipiv_shift(rfb_ctx *ctx, int64_t *ipiv_dev, int64_t n, int64_t shift);
/* rfb_trsm_lunn: B (k x nrhs) <- U^-1 B with U the (non-unit) upper triangle of the k x k block at U:
 * the back-substitution leg of `ldiv!(F::LU, B)` (LinearAlgebra; src/lu.jl:62 for the NotIPIV overload). */
int rfb_trsm_lunn_f64(rfb_ctx *ctx, const double *U, int64_t k, double *B, int64_t nrhs, int64_t lda);
int rfb_trsm_lunn_f32(rfb_ctx *ctx, const float *U, int64_t k, float *B, int64_t nrhs, int64_t lda);

/* ---- consumer of the factorization: `ldiv!(F, B)` for a square LU --------
 * B (n x nrhs) <- U^-1 L^-1 P B with the packed factors / pivots produced by rfb_lu_*.  Host mode copies
 * factors, pivots and B in and B out; device mode works in place and needs ldb == lda (the kernels share
 * one leading dimension).  A singular U (info > 0) gives Inf/NaN like LAPACK getrs, no error. */
/* ipiv == NULL: the factorization is unpivoted (NotIPIV), no row interchanges are applied. */
int rfb_solve_f64(rfb_ctx *ctx, const double *LU, int64_t n, int64_t lda, const int64_t *ipiv, double *B,
                  int64_t nrhs, int64_t ldb, const rfb_opts *opts);
int rfb_solve_f32(rfb_ctx *ctx, const float *LU, int64_t n, int64_t lda, const int64_t *ipiv, float *B,
                  int64_t nrhs, int64_t ldb, const rfb_opts *opts);

/* Device-resident factors (`lu` followed by `ldiv!` without uploading the factors a second time).  A host-mode rfb_lu_* with
 * opts->keep_factors = 1 leaves the square factorization in the context's staging buffer; rfb_kept_id returns its identifier
 * (0 = nothing resident: any later host-mode call on the context that needs the staging buffer drops it);
 * rfb_solve_kept_*(ctx, id, B, nrhs, ldb): B (host, n x nrhs) <- U^-1 L^-1 P B like rfb_solve_*, RFB_ERR_ARG if `id` is stale. */
int rfb_kept_id(rfb_ctx *ctx, int64_t *id);
int rfb_solve_kept_f64(rfb_ctx *ctx, int64_t id, double *B, int64_t nrhs, int64_t ldb);
int rfb_solve_kept_f32(rfb_ctx *ctx, int64_t id, float *B, int64_t nrhs, int64_t ldb);

/* ---- pivot = Val(false) and the butterfly solver ---------------------------
 * The NoPivot factorization itself is rfb_lu_* with opts->no_pivot = 1; rfb_solve_* with ipiv == NULL is
 * `ldiv!(F::LU{..,NotIPIV}, B)` (src/lu.jl:60-64).
 * rfb_panel_getrf_nopiv: src/lu.jl:290-338 with Pivot = false on an m x n device panel (n <= 64, m >= n);
 *   a zero pivot at local column k sets *info_dev = -(col_offset + k + 1) if it was 0.
 * rfb_butterfly_mul: `🦋mul!(A, uv)` (src/butterflylu.jl:93-113): the n x n device matrix (n % 4 == 0)
 *   becomes U' A V for the two-level random butterflies held in uv_dev (4 n values laid out as the
 *   reference's `uv`: U1 | V1 | U2 | V2 | U | V with lengths n/2, n/2, n/2, n/2, n, n).
 * rfb_butterfly_vec: which = 0: B <- U' B (src/butterflylu.jl:50); which = 1: B <- V B (:52); B is n x nrhs.
 * rfb_butterfly_solve: the whole `🦋solve!` (src/butterflylu.jl:45-55) for A x = b.  Host mode: A (n x n,
 *   NOT modified), B (n x nrhs, overwritten with the solution) and uv (4 * npad values, npad = n rounded up
 *   to a multiple of 4 -- the library pads like `pad!`, :180-197) are host pointers.  Device mode: in place
 *   on device buffers, n % 4 == 0, ldb == lda; A is overwritten with the factors of U' A V.
 *   *info is the NoPivot factorization's info (0, or -k for a zero pivot at column k). */
int rfb_panel_getrf_nopiv_f64(rfb_ctx *ctx, double *A, int64_t m, int64_t n, int64_t lda, int64_t *info_dev,
                              int64_t col_offset);
int rfb_panel_getrf_nopiv_f32(rfb_ctx *ctx, float *A, int64_t m, int64_t n, int64_t lda, int64_t *info_dev,
                              int64_t col_offset);
int rfb_butterfly_mul_f64(rfb_ctx *ctx, double *A, int64_t n, int64_t lda, const double *uv_dev);
int rfb_butterfly_mul_f32(rfb_ctx *ctx, float *A, int64_t n, int64_t lda, const float *uv_dev);
int rfb_butterfly_vec_f64(rfb_ctx *ctx, double *B, int64_t n, int64_t nrhs, int64_t ldb, const double *uv_dev,
                          int which);
int rfb_butterfly_vec_f32(rfb_ctx *ctx, float *B, int64_t n, int64_t nrhs, int64_t ldb, const float *uv_dev,
                          int which);
int rfb_butterfly_solve_f64(rfb_ctx *ctx, const double *A, int64_t n, int64_t lda, double *B, int64_t nrhs,
                            int64_t ldb, const double *uv, int64_t *info, const rfb_opts *opts);
int rfb_butterfly_solve_f32(rfb_ctx *ctx, const float *A, int64_t n, int64_t lda, float *B, int64_t nrhs,
                            int64_t ldb, const float *uv, int64_t *info, const rfb_opts *opts);

/* ---- batched factorization -------------------------------------------------
 * `lu!` on each of `batch` m x n matrices A + b * stride_a (column-major, shared lda): ipiv holds
 * batch * min(m, n) pivots (matrix b at offset b * min(m, n), 1-based, local to its matrix), info holds
 * batch words.  The reference has no batched call; its motivating workload is (README.md:34-35, many small
 * Jacobians), and each matrix gets exactly what `lu!` gives it.  Small pivoted matrices (n <= 64, m <= 128)
 * are factored by one launch with one CTA per matrix running the unblocked loop the reference uses below
 * its threshold (src/lu.jl:125-126, :290-338); larger ones run the recursive driver one after another. */
int rfb_lu_batched_f64(rfb_ctx *ctx, double *A, int64_t m, int64_t n, int64_t lda, int64_t stride_a, int64_t batch,
                       int64_t *ipiv, int64_t *info, const rfb_opts *opts);
int rfb_lu_batched_f32(rfb_ctx *ctx, float *A, int64_t m, int64_t n, int64_t lda, int64_t stride_a, int64_t batch,
                       int64_t *ipiv, int64_t *info, const rfb_opts *opts);

/* ---- building blocks of the multi-GPU driver (1-D block-cyclic columns) ----
 * The distributed recursion runs on the host (recursivefactorization.jl_b200/dist_lu.py, one process
 * per GPU); each rank holds a full-size column-major buffer in which its own block columns and the
 * received L panels are valid, and calls:
 * rfb_lu_range: reckernel! (src/lu.jl:189-263) on columns [c0, c0+n) of the root matrix (rows
 *   c0..m), pivots and info in GLOBAL coordinates, exchange lists recorded for rfb_laswp_range.
 * rfb_laswp_range: apply_permutation! (src/lu.jl:164-188) with pivots [k0, k1) to columns
 *   [col0, col0+ncols) (rows >= k0) of the root matrix, list-driven when the lists of those
 *   pivots are present (factored or received on this rank), ipiv-driven otherwise.
 * rfb_perm_buffers: caller-owned device arrays for the exchange lists (dst, src: 2*cap int32,
 *   width: cap int32) so that they can be broadcast with the panel; resets them.
 */
int rfb_lu_range_f64(rfb_ctx *ctx, double *A_root, int64_t m, int64_t lda, int64_t c0, int64_t n,
                     int64_t *ipiv_dev, int64_t *info_dev, const rfb_opts *opts);
int rfb_lu_range_f32(rfb_ctx *ctx, float *A_root, int64_t m, int64_t lda, int64_t c0, int64_t n,
                     int64_t *ipiv_dev, int64_t *info_dev, const rfb_opts *opts);
int rfb_laswp_range_f64(rfb_ctx *ctx, double *A_root, int64_t lda, int64_t col0, int64_t ncols,
                        int64_t k0, int64_t k1, const int64_t *ipiv_dev, int use_lists);
int rfb_laswp_range_f32(rfb_ctx *ctx, float *A_root, int64_t lda, int64_t col0, int64_t ncols,
                        int64_t k0, int64_t k1, const int64_t *ipiv_dev, int use_lists);
int rfb_perm_buffers(rfb_ctx *ctx, int32_t *dst_dev, int32_t *src_dev, int32_t *width_dev, int64_t cap);
int rfb_perm_buffers_release(rfb_ctx *ctx);   /* synchronises, then forgets caller-owned list arrays (before freeing them) */
int rfb_copy2d(rfb_ctx *ctx, void *dst_dev, size_t dpitch, const void *src_dev, size_t spitch,
               size_t width_bytes, size_t height);                               /* async d2d      */
/* enqueue on a caller-provided CUDA stream (e.g. torch's current stream); NULL restores the own one */
int rfb_set_stream(rfb_ctx *ctx, void *cuda_stream);

/* ---- multi-GPU whole path (SURVEY.md section 8b "rfb_lu_f64_mg", section 8e; BASELINE config "32768x32768 ... across 8xB200") ----
 * ONE n x n matrix factored by G GPUs of one node: block columns of width `block` distributed 1-D block-cyclic (block column
 * J on rank J mod G), each factored block column + its pivots broadcast from its owner with ncclBroadcast, L replicated.
 * The schedule is C++ inside the library (csrc/rfb_mg.cu); libnccl is loaded with dlopen on first use.
 *
 * Two ways in:
 *   one process, G devices (what a Julia caller of `lu!` uses; src/lu.jl:97-130 is the call this replaces):
 *       rfb_mg_create_all(&mg, G, devices ( NULL = 0..G-1 ));
 *       rfb_mg_lu_f64(mg, A_host, n, lda, ipiv, &info, block ( 0 = 512 ));      // upload, factor, download; A, ipiv, info as rfb_lu_f64
 *       rfb_mg_destroy(mg);                    (rfb_lu_f64_mg = the three calls in one, pays the NCCL start-up every time)
 *   one process per GPU (torchrun-style launchers; rank 0 makes the id and the launcher's own transport distributes it):
 *       rfb_mg_unique_id(id128);  rfb_mg_create_rank(&mg, device, rank, nranks, id128);
 *       rfb_mg_setup(mg, n, block, is_f32);  rfb_mg_load_block(mg, 0, j, src, ld, kind) for the owned block columns;
 *       rfb_mg_factor(mg);  rfb_mg_sync(mg, &ms);  rfb_mg_store_block / rfb_mg_get_pivots / rfb_mg_get_info.
 * `lr` is the LOCAL rank index (0..G-1 in a one-process handle, always 0 in a per-GPU handle).  Only square matrices.
 */
typedef struct rfb_mg rfb_mg;
int rfb_mg_unique_id(void *id128);                                   /* ncclGetUniqueId: 128 bytes                      */
int rfb_mg_create_rank(rfb_mg **out, int device, int rank, int nranks, const void *id128);
int rfb_mg_create_all(rfb_mg **out, int ngpus, const int *devices);
int rfb_mg_destroy(rfb_mg *mg);
const char *rfb_mg_last_error(rfb_mg *mg);
int rfb_mg_setup(rfb_mg *mg, int64_t n, int64_t block, int is_f32);  /* (re)allocates replica + own block columns       */
int rfb_mg_owner_of(int64_t block, int world);                       /* rank that owns block column `block` (no handle needed) */
int rfb_mg_local_ranks(rfb_mg *mg, int *count, int *world);
int rfb_mg_rank_ctx(rfb_mg *mg, int lr, rfb_ctx **ctx, int *global_rank);   /* the rank's single-GPU context (same device) */
int rfb_mg_block_ptr(rfb_mg *mg, int lr, int64_t j, void **dev_ptr); /* owned block column j: n x w, leading dimension n */
/* copy block column j in (src_kind 0: host memory, 1: device memory) on the rank's copy stream; the factorization waits for it */
int rfb_mg_load_block(rfb_mg *mg, int lr, int64_t j, const void *src, int64_t ld, int src_kind);
int rfb_mg_store_block(rfb_mg *mg, int lr, int64_t j, void *dst_host, int64_t ld);   /* after the factorization, async       */
int rfb_mg_factor(rfb_mg *mg);                 /* runs the schedule on every local rank; returns when all of it is enqueued */
int rfb_mg_sync(rfb_mg *mg, float *ms);        /* waits; *ms = device time of the last factorization, max over local ranks  */
int rfb_mg_get_pivots(rfb_mg *mg, int64_t *ipiv_host);               /* n pivots, 1-based global rows (every rank holds all) */
int rfb_mg_get_info(rfb_mg *mg, int64_t *info);                      /* global info (collective in per-GPU handles)         */
int rfb_mg_stats(rfb_mg *mg, int64_t *bcast_bytes_per_rank, int64_t *launches);
/* host-scheduler statistics of local rank lr's last factorization: [0] bulk slices, [1] critical-path enqueues, [2] us with an idle
 * compute stream and nothing runnable, [3] us of the whole schedule loop, [4] kernels launched so far, [5] device us inside the
 * owned block columns' critical sections (last contributions + factorization), [6] device us of their publications */
int rfb_mg_sched_stats(rfb_mg *mg, int lr, int64_t out[8]);
int rfb_mg_lu_f64(rfb_mg *mg, double *A_host, int64_t n, int64_t lda, int64_t *ipiv, int64_t *info, int64_t block);
int rfb_mg_lu_f32(rfb_mg *mg, float *A_host, int64_t n, int64_t lda, int64_t *ipiv, int64_t *info, int64_t block);
int rfb_lu_f64_mg(const int *devices, int ngpus, double *A_host, int64_t n, int64_t lda, int64_t *ipiv, int64_t *info,
                  int64_t block);
int rfb_lu_f32_mg(const int *devices, int ngpus, float *A_host, int64_t n, int64_t lda, int64_t *ipiv, int64_t *info,
                  int64_t block);
/* dry run of one rank's schedule (no GPU, no NCCL), 5 int64 per operation -- see csrc/rfb_mg.cu; used by the CPU tests, which
 * replay it with the oracle's kernels over gloo and compare with the oracle's own LU */
int rfb_mg_trace(int64_t n, int64_t block, int rank, int world, int64_t *ops, int64_t cap, int64_t *count);

/* ---- host-driver trace (no GPU needed) -------------------------------------------------------------------
 * Runs the host recursion of rfb_lu_* (twin of src/lu.jl:97-156 and :189-263) for an m x n matrix WITHOUT launching
 * anything and returns the sequence of operations it would enqueue, 8 int64 per operation:
 *   [0] op: 1 panel getrf (src/lu.jl:290-338), 2 unpivoted panel, 3 row interchanges (:164-188), 4 unit-lower TRSM
 *       (:235, :153), 5 Schur update (:265-284), 6 early download of a finished tile (host mode), 7 identity ipiv fill
 *   [1],[2] row, column of the first operand (panel / swapped block / L / C / downloaded tile); [6],[7] of the second
 *       (B of the TRSM, A of the update, whose B is at row [7], column [2]);  [3],[4],[5] sizes: panel m, n, column
 *       offset; laswp ncols, first pivot, one past the last pivot; TRSM k, nrhs; update m, n, k; download rows, cols.
 * `pinned_host` selects the schedule used for page-locked host matrices: eager interchanges + early downloads of
 *   finished tiles (pinned_host = 1: the library default, environment RFB_EARLY_DOWNLOAD = 0 off / 1 row bands at
 *   the right spine / 2 tiles; pinned_host = 10 + mode forces a mode).
 * Used by the CPU tests to replay the schedule with the oracle's kernels and compare with the oracle's own LU. */
int rfb_trace_lu(int is_f32, int64_t m, int64_t n, int64_t lda, const rfb_opts *opts, int pinned_host, int64_t *ops,
                 int64_t cap, int64_t *count);

/* ---- memory / stream plumbing ------------------------------------------------------------- */
int rfb_malloc(rfb_ctx *ctx, void **dev_ptr, size_t bytes);
int rfb_free(rfb_ctx *ctx, void *dev_ptr);
int rfb_host_alloc(rfb_ctx *ctx, void **host_ptr, size_t bytes);   /* pinned host memory       */
int rfb_host_free(rfb_ctx *ctx, void *host_ptr);
int rfb_h2d(rfb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);   /* async        */
int rfb_d2h(rfb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);   /* async        */
int rfb_d2d(rfb_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes);    /* async        */
int rfb_memset(rfb_ctx *ctx, void *dst_dev, int value, size_t bytes);           /* async        */
int rfb_sync(rfb_ctx *ctx);

/* ---- measurement helpers (CUDA events on the context stream) ------------------------------- */
int rfb_timer_start(rfb_ctx *ctx);
int rfb_timer_stop(rfb_ctx *ctx, float *ms);              /* synchronises                      */
int rfb_launch_count(rfb_ctx *ctx, int64_t *count);       /* kernels launched so far by ctx    */
/* accumulated device time per kernel class since the last reset, when profiling is enabled
 * (classes: 0 panel, 1 laswp, 2 trsm-diagonal, 3 gemm, 4 other) */
int rfb_profile_enable(rfb_ctx *ctx, int on);
int rfb_profile_read(rfb_ctx *ctx, double ms_by_class[8], int64_t launches_by_class[8],
                     double work_by_class[8]);  /* algorithmic flops (bytes for laswp) issued */
/* register-only DMMA (mma.sync m8n8k4 f64) throughput: the FP64 roofline denominator */
int rfb_bench_dmma_peak(rfb_ctx *ctx, int iters, double *tflops);
/* smem-resident tcgen05.mma kind::tf32 (128 x 128 x 8, FP32 accumulate in TMEM) throughput, one CTA per SM: the
 * tensor roofline denominator of the Float32 path (dense TF32 TFLOP/s; the 3xTF32 mode issues 3 MMAs per product) */
int rfb_bench_tf32_peak(rfb_ctx *ctx, int iters, double *tflops);
/* plain device copy bandwidth (read + write bytes / time) */
int rfb_bench_copy(rfb_ctx *ctx, size_t bytes, int iters, double *gbs);

#ifdef __cplusplus
}
#endif
#endif /* RFB200_H */
