// rfb_internal.h -- shared declarations of librfb200 (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <set>
#include <map>

#include "../../include/rfb200.h"

// ---------------------------------------------------------------------------------------------
// Panel exchange workspace (K1).  One slot per CTA and step parity.  Every 8-byte word is
// { epoch tag : 32 | payload : 32 } (the NCCL "LL" idea), so a reader never needs a fence and never
// depends on 16-byte atomicity: it polls until the tags of all words it needs match the step.
// ---------------------------------------------------------------------------------------------
constexpr int RFB_MAX_PANEL_CTAS = 768;   // 148 SMs x up to 5 co-resident CTAs of the narrow kernels
constexpr int RFB_MAX_NB = 64;            // widest panel one launch factors

struct alignas(128) RfbPanelHeader {   // one 128-byte line per CTA: 148 pollers do not pile up on one L2 line
    ulonglong2 h[2];                   // four tagged words: key lo, key hi, logical row, spare
    ulonglong2 pad[6];
};

struct RfbPanelXchg {
    // header[parity][cta].h = tagged { |candidate| bits lo, hi } and { logical row, 0 }
    RfbPanelHeader header[2][RFB_MAX_PANEL_CTAS];
    // row[parity][cta][j] = tagged { lo, hi } halves of the candidate row's value in window column j
    // slot [RFB_MAX_NB] carries the reciprocal of the candidate's own value (the pivot's, if it wins)
    ulonglong2 row[2][RFB_MAX_PANEL_CTAS][RFB_MAX_NB + 2];
    unsigned int error_flag;              // set by a kernel whose poll loop gave up
    unsigned int pad[3];                  // pad[0]: "diagonal block loaded" counter of the unpivoted panel kernel
};

// Dry-run trace of the host driver (rfb_trace_lu): the launchers record what they WOULD launch instead of
// launching it, so the recursion of rfb_api.cu can be replayed and checked on a machine without a GPU.
enum RfbTraceOpCode { RFB_T_PANEL = 1, RFB_T_PANEL_NOPIV = 2, RFB_T_LASWP = 3, RFB_T_TRSM_LOWER = 4, RFB_T_GEMM = 5,
                      RFB_T_DOWNLOAD = 6, RFB_T_IOTA = 7 };
struct RfbTraceOp {
    int64_t v[8];   // v[0] = op code; operands are (row, col) offsets into the traced matrix and sizes (see rfb200.h)
};

enum RfbKernelClass { RFB_KC_PANEL = 0, RFB_KC_LASWP = 1, RFB_KC_TRSM = 2, RFB_KC_GEMM = 3, RFB_KC_OTHER = 4, RFB_KC_COUNT = 8 };

struct rfb_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t mem_bytes = 0;
    cudaStream_t stream = nullptr;        // stream everything is enqueued on
    cudaStream_t own_stream = nullptr;    // the stream created with the context
    cudaStream_t copy_stream = nullptr;   // host -> device uploads of host-mode calls
    cudaStream_t down_stream = nullptr;   // device -> host early downloads (its own stream: the two copy engines overlap)
    int early_mode = 2;                   // early download of pinned host matrices: 0 off, 1 row bands, 2 tiles (rfb_api.cu)
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_sync = nullptr;
    std::string last_error;
    rfb_opts default_opts = {};           // used by kernel-level ABI calls

    // workspaces
    RfbPanelXchg *xchg = nullptr;         // device
    uint32_t panel_epoch = 1;             // next unused epoch (host mirror)
    int64_t *d_ipiv = nullptr;            // device pivots for HOST mem_space calls
    size_t d_ipiv_cap = 0;
    int64_t *d_info = nullptr;            // device info word
    void *d_mat = nullptr;                // device matrix for HOST mem_space calls (grow-only)
    size_t d_mat_cap = 0;
    // factors kept on the device by a host-mode rfb_lu_* with opts->keep_factors (they live in d_mat / d_ipiv)
    struct { bool valid = false; bool f32 = false, nopiv = false; int64_t n = 0, ldd = 0, id = 0; } kept;
    int64_t kept_counter = 0;
    void *d_rhs = nullptr;                // right-hand sides of rfb_solve_kept_* (grow-only)
    size_t d_rhs_cap = 0;
    int64_t *h_pinned = nullptr;          // pinned scratch (info + small results)
    int64_t *d_binfo = nullptr;           // device info words of host-mode batched calls (grow-only)
    size_t d_binfo_cap = 0;
    // per-panel row-exchange lists written by K1 and consumed by the list-driven K2 (absolute rows)
    int *perm_dst = nullptr, *perm_src = nullptr, *perm_width = nullptr;
    size_t perm_cap = 0;                  // columns the three arrays are sized for
    bool perm_external = false;           // arrays belong to the caller (rfb_perm_buffers)
    // node-level row interchange (laswp.cu): net permutation of the chunk being applied
    // (one set per lane: the multi-GPU driver runs interchanges on its compute stream and on its replica stream at once)
    // lanes 0 / 1: the two streams of a driver; lanes 2..: cached compositions of the multi-GPU driver (one per ring slot)
    static constexpr int kLanes = 10;
    int lane = 0;
    int *net_meta_[kLanes] = {}, *net_srcmap_[kLanes] = {}, *net_clist_[kLanes] = {};
    int64_t net_key[kLanes] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
    int *net_meta() const { return net_meta_[lane]; }
    int *net_srcmap() const { return net_srcmap_[lane]; }
    int *net_clist() const { return net_clist_[lane]; }
    int gemm_reduce_epilogue = 1;         // K4 Float64: full tiles leave as TMA bulk f64 reduce-adds (default; env RFB_GEMM_EPILOGUE=0: SM-side read-modify-write)
    int64_t laswp_net_min = 512;          // pivot ranges at least this long take the node-level path (env RFB_LASWP_NET_MIN)
    int64_t laswp_net_cap = 0;            // pivots per chunk, 0 = kernel default (env RFB_LASWP_NET_CAP; tests shrink it)

    std::vector<cudaEvent_t> up_events;   // upload-chunk events of host-mode calls (reused)

    // statistics
    int64_t launches = 0;
    bool profiling = false;
    double prof_ms[RFB_KC_COUNT] = {0};
    int64_t prof_launches[RFB_KC_COUNT] = {0};
    double prof_work[RFB_KC_COUNT] = {0};  // algorithmic flops (panel, trsm, gemm) or bytes (laswp)
    std::vector<cudaEvent_t> prof_events; // pairs (start, stop) pending
    std::vector<int> prof_classes;

    // co-resident CTA capacity of each panel-kernel instantiation (cooperative launch limit)
    std::map<const void *, int> panel_capacity;
    // (kernel, cluster size) -> can this device co-schedule one such cluster (DSMEM panel exchange)
    std::map<std::pair<const void *, int>, bool> cluster_ok;
    // kernels whose dynamic shared memory limit has been raised on this device
    std::set<const void *> smem_configured;

    // TMA
    void *encode_tiled = nullptr;         // cuTensorMapEncodeTiled via cudaGetDriverEntryPoint

    // dry run (rfb_trace_lu): no CUDA calls at all, launches are recorded
    bool dry_run = false;
    std::vector<RfbTraceOp> trace;
    const char *trace_base = nullptr;     // fake base address of the traced matrix
    int64_t trace_lda = 0;
    size_t trace_elt = 8;
    void rec(int op, const void *p0, const void *p1, const void *p2, int64_t a, int64_t b, int64_t c) {
        auto rc = [&](const void *p, int64_t &r, int64_t &cc) {
            if (!p) { r = cc = -1; return; }
            const int64_t off = (int64_t)((reinterpret_cast<const char *>(p) - trace_base) / (int64_t)trace_elt);
            r = off % trace_lda; cc = off / trace_lda;
        };
        RfbTraceOp o{};
        o.v[0] = op;
        int64_t r, cc;
        rc(p0, r, cc); o.v[1] = r; o.v[2] = cc;
        if (p1) { rc(p1, r, cc); o.v[6] = r; o.v[7] = cc; }
        (void)p2;
        o.v[3] = a; o.v[4] = b; o.v[5] = c;
        trace.push_back(o);
    }

    int fail(int code, const char *fmt, ...);
};

#define RFB_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return (ctx)->fail(RFB_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__,       \
                               __LINE__, cudaGetErrorString(e__));                            \
    } while (0)

#define RFB_TRY(expr)                      \
    do {                                   \
        int rc__ = (expr);                 \
        if (rc__ != RFB_OK) return rc__;   \
    } while (0)

// Raise the dynamic shared memory limit of `func` once per context.
static inline int rfb_ensure_smem(rfb_ctx *ctx, const void *func, size_t bytes) {
    if (ctx->smem_configured.count(func)) return RFB_OK;
    RFB_CUDA(ctx, cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    ctx->smem_configured.insert(func);
    return RFB_OK;
}

// RAII-less profiling hooks: bracket each launch with events when profiling is on.
struct RfbLaunchScope {
    rfb_ctx *ctx;
    int cls;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    RfbLaunchScope(rfb_ctx *c, int k, double work = 0.0) : ctx(c), cls(k) {
        ctx->launches++;
        ctx->prof_launches[cls]++;
        ctx->prof_work[cls] += work;
        if (ctx->profiling) {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, ctx->stream);
        }
    }
    ~RfbLaunchScope() {
        if (ctx->profiling) {
            cudaEventRecord(e1, ctx->stream);
            ctx->prof_events.push_back(e0);
            ctx->prof_events.push_back(e1);
            ctx->prof_classes.push_back(cls);
        }
    }
};

// ---- kernel launchers (each enqueues on ctx->stream, returns RFB_* status) --------------------
template <typename T>
int rfb_launch_panel(rfb_ctx *ctx, T *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv_dev,
                     int64_t ipiv_add, int64_t *info_dev, int64_t col_offset, int64_t perm_row0 = -1);
// batched small-matrix LU: one CTA per matrix, n <= 64, m <= 128 (panel_batched_*.cu); info must be pre-zeroed
template <typename T>
int rfb_launch_panel_batched(rfb_ctx *ctx, T *A, int64_t m, int64_t n, int64_t lda, int64_t stride_a, int64_t batch,
                             int64_t *ipiv_dev, int64_t *info_dev);
// K1' (panel_nopiv.cu): src/lu.jl:290-338 with Pivot = false; negative info on a zero pivot
template <typename T>
int rfb_launch_panel_nopiv(rfb_ctx *ctx, T *A, int64_t m, int64_t n, int64_t lda, int64_t *info_dev, int64_t col_offset);
int rfb_launch_iota(rfb_ctx *ctx, int64_t *p_dev, int64_t n, int64_t first);
// butterfly.cu: 🦋mul! (src/butterflylu.jl:93-113) and the factored U' b / V b products (:50-52)
template <typename T>
int rfb_launch_butterfly_mul(rfb_ctx *ctx, T *A, int64_t M, int64_t lda, const T *uv);
template <typename T>
int rfb_launch_butterfly_vec(rfb_ctx *ctx, T *B, int64_t M, int64_t nrhs, int64_t ldb, const T *uv, int which);
template <typename T>
int rfb_launch_set_diag(rfb_ctx *ctx, T *A, int64_t lda, int64_t i0, int64_t i1, T value);
// widest leaf (64, 32 or 16 columns) whose one-row-per-thread cooperative grid can hold m rows; 0 = none
template <typename T>
int rfb_panel_leaf_for_rows(rfb_ctx *ctx, int64_t m);
// list-driven row interchange: applies the exchange lists of the panels covering pivots [k0, k1)
// to the (rows >= k0) x ncols block whose first row is absolute row k0
template <typename T>
int rfb_launch_laswp_lists(rfb_ctx *ctx, T *A, int64_t ncols, int64_t lda, int64_t k0, int64_t k1, int64_t row_bound,
                           int64_t cache_key = -1);
template <typename T>
int rfb_launch_laswp(rfb_ctx *ctx, T *A, int64_t ncols, int64_t lda, const int64_t *ipiv_dev,
                     int64_t npiv, int64_t ipiv_sub);
template <typename T>
int rfb_launch_trsm(rfb_ctx *ctx, const T *L, int64_t k, T *B, int64_t nrhs, int64_t lda,
                    const rfb_opts *opts);
template <typename T>
int rfb_launch_trsm_upper(rfb_ctx *ctx, const T *U, int64_t k, T *B, int64_t nrhs, int64_t lda, const rfb_opts *opts);
template <typename T>
int rfb_launch_gemm(rfb_ctx *ctx, T *C, const T *A, const T *B, int64_t m, int64_t n, int64_t k,
                    int64_t lda, const rfb_opts *opts);
// K4 for n <= 8 right-hand-side columns (gemm_skinny.cu): HBM-bound GEMV-shaped update, deterministic
constexpr int RFB_SKINNY_MAX_N = 8;
template <typename T>
int rfb_launch_gemm_skinny(rfb_ctx *ctx, T *C, const T *A, const T *B, int64_t m, int64_t n, int64_t k, int64_t lda);
int rfb_launch_ipiv_shift(rfb_ctx *ctx, int64_t *ipiv_dev, int64_t n, int64_t shift);
int rfb_run_dmma_peak(rfb_ctx *ctx, int iters, double *tflops);
int rfb_run_tf32_peak(rfb_ctx *ctx, int iters, double *tflops);
int rfb_run_copy_bench(rfb_ctx *ctx, size_t bytes, int iters, double *gbs);

// src/lu.jl:158-162
template <typename T>
static inline int64_t rfb_nsplit(int64_t n) {
    int64_t k = 128 / (int64_t)sizeof(T);
    if (k < 2) k = 2;
    int64_t k2 = k / 2;
    return n >= k ? ((n + k2) / k) * k2 : n / 2;
}
