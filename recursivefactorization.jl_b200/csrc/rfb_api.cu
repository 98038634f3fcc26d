// rfb_api.cu -- C ABI entry points, context management and the host-side recursion.
//
// The recursion is a twin of the reference driver: lu! (src/lu.jl:97-130), _recurse! with its fat
// tail (:145-156) and reckernel! (:189-263).  Differences that do not change the result:
//   * pivots and info are produced in GLOBAL coordinates by the kernels, so the reference's
//     `P2 .+= n1` (:256-260) and `info += n1` (:248-255) fix-ups have nothing left to do;
//   * recursion stops at `leaf_width` columns (default 64) instead of blocksize 8/16 (:101): one
//     K1 launch factors the whole leaf panel.
#include <algorithm>
#include <cstdarg>
#include <cstdlib>

#include "rfb_internal.h"

int rfb_ctx::fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error = buf;
    return code;
}

namespace {

struct LuPlan {
    int leaf;
    int64_t rows = 0;      // row count of the root matrix (bound of every row an exchange list can name)
    bool pivot = true;     // false: pivot = Val(false) (src/lu.jl:27-65), no interchanges, negative info
    bool lists;            // K1 emits row-exchange lists and K2 consumes them (default)
    const rfb_opts *opts;
    // early download (host mode): rows [0, n1) of the root are final long before the factorization ends
    void *host_A = nullptr;            // caller's matrix (host), nullptr in device mode
    int64_t host_lda = 0, host_m = 0;
    int early_mode = 0;                // 0 = one download at the end, 1 = row bands at the right spine, 2 = finished tiles
    int64_t early_rows = 0;            // mode 1: rows [0, early_rows) of ALL columns were already sent back (download stream)
    bool tiles_done = false;           // mode 2: every element has been sent back tile by tile
    int64_t n_total = 0;               // columns of the whole matrix
    // pipelined upload (host mode): column chunk i is resident once up_events[i] has fired
    std::vector<cudaEvent_t> *up_events = nullptr;
    const std::vector<int64_t> *up_bounds = nullptr;   // up_bounds[i] = first column NOT covered by chunks 0..i
    int up_waited = -1;    // last chunk the compute stream already waits for
};

// The compute stream is about to touch columns [0, ncols): make it wait for their upload.
static int need_cols(rfb_ctx *ctx, LuPlan &plan, int64_t ncols) {
    if (!plan.up_events || ncols <= 0) return RFB_OK;
    const int nchunks = (int)plan.up_events->size();
    while (plan.up_waited + 1 < nchunks && (plan.up_waited < 0 || (*plan.up_bounds)[plan.up_waited] < ncols)) {
        plan.up_waited++;
        RFB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, (*plan.up_events)[plan.up_waited], 0));
    }
    return RFB_OK;
}

// apply_permutation! (src/lu.jl:164-188) with pivots [k0, k0 + np) on the block whose first row is
// absolute row k0
template <typename T>
int lu_swap(rfb_ctx *ctx, T *A, int64_t ncols, int64_t lda, const int64_t *ipiv, int64_t k0, int64_t np,
            const LuPlan &plan) {
    if (plan.lists) return rfb_launch_laswp_lists<T>(ctx, A, ncols, lda, k0, k0 + np, plan.rows);
    return rfb_launch_laswp<T>(ctx, A, ncols, lda, ipiv + k0, np, k0);
}

// reckernel! (src/lu.jl:189-263) on columns [c0, c0 + n) of the root matrix; the node's block is
// rows [c0, m) (the diagonal block starts at row c0 == column c0).
//
// eager_left < 0: the reference's own order -- the interchanges of the right half reach the columns to
//   their left when the node finishes (`apply_permutation!(P2, A21)`, :246).
// eager_left >= 0 (pinned host matrices): every finished subtree of at most kEagerUnit columns (a "unit") applies its
//   interchanges to ALL columns [eager_left, c0) on its left at once, so :246 has nothing left to do at the
//   nodes above it.  Each column still receives every later pivot exactly once and in pivot order, so the
//   result is identical; what changes is WHEN elements become final, and final elements travel back to the host
//   while the factorization is still running (download stream) instead of waiting for its end:
//     mode 2 (default, "tiles"): when unit [c0, c0 + w) has applied its interchanges, rows [c0, c0 + w) of columns
//       [0, c0 + w) are final (L on the left, the unit's own L\U block; every later pivot only touches rows >= c0 + w);
//       when ANY node above unit level has done its swap + TRSM, its U12 block -- rows [c0, c0 + n1) of columns
//       [c0 + n1, c0 + n) -- is final (later pivots live below it, ancestors only read it).  Row r and column j > r's
//       unit meet in exactly one such node (the lowest common ancestor of their units), so the tiles cover the matrix
//       exactly once; the last unit takes the rows below the square part with it.  The L half of the matrix -- which a
//       row-band scheme can only release when the right spine passes -- leaves evenly, unit by unit, and only the last
//       unit's band (w x n) is still to be sent when the last panel finishes.
//     mode 1 ("row bands", round 1): after the swap + TRSM of a node on the right spine, rows [c0, c0 + n1) of every
//       column are final; the bands halve while the PCIe time of what is left does not shrink as fast as the compute
//       time, so the last quarter of the rows arrives ~6 ms after the last panel at 16384^2.
constexpr int64_t kEagerUnit = 512;
constexpr int64_t kEarlyRowsMin = 256;

// Early download (host mode): rows [r0, r0 + nr) x columns [j0, j0 + nc) are final on the device.
template <typename T>
int early_download(rfb_ctx *ctx, LuPlan &plan, T *root, int64_t lda, int64_t r0, int64_t nr, int64_t j0, int64_t nc) {
    if (nr <= 0 || nc <= 0) return RFB_OK;
    if (ctx->dry_run) {
        RfbTraceOp o{};
        o.v[0] = RFB_T_DOWNLOAD; o.v[1] = r0; o.v[2] = j0; o.v[3] = nr; o.v[4] = nc;
        ctx->trace.push_back(o);
        return RFB_OK;
    }
    RFB_CUDA(ctx, cudaEventRecord(ctx->ev_sync, ctx->stream));
    RFB_CUDA(ctx, cudaStreamWaitEvent(ctx->down_stream, ctx->ev_sync, 0));
    RFB_CUDA(ctx, cudaMemcpy2DAsync(reinterpret_cast<T *>(plan.host_A) + r0 + j0 * plan.host_lda, sizeof(T) * plan.host_lda,
                                    root + r0 + j0 * lda, sizeof(T) * lda, sizeof(T) * nr, nc, cudaMemcpyDeviceToHost,
                                    ctx->down_stream));
    return RFB_OK;
}

template <typename T>
int lu_rec(rfb_ctx *ctx, T *root, int64_t m, int64_t lda, int64_t c0, int64_t n, int64_t *ipiv, int64_t *info,
           LuPlan &plan, int64_t eager_left = -1) {
    T *A = root + c0 + c0 * lda;          // top-left of the node
    const int64_t mm = m - c0;            // rows of the node
    if (eager_left >= 0 && n <= kEagerUnit) {
        RFB_TRY(lu_rec<T>(ctx, root, m, lda, c0, n, ipiv, info, plan, -1));
        if (plan.pivot && c0 > eager_left)
            RFB_TRY(lu_swap<T>(ctx, root + c0 + eager_left * lda, c0 - eager_left, lda, ipiv, c0, n, plan));
        if (plan.host_A && plan.early_mode == 2 && eager_left == 0) {
            // the unit's band: L of all columns on its left + its own L\U block; the last unit also carries the rows
            // below the square part (tall matrices) and closes the tiling
            const bool last = c0 + n == plan.n_total;
            RFB_TRY(early_download<T>(ctx, plan, root, lda, c0, last ? m - c0 : n, 0, c0 + n));
            if (last) plan.tiles_done = true;
        }
        return RFB_OK;
    }
    if (n <= plan.leaf) {                 // :192-195 leaf -> K1
        RFB_TRY(need_cols(ctx, plan, c0 + n));
        if (!plan.pivot) return rfb_launch_panel_nopiv<T>(ctx, A, mm, n, lda, info, c0);
        return rfb_launch_panel<T>(ctx, A, mm, n, lda, ipiv + c0, c0, info, c0, plan.lists ? c0 : -1);
    }
    const int64_t n1 = rfb_nsplit<T>(n), n2 = n - n1;                                   // :196-198
    RFB_TRY(lu_rec<T>(ctx, root, m, lda, c0, n1, ipiv, info, plan, eager_left));        // :229
    T *AR = A + n1 * lda;
    RFB_TRY(need_cols(ctx, plan, c0 + n));
    if (plan.pivot) RFB_TRY(lu_swap<T>(ctx, AR, n2, lda, ipiv, c0, n1, plan));          // :233
    RFB_TRY(rfb_launch_trsm<T>(ctx, A, n1, AR, n2, lda, plan.opts));                    // :235
    if (plan.host_A && eager_left == 0 && plan.early_mode == 2) {
        RFB_TRY(early_download<T>(ctx, plan, root, lda, c0, n1, c0 + n1, n2));          // this node's U12 is final
    } else if (plan.host_A && eager_left == 0 && plan.early_mode == 1 && c0 == plan.early_rows && c0 + n == plan.n_total &&
               n1 >= kEarlyRowsMin) {
        // right-spine node, pinned host matrix: rows [c0, c0 + n1) of ALL columns are final now (L and U11 on the
        // left, U12 on the right; everything still to come touches rows >= c0 + n1 only).
        RFB_TRY(early_download<T>(ctx, plan, root, lda, c0, n1, 0, plan.n_total));
        plan.early_rows = c0 + n1;
    }
    RFB_TRY(rfb_launch_gemm<T>(ctx, AR + n1, A + n1, AR, mm - n1, n2, n1, lda, plan.opts));   // :240
    RFB_TRY(lu_rec<T>(ctx, root, m, lda, c0 + n1, n2, ipiv, info, plan, eager_left));   // :244
    if (!plan.pivot || eager_left >= 0) return RFB_OK;                                  // (eager: already applied)
    return lu_swap<T>(ctx, A + n1, n1, lda, ipiv, c0 + n1, n2, plan);                   // :246
}

template <typename T>
int lu_device(rfb_ctx *ctx, T *dA, int64_t m, int64_t n, int64_t lda, int64_t *d_ipiv, int64_t *d_info,
              const rfb_opts *opts, std::vector<cudaEvent_t> *up_events = nullptr, const std::vector<int64_t> *up_bounds = nullptr,
              T *host_A = nullptr, int64_t host_lda = 0, int64_t *early_rows = nullptr) {
    LuPlan plan;
    // Float32 arithmetic of the trailing update: resolve RFB_F32_AUTO once per factorization (include/rfb200.h)
    rfb_opts resolved = opts ? *opts : rfb_opts{};
    if (sizeof(T) == 4 && resolved.f32_mode == RFB_F32_AUTO)
        resolved.f32_mode = (m < n ? m : n) >= 4096 ? RFB_F32_TF32X3 : RFB_F32_FP32;
    opts = &resolved;
    plan.opts = opts;
    plan.up_events = up_events;
    plan.up_bounds = up_bounds;
    plan.host_A = host_A;
    plan.early_mode = host_A ? ctx->early_mode : 0;
    plan.host_lda = host_lda;
    plan.host_m = m;
    plan.rows = m;
    plan.n_total = n;
    plan.leaf = (opts && opts->leaf_width > 0) ? opts->leaf_width : 64;
    plan.pivot = !(opts && opts->no_pivot);
    if (plan.leaf != 8 && plan.leaf != 16 && plan.leaf != 32 && plan.leaf != 64)
        return ctx->fail(RFB_ERR_ARG, "leaf_width must be 8, 16, 32 or 64 (got %d)", plan.leaf);
    if (plan.pivot) {   // very tall matrices: narrow the leaf until one row per thread fits the cooperative grid
        const int fit = rfb_panel_leaf_for_rows<T>(ctx, m);
        if (fit == 0)
            return ctx->fail(RFB_ERR_UNSUPPORTED, "%lld rows exceed the panel kernel's one-row-per-thread capacity", (long long)m);
        if (plan.leaf > fit) plan.leaf = fit;
    }
    if (!ctx->dry_run) RFB_CUDA(ctx, cudaMemsetAsync(d_info, 0, sizeof(int64_t), ctx->stream));
    const int64_t mn = m < n ? m : n;
    if (mn == 0) return RFB_OK;
    plan.lists = plan.pivot && !(opts && opts->laswp_path == 1);
    if (!plan.pivot && d_ipiv) RFB_TRY(rfb_launch_iota(ctx, d_ipiv, mn, 1));               // :107-113
    if (plan.lists && !ctx->dry_run) {
        if ((size_t)mn > ctx->perm_cap) {
            if (ctx->perm_external) return ctx->fail(RFB_ERR_ARG, "caller-provided exchange-list buffers are too small");
            RFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->perm_dst); cudaFree(ctx->perm_src); cudaFree(ctx->perm_width);
            ctx->perm_dst = ctx->perm_src = ctx->perm_width = nullptr;
            ctx->perm_cap = 0;
            const size_t cap = (size_t)mn + 64;
            if (cudaMalloc(&ctx->perm_dst, 2 * cap * sizeof(int)) != cudaSuccess ||
                cudaMalloc(&ctx->perm_src, 2 * cap * sizeof(int)) != cudaSuccess ||
                cudaMalloc(&ctx->perm_width, cap * sizeof(int)) != cudaSuccess) {
                cudaGetLastError();
                return ctx->fail(RFB_ERR_NOMEM, "cannot allocate the row-exchange lists");
            }
            ctx->perm_cap = cap;
        }
        RFB_CUDA(ctx, cudaMemsetAsync(ctx->perm_dst, 0xFF, 2 * (size_t)mn * sizeof(int), ctx->stream));
        RFB_CUDA(ctx, cudaMemsetAsync(ctx->perm_width, 0, (size_t)mn * sizeof(int), ctx->stream));
    }
    RFB_TRY(lu_rec<T>(ctx, dA, m, lda, 0, mn, d_ipiv, d_info, plan, (host_A && m >= n && plan.early_mode != 0) ? 0 : -1));   // :147
    if (m < n) {                                                                        // :148-154
        T *AR = dA + m * lda;
        RFB_TRY(need_cols(ctx, plan, n));
        if (plan.pivot) RFB_TRY(lu_swap<T>(ctx, AR, n - m, lda, d_ipiv, 0, mn, plan));
        RFB_TRY(rfb_launch_trsm<T>(ctx, dA, m, AR, n - m, lda, opts));
    }
    if (early_rows) *early_rows = plan.tiles_done ? m : plan.early_rows;     // tiles: nothing is left to download
    return RFB_OK;
}

// reckernel! on a column range of the root (multi-GPU building block); needs the exchange-list arrays
template <typename T>
int lu_range(rfb_ctx *ctx, T *A_root, int64_t m, int64_t lda, int64_t c0, int64_t n, int64_t *ipiv, int64_t *info,
             const rfb_opts *opts) {
    if (!A_root || !ipiv || !info) return ctx->fail(RFB_ERR_ARG, "rfb_lu_range: null pointer");
    if (c0 < 0 || n < 0 || c0 + n > m || lda < m) return ctx->fail(RFB_ERR_ARG, "rfb_lu_range: bad range/lda");
    if (n == 0) return RFB_OK;
    if (opts && opts->no_pivot)
        return ctx->fail(RFB_ERR_UNSUPPORTED, "rfb_lu_range: pivot = Val(false) is not distributed (use rfb_lu_* with no_pivot)");
    RFB_CUDA(ctx, cudaSetDevice(ctx->device));
    LuPlan plan;
    plan.opts = opts;
    plan.rows = m;
    plan.leaf = (opts && opts->leaf_width > 0) ? opts->leaf_width : 64;
    if (plan.leaf != 8 && plan.leaf != 16 && plan.leaf != 32 && plan.leaf != 64)
        return ctx->fail(RFB_ERR_ARG, "leaf_width must be 8, 16, 32 or 64 (got %d)", plan.leaf);
    {
        const int fit = rfb_panel_leaf_for_rows<T>(ctx, m - c0);
        if (fit == 0)
            return ctx->fail(RFB_ERR_UNSUPPORTED, "%lld rows exceed the panel kernel's one-row-per-thread capacity", (long long)(m - c0));
        if (plan.leaf > fit) plan.leaf = fit;
    }
    plan.lists = !(opts && opts->laswp_path == 1);
    if (plan.lists && (ctx->perm_dst == nullptr || (size_t)(c0 + n) > ctx->perm_cap))
        return ctx->fail(RFB_ERR_ARG, "rfb_lu_range: exchange-list buffers missing or too small (rfb_perm_buffers)");
    return lu_rec<T>(ctx, A_root, m, lda, c0, n, ipiv, info, plan);
}

template <typename T>
int laswp_range(rfb_ctx *ctx, T *A_root, int64_t lda, int64_t col0, int64_t ncols, int64_t k0, int64_t k1,
                const int64_t *ipiv, int use_lists) {
    if (!A_root) return ctx->fail(RFB_ERR_ARG, "rfb_laswp_range: null matrix");
    if (ncols <= 0 || k1 <= k0) return RFB_OK;
    T *A = A_root + k0 + col0 * lda;
    if (use_lists) {
        if (ctx->perm_dst == nullptr || (size_t)k1 > ctx->perm_cap)
            return ctx->fail(RFB_ERR_ARG, "rfb_laswp_range: exchange lists missing");
        return rfb_launch_laswp_lists<T>(ctx, A, ncols, lda, k0, k1, lda);
    }
    if (!ipiv) return ctx->fail(RFB_ERR_ARG, "rfb_laswp_range: null ipiv");
    return rfb_launch_laswp<T>(ctx, A, ncols, lda, ipiv + k0, k1 - k0, k0);
}

// ldiv!(F::LU, B): B <- U^-1 L^-1 P B
template <typename T>
int solve_entry(rfb_ctx *ctx, const T *LU, int64_t n, int64_t lda, const int64_t *ipiv, T *B, int64_t nrhs, int64_t ldb,
                const rfb_opts *opts) {
    if (!ctx) return RFB_ERR_ARG;
    if (n < 0 || nrhs < 0) return ctx->fail(RFB_ERR_ARG, "negative dimension");
    if (n == 0 || nrhs == 0) return RFB_OK;
    if (!LU || !B) return ctx->fail(RFB_ERR_ARG, "null pointer");   // ipiv == NULL: NotIPIV (src/lu.jl:60-64)
    if (lda < n || ldb < n) return ctx->fail(RFB_ERR_ARG, "leading dimension smaller than n");
    RFB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int space = opts ? opts->mem_space : RFB_MEM_HOST;
    if (space == RFB_MEM_DEVICE) {
        if (ldb != lda) return ctx->fail(RFB_ERR_UNSUPPORTED, "device-mode solve needs ldb == lda (one shared leading dimension)");
        if (ipiv) RFB_TRY(rfb_launch_laswp<T>(ctx, B, nrhs, ldb, ipiv, n, 0));
        RFB_TRY(rfb_launch_trsm<T>(ctx, LU, n, B, nrhs, lda, opts));
        return rfb_launch_trsm_upper<T>(ctx, LU, n, B, nrhs, lda, opts);
    }
    ctx->kept.valid = false;                       // the staging buffer is about to be reused
    const int64_t ldd = (n + 3) & ~int64_t(3);
    const size_t need = sizeof(T) * (size_t)ldd * (size_t)(n + nrhs);
    if (need > ctx->d_mat_cap) {
        if (ctx->d_mat) cudaFree(ctx->d_mat);
        ctx->d_mat = nullptr;
        ctx->d_mat_cap = 0;
        if (cudaMalloc(&ctx->d_mat, need) != cudaSuccess) {
            cudaGetLastError();
            return ctx->fail(RFB_ERR_NOMEM, "cannot allocate %zu bytes of device memory for the solve", need);
        }
        ctx->d_mat_cap = need;
    }
    if (ipiv && (size_t)n > ctx->d_ipiv_cap) {
        if (ctx->d_ipiv) cudaFree(ctx->d_ipiv);
        ctx->d_ipiv = nullptr;
        ctx->d_ipiv_cap = 0;
        if (cudaMalloc(&ctx->d_ipiv, sizeof(int64_t) * (size_t)n) != cudaSuccess) {
            cudaGetLastError();
            return ctx->fail(RFB_ERR_NOMEM, "cannot allocate the device pivot vector");
        }
        ctx->d_ipiv_cap = (size_t)n;
    }
    T *dLU = reinterpret_cast<T *>(ctx->d_mat);
    T *dB = dLU + (size_t)ldd * n;
    RFB_CUDA(ctx, cudaMemcpy2DAsync(dLU, sizeof(T) * ldd, LU, sizeof(T) * lda, sizeof(T) * n, n, cudaMemcpyHostToDevice, ctx->stream));
    RFB_CUDA(ctx, cudaMemcpy2DAsync(dB, sizeof(T) * ldd, B, sizeof(T) * ldb, sizeof(T) * n, nrhs, cudaMemcpyHostToDevice, ctx->stream));
    if (ipiv) {
        RFB_CUDA(ctx, cudaMemcpyAsync(ctx->d_ipiv, ipiv, sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
        RFB_TRY(rfb_launch_laswp<T>(ctx, dB, nrhs, ldd, ctx->d_ipiv, n, 0));
    }
    RFB_TRY(rfb_launch_trsm<T>(ctx, dLU, n, dB, nrhs, ldd, opts));
    RFB_TRY(rfb_launch_trsm_upper<T>(ctx, dLU, n, dB, nrhs, ldd, opts));
    RFB_CUDA(ctx, cudaMemcpy2DAsync(B, sizeof(T) * ldb, dB, sizeof(T) * ldd, sizeof(T) * n, nrhs, cudaMemcpyDeviceToHost, ctx->stream));
    RFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RFB_OK;
}

// 🦋solve!(🦋workspace(A, b)) (src/butterflylu.jl:20-55): pad to a multiple of 4 (:34-38, :180-197),
// A <- U' A V (:47), NoPivot recursive LU (:48), tmp = U' b (:50), ldiv!(F, tmp) (:51, the NotIPIV
// overload src/lu.jl:60-64), x = V tmp (:52), out = x[1:n] (:53).
template <typename T>
int butterfly_solve_entry(rfb_ctx *ctx, const T *A, int64_t n, int64_t lda, T *B, int64_t nrhs, int64_t ldb,
                          const T *uv, int64_t *info, const rfb_opts *opts) {
    if (!ctx) return RFB_ERR_ARG;
    if (n < 0 || nrhs < 0) return ctx->fail(RFB_ERR_ARG, "negative dimension");
    if (!info) return ctx->fail(RFB_ERR_ARG, "info is null");
    const int space = opts ? opts->mem_space : RFB_MEM_HOST;
    if (n == 0) {
        if (space == RFB_MEM_HOST) *info = 0;
        return RFB_OK;
    }
    if (!A || !uv || (nrhs > 0 && !B)) return ctx->fail(RFB_ERR_ARG, "null pointer");
    if (lda < n || (nrhs > 0 && ldb < n)) return ctx->fail(RFB_ERR_ARG, "leading dimension smaller than n");
    if (n > 0x7ffffff0LL) return ctx->fail(RFB_ERR_UNSUPPORTED, "dimension exceeds int32");
    RFB_CUDA(ctx, cudaSetDevice(ctx->device));
    rfb_opts o = opts ? *opts : rfb_opts{};
    o.no_pivot = 1;
    o.mem_space = RFB_MEM_DEVICE;
    if (space == RFB_MEM_DEVICE) {
        // in place on device buffers: the caller has already padded (n % 4 == 0) and shares one leading dimension
        if (n % 4 != 0) return ctx->fail(RFB_ERR_UNSUPPORTED, "device-mode butterfly solve needs n %% 4 == 0 (pad first, src/butterflylu.jl:180-197)");
        if (nrhs > 0 && ldb != lda) return ctx->fail(RFB_ERR_UNSUPPORTED, "device-mode butterfly solve needs ldb == lda");
        T *dA = const_cast<T *>(A);
        RFB_TRY(rfb_launch_butterfly_mul<T>(ctx, dA, n, lda, uv));
        RFB_TRY(lu_device<T>(ctx, dA, n, n, lda, nullptr, info, &o));
        RFB_TRY(rfb_launch_butterfly_vec<T>(ctx, B, n, nrhs, ldb, uv, 0));
        RFB_TRY(rfb_launch_trsm<T>(ctx, dA, n, B, nrhs, lda, &o));
        RFB_TRY(rfb_launch_trsm_upper<T>(ctx, dA, n, B, nrhs, lda, &o));
        return rfb_launch_butterfly_vec<T>(ctx, B, n, nrhs, ldb, uv, 1);
    }
    if (space != RFB_MEM_HOST) return ctx->fail(RFB_ERR_ARG, "unknown mem_space %d", space);
    ctx->kept.valid = false;
    const int64_t np = (n % 4) ? n + (4 - n % 4) : n;      // padded size (:34-38)
    const int64_t ldd = np;
    const size_t need = sizeof(T) * ((size_t)ldd * (size_t)(np + nrhs) + 4 * (size_t)np);
    if (need > ctx->d_mat_cap) {
        if (ctx->d_mat) cudaFree(ctx->d_mat);
        ctx->d_mat = nullptr;
        ctx->d_mat_cap = 0;
        if (cudaMalloc(&ctx->d_mat, need) != cudaSuccess) {
            cudaGetLastError();
            return ctx->fail(RFB_ERR_NOMEM, "cannot allocate %zu bytes of device memory for the butterfly solve", need);
        }
        ctx->d_mat_cap = need;
    }
    T *dA = reinterpret_cast<T *>(ctx->d_mat);
    T *dB = dA + (size_t)ldd * np;
    T *duv = dB + (size_t)ldd * nrhs;
    if (np != n || nrhs > 0) RFB_CUDA(ctx, cudaMemsetAsync(dA, 0, sizeof(T) * (size_t)ldd * (size_t)(np + nrhs), ctx->stream));
    RFB_CUDA(ctx, cudaMemcpy2DAsync(dA, sizeof(T) * ldd, A, sizeof(T) * lda, sizeof(T) * n, n, cudaMemcpyHostToDevice, ctx->stream));
    if (np != n) RFB_TRY(rfb_launch_set_diag<T>(ctx, dA, ldd, n, np, T(1)));     // identity corner of pad! (:193-195)
    if (nrhs > 0)
        RFB_CUDA(ctx, cudaMemcpy2DAsync(dB, sizeof(T) * ldd, B, sizeof(T) * ldb, sizeof(T) * n, nrhs, cudaMemcpyHostToDevice, ctx->stream));
    RFB_CUDA(ctx, cudaMemcpyAsync(duv, uv, sizeof(T) * 4 * (size_t)np, cudaMemcpyHostToDevice, ctx->stream));
    RFB_TRY(rfb_launch_butterfly_mul<T>(ctx, dA, np, ldd, duv));
    RFB_TRY(lu_device<T>(ctx, dA, np, np, ldd, nullptr, ctx->d_info, &o));
    if (nrhs > 0) {
        RFB_TRY(rfb_launch_butterfly_vec<T>(ctx, dB, np, nrhs, ldd, duv, 0));
        RFB_TRY(rfb_launch_trsm<T>(ctx, dA, np, dB, nrhs, ldd, &o));
        RFB_TRY(rfb_launch_trsm_upper<T>(ctx, dA, np, dB, nrhs, ldd, &o));
        RFB_TRY(rfb_launch_butterfly_vec<T>(ctx, dB, np, nrhs, ldd, duv, 1));
        RFB_CUDA(ctx, cudaMemcpy2DAsync(B, sizeof(T) * ldb, dB, sizeof(T) * ldd, sizeof(T) * n, nrhs, cudaMemcpyDeviceToHost, ctx->stream));
    }
    RFB_CUDA(ctx, cudaMemcpyAsync(&ctx->h_pinned[0], ctx->d_info, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    RFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *info = ctx->h_pinned[0];
    return RFB_OK;
}

// lu! applied to each matrix of a strided batch (the reference's sweet spot is many
// small factorizations, README.md:34-35).  Small pivoted matrices (n <= 64, m <= 128) take ONE launch with one
// CTA per matrix -- the unblocked loop the reference itself uses below its threshold (src/lu.jl:125-126);
// anything else runs the recursive driver matrix by matrix on the stream.
template <typename T>
int lu_batched_device(rfb_ctx *ctx, T *dA, int64_t m, int64_t n, int64_t lda, int64_t stride_a, int64_t batch,
                      int64_t *d_ipiv, int64_t *d_info, const rfb_opts *opts) {
    const int64_t mn = m < n ? m : n;
    const bool nopiv = opts && opts->no_pivot;
    RFB_CUDA(ctx, cudaMemsetAsync(d_info, 0, sizeof(int64_t) * (size_t)batch, ctx->stream));
    if (mn == 0) return RFB_OK;
    if (!nopiv && n <= RFB_MAX_NB && m <= 128)
        return rfb_launch_panel_batched<T>(ctx, dA, m, n, lda, stride_a, batch, d_ipiv, d_info);
    for (int64_t b = 0; b < batch; ++b)
        RFB_TRY(lu_device<T>(ctx, dA + b * stride_a, m, n, lda, d_ipiv ? d_ipiv + b * mn : nullptr, d_info + b, opts));
    return RFB_OK;
}

template <typename T>
int lu_batched_entry(rfb_ctx *ctx, T *A, int64_t m, int64_t n, int64_t lda, int64_t stride_a, int64_t batch, int64_t *ipiv,
                     int64_t *info, const rfb_opts *opts) {
    if (!ctx) return RFB_ERR_ARG;
    if (m < 0 || n < 0 || batch < 0) return ctx->fail(RFB_ERR_ARG, "negative dimension");
    if (batch == 0) return RFB_OK;
    if (!info) return ctx->fail(RFB_ERR_ARG, "info is null");
    if (lda < (m > 1 ? m : 1)) return ctx->fail(RFB_ERR_ARG, "lda %lld < max(1, m)", (long long)lda);
    if (batch > 1 && stride_a < lda * n) return ctx->fail(RFB_ERR_ARG, "batch stride %lld < lda * n (matrices overlap)", (long long)stride_a);
    const int64_t mn = m < n ? m : n;
    const bool nopiv = opts && opts->no_pivot;
    if (mn > 0 && (!A || (!ipiv && !nopiv))) return ctx->fail(RFB_ERR_ARG, "A or ipiv is null");
    if (m > 0x7fffffffLL || n > 0x7fffffffLL) return ctx->fail(RFB_ERR_UNSUPPORTED, "dimension exceeds int32");
    RFB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int space = opts ? opts->mem_space : RFB_MEM_HOST;
    if (space == RFB_MEM_DEVICE) return lu_batched_device<T>(ctx, A, m, n, lda, stride_a, batch, ipiv, info, opts);
    if (space != RFB_MEM_HOST) return ctx->fail(RFB_ERR_ARG, "unknown mem_space %d", space);
    if (mn == 0) {
        for (int64_t b = 0; b < batch; ++b) info[b] = 0;
        return RFB_OK;
    }
    ctx->kept.valid = false;
    const int64_t ldd = (m + 1) & ~int64_t(1);
    const int64_t sdd = ldd * n;
    const size_t need = sizeof(T) * (size_t)sdd * (size_t)batch;
    if (need > ctx->d_mat_cap) {
        if (ctx->d_mat) cudaFree(ctx->d_mat);
        ctx->d_mat = nullptr;
        ctx->d_mat_cap = 0;
        if (cudaMalloc(&ctx->d_mat, need) != cudaSuccess) {
            cudaGetLastError();
            return ctx->fail(RFB_ERR_NOMEM, "cannot allocate %zu bytes of device memory for the batch", need);
        }
        ctx->d_mat_cap = need;
    }
    if (!nopiv && (size_t)(mn * batch) > ctx->d_ipiv_cap) {
        if (ctx->d_ipiv) cudaFree(ctx->d_ipiv);
        ctx->d_ipiv = nullptr;
        ctx->d_ipiv_cap = 0;
        if (cudaMalloc(&ctx->d_ipiv, sizeof(int64_t) * (size_t)(mn * batch)) != cudaSuccess) {
            cudaGetLastError();
            return ctx->fail(RFB_ERR_NOMEM, "cannot allocate the device pivot vectors");
        }
        ctx->d_ipiv_cap = (size_t)(mn * batch);
    }
    if ((size_t)batch > ctx->d_binfo_cap) {
        if (ctx->d_binfo) cudaFree(ctx->d_binfo);
        ctx->d_binfo = nullptr;
        ctx->d_binfo_cap = 0;
        if (cudaMalloc(&ctx->d_binfo, sizeof(int64_t) * (size_t)batch) != cudaSuccess) {
            cudaGetLastError();
            return ctx->fail(RFB_ERR_NOMEM, "cannot allocate the device info vector");
        }
        ctx->d_binfo_cap = (size_t)batch;
    }
    T *dA = reinterpret_cast<T *>(ctx->d_mat);
    // a dense batch (lda == m, stride == m * n, m even) is one 1-D copy; otherwise one 2-D copy per matrix
    const bool dense = (lda == ldd && stride_a == sdd);
    if (dense) {
        RFB_CUDA(ctx, cudaMemcpyAsync(dA, A, need, cudaMemcpyHostToDevice, ctx->stream));
    } else if (lda == m && stride_a == m * n) {   // contiguous on the host, padded on the device: one 2-D copy over all columns
        RFB_CUDA(ctx, cudaMemcpy2DAsync(dA, sizeof(T) * ldd, A, sizeof(T) * m, sizeof(T) * m, (size_t)(n * batch), cudaMemcpyHostToDevice, ctx->stream));
    } else {
        for (int64_t b = 0; b < batch; ++b)
            RFB_CUDA(ctx, cudaMemcpy2DAsync(dA + b * sdd, sizeof(T) * ldd, A + b * stride_a, sizeof(T) * lda, sizeof(T) * m, n,
                                            cudaMemcpyHostToDevice, ctx->stream));
    }
    RFB_TRY(lu_batched_device<T>(ctx, dA, m, n, ldd, sdd, batch, nopiv ? nullptr : ctx->d_ipiv, ctx->d_binfo, opts));
    if (dense) {
        RFB_CUDA(ctx, cudaMemcpyAsync(A, dA, need, cudaMemcpyDeviceToHost, ctx->stream));
    } else if (lda == m && stride_a == m * n) {
        RFB_CUDA(ctx, cudaMemcpy2DAsync(A, sizeof(T) * m, dA, sizeof(T) * ldd, sizeof(T) * m, (size_t)(n * batch), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        for (int64_t b = 0; b < batch; ++b)
            RFB_CUDA(ctx, cudaMemcpy2DAsync(A + b * stride_a, sizeof(T) * lda, dA + b * sdd, sizeof(T) * ldd, sizeof(T) * m, n,
                                            cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (nopiv) {
        if (ipiv) for (int64_t b = 0; b < batch; ++b) for (int64_t i = 0; i < mn; ++i) ipiv[b * mn + i] = i + 1;
    } else {
        RFB_CUDA(ctx, cudaMemcpyAsync(ipiv, ctx->d_ipiv, sizeof(int64_t) * (size_t)(mn * batch), cudaMemcpyDeviceToHost, ctx->stream));
    }
    RFB_CUDA(ctx, cudaMemcpyAsync(info, ctx->d_binfo, sizeof(int64_t) * (size_t)batch, cudaMemcpyDeviceToHost, ctx->stream));
    RFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RFB_OK;
}

template <typename T>
int lu_entry(rfb_ctx *ctx, T *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv, int64_t *info,
             const rfb_opts *opts) {
    if (!ctx) return RFB_ERR_ARG;
    if (m < 0 || n < 0) return ctx->fail(RFB_ERR_ARG, "negative dimension (%lld x %lld)", (long long)m, (long long)n);
    if (lda < (m > 1 ? m : 1)) return ctx->fail(RFB_ERR_ARG, "lda %lld < max(1, m) with m = %lld", (long long)lda, (long long)m);
    if (!info) return ctx->fail(RFB_ERR_ARG, "info is null");
    const int64_t mn = m < n ? m : n;
    const bool nopiv = opts && opts->no_pivot;        // ipiv may be NULL then (NotIPIV, src/lu.jl:27-32)
    if (mn > 0 && (!A || (!ipiv && !nopiv))) return ctx->fail(RFB_ERR_ARG, "A or ipiv is null");
    if (m > 0x7fffffffLL || n > 0x7fffffffLL) return ctx->fail(RFB_ERR_UNSUPPORTED, "dimension exceeds int32");
    RFB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int space = opts ? opts->mem_space : RFB_MEM_HOST;
    if (mn == 0) {
        if (space == RFB_MEM_HOST) *info = 0;
        else RFB_CUDA(ctx, cudaMemsetAsync(info, 0, sizeof(int64_t), ctx->stream));
        return RFB_OK;
    }
    if (space == RFB_MEM_DEVICE) return lu_device<T>(ctx, A, m, n, lda, ipiv, info, opts);
    if (space != RFB_MEM_HOST) return ctx->fail(RFB_ERR_ARG, "unknown mem_space %d", space);

    // host pointers: stage through a grow-only device buffer with a dense leading dimension
    ctx->kept.valid = false;
    const int64_t ldd = (m + 1) & ~int64_t(1);          // even => 16-byte aligned columns for f64
    const size_t need = sizeof(T) * (size_t)ldd * (size_t)n;
    if (need > ctx->d_mat_cap) {
        if (ctx->d_mat) cudaFree(ctx->d_mat);
        ctx->d_mat = nullptr;
        ctx->d_mat_cap = 0;
        if (cudaMalloc(&ctx->d_mat, need) != cudaSuccess) {
            cudaGetLastError();
            return ctx->fail(RFB_ERR_NOMEM, "cannot allocate %zu bytes of device memory for the matrix", need);
        }
        ctx->d_mat_cap = need;
    }
    if (!nopiv && (size_t)mn > ctx->d_ipiv_cap) {
        if (ctx->d_ipiv) cudaFree(ctx->d_ipiv);
        ctx->d_ipiv = nullptr;
        ctx->d_ipiv_cap = 0;
        if (cudaMalloc(&ctx->d_ipiv, sizeof(int64_t) * (size_t)mn) != cudaSuccess) {
            cudaGetLastError();
            return ctx->fail(RFB_ERR_NOMEM, "cannot allocate the device pivot vector");
        }
        ctx->d_ipiv_cap = (size_t)mn;
    }
    T *dA = reinterpret_cast<T *>(ctx->d_mat);
    // Pipelined upload: column chunks go up on the copy stream in order; the factorization is
    // left-looking, so it starts as soon as the first chunk is resident and the rest of the upload
    // hides behind the work on the left columns.
    // The first panel only waits for the first 8 MB, and the next 15 chunks stay at 8 MB too: at the start a 64-column
    // leaf step costs about what its columns take to upload (~150 us at 16384 rows), and a chunk is only usable when
    // ALL of it has arrived -- doubling the chunks right away made each of the first steps wait for a chunk twice as
    // long as the one before (~0.5 ms in total at 16384^2).  After that the factorization is far ahead of its need
    // and the chunks grow geometrically to 64 MB.
    std::vector<int64_t> bounds;
    {
        const size_t col_bytes = sizeof(T) * (size_t)ldd;
        size_t target = size_t(8) << 20;
        int64_t j = 0;
        while (j < n) {
            const int64_t nc = std::max<int64_t>(64, (int64_t)(target / col_bytes));
            j = std::min<int64_t>(n, j + nc);
            bounds.push_back(j);
            if (bounds.size() >= 16 && target < (size_t(64) << 20)) target *= 2;
        }
    }
    const int nchunks = (int)bounds.size();
    while ((int)ctx->up_events.size() < nchunks) {
        cudaEvent_t e;
        RFB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->up_events.push_back(e);
    }
    RFB_CUDA(ctx, cudaEventRecord(ctx->ev_sync, ctx->stream));          // the staging buffer is free again
    RFB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_sync, 0));
    for (int c = 0; c < nchunks; ++c) {
        const int64_t j0 = c ? bounds[c - 1] : 0, nc = bounds[c] - j0;
        RFB_CUDA(ctx, cudaMemcpy2DAsync(dA + j0 * ldd, sizeof(T) * ldd, A + j0 * lda, sizeof(T) * lda, sizeof(T) * m,
                                        nc, cudaMemcpyHostToDevice, ctx->copy_stream));
        RFB_CUDA(ctx, cudaEventRecord(ctx->up_events[c], ctx->copy_stream));
    }
    std::vector<cudaEvent_t> evs(ctx->up_events.begin(), ctx->up_events.begin() + nchunks);
    int64_t early_rows = 0;
    // the early download is only worth it (and only asynchronous) when the caller's matrix is page-locked
    cudaPointerAttributes pattr;
    const bool pinned = cudaPointerGetAttributes(&pattr, A) == cudaSuccess && pattr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    {
        const int rc = lu_device<T>(ctx, dA, m, n, ldd, nopiv ? nullptr : ctx->d_ipiv, ctx->d_info, opts, &evs, &bounds,
                                    (pinned && m >= n) ? A : nullptr, lda, &early_rows);
        if (rc != RFB_OK) {
            // uploads from / early downloads into the caller's matrix may still be in flight: the library never keeps a
            // host pointer past the call, so drain both streams before reporting the failure
            cudaStreamSynchronize(ctx->copy_stream);
            cudaStreamSynchronize(ctx->down_stream);
            cudaStreamSynchronize(ctx->stream);
            cudaGetLastError();
            return rc;
        }
    }
    RFB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, evs[nchunks - 1], 0));   // (already implied; keeps the order explicit)
    // download what the early copies (rows [0, early_rows) of all columns; everything in tile mode) did not cover
    if (early_rows > 0) {
        if (early_rows < m)
            RFB_CUDA(ctx, cudaMemcpy2DAsync(A + early_rows, sizeof(T) * lda, dA + early_rows, sizeof(T) * ldd,
                                            sizeof(T) * (m - early_rows), n, cudaMemcpyDeviceToHost, ctx->stream));
        RFB_CUDA(ctx, cudaEventRecord(ctx->ev_sync, ctx->down_stream));
        RFB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_sync, 0));    // the early copies must be done before we return
    } else {
        RFB_CUDA(ctx, cudaMemcpy2DAsync(A, sizeof(T) * lda, dA, sizeof(T) * ldd, sizeof(T) * m, n,
                                        cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (nopiv) {
        if (ipiv) for (int64_t i = 0; i < mn; ++i) ipiv[i] = i + 1;            // src/lu.jl:107-113
    } else {
        RFB_CUDA(ctx, cudaMemcpyAsync(ipiv, ctx->d_ipiv, sizeof(int64_t) * mn, cudaMemcpyDeviceToHost, ctx->stream));
    }
    RFB_CUDA(ctx, cudaMemcpyAsync(&ctx->h_pinned[0], ctx->d_info, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    RFB_CUDA(ctx, cudaMemcpyAsync(&ctx->h_pinned[1], &ctx->xchg->error_flag, sizeof(unsigned int),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    RFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *info = ctx->h_pinned[0];
    if ((unsigned int)ctx->h_pinned[1] != 0) {
        cudaMemset(&ctx->xchg->error_flag, 0, sizeof(unsigned int));   // report once, keep the context usable
        return ctx->fail(RFB_ERR_INTERNAL, "device-side protocol error (flag set: 1 = panel exchange timed out, 2 = row-exchange lists incomplete)");
    }
    if (opts && opts->keep_factors && m == n) {          // the factors stay where they are for rfb_solve_kept_*
        ctx->kept.valid = true;
        ctx->kept.f32 = sizeof(T) == 4;
        ctx->kept.nopiv = nopiv;
        ctx->kept.n = n;
        ctx->kept.ldd = ldd;
        ctx->kept.id = ++ctx->kept_counter;
    }
    return RFB_OK;
}

// ldiv!(F, B) with the factors a host-mode rfb_lu_* (keep_factors) left on the device: only B travels
template <typename T>
int solve_kept(rfb_ctx *ctx, int64_t id, T *B, int64_t nrhs, int64_t ldb) {
    if (!ctx) return RFB_ERR_ARG;
    if (!ctx->kept.valid || ctx->kept.id != id || ctx->kept.f32 != (sizeof(T) == 4))
        return ctx->fail(RFB_ERR_ARG, "rfb_solve_kept: no factors with id %lld are resident (a later call reused the staging buffer)", (long long)id);
    const int64_t n = ctx->kept.n, ldd = ctx->kept.ldd;
    if (nrhs < 0) return ctx->fail(RFB_ERR_ARG, "negative nrhs");
    if (nrhs == 0 || n == 0) return RFB_OK;
    if (!B || ldb < n) return ctx->fail(RFB_ERR_ARG, "B is null or ldb < n");
    RFB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t need = sizeof(T) * (size_t)ldd * (size_t)nrhs;
    if (need > ctx->d_rhs_cap) {
        if (ctx->d_rhs) cudaFree(ctx->d_rhs);
        ctx->d_rhs = nullptr;
        ctx->d_rhs_cap = 0;
        if (cudaMalloc(&ctx->d_rhs, need) != cudaSuccess) {
            cudaGetLastError();
            return ctx->fail(RFB_ERR_NOMEM, "cannot allocate %zu bytes for the right-hand sides", need);
        }
        ctx->d_rhs_cap = need;
    }
    const T *dLU = reinterpret_cast<const T *>(ctx->d_mat);
    T *dB = reinterpret_cast<T *>(ctx->d_rhs);
    RFB_CUDA(ctx, cudaMemcpy2DAsync(dB, sizeof(T) * ldd, B, sizeof(T) * ldb, sizeof(T) * n, nrhs, cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->kept.nopiv) RFB_TRY(rfb_launch_laswp<T>(ctx, dB, nrhs, ldd, ctx->d_ipiv, n, 0));
    RFB_TRY(rfb_launch_trsm<T>(ctx, dLU, n, dB, nrhs, ldd, &ctx->default_opts));
    RFB_TRY(rfb_launch_trsm_upper<T>(ctx, dLU, n, dB, nrhs, ldd, &ctx->default_opts));
    RFB_CUDA(ctx, cudaMemcpy2DAsync(B, sizeof(T) * ldb, dB, sizeof(T) * ldd, sizeof(T) * n, nrhs, cudaMemcpyDeviceToHost, ctx->stream));
    RFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RFB_OK;
}

// A host-mode call that fails after it has enqueued asynchronous copies from / into the caller's buffers must not return
// while a DMA can still touch them (the library never keeps a host pointer past a call): drain every stream of the context.
static int drained(rfb_ctx *ctx, int rc) {
    if (rc != RFB_OK && ctx && ctx->stream && !ctx->dry_run) {
        cudaSetDevice(ctx->device);
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        if (ctx->down_stream) cudaStreamSynchronize(ctx->down_stream);
        cudaStreamSynchronize(ctx->stream);
        cudaGetLastError();
    }
    return rc;
}

}  // namespace

extern "C" {

int rfb_version(void) { return 100; }

int rfb_create(rfb_ctx **out, int device) {
    if (!out) return RFB_ERR_ARG;
    *out = nullptr;
    rfb_ctx *ctx = new rfb_ctx();
    *out = ctx;   // returned even on failure so that rfb_last_error works; caller still destroys it
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return ctx->fail(RFB_ERR_CUDA, "no CUDA device available (%s); librfb200 has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return ctx->fail(RFB_ERR_ARG, "device %d out of range [0, %d)", device, ndev);
    ctx->device = device;
    RFB_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    RFB_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->mem_bytes = prop.totalGlobalMem;
    if (prop.major != 10)
        return ctx->fail(RFB_ERR_UNSUPPORTED, "device %d is sm_%d%d; librfb200 is built for sm_100a only", device,
                         prop.major, prop.minor);
    RFB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    RFB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    RFB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
    if (const char *e = getenv("RFB_EARLY_DOWNLOAD")) ctx->early_mode = atoi(e);      // A/B switch, see lu_rec
    if (ctx->early_mode < 0 || ctx->early_mode > 2) ctx->early_mode = 2;
    RFB_CUDA(ctx, cudaEventCreate(&ctx->ev_start));
    RFB_CUDA(ctx, cudaEventCreate(&ctx->ev_stop));
    RFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_sync, cudaEventDisableTiming));
    RFB_CUDA(ctx, cudaMalloc(&ctx->xchg, sizeof(RfbPanelXchg)));
    RFB_CUDA(ctx, cudaMemset(ctx->xchg, 0, sizeof(RfbPanelXchg)));
    if (const char *e = getenv("RFB_GEMM_EPILOGUE")) ctx->gemm_reduce_epilogue = atoi(e) != 0;
    if (const char *e = getenv("RFB_LASWP_NET_MIN")) ctx->laswp_net_min = atoll(e);
    if (const char *e = getenv("RFB_LASWP_NET_CAP")) ctx->laswp_net_cap = atoll(e);
    for (int l = 0; l < rfb_ctx::kLanes; ++l) {
        RFB_CUDA(ctx, cudaMalloc(&ctx->net_meta_[l], 64));
        RFB_CUDA(ctx, cudaMemset(ctx->net_meta_[l], 0, 64));
        RFB_CUDA(ctx, cudaMalloc(&ctx->net_srcmap_[l], sizeof(int) * 8192));
        RFB_CUDA(ctx, cudaMalloc(&ctx->net_clist_[l], sizeof(int) * 2 * 8192));
    }
    RFB_CUDA(ctx, cudaMalloc(&ctx->d_info, 64));
    RFB_CUDA(ctx, cudaMemset(ctx->d_info, 0, 64));
    RFB_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_pinned, 64, cudaHostAllocDefault));
    memset(ctx->h_pinned, 0, 64);
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ctx->encode_tiled, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        ctx->encode_tiled = nullptr;   // the TMA GEMM path then reports "not handled"
    }
    return RFB_OK;
}

int rfb_destroy(rfb_ctx *ctx) {
    if (!ctx) return RFB_OK;
    if (ctx->stream) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->up_events) cudaEventDestroy(e);
    if (ctx->xchg) cudaFree(ctx->xchg);
    for (int l = 0; l < rfb_ctx::kLanes; ++l) {
        if (ctx->net_meta_[l]) cudaFree(ctx->net_meta_[l]);
        if (ctx->net_srcmap_[l]) cudaFree(ctx->net_srcmap_[l]);
        if (ctx->net_clist_[l]) cudaFree(ctx->net_clist_[l]);
    }
    if (ctx->d_info) cudaFree(ctx->d_info);
    if (ctx->d_ipiv) cudaFree(ctx->d_ipiv);
    if (ctx->d_binfo) cudaFree(ctx->d_binfo);
    if (ctx->d_mat) cudaFree(ctx->d_mat);
    if (ctx->d_rhs) cudaFree(ctx->d_rhs);
    if (!ctx->perm_external) {
        if (ctx->perm_dst) cudaFree(ctx->perm_dst);
        if (ctx->perm_src) cudaFree(ctx->perm_src);
        if (ctx->perm_width) cudaFree(ctx->perm_width);
    }
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->ev_sync) cudaEventDestroy(ctx->ev_sync);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->down_stream) cudaStreamDestroy(ctx->down_stream);
    delete ctx;
    return RFB_OK;
}

const char *rfb_last_error(rfb_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int rfb_device_info(rfb_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *mem_bytes) {
    if (!ctx) return RFB_ERR_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (mem_bytes) *mem_bytes = ctx->mem_bytes;
    return RFB_OK;
}

int rfb_set_default_opts(rfb_ctx *ctx, const rfb_opts *opts) {
    if (!ctx) return RFB_ERR_ARG;
    if (opts) ctx->default_opts = *opts;
    else ctx->default_opts = rfb_opts{};
    return RFB_OK;
}

int rfb_set_early_download(rfb_ctx *ctx, int mode) {
    if (!ctx) return RFB_ERR_ARG;
    if (mode < 0 || mode > 2) return ctx->fail(RFB_ERR_ARG, "early-download mode must be 0, 1 or 2 (got %d)", mode);
    ctx->early_mode = mode;
    return RFB_OK;
}

int rfb_trace_lu(int is_f32, int64_t m, int64_t n, int64_t lda, const rfb_opts *opts, int pinned_host, int64_t *ops,
                 int64_t cap, int64_t *count) {
    if (!count || m < 0 || n < 0 || lda < (m > 1 ? m : 1)) return RFB_ERR_ARG;
    rfb_ctx ctx;                                   // never touches CUDA: dry_run short-circuits every launcher
    ctx.dry_run = true;
    ctx.trace_lda = lda;
    ctx.trace_elt = is_f32 ? 4 : 8;
    char *base = reinterpret_cast<char *>(uintptr_t(1) << 40);       // fake address, never dereferenced
    ctx.trace_base = base;
    if (const char *e = getenv("RFB_EARLY_DOWNLOAD")) ctx.early_mode = atoi(e);
    if (pinned_host >= 10) ctx.early_mode = pinned_host - 10;
    if (ctx.early_mode < 0 || ctx.early_mode > 2) ctx.early_mode = 2;
    int64_t *fake_piv = reinterpret_cast<int64_t *>(uintptr_t(1) << 39);
    int rc;
    if (is_f32) {
        float *A = reinterpret_cast<float *>(base);
        rc = lu_device<float>(&ctx, A, m, n, lda, (opts && opts->no_pivot) ? nullptr : fake_piv, fake_piv, opts, nullptr, nullptr,
                              pinned_host ? A : nullptr, lda, nullptr);
    } else {
        double *A = reinterpret_cast<double *>(base);
        rc = lu_device<double>(&ctx, A, m, n, lda, (opts && opts->no_pivot) ? nullptr : fake_piv, fake_piv, opts, nullptr, nullptr,
                               pinned_host ? A : nullptr, lda, nullptr);
    }
    if (rc != RFB_OK) return rc;
    *count = (int64_t)ctx.trace.size();
    if (ops) {
        const int64_t nout = *count < cap ? *count : cap;
        for (int64_t i = 0; i < nout; ++i)
            for (int j = 0; j < 8; ++j) ops[8 * i + j] = ctx.trace[(size_t)i].v[j];
    }
    return RFB_OK;
}

int rfb_lu_f64(rfb_ctx *ctx, double *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv, int64_t *info,
               const rfb_opts *opts) {
    return lu_entry<double>(ctx, A, m, n, lda, ipiv, info, opts);
}

int rfb_lu_range_f64(rfb_ctx *ctx, double *A_root, int64_t m, int64_t lda, int64_t c0, int64_t n, int64_t *ipiv_dev,
                     int64_t *info_dev, const rfb_opts *opts) {
    if (!ctx) return RFB_ERR_ARG;
    return lu_range<double>(ctx, A_root, m, lda, c0, n, ipiv_dev, info_dev, opts);
}
int rfb_lu_range_f32(rfb_ctx *ctx, float *A_root, int64_t m, int64_t lda, int64_t c0, int64_t n, int64_t *ipiv_dev,
                     int64_t *info_dev, const rfb_opts *opts) {
    if (!ctx) return RFB_ERR_ARG;
    return lu_range<float>(ctx, A_root, m, lda, c0, n, ipiv_dev, info_dev, opts);
}
int rfb_laswp_range_f64(rfb_ctx *ctx, double *A_root, int64_t lda, int64_t col0, int64_t ncols, int64_t k0, int64_t k1,
                        const int64_t *ipiv_dev, int use_lists) {
    if (!ctx) return RFB_ERR_ARG;
    return laswp_range<double>(ctx, A_root, lda, col0, ncols, k0, k1, ipiv_dev, use_lists);
}
int rfb_laswp_range_f32(rfb_ctx *ctx, float *A_root, int64_t lda, int64_t col0, int64_t ncols, int64_t k0, int64_t k1,
                        const int64_t *ipiv_dev, int use_lists) {
    if (!ctx) return RFB_ERR_ARG;
    return laswp_range<float>(ctx, A_root, lda, col0, ncols, k0, k1, ipiv_dev, use_lists);
}
int rfb_perm_buffers(rfb_ctx *ctx, int32_t *dst_dev, int32_t *src_dev, int32_t *width_dev, int64_t cap) {
    if (!ctx) return RFB_ERR_ARG;
    if (!dst_dev || !src_dev || !width_dev || cap <= 0) return ctx->fail(RFB_ERR_ARG, "rfb_perm_buffers: null buffer or cap <= 0");
    if (!ctx->perm_external) {
        cudaFree(ctx->perm_dst); cudaFree(ctx->perm_src); cudaFree(ctx->perm_width);
    }
    ctx->perm_dst = dst_dev; ctx->perm_src = src_dev; ctx->perm_width = width_dev;
    ctx->perm_cap = (size_t)cap;
    ctx->perm_external = true;
    RFB_CUDA(ctx, cudaMemsetAsync(ctx->perm_dst, 0xFF, 2 * (size_t)cap * sizeof(int), ctx->stream));
    RFB_CUDA(ctx, cudaMemsetAsync(ctx->perm_src, 0xFF, 2 * (size_t)cap * sizeof(int), ctx->stream));
    RFB_CUDA(ctx, cudaMemsetAsync(ctx->perm_width, 0, (size_t)cap * sizeof(int), ctx->stream));
    return RFB_OK;
}
int rfb_perm_buffers_release(rfb_ctx *ctx) {
    if (!ctx) return RFB_ERR_ARG;
    RFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->perm_external) {
        ctx->perm_dst = ctx->perm_src = ctx->perm_width = nullptr;
        ctx->perm_cap = 0;
        ctx->perm_external = false;
    }
    return RFB_OK;
}
int rfb_copy2d(rfb_ctx *ctx, void *dst_dev, size_t dpitch, const void *src_dev, size_t spitch, size_t width_bytes,
               size_t height) {
    if (!ctx) return RFB_ERR_ARG;
    if (width_bytes == 0 || height == 0) return RFB_OK;
    RFB_CUDA(ctx, cudaMemcpy2DAsync(dst_dev, dpitch, src_dev, spitch, width_bytes, height, cudaMemcpyDeviceToDevice, ctx->stream));
    return RFB_OK;
}
int rfb_set_stream(rfb_ctx *ctx, void *cuda_stream) {
    if (!ctx) return RFB_ERR_ARG;
    RFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return RFB_OK;
}
int rfb_lu_f32(rfb_ctx *ctx, float *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv, int64_t *info,
               const rfb_opts *opts) {
    return lu_entry<float>(ctx, A, m, n, lda, ipiv, info, opts);
}

#define RFB_CHECK_CTX(ctx) \
    if (!(ctx)) return RFB_ERR_ARG

int rfb_panel_getrf_f64(rfb_ctx *ctx, double *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv_dev,
                        int64_t ipiv_add, int64_t *info_dev, int64_t col_offset) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_panel<double>(ctx, A, m, n, lda, ipiv_dev, ipiv_add, info_dev, col_offset);
}
int rfb_panel_getrf_f32(rfb_ctx *ctx, float *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv_dev,
                        int64_t ipiv_add, int64_t *info_dev, int64_t col_offset) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_panel<float>(ctx, A, m, n, lda, ipiv_dev, ipiv_add, info_dev, col_offset);
}
int rfb_laswp_f64(rfb_ctx *ctx, double *A, int64_t ncols, int64_t lda, const int64_t *ipiv_dev, int64_t npiv,
                  int64_t ipiv_sub) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_laswp<double>(ctx, A, ncols, lda, ipiv_dev, npiv, ipiv_sub);
}
int rfb_laswp_f32(rfb_ctx *ctx, float *A, int64_t ncols, int64_t lda, const int64_t *ipiv_dev, int64_t npiv,
                  int64_t ipiv_sub) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_laswp<float>(ctx, A, ncols, lda, ipiv_dev, npiv, ipiv_sub);
}
int rfb_trsm_llnu_f64(rfb_ctx *ctx, const double *L, int64_t k, double *B, int64_t nrhs, int64_t lda) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_trsm<double>(ctx, L, k, B, nrhs, lda, &ctx->default_opts);
}
int rfb_trsm_llnu_f32(rfb_ctx *ctx, const float *L, int64_t k, float *B, int64_t nrhs, int64_t lda) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_trsm<float>(ctx, L, k, B, nrhs, lda, &ctx->default_opts);
}
int rfb_gemm_nn_sub_f64(rfb_ctx *ctx, double *C, const double *A, const double *B, int64_t m, int64_t n,
                        int64_t k, int64_t lda) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_gemm<double>(ctx, C, A, B, m, n, k, lda, &ctx->default_opts);
}
int rfb_gemm_nn_sub_f32(rfb_ctx *ctx, float *C, const float *A, const float *B, int64_t m, int64_t n, int64_t k,
                        int64_t lda) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_gemm<float>(ctx, C, A, B, m, n, k, lda, &ctx->default_opts);
}
int rfb_trsm_lunn_f64(rfb_ctx *ctx, const double *U, int64_t k, double *B, int64_t nrhs, int64_t lda) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_trsm_upper<double>(ctx, U, k, B, nrhs, lda, &ctx->default_opts);
}
int rfb_trsm_lunn_f32(rfb_ctx *ctx, const float *U, int64_t k, float *B, int64_t nrhs, int64_t lda) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_trsm_upper<float>(ctx, U, k, B, nrhs, lda, &ctx->default_opts);
}
int rfb_solve_f64(rfb_ctx *ctx, const double *LU, int64_t n, int64_t lda, const int64_t *ipiv, double *B, int64_t nrhs,
                  int64_t ldb, const rfb_opts *opts) {
    return drained(ctx, solve_entry<double>(ctx, LU, n, lda, ipiv, B, nrhs, ldb, opts));
}
int rfb_solve_f32(rfb_ctx *ctx, const float *LU, int64_t n, int64_t lda, const int64_t *ipiv, float *B, int64_t nrhs,
                  int64_t ldb, const rfb_opts *opts) {
    return drained(ctx, solve_entry<float>(ctx, LU, n, lda, ipiv, B, nrhs, ldb, opts));
}
int rfb_kept_id(rfb_ctx *ctx, int64_t *id) {
    if (!ctx || !id) return RFB_ERR_ARG;
    *id = ctx->kept.valid ? ctx->kept.id : 0;
    return RFB_OK;
}
int rfb_solve_kept_f64(rfb_ctx *ctx, int64_t id, double *B, int64_t nrhs, int64_t ldb) { return drained(ctx, solve_kept<double>(ctx, id, B, nrhs, ldb)); }
int rfb_solve_kept_f32(rfb_ctx *ctx, int64_t id, float *B, int64_t nrhs, int64_t ldb) { return drained(ctx, solve_kept<float>(ctx, id, B, nrhs, ldb)); }
int rfb_panel_getrf_nopiv_f64(rfb_ctx *ctx, double *A, int64_t m, int64_t n, int64_t lda, int64_t *info_dev, int64_t col_offset) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_panel_nopiv<double>(ctx, A, m, n, lda, info_dev, col_offset);
}
int rfb_panel_getrf_nopiv_f32(rfb_ctx *ctx, float *A, int64_t m, int64_t n, int64_t lda, int64_t *info_dev, int64_t col_offset) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_panel_nopiv<float>(ctx, A, m, n, lda, info_dev, col_offset);
}
int rfb_butterfly_mul_f64(rfb_ctx *ctx, double *A, int64_t n, int64_t lda, const double *uv_dev) {
    RFB_CHECK_CTX(ctx);
    if (!A || !uv_dev) return ctx->fail(RFB_ERR_ARG, "null pointer");
    return rfb_launch_butterfly_mul<double>(ctx, A, n, lda, uv_dev);
}
int rfb_butterfly_mul_f32(rfb_ctx *ctx, float *A, int64_t n, int64_t lda, const float *uv_dev) {
    RFB_CHECK_CTX(ctx);
    if (!A || !uv_dev) return ctx->fail(RFB_ERR_ARG, "null pointer");
    return rfb_launch_butterfly_mul<float>(ctx, A, n, lda, uv_dev);
}
int rfb_butterfly_vec_f64(rfb_ctx *ctx, double *B, int64_t n, int64_t nrhs, int64_t ldb, const double *uv_dev, int which) {
    RFB_CHECK_CTX(ctx);
    if (!B || !uv_dev || (which != 0 && which != 1)) return ctx->fail(RFB_ERR_ARG, "null pointer or which not in {0, 1}");
    return rfb_launch_butterfly_vec<double>(ctx, B, n, nrhs, ldb, uv_dev, which);
}
int rfb_butterfly_vec_f32(rfb_ctx *ctx, float *B, int64_t n, int64_t nrhs, int64_t ldb, const float *uv_dev, int which) {
    RFB_CHECK_CTX(ctx);
    if (!B || !uv_dev || (which != 0 && which != 1)) return ctx->fail(RFB_ERR_ARG, "null pointer or which not in {0, 1}");
    return rfb_launch_butterfly_vec<float>(ctx, B, n, nrhs, ldb, uv_dev, which);
}
int rfb_butterfly_solve_f64(rfb_ctx *ctx, const double *A, int64_t n, int64_t lda, double *B, int64_t nrhs, int64_t ldb,
                            const double *uv, int64_t *info, const rfb_opts *opts) {
    return drained(ctx, butterfly_solve_entry<double>(ctx, A, n, lda, B, nrhs, ldb, uv, info, opts));
}
int rfb_butterfly_solve_f32(rfb_ctx *ctx, const float *A, int64_t n, int64_t lda, float *B, int64_t nrhs, int64_t ldb,
                            const float *uv, int64_t *info, const rfb_opts *opts) {
    return drained(ctx, butterfly_solve_entry<float>(ctx, A, n, lda, B, nrhs, ldb, uv, info, opts));
}
int rfb_lu_batched_f64(rfb_ctx *ctx, double *A, int64_t m, int64_t n, int64_t lda, int64_t stride_a, int64_t batch,
                       int64_t *ipiv, int64_t *info, const rfb_opts *opts) {
    return drained(ctx, lu_batched_entry<double>(ctx, A, m, n, lda, stride_a, batch, ipiv, info, opts));
}
int rfb_lu_batched_f32(rfb_ctx *ctx, float *A, int64_t m, int64_t n, int64_t lda, int64_t stride_a, int64_t batch,
                       int64_t *ipiv, int64_t *info, const rfb_opts *opts) {
    return drained(ctx, lu_batched_entry<float>(ctx, A, m, n, lda, stride_a, batch, ipiv, info, opts));
}
int rfb_ipiv_shift(rfb_ctx *ctx, int64_t *ipiv_dev, int64_t n, int64_t shift) {
    RFB_CHECK_CTX(ctx);
    return rfb_launch_ipiv_shift(ctx, ipiv_dev, n, shift);
}

int rfb_malloc(rfb_ctx *ctx, void **dev_ptr, size_t bytes) {
    RFB_CHECK_CTX(ctx);
    if (!dev_ptr) return ctx->fail(RFB_ERR_ARG, "dev_ptr is null");
    *dev_ptr = nullptr;
    cudaSetDevice(ctx->device);
    if (cudaMalloc(dev_ptr, bytes ? bytes : 16) != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(RFB_ERR_NOMEM, "cudaMalloc(%zu) failed", bytes);
    }
    return RFB_OK;
}
int rfb_free(rfb_ctx *ctx, void *dev_ptr) {
    RFB_CHECK_CTX(ctx);
    RFB_CUDA(ctx, cudaFree(dev_ptr));
    return RFB_OK;
}
int rfb_host_alloc(rfb_ctx *ctx, void **host_ptr, size_t bytes) {
    RFB_CHECK_CTX(ctx);
    if (!host_ptr) return ctx->fail(RFB_ERR_ARG, "host_ptr is null");
    *host_ptr = nullptr;
    if (cudaHostAlloc(host_ptr, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(RFB_ERR_NOMEM, "cudaHostAlloc(%zu) failed", bytes);
    }
    return RFB_OK;
}
int rfb_host_free(rfb_ctx *ctx, void *host_ptr) {   // ctx may be NULL (the block outlives its context)
    if (!ctx) return cudaFreeHost(host_ptr) == cudaSuccess ? RFB_OK : RFB_ERR_CUDA;
    RFB_CUDA(ctx, cudaFreeHost(host_ptr));
    return RFB_OK;
}
int rfb_h2d(rfb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    RFB_CHECK_CTX(ctx);
    RFB_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return RFB_OK;
}
int rfb_d2h(rfb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    RFB_CHECK_CTX(ctx);
    RFB_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return RFB_OK;
}
int rfb_d2d(rfb_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes) {
    RFB_CHECK_CTX(ctx);
    RFB_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return RFB_OK;
}
int rfb_memset(rfb_ctx *ctx, void *dst_dev, int value, size_t bytes) {
    RFB_CHECK_CTX(ctx);
    RFB_CUDA(ctx, cudaMemsetAsync(dst_dev, value, bytes, ctx->stream));
    return RFB_OK;
}
int rfb_sync(rfb_ctx *ctx) {
    RFB_CHECK_CTX(ctx);
    RFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned int flag = 0;
    RFB_CUDA(ctx, cudaMemcpy(&flag, &ctx->xchg->error_flag, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag) {
        cudaMemset(&ctx->xchg->error_flag, 0, sizeof(unsigned int));       // report once, keep the context usable
        return ctx->fail(RFB_ERR_INTERNAL, "device-side protocol error (flag set: 1 = panel exchange timed out, 2 = row-exchange lists incomplete)");
    }
    return RFB_OK;
}

int rfb_timer_start(rfb_ctx *ctx) {
    RFB_CHECK_CTX(ctx);
    RFB_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    return RFB_OK;
}
int rfb_timer_stop(rfb_ctx *ctx, float *ms) {
    RFB_CHECK_CTX(ctx);
    RFB_CUDA(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
    RFB_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop));
    float t = 0;
    RFB_CUDA(ctx, cudaEventElapsedTime(&t, ctx->ev_start, ctx->ev_stop));
    if (ms) *ms = t;
    return RFB_OK;
}
int rfb_launch_count(rfb_ctx *ctx, int64_t *count) {
    RFB_CHECK_CTX(ctx);
    if (count) *count = ctx->launches;
    return RFB_OK;
}
int rfb_profile_enable(rfb_ctx *ctx, int on) {
    RFB_CHECK_CTX(ctx);
    RFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    ctx->prof_events.clear();
    ctx->prof_classes.clear();
    for (int i = 0; i < RFB_KC_COUNT; ++i) { ctx->prof_ms[i] = 0; ctx->prof_launches[i] = 0; ctx->prof_work[i] = 0; }
    ctx->profiling = on != 0;
    return RFB_OK;
}
int rfb_profile_read(rfb_ctx *ctx, double ms_by_class[8], int64_t launches_by_class[8], double work_by_class[8]) {
    RFB_CHECK_CTX(ctx);
    RFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < ctx->prof_classes.size(); ++i) {
        float t = 0;
        cudaEventElapsedTime(&t, ctx->prof_events[2 * i], ctx->prof_events[2 * i + 1]);
        ctx->prof_ms[ctx->prof_classes[i]] += t;
        cudaEventDestroy(ctx->prof_events[2 * i]);
        cudaEventDestroy(ctx->prof_events[2 * i + 1]);
    }
    ctx->prof_events.clear();
    ctx->prof_classes.clear();
    for (int i = 0; i < RFB_KC_COUNT; ++i) {
        if (ms_by_class) ms_by_class[i] = ctx->prof_ms[i];
        if (launches_by_class) launches_by_class[i] = ctx->prof_launches[i];
        if (work_by_class) work_by_class[i] = ctx->prof_work[i];
    }
    return RFB_OK;
}
int rfb_bench_dmma_peak(rfb_ctx *ctx, int iters, double *tflops) {
    RFB_CHECK_CTX(ctx);
    if (!tflops) return ctx->fail(RFB_ERR_ARG, "tflops is null");
    return rfb_run_dmma_peak(ctx, iters, tflops);
}
int rfb_bench_tf32_peak(rfb_ctx *ctx, int iters, double *tflops) {
    RFB_CHECK_CTX(ctx);
    if (!tflops) return ctx->fail(RFB_ERR_ARG, "tflops is null");
    return rfb_run_tf32_peak(ctx, iters, tflops);
}
int rfb_bench_copy(rfb_ctx *ctx, size_t bytes, int iters, double *gbs) {
    RFB_CHECK_CTX(ctx);
    if (!gbs) return ctx->fail(RFB_ERR_ARG, "gbs is null");
    return rfb_run_copy_bench(ctx, bytes, iters, gbs);
}

}  // extern "C"
