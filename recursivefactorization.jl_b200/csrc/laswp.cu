// laswp.cu -- K2: row interchanges (src/lu.jl:164-188, apply_permutation!).
//
// Sequential-swap semantics are kept exactly: for i = 0..npiv-1 swap rows i and ipiv[i]-1-sub of
// the column block; a later swap sees the effect of the earlier ones and ipiv values may repeat.
#include <cstdlib>

#include "rfb_internal.h"

namespace {

constexpr int kLaswpThreads = 128;
constexpr int kPivChunk = 1024;

// One thread per column, pivots staged through shared memory in chunks.  Swaps of one column are
// independent of every other column (the reference's threaded form :164-175 makes the same cut).
template <typename T>
__global__ void __launch_bounds__(kLaswpThreads)
laswp_ipiv_kernel(T *__restrict__ A, long long ncols, long long lda, const long long *__restrict__ ipiv,
                  int npiv, long long sub) {
    __shared__ int s_piv[kPivChunk];
    const long long col = (long long)blockIdx.x * kLaswpThreads + threadIdx.x;
    T *c = A + (col < ncols ? col : 0) * lda;
    for (int c0 = 0; c0 < npiv; c0 += kPivChunk) {
        const int cn = npiv - c0 < kPivChunk ? npiv - c0 : kPivChunk;
        __syncthreads();
        for (int i = threadIdx.x; i < cn; i += kLaswpThreads) s_piv[i] = (int)(ipiv[c0 + i] - 1 - sub);
        __syncthreads();
        if (col < ncols) {
            for (int i = 0; i < cn; ++i) {
                const int r = s_piv[i];
                const int p = c0 + i;
                if (r != p) {                       // serial form skips i' == i (:180)
                    const T t = c[p];
                    c[p] = c[r];
                    c[r] = t;
                }
            }
        }
    }
}

// List-driven form used by the whole-path driver.  K1 already composed the interchanges of each
// panel into a permutation ("row r ends up at row d", at most 2*NB rows per panel), so within one
// panel all reads can be issued before all writes: one warp owns one column, its 32 lanes gather
// up to 128 values in parallel, then scatter them.  Panels are applied in order, which keeps the
// sequential-swap semantics across panels; the pivot rows of a panel are contiguous, so half of
// every gather/scatter is coalesced down the column.  Replaces 2*npiv dependent memory round
// trips per thread of the ipiv-driven kernel by one round trip per panel.
constexpr int kListWarps = 8;
template <typename T>
__global__ void __launch_bounds__(kListWarps * 32)
laswp_list_kernel(T *__restrict__ A, long long ncols, long long lda, const int *__restrict__ perm_dst,
                  const int *__restrict__ perm_src, const int *__restrict__ perm_width, int k0, int k1) {
    const long long colidx = (long long)blockIdx.x * kListWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (colidx >= ncols) return;
    T *col = A + colidx * lda - k0;                 // col[r] addresses absolute row r
    int c = k0;
    int w = perm_width[c];
    while (c < k1 && w > 0) {
        const int nslots = 2 * w;
        int d[4], s[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = lane + 32 * e;
            d[e] = idx < nslots ? perm_dst[2 * c + idx] : -1;
            s[e] = idx < nslots ? perm_src[2 * c + idx] : -1;
        }
        const int cn = c + w;
        const int wn = cn < k1 ? perm_width[cn] : 0;            // prefetch the next panel's width
        T v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (d[e] >= 0 && d[e] != s[e]) ? col[s[e]] : T(0);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (d[e] >= 0 && d[e] != s[e]) col[d[e]] = v[e];
        __syncwarp();
        c = cn;
        w = wn;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Node-level form (round 2).  The list kernel above makes one dependent gather -> scatter round trip per PANEL of the
// node (128 at the root of a 16384^2 factorization) and was latency-bound at 10 % of HBM bandwidth.  Here the panels'
// lists are first composed ON DEVICE into the node's net permutation (one small single-CTA kernel), and the apply
// kernel then moves every column in ONE pass:
//   * rows [k0, k0 + n1) -- the pivot block, every row of which is a destination -- are read and written as one
//     contiguous, fully coalesced segment of the column, staged in shared memory;
//   * the (at most n1) rows below the block that take part are gathered / scattered sector by sector.
// A row below the block only ever receives content that started inside the block (a sequential swap i <-> p[i]
// never touches row i again), so the net permutation has three kinds of entries only: block <- block, block <- below
// (gather) and below <- block (scatter); the kernel issues ALL gathers of a column before any scatter.
// Pivot ranges longer than kNetCap are applied in consecutive chunks (each is a valid sub-sequence of swaps).
constexpr int kNetCap = 8192;            // pivots per chunk: 64 KB (f64) of staged column per CTA
constexpr int kComposeThreads = 1024;
constexpr int kMaxPanelsPerChunk = kNetCap / 8;   // narrowest leaf is 8 columns
constexpr int kNetMaxRows = 45056;       // rows the composition can track in shared memory (176 KB of int32)

// meta[0] cursor (next pivot to compose), meta[1] chunk k0, meta[2] chunk n1, meta[3] nC (scatter entries)
__global__ void __launch_bounds__(kComposeThreads)
laswp_compose_kernel(const int *__restrict__ perm_dst, const int *__restrict__ perm_src, const int *__restrict__ perm_width,
                     int k0, int k1, int row_bound, int cap, int first_round, int stage_lists, int *__restrict__ meta,
                     int *__restrict__ srcmap,
                     int *__restrict__ clist, unsigned int *__restrict__ error_flag) {
    extern __shared__ int cur[];                         // cur[r - start] = original row whose content is now at row r
    __shared__ int pstart[kMaxPanelsPerChunk + 1];
    __shared__ int wsum[kComposeThreads / 32];
    __shared__ int s_np, s_end, s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int start = first_round ? k0 : meta[0];
    if (start >= k1) {                                   // nothing left: make the apply kernel a no-op
        if (tid == 0) { meta[0] = k1; meta[1] = k1; meta[2] = 0; meta[3] = 0; }
        return;
    }
    const int lim = min(k1, start + cap);
    // panel starts inside [start, lim): positions whose width is non-zero (the arrays are cleared once per
    // factorization and every column starts at most one panel); ordered compaction into pstart[]
    if (tid == 0) { s_np = 0; s_end = start; s_total = 0; }
    __syncthreads();
    for (int base = start; base < lim; base += kComposeThreads) {
        const int c = base + tid;
        const int w = c < lim ? perm_width[c] : 0;
        const bool is_start = w > 0 && c + w <= lim;     // a panel that would cross the chunk limit waits for the next chunk
        const unsigned int bal = __ballot_sync(0xffffffffu, is_start);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int off = s_np;
        for (int i = 0; i < warp; ++i) off += wsum[i];
        if (is_start) {
            const int slot = off + __popc(bal & ((1u << lane) - 1));
            if (slot < kMaxPanelsPerChunk) pstart[slot] = c;
            atomicMax(&s_end, c + w);
            atomicAdd(&s_total, w);
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int i = 0; i < kComposeThreads / 32; ++i) t += wsum[i]; s_np += t; }
        __syncthreads();
    }
    const int np = min(s_np, kMaxPanelsPerChunk), end = s_end;
    // the panels must tile [start, end) without a gap (a missing list means the caller asked for pivots this rank does not hold)
    if (np == 0 || s_total != end - start || pstart[0] != start) {
        if (tid == 0) { meta[0] = k1; meta[1] = k1; meta[2] = 0; meta[3] = 0; atomicExch(error_flag, 2u); }
        return;
    }
    const int R = row_bound - start;
    for (int r = tid; r < R; r += kComposeThreads) cur[r] = start + r;
    const int n1 = end - start;
    if (stage_lists) {
        // the chunk's lists fit beside `cur`: stage them with all threads (coalesced), then ONE group of 128 threads walks the
        // panels out of shared memory with a 4-warp named barrier -- ~100 cycles per panel instead of two 32-warp barriers
        int *sdst = cur + R, *ssrc = sdst + 2 * n1;
        for (int i = tid; i < 2 * n1; i += kComposeThreads) {
            sdst[i] = perm_dst[2 * start + i];
            ssrc[i] = perm_src[2 * start + i];
        }
        __syncthreads();
        if (tid < 128) {
            for (int p = 0; p < np; ++p) {
                const int c = pstart[p];
                const int w = (p + 1 < np ? pstart[p + 1] : end) - c;
                int d = -1, sv = -1, t = 0;
                if (tid < 2 * w) { d = sdst[2 * (c - start) + tid]; sv = ssrc[2 * (c - start) + tid]; }
                const bool act = d >= 0 && d != sv;
                if (act) t = cur[sv - start];
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (act) cur[d - start] = t;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
        __syncthreads();
    } else {
    // sequential over panels, parallel within one: eight groups of 128 threads take the panels round-robin and keep
    // their next panel's entries prefetched, so the global-load latency hides behind the seven panels in between
    const int group = tid >> 7, e = tid & 127;
    int pd = -1, ps = -1;
    auto prefetch = [&](int p) {
        pd = ps = -1;
        if (p < np) {
            const int c = pstart[p];
            const int w = (p + 1 < np ? pstart[p + 1] : end) - c;
            if (e < 2 * w) { pd = perm_dst[2 * c + e]; ps = perm_src[2 * c + e]; }
        }
    };
    prefetch(group);
    __syncthreads();
    for (int p = 0; p < np; ++p) {
        const bool mine = (p & 7) == group;
        const bool act = mine && pd >= 0 && pd != ps;
        int t = 0;
        if (act) t = cur[ps - start];
        __syncthreads();
        if (act) cur[pd - start] = t;
        if (mine) prefetch(p + 8);
        __syncthreads();
    }
    }
    for (int i = tid; i < n1; i += kComposeThreads) srcmap[i] = cur[i];
    // below-block rows whose content changed: ordered compaction of (dst, src) pairs
    int total = 0;
    for (int base = n1; base < R; base += kComposeThreads) {
        const int r = base + tid;
        const bool moved = r < R && cur[r] != start + r;
        const unsigned int bal = __ballot_sync(0xffffffffu, moved);
        __syncthreads();
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int off = total, all = 0;
        for (int i = 0; i < kComposeThreads / 32; ++i) { if (i < warp) off += wsum[i]; all += wsum[i]; }
        if (moved) {
            const int slot = off + __popc(bal & ((1u << lane) - 1));
            clist[2 * slot] = start + r;
            clist[2 * slot + 1] = cur[r];
        }
        total += all;
    }
    if (tid == 0) { meta[0] = end; meta[1] = start; meta[2] = n1; meta[3] = total; }
}

// One column per group of GS threads, THREADS / GS columns per CTA, VPT block rows per thread.
template <typename T, int THREADS, int GS, int VPT>
__global__ void __launch_bounds__(THREADS)
laswp_net_kernel(T *__restrict__ A, long long ncols, long long lda, int abs_row0, const int *__restrict__ meta,
                 const int *__restrict__ srcmap, const int *__restrict__ clist) {
    extern __shared__ __align__(16) unsigned char net_smem[];
    constexpr int CPB = THREADS / GS;
    const int k0 = meta[1], n1 = meta[2], nC = meta[3];
    if (n1 <= 0) return;
    const int g = threadIdx.x / GS, lt = threadIdx.x % GS;
    const long long colidx = (long long)blockIdx.x * CPB + g;
    const bool live = colidx < ncols;
    T *Xs = reinterpret_cast<T *>(net_smem) + (size_t)g * (size_t)(GS * VPT);
    T *col = A + (live ? colidx : 0) * lda - abs_row0;            // col[r] addresses absolute row r
    T val[VPT];
    int sm[VPT];
    const int kend = k0 + n1;
#pragma unroll
    for (int t = 0; t < VPT; ++t) {
        const int i = lt + t * GS;
        sm[t] = (live && i < n1) ? srcmap[i] : -1;
    }
#pragma unroll
    for (int t = 0; t < VPT; ++t) {                               // the block itself: coalesced
        const int i = lt + t * GS;
        if (sm[t] >= 0) Xs[i] = col[k0 + i];
    }
#pragma unroll
    for (int t = 0; t < VPT; ++t)                                 // block <- below: gathers, all in flight together
        if (sm[t] >= kend) val[t] = col[sm[t]];
    __syncthreads();
#pragma unroll
    for (int t = 0; t < VPT; ++t) {
        const int i = lt + t * GS;
        if (sm[t] >= 0) {
            if (sm[t] < kend) val[t] = Xs[sm[t] - k0];            // block <- block
            col[k0 + i] = val[t];                                 // coalesced write of the block
        }
    }
    __syncthreads();                                              // every gathered value has been consumed
    if (live)
        for (int j = lt; j < nC; j += GS) col[clist[2 * j]] = Xs[clist[2 * j + 1] - k0];   // below <- block
}

__global__ void ipiv_shift_kernel(long long *ipiv, long long n, long long shift) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ipiv[i] += shift;
}

}  // namespace

template <typename T>
int rfb_launch_laswp(rfb_ctx *ctx, T *A, int64_t ncols, int64_t lda, const int64_t *ipiv_dev,
                     int64_t npiv, int64_t ipiv_sub) {
    if (ncols <= 0 || npiv <= 0) return RFB_OK;
    if (ctx->dry_run) { ctx->rec(RFB_T_LASWP, A, nullptr, nullptr, ncols, ipiv_sub, ipiv_sub + npiv); return RFB_OK; }
    const unsigned int blocks = (unsigned int)((ncols + kLaswpThreads - 1) / kLaswpThreads);
    RfbLaunchScope scope(ctx, RFB_KC_LASWP, 4.0 * sizeof(T) * (double)npiv * (double)ncols);
    laswp_ipiv_kernel<T><<<blocks, kLaswpThreads, 0, ctx->stream>>>(A, ncols, lda, (const long long *)ipiv_dev,
                                                                     (int)npiv, ipiv_sub);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

template <typename T, int THREADS, int GS, int VPT>
static int launch_net(rfb_ctx *ctx, T *A, int64_t ncols, int64_t lda, int abs_row0) {
    constexpr int CPB = THREADS / GS;
    constexpr size_t smem = sizeof(T) * (size_t)CPB * GS * VPT;
    auto kern = laswp_net_kernel<T, THREADS, GS, VPT>;
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)kern, smem));
    kern<<<(unsigned int)((ncols + CPB - 1) / CPB), THREADS, smem, ctx->stream>>>(A, ncols, lda, abs_row0, ctx->net_meta(),
                                                                                 ctx->net_srcmap(), ctx->net_clist());
    RFB_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return RFB_OK;
}

// `row_bound`: one past the last absolute row any list of [k0, k1) can name (the root's row count, or lda).
template <typename T>
int rfb_launch_laswp_lists(rfb_ctx *ctx, T *A, int64_t ncols, int64_t lda, int64_t k0, int64_t k1, int64_t row_bound,
                           int64_t cache_key) {
    if (ncols <= 0 || k1 <= k0) return RFB_OK;
    if (ctx->dry_run) { ctx->rec(RFB_T_LASWP, A, nullptr, nullptr, ncols, k0, k1); return RFB_OK; }
    const int64_t np = k1 - k0;
    const int64_t cap = ctx->laswp_net_cap > 0 && ctx->laswp_net_cap <= kNetCap ? ctx->laswp_net_cap : kNetCap;
    RfbLaunchScope scope(ctx, RFB_KC_LASWP, 4.0 * sizeof(T) * (double)np * (double)ncols);
    if (np >= ctx->laswp_net_min && ctx->net_meta() != nullptr && row_bound > k0 && row_bound - k0 <= kNetMaxRows) {
        // node-level path: compose the panels' lists into the net permutation, then one pass per column
        const int64_t rounds = np <= cap ? 1 : (np + (cap - RFB_MAX_NB) - 1) / (cap - RFB_MAX_NB);
        const int64_t chunk = np < cap ? np : cap;                // upper bound of a chunk's pivot count
        size_t csmem = sizeof(int) * (size_t)(row_bound - k0);
        const int stage_lists = csmem + 4 * sizeof(int) * (size_t)chunk <= sizeof(int) * (size_t)kNetMaxRows ? 1 : 0;
        if (stage_lists) csmem += 4 * sizeof(int) * (size_t)chunk;
        RFB_TRY(rfb_ensure_smem(ctx, (const void *)laswp_compose_kernel, sizeof(int) * (size_t)kNetMaxRows));
        // `cache_key` >= 0 (single-chunk ranges only): the composition of this lane is kept and reused by later calls with the
        // same key -- the multi-GPU driver applies one block column's pivots to many column groups, one composition serves all
        const bool cached = cache_key >= 0 && rounds == 1 && ctx->net_key[ctx->lane] == cache_key;
        if (!(cache_key >= 0 && rounds == 1)) ctx->net_key[ctx->lane] = -1;
        for (int64_t r = 0; r < rounds; ++r) {
            if (!cached) {
                laswp_compose_kernel<<<1, kComposeThreads, csmem, ctx->stream>>>(ctx->perm_dst, ctx->perm_src, ctx->perm_width, (int)k0,
                                                                                (int)k1, (int)row_bound, (int)cap, r == 0 ? 1 : 0, stage_lists,
                                                                                ctx->net_meta(), ctx->net_srcmap(), ctx->net_clist(),
                                                                                &ctx->xchg->error_flag);
                RFB_CUDA(ctx, cudaGetLastError());
                ctx->launches++;
                if (cache_key >= 0 && rounds == 1) ctx->net_key[ctx->lane] = cache_key;
            }
            int rc;
            if (chunk <= 256) rc = launch_net<T, 256, 256, 1>(ctx, A, ncols, lda, (int)k0);
            else if (chunk <= 512) rc = launch_net<T, 256, 256, 2>(ctx, A, ncols, lda, (int)k0);
            else if (chunk <= 1024) rc = launch_net<T, 256, 256, 4>(ctx, A, ncols, lda, (int)k0);
            else if (chunk <= 2048) rc = launch_net<T, 256, 256, 8>(ctx, A, ncols, lda, (int)k0);
            else if (chunk <= 4096) rc = launch_net<T, 512, 512, 8>(ctx, A, ncols, lda, (int)k0);
            else rc = launch_net<T, 512, 512, 16>(ctx, A, ncols, lda, (int)k0);
            RFB_TRY(rc);
        }
        return RFB_OK;
    }
    const unsigned int blocks = (unsigned int)((ncols + kListWarps - 1) / kListWarps);
    laswp_list_kernel<T><<<blocks, kListWarps * 32, 0, ctx->stream>>>(A, ncols, lda, ctx->perm_dst, ctx->perm_src,
                                                                      ctx->perm_width, (int)k0, (int)k1);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

int rfb_launch_ipiv_shift(rfb_ctx *ctx, int64_t *ipiv_dev, int64_t n, int64_t shift) {
    if (n <= 0) return RFB_OK;
    RfbLaunchScope scope(ctx, RFB_KC_OTHER);
    ipiv_shift_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, ctx->stream>>>((long long *)ipiv_dev, n, shift);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

template int rfb_launch_laswp<double>(rfb_ctx *, double *, int64_t, int64_t, const int64_t *, int64_t, int64_t);
template int rfb_launch_laswp<float>(rfb_ctx *, float *, int64_t, int64_t, const int64_t *, int64_t, int64_t);
template int rfb_launch_laswp_lists<double>(rfb_ctx *, double *, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t);
template int rfb_launch_laswp_lists<float>(rfb_ctx *, float *, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t);
