// trsm.cu -- K3: B <- unitlower(L)^-1 B, left side, in place.
//
// Replaces TriangularSolve.ldiv!(UnitLowerTriangular(A11), A12, thread) at src/lu.jl:235 and :153
// (TriangularSolve.jl is an un-vendored dependency of the reference; its contract is forward
// substitution with an implied unit diagonal that reads only the strict lower triangle).
//
// Blocked recursively on the host: diagonal TB x TB blocks are solved by the kernel below, the
// off-diagonal work (all but a TB/k fraction of the flops) is the K4 GEMM.
#include "rfb_internal.h"

namespace {

// Diagonal block solve.  One thread owns one right-hand-side column and keeps its TB values in
// registers for the whole solve; L is staged once in shared memory and read as warp-wide
// broadcasts.  Memory is touched in exactly two round trips: all TB loads of the L tile are in
// flight together, then all TB loads of the thread's own column (each thread reads and writes whole
// 32-byte sectors of its column, and the 128-byte lines are shared through L1, so going straight
// from global memory to registers costs no extra DRAM traffic).
template <typename T, int TB, int COLS>
__global__ void __launch_bounds__(COLS)
trsm_diag_kernel(const T *__restrict__ L, int kb, T *__restrict__ B, long long nrhs, long long lda) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sL = reinterpret_cast<T *>(smem_raw);          // [TB cols][TB rows], zero outside strict lower
    static_assert(COLS == TB, "the L tile load assumes one thread per tile row");
    const int tid = threadIdx.x;
    const long long col = (long long)blockIdx.x * COLS + tid;
    const bool active = col < nrhs;
    {
        T tl[TB];
#pragma unroll
        for (int c = 0; c < TB; ++c) tl[c] = (tid < kb && c < kb && tid > c) ? L[tid + (long long)c * lda] : T(0);
#pragma unroll
        for (int c = 0; c < TB; ++c) sL[c * TB + tid] = tl[c];
    }
    T *xcol = B + (active ? col : 0) * lda;
    T x[TB];
#pragma unroll
    for (int r = 0; r < TB; ++r) x[r] = (active && r < kb) ? xcol[r] : T(0);
    __syncthreads();
#pragma unroll
    for (int c = 0; c < TB - 1; ++c) {
        const T nxc = -x[c];
#pragma unroll
        for (int r = c + 1; r < TB; ++r) x[r] = fma(sL[c * TB + r], nxc, x[r]);
    }
    if (active) {
#pragma unroll
        for (int r = 0; r < TB; ++r)
            if (r < kb) xcol[r] = x[r];
    }
}

// Block solve for up to 256 x 256 triangles in ONE launch (left-looking over 64 x 64 sub-blocks).
// A thread-per-column solve has only nrhs threads of parallelism (8192 columns = 1.7 warps per SM)
// and crawls on exposed shared-memory latency, so the rows are split as well: a CTA of 8 warps
// handles 32 right-hand-side columns (lane = column), warp g owns rows 8g..8g+7 of the current
// 64-row sub-block in registers.  For sub-block row s: subtract the contributions of the solved
// sub-blocks t < s (x_s -= L_st x_t; x_t stays in shared memory, L_st streams through a shared tile
// whose successor is prefetched into registers during the FMAs), then solve the 64 x 64 diagonal
// triangle in eight 8-row phases (owning warp solves its 8 x 8 triangle and publishes it, the warps
// below update).  Replaces, per 256 rows, 4 diagonal launches + 3 half-empty small GEMM launches.
constexpr int kSub = 64;
constexpr int kBlkCols = 32;
constexpr int kBlkThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kBlkThreads)
trsm_block_kernel(const T *__restrict__ L, int kb, T *__restrict__ B, long long nrhs, long long lda) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sL = reinterpret_cast<T *>(smem_raw);                    // [64 cols][64 rows]
    T *sX = sL + kSub * kSub;                                   // [4 sub-blocks][64 rows][32 cols]
    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
    const long long col = (long long)blockIdx.x * kBlkCols + lane;
    const bool active = col < nrhs;
    T *xcol = B + (active ? col : 0) * lda;
    const int nsub = (kb + kSub - 1) / kSub;
    const int nblocks = nsub * (nsub + 1) / 2;

    // L block idx -> (s, t <= s) in row-major order of s; each thread carries 16 elements of it
    T tl[16];
    auto fetch = [&](int idx) {
        int s = 0, rem = idx;
        while (rem > s) { rem -= s + 1; ++s; }
        const int t = rem;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int e = tid + kBlkThreads * i, r = e & 63, c = e >> 6;
            const int gr = s * kSub + r, gc = t * kSub + c;
            tl[i] = (gr < kb && gc < kb) ? L[gr + (long long)gc * lda] : T(0);
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int e = tid + kBlkThreads * i;
            sL[(e >> 6) * kSub + (e & 63)] = tl[i];
        }
    };

    fetch(0);
    int idx = 0;
    for (int s = 0; s < nsub; ++s) {
        const int r0 = s * kSub + 8 * g;
        T x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = (active && r0 + q < kb) ? xcol[r0 + q] : T(0);
        for (int t = 0; t <= s; ++t, ++idx) {
            __syncthreads();                          // everyone is done with the previous L block
            stash();
            if (idx + 1 < nblocks) fetch(idx + 1);    // next block's loads fly during the FMAs below
            __syncthreads();
            if (t < s) {                              // x_s -= L_st x_t
                const T *xt = sX + t * kSub * kBlkCols + lane;
#pragma unroll 8
                for (int j = 0; j < kSub; ++j) {
                    const T nv = -xt[j * kBlkCols];
                    const T *lj = sL + j * kSub + 8 * g;
#pragma unroll
                    for (int q = 0; q < 8; ++q) x[q] = fma(lj[q], nv, x[q]);
                }
            } else {                                  // 64 x 64 diagonal triangle, eight 8-row phases
                T *xs = sX + s * kSub * kBlkCols + lane;
#pragma unroll 1
                for (int p = 0; p < 8; ++p) {
                    if (g == p) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const T nxc = -x[c];
                            const T *lc = sL + (8 * p + c) * kSub + 8 * p;
#pragma unroll
                            for (int q = c + 1; q < 8; ++q) x[q] = fma(lc[q], nxc, x[q]);
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) xs[(8 * p + q) * kBlkCols] = x[q];
                    }
                    __syncthreads();
                    if (g > p) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const T nv = -xs[(8 * p + c) * kBlkCols];
                            const T *lc = sL + (8 * p + c) * kSub + 8 * g;
#pragma unroll
                            for (int q = 0; q < 8; ++q) x[q] = fma(lc[q], nv, x[q]);
                        }
                    }
                }
                if (active) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (r0 + q < kb) xcol[r0 + q] = x[q];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Float64 block solve, round 2: warp-independent, DMMA-fed (trsm_dmma_kernel).
//
// The kernel above spends a 256-row block in ~50 us whatever the number of right-hand sides: ten 64 x 64 sub-block
// steps, each with block-wide barriers, 512 dependent shared-memory broadcasts per thread and the FP64 pipe 21 % busy
// (profiles/r01_trsm_block_kernel.txt) -- and a 16384^2 LU runs 384 such launches back to back.  Here
//   * every WARP owns 8 right-hand-side columns and walks the whole triangle on its own: no barrier inside a
//     sub-block step, only one per staged 64 x 64 tile of L (ten per 256 rows);
//   * the unknowns live in registers in the DMMA accumulator layout (8 m-tiles x 2 doubles); an off-diagonal sub-block
//     x_s -= L_st x_t is 128 mma.sync.m8n8k4.f64 per warp, with L fragments read conflict-free from a padded shared
//     tile and x_t fragments from the warp's own staging slot (stored negated, so the MMA accumulates the subtraction);
//   * a diagonal 64 x 64 triangle goes in eight 8-row micro-blocks: forward substitution inside the micro-block with
//     warp shuffles (the dependent chain: 7 steps of shuffle + FMA), then the rows below take the micro-block's
//     contribution as two DMMAs per m-tile;
//   * L tiles are prefetched one ahead with cp.async into a double buffer shared by the CTA's warps.
// Substitution order per unknown is still "all earlier unknowns, in order"; only the association inside the 4-term
// DMMA dot products differs from the FMA chain, so results agree with the oracle to rounding, not bit for bit.
constexpr int kLdTile = 72;      // padded column stride (doubles): fragment reads of 4 columns x 8 rows hit all 32 banks twice

__device__ __forceinline__ void trsm_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void trsm_cp_async8(double *dst, const double *src, int src_bytes) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ void trsm_cp_async16(double *dst, const double *src, int src_bytes) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}

// W consumer warps (8 right-hand-side columns each) + ONE producer warp that does nothing but stream the 64 x 64 tiles
// of L into the double buffer (v1 let the consumers copy: more than half of their instructions were address arithmetic
// of the copy loop, profiles/r02_trsm_dmma_kernel_v1.txt).
template <int W>
__global__ void __launch_bounds__((W + 1) * 32)
trsm_dmma_kernel(const double *__restrict__ L, int kb, double *__restrict__ B, long long nrhs, long long lda, int aligned16) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sL = reinterpret_cast<double *>(smem_raw);               // [2][64 cols][kLdTile]
    double *sXall = sL + 2 * 64 * kLdTile;                           // [W][4 sub-blocks][64 rows][8 cols], NEGATED unknowns
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == W;
    const int g = lane >> 2, q = lane & 3;
    double *sx = sXall + (producer ? 0 : warp) * (4 * 64 * 8);
    const long long col0 = (long long)blockIdx.x * (W * 8) + warp * 8;
    const int nsub = (kb + 63) >> 6;
    const int ntiles = nsub * (nsub + 1) / 2;

    auto load_tile = [&](int s, int t, int buf) {                    // producer warp only
        double *dstb = sL + buf * 64 * kLdTile;
        if (aligned16) {                                             // lane = row pair, one 16-byte copy per column
            const int r = 2 * lane, gr = s * 64 + r;
            const double *src = L + gr + (long long)(t * 64) * lda;
            const int rb = gr + 1 < kb ? 16 : (gr < kb ? 8 : 0);
#pragma unroll 8
            for (int c = 0; c < 64; ++c)
                trsm_cp_async16(dstb + c * kLdTile + r, rb && t * 64 + c < kb ? src + (long long)c * lda : L, t * 64 + c < kb ? rb : 0);
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = lane + 32 * h, gr = s * 64 + r;
                const double *src = L + gr + (long long)(t * 64) * lda;
#pragma unroll 8
                for (int c = 0; c < 64; ++c) {
                    const bool ok = gr < kb && t * 64 + c < kb;
                    trsm_cp_async8(dstb + c * kLdTile + r, ok ? src + (long long)c * lda : L, ok ? 8 : 0);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto load_x = [&](int s, double (&x)[8][2]) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = s * 64 + 8 * i + g;
                const long long col = col0 + 2 * q + e;
                x[i][e] = (row < kb && col < nrhs) ? B[row + col * lda] : 0.0;
            }
    };

    if (producer) load_tile(0, 0, 0);
    double xc[8][2], xn[8][2];
    if (!producer) load_x(0, xn);
    int idx = 0;
    for (int s = 0; s < nsub; ++s) {
        if (!producer) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { xc[i][0] = xn[i][0]; xc[i][1] = xn[i][1]; }
        }
        for (int t = 0; t <= s; ++t, ++idx) {
            if (producer) asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                                   // tile idx has landed; every warp is done with tile idx - 1
            if (producer) {
                if (idx + 1 < ntiles) {
                    const int s2 = (t == s) ? s + 1 : s, t2 = (t == s) ? 0 : t + 1;
                    load_tile(s2, t2, (idx + 1) & 1);
                }
                continue;
            }
            const double *tL = sL + (idx & 1) * 64 * kLdTile;
            if (t < s) {                                       // x_s -= L_st x_t
                const double *xt = sx + t * (64 * 8);
#pragma unroll 4
                for (int kk = 0; kk < 16; ++kk) {
                    const double b = xt[(4 * kk + q) * 8 + g];
                    const double *ta = tL + (4 * kk + q) * kLdTile + g;
#pragma unroll
                    for (int i = 0; i < 8; ++i) trsm_dmma(xc[i][0], xc[i][1], ta[8 * i], b);
                }
            } else {                                           // the 64 x 64 unit-lower triangle
                if (s + 1 < nsub) load_x(s + 1, xn);           // next sub-block's right-hand sides fly during this triangle
                double *xs = sx + s * (64 * 8);
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    // rows 8p .. 8p+7: this lane's row is 8p + g; it needs L[8p+g][8p+r] for r < g
                    double lrow[7];
#pragma unroll
                    for (int r = 0; r < 7; ++r) lrow[r] = tL[(8 * p + r) * kLdTile + 8 * p + g];
#pragma unroll
                    for (int r = 0; r < 7; ++r) {
                        const double x0 = __shfl_sync(0xffffffffu, xc[p][0], (r << 2) | q);
                        const double x1 = __shfl_sync(0xffffffffu, xc[p][1], (r << 2) | q);
                        if (g > r) {
                            xc[p][0] = fma(-lrow[r], x0, xc[p][0]);
                            xc[p][1] = fma(-lrow[r], x1, xc[p][1]);
                        }
                    }
                    // publish -x_p in the B-fragment layout, then the rows below take its contribution
                    *reinterpret_cast<double2 *>(xs + (8 * p + g) * 8 + 2 * q) = make_double2(-xc[p][0], -xc[p][1]);
                    __syncwarp();
                    if (p < 7) {
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            const double b = xs[(8 * p + 4 * kk + q) * 8 + g];
                            const double *ta = tL + (8 * p + 4 * kk + q) * kLdTile + g;
#pragma unroll
                            for (int i = p + 1; i < 8; ++i) trsm_dmma(xc[i][0], xc[i][1], ta[8 * i], b);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int row = s * 64 + 8 * i + g;
                        const long long col = col0 + 2 * q + e;
                        if (row < kb && col < nrhs) B[row + col * lda] = xc[i][e];
                    }
            }
        }
    }
}

template <int W>
int launch_dmma(rfb_ctx *ctx, const double *L, int kb, double *B, int64_t nrhs, int64_t lda) {
    constexpr size_t smem = sizeof(double) * (2 * 64 * kLdTile + (size_t)W * 4 * 64 * 8);
    auto kern = trsm_dmma_kernel<W>;
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)kern, smem));
    RfbLaunchScope scope(ctx, RFB_KC_TRSM, (double)kb * (double)kb * (double)nrhs);
    const int aligned16 = ((reinterpret_cast<uintptr_t>(L) & 15) == 0 && (lda & 1) == 0) ? 1 : 0;
    kern<<<(unsigned int)((nrhs + W * 8 - 1) / (W * 8)), (W + 1) * 32, smem, ctx->stream>>>(L, kb, B, nrhs, lda, aligned16);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

template <typename T>
int launch_block_legacy(rfb_ctx *ctx, const T *L, int kb, T *B, int64_t nrhs, int64_t lda);

template <typename T>
int launch_block(rfb_ctx *ctx, const T *L, int kb, T *B, int64_t nrhs, int64_t lda, bool legacy) {
    if constexpr (sizeof(T) == 8) {
        if (!legacy) {
            // 8 warps x 8 columns per CTA once that still gives every SM a CTA; 4 warps otherwise (more CTAs in flight)
            if (nrhs >= (int64_t)ctx->sm_count * 48) return launch_dmma<8>(ctx, L, kb, B, nrhs, lda);
            return launch_dmma<4>(ctx, L, kb, B, nrhs, lda);
        }
    }
    return launch_block_legacy<T>(ctx, L, kb, B, nrhs, lda);
}




template <typename T>
int launch_block_legacy(rfb_ctx *ctx, const T *L, int kb, T *B, int64_t nrhs, int64_t lda) {
    constexpr size_t smem = sizeof(T) * (kSub * kSub + 4 * kSub * kBlkCols);
    auto kern = trsm_block_kernel<T>;
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)kern, smem));
    RfbLaunchScope scope(ctx, RFB_KC_TRSM, (double)kb * (double)kb * (double)nrhs);
    kern<<<(unsigned int)((nrhs + kBlkCols - 1) / kBlkCols), kBlkThreads, smem, ctx->stream>>>(L, kb, B, nrhs, lda);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

// Upper-triangular, non-unit twin of trsm_block_kernel for the back substitution of an LU solve
// (`ldiv!(UpperTriangular(F.factors), B)`, src/lu.jl:62): sub-blocks and 8-row phases run bottom-up,
// the owning warp divides by the diagonal.  Padding rows of a ragged last sub-block get a unit diagonal.
template <typename T>
__global__ void __launch_bounds__(kBlkThreads)
trsm_upper_block_kernel(const T *__restrict__ U, int kb, T *__restrict__ B, long long nrhs, long long lda) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sL = reinterpret_cast<T *>(smem_raw);                    // [64 cols][64 rows] of the current U block
    T *sX = sL + kSub * kSub;                                   // [4 sub-blocks][64 rows][32 cols]
    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
    const long long col = (long long)blockIdx.x * kBlkCols + lane;
    const bool active = col < nrhs;
    T *xcol = B + (active ? col : 0) * lda;
    const int nsub = (kb + kSub - 1) / kSub;

    T tl[16];
    auto fetch = [&](int s, int t) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int e = tid + kBlkThreads * i, r = e & 63, c = e >> 6;
            const int gr = s * kSub + r, gc = t * kSub + c;
            T v = (gr < kb && gc < kb) ? U[gr + (long long)gc * lda] : T(0);
            if (gr == gc && gr >= kb) v = T(1);
            tl[i] = v;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int e = tid + kBlkThreads * i;
            sL[(e >> 6) * kSub + (e & 63)] = tl[i];
        }
    };

    fetch(nsub - 1, nsub - 1);
    for (int s = nsub - 1; s >= 0; --s) {
        const int r0 = s * kSub + 8 * g;
        T x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = (active && r0 + q < kb) ? xcol[r0 + q] : T(0);
        for (int t = nsub - 1; t >= s; --t) {
            __syncthreads();
            stash();
            if (t > s) fetch(s, t - 1);
            else if (s > 0) fetch(s - 1, nsub - 1);
            __syncthreads();
            if (t > s) {                              // x_s -= U_st x_t
                const T *xt = sX + t * kSub * kBlkCols + lane;
#pragma unroll 8
                for (int j = 0; j < kSub; ++j) {
                    const T nv = -xt[j * kBlkCols];
                    const T *lj = sL + j * kSub + 8 * g;
#pragma unroll
                    for (int q = 0; q < 8; ++q) x[q] = fma(lj[q], nv, x[q]);
                }
            } else {                                  // diagonal 64 x 64 upper triangle, bottom-up phases
                T *xs = sX + s * kSub * kBlkCols + lane;
#pragma unroll 1
                for (int p = 7; p >= 0; --p) {
                    if (g == p) {
#pragma unroll
                        for (int c = 7; c >= 0; --c) {
                            const T *lc = sL + (8 * p + c) * kSub + 8 * p;
                            x[c] = x[c] / lc[c];
                            const T nxc = -x[c];
#pragma unroll
                            for (int q = 0; q < c; ++q) x[q] = fma(lc[q], nxc, x[q]);
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) xs[(8 * p + q) * kBlkCols] = x[q];
                    }
                    __syncthreads();
                    if (g < p) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const T nv = -xs[(8 * p + c) * kBlkCols];
                            const T *lc = sL + (8 * p + c) * kSub + 8 * g;
#pragma unroll
                            for (int q = 0; q < 8; ++q) x[q] = fma(lc[q], nv, x[q]);
                        }
                    }
                }
                if (active) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (r0 + q < kb) xcol[r0 + q] = x[q];
                }
            }
        }
    }
}

template <typename T>
int launch_upper_block(rfb_ctx *ctx, const T *U, int kb, T *B, int64_t nrhs, int64_t lda) {
    constexpr size_t smem = sizeof(T) * (kSub * kSub + 4 * kSub * kBlkCols);
    auto kern = trsm_upper_block_kernel<T>;
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)kern, smem));
    RfbLaunchScope scope(ctx, RFB_KC_TRSM, (double)kb * (double)kb * (double)nrhs);
    kern<<<(unsigned int)((nrhs + kBlkCols - 1) / kBlkCols), kBlkThreads, smem, ctx->stream>>>(U, kb, B, nrhs, lda);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

template <typename T>
int trsm_upper_rec(rfb_ctx *ctx, const T *U, int64_t k, T *B, int64_t nrhs, int64_t lda, const rfb_opts *opts) {
    if (k <= 256) return launch_upper_block<T>(ctx, U, (int)k, B, nrhs, lda);
    int64_t k1 = ((k / 2 + 255) / 256) * 256;
    if (k1 >= k) k1 = ((k - 1) / 256) * 256;
    RFB_TRY(trsm_upper_rec<T>(ctx, U + k1 + k1 * lda, k - k1, B + k1, nrhs, lda, opts));        // bottom block first
    RFB_TRY(rfb_launch_gemm<T>(ctx, B, U + k1 * lda, B + k1, k1, nrhs, k - k1, lda, opts));      // B1 -= U12 X2
    return trsm_upper_rec<T>(ctx, U, k1, B, nrhs, lda, opts);
}

template <typename T, int TB>
int launch_diag(rfb_ctx *ctx, const T *L, int kb, T *B, int64_t nrhs, int64_t lda) {
    constexpr int COLS = TB;
    constexpr size_t smem = sizeof(T) * (TB * TB);
    auto kern = trsm_diag_kernel<T, TB, COLS>;
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)kern, smem));
    RfbLaunchScope scope(ctx, RFB_KC_TRSM, (double)kb * (double)kb * (double)nrhs);
    kern<<<(unsigned int)((nrhs + COLS - 1) / COLS), COLS, smem, ctx->stream>>>(L, kb, B, nrhs, lda);
    RFB_CUDA(ctx, cudaGetLastError());
    return RFB_OK;
}

template <typename T>
int trsm_rec(rfb_ctx *ctx, const T *L, int64_t k, T *B, int64_t nrhs, int64_t lda, int tb,
             const rfb_opts *opts) {
    if (k <= tb) {
        if (tb == 32) return launch_diag<T, 32>(ctx, L, (int)k, B, nrhs, lda);
        if (tb == 64) return launch_diag<T, 64>(ctx, L, (int)k, B, nrhs, lda);
        return launch_block<T>(ctx, L, (int)k, B, nrhs, lda, opts && opts->trsm_block == 1);
    }
    // split at a multiple of the diagonal block nearest to k/2
    int64_t k1 = ((k / 2 + tb - 1) / tb) * tb;
    if (k1 >= k) k1 = ((k - 1) / tb) * tb;
    RFB_TRY(trsm_rec<T>(ctx, L, k1, B, nrhs, lda, tb, opts));
    RFB_TRY(rfb_launch_gemm<T>(ctx, B + k1, L + k1, B, k - k1, nrhs, k1, lda, opts));
    return trsm_rec<T>(ctx, L + k1 + k1 * lda, k - k1, B + k1, nrhs, lda, tb, opts);
}

}  // namespace

template <typename T>
int rfb_launch_trsm(rfb_ctx *ctx, const T *L, int64_t k, T *B, int64_t nrhs, int64_t lda,
                    const rfb_opts *opts) {
    if (k <= 0 || nrhs <= 0) return RFB_OK;
    if (ctx->dry_run) { ctx->rec(RFB_T_TRSM_LOWER, L, B, nullptr, k, nrhs, 0); return RFB_OK; }
    int tb = 256;                                    // default: fused 256-row block solve
    if (opts && (opts->trsm_block == 32 || opts->trsm_block == 64 || opts->trsm_block == 128)) tb = opts->trsm_block;
    return trsm_rec<T>(ctx, L, k, B, nrhs, lda, tb, opts);
}

template int rfb_launch_trsm<double>(rfb_ctx *, const double *, int64_t, double *, int64_t, int64_t, const rfb_opts *);
template int rfb_launch_trsm<float>(rfb_ctx *, const float *, int64_t, float *, int64_t, int64_t, const rfb_opts *);

template <typename T>
int rfb_launch_trsm_upper(rfb_ctx *ctx, const T *U, int64_t k, T *B, int64_t nrhs, int64_t lda, const rfb_opts *opts) {
    if (k <= 0 || nrhs <= 0) return RFB_OK;
    return trsm_upper_rec<T>(ctx, U, k, B, nrhs, lda, opts);
}
template int rfb_launch_trsm_upper<double>(rfb_ctx *, const double *, int64_t, double *, int64_t, int64_t, const rfb_opts *);
template int rfb_launch_trsm_upper<float>(rfb_ctx *, const float *, int64_t, float *, int64_t, int64_t, const rfb_opts *);
