// gemm_tc32.cu -- K4', Float32 trailing update  C <- C - A * B  (src/lu.jl:265-284) on the 5th-gen
// tensor cores: tcgen05.mma kind::tf32 with FP32 accumulators in TMEM, operands fed by TMA.
//
// Plain TF32 (10-bit mantissa) would miss the reference's own bound (20*n*eps(Float32)),
// so every operand is split in-kernel into  x = hi + lo  (hi = x with the 13 low mantissa bits
// cleared, lo = x - hi, both exact) and three MMAs per k-step accumulate
//     hi(A)*hi(B) + lo(A)*hi(B) + hi(A)*lo(B)
// into the same TMEM accumulator ("3xTF32"; the dropped lo*lo term is ~2^-22 relative).
//
// Warp roles (192 threads, one 128 x 128 output tile per CTA, 3-stage ring of K = 32 slices):
//   warp 0      TMA producer: A slice as four [32 k][32 m] boxes (A is column-major => MN-major
//               operand, swizzle 128B_ATOM_32B), B slice as one [128 n][32 k] box (K-major, swizzle 128B)
//   warps 2..5  split workers: rewrite hi in place, write lo to the twin buffers, fence to the async
//               proxy, signal the MMA warp; afterwards the same warps are the epilogue
//   warp 1      TMEM allocation + one elected lane issuing 12 tcgen05.mma per slice, tcgen05.commit
//               releasing the ring slot / signalling the epilogue
//   epilogue    tcgen05.ld (32 lanes x 16 columns per instruction), coalesced read-modify-write of C
//               (thread = row, so a warp touches 32 consecutive rows of one column per access)
#include <cuda.h>

#include "rfb_internal.h"

namespace {

constexpr int CBM = 128, CBN = 128, CBK = 32, CSTAGES = 3;
constexpr int CTHREADS = 192;
constexpr int kTileBytes = CBM * CBK * 4;                       // 16 KB (A slice == B slice)
constexpr int kStage = 4 * kTileBytes;                          // A_hi | B_hi | A_lo | B_lo
constexpr size_t kTc32Smem = (size_t)CSTAGES * kStage + 1024 + 256;
constexpr int kTmemCols = 256;   // two 128-column FP32 accumulators

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// shared-memory matrix descriptor (sm_100 format): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout type << 61 (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B)
__device__ __forceinline__ unsigned long long make_desc(unsigned int saddr, unsigned int lbo, unsigned int sbo,
                                                        unsigned int layout) {
    unsigned long long d = (unsigned long long)((saddr & 0x3FFFFu) >> 4);
    d |= (unsigned long long)(lbo >> 4) << 16;
    d |= (unsigned long long)(sbo >> 4) << 32;
    d |= 1ull << 46;
    d |= (unsigned long long)layout << 61;
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, A MN-major (column-major A), B K-major, N = 128, M = 128
constexpr unsigned int kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((CBN >> 3) << 17) | ((CBM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(unsigned int d_tmem, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned int accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(CTHREADS, 1)
gemm_f32_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   float *__restrict__ C, int M, int N, int K, long long lda, int tiles_m, int tiles_n) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned long long *full = reinterpret_cast<unsigned long long *>(base + (size_t)CSTAGES * kStage);
    unsigned long long *conv = full + CSTAGES;
    unsigned long long *empty = conv + CSTAGES;
    unsigned long long *accf = empty + CSTAGES;
    unsigned int *tmem_slot = reinterpret_cast<unsigned int *>(accf + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int pid_m, pid_n;
    {
        const int pid = blockIdx.x, in_group = 16 * tiles_n, group = pid / in_group, first_m = group * 16;
        const int gsz = min(tiles_m - first_m, 16);
        pid_m = first_m + (pid % in_group) % gsz;
        pid_n = (pid % in_group) / gsz;
    }
    const int m0 = pid_m * CBM, n0 = pid_n * CBN;
    const int KT = (K + CBK - 1) / CBK;

    if (tid == 0) {
        for (int s = 0; s < CSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], 4); mbar_init(&empty[s], 1); }
        mbar_init(accf, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: 128 lanes x 128 fp32 columns for the accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned int tmem = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % CSTAGES;
                const int use = kt / CSTAGES;
                if (use > 0) mbar_wait(&empty[s], (unsigned int)(use - 1) & 1u);
                mbar_expect_tx(&full[s], 2 * kTileBytes);
                unsigned char *st = base + (size_t)s * kStage;
#pragma unroll
                for (int b = 0; b < 4; ++b) tma_load_2d(st + b * 4096, &mapA, m0 + b * 32, kt * CBK, &full[s]);
                tma_load_2d(st + kTileBytes, &mapB, kt * CBK, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % CSTAGES;
                mbar_wait(&conv[s], (unsigned int)(kt / CSTAGES) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned int st = smem_u32(base + (size_t)s * kStage);
                const unsigned int a_hi = st, b_hi = st + kTileBytes, a_lo = st + 2 * kTileBytes, b_lo = st + 3 * kTileBytes;
#pragma unroll
                for (int ks = 0; ks < CBK / 8; ++ks) {
                    // A (MN-major TF32: the only legal layout is SWIZZLE_128B_BASE32B, verified with
                    //   scripts/umma_probe.cu): k-rows dense at 128 B, 32-byte chunks XOR-ed with k % 4;
                    //   LBO = 4096 B between 32-row M chunks, SBO = 512 B between 4-row K groups, k-step = +1024 B
                    // B (K-major, SWIZZLE_128B): k-step = +32 B inside the swizzled row; SBO = 1024 B between 8-row N groups
                    const unsigned long long dah = make_desc(a_hi + ks * 1024, 4096, 512, 1);
                    const unsigned long long dal = make_desc(a_lo + ks * 1024, 4096, 512, 1);
                    const unsigned long long dbh = make_desc(b_hi + ks * 32, 16, 1024, 2);
                    const unsigned long long dbl = make_desc(b_lo + ks * 32, 16, 1024, 2);
                    // The tensor core adds into its FP32 accumulator with truncation, so every accumulation
                    // step of a large running sum costs a biased half-ulp.  The big term and the two
                    // correction terms (2^-11 smaller) therefore go to separate accumulators: the big sum is
                    // touched once per k-step instead of three times and the small sum's truncation is negligible.
                    umma_tf32(tmem, dah, dbh, (kt | ks) != 0);
                    umma_tf32(tmem + CBN, dal, dbh, (kt | ks) != 0);
                    umma_tf32(tmem + CBN, dah, dbl, 1u);
                }
                umma_commit(&empty[s]);                  // slot reusable once these MMAs have read it
            }
            umma_commit(accf);                           // accumulator complete
        }
    } else {
        // ===== split workers (warps 2..5), then epilogue =====
        const int wt = tid - 64;                         // 0..127
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % CSTAGES;
            mbar_wait(&full[s], (unsigned int)(kt / CSTAGES) & 1u);
            float4 *hi = reinterpret_cast<float4 *>(base + (size_t)s * kStage);
            float4 *lo = reinterpret_cast<float4 *>(base + (size_t)s * kStage + 2 * kTileBytes);
#pragma unroll 4
            for (int i = wt; i < 2 * kTileBytes / 16; i += 128) {
                const float4 v = hi[i];
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                hi[i] = h;
                lo[i] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&conv[s]);
        }
        // epilogue: TMEM -> registers -> C
        mbar_wait(accf, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int r = m0 + q * 32 + lane;
        const unsigned int taddr = tmem + ((unsigned int)(q * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < CBN; c0 += 16) {
            unsigned int v[16], w[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr + (unsigned int)c0) : "memory");
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                  "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                : "r"(taddr + (unsigned int)(CBN + c0)) : "memory");
            float cv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int c = n0 + c0 + u;
                cv[u] = (r < M && c < N) ? C[r + (long long)c * lda] : 0.f;
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int c = n0 + c0 + u;
                if (r < M && c < N) C[r + (long long)c * lda] = cv[u] - (__uint_as_float(v[u]) + __uint_as_float(w[u]));
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// Tensor-pipe peak: every CTA issues `iters` back-to-back 128 x 128 x 8 kind::tf32 MMAs on operands that sit in shared
// memory (same descriptors / layouts as the GEMM above), accumulating in TMEM; nothing is loaded, nothing is stored.
__global__ void __launch_bounds__(128, 1) tf32_peak_kernel(int iters, float *sink) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned long long *done = reinterpret_cast<unsigned long long *>(base + 2 * kTileBytes);
    unsigned int *tmem_slot = reinterpret_cast<unsigned int *>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *f = reinterpret_cast<float *>(base);
    for (int i = tid; i < 2 * kTileBytes / 4; i += 128) f[i] = 1.0f + (float)(i & 7) * 0.125f;
    if (tid == 0) {
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned int tmem = *tmem_slot;
    if (warp == 0 && lane == 0) {
        const unsigned int a = smem_u32(base), b = a + kTileBytes;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                umma_tf32(tmem, make_desc(a + ks * 1024, 4096, 512, 1), make_desc(b + ks * 32, 16, 1024, 2), (it | ks) != 0);
        }
        umma_commit(done);
    }
    mbar_wait(done, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (sink && tid == 0 && iters < 0) sink[blockIdx.x] = 0.f;      // (keeps `sink` alive; never taken)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_map_f32(rfb_ctx *ctx, CUtensorMap *map, const float *ptr, uint64_t d0, uint64_t d1, uint64_t stride1_bytes,
                  uint32_t box0, uint32_t box1, CUtensorMapSwizzle swz) {
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {stride1_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

int rfb_launch_gemm_f32_tc(rfb_ctx *ctx, float *C, const float *A, const float *B, int64_t m, int64_t n, int64_t k,
                           int64_t lda, bool *handled) {
    *handled = false;
    if (!ctx->encode_tiled) return RFB_OK;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (lda & 3)) return RFB_OK;
    CUtensorMap mapA, mapB;
    if (!make_map_f32(ctx, &mapA, A, (uint64_t)m, (uint64_t)k, (uint64_t)lda * 4, 32, CBK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return RFB_OK;
    if (!make_map_f32(ctx, &mapB, B, (uint64_t)k, (uint64_t)n, (uint64_t)lda * 4, CBK, CBN, CU_TENSOR_MAP_SWIZZLE_128B))
        return RFB_OK;
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)gemm_f32_tc_kernel, kTc32Smem));
    const int tiles_m = (int)((m + CBM - 1) / CBM), tiles_n = (int)((n + CBN - 1) / CBN);
    RfbLaunchScope scope(ctx, RFB_KC_GEMM, 2.0 * (double)m * (double)n * (double)k);
    gemm_f32_tc_kernel<<<(unsigned int)(tiles_m * tiles_n), CTHREADS, kTc32Smem, ctx->stream>>>(
        mapA, mapB, C, (int)m, (int)n, (int)k, lda, tiles_m, tiles_n);
    RFB_CUDA(ctx, cudaGetLastError());
    *handled = true;
    return RFB_OK;
}

int rfb_run_tf32_peak(rfb_ctx *ctx, int iters, double *tflops) {
    if (iters <= 0) iters = 4000;
    constexpr size_t smem = 2 * kTileBytes + 1024 + 64;
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)tf32_peak_kernel, smem));
    cudaEvent_t e0, e1;
    RFB_CUDA(ctx, cudaEventCreate(&e0));
    RFB_CUDA(ctx, cudaEventCreate(&e1));
    const int blocks = ctx->sm_count;
    tf32_peak_kernel<<<blocks, 128, smem, ctx->stream>>>(iters / 10 + 1, nullptr);      // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        RFB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        tf32_peak_kernel<<<blocks, 128, smem, ctx->stream>>>(iters, nullptr);
        RFB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        RFB_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        RFB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double flops = (double)blocks * (double)iters * 4.0 * 2.0 * CBM * CBN * 8.0;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    ctx->launches += 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    RFB_CUDA(ctx, cudaGetLastError());
    *tflops = best;
    return RFB_OK;
}
