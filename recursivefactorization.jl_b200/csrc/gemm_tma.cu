// gemm_tma.cu -- K4, TMA-fed FP64 tensor-core (DMMA) kernel for  C <- C - A * B  (src/lu.jl:265-284).
//
// sm_100a has no tcgen05 FP64 kind, so FP64 tensor math is the warp-level mma.sync.m8n8k4.f64
// (SASS DMMA); what IS Blackwell/Hopper-native here is the data path:
//   * A (m x k) and B (k x n) tiles are fetched by the TMA engine (cp.async.bulk.tensor.2d, SASS
//     UTMALDG) straight from the column-major views into 128-byte-swizzled shared tiles, completion
//     signalled on mbarriers -- no registers, no LSU instructions, zero-fill for ragged edges;
//   * one elected producer lane runs 3 k-tiles ahead of the eight DMMA consumer warps through a
//     4-stage full/empty mbarrier ring;
//   * the swizzle is matched by a permuted fragment map (which matrix row/column a DMMA fragment
//     lane owns is free to choose), making every ld.shared.f64 of the inner loop conflict-free
//     without padding -- padding is impossible with TMA's dense boxes;
//   * tiles are rasterised in groups of 8 row-tiles so that a wave of 148 CTAs shares its A and B
//     panels through the 126 MB L2.
// Requirements: 16-byte aligned views and even lda (TMA global address / stride rules); anything
// else goes to the generic cp.async kernel in gemm.cu.
#include <cuda.h>

#include "rfb_internal.h"

namespace {

constexpr int TBM = 128, TBN = 128, TBK = 16, TSTAGES = 4;
constexpr int TCONSUMER_WARPS = 8;
// 8 warps = 2 per SM sub-partition.  A 9th (producer-only) warp would put 3 warps on one
// sub-partition and cap every thread at 65536/4/(3*32) = 170 registers (ptxas then spills the DMMA
// accumulators), so the TMA producer is lane 0 of warp 0, running STAGES-1 tiles ahead.
constexpr int TTHREADS = TCONSUMER_WARPS * 32;
constexpr int kABoxRows = 16;                                  // one A box = 16 rows x TBK cols = 2 KB
constexpr int kABoxes = TBM / kABoxRows;
constexpr int kStageABytes = TBM * TBK * 8;                    // 16 KB
constexpr int kStageBBytes = TBN * TBK * 8;                    // 16 KB
constexpr int kStageBytes = kStageABytes + kStageBBytes;
constexpr size_t kTmaSmem = (size_t)TSTAGES * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
// Rasterisation group: 8 row-tiles.  With K = 8192 one A row-tile panel is 8.4 MB, so a group's A panels
// (67 MB) stay L2-resident while the CTAs sweep the column tiles; 16 row-tiles (134 MB > 126 MB L2) made
// every wave re-read A from HBM (ncu: 19.1 GB read for 2.1 GB of operands, profiles/r01_gemm_f64_tma_*).
constexpr int kGroupM = 8;

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
__device__ __forceinline__ void dmma_884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// fragment-lane -> matrix-index maps that make the swizzled shared reads conflict-free
__device__ __forceinline__ int a_row_in_box(int g, int t) { return (g & 1) + ((g >> 1) & 1) * 8 + (g >> 2) * 2 + t * 4; }
__device__ __forceinline__ int b_col_in_tile(int f) { return (f & 3) * 2 + (f >> 2); }

// REDUCE_EPI: full 128 x 128 tiles leave through the TMA engine as bulk f64 reduce-adds into C (SASS
// UBLKRED.G.S.ADD.F64.RN, one 1 KB column segment per operation): the SM parks -acc in shared memory and never
// reads C -- the addition C + (-acc) is done once, round-to-nearest, at L2 (bit-identical to C - acc).  Ragged
// edge tiles keep the read-modify-write below (a bulk operation needs 16-byte sizes and cannot be predicated per row).
template <bool REDUCE_EPI>
__global__ void __launch_bounds__(TTHREADS, 1)
gemm_f64_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    double *__restrict__ C, int M, int N, int K, long long lda, int tiles_m, int tiles_n) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned long long *full = reinterpret_cast<unsigned long long *>(base + (size_t)TSTAGES * kStageBytes);
    unsigned long long *empty = full + TSTAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // grouped rasterisation (L2 reuse)
    int pid_m, pid_n;
    {
        const int pid = blockIdx.x;
        const int in_group = kGroupM * tiles_n;
        const int group = pid / in_group;
        const int first_m = group * kGroupM;
        const int gsz = min(tiles_m - first_m, kGroupM);
        pid_m = first_m + (pid % in_group) % gsz;
        pid_n = (pid % in_group) / gsz;
    }
    const int m0 = pid_m * TBM, n0 = pid_n * TBN;
    const int KT = (K + TBK - 1) / TBK;

    if (tid == 0) {
        for (int s = 0; s < TSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TCONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // Pull this CTA's C tile into L2 while the main loop runs: the epilogue's read-modify-write then
    // starts from L2 instead of paying an HBM round trip after the last MMA (matters for K <= 1024).
    {
        const int rr = (tid & 7) * 16;                 // 8 threads cover the 128 rows of a column (16 doubles = 128 B each)
#pragma unroll
        for (int c = tid >> 3; c < TBN; c += TTHREADS / 8) {
            if (m0 + rr < M && n0 + c < N) {
                const double *p = C + (m0 + rr) + (long long)(n0 + c) * lda;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            }
        }
    }
    const bool producer = (warp == 0 && lane == 0);
    auto produce = [&](int nk) {          // fill the ring slot of k-tile nk (one elected lane)
        const int s2 = nk % TSTAGES;
        const int use = nk / TSTAGES;
        if (use > 0) mbar_wait(&empty[s2], (unsigned int)(use - 1) & 1u);   // consumers drained the previous use
        mbar_expect_tx(&full[s2], kStageBytes);
        unsigned char *dA = base + (size_t)s2 * kStageBytes;
        unsigned char *dB = dA + kStageABytes;
#pragma unroll
        for (int b = 0; b < kABoxes; ++b)
            tma_load_2d(dA + b * (kABoxRows * TBK * 8), &mapA, m0 + b * kABoxRows, nk * TBK, &full[s2]);
        tma_load_2d(dB, &mapB, nk * TBK, n0, &full[s2]);
    };
    if (producer) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        for (int nk = 0; nk < TSTAGES - 1 && nk < KT; ++nk) produce(nk);
    }
    __syncwarp();

    // ===== DMMA consumers: 8 warps as 2 (m) x 4 (n), warp tile 64 x 32 =====
    const int g = lane >> 2, q = lane & 3;
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;

    // per-thread shared offsets (in doubles) inside one stage
    // A box b: [k 0..15][16 rows], 16-byte chunk c=(row>>1) stored at c ^ (k & 7)
    int a_off[2][2];        // [t][ks & 1] without the b and ks terms
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int r = a_row_in_box(g, t);
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
            const int k7 = kp * 4 + q;
            a_off[t][kp] = q * 16 + ((((r >> 1) ^ k7) & 7) << 1) + (r & 1);
        }
    }
    // B tile: [n 0..127][16 k], chunk c=(k>>1) stored at c ^ (n & 7);  n & 7 == b_col_in_tile(g)
    const int bn = b_col_in_tile(g);
    int b_off[4];           // [ks] without the j term
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) b_off[ks] = (wn + bn) * 16 + ((((ks * 2 + (q >> 1)) ^ bn) & 7) << 1) + (q & 1);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % TSTAGES;
        const unsigned int ph = (unsigned int)(kt / TSTAGES) & 1u;
        if (producer && kt + TSTAGES - 1 < KT) produce(kt + TSTAGES - 1);
        __syncwarp();
        mbar_wait(&full[s], ph);
        const double *tA = reinterpret_cast<const double *>(base + (size_t)s * kStageBytes) + (wm / kABoxRows) * (kABoxRows * TBK);
        const double *tB = reinterpret_cast<const double *>(base + (size_t)s * kStageBytes + kStageABytes);
#pragma unroll
        for (int ks = 0; ks < TBK / 4; ++ks) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                a[i] = tA[(i >> 1) * (kABoxRows * TBK) + ks * 64 + a_off[i & 1][ks & 1]];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = tB[j * 128 + b_off[ks]];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // epilogue: C = C - acc (src/lu.jl:269-273).  The accumulators are first parked in the (now idle)
    // pipeline buffers as a dense column-major 128 x 128 tile -- undoing the fragment permutation --
    // so that the read-modify-write of C runs as fully coalesced 16-byte accesses with 8 independent
    // loads in flight per thread (a direct per-fragment RMW serialises 64 load->store round trips).
    __syncthreads();                                   // every warp is done reading the last stage
    double *sC = reinterpret_cast<double *>(base);     // [n 0..127][m 0..127]
    if (REDUCE_EPI && m0 + TBM <= M && n0 + TBN <= N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = wm + (i >> 1) * kABoxRows + a_row_in_box(g, i & 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = wn + j * 8 + b_col_in_tile(2 * q + e);
                    sC[c * TBM + r] = -acc[i][j][e];
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk engine
        __syncthreads();
        if (tid < TBN) {                               // one 1 KB column segment per thread
            double *gp = C + m0 + (long long)(n0 + tid) * lda;
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                         ::"l"(gp), "r"(smem_u32(sC + tid * TBM)), "r"(TBM * 8) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // performed before the CTA (and the kernel) ends
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = wm + (i >> 1) * kABoxRows + a_row_in_box(g, i & 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = wn + j * 8 + b_col_in_tile(2 * q + e);
                sC[c * TBM + r] = acc[i][j][e];
            }
        }
    }
    __syncthreads();
    {
        const int rp = (tid & 63) * 2;                 // row pair inside the tile
        const int cb = tid >> 6;                       // 0..3
        const int gr = m0 + rp;
#pragma unroll
        for (int it0 = 0; it0 < TBN / 4; it0 += 8) {
            double2 cv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = cb + 4 * (it0 + u);
                const int gc = n0 + c;
                cv[u] = make_double2(0.0, 0.0);
                if (gc < N) {
                    const double *p = C + gr + (long long)gc * lda;
                    if (gr + 1 < M) cv[u] = *reinterpret_cast<const double2 *>(p);
                    else if (gr < M) cv[u].x = *p;
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = cb + 4 * (it0 + u);
                const int gc = n0 + c;
                if (gc < N) {
                    const double2 a2 = *reinterpret_cast<const double2 *>(sC + c * TBM + rp);
                    double *p = C + gr + (long long)gc * lda;
                    if (gr + 1 < M) *reinterpret_cast<double2 *>(p) = make_double2(cv[u].x - a2.x, cv[u].y - a2.y);
                    else if (gr < M) *p = cv[u].x - a2.x;
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_map(rfb_ctx *ctx, CUtensorMap *map, const double *ptr, uint64_t d0, uint64_t d1, uint64_t stride1_bytes,
              uint32_t box0, uint32_t box1) {
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {stride1_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace

int rfb_launch_gemm_f64_tma(rfb_ctx *ctx, double *C, const double *A, const double *B, int64_t m, int64_t n,
                            int64_t k, int64_t lda, bool *handled) {
    *handled = false;
    if (!ctx->encode_tiled) return RFB_OK;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
        (reinterpret_cast<uintptr_t>(C) & 15) || (lda & 1))
        return RFB_OK;
    if (lda * 8 >= (int64_t(1) << 40)) return RFB_OK;
    CUtensorMap mapA, mapB;
    if (!make_map(ctx, &mapA, A, (uint64_t)m, (uint64_t)k, (uint64_t)lda * 8, kABoxRows, TBK)) return RFB_OK;
    if (!make_map(ctx, &mapB, B, (uint64_t)k, (uint64_t)n, (uint64_t)lda * 8, TBK, TBN)) return RFB_OK;
    const int tiles_m = (int)((m + TBM - 1) / TBM), tiles_n = (int)((n + TBN - 1) / TBN);
    if (ctx->gemm_reduce_epilogue) {
        RFB_TRY(rfb_ensure_smem(ctx, (const void *)gemm_f64_tma_kernel<true>, kTmaSmem));
        RfbLaunchScope scope(ctx, RFB_KC_GEMM, 2.0 * (double)m * (double)n * (double)k);
        gemm_f64_tma_kernel<true><<<(unsigned int)(tiles_m * tiles_n), TTHREADS, kTmaSmem, ctx->stream>>>(
            mapA, mapB, C, (int)m, (int)n, (int)k, lda, tiles_m, tiles_n);
        RFB_CUDA(ctx, cudaGetLastError());
        *handled = true;
        return RFB_OK;
    }
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)gemm_f64_tma_kernel<false>, kTmaSmem));
    RfbLaunchScope scope(ctx, RFB_KC_GEMM, 2.0 * (double)m * (double)n * (double)k);
    gemm_f64_tma_kernel<false><<<(unsigned int)(tiles_m * tiles_n), TTHREADS, kTmaSmem, ctx->stream>>>(
        mapA, mapB, C, (int)m, (int)n, (int)k, lda, tiles_m, tiles_n);
    RFB_CUDA(ctx, cudaGetLastError());
    *handled = true;
    return RFB_OK;
}
