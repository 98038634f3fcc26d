// panel_impl.cuh -- K1: pivoted LU of a tall m x n panel (n <= 64) in ONE launch.
//
// Replaces the reference's leaf `_generic_lufact!` (src/lu.jl:290-338) *and* the bottom levels of
// `reckernel!` (src/lu.jl:189-263) whose nodes are narrower than the leaf width: the result is the
// same factorization (first-strict-max pivot :296-305, reciprocal scaling :317-320, info rule
// :321-327, rank-1 updates :330-334) with LAPACK sequential-swap pivots.
//
// B200 design (not a translation of the CPU loop nest):
//   * rows are distributed over the CTAs of a cooperative grid, ONE ROW PER THREAD, and the whole
//     row (n <= 64 values) lives in registers for the entire panel: the panel is read from HBM
//     once and written once (algorithmic bytes 2*s*m*n), every rank-1 update is register FMAs;
//   * implicit pivoting: rows never move during the panel.  Each thread tracks the LOGICAL
//     position its row would have under LAPACK's sequential swaps (`logpos`); the argmax tie-break
//     uses the logical position, so pivots are bit-identical to the swapping algorithm; rows are
//     scattered to their final position in the single write-back at the end;
//   * one exchange per column and NO grid barrier: each CTA reduces its candidate with warp
//     shuffles, the winning thread publishes {|v|, logpos} and its row to a per-CTA slot in L2
//     (every 16-byte word carries its epoch tag, so readers just poll for the tag; two parities
//     make the slots reusable without a second sync); every CTA then reduces the G headers
//     redundantly and reads the winner's row.
#include <cooperative_groups.h>
#include <cstdlib>

#include "rfb_internal.h"

namespace {

constexpr unsigned int kNone = 0xFFFFFFFFu;
constexpr unsigned int kSpinLimit = 1u << 22;

// Exchange accesses are *strong* relaxed gpu-scope operations on 8-byte self-tagged words
// ({epoch : 32 | payload : 32}); two words travel in one 16-byte vector access but each half is
// validated on its own, so nothing depends on 16-byte single-copy atomicity.  (Weak st.cg / ld.cv
// look faster in a microbenchmark but are not a legal communication pair: with them this kernel
// read stale rows.)  Measured on B200 (scripts/xchg_bench*.cu): one 128-byte line per header and
// a warp-wide row store take the two-phase exchange from 1.1-3.6 us to under 1 us for 2..128 CTAs.
__device__ __forceinline__ void st_tagged(ulonglong2 *p, unsigned int epoch, unsigned int lo, unsigned int hi) {
    const unsigned long long a = ((unsigned long long)epoch << 32) | lo, b = ((unsigned long long)epoch << 32) | hi;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
// returns true when both halves carry `epoch`; lo/hi are the payloads
__device__ __forceinline__ bool ld_tagged(const ulonglong2 *p, unsigned int epoch, unsigned int &lo, unsigned int &hi) {
    unsigned long long a, b;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    lo = (unsigned int)a;
    hi = (unsigned int)b;
    return (unsigned int)(a >> 32) == epoch && (unsigned int)(b >> 32) == epoch;
}
__device__ __forceinline__ unsigned int ld_flag(const unsigned int *p) {
    unsigned int r;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}

__device__ __forceinline__ unsigned long long to_bits(double v) { return (unsigned long long)__double_as_longlong(v); }
__device__ __forceinline__ unsigned long long to_bits(float v) { return (unsigned long long)__float_as_uint(v); }
template <typename T> __device__ __forceinline__ T from_bits(unsigned long long b);
template <> __device__ __forceinline__ double from_bits<double>(unsigned long long b) { return __longlong_as_double((long long)b); }
template <> __device__ __forceinline__ float from_bits<float>(unsigned long long b) { return __uint_as_float((unsigned int)b); }
__device__ __forceinline__ double rcp_rn(double v) { return __drcp_rn(v); }
__device__ __forceinline__ float rcp_rn(float v) { return __frcp_rn(v); }

// Candidate ordering of src/lu.jl:299-304: larger |v| wins; among equal |v| the FIRST (lowest
// logical row) wins.  key == 0 encodes "not greater than the initial amax = 0" (zeros and NaNs).
struct Cand {
    unsigned long long key;
    unsigned int lp;
    unsigned int src;
};
__device__ __forceinline__ bool better(unsigned long long k1, unsigned int lp1, unsigned long long k2, unsigned int lp2) {
    return k1 > k2 || (k1 == k2 && lp1 < lp2);
}
// Warp argmax with three redux.sync instead of a 5-round shuffle butterfly: max of the key's high
// word, max of the low word among the lanes that hold that high word, then the lowest logical row
// among the lanes that hold the full key.
__device__ __forceinline__ Cand warp_best(Cand c) {
    const unsigned int hi = (unsigned int)(c.key >> 32), lo = (unsigned int)c.key;
    const unsigned int mhi = __reduce_max_sync(0xffffffffu, hi);
    const bool in1 = hi == mhi;
    const unsigned int mlo = __reduce_max_sync(0xffffffffu, in1 ? lo : 0u);
    const bool in2 = in1 && lo == mlo;
    const unsigned int mlp = __reduce_min_sync(0xffffffffu, in2 ? c.lp : kNone);
    const unsigned int who = __ballot_sync(0xffffffffu, in2 && c.lp == mlp);
    Cand r;
    r.key = ((unsigned long long)mhi << 32) | mlo;
    r.lp = mlp;
    r.src = __shfl_sync(0xffffffffu, c.src, who ? (__ffs(who) - 1) : 0);
    return r;
}

template <int WARPS>
struct ReduceBuf {
    unsigned long long key[WARPS];
    unsigned int lp[WARPS];
    unsigned int src[WARPS];
};

// One __syncthreads.  Every thread returns the block-wide best candidate.
template <int WARPS>
__device__ __forceinline__ Cand block_best(Cand c, ReduceBuf<WARPS> &buf, int warp, int lane) {
    c = warp_best(c);
    if (lane == 0) { buf.key[warp] = c.key; buf.lp[warp] = c.lp; buf.src[warp] = c.src; }
    __syncthreads();
    Cand b{buf.key[0], buf.lp[0], buf.src[0]};
#pragma unroll
    for (int w = 1; w < WARPS; ++w) {
        unsigned long long k = buf.key[w];
        unsigned int l = buf.lp[w];
        if (better(k, l, b.key, b.lp)) { b.key = k; b.lp = l; b.src = buf.src[w]; }
    }
    return b;
}

template <typename T, int NB, int WARPS>
struct PanelShared {
    alignas(16) T u[2][NB];
    T rinv[2];          // reciprocal of the pivot (1 for an exactly-zero pivot), computed by the winner
    T pub[NB + 1];      // staging of the CTA winner's row window (+ its reciprocal) for the warp-wide publish
    ReduceBuf<WARPS> loc[2];
    ReduceBuf<WARPS> glb[2];
};

// One CTA's candidate of one step in the cluster (DSMEM) exchange.
template <typename T, int NB>
struct ClusterSlot {
    unsigned long long key;
    unsigned int lp;
    unsigned int pad;
    T rinv;
    T row[NB];
};

// Row-exchange list of one panel (consumed by the list-driven laswp, laswp.cu): which rows of the
// panel changed place.  Slot k < n: pivot row k (dst = row0 + k).  Slot n + k: the row that the
// k-th interchange displaced and that still sits there at the end.  Unused slots keep dst = -1
// (the arrays are memset to 0xFF once per factorization).  Rows are absolute (root row 0).
struct PanelPermOut {
    int *dst;      // [2 n] slots of this panel, or nullptr
    int *src;
    int *width;    // width[0] = n
    int row0;      // absolute row of the panel's first row
};

// The column loop is a REAL loop (not unrolled over k): the per-thread row is kept as a sliding
// register window -- reg[j] always holds column k + j, and the rank-1 update writes its result one
// register to the left (reg[j-1] = reg[j] - l * u[j]) -- so every step executes the same few
// hundred instructions out of the instruction cache.  (A version unrolled over k, 51k SASS
// instructions for NB = 64, spent ~3 us per column in instruction-fetch stalls.)  Finished values
// leave the window through a shared-memory tile fin[column][thread] and are written to global
// memory once, at the row's final position.
//
// BATCHED = true turns the same kernel into the batched small-matrix LU (the
// reference's own small-matrix path is this unblocked loop, src/lu.jl:125-126): one CTA per matrix
// (blockIdx.x = matrix, up to THREADS rows and NB columns, fat shapes included -- the loop runs
// min(m, n) pivot steps over all n columns), no inter-CTA exchange.
//
// CLUSTER = true: the whole panel (up to 16 x THREADS rows) is ONE thread-block cluster and the per-column
// exchange goes through DISTRIBUTED SHARED MEMORY instead of L2: each CTA's winner leaves {key, row, 1/pivot}
// in its own shared-memory slot, one barrier.cluster (~380 cycles) makes all slots visible, every CTA reads
// the <= 16 headers remotely (~215 cycles), reduces, and reads the winning row remotely.  That replaces three
// dependent L2 round trips (~700 cycles each on this part) per column; two slot parities make the slots
// reusable with that single cluster barrier per column.
template <typename T, int NB, int THREADS, bool BATCHED = false, bool CLUSTER = false>
__global__ void __launch_bounds__(THREADS, 1)
panel_kernel(T *__restrict__ A, int m, int n, long long lda, long long *__restrict__ ipiv,
             long long ipiv_add, long long *__restrict__ info, long long col_offset,
             RfbPanelXchg *__restrict__ x, unsigned int epoch_base, PanelPermOut perm,
             long long batch_stride_a, long long batch_stride_p) {
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char panel_smem[];
    T *fin = reinterpret_cast<T *>(panel_smem);                 // [NB][THREADS]
    __shared__ PanelShared<T, NB, WARPS> sh;
    __shared__ ClusterSlot<T, NB> cslot[CLUSTER ? 2 : 1];       // this CTA's candidate, read remotely by the cluster

    const int G = BATCHED ? 1 : (int)gridDim.x, bid = BATCHED ? 0 : (int)blockIdx.x, tid = threadIdx.x;
    if (BATCHED) {
        A += (long long)blockIdx.x * batch_stride_a;
        ipiv += (long long)blockIdx.x * batch_stride_p;
        info += blockIdx.x;
    }
    const int npiv = m < n ? m : n;                             // pivot steps (length(ipiv), src/lu.jl:292)
    const int lane = tid & 31, warp = tid >> 5;
    const int row = bid * THREADS + tid;

    bool alive = row < m;
    unsigned int logpos = (unsigned int)row;
    int finalpos = 0, dslot = 0;
    bool bailed = false;

    T reg[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) reg[j] = (alive && j < n) ? A[row + (long long)j * lda] : T(0);
    if (tid < NB) { sh.u[0][tid] = T(0); sh.u[1][tid] = T(0); }
    __syncthreads();

    // Deferred window update: after pivot row k arrives a thread only computes its multiplier and its NEXT
    // column value (cur0, one FMA) -- that is all the next pivot search needs -- and goes straight into the next
    // reduction; the other up-to-62 FMAs of the rank-1 update (same formula, so reg[0] reproduces cur0 bit for bit)
    // run after the next header is on its way, in the shadow of the header round trip instead of in front of it.
    T cur0 = reg[0];
    T pend_nl = T(0);
    bool pending = false;

#pragma unroll 1
    for (int k = 0; k < npiv; ++k) {
        const int par = k & 1;
        const int rem = n - k;                                   // live width of the window
        const unsigned int epoch = epoch_base + (unsigned int)k;

        // -- local candidate (src/lu.jl:296-305) ------------------------------------------------
        const T av = fabs(cur0);
        // Every candidate computes the reciprocal of ITS OWN value now: the ~7 dependent FP64 operations of
        // the correctly rounded reciprocal run in the shadow of the reduction / exchange below instead of
        // on the critical path after it (the winner's value is the one that gets used, :317-320).
        const T myrinv = (cur0 != T(0)) ? rcp_rn(cur0) : T(1);
        Cand c;
        c.key = (alive && av > T(0)) ? to_bits(av) : 0ull;
        c.lp = alive ? logpos : kNone;
        c.src = (unsigned int)bid;
        const Cand cb = block_best<WARPS>(c, sh.loc[par], warp, lane);
        const bool cta_winner = alive && logpos == cb.lp;

        // -- L2 exchange: this CTA's header goes out before anything else -----------------------------
        if (!CLUSTER && G > 1) {
            if (cta_winner) {
                st_tagged(&x->header[par][bid].h[0], epoch, (unsigned int)cb.key, (unsigned int)(cb.key >> 32));
                st_tagged(&x->header[par][bid].h[1], epoch, cb.lp, 0u);
            } else if (cb.lp == kNone && tid == 0) {
                st_tagged(&x->header[par][bid].h[0], epoch, 0u, 0u);
                st_tagged(&x->header[par][bid].h[1], epoch, kNone, 0u);
            }
        }
        // -- deferred rank-1 update of step k-1: slide the window, in chunks of 8 columns; a chunk runs iff it still
        //    held live columns at step k-1 (warp-uniform; the publisher wrote exactly the same chunks, zeros beyond)
        if (pending) {
            const int remp = rem + 1;
#pragma unroll
            for (int c8 = 0; c8 < NB; c8 += 8)
                if (c8 <= remp) {
#pragma unroll
                    for (int j = (c8 == 0 ? 1 : c8); j < c8 + 8; ++j) reg[j - 1] = fma(pend_nl, sh.u[par ^ 1][j], reg[j]);
                }
            reg[NB - 1] = T(0);
            pending = false;
        }

        Cand wb;
        if (CLUSTER) {
            namespace cg = cooperative_groups;
            cg::cluster_group cluster = cg::this_cluster();
            ClusterSlot<T, NB> &mine = cslot[CLUSTER ? par : 0];
            if (cta_winner) {
                mine.key = cb.key;
                mine.lp = cb.lp;
                mine.rinv = myrinv;
#pragma unroll
                for (int c8 = 0; c8 < NB; c8 += 8)
                    if (c8 <= rem) {
#pragma unroll
                        for (int j = c8; j < c8 + 8; ++j) mine.row[j] = reg[j];
                    }
            } else if (cb.lp == kNone && tid == 0) {
                mine.key = 0ull;
                mine.lp = kNone;
            }
            cluster.sync();                                   // every CTA's slot of this parity is complete and visible
            Cand g{0ull, kNone, 0u};
            if (tid < G) {
                const ClusterSlot<T, NB> *rs = cluster.map_shared_rank(&mine, tid);
                g.key = rs->key;
                g.lp = rs->lp;
                g.src = (unsigned int)tid;
            }
            wb = block_best<WARPS>(g, sh.glb[par], warp, lane);
            const ClusterSlot<T, NB> *ws = cluster.map_shared_rank(&mine, wb.src);
            if (tid < NB) sh.u[par][tid] = ws->row[tid];
            else if (tid == NB) sh.rinv[par] = ws->rinv;
            __syncthreads();
        } else if (G == 1) {
            wb = cb;
            if (cta_winner) {
#pragma unroll
                for (int c8 = 0; c8 < NB; c8 += 8)
                    if (c8 <= rem) {
#pragma unroll
                        for (int j = c8; j < c8 + 8; ++j) sh.u[par][j] = reg[j];
                    }
                sh.rinv[par] = myrinv;
            } else if (cb.lp == kNone && tid == 0) {
                sh.rinv[par] = T(1);                  // no candidate at all (zeros / NaNs only): u stays as it is
            }
            __syncthreads();
        } else {
            // -- publish this CTA's candidate (header first, then its row window) ---------------
            if (cta_winner) {
#pragma unroll
                for (int c8 = 0; c8 < NB; c8 += 8)
                    if (c8 < rem) {
#pragma unroll
                        for (int j = c8; j < c8 + 8; ++j) sh.pub[j] = reg[j];
                    }
                sh.pub[NB] = myrinv;
            }
            if (__ballot_sync(0xffffffffu, cta_winner)) {       // the winner's warp stores the row together
                __syncwarp();
#pragma unroll
                for (int j = lane; j < NB; j += 32)
                    if (j < rem) {
                        const unsigned long long bits = to_bits(sh.pub[j]);
                        st_tagged(&x->row[par][bid][j], epoch, (unsigned int)bits, (unsigned int)(bits >> 32));
                    }
                if (lane == 0) {
                    const unsigned long long bits = to_bits(sh.pub[NB]);
                    st_tagged(&x->row[par][bid][RFB_MAX_NB], epoch, (unsigned int)bits, (unsigned int)(bits >> 32));
                }
            }
            // -- gather all G headers, reduce redundantly in every CTA --------------------------
            Cand g{0ull, kNone, 0u};
            for (int cta = tid; cta < G; cta += THREADS) {
                unsigned int klo, khi, hl, spare;
                unsigned int spins = 0;
                while (true) {
                    const bool ok0 = ld_tagged(&x->header[par][cta].h[0], epoch, klo, khi);
                    const bool ok1 = ld_tagged(&x->header[par][cta].h[1], epoch, hl, spare);
                    if (ok0 && ok1) break;
                    if (bailed) break;
                    if ((++spins & 1023u) == 0 && (spins > kSpinLimit || ld_flag(&x->error_flag))) {
                        atomicExch(&x->error_flag, 1u);
                        bailed = true;
                        break;
                    }
                }
                const unsigned long long hk = ((unsigned long long)khi << 32) | klo;
                if (better(hk, hl, g.key, g.lp)) { g.key = hk; g.lp = hl; g.src = (unsigned int)cta; }
            }
            wb = block_best<WARPS>(g, sh.glb[par], warp, lane);
            // -- fetch the winning row window -----------------------------------------------------
            if (tid < NB) {
                T val = T(0);
                if (tid < rem) {
                    unsigned int vlo, vhi;
                    unsigned int spins = 0;
                    while (true) {
                        if (ld_tagged(&x->row[par][wb.src][tid], epoch, vlo, vhi)) break;
                        if (bailed) break;
                        if ((++spins & 1023u) == 0 && (spins > kSpinLimit || ld_flag(&x->error_flag))) {
                            atomicExch(&x->error_flag, 1u);
                            bailed = true;
                            break;
                        }
                    }
                    val = from_bits<T>(((unsigned long long)vhi << 32) | vlo);
                }
                sh.u[par][tid] = val;
            } else if (tid == NB) {                   // one more thread fetches the winner's reciprocal
                T val = T(1);
                if (wb.lp != kNone) {
                    unsigned int vlo, vhi;
                    unsigned int spins = 0;
                    while (true) {
                        if (ld_tagged(&x->row[par][wb.src][RFB_MAX_NB], epoch, vlo, vhi)) break;
                        if (bailed) break;
                        if ((++spins & 1023u) == 0 && (spins > kSpinLimit || ld_flag(&x->error_flag))) {
                            atomicExch(&x->error_flag, 1u);
                            bailed = true;
                            break;
                        }
                    }
                    val = from_bits<T>(((unsigned long long)vhi << 32) | vlo);
                }
                sh.rinv[par] = val;
            }
            __syncthreads();
        }

        // -- eliminate (src/lu.jl:307-334) ------------------------------------------------------
        const T pv = sh.u[par][0];
        if (alive && logpos == wb.lp) {
            alive = false;            // this row is pivot row k: frozen; its window is row k of U
            finalpos = k;
#pragma unroll
            for (int c8 = 0; c8 < NB; c8 += 8)
                if (c8 < rem) {
#pragma unroll
                    for (int j = c8; j < c8 + 8; ++j)
                        if (j < rem) fin[(k + j) * THREADS + tid] = reg[j];
                }
        } else if (alive) {
            if (logpos == (unsigned int)k) { logpos = wb.lp; dslot = k; }   // the swap k <-> kp, on the index
            T l = reg[0];
            if (pv != T(0)) l *= sh.rinv[par];               // reciprocal scaling (:317-320)
            fin[k * THREADS + tid] = l;
            pend_nl = -l;
            if (G == 1 && !CLUSTER) {
                // no exchange latency to hide behind: update the whole window now (one CTA / batched kernel)
#pragma unroll
                for (int c8 = 0; c8 < NB; c8 += 8)
                    if (c8 <= rem) {
#pragma unroll
                        for (int j = (c8 == 0 ? 1 : c8); j < c8 + 8; ++j) reg[j - 1] = fma(pend_nl, sh.u[par][j], reg[j]);
                    }
                reg[NB - 1] = T(0);
                cur0 = reg[0];
            } else {
                cur0 = fma(pend_nl, sh.u[par][1], reg[1]);   // column k+1 after step k (:330-334); the rest is deferred
                pending = true;
            }
        }
        if (bid == 0 && tid == 0) {
            ipiv[k] = (long long)wb.lp + 1 + ipiv_add;
            if (pv == T(0) && *info == 0) *info = col_offset + k + 1;   // (:321-327)
        }
    }

    // -- single write-back, rows land at their final (swapped) position --------------------------
    if (row < m) {
        const long long dst = alive ? (long long)logpos : (long long)finalpos;
        for (int j = 0; j < n; ++j) A[dst + (long long)j * lda] = fin[j * THREADS + tid];
        if (perm.dst != nullptr) {
            if (!alive) {
                perm.dst[finalpos] = perm.row0 + finalpos;
                perm.src[finalpos] = perm.row0 + row;
            } else if (logpos != (unsigned int)row) {
                perm.dst[n + dslot] = perm.row0 + (int)logpos;
                perm.src[n + dslot] = perm.row0 + row;
            }
        }
    }
    if (perm.width != nullptr && bid == 0 && tid == 0) perm.width[0] = n;
    if (CLUSTER) cooperative_groups::this_cluster().sync();     // nobody may exit while its slots can still be read
}

#ifdef RFB_PANEL_BATCHED
// ---- batched small-matrix launcher (panel_batched_f64.cu / _f32.cu) ---------------------------------
template <typename T, int NB, int THREADS>
int launch_batched_inst(rfb_ctx *ctx, T *A, int m, int n, int64_t lda, int64_t stride_a, int64_t batch, int64_t *ipiv,
                        int64_t *info) {
    auto kern = panel_kernel<T, NB, THREADS, true>;
    constexpr size_t smem = sizeof(T) * NB * THREADS;
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)kern, smem));
    const int64_t mn = m < n ? m : n;
    const double flops = mn == n ? (double)m * n * n - (double)n * n * n / 3.0 : (double)n * m * m - (double)m * m * m / 3.0;
    RfbLaunchScope scope(ctx, RFB_KC_PANEL, flops * (double)batch);
    PanelPermOut perm{nullptr, nullptr, nullptr, 0};
    for (int64_t b0 = 0; b0 < batch; b0 += 65535 * 32) {       // (grid.x limit is 2^31-1; chunk anyway)
        const int64_t nb = batch - b0 < 65535 * 32 ? batch - b0 : 65535 * 32;
        kern<<<(unsigned)nb, THREADS, smem, ctx->stream>>>(A + b0 * stride_a, m, n, (long long)lda, (long long *)(ipiv + b0 * mn), 0ll,
                                                          (long long *)(info + b0), 0ll, nullptr, 0u, perm, (long long)stride_a,
                                                          (long long)mn);
        RFB_CUDA(ctx, cudaGetLastError());
    }
    return RFB_OK;
}

template <typename T, int THREADS>
int launch_batched_threads(rfb_ctx *ctx, T *A, int m, int n, int64_t lda, int64_t stride_a, int64_t batch, int64_t *ipiv,
                           int64_t *info) {
    if (n <= 16) return launch_batched_inst<T, 16, THREADS>(ctx, A, m, n, lda, stride_a, batch, ipiv, info);
    if (n <= 32) return launch_batched_inst<T, 32, THREADS>(ctx, A, m, n, lda, stride_a, batch, ipiv, info);
    return launch_batched_inst<T, 64, THREADS>(ctx, A, m, n, lda, stride_a, batch, ipiv, info);
}

}  // namespace

// One CTA per matrix; needs n <= 64 and m <= 128 (the caller falls back to the recursive driver otherwise).
// info[b] must be zeroed by the caller (the kernel only writes a first zero pivot).
template <typename T>
int rfb_launch_panel_batched(rfb_ctx *ctx, T *A, int64_t m, int64_t n, int64_t lda, int64_t stride_a, int64_t batch,
                             int64_t *ipiv_dev, int64_t *info_dev) {
    if (batch <= 0 || m <= 0 || n <= 0) return RFB_OK;
    if (n > RFB_MAX_NB || m > 128) return ctx->fail(RFB_ERR_UNSUPPORTED, "batched kernel handles up to 128 x 64");
    if (m <= 32) return launch_batched_threads<T, 32>(ctx, A, (int)m, (int)n, lda, stride_a, batch, ipiv_dev, info_dev);
    if (m <= 64) return launch_batched_threads<T, 64>(ctx, A, (int)m, (int)n, lda, stride_a, batch, ipiv_dev, info_dev);
    return launch_batched_threads<T, 128>(ctx, A, (int)m, (int)n, lda, stride_a, batch, ipiv_dev, info_dev);
}
template int rfb_launch_panel_batched<RFB_PANEL_T>(rfb_ctx *, RFB_PANEL_T *, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t *, int64_t *);

#else  // !RFB_PANEL_BATCHED
// How many CTAs of this instantiation can be co-resident (what a cooperative launch accepts).
template <typename T, int NB, int THREADS>
int panel_capacity(rfb_ctx *ctx) {
    auto kern = panel_kernel<T, NB, THREADS>;
    auto it = ctx->panel_capacity.find((const void *)kern);
    if (it != ctx->panel_capacity.end()) return it->second;
    constexpr size_t smem = sizeof(T) * NB * THREADS;
    if (rfb_ensure_smem(ctx, (const void *)kern, smem) != RFB_OK) return 0;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem) != cudaSuccess) { cudaGetLastError(); per_sm = 1; }
    int cap = per_sm * ctx->sm_count;
    if (cap > RFB_MAX_PANEL_CTAS) cap = RFB_MAX_PANEL_CTAS;
    ctx->panel_capacity[(const void *)kern] = cap;
    return cap;
}

template <typename T, int NB, int THREADS>
int launch_panel_inst(rfb_ctx *ctx, T *A, int m, int n, int64_t lda, int64_t *ipiv, int64_t ipiv_add,
                      int64_t *info, int64_t col_offset, int G, PanelPermOut perm) {
    auto kern = panel_kernel<T, NB, THREADS>;
    if (G > panel_capacity<T, NB, THREADS>(ctx))
        return ctx->fail(RFB_ERR_UNSUPPORTED, "panel of %d rows x %d columns needs %d co-resident CTAs, the device holds %d", m, n, G,
                         panel_capacity<T, NB, THREADS>(ctx));
    long long lda_ = lda, add_ = ipiv_add, off_ = col_offset;
    long long *ipiv_ = (long long *)ipiv, *info_ = (long long *)info;
    RfbPanelXchg *x = ctx->xchg;
    unsigned int epoch = ctx->panel_epoch;
    long long zero_ = 0;
    void *args[] = {&A, &m, &n, &lda_, &ipiv_, &add_, &info_, &off_, &x, &epoch, &perm, &zero_, &zero_};
    constexpr size_t smem = sizeof(T) * NB * THREADS;
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)kern, smem));
    RfbLaunchScope scope(ctx, RFB_KC_PANEL, (double)m * n * n - (double)n * n * n / 3.0);
    if (G == 1) {
        RFB_CUDA(ctx, cudaLaunchKernel((const void *)kern, dim3(1), dim3(THREADS), args, smem, ctx->stream));
    } else {
        RFB_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)kern, dim3(G), dim3(THREADS), args, smem, ctx->stream));
    }
    return RFB_OK;
}

// Single-cluster launch (DSMEM exchange): G <= 16 CTAs of 256 threads.  *handled = false when this device /
// driver cannot co-schedule such a cluster (the caller then uses the L2 exchange).
template <typename T, int NB>
int launch_panel_cluster_inst(rfb_ctx *ctx, T *A, int m, int n, int64_t lda, int64_t *ipiv, int64_t ipiv_add,
                              int64_t *info, int64_t col_offset, int G, PanelPermOut perm, bool *handled) {
    constexpr int THREADS = 256;
    auto kern = panel_kernel<T, NB, THREADS, false, true>;
    constexpr size_t smem = sizeof(T) * NB * THREADS;
    *handled = false;
    RFB_TRY(rfb_ensure_smem(ctx, (const void *)kern, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const auto key = std::make_pair((const void *)kern, G);
    auto it = ctx->cluster_ok.find(key);
    if (it == ctx->cluster_ok.end()) {
        bool ok = true;
        if (G > 8 && cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ok = false;
        int nclusters = 0;
        if (ok && (cudaOccupancyMaxActiveClusters(&nclusters, (const void *)kern, &cfg) != cudaSuccess || nclusters < 1)) ok = false;
        cudaGetLastError();
        it = ctx->cluster_ok.emplace(key, ok).first;
    }
    if (!it->second) return RFB_OK;
    long long lda_ = lda, add_ = ipiv_add, off_ = col_offset, zero_ = 0;
    long long *ipiv_ = (long long *)ipiv, *info_ = (long long *)info;
    RfbPanelXchg *x = ctx->xchg;
    unsigned int epoch = 0;
    void *args[] = {&A, &m, &n, &lda_, &ipiv_, &add_, &info_, &off_, &x, &epoch, &perm, &zero_, &zero_};
    RfbLaunchScope scope(ctx, RFB_KC_PANEL, (double)m * n * n - (double)n * n * n / 3.0);
    RFB_CUDA(ctx, cudaLaunchKernelExC(&cfg, (const void *)kern, args));
    *handled = true;
    return RFB_OK;
}

template <typename T>
int launch_panel_cluster(rfb_ctx *ctx, T *A, int m, int n, int64_t lda, int64_t *ipiv, int64_t ipiv_add, int64_t *info,
                         int64_t col_offset, int G, PanelPermOut perm, bool *handled) {
    if (n <= 16) return launch_panel_cluster_inst<T, 16>(ctx, A, m, n, lda, ipiv, ipiv_add, info, col_offset, G, perm, handled);
    if (n <= 32) return launch_panel_cluster_inst<T, 32>(ctx, A, m, n, lda, ipiv, ipiv_add, info, col_offset, G, perm, handled);
    return launch_panel_cluster_inst<T, 64>(ctx, A, m, n, lda, ipiv, ipiv_add, info, col_offset, G, perm, handled);
}

template <typename T, int THREADS>
int launch_panel_threads(rfb_ctx *ctx, T *A, int m, int n, int64_t lda, int64_t *ipiv, int64_t ipiv_add,
                         int64_t *info, int64_t col_offset, int G, PanelPermOut perm) {
    if (n <= 16) return launch_panel_inst<T, 16, THREADS>(ctx, A, m, n, lda, ipiv, ipiv_add, info, col_offset, G, perm);
    if (n <= 32) return launch_panel_inst<T, 32, THREADS>(ctx, A, m, n, lda, ipiv, ipiv_add, info, col_offset, G, perm);
    return launch_panel_inst<T, 64, THREADS>(ctx, A, m, n, lda, ipiv, ipiv_add, info, col_offset, G, perm);
}

}  // namespace

template <typename T>
int rfb_launch_panel(rfb_ctx *ctx, T *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv_dev,
                     int64_t ipiv_add, int64_t *info_dev, int64_t col_offset, int64_t perm_row0) {
    if (n <= 0 || m <= 0) return RFB_OK;
    if (ctx->dry_run) { ctx->rec(RFB_T_PANEL, A, nullptr, nullptr, m, n, col_offset); return RFB_OK; }
    PanelPermOut perm{nullptr, nullptr, nullptr, 0};
    if (perm_row0 >= 0 && ctx->perm_dst != nullptr) {     // whole-path driver: also emit the exchange list
        perm.dst = ctx->perm_dst + 2 * perm_row0;
        perm.src = ctx->perm_src + 2 * perm_row0;
        perm.width = ctx->perm_width + perm_row0;
        perm.row0 = (int)perm_row0;
    }
    if (n > RFB_MAX_NB) return ctx->fail(RFB_ERR_UNSUPPORTED, "panel width %lld > %d", (long long)n, RFB_MAX_NB);
    if (m < n) return ctx->fail(RFB_ERR_ARG, "panel needs m >= n (got %lld x %lld)", (long long)m, (long long)n);
    // epoch hygiene: tags are 32 bit; restart the sequence (and wipe stale tags) long before wrap
    if (ctx->panel_epoch > 0xF0000000u) {
        RFB_CUDA(ctx, cudaMemsetAsync(ctx->xchg, 0, sizeof(RfbPanelXchg), ctx->stream));
        ctx->panel_epoch = 1;
    }
    int rc;
    const int64_t g128 = (m + 127) / 128, g256 = (m + 255) / 256;
    static const int force_threads = getenv("RFB_PANEL_THREADS") ? atoi(getenv("RFB_PANEL_THREADS")) : 0;   // tuning aid
    // 128-thread CTAs spread over more SMs; the wide (64-column) kernel fits one 256-thread CTA per SM,
    // which is what lets a 64-column panel reach 148 x 256 rows; narrower panels fit several CTAs per SM
    // A/B switch, default OFF: measured on B200 (run 15, profiles/r01_panel_cluster_m4096.txt) the DSMEM exchange is
    // no faster than the L2 one (2.09 vs 2.08 us per column at 4096 rows, 1.95 vs 1.80 at 512): a column costs
    // ~4000 cycles spread evenly over the in-CTA stages (reduce, publish, fetch, eliminate), not the medium.
    static const int use_cluster = getenv("RFB_PANEL_CLUSTER") ? atoi(getenv("RFB_PANEL_CLUSTER")) : 0;
    if (g256 == 1 && m > 128 && force_threads != 128) {
        // up to 256 rows: one 256-thread CTA, no inter-CTA exchange at all
        rc = launch_panel_threads<T, 256>(ctx, A, (int)m, (int)n, lda, ipiv_dev, ipiv_add, info_dev, col_offset, 1, perm);
        if (rc != RFB_OK) return rc;
        ctx->panel_epoch += (unsigned int)n;
        return rc;
    }
    if (use_cluster && g256 >= 2 && g256 <= 16 && force_threads == 0) {
        // up to 4096 rows: one thread-block cluster, exchange through distributed shared memory
        bool handled = false;
        RFB_TRY(launch_panel_cluster<T>(ctx, A, (int)m, (int)n, lda, ipiv_dev, ipiv_add, info_dev, col_offset, (int)g256, perm, &handled));
        if (handled) return RFB_OK;
    }
    int cap128 = n <= 16 ? panel_capacity<T, 16, 128>(ctx) : n <= 32 ? panel_capacity<T, 32, 128>(ctx) : panel_capacity<T, 64, 128>(ctx);
    if (g128 <= cap128 && force_threads != 256)
        rc = launch_panel_threads<T, 128>(ctx, A, (int)m, (int)n, lda, ipiv_dev, ipiv_add, info_dev, col_offset, (int)g128, perm);
    else
        rc = launch_panel_threads<T, 256>(ctx, A, (int)m, (int)n, lda, ipiv_dev, ipiv_add, info_dev, col_offset, (int)g256, perm);
    if (rc != RFB_OK) return rc;
    ctx->panel_epoch += (unsigned int)n;
    return rc;
}

template <typename T>
int rfb_panel_leaf_for_rows(rfb_ctx *ctx, int64_t m) {
    if (ctx->dry_run) return m <= 148 * 256 ? 64 : (m <= 2 * 148 * 256 ? 32 : 16);   // nominal B200 capacities
    const int64_t g128 = (m + 127) / 128, g256 = (m + 255) / 256;
    if (g128 <= panel_capacity<T, 64, 128>(ctx) || g256 <= panel_capacity<T, 64, 256>(ctx)) return 64;
    if (g128 <= panel_capacity<T, 32, 128>(ctx) || g256 <= panel_capacity<T, 32, 256>(ctx)) return 32;
    if (g128 <= panel_capacity<T, 16, 128>(ctx) || g256 <= panel_capacity<T, 16, 256>(ctx)) return 16;
    return 0;
}
template int rfb_panel_leaf_for_rows<RFB_PANEL_T>(rfb_ctx *, int64_t);

template int rfb_launch_panel<RFB_PANEL_T>(rfb_ctx *, RFB_PANEL_T *, int64_t, int64_t, int64_t, int64_t *, int64_t, int64_t *, int64_t, int64_t);
#endif  // RFB_PANEL_BATCHED
