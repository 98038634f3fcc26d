// rfb_mg.cu -- multi-GPU recursive LU behind the C ABI (SURVEY.md section 8b "rfb_lu_f64_mg", section 8e).
//
// The reference has no distributed path.  Here ONE n x n matrix is factored by G GPUs of one node:
//   * block columns of width nb are distributed 1-D block-cyclic (block column J on rank J mod G);
//   * a block column is factored by its owner with the single-GPU path -- reckernel! (src/lu.jl:189-263) on the column
//     range, rfb_lu_range -- i.e. the Toledo recursion runs INSIDE every block column;
//   * the factored panel (rows below its diagonal block included) + its pivots + its row-exchange lists are broadcast from
//     the owner with ncclBroadcast, one collective per block column;
//   * ACROSS block columns the order is right-looking: when panel b has arrived, every rank applies b's step of
//     reckernel! to the block columns it owns to the right of b -- row swaps (:233), A12 <- L11^-1 A12 (:235),
//     A22 -= L21 A12 (:240) -- and b's pivots to its own finished columns on the left (:246).
//
// Round 1 ran the Toledo recursion over block columns as well (node-level TRSM / GEMM with the whole left half as the
// inner dimension).  Measured on 8 GPUs (profiles/r02_bench_dist8gpu_32768_tree_schedule.json) that costs the critical
// path dearly: the update of the first block column of a node's right half has the node's whole left half as inner
// dimension (up to n/2 columns, a chain of n1/256 dependent diagonal solves) and cannot start before the left half's LAST
// panel has arrived, and every per-block-column TRSM repeats that latency chain.  In the right-looking order only panel
// b's own contribution (inner dimension nb) sits between the arrival of panel b and the factorization of block column
// b + 1; the pivots are the same (same exact-arithmetic algorithm), the summation is grouped by block column.
// It also needs no n x n replica of L: received panels live in a small ring.
//
// What runs where:
//   * the schedule is C++, behind the C ABI; no torch / Python in the product path.  Two ways in: one process with G devices
//     (rfb_mg_create_all: ncclCommInitRank per device inside one group, one host thread per device) -- what a Julia caller of
//     lu! uses -- or one process per GPU (rfb_mg_create_rank with a shared ncclUniqueId) -- what torchrun / bench.py uses.
//     libnccl is loaded with dlopen, so the single-GPU library has no NCCL dependency.
//   * communication has its own (high-priority) stream per rank: pack -> ncclBroadcast -> unpack into the ring run there,
//     ordered against the compute stream by events, so a rank's GEMMs overlap its own broadcasts.
//   * a host scheduler per rank keeps the compute stream fed by NEED: when panel j - 1 has ARRIVED (event query) and the rank
//     owns block column j, the contributions block column j still lacks and its factorization go out back to back (the
//     critical path, look-ahead); otherwise a bounded slice (~0.5 ms) of the oldest pending contribution to the owned block
//     columns nearest to the critical path, at most two slices in flight.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdlib>
#include <dlfcn.h>
#include <mutex>
#include <thread>

#include <nccl.h>

#include "rfb_internal.h"

namespace {

// ---- libnccl, loaded on first use ---------------------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t *, int, ncclUniqueId, int, ncclConfig_t *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("RFB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("cannot load libnccl: ") + dlerror(); return; }
        // global scope first: a preloaded interposer (profilers, the harness' own NCCL hooks) must still see our calls
        auto sym = [&](const char *n) { void *p = dlsym(RTLD_DEFAULT, n); return p ? p : dlsym(api.handle, n); };
#define RFB_NCCL_SYM(field, name)                                                  \
        api.field = reinterpret_cast<decltype(api.field)>(sym(name));              \
        if (!api.field && api.error.empty()) api.error = std::string("libnccl lacks ") + name;
        RFB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        RFB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        RFB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        RFB_NCCL_SYM(CommAbort, "ncclCommAbort")
        RFB_NCCL_SYM(Broadcast, "ncclBroadcast")
        RFB_NCCL_SYM(AllReduce, "ncclAllReduce")
        RFB_NCCL_SYM(GroupStart, "ncclGroupStart")
        RFB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        RFB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef RFB_NCCL_SYM
        api.CommInitRankConfig = reinterpret_cast<decltype(api.CommInitRankConfig)>(sym("ncclCommInitRankConfig"));   // optional
    });
    return &api;
}

// ---- the block-column layout (pure host logic; shared by the GPU run and the dry-run trace) ------------------------------------
struct MgPlan {
    int64_t n = 0, nb = 0;
    int nblk = 0, world = 1;
    int64_t col0(int b) const { return (int64_t)b * nb; }
    int64_t width(int b) const { return std::min<int64_t>(n, (int64_t)(b + 1) * nb) - (int64_t)b * nb; }
    // block-cyclic with alternating direction (0 1 .. G-1 | G-1 .. 1 0 | 0 1 ..): block column j receives j contributions, so a
    // plain cyclic map gives the last rank ~15 % more update work than the first at 8 blocks per rank; the snake evens it out,
    // and at every turn the owner of block column b also owns b + 1 (no hand-over on the critical path there)
    // Ownership goes in pairs of block columns (kSuper): the owner of an even block column also owns the next one, whose
    // critical-path hand-over then needs no broadcast at all (the panel is read where it was factored).
    static constexpr int kSuper = 2;
    int owner(int b) const { const int sb = b / kSuper, c = sb / world, p = sb % world; return (c & 1) ? world - 1 - p : p; }
    void build(int64_t n_, int64_t nb_, int world_) {
        n = n_; nb = nb_; world = world_;
        nblk = (int)((n + nb - 1) / nb);
    }
};

enum MgTraceCode { MG_T_UPDATE = 1, MG_T_FACTOR = 2, MG_T_BCAST = 3, MG_T_SWAP_LEFT = 4 };

struct MgRank;

}  // namespace

struct rfb_mg {
    std::vector<MgRank *> ranks;       // local ranks (G in one-process mode, 1 in one-process-per-GPU mode)
    int world = 1;
    bool all_mode = false;
    std::string last_error;
    int64_t n = 0, nb = 0;
    bool f32 = false;
    float last_ms = 0;
};

namespace {

struct MgRank {
    rfb_mg *mg = nullptr;
    int rank = 0, world = 1, device = 0;
    rfb_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    cudaStream_t s_comp = nullptr, s_L = nullptr, s_copy = nullptr;
    MgPlan plan;
    bool f32 = false;
    size_t es = 8;
    std::vector<int> own;
    std::vector<int64_t> lcol;
    int64_t ncl = 0;
    char *A = nullptr, *L = nullptr, *stage = nullptr;
    int64_t *ipiv = nullptr, *info = nullptr;
    int *pdst = nullptr, *psrc = nullptr, *pwidth = nullptr;
    std::vector<cudaEvent_t> ev_blk, ev_fact, ev_up;
    std::vector<char> up_pending;
    cudaEvent_t ev_sent = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_final = nullptr, ev_chunk[2] = {nullptr, nullptr};
    rfb_opts opts = {};
    // Bulk slicing: one slice = panel b's contribution to as many consecutive owned block columns as fit `slice_us` of estimated
    // work (at least one, at most merge_max); wider slices make better GEMM shapes, shorter ones bound how long the critical
    // path can wait behind bulk work already in flight (two slices).  Env: RFB_MG_SLICE_US, RFB_MG_MERGE.
    int merge_max = 16;
    double slice_us = 500.0, alert_us = 350.0;
    bool slice_env = false;
    int status = RFB_OK;
    std::string error;
    int64_t bcast_bytes = 0;
    // scheduler statistics of the last factorization (host side): slices, critical-path enqueues, time with nothing to enqueue
    int64_t st_slices = 0, st_crit = 0, st_crit_late_us = 0;
    double st_idle_us = 0, st_wall_us = 0;
    // dry run
    bool dry = false;
    std::vector<int64_t> *trace = nullptr;

    int fail(int code, const char *fmt, ...) {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        error = buf;
        status = code;
        return code;
    }
    char *Aj(int64_t r, int64_t lc) const { return A + ((size_t)r + (size_t)lc * (size_t)plan.n) * es; }
    // received panels: a ring of kRing slots, each n x nb with leading dimension n (the kernels share one lda with A);
    // panel b sits in slot b % kRing, element (row r, column c of the panel) at Pn(b, r, c)
    static constexpr int kRing = 6;
    char *Pn(int b, int64_t r, int64_t c) const {
        return L + ((size_t)(b % kRing) * (size_t)plan.n * (size_t)plan.nb + (size_t)r + (size_t)c * (size_t)plan.n) * es;
    }
    cudaEvent_t ev_slot[kRing] = {};
    // critical-path timing of the last factorization (timing events around each owned block column's update + factorization,
    // and around its publication): where the per-block-column chain time goes
    std::vector<cudaEvent_t> tev_c0, tev_c1, tev_p0, tev_p1;
    void rec(int code, int64_t a, int64_t b, int64_t c, int64_t d) {
        trace->push_back(code); trace->push_back(a); trace->push_back(b); trace->push_back(c); trace->push_back(d);
    }
};

#define MG_CUDA(r, call)                                                                                       \
    do {                                                                                                       \
        cudaError_t e__ = (call);                                                                              \
        if (e__ != cudaSuccess)                                                                                \
            return (r)->fail(RFB_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)
#define MG_NCCL(r, call)                                                                                       \
    do {                                                                                                       \
        ncclResult_t e__ = (call);                                                                             \
        if (e__ != ncclSuccess)                                                                                \
            return (r)->fail(RFB_ERR_NCCL, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, nccl_api()->GetErrorString(e__)); \
    } while (0)
#define MG_TRY(r, expr)                                                                       \
    do {                                                                                      \
        int rc__ = (expr);                                                                    \
        if (rc__ != RFB_OK) {                                                                 \
            if ((r)->status == RFB_OK) (r)->fail(rc__, "%s", (r)->ctx ? (r)->ctx->last_error.c_str() : "error"); \
            return rc__;                                                                      \
        }                                                                                     \
    } while (0)

// ---- the per-rank scheduler -----------------------------------------------------------------------------------------------
template <typename T>
struct MgSched {
    MgRank *r;
    const MgPlan &P;
    int comm_cursor = 0;                 // next block whose publish goes onto the communication stream
    std::vector<char> factored;          // own blocks: factorization enqueued
    std::vector<int> next_src;           // own blocks: next source block whose contribution is still to be applied (== j: ready to factor)
    std::vector<int> users_left;         // per source block: own blocks to its right that have not taken its contribution yet
    std::vector<char> touched;
    int last_arrived = -1;               // highest block column whose panel is known to have arrived
    int next_left = 0;                   // next source block whose pivots go to the rank's finished columns on the left (:246)
    int chunks = 0;

    explicit MgSched(MgRank *rank) : r(rank), P(rank->plan) {
        factored.assign(P.nblk, 0);
        next_src.assign(P.nblk, 0);
        users_left.assign(P.nblk, 0);
        touched.assign(P.nblk, 0);
        for (int b = 0; b < P.nblk; ++b)
            for (int j : r->own)
                if (j > b) users_left[b]++;
    }

    bool recorded(int blk) const { return comm_cursor > blk; }
    bool arrived(int blk) const {
        if (!recorded(blk)) return false;
        if (r->dry) return true;
        return cudaEventQuery(r->ev_blk[blk]) == cudaSuccess;
    }
    bool slot_is_free(int b) const { return b < 0 || users_left[b] == 0; }   // every user of panel b is at least enqueued

    // ---- communication stream -----------------------------------------------------------------------------------------
    int publish(int b) {
        const int root = P.owner(b);
        const int64_t c0 = P.col0(b), w = P.width(b), rows = P.n - c0;
        if (r->dry) { r->rec(MG_T_BCAST, b, root, c0, w); return RFB_OK; }
        const size_t es = r->es;
        const size_t pbytes = (size_t)rows * (size_t)w * es;
        const size_t off_piv = pbytes, off_dst = pbytes + 8 * (size_t)w, off_src = pbytes + 16 * (size_t)w, off_w = pbytes + 24 * (size_t)w;
        const size_t total = pbytes + 28 * (size_t)w;
        cudaStream_t s = r->s_L;
        // the ring slot of panel b held panel b - kRing: its last reader on the compute stream must be done
        if (b >= MgRank::kRing && !r->own.empty() && r->own.back() > b - MgRank::kRing)
            MG_CUDA(r, cudaStreamWaitEvent(s, r->ev_slot[b % MgRank::kRing], 0));
        if (r->rank == root) {
            MG_CUDA(r, cudaStreamWaitEvent(s, r->ev_fact[b], 0));
            MG_CUDA(r, cudaEventRecord(r->tev_p0[b], s));
            if (P.world > 1) {
                MG_CUDA(r, cudaMemcpy2DAsync(r->stage, rows * es, r->Aj(c0, r->lcol[b]), P.n * es, rows * es, w, cudaMemcpyDeviceToDevice, s));
                MG_CUDA(r, cudaMemcpyAsync(r->stage + off_piv, r->ipiv + c0, 8 * w, cudaMemcpyDeviceToDevice, s));
                MG_CUDA(r, cudaMemcpyAsync(r->stage + off_dst, r->pdst + 2 * c0, 8 * w, cudaMemcpyDeviceToDevice, s));
                MG_CUDA(r, cudaMemcpyAsync(r->stage + off_src, r->psrc + 2 * c0, 8 * w, cudaMemcpyDeviceToDevice, s));
                MG_CUDA(r, cudaMemcpyAsync(r->stage + off_w, r->pwidth + c0, 4 * w, cudaMemcpyDeviceToDevice, s));
            }
        }
        if (P.world > 1) {
            MG_NCCL(r, nccl_api()->Broadcast(r->stage, r->stage, total, ncclUint8, root, r->comm, s));
            r->bcast_bytes += (int64_t)total;
            // the owner's next BULK work must not take the SMs before its send has gone out -- unless the owner also owns the next
            // block column: then the critical path continues right here and must not wait for this broadcast
            if (r->rank == root && !(b + 1 < P.nblk && P.owner(b + 1) == r->rank)) {
                MG_CUDA(r, cudaEventRecord(r->ev_sent, s));
                MG_CUDA(r, cudaStreamWaitEvent(r->s_comp, r->ev_sent, 0));
            }
            MG_CUDA(r, cudaMemcpy2DAsync(r->Pn(b, c0, 0), P.n * es, r->stage, rows * es, rows * es, w, cudaMemcpyDeviceToDevice, s));
            if (r->rank != root) {
                MG_CUDA(r, cudaMemcpyAsync(r->ipiv + c0, r->stage + off_piv, 8 * w, cudaMemcpyDeviceToDevice, s));
                MG_CUDA(r, cudaMemcpyAsync(r->pdst + 2 * c0, r->stage + off_dst, 8 * w, cudaMemcpyDeviceToDevice, s));
                MG_CUDA(r, cudaMemcpyAsync(r->psrc + 2 * c0, r->stage + off_src, 8 * w, cudaMemcpyDeviceToDevice, s));
                MG_CUDA(r, cudaMemcpyAsync(r->pwidth + c0, r->stage + off_w, 4 * w, cudaMemcpyDeviceToDevice, s));
            }
        } else {
            MG_CUDA(r, cudaMemcpy2DAsync(r->Pn(b, c0, 0), P.n * es, r->Aj(c0, r->lcol[b]), P.n * es, rows * es, w, cudaMemcpyDeviceToDevice, s));
        }
        return RFB_OK;
    }

    int advance_comm() {
        while (comm_cursor < P.nblk) {
            const int b = comm_cursor;
            if (P.owner(b) == r->rank && !factored[b]) break;
            if (!r->dry && !slot_is_free(b - MgRank::kRing)) break;         // the ring is full: the compute side has to catch up first
            MG_TRY(r, publish(b));
            if (!r->dry) MG_CUDA(r, cudaEventRecord(r->ev_blk[b], r->s_L));
            if (!r->dry && P.owner(b) == r->rank) MG_CUDA(r, cudaEventRecord(r->tev_p1[b], r->s_L));
            comm_cursor++;
        }
        return RFB_OK;
    }

    // ---- compute stream -----------------------------------------------------------------------------------------------
    // panel b's step of reckernel! (src/lu.jl:233-240) on `cnt` consecutive owned block columns starting at block j
    // `local`: read panel b where this rank factored it (its own storage) instead of from the ring, and do not wait for its
    // publication -- only for the block column right after it, before any later pivots have touched those rows
    int contribute(int b, size_t q, int cnt, bool local = false) {             // q: position of the first target in r->own
        const int64_t c0 = P.col0(b), w = P.width(b);
        const int j = r->own[q];
        if (r->dry) {
            for (int g = 0; g < cnt; ++g) { const int jj = r->own[q + g]; r->rec(MG_T_UPDATE, c0, w, jj, 1); next_src[jj] = b + 1; users_left[b]--; }
            return RFB_OK;
        }
        int64_t ncols = 0;
        for (int g = 0; g < cnt; ++g) {
            const int jj = r->own[q + g];
            ncols += P.width(jj);
            if (!touched[jj]) {
                touched[jj] = 1;
                if (r->up_pending[jj]) { MG_CUDA(r, cudaStreamWaitEvent(r->s_comp, r->ev_up[jj], 0)); r->up_pending[jj] = 0; }
            }
        }
        if (!local) MG_CUDA(r, cudaStreamWaitEvent(r->s_comp, r->ev_blk[b], 0));
        rfb_ctx *ctx = r->ctx;
        T *Ablk = reinterpret_cast<T *>(r->Aj(c0, r->lcol[j]));                 // rows c0.. of the target columns
        const T *Lb = reinterpret_cast<const T *>(local ? r->Aj(c0, r->lcol[b]) : r->Pn(b, c0, 0));   // panel b, rows c0..
        ctx->lane = 2 + b % MgRank::kRing;                  // block column b's composed interchanges: built once, used by every slice
        const int rc_swap = rfb_launch_laswp_lists<T>(ctx, Ablk, ncols, P.n, c0, c0 + w, P.n, b);                       // :233
        ctx->lane = 0;
        MG_TRY(r, rc_swap);
        MG_TRY(r, rfb_launch_trsm<T>(ctx, Lb, w, Ablk, ncols, P.n, &r->opts));                                         // :235
        MG_TRY(r, rfb_launch_gemm<T>(ctx, Ablk + w, Lb + w, Ablk, P.n - c0 - w, ncols, w, P.n, &r->opts));              // :240
        for (int g = 0; g < cnt; ++g) { next_src[r->own[q + g]] = b + 1; users_left[b]--; }
        if (users_left[b] == 0) MG_CUDA(r, cudaEventRecord(r->ev_slot[b % MgRank::kRing], r->s_comp));   // the slot may be overwritten
        return RFB_OK;
    }

    // src/lu.jl:246 for source block b: its pivots applied to the rank's finished columns on its left (rows below b's diagonal)
    int swap_left(int b) {
        const int64_t c0 = P.col0(b), w = P.width(b);
        if (r->dry) { r->rec(MG_T_SWAP_LEFT, 0, c0, c0, c0 + w); return RFB_OK; }
        int64_t ncols = 0;
        for (int j : r->own)
            if (j < b) ncols += P.width(j);
        if (ncols == 0) return RFB_OK;
        MG_CUDA(r, cudaStreamWaitEvent(r->s_comp, r->ev_blk[b], 0));
        r->ctx->lane = 2 + b % MgRank::kRing;
        const int rc_swap = rfb_launch_laswp_lists<T>(r->ctx, reinterpret_cast<T *>(r->Aj(c0, 0)), ncols, P.n, c0, c0 + w, P.n, b);
        r->ctx->lane = 0;
        MG_TRY(r, rc_swap);
        return RFB_OK;
    }

    int factor(int j) {
        const int64_t c0 = P.col0(j), w = P.width(j);
        factored[j] = 1;
        if (r->dry) { r->rec(MG_T_FACTOR, j, c0, w, 0); return RFB_OK; }
        if (!touched[j]) {
            touched[j] = 1;
            if (r->up_pending[j]) { MG_CUDA(r, cudaStreamWaitEvent(r->s_comp, r->ev_up[j], 0)); r->up_pending[j] = 0; }
        }
        // the block lives at local column lcol[j]: shift the base so that (row c0, column c0) of the "root" view is it
        char *root = r->A + ((size_t)(r->lcol[j] - c0) * (size_t)P.n) * r->es;   // (pointer arithmetic only; never dereferenced outside the block)
        int rc;
        if (sizeof(T) == 8) rc = rfb_lu_range_f64(r->ctx, reinterpret_cast<double *>(root), P.n, P.n, c0, w, r->ipiv, r->info, &r->opts);
        else rc = rfb_lu_range_f32(r->ctx, reinterpret_cast<float *>(root), P.n, P.n, c0, w, r->ipiv, r->info, &r->opts);
        MG_TRY(r, rc);
        MG_CUDA(r, cudaEventRecord(r->ev_fact[j], r->s_comp));
        MG_CUDA(r, cudaEventRecord(r->tev_c1[j], r->s_comp));
        return RFB_OK;
    }

    // estimated device time (us) of panel b's contribution to ncols columns
    double est_us(int b, int64_t ncols) const {
        const double w = (double)P.width(b), below = (double)(P.n - P.col0(b)) - w;
        const double rate = sizeof(T) == 8 ? 26e6 : 60e6;                        // flops per microsecond
        return 60.0 + 2.0 * below * (double)ncols * w / rate + w * w * (double)ncols / rate;
    }

    int chunks_in_flight() {
        int c = 0;
        for (int i = 0; i < 2; ++i)
            if (chunks > i && cudaEventQuery(r->ev_chunk[(chunks - 1 - i) & 1]) != cudaSuccess) c++;
        return c;
    }

    int run() {
        using clock = std::chrono::steady_clock;
        auto last_progress = clock::now();
        const auto t_begin = last_progress;
        r->st_slices = r->st_crit = 0;
        r->st_idle_us = 0;
        size_t own_pos = 0;                                          // index into r->own of the next block to factor
        auto idle_since = clock::now();
        bool idling = false;
        auto progress = [&] {
            if (idling) { r->st_idle_us += std::chrono::duration<double, std::micro>(clock::now() - idle_since).count(); idling = false; }
            last_progress = clock::now();
        };
        while (true) {
            MG_TRY(r, advance_comm());
            while (own_pos < r->own.size() && factored[r->own[own_pos]]) own_pos++;
            const bool own_done = own_pos >= r->own.size();
            if (own_done && comm_cursor >= P.nblk && next_left >= P.nblk) break;
            // 1. critical path: the next owned block column only waits for contributions whose panels have all arrived
            if (!own_done) {
                const int jn = r->own[own_pos];
                // ready when panel jn - 1 has arrived -- or was factored right here (then nothing older can be missing either)
                // (every older panel must at least be ON ITS WAY into the ring -- recorded -- for the stream waits to mean anything:
                //  with a full ring the communication cursor can lag behind this rank's own factorizations)
                const bool pred_local = jn > 0 && P.owner(jn - 1) == r->rank && factored[jn - 1] && (jn < 2 || recorded(jn - 2));
                if (jn == 0 || pred_local || arrived(jn - 1)) {
                    progress();
                    if (!r->dry) MG_CUDA(r, cudaEventRecord(r->tev_c0[jn], r->s_comp));
                    while (next_src[jn] < jn) {
                        const int b = next_src[jn];
                        MG_TRY(r, contribute(b, own_pos, 1, pred_local && b == jn - 1 && !arrived(b)));
                    }
                    MG_TRY(r, factor(jn));
                    r->st_crit++;
                    continue;
                }
            }
            // how far the critical path is from this rank: d panels still have to arrive before its next block column is ready.
            // d >= 2: at least one whole block column is factored elsewhere first -> wide slices, two in flight;
            // d == 1: the next arrival makes this rank critical -> short slices, one in flight (it must react at once)
            while (last_arrived + 1 < P.nblk && arrived(last_arrived + 1)) last_arrived++;
            const int dist = own_done ? 1 << 20 : r->own[own_pos] - 1 - last_arrived;
            const bool alert = dist <= 1 && !own_done;
            const double slice_target = alert ? std::min(r->slice_us, r->alert_us) : r->slice_us;
            const bool room = r->dry || chunks_in_flight() < (alert ? 1 : 2);
            // 2. bulk: the oldest pending contribution to the owned block columns nearest to the critical path
            bool did = false;
            if (room) {
                for (size_t q = own_pos; q < r->own.size() && !did; ++q) {
                    const int j = r->own[q], b = next_src[j];
                    if (b >= j || !arrived(b)) continue;
                    int cnt = 1;
                    int64_t ncols = P.width(j);
                    for (size_t q2 = q + 1; q2 < r->own.size() && cnt < r->merge_max; ++q2) {
                        const int j2 = r->own[q2];
                        if (next_src[j2] != b || est_us(b, ncols + P.width(j2)) > slice_target) break;
                        ncols += P.width(j2);
                        cnt++;
                    }
                    progress();
                    MG_TRY(r, contribute(b, q, cnt));
                    did = true;
                }
                // 3. the pivots of arrived panels, in order, to the finished columns on their left
                if (!did && next_left < P.nblk && arrived(next_left)) {
                    progress();
                    MG_TRY(r, swap_left(next_left));
                    next_left++;
                    did = true;
                }
                if (did) {
                    if (!r->dry) { MG_CUDA(r, cudaEventRecord(r->ev_chunk[chunks & 1], r->s_comp)); chunks++; }
                    r->st_slices++;
                    continue;
                }
            }
            if (r->dry) return r->fail(RFB_ERR_INTERNAL, "multi-GPU schedule cannot make progress (rank %d, cursor %d)", r->rank, comm_cursor);
            if (room && !idling && chunks_in_flight() == 0) { idling = true; idle_since = clock::now(); }   // nothing runnable at all
            if (std::chrono::duration<double>(clock::now() - last_progress).count() > 60.0)
                return r->fail(RFB_ERR_INTERNAL, "multi-GPU schedule stalled for 60 s (rank %d, communication cursor %d)", r->rank, comm_cursor);
            std::this_thread::yield();
        }
        r->st_wall_us = std::chrono::duration<double, std::micro>(clock::now() - t_begin).count();
        return RFB_OK;
    }
};

// ---- rank set-up / tear-down ---------------------------------------------------------------------------------------------------
int rank_free_problem(MgRank *r) {
    if (!r->ctx) return RFB_OK;
    cudaSetDevice(r->device);
    cudaDeviceSynchronize();
    for (void *p : {(void *)r->A, (void *)r->L, (void *)r->stage, (void *)r->ipiv, (void *)r->info, (void *)r->pdst, (void *)r->psrc, (void *)r->pwidth})
        if (p) cudaFree(p);
    r->A = r->L = r->stage = nullptr;
    r->ipiv = r->info = nullptr;
    r->pdst = r->psrc = r->pwidth = nullptr;
    for (auto *v : {&r->ev_blk, &r->ev_fact, &r->ev_up, &r->tev_c0, &r->tev_c1, &r->tev_p0, &r->tev_p1}) {
        for (cudaEvent_t e : *v) if (e) cudaEventDestroy(e);
        v->clear();
    }
    r->ctx->perm_dst = r->ctx->perm_src = r->ctx->perm_width = nullptr;
    r->ctx->perm_cap = 0;
    r->ctx->perm_external = false;
    return RFB_OK;
}

int rank_setup(MgRank *r, int64_t n, int64_t nb, bool f32) {
    MG_CUDA(r, cudaSetDevice(r->device));
    rank_free_problem(r);
    r->plan.build(n, nb, r->world);
    r->f32 = f32;
    r->es = f32 ? 4 : 8;
    {
        // Balance of this problem on this many ranks: bulk = one rank's share of the GEMM-shaped work at a typical rate; critical
        // path = every block column's pivot chain (one all-CTA exchange per column) + its in-block updates + the hand-over.  The
        // more bulk-bound, the wider the slices (better GEMM shapes, fixed per-slice costs amortised); the more
        // critical-path-bound, the shorter (the critical path waits for at most two slices in flight).
        const double t_bulk = (2.0 / 3.0) * (double)n * (double)n * (double)n / r->world / (f32 ? 60e12 : 28e12);
        const double t_crit = (double)r->plan.nblk * ((double)nb * 2.4e-6 + 1.2e-3);
        const double ratio = t_bulk / t_crit;
        if (!r->slice_env) r->slice_us = ratio > 2.0 ? 3000.0 : 1300.0;
    }
    r->opts = rfb_opts{};
    r->opts.mem_space = RFB_MEM_DEVICE;
    if (f32) r->opts.f32_mode = n >= 4096 ? RFB_F32_TF32X3 : RFB_F32_FP32;      // RFB_F32_AUTO, resolved like rfb_lu_f32 does
    const MgPlan &P = r->plan;
    r->own.clear();
    r->lcol.assign(P.nblk, -1);
    r->ncl = 0;
    for (int j = 0; j < P.nblk; ++j)
        if (P.owner(j) == r->rank) { r->own.push_back(j); r->lcol[j] = r->ncl; r->ncl += P.width(j); }
    const size_t nn = (size_t)n;
    auto alloc = [&](void **p, size_t bytes) { return cudaMalloc(p, bytes ? bytes : 16) == cudaSuccess; };
    bool ok = alloc((void **)&r->L, (size_t)MgRank::kRing * nn * (size_t)nb * r->es) &&
              alloc((void **)&r->A, nn * (size_t)std::max<int64_t>(r->ncl, 1) * r->es) &&
              alloc((void **)&r->stage, nn * (size_t)nb * r->es + 28 * (size_t)nb + 256) && alloc((void **)&r->ipiv, 8 * (nn + 64)) &&
              alloc((void **)&r->info, 64) && alloc((void **)&r->pdst, 4 * (2 * nn + 128)) && alloc((void **)&r->psrc, 4 * (2 * nn + 128)) &&
              alloc((void **)&r->pwidth, 4 * (nn + 64));
    if (!ok) {
        cudaGetLastError();
        rank_free_problem(r);
        return r->fail(RFB_ERR_NOMEM, "rank %d: cannot allocate the own block columns (%zu bytes) and the panel ring", r->rank,
                       nn * (size_t)std::max<int64_t>(r->ncl, 1) * r->es);
    }
    r->ctx->perm_dst = r->pdst; r->ctx->perm_src = r->psrc; r->ctx->perm_width = r->pwidth;
    r->ctx->perm_cap = nn + 64;
    r->ctx->perm_external = true;
    r->ev_blk.assign(P.nblk, nullptr);
    r->ev_fact.assign(P.nblk, nullptr);
    r->ev_up.assign(P.nblk, nullptr);
    r->tev_c0.assign(P.nblk, nullptr); r->tev_c1.assign(P.nblk, nullptr);
    r->tev_p0.assign(P.nblk, nullptr); r->tev_p1.assign(P.nblk, nullptr);
    r->up_pending.assign(P.nblk, 0);
    for (int j = 0; j < P.nblk; ++j) {
        MG_CUDA(r, cudaEventCreateWithFlags(&r->ev_blk[j], cudaEventDisableTiming));
        if (P.owner(j) == r->rank) {
            MG_CUDA(r, cudaEventCreateWithFlags(&r->ev_fact[j], cudaEventDisableTiming));
            MG_CUDA(r, cudaEventCreateWithFlags(&r->ev_up[j], cudaEventDisableTiming));
            MG_CUDA(r, cudaEventCreate(&r->tev_c0[j])); MG_CUDA(r, cudaEventCreate(&r->tev_c1[j]));
            MG_CUDA(r, cudaEventCreate(&r->tev_p0[j])); MG_CUDA(r, cudaEventCreate(&r->tev_p1[j]));
        }
    }
    MG_CUDA(r, cudaMemset(r->info, 0, 64));
    MG_CUDA(r, cudaDeviceSynchronize());
    return RFB_OK;
}

int rank_create(MgRank *r) {
    rfb_ctx *ctx = nullptr;
    int rc = rfb_create(&ctx, r->device);
    r->ctx = ctx;
    if (rc != RFB_OK) return r->fail(rc, "%s", ctx ? ctx->last_error.c_str() : "rfb_create failed");
    r->s_comp = ctx->own_stream;
    int lo = 0, hi = 0;
    MG_CUDA(r, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    MG_CUDA(r, cudaStreamCreateWithPriority(&r->s_L, cudaStreamNonBlocking, hi));
    MG_CUDA(r, cudaStreamCreateWithFlags(&r->s_copy, cudaStreamNonBlocking));
    MG_CUDA(r, cudaEventCreateWithFlags(&r->ev_sent, cudaEventDisableTiming));
    MG_CUDA(r, cudaEventCreateWithFlags(&r->ev_final, cudaEventDisableTiming));
    MG_CUDA(r, cudaEventCreateWithFlags(&r->ev_chunk[0], cudaEventDisableTiming));
    MG_CUDA(r, cudaEventCreateWithFlags(&r->ev_chunk[1], cudaEventDisableTiming));
    for (int i = 0; i < MgRank::kRing; ++i) MG_CUDA(r, cudaEventCreateWithFlags(&r->ev_slot[i], cudaEventDisableTiming));
    MG_CUDA(r, cudaEventCreate(&r->ev_t0));
    MG_CUDA(r, cudaEventCreate(&r->ev_t1));
    if (const char *e = getenv("RFB_MG_MERGE")) r->merge_max = std::max(1, atoi(e));
    if (const char *e = getenv("RFB_MG_SLICE_US")) { r->slice_us = atof(e); r->slice_env = true; }
    if (const char *e = getenv("RFB_MG_ALERT_US")) r->alert_us = atof(e);
    return RFB_OK;
}

int rank_comm_init(MgRank *r, const ncclUniqueId &id) {
    NcclApi *api = nccl_api();
    if (!api->error.empty()) return r->fail(RFB_ERR_NCCL, "%s", api->error.c_str());
    MG_CUDA(r, cudaSetDevice(r->device));
    if (api->CommInitRankConfig) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        // the broadcast kernel waits for the owner while GEMM waves run beside it: keep its footprint to a few SMs
        const char *e = getenv("RFB_MG_NCCL_MAX_CTAS");
        cfg.maxCTAs = e ? atoi(e) : 16;     // measured on 8 GPUs: 16 -> 188 ms, 8 -> 195 ms per 32768^2 factorization
        cfg.minCTAs = 1;
        MG_NCCL(r, api->CommInitRankConfig(&r->comm, r->world, id, r->rank, &cfg));
    } else {
        MG_NCCL(r, api->CommInitRank(&r->comm, r->world, id, r->rank));
    }
    return RFB_OK;
}

void rank_destroy(MgRank *r) {
    if (!r) return;
    if (r->ctx) {
        cudaSetDevice(r->device);
        rank_free_problem(r);
        if (r->comm) nccl_api()->CommDestroy(r->comm);
        for (cudaEvent_t e : {r->ev_sent, r->ev_final, r->ev_chunk[0], r->ev_chunk[1], r->ev_t0, r->ev_t1})
            if (e) cudaEventDestroy(e);
        for (int i = 0; i < MgRank::kRing; ++i) if (r->ev_slot[i]) cudaEventDestroy(r->ev_slot[i]);
        if (r->s_L) cudaStreamDestroy(r->s_L);
        if (r->s_copy) cudaStreamDestroy(r->s_copy);
        rfb_destroy(r->ctx);
    }
    delete r;
}

// one factorization on one rank: reset, schedule, close the timer; returns after everything is ENQUEUED
template <typename T>
int rank_factor(MgRank *r) {
    MG_CUDA(r, cudaSetDevice(r->device));
    const MgPlan &P = r->plan;
    r->status = RFB_OK;
    r->ctx->stream = r->s_comp;
    r->ctx->lane = 0;
    for (int i = 0; i < rfb_ctx::kLanes; ++i) r->ctx->net_key[i] = -1;
    MG_CUDA(r, cudaEventRecord(r->ev_t0, r->s_comp));
    MG_CUDA(r, cudaMemsetAsync(r->info, 0, 64, r->s_comp));
    MG_CUDA(r, cudaMemsetAsync(r->pdst, 0xFF, 4 * (2 * (size_t)P.n + 128), r->s_comp));
    MG_CUDA(r, cudaMemsetAsync(r->psrc, 0xFF, 4 * (2 * (size_t)P.n + 128), r->s_comp));
    MG_CUDA(r, cudaMemsetAsync(r->pwidth, 0, 4 * ((size_t)P.n + 64), r->s_comp));
    // the replica stream must not unpack into the lists before they are cleared
    MG_CUDA(r, cudaEventRecord(r->ev_sent, r->s_comp));
    MG_CUDA(r, cudaStreamWaitEvent(r->s_L, r->ev_sent, 0));
    MgSched<T> sched(r);
    int rc = sched.run();
    if (rc != RFB_OK) {
        if (r->comm && r->world > 1) nccl_api()->CommAbort(r->comm), r->comm = nullptr;     // never leave a collective hanging on the device
        return rc;
    }
    // the factorization is complete when BOTH streams are: the compute stream joins the communication stream, then closes the timer
    MG_CUDA(r, cudaEventRecord(r->ev_sent, r->s_L));
    MG_CUDA(r, cudaStreamWaitEvent(r->s_comp, r->ev_sent, 0));
    MG_CUDA(r, cudaEventRecord(r->ev_t1, r->s_comp));
    MG_CUDA(r, cudaEventRecord(r->ev_final, r->s_comp));
    return RFB_OK;
}

int rank_sync(MgRank *r, float *ms) {
    MG_CUDA(r, cudaSetDevice(r->device));
    MG_CUDA(r, cudaStreamSynchronize(r->s_comp));
    MG_CUDA(r, cudaStreamSynchronize(r->s_L));
    MG_CUDA(r, cudaStreamSynchronize(r->s_copy));
    unsigned int flag = 0;
    MG_CUDA(r, cudaMemcpy(&flag, &r->ctx->xchg->error_flag, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag) {
        cudaMemset(&r->ctx->xchg->error_flag, 0, sizeof(unsigned int));
        return r->fail(RFB_ERR_INTERNAL, "rank %d: device-side protocol error (flag %u)", r->rank, flag);
    }
    if (ms) {
        *ms = 0;
        if (cudaEventQuery(r->ev_t1) == cudaSuccess && cudaEventElapsedTime(ms, r->ev_t0, r->ev_t1) != cudaSuccess) { cudaGetLastError(); *ms = 0; }
    }
    return RFB_OK;
}

int mg_fail(rfb_mg *mg, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    mg->last_error = buf;
    return code;
}

int collect(rfb_mg *mg) {
    for (MgRank *r : mg->ranks)
        if (r->status != RFB_OK) { mg->last_error = r->error; return r->status; }
    return RFB_OK;
}

// run fn(rank) on every local rank: inline for one rank, one host thread per device otherwise
template <typename F>
int for_ranks(rfb_mg *mg, F fn) {
    if (mg->ranks.size() == 1) { fn(mg->ranks[0]); return collect(mg); }
    std::vector<std::thread> th;
    for (MgRank *r : mg->ranks) th.emplace_back([r, &fn] { fn(r); });
    for (auto &t : th) t.join();
    return collect(mg);
}

MgRank *local(rfb_mg *mg, int lr) { return (mg && lr >= 0 && lr < (int)mg->ranks.size()) ? mg->ranks[lr] : nullptr; }

template <typename T>
int mg_lu_host(rfb_mg *mg, T *A, int64_t n, int64_t lda, int64_t *ipiv, int64_t *info, int64_t nb) {
    if (!mg->all_mode) return mg_fail(mg, RFB_ERR_UNSUPPORTED, "rfb_mg_lu_*: whole-matrix host entry needs a one-process handle (rfb_mg_create_all)");
    if (n < 0 || lda < std::max<int64_t>(n, 1) || !info) return mg_fail(mg, RFB_ERR_ARG, "bad n / lda / info");
    if (n == 0) { *info = 0; return RFB_OK; }
    if (!A || !ipiv) return mg_fail(mg, RFB_ERR_ARG, "A or ipiv is null");
    if (nb <= 0) nb = n < 8192 ? 256 : (mg->world >= 8 ? 256 : (mg->world >= 5 ? 512 : 1024));   // measured at 32768^2, see bench.py
    if (nb % 64) return mg_fail(mg, RFB_ERR_ARG, "block width must be a multiple of 64");
    int rc = rfb_mg_setup(mg, n, nb, sizeof(T) == 4);
    if (rc != RFB_OK) return rc;
    // upload (each rank its own block columns, in block order: the factorization starts on the first ones), factor, download
    rc = for_ranks(mg, [&](MgRank *r) {
        cudaSetDevice(r->device);
        for (int j : r->own) {
            const int64_t c0 = r->plan.col0(j), w = r->plan.width(j);
            if (cudaMemcpy2DAsync(r->Aj(0, r->lcol[j]), n * sizeof(T), A + c0 * lda, lda * sizeof(T), n * sizeof(T), w, cudaMemcpyHostToDevice,
                                  r->s_copy) != cudaSuccess ||
                cudaEventRecord(r->ev_up[j], r->s_copy) != cudaSuccess) { r->fail(RFB_ERR_CUDA, "upload of block %d failed", j); return; }
            r->up_pending[j] = 1;
        }
        if (rank_factor<T>(r) != RFB_OK) return;
        if (cudaStreamWaitEvent(r->s_copy, r->ev_final, 0) != cudaSuccess) { r->fail(RFB_ERR_CUDA, "stream wait failed"); return; }
        for (int j : r->own) {
            const int64_t c0 = r->plan.col0(j), w = r->plan.width(j);
            if (cudaMemcpy2DAsync(A + c0 * lda, lda * sizeof(T), r->Aj(0, r->lcol[j]), n * sizeof(T), n * sizeof(T), w, cudaMemcpyDeviceToHost,
                                  r->s_copy) != cudaSuccess) { r->fail(RFB_ERR_CUDA, "download of block %d failed", j); return; }
        }
        float ms = 0;
        rank_sync(r, &ms);
    });
    if (rc != RFB_OK) return rc;
    rc = rfb_mg_get_pivots(mg, ipiv);
    if (rc != RFB_OK) return rc;
    return rfb_mg_get_info(mg, info);
}

}  // namespace

// =================================================================================================================================
extern "C" {

const char *rfb_mg_last_error(rfb_mg *mg) { return mg ? mg->last_error.c_str() : "null handle"; }

int rfb_mg_unique_id(void *id128) {
    if (!id128) return RFB_ERR_ARG;
    NcclApi *api = nccl_api();
    if (!api->error.empty()) return RFB_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return RFB_ERR_NCCL;
    memcpy(id128, &id, sizeof(id));
    return RFB_OK;
}

int rfb_mg_create_rank(rfb_mg **out, int device, int rank, int nranks, const void *id128) {
    if (!out) return RFB_ERR_ARG;
    rfb_mg *mg = new rfb_mg();
    *out = mg;
    if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) return mg_fail(mg, RFB_ERR_ARG, "bad rank / nranks / id");
    mg->world = nranks;
    MgRank *r = new MgRank();
    r->mg = mg; r->rank = rank; r->world = nranks; r->device = device;
    mg->ranks.push_back(r);
    if (rank_create(r) != RFB_OK) return collect(mg);
    if (nranks > 1) {
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        if (rank_comm_init(r, id) != RFB_OK) return collect(mg);
    }
    return RFB_OK;
}

int rfb_mg_create_all(rfb_mg **out, int ngpus, const int *devices) {
    if (!out) return RFB_ERR_ARG;
    rfb_mg *mg = new rfb_mg();
    *out = mg;
    if (ngpus < 1) return mg_fail(mg, RFB_ERR_ARG, "ngpus must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < ngpus) {
        cudaGetLastError();
        return mg_fail(mg, RFB_ERR_CUDA, "%d GPUs requested, %d visible; librfb200 has no CPU fallback", ngpus, ndev);
    }
    mg->world = ngpus;
    mg->all_mode = true;
    for (int i = 0; i < ngpus; ++i) {
        MgRank *r = new MgRank();
        r->mg = mg; r->rank = i; r->world = ngpus; r->device = devices ? devices[i] : i;
        mg->ranks.push_back(r);
        if (rank_create(r) != RFB_OK) return collect(mg);
    }
    if (ngpus > 1) {
        NcclApi *api = nccl_api();
        if (!api->error.empty()) return mg_fail(mg, RFB_ERR_NCCL, "%s", api->error.c_str());
        ncclUniqueId id;
        if (api->GetUniqueId(&id) != ncclSuccess) return mg_fail(mg, RFB_ERR_NCCL, "ncclGetUniqueId failed");
        api->GroupStart();                                    // one process, G devices: all communicators inside one group
        for (MgRank *r : mg->ranks)
            if (rank_comm_init(r, id) != RFB_OK) break;
        if (api->GroupEnd() != ncclSuccess && collect(mg) == RFB_OK) return mg_fail(mg, RFB_ERR_NCCL, "ncclGroupEnd failed");
        if (collect(mg) != RFB_OK) return collect(mg);
    }
    return RFB_OK;
}

int rfb_mg_destroy(rfb_mg *mg) {
    if (!mg) return RFB_OK;
    for (MgRank *r : mg->ranks) rank_destroy(r);
    delete mg;
    return RFB_OK;
}

int rfb_mg_setup(rfb_mg *mg, int64_t n, int64_t nb, int is_f32) {
    if (!mg) return RFB_ERR_ARG;
    if (n <= 0 || nb < 64 || nb % 64) return mg_fail(mg, RFB_ERR_ARG, "need n > 0 and a block width that is a positive multiple of 64");
    if (n > 0x7ffffff0LL) return mg_fail(mg, RFB_ERR_UNSUPPORTED, "dimension exceeds int32");
    if (mg->n == n && mg->nb == nb && mg->f32 == (is_f32 != 0) && !mg->ranks.empty() && mg->ranks[0]->L) return RFB_OK;   // buffers are reused
    mg->n = n; mg->nb = nb; mg->f32 = is_f32 != 0;
    return for_ranks(mg, [&](MgRank *r) { rank_setup(r, n, nb, is_f32 != 0); });
}

int rfb_mg_owner_of(int64_t block, int world) {
    if (block < 0 || world < 1) return -1;
    MgPlan p;
    p.world = world;
    return p.owner((int)block);
}

int rfb_mg_local_ranks(rfb_mg *mg, int *count, int *world) {
    if (!mg) return RFB_ERR_ARG;
    if (count) *count = (int)mg->ranks.size();
    if (world) *world = mg->world;
    return RFB_OK;
}

int rfb_mg_rank_ctx(rfb_mg *mg, int lr, rfb_ctx **ctx, int *global_rank) {
    MgRank *r = local(mg, lr);
    if (!r) return RFB_ERR_ARG;
    if (ctx) *ctx = r->ctx;
    if (global_rank) *global_rank = r->rank;
    return RFB_OK;
}

int rfb_mg_block_ptr(rfb_mg *mg, int lr, int64_t j, void **dev_ptr) {
    MgRank *r = local(mg, lr);
    if (!r || !dev_ptr || !r->A) return RFB_ERR_ARG;
    if (j < 0 || j >= r->plan.nblk || r->plan.owner((int)j) != r->rank) return mg_fail(mg, RFB_ERR_ARG, "block %lld is not owned by rank %d", (long long)j, r->rank);
    *dev_ptr = r->Aj(0, r->lcol[j]);
    return RFB_OK;
}

// src_kind: 0 host (H2D), 1 device (D2D); the copy runs on the rank's copy stream and the factorization waits for it
int rfb_mg_load_block(rfb_mg *mg, int lr, int64_t j, const void *src, int64_t ld, int src_kind) {
    MgRank *r = local(mg, lr);
    if (!r || !src || !r->A) return RFB_ERR_ARG;
    if (j < 0 || j >= r->plan.nblk || r->plan.owner((int)j) != r->rank) return mg_fail(mg, RFB_ERR_ARG, "block %lld is not owned by rank %d", (long long)j, r->rank);
    if (ld < r->plan.n) return mg_fail(mg, RFB_ERR_ARG, "ld < n");
    cudaSetDevice(r->device);
    const size_t es = r->es, n = (size_t)r->plan.n;
    cudaError_t e = cudaMemcpy2DAsync(r->Aj(0, r->lcol[j]), n * es, src, (size_t)ld * es, n * es, (size_t)r->plan.width((int)j),
                                      src_kind ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, r->s_copy);
    if (e == cudaSuccess) e = cudaEventRecord(r->ev_up[j], r->s_copy);
    if (e != cudaSuccess) return mg_fail(mg, RFB_ERR_CUDA, "load of block %lld failed: %s", (long long)j, cudaGetErrorString(e));
    r->up_pending[j] = 1;
    return RFB_OK;
}

int rfb_mg_store_block(rfb_mg *mg, int lr, int64_t j, void *dst_host, int64_t ld) {
    MgRank *r = local(mg, lr);
    if (!r || !dst_host || !r->A) return RFB_ERR_ARG;
    if (j < 0 || j >= r->plan.nblk || r->plan.owner((int)j) != r->rank) return mg_fail(mg, RFB_ERR_ARG, "block %lld is not owned by rank %d", (long long)j, r->rank);
    cudaSetDevice(r->device);
    const size_t es = r->es, n = (size_t)r->plan.n;
    cudaError_t e = cudaStreamWaitEvent(r->s_copy, r->ev_final, 0);
    if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(dst_host, (size_t)ld * es, r->Aj(0, r->lcol[j]), n * es, n * es, (size_t)r->plan.width((int)j), cudaMemcpyDeviceToHost, r->s_copy);
    if (e != cudaSuccess) return mg_fail(mg, RFB_ERR_CUDA, "store of block %lld failed: %s", (long long)j, cudaGetErrorString(e));
    return RFB_OK;
}

int rfb_mg_factor(rfb_mg *mg) {
    if (!mg || mg->ranks.empty() || !mg->ranks[0]->L) return mg ? mg_fail(mg, RFB_ERR_ARG, "rfb_mg_setup first") : RFB_ERR_ARG;
    if (mg->f32) return for_ranks(mg, [](MgRank *r) { rank_factor<float>(r); });
    return for_ranks(mg, [](MgRank *r) { rank_factor<double>(r); });
}

int rfb_mg_sync(rfb_mg *mg, float *ms) {
    if (!mg) return RFB_ERR_ARG;
    float worst = 0;
    for (MgRank *r : mg->ranks) {
        float t = 0;
        if (rank_sync(r, &t) != RFB_OK) return collect(mg);
        worst = std::max(worst, t);
    }
    mg->last_ms = worst;
    if (ms) *ms = worst;      // max over the LOCAL ranks; one-process-per-GPU callers reduce over processes themselves
    return RFB_OK;
}

int rfb_mg_get_pivots(rfb_mg *mg, int64_t *ipiv_host) {
    MgRank *r = local(mg, 0);
    if (!r || !ipiv_host || !r->ipiv) return RFB_ERR_ARG;
    cudaSetDevice(r->device);
    if (cudaStreamSynchronize(r->s_L) != cudaSuccess || cudaStreamSynchronize(r->s_comp) != cudaSuccess ||
        cudaMemcpy(ipiv_host, r->ipiv, 8 * (size_t)r->plan.n, cudaMemcpyDeviceToHost) != cudaSuccess)
        return mg_fail(mg, RFB_ERR_CUDA, "pivot download failed: %s", cudaGetErrorString(cudaGetLastError()));
    return RFB_OK;
}

// global info: the smallest non-zero per-rank value (first exactly-zero pivot column, src/lu.jl:321-327), else 0
int rfb_mg_get_info(rfb_mg *mg, int64_t *info) {
    if (!mg || !info || mg->ranks.empty()) return RFB_ERR_ARG;
    int64_t best = 0;
    for (MgRank *r : mg->ranks) {
        cudaSetDevice(r->device);
        int64_t v = 0;
        if (cudaStreamSynchronize(r->s_comp) != cudaSuccess || cudaMemcpy(&v, r->info, 8, cudaMemcpyDeviceToHost) != cudaSuccess)
            return mg_fail(mg, RFB_ERR_CUDA, "info download failed");
        if (v != 0 && (best == 0 || v < best)) best = v;
    }
    if (!mg->all_mode && mg->world > 1) {                 // one process per GPU: min over the ranks through NCCL
        MgRank *r = mg->ranks[0];
        int64_t *d = reinterpret_cast<int64_t *>(r->info) + 1;
        const int64_t big = INT64_MAX, mine = best == 0 ? big : best;
        if (cudaMemcpy(d, &mine, 8, cudaMemcpyHostToDevice) != cudaSuccess) return mg_fail(mg, RFB_ERR_CUDA, "info upload failed");
        if (nccl_api()->AllReduce(d, d, 1, ncclInt64, ncclMin, r->comm, r->s_L) != ncclSuccess) return mg_fail(mg, RFB_ERR_NCCL, "info all-reduce failed");
        int64_t g = 0;
        if (cudaStreamSynchronize(r->s_L) != cudaSuccess || cudaMemcpy(&g, d, 8, cudaMemcpyDeviceToHost) != cudaSuccess)
            return mg_fail(mg, RFB_ERR_CUDA, "info download failed");
        best = g == big ? 0 : g;
    }
    *info = best;
    return RFB_OK;
}

int rfb_mg_stats(rfb_mg *mg, int64_t *bcast_bytes_per_rank, int64_t *launches) {
    MgRank *r = local(mg, 0);
    if (!r) return RFB_ERR_ARG;
    if (bcast_bytes_per_rank) *bcast_bytes_per_rank = r->bcast_bytes;
    if (launches) { int64_t s = 0; for (MgRank *q : mg->ranks) s += q->ctx->launches; *launches = s; }
    return RFB_OK;
}

// scheduler statistics of local rank lr's last factorization: out[0] bulk slices enqueued, out[1] critical-path enqueues (own
// block columns), out[2] host microseconds with an idle compute stream and nothing runnable, out[3] host microseconds of the whole
// schedule loop, out[4] kernels launched so far by the rank's context
int rfb_mg_sched_stats(rfb_mg *mg, int lr, int64_t out[8]) {
    MgRank *r = local(mg, lr);
    if (!r || !out) return RFB_ERR_ARG;
    for (int i = 0; i < 8; ++i) out[i] = 0;
    out[0] = r->st_slices; out[1] = r->st_crit; out[2] = (int64_t)r->st_idle_us; out[3] = (int64_t)r->st_wall_us;
    out[4] = r->ctx ? r->ctx->launches : 0;
    // [5] device microseconds spent in the owned block columns' critical sections (last contributions + factorization),
    // [6] device microseconds from "factored" to "published here" (pack + broadcast + unpack) of the owned block columns
    if (r->ctx && !r->tev_c0.empty()) {
        cudaSetDevice(r->device);
        double crit = 0, pub = 0;
        for (int j : r->own) {
            float t = 0;
            if (cudaEventElapsedTime(&t, r->tev_c0[j], r->tev_c1[j]) == cudaSuccess) crit += t;
            if (cudaEventElapsedTime(&t, r->tev_p0[j], r->tev_p1[j]) == cudaSuccess) pub += t;
        }
        cudaGetLastError();
        out[5] = (int64_t)(crit * 1e3);
        out[6] = (int64_t)(pub * 1e3);
    }
    return RFB_OK;
}

int rfb_mg_lu_f64(rfb_mg *mg, double *A_host, int64_t n, int64_t lda, int64_t *ipiv, int64_t *info, int64_t block) {
    if (!mg) return RFB_ERR_ARG;
    return mg_lu_host<double>(mg, A_host, n, lda, ipiv, info, block);
}
int rfb_mg_lu_f32(rfb_mg *mg, float *A_host, int64_t n, int64_t lda, int64_t *ipiv, int64_t *info, int64_t block) {
    if (!mg) return RFB_ERR_ARG;
    return mg_lu_host<float>(mg, A_host, n, lda, ipiv, info, block);
}

// convenience: create, factor, destroy (pays the NCCL start-up on every call; keep a handle for repeated use)
int rfb_lu_f64_mg(const int *devices, int ngpus, double *A_host, int64_t n, int64_t lda, int64_t *ipiv, int64_t *info, int64_t block) {
    rfb_mg *mg = nullptr;
    int rc = rfb_mg_create_all(&mg, ngpus, devices);
    if (rc == RFB_OK) rc = rfb_mg_lu_f64(mg, A_host, n, lda, ipiv, info, block);
    if (rc != RFB_OK && mg) fprintf(stderr, "rfb_lu_f64_mg: %s\n", mg->last_error.c_str());
    rfb_mg_destroy(mg);
    return rc;
}
int rfb_lu_f32_mg(const int *devices, int ngpus, float *A_host, int64_t n, int64_t lda, int64_t *ipiv, int64_t *info, int64_t block) {
    rfb_mg *mg = nullptr;
    int rc = rfb_mg_create_all(&mg, ngpus, devices);
    if (rc == RFB_OK) rc = rfb_mg_lu_f32(mg, A_host, n, lda, ipiv, info, block);
    if (rc != RFB_OK && mg) fprintf(stderr, "rfb_lu_f32_mg: %s\n", mg->last_error.c_str());
    rfb_mg_destroy(mg);
    return rc;
}

// Dry run of one rank's schedule (no GPU, no NCCL): 5 int64 per operation, in the order the rank would enqueue them when every
// input is already there:  [1, c0, n1, j, 0] update of block column j by the node whose left half is columns [c0, c0 + n1)
// (:233-240);  [2, j, c0, w, 0] factor block column j;  [3, j, root, c0, w] broadcast of block column j;
// [4, c0, n1, k0, k1] A21 <- P2 A21 (:246): pivots [k0, k1) applied to columns [c0, c0 + n1).
int rfb_mg_trace(int64_t n, int64_t nb, int rank, int world, int64_t *ops, int64_t cap, int64_t *count) {
    if (!count || n <= 0 || nb <= 0 || world < 1 || rank < 0 || rank >= world) return RFB_ERR_ARG;
    MgRank r;
    r.rank = rank; r.world = world; r.dry = true;
    std::vector<int64_t> tr;
    r.trace = &tr;
    r.plan.build(n, nb, world);
    r.lcol.assign(r.plan.nblk, -1);
    for (int j = 0; j < r.plan.nblk; ++j)
        if (r.plan.owner(j) == rank) { r.own.push_back(j); r.lcol[j] = r.ncl; r.ncl += r.plan.width(j); }
    MgSched<double> sched(&r);
    const int rc = sched.run();
    if (rc != RFB_OK) return rc;
    *count = (int64_t)tr.size() / 5;
    if (ops) {
        const int64_t nout = std::min<int64_t>(*count, cap) * 5;
        for (int64_t i = 0; i < nout; ++i) ops[i] = tr[(size_t)i];
    }
    return RFB_OK;
}

}  // extern "C"
