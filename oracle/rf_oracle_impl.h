/*
 * rf_oracle_impl.h -- type-generic body of the CPU oracle (TEST INFRASTRUCTURE, not product).
 *
 * Included twice by rf_oracle.c with
 *     #define RFO_T      double | float
 *     #define RFO_(name) name##_f64 | name##_f32
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  Parity status: see the header of rf_oracle.c.
 *
 * Multiply-subtract is written as an explicit fused multiply-add (RFO_FMA): the reference's
 * @turbo/@tturbo loops (src/lu.jl:268,276,330) compile to FMA on every FMA-capable CPU, and an
 * explicit fma() makes the oracle's bits independent of the host it runs on.
 */

/* ---- src/lu.jl:290-338  _generic_lufact!(A, Val(Pivot), ipiv, info) ------------------------
 * Unblocked right-looking LU of the m x n block A with npiv = length(ipiv) pivot steps.
 * pivot == 0 is Val(false): kp = k, ipiv is not touched (NotIPIV, :27-32), and a zero pivot is
 * reported as NEGATIVE info (the Julia >= 1.11 convention, :24-25 and :323-326).
 * pivot = FIRST index of the strict maximum |A[i,k]| starting from amax = 0 (:296-305);
 * swap rows k,kp over all n columns of the block (:308-315); scale by the RECIPROCAL
 * (:317-320); info = first k with an exactly-zero pivot, factorization continues (:321-327);
 * rank-1 update of the remaining columns (:330-334).  ipiv is 1-based and block-local. */
RFO_CLONES
static int64_t RFO_(rfo_generic_lufact)(RFO_T *A, int64_t m, int64_t n, int64_t lda,
                                        int64_t *ipiv, int64_t npiv, int64_t info, int pivot)
{
    for (int64_t k = 0; k < npiv; ++k) {
        RFO_T *ck = A + k * lda;
        int64_t kp = k;
        if (pivot) {
            RFO_T amax = (RFO_T)0;
            for (int64_t i = k; i < m; ++i) {
                RFO_T absi = ck[i] < 0 ? -ck[i] : ck[i];
                if (absi > amax) { kp = i; amax = absi; }   /* NaN never wins: (NaN > x) is false */
            }
            ipiv[k] = kp + 1;
        }
        if (ck[kp] != (RFO_T)0) {                        /* !iszero: NaN counts as non-zero */
            if (k != kp) {
                for (int64_t j = 0; j < n; ++j) {
                    RFO_T t = A[k + j * lda];
                    A[k + j * lda] = A[kp + j * lda];
                    A[kp + j * lda] = t;
                }
            }
            RFO_T inv = (RFO_T)1 / ck[k];
            for (int64_t i = k + 1; i < m; ++i) ck[i] *= inv;
        } else if (info == 0) {
            info = pivot ? k + 1 : -(k + 1);             /* :321-327 */
        }
        if (k == npiv - 1) break;
        for (int64_t j = k + 1; j < n; ++j) {
            RFO_T *cj = A + j * lda;
            RFO_T akj = cj[k];
            #pragma omp simd
            for (int64_t i = k + 1; i < m; ++i) cj[i] = RFO_FMA(-ck[i], akj, cj[i]);
        }
    }
    return info;
}

/* ---- src/lu.jl:164-188  apply_permutation!(P, A, thread) ------------------------------------
 * Sequential row interchanges i <-> P[i] (1-based, local to A's first row) on an m x ncols
 * block.  The threaded form (:164-175) walks columns in parallel, the serial form (:177-188)
 * walks swaps and skips i' == i; both give the same result. */
static void RFO_(rfo_apply_permutation)(const int64_t *P, int64_t np, RFO_T *A, int64_t ncols,
                                        int64_t lda, int threads)
{
    #pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (int64_t j = 0; j < ncols; ++j) {
        RFO_T *c = A + j * lda;
        for (int64_t i = 0; i < np; ++i) {
            int64_t ip = P[i] - 1;
            RFO_T t = c[i]; c[i] = c[ip]; c[ip] = t;
        }
    }
}

/* ---- src/lu.jl:265-284  schur_complement!(C, A, B, thread) ----------------------------------
 * C[m,n] = C[m,n] + (0 - sum_k A[m,k] B[k,n]): the product is accumulated in its own register
 * and added to C once (:269-273).  Register tiling only, no packing, like the @turbo loop nest;
 * column blocks are spread over threads like @tturbo. */
#define RFO_MR 16          /* rows per register tile: two 8-wide (f64) / one 16-wide (f32) vectors x2 */
#define RFO_NR 6           /* columns per register tile */
typedef RFO_T RFO_(rfo_vec) __attribute__((vector_size(8 * sizeof(RFO_T)), aligned(sizeof(RFO_T))));
RFO_CLONES
static void RFO_(rfo_schur_tile)(RFO_T *C, const RFO_T *A, const RFO_T *B, int64_t m, int64_t nn,
                                 int64_t k, int64_t lda)
{
    /* nn <= RFO_NR columns of C; 12 vector accumulators stay in registers over the whole k loop */
    typedef RFO_(rfo_vec) vec;
    int64_t i0 = 0;
    if (nn == RFO_NR) {
        for (; i0 + RFO_MR <= m; i0 += RFO_MR) {
            vec acc0[RFO_NR], acc1[RFO_NR];
            for (int c = 0; c < RFO_NR; ++c) { acc0[c] = (vec){0}; acc1[c] = (vec){0}; }
            for (int64_t kk = 0; kk < k; ++kk) {
                const vec a0 = *(const vec *)(A + i0 + kk * lda);
                const vec a1 = *(const vec *)(A + i0 + 8 + kk * lda);
                for (int c = 0; c < RFO_NR; ++c) {
                    const RFO_T b = B[kk + c * lda];
                    acc0[c] -= a0 * b;          /* contracted to one fused multiply-add per lane */
                    acc1[c] -= a1 * b;
                }
            }
            for (int c = 0; c < RFO_NR; ++c) {
                vec *c0 = (vec *)(C + i0 + c * lda), *c1 = (vec *)(C + i0 + 8 + c * lda);
                *c0 = acc0[c] + *c0;
                *c1 = acc1[c] + *c1;
            }
        }
    }
    for (; i0 < m; ++i0) {
        for (int c = 0; c < nn; ++c) {
            RFO_T acc = (RFO_T)0;
            for (int64_t kk = 0; kk < k; ++kk) acc = RFO_FMA(-A[i0 + kk * lda], B[kk + c * lda], acc);
            C[i0 + c * lda] = acc + C[i0 + c * lda];
        }
    }
}

static void RFO_(rfo_schur_complement)(RFO_T *C, const RFO_T *A, const RFO_T *B, int64_t m,
                                       int64_t n, int64_t k, int64_t lda, int threads)
{
    int64_t nblk = (n + RFO_NR - 1) / RFO_NR;
    #pragma omp parallel for schedule(dynamic, 2) num_threads(threads) if (threads > 1)
    for (int64_t jb = 0; jb < nblk; ++jb) {
        int64_t j0 = jb * RFO_NR;
        int64_t nn = n - j0 < RFO_NR ? n - j0 : RFO_NR;
        RFO_(rfo_schur_tile)(C + j0 * lda, A, B + j0 * lda, m, nn, k, lda);
    }
}

/* ---- TriangularSolve.ldiv!(UnitLowerTriangular(A11), A12, thread) ---------------------------
 * Call sites src/lu.jl:235 and :153.  TriangularSolve.jl is an un-vendored dependency
 * (Project.toml:23, compat 0.2.5, no Manifest => exact version unpinned).  Its published
 * algorithm for the left/unit-lower case is block forward substitution: solve a diagonal block
 * by substitution, then subtract its contribution from the rows below.  Restated here with
 * block size RFO_TB and the accumulate-then-add update above.  Only the STRICT lower triangle
 * of L is read (its diagonal and upper part hold U). */
#define RFO_TB 64
RFO_CLONES
static void RFO_(rfo_trsm_llnu)(const RFO_T *L, int64_t k, RFO_T *B, int64_t nrhs, int64_t lda,
                                int threads)
{
    for (int64_t b0 = 0; b0 < k; b0 += RFO_TB) {
        int64_t bs = k - b0 < RFO_TB ? k - b0 : RFO_TB;
        #pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
        for (int64_t j = 0; j < nrhs; ++j) {
            RFO_T *x = B + b0 + j * lda;
            for (int64_t c = 0; c < bs; ++c) {
                const RFO_T *l = L + b0 + (b0 + c) * lda;
                RFO_T xc = x[c];
                #pragma omp simd
                for (int64_t r = c + 1; r < bs; ++r) x[r] = RFO_FMA(-l[r], xc, x[r]);
            }
        }
        int64_t rest = k - b0 - bs;
        if (rest > 0)
            RFO_(rfo_schur_complement)(B + b0 + bs, L + b0 + bs + b0 * lda, B + b0, rest, nrhs, bs,
                                       lda, threads);
    }
}

/* ---- src/lu.jl:158-162  nsplit(T, n) --------------------------------------------------------*/
static int64_t RFO_(rfo_nsplit)(int64_t n)
{
    int64_t k = 128 / (int64_t)sizeof(RFO_T);
    if (k < 2) k = 2;
    int64_t k2 = k / 2;
    return n >= k ? ((n + k2) / k) * k2 : n / 2;
}

/* ---- src/lu.jl:189-263  reckernel!(A, Val(Pivot), m, n, ipiv, info, blocksize, thread) ------
 * pivot == 0: no row interchanges (:233, :246 are guarded by `Pivot &&`), ipiv untouched,
 * negative info shifted by -n1 (:249-251). */
static int64_t RFO_(rfo_reckernel)(RFO_T *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv,
                                   int64_t info, int64_t blocksize, int threads, int pivot)
{
    if (n <= (blocksize > 1 ? blocksize : 1))                                   /* :192-195 */
        return RFO_(rfo_generic_lufact)(A, m, n, lda, ipiv, n, info, pivot);
    int64_t n1 = RFO_(rfo_nsplit)(n), n2 = n - n1, m2 = m - n1;                 /* :196-198 */
    RFO_T *AR = A + n1 * lda, *A21 = A + n1, *A22 = A + n1 + n1 * lda;          /* :210-218 */
    int64_t *P1 = ipiv, *P2 = ipiv + n1;                                        /* :221-222 */
    info = RFO_(rfo_reckernel)(A, m, n1, lda, P1, info, blocksize, threads, pivot); /* :229 */
    if (pivot) RFO_(rfo_apply_permutation)(P1, n1, AR, n2, lda, threads);       /* :233 */
    RFO_(rfo_trsm_llnu)(A, n1, AR, n2, lda, threads);                           /* :235 */
    RFO_(rfo_schur_complement)(A22, A21, AR, m2, n2, n1, lda, threads);         /* :240 */
    int64_t previnfo = info;                                                    /* :242 */
    info = RFO_(rfo_reckernel)(A22, m2, n2, lda, P2, info, blocksize, threads, pivot); /* :244 */
    if (pivot) RFO_(rfo_apply_permutation)(P2, n2, A21, n1, lda, threads);      /* :246 */
    if (info != previnfo) info += (info < 0) ? -n1 : n1;                        /* :248-255 */
    if (pivot) for (int64_t i = 0; i < n2; ++i) P2[i] += n1;                    /* :256-260 */
    return info;
}

/* ---- src/lu.jl:97-130 (driver) + :145-156 (_recurse!, fat tail) -----------------------------
 * blocksize <= 0 selects the reference default (length(A) >= 40000 ? 8 : 16, :101);
 * threshold <= 0 selects pick_threshold() for a 64-byte SIMD register, i.e. 48 (:90,:102).
 * Returns info; never throws (checknonsingular, :128, is the caller's job). */
static int64_t RFO_(rfo_lu_impl)(RFO_T *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv,
                                 int64_t blocksize, int64_t threshold, int threads, int pivot)
{
    if (blocksize <= 0) blocksize = (m * n >= 40000) ? 8 : 16;
    if (threshold <= 0) threshold = 48;
    if (threads < 1) threads = 1;
    int64_t mn = m < n ? m : n, info = 0;
    if (mn == 0) return 0;
    if (!pivot && ipiv) for (int64_t i = 0; i < mn; ++i) ipiv[i] = i + 1;       /* :111-113 */
    if (mn > threshold) {                                                       /* :114 */
        info = RFO_(rfo_reckernel)(A, m, mn, lda, ipiv, info, blocksize, threads, pivot); /* :147 */
        if (m < n) {                                                            /* :148-154 */
            RFO_T *AR = A + m * lda;
            if (pivot) RFO_(rfo_apply_permutation)(ipiv, mn, AR, n - m, lda, threads);
            RFO_(rfo_trsm_llnu)(A, m, AR, n - m, lda, threads);
        }
    } else {
        info = RFO_(rfo_generic_lufact)(A, m, n, lda, ipiv, mn, info, pivot);   /* :125-126 */
    }
    return info;
}
int64_t RFO_(rfo_lu)(RFO_T *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv,
                     int64_t blocksize, int64_t threshold, int threads)
{ return RFO_(rfo_lu_impl)(A, m, n, lda, ipiv, blocksize, threshold, threads, 1); }
/* lu!(A, ipiv, Val(false), thread): ipiv may be NULL (NotIPIV) or a user vector that is filled
 * with 1:min(m,n) (:107-113). */
int64_t RFO_(rfo_lu_nopiv)(RFO_T *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv,
                           int64_t blocksize, int64_t threshold, int threads)
{ return RFO_(rfo_lu_impl)(A, m, n, lda, ipiv, blocksize, threshold, threads, 0); }

/* ---- src/lu.jl:60-64  ldiv!(F::LU{T,<:StridedMatrix,<:NotIPIV}, B) ---------------------------
 * B <- U^-1 (L^-1 B) with the packed factors; both legs are TriangularSolve.ldiv! calls (an
 * un-vendored dependency, see rfo_trsm_llnu): forward substitution with the unit lower triangle,
 * then column-oriented back substitution with the upper triangle (division by the diagonal). */
void RFO_(rfo_ldiv_notipiv)(const RFO_T *F, int64_t n, int64_t lda, RFO_T *B, int64_t nrhs, int64_t ldb)
{
    for (int64_t j = 0; j < nrhs; ++j) {
        RFO_T *x = B + j * ldb;
        for (int64_t c = 0; c < n; ++c) {
            const RFO_T *l = F + c * lda;
            RFO_T xc = x[c];
            for (int64_t r = c + 1; r < n; ++r) x[r] = RFO_FMA(-l[r], xc, x[r]);
        }
        for (int64_t c = n - 1; c >= 0; --c) {
            const RFO_T *u = F + c * lda;
            x[c] = x[c] / u[c];
            RFO_T xc = x[c];
            for (int64_t r = 0; r < c; ++r) x[r] = RFO_FMA(-u[r], xc, x[r]);
        }
    }
}

/* ---- src/butterflylu.jl:59-91  🦋mul_level!(A, u, v) ------------------------------------------
 * One butterfly level on the M x N block A (M, N even): with B_u = [D(u1) D(u2); D(u1) -D(u2)]
 * (u1/u2 = halves of u) this is A <- B_u' A B_v, written exactly as the reference's eight
 * additions and the (u * C) * v products (left-to-right, :85-88). */
static void RFO_(rfo_butterfly_level)(RFO_T *A, int64_t M, int64_t N, int64_t lda, const RFO_T *u,
                                      const RFO_T *v)
{
    int64_t Mh = M / 2, Nh = N / 2;
    for (int64_t n = 0; n < Nh; ++n) {
        for (int64_t m = 0; m < Mh; ++m) {
            RFO_T A11 = A[m + n * lda], A21 = A[m + Mh + n * lda];
            RFO_T A12 = A[m + (n + Nh) * lda], A22 = A[m + Mh + (n + Nh) * lda];
            RFO_T T1 = A11 + A12, T2 = A21 + A22, T3 = A11 - A12, T4 = A21 - A22;
            RFO_T C11 = T1 + T2, C21 = T1 - T2, C12 = T3 + T4, C22 = T3 - T4;
            RFO_T u1 = u[m], u2 = u[m + Mh], v1 = v[n], v2 = v[n + Nh];
            A[m + n * lda] = u1 * C11 * v1;
            A[m + Mh + n * lda] = u2 * C21 * v1;
            A[m + (n + Nh) * lda] = u1 * C12 * v2;
            A[m + Mh + (n + Nh) * lda] = u2 * C22 * v2;
        }
    }
}

/* ---- src/butterflylu.jl:93-113  🦋mul!(A, uv) ---------------------------------------------------
 * Two levels: the four quadrants with (U1|U2, V1|V2) = uv[0:M/2], uv[M:3M/2] / uv[M/2:M],
 * uv[3M/2:2M], then the whole matrix with U = uv[2M:3M], V = uv[3M:4M].  M % 4 == 0. */
void RFO_(rfo_butterfly_mul)(RFO_T *A, int64_t M, int64_t lda, const RFO_T *uv)
{
    int64_t Mh = M / 2;
    const RFO_T *U1 = uv, *V1 = uv + Mh, *U2 = uv + M, *V2 = uv + M + Mh;
    RFO_(rfo_butterfly_level)(A, Mh, Mh, lda, U1, V1);
    RFO_(rfo_butterfly_level)(A + Mh, Mh, Mh, lda, U2, V1);
    RFO_(rfo_butterfly_level)(A + Mh * lda, Mh, Mh, lda, U1, V2);
    RFO_(rfo_butterfly_level)(A + Mh + Mh * lda, Mh, Mh, lda, U2, V2);
    RFO_(rfo_butterfly_level)(A, M, M, lda, uv + 2 * M, uv + 3 * M);
}

/* Kernel-level entry points so the tests can check each CUDA kernel against the matching
 * restated loop in isolation. */
int64_t RFO_(rfo_panel)(RFO_T *A, int64_t m, int64_t n, int64_t lda, int64_t *ipiv, int64_t info)
{ return RFO_(rfo_generic_lufact)(A, m, n, lda, ipiv, n < m ? n : m, info, 1); }
int64_t RFO_(rfo_panel_nopiv)(RFO_T *A, int64_t m, int64_t n, int64_t lda, int64_t info)
{ return RFO_(rfo_generic_lufact)(A, m, n, lda, (int64_t *)0, n < m ? n : m, info, 0); }
void RFO_(rfo_laswp)(RFO_T *A, int64_t ncols, int64_t lda, const int64_t *ipiv, int64_t np)
{ RFO_(rfo_apply_permutation)(ipiv, np, A, ncols, lda, 1); }
void RFO_(rfo_trsm)(const RFO_T *L, int64_t k, RFO_T *B, int64_t nrhs, int64_t lda, int threads)
{ RFO_(rfo_trsm_llnu)(L, k, B, nrhs, lda, threads < 1 ? 1 : threads); }
void RFO_(rfo_schur)(RFO_T *C, const RFO_T *A, const RFO_T *B, int64_t m, int64_t n, int64_t k,
                     int64_t lda, int threads)
{ RFO_(rfo_schur_complement)(C, A, B, m, n, k, lda, threads < 1 ? 1 : threads); }
int64_t RFO_(rfo_nsplit_pub)(int64_t n) { return RFO_(rfo_nsplit)(n); }

#undef RFO_MR
#undef RFO_NR
#undef RFO_TB
