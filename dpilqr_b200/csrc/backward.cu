// backward.cu -- Kernel 3: backward Riccati recursion, one CTA per problem (sm_100a, FP64).
//
// Replaces ilqrSolver._backward_pass (reference control.py:116-148).  Per time step, with the
// block structure the reference throws away (A, B block-diagonal per agent; P symmetric):
//
//   phase A   S = B^T (P + mu I)  ->  Q_ux = S A,  Q_uu = L_uu + S B      (S only lives in registers)
//             Q_u = L_u + B^T p,  Q_x = L_x + A^T p
//   then two warp groups run concurrently (named barriers):
//     group 1 (4 warps)   phase C  partial-pivot LU of Q_uu -- the pivoting rule of LAPACK dgetf2
//                                  (the reference calls np.linalg.solve = dgesv, control.py:141-142),
//                                  done without row swaps: pivot rows are marked, one barrier per column
//                         phase D  K = -Q_uu^{-1} Q_ux, d = -Q_uu^{-1} Q_u: one thread per right-hand
//                                  side, the solution vector in registers, right-looking substitution
//     group 2 (the rest)  phase B  Q_xx = L_xx + A^T P A, in place on the upper-triangular blocks of P
//   phase E   pq = Q_ux^T d, z = Q_uu d + Q_u, then Y = Q_uu K + 2 Q_ux (in place over Q_ux)
//   phase F   P <- Q_xx + 1/2 (K^T Y + Y^T K)   (== the reference's symmetrised
//             Q_xx + K^T Q_uu K + K^T Q_ux + Q_ux^T K), upper blocks only
//             p <- Q_x + K^T z + Q_ux^T d
//
// The GEMM-shaped phases E and F run on the FP64 tensor path (mma.sync.m8n8k4.f64, "DMMA") when
// the joint sizes are multiples of 8, otherwise on DFMA register tiles.  P, Q_ux/Y, K, Q_uu and the
// LU factors stay in shared memory for the whole recursion (about 215 kB for 10 Quadcopter12D agents
// -> one CTA per SM); only the stage records stream in and K, d stream out.  Problems too large for
// shared memory keep Q_ux/Y and K in an L2-resident global scratch instead (same code).
#include "kernels.cuh"

namespace dpilqr {

template <int S>
struct TileSize {
    static constexpr int value = (S % 4 == 0) ? 4 : (S % 3 == 0) ? 3 : S;
};

constexpr int kSolveThreads = 256;  // warp group 1: LU factorisation; the remaining warps form group 2

__device__ __forceinline__ void named_barrier(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// D(8x8) += A(8x4) * B(4x8) in FP64 on the tensor path.  Lane l holds A[l/4][l%4], B[l%4][l/4] and
// D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Row stride of the LU work matrix: even (16-byte row alignment for double2 access), at least m + 1, and not a
// multiple of 16 doubles so that consecutive rows start in different banks.
__host__ __device__ constexpr int backward_ldw(int m) { return ((m + 2) & ~1) % 16 == 0 ? ((m + 2) & ~1) + 2 : ((m + 2) & ~1); }

// Reciprocal without the special-case branch of __drcp_rn: hardware seed (about 20 bits) plus two Newton steps.
__device__ __forceinline__ double fast_rcp(double v)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(v));
    double e = fma(-v, x, 1.0);
    x = fma(x, e, x);
    e = fma(-v, x, 1.0);
    return fma(x, e, x);
}

// Phase C of the backward kernel: LU factorisation of the m x m matrix W (row-major, W[r*ldw + c]) with partial
// pivoting (the rule of LAPACK dgetf2, which also scales by the reciprocal pivot), by the kSolveThreads threads of
// warp group 1, with LOOK-AHEAD PIVOTING: the pivot search never sits on the elimination's critical path.
//
//   * Row threads (warps 0..6): TPR threads per row, each owning every TPR-th pair of columns.  The pairs live in
//     registers (static indexing) for the whole factorisation and are mirrored to shared memory after every
//     update; only the pivot row is loaded, and only the pairs that still change.  In round k they eliminate
//     column k-1 and then publish the row's entries of columns k and k+1 into a small side buffer.
//   * The search warp (warp 7) works one column ahead on the side buffer alone: in round k it applies the
//     elimination of column k-1 to column k itself (same operands, same operations => the same bits the row threads
//     produce), takes the arg-max of |.| over the unused rows with warp reductions, and publishes the pivot row of
//     column k together with the reciprocal pivot.
//   * One named barrier per round.  Rows never move: a used pivot row is simply marked, its index goes to order[k].
//
// On return W holds the multipliers l(r, k) in the eliminated positions and the rows of U in the pivot rows.
// Straight-line round body (a lone warp per scheduler pays the full branch latency), rounds not unrolled
// (instruction-cache footprint), out of line for a register allocation of its own.  MT > 0 fixes m at compile time.
template <int MT>
__device__ __noinline__ void lu_lookahead(double *__restrict__ W, double *__restrict__ colbuf, double *__restrict__ rinvbuf,
                                          int *__restrict__ prbuf, int *__restrict__ order, int m_rt, int gt)
{
    const int m = MT > 0 ? MT : m_rt;
    const int ldw = backward_ldw(m);
    const int npair = (m + 1) >> 1;
    const int tpr = (m <= 56) ? 4 : 3;                       // threads per row: rows must fit in warps 0..6
    constexpr int NP = MT > 0 ? ((MT + 1) / 2 + (MT <= 56 ? 3 : 2)) / (MT <= 56 ? 4 : 3) : 11;  // pairs per thread
    const int lane = gt & 31;
    // |v| of a double orders like its bit pattern; +1 so that a live zero still beats a used row (key 0)
    auto pivot_key = [](double v) -> unsigned long long {
        const double av = fabs(v);
        return (av == av) ? (unsigned long long)__double_as_longlong(av) + 1ull : 1ull;
    };
    if ((gt >> 5) == 7) {
        // ------------------------------------------------------------------ search warp
        const int r0 = lane, r1 = lane + 32;
        const bool has0 = r0 < m, has1 = r1 < m;
        bool done0 = !has0, done1 = !has1;
        int pr_prev = 0;
        double rinv_prev = 0.0;
#pragma unroll 1
        for (int k = 0; k < m; ++k) {
            double v0, v1;
            if (k == 0) {
                v0 = has0 ? W[r0 * ldw] : 0.0;
                v1 = has1 ? W[r1 * ldw] : 0.0;
            } else {
                const double *cb = colbuf + ((k - 1) & 1) * 128;  // [0..63]: column k-1, [64..127]: column k
                const double pcur = cb[64 + pr_prev];
                const double a0 = has0 ? cb[r0] : 0.0, b0 = has0 ? cb[64 + r0] : 0.0;
                const double a1 = has1 ? cb[r1] : 0.0, b1 = has1 ? cb[64 + r1] : 0.0;
                v0 = fma(-(a0 * rinv_prev), pcur, b0);
                v1 = fma(-(a1 * rinv_prev), pcur, b1);
            }
            const unsigned long long key0 = done0 ? 0ull : pivot_key(v0);
            const unsigned long long key1 = done1 ? 0ull : pivot_key(v1);
            const bool second = key1 > key0;
            const unsigned long long kmax = second ? key1 : key0;
            const int rsel = second ? r1 : r0;
            const double vsel = second ? v1 : v0;
            const unsigned hi = (unsigned)(kmax >> 32);
            const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
            bool mine = (hi == mhi);
            unsigned bal = __ballot_sync(0xffffffffu, mine);
            if (__popc(bal) > 1) {  // rare: several rows share the top 32 bits
                const unsigned lo = mine ? (unsigned)kmax : 0u;
                const unsigned mlo = __reduce_max_sync(0xffffffffu, lo);
                mine = mine && (lo == mlo);
                bal = __ballot_sync(0xffffffffu, mine);
            }
            const int win = __ffs(bal) - 1;
            const int pr = __shfl_sync(0xffffffffu, rsel, win);
            const double pivot = __shfl_sync(0xffffffffu, vsel, win);
            const double rinv = fast_rcp(pivot);
            if (lane == 0) {
                prbuf[k & 1] = pr;
                rinvbuf[k & 1] = rinv;
                order[k] = pr;
            }
            done0 = done0 || (r0 == pr);
            done1 = done1 || (r1 == pr);
            pr_prev = pr;
            rinv_prev = rinv;
            named_barrier(1, kSolveThreads);
        }
        named_barrier(1, kSolveThreads);
        return;
    }
    // ---------------------------------------------------------------------- row threads
    const int r = gt / tpr, q = gt - r * tpr;
    const bool myrow = r < m;
    double *wrow = W + (myrow ? r : m - 1) * ldw;
    bool mydone = !myrow;
    double2 wreg[NP];  // this thread's column pairs of row r
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int j = q + tpr * i;
        wreg[i] = (j < npair) ? *reinterpret_cast<const double2 *>(wrow + 2 * j) : make_double2(0.0, 0.0);
    }
    double held_mult = 0.0;  // multiplier of the previous elimination, stored one barrier later
    bool held = false;
#pragma unroll 1
    for (int k = 0; k < m; ++k) {
        if (k >= 1) {
            const int kk = k - 1;  // column eliminated in this round
            const int pr = prbuf[kk & 1];
            const double rinv = rinvbuf[kk & 1];
            const double *prow = W + pr * ldw;
            // The multiplier of the previous elimination replaces entry (r, kk-1) only now: every thread of the row
            // has read that entry before the barrier that ended the previous round.
            if (held) wrow[kk - 1] = held_mult;
            mydone = mydone || (r == pr);
            const bool live = !mydone;
            const double mult = wrow[kk] * rinv;
            held = live && (q == 0);
            held_mult = mult;
            __syncwarp();  // all threads of the row have read entry (r, kk): the pair loop may overwrite it
            const int jp0 = (kk + 1) >> 1;
            double2 p2[NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const int j = q + tpr * i;
                p2[i] = make_double2(0.0, 0.0);
                if (live && j >= jp0 && j < npair) p2[i] = *reinterpret_cast<const double2 *>(prow + 2 * j);
            }
            // The pair holding column kk+1 may also rewrite the eliminated entry (r, kk) with rounding noise: the
            // multiplier is stored over it in the next round.
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                wreg[i].x = fma(-mult, p2[i].x, wreg[i].x);
                wreg[i].y = fma(-mult, p2[i].y, wreg[i].y);
            }
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const int j = q + tpr * i;
                if (live && j >= jp0 && j < npair) *reinterpret_cast<double2 *>(wrow + 2 * j) = wreg[i];
            }
        }
        // publish this row's entries of columns k and k+1 (state after eliminating the columns < k) for the search
        __syncwarp();
        if (myrow && q == 1) colbuf[(k & 1) * 128 + r] = wrow[k];
        if (myrow && q == 2 && k + 1 < m) colbuf[(k & 1) * 128 + 64 + r] = wrow[k + 1];
        named_barrier(1, kSolveThreads);
    }
    if (held) wrow[m - 2] = held_mult;
    named_barrier(1, kSolveThreads);
}

struct BackwardSmem {
    size_t Pb, QUU, W, Lp, Up, rdiag, sA, sB, sL, pvec, Qx, pq, Qu, dv, zv, order, keys, rinv, tacc, mbar, mats, total_doubles;
};

// Row stride of the Q_ux / K buffers: room for right-hand side n (Q_u) rounded up to a tile of 8, and
// congruent to 8 modulo 16 doubles so that consecutive rows start 16 banks apart (conflict-free fragments).
__host__ __device__ constexpr int backward_ldn(int n)
{
    return (((n + 8) & ~7) & 15) == 8 ? ((n + 8) & ~7) : ((n + 8) & ~7) + 8;
}

__host__ __device__ constexpr size_t even_up(size_t v) { return (v + 1) & ~(size_t)1; }

// doubles of global (L2-resident) scratch per CTA when a team is too large for shared memory
__host__ __device__ constexpr size_t backward_scratch_per_cta(int m, int n)
{
    return 2 * (size_t)m * backward_ldn(n) + even_up((size_t)m * backward_ldw(m)) + 2 * even_up((size_t)m * m);
}

// Shared-memory carve-up in doubles (everything 16-byte aligned).
__host__ __device__ constexpr BackwardSmem backward_smem(int a, int S, int C, bool mats_in_smem)
{
    const int n = a * S, m = a * C, pairs = a * (a - 1) / 2;
    const size_t nblk = (size_t)a * (a + 1) / 2;
    BackwardSmem L{};
    size_t off = 0;
    L.Pb = off;    off += even_up(nblk * (S * S + 2));
    L.QUU = off;   off += even_up((size_t)m * (m + 4));
    // the LU work matrix and the packed factors move to the global scratch together with Q_ux / K for big teams
    L.W = off;     if (mats_in_smem) off += even_up((size_t)m * backward_ldw(m));
    L.Lp = off;    if (mats_in_smem) off += even_up((size_t)m * m);
    L.Up = off;    if (mats_in_smem) off += even_up((size_t)m * m);
    L.rdiag = off; off += even_up(m);
    L.sA = off;    off += even_up((size_t)a * (S * S + 2));
    L.sB = off;    off += even_up((size_t)a * (S * C + 2));
    L.sL = off;    off += even_up((size_t)n + m + 9 * a + 9 * pairs);
    L.pvec = off;  off += even_up(n);
    L.Qx = off;    off += even_up(n);
    L.pq = off;    off += even_up(n);
    L.Qu = off;    off += even_up(m);
    L.dv = off;    off += even_up(m);
    L.zv = off;    off += even_up(m);
    L.order = off; off += even_up(((size_t)m + 1) / 2 + 1);  // m ints
    L.keys = off;  off += 256;                                 // look-ahead side buffer of the LU
    L.rinv = off;  off += 4;                                   // reciprocal pivots + pivot rows of two rounds
    L.tacc = off;  off += 20;                                  // debug cycle counters
    L.mbar = off;  off += 2;                                   // mbarrier of the stage-record bulk copies
    L.mats = off;
    if (mats_in_smem) off += 2 * (size_t)m * backward_ldn(n);
    L.total_doubles = off;
    return L;
}

template <int S, int C, int AT, bool GLOBAL>
__global__ void __launch_bounds__(512, 1) backward_kernel(const BackwardParams p)
{
    extern __shared__ double smem[];
    const Batch &bt = p.batch;
    if (p.n_active != nullptr && (int)blockIdx.x >= *p.n_active) return;
    const int b = p.active ? p.active[blockIdx.x] : blockIdx.x;
    const int a = AT > 0 ? AT : bt.n_agents;
    const int T = bt.horizon;
    const int n = a * S, m = a * C;
    const int pairs = a * (a - 1) / 2;
    const int nblk = a * (a + 1) / 2;
    constexpr int PBS = S * S + 2;  // +2 doubles de-alias the banks of consecutive blocks
    constexpr int SAS = S * S + 2, SBS = S * C + 2;
    constexpr bool USE_MMA = (AT > 0) && (S % 2 == 0) && ((AT * S) % 8 == 0) && ((AT * C) % 8 == 0);
    const int LDQ = m + 4;
    const int LDW = backward_ldw(m);  // row stride of the LU work matrix
    const int LDN = backward_ldn(n);  // row stride of Q_ux / K: column n carries Q_u / d, columns n+1.. are zero padding
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const StageLayout L = stage_layout(a, S, C);
    const BackwardSmem SM = backward_smem(a, S, C, !GLOBAL);

    double *Pb = smem + SM.Pb;        // [nblk][PBS]   upper-triangular blocks of P
    double *QUU = smem + SM.QUU;      // [m][LDQ]
    double *W = smem + SM.W;          // [m][LDW] row-major LU work matrix
    double *Lp = smem + SM.Lp;        // [m][m] packed unit-lower factor, Lp[k*m + k2] = l(k2, k), k2 > k
    double *Up = smem + SM.Up;        // [m][m] packed upper factor, column-major: Up[c*m + k] = u(k, c), k <= c
    double *rdiag = smem + SM.rdiag;  // [m] 1 / u(k, k)
    double *sA = smem + SM.sA;        // [a][SAS]
    double *sB = smem + SM.sB;        // [a][SBS]
    double *sLx = smem + SM.sL;       // [n]
    double *sLu = sLx + n;            // [m]
    double *sHd = sLu + m;            // [a][9]
    double *sHo = sHd + 9 * a;        // [pairs][9]
    double *pvec = smem + SM.pvec, *Qx = smem + SM.Qx, *pq = smem + SM.pq;
    double *Qu = smem + SM.Qu, *dv = smem + SM.dv, *zv = smem + SM.zv;
    int *order = reinterpret_cast<int *>(smem + SM.order);  // [m] physical pivot row of step k
    double *colbuf = smem + SM.keys;   // [2][2][64] look-ahead side buffer of the LU: two columns of every row
    double *rinvbuf = smem + SM.rinv;  // [2] reciprocal pivots, [2..3] pivot rows (as ints)
    int *prbuf = reinterpret_cast<int *>(smem + SM.rinv + 2);
    double *QUX, *KB;  // [m][LDN] each; QUX becomes Y in phase E
    if constexpr (GLOBAL) {
        double *base = p.scratch + (size_t)blockIdx.x * backward_scratch_per_cta(m, n);
        QUX = base;
        KB = QUX + (size_t)m * LDN;
        W = KB + (size_t)m * LDN;
        Lp = W + even_up((size_t)m * LDW);
        Up = Lp + even_up((size_t)m * m);
    } else {
        QUX = smem + SM.mats;
        KB = QUX + (size_t)m * LDN;
    }

    const int32_t *cidx_b = bt.cost_idx + (int64_t)b * a;
    const double w_ref = bt.weights ? bt.weights[2 * b] : 1.0;
    const double mu = p.mu[b];
    const double *stage_b = p.stage + (int64_t)b * (T + 1) * L.stride;
    double *Kb = p.K + (int64_t)b * T * m * n;
    double *db = p.d + (int64_t)b * T * m;
    int st = 0;
    // optional per-phase cycle counters (debug aid, see dpilqr_debug_backward_timing); kept in shared memory
    long long *tacc = reinterpret_cast<long long *>(smem + SM.tacc) + (tid == 0 ? 0 : 12);
    long long tmark = 0;
    const bool timing = (p.timing != nullptr) && (blockIdx.x == 0) && (tid == 0 || tid == kSolveThreads);
    if (timing) {
        for (int k = 0; k < (tid == 0 ? 12 : 8); ++k) tacc[k] = 0;
    }
    auto tick = [&](int slot) {
        if (timing) {
            const long long now = clock64();
            tacc[slot] += now - tmark;
            tmark = now;
        }
    };

    auto blk_index = [a](int i, int j) { return i * a - (i * (i - 1)) / 2 + (j - i); };  // i <= j
    // Asynchronous copy of one stage record into its (bank-padded) shared-memory home.  With even block sizes one
    // elected thread issues a handful of TMA bulk copies (cp.async.bulk) that complete on an mbarrier; otherwise
    // every thread issues 8-byte LDGSTS copies.
    constexpr bool BULK = ((S * S) % 2 == 0) && ((S * C) % 2 == 0);
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem + SM.mbar);
    int record_phase = 0;
    auto cp_async8 = [](double *dst, const double *src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    auto bulk_copy = [&](double *dst, const double *src, unsigned bytes) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"(mbar) : "memory");
    };
    auto prefetch_record = [&](int t) {  // call after a __syncthreads(): nobody reads the previous record any more
        const double *rec = stage_b + (int64_t)t * L.stride;
        if constexpr (BULK) {
            if (tid == 0) {
                const unsigned tail = (unsigned)(((n + m + 9 * a + 9 * pairs + 1) & ~1) * 8);
                const unsigned total = (unsigned)(a * (S * S + S * C) * 8) + tail;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(total) : "memory");
                for (int i = 0; i < a; ++i) bulk_copy(sA + i * SAS, rec + L.offA + i * S * S, S * S * 8);
                for (int i = 0; i < a; ++i) bulk_copy(sB + i * SBS, rec + L.offB + i * S * C, S * C * 8);
                bulk_copy(sLx, rec + L.offLx, tail);
            }
        } else {
            for (int k = tid; k < a * S * S; k += nthr) cp_async8(sA + (k / (S * S)) * SAS + k % (S * S), rec + L.offA + k);
            for (int k = tid; k < a * S * C; k += nthr) cp_async8(sB + (k / (S * C)) * SBS + k % (S * C), rec + L.offB + k);
            for (int k = tid; k < n + m + 9 * a + 9 * pairs; k += nthr) cp_async8(sLx + k, rec + L.offLx + k);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
    auto wait_record = [&] {
        if constexpr (BULK) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "WAIT_RECORD:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra DONE_RECORD;\n"
                "bra WAIT_RECORD;\n"
                "DONE_RECORD:\n"
                "}\n" ::"r"(mbar), "r"(record_phase & 1) : "memory");
            ++record_phase;
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
    };
    if constexpr (BULK) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    // |v| of a double orders like its bit pattern; +1 so that a live zero still beats a used row (key 0)
    auto pivot_key = [](double v) -> unsigned long long {
        const double av = fabs(v);
        return (av == av) ? (unsigned long long)__double_as_longlong(av) + 1ull : 1ull;
    };

    // ---- terminal condition: p = L_x, P = L_xx at (X[T], u = 0)  (control.py:125-129)
    prefetch_record(T);
    for (int k = tid; k < m * (LDN - n); k += nthr) {  // zero the padding columns once
        const int row = k / (LDN - n), e = k - row * (LDN - n);
        QUX[(size_t)row * LDN + n + e] = 0.0;
        KB[(size_t)row * LDN + n + e] = 0.0;
    }
    wait_record();
    __syncthreads();
    for (int k = tid; k < nblk * S * S; k += nthr) {
        const int blk = k / (S * S), e = k - blk * (S * S);
        const int r = e / S, cc = e - r * S;
        int i = 0, rem = blk;
        while (rem >= a - i) { rem -= a - i; ++i; }
        const int j = i + rem;
        double v = 0.0;
        if (i == j) {
            const double *Qf = bt.Qf + (int64_t)cidx_b[i] * S * S;
            v = w_ref * (Qf[r * S + cc] + Qf[cc * S + r]);
            if (r < 3 && cc < 3) v += sHd[9 * i + r * 3 + cc];
        } else if (r < 3 && cc < 3) {
            v = sHo[9 * pair_index(i, j, a) + r * 3 + cc];
        }
        Pb[(size_t)blk * PBS + e] = v;
    }
    for (int k = tid; k < n; k += nthr) pvec[k] = sLx[k];
    __syncthreads();
    prefetch_record(T - 1);

    if (timing) tmark = clock64();
#pragma unroll 1
    for (int t = T - 1; t >= 0; --t) {
        wait_record();
        __syncthreads();
        tick(0);

        // ---- phase A: Q_ux, Q_uu (S = B^T (P + mu I) in registers), Q_u, Q_x
        for (int it = tid; it < a * a * C; it += nthr) {
            const int g = it % C;
            const int j = (it / C) % a;
            const int i = it / (C * a);
            const double *Bi = sB + i * SBS;
            const bool upper = (i <= j);
            const double *Pblk = Pb + (size_t)(upper ? blk_index(i, j) : blk_index(j, i)) * PBS;
            const int rs = upper ? S : 1, cs = upper ? 1 : S;  // P_ij = (P_ji)^T below the diagonal
            double Srow[S];
#pragma unroll
            for (int sg = 0; sg < S; ++sg) Srow[sg] = 0.0;
#pragma unroll
            for (int r = 0; r < S; ++r) {
                const double bv = Bi[r * C + g];
#pragma unroll
                for (int sg = 0; sg < S; ++sg) {
                    double pv = Pblk[r * rs + sg * cs];
                    if (i == j && r == sg) pv += mu;
                    Srow[sg] = fma(bv, pv, Srow[sg]);
                }
            }
            const double *Aj = sA + j * SAS;
            const double *Bj = sB + j * SBS;
            const int row = i * C + g;
#pragma unroll
            for (int sg2 = 0; sg2 < S; ++sg2) {
                double acc = 0.0;
#pragma unroll
                for (int sg = 0; sg < S; ++sg) acc = fma(Srow[sg], Aj[sg * S + sg2], acc);
                QUX[(size_t)row * LDN + j * S + sg2] = acc;  // L_ux == 0 (cost.py:91)
            }
#pragma unroll
            for (int g2 = 0; g2 < C; ++g2) {
                double acc = 0.0;
#pragma unroll
                for (int sg = 0; sg < S; ++sg) acc = fma(Srow[sg], Bj[sg * C + g2], acc);
                if (i == j) {
                    const double *R = bt.R + (int64_t)cidx_b[i] * C * C;
                    acc += w_ref * (R[g * C + g2] + R[g2 * C + g]);
                }
                QUU[row * LDQ + j * C + g2] = acc;
                W[row * LDW + j * C + g2] = acc;
            }
            if (j == 0) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < S; ++r) acc = fma(Bi[r * C + g], pvec[i * S + r], acc);
                const double qu = sLu[row] + acc;
                Qu[row] = qu;
                QUX[(size_t)row * LDN + n] = qu;  // Q_u rides along as right-hand side n
            }
        }
        for (int col = tid; col < n; col += nthr) {
            const int j = col / S, sg = col - j * S;
            const double *Aj = sA + j * SAS;
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < S; ++r) acc = fma(Aj[r * S + sg], pvec[j * S + r], acc);
            Qx[col] = sLx[col] + acc;
        }
        __syncthreads();
        tick(1);

        if (tid < kSolveThreads) {
            // ================= group 1: phase C, LU with implicit partial pivoting =================
            const int gt = tid;
            lu_lookahead<(AT > 0 ? AT * C : 0)>(W, colbuf, rinvbuf, prbuf, order, m, gt);
            tick(2);
            // pack the factors in pivot order so the substitutions read contiguous memory
            for (int e = gt; e < m * m; e += kSolveThreads) {
                const int k = e / m, x2 = e - k * m;
                const double v = W[order[x2] * LDW + k];
                if (x2 > k) Lp[k * m + x2] = v;   // l(x2, k)
                else Up[k * m + x2] = v;          // u(x2, k): column k, row x2 <= k
                if (x2 == k) {
                    if (!(fabs(v) > 0.0)) st |= DPILQR_ST_SINGULAR;  // exact zero (or NaN) pivot
                    rdiag[k] = __drcp_rn(v);
                }
            }
            if constexpr (USE_MMA) {
                // The blocked forward solve of phase D applies the 8x8 diagonal blocks of the unit lower factor (well
                // conditioned: |l| <= 1) as explicit inverses on the tensor path: invert them here, in place.  One
                // thread per (block, column of the inverse).
                constexpr int M = AT * C, NB = M / 8;
                named_barrier(1, kSolveThreads);
                const int which = gt / (NB * 8), bb = (gt / 8) % NB, j = gt & 7;  // which: 0 = L, 1 = U
                double x[8];
                const bool busy = gt < NB * 8;
                if (busy) {
                    const double *F = (which ? Up : Lp) + (8 * bb) * M + 8 * bb;  // F[c * M + rr] = f(rr, c) of this block
                    if (which == 0) {  // unit lower: solve L x = e_j by forward substitution
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
                            for (int c = 0; c < i; ++c) acc = fma(-F[c * M + i], x[c], acc);
                            x[i] = acc;
                        }
                    } else {  // upper: solve U x = e_j by backward substitution
#pragma unroll
                        for (int i = 7; i >= 0; --i) {
                            double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
                            for (int c = i + 1; c < 8; ++c) acc = fma(-F[c * M + i], x[c], acc);
                            x[i] = acc * rdiag[8 * bb + i];
                        }
                    }
                }
                named_barrier(1, kSolveThreads);
                if (busy) {
                    double *F = which ? Up : Lp;
#pragma unroll
                    for (int i = 0; i < 8; ++i) F[(8 * bb + j) * M + 8 * bb + i] = x[i];  // inverse(i, j), same transposed layout
                }
            }
            tick(3);
        } else {
            // ================= group 2: phase B, Q_xx = L_xx + A^T P A in place (upper blocks) =================
            const int gt = tid - kSolveThreads, gn = nthr - kSolveThreads;
            const int blocks_per_round = gn / S;
            for (int blk0 = (p.debug_mode & 16) ? nblk : 0; blk0 < nblk; blk0 += blocks_per_round) {
                const int blk = blk0 + gt / S;
                const int sg = gt % S;
                const bool live = (gt < blocks_per_round * S) && (blk < nblk);
                double out[S];
                if (live) {
                    int i = 0, rem = blk;
                    while (rem >= a - i) { rem -= a - i; ++i; }
                    const int j = i + rem;
                    const double *Pblk = Pb + (size_t)blk * PBS;
                    const double *Ai = sA + i * SAS, *Aj = sA + j * SAS;
                    double v[S];
#pragma unroll
                    for (int r = 0; r < S; ++r) {
                        double acc = 0.0;
#pragma unroll
                        for (int q = 0; q < S; ++q) acc = fma(Pblk[r * S + q], Aj[q * S + sg], acc);
                        v[r] = acc;
                    }
#pragma unroll
                    for (int r = 0; r < S; ++r) {
                        double acc = 0.0;
#pragma unroll
                        for (int q = 0; q < S; ++q) acc = fma(Ai[q * S + r], v[q], acc);
                        double lxx = 0.0;
                        if (i == j) {
                            const double *Q = bt.Q + (int64_t)cidx_b[i] * S * S;
                            lxx = w_ref * (Q[r * S + sg] + Q[sg * S + r]);
                            if (r < 3 && sg < 3) lxx += sHd[9 * i + r * 3 + sg];
                        } else if (r < 3 && sg < 3) {
                            lxx = sHo[9 * pair_index(i, j, a) + r * 3 + sg];
                        }
                        out[r] = lxx + acc;
                    }
                }
                named_barrier(2, gn);
                if (live) {
                    double *Pblk = Pb + (size_t)blk * PBS;
#pragma unroll
                    for (int r = 0; r < S; ++r) Pblk[r * S + sg] = out[r];
                }
            }
            tick(2);
        }
        __syncthreads();
        tick(5);
        // the stage record of this step is dead from here on: fetch the next one behind phases D, E and F
        if (t > 0) prefetch_record(t - 1);
        tick(11);

        // ---- phase D: K = -Q_uu^{-1} Q_ux, d = -Q_uu^{-1} Q_u.  Right-hand sides 0..n-1 are the columns of Q_ux,
        // right-hand side n is Q_u.  The negated, row-permuted right-hand sides are substituted in place in KB.
        double *Kt = Kb + (int64_t)t * m * n;
        if constexpr (USE_MMA) {
            constexpr int M = AT * C, NB = M / 8;
            // Blocked triangular solves: each warp owns tiles of 8 right-hand sides and needs no other warp.  Every
            // 8x8 block operation -- X_b <- inv(F_bb) X_b on the diagonal, X_b2 -= F(b2, b) X_b off it -- is two
            // m8n8k4 FP64 tensor instructions; the off-diagonal updates of one level are independent and interleave.
            const int fr = lane >> 2, fc = lane & 3;
            for (int nt = warp; nt < (n + 8) / 8; nt += nwarp) {
                double *Xc = KB + 8 * nt;
                // X = -P Q_ux for this warp's 8 right-hand sides: rows in pivot order, negated (lane = 4 rows x 8 cols)
#pragma unroll
                for (int k0 = 0; k0 < M; k0 += 4) {
                    const int k = k0 + (lane >> 3), cc = lane & 7;
                    Xc[(size_t)k * LDN + cc] = -QUX[(size_t)order[k] * LDN + 8 * nt + cc];
                }
                __syncwarp();
#pragma unroll
                for (int dir = 0; dir < 2; ++dir) {  // 0: forward with the unit lower factor, 1: backward with the upper
                    const double *F = dir ? Up : Lp;
#pragma unroll
                    for (int lvl = 0; lvl < NB; ++lvl) {
                        const int blk = dir ? NB - 1 - lvl : lvl;
                        if (dir == 1) {
                            // Upper factor: the diagonal block is substituted through by lanes 0..7 (one right-hand
                            // side each).  Its explicit inverse would be one more tensor instruction pair, but U carries
                            // the conditioning of Q_uu and the inverse costs about a digit of accuracy in K.
                            if (lane < 8) {
                                double x[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) x[j] = Xc[(size_t)(8 * blk + j) * LDN + lane];
#pragma unroll
                                for (int j = 7; j >= 0; --j) {
                                    x[j] *= rdiag[8 * blk + j];
#pragma unroll
                                    for (int j2 = 0; j2 < j; ++j2) x[j2] = fma(-F[(8 * blk + j) * M + 8 * blk + j2], x[j], x[j2]);
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j) Xc[(size_t)(8 * blk + j) * LDN + lane] = x[j];
                            }
                            __syncwarp();
                        } else {
                        double d0 = 0.0, d1 = 0.0;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            dmma_m8n8k4(d0, d1, F[(8 * blk + 4 * ks + fc) * M + 8 * blk + fr],
                                        Xc[(size_t)(8 * blk + 4 * ks + fc) * LDN + fr]);
                        __syncwarp();
                        *reinterpret_cast<double2 *>(Xc + (size_t)(8 * blk + fr) * LDN + 2 * fc) = make_double2(d0, d1);
                        __syncwarp();
                        }
#pragma unroll
                        for (int o = 1; o < NB; ++o) {
                            const int b2 = dir ? blk - o : blk + o;
                            if (b2 >= 0 && b2 < NB) {
                                double *cp = Xc + (size_t)(8 * b2 + fr) * LDN + 2 * fc;
                                double2 c2 = *reinterpret_cast<double2 *>(cp);
#pragma unroll
                                for (int ks = 0; ks < 2; ++ks)
                                    dmma_m8n8k4(c2.x, c2.y, -F[(8 * blk + 4 * ks + fc) * M + 8 * b2 + fr],
                                                Xc[(size_t)(8 * blk + 4 * ks + fc) * LDN + fr]);
                                *reinterpret_cast<double2 *>(cp) = c2;
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        } else {
            for (int col = tid; col <= n; col += nthr) {  // runtime sizes: one thread per right-hand side
                double *xcol = KB + col;
                for (int k = 0; k < m; ++k) xcol[(size_t)k * LDN] = -QUX[(size_t)order[k] * LDN + col];
                for (int k = 0; k < m - 1; ++k) {
                    const double xk = xcol[(size_t)k * LDN];
                    for (int k2 = k + 1; k2 < m; ++k2) xcol[(size_t)k2 * LDN] = fma(-Lp[k * m + k2], xk, xcol[(size_t)k2 * LDN]);
                }
                for (int cc = m - 1; cc >= 0; --cc) {
                    const double xc = xcol[(size_t)cc * LDN] * rdiag[cc];
                    xcol[(size_t)cc * LDN] = xc;
                    for (int k = 0; k < cc; ++k) xcol[(size_t)k * LDN] = fma(-Up[cc * m + k], xc, xcol[(size_t)k * LDN]);
                }
            }
        }
        __syncthreads();
        tick(10);
        for (int e = tid; e < m * n; e += nthr) {  // stream K[t] out, coalesced
            const int k = e / n, col = e - k * n;
            const double kv = KB[(size_t)k * LDN + col];
            if (!isfinite(kv)) st |= DPILQR_ST_NONFINITE;
            Kt[e] = kv;
        }
        for (int k = tid; k < m; k += nthr) {
            const double dk = KB[(size_t)k * LDN + n];
            dv[k] = dk;
            db[(int64_t)t * m + k] = dk;
        }
        __syncthreads();
        tick(4);

        // ---- phase E: pq = Q_ux^T d and z = Q_uu d + Q_u (before Q_ux is overwritten), then Y = Q_uu K + 2 Q_ux
        for (int col = tid; col < n + m; col += nthr) {
            if (col < n) {
                double qd = 0.0;
                for (int k = 0; k < m; ++k) qd = fma(QUX[(size_t)k * LDN + col], dv[k], qd);
                pq[col] = qd;
            } else {
                const int k = col - n;
                double acc = 0.0;
                for (int l = 0; l < m; ++l) acc = fma(QUU[k * LDQ + l], dv[l], acc);
                zv[k] = acc + Qu[k];
            }
        }
        __syncthreads();
        tick(6);
        if constexpr (USE_MMA) {
            constexpr int N = AT * S, M = AT * C, LD = backward_ldn(N);
            constexpr int MT = M / 8, NT = N / 8, KS = M / 4;
            const int fr = lane >> 2, fc = lane & 3;  // fragment coordinates
            for (int tile = warp; tile < MT * NT; tile += nwarp) {
                const int mt = tile / NT, nt = tile - mt * NT;
                double *yp = QUX + (size_t)(8 * mt + fr) * LD + 8 * nt + 2 * fc;
                double c0 = 2.0 * yp[0], c1 = 2.0 * yp[1];
                const double *ap = QUU + (8 * mt + fr) * LDQ + fc;
                const double *bp = KB + (size_t)fc * LD + 8 * nt + fr;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) dmma_m8n8k4(c0, c1, ap[4 * ks], bp[(size_t)4 * ks * LD]);
                yp[0] = c0;
                yp[1] = c1;
            }
        } else {
            for (int col = tid; col < n; col += nthr) {
                for (int k = 0; k < m; ++k) {
                    double acc = 0.0;
                    for (int l = 0; l < m; ++l) acc = fma(QUU[k * LDQ + l], KB[(size_t)l * LDN + col], acc);
                    QUX[(size_t)k * LDN + col] = acc + 2.0 * QUX[(size_t)k * LDN + col];
                }
            }
        }
        __syncthreads();
        tick(7);

        // ---- phase F: P <- Q_xx + 1/2 (K^T Y + Y^T K) on upper blocks; p <- Q_x + K^T z + Q_ux^T d
        const double *Y = QUX;
        if constexpr (USE_MMA) {
            constexpr int N = AT * S, M = AT * C, LD = backward_ldn(N);
            constexpr int NT = N / 8, KS = M / 4;
            constexpr int NTILES = NT * (NT + 1) / 2;
            const int fr = lane >> 2, fc = lane & 3;
            const int q0 = (NTILES * warp) / nwarp, q1 = (NTILES * (warp + 1)) / nwarp;
            int ti = 0, rowstart = 0;  // decode q0 -> (ti, tj) in the row-major upper triangle of tiles
            while (q0 >= rowstart + (NT - ti)) { rowstart += NT - ti; ++ti; }
            int tj = ti + (q0 - rowstart);
            double ak[KS], ay[KS];
            int loaded = -1;
            for (int q = q0; q < q1; ++q) {
                if (loaded != ti) {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        ak[ks] = KB[(size_t)(4 * ks + fc) * LD + 8 * ti + fr];
                        ay[ks] = Y[(size_t)(4 * ks + fc) * LD + 8 * ti + fr];
                    }
                    loaded = ti;
                }
                double c0 = 0.0, c1 = 0.0;
                const double *kp = KB + (size_t)fc * LD + 8 * tj + fr;
                const double *yp = Y + (size_t)fc * LD + 8 * tj + fr;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    dmma_m8n8k4(c0, c1, ak[ks], yp[(size_t)4 * ks * LD]);  // K^T Y
                    dmma_m8n8k4(c0, c1, ay[ks], kp[(size_t)4 * ks * LD]);  // Y^T K
                }
                const int row = 8 * ti + fr, col = 8 * tj + 2 * fc;
                const int bi = row / S, bj = col / S;  // S is even here, so the pair (col, col+1) shares a block
                if (bi <= bj) {
                    double *pblk = Pb + (size_t)blk_index(bi, bj) * PBS;
                    const int rr = row - bi * S, cc = col - bj * S;
                    pblk[rr * S + cc] += 0.5 * c0;
                    pblk[rr * S + cc + 1] += 0.5 * c1;
                    // diagonal blocks are stored in full: an off-diagonal tile inside one also owns the mirrored
                    // entries (their own tile lies below the tile diagonal and is never visited)
                    if (bi == bj && ti != tj) {
                        pblk[cc * S + rr] += 0.5 * c0;
                        pblk[(cc + 1) * S + rr] += 0.5 * c1;
                    }
                }
                if (++tj == NT) { ++ti; tj = ti; }
            }
        } else {
            constexpr int TS = TileSize<S>::value;
            constexpr int TPB = (S / TS) * (S / TS);  // tiles per block
            for (int tile = tid; tile < nblk * TPB; tile += nthr) {
                const int blk = tile / TPB, tt = tile - blk * TPB;
                const int tr = tt / (S / TS), tc = tt - tr * (S / TS);
                int i = 0, rem = blk;
                while (rem >= a - i) { rem -= a - i; ++i; }
                const int j = i + rem;
                const int r0 = i * S + tr * TS, c0 = j * S + tc * TS;
                double acc[TS][TS];
#pragma unroll
                for (int r = 0; r < TS; ++r)
#pragma unroll
                    for (int cc = 0; cc < TS; ++cc) acc[r][cc] = 0.0;
                for (int k = 0; k < m; ++k) {
                    double ki[TS], yi[TS], kj[TS], yj[TS];
#pragma unroll
                    for (int r = 0; r < TS; ++r) {
                        ki[r] = KB[(size_t)k * LDN + r0 + r];
                        yi[r] = Y[(size_t)k * LDN + r0 + r];
                        kj[r] = KB[(size_t)k * LDN + c0 + r];
                        yj[r] = Y[(size_t)k * LDN + c0 + r];
                    }
#pragma unroll
                    for (int r = 0; r < TS; ++r)
#pragma unroll
                        for (int cc = 0; cc < TS; ++cc) acc[r][cc] = fma(ki[r], yj[cc], fma(yi[r], kj[cc], acc[r][cc]));
                }
                double *Pblk = Pb + (size_t)blk * PBS;
#pragma unroll
                for (int r = 0; r < TS; ++r)
#pragma unroll
                    for (int cc = 0; cc < TS; ++cc) {
                        const int e = (tr * TS + r) * S + tc * TS + cc;
                        Pblk[e] = Pblk[e] + 0.5 * acc[r][cc];
                    }
            }
        }
        for (int col = tid; col < n; col += nthr) {
            double acc = 0.0;
            for (int k = 0; k < m; ++k) acc = fma(KB[(size_t)k * LDN + col], zv[k], acc);
            pvec[col] = Qx[col] + acc + pq[col];
        }
        __syncthreads();
        tick(8);
    }
    if (timing) {
        for (int k = 0; k < (tid == 0 ? 12 : 8); ++k) p.timing[(tid == 0 ? 0 : 12) + k] = tacc[k];
    }
    if (st != 0 && p.status) atomicOr(p.status + b, st);
}

struct BackwardPlan {
    size_t smem_bytes;
    int threads;
    int use_global_scratch;
};

static BackwardPlan plan_backward(int a, int s, int c)
{
    BackwardPlan plan;
    const size_t with_mats = backward_smem(a, s, c, true).total_doubles * 8;
    plan.use_global_scratch = (with_mats > 227 * 1024) ? 1 : 0;
    plan.smem_bytes = plan.use_global_scratch ? backward_smem(a, s, c, false).total_doubles * 8 : with_mats;
    plan.threads = 512;
    return plan;
}

int64_t backward_scratch_doubles(int n_problems, int a, int s, int c)
{
    const BackwardPlan plan = plan_backward(a, s, c);
    return plan.use_global_scratch ? (int64_t)n_problems * (int64_t)backward_scratch_per_cta(a * c, a * s) : 0;
}

template <int S, int C, int AT, bool GLOBAL>
static int launch_typed(const BackwardParams &p, int n_blocks, const BackwardPlan &plan, cudaStream_t stream)
{
    auto kernel = backward_kernel<S, C, AT, GLOBAL>;
    static bool attr_set = false;
    if (!attr_set) {
        DPILQR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    kernel<<<n_blocks, plan.threads, plan.smem_bytes, stream>>>(p);
    DPILQR_CUDA(cudaGetLastError());
    return DPILQR_OK;
}

template <int S, int C>
static int launch_generic(const BackwardParams &p, int n_blocks, const BackwardPlan &plan, cudaStream_t stream)
{
    if (plan.use_global_scratch) return launch_typed<S, C, 0, true>(p, n_blocks, plan, stream);
    return launch_typed<S, C, 0, false>(p, n_blocks, plan, stream);
}

int g_backward_debug_mode = 0;
long long *g_backward_timing = nullptr;  // device buffer of 20 counters, set by dpilqr_debug_backward_timing

int launch_backward(const BackwardParams &p_in, int n_blocks, cudaStream_t stream)
{
    if (n_blocks <= 0) return DPILQR_OK;
    BackwardParams p = p_in;
    p.timing = g_backward_timing;
    p.debug_mode = g_backward_debug_mode;
    const Batch &bt = p.batch;
    const int a = bt.n_agents, s = bt.s, c = bt.c;
    if (a * c > 64) {
        set_error("backward kernel: at most 64 joint controls are supported (got %d)", a * c);
        return DPILQR_E_UNSUPPORTED;
    }
    const BackwardPlan plan = plan_backward(a, s, c);
    if (plan.smem_bytes > 227 * 1024) {
        set_error("backward kernel: %d agents x (%d,%d) needs %zu bytes of shared memory (max 232448)", a, s, c, plan.smem_bytes);
        return DPILQR_E_UNSUPPORTED;
    }
    p.use_global_scratch = plan.use_global_scratch;
    if (plan.use_global_scratch && p.scratch == nullptr) {
        set_error("backward kernel: global scratch required for this problem size but none given");
        return DPILQR_E_INVALID;
    }
    if (s == 12 && c == 4) {
        if (a == 10 && !plan.use_global_scratch) return launch_typed<12, 4, 10, false>(p, n_blocks, plan, stream);
        return launch_generic<12, 4>(p, n_blocks, plan, stream);
    }
    if (s == 6 && c == 3) return launch_generic<6, 3>(p, n_blocks, plan, stream);
    if (s == 4 && c == 2) return launch_generic<4, 2>(p, n_blocks, plan, stream);
    if (s == 3 && c == 2) return launch_generic<3, 2>(p, n_blocks, plan, stream);
    if (s == 5 && c == 2) return launch_generic<5, 2>(p, n_blocks, plan, stream);
    set_error("backward kernel: unsupported per-agent dimensions (%d, %d)", s, c);
    return DPILQR_E_UNSUPPORTED;
}

}  // namespace dpilqr
