// lu.cuh -- device routines of the backward kernel's phase C: partial-pivot LU of Q_uu by one warp group
// (reference: np.linalg.solve in ilqrSolver._backward_pass, control.py:141-142).  Shared with tools/lu_bench.cu.
#pragma once
#include <cuda_runtime.h>

namespace dpilqr {

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
    // A row's threads synchronise with __syncwarp(): rows never straddle a warp (with three threads per row a warp
    // takes ten rows and its last two lanes idle)
    const int rows_per_warp = 32 / tpr;
    const int r = (gt >> 5) * rows_per_warp + lane / tpr, q = lane % tpr;
    const bool myrow = r < m && lane < rows_per_warp * tpr;
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

// Blocked form of the same factorisation for m a multiple of 8 (the typed tensor-path kernels): panels of eight
// columns, the same pivoting rule, the same in-place result layout as lu_lookahead.
//
//   * Warp 0 factorises a panel entirely in registers: lane l holds the eight panel entries of row l (and of row
//     l + 32).  Per column: arg-max of |.| over the unused rows with warp reductions (the reciprocal of every
//     candidate is computed speculatively beside the reduction), the pivot row's remaining entries broadcast by
//     shuffles, one multiply and at most seven FMAs per row.  No shared memory and no barrier inside a panel.
//   * U12: one lane per trailing column runs the 8-step forward substitution with the panel's unit-lower block on
//     the pivot rows (28 FMAs, a dependency chain of 7).
//   * The trailing update A22 -= L21 U12 runs on the FP64 tensor path, 8x8 tiles over all rows with the multipliers
//     of used rows masked to zero (rows never move).  The column tile of the NEXT panel is updated first so that
//     warp 0 factorises panel p+1 while the other warps finish the update of panel p (look-ahead).
// NT threads (gt = 0..NT-1) call this together; they synchronise on named barrier 1.
#ifndef DPILQR_LU_LOOKAHEAD_FIRST
#define DPILQR_LU_LOOKAHEAD_FIRST 1
#endif
constexpr bool kLookaheadFirst = DPILQR_LU_LOOKAHEAD_FIRST != 0;

template <int M, int NT, bool TIMED = false>
__device__ __forceinline__ void lu_blocked(double *__restrict__ W, int *__restrict__ order, unsigned *__restrict__ donebuf, int gt,
                                        long long *__restrict__ lt, long long *__restrict__ gl = nullptr)
{
    static_assert(M % 8 == 0 && M <= 64, "blocked LU: m must be a multiple of 8, at most 64");
    constexpr int LDW = backward_ldw(M);
    constexpr int NP = M / 8;
    constexpr int NW = NT / 32;  // warps of the group: warp 0 factorises the panels, the others update
    static_assert(NW >= 2, "blocked LU: at least two warps");
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = gt & 31, warp = gt >> 5;
    const int r0 = lane, r1 = lane + 32;
    const bool has0 = r0 < M, has1 = r1 < M;
    bool done0 = !has0, done1 = !has1;
    // Search key of a row: the top 24 bits of |v| (exponent and 13 mantissa bits) above the row tag 64 | (63 - row).
    // One 32-bit warp maximum then yields pivot magnitude and pivot row at once; entries that agree to 2^-13
    // relative count as tied and the lowest row wins (|l| <= 1 + 2^-13: the stability bound of partial pivoting
    // is unchanged).  A used row has key 0, below every live row (tag bit 64).
    const unsigned tag0 = 64u | (unsigned)(63 - r0), tag1 = 64u | (unsigned)(63 - r1);
    // instrumented build: lt is non-null for every lane of warp 0 (no divergence inside the warp), lane 0 records
    long long tm = (TIMED && lt) ? clock64() : 0;
    auto lap = [&](int slot) {
        if (TIMED && lt) {
            const long long now = clock64();
            if (lane == 0) lt[slot] += now - tm;
            tm = now;
        }
    };
    // finer laps of warp 0 between the panels, into a global buffer (instrumented build): 0 barrier, 1 U12, 2 tile update
    long long tg = 0;
    auto glap = [&](int slot) {
        if (TIMED && gl) {
            const long long now = clock64();
            if (lane == 0 && slot >= 0) atomicAdd(reinterpret_cast<unsigned long long *>(gl) + slot, (unsigned long long)(now - tg));
            tg = now;
        }
    };
#pragma unroll 1
    for (int p = 0; p < NP; ++p) {
        const int c0 = 8 * p;
        if (warp == 0) {
            double x[8], y[8];
            {
                const double2 *s0 = reinterpret_cast<const double2 *>(W + (has0 ? r0 : 0) * LDW + c0);
                const double2 *s1 = reinterpret_cast<const double2 *>(W + (has1 ? r1 : 0) * LDW + c0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double2 v0 = s0[q], v1 = s1[q];
                    x[2 * q] = v0.x, x[2 * q + 1] = v0.y;
                    y[2 * q] = v1.x, y[2 * q + 1] = v1.y;
                }
            }
            // The column loop is software-pipelined by hand (the compiler keeps the source order inside the register-
            // starved recursion loop): the pivot search of column j+1 -- key, speculative reciprocal, warp maximum --
            // is issued as soon as that one column has been updated, and the rest of the rank-1 update runs in its
            // shadow; the pivot-row shuffles of a column are issued as one batch.
            unsigned kmax;
            double rinv_mine;
            auto search = [&](int j) {  // |x[j]|, |y[j]| of the unused rows -> warp maximum of the keys
                const unsigned key0 = done0 ? 0u : (((unsigned)__double2hiint(x[j]) & 0x7fffff80u) | tag0);
                const unsigned key1 = done1 ? 0u : (((unsigned)__double2hiint(y[j]) & 0x7fffff80u) | tag1);
                rinv_mine = fast_rcp(key1 > key0 ? y[j] : x[j]);  // speculative: only the winner's is used
                kmax = __reduce_max_sync(FULL, max(key0, key1));
            };
            search(0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int pr = 63 - (int)(kmax & 63u);  // warp-uniform
                const int win = pr & 31;
                const bool second = pr >= 32;
                const double rinv = __shfl_sync(FULL, rinv_mine, win);
                double pv[8];
#pragma unroll
                for (int jj = j + 1; jj < 8; ++jj) pv[jj] = __shfl_sync(FULL, second ? y[jj] : x[jj], win);
                if (lane == 0) order[c0 + j] = pr;
                done0 = done0 || (pr == r0);
                done1 = done1 || (pr == r1);
                // multipliers: zero for used rows, so the updates below need no predicate
                const double m0 = done0 ? 0.0 : x[j] * rinv, m1 = done1 ? 0.0 : y[j] * rinv;
                if (j + 1 < 8) {
                    x[j + 1] = fma(-m0, pv[j + 1], x[j + 1]);
                    y[j + 1] = fma(-m1, pv[j + 1], y[j + 1]);
                    search(j + 1);
                }
#pragma unroll
                for (int jj = j + 2; jj < 8; ++jj) {
                    x[jj] = fma(-m0, pv[jj], x[jj]);
                    y[jj] = fma(-m1, pv[jj], y[jj]);
                }
                if (!done0) x[j] = m0;
                if (!done1) y[j] = m1;
            }
            if (has0) {
                double2 *d0 = reinterpret_cast<double2 *>(W + r0 * LDW + c0);
#pragma unroll
                for (int q = 0; q < 4; ++q) d0[q] = make_double2(x[2 * q], x[2 * q + 1]);
            }
            if (has1) {
                double2 *d1 = reinterpret_cast<double2 *>(W + r1 * LDW + c0);
#pragma unroll
                for (int q = 0; q < 4; ++q) d1[q] = make_double2(y[2 * q], y[2 * q + 1]);
            }
            const unsigned b0 = __ballot_sync(FULL, done0), b1 = __ballot_sync(FULL, done1);
            if (lane == 0) donebuf[2 * (p & 1)] = b0, donebuf[2 * (p & 1) + 1] = b1;
        }
        lap(p == 0 ? 0 : 2);
        glap(-1);
        named_barrier(1, NT);  // panel p and every earlier trailing update are in shared memory
        glap(0);
        if (p == NP - 1) break;
        // ---- trailing columns: every warp owns whole column tiles -- U12 of the eight columns by lanes 0..7 (the
        // 8-step forward substitution with the panel's unit-lower block on the pivot rows: 28 FMAs, a chain of 7),
        // then the update of all row tiles on the tensor path.  Warp 0 takes the columns of the next panel and goes
        // straight on to factorise it (look-ahead): one barrier per panel.
        const unsigned dm0 = donebuf[2 * (p & 1)], dm1 = donebuf[2 * (p & 1) + 1];
        const int nct = NP - 1 - p;
        // The look-ahead tile goes first and alone: a dependent chain -- warp 0's forward substitution and its two-deep
        // tensor chains -- starves beside warps that issue independent FP64 work on the same sub-partition (measured:
        // 8 -> 100 cycles per dependent FMA, 26 -> 400 per dependent tensor instruction next to three saturating warps).
        // The other warps start their column tiles when warp 0 has finished its own (named barrier 3: warp 0 arrives,
        // the others wait) and work in the shadow of the next panel.
        if constexpr (kLookaheadFirst) {
            if (warp != 0) asm volatile("bar.sync 3, %0;" ::"r"(NT) : "memory");
        }
        for (int ct = warp; ct < nct; ct += (warp == 0 ? nct : NW - 1)) {
            const int col0 = c0 + 8 + 8 * ct;
            if (lane < 8) {
                const int c = col0 + lane;
                double u[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const double *row = W + order[c0 + i] * LDW;
                    double acc = row[c];
#pragma unroll
                    for (int j = 0; j < i; ++j) acc = fma(-row[c0 + j], u[j], acc);
                    u[i] = acc;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) W[order[c0 + i] * LDW + c] = u[i];
            }
            __syncwarp();
            if (warp == 0) glap(1);
            double bv[2];  // rows of U12: the B operand of every row tile
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) bv[kk] = W[order[c0 + 4 * kk + (lane & 3)] * LDW + col0 + (lane >> 2)];
#pragma unroll
            for (int rt = 0; rt < NP; ++rt) {
                const int row = 8 * rt + (lane >> 2);
                const bool rdone = ((row < 32 ? (dm0 >> row) : (dm1 >> (row - 32))) & 1u) != 0;
                double2 *cptr = reinterpret_cast<double2 *>(W + row * LDW + col0 + 2 * (lane & 3));
                double2 cv = *cptr;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const double lv = W[row * LDW + c0 + 4 * kk + (lane & 3)];
                    dmma_m8n8k4(cv.x, cv.y, rdone ? 0.0 : -lv, bv[kk]);  // multipliers of used rows masked
                }
                *cptr = cv;
            }
            __syncwarp();
            if (warp == 0) glap(2);
            if constexpr (kLookaheadFirst) {
                if (warp == 0) asm volatile("bar.arrive 3, %0;" ::"r"(NT) : "memory");
            }
        }
        lap(1);
    }
}

}  // namespace dpilqr
