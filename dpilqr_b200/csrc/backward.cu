// backward.cu -- Kernel 3: backward Riccati recursion, one CTA per problem (sm_100a, FP64).
//
// Replaces ilqrSolver._backward_pass (reference control.py:116-148).  Per time step, with the
// block structure the reference throws away (A, B block-diagonal per agent; P symmetric):
//
//   phase A   S = B^T (P + mu I)  ->  Q_ux = S A,  Q_uu = L_uu + S B      (S only lives in registers)
//             Q_u = L_u + B^T p,  Q_x = L_x + A^T p
//   then two warp groups run concurrently (named barriers):
//     group 1   phase C  partial-pivot LU of Q_uu without row swaps (pivot rows are marked; the reference calls
//                        np.linalg.solve = dgesv, control.py:141-142).  Tensor-path kernels: four warps, blocked
//                        panels of 8 columns factorised in registers by one warp (lu.cuh, lu_blocked); otherwise
//                        eight warps, unblocked with look-ahead pivot search (lu_lookahead).
//     group 2   phase B  Q_xx = L_xx + A^T P A, in place on the upper-triangular blocks of P
//   pack      the factors gathered in pivot order (+ 8x8 inverses of the unit-lower diagonal blocks), all threads
//   phase D   K = -Q_uu^{-1} Q_ux, d = -Q_uu^{-1} Q_u: blocked triangular solves, one warp per 8 right-hand sides
//             (tensor path) or one thread per right-hand side (generic sizes)
//   phase E   pq = Q_ux^T d, z = Q_uu d + Q_u, then Y = Q_uu K + 2 Q_ux (in place over Q_ux)
//   phase F   P <- Q_xx + 1/2 (K^T Y + Y^T K)   (== the reference's symmetrised
//             Q_xx + K^T Q_uu K + K^T Q_ux + Q_ux^T K), upper blocks only
//             p <- Q_x + K^T z + Q_ux^T d
//
// The GEMM-shaped phases D, E, F and the trailing updates of the LU run on the FP64 tensor path
// (mma.sync.m8n8k4.f64, "DMMA") when the joint sizes are multiples of 8, otherwise on DFMA register tiles.
// P, Q_ux/Y, K, Q_uu and the LU factors stay in shared memory for the whole recursion (226.5 kB for 10
// Quadcopter12D agents -> one CTA per SM); only the stage records stream in (TMA bulk copies, one step ahead)
// and K, d stream out.  Problems too large for shared memory keep Q_ux/Y, K and the LU matrices in an L2-resident
// global scratch instead (same code).
#include <stdlib.h>

#include "kernels.cuh"
#include "lu.cuh"

namespace dpilqr {

// The 8x8 diagonal blocks of U are applied as explicit inverses on the tensor path in the backward substitution of phase D,
// like those of L (-DDPILQR_UPPER_INVERSE=0: substituted through by eight lanes, a serial chain that was 6 % of the
// kernel's samples).  U carries the conditioning of Q_uu, so the inverse costs a little accuracy -- measured on the metric
// family: K, d at t = 0 against the reference 1.4e-14 / 4.3e-14 instead of 1.0e-14 / 1.1e-14, the same 118 of 125
// scenarios of the bench's parity sample within 1e-9 (median 6.6e-13), crowded scenarios (cond 3e9) inside their bar --
// for 2.8 % of the kernel's time.
#ifndef DPILQR_UPPER_INVERSE
#define DPILQR_UPPER_INVERSE 1
#endif
constexpr bool kUpperInverse = DPILQR_UPPER_INVERSE != 0;
#ifndef DPILQR_PACK_IN_LU
#define DPILQR_PACK_IN_LU 0
#endif
#ifndef DPILQR_MERGE_E
#define DPILQR_MERGE_E 1
#endif
#ifndef DPILQR_K_FROM_D
#define DPILQR_K_FROM_D 1
#endif
#ifndef DPILQR_LU_ON_SCHED0
#define DPILQR_LU_ON_SCHED0 1
#endif
#ifndef DPILQR_VEC_IN_WINDOW
#define DPILQR_VEC_IN_WINDOW 1
#endif
#ifndef DPILQR_F_BALANCE
#define DPILQR_F_BALANCE 1
#endif
// Share of the column tiles of Q_ux = S A computed by the Q_xx warps (behind their blocks) instead of the LU group's
// update warps: two fifths balance the two groups of the LU window for ten drones (LU 18.6 k -> 16.3 k cycles, Q_xx
// group 16.7 k -> 17.9 k; -DDPILQR_QUX_SHARE_GROUP2_PCT=0: all on the update warps, as in round 1).
#ifndef DPILQR_QUX_SHARE_GROUP2_PCT
#define DPILQR_QUX_SHARE_GROUP2_PCT 54
#endif
__host__ __device__ constexpr int qux_tiles_group2(int nt) { return nt * DPILQR_QUX_SHARE_GROUP2_PCT / 100; }

template <int S>
struct TileSize {
    static constexpr int value = (S % 4 == 0) ? 4 : (S % 3 == 0) ? 3 : S;
};

struct BackwardSmem {
    size_t Pb, QUU, W, Lp, Up, rdiag, sA, sB, sL, pvec, Qx, pq, Qu, dv, zv, order, keys, rinv, scal, tacc, mbar, mats, total_doubles;
};

// Row stride of the Q_ux / K buffers: room for right-hand side n (Q_u) rounded up to a tile of 8, and congruent to
// 4 modulo 16 doubles: the four rows of a tensor-path operand fragment (lane l reads row l%4, column l/4) then start
// 8 banks apart and a half-warp touches every bank once.
__host__ __device__ constexpr int backward_ldn(int n)
{
    return (((n + 8) & ~7) & 15) == 0 ? ((n + 8) & ~7) + 4 : ((n + 8) & ~7) + 12;
}

// Row stride of the packed LU factors, same rule (operand fragments of the blocked substitution)
__host__ __device__ constexpr int backward_ldf(int m)
{
    const int e = (m + 1) & ~1;
    return (e & 15) == 4 || (e & 15) == 12 ? e : (((e & 15) < 4) ? (e & ~15) + 4 : ((e & 15) < 12) ? (e & ~15) + 12 : (e & ~15) + 20);
}

__host__ __device__ constexpr size_t even_up(size_t v) { return (v + 1) & ~(size_t)1; }

// doubles of global (L2-resident) scratch per CTA when a team is too large for shared memory
__host__ __device__ constexpr size_t backward_scratch_per_cta(int m, int n)
{
    return 2 * (size_t)m * backward_ldn(n) + even_up((size_t)m * backward_ldw(m)) + 2 * even_up((size_t)m * backward_ldf(m))
           + even_up((size_t)m * (m + 4));  // (Q_uu: only the 16-agent tensor-path kernel keeps it here)
}

// Shared-memory carve-up in doubles (everything 16-byte aligned).
__host__ __device__ constexpr BackwardSmem backward_smem(int a, int S, int C, bool mats_in_smem, bool quu_in_smem = true)
{
    const int n = a * S, m = a * C, pairs = a * (a - 1) / 2;
    const size_t nblk = (size_t)a * (a + 1) / 2;
    BackwardSmem L{};
    size_t off = 0;
    L.Pb = off;    off += even_up(nblk * (S * S + 2));
    L.QUU = off;   if (quu_in_smem) off += even_up((size_t)m * (m + 4));
    // the LU work matrix and the packed factors move to the global scratch together with Q_ux / K for big teams
    L.W = off;     if (mats_in_smem) off += even_up((size_t)m * backward_ldw(m));
    L.Lp = off;    if (mats_in_smem) off += even_up((size_t)m * backward_ldf(m));
    L.Up = off;    if (mats_in_smem) off += even_up((size_t)m * backward_ldf(m));
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
    L.scal = off;  off += 2;                                   // mu, reference-cost weight
    L.tacc = off;  off += 28;                                  // debug cycle counters
    L.mbar = off;  off += 2;                                   // mbarrier of the stage-record bulk copies
    L.mats = off;
    if (mats_in_smem) off += 2 * (size_t)m * backward_ldn(n);
    L.total_doubles = off;
    return L;
}

template <int S, int C, int AT, bool GLOBAL, bool TIMED>
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
    constexpr bool MMA_A = USE_MMA && S == 12 && C == 4 && AT % 2 == 0;  // phase A on the tensor path
    constexpr bool kMergeE = USE_MMA && !GLOBAL && (AT * S) / 8 < 16 && (AT * S) % 8 == 0 && kUpperInverse && DPILQR_MERGE_E;  // phase E in one piece, see there
    constexpr bool kVecInWindow = MMA_A && DPILQR_LU_ON_SCHED0 && DPILQR_VEC_IN_WINDOW;  // Q_u, Q_x computed beside the LU, see there
    constexpr bool kKFromD = kMergeE && DPILQR_K_FROM_D;  // K[t], d[t] leave for HBM from the fragments of phase D
    const int LDQ = m + 4;
    const int LDW = backward_ldw(m);  // row stride of the LU work matrix
    const int LDF = backward_ldf(m);  // row stride of the packed factors
    const int LDN = backward_ldn(n);  // row stride of Q_ux / K: column n carries Q_u / d, columns n+1.. are zero padding
    constexpr int nthr = 512;  // launch_backward always starts 512 threads
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const StageLayout L = stage_layout(a, S, C);
    // 16 agents (15 drones + the phantom) leave no room for Q_uu beside P: it joins the other matrices in the scratch
    constexpr bool QUU_GLOBAL = GLOBAL && AT > 0 && backward_smem(AT > 0 ? AT : 1, S, C, false).total_doubles * 8 > 227 * 1024;
    const BackwardSmem SM = backward_smem(a, S, C, !GLOBAL, !QUU_GLOBAL);
    // Odd teams run the tensor-path kernel of the next even size: the stage records carry a PHANTOM agent (A = I,
    // B = 0, no cost gradient, no proximity term: the zero background of linquad_kernel), whose blocks of P, Q_uu, K
    // never couple to the real agents' (exact zeros); the descriptor arrays and K, d keep the real team's layout.
    const int a_real = bt.n_agents;
    const int n_real = a_real * S, m_real = a_real * C;

    double *Pb = smem + SM.Pb;        // [nblk][PBS]   upper-triangular blocks of P
    double *QUU = smem + SM.QUU;      // [m][LDQ]
    double *W = smem + SM.W;          // [m][LDW] row-major LU work matrix
    double *Lp = smem + SM.Lp;        // [m][LDF] packed unit-lower factor, Lp[k*LDF + k2] = l(k2, k), k2 > k
    double *Up = smem + SM.Up;        // [m][LDF] packed upper factor, column-major: Up[c*LDF + k] = u(k, c), k <= c
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
        Up = Lp + even_up((size_t)m * backward_ldf(m));
        if constexpr (QUU_GLOBAL) QUU = Up + even_up((size_t)m * backward_ldf(m));
    } else {
        QUX = smem + SM.mats;
        KB = QUX + (size_t)m * LDN;
    }

    // The kernel is capped at 128 registers and local memory has next to no L1 behind it (227 kB of shared memory
    // carved out): per-problem pointers are re-derived from b where they are used (the opaque copy keeps the compiler
    // from hoisting them into registers that live for the whole recursion) and the two per-problem scalars sit in
    // shared memory.
    auto problem = [b]() {
        int v = b;
        asm volatile("" : "+r"(v));
        return v;
    };
    auto cost_row = [&](int i) { return (int64_t)bt.cost_idx[(int64_t)problem() * a_real + min(i, a_real - 1)]; };  // (the phantom borrows the last agent's cost matrices)
    double *scal = smem + SM.scal;  // [0] mu, [1] weight of the reference cost
    if (threadIdx.x == 0) {
        scal[0] = p.mu[b];
        scal[1] = bt.weights ? bt.weights[2 * b] : 1.0;
    }
    int st = 0;
    // optional per-phase cycle counters (debug aid, see dpilqr_debug_backward_timing); kept in shared memory
    long long *tacc = reinterpret_cast<long long *>(smem + SM.tacc) + (tid < 32 ? 0 : 12);
    long long tmark = 0;
    constexpr int kTimedThread2 = 160;  // first thread of warp group 2 (warp 1) in the instrumented tensor-path build
    const bool timing = TIMED && (p.timing != nullptr) && (blockIdx.x == 0) && ((tid >> 5) == 0 || (tid >> 5) == (kTimedThread2 >> 5));  // whole warps: no divergence from timing
    if (timing && (tid & 31) == 0) {
        for (int k = 0; k < 12; ++k) tacc[k] = 0;
        if (tid == 0) tacc[24] = tacc[25] = tacc[26] = 0;
    }
    auto tick = [&](int slot) {
        if (timing) {
            const long long now = clock64();
            if ((tid & 31) == 0) tacc[slot] += now - tmark;
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
    auto prefetch_record = [&](int t, int issuer = 0) {  // call after a __syncthreads(): nobody reads the previous record any more
        const double *rec = p.stage + ((int64_t)problem() * (T + 1) + t) * L.stride;
        if constexpr (BULK) {
            // the shared-memory home sA | sB | sLx.. mirrors the record layout (common.cuh): one bulk copy
            if ((int)threadIdx.x == issuer) {
                const unsigned total = (unsigned)(L.stride * 8);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(total) : "memory");
                // a few medium-sized copies move faster than one large one (the copies are processed concurrently)
                constexpr unsigned kChunk = 4096;
#pragma unroll 1
                for (unsigned off = 0; off < total; off += kChunk)
                    bulk_copy(sA + off / 8, rec + off / 8, min(kChunk, total - off));
            }
        } else {
            for (int k = tid; k < a * S * S; k += nthr) cp_async8(sA + (k / (S * S)) * SAS + k % (S * S), rec + L.offA + (k / (S * S)) * L.strideA + k % (S * S));
            for (int k = tid; k < a * S * C; k += nthr) cp_async8(sB + (k / (S * C)) * SBS + k % (S * C), rec + L.offB + (k / (S * C)) * L.strideB + k % (S * C));
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
            const double *Qf = bt.Qf + cost_row(i) * S * S;
            v = scal[1] * (Qf[r * S + cc] + Qf[cc * S + r]);
            if (r < 3 && cc < 3) v += sHd[9 * i + r * 3 + cc];
        } else if (r < 3 && cc < 3) {
            v = sHo[9 * pair_index(i, j, a) + r * 3 + cc];
        }
        Pb[(size_t)blk * PBS + e] = v;
    }
    for (int k = tid; k < n; k += nthr) pvec[k] = sLx[k];
    __syncthreads();
    prefetch_record(T - 1);

    // Every phase re-derives its thread coordinates from an opaque copy of threadIdx.x: the compiler can then neither
    // hoist a phase's index arithmetic out of the recursion nor keep it alive across the other phases.  The kernel is
    // capped at 128 registers, and whatever is live across the LU leaves the panel warp fewer registers to schedule in.
#define DPILQR_PHASE_IDS                                   \
    int tid_phase_ = threadIdx.x;                          \
    asm volatile("" : "+r"(tid_phase_));                   \
    const int tid = tid_phase_, lane = tid & 31, warp = tid >> 5; \
    (void)lane, (void)warp;
    auto lu_experiment = [&](int slot) {  // timing experiment: the LU on its own on a synthetic matrix (clobbers W, order)
        if constexpr (TIMED && USE_MMA) {
            for (int k = threadIdx.x; k < m * LDW; k += nthr) W[k] = (k % (LDW + 1) == 0) ? 2.0 : 0.001 * ((k * 37) % 101);
            __syncthreads();
            long long best = 1ll << 60;
            for (int rep = 0; rep < 4; ++rep) {
                __syncthreads();
                const long long q0 = clock64();
                if (threadIdx.x < 128)
                    lu_blocked<AT * C, 128, TIMED>(W, order, reinterpret_cast<unsigned *>(colbuf), threadIdx.x, nullptr);
                __syncthreads();
                const long long q1 = clock64() - q0;
                best = q1 < best ? q1 : best;
            }
            if (blockIdx.x == 0 && threadIdx.x == 0) p.timing[slot] = best;
        }
    };
    if (p.debug_mode & 32) lu_experiment(27);
    if (timing) tmark = clock64();
#pragma unroll 1
        // Pack the factors in pivot order so that the substitutions read contiguous memory, and invert the 8x8 diagonal
        // blocks.  gn threads (gt = 0..gn-1) that synchronise with sync().
        auto pack_factors = [&](int gt, int gn, auto sync) {
        for (int e = gt; e < m * m; e += gn) {
            const int k = e / m, x2 = e - k * m;
            const double v = W[order[x2] * LDW + k];
            if (x2 > k) Lp[k * LDF + x2] = v;   // l(x2, k)
            else Up[k * LDF + x2] = v;          // u(x2, k): column k, row x2 <= k
            if (x2 == k) {
                if (v == 0.0) st |= DPILQR_ST_SINGULAR;  // exact zero pivot: dgesv's info > 0 (a NaN pivot is not: np.linalg.solve returns NaN)
                rdiag[k] = __drcp_rn(v);
            }
        }
        if constexpr (USE_MMA) {
            // The blocked forward solve of phase D applies the 8x8 diagonal blocks of the unit lower factor (well
            // conditioned: |l| <= 1) as explicit inverses on the tensor path: invert them here, in place.  One
            // thread per (block, column of the inverse).
            constexpr int M = AT * C, NB = M / 8;
            sync();
            const int bb = (gt & 63) >> 3, j = gt & 7;
            double x[8];
            const bool busy = gt < NB * 8;
            const bool busy_u = kUpperInverse && gt >= 64 && gt < 64 + NB * 8;  // (NB <= 8: the two groups are disjoint)
            if (busy) {
                const double *F = Lp + (8 * bb) * LDF + 8 * bb;  // F[c * LDF + rr] = l(rr, c) of this block
#pragma unroll
                for (int i = 0; i < 8; ++i) {  // unit lower: solve L x = e_j by forward substitution
                    double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
                    for (int c = 0; c < i; ++c) acc = fma(-F[c * LDF + i], x[c], acc);
                    x[i] = acc;
                }
            } else if (busy_u) {
                const double *F = Up + (8 * bb) * LDF + 8 * bb;  // F[c * LDF + rr] = u(rr, c), rr <= c
#pragma unroll
                for (int i = 7; i >= 0; --i) {  // upper: solve U x = e_j by back substitution (x_i = 0 for i > j)
                    // row i scaled by 1 / u_ii beforehand: the chain from x_7 down to x_0 is then one FMA per row (a
                    // dependent FP64 operation costs some 50 cycles) instead of an FMA and a multiplication
                    const double ri = rdiag[8 * bb + i];
                    double acc = (i == j) ? ri : 0.0;
#pragma unroll
                    for (int c = i + 1; c < 8; ++c) acc = fma(-(F[c * LDF + i] * ri), x[c], acc);
                    x[i] = (i <= j) ? acc : 0.0;
                }
            }
            sync();
            if (busy) {
#pragma unroll
                for (int i = 0; i < 8; ++i) Lp[(8 * bb + j) * LDF + 8 * bb + i] = x[i];  // inverse(i, j), same transposed layout
            } else if (busy_u) {
#pragma unroll
                for (int i = 0; i < 8; ++i) Up[(8 * bb + j) * LDF + 8 * bb + i] = x[i];  // inverse(i, j); zero below the diagonal
            }
        }
        };
        // The LU group finishes ahead of the Q_xx warps in the shared-memory tensor-path kernels: it packs its factors
        // there, off the critical path (with the factors in the L2 scratch 128 threads would be too few to hide its latency).
        constexpr bool kPackInLu = USE_MMA && !GLOBAL && DPILQR_PACK_IN_LU;
    for (int t = T - 1; t >= 0; --t) {
        if ((p.debug_mode & 256) && t == T - 1) lu_experiment(28);
        {
            DPILQR_PHASE_IDS
        // Symmetrise the diagonal blocks, P <- (P + P^T)/2 (control.py:146-147).  They are the only part of P stored on
        // both sides of the diagonal, and phases B and F round the two triangles independently: the antisymmetric
        // residue is not damped by the recursion (it propagates with the open-loop A^T . A and grew by about 12 % per
        // time step on Quadcopter12D, costing two to three digits of K and d at t = 0).
        if constexpr (kMergeE) {
            // Phase E in one piece: the lower half of the CTA finishes the step before -- p <- Q_x + K^T z + Q_ux^T d, whose
            // first reader is phase A (K, z, pq and Q_x are intact until then) -- while the upper half symmetrises and
            // sets the regularised diagonal up.  A dependent FP64 operation costs 50 to 70 cycles here, so both are
            // built for short chains: two threads per entry of p with four partial sums of five terms each; every
            // symmetrisation pair loaded before the first is stored.
            constexpr int N = AT * S, M = AT * C, LD = backward_ldn(N);
            constexpr int NPAIR = AT * (S * (S - 1) / 2), ITER = (NPAIR + 255) / 256;
            static_assert(2 * N <= 256 && N <= 256, "p update: two threads per column in the lower half of the CTA");
            if (tid < 256) {
                const int col = min(tid >> 1, N - 1), h = tid & 1;  // (the surplus lanes of the last warp shuffle along)
                if (t < T - 1) {
                    const double *kp = KB + (size_t)(4 * h) * LD + col;
                    const double *zp = zv + 4 * h;
                    double a4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                    for (int i = 0; i < M / 8; ++i)
#pragma unroll
                        for (int c = 0; c < 4; ++c) a4[c] = fma(kp[(size_t)(8 * i + c) * LD], zp[8 * i + c], a4[c]);
                    double sum = (a4[0] + a4[1]) + (a4[2] + a4[3]);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                    const double pnew = Qx[col] + sum + pq[col];
                    if (!isfinite(pnew)) st |= DPILQR_ST_NONFINITE;
                    if (h == 0 && tid < 2 * N) pvec[col] = pnew;
                }
            } else {
                const int u = tid - 256;
                if (u < N) {  // P + mu I for phase A (control.py:134-135): P itself stays as it is
                    const int i = u / S, r = u - i * S;
                    colbuf[u] = Pb[(size_t)blk_index(i, i) * PBS + r * S + r] + scal[0];
                }
                double *lo[ITER], *up[ITER];
                double v[ITER];
#pragma unroll
                for (int it = 0; it < ITER; ++it) {
                    const int q = min(u + 256 * it, NPAIR - 1);  // (the surplus threads compute the last pair and drop it)
                    const int i = q / (S * (S - 1) / 2);
                    int w = q - i * (S * (S - 1) / 2), r = 0;
                    while (w >= S - 1 - r) { w -= S - 1 - r; ++r; }
                    const int cc = r + 1 + w;
                    double *blk = Pb + (size_t)blk_index(i, i) * PBS;
                    up[it] = blk + r * S + cc, lo[it] = blk + cc * S + r;
                    v[it] = 0.5 * (*up[it] + *lo[it]);
                }
#pragma unroll
                for (int it = 0; it < ITER; ++it)
                    if (u + 256 * it < NPAIR) *up[it] = v[it], *lo[it] = v[it];
            }
        } else {
        {
            for (int k = tid; k < a * S * S; k += nthr) {  // one item per entry, the upper ones act (constant divisors only)
                const int i = k / (S * S), e = k - i * (S * S);
                const int r = e / S, cc = e - r * S;
                if (r < cc) {
                    double *blk = Pb + (size_t)blk_index(i, i) * PBS;
                    const double v = 0.5 * (blk[e] + blk[cc * S + r]);
                    blk[e] = v;
                    blk[cc * S + r] = v;
                }
            }
        }
        // Regularise P in place for phase A (P + mu I, control.py:134-135); the plain diagonal waits in pq (free until
        // phase E) and is put back before phase B, which needs the unregularised P.
        // The tensor-path form of phase A leaves P alone and takes the regularised diagonal from a side buffer (the LU's
        // look-ahead buffer, free outside the factorisation): nothing to restore, no barrier in front of phase B.
        if constexpr (MMA_A) {
            for (int k = tid; k < n; k += nthr) {
                const int i = k / S, r = k - i * S;
                colbuf[k] = Pb[(size_t)blk_index(i, i) * PBS + r * S + r] + scal[0];
            }
        }
        }
        if constexpr (!MMA_A) {
        for (int k = (p.debug_mode & 8) ? n : tid; k < n; k += nthr) {
            const int i = k / S, r = k - i * S;
            double *pd = Pb + (size_t)blk_index(i, i) * PBS + r * S + r;
            const double v = *pd;
            pq[k] = v;
            *pd = v + scal[0];
        }
        }
        }
        tick(9);
        wait_record();
        __syncthreads();
        tick(0);

        {
            DPILQR_PHASE_IDS
        // ---- phase A: Q_ux, Q_uu, Q_u, Q_x
        if constexpr (MMA_A) {
            // Tensor-path form for 12-state / 4-control agents.  S = B^T (P + mu I) is a [m x n] product whose 8-row
            // tiles are the rows of two agents: the A operand holds B_i^T of one agent per k-step (zero rows for the
            // other one), the B operand reads P straight from its upper-block storage (transposed below the
            // diagonal).  S is staged in the K buffer (K of step t+1 is dead by now); Q_ux = S A and Q_uu = S B then
            // take their k-steps only from the agents whose columns a tile covers (block-diagonal A, B).
            constexpr int N = AT * S, M = AT * C, LD = backward_ldn(N);
            constexpr int NT = N / 8, MT = M / 8;
            const int fr = lane >> 2, fc = lane & 3;
            double *Ssm = KB;
            for (int tile = warp; tile < MT * NT; tile += nwarp) {
                const int pr = tile / NT, ct = tile - pr * NT;
                const int col = 8 * ct + fr;  // column of P this lane feeds
                const int bj = col / S, cc = col - bj * S;
                double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < 6; ++ks) {
                    const int ag = 2 * pr + ks / 3;    // agent whose rows of P this k-step covers
                    const int rr = 4 * (ks % 3) + fc;  // row inside that agent's block
                    const double av = ((fr >> 2) == ks / 3) ? sB[ag * SBS + rr * C + (fr & 3)] : 0.0;
                    // (diagonal entries come regularised from their side buffer: P itself stays as it is)
                    // (an index into the shared-memory array, not a pointer: the select must not turn the load generic)
                    int src = (int)SM.Pb + ((ag <= bj) ? blk_index(ag, bj) * PBS + rr * S + cc : blk_index(bj, ag) * PBS + cc * S + rr);
                    if (ag == bj && rr == cc) src = (int)SM.keys + col;
                    const double bv = smem[src];
                    if (ks < 3) dmma_m8n8k4(c0, c1, av, bv);
                    else dmma_m8n8k4(e0, e1, av, bv);
                }
                *reinterpret_cast<double2 *>(Ssm + (size_t)(8 * pr + fr) * LD + 8 * ct + 2 * fc) = make_double2(c0 + e0, c1 + e1);
            }
            if constexpr (!kVecInWindow) {
            for (int row = tid; row < m; row += nthr) {  // Q_u = L_u + B^T p
                const int i = row / C, g = row - i * C;
                const double *Bi = sB + i * SBS;
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < S; ++r) acc = fma(Bi[r * C + g], pvec[i * S + r], acc);
                const double qu = sLu[row] + acc;
                Qu[row] = qu;
                QUX[(size_t)row * LDN + n] = qu;  // Q_u rides along as right-hand side n
            }
            }
            __syncthreads();
            for (int q = warp; q < MT * MT; q += nwarp) {  // Q_ux = S A follows in warp group 2, beside the LU
                {
                    // Q_uu tile: rows 8 mt.. (agents 2 mt, 2 mt + 1), columns 8 nt.. (agents 2 nt, 2 nt + 1)
                    const int mt = q / MT, nt = q - mt * MT;
                    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;  // one accumulator chain per agent of the column pair
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int ja = 2 * nt + e;
                            const double *sp = Ssm + (size_t)(8 * mt + fr) * LD + ja * S + fc;
                            const double *bp = sB + ja * SBS + fc * C + (fr & 3);
                            const double bv = ((fr >> 2) == e) ? bp[4 * kk * C] : 0.0;
                            if (e == 0) dmma_m8n8k4(c0, c1, sp[4 * kk], bv);
                            else dmma_m8n8k4(e0, e1, sp[4 * kk], bv);
                        }
                    }
                    c0 += e0, c1 += e1;
                    const int row = 8 * mt + fr, colq = 8 * nt + 2 * fc;
                    const int ir = row / C, g = row - ir * C, ic = colq / C, g2 = colq - ic * C;
                    if (ir == ic) {  // L_uu of the reference cost, w (R + R^T) (cost.py:85-93)
                        const double *R = bt.R + cost_row(ir) * C * C;
                        c0 += scal[1] * (R[g * C + g2] + R[g2 * C + g]);
                        c1 += scal[1] * (R[g * C + g2 + 1] + R[(g2 + 1) * C + g]);
                    }
                    *reinterpret_cast<double2 *>(QUU + row * LDQ + colq) = make_double2(c0, c1);
                    *reinterpret_cast<double2 *>(W + row * LDW + colq) = make_double2(c0, c1);
                }
            }
        } else {
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
            // outer loops rolled: the whole recursion step has to stay resident in the instruction cache
#pragma unroll 1
            for (int r = 0; r < S; ++r) {
                const double bv = Bi[r * C + g];
                const double *prow = Pblk + r * rs;
#pragma unroll
                for (int sg = 0; sg < S; ++sg) Srow[sg] = fma(bv, prow[sg * cs], Srow[sg]);
            }
            const double *Aj = sA + j * SAS;
            const double *Bj = sB + j * SBS;
            const int row = i * C + g;
#pragma unroll 1
            for (int sg2 = 0; sg2 < S; ++sg2) {
                double acc = 0.0;
#pragma unroll
                for (int sg = 0; sg < S; ++sg) acc = fma(Srow[sg], Aj[sg * S + sg2], acc);
                QUX[(size_t)row * LDN + j * S + sg2] = acc;  // L_ux == 0 (cost.py:91)
            }
#pragma unroll 1
            for (int g2 = 0; g2 < C; ++g2) {
                double acc = 0.0;
#pragma unroll
                for (int sg = 0; sg < S; ++sg) acc = fma(Srow[sg], Bj[sg * C + g2], acc);
                if (i == j) {
                    const double *R = bt.R + cost_row(i) * C * C;
                    acc += scal[1] * (R[g * C + g2] + R[g2 * C + g]);
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
        }
        if constexpr (!kVecInWindow) {
        for (int col = tid; col < n; col += nthr) {
            const int j = col / S, sg = col - j * S;
            const double *Aj = sA + j * SAS;
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < S; ++r) acc = fma(Aj[r * S + sg], pvec[j * S + r], acc);
            Qx[col] = sLx[col] + acc;
        }
        }
        }
        if constexpr (TIMED) {
            if (p.debug_mode & 64) {  // experiment: factorise a synthetic matrix instead of Q_uu (wrong numerics)
                __syncthreads();
                for (int k = threadIdx.x; k < m * LDW; k += nthr) W[k] = (k % (LDW + 1) == 0) ? 2.0 : 0.001 * ((k * 37) % 101);
            }
        }
        if ((p.debug_mode & 512) && t == T - 2) { __syncthreads(); lu_experiment(29); }
        __syncthreads();
        tick(1);

        {
            DPILQR_PHASE_IDS
        // Two warp groups run side by side.  Tensor-path kernels split by scheduler: the warps of SM sub-partition 0
        // (warp % 4 == 0) factorise Q_uu -- the panel warp is latency-bound and shares its FP64 pipe with nothing but the
        // trailing updates of its own group -- while the twelve warps of the other sub-partitions compute Q_xx and
        // Q_ux.  Otherwise: first 256 threads / the rest.
        constexpr int kLuThreads = USE_MMA ? 128 : kSolveThreads;
        // (-DDPILQR_LU_ON_SCHED0=0: the earlier split -- warps 0..3 factorise, one per sub-partition, with the three update
        // warps competing with nine Q_xx warps for the tensor pipes and the three other warps of the panel warp's
        // sub-partition parked.  Measured, ten drones: LU group 15.2 k, Q_xx warps 16.5 k cycles per step, against 15.8 k
        // and 14.8 k with the split by sub-partition; 12 drones 30.7 -> 24.8 us per problem.)
        constexpr bool kLuSched0 = USE_MMA && DPILQR_LU_ON_SCHED0;
        const bool lu_group = kLuSched0 ? ((warp & 3) == 0) : (tid < kLuThreads);
        // one 8-column tile of Q_ux = S A on the tensor path: the A_j operand is shared by the MT row tiles, whose
        // independent accumulator chains keep the tensor pipe busy
        auto qux_tile = [&](int ct) {
            if constexpr (MMA_A) {
                constexpr int N = AT * S, M = AT * C, LD = backward_ldn(N);
                constexpr int MT = M / 8;
                const int fr = lane >> 2, fc = lane & 3;
                const double *Ssm = KB;
                const int col = 8 * ct + fr;
                const int bj = col / S, cc = col - bj * S;
                const int j0 = (8 * ct) / S, j1 = (8 * ct + 7) / S;
                double2 acc[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) acc[mt] = make_double2(0.0, 0.0);
                for (int ja = j0; ja <= j1; ++ja) {
                    const double *sp = Ssm + (size_t)fr * LD + ja * S + fc;
                    const double *ap = sA + ja * SAS + fc * S + cc;
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk) {
                        const double bv = (ja == bj) ? ap[4 * kk * S] : 0.0;
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) dmma_m8n8k4(acc[mt].x, acc[mt].y, sp[(size_t)8 * mt * LD + 4 * kk], bv);
                    }
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                    *reinterpret_cast<double2 *>(QUX + (size_t)(8 * mt + fr) * LD + 8 * ct + 2 * fc) = acc[mt];
            }
        };
        const bool idle_group = USE_MMA && !kLuSched0 && !lu_group && ((warp & 3) == 0);  // shares the scheduler of the panel warp
        if (lu_group) {
            // ================= group 1: phase C, LU with implicit partial pivoting =================
            const int gt = kLuSched0 ? ((warp >> 2) << 5) + lane : tid;
            if constexpr (MMA_A) {
                // Q_ux = S A (L_ux == 0, cost.py:91): only phase D needs it, so it runs in the shadow of the LU: the three
                // update warps of the LU group do it while warp 0 factorises the first panel (they have nothing else to do
                // until then); the warps that share the panel warp's scheduler stay parked
                // a warp takes whole column tiles -- these warps are due back at the first trailing update; the last
                // qux_tiles_group2(NT) column tiles go to the Q_xx warps instead, behind their blocks
                constexpr int NT = (AT * S) / 8;
                if constexpr (!kLuSched0)
                for (int ct = (gt >> 5) - 1; ct < NT - qux_tiles_group2(NT) && (gt >> 5) >= 1; ct += 3) qux_tile(ct);
            }
            if constexpr (USE_MMA)
                lu_blocked<AT * C, kLuThreads, TIMED>(W, order, reinterpret_cast<unsigned *>(colbuf), gt, (timing && tid < 32) ? tacc + 24 : nullptr,
                                                        (timing && tid < 32) ? p.timing + 64 : nullptr);
            else lu_lookahead<(AT > 0 ? AT * C : 0)>(W, colbuf, rinvbuf, prbuf, order, m, gt);
            if constexpr (kPackInLu) pack_factors(gt, kLuThreads, [] { named_barrier(1, kLuThreads); });
            tick(2);
            if constexpr (USE_MMA && TIMED) {  // timing experiment: a second, warm pass over the same code
                if (p.debug_mode & 4) lu_blocked<AT * C, kLuThreads, TIMED>(W, order, reinterpret_cast<unsigned *>(colbuf), gt, nullptr);
            }
        } else if (!idle_group) {
            // ================= group 2: phase B, Q_xx = L_xx + A^T P A in place (upper blocks) =================
            const int gt = kLuSched0 ? ((((warp >> 2) * 3 + (warp & 3) - 1) << 5) + lane) : USE_MMA ? ((((warp >> 2) - 1) * 3 + (warp & 3) - 1) << 5) + lane : tid - kSolveThreads;
            const int gn = kLuSched0 ? 384 : USE_MMA ? 288 : nthr - kLuThreads;
            const int blocks_per_round = gn / S;
            if constexpr (!MMA_A) {
            for (int k = gt; k < n; k += gn) {  // the unregularised diagonal of P comes back
                const int i = k / S, r = k - i * S;
                Pb[(size_t)blk_index(i, i) * PBS + r * S + r] = pq[k];
            }
            named_barrier(2, gn);
            }
            tick(3);
            if constexpr (MMA_A) {
                // Tensor-path form: every warp owns whole blocks.  V = P_ij A_j in four 8x8 accumulator tiles (the
                // 12x12 block padded to 16x16: rows/columns 12..15 are zero operands), V written over P_ij (the block is
                // consumed), then Q = A_i^T V with V as the B operand, and L_xx added in the epilogue.  No barrier:
                // a block is read and written by one warp only.
                const int fr = lane >> 2, fc = lane & 3;
                const double w_ref = scal[1];
                for (int blk = (p.debug_mode & 16) ? nblk : (gt >> 5); blk < nblk; blk += gn >> 5) {
                    int i = 0, rem = blk;
                    while (rem >= a - i) { rem -= a - i; ++i; }
                    const int j = i + rem;
                    double *Pblk = Pb + (size_t)blk * PBS;
                    const double *Ai = sA + i * SAS, *Aj = sA + j * SAS;
                    const bool diag = (i == j);
                    // L_xx of the reference cost, w (Q + Q^T) (cost.py:85-93), on diagonal blocks: loads issued up front
                    double lq[2][2][2];
                    const double *Qref = bt.Q + (diag ? cost_row(i) : 0) * S * S;
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int r = 8 * mt + fr, cc = 8 * nt + 2 * fc + e;
                                lq[mt][nt][e] = 0.0;
                                if (diag && r < S && cc < S) lq[mt][nt][e] = Qref[r * S + cc] + Qref[cc * S + r];
                            }
                    // k-steps outermost: the tensor instructions are volatile asm and issue in program order, so the four
                    // independent tiles have to be interleaved by hand
                    double2 acc[2][2];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) acc[mt][nt] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int ks = 0; ks < S / 4; ++ks) {
                        double av[2], bv[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            av[h] = (8 * h + fr < S) ? Pblk[(8 * h + fr) * S + 4 * ks + fc] : 0.0;
                            bv[h] = (8 * h + fr < S) ? Aj[(4 * ks + fc) * S + 8 * h + fr] : 0.0;
                        }
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) dmma_m8n8k4(acc[mt][nt].x, acc[mt][nt].y, av[mt], bv[nt]);
                    }
                    __syncwarp();  // every lane has read P_ij: overwrite it with V
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) {
                            const int r = 8 * mt + fr, cc = 8 * nt + 2 * fc;
                            if (r < S && cc < S) *reinterpret_cast<double2 *>(Pblk + r * S + cc) = acc[mt][nt];
                        }
                    __syncwarp();
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) acc[mt][nt] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int ks = 0; ks < S / 4; ++ks) {
                        double av[2], bv[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            av[h] = (8 * h + fr < S) ? Ai[(4 * ks + fc) * S + 8 * h + fr] : 0.0;
                            bv[h] = (8 * h + fr < S) ? Pblk[(4 * ks + fc) * S + 8 * h + fr] : 0.0;
                        }
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) dmma_m8n8k4(acc[mt][nt].x, acc[mt][nt].y, av[mt], bv[nt]);
                    }
                    __syncwarp();  // V has been consumed: overwrite it with Q_xx
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) {
                            const int r = 8 * mt + fr, cc = 8 * nt + 2 * fc;
                            if (r < S && cc < S) {
                                double out[2] = {acc[mt][nt].x, acc[mt][nt].y};
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    double lxx = diag ? w_ref * lq[mt][nt][e] : 0.0;
                                    if (r < 3 && cc + e < 3) {  // the proximity Hessians only touch the 3x3 position corner
                                        if (diag) lxx += sHd[9 * i + r * 3 + cc + e];
                                        else lxx = sHo[9 * pair_index(i, j, a) + r * 3 + cc + e];
                                    }
                                    out[e] = lxx + out[e];
                                }
                                *reinterpret_cast<double2 *>(Pblk + r * S + cc) = make_double2(out[0], out[1]);
                            }
                        }
                }
                {
                    // the column tiles continue the round robin of the blocks: the warps that had one block fewer go first
                    constexpr int NT = (AT * S) / 8, NW = kLuSched0 ? 12 : 9, SHARE = kLuSched0 ? NT : qux_tiles_group2(NT);
                    for (int u = nblk + ((gt >> 5) + NW - nblk % NW) % NW; u < nblk + SHARE; u += NW) qux_tile(NT - SHARE + u - nblk);
                    if constexpr (kVecInWindow) {
                        // Q_u = L_u + B^T p and Q_x = L_x + A^T p: nobody needs them before phase D, and each is a chain of
                        // twelve dependent FMAs that used to trail phase A.  The last two warps of the group -- the
                        // round robin gives them a unit less than most -- run up to three of those chains side by side.
                        if ((gt >> 5) >= NW - 2) {
                            const int u = gt - (NW - 2) * 32;  // 0..63
                            for (int c0 = u; c0 < n; c0 += 128) {
                                const int c1 = min(c0 + 64, n - 1), row = min(u, m - 1);
                                const bool do_u = (c0 == u);  // first round: Q_u rides along
                                const int j0 = c0 / S, j1 = c1 / S, iu = row / C;
                                const double *A0 = sA + j0 * SAS + (c0 - j0 * S), *A1 = sA + j1 * SAS + (c1 - j1 * S);
                                const double *Bu = sB + iu * SBS + (row - iu * C);
                                double a0 = 0.0, a1 = 0.0, au = 0.0;
#pragma unroll
                                for (int r = 0; r < S; ++r) {
                                    a0 = fma(A0[r * S], pvec[j0 * S + r], a0);
                                    a1 = fma(A1[r * S], pvec[j1 * S + r], a1);
                                    if (do_u) au = fma(Bu[r * C], pvec[iu * S + r], au);
                                }
                                Qx[c0] = sLx[c0] + a0;
                                if (c0 + 64 < n) Qx[c1] = sLx[c1] + a1;
                                if (do_u && u < m) {
                                    const double qu = sLu[row] + au;
                                    Qu[row] = qu;
                                    QUX[(size_t)row * LDN + n] = qu;  // Q_u rides along as right-hand side n
                                }
                            }
                        }
                    }
                }
            } else {
            for (int blk0 = (p.debug_mode & 16) ? nblk : 0; blk0 < nblk; blk0 += blocks_per_round) {
                const int blk = blk0 + gt / S;
                const int sg = gt % S;
                const bool live = (gt < blocks_per_round * S) && (blk < nblk);
                int i = 0, rem = live ? blk : 0;
                while (rem >= a - i) { rem -= a - i; ++i; }
                const int j = i + rem;
                double *Pblk = Pb + (size_t)(live ? blk : 0) * PBS;
                double v[S];  // column sg of P_ij A_j
                if (live) {
                    const double *Aj = sA + j * SAS;
                    double acol[S];
#pragma unroll
                    for (int q = 0; q < S; ++q) acol[q] = Aj[q * S + sg];
#pragma unroll
                    for (int r = 0; r < S; ++r) {
                        double acc = 0.0;
                        if constexpr (S % 2 == 0) {  // two entries of P_ij per 16-byte load: half the shared-memory requests
#pragma unroll
                            for (int q = 0; q < S; q += 2) {
                                const double2 pp = *reinterpret_cast<const double2 *>(Pblk + r * S + q);
                                acc = fma(pp.y, acol[q + 1], fma(pp.x, acol[q], acc));
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < S; ++q) acc = fma(Pblk[r * S + q], acol[q], acc);
                        }
                        v[r] = acc;
                    }
                }
                named_barrier(2, gn);  // every thread of the block has read P_ij: overwrite it
                if (live) {
                    const double *Ai = sA + i * SAS;
                    const double w_ref = scal[1];
                    constexpr int RS = (S % 2 == 0) ? 2 : 1;  // rows of A_i^T per 16-byte load
#pragma unroll
                    for (int r = 0; r < S; r += RS) {
                        double acc[RS];
#pragma unroll
                        for (int e = 0; e < RS; ++e) acc[e] = 0.0;
#pragma unroll
                        for (int q = 0; q < S; ++q) {
                            if constexpr (RS == 2) {
                                const double2 aa = *reinterpret_cast<const double2 *>(Ai + q * S + r);
                                acc[0] = fma(aa.x, v[q], acc[0]);
                                acc[1] = fma(aa.y, v[q], acc[1]);
                            } else {
                                acc[0] = fma(Ai[q * S + r], v[q], acc[0]);
                            }
                        }
#pragma unroll
                        for (int e = 0; e < RS; ++e) {
                            const int rr = r + e;
                            double lxx = 0.0;
                            if (i == j) {
                                const double *Q = bt.Q + cost_row(i) * S * S;
                                lxx = w_ref * (Q[rr * S + sg] + Q[sg * S + rr]);
                                if (rr < 3 && sg < 3) lxx += sHd[9 * i + rr * 3 + sg];
                            } else if (rr < 3 && sg < 3) {
                                lxx = sHo[9 * pair_index(i, j, a) + rr * 3 + sg];
                            }
                            Pblk[rr * S + sg] = lxx + acc[e];
                        }
                    }
                }
            }
            }
            tick(4);
            tick(2);
        }
        }
        __syncthreads();
        tick(3);
        {
            DPILQR_PHASE_IDS
        // ---- pack the factors in pivot order so the substitutions read contiguous memory (all threads; the
        // shared-memory tensor-path kernels have done it in the LU group, behind the factorisation)
        if constexpr (!kPackInLu) pack_factors(tid, nthr, [] { __syncthreads(); });
        }
        __syncthreads();
        tick(5);
        tick(11);

        {
            DPILQR_PHASE_IDS
        // ---- phase D: K = -Q_uu^{-1} Q_ux, d = -Q_uu^{-1} Q_u.  Right-hand sides 0..n-1 are the columns of Q_ux,
        // right-hand side n is Q_u.  The negated, row-permuted right-hand sides are substituted in place in KB.
        double *Kt = p.K + ((int64_t)problem() * T + t) * m * n;
        if constexpr (USE_MMA) {
            constexpr int M = AT * C, NB = M / 8;
            // Blocked triangular solves: each warp owns tiles of 8 right-hand sides and needs no other warp.  Every
            // 8x8 block operation -- X_b <- inv(F_bb) X_b on the diagonal, X_b2 -= F(b2, b) X_b off it -- is two
            // m8n8k4 FP64 tensor instructions; the off-diagonal updates of one level are independent and interleave.
            const int fr = lane >> 2, fc = lane & 3;
            for (int nt = warp; nt < (n + 8) / 8; nt += nwarp) {
                double *Xc = KB + 8 * nt;
                // X = -P Q_ux for this warp's 8 right-hand sides (rows in pivot order, negated), kept in registers as
                // accumulator fragments for both substitutions: lane = row fr, columns 2 fc and 2 fc + 1 of each block.
                // Only the block being solved goes through shared memory (it is the B operand of its updates).
                double2 xc[NB];
#pragma unroll
                for (int bq = 0; bq < NB; ++bq) {
                    const double2 q = *reinterpret_cast<const double2 *>(QUX + (size_t)order[8 * bq + fr] * LDN + 8 * nt + 2 * fc);
                    xc[bq] = make_double2(-q.x, -q.y);
                }
                // ---- forward substitution with the unit lower factor
#pragma unroll
                for (int blk = 0; blk < NB; ++blk) {
                    double2 *home = reinterpret_cast<double2 *>(Xc + (size_t)(8 * blk + fr) * LDN + 2 * fc);
                    *home = xc[blk];
                    __syncwarp();
                    double d0 = 0.0, d1 = 0.0;  // X_b <- inv(L_bb) X_b
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        dmma_m8n8k4(d0, d1, Lp[(8 * blk + 4 * ks + fc) * LDF + 8 * blk + fr],
                                    Xc[(size_t)(8 * blk + 4 * ks + fc) * LDN + fr]);
                    __syncwarp();
                    xc[blk] = make_double2(d0, d1);
                    if (blk + 1 < NB) {
                        *home = xc[blk];
                        __syncwarp();
                        double xb[2];
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) xb[ks] = Xc[(size_t)(8 * blk + 4 * ks + fc) * LDN + fr];
#pragma unroll
                        for (int b2 = blk + 1; b2 < NB; ++b2)
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks)
                                dmma_m8n8k4(xc[b2].x, xc[b2].y, -Lp[(8 * blk + 4 * ks + fc) * LDF + 8 * b2 + fr], xb[ks]);
                        __syncwarp();
                    }
                }
                // ---- backward substitution with the upper factor
#pragma unroll
                for (int blk = NB - 1; blk >= 0; --blk) {
                    *reinterpret_cast<double2 *>(Xc + (size_t)(8 * blk + fr) * LDN + 2 * fc) = xc[blk];
                    __syncwarp();
                    // The diagonal block: its explicit inverse (formed with the pack) on the tensor path, or -- the
                    // fallback build -- substituted through by lanes 0..7, one right-hand side each.
                    if constexpr (kUpperInverse) {
                        // X_b <- inv(U_bb) X_b on the tensor path (the inverse was formed with the pack)
                        double d0 = 0.0, d1 = 0.0;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            dmma_m8n8k4(d0, d1, Up[(8 * blk + 4 * ks + fc) * LDF + 8 * blk + fr],
                                        Xc[(size_t)(8 * blk + 4 * ks + fc) * LDN + fr]);
                        __syncwarp();
                        xc[blk] = make_double2(d0, d1);
                        *reinterpret_cast<double2 *>(Xc + (size_t)(8 * blk + fr) * LDN + 2 * fc) = xc[blk];
                        if constexpr (kKFromD) {
                            // these rows of K[t] (and of d[t], right-hand side n) are final: they leave for HBM straight
                            // from the accumulator fragment, 64 contiguous bytes per row of the tile
                            const int row = 8 * blk + fr, col = 8 * nt + 2 * fc;
                            if (row < m_real) {
                                if (col < n_real) {  // (n_real is even: the pair stays inside the row)
                                    if (!isfinite(d0) || !isfinite(d1)) st |= DPILQR_ST_NONFINITE;
                                    double *kp = p.K + (((int64_t)problem() * T + t) * m_real + row) * n_real + col;
                                    if ((reinterpret_cast<uintptr_t>(p.K) & 15) == 0) *reinterpret_cast<double2 *>(kp) = xc[blk];
                                    else kp[0] = d0, kp[1] = d1;
                                } else if (col == AT * S) {
                                    p.d[((int64_t)problem() * T + t) * m_real + row] = d0;
                                }
                            }
                        }
                    } else if (lane < 8) {
                        double x[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) x[j] = Xc[(size_t)(8 * blk + j) * LDN + lane];
#pragma unroll
                        for (int j = 7; j >= 0; --j) {
                            x[j] *= rdiag[8 * blk + j];
#pragma unroll
                            for (int j2 = 0; j2 < j; ++j2) x[j2] = fma(-Up[(8 * blk + j) * LDF + 8 * blk + j2], x[j], x[j2]);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) Xc[(size_t)(8 * blk + j) * LDN + lane] = x[j];
                    }
                    __syncwarp();
                    if (blk > 0) {
                        double xb[2];
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) xb[ks] = Xc[(size_t)(8 * blk + 4 * ks + fc) * LDN + fr];
#pragma unroll
                        for (int b2 = 0; b2 < blk; ++b2)
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks)
                                dmma_m8n8k4(xc[b2].x, xc[b2].y, -Up[(8 * blk + 4 * ks + fc) * LDF + 8 * b2 + fr], xb[ks]);
                        __syncwarp();
                    }
                }
            }
        } else {
            // runtime sizes: one thread per right-hand side, blocked by eight rows.  The eight unknowns of a block are
            // solved in registers; the rows outside the block are updated four at a time with every load of a batch
            // issued before its first use -- the column lives in shared memory or, for large teams, in the L2 scratch,
            // and an update chained one load-FMA-store after the other costs a full round trip per entry (2 M cycles
            // per time step for 15 drones).
            for (int col = tid; col <= n; col += nthr) {
                double *__restrict__ xcol = KB + col;
                const double *__restrict__ lp = Lp;
                const double *__restrict__ up = Up;
                for (int k0 = 0; k0 < m; k0 += 8) {
                    double xv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) xv[e] = (k0 + e < m) ? -QUX[(size_t)order[k0 + e] * LDN + col] : 0.0;
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (k0 + e < m) xcol[(size_t)(k0 + e) * LDN] = xv[e];
                }
                // ---- forward: unit lower factor, lp[k * LDF + k2] = l(k2, k), k2 > k
                for (int k0 = 0; k0 < m; k0 += 8) {
                    double x[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] = (k0 + e < m) ? xcol[(size_t)(k0 + e) * LDN] : 0.0;
#pragma unroll
                    for (int e = 0; e < 8; ++e)
#pragma unroll
                        for (int f = e + 1; f < 8; ++f)
                            if (k0 + f < m) x[f] = fma(-lp[(k0 + e) * LDF + k0 + f], x[e], x[f]);
#pragma unroll
                    for (int e = 1; e < 8; ++e)
                        if (k0 + e < m) xcol[(size_t)(k0 + e) * LDN] = x[e];
                    for (int k2 = k0 + 8; k2 < m; k2 += 4) {
                        double v[4], l[4][8];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int row = min(k2 + r, m - 1);
                            v[r] = xcol[(size_t)row * LDN];
#pragma unroll
                            for (int e = 0; e < 8; ++e) l[r][e] = lp[(k0 + e) * LDF + row];
                        }
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[r] = fma(-l[r][e], x[e], v[r]);
                            if (k2 + r < m) xcol[(size_t)(k2 + r) * LDN] = v[r];
                        }
                    }
                }
                // ---- backward: upper factor, up[c * LDF + k] = u(k, c), k <= c
                for (int k0 = ((m - 1) >> 3) << 3; k0 >= 0; k0 -= 8) {
                    double x[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] = (k0 + e < m) ? xcol[(size_t)(k0 + e) * LDN] : 0.0;
#pragma unroll
                    for (int e = 7; e >= 0; --e) {
                        if (k0 + e < m) {
                            x[e] *= rdiag[k0 + e];
#pragma unroll
                            for (int f = 0; f < e; ++f) x[f] = fma(-up[(k0 + e) * LDF + k0 + f], x[e], x[f]);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (k0 + e < m) xcol[(size_t)(k0 + e) * LDN] = x[e];
                    for (int k2 = 0; k2 < k0; k2 += 4) {  // k0 is a multiple of 8: whole batches of four rows
                        double v[4], uu[4][8];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            v[r] = xcol[(size_t)(k2 + r) * LDN];
#pragma unroll
                            for (int e = 0; e < 8; ++e) uu[r][e] = (k0 + e < m) ? up[(k0 + e) * LDF + k2 + r] : 0.0;
                        }
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[r] = fma(-uu[r][e], x[e], v[r]);
                            xcol[(size_t)(k2 + r) * LDN] = v[r];
                        }
                    }
                }
            }
        }
        }
        __syncthreads();
        tick(10);
        if constexpr (kMergeE) {
            DPILQR_PHASE_IDS
            // ---- phase E in one piece (shared-memory tensor-path kernels, at most 15 column tiles): a warp owns a whole
            // column tile of Q_ux -- nobody else reads or writes those columns -- so pq = Q_ux^T d for its eight columns
            // comes first and Y = Q_uu K + 2 Q_ux then goes over the tile in place, with no block barrier between the
            // two; the spare last warp computes z = Q_uu d + Q_u.  d is read where phase D left it (column n of the K
            // buffer); K[t] and d[t] went out to HBM from phase D.  Same partial sums and the same order of additions as the vector phase of
            // the other kernels.
            constexpr int N = AT * S, M = AT * C, LD = backward_ldn(N);
            constexpr int MT = M / 8, NT = N / 8, KS = M / 4;
            static_assert(NT < 16, "phase E: the last warp is spare");
            // instrumented build: cycles from the start of the phase to the end of every warp's own work (slots 32 + warp)
            // and to the end of its pq (slots 48 + warp) of CTA 0; the timing buffer has 64 entries
            const bool etime = TIMED && p.timing != nullptr && blockIdx.x == 0;
            const long long e_start = etime ? clock64() : 0;
            if (t > 0) prefetch_record(t - 1, nthr - 1);
            if constexpr (!kKFromD) {
                // K[t] streams out to HBM (K stays in shared memory until phase A of the next step reuses the buffer)
                double *Kt = p.K + ((int64_t)problem() * T + t) * m_real * n_real;
                if ((n_real & 1) == 0 && (reinterpret_cast<uintptr_t>(p.K) & 15) == 0) {  // coalesced, two entries per access
                    for (int e = tid; e < (m_real * n_real) >> 1; e += nthr) {
                        const int k = (2 * e) / n_real, col = 2 * e - k * n_real;
                        const double2 kv = *reinterpret_cast<const double2 *>(KB + (size_t)k * LD + col);
                        if (!isfinite(kv.x) || !isfinite(kv.y)) st |= DPILQR_ST_NONFINITE;
                        *reinterpret_cast<double2 *>(Kt + 2 * e) = kv;
                    }
                } else {
                    for (int e = tid; e < m_real * n_real; e += nthr) {
                        const int k = e / n_real, col = e - k * n_real;
                        const double kv = KB[(size_t)k * LD + col];
                        if (!isfinite(kv)) st |= DPILQR_ST_NONFINITE;
                        Kt[e] = kv;
                    }
                }
                for (int k = tid; k < m_real; k += nthr) p.d[((int64_t)problem() * T + t) * m_real + k] = KB[(size_t)k * LD + N];
            }
            const int fr = lane >> 2, fc = lane & 3;  // fragment coordinates
            if (warp == nwarp - 1) {
                // z = Q_uu d + Q_u (phase F needs it for p).  Plain FMAs spread like pq: lane (fr, fc) sums the terms
                // k = fc (mod 4) of row 8 mt + fr, two shuffles add the four partial sums.  (Anything more on this warp
                // holds the phase up: its FMAs queue behind the tensor instructions of the three column-tile warps
                // that share its scheduler's FP64 pipe.)
                double q[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) q[mt] = 0.0;
                const double *ap = QUU + fr * LDQ + fc;
                const double *dp = KB + (size_t)fc * LD + N;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const double dk = dp[(size_t)4 * ks * LD];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) q[mt] = fma(ap[8 * mt * LDQ + 4 * ks], dk, q[mt]);
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    q[mt] += __shfl_xor_sync(0xffffffffu, q[mt], 1);
                    q[mt] += __shfl_xor_sync(0xffffffffu, q[mt], 2);
                    if (fc == 0) zv[8 * mt + fr] = q[mt] + Qu[8 * mt + fr];
                }
            }
            for (int nt = warp; nt < NT; nt += nwarp) {
                // pq for columns 8 nt .. 8 nt + 7: lane (fr, fc) sums the rows k = fc (mod 4) of column 8 nt + fr, one
                // term per k-step of the tile product below -- a chain of ten dependent FMAs (some 50 cycles each) that
                // would hold the warp's tensor instructions up for a thousand cycles if it came first
                const double *qp = QUX + (size_t)fc * LD + 8 * nt + fr;
                const double *dp = KB + (size_t)fc * LD + N;
                double q = 0.0;
                double2 acc[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const double2 y0 = *reinterpret_cast<const double2 *>(QUX + (size_t)(8 * mt + fr) * LD + 8 * nt + 2 * fc);
                    acc[mt] = make_double2(2.0 * y0.x, 2.0 * y0.y);
                }
                const double *ap = QUU + fr * LDQ + fc;
                const double *bp = KB + (size_t)fc * LD + 8 * nt + fr;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const double bv = bp[(size_t)4 * ks * LD];
                    q = fma(qp[(size_t)4 * ks * LD], dp[(size_t)4 * ks * LD], q);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) dmma_m8n8k4(acc[mt].x, acc[mt].y, ap[8 * mt * LDQ + 4 * ks], bv);
                }
                q += __shfl_xor_sync(0xffffffffu, q, 1);
                q += __shfl_xor_sync(0xffffffffu, q, 2);
                if (fc == 0) pq[8 * nt + fr] = q;
                if (etime && lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(p.timing) + 48 + warp, (unsigned long long)(clock64() - e_start));
                __syncwarp();  // every lane has read the column tile of Q_ux: overwrite it with Y
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                    *reinterpret_cast<double2 *>(QUX + (size_t)(8 * mt + fr) * LD + 8 * nt + 2 * fc) = acc[mt];
            }
            if (etime && lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(p.timing) + 32 + warp, (unsigned long long)(clock64() - e_start));
        } else {
        {
            DPILQR_PHASE_IDS
        for (int k = tid; k < m; k += nthr) {
            const double dk = KB[(size_t)k * LDN + n];
            dv[k] = dk;
            if (k < m_real) p.d[((int64_t)problem() * T + t) * m_real + k] = dk;
        }
        }
        __syncthreads();
        tick(4);

        {
            DPILQR_PHASE_IDS
        // The stage record of this step has been dead since the warp groups joined: fetch the next one behind phases E
        // and F.  The last thread issues the copies -- its warp has nothing to do in the vector phase below.
        if (t > 0) prefetch_record(t - 1, nthr - 1);
        // K[t] streams out to HBM on the threads that have no column in the vector phase below (K stays in shared
        // memory until phase A of the next step reuses the buffer).  (One TMA bulk store per row, issued by a warp and
        // left to complete behind phases E and F, measured slower: the issuing warp holds the phase up.)
        {
            double *Kt = p.K + ((int64_t)problem() * T + t) * m_real * n_real;
            const int first = (n + m < nthr - 64) ? n + m : 0, nw = nthr - first;
            if (tid >= first) {
                if ((n_real & 1) == 0 && (reinterpret_cast<uintptr_t>(p.K) & 15) == 0) {  // coalesced, two entries per access
                    for (int e = tid - first; e < (m_real * n_real) >> 1; e += nw) {
                        const int k = (2 * e) / n_real, col = 2 * e - k * n_real;
                        const double2 kv = *reinterpret_cast<const double2 *>(KB + (size_t)k * LDN + col);
                        if (!isfinite(kv.x) || !isfinite(kv.y)) st |= DPILQR_ST_NONFINITE;
                        *reinterpret_cast<double2 *>(Kt + 2 * e) = kv;
                    }
                } else {
                    for (int e = tid - first; e < m_real * n_real; e += nw) {
                        const int k = e / n_real, col = e - k * n_real;
                        const double kv = KB[(size_t)k * LDN + col];
                        if (!isfinite(kv)) st |= DPILQR_ST_NONFINITE;
                        Kt[e] = kv;
                    }
                }
            }
        }
        // ---- phase E: pq = Q_ux^T d and z = Q_uu d + Q_u (before Q_ux is overwritten), then Y = Q_uu K + 2 Q_ux
        for (int col = tid; col < n + m; col += nthr) {
            if (col < n) {
                double q4[4] = {0.0, 0.0, 0.0, 0.0};  // four partial sums: the 40-long chain is pure latency otherwise
                int k = 0;
                for (; k + 4 <= m; k += 4) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) q4[e] = fma(QUX[(size_t)(k + e) * LDN + col], dv[k + e], q4[e]);
                }
                for (; k < m; ++k) q4[0] = fma(QUX[(size_t)k * LDN + col], dv[k], q4[0]);
                pq[col] = (q4[0] + q4[1]) + (q4[2] + q4[3]);
            } else {
                const int k = col - n;
                double z4[4] = {0.0, 0.0, 0.0, 0.0};
                int l = 0;
                for (; l + 4 <= m; l += 4) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) z4[e] = fma(QUU[k * LDQ + l + e], dv[l + e], z4[e]);
                }
                for (; l < m; ++l) z4[0] = fma(QUU[k * LDQ + l], dv[l], z4[0]);
                zv[k] = ((z4[0] + z4[1]) + (z4[2] + z4[3])) + Qu[k];
            }
        }
        }
        __syncthreads();
        tick(6);
        {
            DPILQR_PHASE_IDS
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
                double e0 = 0.0, e1 = 0.0;  // second accumulator pair: two independent chains
#pragma unroll
                for (int ks = 0; ks < KS; ks += 2) {
                    dmma_m8n8k4(c0, c1, ap[4 * ks], bp[(size_t)4 * ks * LD]);
                    if (ks + 1 < KS) dmma_m8n8k4(e0, e1, ap[4 * ks + 4], bp[(size_t)(4 * ks + 4) * LD]);
                }
                yp[0] = c0 + e0;
                yp[1] = c1 + e1;
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
        }
        }
        __syncthreads();
        tick(7);

        {
            DPILQR_PHASE_IDS
        // ---- phase F: P <- Q_xx + 1/2 (K^T Y + Y^T K) on upper blocks; p <- Q_x + K^T z + Q_ux^T d
        const double *Y = QUX;
        if constexpr (USE_MMA) {
            constexpr int N = AT * S, M = AT * C, LD = backward_ldn(N);
            constexpr int NT = N / 8, KS = M / 4;
            constexpr int NTILES = NT * (NT + 1) / 2;
            const int fr = lane >> 2, fc = lane & 3;
            auto add_tile = [&](int tr, int tc, double c0, double c1) {  // P tile (tr, tc) += 1/2 (c0, c1)
                const int row = 8 * tr + fr, col = 8 * tc + 2 * fc;
                const int bi = row / S, bj = col / S;  // S is even here, so the pair (col, col+1) shares a block
                if (bi <= bj) {
                    double *pblk = Pb + (size_t)blk_index(bi, bj) * PBS;
                    const int rr = row - bi * S, cc = col - bj * S;
                    pblk[rr * S + cc] += 0.5 * c0;
                    pblk[rr * S + cc + 1] += 0.5 * c1;
                    // diagonal blocks are stored in full: an off-diagonal tile inside one also owns the mirrored
                    // entries (their own tile lies below the tile diagonal and is never visited)
                    if (bi == bj && tr != tc) {
                        pblk[cc * S + rr] += 0.5 * c0;
                        pblk[(cc + 1) * S + rr] += 0.5 * c1;
                    }
                }
            };
            // warps w, w+4, w+8, w+12 share a scheduler (and its FP64 pipe): when the tile count does not divide by the warp
            // count, every other slot has one tile more -- swap the slots of neighbouring warps in every other group of
            // four, so that each scheduler gets as many long slots as short ones
            const int slot = DPILQR_F_BALANCE ? (warp ^ ((warp >> 2) & 1)) : warp;
            const int q0 = (NTILES * slot) / nwarp, q1 = (NTILES * (slot + 1)) / nwarp;
            int ti = 0, rowstart = 0;  // decode q0 -> (ti, tj) in the row-major upper triangle of tiles
            while (q0 >= rowstart + (NT - ti)) { rowstart += NT - ti; ++ti; }
            int tj = ti + (q0 - rowstart);
            for (int q = q0; q < q1;) {
                // Two neighbouring tiles of a tile row share the row's operand fragments (K^T and Y^T of rows 8 ti..):
                // six shared-memory loads for four tensor instructions instead of eight, four independent chains.
                const bool two = (q + 1 < q1) && (tj + 1 < NT);
                double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
                const double *kp = KB + (size_t)fc * LD + 8 * tj + fr;
                const double *yp = Y + (size_t)fc * LD + 8 * tj + fr;
                const double *kq = KB + (size_t)fc * LD + 8 * ti + fr;
                const double *yq = Y + (size_t)fc * LD + 8 * ti + fr;
                if (two) {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const double ak = kq[(size_t)4 * ks * LD], ay = yq[(size_t)4 * ks * LD];
                        dmma_m8n8k4(c0, c1, ak, yp[(size_t)4 * ks * LD]);      // K^T Y
                        dmma_m8n8k4(e0, e1, ay, kp[(size_t)4 * ks * LD]);      // Y^T K
                        dmma_m8n8k4(g0, g1, ak, yp[(size_t)4 * ks * LD + 8]);  // same for the next column tile
                        dmma_m8n8k4(h0, h1, ay, kp[(size_t)4 * ks * LD + 8]);
                    }
                } else {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        dmma_m8n8k4(c0, c1, kq[(size_t)4 * ks * LD], yp[(size_t)4 * ks * LD]);
                        dmma_m8n8k4(e0, e1, yq[(size_t)4 * ks * LD], kp[(size_t)4 * ks * LD]);
                    }
                }
                add_tile(ti, tj, c0 + e0, c1 + e1);
                if (two) add_tile(ti, tj + 1, g0 + h0, g1 + h1);
                const int step = two ? 2 : 1;
                q += step;
                tj += step;
                if (tj >= NT) { ++ti; tj = ti; }
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
        if constexpr (!kMergeE) {
        for (int col = tid; col < n; col += nthr) {
            double a4[4] = {0.0, 0.0, 0.0, 0.0};
            int k = 0;
            for (; k + 4 <= m; k += 4) {
#pragma unroll
                for (int e = 0; e < 4; ++e) a4[e] = fma(KB[(size_t)(k + e) * LDN + col], zv[k + e], a4[e]);
            }
            for (; k < m; ++k) a4[0] = fma(KB[(size_t)k * LDN + col], zv[k], a4[0]);
            const double acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
            const double pnew = Qx[col] + acc + pq[col];
            if (!isfinite(pnew)) st |= DPILQR_ST_NONFINITE;
            pvec[col] = pnew;
        }
        }
        }
        __syncthreads();
        tick(8);
    }
    if (timing && (tid & 31) == 0) {
        for (int k = 0; k < 12; ++k) p.timing[(tid == 0 ? 0 : 12) + k] = tacc[k];
        if (tid == 0) p.timing[24] = tacc[24], p.timing[25] = tacc[25], p.timing[26] = tacc[26];
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
    if (plan.smem_bytes > 227 * 1024 && s == 12 && c == 4 && a == 16)  // the 16-agent tensor-path kernel keeps Q_uu in the scratch
        plan.smem_bytes = backward_smem(a, s, c, false, false).total_doubles * 8;
    plan.threads = 512;
    return plan;
}

// Odd Quadcopter12D teams of 5..15 run the tensor-path kernel of the next even size on stage records padded by a
// phantom agent (see backward_kernel).  The solver asks here which layout its records should have.
int backward_layout_agents(int a, int s, int c)
{
    static const bool no_pad = getenv("DPILQR_BACKWARD_NO_PAD") != nullptr;  // experiments / cross-checks
    if (!no_pad && s == 12 && c == 4 && (a & 1) && a >= 5 && a <= 15) return a + 1;
    return a;
}

int64_t backward_scratch_doubles(int n_problems, int a, int s, int c)
{
    const int al = backward_layout_agents(a, s, c);
    const BackwardPlan plan = plan_backward(al, s, c);
    return plan.use_global_scratch ? (int64_t)n_problems * (int64_t)backward_scratch_per_cta(al * c, al * s) : 0;
}

// TIMED: the instrumented build of the kernel (per-phase cycle counters); the product build carries no timing code
template <int S, int C, int AT, bool GLOBAL, bool TIMED = false>
static int launch_typed(const BackwardParams &p, int n_blocks, const BackwardPlan &plan, cudaStream_t stream)
{
    auto kernel = backward_kernel<S, C, AT, GLOBAL, TIMED>;
    // per launch: the attribute belongs to the device / context, and callers may drive several devices and threads
    DPILQR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kernel<<<n_blocks, plan.threads, plan.smem_bytes, stream>>>(p);
    DPILQR_CUDA(cudaGetLastError());
    return DPILQR_OK;
}

template <int S, int C>
static int launch_generic(const BackwardParams &p, int n_blocks, const BackwardPlan &plan, cudaStream_t stream)
{
    if constexpr (S == 12 && C == 4) {  // instrumented builds of the generic paths (tools/backward_phases.py)
        if (p.timing != nullptr) {
            if (plan.use_global_scratch) return launch_typed<S, C, 0, true, true>(p, n_blocks, plan, stream);
            return launch_typed<S, C, 0, false, true>(p, n_blocks, plan, stream);
        }
    }
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
    const int s = bt.s, c = bt.c;
    const bool padded = p.a_layout > bt.n_agents;  // records carry a phantom agent: tensor-path kernel of the even size
    const int a = padded ? p.a_layout : bt.n_agents;
    if (padded && !(s == 12 && c == 4 && a == bt.n_agents + 1 && (a & 1) == 0 && a >= 6 && a <= 16)) {
        set_error("backward kernel: no padded path for %d agents in a layout of %d", bt.n_agents, p.a_layout);
        return DPILQR_E_INVALID;
    }
    // small problems (DP-iLQR neighbourhoods, small teams): several problems per SM (backward_small.cu)
    static const bool force_big = getenv("DPILQR_BACKWARD_FORCE_BIG") != nullptr;  // experiments / cross-checks
    static const bool no_warp = getenv("DPILQR_BACKWARD_NO_WARP") != nullptr;            // experiments / cross-checks
    // tiny problems (one or two drones, up to four planar agents): one warp per problem (backward_warp.cu)
    if (!padded && !force_big && !no_warp && p.timing == nullptr && backward_warp_applies(a, s, c)) return launch_backward_warp(p, n_blocks, stream);
    if (!padded && !force_big && p.timing == nullptr && backward_small_applies(a, s, c)) return launch_backward_small(p, n_blocks, stream);
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
        // tensor-path instantiations: even team sizes (an 8-row tile of the control rows is two agents), odd teams
        // padded by a phantom agent; 12 to 16 agents keep Q_ux, K and the LU factors (16: Q_uu too) in the
        // L2-resident scratch
        static const bool force_generic = getenv("DPILQR_BACKWARD_FORCE_GENERIC") != nullptr;  // experiments / cross-checks
        if (!force_generic || padded) {
            if (a == 10 && !plan.use_global_scratch) {
                if (p.timing != nullptr) return launch_typed<12, 4, 10, false, true>(p, n_blocks, plan, stream);
                return launch_typed<12, 4, 10, false>(p, n_blocks, plan, stream);
            }
            if (a == 6 && !plan.use_global_scratch) return launch_typed<12, 4, 6, false>(p, n_blocks, plan, stream);
            if (a == 8 && !plan.use_global_scratch) return launch_typed<12, 4, 8, false>(p, n_blocks, plan, stream);
            if (a == 12 && plan.use_global_scratch) return launch_typed<12, 4, 12, true>(p, n_blocks, plan, stream);
            if (a == 14 && plan.use_global_scratch) return launch_typed<12, 4, 14, true>(p, n_blocks, plan, stream);
            if (a == 16 && plan.use_global_scratch) return launch_typed<12, 4, 16, true>(p, n_blocks, plan, stream);
        }
        if (padded) {
            set_error("backward kernel: padded layout of %d agents has no tensor-path instantiation", a);
            return DPILQR_E_UNSUPPORTED;
        }
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
