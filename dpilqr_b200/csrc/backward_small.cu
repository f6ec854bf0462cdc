// backward_small.cu -- Kernel 3s: backward Riccati recursion for SMALL problems (joint state n <= 64, joint control
// m <= 32), 256 threads per problem, several problems per SM (sm_100a, FP64).
//
// Same recursion as backward.cu (reference ilqrSolver._backward_pass, control.py:116-148); this is the kernel of the
// DP-iLQR sub-problems -- nine in ten neighbourhoods hold one to three agents -- and of small
// teams in general, for which one 512-thread CTA per SM (backward.cu) leaves the machine idle.  Everything of a
// problem lives in shared memory for the whole recursion (P dense, updated in place; 5 to 70 kB per problem, so 3 to 16
// problems share an SM); per time step
//   A   S = B^T (P + mu I) (block-diagonal B: C rows per agent),  Q_ux = S A,  Q_uu = L_uu + S B,  Q_u, Q_x
//   B   Q_xx = L_xx + A^T P A on the upper blocks, one thread per (block, column), in place (whole blocks per round)
//   C   LU of Q_uu with partial pivoting by ONE warp, a row per lane (m <= 32): exact arg-max pivot (two redux steps
//       on the magnitude's bit pattern) -- runs beside phase B
//   D   K = -Q_uu^{-1} Q_ux, d = -Q_uu^{-1} Q_u: one thread per right-hand side, four rows at a time in registers
//   E   Y = Q_uu K + 2 Q_ux,  z = Q_uu d + Q_u,  pq = Q_ux^T d
//   F   P <- Q_xx + 1/2 (K^T Y + Y^T K) on 4x4 register tiles of the upper triangle, mirrored;  p <- Q_x + K^T z + pq
// (== the reference's symmetrised Q_xx + K^T Q_uu K + K^T Q_ux + Q_ux^T K).  Stage records arrive by cp.async a step
// ahead; K, d stream out coalesced.
#include "kernels.cuh"
#include "lu.cuh"

namespace dpilqr {

struct SmallSmem {
    int P0, QUX, KB, QUU, W, rec, pvec, Qx, pq, Qu, dv, zv, order, total_doubles, ldp, ldn, ldq;
};

__host__ __device__ inline SmallSmem small_smem(int a, int s, int c)
{
    const int n = a * s, m = a * c, pairs = a * (a - 1) / 2;
    auto even = [](int v) { return (v + 1) & ~1; };
    SmallSmem L{};
    L.ldp = even(n) + 2;      // row stride of P (even: 16-byte rows)
    L.ldn = even(n + 1) + 2;  // row stride of Q_ux / K: column n carries Q_u / d
    L.ldq = m | 1;            // row stride of Q_uu and of the LU factors: odd, so a column walks distinct banks
    int off = 0;
    L.P0 = off;   off += n * L.ldp;  // P in place: Q_xx overwrites it block by block (round 2: was double-buffered)
    L.QUX = off;  off += m * L.ldn;
    L.KB = off;   off += m * L.ldn;
    L.QUU = off;  off += even(m * L.ldq);
    L.W = off;    off += even(m * L.ldq);
    L.rec = off;  off += even(stage_layout(a, s, c).stride);
    L.pvec = off; off += even(n);
    L.Qx = off;   off += even(n);
    L.pq = off;   off += even(n);
    L.Qu = off;   off += even(m);
    L.dv = off;   off += even(m);
    L.zv = off;   off += even(m);
    L.order = off; off += even((m + 1) / 2 + 1);
    L.total_doubles = off + 2;
    (void)pairs;
    return L;
}

#ifndef DPILQR_SMALL_THREADS
#define DPILQR_SMALL_THREADS 128
#endif
constexpr int kSmallThreads = DPILQR_SMALL_THREADS;

template <int S, int C, int MMAX>
__global__ void __launch_bounds__(kSmallThreads) backward_small_kernel(const BackwardParams p)
{
    extern __shared__ __align__(16) double smem[];
    const Batch &bt = p.batch;
    if (p.n_active != nullptr && (int)blockIdx.x >= *p.n_active) return;
    const int b = p.active ? p.active[blockIdx.x] : blockIdx.x;
    const int a = bt.n_agents, T = bt.horizon;
    const int n = a * S, m = a * C;
    const int nblk = a * (a + 1) / 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nthr = kSmallThreads;
    const StageLayout L = stage_layout(a, S, C);
    const SmallSmem SM = small_smem(a, S, C);
    const int LDP = SM.ldp, LDN = SM.ldn, LDQ = SM.ldq;
    double *Pcur = smem + SM.P0, *Pnxt = Pcur;  // one buffer: phase B reads a block before it overwrites it
    double *QUX = smem + SM.QUX;  // [m][LDN]  S, then Q_ux (col n: Q_u), then Y
    double *KB = smem + SM.KB;    // [m][LDN]  S staging, then K (col n: d)
    double *QUU = smem + SM.QUU;  // [m][LDQ]
    double *W = smem + SM.W;      // [m][LDQ]  LU factors: multipliers below, U in the pivot rows
    double *rec = smem + SM.rec;
    double *sA = rec + L.offA, *sB = rec + L.offB, *sLx = rec + L.offLx, *sLu = rec + L.offLu, *sHd = rec + L.offHd, *sHo = rec + L.offHo;
    double *pvec = smem + SM.pvec, *Qx = smem + SM.Qx, *pq = smem + SM.pq;
    double *Qu = smem + SM.Qu, *dv = smem + SM.dv, *zv = smem + SM.zv;
    int *order = reinterpret_cast<int *>(smem + SM.order);
    const double mu = p.mu[b];
    const double w_ref = bt.weights ? bt.weights[2 * b] : 1.0;
    const int32_t *cidx = bt.cost_idx + (int64_t)b * a;
    int st = 0;

    auto fetch_record = [&](int t) {
        const double *src = p.stage + ((int64_t)b * (T + 1) + t) * L.stride;
        for (int k = tid; k < L.stride / 2; k += nthr)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(rec + 2 * k)), "l"(src + 2 * k) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto wait_record = [] { asm volatile("cp.async.wait_all;" ::: "memory"); };

    // ---- terminal condition: p = L_x, P = L_xx at (X[T], u = 0)  (control.py:125-129)
    fetch_record(T);
    wait_record();
    __syncthreads();
    for (int k = tid; k < n * n; k += nthr) {
        const int row = k / n, col = k - row * n;
        const int i = row / S, r = row - i * S, j = col / S, cc = col - j * S;
        double v = 0.0;
        if (i == j) {
            const double *Qf = bt.Qf + (int64_t)cidx[i] * S * S;
            v = w_ref * (Qf[r * S + cc] + Qf[cc * S + r]);
            if (r < 3 && cc < 3) v += sHd[9 * i + r * 3 + cc];
        } else if (r < 3 && cc < 3) {
            v = (i < j) ? sHo[9 * pair_index(i, j, a) + r * 3 + cc] : sHo[9 * pair_index(j, i, a) + cc * 3 + r];
        }
        Pcur[row * LDP + col] = v;
    }
    for (int k = tid; k < n; k += nthr) pvec[k] = sLx[k];
    __syncthreads();
    fetch_record(T - 1);

#pragma unroll 1
    for (int t = T - 1; t >= 0; --t) {
        wait_record();
        __syncthreads();
        // ================= phase A1: S = B^T (P + mu I) -> KB (staging) =================
        // item (agent i, column col): the C rows of S that belong to agent i
        for (int it = tid; it < a * n; it += nthr) {
            const int i = it / n, col = it - i * n;
            const double *Bi = sB + i * L.strideB;
            const double *Pc = Pcur + (size_t)(i * S) * LDP + col;
            double acc[C];
#pragma unroll
            for (int g = 0; g < C; ++g) acc[g] = 0.0;
#pragma unroll
            for (int r = 0; r < S; ++r) {
                const double pv = Pc[r * LDP] + ((i * S + r == col) ? mu : 0.0);
#pragma unroll
                for (int g = 0; g < C; ++g) acc[g] = fma(Bi[r * C + g], pv, acc[g]);
            }
#pragma unroll
            for (int g = 0; g < C; ++g) KB[(i * C + g) * LDN + col] = acc[g];
        }
        __syncthreads();
        // ================= phase A2: Q_ux = S A, Q_uu = L_uu + S B, Q_u, Q_x =================
        for (int it = tid; it < m * a; it += nthr) {  // item (row of S, agent j)
            const int row = it / a, j = it - row * a;
            const double *Srow = KB + row * LDN + j * S;
            const double *Aj = sA + j * L.strideA, *Bj = sB + j * L.strideB;
            double sv[S];
#pragma unroll
            for (int r = 0; r < S; ++r) sv[r] = Srow[r];
#pragma unroll
            for (int c2 = 0; c2 < S; ++c2) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < S; ++r) acc = fma(sv[r], Aj[r * S + c2], acc);
                QUX[row * LDN + j * S + c2] = acc;  // L_ux == 0 (cost.py:91)
            }
            const int i = row / C, g = row - i * C;
#pragma unroll
            for (int g2 = 0; g2 < C; ++g2) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < S; ++r) acc = fma(sv[r], Bj[r * C + g2], acc);
                if (i == j) {
                    const double *R = bt.R + (int64_t)cidx[i] * C * C;
                    acc += w_ref * (R[g * C + g2] + R[g2 * C + g]);
                }
                QUU[row * LDQ + j * C + g2] = acc;
                W[row * LDQ + j * C + g2] = acc;
            }
            if (j == 0) {
                const double *Bi = sB + i * L.strideB;
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < S; ++r) acc = fma(Bi[r * C + g], pvec[i * S + r], acc);
                const double qu = sLu[row] + acc;
                Qu[row] = qu;
                QUX[row * LDN + n] = qu;  // Q_u rides along as right-hand side n
            }
        }
        for (int col = tid; col < n; col += nthr) {
            const int j = col / S, sg = col - j * S;
            const double *Aj = sA + j * L.strideA;
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < S; ++r) acc = fma(Aj[r * S + sg], pvec[j * S + r], acc);
            Qx[col] = sLx[col] + acc;
        }
        __syncthreads();
        // ================= phase C (warp 0) beside phase B (the other warps) =================
        if (warp == 0) {
            // LU of Q_uu (in W) with partial pivoting, a row per lane.  Rows never move: a used pivot row is marked and
            // its index recorded in order[k] (reference: dgesv behind np.linalg.solve, control.py:141).
            bool used = lane >= m;
            double *myrow = W + lane * LDQ;
            for (int k = 0; k < m; ++k) {
                const double wk = (lane < m) ? myrow[k] : 0.0;
                // exact arg-max of |w[k]| over the unused rows: two 32-bit reductions on the bit pattern
                const double av = fabs(wk);
                const unsigned long long key = (used || !(av == av)) ? 0ull : (unsigned long long)__double_as_longlong(av) + 1ull;
                const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
                const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
                const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
                const unsigned win = __ballot_sync(0xffffffffu, !used && hi == mhi && lo == mlo);
                // (all keys zero: every remaining entry of the column is NaN -- take the first unused row)
                const unsigned cand = win ? win : __ballot_sync(0xffffffffu, !used);
                const int pr = __ffs(cand) - 1;
                if (lane == 0) order[k] = pr;
                const double piv = __shfl_sync(0xffffffffu, wk, pr);
                const double rinv = 1.0 / piv;  // dgetf2 scales the column by the reciprocal pivot
                if (!used && lane != pr) {
                    const double l = wk * rinv;
                    myrow[k] = l;
                    const double *prow = W + pr * LDQ;
                    for (int cc = k + 1; cc < m; ++cc) myrow[cc] = fma(-l, prow[cc], myrow[cc]);
                }
                if (lane == pr) used = true;
                __syncwarp();
            }
        } else {
            // phase B: Q_xx = L_xx + A^T P A on the upper blocks, IN PLACE (mirrored below the diagonal).  A round takes
            // whole blocks (one thread per block column): every thread of the round reads its block (V = P_ij A_j, a
            // column in registers), the group meets at a named barrier, then the columns of Q_xx overwrite the block
            // (and its mirror image, which phase B never reads).
            constexpr int gn = nthr - 32;
            constexpr int blocks_per_round = gn / S;
            const int gt = tid - 32;
            const bool worker = gt < blocks_per_round * S;
            for (int blk0 = 0; blk0 < nblk; blk0 += blocks_per_round) {
                const int blk = blk0 + gt / S, sg = gt % S;
                const bool live = worker && blk < nblk;
                int i = 0, rem = live ? blk : 0;
                while (rem >= a - i) { rem -= a - i; ++i; }
                const int j = i + rem;
                double v[S];
                if (live) {
                    const double *Pblk = Pcur + (size_t)(i * S) * LDP + j * S;
                    const double *Aj = sA + j * L.strideA;
                    double acol[S];
#pragma unroll
                    for (int q = 0; q < S; ++q) acol[q] = Aj[q * S + sg];
#pragma unroll
                    for (int r = 0; r < S; ++r) {
                        double acc = 0.0;
#pragma unroll
                        for (int q = 0; q < S; ++q) acc = fma(Pblk[r * LDP + q], acol[q], acc);
                        v[r] = acc;
                    }
                }
                asm volatile("bar.sync 2, %0;" ::"n"(gn) : "memory");  // every block of the round has been read
                if (live) {
                    const double *Ai = sA + i * L.strideA;
#pragma unroll
                    for (int r = 0; r < S; ++r) {
                        double acc = 0.0;
#pragma unroll
                        for (int q = 0; q < S; ++q) acc = fma(Ai[q * S + r], v[q], acc);
                        double lxx = 0.0;
                        if (i == j) {
                            const double *Q = bt.Q + (int64_t)cidx[i] * S * S;
                            lxx = w_ref * (Q[r * S + sg] + Q[sg * S + r]);
                            if (r < 3 && sg < 3) lxx += sHd[9 * i + r * 3 + sg];
                        } else if (r < 3 && sg < 3) {
                            lxx = sHo[9 * pair_index(i, j, a) + r * 3 + sg];
                        }
                        const double q = lxx + acc;
                        Pnxt[(size_t)(i * S + r) * LDP + j * S + sg] = q;
                        if (i != j) Pnxt[(size_t)(j * S + sg) * LDP + i * S + r] = q;
                    }
                }
                // (no barrier before the next round: it reads upper blocks nobody has written -- the mirrors land below the diagonal)
            }
        }
        __syncthreads();
        if (t > 0) fetch_record(t - 1);  // the record of this step is consumed
        // ================= phase D: K = -Q_uu^{-1} [Q_ux | Q_u], one thread per right-hand side =================
        // The column lives in shared memory (KB[:, col]); up to FOUR lanes share a column: each solves the 4x4 diagonal block of
        // the step in registers (redundantly) and takes every fourth row of the update below / above it, so the
        // dependent chain through shared memory is m/4 blocks long and each link a quarter of the rows.
        // lanes per column: as many (1, 2 or 4) as let all the right-hand sides go in one pass
        const int lpc = (m <= 8) ? 1 : ((n + 1) * 4 <= nthr) ? 4 : ((n + 1) * 2 <= nthr) ? 2 : 1;
        for (int cbase = 0; cbase <= n; cbase += nthr / lpc) {  // uniform trip count: every lane reaches the warp barriers
            const int col = cbase + tid / lpc, sub = tid % lpc;
            const bool live = col <= n;
            double *xc = KB + (live ? col : 0);
            if (live)
                for (int k = sub; k < m; k += lpc) xc[k * LDN] = -QUX[order[k] * LDN + col];
            __syncwarp();
            for (int k0 = 0; k0 < m; k0 += 4) {  // forward: unit lower factor, rows in pivot order
                double x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = (live && k0 + e < m) ? xc[(k0 + e) * LDN] : 0.0;
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int f = e + 1; f < 4; ++f)
                        if (k0 + f < m) x[f] = fma(-W[order[k0 + f] * LDQ + k0 + e], x[e], x[f]);
                __syncwarp();  // every lane of the quad has read the block before it is rewritten
                if (live) {
#pragma unroll
                    for (int e = 1; e < 4; ++e)
                        if (e % lpc == sub && k0 + e < m) xc[(k0 + e) * LDN] = x[e];
                    for (int k2 = k0 + 4 + sub; k2 < m; k2 += lpc) {
                        const double *lrow = W + order[k2] * LDQ + k0;
                        double v = xc[k2 * LDN];
#pragma unroll
                        for (int e = 0; e < 4; ++e) v = fma(-lrow[e], x[e], v);
                        xc[k2 * LDN] = v;
                    }
                }
                __syncwarp();
            }
            for (int k0 = ((m - 1) >> 2) << 2; k0 >= 0; k0 -= 4) {  // backward: upper factor
                double x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = (live && k0 + e < m) ? xc[(k0 + e) * LDN] : 0.0;
#pragma unroll
                for (int e = 3; e >= 0; --e) {
                    if (k0 + e < m) {
                        const double *urow = W + order[k0 + e] * LDQ;
                        if (col == 0 && sub == 0 && urow[k0 + e] == 0.0) st |= DPILQR_ST_SINGULAR;  // exact zero pivot: dgesv's info > 0
                        x[e] = x[e] / urow[k0 + e];
#pragma unroll
                        for (int f = 0; f < e; ++f) x[f] = fma(-W[order[k0 + f] * LDQ + k0 + e], x[e], x[f]);
                    }
                }
                __syncwarp();
                if (live) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (e % lpc == sub && k0 + e < m) {
                            xc[(k0 + e) * LDN] = x[e];
                            if (!isfinite(x[e])) st |= DPILQR_ST_NONFINITE;
                        }
                    }
                    for (int k2 = sub; k2 < k0; k2 += lpc) {
                        const double *urow = W + order[k2] * LDQ + k0;
                        double v = xc[k2 * LDN];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (k0 + e < m) v = fma(-urow[e], x[e], v);
                        xc[k2 * LDN] = v;
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- stream K[t], d[t] out; d into shared memory
        {
            double *Kt = p.K + ((int64_t)b * T + t) * m * n;
            for (int e = tid; e < m * n; e += nthr) {
                const int k = e / n, col = e - k * n;
                Kt[e] = KB[k * LDN + col];
            }
            for (int k = tid; k < m; k += nthr) {
                const double dk = KB[k * LDN + n];
                dv[k] = dk;
                p.d[((int64_t)b * T + t) * m + k] = dk;
            }
        }
        __syncthreads();
        // ================= phase E: pq = Q_ux^T d, z = Q_uu d + Q_u, Y = Q_uu K + 2 Q_ux (in place over Q_ux) ==========
        for (int col = tid; col < n + m; col += nthr) {
            if (col < n) {
                double acc = 0.0;
                for (int k = 0; k < m; ++k) acc = fma(QUX[k * LDN + col], dv[k], acc);
                pq[col] = acc;
            } else {
                const int k = col - n;
                double acc = 0.0;
                for (int l = 0; l < m; ++l) acc = fma(QUU[k * LDQ + l], dv[l], acc);
                zv[k] = acc + Qu[k];
            }
        }
        __syncthreads();
        // Y in 1 x 4 register tiles (row k, four columns), in place over Q_ux: a tile reads K, Q_uu and its own four
        // entries of Q_ux only
        {
            const int n4 = (n + 3) >> 2;
            for (int it = tid; it < m * n4; it += nthr) {
                const int k = it / n4, c0 = 4 * (it - k * n4);
                double y[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) y[e] = (c0 + e < n) ? 2.0 * QUX[k * LDN + c0 + e] : 0.0;
                const double *qrow = QUU + k * LDQ;
                for (int l = 0; l < m; ++l) {
                    const double q = qrow[l];
                    const double *kr = KB + l * LDN + c0;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (c0 + e < n) y[e] = fma(q, kr[e], y[e]);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (c0 + e < n) QUX[k * LDN + c0 + e] = y[e];
            }
        }
        __syncthreads();
        // ================= phase F: P <- Q_xx + 1/2 (K^T Y + Y^T K), p <- Q_x + K^T z + pq =================
        {
            const double *Y = QUX;
            const int nt = (n + 3) >> 2;  // 4x4 tiles; upper triangle of tiles
            const int ntiles = nt * (nt + 1) / 2;
            for (int tile = tid; tile < ntiles; tile += nthr) {
                int ti = 0, rem = tile;
                while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
                const int tj = ti + rem;
                const int r0 = 4 * ti, c0 = 4 * tj;
                double acc[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) acc[r][cc] = 0.0;
                const bool full = (r0 + 4 <= n) && (c0 + 4 <= n);
                for (int k = 0; k < m; ++k) {
                    double ki[4], yi[4], kj[4], yj[4];
                    if (full) {  // 16-byte loads: the rows of K and Y are 16-byte aligned, tiles start at multiples of four
                        const double2 a0 = *reinterpret_cast<const double2 *>(KB + k * LDN + r0), a1 = *reinterpret_cast<const double2 *>(KB + k * LDN + r0 + 2);
                        const double2 b0 = *reinterpret_cast<const double2 *>(Y + k * LDN + r0), b1 = *reinterpret_cast<const double2 *>(Y + k * LDN + r0 + 2);
                        const double2 c0v = *reinterpret_cast<const double2 *>(KB + k * LDN + c0), c1v = *reinterpret_cast<const double2 *>(KB + k * LDN + c0 + 2);
                        const double2 d0 = *reinterpret_cast<const double2 *>(Y + k * LDN + c0), d1 = *reinterpret_cast<const double2 *>(Y + k * LDN + c0 + 2);
                        ki[0] = a0.x; ki[1] = a0.y; ki[2] = a1.x; ki[3] = a1.y;
                        yi[0] = b0.x; yi[1] = b0.y; yi[2] = b1.x; yi[3] = b1.y;
                        kj[0] = c0v.x; kj[1] = c0v.y; kj[2] = c1v.x; kj[3] = c1v.y;
                        yj[0] = d0.x; yj[1] = d0.y; yj[2] = d1.x; yj[3] = d1.y;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            ki[e] = (r0 + e < n) ? KB[k * LDN + r0 + e] : 0.0;
                            yi[e] = (r0 + e < n) ? Y[k * LDN + r0 + e] : 0.0;
                            kj[e] = (c0 + e < n) ? KB[k * LDN + c0 + e] : 0.0;
                            yj[e] = (c0 + e < n) ? Y[k * LDN + c0 + e] : 0.0;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) acc[r][cc] = fma(ki[r], yj[cc], fma(yi[r], kj[cc], acc[r][cc]));
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int row = r0 + r, col = c0 + cc;
                        if (row < n && col < n && row <= col) {
                            const double v = Pnxt[(size_t)row * LDP + col] + 0.5 * acc[r][cc];
                            Pnxt[(size_t)row * LDP + col] = v;
                            Pnxt[(size_t)col * LDP + row] = v;
                        }
                    }
            }
            for (int col = tid; col < n; col += nthr) {
                double acc = 0.0;
                for (int k = 0; k < m; ++k) acc = fma(KB[k * LDN + col], zv[k], acc);
                const double pnew = Qx[col] + acc + pq[col];
                if (!isfinite(pnew)) st |= DPILQR_ST_NONFINITE;
                pvec[col] = pnew;
            }
        }
        // (the barrier at the top of the next step orders phase F before its readers)
    }
    if (st != 0 && p.status) atomicOr(p.status + b, st);
}

template <int S, int C>
int launch_small_sc(const BackwardParams &p, int n_blocks, cudaStream_t stream)
{
    const int a = p.batch.n_agents, m = a * C;
    const size_t smem = (size_t)small_smem(a, S, C).total_doubles * 8;
    auto go = [&](auto kernel) -> int {
        if (smem > 48 * 1024) DPILQR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))  /* the ceiling, not this launch's need: concurrent callers must not lower it under each other */;
        kernel<<<n_blocks, kSmallThreads, smem, stream>>>(p);
        DPILQR_CUDA(cudaGetLastError());
        return DPILQR_OK;
    };
    if (m <= 8) return go(backward_small_kernel<S, C, 8>);
    if (m <= 16) return go(backward_small_kernel<S, C, 16>);
    if (m <= 24) return go(backward_small_kernel<S, C, 24>);
    return go(backward_small_kernel<S, C, 32>);
}

#ifdef DPILQR_SMALL_S
// one translation unit per (S, C) size class: -DDPILQR_SMALL_S=.. -DDPILQR_SMALL_C=..
template int launch_small_sc<DPILQR_SMALL_S, DPILQR_SMALL_C>(const BackwardParams &, int, cudaStream_t);
#else
extern template int launch_small_sc<12, 4>(const BackwardParams &, int, cudaStream_t);
extern template int launch_small_sc<6, 3>(const BackwardParams &, int, cudaStream_t);
extern template int launch_small_sc<4, 2>(const BackwardParams &, int, cudaStream_t);
extern template int launch_small_sc<3, 2>(const BackwardParams &, int, cudaStream_t);
extern template int launch_small_sc<5, 2>(const BackwardParams &, int, cudaStream_t);

bool backward_small_applies(int a, int s, int c)
{
    return a * s <= 64 && a * c <= 32 && small_smem(a, s, c).total_doubles * 8 <= 110 * 1024;
}

int launch_backward_small(const BackwardParams &p, int n_blocks, cudaStream_t stream)
{
    const int s = p.batch.s, c = p.batch.c;
    if (s == 12 && c == 4) return launch_small_sc<12, 4>(p, n_blocks, stream);
    if (s == 6 && c == 3) return launch_small_sc<6, 3>(p, n_blocks, stream);
    if (s == 4 && c == 2) return launch_small_sc<4, 2>(p, n_blocks, stream);
    if (s == 3 && c == 2) return launch_small_sc<3, 2>(p, n_blocks, stream);
    if (s == 5 && c == 2) return launch_small_sc<5, 2>(p, n_blocks, stream);
    set_error("backward kernel: unsupported per-agent dimensions (%d, %d)", s, c);
    return DPILQR_E_UNSUPPORTED;
}
#endif

}  // namespace dpilqr
