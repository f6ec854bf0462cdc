// backward_warp.cu -- Kernel 3w: backward Riccati recursion for TINY problems (joint state n <= 32, joint control
// m <= 16): ONE WARP PER PROBLEM, no block barrier anywhere (sm_100a, FP64).
//
// Same recursion as backward.cu / backward_small.cu (reference ilqrSolver._backward_pass, control.py:116-148).  Nine in
// ten DP-iLQR sub-problems are one or two agents (neighbourhood sizes 1 and 2); a CTA per problem
// spends its time in barriers there.  Here lane j owns COLUMN j of everything that has n columns (P, Q_ux, K, Y, Q_xx);
// the matrices of a problem live in the warp's private slice of shared memory, the lane's own column of the working
// matrices in registers, and the phases of a step are separated by __syncwarp() only.  A CTA is four independent
// warps; an SM holds up to twenty of them, whose latency chains (LU, substitutions) interleave.  Per time step:
//   A   S = B^T (P + mu I),  Q_ux = S A,  Q_uu = L_uu + S B,  Q_u, Q_x
//   C   LU of Q_uu with partial pivoting, a row per lane IN REGISTERS, pivot row broadcast by shuffles (exact arg-max)
//   B   Q_xx = L_xx + A^T P A, a column per lane, block row by block row (A is block diagonal)
//   D   K = -Q_uu^{-1} [Q_ux | Q_u]: a right-hand side per lane, the column in registers
//   E   Y = Q_uu K + 2 Q_ux,  z = Q_uu d + Q_u,  pq = Q_ux^T d
//   F   T = Q_xx + 1/2 (K^T Y + Y^T K),  P <- (T + T^T)/2 (control.py:146-147),  p <- Q_x + K^T z + pq
// The arithmetic of every entry follows backward_small.cu operation for operation (except that both triangles of P
// are computed and averaged instead of one being mirrored).
#include "kernels.cuh"

namespace dpilqr {

constexpr int kWarpKernelWarps = 4;

struct WarpSmem {
    int P, Tm, rec, QUX, KB, QUU, W, RR, pvec, Qx, pq, zv, order, total_doubles, ldp, ldn, ldq;
};

__host__ __device__ inline WarpSmem warp_smem(int a, int s, int c)
{
    const int n = a * s, m = a * c;
    auto even = [](int v) { return (v + 1) & ~1; };
    WarpSmem L{};
    L.ldp = n | 1;        // odd strides: a column walks distinct banks
    L.ldn = (n + 1) | 1;  // Q_ux / K / Y rows: column n carries Q_u / d
    L.ldq = m | 1;
    int off = 0;
    L.rec = off;   off += even(stage_layout(a, s, c).stride);  // first: 16-byte aligned for cp.async
    L.P = off;     off += even(n * L.ldp);
    L.Tm = off;    off += even(n * L.ldp);
    L.QUX = off;   off += even(m * L.ldn);
    L.KB = off;    off += even(m * L.ldn);
    L.QUU = off;   off += even(m * L.ldq);
    L.W = off;     off += even(m * L.ldq);
    L.RR = off;    off += even(a * c * c);
    L.pvec = off;  off += even(n);
    L.Qx = off;    off += even(n);
    L.pq = off;    off += even(n);
    L.zv = off;    off += even(m);
    L.order = off; off += even((m + 1) / 2 + 1);
    L.total_doubles = off;
    return L;
}

template <int S, int C, int A>
__global__ void __launch_bounds__(32 * kWarpKernelWarps, 4) backward_warp_kernel(const BackwardParams p)
{
    constexpr int MMAX = A * C;
    extern __shared__ __align__(16) double smem_all[];
    const Batch &bt = p.batch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * kWarpKernelWarps + warp;
    const int count = p.n_active ? min(*p.n_active, p.n_launch) : p.n_launch;
    if (slot >= count) return;  // whole warps leave: nothing below synchronises across warps
    const int b = p.active ? p.active[slot] : slot;
    constexpr int a = A, n = A * S, m = A * C;  // compile-time sizes: static strides, unrolled loops
    const int T = bt.horizon;
    const StageLayout L = stage_layout(a, S, C);
    const WarpSmem SM = warp_smem(a, S, C);
    double *smem = smem_all + (size_t)warp * SM.total_doubles;
    constexpr int LDP = n | 1, LDN = (n + 1) | 1, LDQ = m | 1;
    double *P = smem + SM.P, *Tm = smem + SM.Tm, *rec = smem + SM.rec;
    double *QUX = smem + SM.QUX, *KB = smem + SM.KB, *Ysm = QUX, *QUU = smem + SM.QUU, *W = smem + SM.W;
    double *RR = smem + SM.RR;
    double *pvec = smem + SM.pvec, *Qx = smem + SM.Qx, *pq = smem + SM.pq, *zv = smem + SM.zv;
    int *order = reinterpret_cast<int *>(smem + SM.order);
    const double *sA = rec + L.offA, *sB = rec + L.offB, *sLx = rec + L.offLx, *sLu = rec + L.offLu, *sHd = rec + L.offHd, *sHo = rec + L.offHo;
    const double mu = p.mu[b];
    const double w_ref = bt.weights ? bt.weights[2 * b] : 1.0;
    const int32_t *cidx = bt.cost_idx + (int64_t)b * a;
    constexpr unsigned FULL = 0xffffffffu;
    int st = 0;

    const bool act = lane < n;           // this lane owns column `lane`
    const int col = act ? lane : 0;
    const int cj = col / S, cc = col - cj * S;  // agent and state index of the column
    const double *Qc = bt.Q + (int64_t)cidx[cj] * S * S;  // reference-cost Hessian of the column's agent

    auto fetch_record = [&](int t) {
        const double *src = p.stage + ((int64_t)b * (T + 1) + t) * L.stride;
        for (int k = lane; k < L.stride / 2; k += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(rec + 2 * k)), "l"(src + 2 * k) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto wait_record = [] { asm volatile("cp.async.wait_all;" ::: "memory"); };

    // ---- constants of the problem: w (R + R^T) per agent
    fetch_record(T);
    for (int k = lane; k < a * C * C; k += 32) {
        const int i = k / (C * C), e = k - i * (C * C), g = e / C, g2 = e - g * C;
        const double *R = bt.R + (int64_t)cidx[i] * C * C;
        RR[k] = w_ref * (R[g * C + g2] + R[g2 * C + g]);
    }
    wait_record();
    __syncwarp();
    // ---- terminal condition: p = L_x, P = L_xx at (X[T], u = 0)  (control.py:125-129)
    if (act) {
        const double *Qf = bt.Qf + (int64_t)cidx[cj] * S * S;
        for (int row = 0; row < n; ++row) {
            const int i = row / S, r = row - i * S;
            double v = 0.0;
            if (i == cj) {
                v = w_ref * (Qf[r * S + cc] + Qf[cc * S + r]);
                if (r < 3 && cc < 3) v += sHd[9 * i + r * 3 + cc];
            } else if (r < 3 && cc < 3) {
                v = (i < cj) ? sHo[9 * pair_index(i, cj, a) + r * 3 + cc] : sHo[9 * pair_index(cj, i, a) + cc * 3 + r];
            }
            P[row * LDP + col] = v;
        }
        pvec[col] = sLx[col];
    }
    __syncwarp();
    fetch_record(T - 1);

#pragma unroll 1
    for (int t = T - 1; t >= 0; --t) {
        wait_record();
        __syncwarp();
        // ================= phase A1: S = B^T (P + mu I) -> KB (staging), column `col` =================
        if (act) {
#pragma unroll
            for (int i = 0; i < a; ++i) {
                const double *Bi = sB + i * L.strideB;
                const double *Pc = P + (size_t)(i * S) * LDP + col;
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
        }
        __syncwarp();
        // ================= phase A2: Q_ux = S A, Q_x (column per lane); Q_uu = L_uu + S B, Q_u =================
        double acol[S];  // column cc of A_cj: lives through phase B
        {
            const double *Aj = sA + cj * L.strideA;
#pragma unroll
            for (int r = 0; r < S; ++r) acol[r] = Aj[r * S + cc];
        }
        if (act) {
#pragma unroll
            for (int row = 0; row < m; ++row) {
                const double *Srow = KB + row * LDN + cj * S;
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < S; ++r) acc = fma(Srow[r], acol[r], acc);
                QUX[row * LDN + col] = acc;  // L_ux == 0 (cost.py:91)
            }
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < S; ++r) acc = fma(acol[r], pvec[cj * S + r], acc);
            Qx[col] = sLx[col] + acc;
        }
        for (int item = lane; item < m * m; item += 32) {
            const int row = item / m, c2 = item - row * m;
            const int j2 = c2 / C, g2 = c2 - j2 * C;
            const double *Srow = KB + row * LDN + j2 * S;
            const double *Bj = sB + j2 * L.strideB;
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < S; ++r) acc = fma(Srow[r], Bj[r * C + g2], acc);
            const int i = row / C, g = row - i * C;
            if (i == j2) acc += RR[i * C * C + g * C + g2];
            QUU[row * LDQ + c2] = acc;
        }
        if (lane < m) {
            const int i = lane / C, g = lane - i * C;
            const double *Bi = sB + i * L.strideB;
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < S; ++r) acc = fma(Bi[r * C + g], pvec[i * S + r], acc);
            QUX[lane * LDN + n] = sLu[lane] + acc;  // Q_u rides along as right-hand side n
        }
        __syncwarp();
        // ================= phase C: LU of Q_uu, row `lane` in registers =================
        {
            double w[MMAX];
#pragma unroll
            for (int k = 0; k < MMAX; ++k) w[k] = (lane < m && k < m) ? QUU[lane * LDQ + k] : 0.0;
            bool used = lane >= m;
#pragma unroll
            for (int k = 0; k < MMAX; ++k) {
                if (k < m) {  // warp-uniform
                    // exact arg-max of |w[k]| over the unused rows: two 32-bit reductions on the bit pattern
                    const double av = fabs(w[k]);
                    const unsigned long long key = (used || !(av == av)) ? 0ull : (unsigned long long)__double_as_longlong(av) + 1ull;
                    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
                    const unsigned mhi = __reduce_max_sync(FULL, hi);
                    const unsigned mlo = __reduce_max_sync(FULL, hi == mhi ? lo : 0u);
                    const unsigned win = __ballot_sync(FULL, !used && hi == mhi && lo == mlo);
                    // (all keys zero: every remaining entry of the column is NaN -- take the first unused row)
                    const unsigned cand = win ? win : __ballot_sync(FULL, !used);
                    const int pr = __ffs(cand) - 1;
                    if (lane == 0) order[k] = pr;
                    const double piv = __shfl_sync(FULL, w[k], pr);
                    if (piv == 0.0) st |= DPILQR_ST_SINGULAR;  // exact zero pivot: dgesv's info > 0
                    const double rinv = 1.0 / piv;  // dgetf2 scales the column by the reciprocal pivot
                    const bool elim = !used && lane != pr;
                    const double l = w[k] * rinv;
                    if (elim) w[k] = l;
#pragma unroll
                    for (int c2 = k + 1; c2 < MMAX; ++c2) {
                        if (c2 < m) {
                            const double prow = __shfl_sync(FULL, w[c2], pr);
                            if (elim) w[c2] = fma(-l, prow, w[c2]);
                        }
                    }
                    if (lane == pr) used = true;
                }
            }
            if (lane < m) {
#pragma unroll
                for (int k = 0; k < MMAX; ++k)
                    if (k < m) W[lane * LDQ + k] = w[k];
            }
        }
        // ================= phase B: Q_xx = L_xx + A^T P A -> T, column `col`, block row by block row =================
        if (act) {
#pragma unroll
            for (int i = 0; i < a; ++i) {
                const double *Pblk = P + (size_t)(i * S) * LDP + cj * S;
                const double *Ai = sA + i * L.strideA;
                double v[S];
#pragma unroll
                for (int r = 0; r < S; ++r) {
                    double acc = 0.0;
#pragma unroll
                    for (int q = 0; q < S; ++q) acc = fma(Pblk[r * LDP + q], acol[q], acc);
                    v[r] = acc;
                }
#pragma unroll
                for (int r = 0; r < S; ++r) {
                    double acc = 0.0;
#pragma unroll
                    for (int q = 0; q < S; ++q) acc = fma(Ai[q * S + r], v[q], acc);
                    double lxx = 0.0;
                    if (i == cj) {
                        lxx = w_ref * (Qc[r * S + cc] + Qc[cc * S + r]);
                        if (r < 3 && cc < 3) lxx += sHd[9 * i + r * 3 + cc];
                    } else if (r < 3 && cc < 3) {
                        lxx = (i < cj) ? sHo[9 * pair_index(i, cj, a) + r * 3 + cc] : sHo[9 * pair_index(cj, i, a) + cc * 3 + r];
                    }
                    Tm[(size_t)(i * S + r) * LDP + col] = lxx + acc;
                }
            }
        }
        __syncwarp();
        if (t > 0) fetch_record(t - 1);  // the record of this step is consumed
        // ================= phase D: K = -Q_uu^{-1} [Q_ux | Q_u], a right-hand side per lane =================
        double *Kt = p.K + ((int64_t)b * T + t) * m * n;
        for (int rhs = lane; rhs <= n; rhs += 32) {
            double x[MMAX];
#pragma unroll
            for (int k = 0; k < MMAX; ++k) x[k] = (k < m) ? -QUX[order[k] * LDN + rhs] : 0.0;
#pragma unroll
            for (int k = 0; k < MMAX; ++k) {  // forward: unit lower factor, rows in pivot order
#pragma unroll
                for (int k2 = k + 1; k2 < MMAX; ++k2)
                    if (k2 < m) x[k2] = fma(-W[order[k2] * LDQ + k], x[k], x[k2]);
            }
#pragma unroll
            for (int k = MMAX - 1; k >= 0; --k) {  // backward: upper factor
                if (k < m) {
                    x[k] = x[k] / W[order[k] * LDQ + k];
                    if (!isfinite(x[k])) st |= DPILQR_ST_NONFINITE;
#pragma unroll
                    for (int k2 = 0; k2 < k; ++k2) x[k2] = fma(-W[order[k2] * LDQ + k], x[k], x[k2]);
                }
            }
#pragma unroll
            for (int k = 0; k < MMAX; ++k) {
                if (k < m) {
                    KB[k * LDN + rhs] = x[k];
                    if (rhs < n) Kt[k * n + rhs] = x[k];
                    else p.d[((int64_t)b * T + t) * m + k] = x[k];
                }
            }
        }
        __syncwarp();
        // ================= phase E: pq = Q_ux^T d, z = Q_uu d + Q_u, Y = Q_uu K + 2 Q_ux =================
        double x[MMAX], y[MMAX];  // columns `col` of K and Y
#pragma unroll
        for (int k = 0; k < MMAX; ++k) x[k] = (k < m) ? KB[k * LDN + col] : 0.0;
        if (act) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < m; ++k) acc = fma(QUX[k * LDN + col], KB[k * LDN + n], acc);
            pq[col] = acc;
        }
        if (lane < m) {
            double acc = 0.0;
#pragma unroll
            for (int l = 0; l < m; ++l) acc = fma(QUU[lane * LDQ + l], KB[l * LDN + n], acc);
            zv[lane] = acc + QUX[lane * LDN + n];
        }
#pragma unroll
        for (int k = 0; k < MMAX; ++k) {
            y[k] = 0.0;
            if (k < m) {
                double acc = 2.0 * QUX[k * LDN + col];
#pragma unroll
                for (int l = 0; l < MMAX; ++l)
                    if (l < m) acc = fma(QUU[k * LDQ + l], x[l], acc);
                y[k] = acc;
                if (act) Ysm[k * LDN + col] = acc;
            }
        }
        __syncwarp();
        // ================= phase F: T += 1/2 (K^T Y + Y^T K); P <- (T + T^T)/2; p <- Q_x + K^T z + pq =================
        if (act) {
#pragma unroll 4
            for (int row = 0; row < n; ++row) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < MMAX; ++k)
                    if (k < m) acc = fma(KB[k * LDN + row], y[k], fma(Ysm[k * LDN + row], x[k], acc));
                Tm[(size_t)row * LDP + col] += 0.5 * acc;
            }
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < MMAX; ++k)
                if (k < m) acc = fma(x[k], zv[k], acc);
            const double pnew = Qx[col] + acc + pq[col];
            if (!isfinite(pnew)) st |= DPILQR_ST_NONFINITE;
            pvec[col] = pnew;
        }
        __syncwarp();
        if (act) {
#pragma unroll 4
            for (int row = 0; row < n; ++row) P[(size_t)row * LDP + col] = 0.5 * (Tm[(size_t)row * LDP + col] + Tm[(size_t)col * LDP + row]);
        }
        // (the __syncwarp at the top of the next step orders this before its readers)
    }
    wait_record();
    st = __reduce_or_sync(FULL, st);
    if (lane == 0 && st != 0 && p.status) atomicOr(p.status + b, st);
}

// instantiated team sizes per size class: what the DP-iLQR neighbourhoods and small planar teams need
template <int S, int C>
constexpr int warp_max_agents()
{
    int a = 1;
    while ((a + 1) * S <= 32 && (a + 1) * C <= 16 && a + 1 <= 4) ++a;
    return a;
}

template <int S, int C, int A>
int launch_warp_one(const BackwardParams &p, int n_blocks, cudaStream_t stream)
{
    if constexpr (A <= warp_max_agents<S, C>()) {
        const size_t smem = (size_t)warp_smem(A, S, C).total_doubles * 8 * kWarpKernelWarps;
        const int grid = (n_blocks + kWarpKernelWarps - 1) / kWarpKernelWarps;
        auto kernel = backward_warp_kernel<S, C, A>;
        if (smem > 48 * 1024) DPILQR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))  /* the ceiling, not this launch's need: concurrent callers must not lower it under each other */;
        kernel<<<grid, 32 * kWarpKernelWarps, smem, stream>>>(p);
        DPILQR_CUDA(cudaGetLastError());
        return DPILQR_OK;
    } else {
        set_error("backward warp kernel: %d agents of (%d, %d) not instantiated", A, S, C);
        return DPILQR_E_UNSUPPORTED;
    }
}

template <int S, int C>
int launch_warp_sc(const BackwardParams &p_in, int n_blocks, cudaStream_t stream)
{
    BackwardParams p = p_in;
    p.n_launch = n_blocks;
    switch (p.batch.n_agents) {
    case 1: return launch_warp_one<S, C, 1>(p, n_blocks, stream);
    case 2: return launch_warp_one<S, C, 2>(p, n_blocks, stream);
    case 3: return launch_warp_one<S, C, 3>(p, n_blocks, stream);
    case 4: return launch_warp_one<S, C, 4>(p, n_blocks, stream);
    default: break;
    }
    set_error("backward warp kernel: %d agents not instantiated", p.batch.n_agents);
    return DPILQR_E_UNSUPPORTED;
}

template <int S, int C>
bool warp_applies_sc(int a)
{
    return a <= warp_max_agents<S, C>() && (size_t)warp_smem(a, S, C).total_doubles * 8 * kWarpKernelWarps <= 200 * 1024;
}

bool backward_warp_applies(int a, int s, int c)
{
    if (s == 12 && c == 4) return warp_applies_sc<12, 4>(a);
    if (s == 6 && c == 3) return warp_applies_sc<6, 3>(a);
    if (s == 4 && c == 2) return warp_applies_sc<4, 2>(a);
    if (s == 3 && c == 2) return warp_applies_sc<3, 2>(a);
    if (s == 5 && c == 2) return warp_applies_sc<5, 2>(a);
    return false;
}

int launch_backward_warp(const BackwardParams &p, int n_blocks, cudaStream_t stream)
{
    const int s = p.batch.s, c = p.batch.c;
    if (s == 12 && c == 4) return launch_warp_sc<12, 4>(p, n_blocks, stream);
    if (s == 6 && c == 3) return launch_warp_sc<6, 3>(p, n_blocks, stream);
    if (s == 4 && c == 2) return launch_warp_sc<4, 2>(p, n_blocks, stream);
    if (s == 3 && c == 2) return launch_warp_sc<3, 2>(p, n_blocks, stream);
    if (s == 5 && c == 2) return launch_warp_sc<5, 2>(p, n_blocks, stream);
    set_error("backward kernel: unsupported per-agent dimensions (%d, %d)", s, c);
    return DPILQR_E_UNSUPPORTED;
}

}  // namespace dpilqr
