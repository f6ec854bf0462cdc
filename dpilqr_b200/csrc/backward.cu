// backward.cu -- Kernel 3: backward Riccati recursion, one CTA per problem (sm_100a, FP64).
//
// Replaces ilqrSolver._backward_pass (reference control.py:116-148).  Per time step, with the
// block structure the reference throws away (A, B block-diagonal per agent; P symmetric):
//
//   phase A   S = B^T (P + mu I)  ->  Q_ux = S A,  Q_uu = L_uu + S B      (S only lives in registers)
//             Q_u = L_u + B^T p,  Q_x = L_x + A^T p
//   phase B   Q_xx = L_xx + A^T P A, in place on the upper-triangular 12x12 blocks of P
//   phase C   partial-pivot LU of Q_uu (same pivoting rule as LAPACK dgetf2; the reference calls
//             np.linalg.solve = dgesv, control.py:141-142; no definiteness is assumed)
//   phase D   K = -Q_uu^{-1} Q_ux, d = -Q_uu^{-1} Q_u: one thread per right-hand-side column
//   phase E   Y = Q_uu K + 2 Q_ux (in place over Q_ux), z = Q_uu d + Q_u, Q_ux^T d
//   phase F   P <- Q_xx + 1/2 (K^T Y + Y^T K)   (== the reference's symmetrised
//             Q_xx + K^T Q_uu K + K^T Q_ux + Q_ux^T K), upper blocks only, register tiles
//             p <- Q_x + K^T z + Q_ux^T d
//
// P, Q_ux/Y, K, Q_uu and its LU factors stay in shared memory for the whole recursion (about
// 210 kB for 10 Quadcopter12D agents -> one CTA per SM); only the stage records stream in and
// K, d stream out.  Problems too large for shared memory keep Q_ux/Y and K in an L2-resident
// global scratch instead (same code, generic pointers).
#include "kernels.cuh"

namespace dpilqr {

template <int S>
struct TileSize {
    static constexpr int value = (S % 4 == 0) ? 4 : (S % 3 == 0) ? 3 : S;
};

__host__ __device__ inline int pblock_stride(int s) { return s * s + 2; }  // +2 doubles: de-alias banks across blocks

template <int S, int C, int AT>
__global__ void __launch_bounds__(512, 1) backward_kernel(const BackwardParams p)
{
    extern __shared__ double smem[];
    const Batch &bt = p.batch;
    if (p.n_active != nullptr && (int)blockIdx.x >= *p.n_active) return;
    const int b = p.active ? p.active[blockIdx.x] : blockIdx.x;
    const int a = AT > 0 ? AT : bt.n_agents;
    const int T = bt.horizon;
    const int n = a * S, m = a * C;
    const int nblk = a * (a + 1) / 2;
    constexpr int PBS = S * S + 2;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    const StageLayout L = stage_layout(a, S, C);

    // ---- shared memory carve-up (doubles)
    double *Pb = smem;                           // [nblk][PBS]  upper-triangular blocks of P
    double *QUU = Pb + (size_t)nblk * PBS;       // [m][m]
    double *LU = QUU + m * m;                    // [m][m]
    double *rec = LU + m * m;                    // [L.stride]   current stage record
    double *pvec = rec + L.stride;               // [n]
    double *Qx = pvec + n;                       // [n]
    double *pq = Qx + n;                         // [n]   Q_ux^T d
    double *Qu = pq + n;                         // [m]
    double *dv = Qu + m;                         // [m]
    double *zv = dv + m;                         // [m]
    int *perm = reinterpret_cast<int *>(zv + m); // [m]
    int *pivrow = perm + m;                      // [1] (+pad)
    double *after = reinterpret_cast<double *>(pivrow + 2 + (m & 1));
    double *QUX, *KB;                            // [m][n] each
    if (p.use_global_scratch) {
        QUX = p.scratch + (size_t)blockIdx.x * 2 * m * n;
        KB = QUX + (size_t)m * n;
    } else {
        QUX = after;
        KB = QUX + (size_t)m * n;
    }

    const int32_t *cidx_b = bt.cost_idx + (int64_t)b * a;
    const double w_ref = bt.weights ? bt.weights[2 * b] : 1.0;
    const double mu = p.mu[b];
    const double *stage_b = p.stage + (int64_t)b * (T + 1) * L.stride;
    double *Kb = p.K + (int64_t)b * T * m * n;
    double *db = p.d + (int64_t)b * T * m;
    int st = 0;

    auto blk_index = [a](int i, int j) { return i * a - (i * (i - 1)) / 2 + (j - i); };  // i <= j
    auto Pij = [&](int i, int j, int r, int cc) -> double {
        return (i <= j) ? Pb[(size_t)blk_index(i, j) * PBS + r * S + cc] : Pb[(size_t)blk_index(j, i) * PBS + cc * S + r];
    };

    // ---- terminal condition: p = L_x, P = L_xx at (X[T], u = 0)  (control.py:125-129)
    for (int k = tid; k < L.stride; k += nthr) rec[k] = stage_b[(int64_t)T * L.stride + k];
    __syncthreads();
    for (int k = tid; k < nblk * S * S; k += nthr) {
        const int blk = k / (S * S), e = k - blk * (S * S);
        const int r = e / S, cc = e - r * S;
        // decode blk -> (i, j)
        int i = 0, rem = blk;
        while (rem >= a - i) { rem -= a - i; ++i; }
        const int j = i + rem;
        double v = 0.0;
        if (i == j) {
            const double *Qf = bt.Qf + (int64_t)cidx_b[i] * S * S;
            v = w_ref * (Qf[r * S + cc] + Qf[cc * S + r]);
            if (r < 3 && cc < 3) v += rec[L.offHd + 9 * i + r * 3 + cc];
        } else if (r < 3 && cc < 3) {
            v = rec[L.offHo + 9 * pair_index(i, j, a) + r * 3 + cc];
        }
        Pb[(size_t)blk * PBS + e] = v;
    }
    for (int k = tid; k < n; k += nthr) pvec[k] = rec[L.offLx + k];
    __syncthreads();

    for (int t = T - 1; t >= 0; --t) {
        // ---- stream in the stage record of step t
        for (int k = tid; k < L.stride; k += nthr) rec[k] = stage_b[(int64_t)t * L.stride + k];
        __syncthreads();
        const double *sA = rec + L.offA, *sB = rec + L.offB;

        // ---- phase A: Q_ux, Q_uu (S = B^T (P + mu I) in registers), Q_u, Q_x
        for (int it = tid; it < a * a * C; it += nthr) {
            const int g = it % C;
            const int j = (it / C) % a;
            const int i = it / (C * a);
            const double *Bi = sB + i * S * C;
            double Srow[S];
#pragma unroll
            for (int sg = 0; sg < S; ++sg) Srow[sg] = 0.0;
#pragma unroll
            for (int r = 0; r < S; ++r) {
                const double bv = Bi[r * C + g];
#pragma unroll
                for (int sg = 0; sg < S; ++sg) {
                    double pv = Pij(i, j, r, sg);
                    if (i == j && r == sg) pv += mu;
                    Srow[sg] = fma(bv, pv, Srow[sg]);
                }
            }
            const double *Aj = sA + j * S * S;
            const double *Bj = sB + j * S * C;
            const int row = i * C + g;
#pragma unroll
            for (int sg2 = 0; sg2 < S; ++sg2) {
                double acc = 0.0;
#pragma unroll
                for (int sg = 0; sg < S; ++sg) acc = fma(Srow[sg], Aj[sg * S + sg2], acc);
                QUX[(size_t)row * n + j * S + sg2] = acc;  // L_ux == 0 (cost.py:91)
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
                QUU[row * m + j * C + g2] = acc;
            }
            if (j == 0) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < S; ++r) acc = fma(Bi[r * C + g], pvec[i * S + r], acc);
                Qu[row] = rec[L.offLu + row] + acc;
            }
        }
        for (int col = tid; col < n; col += nthr) {
            const int j = col / S, sg = col - j * S;
            const double *Aj = sA + j * S * S;
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < S; ++r) acc = fma(Aj[r * S + sg], pvec[j * S + r], acc);
            Qx[col] = rec[L.offLx + col] + acc;
        }
        __syncthreads();

        // ---- phase B: Q_xx = L_xx + A^T P A in place (upper blocks); item = (block, column)
        {
            const int blocks_per_round = nthr / S;
            for (int blk0 = 0; blk0 < nblk; blk0 += blocks_per_round) {
                const int blk = blk0 + tid / S;
                const int sg = tid % S;
                const bool live = (tid < blocks_per_round * S) && (blk < nblk);
                double out[S];
                if (live) {
                    int i = 0, rem = blk;
                    while (rem >= a - i) { rem -= a - i; ++i; }
                    const int j = i + rem;
                    const double *Pblk = Pb + (size_t)blk * PBS;
                    const double *Ai = sA + i * S * S, *Aj = sA + j * S * S;
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
                            if (r < 3 && sg < 3) lxx += rec[L.offHd + 9 * i + r * 3 + sg];
                        } else if (r < 3 && sg < 3) {
                            lxx = rec[L.offHo + 9 * pair_index(i, j, a) + r * 3 + sg];
                        }
                        out[r] = lxx + acc;
                    }
                }
                __syncthreads();
                if (live) {
                    double *Pblk = Pb + (size_t)blk * PBS;
#pragma unroll
                    for (int r = 0; r < S; ++r) Pblk[r * S + sg] = out[r];
                }
            }
        }

        // ---- phase C: LU = P_perm * Q_uu with partial pivoting
        for (int k = tid; k < m * m; k += nthr) LU[k] = QUU[k];
        for (int k = tid; k < m; k += nthr) perm[k] = k;
        __syncthreads();
        for (int k = 0; k < m; ++k) {
            if (warp == 0) {
                double best = -1.0;
                int brow = k;
                for (int r = k + lane; r < m; r += 32) {
                    const double v = fabs(LU[r * m + k]);
                    if (v > best || !(v == v)) { best = (v == v) ? v : INFINITY; brow = r; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, off);
                    const int orow = __shfl_xor_sync(0xffffffffu, brow, off);
                    if (ov > best || (ov == best && orow < brow)) { best = ov; brow = orow; }
                }
                if (lane == 0) pivrow[0] = brow;
            }
            __syncthreads();
            const int pr = pivrow[0];
            if (pr != k) {
                for (int col = tid; col < m; col += nthr) {
                    const double tmp = LU[k * m + col];
                    LU[k * m + col] = LU[pr * m + col];
                    LU[pr * m + col] = tmp;
                }
                if (tid == 0) { const int tmp = perm[k]; perm[k] = perm[pr]; perm[pr] = tmp; }
            }
            __syncthreads();
            const double pivot = LU[k * m + k];
            if (pivot == 0.0) st |= DPILQR_ST_SINGULAR;
            const int rem = m - k - 1;
            for (int r = k + 1 + tid; r < m; r += nthr) LU[r * m + k] /= pivot;
            __syncthreads();
            for (int idx = tid; idx < rem * rem; idx += nthr) {
                const int r = k + 1 + idx / rem, cc = k + 1 + idx % rem;
                LU[r * m + cc] = fma(-LU[r * m + k], LU[k * m + cc], LU[r * m + cc]);
            }
            __syncthreads();
        }

        // ---- phase D: solve, one thread per right-hand side (columns of Q_ux, then Q_u)
        double *Kt = Kb + (int64_t)t * m * n;
        if constexpr (AT > 0) {
            constexpr int M = AT * C;
            for (int col = tid; col <= n; col += nthr) {
                double x[M];
#pragma unroll
                for (int k = 0; k < M; ++k) x[k] = (col < n) ? QUX[(size_t)perm[k] * n + col] : Qu[perm[k]];
#pragma unroll
                for (int k = 1; k < M; ++k) {
                    double acc = x[k];
#pragma unroll
                    for (int l = 0; l < k; ++l) acc = fma(-LU[k * M + l], x[l], acc);
                    x[k] = acc;
                }
#pragma unroll
                for (int k = M - 1; k >= 0; --k) {
                    double acc = x[k];
#pragma unroll
                    for (int l = k + 1; l < M; ++l) acc = fma(-LU[k * M + l], x[l], acc);
                    x[k] = acc / LU[k * M + k];
                }
                if (col < n) {
#pragma unroll
                    for (int k = 0; k < M; ++k) {
                        const double kv = -x[k];
                        if (!isfinite(kv)) st |= DPILQR_ST_NONFINITE;
                        KB[(size_t)k * n + col] = kv;
                        Kt[(size_t)k * n + col] = kv;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < M; ++k) {
                        dv[k] = -x[k];
                        db[(int64_t)t * m + k] = -x[k];
                    }
                }
            }
        } else {
            for (int col = tid; col <= n; col += nthr) {
                // the Q_u column uses zv as its work vector
                auto X = [&](int k) -> double & { return col < n ? KB[(size_t)k * n + col] : zv[k]; };
                for (int k = 0; k < m; ++k) {
                    double acc = (col < n) ? QUX[(size_t)perm[k] * n + col] : Qu[perm[k]];
                    for (int l = 0; l < k; ++l) acc = fma(-LU[k * m + l], X(l), acc);
                    X(k) = acc;
                }
                for (int k = m - 1; k >= 0; --k) {
                    double acc = X(k);
                    for (int l = k + 1; l < m; ++l) acc = fma(-LU[k * m + l], X(l), acc);
                    X(k) = acc / LU[k * m + k];
                }
                if (col < n) {
                    for (int k = 0; k < m; ++k) {
                        const double kv = -X(k);
                        if (!isfinite(kv)) st |= DPILQR_ST_NONFINITE;
                        KB[(size_t)k * n + col] = kv;
                        Kt[(size_t)k * n + col] = kv;
                    }
                } else {
                    for (int k = 0; k < m; ++k) {
                        const double dk = -zv[k];
                        dv[k] = dk;
                        db[(int64_t)t * m + k] = dk;
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase E: Y = Q_uu K + 2 Q_ux (in place), pq = Q_ux^T d, z = Q_uu d + Q_u
        for (int col = tid; col <= n; col += nthr) {
            if (col < n) {
                double qd = 0.0;
                for (int k = 0; k < m; ++k) qd = fma(QUX[(size_t)k * n + col], dv[k], qd);
                pq[col] = qd;
                for (int k = 0; k < m; ++k) {
                    double acc = 0.0;
                    for (int l = 0; l < m; ++l) acc = fma(QUU[k * m + l], KB[(size_t)l * n + col], acc);
                    QUX[(size_t)k * n + col] = acc + 2.0 * QUX[(size_t)k * n + col];
                }
            } else {
                for (int k = 0; k < m; ++k) {
                    double acc = 0.0;
                    for (int l = 0; l < m; ++l) acc = fma(QUU[k * m + l], dv[l], acc);
                    zv[k] = acc + Qu[k];
                }
            }
        }
        __syncthreads();

        // ---- phase F: P <- Q_xx + 1/2 (K^T Y + Y^T K) on upper blocks; p <- Q_x + K^T z + Q_ux^T d
        {
            constexpr int TS = TileSize<S>::value;
            constexpr int TPB = (S / TS) * (S / TS);  // tiles per block
            const double *Y = QUX;
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
                        ki[r] = KB[(size_t)k * n + r0 + r];
                        yi[r] = Y[(size_t)k * n + r0 + r];
                        kj[r] = KB[(size_t)k * n + c0 + r];
                        yj[r] = Y[(size_t)k * n + c0 + r];
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
            for (int col = tid; col < n; col += nthr) {
                double acc = 0.0;
                for (int k = 0; k < m; ++k) acc = fma(KB[(size_t)k * n + col], zv[k], acc);
                pvec[col] = Qx[col] + acc + pq[col];
            }
        }
        __syncthreads();
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
    const StageLayout L = stage_layout(a, s, c);
    const int n = a * s, m = a * c;
    const size_t nblk = (size_t)a * (a + 1) / 2;
    size_t base = nblk * pblock_stride(s) + 2 * (size_t)m * m + L.stride + 3 * (size_t)n + 3 * (size_t)m;
    base = base * 8 + ((size_t)m + 2 + (m & 1)) * 4;
    const size_t mats = 2 * (size_t)m * n * 8;
    BackwardPlan plan;
    plan.use_global_scratch = (base + mats > 227 * 1024) ? 1 : 0;
    plan.smem_bytes = plan.use_global_scratch ? base : base + mats;
    plan.threads = n <= 24 ? 128 : (n <= 60 ? 256 : 512);
    return plan;
}

int64_t backward_scratch_doubles(int n_problems, int a, int s, int c)
{
    const BackwardPlan plan = plan_backward(a, s, c);
    return plan.use_global_scratch ? (int64_t)n_problems * 2 * (a * c) * (a * s) : 0;
}

template <int S, int C, int AT>
static int launch_typed(const BackwardParams &p, int n_blocks, const BackwardPlan &plan, cudaStream_t stream)
{
    auto kernel = backward_kernel<S, C, AT>;
    static bool attr_set = false;
    if (!attr_set) {
        DPILQR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    kernel<<<n_blocks, plan.threads, plan.smem_bytes, stream>>>(p);
    DPILQR_CUDA(cudaGetLastError());
    return DPILQR_OK;
}

int launch_backward(const BackwardParams &p_in, int n_blocks, cudaStream_t stream)
{
    if (n_blocks <= 0) return DPILQR_OK;
    BackwardParams p = p_in;
    const Batch &bt = p.batch;
    const int a = bt.n_agents, s = bt.s, c = bt.c;
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
        if (a == 10) return launch_typed<12, 4, 10>(p, n_blocks, plan, stream);
        return launch_typed<12, 4, 0>(p, n_blocks, plan, stream);
    }
    if (s == 6 && c == 3) return launch_typed<6, 3, 0>(p, n_blocks, plan, stream);
    if (s == 4 && c == 2) return launch_typed<4, 2, 0>(p, n_blocks, plan, stream);
    if (s == 3 && c == 2) return launch_typed<3, 2, 0>(p, n_blocks, plan, stream);
    if (s == 5 && c == 2) return launch_typed<5, 2, 0>(p, n_blocks, plan, stream);
    set_error("backward kernel: unsupported per-agent dimensions (%d, %d)", s, c);
    return DPILQR_E_UNSUPPORTED;
}

}  // namespace dpilqr
