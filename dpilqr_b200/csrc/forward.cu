// forward.cu -- Kernel 1: batched rollout + line search (sm_100a).
//
// One CTA per problem evaluates EVERY alpha candidate of the line search in one launch:
//   u_t = U[t] + K[t] (x_t - X[t]) + alpha d[t],  x_{t+1} = RK4(x_t, u_t),  J += cost(x_t, u_t)
// which replaces the reference's sequential ilqrSolver._rollout / _forward_pass loop
// (reference control.py:80-114, <=10 passes per iteration) and, per step, the Python loops in
// MultiDynamicalModel.__call__ (dynamics.py:159-171) and GameCost.__call__ (cost.py:197-206,
// 79-83, 117-133).
//
// Work decomposition inside the CTA, per time step:
//   gain phase   : warp = 8 gain rows x 4 column lanes; every K[t] element is read from HBM
//                  exactly once (32-byte sectors fully used) and applied to all candidates from
//                  registers; 2-stage shuffle reduction.
//   agent phase  : one thread per (candidate, agent): reference cost, then the RK4 step with the
//                  state in registers.
//   pair phase   : one thread per (candidate, agent pair): proximity penalty.
//   sum phase    : one thread per candidate adds the step cost in the reference's summation
//                  order (agents ascending; pairs in NumPy pairwise-sum order; cost.py:206).
#include "cost.cuh"
#include "kernels.cuh"

namespace dpilqr {

// doubles in front of the (16-byte aligned) mbarrier + gain staging area of the dynamic shared memory
__host__ __device__ inline size_t forward_prefix_doubles(int a, int s, int c, int NA)
{
    const int n = a * s, m = a * c, pairs = a * (a - 1) / 2;
    const size_t doubles = (size_t)3 * NA * n + (size_t)NA * m + n + 2 * m + (size_t)NA * a + (size_t)NA * (pairs > 0 ? pairs : 1) + NA;
    return (doubles + 1) & ~(size_t)1;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 128 ? 3 : 1) forward_kernel(const ForwardParams p)
{
    extern __shared__ double smem[];
    const Batch &bt = p.batch;
    if (p.n_active != nullptr && (int)blockIdx.x >= *p.n_active) return;
    const int b = p.active ? p.active[blockIdx.x] : blockIdx.x;
    const int a = bt.n_agents, s = bt.s, c = bt.c, T = bt.horizon;
    const int n = a * s, m = a * c, pairs = a * (a - 1) / 2;
    const int NA = p.n_alpha;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;

    // shared-memory carve-up
    double *xbuf0 = smem;                 // [NA][n]
    double *xbuf1 = xbuf0 + NA * n;       // [NA][n]
    double *dx = xbuf1 + NA * n;          // [NA][n]
    double *ucur = dx + NA * n;           // [NA][m]
    double *xref = ucur + NA * m;         // [n]
    double *uref = xref + n;              // [m]
    double *dref = uref + m;              // [m]
    double *refc = dref + m;              // [NA][a]
    double *proxc = refc + NA * a;        // [NA][max(pairs,1)]
    double *Jacc = proxc + NA * (pairs > 0 ? pairs : 1);  // [NA]
    double *mbar_slot = smem + forward_prefix_doubles(a, s, c, NA);  // [2] mbarrier of the K[t] bulk copies
    double *Ks = mbar_slot + 2;                                       // [m][n] gains of the current step (only with K)

    const int slot = p.slot ? p.slot[b] : 0;
    const double *Xb = p.X + (int64_t)b * p.x_stride + (int64_t)slot * p.x_slot_stride;
    const double *Ub = p.U + (int64_t)b * p.u_stride + (int64_t)slot * p.u_slot_stride;
    const double *Kb = p.K ? p.K + (int64_t)b * T * m * n : nullptr;
    const double *db = p.d ? p.d + (int64_t)b * T * m : nullptr;
    double *Xcb = p.Xc + (int64_t)b * p.xc_stride;
    double *Ucb = p.Uc + (int64_t)b * p.uc_stride;

    const int32_t *model_b = bt.model + (int64_t)b * a;
    const int32_t *ndims_b = bt.n_dims + (int64_t)b * a;
    const int32_t *cidx_b = bt.cost_idx + (int64_t)b * a;
    const double *xf_b = bt.xf + (int64_t)b * n;
    const bool has_prox = (a > 1) && (bt.has_prox == nullptr || bt.has_prox[b] != 0);
    const double radius = has_prox ? bt.radius[b] : 0.0;
    const double w_ref = bt.weights ? bt.weights[2 * b] : 1.0;
    const double w_prox = bt.weights ? bt.weights[2 * b + 1] : 200.0;
    // ProximityCost.__call__ uses the planar distance whenever all n_dims agree (cost.py:122-123)
    bool uniform_dims = true;
    for (int i = 1; i < a; ++i) uniform_dims = uniform_dims && (ndims_b[i] == ndims_b[0]);

    // optional per-phase cycle counters of CTA 0 / thread 0 (debug aid, see dpilqr_debug_backward_timing)
    long long tacc[6] = {0, 0, 0, 0, 0, 0};
    long long tmark = 0;
    const bool timing = (p.timing != nullptr) && (blockIdx.x == 0) && (tid == 0);
    auto tick = [&](int slot) {
        if (timing) {
            const long long now = clock64();
            tacc[slot] += now - tmark;
            tmark = now;
        }
    };
    for (int k = tid; k < NA * n; k += nthr) xbuf0[k] = Xb[k % n];  // X_next[0] = X[0]
    if (tid < NA) Jacc[tid] = 0.0;
    __syncthreads();

    // K[t] (m*n contiguous doubles) is staged into shared memory by one TMA bulk copy per step, issued a step
    // ahead right after the gain phase has consumed the previous one, so its HBM latency hides behind the RK4.
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(mbar_slot);
    const unsigned k_bytes = (unsigned)(m * n * sizeof(double));
    auto fetch_gains = [&](int t) {  // thread 0 only, after a __syncthreads()
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(k_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((unsigned)__cvta_generic_to_shared(Ks)), "l"(Kb + (int64_t)t * m * n), "r"(k_bytes), "r"(mbar) : "memory");
    };
    if (Kb) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) fetch_gains(0);
    }
    if (timing) tmark = clock64();
    for (int t = 0; t <= T; ++t) {
        double *xcur = (t & 1) ? xbuf1 : xbuf0;
        double *xnxt = (t & 1) ? xbuf0 : xbuf1;
        const bool terminal = (t == T);

        // ---- load reference step, form dx, stream the candidate states out
        if (!terminal) {
            if (Kb) {
                for (int k = tid; k < n; k += nthr) xref[k] = Xb[(int64_t)t * n + k];
                for (int k = tid; k < m; k += nthr) dref[k] = db[(int64_t)t * m + k];
            }
            for (int k = tid; k < m; k += nthr) uref[k] = Ub[(int64_t)t * m + k];
        }
        for (int k = tid; k < NA * n; k += nthr) {
            const int al = k / n, j = k - al * n;
            Xcb[((int64_t)al * (T + 1) + t) * n + j] = xcur[k];
        }
        __syncthreads();
        tick(0);
        if (!terminal) {
            if (Kb) {
                for (int k = tid; k < NA * n; k += nthr) dx[k] = xcur[k] - xref[k % n];
                __syncthreads();
                // ---- gain phase: u = U[t] + (K[t] dx + alpha d[t])
                asm volatile(
                    "{\n"
                    ".reg .pred p;\n"
                    "WAIT_GAINS:\n"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                    "@p bra DONE_GAINS;\n"
                    "bra WAIT_GAINS;\n"
                    "DONE_GAINS:\n"
                    "}\n" ::"r"(mbar), "r"(t & 1) : "memory");
                const double *Kt = Ks;
                if ((m & 7) == 0 && (n & 3) == 0) {
                    // FP64 tensor path: dU (m x NA) = K[t] (m x n) * dx^T (n x NA) in 8x8 tiles, k-steps of 4
                    // (candidates beyond NA are zero columns).  Lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4)+{0,1}].
                    const int fr = lane >> 2, fc = lane & 3;
                    const int n_tiles = (m >> 3) * ((NA + 7) >> 3);
                    for (int tile = warp; tile < n_tiles; tile += nwarp) {
                        const int mt = tile % (m >> 3), nt = tile / (m >> 3);
                        const double *ap = Kt + (size_t)(8 * mt + fr) * n + fc;
                        const int al_b = 8 * nt + fr;                 // candidate of this lane's B fragment
                        const double *bp = dx + (size_t)(al_b < NA ? al_b : 0) * n + fc;
                        const bool bvalid = al_b < NA;
                        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;  // two accumulator pairs: independent chains
                        for (int ks = 0; ks < (n >> 2); ks += 2) {
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                         : "+d"(c0), "+d"(c1) : "d"(ap[4 * ks]), "d"(bvalid ? bp[4 * ks] : 0.0));
                            if (ks + 1 < (n >> 2))
                                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                             : "+d"(e0), "+d"(e1) : "d"(ap[4 * ks + 4]), "d"(bvalid ? bp[4 * ks + 4] : 0.0));
                        }
                        const int r = 8 * mt + fr;
                        const int al0 = 8 * nt + 2 * fc;
                        if (al0 < NA) ucur[al0 * m + r] = uref[r] + ((c0 + e0) + p.alpha[al0] * dref[r]);
                        if (al0 + 1 < NA) ucur[(al0 + 1) * m + r] = uref[r] + ((c1 + e1) + p.alpha[al0 + 1] * dref[r]);
                    }
                } else {
                const int q = lane & 3, rr = lane >> 2;
                for (int r0 = warp * 8; r0 < m; r0 += nwarp * 8) {
                    const int r = r0 + rr;
                    double acc[kMaxAlpha];
#pragma unroll
                    for (int al = 0; al < kMaxAlpha; ++al) acc[al] = 0.0;
                    if (r < m) {
                        const double *Krow = Kt + (int64_t)r * n;
                        for (int j = q; j < n; j += 4) {
                            const double kv = Krow[j];
#pragma unroll
                            for (int al = 0; al < kMaxAlpha; ++al)
                                if (al < NA) acc[al] = fma(kv, dx[al * n + j], acc[al]);
                        }
                    }
#pragma unroll
                    for (int al = 0; al < kMaxAlpha; ++al) {
                        if (al < NA) {
                            double v = acc[al];
                            v += __shfl_xor_sync(0xffffffffu, v, 1);
                            v += __shfl_xor_sync(0xffffffffu, v, 2);
                            if (q == 0 && r < m) ucur[al * m + r] = uref[r] + (v + p.alpha[al] * dref[r]);
                        }
                    }
                }
                }
            } else {
                for (int k = tid; k < NA * m; k += nthr) ucur[k] = uref[k % m];
            }
            __syncthreads();
            if (Kb && tid == 0 && t + 1 < T) fetch_gains(t + 1);
            tick(1);
            for (int k = tid; k < NA * m; k += nthr) {
                const int al = k / m, r = k - al * m;
                Ucb[((int64_t)al * T + t) * m + r] = ucur[k];
            }
            tick(2);
        }

        // ---- agent phase: reference cost at (x_t, u_t), then x_{t+1} = RK4(x_t, u_t)
        for (int item = tid; item < NA * a; item += nthr) {
            const int al = item / a, i = item - al * a;
            const int model = model_b[i];
            const int ci = cidx_b[i];
            const double *xi = xcur + al * n + i * s;
            const double *ui = ucur + al * m + i * c;
            double *xo = xnxt + al * n + i * s;
            dispatch_model(model, [&]<int M>() {
                constexpr int NX = model_nx(M), NU = model_nu(M);
                double x[NX], u[NU];
#pragma unroll
                for (int k = 0; k < NX; ++k) x[k] = xi[k];
#pragma unroll
                for (int k = 0; k < NU; ++k) u[k] = terminal ? 0.0 : ui[k];
                const double *Qm = (terminal ? bt.Qf : bt.Q) + (int64_t)ci * NX * NX;
                const double *Rm = bt.R + (int64_t)ci * NU * NU;
                refc[al * a + i] = reference_cost<M>(x, u, xf_b + i * s, Qm, Rm, terminal);
                if (!terminal) {
                    model_step<M>(bt.dt, x, u);
#pragma unroll
                    for (int k = 0; k < NX; ++k) xo[k] = x[k];
                }
            });
        }
        tick(3);
        // ---- pair phase: fmin(0, dist - radius)^2  (reference cost.py:117-133, util.py:48-87)
        if (has_prox) {
            for (int item = tid; item < NA * pairs; item += nthr) {
                const int al = item / pairs, pr = item - al * pairs;
                // decode pair index -> (i, j), itertools.combinations order
                int i = 0, rem = pr;
                while (rem >= a - 1 - i) { rem -= a - 1 - i; ++i; }
                const int j = i + 1 + rem;
                const int nd = uniform_dims ? 2 : min(ndims_b[i], ndims_b[j]);
                proxc[al * pairs + pr] = pair_penalty(xcur + al * n + i * s, xcur + al * n + j * s, nd, radius);
            }
        }
        __syncthreads();
        tick(4);
        // ---- sum phase, reference order: PROX_WEIGHT * prox + REF_WEIGHT * ref_total (cost.py:206)
        if (tid < NA) {
            double ref_total = 0.0;
            for (int i = 0; i < a; ++i) ref_total += refc[tid * a + i];
            const double prox = has_prox ? numpy_pairwise_sum(proxc + tid * pairs, pairs) : 0.0;
            Jacc[tid] += w_prox * prox + w_ref * ref_total;
        }
        tick(5);
        // next iteration's first __syncthreads orders these reads before refc/proxc are rewritten
    }
    if (timing) {
        for (int k = 0; k < 6; ++k) p.timing[24 + k] = tacc[k];
    }
    __syncthreads();
    if (tid < NA) p.Jc[(int64_t)b * p.jc_stride + tid] = Jacc[tid];
}

static size_t forward_smem_bytes(int a, int s, int c, int NA, bool with_gains)
{
    const size_t doubles = forward_prefix_doubles(a, s, c, NA) + 2 + (with_gains ? (size_t)(a * c) * (a * s) : 0);
    return doubles * sizeof(double);
}

int launch_forward(const ForwardParams &p_in, int n_blocks, cudaStream_t stream)
{
    ForwardParams p = p_in;
    p.timing = g_backward_timing;
    const Batch &bt = p.batch;
    if (p.n_alpha < 1 || p.n_alpha > kMaxAlpha) {
        set_error("n_alpha must be in 1..%d (got %d)", kMaxAlpha, p.n_alpha);
        return DPILQR_E_INVALID;
    }
    if (n_blocks <= 0) return DPILQR_OK;
    const size_t smem = forward_smem_bytes(bt.n_agents, bt.s, bt.c, p.n_alpha, p.K != nullptr);
    if (smem > 227 * 1024) {
        set_error("forward kernel: problem too large for shared memory (%zu bytes)", smem);
        return DPILQR_E_UNSUPPORTED;
    }
    static bool attr_set = false;
    if (!attr_set) {
        DPILQR_CUDA(cudaFuncSetAttribute(forward_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DPILQR_CUDA(cudaFuncSetAttribute(forward_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int items = p.n_alpha * bt.n_agents;
    if (items <= 128) forward_kernel<128><<<n_blocks, 128, smem, stream>>>(p);
    else forward_kernel<256><<<n_blocks, 256, smem, stream>>>(p);
    DPILQR_CUDA(cudaGetLastError());
    return DPILQR_OK;
}

}  // namespace dpilqr
