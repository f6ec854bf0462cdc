// rollout.cuh -- Kernel 1: batched rollout + line search (sm_100a), one instantiation per model / size class.
//
//   u_t = U[t] + K[t] (x_t - X[t]) + alpha d[t],   x_{t+1} = RK4(x_t, u_t),   J += cost(x_t, u_t)
//
// replaces the reference's sequential ilqrSolver._rollout / _forward_pass (reference control.py:80-114) and, per
// step, the Python loops in MultiDynamicalModel.__call__ (dynamics.py:159-171) and GameCost.__call__
// (cost.py:197-206, 79-83, 117-133).
//
// Work decomposition.  A *group* is one problem with the NA line-search candidates of this launch (NA = 1, 2, 7 in
// the solver's staged search, 1..10 through the C ABI); a *rollout* is one (problem, candidate).  A CTA carries G
// groups (G is chosen by the launcher so that the grid fills the machine: many groups per CTA for thousands of
// problems, one group per CTA -- minimum latency -- for the stragglers of a solve) and walks them through the
// horizon in lock step.  Per time step:
//   P1  dx = x_t - X[t] per rollout; proximity penalties, one thread per (rollout, agent pair); the cost sum of the
//       previous step in the reference's summation order (agents ascending, pairs in NumPy pairwise order)
//   P2  gain phase: the rows of all the CTA's K[t] matrices are dealt to the warps eight at a time; a quad of lanes
//       walks one row with 16-byte loads straight from global memory (every K element is read once per launch and
//       feeds NA accumulators) and reduces with two shuffles
//   P3  agent phase: one thread per (rollout, agent) -- reference cost, then the 5-sub-step RK4 with the state in
//       registers; the candidate trajectories stream out from registers
// with three block barriers per step.  X[t], U[t], d[t] of the next step arrive by cp.async and the next step's
// gains are pulled into L2 by one bulk prefetch per group while the agents integrate.
#pragma once
#include "cost.cuh"
#include "kernels.cuh"

namespace dpilqr {

constexpr int kRolloutMaxThreads = 256;

struct RolloutSmem {
    size_t xcur, dx, ucur, refc, proxc, Jacc, xref, uref, dref, radius, wref, wprox, ints, total_doubles;
};

// Shared-memory carve-up in doubles for G groups of NA candidates
__host__ __device__ inline RolloutSmem rollout_smem(int a, int s, int c, int G, int NA)
{
    const size_t n = (size_t)a * s, m = (size_t)a * c, P = a > 1 ? (size_t)a * (a - 1) / 2 : 1, R = (size_t)G * NA;
    auto even = [](size_t v) { return (v + 1) & ~(size_t)1; };
    RolloutSmem L{};
    size_t off = 0;
    L.xcur = off;   off += even(R * n);
    L.dx = off;     off += even(R * n);
    L.ucur = off;   off += even(R * m);
    L.refc = off;   off += even(2 * R * a);
    L.proxc = off;  off += even(2 * R * P);
    L.Jacc = off;   off += even(R);
    L.xref = off;   off += even(G * n);
    L.uref = off;   off += even(G * m);
    L.dref = off;   off += even(G * m);
    L.radius = off; off += even(G);
    L.wref = off;   off += even(G);
    L.wprox = off;  off += even(G);
    // ints: problem index, slot, flags per group; n_dims per (group, agent); pair table (i, j) as bytes
    L.ints = off;   off += even(((size_t)3 * G + (size_t)G * a + (P + 1) / 2 + 2) / 2 + 1);
    L.total_doubles = off;
    return L;
}

template <int MC, int NAMAX, bool GAINS>
__global__ void __launch_bounds__(kRolloutMaxThreads, 2) rollout_kernel(const ForwardParams p)
{
    extern __shared__ __align__(16) double smem[];
    const Batch &bt = p.batch;
    constexpr int S = class_nx(MC), C = class_nu(MC);
    const int a = bt.n_agents, T = bt.horizon;
    const int n = a * S, m = a * C, pairs = a * (a - 1) / 2, P = pairs > 0 ? pairs : 1;
    const int NA = p.n_alpha, G = p.groups_per_cta;
    const int count = p.n_active ? min(*p.n_active, p.n_list) : p.n_list;
    const int g0 = blockIdx.x * G;
    if (g0 >= count) return;
    const int Ge = min(G, count - g0);  // groups of this CTA
    const int R = Ge * NA;              // rollouts of this CTA
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;

    const RolloutSmem SM = rollout_smem(a, S, C, G, NA);
    double *xcur = smem + SM.xcur;    // [R][n]
    double *dx = smem + SM.dx;        // [R][n]
    double *ucur = smem + SM.ucur;    // [R][m]
    double *refc = smem + SM.refc;    // [2][R][a]   by step parity
    double *proxc = smem + SM.proxc;  // [2][R][P]
    double *Jacc = smem + SM.Jacc;    // [R]
    double *xref = smem + SM.xref;    // [Ge][n]
    double *uref = smem + SM.uref;    // [Ge][m]
    double *dref = smem + SM.dref;    // [Ge][m]
    double *g_radius = smem + SM.radius, *g_wref = smem + SM.wref, *g_wprox = smem + SM.wprox;
    int *g_prob = reinterpret_cast<int *>(smem + SM.ints);  // [G] problem index
    int *g_slot = g_prob + G;                               // [G] trajectory slot
    int *g_flags = g_slot + G;                              // [G] bit 0: proximity term, bit 1: planar distance for all pairs
    int *g_ndims = g_flags + G;                             // [G][a]
    unsigned char *pair_ij = reinterpret_cast<unsigned char *>(g_ndims + (size_t)G * a);  // [P][2]

    // ---- per-group constants
    for (int g = tid; g < Ge; g += nthr) {
        const int b = p.active ? p.active[g0 + g] : g0 + g;
        g_prob[g] = b;
        g_slot[g] = p.slot ? p.slot[b] : 0;
        const bool has_prox = (a > 1) && (bt.has_prox == nullptr || bt.has_prox[b] != 0);
        // ProximityCost.__call__ uses the planar distance whenever all n_dims agree (cost.py:122-123)
        bool uniform_dims = true;
        const int32_t *nd = bt.n_dims + (int64_t)b * a;
        for (int i = 0; i < a; ++i) {
            g_ndims[g * a + i] = nd[i];
            uniform_dims = uniform_dims && (nd[i] == nd[0]);
        }
        g_flags[g] = (has_prox ? 1 : 0) | (uniform_dims ? 2 : 0);
        g_radius[g] = has_prox ? bt.radius[b] : 0.0;
        g_wref[g] = bt.weights ? bt.weights[2 * b] : 1.0;
        g_wprox[g] = bt.weights ? bt.weights[2 * b + 1] : 200.0;
    }
    for (int pr = tid; pr < pairs; pr += nthr) {  // itertools.combinations order (reference util.py:58)
        int i = 0, rem = pr;
        while (rem >= a - 1 - i) { rem -= a - 1 - i; ++i; }
        pair_ij[2 * pr] = (unsigned char)i;
        pair_ij[2 * pr + 1] = (unsigned char)(i + 1 + rem);
    }
    __syncthreads();

    auto x_traj = [&](int g) { return p.X + (int64_t)g_prob[g] * p.x_stride + (int64_t)g_slot[g] * p.x_slot_stride; };
    auto u_traj = [&](int g) { return p.U + (int64_t)g_prob[g] * p.u_stride + (int64_t)g_slot[g] * p.u_slot_stride; };
    auto cp_async8 = [](double *dst, const double *src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    // X[t], U[t], d[t] of every group -> shared memory (asynchronously; consumed one step later)
    auto fetch_refs = [&](int t) {
        if (t < T) {
            if constexpr (GAINS) {
                for (int k = tid; k < Ge * n; k += nthr) {
                    const int g = k / n, j = k - g * n;
                    cp_async8(xref + k, x_traj(g) + (int64_t)t * n + j);
                }
                for (int k = tid; k < Ge * m; k += nthr) {
                    const int g = k / m, j = k - g * m;
                    cp_async8(dref + k, p.d + ((int64_t)g_prob[g] * T + t) * m + j);
                }
            }
            for (int k = tid; k < Ge * m; k += nthr) {
                const int g = k / m, j = k - g * m;
                cp_async8(uref + k, u_traj(g) + (int64_t)t * m + j);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto prefetch_gains = [&](int t) {  // one bulk L2 prefetch per group
        if constexpr (GAINS) {
            if (p.prefetch && t < T && tid < Ge) {
                const double *src = p.K + ((int64_t)g_prob[tid] * T + t) * m * n;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((unsigned)(m * n * 8)) : "memory");
            }
        }
    };

    fetch_refs(0);
    prefetch_gains(0);
    for (int k = tid; k < R * n; k += nthr) {  // X_next[0] = X[0]
        const int r = k / n, j = k - r * n;
        xcur[k] = x_traj(r / NA)[j];
    }
    for (int r = tid; r < R; r += nthr) Jacc[r] = 0.0;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    // ---- the agent this thread integrates (fixed for the whole horizon)
    const bool is_agent = tid < R * a;
    const int ag_r = is_agent ? tid / a : 0;             // rollout
    const int ag_i = is_agent ? tid - ag_r * a : 0;      // agent
    const int ag_g = ag_r / NA, ag_al = ag_r - ag_g * NA;
    const int ag_b = g_prob[ag_g];
    const int ag_model = (MC >= kMixed4) ? bt.model[(int64_t)ag_b * a + ag_i] : MC;
    const int64_t ag_ci = bt.cost_idx[(int64_t)ag_b * a + ag_i];
    const double *ag_xf = bt.xf + (int64_t)ag_b * n + ag_i * S;
    double *ag_Xout = p.Xc + (int64_t)ag_b * p.xc_stride + (int64_t)(p.alpha_first + ag_al) * (T + 1) * n + ag_i * S;
    double *ag_Uout = p.Uc + (int64_t)ag_b * p.uc_stride + (int64_t)(p.alpha_first + ag_al) * T * m + ag_i * C;

    auto sum_step = [&](int par) {  // PROX_WEIGHT * prox + REF_WEIGHT * ref_total (cost.py:206), one thread per rollout
        for (int r = tid; r < R; r += nthr) {
            const int g = r / NA;
            const double *rc = refc + ((size_t)par * R + r) * a;
            double ref_total = 0.0;
            for (int i = 0; i < a; ++i) ref_total += rc[i];
            const double prox = (g_flags[g] & 1) ? numpy_pairwise_sum(proxc + ((size_t)par * R + r) * P, pairs) : 0.0;
            Jacc[r] += g_wprox[g] * prox + g_wref[g] * ref_total;
        }
    };

#pragma unroll 1
    for (int t = 0; t <= T; ++t) {
        const bool terminal = (t == T);
        const int par = t & 1;
        // ================= P1: cost sum of step t-1, dx, proximity penalties =================
        if (t > 0) sum_step(par ^ 1);
        if constexpr (GAINS) {
            if (!terminal) {
                for (int k = tid; k < R * n; k += nthr) {
                    const int r = k / n, j = k - r * n;
                    dx[k] = xcur[k] - xref[(r / NA) * n + j];
                }
            }
        }
        for (int k = tid; k < R * pairs; k += nthr) {  // fmin(0, dist - radius)^2 (cost.py:117-133, util.py:48-87)
            const int r = k / pairs, pr = k - r * pairs, g = r / NA;
            if (g_flags[g] & 1) {
                const int i = pair_ij[2 * pr], j = pair_ij[2 * pr + 1];
                const int nd = (g_flags[g] & 2) ? 2 : min(g_ndims[g * a + i], g_ndims[g * a + j]);
                proxc[((size_t)par * R + r) * P + pr] = pair_penalty(xcur + (size_t)r * n + i * S, xcur + (size_t)r * n + j * S, nd, g_radius[g]);
            }
        }
        __syncthreads();
        // ================= P2: controls of this step =================
        if (!terminal) {
            if constexpr (GAINS) {
                const int q = lane & 3;
                const int rows_total = Ge * m;
                for (int task = warp; task * 8 < rows_total; task += nwarp) {
                    const int grow = task * 8 + (lane >> 2);
                    const bool live = grow < rows_total;
                    const int g = live ? grow / m : 0, row = live ? grow - g * m : 0;
                    const double *Krow = p.K + (((int64_t)g_prob[g] * T + t) * m + row) * n;
                    const double *dxg = dx + (size_t)g * NA * n;
                    double acc[NAMAX];
#pragma unroll
                    for (int al = 0; al < NAMAX; ++al) acc[al] = 0.0;
                    if (live) {
                        if ((n & 1) == 0) {
#pragma unroll 4
                            for (int col = 2 * q; col < n; col += 8) {
                                const double2 kv = __ldg(reinterpret_cast<const double2 *>(Krow + col));
#pragma unroll
                                for (int al = 0; al < NAMAX; ++al) {
                                    if (al < NA) {
                                        const double2 dv = *reinterpret_cast<const double2 *>(dxg + (size_t)al * n + col);
                                        acc[al] = fma(kv.y, dv.y, fma(kv.x, dv.x, acc[al]));
                                    }
                                }
                            }
                        } else {
#pragma unroll 4
                            for (int col = q; col < n; col += 4) {
                                const double kv = __ldg(Krow + col);
#pragma unroll
                                for (int al = 0; al < NAMAX; ++al)
                                    if (al < NA) acc[al] = fma(kv, dxg[(size_t)al * n + col], acc[al]);
                            }
                        }
                    }
#pragma unroll
                    for (int al = 0; al < NAMAX; ++al) {
                        if (al < NA) {
                            double v = acc[al];
                            v += __shfl_xor_sync(0xffffffffu, v, 1);
                            v += __shfl_xor_sync(0xffffffffu, v, 2);
                            if (live && (al & 3) == q)
                                ucur[((size_t)g * NA + al) * m + row] = uref[g * m + row] + (v + p.alpha[al] * dref[g * m + row]);
                        }
                    }
                }
            } else {
                for (int k = tid; k < R * m; k += nthr) {
                    const int r = k / m, j = k - r * m;
                    ucur[k] = uref[(r / NA) * m + j];
                }
            }
            __syncthreads();
            // the reference rows of this step are consumed: fetch the next ones behind the integration
            fetch_refs(t + 1);
            prefetch_gains(t + 1);
        }
        // ================= P3: reference cost at (x_t, u_t), then x_{t+1} = RK4(x_t, u_t) =================
        if (is_agent) {
            dispatch_class<MC>(ag_model, [&]<int M>() {
                constexpr int NX = model_nx(M), NU = model_nu(M);
                static_assert(NX == S && NU == C, "model does not belong to this size class");
                double *xs = xcur + (size_t)ag_r * n + ag_i * S;
                const double *us = ucur + (size_t)ag_r * m + ag_i * C;
                double x[NX], u[NU];
                double *xo = ag_Xout + (int64_t)t * n;
                if constexpr (NX % 2 == 0) {
#pragma unroll
                    for (int k = 0; k < NX; k += 2) {
                        const double2 v = *reinterpret_cast<const double2 *>(xs + k);
                        x[k] = v.x; x[k + 1] = v.y;
                        *reinterpret_cast<double2 *>(xo + k) = v;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < NX; ++k) { x[k] = xs[k]; xo[k] = x[k]; }
                }
                if (!terminal) {
                    double *uo = ag_Uout + (int64_t)t * m;
#pragma unroll
                    for (int k = 0; k < NU; ++k) { u[k] = us[k]; uo[k] = u[k]; }
                } else {
#pragma unroll
                    for (int k = 0; k < NU; ++k) u[k] = 0.0;
                }
                const double *Qm = (terminal ? bt.Qf : bt.Q) + ag_ci * NX * NX;
                const double *Rm = bt.R + ag_ci * NU * NU;
                refc[((size_t)par * R + ag_r) * a + ag_i] = reference_cost<M>(x, u, ag_xf, Qm, Rm, terminal);
                if (!terminal) {
                    model_step<M>(bt.dt, x, u);
                    if constexpr (NX % 2 == 0) {
#pragma unroll
                        for (int k = 0; k < NX; k += 2) *reinterpret_cast<double2 *>(xs + k) = make_double2(x[k], x[k + 1]);
                    } else {
#pragma unroll
                        for (int k = 0; k < NX; ++k) xs[k] = x[k];
                    }
                }
            });
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
    }
    sum_step(T & 1);
    __syncthreads();
    for (int r = tid; r < R; r += nthr) {
        const int g = r / NA, al = r - g * NA;
        p.Jc[(int64_t)g_prob[g] * p.jc_stride + p.alpha_first + al] = Jacc[r];
    }
}

// ---- launch plan ---------------------------------------------------------------------------------------------
struct RolloutPlan {
    int G, threads, grid, prefetch;
    size_t smem;
};

inline RolloutPlan plan_rollout(int a, int s, int c, int T, int NA, int n_list, int expected, bool gains)
{
    RolloutPlan plan{};
    const int per_group = NA * a;                     // agent threads of one group
    const int row_tasks_per_group = (a * c + 7) / 8;  // gain-phase warp tasks of one group
    int Gmax = kRolloutMaxThreads / per_group;
    if (Gmax < 1) Gmax = 1;
    // fill the machine: about two CTAs per SM; stragglers get a CTA (and its warps' gain tasks) to themselves
    const int slots = 148 * 2;
    int G = (expected + slots - 1) / slots;
    if (G < 1) G = 1;
    if (G > Gmax) G = Gmax;
    while (G > 1 && rollout_smem(a, s, c, G, NA).total_doubles * 8 > 100 * 1024) --G;  // keep two CTAs per SM
    int threads = ((G * per_group + 31) / 32) * 32;
    int gain_warps = G * row_tasks_per_group;
    if (gain_warps > 8) gain_warps = 8;
    if (gains && threads < 32 * gain_warps) threads = 32 * gain_warps;
    if (threads > kRolloutMaxThreads) threads = kRolloutMaxThreads;
    plan.G = G;
    plan.threads = threads;
    plan.grid = (n_list + G - 1) / G;
    plan.smem = rollout_smem(a, s, c, G, NA).total_doubles * 8;
    // pull K[t+1] into L2 a step ahead unless the gains of one step of the whole list would flood the L2 (126 MB)
    plan.prefetch = gains && ((double)expected * a * c * a * s * 8.0 < 48e6) ? 1 : 0;
    (void)T;
    return plan;
}

template <int MC>
int launch_rollout_class(const ForwardParams &p, int expected, cudaStream_t stream);

template <int MC, int NAMAX, bool GAINS>
int launch_rollout_one(ForwardParams p, int expected, cudaStream_t stream)
{
    const Batch &bt = p.batch;
    const RolloutPlan plan = plan_rollout(bt.n_agents, bt.s, bt.c, bt.horizon, p.n_alpha, p.n_list, expected, GAINS);
    if (p.n_alpha * bt.n_agents > kRolloutMaxThreads) {
        set_error("rollout kernel: %d candidates x %d agents exceed %d threads", p.n_alpha, bt.n_agents, kRolloutMaxThreads);
        return DPILQR_E_UNSUPPORTED;
    }
    if (plan.smem > 227 * 1024) {
        set_error("rollout kernel: problem too large for shared memory (%zu bytes)", plan.smem);
        return DPILQR_E_UNSUPPORTED;
    }
    p.groups_per_cta = plan.G;
    p.prefetch = plan.prefetch;
    auto kernel = rollout_kernel<MC, NAMAX, GAINS>;
    if (plan.smem > 48 * 1024)
        DPILQR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    kernel<<<plan.grid, plan.threads, plan.smem, stream>>>(p);
    DPILQR_CUDA(cudaGetLastError());
    return DPILQR_OK;
}

// Definition used by the per-model translation units (rollout_m*.cu)
#define DPILQR_DEFINE_ROLLOUT_CLASS(MC)                                                                   \
    template <>                                                                                           \
    int launch_rollout_class<MC>(const ForwardParams &p, int expected, cudaStream_t stream)               \
    {                                                                                                     \
        if (p.K == nullptr) return launch_rollout_one<MC, 1, false>(p, expected, stream);                 \
        if (p.n_alpha == 1) return launch_rollout_one<MC, 1, true>(p, expected, stream);                  \
        if (p.n_alpha == 2) return launch_rollout_one<MC, 2, true>(p, expected, stream);                  \
        if (p.n_alpha <= 7) return launch_rollout_one<MC, 7, true>(p, expected, stream);                  \
        return launch_rollout_one<MC, 10, true>(p, expected, stream);                                     \
    }

}  // namespace dpilqr
