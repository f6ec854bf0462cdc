// rollout.cuh -- Kernel 1: batched rollout + line search (sm_100a), one instantiation per model / size class.
//
//   u_t = U[t] + K[t] (x_t - X[t]) + alpha d[t],   x_{t+1} = RK4(x_t, u_t),   J += cost(x_t, u_t)
//
// replaces the reference's sequential ilqrSolver._rollout / _forward_pass (reference control.py:80-114) and, per
// step, the Python loops in MultiDynamicalModel.__call__ (dynamics.py:159-171) and GameCost.__call__
// (cost.py:197-206, 79-83, 117-133).
//
// Work decomposition.  A *group* is one problem with a chunk of NA of the launch's line-search candidates (the
// launch's candidates -- 1, 2, 7 in the solver's staged search, 1..10 through the C ABI -- are cut into equal chunks
// so that a group fits the CTA; slots past the last candidate are computed and discarded); a *rollout* is one
// (problem, candidate).  A CTA carries G groups (G is chosen by the launcher so that the grid fills the machine: many
// groups per CTA for thousands of problems, one group per CTA -- minimum latency -- for the stragglers of a solve) and
// walks them through the horizon in lock step.  Per time step:
//   P1  dx = x_t - X[t] per rollout; proximity penalties, one thread per (rollout, agent pair); the cost sum of the
//       previous step in the reference's summation order (agents ascending, pairs in NumPy pairwise order)
//   P2  gain phase: the rows of all the CTA's K[t] matrices are dealt to the warps eight at a time; a quad of lanes
//       walks one row with 16-byte loads straight from global memory (every K element is read once per launch and
//       feeds NA accumulators) and reduces with two shuffles
//   P3  agent phase: reference cost, then the 5-sub-step RK4 with the state in registers.  One thread per
//       (rollout, agent) -- except Quadcopter12D, which a TEAM of three warps integrates, 32 agents at a time, each
//       warp owning a role (models.cuh, quad12_step_team): the rollout is bound by the latency of 20 dependent ODE
//       evaluations per step, and a batch of 4096 ten-drone problems is only 41 k agents, a seventh of the threads
//       the GPU holds.
// Bounded line search (solver path only, ForwardParams::J_bound): every stage cost is >= 0, so a candidate whose
// accumulated cost has passed the problem's best cost J* is rejected whatever follows; it stops integrating (its
// lanes idle like the phantom lanes of a partly filled warp) and a CTA whose candidates have all stopped leaves.
// Runaway candidates -- whose huge angles would send every sin/cos of their warp through the slow library path --
// are gone after a few steps.
// with three block barriers per step.  X[t], U[t], d[t] of the next step arrive by cp.async and the next step's
// gains are pulled into L2 by one bulk prefetch per group while the agents integrate.
#pragma once
#include "cost.cuh"
#include <stdlib.h>

#include "kernels.cuh"

namespace dpilqr {

constexpr int kRolloutMaxThreads = 256;
constexpr int kRolloutMaxTeams = 2;  // team mode: 2 x 32 agents per CTA

__host__ __device__ constexpr bool rollout_team_mode(int mc) { return mc == kQuad12D; }

struct RolloutSmem {
    size_t xcur, dx, ucur, refc, proxc, Jacc, xref, uref, dref, radius, wref, wprox, scratch, ints, gains, total_doubles;
};

// Shared-memory carve-up in doubles for G groups of NA candidates
__host__ __device__ inline RolloutSmem rollout_smem(int a, int s, int c, int G, int NA, bool team, bool stage_gains = false)
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
    L.scratch = off; if (team) off += ((R * a + 31) / 32) * (kTeamScratch + 3 * 32);  // per team: exchange + cost partials
    // ints: problem index, slot, flags, first candidate per group; n_dims per (group, agent); group and candidate
    // per rollout; pair table (i, j) as bytes
    L.ints = off;   off += even(((size_t)4 * G + (size_t)G * a + 3 * R + (P + 1) / 2 + 2) / 2 + 1);
    L.gains = off;  if (stage_gains) off += 2 + (size_t)G * m * n;  // mbarrier + K[t] of every group
    L.total_doubles = off;
    return L;
}

#ifndef DPILQR_ROLLOUT_MINBLOCKS
#define DPILQR_ROLLOUT_MINBLOCKS 2
#endif
#ifndef DPILQR_ROLLOUT_TEAM_THREADS
#define DPILQR_ROLLOUT_TEAM_THREADS 256  // experiment: 192 = exactly two teams, three CTAs per SM (-DDPILQR_ROLLOUT_TEAM_MINBLOCKS=3)
#endif
#ifndef DPILQR_ROLLOUT_TEAM_MINBLOCKS
#define DPILQR_ROLLOUT_TEAM_MINBLOCKS DPILQR_ROLLOUT_MINBLOCKS
#endif
__host__ __device__ constexpr int rollout_max_threads(int mc) { return rollout_team_mode(mc) ? DPILQR_ROLLOUT_TEAM_THREADS : kRolloutMaxThreads; }
__host__ __device__ constexpr int rollout_min_blocks(int mc) { return rollout_team_mode(mc) ? DPILQR_ROLLOUT_TEAM_MINBLOCKS : DPILQR_ROLLOUT_MINBLOCKS; }

template <int MC, int NAMAX, bool GAINS>
__global__ void __launch_bounds__(rollout_max_threads(MC), rollout_min_blocks(MC)) rollout_kernel(const ForwardParams p)
{
    extern __shared__ __align__(16) double smem[];
    const Batch &bt = p.batch;
    constexpr int S = class_nx(MC), C = class_nu(MC);
    constexpr bool TEAM = rollout_team_mode(MC);
    const int a = bt.n_agents, T = bt.horizon;
    const int n = a * S, m = a * C, pairs = a * (a - 1) / 2, P = pairs > 0 ? pairs : 1;
    const int NA = p.chunk_alpha, G = p.groups_per_cta, n_chunks = p.n_chunks;
    const int count = (p.n_active ? min(*p.n_active, p.n_list) : p.n_list) * n_chunks;  // groups of the launch
    const int g0 = blockIdx.x * G;
    if (g0 >= count) return;
    const int Ge = min(G, count - g0);  // groups of this CTA
    const int R = Ge * NA;              // rollouts of this CTA
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;

    // Offsets of the carve-up, made opaque: left to itself the compiler re-derives them (a hundred 64-bit integer
    // instructions) at every use inside the time loop instead of spending registers on them.
    int o_xcur, o_dx, o_ucur, o_refc, o_proxc, o_Jacc, o_xref, o_uref, o_dref, o_radius, o_scratch, o_ints, o_gains;
    {
        const RolloutSmem SM = rollout_smem(a, S, C, G, NA, TEAM, GAINS && p.stage_gains);
        o_xcur = (int)SM.xcur; o_dx = (int)SM.dx; o_ucur = (int)SM.ucur; o_refc = (int)SM.refc; o_proxc = (int)SM.proxc;
        o_Jacc = (int)SM.Jacc; o_xref = (int)SM.xref; o_uref = (int)SM.uref; o_dref = (int)SM.dref; o_radius = (int)SM.radius;
        o_scratch = (int)SM.scratch; o_ints = (int)SM.ints; o_gains = (int)SM.gains;
        asm volatile("" : "+r"(o_xcur), "+r"(o_dx), "+r"(o_ucur), "+r"(o_refc), "+r"(o_proxc), "+r"(o_Jacc));
        asm volatile("" : "+r"(o_xref), "+r"(o_uref), "+r"(o_dref), "+r"(o_radius), "+r"(o_scratch), "+r"(o_ints), "+r"(o_gains));
    }
    const int Gpad = (G + 1) & ~1;
    double *xcur = smem + o_xcur;     // [R][n]
    double *dx = smem + o_dx;         // [R][n]
    double *ucur = smem + o_ucur;     // [R][m]
    double *refc = smem + o_refc;     // [2][R][a]   by step parity
    double *proxc = smem + o_proxc;   // [2][R][P]
    double *Jacc = smem + o_Jacc;     // [R]
    double *xref = smem + o_xref;     // [Ge][n]
    double *uref = smem + o_uref;     // [Ge][m]
    double *dref = smem + o_dref;     // [Ge][m]
    double *g_radius = smem + o_radius, *g_wref = g_radius + Gpad, *g_wprox = g_wref + Gpad;
    int *g_prob = reinterpret_cast<int *>(smem + o_ints);   // [G] problem index
    int *g_slot = g_prob + G;                               // [G] trajectory slot
    int *g_flags = g_slot + G;                              // [G] bit 0: proximity term, bit 1: planar distance for all pairs
    int *g_abase = g_flags + G;                             // [G] first candidate of the group's chunk
    int *g_ndims = g_abase + G;                             // [G][a]
    int *r_group = g_ndims + (size_t)G * a;                 // [G * NA] group of a rollout
    int *r_cand = r_group + (size_t)G * NA;                 // [G * NA] candidate slot of a rollout in the launch
    int *r_dead = r_cand + (size_t)G * NA;                  // [G * NA] stopped by the bounded line search
    unsigned char *pair_ij = reinterpret_cast<unsigned char *>(r_dead + (size_t)G * NA);  // [P][2]

    // ---- per-group constants
    for (int g = tid; g < Ge; g += nthr) {
        const int li = (g0 + g) / n_chunks;
        const int b = p.active ? p.active[li] : li;
        g_prob[g] = b;
        g_abase[g] = ((g0 + g) - li * n_chunks) * NA;
        for (int al = 0; al < NA; ++al) {
            r_group[g * NA + al] = g;
            r_cand[g * NA + al] = ((g0 + g) - li * n_chunks) * NA + al;
            r_dead[g * NA + al] = 0;
        }
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
    auto fetch_refs = [&](int t) {  // a warp takes whole groups: no per-element index arithmetic
        if (t < T) {
            for (int g = warp; g < Ge; g += nwarp) {
                if constexpr (GAINS) {
                    const double *xs = x_traj(g) + (int64_t)t * n;
                    const double *ds = p.d + ((int64_t)g_prob[g] * T + t) * m;
                    for (int j = lane; j < n; j += 32) cp_async8(xref + g * n + j, xs + j);
                    for (int j = lane; j < m; j += 32) cp_async8(dref + g * m + j, ds + j);
                }
                const double *us = u_traj(g) + (int64_t)t * m;
                for (int j = lane; j < m; j += 32) cp_async8(uref + g * m + j, us + j);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto prefetch_gains = [&](int t) {  // one bulk L2 prefetch per group
        if constexpr (GAINS) {
            if (p.prefetch == 1 && t < T && tid < Ge) {
                const double *src = p.K + ((int64_t)g_prob[tid] * T + t) * m * n;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((unsigned)(m * n * 8)) : "memory");
            } else if (p.prefetch == 2 && t < T) {  // one prefetch instruction per 128-byte line
                const int lines = (m * n * 8 + 127) >> 7;
                for (int g = warp; g < Ge; g += nwarp) {
                    const char *src = reinterpret_cast<const char *>(p.K + ((int64_t)g_prob[g] * T + t) * m * n);
                    for (int l = lane; l < lines; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + ((size_t)l << 7)));
                }
            }
        }
    };
    // Few groups per CTA (the stragglers of a solve, small teams): K[t] of every group is staged in shared memory by
    // TMA bulk copies issued a step ahead, so that the gain phase -- on the serial path x_t -> u_t -> x_{t+1} -- reads
    // shared memory instead of waiting for L2 / HBM.
    const bool staged = GAINS && p.stage_gains;
    double *Ks = smem + o_gains + 2;  // [Ge][m][n]
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem + o_gains);
    // (warp 0, after a block barrier behind the last reads of the previous K.  Lane 0 arms the barrier -- its arrival is
    // the only one, so the phase cannot complete before the expected byte count is posted -- then the lanes issue the
    // copies of their groups side by side: with dozens of small problems per CTA one thread issuing every copy was a
    // quarter of the step.)
    auto stage_gains = [&](int t) {
        if constexpr (GAINS) {
            if (staged && t < T && warp == 0) {
                const unsigned bytes = (unsigned)(m * n * 8);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (lane == 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes * Ge) : "memory");
                __syncwarp();
                for (int g = lane; g < Ge; g += 32)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"((unsigned)__cvta_generic_to_shared(Ks + (size_t)g * m * n)),
                                   "l"(p.K + ((int64_t)g_prob[g] * T + t) * m * n), "r"(bytes), "r"(mbar) : "memory");
            }
        }
    };
    if constexpr (GAINS) {
        if (staged) {
            if (tid == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            stage_gains(0);
        }
    }
    // candidate slot of rollout r in the launch (slots past the last candidate are computed but never stored)
    auto rollout_cand = [&](int r) { return r_cand[r]; };

    fetch_refs(0);
    if (!staged) prefetch_gains(0);
    for (int r = warp; r < R; r += nwarp) {  // X_next[0] = X[0]
        const double *src = x_traj(r_group[r]);
        for (int j = lane; j < n; j += 32) xcur[r * n + j] = src[j];
    }
    for (int r = tid; r < R; r += nthr) Jacc[r] = 0.0;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    // ---- one thread per (rollout, agent) [thread mode]: the agent this thread integrates, fixed for the horizon
    const bool is_agent = !TEAM && tid < R * a;
    const int ag_r = is_agent ? tid / a : 0;          // rollout
    const int ag_i = is_agent ? tid - ag_r * a : 0;   // agent
    const int ag_b = g_prob[r_group[ag_r]];
    const int ag_model = (MC >= kMixed4) ? bt.model[(int64_t)ag_b * a + ag_i] : MC;
    const int64_t ag_ci = bt.cost_idx[(int64_t)ag_b * a + ag_i];
    const double *ag_xf = bt.xf + (int64_t)ag_b * n + ag_i * S;
    const int ag_cand = rollout_cand(ag_r);
    const bool ag_store = ag_cand < p.n_alpha;
    double *ag_Xout = p.Xc + (int64_t)ag_b * p.xc_stride + (int64_t)(p.alpha_first + ag_cand) * (T + 1) * n + ag_i * S;
    double *ag_Uout = p.Uc + (int64_t)ag_b * p.uc_stride + (int64_t)(p.alpha_first + ag_cand) * T * m + ag_i * C;

    // ---- team mode: the agent of this lane (the same for the three warps of a team)
    const int tm_slot = (warp / kTeamRoles) * 32 + lane;
    const bool tm_real = TEAM && tm_slot < R * a;
    const int tm_r = tm_real ? tm_slot / a : 0, tm_i = tm_real ? tm_slot - tm_r * a : 0;

    auto sum_step = [&](int par) {  // PROX_WEIGHT * prox + REF_WEIGHT * ref_total (cost.py:206), one thread per rollout
        for (int r = tid; r < R; r += nthr) {
            const int g = r_group[r];
            const double *rc = refc + ((size_t)par * R + r) * a;
            double ref_total = 0.0;
            for (int i = 0; i < a; ++i) ref_total += rc[i];
            const double prox = (g_flags[g] & 1) ? numpy_pairwise_sum(proxc + ((size_t)par * R + r) * P, pairs) : 0.0;
            if (!r_dead[r]) {
                const double J = Jacc[r] + (g_wprox[g] * prox + g_wref[g] * ref_total);
                Jacc[r] = J;
                if (p.J_bound != nullptr && p.alpha_first + r_cand[r] != p.exempt_cand && J > p.J_bound[g_prob[g]]) r_dead[r] = 1;
            }
        }
    };

    // optional per-phase cycle counters of CTA 0 / thread 0 (debug aid, see dpilqr_debug_backward_timing)
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tmark = 0;
    const bool timing = (p.timing != nullptr) && (blockIdx.x == 0) && (tid == 0);
    auto tick = [&](int slot) {
        if (timing) {
            const long long now = clock64();
            tacc[slot] += now - tmark;
            tmark = now;
        }
    };
    if (timing) tmark = clock64();
#pragma unroll 1
    for (int t = 0; t <= T; ++t) {
        const bool terminal = (t == T);
        const int par = t & 1;
        // ================= P1: cost sum of step t-1, dx, proximity penalties =================
        if (t > 0) sum_step(par ^ 1);
        for (int r = warp; r < R; r += nwarp) {  // a warp takes whole rollouts: no per-element index arithmetic
            const int g = r_group[r];
            const double *xr = xcur + r * n;
            bool store = false;
            double *xo = nullptr;
            if constexpr (TEAM) {  // the candidate states stream out from shared memory, coalesced
                const int cand = r_cand[r];
                store = cand < p.n_alpha;
                xo = p.Xc + (int64_t)g_prob[g] * p.xc_stride + ((int64_t)(p.alpha_first + cand) * (T + 1) + t) * n;
            }
            for (int j = lane; j < n; j += 32) {
                const double v = xr[j];
                if constexpr (GAINS) {
                    if (!terminal) dx[r * n + j] = v - xref[g * n + j];
                }
                if (store) xo[j] = v;
            }
            if (g_flags[g] & 1) {  // fmin(0, dist - radius)^2 (cost.py:117-133, util.py:48-87)
                const bool planar = (g_flags[g] & 2) != 0;
                const double radius = g_radius[g];
                double *pc = proxc + ((size_t)par * R + r) * P;
                for (int pr = lane; pr < pairs; pr += 32) {
                    const int i = pair_ij[2 * pr], j = pair_ij[2 * pr + 1];
                    const int nd = planar ? 2 : min(g_ndims[g * a + i], g_ndims[g * a + j]);
                    pc[pr] = pair_penalty(xr + i * S, xr + j * S, nd, radius);
                }
            }
        }
        tick(0);
        __syncthreads();
        tick(1);
        // ================= P2: controls of this step =================
        if (!terminal) {
            if constexpr (GAINS) {
                if (staged) {
                    asm volatile(
                        "{\n"
                        ".reg .pred p;\n"
                        "WAIT_GAINS:\n"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                        "@p bra DONE_GAINS;\n"
                        "bra WAIT_GAINS;\n"
                        "DONE_GAINS:\n"
                        "}\n" ::"r"(mbar), "r"(t & 1) : "memory");
                }
                const int q = lane & 3;
                const int rows_total = Ge * m;
                for (int task = warp; task * 8 < rows_total; task += nwarp) {
                    const int grow = task * 8 + (lane >> 2);
                    const bool live = grow < rows_total;
                    const int g = live ? grow / m : 0, row = live ? grow - g * m : 0;
                    const double *Krow = staged ? Ks + ((size_t)g * m + row) * n : p.K + (((int64_t)g_prob[g] * T + t) * m + row) * n;
                    const double *dxg = dx + (size_t)g * NA * n;
                    double acc[NAMAX];
#pragma unroll
                    for (int al = 0; al < NAMAX; ++al) acc[al] = 0.0;
                    if (live) {
                        if ((n & 1) == 0) {
                            // all the loads of a batch of 8 column pairs are issued before the first one is used
                            // (an HBM round trip per batch, not per load)
                            constexpr int KV = 8;
                            for (int c0 = 2 * q; c0 < n; c0 += 8 * KV) {
                                double2 kv[KV];
#pragma unroll
                                for (int j = 0; j < KV; ++j) {
                                    const int col = c0 + 8 * j;
                                    kv[j] = (col < n) ? *reinterpret_cast<const double2 *>(Krow + col) : make_double2(0.0, 0.0);
                                }
#pragma unroll
                                for (int j = 0; j < KV; ++j) {
                                    const int col = c0 + 8 * j;
                                    if (col < n) {
#pragma unroll
                                        for (int al = 0; al < NAMAX; ++al) {
                                            if (al < NA) {
                                                const double2 dv = *reinterpret_cast<const double2 *>(dxg + (size_t)al * n + col);
                                                acc[al] = fma(kv[j].y, dv.y, fma(kv[j].x, dv.x, acc[al]));
                                            }
                                        }
                                    }
                                }
                            }
                        } else {
#pragma unroll 4
                            for (int col = q; col < n; col += 4) {
                                const double kv = Krow[col];
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
                            if (live && (al & 3) == q) {
                                const int cand = min(r_cand[g * NA + al], kMaxAlpha - 1);
                                ucur[((size_t)g * NA + al) * m + row] = uref[g * m + row] + (v + p.alpha[cand] * dref[g * m + row]);
                            }
                        }
                    }
                }
            } else {
                for (int r = warp; r < R; r += nwarp)
                    for (int j = lane; j < m; j += 32) ucur[r * m + j] = uref[r_group[r] * m + j];
            }
            tick(2);
            __syncthreads();
            tick(3);
            // the reference rows of this step are consumed: fetch the next ones behind the integration
            fetch_refs(t + 1);
            if (staged) stage_gains(t + 1);
            else prefetch_gains(t + 1);
            if constexpr (TEAM) {
                for (int r = warp; r < R; r += nwarp) {
                    const int cand = r_cand[r];
                    if (cand < p.n_alpha) {
                        double *uo = p.Uc + (int64_t)g_prob[r_group[r]] * p.uc_stride + ((int64_t)(p.alpha_first + cand) * T + t) * m;
                        for (int j = lane; j < m; j += 32) uo[j] = ucur[r * m + j];
                    }
                }
            }
        }
        tick(4);
        // ================= P3: reference cost at (x_t, u_t), then x_{t+1} = RK4(x_t, u_t) =================
        if constexpr (TEAM) {
            // Quadcopter12D: a team of three warps per 32 agents, one role per warp (models.cuh)
            const int n_ag = R * a;
            const int team = warp / kTeamRoles, role = warp - team * kTeamRoles;
            if (team * 32 < n_ag) {
                const int r = tm_r, i = tm_i;
                const bool real = tm_real && !r_dead[r];
                const int b = g_prob[r_group[r]];
                double *xs_r = xcur + (size_t)r * n + i * S;
                const double *xf = bt.xf + (int64_t)b * n + i * S;
                const int64_t ci = bt.cost_idx[(int64_t)b * a + i];
                double *tscr = smem + o_scratch + team * (kTeamScratch + 3 * 32);
                double *partial = tscr + kTeamScratch;  // [3][32] cost partials of the roles
                const int off = team_role_offset(role), nc = team_role_count(role);
                double x[6], u[4];
#pragma unroll
                for (int k = 0; k < 6; ++k) x[k] = (real && k < nc) ? xs_r[off + k] : 0.0;
                u[0] = u[1] = u[2] = u[3] = 0.0;
                if (!terminal && real) {
                    const double *us = ucur + (size_t)r * m + i * C;
                    const double2 ua = *reinterpret_cast<const double2 *>(us);
                    const double2 ub = *reinterpret_cast<const double2 *>(us + 2);
                    u[0] = ua.x; u[1] = ua.y; u[2] = ub.x; u[3] = ub.y;
                }
                // reference cost (cost.py:79-83): every role sums the columns of its state slice, role 2 adds the
                // control term, role 0 adds the three partials up
                double part = 0.0;
                if (real) {
                    const double *Qm = (terminal ? bt.Qf : bt.Q) + ci * 144;
                    double e[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) e[k] = xs_r[k] - xf[k];
                    for (int jj = 0; jj < nc; ++jj) {
                        const int j = off + jj;
                        double v = 0.0;
#pragma unroll
                        for (int k = 0; k < 12; ++k) v = fma(e[k], Qm[k * 12 + j], v);
                        part = fma(v, xs_r[j] - xf[j], part);
                    }
                    if (role == 2 && !terminal) {
                        const double *Rm = bt.R + ci * 16;
                        double cu = 0.0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            double v = 0.0;
#pragma unroll
                            for (int k = 0; k < 4; ++k) v = fma(u[k], Rm[k * 4 + j], v);
                            cu = fma(v, u[j], cu);
                        }
                        part += cu;
                    }
                }
                partial[role * 32 + lane] = part;
                team_barrier(1 + team);
                if (role == 0 && real) refc[((size_t)par * R + r) * a + i] = (partial[lane] + partial[32 + lane]) + partial[64 + lane];
                tick(5);
                if (!terminal) {
                    if (role == 0) quad12_step_team<0>(bt.dt, lane, 1 + team, x, u, tscr);
                    else if (role == 1) quad12_step_team<1>(bt.dt, lane, 1 + team, x, u, tscr);
                    else quad12_step_team<2>(bt.dt, lane, 1 + team, x, u, tscr);
                    if (real) {
#pragma unroll
                        for (int k = 0; k < 6; ++k)
                            if (k < nc) xs_r[off + k] = x[k];
                    }
                }
            }
        } else if (is_agent && !r_dead[ag_r]) {
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
                        if (ag_store) *reinterpret_cast<double2 *>(xo + k) = v;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < NX; ++k) { x[k] = xs[k]; if (ag_store) xo[k] = x[k]; }
                }
                if (!terminal) {
                    double *uo = ag_Uout + (int64_t)t * m;
#pragma unroll
                    for (int k = 0; k < NU; ++k) { u[k] = us[k]; if (ag_store) uo[k] = u[k]; }
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
        tick(6);
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (p.J_bound != nullptr) {  // every candidate of the CTA stopped: nothing left to do
            int alive = 0;
            for (int r = tid; r < R; r += nthr) alive |= !r_dead[r];
            if (!__syncthreads_or(alive)) {
                if constexpr (GAINS) {
                    if (staged && t + 1 < T) {  // the bulk copy of the next step's gains is in flight: let it land before leaving
                        asm volatile(
                            "{\n"
                            ".reg .pred p;\n"
                            "WAIT_LAST_GAINS:\n"
                            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                            "@p bra DONE_LAST_GAINS;\n"
                            "bra WAIT_LAST_GAINS;\n"
                            "DONE_LAST_GAINS:\n"
                            "}\n" ::"r"(mbar), "r"((t + 1) & 1) : "memory");
                    }
                }
                break;
            }
        } else {
            __syncthreads();
        }
        tick(7);
    }
    if (timing) {
        for (int k = 0; k < 8; ++k) p.timing[24 + k] = tacc[k];
    }
    sum_step(T & 1);  // (harmless after an early exit: every candidate is stopped)
    __syncthreads();
    for (int r = tid; r < R; r += nthr) {
        const int cand = rollout_cand(r);
        if (cand < p.n_alpha) p.Jc[(int64_t)g_prob[r_group[r]] * p.jc_stride + p.alpha_first + cand] = r_dead[r] ? DPILQR_J_ABORTED : Jacc[r];
    }
}

// ---- launch plan ---------------------------------------------------------------------------------------------
struct RolloutPlan {
    int G, threads, grid, prefetch, chunk_alpha, n_chunks, stage_gains;
    size_t smem;
};

inline RolloutPlan plan_rollout(int a, int s, int c, int n_alpha, int n_list, int expected, bool gains, bool team)
{
    RolloutPlan plan{};
    // agents a CTA can integrate at once: one thread each, or 32 per team of three warps
    const int agent_slots = team ? 32 * kRolloutMaxTeams : kRolloutMaxThreads;
    // candidates per group: as many as fit the CTA, in equal chunks
    int fit = agent_slots / a;
    if (fit < 1) fit = 1;
    const int n_chunks = (n_alpha + fit - 1) / fit;
    const int NA = (n_alpha + n_chunks - 1) / n_chunks;
    const int row_tasks_per_group = (a * c + 7) / 8;  // gain-phase warp tasks of one group
    int Gmax = agent_slots / (NA * a);
    if (Gmax < 1) Gmax = 1;
    // fill the machine: about three CTAs per SM; stragglers get a CTA (and its warps' gain tasks) to themselves
    const int slots = 148 * 3;
    int G = (int)(((int64_t)expected * n_chunks + slots - 1) / slots);
    if (G < 1) G = 1;
    if (G > Gmax) G = Gmax;
    while (G > 1 && rollout_smem(a, s, c, G, NA, team).total_doubles * 8 > 72 * 1024) --G;  // keep three CTAs per SM
    const int agents = G * NA * a;
    int threads = team ? ((agents + 31) / 32) * kTeamThreads : ((agents + 31) / 32) * 32;
    int gain_warps = G * row_tasks_per_group;
    if (gain_warps > 8) gain_warps = 8;
    if (gains && threads < 32 * gain_warps) threads = 32 * gain_warps;
    const int max_threads = team ? DPILQR_ROLLOUT_TEAM_THREADS : kRolloutMaxThreads;
    if (threads > max_threads) threads = max_threads;
    plan.G = G;
    plan.threads = threads;
    plan.chunk_alpha = NA;
    plan.n_chunks = n_chunks;
    plan.grid = (int)(((int64_t)n_list * n_chunks + G - 1) / G);
    plan.stage_gains = gains && ((size_t)G * a * c * a * s * 8 <= 48 * 1024) ? 1 : 0;
    plan.smem = rollout_smem(a, s, c, G, NA, team, plan.stage_gains != 0).total_doubles * 8;
    // pull K[t+1] into L2 a step ahead unless the gains of one step of the whole list would flood the L2 (126 MB)
    plan.prefetch = gains && ((double)expected * a * c * a * s * 8.0 < 48e6) ? 1 : 0;
    if (const char *env = getenv("DPILQR_ROLLOUT_PREFETCH")) plan.prefetch = gains ? atoi(env) : 0;  // experiments
    return plan;
}

template <int MC>
int launch_rollout_class(const ForwardParams &p, int expected, cudaStream_t stream);

template <int MC, int NAMAX, bool GAINS>
int launch_rollout_one(ForwardParams p, const RolloutPlan &plan, cudaStream_t stream)
{
    if (plan.smem > 227 * 1024) {
        set_error("rollout kernel: problem too large for shared memory (%zu bytes)", plan.smem);
        return DPILQR_E_UNSUPPORTED;
    }
    p.groups_per_cta = plan.G;
    p.prefetch = plan.prefetch;
    p.chunk_alpha = plan.chunk_alpha;
    p.n_chunks = plan.n_chunks;
    p.stage_gains = plan.stage_gains;
    auto kernel = rollout_kernel<MC, NAMAX, GAINS>;
    if (plan.smem > 48 * 1024)
        // (the ceiling, not this launch's need: concurrent callers must not lower it under each other)
        DPILQR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kernel<<<plan.grid, plan.threads, plan.smem, stream>>>(p);
    DPILQR_CUDA(cudaGetLastError());
    return DPILQR_OK;
}

// Definition used by the per-model translation units (rollout_inst.cu)
#define DPILQR_DEFINE_ROLLOUT_CLASS(MC)                                                                             \
    template <>                                                                                                     \
    int launch_rollout_class<MC>(const ForwardParams &p, int expected, cudaStream_t stream)                         \
    {                                                                                                               \
        const Batch &bt = p.batch;                                                                                  \
        const RolloutPlan plan = plan_rollout(bt.n_agents, bt.s, bt.c, p.n_alpha, p.n_list, expected, p.K != nullptr, \
                                              rollout_team_mode(MC));                                               \
        if (plan.chunk_alpha * bt.n_agents > (rollout_team_mode(MC) ? 32 * kRolloutMaxTeams : kRolloutMaxThreads)) { \
            set_error("rollout kernel: %d agents do not fit a CTA", bt.n_agents);                                   \
            return DPILQR_E_UNSUPPORTED;                                                                            \
        }                                                                                                           \
        if (p.K == nullptr) return launch_rollout_one<MC, 1, false>(p, plan, stream);                               \
        if (plan.chunk_alpha == 1) return launch_rollout_one<MC, 1, true>(p, plan, stream);                         \
        if (plan.chunk_alpha == 2) return launch_rollout_one<MC, 2, true>(p, plan, stream);                         \
        if (plan.chunk_alpha <= 4) return launch_rollout_one<MC, 4, true>(p, plan, stream);                         \
        if (plan.chunk_alpha <= 6) return launch_rollout_one<MC, 6, true>(p, plan, stream);                         \
        return launch_rollout_one<MC, 10, true>(p, plan, stream);                                                   \
    }

}  // namespace dpilqr
