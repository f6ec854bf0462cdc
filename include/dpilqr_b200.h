/*
 * dpilqr_b200.h -- C ABI of libdpilqr_b200.so (hand-written sm_100a CUDA kernels for
 * the batched iLQR hot path of labicon/dp-ilqr).
 *
 * Conventions (all entry points):
 *   - plain C types only; every array is row-major, contiguous, float64 unless noted;
 *   - "device" pointers are CUDA device pointers (e.g. torch.Tensor.data_ptr()),
 *     "host" pointers are ordinary host memory; the caller owns all memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); device entry
 *     points are asynchronous on it unless stated otherwise;
 *   - return value 0 = success, negative = DPILQR_E_* (no exceptions cross the ABI);
 *     dpilqr_last_error() returns a human-readable message for the calling thread;
 *   - numerical trouble inside a batch never aborts the batch: it is reported per problem
 *     in a `status` array (DPILQR_ST_* bits), the way the reference would have raised or
 *     produced NaN for that one problem.
 *
 * Each entry point names the reference interface (file:line under the reference repo)
 * it replaces.  A "problem" is one (sub)problem of the reference: `a` agents with uniform
 * per-agent state/control sizes (s, c) (reference dynamics.py:165-166), joint sizes
 * n = a*s, m = a*c, horizon T (the reference's N), time step dt.
 */
#ifndef DPILQR_B200_H
#define DPILQR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- model enum: reference bbdynamicswrap.pyx:8-16, Bike5D (dynamics.py:254) appended ---- */
enum {
    DPILQR_MODEL_DOUBLE_INT_4D = 0,
    DPILQR_MODEL_DOUBLE_INT_6D = 1,
    DPILQR_MODEL_CAR_3D = 2,
    DPILQR_MODEL_UNICYCLE_4D = 3,
    DPILQR_MODEL_QUADCOPTER_6D = 4,
    DPILQR_MODEL_HUMAN_6D = 5,
    DPILQR_MODEL_HUMAN_LIN_6D = 6,
    DPILQR_MODEL_QUADCOPTER_12D = 7,
    DPILQR_MODEL_BIKE_5D = 8,
    DPILQR_MODEL_COUNT = 9
};

/* ---- error codes ---- */
enum {
    DPILQR_OK = 0,
    DPILQR_E_INVALID = -1,     /* bad argument / inconsistent descriptor */
    DPILQR_E_UNSUPPORTED = -2, /* problem shape not supported by any kernel */
    DPILQR_E_CUDA = -3,        /* CUDA runtime error (see dpilqr_last_error) */
    DPILQR_E_NO_DEVICE = -4    /* no CUDA device: there is no CPU fallback */
};

/* ---- per-problem status bits ---- */
enum {
    DPILQR_ST_OK = 0,
    DPILQR_ST_NONFINITE = 1,    /* a cost or gain became NaN/Inf */
    DPILQR_ST_SINGULAR = 2,     /* exact zero pivot in Q_uu (reference: LinAlgError, control.py:141) */
    DPILQR_ST_POINT_NDIM = 4,   /* reference would hit `assert point_a.ndim == point_b.ndim` (cost.py:279) */
    DPILQR_ST_CONVERGED = 16,   /* |dJ/J| < tol on the accepting candidate (control.py:184-185) */
    DPILQR_ST_LS_FAILED = 32,   /* all candidates rejected -> bail out (control.py:195-198) */
    DPILQR_ST_ITER_LIMIT = 64,  /* n_lqr_iter reached */
    DPILQR_ST_TIME_LIMIT = 128  /* t_kill wall-clock budget hit (control.py:213-218, batch-wide) */
};

/*
 * One batch ("bin") of problems sharing (a, s, c, T, dt).  All pointers are DEVICE pointers.
 * Replaces the object graph ilqrProblem -> MultiDynamicalModel + GameCost(ReferenceCost[], ProximityCost)
 * (reference problem.py:15-24, dynamics.py:133-146, cost.py:37-66,110-115,174-191).
 */
typedef struct dpilqr_batch {
    int32_t n_problems;      /* B */
    int32_t n_agents;        /* a */
    int32_t s;               /* per-agent state size  (uniform, from agent 0) */
    int32_t c;               /* per-agent control size */
    int32_t horizon;         /* T */
    int32_t n_cost;          /* rows in the Q/R/Qf tables */
    double dt;
    const int32_t *model;    /* [B][a] DPILQR_MODEL_* per agent */
    const int32_t *n_dims;   /* [B][a] 2 or 3: ProximityCost.n_dims (cost.py:111-114) */
    const int32_t *cost_idx; /* [B][a] row of the cost tables used by this agent */
    const double *Q;         /* [n_cost][s][s] ReferenceCost.Q  (dense, may be asymmetric) */
    const double *R;         /* [n_cost][c][c] ReferenceCost.R */
    const double *Qf;        /* [n_cost][s][s] ReferenceCost.Qf */
    const double *xf;        /* [B][n]  goal states */
    const double *radius;    /* [B]     ProximityCost.radius; ignored when a == 1 */
    const double *weights;   /* [B][2]  GameCost.REF_WEIGHT, GameCost.PROX_WEIGHT (cost.py:185-186) */
    const int32_t *has_prox; /* [B] 0: no proximity term (bare ReferenceCost or GameCost(.., None)) */
    int32_t model_hint;      /* 1 + DPILQR_MODEL_* when EVERY agent of EVERY problem runs that model (selects the rollout
                                kernel compiled for it alone); 0 (a zeroed struct): per-agent dispatch in the size class */
    int32_t reserved;        /* 0 */
} dpilqr_batch;

/* ---- solver options: arguments of ilqrSolver.solve (reference control.py:150) ---- */
typedef struct dpilqr_solve_opts {
    int32_t n_lqr_iter; /* default 50 */
    int32_t n_alpha;    /* N_LS_ITER, 1..10 (control.py:51); alphas are the float32 table of control.py:162 */
    double tol;         /* default 1e-3 */
    double t_kill;      /* <= 0: none.  Wall-clock seconds for the whole batch (documented deviation) */
    int32_t record_trace; /* 1: fill the per-iteration trace arrays */
    int32_t profile;      /* 1: time every kernel launch with CUDA events (see dpilqr_get_profile) */
    int32_t bounded_search; /* 1: a line-search candidate stops rolling out as soon as its accumulated cost exceeds the
                               best cost J* -- it is rejected already, because every remaining term is >= 0 -- and its
                               cost reads DPILQR_J_ABORTED.  ONLY valid when (Q+Q^T), (R+R^T), (Qf+Qf^T) are positive
                               semi-definite and both GameCost weights are >= 0 (the caller checks; the Python front
                               door does).  Accepted steps, iteration counts and every returned number are unchanged:
                               the last candidate, whose cost ilqrSolver.solve returns after a failed search
                               (control.py:225), is always rolled out in full. */
    int32_t reserved;
} dpilqr_solve_opts;

/* cost reported for a candidate stopped by bounded_search: "rejected, at least J*" */
#define DPILQR_J_ABORTED 1.7976931348623157e308

/* kernel kinds of the solve loop, index into dpilqr_profile */
enum {
    DPILQR_K_ROLLOUT = 0,    /* warm-start rollout (kernel 1 without gains) */
    DPILQR_K_LINQUAD = 1,    /* kernel 2 */
    DPILQR_K_BACKWARD = 2,   /* kernel 3 */
    DPILQR_K_LINESEARCH = 3, /* kernel 1 with gains, all candidates */
    DPILQR_K_SELECT = 4,     /* accept / regularisation / compaction */
    DPILQR_K_BACKWARD_FULL = 5, /* the kernel-3 launches with at least one problem per SM (counted in kind 2 as well) */
    DPILQR_K_COUNT = 6
};

/* accumulated since the last reset: device milliseconds, number of launches and number of problems
 * processed ("units") per kernel kind, measured with CUDA events on the launching stream */
typedef struct dpilqr_profile {
    double ms[DPILQR_K_COUNT];
    int64_t launches[DPILQR_K_COUNT];
    int64_t units[DPILQR_K_COUNT];
} dpilqr_profile;

const char *dpilqr_last_error(void);
int dpilqr_version(void);
/* number of visible CUDA devices, or DPILQR_E_NO_DEVICE */
int dpilqr_device_count(void);

/* per-model sizes (reference dynamics.py:205-256) */
int dpilqr_model_nx(int model);
int dpilqr_model_nu(int model);

/* doubles per (problem, time step) of the linearise/quadraticise record, see DESIGN.md */
int64_t dpilqr_stage_stride(int n_agents, int s, int c);
/* device workspace (bytes) needed by dpilqr_solve_batch for this batch shape */
int64_t dpilqr_workspace_bytes(int n_problems, int n_agents, int s, int c, int horizon, int n_alpha);

/* ------------------------------------------------------------------------------------------
 * Single-agent dynamics, batched over `count` independent (x, u) samples of one model.
 * Replace bbdynamicswrap.f / integrate / linearize (reference bbdynamicswrap.pyx:61,93,125,
 * bbdynamics.cpp:39-711) and SymbolicModel for Bike5D (dynamics.py:95-114,254-277).
 *   x [count][nx], u [count][nu] -> xdot / x_new [count][nx], A [count][nx][nx], B [count][nx][nu]
 * ------------------------------------------------------------------------------------------ */
int dpilqr_f(int model, int64_t count, const double *x, const double *u, double *xdot, void *stream);
int dpilqr_integrate(int model, double dt, int64_t count, const double *x, const double *u, double *x_new, void *stream);
int dpilqr_linearize(int model, double dt, int64_t count, const double *x, const double *u, double *A, double *B, void *stream);

/* ------------------------------------------------------------------------------------------
 * Kernel 1 -- rollout + line search: replaces ilqrSolver._rollout / _forward_pass
 * (reference control.py:80-114) for every problem and every alpha candidate in ONE launch.
 *   X [B][T+1][n], U [B][T][m]: current trajectories (only X[:,0] is read when K == NULL)
 *   K [B][T][m][n], d [B][T][m]: gains (NULL, NULL => plain rollout of U, n_alpha must be 1)
 *   alphas [n_alpha] (host pointer; promoted float32 values, control.py:162)
 *   Xc [B][n_alpha][T+1][n], Uc [B][n_alpha][T][m], Jc [B][n_alpha]: candidates
 * ------------------------------------------------------------------------------------------ */
int dpilqr_rollout_linesearch(const dpilqr_batch *batch, const double *X, const double *U, const double *K,
                              const double *d, const double *alphas, int n_alpha, double *Xc, double *Uc,
                              double *Jc, void *stream);

/* ------------------------------------------------------------------------------------------
 * Kernel 2 -- fused linearise + quadraticise: replaces MultiDynamicalModel.linearize
 * (dynamics.py:173-186), GameCost.quadraticize (cost.py:208-239), ProximityCost.quadraticize
 * (cost.py:135-171) and quadraticize_distance (cost.py:269-315) for all T+1 steps at once.
 *   stage [B][T+1][dpilqr_stage_stride]: structured (block) output, never the dense n x n.
 *   status [B]: DPILQR_ST_POINT_NDIM / NONFINITE are OR-ed in.
 * ------------------------------------------------------------------------------------------ */
int dpilqr_linearize_quadraticize(const dpilqr_batch *batch, const double *X, const double *U, double *stage,
                                  int32_t *status, void *stream);

/* GameCost value at `rows` arbitrary points per problem: replaces Cost.__call__ (reference cost.py:79-83,
 * 117-133, 197-206).  X [B][rows][n], U [B][rows][m] (ignored when terminal != 0), L [B][rows]. */
int dpilqr_game_cost(const dpilqr_batch *batch, int64_t rows, const double *X, const double *U, int terminal,
                     double *L, void *stream);

/* Dense views of one stage record, for tests / drop-in hooks (cost.quadraticize, dynamics.linearize):
 * A [B][T+1][n][n], Bm [B][T+1][n][m], Lx [B][T+1][n], Lu [B][T+1][m], Lxx [B][T+1][n][n], Luu [B][T+1][m][m]
 * (any output may be NULL). */
int dpilqr_stage_to_dense(const dpilqr_batch *batch, const double *stage, double *A, double *Bm, double *Lx,
                          double *Lu, double *Lxx, double *Luu, void *stream);

/* ------------------------------------------------------------------------------------------
 * Kernel 3 -- backward Riccati recursion, one CTA per problem: replaces
 * ilqrSolver._backward_pass (reference control.py:116-148).
 *   mu [B] regularisation (control.py:123); K [B][T][m][n], d [B][T][m] out; status [B].
 * ------------------------------------------------------------------------------------------ */
int dpilqr_backward(const dpilqr_batch *batch, const double *stage, const double *mu, double *K, double *d,
                    int32_t *status, void *stream);

/* ------------------------------------------------------------------------------------------
 * Kernel 4 -- interaction graph: replaces define_inter_graph_threshold
 * (reference distributed.py:224-247, util.py:48-61) for `n_scen` scenarios at once.
 *   X [n_scen][rows][a*s]; radius [n_scen]; adj [n_scen][a] uint64 bit masks (bit j of adj[k][i]
 *   set iff agent j is in agent i's neighbourhood, self included).  a <= 64.  Bit-exact.
 * ------------------------------------------------------------------------------------------ */
int dpilqr_inter_graph(const double *X, int64_t n_scen, int rows, int n_agents, int s, const double *radius,
                       uint64_t *adj, void *stream);

/* ------------------------------------------------------------------------------------------
 * Scenario generation: replaces random_setup(n_agents, n_states, n_d=n_d, random=True, var=var,
 * energy=energy) (reference util.py:165-195 with randomize_locs :125-132, normalize_energy
 * :203-217, compute_energy :198-200) for the seeds first_seed .. first_seed + count - 1, each
 * scenario as if preceded by np.random.seed(seed).  Bit-identical to the host path (the kernel
 * runs NumPy's legacy MT19937 stream and NumPy's reduction orders).  energy <= 0: no normalisation.
 *   x0, xf [count][n_agents * n_states] out (positions in the first n_d states of each agent).
 * ------------------------------------------------------------------------------------------ */
int dpilqr_random_setup(int64_t first_seed, int64_t count, int n_agents, int n_states, int n_d, double var, double energy,
                        double *x0, double *xf, void *stream);

/* ------------------------------------------------------------------------------------------
 * Whole solve for a batch: replaces ilqrSolver.solve (reference control.py:150-225) including
 * the regularisation schedule (control.py:227-237).  Synchronous on `stream` at return.
 *   x0 [B][n], U0 [B][T][m] in; X [B][T+1][n], U [B][T][m] out;
 *   J [B]: cost of the LAST candidate tried (control.py:225); J_star [B]: best accepted cost;
 *   iters [B]: backward passes executed; status [B];
 *   trace_* (may be NULL unless opts->record_trace): [B][n_lqr_iter] accepted alpha index (-1: failed
 *   search, -2: not executed), mu used, and [B][n_lqr_iter][n_alpha] candidate costs.
 *   workspace: device buffer of dpilqr_workspace_bytes(...).
 * Returns the total number of iterations (backward passes) executed, or a negative error.
 * ------------------------------------------------------------------------------------------ */
int64_t dpilqr_solve_batch(const dpilqr_batch *batch, const dpilqr_solve_opts *opts, const double *x0,
                           const double *U0, double *X, double *U, double *J, double *J_star, int32_t *iters,
                           int32_t *status, int32_t *trace_alpha, double *trace_mu, double *trace_J,
                           void *workspace, int64_t workspace_bytes, void *stream);

/* Same, with HOST buffers for everything (descriptor arrays included): device memory is allocated
 * and cached inside the library, inputs are copied host->device and results device->host on every
 * call.  This is the call a non-Python host would bind. */
int64_t dpilqr_solve_batch_host(const dpilqr_batch *host_batch, const dpilqr_solve_opts *opts, const double *x0,
                                const double *U0, double *X, double *U, double *J, double *J_star,
                                int32_t *iters, int32_t *status, int32_t *trace_alpha, double *trace_mu,
                                double *trace_J, int device);
/* copy (and optionally reset) the accumulated per-kernel timings of solves run with opts->profile = 1 */
int dpilqr_get_profile(dpilqr_profile *out, int reset);
/* Debug aid: when set to a device buffer of 32 int64, backward launches run the instrumented build of the kernel
 * and CTA 0 writes its per-phase cycle counts there (slots 0..11: warp 0; 12..23: first warp of the Q_xx group;
 * 24..29: LU detail and experiments); CTA 0 of every rollout / line-search launch writes slots 24..29 (see
 * tools/backward_phases.py, tools/forward_phases.py).  NULL switches it off. */
int dpilqr_debug_backward_timing(long long *device_counters);
/* release the cached device memory of dpilqr_solve_batch_host */
int dpilqr_release_cache(void);

#ifdef __cplusplus
}
#endif
#endif /* DPILQR_B200_H */
