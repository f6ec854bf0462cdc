// kernels.cuh -- parameter blocks and host launchers shared by the translation units.
#pragma once
#include "common.cuh"

namespace dpilqr {

constexpr int kMaxAlpha = 10;

// Trajectories are addressed as  base + b * stride + slot[b] * slot_stride  so the solver loop can
// keep every problem's current trajectory inside the ping-pong candidate buffers without copies.
//
// One launch of the rollout kernel evaluates, for every problem of a list, the `n_alpha` line-search candidates
// alpha[0..n_alpha-1]; candidate k lands in output slot alpha_first + k (the solver's staged line search launches
// the kernel up to three times per iteration with different candidate ranges and shrinking lists).
struct ForwardParams {
    Batch batch;
    const double *X;  // current trajectories (only X[:, 0] is read when K == nullptr)
    const double *U;
    const double *K;  // may be null (plain rollout)
    const double *d;
    double *Xc;       // candidates out
    double *Uc;
    double *Jc;
    int64_t x_stride, u_stride;               // per-problem strides of X / U (doubles)
    int64_t xc_stride, uc_stride, jc_stride;  // per-problem strides of the candidate outputs
    int64_t x_slot_stride, u_slot_stride;
    const int32_t *slot;      // may be null
    const int32_t *active;    // may be null: list of problem indices
    const int32_t *n_active;  // may be null: device-side length of `active`
    int n_list;               // host-side length of the list (upper bound when n_active is given)
    int n_alpha;              // candidates per problem in this launch
    int alpha_first;          // output slot of candidate 0 of this launch
    int uniform_model;        // model id shared by every agent of the batch, or -1 (mixed team / unknown)
    int groups_per_cta;       // filled in by the launcher
    int prefetch;             // filled in by the launcher: L2 prefetch of the next step's gains
    int chunk_alpha;          // filled in by the launcher: candidates per group (the launch's candidates in equal chunks)
    int n_chunks;             // filled in by the launcher
    int stage_gains;          // filled in by the launcher: K[t] of the CTA's groups staged in shared memory by TMA, a step ahead
    const double *J_bound;    // may be null: per-problem cost bound of the bounded line search (the solver's J*)
    int exempt_cand;          // output slot never stopped by the bound (the solver's last candidate), or -1
    double alpha[kMaxAlpha];
    long long *timing;        // optional debug cycle counters of CTA 0 / thread 0 (slots 24..31 of the dpilqr_debug_backward_timing buffer)
};

// expected_list: the launcher sizes the groups per CTA for this many problems (the grid always covers n_list)
int launch_forward(const ForwardParams &p, int expected_list, cudaStream_t stream);

struct LinQuadParams {
    Batch batch;
    const double *X;
    const double *U;
    double *stage;
    int32_t *status;
    int64_t x_stride, u_stride, x_slot_stride, u_slot_stride;
    const int32_t *slot;
    const int32_t *active;
    const int32_t *n_active;
    int n_blocks_per_problem;
    int records_per_cta;  // consecutive stage records a CTA builds in shared memory and stores in one bulk copy
    int a_layout;         // agents of the record LAYOUT (0: the batch's): padded by a phantom agent for odd teams (backward.cu)
};

struct BackwardParams {
    Batch batch;
    const double *stage;
    const double *mu;
    double *K;
    double *d;
    int32_t *status;
    const int32_t *active;
    const int32_t *n_active;
    double *scratch;  // global scratch for the big-problem path (2*m*n doubles per CTA)
    int a_layout;     // agents of the stage-record layout (0: the batch's); n_agents + 1: records padded by a phantom agent
    int n_launch;     // problems of this launch (kernels whose grid is rounded up to whole CTAs of several problems)
    int use_global_scratch;
    long long *timing;  // optional: 20 per-phase cycle counters written by CTA 0 (debug aid)
    int debug_mode;     // timing experiments of the instrumented build only (wrong numerics), env DPILQR_DEBUG_BACKWARD_MODE:
                        // 4 a second LU call, 8 no regularisation pass, 16 no Q_xx, 64 LU of a synthetic matrix,
                        // 32 / 256 / 512 the LU alone before the recursion / at the top of / inside a step
};
extern long long *g_backward_timing;
extern int g_backward_debug_mode;

int launch_linquad(const LinQuadParams &p, int n_problems, cudaStream_t stream);
int launch_stage_to_dense(const Batch &bt, const double *stage, double *A, double *Bm, double *Lx, double *Lu,
                          double *Lxx, double *Luu, cudaStream_t stream);
int launch_backward(const BackwardParams &p, int n_blocks, cudaStream_t stream);
int backward_layout_agents(int a, int s, int c);   // agents of the stage-record layout the solver's backward kernel wants
bool backward_small_applies(int a, int s, int c);
bool backward_warp_applies(int a, int s, int c);  // tiny problems: one warp per problem (backward_warp.cu)
int launch_backward_warp(const BackwardParams &p, int n_blocks, cudaStream_t stream);  // backward_small.cu: n <= 64, m <= 32
int launch_backward_small(const BackwardParams &p, int n_blocks, cudaStream_t stream);
int64_t backward_scratch_doubles(int n_problems, int a, int s, int c);
int launch_random_setup(int64_t first_seed, int64_t count, int a, int s, int n_d, double var, double energy, double *x0,
                        double *xf, cudaStream_t stream);  // scenario.cu
int launch_inter_graph(const double *X, int64_t n_scen, int rows, int a, int s, const double *radius, uint64_t *adj,
                       cudaStream_t stream);
int launch_game_cost(const Batch &bt, int64_t rows, const double *X, const double *U, int terminal, double *L,
                     cudaStream_t stream);
int launch_dynamics(int mode, int model, double dt, int64_t count, const double *x, const double *u, double *out0,
                    double *out1, cudaStream_t stream);

}  // namespace dpilqr
