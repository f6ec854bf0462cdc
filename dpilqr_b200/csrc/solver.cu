// solver.cu -- C ABI of libdpilqr_b200.so and the batched iLQR driver loop (sm_100a).
//
// dpilqr_solve_batch replaces ilqrSolver.solve (reference control.py:150-225) for a whole batch:
// the trajectories of every problem live in two ping-pong candidate buffers on the device; an
// iteration is  linearise/quadraticise -> backward Riccati -> all-candidate line search ->
// select, with the accept / converge / bail-out decisions and the regularisation schedule
// (reference control.py:179-211, 227-237) taken per problem by the select kernel.  Finished
// problems leave a compacted active list; the host only reads back one integer per iteration.
#include <stdarg.h>
#include <stdlib.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <vector>

#include "kernels.cuh"

namespace dpilqr {

// float32 values of 1.1 ** (-k**2), k = 0..9, promoted to double (reference control.py:162)
const double kAlphaTable[10] = {0x1p+0,         0x1.d1745cp-1, 0x1.5db3eep-1, 0x1.b246ap-2,  0x1.bdb44ep-3,
                                0x1.7a0b5p-4,   0x1.09011ap-5, 0x1.330c94p-7, 0x1.260538p-9, 0x1.d15cep-12};

// ---- error plumbing ---------------------------------------------------------------------------
static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t err, const char *what)
{
    if (err == cudaSuccess) return 0;
    set_error("CUDA error %s (%s) in %s", cudaGetErrorName(err), cudaGetErrorString(err), what);
    return (err == cudaErrorNoDevice || err == cudaErrorInsufficientDriver) ? DPILQR_E_NO_DEVICE : DPILQR_E_CUDA;
}

int validate_batch(const dpilqr_batch *b)
{
    if (!b) { set_error("null batch descriptor"); return DPILQR_E_INVALID; }
    if (b->n_problems < 0 || b->n_agents < 1 || b->s < 2 || b->c < 1 || b->horizon < 1 || b->n_cost < 1) {
        set_error("invalid batch shape: B=%d a=%d s=%d c=%d T=%d n_cost=%d", b->n_problems, b->n_agents, b->s, b->c,
                  b->horizon, b->n_cost);
        return DPILQR_E_INVALID;
    }
    if (!(b->dt > 0.0)) { set_error("dt must be positive"); return DPILQR_E_INVALID; }
    if (!b->model || !b->n_dims || !b->cost_idx || !b->Q || !b->R || !b->Qf || !b->xf) {
        set_error("batch descriptor has null arrays");
        return DPILQR_E_INVALID;
    }
    if (b->n_agents > 1 && !b->radius) { set_error("radius array required for multi-agent problems"); return DPILQR_E_INVALID; }
    bool known = false;
    for (int mdl = 0; mdl < kModelCount; ++mdl) known = known || (model_nx(mdl) == b->s && model_nu(mdl) == b->c);
    if (!known) { set_error("no model has per-agent dimensions (%d, %d)", b->s, b->c); return DPILQR_E_UNSUPPORTED; }
    return 0;
}

// ---- workspace layout --------------------------------------------------------------------------
struct Workspace {
    double *candX[2], *candU[2];
    double *Jc, *K, *d, *stage, *scratch, *mu, *delta, *Jstar, *Jlast;
    int32_t *slot, *parity, *active[2], *n_active, *flag, *ls_list[2], *ls_count;
    int64_t bytes;
};

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static Workspace carve(char *base, int B, int a, int s, int c, int T, int NA)
{
    Workspace w;
    const int64_t n = (int64_t)a * s, m = (int64_t)a * c;
    const StageLayout L = stage_layout(backward_layout_agents(a, s, c), s, c);  // (odd teams: records padded by a phantom agent)
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        char *ptr = base ? base + off : nullptr;
        off = align_up(off + bytes, 256);
        return ptr;
    };
    for (int k = 0; k < 2; ++k) {
        w.candX[k] = (double *)take((int64_t)B * NA * (T + 1) * n * 8);
        w.candU[k] = (double *)take((int64_t)B * NA * T * m * 8);
    }
    w.Jc = (double *)take((int64_t)B * NA * 8);
    w.K = (double *)take((int64_t)B * T * m * n * 8);
    w.d = (double *)take((int64_t)B * T * m * 8);
    w.stage = (double *)take((int64_t)B * (T + 1) * L.stride * 8);
    w.scratch = (double *)take(backward_scratch_doubles(B, a, s, c) * 8);
    w.mu = (double *)take((int64_t)B * 8);
    w.delta = (double *)take((int64_t)B * 8);
    w.Jstar = (double *)take((int64_t)B * 8);
    w.Jlast = (double *)take((int64_t)B * 8);
    w.slot = (int32_t *)take((int64_t)B * 4);
    w.parity = (int32_t *)take((int64_t)B * 4);
    w.active[0] = (int32_t *)take((int64_t)B * 4);
    w.active[1] = (int32_t *)take((int64_t)B * 4);
    w.ls_list[0] = (int32_t *)take((int64_t)B * 4);
    w.ls_list[1] = (int32_t *)take((int64_t)B * 4);
    w.n_active = (int32_t *)take(256);
    w.flag = (int32_t *)take(256);
    w.ls_count = (int32_t *)take(256);
    w.bytes = off;
    return w;
}

// ---- small kernels of the driver loop ------------------------------------------------------------
__global__ void init_state_kernel(int B, double *mu, double *delta, double *Jstar, double *Jlast, const double *J0,
                                  int64_t j_stride, int32_t *slot, int32_t *parity, int32_t *active, int32_t *n_active,
                                  int32_t *iters, int32_t *status, int32_t *trace_alpha, double *trace_mu,
                                  double *trace_J, int n_iter, int NA)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *n_active = B;
    if (b >= B) return;
    mu[b] = 1.0;     // _reset_regularization, control.py:227-230
    delta[b] = 2.0;  // DELTA_0
    Jstar[b] = J0[(int64_t)b * j_stride];
    Jlast[b] = J0[(int64_t)b * j_stride];
    slot[b] = 0;
    parity[b] = 1;
    active[b] = b;
    iters[b] = 0;
    status[b] = 0;
    if (trace_alpha) {
        for (int i = 0; i < n_iter; ++i) {
            trace_alpha[(int64_t)b * n_iter + i] = -2;
            trace_mu[(int64_t)b * n_iter + i] = 0.0;
            for (int k = 0; k < NA; ++k) trace_J[((int64_t)b * n_iter + i) * NA + k] = __longlong_as_double(0x7ff8000000000000ll);
        }
    }
}

// Line-search decision + regularisation schedule + active-list compaction (one CTA).
__global__ void __launch_bounds__(1024) select_kernel(int iter, int n_iter, int NA, double tol, int write_parity,
                                                      const double *Jc, double *mu, double *delta, double *Jstar,
                                                      double *Jlast, int32_t *slot, int32_t *parity,
                                                      const int32_t *active_in, int32_t *active_out, int32_t *n_active,
                                                      int32_t *iters, int32_t *status, int32_t *trace_alpha,
                                                      double *trace_mu, double *trace_J)
{
    __shared__ int warp_counts[32];
    __shared__ int chunk_base;
    const int count = *n_active;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) chunk_base = 0;
    __syncthreads();
    for (int base = 0; base < count; base += blockDim.x) {
        const int idx = base + tid;
        bool keep = false;
        int b = -1;
        if (idx < count) {
            b = active_in[idx];
            const double *J = Jc + (int64_t)b * NA;
            const double Js = Jstar[b];
            if (trace_alpha) {
                trace_mu[(int64_t)b * n_iter + iter] = mu[b];
                for (int k = 0; k < NA; ++k) trace_J[((int64_t)b * n_iter + iter) * NA + k] = J[k];
            }
            int acc = -1;
            for (int k = 0; k < NA; ++k) {
                if (J[k] < Js) { acc = k; break; }  // first improving candidate wins (control.py:179-193)
            }
            int st = 0;
            iters[b] = iter + 1;
            if (acc >= 0) {
                const double Jn = J[acc];
                const bool converged = fabs((Js - Jn) / Js) < tol;
                Jstar[b] = Jn;
                Jlast[b] = Jn;
                slot[b] = acc;
                parity[b] = write_parity;
                // _decrease_regularization (control.py:232-237)
                const double dl = fmin(1.0, delta[b]) / 2.0;
                double muv = mu[b] * dl;
                if (muv <= 1e-6) muv = 0.0;
                delta[b] = dl;
                mu[b] = muv;
                if (converged) st |= DPILQR_ST_CONVERGED;
                else if (iter + 1 >= n_iter) st |= DPILQR_ST_ITER_LIMIT;
                else keep = true;
            } else {
                Jlast[b] = J[NA - 1];  // J of the last candidate tried (control.py:225)
                st |= DPILQR_ST_LS_FAILED;
                bool finite = true;
                for (int k = 0; k < NA; ++k) finite = finite && isfinite(J[k]);
                if (!finite) st |= DPILQR_ST_NONFINITE;
            }
            if (trace_alpha) trace_alpha[(int64_t)b * n_iter + iter] = acc;
            if (st) atomicOr(status + b, st);
        }
        // ordered compaction of the survivors
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_counts[warp] = __popc(ballot);
        __syncthreads();
        int prefix = 0, total = 0;
        const int nwarps = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nwarps; ++w) {
            if (w < warp) prefix += warp_counts[w];
            total += warp_counts[w];
        }
        if (keep) active_out[chunk_base + prefix + __popc(ballot & ((1u << lane) - 1))] = b;
        __syncthreads();
        if (tid == 0) chunk_base += total;
        __syncthreads();
    }
    if (tid == 0) *n_active = chunk_base;
}

// Staged line search (the reference tries the candidates one after the other and stops at the first improvement,
// control.py:179-193; 64 % of the iterations of the metric batch accept the first one): after the candidates
// [k0, k1) have been rolled out for the problems of `list_in`, the problems none of them improved move on to
// `list_out` (ordered compaction); for the others the costs of the candidates never tried are marked NaN.
__global__ void __launch_bounds__(1024) stage_select_kernel(int k0, int k1, int NA, const double *Jstar, double *Jc,
                                                            const int32_t *list_in, const int32_t *count_in,
                                                            int32_t *list_out, int32_t *count_out)
{
    __shared__ int warp_counts[32];
    __shared__ int chunk_base;
    const int count = *count_in;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) chunk_base = 0;
    __syncthreads();
    for (int base = 0; base < count; base += blockDim.x) {
        const int idx = base + tid;
        bool keep = false;
        int b = -1;
        if (idx < count) {
            b = list_in[idx];
            double *J = Jc + (int64_t)b * NA;
            const double Js = Jstar[b];
            bool improved = false;
            for (int k = k0; k < k1; ++k) improved = improved || (J[k] < Js);
            keep = !improved;
            if (improved)
                for (int k = k1; k < NA; ++k) J[k] = __longlong_as_double(0x7ff8000000000000ll);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_counts[warp] = __popc(ballot);
        __syncthreads();
        int prefix = 0, total = 0;
        const int nwarps = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nwarps; ++w) {
            if (w < warp) prefix += warp_counts[w];
            total += warp_counts[w];
        }
        if (keep) list_out[chunk_base + prefix + __popc(ballot & ((1u << lane) - 1))] = b;
        __syncthreads();
        if (tid == 0) chunk_base += total;
        __syncthreads();
    }
    if (tid == 0) *count_out = chunk_base;
}

__global__ void mark_time_limit_kernel(const int32_t *active, const int32_t *n_active, int32_t *status)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < *n_active) atomicOr(status + active[idx], DPILQR_ST_TIME_LIMIT);
}

__global__ void gather_result_kernel(int B, int64_t xlen, int64_t ulen, int NA, const double *candX0,
                                     const double *candX1, const double *candU0, const double *candU1,
                                     const int32_t *slot, const int32_t *parity, const double *Jlast,
                                     const double *Jstar, double *X, double *U, double *J, double *Jstar_out)
{
    const int b = blockIdx.x;
    const double *sx = (parity[b] ? candX1 : candX0) + ((int64_t)b * NA + slot[b]) * xlen;
    const double *su = (parity[b] ? candU1 : candU0) + ((int64_t)b * NA + slot[b]) * ulen;
    for (int64_t k = threadIdx.x; k < xlen; k += blockDim.x) X[(int64_t)b * xlen + k] = sx[k];
    for (int64_t k = threadIdx.x; k < ulen; k += blockDim.x) U[(int64_t)b * ulen + k] = su[k];
    if (threadIdx.x == 0) {
        if (J) J[b] = Jlast[b];
        if (Jstar_out) Jstar_out[b] = Jstar[b];
    }
}

// ---- optional per-kernel timing (CUDA events on the launching stream) ---------------------------
static std::mutex g_profile_lock;
static dpilqr_profile g_profile = {};

static int sm_count()
{
    static int count = 0;
    if (count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || count <= 0) count = 148;
    }
    return count;
}

struct LaunchTimer {
    bool on;
    cudaStream_t stream;
    struct Rec { int kind; int units; cudaEvent_t e0, e1; };
    std::vector<Rec> recs;
    LaunchTimer(bool enable, cudaStream_t s) : on(enable), stream(s) {}
    LaunchTimer(const LaunchTimer &) = delete;
    LaunchTimer &operator=(const LaunchTimer &) = delete;
    ~LaunchTimer()  // an error path that never reached flush(): the events still go back
    {
        for (auto &r : recs) {
            cudaEventDestroy(r.e0);
            cudaEventDestroy(r.e1);
        }
    }
    void begin(int kind, int units)
    {
        if (!on) return;
        Rec r{kind, units, nullptr, nullptr};
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, stream);
        recs.push_back(r);
    }
    void end()
    {
        if (on) cudaEventRecord(recs.back().e1, stream);
    }
    void flush()  // call after the stream has been synchronised
    {
        if (!on) return;
        std::lock_guard<std::mutex> guard(g_profile_lock);
        for (auto &r : recs) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
                g_profile.ms[r.kind] += ms;
                g_profile.launches[r.kind] += 1;
                g_profile.units[r.kind] += r.units;
                if (r.kind == DPILQR_K_BACKWARD && r.units >= sm_count()) {  // launches that fill the machine
                    g_profile.ms[DPILQR_K_BACKWARD_FULL] += ms;
                    g_profile.launches[DPILQR_K_BACKWARD_FULL] += 1;
                    g_profile.units[DPILQR_K_BACKWARD_FULL] += r.units;
                }
            }
            cudaEventDestroy(r.e0);
            cudaEventDestroy(r.e1);
        }
        recs.clear();
    }
};

constexpr int kStagedSearchMinProblems = 600;

// ---- several solves in flight on one device ------------------------------------------------------
// A solve has a BULK (thousands of active problems: most launches fill the machine) and a TAIL (the few problems that
// need many more iterations than the rest: tens of launches of a handful of CTAs, each costing its full latency with
// most SMs idle -- a tenth of the time of the metric batch for 3 % of its work); and even in the bulk the line-search
// launches are bound by the latency of their 51-step chains, not by the machine.  Host threads may therefore call the
// solver concurrently (one stream each): their launches interleave freely, and the tail of a solve moves to a
// high-priority stream so that its few CTAs are scheduled ahead of the other callers' full grids and cost their work
// instead of their latency.  A single caller sees no difference.  (Measured on the metric batch, ms per step: one
// solve at a time 340; two in flight taking turns for the bulk 311, interleaving freely 301; three 297; four 293.)
constexpr int kBulkMinProblems = kStagedSearchMinProblems;

// Per host thread and device: the high-priority stream of the tail, the events that tie it to the caller's stream, and
// the pinned word the active count is read back into (never freed: a few bytes and one stream per calling thread).
struct ThreadContext {
    int device = -1;
    cudaStream_t tail = nullptr, host = nullptr;
    cudaEvent_t ev = nullptr;
    int32_t *h_count = nullptr;
};
static thread_local ThreadContext t_ctx;

static int thread_context(ThreadContext **out)
{
    int device = -1;
    DPILQR_CUDA(cudaGetDevice(&device));
    if (t_ctx.device != device) {
        int lo = 0, hi = 0;
        DPILQR_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi is the numerically lowest = greatest priority
        DPILQR_CUDA(cudaStreamCreateWithPriority(&t_ctx.tail, cudaStreamNonBlocking, hi));
        DPILQR_CUDA(cudaStreamCreateWithFlags(&t_ctx.host, cudaStreamNonBlocking));
        DPILQR_CUDA(cudaEventCreateWithFlags(&t_ctx.ev, cudaEventDisableTiming));
        if (!t_ctx.h_count) DPILQR_CUDA(cudaMallocHost(&t_ctx.h_count, 64));
        t_ctx.device = device;
    }
    *out = &t_ctx;
    return DPILQR_OK;
}

// ---- the driver loop -----------------------------------------------------------------------------
static int64_t solve_device(const dpilqr_batch *batch, const dpilqr_solve_opts *opts, const double *x0,
                            const double *U0, double *X, double *U, double *J, double *J_star, int32_t *iters,
                            int32_t *status, int32_t *trace_alpha, double *trace_mu, double *trace_J, void *workspace,
                            int64_t workspace_bytes, cudaStream_t user_stream)
{
    int rc = validate_batch(batch);
    if (rc) return rc;
    cudaStream_t stream = user_stream;
    if (!opts || !x0 || !U0 || !X || !U || !iters || !status || !workspace) {
        set_error("dpilqr_solve_batch: null argument");
        return DPILQR_E_INVALID;
    }
    const int B = batch->n_problems, a = batch->n_agents, s = batch->s, c = batch->c, T = batch->horizon;
    const int NA = opts->n_alpha, n_iter = opts->n_lqr_iter;
    if (NA < 1 || NA > kMaxAlpha || n_iter < 0) {
        set_error("n_alpha must be 1..10 and n_lqr_iter >= 0");
        return DPILQR_E_INVALID;
    }
    if (opts->record_trace && (!trace_alpha || !trace_mu || !trace_J)) {
        set_error("record_trace set but trace arrays are null");
        return DPILQR_E_INVALID;
    }
    if (B == 0) return 0;
    const int64_t n = (int64_t)a * s, m = (int64_t)a * c;
    Workspace w = carve((char *)workspace, B, a, s, c, T, NA);
    if (w.bytes > workspace_bytes) {
        set_error("workspace too small: need %lld bytes, got %lld", (long long)w.bytes, (long long)workspace_bytes);
        return DPILQR_E_INVALID;
    }
    if (!opts->record_trace) trace_alpha = nullptr;
    const int64_t xlen = (T + 1) * n, ulen = T * m;
    ThreadContext *ctx = nullptr;
    if ((rc = thread_context(&ctx))) return rc;

    // rollout of the warm start into candidate buffer 1, slot 0 (control.py:164)
    ForwardParams fp{};
    fp.batch = *batch;
    fp.X = x0; fp.x_stride = n; fp.U = U0; fp.u_stride = ulen;
    fp.K = nullptr; fp.d = nullptr;
    fp.Xc = w.candX[1]; fp.Uc = w.candU[1]; fp.Jc = w.Jc;
    fp.xc_stride = NA * xlen; fp.uc_stride = NA * ulen; fp.jc_stride = NA;
    fp.n_alpha = 1;
    fp.alpha[0] = 0.0;
    fp.n_list = B;
    fp.uniform_model = batch->model_hint - 1;
    fp.exempt_cand = -1;
    LaunchTimer timer(opts->profile != 0, stream);
    timer.begin(DPILQR_K_ROLLOUT, B);
    rc = launch_forward(fp, B, stream);
    timer.end();
    if (rc) return rc;
    init_state_kernel<<<(B + 255) / 256, 256, 0, stream>>>(B, w.mu, w.delta, w.Jstar, w.Jlast, w.Jc, NA, w.slot, w.parity,
                                                          w.active[0], w.n_active, iters, status, trace_alpha, trace_mu,
                                                          trace_J, n_iter, NA);
    DPILQR_CUDA(cudaGetLastError());

    int32_t *h_count = ctx->h_count;
    *h_count = B;
    int64_t total_iters = 0;
    int n_act = B;
    bool in_tail = false;
    const auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < n_iter && n_act > 0; ++it) {
        if (!in_tail && n_act < kBulkMinProblems) {
            // the tail: move to the high-priority stream
            if ((rc = check_cuda(cudaEventRecord(ctx->ev, stream), "tail event"))) break;
            if ((rc = check_cuda(cudaStreamWaitEvent(ctx->tail, ctx->ev, 0), "tail wait"))) break;
            stream = ctx->tail;
            timer.stream = stream;
            in_tail = true;
        }
        const int cur = (it + 1) & 1, nxt = it & 1;
        const int32_t *act = w.active[it & 1];
        int32_t *act_out = w.active[(it + 1) & 1];

        LinQuadParams lq{};
        lq.batch = *batch;
        lq.X = w.candX[cur]; lq.U = w.candU[cur];
        lq.x_stride = NA * xlen; lq.u_stride = NA * ulen; lq.x_slot_stride = xlen; lq.u_slot_stride = ulen;
        lq.slot = w.slot; lq.active = act; lq.n_active = w.n_active;
        lq.stage = w.stage; lq.status = status;
        lq.a_layout = backward_layout_agents(a, s, c);
        timer.begin(DPILQR_K_LINQUAD, n_act);
        rc = launch_linquad(lq, n_act, stream);
        timer.end();
        if (rc) break;

        BackwardParams bp{};
        bp.batch = *batch;
        bp.stage = w.stage; bp.mu = w.mu; bp.K = w.K; bp.d = w.d; bp.status = status;
        bp.active = act; bp.n_active = w.n_active; bp.scratch = w.scratch;
        bp.a_layout = lq.a_layout;
        timer.begin(DPILQR_K_BACKWARD, n_act);
        rc = launch_backward(bp, n_act, stream);
        timer.end();
        if (rc) break;

        // staged line search: candidate 0 for every active problem, candidates 1..2 for those it did not improve,
        // the rest for those still without an improvement (lists and counts stay on the device; the grids are
        // sized for the host-side bound n_act and surplus CTAs leave at once)
        ForwardParams ls{};
        ls.batch = *batch;
        ls.X = w.candX[cur]; ls.U = w.candU[cur];
        ls.x_stride = NA * xlen; ls.u_stride = NA * ulen; ls.x_slot_stride = xlen; ls.u_slot_stride = ulen;
        ls.slot = w.slot;
        ls.K = w.K; ls.d = w.d;
        ls.Xc = w.candX[nxt]; ls.Uc = w.candU[nxt]; ls.Jc = w.Jc;
        ls.xc_stride = NA * xlen; ls.uc_stride = NA * ulen; ls.jc_stride = NA;
        ls.n_list = n_act;
        ls.uniform_model = batch->model_hint - 1;
        ls.J_bound = opts->bounded_search ? w.Jstar : nullptr;
        ls.exempt_cand = NA - 1;
        // (for a few hundred problems the machine is far from full and a launch costs its latency whatever it
        // carries: all the candidates go in one launch then)
        const bool staged = n_act >= kStagedSearchMinProblems;
        // stage boundaries: candidate 0 | 1..2 | 3..4 | the rest (accepted-index histogram of the metric batch: 56 %, 18 %,
        // 12 %, 14 % incl. failures).  With several solves in flight the latency of one more launch is hidden and the
        // rollouts it saves are not.
        static const bool four_stages = getenv("DPILQR_LS_THREE_STAGES") == nullptr;
        constexpr int kMaxStages = 4;
        int bounds[kMaxStages + 1] = {0, NA, NA, NA, NA};
        if (staged) {
            bounds[1] = NA < 1 ? NA : 1;
            bounds[2] = NA < 3 ? NA : 3;
            bounds[3] = four_stages ? (NA < 5 ? NA : 5) : NA;
        }
        const double expect[kMaxStages] = {1.0, 0.45, 0.27, 0.16};  // share of the active problems that reaches each stage (metric batch)
        const int32_t *list_in = act;
        const int32_t *count_in = w.n_active;
        timer.begin(DPILQR_K_LINESEARCH, n_act);
        for (int stg = 0; stg < kMaxStages && !rc; ++stg) {
            const int k0 = bounds[stg], k1 = bounds[stg + 1];
            if (k1 <= k0) continue;
            ls.active = list_in; ls.n_active = count_in;
            ls.alpha_first = k0; ls.n_alpha = k1 - k0;
            for (int k = k0; k < k1; ++k) ls.alpha[k - k0] = kAlphaTable[k];
            int expected = (int)(expect[stg] * n_act) + 1;
            rc = launch_forward(ls, expected, stream);
            if (rc) break;
            if (k1 < NA) {
                int32_t *list_out = w.ls_list[stg & 1];
                int32_t *count_out = w.ls_count + stg;
                stage_select_kernel<<<1, 1024, 0, stream>>>(k0, k1, NA, w.Jstar, w.Jc, list_in, count_in, list_out, count_out);
                rc = check_cuda(cudaGetLastError(), "stage_select_kernel");
                list_in = list_out;
                count_in = count_out;
            }
        }
        timer.end();
        if (rc) break;

        timer.begin(DPILQR_K_SELECT, n_act);
        select_kernel<<<1, 1024, 0, stream>>>(it, n_iter, NA, opts->tol, nxt, w.Jc, w.mu, w.delta, w.Jstar, w.Jlast,
                                              w.slot, w.parity, act, act_out, w.n_active, iters, status, trace_alpha,
                                              trace_mu, trace_J);
        timer.end();
        if ((rc = check_cuda(cudaGetLastError(), "select_kernel"))) break;
        total_iters += n_act;
        if ((rc = check_cuda(cudaMemcpyAsync(h_count, w.n_active, sizeof(int32_t), cudaMemcpyDeviceToHost, stream), "count D2H"))) break;
        if ((rc = check_cuda(cudaStreamSynchronize(stream), "iteration sync"))) break;
        n_act = *h_count;
        if (opts->t_kill > 0.0 && n_act > 0) {
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (el > opts->t_kill) {  // control.py:213-218, applied to the whole batch
                mark_time_limit_kernel<<<(n_act + 255) / 256, 256, 0, stream>>>(act_out, w.n_active, status);
                break;
            }
        }
    }
    if (rc) {
        cudaStreamSynchronize(stream);
        timer.flush();
        return rc;
    }
    gather_result_kernel<<<B, 128, 0, stream>>>(B, xlen, ulen, NA, w.candX[0], w.candX[1], w.candU[0], w.candU[1], w.slot,
                                               w.parity, w.Jlast, w.Jstar, X, U, J, J_star);
    DPILQR_CUDA(cudaGetLastError());
    if (stream != user_stream) {  // whatever the caller enqueues next on its own stream comes after the tail
        DPILQR_CUDA(cudaEventRecord(ctx->ev, stream));
        DPILQR_CUDA(cudaStreamWaitEvent(user_stream, ctx->ev, 0));
    }
    DPILQR_CUDA(cudaStreamSynchronize(stream));
    timer.flush();
    return total_iters;
}

// ---- cached device memory for the host-buffer entry point --------------------------------------
// A small pool of arenas, one per call in flight (host threads may call concurrently, see above).
struct HostArena {
    int device = -1;
    void *buf = nullptr;
    int64_t bytes = 0;
    bool in_use = false;
};
constexpr int kMaxArenas = 4;
struct HostCache {
    std::mutex lock;
    std::condition_variable cv;
    HostArena arena[kMaxArenas];
};
static HostCache g_cache;

// Borrow an arena of at least `bytes` on `device`: a free one that fits, else a free one re-allocated, else wait.
static int arena_acquire(int device, int64_t bytes, HostArena **out)
{
    std::unique_lock<std::mutex> lk(g_cache.lock);
    for (;;) {
        HostArena *pick = nullptr;
        for (auto &ar : g_cache.arena)
            if (!ar.in_use && ar.device == device && ar.bytes >= bytes && (!pick || ar.bytes < pick->bytes)) pick = &ar;
        if (!pick)
            for (auto &ar : g_cache.arena)
                if (!ar.in_use && !ar.buf) { pick = &ar; break; }
        if (!pick)
            for (auto &ar : g_cache.arena)
                if (!ar.in_use) { pick = &ar; break; }
        if (pick) {
            if (pick->device != device || pick->bytes < bytes) {
                if (pick->buf) { cudaSetDevice(pick->device); cudaFree(pick->buf); cudaSetDevice(device); }
                pick->buf = nullptr; pick->bytes = 0; pick->device = device;
                DPILQR_CUDA(cudaMalloc(&pick->buf, bytes));
                pick->bytes = bytes;
            }
            pick->in_use = true;
            *out = pick;
            return DPILQR_OK;
        }
        g_cache.cv.wait(lk);
    }
}

struct ArenaLease {
    HostArena *arena = nullptr;
    ~ArenaLease()
    {
        if (!arena) return;
        {
            std::lock_guard<std::mutex> lk(g_cache.lock);
            arena->in_use = false;
        }
        g_cache.cv.notify_one();
    }
};

}  // namespace dpilqr

using namespace dpilqr;

extern "C" {

const char *dpilqr_last_error(void) { return g_error; }
int dpilqr_version(void) { return 200; }

int dpilqr_device_count(void)
{
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        set_error("no CUDA device visible (%s); dpilqr_b200 has no CPU fallback", cudaGetErrorString(err));
        cudaGetLastError();
        return DPILQR_E_NO_DEVICE;
    }
    return count;
}

int dpilqr_model_nx(int model) { return model_nx(model); }
int dpilqr_model_nu(int model) { return model_nu(model); }

int64_t dpilqr_stage_stride(int n_agents, int s, int c) { return stage_layout(n_agents, s, c).stride; }

int64_t dpilqr_workspace_bytes(int n_problems, int n_agents, int s, int c, int horizon, int n_alpha)
{
    return carve(nullptr, n_problems, n_agents, s, c, horizon, n_alpha).bytes;
}

int dpilqr_f(int model, int64_t count, const double *x, const double *u, double *xdot, void *stream)
{
    return launch_dynamics(0, model, 0.0, count, x, u, xdot, nullptr, (cudaStream_t)stream);
}

int dpilqr_integrate(int model, double dt, int64_t count, const double *x, const double *u, double *x_new, void *stream)
{
    return launch_dynamics(1, model, dt, count, x, u, x_new, nullptr, (cudaStream_t)stream);
}

int dpilqr_linearize(int model, double dt, int64_t count, const double *x, const double *u, double *A, double *B,
                     void *stream)
{
    return launch_dynamics(2, model, dt, count, x, u, A, B, (cudaStream_t)stream);
}

int dpilqr_rollout_linesearch(const dpilqr_batch *batch, const double *X, const double *U, const double *K,
                              const double *d, const double *alphas, int n_alpha, double *Xc, double *Uc, double *Jc,
                              void *stream)
{
    int rc = validate_batch(batch);
    if (rc) return rc;
    if ((K == nullptr) != (d == nullptr)) { set_error("K and d must both be given or both be null"); return DPILQR_E_INVALID; }
    if (n_alpha < 1 || n_alpha > kMaxAlpha) { set_error("n_alpha must be 1..10"); return DPILQR_E_INVALID; }
    const int64_t n = (int64_t)batch->n_agents * batch->s, m = (int64_t)batch->n_agents * batch->c;
    const int T = batch->horizon;
    ForwardParams fp{};
    fp.batch = *batch;
    fp.X = X; fp.U = U; fp.K = K; fp.d = d;
    fp.x_stride = K ? (T + 1) * n : n;  // plain rollout: X is x0 [B][n]
    fp.u_stride = T * m;
    fp.Xc = Xc; fp.Uc = Uc; fp.Jc = Jc;
    fp.xc_stride = (int64_t)n_alpha * (T + 1) * n; fp.uc_stride = (int64_t)n_alpha * T * m; fp.jc_stride = n_alpha;
    fp.n_list = batch->n_problems;
    fp.uniform_model = batch->model_hint - 1;
    fp.alpha_first = 0;
    fp.exempt_cand = -1;
    fp.n_alpha = n_alpha;
    for (int k = 0; k < n_alpha; ++k) fp.alpha[k] = alphas ? alphas[k] : kAlphaTable[k];
    return launch_forward(fp, batch->n_problems, (cudaStream_t)stream);
}

int dpilqr_linearize_quadraticize(const dpilqr_batch *batch, const double *X, const double *U, double *stage,
                                  int32_t *status, void *stream)
{
    int rc = validate_batch(batch);
    if (rc) return rc;
    const int64_t n = (int64_t)batch->n_agents * batch->s, m = (int64_t)batch->n_agents * batch->c;
    LinQuadParams lq{};
    lq.batch = *batch;
    lq.X = X; lq.U = U; lq.stage = stage; lq.status = status;
    lq.x_stride = (batch->horizon + 1) * n; lq.u_stride = batch->horizon * m;
    return launch_linquad(lq, batch->n_problems, (cudaStream_t)stream);
}

int dpilqr_stage_to_dense(const dpilqr_batch *batch, const double *stage, double *A, double *Bm, double *Lx, double *Lu,
                          double *Lxx, double *Luu, void *stream)
{
    int rc = validate_batch(batch);
    if (rc) return rc;
    return launch_stage_to_dense(*batch, stage, A, Bm, Lx, Lu, Lxx, Luu, (cudaStream_t)stream);
}

int dpilqr_game_cost(const dpilqr_batch *batch, int64_t rows, const double *X, const double *U, int terminal, double *L,
                     void *stream)
{
    int rc = validate_batch(batch);
    if (rc) return rc;
    if (batch->n_agents > 16) { set_error("dpilqr_game_cost supports at most 16 agents"); return DPILQR_E_UNSUPPORTED; }
    return launch_game_cost(*batch, rows, X, U, terminal, L, (cudaStream_t)stream);
}

int dpilqr_backward(const dpilqr_batch *batch, const double *stage, const double *mu, double *K, double *d,
                    int32_t *status, void *stream)
{
    int rc = validate_batch(batch);
    if (rc) return rc;
    BackwardParams bp{};
    bp.batch = *batch;
    bp.stage = stage; bp.mu = mu; bp.K = K; bp.d = d; bp.status = status;
    const int64_t scratch = backward_scratch_doubles(batch->n_problems, batch->n_agents, batch->s, batch->c);
    double *tmp = nullptr;
    if (scratch > 0) {
        DPILQR_CUDA(cudaMallocAsync(&tmp, scratch * 8, (cudaStream_t)stream));
        bp.scratch = tmp;
    }
    rc = launch_backward(bp, batch->n_problems, (cudaStream_t)stream);
    if (tmp) cudaFreeAsync(tmp, (cudaStream_t)stream);
    return rc;
}

int dpilqr_inter_graph(const double *X, int64_t n_scen, int rows, int n_agents, int s, const double *radius,
                       uint64_t *adj, void *stream)
{
    return launch_inter_graph(X, n_scen, rows, n_agents, s, radius, adj, (cudaStream_t)stream);
}

int dpilqr_random_setup(int64_t first_seed, int64_t count, int n_agents, int n_states, int n_d, double var, double energy,
                        double *x0, double *xf, void *stream)
{
    return launch_random_setup(first_seed, count, n_agents, n_states, n_d, var, energy, x0, xf, (cudaStream_t)stream);
}

int64_t dpilqr_solve_batch(const dpilqr_batch *batch, const dpilqr_solve_opts *opts, const double *x0, const double *U0,
                           double *X, double *U, double *J, double *J_star, int32_t *iters, int32_t *status,
                           int32_t *trace_alpha, double *trace_mu, double *trace_J, void *workspace,
                           int64_t workspace_bytes, void *stream)
{
    return solve_device(batch, opts, x0, U0, X, U, J, J_star, iters, status, trace_alpha, trace_mu, trace_J, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}

int64_t dpilqr_solve_batch_host(const dpilqr_batch *hb, const dpilqr_solve_opts *opts, const double *x0,
                                const double *U0, double *X, double *U, double *J, double *J_star, int32_t *iters,
                                int32_t *status, int32_t *trace_alpha, double *trace_mu, double *trace_J, int device)
{
    int rc = validate_batch(hb);
    if (rc) return rc;
    if (!opts) { set_error("null opts"); return DPILQR_E_INVALID; }
    if ((rc = dpilqr_device_count()) < 0) return rc;
    DPILQR_CUDA(cudaSetDevice(device));
    const int B = hb->n_problems, a = hb->n_agents, s = hb->s, c = hb->c, T = hb->horizon, NA = opts->n_alpha;
    const int64_t n = (int64_t)a * s, m = (int64_t)a * c;
    const int n_iter = opts->n_lqr_iter;
    const bool tr = opts->record_trace != 0;
    // one device arena: descriptor arrays + inputs + outputs + solver workspace
    int64_t off = 0;
    auto reserve = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes, 256); return o; };
    const int64_t o_model = reserve((int64_t)B * a * 4), o_ndims = reserve((int64_t)B * a * 4), o_cidx = reserve((int64_t)B * a * 4);
    const int64_t o_Q = reserve((int64_t)hb->n_cost * s * s * 8), o_R = reserve((int64_t)hb->n_cost * c * c * 8), o_Qf = reserve((int64_t)hb->n_cost * s * s * 8);
    const int64_t o_xf = reserve(B * n * 8), o_rad = reserve((int64_t)B * 8), o_w = reserve((int64_t)B * 16), o_hp = reserve((int64_t)B * 4);
    const int64_t o_x0 = reserve(B * n * 8), o_U0 = reserve(B * T * m * 8);
    const int64_t o_X = reserve(B * (T + 1) * n * 8), o_U = reserve(B * T * m * 8), o_J = reserve((int64_t)B * 8), o_Js = reserve((int64_t)B * 8);
    const int64_t o_it = reserve((int64_t)B * 4), o_st = reserve((int64_t)B * 4);
    const int64_t o_ta = reserve(tr ? (int64_t)B * n_iter * 4 : 0), o_tm = reserve(tr ? (int64_t)B * n_iter * 8 : 0);
    const int64_t o_tj = reserve(tr ? (int64_t)B * n_iter * NA * 8 : 0);
    const int64_t ws_bytes = dpilqr_workspace_bytes(B, a, s, c, T, NA);
    const int64_t o_ws = reserve(ws_bytes);

    ArenaLease lease;
    if ((rc = arena_acquire(device, off, &lease.arena))) return rc;
    char *base = (char *)lease.arena->buf;
    ThreadContext *ctx = nullptr;
    if ((rc = thread_context(&ctx))) return rc;
    cudaStream_t stream = ctx->host;  // one stream per calling thread: concurrent callers overlap
    auto h2d = [&](int64_t o, const void *src, int64_t bytes) -> int {
        if (!src || bytes == 0) return 0;
        return check_cuda(cudaMemcpyAsync(base + o, src, bytes, cudaMemcpyHostToDevice, stream), "H2D");
    };
    if ((rc = h2d(o_model, hb->model, (int64_t)B * a * 4))) return rc;
    if ((rc = h2d(o_ndims, hb->n_dims, (int64_t)B * a * 4))) return rc;
    if ((rc = h2d(o_cidx, hb->cost_idx, (int64_t)B * a * 4))) return rc;
    if ((rc = h2d(o_Q, hb->Q, (int64_t)hb->n_cost * s * s * 8))) return rc;
    if ((rc = h2d(o_R, hb->R, (int64_t)hb->n_cost * c * c * 8))) return rc;
    if ((rc = h2d(o_Qf, hb->Qf, (int64_t)hb->n_cost * s * s * 8))) return rc;
    if ((rc = h2d(o_xf, hb->xf, B * n * 8))) return rc;
    if ((rc = h2d(o_rad, hb->radius, (int64_t)B * 8))) return rc;
    if ((rc = h2d(o_w, hb->weights, (int64_t)B * 16))) return rc;
    if ((rc = h2d(o_hp, hb->has_prox, (int64_t)B * 4))) return rc;
    if ((rc = h2d(o_x0, x0, B * n * 8))) return rc;
    if ((rc = h2d(o_U0, U0, B * T * m * 8))) return rc;
    dpilqr_batch db = *hb;
    db.model = (const int32_t *)(base + o_model);
    db.n_dims = (const int32_t *)(base + o_ndims);
    db.cost_idx = (const int32_t *)(base + o_cidx);
    db.Q = (const double *)(base + o_Q);
    db.R = (const double *)(base + o_R);
    db.Qf = (const double *)(base + o_Qf);
    db.xf = (const double *)(base + o_xf);
    db.radius = hb->radius ? (const double *)(base + o_rad) : nullptr;
    db.weights = hb->weights ? (const double *)(base + o_w) : nullptr;
    db.has_prox = hb->has_prox ? (const int32_t *)(base + o_hp) : nullptr;
    const int64_t total = solve_device(&db, opts, (const double *)(base + o_x0), (const double *)(base + o_U0),
                                       (double *)(base + o_X), (double *)(base + o_U), (double *)(base + o_J),
                                       (double *)(base + o_Js), (int32_t *)(base + o_it), (int32_t *)(base + o_st),
                                       tr ? (int32_t *)(base + o_ta) : nullptr, tr ? (double *)(base + o_tm) : nullptr,
                                       tr ? (double *)(base + o_tj) : nullptr, base + o_ws, ws_bytes, stream);
    if (total < 0) return total;
    auto d2h = [&](void *dst, int64_t o, int64_t bytes) -> int {
        if (!dst || bytes == 0) return 0;
        return check_cuda(cudaMemcpyAsync(dst, base + o, bytes, cudaMemcpyDeviceToHost, stream), "D2H");
    };
    if ((rc = d2h(X, o_X, B * (T + 1) * n * 8))) return rc;
    if ((rc = d2h(U, o_U, B * T * m * 8))) return rc;
    if ((rc = d2h(J, o_J, (int64_t)B * 8))) return rc;
    if ((rc = d2h(J_star, o_Js, (int64_t)B * 8))) return rc;
    if ((rc = d2h(iters, o_it, (int64_t)B * 4))) return rc;
    if ((rc = d2h(status, o_st, (int64_t)B * 4))) return rc;
    if (tr) {
        if ((rc = d2h(trace_alpha, o_ta, (int64_t)B * n_iter * 4))) return rc;
        if ((rc = d2h(trace_mu, o_tm, (int64_t)B * n_iter * 8))) return rc;
        if ((rc = d2h(trace_J, o_tj, (int64_t)B * n_iter * NA * 8))) return rc;
    }
    DPILQR_CUDA(cudaStreamSynchronize(stream));
    return total;
}

int dpilqr_get_profile(dpilqr_profile *out, int reset)
{
    std::lock_guard<std::mutex> guard(g_profile_lock);
    if (out) *out = g_profile;
    if (reset) g_profile = dpilqr_profile{};
    return 0;
}

int dpilqr_debug_backward_timing(long long *device_counters)
{
    g_backward_timing = device_counters;
    const char *mode = getenv("DPILQR_DEBUG_BACKWARD_MODE");  // timing experiments only, see kernels.cuh
    g_backward_debug_mode = (device_counters && mode) ? atoi(mode) : 0;
    return 0;
}

int dpilqr_release_cache(void)
{
    std::lock_guard<std::mutex> guard(g_cache.lock);
    for (auto &ar : g_cache.arena) {
        if (ar.in_use || !ar.buf) continue;
        cudaSetDevice(ar.device);
        cudaFree(ar.buf);
        ar.buf = nullptr;
        ar.bytes = 0;
        ar.device = -1;
    }
    return 0;
}

}  // extern "C"
