// cost_api.cu -- GameCost value at arbitrary (x, u) points: the drop-in hook behind
// Cost.__call__ (reference cost.py:79-83, 117-133, 197-206).  One thread per point; it reuses
// the device functions of the rollout kernel so both evaluate the cost identically.
#include "cost.cuh"
#include "kernels.cuh"

namespace dpilqr {

__global__ void __launch_bounds__(128) game_cost_kernel(const Batch bt, int64_t rows, const double *__restrict__ X,
                                                        const double *__restrict__ U, int terminal,
                                                        double *__restrict__ Lout)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (int64_t)bt.n_problems * rows) return;
    const int b = (int)(k / rows);
    const int a = bt.n_agents, s = bt.s, c = bt.c;
    const int n = a * s, m = a * c;
    const double *x = X + k * n;
    const double *u = U ? U + k * m : nullptr;
    const int32_t *ndims_b = bt.n_dims + (int64_t)b * a;
    const bool has_prox = (a > 1) && (bt.has_prox == nullptr || bt.has_prox[b] != 0);
    const double w_ref = bt.weights ? bt.weights[2 * b] : 1.0;
    const double w_prox = bt.weights ? bt.weights[2 * b + 1] : 200.0;
    double ref_total = 0.0;
    for (int i = 0; i < a; ++i) {
        const int ci = bt.cost_idx[(int64_t)b * a + i];
        dispatch_model(bt.model[(int64_t)b * a + i], [&]<int M>() {
            constexpr int NX = model_nx(M), NU = model_nu(M);
            double xs[NX], us[NU];
#pragma unroll
            for (int q = 0; q < NX; ++q) xs[q] = x[i * s + q];
#pragma unroll
            for (int q = 0; q < NU; ++q) us[q] = (terminal || !u) ? 0.0 : u[i * c + q];
            const double *Qm = (terminal ? bt.Qf : bt.Q) + (int64_t)ci * NX * NX;
            ref_total += reference_cost<M>(xs, us, bt.xf + (int64_t)b * n + i * s, Qm, bt.R + (int64_t)ci * NU * NU, terminal != 0);
        });
    }
    double prox = 0.0;
    if (has_prox) {
        bool uniform_dims = true;
        for (int i = 1; i < a; ++i) uniform_dims = uniform_dims && (ndims_b[i] == ndims_b[0]);
        double pc[128];
        const int pairs = a * (a - 1) / 2;
        if (pairs <= 128) {
            int pr = 0;
            for (int i = 0; i < a; ++i)
                for (int j = i + 1; j < a; ++j, ++pr)
                    pc[pr] = pair_penalty(x + i * s, x + j * s, uniform_dims ? 2 : min(ndims_b[i], ndims_b[j]), bt.radius[b]);
            prox = numpy_pairwise_sum(pc, pairs);
        } else {
            prox = __longlong_as_double(0x7ff8000000000000ll);  // > 16 agents: not supported by this hook
        }
    }
    Lout[k] = w_prox * prox + w_ref * ref_total;
}

int launch_game_cost(const Batch &bt, int64_t rows, const double *X, const double *U, int terminal, double *L,
                     cudaStream_t stream)
{
    const int64_t total = (int64_t)bt.n_problems * rows;
    if (total <= 0) return DPILQR_OK;
    game_cost_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(bt, rows, X, U, terminal, L);
    DPILQR_CUDA(cudaGetLastError());
    return DPILQR_OK;
}

}  // namespace dpilqr
