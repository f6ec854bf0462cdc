// dynamics_api.cu -- batched single-agent entry points f / integrate / linearize (sm_100a).
//
// These keep the four names of the reference's native module alive (reference
// bbdynamicswrap.pyx:61-164: f, integrate, linearize, Model) on top of the device model
// library; one thread per (x, u) sample.
#include "kernels.cuh"

namespace dpilqr {

template <int MODE>  // 0: f, 1: integrate, 2: linearize
__global__ void __launch_bounds__(128) dynamics_kernel(int model, double dt, int64_t count, const double *__restrict__ x,
                                                       const double *__restrict__ u, double *__restrict__ out0,
                                                       double *__restrict__ out1)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    dispatch_model(model, [&]<int M>() {
        constexpr int NX = model_nx(M), NU = model_nu(M);
        double xs[NX], us[NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) xs[i] = x[k * NX + i];
#pragma unroll
        for (int i = 0; i < NU; ++i) us[i] = u[k * NU + i];
        if constexpr (MODE == 0) {
            double xd[NX];
            model_f<M>(xs, us, xd);
#pragma unroll
            for (int i = 0; i < NX; ++i) out0[k * NX + i] = xd[i];
        } else if constexpr (MODE == 1) {
            model_step<M>(dt, xs, us);
#pragma unroll
            for (int i = 0; i < NX; ++i) out0[k * NX + i] = xs[i];
        } else {
            double *A = out0 + k * NX * NX;
            double *B = out1 + k * NX * NU;
#pragma unroll
            for (int r = 0; r < NX; ++r) {
#pragma unroll
                for (int cc = 0; cc < NX; ++cc) A[r * NX + cc] = (r == cc) ? 1.0 : 0.0;
#pragma unroll
                for (int cc = 0; cc < NU; ++cc) B[r * NU + cc] = 0.0;
            }
            EulerDenseSink sink{A, B, NX, NU, dt};
            model_jacobian<M>(xs, us, sink);
        }
    });
}

int launch_dynamics(int mode, int model, double dt, int64_t count, const double *x, const double *u, double *out0,
                    double *out1, cudaStream_t stream)
{
    if (model < 0 || model >= kModelCount) {
        set_error("unknown model id %d", model);
        return DPILQR_E_INVALID;
    }
    if (count <= 0) return DPILQR_OK;
    const unsigned blocks = (unsigned)((count + 127) / 128);
    if (mode == 0) dynamics_kernel<0><<<blocks, 128, 0, stream>>>(model, dt, count, x, u, out0, out1);
    else if (mode == 1) dynamics_kernel<1><<<blocks, 128, 0, stream>>>(model, dt, count, x, u, out0, out1);
    else dynamics_kernel<2><<<blocks, 128, 0, stream>>>(model, dt, count, x, u, out0, out1);
    DPILQR_CUDA(cudaGetLastError());
    return DPILQR_OK;
}

}  // namespace dpilqr
