// graph.cu -- Kernel 4: DP-iLQR interaction graph (sm_100a).
//
// Replaces define_inter_graph_threshold (reference distributed.py:224-247) together with the
// planar pairwise distance it is built on (reference util.py:48-61), for many scenarios per
// launch.  One CTA per scenario, one thread per agent pair; neighbourhoods come back as
// bit masks and are bit-exact with the reference: the distance is formed from individually
// rounded operations (no FMA contraction) exactly like np.linalg.norm on two coordinates, and the
// test is the strict `<` against 2*radius on the rows 0, step, 2*step, ... with
// step = max(rows // 10, 1).
#include "kernels.cuh"

namespace dpilqr {

__global__ void __launch_bounds__(128) inter_graph_kernel(const double *__restrict__ X, int rows, int a, int s,
                                                          const double *__restrict__ radius,
                                                          unsigned long long *__restrict__ adj)
{
    __shared__ unsigned long long nb[64];
    const int64_t k = blockIdx.x;
    const int n = a * s;
    const double *Xk = X + k * (int64_t)rows * n;
    const double planning_radius = 2 * radius[k];
    const int step = max(rows / 10, 1);
    if (threadIdx.x < a) nb[threadIdx.x] = 1ull << threadIdx.x;
    __syncthreads();
    const int pairs = a * (a - 1) / 2;
    for (int pr = threadIdx.x; pr < pairs; pr += blockDim.x) {
        int i = 0, rem = pr;
        while (rem >= a - 1 - i) { rem -= a - 1 - i; ++i; }
        const int j = i + 1 + rem;
        bool close = false;
        for (int row = 0; row < rows && !close; row += step) {
            const double *xr = Xk + (int64_t)row * n;
            const double dx = __dsub_rn(xr[i * s], xr[j * s]);
            const double dy = __dsub_rn(xr[i * s + 1], xr[j * s + 1]);
            const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            close = dist < planning_radius;
        }
        if (close) {
            atomicOr(&nb[i], 1ull << j);
            atomicOr(&nb[j], 1ull << i);
        }
    }
    __syncthreads();
    if (threadIdx.x < a) adj[k * a + threadIdx.x] = nb[threadIdx.x];
}

int launch_inter_graph(const double *X, int64_t n_scen, int rows, int a, int s, const double *radius, uint64_t *adj,
                       cudaStream_t stream)
{
    if (a < 2) {
        // reference util.py:55-56 raises ValueError for a single agent
        set_error("Can't compute pairwise distance for one agent.");
        return DPILQR_E_INVALID;
    }
    if (a > 64 || s < 2 || rows < 1) {
        set_error("inter_graph: need 2 <= agents <= 64, s >= 2, rows >= 1 (got a=%d s=%d rows=%d)", a, s, rows);
        return DPILQR_E_INVALID;
    }
    if (n_scen <= 0) return DPILQR_OK;
    inter_graph_kernel<<<(unsigned)n_scen, 128, 0, stream>>>(X, rows, a, s, radius,
                                                             reinterpret_cast<unsigned long long *>(adj));
    DPILQR_CUDA(cudaGetLastError());
    return DPILQR_OK;
}

}  // namespace dpilqr
