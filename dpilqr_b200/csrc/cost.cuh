// cost.cuh -- device-side cost evaluation helpers shared by the rollout and the cost hooks.
#pragma once
#include "common.cuh"

namespace dpilqr {

// NumPy's pairwise summation (numpy/core/src/umath/loops_utils.h, used by ndarray.sum), needed to
// reproduce `pair_costs.sum()` of reference cost.py:132-133 in the same order.  The recursion of the
// original is unrolled at compile time (depth 5 covers 4096 addends) so no device stack is needed.
__device__ __forceinline__ double numpy_pairwise_block(const double *a, int n)  // n <= 128
{
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    double r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] += a[i + k];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

// More than 128 addends (teams of 17 agents and up): NumPy splits the range in halves (the first a multiple of 8)
// and recurses.  Out of line and truly recursive -- the compile-time unrolled form inlined dozens of copies of the
// block loop into every kernel that sums pair costs.
static __device__ __noinline__ double numpy_pairwise_recursive(const double *a, int n)
{
    if (n <= 128) return numpy_pairwise_block(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return numpy_pairwise_recursive(a, n2) + numpy_pairwise_recursive(a + n2, n - n2);
}

__device__ __forceinline__ double numpy_pairwise_sum(const double *a, int n)
{
    return n <= 128 ? numpy_pairwise_block(a, n) : numpy_pairwise_recursive(a, n);
}

// (x - xf) Q (x - xf)^T + u R u^T  (terminal: Qf, no control term) -- reference cost.py:79-83
template <int M>
__device__ __forceinline__ double reference_cost(const double (&x)[model_nx(M)], const double (&u)[model_nu(M)],
                                                 const double *__restrict__ xf, const double *__restrict__ Q,
                                                 const double *__restrict__ R, bool terminal)
{
    constexpr int NX = model_nx(M), NU = model_nu(M);
    double e[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) e[i] = x[i] - xf[i];
    double cx = 0.0;
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < NX; ++i) v += e[i] * Q[i * NX + j];
        cx += v * e[j];
    }
    if (terminal) return cx;
    double cu = 0.0;
#pragma unroll
    for (int j = 0; j < NU; ++j) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < NU; ++i) v += u[i] * R[i * NU + j];
        cu += v * u[j];
    }
    return cx + cu;
}

// fmin(0, dist - radius)^2 over the first nd coordinates (reference cost.py:117-133, util.py:48-87);
// the squared norm is accumulated with individually rounded operations like np.linalg.norm.
__device__ __forceinline__ double pair_penalty(const double *xi, const double *xj, int nd, double radius)
{
    const double ddx = xi[0] - xj[0], ddy = xi[1] - xj[1];
    double sq = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));
    if (nd > 2) {
        const double ddz = xi[2] - xj[2];
        sq = __dadd_rn(sq, __dmul_rn(ddz, ddz));
    }
    const double gap = fmin(0.0, sqrt(sq) - radius);
    return gap * gap;
}

}  // namespace dpilqr
