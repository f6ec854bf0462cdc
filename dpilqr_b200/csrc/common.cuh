// common.cuh -- shared layouts and helpers for the dpilqr_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dpilqr_b200.h"
#include "models.cuh"

namespace dpilqr {

// ------------------------------------------------------------------------------------------
// Stage record: the structured output of the fused linearise+quadraticise kernel for one
// (problem, time step).  Never the dense n x n matrices the reference materialises
// (reference util.py:229-236, cost.py:147-148).  Offsets in doubles:
//   A   [a][s][s]   Euler-discretised per-agent state Jacobians
//   B   [a][s][c]   per-agent control Jacobians
//   Lx  [n]         cost gradient wrt x  (reference weights applied)
//   Lu  [m]         cost gradient wrt u
//   Hd  [a][3][3]   PROX_WEIGHT * sum_j H_ij   (diagonal-block proximity Hessian)
//   Ho  [pairs][3][3]  -PROX_WEIGHT * H_ij     (off-diagonal block for pair i<j)
// The constant reference-cost Hessians (Q+Q^T, R+R^T, Qf+Qf^T) are NOT stored per step.
// ------------------------------------------------------------------------------------------
struct StageLayout {
    int a, s, c, n, m, pairs;
    int offA, offB, offLx, offLu, offHd, offHo, stride;
    int strideA, strideB;  // doubles between the A (B) blocks of consecutive agents
};

__host__ __device__ inline StageLayout stage_layout(int a, int s, int c)
{
    StageLayout L;
    L.a = a; L.s = s; L.c = c;
    L.n = a * s; L.m = a * c;
    L.pairs = a * (a - 1) / 2;
    // The per-agent blocks carry two doubles of padding: the backward kernel keeps a record in shared memory in
    // exactly this layout (one TMA bulk copy per record) and the padding de-aliases the banks of consecutive blocks.
    L.strideA = s * s + 2;
    L.strideB = s * c + 2;
    L.offA = 0;
    L.offB = L.offA + a * L.strideA;
    L.offLx = L.offB + a * L.strideB;
    L.offLu = L.offLx + L.n;
    L.offHd = L.offLu + L.m;
    L.offHo = L.offHd + 9 * a;
    int end = L.offHo + 9 * L.pairs;
    L.stride = (end + 1) & ~1;  // keep every record 16-byte aligned
    return L;
}

// index of the pair (i, j), i < j, in itertools.combinations order (reference util.py:58)
__host__ __device__ inline int pair_index(int i, int j, int a) { return i * a - i * (i + 1) / 2 + (j - i - 1); }

// Device-side copy of the batch descriptor (plain struct, passed by value to kernels).
using Batch = dpilqr_batch;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int check_cuda(cudaError_t err, const char *what);
int validate_batch(const dpilqr_batch *b);

#define DPILQR_CUDA(call)                                      \
    do {                                                       \
        int _rc = ::dpilqr::check_cuda((call), #call);         \
        if (_rc != 0) return _rc;                              \
    } while (0)

// The float32 line-search table of reference control.py:162, as doubles.
extern const double kAlphaTable[10];

}  // namespace dpilqr
