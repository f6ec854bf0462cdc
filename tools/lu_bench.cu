// lu_bench.cu -- the LU routine of the backward kernel in isolation, with per-segment cycle counters.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int kSolveThreads = 256;
__device__ long long g_dbg[4];
__device__ int g_probe;
__host__ __device__ constexpr int backward_ldw(int m) { return ((m + 2) & ~1) % 16 == 0 ? ((m + 2) & ~1) + 2 : ((m + 2) & ~1); }
__device__ __forceinline__ void named_barrier(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// Reciprocal without the special-case branch of __drcp_rn: hardware seed (about 20 bits) plus two Newton steps.
__device__ __forceinline__ double fast_rcp(double v)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(v));
    double e = fma(-v, x, 1.0);
    x = fma(x, e, x);
    e = fma(-v, x, 1.0);
    return fma(x, e, x);
}

// Phase C of the backward kernel: LU factorisation of the m x m matrix W (row-major, W[r*ldw + c]) with partial
// pivoting, by the kSolveThreads threads of warp group 1.  Four threads per row, each owning every fourth pair of
// columns (double2 accesses), everything in shared memory, ONE named barrier per column.  While eliminating
// column k the thread that owns the entry of column k+1 publishes its pivot-search key together with its
// reciprocal (computed speculatively, off the critical path).  After the barrier each warp finds the arg-max on
// its own with warp reductions and picks the matching reciprocal up (LAPACK dgetf2 also scales by the reciprocal
// pivot).  Rows never move: a used pivot row is simply marked (key 0) and its index recorded in order[k].  On
// return W holds the multipliers l(r, k) in the eliminated positions and the rows of U in the pivot rows.
// The body of the column loop is branch-free straight-line code (a lone warp per scheduler pays the full branch
// latency), the loop itself is not unrolled (instruction-cache footprint); kept out of line for a register
// allocation of its own.  MT > 0 fixes m at compile time.
template <int MT>
__device__ __noinline__ void lu_implicit_pivoting(double *__restrict__ W, unsigned long long *__restrict__ keybuf,
                                                 double *__restrict__ rinvbuf, int *__restrict__ order, int m_rt, int gt)
{
    const int m = MT > 0 ? MT : m_rt;
    const int ldw = backward_ldw(m);
    const int npair = (m + 1) >> 1;
    constexpr int NP = MT > 0 ? ((MT + 1) / 2 + 3) / 4 : 8;  // column pairs per thread
    const int lane = gt & 31;
    const int r = gt >> 2, q = gt & 3;
    const bool myrow = r < m;
    double *wrow = W + (myrow ? r : m - 1) * ldw;
    // |v| of a double orders like its bit pattern; +1 so that a live zero still beats a used row (key 0)
    auto pivot_key = [](double v) -> unsigned long long {
        const double av = fabs(v);
        return (av == av) ? (unsigned long long)__double_as_longlong(av) + 1ull : 1ull;
    };
    if (gt < 128) keybuf[gt] = 0ull;  // rows >= m never compete
    named_barrier(1, kSolveThreads);
    if (q == 0 && myrow) {
        const double v = wrow[0];
        keybuf[r] = pivot_key(v);
        rinvbuf[r] = fast_rcp(v);
    }
    named_barrier(1, kSolveThreads);
    bool mydone = !myrow;
    double2 wreg[NP];  // this thread's column pairs of row r
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int j = q + 4 * i;
        wreg[i] = (j < npair) ? *reinterpret_cast<const double2 *>(wrow + 2 * j) : make_double2(0.0, 0.0);
    }
    double held_mult = 0.0;  // multiplier of the previous step, stored one barrier later
    bool held = false;
#pragma unroll 1
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, tm = clock64();
    for (int k = 0; k < m; ++k) {
        const unsigned long long *cur = keybuf + (k & 1) * 64;
        const unsigned long long key0 = cur[lane];
        const unsigned long long key1 = cur[lane + 32];
        const unsigned long long kmax = key1 > key0 ? key1 : key0;
        const int rsel = key1 > key0 ? lane + 32 : lane;
        const unsigned hi = (unsigned)(kmax >> 32);
        const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
        bool mine = (hi == mhi);
        unsigned bal = __ballot_sync(0xffffffffu, mine);
        if (__popc(bal) > 1) {  // rare: several rows share the top 32 bits
            const unsigned lo = mine ? (unsigned)kmax : 0u;
            const unsigned mlo = __reduce_max_sync(0xffffffffu, lo);
            mine = mine && (lo == mlo);
            bal = __ballot_sync(0xffffffffu, mine);
        }
        const int pr = __shfl_sync(0xffffffffu, rsel, __ffs(bal) - 1);
        { long long now = clock64(); c0 += now - tm; tm = now; }
        const double rinv = rinvbuf[(k & 1) * 64 + pr];
        const double *prow = W + pr * ldw;
        if (gt == 0) order[k] = pr;
        // The multiplier of step k-1 replaces the eliminated entry (r, k-1) only now: every thread of the row has
        // read that entry before the barrier that ended step k-1.
        if (held) wrow[k - 1] = held_mult;
        mydone = mydone || (r == pr);
        const bool live = !mydone;
        const double mult = wrow[k] * rinv;
        held = live && (q == 0);
        held_mult = mult;
        { long long now = clock64(); c1 += now - tm; tm = now; }
        __syncwarp();  // all four threads of the row have read entry (r, k): the pair loop below may overwrite it
        // Column k+1 first, by all four threads of the row alike (no divergence): the next pivot search needs its
        // key and the speculative reciprocal as early as possible.  The pair loop recomputes the same value.
        if (k + 1 < m) {
            const double v = live ? fma(-mult, prow[k + 1], wrow[k + 1]) : 0.0;
            const unsigned long long key = live ? pivot_key(v) : 0ull;
            const double vr = fast_rcp(v);
            if (myrow && q == 1) {
                keybuf[((k + 1) & 1) * 64 + r] = key;
                rinvbuf[((k + 1) & 1) * 64 + r] = vr;
            }
        }
        // Pair loop.  The thread's own pairs live in registers (static indexing) for the whole factorisation and
        // are mirrored to shared memory after every update; only the pivot row is loaded, and only the pairs that
        // still change (j >= jp0), so the shared-memory traffic shrinks with the active sub-matrix.  The pair
        // holding column k+1 may also rewrite the eliminated entry (r, k) with rounding noise: the multiplier is
        // stored over it at the next step.
        const int jp0 = (k + 1) >> 1;
        double2 p2[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const int j = q + 4 * i;
            p2[i] = make_double2(0.0, 0.0);
            if (live && j >= jp0 && j < npair) p2[i] = *reinterpret_cast<const double2 *>(prow + 2 * j);
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            wreg[i].x = fma(-mult, p2[i].x, wreg[i].x);
            wreg[i].y = fma(-mult, p2[i].y, wreg[i].y);
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const int j = q + 4 * i;
            if (live && j >= jp0 && j < npair) *reinterpret_cast<double2 *>(wrow + 2 * j) = wreg[i];
        }
        { long long now = clock64(); c2 += now - tm; tm = now; }
        named_barrier(1, kSolveThreads);
        { long long now = clock64(); c3 += now - tm; tm = now; }
    }
    if (gt == g_probe) { g_dbg[0] = c0; g_dbg[1] = c1; g_dbg[2] = c2; g_dbg[3] = c3; }
    if (held) wrow[m - 1] = held_mult;
    named_barrier(1, kSolveThreads);
}


__global__ void k_lu(const double *A, long long *out, int m, int reps)
{
    extern __shared__ double smem[];
    const int ldw = backward_ldw(m);
    double *W = smem;
    unsigned long long *keybuf = reinterpret_cast<unsigned long long *>(smem + m * ldw);
    double *rinvbuf = smem + m * ldw + 128;
    int *order = reinterpret_cast<int *>(smem + m * ldw + 256);
    const int tid = threadIdx.x;
    long long total = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int k = tid; k < m * m; k += blockDim.x) W[(k / m) * ldw + k % m] = A[k];
        __syncthreads();
        long long t0 = clock64();
        if (tid < kSolveThreads) lu_implicit_pivoting<40>(W, keybuf, rinvbuf, order, m, tid);
        long long t1 = clock64();
        __syncthreads();
        total += t1 - t0;
    }
    if (tid == 0) { out[0] = total; out[1] = order[0] + order[39]; }
}

int main(int argc, char **argv)
{
    const int m = 40, reps = 50;
    const int ldw = backward_ldw(m);
    double *hA = (double *)malloc(m * m * 8);
    srand(1);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) hA[i * m + j] = (double)rand() / RAND_MAX - 0.5 + (i == j ? 3.0 : 0.0);
    double *dA;
    long long *dout, h[2], dbg[4];
    cudaMalloc(&dA, m * m * 8);
    cudaMalloc(&dout, 16);
    cudaMemcpy(dA, hA, m * m * 8, cudaMemcpyHostToDevice);
    for (int probe : {0, 33, 100, 200}) {
        cudaMemcpyToSymbol(g_probe, &probe, 4);
        for (int pass = 0; pass < 2; ++pass) k_lu<<<1, 512, (m * ldw + 512) * 8>>>(dA, dout, m, reps);
        cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
        cudaMemcpyFromSymbol(dbg, g_dbg, 32);
        printf("probe thread %3d: %.0f cycles per LU (%.0f per column); per column: search %.0f, mult %.0f, eliminate %.0f, barrier %.0f  [%s]\n",
               probe, (double)h[0] / reps, (double)h[0] / reps / m, dbg[0] / 40.0, dbg[1] / 40.0, dbg[2] / 40.0, dbg[3] / 40.0,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
