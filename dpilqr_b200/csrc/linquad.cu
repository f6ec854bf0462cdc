// linquad.cu -- Kernel 2: fused linearisation + quadraticisation (sm_100a).
//
// One thread per (problem, time step, agent) emits that agent's slice of the stage record
// (see StageLayout in common.cuh): Euler-discretised Jacobian blocks A_i, B_i, the cost
// gradients L_x, L_u and the 3x3 proximity Hessian blocks.  This replaces, for all T+1 steps
// of all problems at once,
//   MultiDynamicalModel.linearize + uniform_block_diag   reference dynamics.py:173-186, util.py:229-236
//   ReferenceCost.quadraticize                           reference cost.py:85-101
//   ProximityCost.quadraticize / quadraticize_distance   reference cost.py:135-171, 269-315
//   GameCost.quadraticize                                reference cost.py:208-239
// The proximity terms of agent i are accumulated over the other agents in ascending order,
// which is the order in which the reference's pair loop touches agent i's entries.
#include <stdlib.h>

#include "kernels.cuh"

namespace dpilqr {

// Gradient (3) and Hessian (3x3, symmetric) of the thresholded distance penalty for one pair,
// point a = lower agent index, point b = higher (reference cost.py:269-315).
__device__ __forceinline__ bool pair_quadratic(const double *pa, const double *pb, int nd, double radius,
                                               double (&g)[3], double (&H)[6], bool &ndim_mismatch)
{
    const double ax = pa[0], ay = pa[1], az = nd > 2 ? pa[2] : 0.0;
    const double bx = pb[0], by = pb[1], bz = nd > 2 ? pb[2] : 0.0;
    ndim_mismatch = ((az == 0.0) != (bz == 0.0));  // Point.ndim, reference util.py:28-30
    const double dx = ax - bx, dy = ay - by, dz = az - bz;
    const double dist = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    if (dist > radius) return false;
    const double gf = 2.0 * (dist - radius) / dist;
    g[0] = gf * dx; g[1] = gf * dy; g[2] = gf * dz;
    // cross terms use the distance recomputed by the cancelling formula (cost.py:294-303)
    const double ha = __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
    const double hb = __dadd_rn(__dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by)), __dmul_rn(bz, bz));
    const double ab = __dadd_rn(__dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)), __dmul_rn(az, bz));
    const double d2 = sqrt(__dsub_rn(__dadd_rn(ha, hb), __dmul_rn(2.0, ab)));
    const double cross = 2.0 * radius / (d2 * d2 * d2);
    const double dist3 = dist * dist * dist;
    const double tr = 2.0 * radius;
    H[0] = tr * (dx * dx) / dist3 - tr / dist + 2.0;  // xx
    H[3] = tr * (dy * dy) / dist3 - tr / dist + 2.0;  // yy
    H[5] = tr * (dz * dz) / dist3 - tr / dist + 2.0;  // zz
    H[1] = (dx * dy) * cross;                         // xy
    H[2] = (dx * dz) * cross;                         // xz
    H[4] = (dy * dz) * cross;                         // yz
    return true;
}

// A CTA builds a GROUP of consecutive stage records of one problem (p.records_per_cta of them) in shared memory and
// streams the group out with TMA bulk stores: records of a problem are contiguous in HBM, so the group leaves as one
// fully coalesced run of up to 63 kB.  (The first version wrote every agent's 146-double Jacobian block straight from
// its thread: neighbouring threads 1168 bytes apart, every store instruction 32 sectors -- 0.21 of the HBM roofline.)
//   1  all threads lay the zero background of the group (16-byte shared-memory stores)
//   2a the unit diagonals of the A blocks; one thread per (record, agent pair) evaluates the pair's penalty gradient /
//      Hessian once, into a table; the off-diagonal Hessian block of the pair goes straight into the record
//   2b one thread per (record, agent): proximity sums, cost gradients, Jacobian non-zeros on top of the background
//   3  fence to the async proxy, one thread issues cp.async.bulk shared -> global and waits for the read side
constexpr int kLinquadThreads = 128;

struct LinquadSmem {
    int recs, ptab, total_doubles;
};

// (al: agents of the record layout, a: real agents -- see LinQuadParams::a_layout)
__host__ __device__ inline LinquadSmem linquad_smem(int al, int a, int s, int c, int rpc)
{
    const StageLayout L = stage_layout(al, s, c);
    auto even = [](int v) { return (v + 1) & ~1; };
    LinquadSmem M{};
    int off = 0;
    M.recs = off; off += rpc * L.stride;
    M.ptab = off; off += even(rpc * (a * (a - 1) / 2) * 10);  // [rec][pair]: inside flag, g[3], H[6]
    M.total_doubles = off;
    return M;
}

__global__ void __launch_bounds__(kLinquadThreads) linquad_kernel(const LinQuadParams p)
{
    extern __shared__ __align__(16) double lq_smem[];
    const Batch &bt = p.batch;
    const int a = bt.n_agents, s = bt.s, c = bt.c, T = bt.horizon;
    const int n = a * s, m = a * c;
    const int prob_slot = blockIdx.x / p.n_blocks_per_problem;
    if (p.n_active != nullptr && prob_slot >= *p.n_active) return;
    const int b = p.active ? p.active[prob_slot] : prob_slot;
    const int RPC = p.records_per_cta;
    const int t0 = (blockIdx.x % p.n_blocks_per_problem) * RPC;
    const int nrec = min(RPC, T + 1 - t0);
    // The records may be laid out for one agent more than the team has (a phantom agent with A = I, B = 0 and no cost
    // terms -- exactly the background below -- lets odd teams use the tensor-path backward kernel of the even size)
    const int al = p.a_layout > 0 ? p.a_layout : a;
    const StageLayout L = stage_layout(al, s, c);
    const LinquadSmem SM = linquad_smem(al, a, s, c, RPC);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int pairs = a * (a - 1) / 2;
    double *ptab = lq_smem + SM.ptab;
    const int slot = p.slot ? p.slot[b] : 0;
    const double *xbase = p.X + (int64_t)b * p.x_stride + (int64_t)slot * p.x_slot_stride + (int64_t)t0 * n;
    const double *ubase = p.U + (int64_t)b * p.u_stride + (int64_t)slot * p.u_slot_stride + (int64_t)t0 * m;
    const double *xf = bt.xf + (int64_t)b * n;

    // ---- 1: zero background
    {
        const int total2 = nrec * L.stride / 2;  // the stride is even
        const double2 zero2 = make_double2(0.0, 0.0);
        for (int e = tid; e < total2; e += nthr) reinterpret_cast<double2 *>(lq_smem)[e] = zero2;
    }
    __syncthreads();

    // ---- 2a: A_i = I; one thread per (record, agent pair)
    for (int k = tid; k < nrec * L.n; k += nthr) {  // (L.n: the phantom agent's block too)
        const int tl = k / L.n, row = k - tl * L.n;
        const int i = row / s, r = row - i * s;
        lq_smem[(size_t)tl * L.stride + L.offA + i * L.strideA + r * s + r] = 1.0;
    }
    int st = 0;
    const bool has_prox = (a > 1) && (bt.has_prox == nullptr || bt.has_prox[b] != 0);
    const double w_prox = bt.weights ? bt.weights[2 * b + 1] : 200.0;
    const double w_ref = bt.weights ? bt.weights[2 * b] : 1.0;
    if (has_prox) {
        const double radius = bt.radius[b];
        const int32_t *ndims_b = bt.n_dims + (int64_t)b * a;
        for (int item = tid; item < nrec * pairs; item += nthr) {
            const int tl = item / pairs, pr = item - tl * pairs;
            int i = 0, rem = pr;
            while (rem >= a - 1 - i) { rem -= a - 1 - i; ++i; }
            const int j = i + 1 + rem;
            const double *xt = xbase + (size_t)tl * n;
            const int nd = min(ndims_b[i], ndims_b[j]);
            double g[3] = {0.0, 0.0, 0.0}, H[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            bool mismatch;
            const bool inside = pair_quadratic(xt + i * s, xt + j * s, nd, radius, g, H, mismatch);
            if (mismatch) st |= DPILQR_ST_POINT_NDIM;
            if (nd < 3) { g[2] = 0.0; H[2] = 0.0; H[4] = 0.0; H[5] = 0.0; }
            double *row = ptab + (size_t)item * 10;
            row[0] = inside ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) row[1 + k] = g[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) row[4 + k] = H[k];
            if (inside) {  // off-diagonal block of pair (i, j); zero (the background) outside the radius
                double *Ho = lq_smem + (size_t)tl * L.stride + L.offHo + 9 * pair_index(i, j, al);
                const double h[6] = {-w_prox * H[0], -w_prox * H[1], -w_prox * H[2], -w_prox * H[3], -w_prox * H[4], -w_prox * H[5]};
                Ho[0] = h[0]; Ho[1] = h[1]; Ho[2] = h[2];
                Ho[3] = h[1]; Ho[4] = h[3]; Ho[5] = h[4];
                Ho[6] = h[2]; Ho[7] = h[4]; Ho[8] = h[5];
            }
        }
    }
    __syncthreads();

    // ---- 2b: one thread per (record, agent): proximity sums, cost gradients, Jacobian non-zeros
    if (tid < nrec * a) {
        const int tl = tid / a, i = tid - tl * a;
        const bool terminal = (t0 + tl == T);
        double *rec = lq_smem + (size_t)tl * L.stride;
        double gsum[3] = {0.0, 0.0, 0.0};
        double hsum[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (has_prox) {  // the pairs of agent i, the other agent ascending (the order of the reference's pair loop)
            for (int j = 0; j < a; ++j) {
                if (j == i) continue;
                const int lo = i < j ? i : j, hi = i < j ? j : i;
                const double *row = ptab + ((size_t)tl * pairs + pair_index(lo, hi, a)) * 10;
                const double sign = (i == lo) ? 1.0 : -1.0;
                if (row[0] != 0.0) {  // the reference adds a (zero) pair contribution even when outside the radius
#pragma unroll
                    for (int k = 0; k < 3; ++k) gsum[k] += sign * row[1 + k];
#pragma unroll
                    for (int k = 0; k < 6; ++k) hsum[k] += row[4 + k];
                }
            }
        }
        double *Hd = rec + L.offHd + 9 * i;
        Hd[0] = w_prox * hsum[0]; Hd[1] = w_prox * hsum[1]; Hd[2] = w_prox * hsum[2];
        Hd[3] = w_prox * hsum[1]; Hd[4] = w_prox * hsum[3]; Hd[5] = w_prox * hsum[4];
        Hd[6] = w_prox * hsum[2]; Hd[7] = w_prox * hsum[4]; Hd[8] = w_prox * hsum[5];
        const double *xt = xbase + (size_t)tl * n + i * s;
        const double *et = xf + i * s;
        const double *ut = ubase + (size_t)tl * m + i * c;  // (not read for the terminal record)
        const int64_t ci = bt.cost_idx[(int64_t)b * a + i];
        dispatch_model(bt.model[(int64_t)b * a + i], [&]<int M>() {
            constexpr int NX = model_nx(M), NU = model_nu(M);
            double x[NX], u[NU], e[NX];
#pragma unroll
            for (int k = 0; k < NX; ++k) { x[k] = xt[k]; e[k] = x[k] - et[k]; }
#pragma unroll
            for (int k = 0; k < NU; ++k) u[k] = terminal ? 0.0 : ut[k];
            const double *Qm = (terminal ? bt.Qf : bt.Q) + ci * NX * NX;
            const double *Rm = bt.R + ci * NU * NU;
            double *Lx = rec + L.offLx + i * s;
            double *Lu = rec + L.offLu + i * c;
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < NX; ++k) v += e[k] * (Qm[k * NX + j] + Qm[j * NX + k]);
                double lx = w_ref * v;
                if (has_prox && j < 3) lx += w_prox * gsum[j];
                if (!isfinite(lx)) st |= DPILQR_ST_NONFINITE;
                Lx[j] = lx;
            }
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                double v = 0.0;
                if (!terminal) {
#pragma unroll
                    for (int k = 0; k < NU; ++k) v += u[k] * (Rm[k * NU + j] + Rm[j * NU + k]);
                }
                Lu[j] = w_ref * v;
            }
            if (!terminal) {
                EulerDenseSink sink{rec + L.offA + i * L.strideA, rec + L.offB + i * L.strideB, NX, NU, bt.dt};
                model_jacobian<M>(x, u, sink);
            }
        });
    }
    if (st != 0 && p.status) atomicOr(p.status + b, st);

    // ---- 3: the group leaves in bulk stores
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        double *dst = p.stage + ((int64_t)b * (T + 1) + t0) * L.stride;
        const unsigned bytes = (unsigned)(nrec * L.stride * 8);
        // several medium-sized copies move faster than one large one: they are processed concurrently
        constexpr unsigned kChunk = 4096;
        const unsigned src = (unsigned)__cvta_generic_to_shared(lq_smem);
#pragma unroll 1
        for (unsigned off = 0; off < bytes; off += kChunk)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<char *>(dst) + off),
                         "r"(src + off), "r"(min(kChunk, bytes - off)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory may be released once it has been read
    }
}

int launch_linquad(const LinQuadParams &p_in, int n_problems, cudaStream_t stream)
{
    if (n_problems <= 0) return DPILQR_OK;
    LinQuadParams p = p_in;
    const Batch &bt = p.batch;
    const int a = bt.n_agents, s = bt.s, c = bt.c;
    static const int env_threads = getenv("DPILQR_LQ_THREADS") ? atoi(getenv("DPILQR_LQ_THREADS")) : 0;
    static const int env_kb = getenv("DPILQR_LQ_KB") ? atoi(getenv("DPILQR_LQ_KB")) : 0;
    const int threads = env_threads ? env_threads : kLinquadThreads;
    const int al = p.a_layout > 0 ? p.a_layout : a;
    const size_t one = (size_t)linquad_smem(al, a, s, c, 1).total_doubles * 8;
    if (one > 200 * 1024 || a > threads) {
        set_error("linearise/quadraticise kernel: a stage record of %d agents does not fit shared memory", a);
        return DPILQR_E_UNSUPPORTED;
    }
    // records per CTA: three CTAs of up to 74 kB per SM (measured: 1.56 ms for 4096 ten-drone problems against 1.9 to
    // 2.7 ms with one, two or four records per CTA, with 64 or 256 threads, or with the inputs staged in shared memory)
    int rpc = (int)(((env_kb ? env_kb : 74) * 1024) / one);
    if (rpc > threads / a) rpc = threads / a;
    if (rpc > bt.horizon + 1) rpc = bt.horizon + 1;
    if (rpc < 1) rpc = 1;
    // groups of equal size: the last CTA of a problem does not run nearly empty
    const int groups = (bt.horizon + 1 + rpc - 1) / rpc;
    rpc = (bt.horizon + 1 + groups - 1) / groups;
    p.records_per_cta = rpc;
    p.n_blocks_per_problem = groups;
    const size_t smem = (size_t)linquad_smem(al, a, s, c, rpc).total_doubles * 8;
    if (smem > 48 * 1024) DPILQR_CUDA(cudaFuncSetAttribute(linquad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))  /* the ceiling, not this launch's need: concurrent callers must not lower it under each other */;
    linquad_kernel<<<n_problems * p.n_blocks_per_problem, threads, smem, stream>>>(p);
    DPILQR_CUDA(cudaGetLastError());
    return DPILQR_OK;
}

// ------------------------------------------------------------------------------------------
// Dense views of the stage records, for the drop-in hooks (cost.quadraticize / dynamics.linearize)
// and the parity tests.  Not on the solve path.
// ------------------------------------------------------------------------------------------
__global__ void stage_to_dense_kernel(const Batch bt, const double *stage, double *A, double *Bm, double *Lx,
                                      double *Lu, double *Lxx, double *Luu)
{
    const int a = bt.n_agents, s = bt.s, c = bt.c, T = bt.horizon;
    const int n = a * s, m = a * c;
    const StageLayout L = stage_layout(a, s, c);
    const int64_t rec_id = blockIdx.x;  // b * (T+1) + t
    const int b = (int)(rec_id / (T + 1)), t = (int)(rec_id % (T + 1));
    const bool terminal = (t == T);
    const double *rec = stage + rec_id * L.stride;
    const double w_ref = bt.weights ? bt.weights[2 * b] : 1.0;
    for (int k = threadIdx.x; k < n * n; k += blockDim.x) {
        const int r = k / n, col = k % n;
        const int i = r / s, ri = r % s, j = col / s, cj = col % s;
        if (A) A[rec_id * n * n + k] = (i == j && !terminal) ? rec[L.offA + i * L.strideA + ri * s + cj] : 0.0;
        if (Lxx) {
            double v = 0.0;
            if (i == j) {
                const int ci = bt.cost_idx[(int64_t)b * a + i];
                const double *Qm = (terminal ? bt.Qf : bt.Q) + (int64_t)ci * s * s;
                v = w_ref * (Qm[ri * s + cj] + Qm[cj * s + ri]);
                if (ri < 3 && cj < 3) v += rec[L.offHd + 9 * i + ri * 3 + cj];
            } else if (ri < 3 && cj < 3) {
                const int lo = i < j ? i : j, hi = i < j ? j : i;
                const double *Ho = rec + L.offHo + 9 * pair_index(lo, hi, a);
                v = (i < j) ? Ho[ri * 3 + cj] : Ho[cj * 3 + ri];
            }
            Lxx[rec_id * n * n + k] = v;
        }
    }
    for (int k = threadIdx.x; k < n * m; k += blockDim.x) {
        const int r = k / m, col = k % m;
        const int i = r / s, ri = r % s, j = col / c, cj = col % c;
        if (Bm) Bm[rec_id * n * m + k] = (i == j && !terminal) ? rec[L.offB + i * L.strideB + ri * c + cj] : 0.0;
    }
    for (int k = threadIdx.x; k < m * m; k += blockDim.x) {
        const int r = k / m, col = k % m;
        const int i = r / c, ri = r % c, j = col / c, cj = col % c;
        if (Luu) {
            double v = 0.0;
            if (i == j && !terminal) {
                const int ci = bt.cost_idx[(int64_t)b * a + i];
                const double *Rm = bt.R + (int64_t)ci * c * c;
                v = w_ref * (Rm[ri * c + cj] + Rm[cj * c + ri]);
            }
            Luu[rec_id * m * m + k] = v;
        }
    }
    for (int k = threadIdx.x; k < n; k += blockDim.x)
        if (Lx) Lx[rec_id * n + k] = rec[L.offLx + k];
    for (int k = threadIdx.x; k < m; k += blockDim.x)
        if (Lu) Lu[rec_id * m + k] = rec[L.offLu + k];
}

int launch_stage_to_dense(const Batch &bt, const double *stage, double *A, double *Bm, double *Lx, double *Lu,
                          double *Lxx, double *Luu, cudaStream_t stream)
{
    const int64_t recs = (int64_t)bt.n_problems * (bt.horizon + 1);
    if (recs <= 0) return DPILQR_OK;
    stage_to_dense_kernel<<<(unsigned)recs, 128, 0, stream>>>(bt, stage, A, Bm, Lx, Lu, Lxx, Luu);
    DPILQR_CUDA(cudaGetLastError());
    return DPILQR_OK;
}

}  // namespace dpilqr
