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

// A CTA builds a GROUP of consecutive stage records of one problem (p.records_per_cta of them: as many as 64 kB of
// shared memory and 128 threads hold) in shared memory and streams the group out with one TMA bulk store: records of a
// problem are contiguous in HBM, so the store is a single fully coalesced run of up to 64 kB.  (The first version wrote
// every agent's 146-double Jacobian block straight from its thread: neighbouring threads 1168 bytes apart, every
// store instruction 32 sectors -- 0.21 of the HBM roofline.)
//   1  all threads lay the identity / zero background of the group (16-byte shared-memory stores)
//   2  one thread per (record, agent) computes its slice -- proximity terms, cost gradients, Jacobian non-zeros --
//      on top of it
//   3  fence to the async proxy, one thread issues cp.async.bulk shared -> global and waits for the read side
__global__ void __launch_bounds__(128) linquad_kernel(const LinQuadParams p)
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
    const StageLayout L = stage_layout(a, s, c);
    const int tid = threadIdx.x, nthr = blockDim.x;

    // ---- 1: background of every record of the group: A_i = I, everything else zero
    {
        const int total2 = nrec * L.stride / 2;  // the stride is even
        const int ss = s * s;
        for (int e = tid; e < total2; e += nthr) {
            const int off = (2 * e) % L.stride;
            double v0 = 0.0, v1 = 0.0;
            if (off < L.offB) {
                const int r0 = off % L.strideA, r1 = r0 + 1;  // strideA is even iff s*s is: a pair never straddles two blocks then
                v0 = (r0 < ss && r0 / s == r0 % s) ? 1.0 : 0.0;
                v1 = (r1 < ss && r1 / s == r1 % s) ? 1.0 : 0.0;
                if ((L.strideA & 1) != 0) {  // odd block stride: locate both entries on their own
                    const int q1 = (off + 1) % L.strideA;
                    v1 = (off + 1 < L.offB && q1 < ss && q1 / s == q1 % s) ? 1.0 : 0.0;
                }
            }
            *reinterpret_cast<double2 *>(lq_smem + 2 * e) = make_double2(v0, v1);
        }
    }
    __syncthreads();

    // ---- 2: one thread per (record, agent)
    int st = 0;
    if (tid < nrec * a) {
    const int tl = tid / a, i = tid - tl * a;
    const int t = t0 + tl;
    const bool terminal = (t == T);
    const int slot = p.slot ? p.slot[b] : 0;
    const double *xt = p.X + (int64_t)b * p.x_stride + (int64_t)slot * p.x_slot_stride + (int64_t)t * n;
    const double *ut = terminal ? nullptr
                                : p.U + (int64_t)b * p.u_stride + (int64_t)slot * p.u_slot_stride + (int64_t)t * m;
    double *rec = lq_smem + (size_t)tl * L.stride;

    const int32_t *ndims_b = bt.n_dims + (int64_t)b * a;
    const int model = bt.model[(int64_t)b * a + i];
    const int ci = bt.cost_idx[(int64_t)b * a + i];
    const bool has_prox = (a > 1) && (bt.has_prox == nullptr || bt.has_prox[b] != 0);
    const double w_ref = bt.weights ? bt.weights[2 * b] : 1.0;
    const double w_prox = bt.weights ? bt.weights[2 * b + 1] : 200.0;

    // ---- proximity terms of agent i (position = first coordinates of each agent's state)
    double gsum[3] = {0.0, 0.0, 0.0};
    double hsum[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (has_prox) {
        const double radius = bt.radius[b];
        const int nd_i = ndims_b[i];
        for (int j = 0; j < a; ++j) {
            if (j == i) continue;
            const int lo = i < j ? i : j, hi = i < j ? j : i;
            const int nd = min(nd_i, ndims_b[j]);
            double g[3] = {0.0, 0.0, 0.0}, H[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            bool mismatch;
            const bool inside = pair_quadratic(xt + lo * s, xt + hi * s, nd, radius, g, H, mismatch);
            if (mismatch) st |= DPILQR_ST_POINT_NDIM;
            if (nd < 3) { g[2] = 0.0; H[2] = 0.0; H[4] = 0.0; H[5] = 0.0; }
            const double sign = (i == lo) ? 1.0 : -1.0;
            // the reference adds a (zero) pair contribution even when outside the radius
            if (inside) {
#pragma unroll
                for (int k = 0; k < 3; ++k) gsum[k] += sign * g[k];
#pragma unroll
                for (int k = 0; k < 6; ++k) hsum[k] += H[k];
            }
            if (j > i && inside) {  // off-diagonal block of pair (i, j); zero (the background) outside the radius
                double *Ho = rec + L.offHo + 9 * pair_index(i, j, a);
                const double h[6] = {-w_prox * H[0], -w_prox * H[1], -w_prox * H[2], -w_prox * H[3], -w_prox * H[4], -w_prox * H[5]};
                Ho[0] = h[0]; Ho[1] = h[1]; Ho[2] = h[2];
                Ho[3] = h[1]; Ho[4] = h[3]; Ho[5] = h[4];
                Ho[6] = h[2]; Ho[7] = h[4]; Ho[8] = h[5];
            }
        }
    }
    {
        double *Hd = rec + L.offHd + 9 * i;
        Hd[0] = w_prox * hsum[0]; Hd[1] = w_prox * hsum[1]; Hd[2] = w_prox * hsum[2];
        Hd[3] = w_prox * hsum[1]; Hd[4] = w_prox * hsum[3]; Hd[5] = w_prox * hsum[4];
        Hd[6] = w_prox * hsum[2]; Hd[7] = w_prox * hsum[4]; Hd[8] = w_prox * hsum[5];
    }

    // ---- reference-cost gradients and dynamics Jacobians
    dispatch_model(model, [&]<int M>() {
        constexpr int NX = model_nx(M), NU = model_nu(M);
        double x[NX], u[NU], e[NX];
#pragma unroll
        for (int k = 0; k < NX; ++k) { x[k] = xt[i * s + k]; e[k] = x[k] - bt.xf[(int64_t)b * n + i * s + k]; }
#pragma unroll
        for (int k = 0; k < NU; ++k) u[k] = terminal ? 0.0 : ut[i * c + k];
        const double *Qm = (terminal ? bt.Qf : bt.Q) + (int64_t)ci * NX * NX;
        const double *Rm = bt.R + (int64_t)ci * NU * NU;
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

    // ---- 3: the group leaves in one bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        double *dst = p.stage + ((int64_t)b * (T + 1) + t0) * L.stride;
        const unsigned bytes = (unsigned)(nrec * L.stride * 8);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                     "r"((unsigned)__cvta_generic_to_shared(lq_smem)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory may be released once it has been read
    }
}

int launch_linquad(const LinQuadParams &p_in, int n_problems, cudaStream_t stream)
{
    if (n_problems <= 0) return DPILQR_OK;
    LinQuadParams p = p_in;
    const Batch &bt = p.batch;
    const StageLayout L = stage_layout(bt.n_agents, bt.s, bt.c);
    const int threads = 128;
    const size_t rec_bytes = (size_t)L.stride * 8;
    if (rec_bytes > 200 * 1024 || bt.n_agents > threads) {
        set_error("linearise/quadraticise kernel: a stage record of %d agents does not fit shared memory", bt.n_agents);
        return DPILQR_E_UNSUPPORTED;
    }
    int rpc = (int)((64 * 1024) / rec_bytes);
    if (rpc > threads / bt.n_agents) rpc = threads / bt.n_agents;
    if (rpc > bt.horizon + 1) rpc = bt.horizon + 1;
    if (rpc < 1) rpc = 1;
    // groups of equal size: the last CTA of a problem does not run nearly empty
    const int groups = (bt.horizon + 1 + rpc - 1) / rpc;
    rpc = (bt.horizon + 1 + groups - 1) / groups;
    p.records_per_cta = rpc;
    p.n_blocks_per_problem = groups;
    const size_t smem = rpc * rec_bytes;
    if (smem > 48 * 1024) DPILQR_CUDA(cudaFuncSetAttribute(linquad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
