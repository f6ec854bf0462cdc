// scenario.cu -- scenario generation on the device (sm_100a): random_setup(..., random=True) of the reference
// (util.py:165-195: randomize_locs :125-132, normalize_energy :203-217, compute_energy :198-200) for a whole batch of
// seeds at once, BIT-IDENTICAL to the host path seeded with np.random.seed(seed).
//
// One thread per scenario runs NumPy's legacy generator itself: MT19937 seeded by init_genrand(seed) (numpy
// _legacy_seeding -> mt19937_seed), doubles drawn as (a >> 5, b >> 6) pairs of 32-bit outputs (random_double),
// uniform(-1, 1) as low + range * r, in the order the reference consumes them (initial positions agent by agent, then
// the goals).  The reductions follow NumPy's orders: mean(0) adds the agents one after the other, the norm of a row adds
// its squares left to right, ndarray.sum() is the pairwise summation of cost.cuh.  Every operation is individually
// rounded (no FMA contraction), like the host's.  The `random=False` variant (rejection by mutual repulsion,
// util.py:141-147) stays on the host.
#include "cost.cuh"
#include "kernels.cuh"

namespace dpilqr {

constexpr int kMtN = 624, kMtM = 397;

__global__ void __launch_bounds__(64) random_setup_kernel(int64_t first_seed, int64_t count, int a, int s, int n_d,
                                                          double var, double energy, double *__restrict__ x0,
                                                          double *__restrict__ xf)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    unsigned mt[kMtN];
    unsigned v = (unsigned)((uint64_t)(first_seed + idx) & 0xffffffffull);
    for (int i = 0; i < kMtN; ++i) {
        mt[i] = v;
        v = 1812433253u * (v ^ (v >> 30)) + (unsigned)(i + 1);
    }
    // outputs 0 .. 226 of the first generation only read words of the seeding (i + 397 < 624): no in-place twist needed
    auto output = [&](int i) {
        const unsigned y = (mt[i] & 0x80000000u) | (mt[i + 1] & 0x7fffffffu);
        unsigned z = mt[i + kMtM] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        z ^= z >> 11;
        z ^= (z << 7) & 0x9d2c5680u;
        z ^= (z << 15) & 0xefc60000u;
        z ^= z >> 18;
        return z;
    };
    const int per_set = a * n_d;
    const int n = a * s;
    for (int set = 0; set < 2; ++set) {
        double *out = (set == 0 ? x0 : xf) + idx * n;
        double pos[64 * 3];
        for (int k = 0; k < per_set; ++k) {
            const int d = set * per_set + k;
            const double hi = (double)(output(2 * d) >> 5), lo = (double)(output(2 * d + 1) >> 6);
            const double r = __ddiv_rn(__dadd_rn(__dmul_rn(hi, 67108864.0), lo), 9007199254740992.0);
            pos[k] = __dmul_rn(var, __dadd_rn(-1.0, __dmul_rn(2.0, r)));  // var * uniform(-1, 1)
        }
        if (energy > 0.0) {  // normalize_energy: centre, then scale the summed distance from the origin to `energy`
            for (int c = 0; c < n_d; ++c) {
                double acc = pos[c];
                for (int i = 1; i < a; ++i) acc = __dadd_rn(acc, pos[i * n_d + c]);
                const double center = __ddiv_rn(acc, (double)a);
                for (int i = 0; i < a; ++i) pos[i * n_d + c] = __dsub_rn(pos[i * n_d + c], center);
            }
            double nrm[64];
            for (int i = 0; i < a; ++i) {
                double sq = __dmul_rn(pos[i * n_d], pos[i * n_d]);
                for (int c = 1; c < n_d; ++c) sq = __dadd_rn(sq, __dmul_rn(pos[i * n_d + c], pos[i * n_d + c]));
                nrm[i] = __dsqrt_rn(sq);
            }
            const double scale = __ddiv_rn(energy, numpy_pairwise_sum(nrm, a));
            for (int k = 0; k < per_set; ++k) pos[k] = __dmul_rn(pos[k], scale);
        }
        for (int i = 0; i < a; ++i)
            for (int c = 0; c < s; ++c) out[i * s + c] = (c < n_d) ? pos[i * n_d + c] : 0.0;
    }
}

int launch_random_setup(int64_t first_seed, int64_t count, int a, int s, int n_d, double var, double energy, double *x0,
                        double *xf, cudaStream_t stream)
{
    if (count <= 0) return DPILQR_OK;
    if (a < 1 || a > 64 || n_d < 1 || n_d > 3 || s < n_d || 4 * a * n_d > 227) {
        set_error("random_setup: need 1 <= agents <= 64, 1 <= n_d <= 3 <= ... <= s and at most 56 coordinates per set (got a=%d s=%d n_d=%d)", a, s, n_d);
        return DPILQR_E_INVALID;
    }
    if (first_seed < 0 || first_seed + count - 1 > 0xffffffffll) {
        set_error("random_setup: seeds must lie in [0, 2^32) like np.random.seed's");
        return DPILQR_E_INVALID;
    }
    if (!x0 || !xf) { set_error("random_setup: null output"); return DPILQR_E_INVALID; }
    random_setup_kernel<<<(unsigned)((count + 63) / 64), 64, 0, stream>>>(first_seed, count, a, s, n_d, var, energy, x0, xf);
    DPILQR_CUDA(cudaGetLastError());
    return DPILQR_OK;
}

}  // namespace dpilqr
