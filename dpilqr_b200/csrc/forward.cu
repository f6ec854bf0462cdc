// forward.cu -- dispatch of Kernel 1 (rollout + line search, rollout.cuh) to the instantiation of the batch's
// model: a batch whose agents all run one model gets the kernel compiled for that model alone; teams mixing the
// models of one size class (zero-padded heterogeneous teams) get the class kernel that switches per agent.
#include "rollout.cuh"

namespace dpilqr {

int launch_forward(const ForwardParams &p_in, int expected_list, cudaStream_t stream)
{
    ForwardParams p = p_in;
    p.timing = g_backward_timing;
    const Batch &bt = p.batch;
    if (p.n_alpha < 1 || p.n_alpha > kMaxAlpha) {
        set_error("n_alpha must be in 1..%d (got %d)", kMaxAlpha, p.n_alpha);
        return DPILQR_E_INVALID;
    }
    if (p.n_list <= 0) return DPILQR_OK;
    if (((uintptr_t)p.K | (uintptr_t)p.Xc | (uintptr_t)p.X) & 15) {
        set_error("rollout kernel: K, X and the candidate buffers must be 16-byte aligned");
        return DPILQR_E_INVALID;
    }
    if (bt.n_agents > 255) {
        set_error("rollout kernel: at most 255 agents");
        return DPILQR_E_UNSUPPORTED;
    }
    if (expected_list < 1) expected_list = 1;
    if (expected_list > p.n_list) expected_list = p.n_list;
    int mc = -1;
    if (bt.s == 12 && bt.c == 4) mc = kQuad12D;
    else if (bt.s == 3 && bt.c == 2) mc = kCar3D;
    else if (bt.s == 5 && bt.c == 2) mc = kBike5D;
    else if (bt.s == 4 && bt.c == 2) mc = (p.uniform_model == kDoubleInt4D || p.uniform_model == kUnicycle4D) ? p.uniform_model : kMixed4;
    else if (bt.s == 6 && bt.c == 3)
        mc = (p.uniform_model == kDoubleInt6D || p.uniform_model == kQuad6D || p.uniform_model == kHuman6D || p.uniform_model == kHumanLin6D)
                 ? p.uniform_model : kMixed6;
    if (mc < 0) {
        set_error("rollout kernel: unsupported per-agent dimensions (%d, %d)", bt.s, bt.c);
        return DPILQR_E_UNSUPPORTED;
    }
    switch (mc) {
    case kDoubleInt4D: return launch_rollout_class<kDoubleInt4D>(p, expected_list, stream);
    case kDoubleInt6D: return launch_rollout_class<kDoubleInt6D>(p, expected_list, stream);
    case kCar3D: return launch_rollout_class<kCar3D>(p, expected_list, stream);
    case kUnicycle4D: return launch_rollout_class<kUnicycle4D>(p, expected_list, stream);
    case kQuad6D: return launch_rollout_class<kQuad6D>(p, expected_list, stream);
    case kHuman6D: return launch_rollout_class<kHuman6D>(p, expected_list, stream);
    case kHumanLin6D: return launch_rollout_class<kHumanLin6D>(p, expected_list, stream);
    case kQuad12D: return launch_rollout_class<kQuad12D>(p, expected_list, stream);
    case kBike5D: return launch_rollout_class<kBike5D>(p, expected_list, stream);
    case kMixed4: return launch_rollout_class<kMixed4>(p, expected_list, stream);
    case kMixed6: return launch_rollout_class<kMixed6>(p, expected_list, stream);
    default: break;
    }
    return DPILQR_E_UNSUPPORTED;
}

}  // namespace dpilqr
