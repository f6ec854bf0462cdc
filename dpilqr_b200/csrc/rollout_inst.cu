// rollout_inst.cu -- one translation unit per model / size class of the rollout kernel (rollout.cuh); compiled
// once per class with -DDPILQR_ROLLOUT_CLASS=<id> so the nine ODE bodies build in parallel and never share a kernel.
#include "rollout.cuh"

#ifndef DPILQR_ROLLOUT_CLASS
#error "compile with -DDPILQR_ROLLOUT_CLASS=<model id | 100 | 101>"
#endif

namespace dpilqr {
DPILQR_DEFINE_ROLLOUT_CLASS(DPILQR_ROLLOUT_CLASS)
}  // namespace dpilqr
