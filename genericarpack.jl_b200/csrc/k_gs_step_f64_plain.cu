// Step-kernel instantiations for one element type (see k_gs_step.cuh); one translation unit per type so that
// the per-NCT instantiations compile in parallel.
#include "k_gs_step.cuh"

namespace ga {
int launch_step_f64_plain(ga_ctx* c, int j) { return launch_step_tp<double, true>(c, j); }
}  // namespace ga
