// Dispatcher of the persistent Gram-Schmidt step kernel (k_gs_step.cuh; instantiated per element type in
// k_gs_step_*.cu).
#include "ga_ctx.cuh"

namespace ga {

int launch_step_f64(ga_ctx* c, int j);
int launch_step_f64_plain(ga_ctx* c, int j);
int launch_step_f32(ga_ctx* c, int j);
int launch_step_c64(ga_ctx* c, int j);
int launch_step_dd(ga_ctx* c, int j);

// All Gram-Schmidt phases of Lanczos step j (j basis columns) on c->resid, decisions on the device.
int launch_gs_step(ga_ctx* c, int j) {
  if (j < 1 || j > kMaxNcv) return set_err(c, GA_ERR_ARG, "gs: j out of range");
  switch (c->dtype) {
    case GA_F64: return c->dot_plain ? launch_step_f64_plain(c, j) : launch_step_f64(c, j);
    case GA_C64: return launch_step_c64(c, j);
    case GA_F32: return launch_step_f32(c, j);
    default: return launch_step_dd(c, j);
  }
}

}  // namespace ga
