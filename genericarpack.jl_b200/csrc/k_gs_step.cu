// Dispatcher of the persistent Gram-Schmidt step kernel (k_gs_step.cuh; instantiated per element type in
// k_gs_step_*.cu).
#include <string>

#include "ga_ctx.cuh"

namespace ga {

int launch_step_f64(ga_ctx* c, int j);
int launch_step_f64_plain(ga_ctx* c, int j);
int launch_step_f32(ga_ctx* c, int j);
int launch_step_c64(ga_ctx* c, int j);
int launch_step_dd(ga_ctx* c, int j);

// Tensor map of this context's basis V for tiles of `rows` rows x j columns (the 1 KiB-segment variant of the step
// kernel moves a tile with one cp.async.bulk.tensor instruction).  V is n_pad x ncv column-major: dimension 0 = rows
// (contiguous), dimension 1 = columns (stride n_pad elements).  16-byte element types are described as pairs of doubles.
// The encoder is a driver entry point, fetched at run time (no -lcuda link dependency).
const CUtensorMap* gs_tensor_map(ga_ctx* c, int j, int rows) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
      set_err(c, GA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
      return nullptr;
    }
    encode = (encode_fn)fn;
  }
  if (c->gs_maps.empty()) { c->gs_maps.resize(kMaxNcv + 1); c->gs_maps_ok.assign(kMaxNcv + 1, 0); }
  if (c->gs_maps_ok[(size_t)j]) return &c->gs_maps[(size_t)j];
  const bool f32 = c->dtype == GA_F32;
  const cuuint64_t per = f32 ? 1 : (cuuint64_t)(c->esize / 8);          // map elements per vector element
  const cuuint64_t gdim[2] = {(cuuint64_t)c->n_pad * per, (cuuint64_t)c->ncv};
  const cuuint64_t gstr[1] = {(cuuint64_t)c->n_pad * (cuuint64_t)c->esize};   // bytes between columns
  const cuuint32_t box[2] = {(cuuint32_t)((cuuint64_t)rows * per), (cuuint32_t)j};
  const cuuint32_t est[2] = {1, 1};
  if (box[0] > 256 || box[1] > 256) { set_err(c, GA_ERR_ARG, "gs: tensor-map box too large"); return nullptr; }
  const CUresult r = encode(&c->gs_maps[(size_t)j], f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, c->V,
                            gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_err(c, GA_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r)); return nullptr; }
  c->gs_maps_ok[(size_t)j] = 1;
  return &c->gs_maps[(size_t)j];
}

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
