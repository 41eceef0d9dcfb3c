// extern "C" entry points of the coupling kernels + kernel selection.
#include "bgx_common.cuh"

namespace bgx {
int affine_coupling_simt(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_packed_mlp*, float, int,
                         cudaStream_t);
int spline_coupling_simt(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_spline_cfg*, int,
                         cudaStream_t);
bool spline_tc_eligible(const bgx_packed_mlp*, const bgx_spline_cfg*, int);
void tc_set_trace(unsigned long long*, int);
int spline_coupling_tc(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_spline_cfg*, int, int*,
                       cudaStream_t);
}  // namespace bgx

extern "C" int bgx_affine_coupling(const bgx_coupling_io* io, const bgx_packed_mlp* shift,
                                   const bgx_packed_mlp* scale, float log_alpha, int flags, void* stream) {
  return bgx::affine_coupling_simt(io, shift, scale, log_alpha, flags, (cudaStream_t)stream);
}

extern "C" int bgx_spline_coupling(const bgx_coupling_io* io, const bgx_packed_mlp* params_net,
                                   const bgx_spline_cfg* cfg, int flags, void* stream) {
  if (!io || !params_net || !cfg) return BGX_ERR_INVALID;
  if (!(flags & BGX_FLAG_FORCE_SIMT) && bgx::spline_tc_eligible(params_net, cfg, 0))
    return bgx::spline_coupling_tc(io, params_net, cfg, flags, cfg->status, (cudaStream_t)stream);
  return bgx::spline_coupling_simt(io, params_net, cfg, flags, (cudaStream_t)stream);
}

extern "C" int bgx_debug_set_trace(uint64_t* device_buffer, int capacity) {
  bgx::tc_set_trace((unsigned long long*)device_buffer, capacity);
  return BGX_OK;
}
