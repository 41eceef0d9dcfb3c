// extern "C" entry points of the coupling kernels + kernel selection.
#include "bgx_common.cuh"

namespace bgx {
int affine_coupling_simt(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_packed_mlp*, float, int,
                         cudaStream_t);
int spline_coupling_simt(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_spline_cfg*, int,
                         cudaStream_t);
bool spline_tc_eligible(const bgx_packed_mlp*, const bgx_spline_cfg*, int);
bool spline_tc2_eligible(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_spline_cfg*, int);
int spline_coupling_tc2(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_spline_cfg*, int, int*, cudaStream_t);
int spline_pair_eligible(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_spline_cfg*, int);
int spline_coupling_pair(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_spline_cfg*, int, int, int*, cudaStream_t);
bool affine_pair_eligible(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_packed_mlp*, int);
int affine_coupling_pair(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_packed_mlp*, float, int, int*, cudaStream_t);
void tc_set_trace(unsigned long long*, int);
bool affine_tc_eligible(const bgx_packed_mlp*, const bgx_packed_mlp*, int);
bool affine_tc2_eligible(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_packed_mlp*, int);
int affine_coupling_tc2(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_packed_mlp*, float, int, int*,
                        cudaStream_t);
int affine_coupling_tc(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_packed_mlp*, float, int, int*,
                       cudaStream_t);
int spline_coupling_tc(const bgx_coupling_io*, const bgx_packed_mlp*, const bgx_spline_cfg*, int, int*,
                       cudaStream_t);
}  // namespace bgx

static long long g_kernel_count[BGX_KERNEL_IDS] = {};
extern "C" int64_t bgx_kernel_count(int id) { return (id >= 0 && id < BGX_KERNEL_IDS) ? g_kernel_count[id] : -1; }

static int* g_status = nullptr;   // device flag raised by the tensor-core kernels on a pipeline timeout

extern "C" int bgx_set_status_buffer(int32_t* device_int) {
  g_status = device_int;
  return BGX_OK;
}

extern "C" int bgx_affine_coupling(const bgx_coupling_io* io, const bgx_packed_mlp* shift,
                                   const bgx_packed_mlp* scale, float log_alpha, int flags, void* stream) {
  if (!io) return BGX_ERR_INVALID;
  if (!(flags & (BGX_FLAG_FORCE_SIMT | BGX_FLAG_PREFER_PAIR)) && bgx::affine_tc2_eligible(io, shift, scale, flags)) {
    ++g_kernel_count[BGX_KERNEL_AFFINE_TC2];
    return bgx::affine_coupling_tc2(io, shift, scale, log_alpha, flags, g_status, (cudaStream_t)stream);
  }
  if (!(flags & BGX_FLAG_FORCE_SIMT) && bgx::affine_pair_eligible(io, shift, scale, flags) &&
      (!bgx::affine_tc_eligible(shift, scale, flags) || (flags & BGX_FLAG_PREFER_PAIR))) {
    ++g_kernel_count[BGX_KERNEL_AFFINE_PAIR_WIDE];
    return bgx::affine_coupling_pair(io, shift, scale, log_alpha, flags, g_status, (cudaStream_t)stream);
  }
  if (!(flags & BGX_FLAG_FORCE_SIMT) && bgx::affine_tc_eligible(shift, scale, flags)) {
    ++g_kernel_count[BGX_KERNEL_AFFINE_TC];
    return bgx::affine_coupling_tc(io, shift, scale, log_alpha, flags, g_status, (cudaStream_t)stream);
  }
  ++g_kernel_count[BGX_KERNEL_AFFINE_SIMT];
  return bgx::affine_coupling_simt(io, shift, scale, log_alpha, flags, (cudaStream_t)stream);
}

extern "C" int bgx_spline_coupling(const bgx_coupling_io* io, const bgx_packed_mlp* params_net,
                                   const bgx_spline_cfg* cfg, int flags, void* stream) {
  if (!io || !params_net || !cfg) return BGX_ERR_INVALID;
  int* status = cfg->status ? cfg->status : g_status;
  const bool tc_ok = !(flags & BGX_FLAG_FORCE_SIMT);
  const int pair_mode = (tc_ok && !(flags & BGX_FLAG_NO_PAIR)) ? bgx::spline_pair_eligible(io, params_net, cfg, flags) : 0;
  const bool tc2_ok = tc_ok && !(flags & BGX_FLAG_FORCE_WIDE) && bgx::spline_tc2_eligible(io, params_net, cfg, flags);
  if (pair_mode) {      // (measured at D = 66: 0.849 ms per launch vs 0.881 ms for the two-CTAs-per-SM kernel)
    ++g_kernel_count[pair_mode == 2 ? BGX_KERNEL_SPLINE_PAIR_WIDE : BGX_KERNEL_SPLINE_PAIR];
    return bgx::spline_coupling_pair(io, params_net, cfg, flags, pair_mode, status, (cudaStream_t)stream);
  }
  if (tc2_ok) {
    ++g_kernel_count[BGX_KERNEL_SPLINE_TC2];
    return bgx::spline_coupling_tc2(io, params_net, cfg, flags, status, (cudaStream_t)stream);
  }
  if (!(flags & BGX_FLAG_FORCE_SIMT) && bgx::spline_tc_eligible(params_net, cfg, 0)) {
    ++g_kernel_count[BGX_KERNEL_SPLINE_TC];
    return bgx::spline_coupling_tc(io, params_net, cfg, flags, cfg->status ? cfg->status : g_status,
                                   (cudaStream_t)stream);
  }
  ++g_kernel_count[BGX_KERNEL_SPLINE_SIMT];
  return bgx::spline_coupling_simt(io, params_net, cfg, flags, (cudaStream_t)stream);
}

extern "C" int bgx_debug_set_trace(uint64_t* device_buffer, int capacity) {
  bgx::tc_set_trace((unsigned long long*)device_buffer, capacity);
  return BGX_OK;
}
