// Argument structures shared by the SIMT and the tensor-core coupling kernels.
#pragma once
#include "bgx_common.cuh"

namespace bgx {

struct DevMlp {
  int n_layers, act;
  int K[BGX_MAX_LAYERS], N[BGX_MAX_LAYERS], Kp[BGX_MAX_LAYERS], Np[BGX_MAX_LAYERS];
  const float* Wt[BGX_MAX_LAYERS];
  const float* bias[BGX_MAX_LAYERS];
  const int* in_map;
  float pscale, pleft;
};

struct Segs {
  int n;
  const float* ptr[BGX_MAX_SEGS];
  int width[BGX_MAX_SEGS];
  int stride[BGX_MAX_SEGS];
};

struct CouplingArgs {
  long long B;
  Segs cond, tin, tout;
  int D_t;
  DevMlp net0, net1;  // affine: shift, scale ; spline: params_net, -
  int has0, has1;
  float alpha;
  int flags;
  int dpp, pstride;   // spline column layout
  SplineParams sp;
  const float* dlogp_in;
  float* dlogp_out;
  int hb;             // rows of each activation buffer
};

__device__ __forceinline__ const float* seg_addr(const Segs& s, long long row, int col) {
  int i = 0;
  while (i < s.n - 1 && col >= s.width[i]) {
    col -= s.width[i];
    ++i;
  }
  return s.ptr[i] + row * (long long)s.stride[i] + col;
}


// layer-0 input element (global row, input column k) with WrapPeriodic folded in
__device__ __forceinline__ float load_cond(const Segs& cond, const DevMlp& net, long long B, long long row,
                                           int k) {
  if (row >= B || k >= net.K[0]) return 0.f;
  int code = net.in_map[k];
  int col = code & 0xffffff, kind = code >> 24;
  float v = __ldg(seg_addr(cond, row, col));
  if (kind == 0) return v;
  float arg = (v - net.pleft) * net.pscale;
  return kind == 1 ? cosf(arg) : sinf(arg);
}

void mlp_to_dev(const bgx_packed_mlp* p, DevMlp& d);
int coupling_fill_io(const bgx_coupling_io* io, CouplingArgs& a, int& d_c, int& d_t);
void spline_params_from_cfg(const bgx_spline_cfg* cfg, SplineParams& sp);

}  // namespace bgx
