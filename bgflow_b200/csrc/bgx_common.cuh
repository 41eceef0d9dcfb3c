// Shared helpers for the bgflow_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bgflow_b200.h"

namespace bgx {

extern thread_local cudaError_t g_last_error;
extern long long g_launches;

inline int check(cudaError_t e) {
  if (e != cudaSuccess) {
    g_last_error = e;
    return BGX_ERR_CUDA;
  }
  return BGX_OK;
}
inline int post_launch() {
  ++g_launches;
  return check(cudaPeekAtLastError());
}

// cudaFuncSetAttribute, occupancy figures and the SM count are properties of (function, DEVICE): the launch helpers
// keep their once-flags per device, indexed by device_slot() (one process may drive several GPUs through the Python
// mirror's device guard).
constexpr int BGX_MAX_DEVICES = 64;
inline int device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= BGX_MAX_DEVICES) d = 0;
  return d;
}
inline int device_sm_count(int* out) {
  static int n[BGX_MAX_DEVICES] = {};
  const int d = device_slot();
  if (!n[d]) {
    int rc = check(cudaDeviceGetAttribute(&n[d], cudaDevAttrMultiProcessorCount, d));
    if (rc) return rc;
  }
  *out = n[d];
  return BGX_OK;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case BGX_ACT_RELU: return fmaxf(x, 0.f);
    case BGX_ACT_SILU: return x / (1.f + expf(-x));
    case BGX_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Runtime description of the RQ spline (device copy of bgx_spline_cfg + derived constants).
struct SplineParams {
  int K;
  float left, right, bottom, top;
  float min_w, min_h, min_d;
  float beta, inv_beta;
  int* oob;
};

__device__ __forceinline__ float softplus_beta(float x, float beta, float inv_beta) {
  // torch.nn.functional.softplus(x, beta) with its threshold=20
  float bx = beta * x;
  return bx > 20.f ? x : log1pf(expf(bx)) * inv_beta;
}

// One (sample, dim) evaluation of the rational-quadratic spline.
//   p points at this dim's K widths, K heights, K+1 slopes (unnormalised), element stride `ps`.
//   ROOT = true : nflows inverse=True  (bgflow forward, quadratic-root branch)
//   ROOT = false: nflows inverse=False (bgflow inverse, direct evaluation)
// Restates SURVEY.md A.5 / oracle/flows.py::rational_quadratic_spline.
template <bool ROOT>
__device__ __forceinline__ void rqs_eval(const float* __restrict__ p, int ps, const SplineParams& sp,
                                         float x, float& y, float& lad) {
  const int K = sp.K;
  const float* W = p;
  const float* H = p + K * ps;
  const float* S = p + 2 * K * ps;
  float mw = W[0], mh = H[0];
  for (int k = 1; k < K; ++k) {
    mw = fmaxf(mw, W[k * ps]);
    mh = fmaxf(mh, H[k * ps]);
  }
  float sw = 0.f, sh = 0.f;
  for (int k = 0; k < K; ++k) {
    sw += expf(W[k * ps] - mw);
    sh += expf(H[k * ps] - mh);
  }
  const float cw_scale = (1.f - sp.min_w * K) / sw;
  const float ch_scale = (1.f - sp.min_h * K) / sh;
  const float wx = sp.right - sp.left, hy = sp.top - sp.bottom;
  // walk the knots; remember the bin that contains x on the searched axis
  float cumw = 0.f, cumh = 0.f;
  float kw_lo = sp.left, kh_lo = sp.bottom;  // knots k
  float bw_lo = sp.left, bw_hi = sp.right, bh_lo = sp.bottom, bh_hi = sp.top;
  int bin = 0;
  for (int k = 0; k < K; ++k) {
    cumw += sp.min_w + cw_scale * expf(W[k * ps] - mw);
    cumh += sp.min_h + ch_scale * expf(H[k * ps] - mh);
    float kw_hi = (k == K - 1) ? sp.right : fmaf(wx, cumw, sp.left);
    float kh_hi = (k == K - 1) ? sp.top : fmaf(hy, cumh, sp.bottom);
    const float knot = ROOT ? kh_lo : kw_lo;
    if (k == 0 || x >= knot) {
      bin = k;
      bw_lo = kw_lo; bw_hi = kw_hi; bh_lo = kh_lo; bh_hi = kh_hi;
    }
    kw_lo = kw_hi;
    kh_lo = kh_hi;
  }
  const float w = bw_hi - bw_lo, h = bh_hi - bh_lo;
  const float delta = h / w;
  const float d0 = sp.min_d + softplus_beta(S[bin * ps], sp.beta, sp.inv_beta);
  const float d1 = sp.min_d + softplus_beta(S[(bin + 1) * ps], sp.beta, sp.inv_beta);
  const float s = d0 + d1 - 2.f * delta;
  if (ROOT) {
    const float q = x - bh_lo;
    const float a = q * s + h * (delta - d0);
    const float b = h * d0 - q * s;
    const float c = -delta * q;
    const float disc = fmaxf(b * b - 4.f * a * c, 0.f);
    const float root = (2.f * c) / (-b - sqrtf(disc));
    y = fmaf(root, w, bw_lo);
    const float t1 = root * (1.f - root);
    const float den = delta + s * t1;
    const float omr = 1.f - root;
    const float num = delta * delta * (d1 * root * root + 2.f * delta * t1 + d0 * omr * omr);
    lad = -(logf(num) - 2.f * logf(den));
  } else {
    const float th = (x - bw_lo) / w;
    const float t1 = th * (1.f - th);
    const float den = delta + s * t1;
    y = bh_lo + h * (delta * th * th + d0 * t1) / den;
    const float omt = 1.f - th;
    const float num = delta * delta * (d1 * th * th + 2.f * delta * t1 + d0 * omt * omt);
    lad = logf(num) - 2.f * logf(den);
  }
}

}  // namespace bgx
