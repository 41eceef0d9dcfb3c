// Geometry helpers shared by the internal-coordinate kernels (bgx_ic.cu: global z-matrix,
// bgx_ic_rel.cu: relative / mixed transforms).
#pragma once
#include "bgx_common.cuh"

namespace bgx {

constexpr int BT = 128;       // samples (threads) per CTA
constexpr int LDT = BT + 1;
constexpr float PI_F = 3.14159265358979323846f;
constexpr float TWO_PI_F = 6.28318530717958647692f;

struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float norm_c(V3 a, float eps) { return fmaxf(sqrtf(dot(a, a)), eps); }
// 1 / max(|a|, eps) with one MUFU.RSQ (eps^2 = 1e-14 is representable)
__device__ __forceinline__ float inv_norm_c(V3 a, float eps2) { return rsqrtf(fmaxf(dot(a, a), eps2)); }

__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// atan2 with a degree-15 odd minimax polynomial on [0, 1] (max error 1.3e-7 rad in fp32, fitted
// against numpy's arctan) + octant fix-ups: ~25 instructions instead of libm's ~70.
__device__ __forceinline__ float atan2_fast(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = mx > 0.f ? mn * rcp_approx(mx) : 0.f;
  const float s = t * t;
  float p = -4.0545426679e-03f;
  p = fmaf(p, s, 2.1862866834e-02f);
  p = fmaf(p, s, -5.5912191968e-02f);
  p = fmaf(p, s, 9.6421871900e-02f);
  p = fmaf(p, s, -1.3908625481e-01f);
  p = fmaf(p, s, 1.9946564817e-01f);
  p = fmaf(p, s, -3.3329860709e-01f);
  p = fmaf(p, s, 9.9999933556e-01f);
  float r = t * p;
  r = ay > ax ? 1.57079632679489661923f - r : r;
  r = x < 0.f ? 3.14159265358979323846f - r : r;
  return y < 0.f ? -r : r;
}

template <bool SMEM>
struct PosStore {
  float* base;
  long long stride_c;  // distance between consecutive coordinates
  __device__ __forceinline__ V3 get(int atom) const {
    const float* p = base + (long long)(3 * atom) * stride_c;
    return {p[0], p[stride_c], p[2 * stride_c]};
  }
  __device__ __forceinline__ void set(int atom, V3 v) const {
    float* p = base + (long long)(3 * atom) * stride_c;
    p[0] = v.x;
    p[stride_c] = v.y;
    p[2 * stride_c] = v.z;
  }
};

// division-free walk of thread t over the elements e = t, t+BT, ... of a [rows x W] block
template <typename F>
__device__ __forceinline__ void walk_block(int t, int W, int rows, F&& body) {
  int m = t / W, c = t - m * W;
  const int dm = BT / W, dc = BT - dm * W;
  while (m < rows) {
    body(m, c);
    m += dm; c += dc;
    if (c >= W) { c -= W; ++m; }
  }
}

// The same walk for global -> shared staging: UNROLL loads are issued before the first of their
// values is consumed, so a thread keeps UNROLL requests in flight (one-at-a-time loads leave the
// copy phase bound by DRAM latency, not bandwidth).
template <int UNROLL = 8, typename L, typename S>
__device__ __forceinline__ void walk_block_ld(int t, int W, int rows, L&& load, S&& store) {
  int m = t / W, c = t - m * W;
  const int dm = BT / W, dc = BT - dm * W;
  while (m < rows) {
    int ms[UNROLL], cs[UNROLL];
    float v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      ms[u] = m; cs[u] = c;
      m += dm; c += dc;
      if (c >= W) { c -= W; ++m; }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (ms[u] < rows) v[u] = load(ms[u], cs[u]);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (ms[u] < rows) store(ms[u], cs[u], v[u]);
  }
}

// same walk for kernels whose CTA covers fewer samples than threads
template <typename F>
__device__ __forceinline__ void walk_block_f(int t, int W, int rows, F&& body) {
  walk_block(t, W, rows, body);
}

}  // namespace bgx
