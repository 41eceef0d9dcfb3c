// Special functions of the tensor-core epilogues, written once for device and host: on the device
// the MUFU approximations (ex2 / lg2 / rcp / sqrt, a few ulp — far inside the stated parity
// tolerance), on the host libm, so that tests/native/ can compile the very same spline arithmetic
// with g++ and check it against the oracle without a GPU.  The product only calls these in kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef BGX_HD
#if defined(__CUDACC__)
#define BGX_HD __host__ __device__ __forceinline__
#else
#define BGX_HD inline
#endif
#endif

namespace bgx {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

BGX_HD float ex2_fast(float x) {
#if defined(__CUDA_ARCH__)
  float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
  return exp2f(x);
#endif
}
BGX_HD float lg2_fast(float x) {
#if defined(__CUDA_ARCH__)
  float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
  return log2f(x);
#endif
}
BGX_HD float rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
  float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
  return 1.0f / x;
#endif
}
BGX_HD float sqrt_fast(float x) {
#if defined(__CUDA_ARCH__)
  float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#else
  return sqrtf(x);
#endif
}

// ---- packed fp32 pairs (sm_100: add / mul / fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2; one issue slot, two lanes in an
// aligned register pair).  Host build: two floats.
#if defined(__CUDA_ARCH__)
struct F2 {
  uint64_t v;
};
BGX_HD F2 f2(float lo, float hi) {
  F2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
BGX_HD float lo(F2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  (void)y;
  return x;
}
BGX_HD float hi(F2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  (void)x;
  return y;
}
BGX_HD F2 add2(F2 a, F2 b) {
  F2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
BGX_HD F2 mul2(F2 a, F2 b) {
  F2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
BGX_HD F2 fma2(F2 a, F2 b, F2 c) {
  F2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
#else
struct F2 {
  float x, y;
};
BGX_HD F2 f2(float lo, float hi) { return {lo, hi}; }
BGX_HD float lo(F2 a) { return a.x; }
BGX_HD float hi(F2 a) { return a.y; }
BGX_HD F2 add2(F2 a, F2 b) { return {a.x + b.x, a.y + b.y}; }
BGX_HD F2 mul2(F2 a, F2 b) { return {a.x * b.x, a.y * b.y}; }
BGX_HD F2 fma2(F2 a, F2 b, F2 c) { return {fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#endif
BGX_HD F2 bc2(float s) { return f2(s, s); }
BGX_HD F2 neg2(F2 a) { return f2(-lo(a), -hi(a)); }
BGX_HD F2 sub2(F2 a, F2 b) { return fma2(b, bc2(-1.f), a); }     // a - b exactly (b * -1 is exact)
template <typename Fn>
BGX_HD F2 map2(F2 a, Fn fn) { return f2(fn(lo(a)), fn(hi(a))); }
BGX_HD F2 sel2(bool ca, bool cb, F2 t, F2 f) { return f2(ca ? lo(t) : lo(f), cb ? hi(t) : hi(f)); }

}  // namespace bgx
