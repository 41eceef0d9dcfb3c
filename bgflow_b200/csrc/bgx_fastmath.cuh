// Special functions of the tensor-core epilogues, written once for device and host: on the device
// the MUFU approximations (ex2 / lg2 / rcp / sqrt, a few ulp — far inside the stated parity
// tolerance), on the host libm, so that tests/native/ can compile the very same spline arithmetic
// with g++ and check it against the oracle without a GPU.  The product only calls these in kernels.
#pragma once
#include <math.h>

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

}  // namespace bgx
