// Column-wise CDF maps of bgflow's IC-domain layers, written once for host and device.
//
//   CDFTransform._forward / _inverse        bgflow/nn/flow/cdf.py:29-46
//   TruncatedNormalDistribution.cdf/icdf/log_prob   bgflow/distribution/normal.py:185-196
//   torch.distributions.Normal / Uniform (as used by factory/icmarginals.py:39-77; SloppyUniform
//   = Uniform with a tolerance-widened support, distribution/distributions.py:71-97)
//
// The functions below are __host__ __device__ so that tests/native/ can compile the very same
// arithmetic with g++ and check it against the oracle without a GPU; the product only ever calls
// them from kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#include "bgflow_b200.h"

#ifndef BGX_HD
#if defined(__CUDACC__)
#define BGX_HD __host__ __device__ __forceinline__
#else
#define BGX_HD inline
#endif
#endif

// Device code uses the MUFU approximations (lg2 / rcp / sqrt: a few ulp) for the logarithm, the
// divisions and the square root of the inverse normal cdf; the host build (tests/native) uses libm.
#if defined(__CUDA_ARCH__)
#define BGX_CDF_LOG(x) __logf(x)
#define BGX_CDF_DIV(a, b) __fdividef((a), (b))
#define BGX_CDF_SQRT(x) __fsqrt_rn(x)   /* rare branches only */
#else
#define BGX_CDF_LOG(x) logf(x)
#define BGX_CDF_DIV(a, b) ((a) / (b))
#define BGX_CDF_SQRT(x) sqrtf(x)
#endif

namespace bgx {

// indices into bgx_cdf_col::p
enum { CP_LOC = 0, CP_SCALE = 1, CP_CDF_LO = 2, CP_Z = 3, CP_LOGNORM = 4, CP_INV_SCALE = 5, CP_HIGH = 6 };

constexpr float CDF_INV_SQRT2 = 0.70710678118654752440f;

// Phi(z) through erfc: no cancellation in the lower tail (the reference's 0.5 (1 + erf(z / sqrt 2))
// loses all digits below z ~ -5 in fp32; both agree to fp32 rounding elsewhere).
BGX_HD float std_normal_cdf(float z) { return 0.5f * erfcf(-z * CDF_INV_SQRT2); }

// Phi^-1(p) = sqrt(2) erfinv(2p - 1).  Main branches: M. Giles' single-precision erfinv polynomials in
// w = -log(4 p (1 - p)) (= -log((1 - x)(1 + x)), formed from p directly so that small p keeps its
// relative precision, which torch's erfinv(2p - 1) does not): w < 5 covers 0.0017 < p < 0.9983, so
// nearly every warp runs one 9-term polynomial; 5 <= w < 16 the tail polynomial in sqrt(w).  Beyond
// (p < 2.8e-8, reachable only with eps = None): Wichura's AS241 PPND7 far-tail branch.  Relative
// error < 3e-7 over (0, 1), checked against scipy.special.ndtri in tests/test_native_math.py.
BGX_HD float std_normal_icdf(float p) {
  const float x = 2.0f * p - 1.0f;
  float w = -BGX_CDF_LOG(4.0f * p * (1.0f - p));
  if (w < 5.0f) {
    w -= 2.5f;
    float q = 2.81022636e-08f;
    q = fmaf(q, w, 3.43273939e-07f);
    q = fmaf(q, w, -3.5233877e-06f);
    q = fmaf(q, w, -4.39150654e-06f);
    q = fmaf(q, w, 0.00021858087f);
    q = fmaf(q, w, -0.00125372503f);
    q = fmaf(q, w, -0.00417768164f);
    q = fmaf(q, w, 0.246640727f);
    q = fmaf(q, w, 1.50140941f);
    return 1.41421356237309504880f * x * q;
  }
  if (w < 16.0f) {
    w = BGX_CDF_SQRT(w) - 3.0f;
    float q = -0.000200214257f;
    q = fmaf(q, w, 0.000100950558f);
    q = fmaf(q, w, 0.00134934322f);
    q = fmaf(q, w, -0.00367342844f);
    q = fmaf(q, w, 0.00573950773f);
    q = fmaf(q, w, -0.0076224613f);
    q = fmaf(q, w, 0.00943887047f);
    q = fmaf(q, w, 1.00167406f);
    q = fmaf(q, w, 2.83297682f);
    return 1.41421356237309504880f * x * q;
  }
  if (!(p > 0.0f)) return -INFINITY;
  if (!(p < 1.0f)) return INFINITY;
  float r = x < 0.f ? p : 1.0f - p;
  r = BGX_CDF_SQRT(-BGX_CDF_LOG(r));
  float z;
  if (r <= 5.0f) {
    r -= 1.6f;
    z = BGX_CDF_DIV(((1.7023821103e-01f * r + 1.3067284816e+00f) * r + 2.7568153900e+00f) * r + 1.4234372777e+00f,
                    (1.2021132975e-01f * r + 7.3700164250e-01f) * r + 1.0f);
  } else {
    r -= 5.0f;
    z = BGX_CDF_DIV(((1.7337203997e-02f * r + 4.2868294337e-01f) * r + 3.0812263860e+00f) * r + 6.6579051150e+00f,
                    (1.2258202635e-02f * r + 2.4197894225e-01f) * r + 1.0f);
  }
  return x < 0.f ? -z : z;
}

struct CdfClamp {
  float lo, hi;      // cdf values are clamped to [eps, 1 - eps] (cdf.py:31-32,40-41)
  float ld_min;      // log-dets are clamped from below at -1/eps (cdf.py:34-35,44-45)
};

// x -> u = cdf(x), logdet = log_prob(x)            (CDFTransform._forward)
BGX_HD void cdf_forward(const bgx_cdf_col& c, const CdfClamp& k, float x, float& u, float& ld) {
  float lp;
  if (c.kind == BGX_DIST_NONE) {
    u = x; ld = 0.f;
    return;
  }
  if (c.kind == BGX_DIST_UNIFORM) {
    u = (x - c.p[CP_LOC]) * c.p[CP_INV_SCALE];
    u = fminf(fmaxf(u, 0.f), 1.f);                                    // torch Uniform.cdf clamps
    lp = (x >= c.p[CP_LOC] && x < c.p[CP_HIGH]) ? -c.p[CP_LOGNORM] : -INFINITY;
  } else {
    const float z = (x - c.p[CP_LOC]) * c.p[CP_INV_SCALE];
    u = std_normal_cdf(z);
    if (c.kind == BGX_DIST_TRUNCNORMAL) u = BGX_CDF_DIV(u - c.p[CP_CDF_LO], c.p[CP_Z]);
    lp = -0.5f * z * z - c.p[CP_LOGNORM];
  }
  u = fminf(fmaxf(u, k.lo), k.hi);
  ld = fmaxf(lp, k.ld_min);
}

// u -> x = icdf(clamp(u)), logdet = -log_prob(x)   (CDFTransform._inverse)
BGX_HD void cdf_inverse(const bgx_cdf_col& c, const CdfClamp& k, float u, float& x, float& ld) {
  if (c.kind == BGX_DIST_NONE) {
    x = u; ld = 0.f;
    return;
  }
  u = fminf(fmaxf(u, k.lo), k.hi);
  float lp;
  if (c.kind == BGX_DIST_UNIFORM) {
    x = fmaf(u, c.p[CP_SCALE], c.p[CP_LOC]);
    lp = (x >= c.p[CP_LOC] && x < c.p[CP_HIGH]) ? -c.p[CP_LOGNORM] : -INFINITY;
  } else {
    const float r = (c.kind == BGX_DIST_TRUNCNORMAL) ? fmaf(c.p[CP_Z], u, c.p[CP_CDF_LO]) : u;
    const float z = std_normal_icdf(r);
    x = fmaf(z, c.p[CP_SCALE], c.p[CP_LOC]);
    lp = -0.5f * z * z - c.p[CP_LOGNORM];
  }
  ld = fmaxf(-lp, k.ld_min);
}

// Host side of bgx_cdf_col_init: per-column constants derived in double precision.
inline int cdf_col_init_host(int32_t kind, double a, double b, double lower, double upper, bgx_cdf_col* out) {
  if (!out) return BGX_ERR_INVALID;
  for (int i = 0; i < 7; ++i) out->p[i] = 0.f;
  out->kind = kind;
  const double half_log_2pi = 0.91893853320467274178;
  switch (kind) {
    case BGX_DIST_NONE:
      return BGX_OK;
    case BGX_DIST_NORMAL:
      if (!(b > 0.0)) return BGX_ERR_INVALID;
      out->p[CP_LOC] = (float)a; out->p[CP_SCALE] = (float)b;
      out->p[CP_CDF_LO] = 0.f; out->p[CP_Z] = 1.f;
      out->p[CP_LOGNORM] = (float)(::log(b) + half_log_2pi);
      out->p[CP_INV_SCALE] = (float)(1.0 / b);
      return BGX_OK;
    case BGX_DIST_TRUNCNORMAL: {
      if (!(b > 0.0) || !(upper > lower)) return BGX_ERR_INVALID;
      // normal.py:148-151: Phi((lower - mu) / sigma), Phi((upper - mu) / sigma)
      const double lo = 0.5 * ::erfc(-((lower - a) / b) * 0.70710678118654752440);
      const double hi = 0.5 * ::erfc(-((upper - a) / b) * 0.70710678118654752440);
      const double Z = hi - lo;
      if (!(Z > 0.0)) return BGX_ERR_INVALID;
      out->p[CP_LOC] = (float)a; out->p[CP_SCALE] = (float)b;
      out->p[CP_CDF_LO] = (float)lo; out->p[CP_Z] = (float)Z;
      out->p[CP_LOGNORM] = (float)(::log(Z * b) + half_log_2pi);   // normal.py:195
      out->p[CP_INV_SCALE] = (float)(1.0 / b);
      return BGX_OK;
    }
    case BGX_DIST_UNIFORM:
      if (!(b > a)) return BGX_ERR_INVALID;
      out->p[CP_LOC] = (float)a; out->p[CP_SCALE] = (float)(b - a);
      out->p[CP_LOGNORM] = (float)::log(b - a);
      out->p[CP_INV_SCALE] = (float)(1.0 / (b - a));
      out->p[CP_HIGH] = (float)b;
      return BGX_OK;
    default:
      return BGX_ERR_INVALID;
  }
}

}  // namespace bgx
