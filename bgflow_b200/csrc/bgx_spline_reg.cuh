// Register-resident rational-quadratic spline evaluation (8 bins) shared by the tensor-core
// spline kernels; __host__ __device__, so tests/native/ checks this very arithmetic against the
// oracle on the CPU (tests/test_native_math.py).  Same algorithm as rqs_eval (bgx_common.cuh) / oracle.flows.rational_quadratic_spline.
#pragma once
#include "bgx_fastmath.cuh"

namespace bgx {

constexpr int NB = 8;            // spline bins handled by the tensor-core kernels
constexpr int PS = 3 * NB + 1;   // parameters per transformed dim

// per-kernel constants of the spline (uniform across threads)
struct SplineK {
  float left, right, bottom, top;
  float wscale, hscale;   // (right-left)*(1-min_w*K), (top-bottom)*(1-min_h*K)
  float wstep, hstep;     // (right-left)*min_w, (top-bottom)*min_h
  float min_d, beta_l2e, ln2_over_beta, beta;
};

BGX_HD float softplus_fast(float s, const SplineK& c) {
  // softplus(s) >= s, and lg2(1 + 2^t) -> t for large t: the max() only matters once the exponent
  // clamp (t > 64, i.e. slopes above ~44) bites, where torch's thresholded softplus returns s too
  const float v = lg2_fast(1.f + ex2_fast(fminf(s * c.beta_l2e, 64.f))) * c.ln2_over_beta;
  return fmaxf(v, s);
}

// One (sample, dim) spline evaluation with the 25 parameters in registers (NB = 8 bins).
// Same algorithm as rqs_eval (bgx_common.cuh); the knots are formed from prefix sums of the
// softmax numerators instead of a running sum of the normalised bins.
// FASTLOG: log-det through MUFU lg2 (absolute error <= 2^-22 ln 2 per evaluation) instead of logf,
// binary instead of linear bin search.
template <bool ROOT, bool FASTLOG = false>
BGX_HD void rqs_eval_reg(const float (&p)[PS], const SplineK& c, float x, float& y,
                                             float& lad) {
  const float mw = fmaxf(fmaxf(fmaxf(p[0], p[1]), fmaxf(p[2], p[3])), fmaxf(fmaxf(p[4], p[5]), fmaxf(p[6], p[7])));
  const float mh = fmaxf(fmaxf(fmaxf(p[8], p[9]), fmaxf(p[10], p[11])), fmaxf(fmaxf(p[12], p[13]), fmaxf(p[14], p[15])));
  const float nmw = -mw * LOG2E, nmh = -mh * LOG2E;
  float pw[NB], ph[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    pw[k] = ex2_fast(fmaf(p[k], LOG2E, nmw));
    ph[k] = ex2_fast(fmaf(p[NB + k], LOG2E, nmh));
  }
#pragma unroll
  for (int k = 1; k < NB; ++k) {
    pw[k] += pw[k - 1];
    ph[k] += ph[k - 1];
  }
  const float aw = c.wscale * rcp_fast(pw[NB - 1]);
  const float ah = c.hscale * rcp_fast(ph[NB - 1]);
  // knots 1..NB-1 (knot 0 = left/bottom, knot NB = right/top exactly)
  float kw[NB + 1], kh[NB + 1];
  kw[0] = c.left; kh[0] = c.bottom; kw[NB] = c.right; kh[NB] = c.top;
#pragma unroll
  for (int k = 1; k < NB; ++k) {
    kw[k] = fmaf(aw, pw[k - 1], fmaf(c.wstep, (float)k, c.left));
    kh[k] = fmaf(ah, ph[k - 1], fmaf(c.hstep, (float)k, c.bottom));
  }
  float w_lo, w_hi, h_lo, h_hi, s0, s1;
  if (FASTLOG) {
    // binary search over the 8 bins: 3 dependent compares, 30 selects (the linear walk below needs
    // 7 compares and 42 selects in a 7-deep chain); same bin: the largest k with x >= knot[k]
    const float* sl = p + 2 * NB;
    const bool c4 = x >= (ROOT ? kh[4] : kw[4]);
    float qw[5], qh[5], qs[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      qw[q] = c4 ? kw[4 + q] : kw[q];
      qh[q] = c4 ? kh[4 + q] : kh[q];
      qs[q] = c4 ? sl[4 + q] : sl[q];
    }
    const bool c2 = x >= (ROOT ? qh[2] : qw[2]);
    float bw[3], bh[3], bs[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      bw[q] = c2 ? qw[2 + q] : qw[q];
      bh[q] = c2 ? qh[2 + q] : qh[q];
      bs[q] = c2 ? qs[2 + q] : qs[q];
    }
    const bool c1 = x >= (ROOT ? bh[1] : bw[1]);
    w_lo = c1 ? bw[1] : bw[0]; w_hi = c1 ? bw[2] : bw[1];
    h_lo = c1 ? bh[1] : bh[0]; h_hi = c1 ? bh[2] : bh[1];
    s0 = c1 ? bs[1] : bs[0]; s1 = c1 ? bs[2] : bs[1];
  } else {
  w_lo = kw[0]; w_hi = kw[1]; h_lo = kh[0]; h_hi = kh[1]; s0 = p[2 * NB]; s1 = p[2 * NB + 1];
#pragma unroll
  for (int k = 1; k < NB; ++k) {
    const bool in = x >= (ROOT ? kh[k] : kw[k]);
    w_lo = in ? kw[k] : w_lo;
    w_hi = in ? kw[k + 1] : w_hi;
    h_lo = in ? kh[k] : h_lo;
    h_hi = in ? kh[k + 1] : h_hi;
    s0 = in ? p[2 * NB + k] : s0;
    s1 = in ? p[2 * NB + k + 1] : s1;
  }
  }
  const float w = w_hi - w_lo, h = h_hi - h_lo;
  const float rw = rcp_fast(w);
  const float delta = h * rw;
  const float d0 = c.min_d + softplus_fast(s0, c);
  const float d1 = c.min_d + softplus_fast(s1, c);
  const float s = d0 + d1 - 2.f * delta;
  float th;
  if (ROOT) {
    const float q = x - h_lo;
    const float qs = q * s;
    const float a = fmaf(h, delta - d0, qs);
    const float b = fmaf(h, d0, -qs);
    const float cc = -delta * q;
    const float disc = fmaxf(fmaf(b, b, -4.f * a * cc), 0.f);
    th = 2.f * cc * rcp_fast(-b - sqrt_fast(disc));
    y = fmaf(th, w, w_lo);
  } else {
    th = (x - w_lo) * rw;
  }
  const float omt = 1.f - th;
  const float t1 = th * omt;
  const float den = fmaf(s, t1, delta);
  const float rden = rcp_fast(den);
  if (!ROOT) y = fmaf(h * fmaf(delta * th, th, d0 * t1), rden, h_lo);
  const float num = delta * delta * fmaf(d1 * th, th, fmaf(2.f * delta, t1, d0 * omt * omt));
  const float l = FASTLOG ? LN2 * lg2_fast(num * rden * rden) : logf(num * rden * rden);
  lad = ROOT ? -l : l;
}


}  // namespace bgx
