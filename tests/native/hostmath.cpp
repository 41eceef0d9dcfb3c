// Test harness: compiles the __host__ __device__ arithmetic of the CDF-map kernels
// (bgflow_b200/csrc/bgx_cdf_math.cuh) with g++ so that tests/test_native_math.py can check the very
// same code against the oracle without a GPU.  Never part of the product library.
#include "bgx_cdf_math.cuh"

extern "C" float hm_std_normal_icdf(float p) { return bgx::std_normal_icdf(p); }
extern "C" float hm_std_normal_cdf(float z) { return bgx::std_normal_cdf(z); }

extern "C" int hm_cdf_apply(int kind, double a, double b, double lower, double upper, int inverse, float clamp_lo,
                            float clamp_hi, float ld_min, int n, const float* in, float* out, float* ld) {
  bgx_cdf_col col;
  int rc = bgx::cdf_col_init_host(kind, a, b, lower, upper, &col);
  if (rc) return rc;
  const bgx::CdfClamp k = {clamp_lo, clamp_hi, ld_min};
  for (int i = 0; i < n; ++i) {
    if (inverse) bgx::cdf_inverse(col, k, in[i], out[i], ld[i]);
    else bgx::cdf_forward(col, k, in[i], out[i], ld[i]);
  }
  return 0;
}

// ---- register spline evaluation of the tensor-core kernels (bgx_spline_reg.cuh), host build
#include "bgx_spline_reg.cuh"

// params: n x 25 floats per evaluation ([W(8) H(8) S(9)], the kernels' dim-major layout); x, y, lad: n
extern "C" int hm_rqs_eval(int root, int fast, float left, float right, float bottom, float top, float min_w,
                           float min_h, float min_d, int identity_init, int n, const float* params,
                           const float* x, float* y, float* lad) {
  bgx::SplineK c;
  const float wx = right - left, hy = top - bottom;
  const float beta = identity_init ? (float)(0.6931471805599453 / (1.0 - (double)min_d)) : 1.f;
  c.left = left; c.right = right; c.bottom = bottom; c.top = top;
  c.wscale = wx * (1.f - min_w * bgx::NB); c.hscale = hy * (1.f - min_h * bgx::NB);
  c.wstep = wx * min_w; c.hstep = hy * min_h;
  c.min_d = min_d; c.beta = beta; c.beta_l2e = beta * bgx::LOG2E; c.ln2_over_beta = bgx::LN2 / beta;
  for (int i = 0; i < n; ++i) {
    float p[bgx::PS];
    for (int k = 0; k < bgx::PS; ++k) p[k] = params[i * bgx::PS + k];
    const float xi = fminf(fmaxf(x[i], left), right);
    if (root) {
      if (fast) bgx::rqs_eval_reg<true, true>(p, c, xi, y[i], lad[i]);
      else bgx::rqs_eval_reg<true, false>(p, c, xi, y[i], lad[i]);
    } else {
      if (fast) bgx::rqs_eval_reg<false, true>(p, c, xi, y[i], lad[i]);
      else bgx::rqs_eval_reg<false, false>(p, c, xi, y[i], lad[i]);
    }
  }
  return 0;
}

// ---- the packed two-dims-per-thread evaluation (bgx_spline_reg2.cuh), host build: pairs (2i, 2i+1)
#include "bgx_spline_reg2.cuh"

extern "C" int hm_rqs_eval2(int root, float left, float right, float bottom, float top, float min_w, float min_h,
                            float min_d, int identity_init, int n, const float* params, const float* x, float* y,
                            float* lad) {
  bgx::SplineK c;
  const float wx = right - left, hy = top - bottom;
  const float beta = identity_init ? (float)(0.6931471805599453 / (1.0 - (double)min_d)) : 1.f;
  c.left = left; c.right = right; c.bottom = bottom; c.top = top;
  c.wscale = wx * (1.f - min_w * bgx::NB); c.hscale = hy * (1.f - min_h * bgx::NB);
  c.wstep = wx * min_w; c.hstep = hy * min_h;
  c.min_d = min_d; c.beta = beta; c.beta_l2e = beta * bgx::LOG2E; c.ln2_over_beta = bgx::LN2 / beta;
  for (int i = 0; i + 1 < n; i += 2) {
    bgx::F2 p[bgx::PS];
    for (int k = 0; k < bgx::PS; ++k) p[k] = bgx::f2(params[i * bgx::PS + k], params[(i + 1) * bgx::PS + k]);
    const bgx::F2 xi = bgx::f2(fminf(fmaxf(x[i], left), right), fminf(fmaxf(x[i + 1], left), right));
    bgx::F2 yy, ll;
    if (root) bgx::rqs_eval_reg2<true>(p, c, xi, yy, ll);
    else bgx::rqs_eval_reg2<false>(p, c, xi, yy, ll);
    y[i] = bgx::lo(yy); y[i + 1] = bgx::hi(yy);
    lad[i] = bgx::lo(ll); lad[i + 1] = bgx::hi(ll);
  }
  return 0;
}
