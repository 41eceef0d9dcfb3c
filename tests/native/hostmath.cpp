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
