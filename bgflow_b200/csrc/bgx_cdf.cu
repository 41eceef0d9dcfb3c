// IC-domain maps: CDFTransform (bgflow/nn/flow/cdf.py:13-46) over per-column marginals, every
// field of the flow state in ONE launch (the reference runs one WrapFlow(InverseFlow(CDFTransform))
// per field, generator_builder.py:443-459: ~10 elementwise passes + a row sum each).
//
// HBM-bound: 8 B per element + 8 B per row of dlogp.  A CTA owns 128 consecutive rows; for each
// segment a thread owns one column and every k-th row (column constants in registers, accesses of a
// pass contiguous for dense tensors), writes the mapped value straight back and parks the element's
// log-det in shared memory, column-major (leading dim 129: conflict-free for both phases); then
// thread r sums row r in a fixed order, so dlogp is deterministic.
#include <cmath>

#include "bgx_cdf_math.cuh"
#include "bgx_common.cuh"

namespace bgx {

constexpr int CT = 128;          // rows (threads) per CTA
constexpr int CLD = CT + 1;
constexpr int CU = 8;            // loads in flight per thread
constexpr int CCHUNK = 44;       // columns staged per pass: 44 * 129 * 4 B = 22.7 KB (10 CTAs / SM)

struct CdfArgs {
  long long B;
  int n_seg;
  bgx_seg in[BGX_MAX_SEGS];
  bgx_seg out[BGX_MAX_SEGS];
  const bgx_cdf_col* cols;
  CdfClamp clamp;
  const float* dlogp_in;
  float* dlogp_out;
};

template <bool INVERSE>
__global__ void __launch_bounds__(CT) cdf_map_kernel(const CdfArgs a) {
  __shared__ float ld_s[CCHUNK * CLD];
  const int t = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * CT;
  const int nrow = (int)min((long long)CT, a.B - row0);
  float acc = 0.f;
  int used = 0;                   // columns currently parked in ld_s
  int col_base = 0;               // index of the segment's first column in a.cols
  auto flush = [&]() {
    __syncthreads();
    if (t < nrow)
      for (int c = 0; c < used; ++c) acc += ld_s[c * CLD + t];
    __syncthreads();
    used = 0;
  };
  for (int s = 0; s < a.n_seg; ++s) {
    const int W = a.in[s].width;
    for (int c0 = 0; c0 < W; c0 += CCHUNK) {
      const int cw = min(CCHUNK, W - c0);
      if (used + cw > CCHUNK) flush();
      const int si = a.in[s].stride, so = a.out[s].stride;       // (128 rows x stride fits 32 bits: checked on the host)
      const float* gin = a.in[s].ptr + row0 * (long long)si + c0;
      float* gout = const_cast<float*>(a.out[s].ptr) + row0 * (long long)so + c0;
      const bgx_cdf_col* cols = a.cols + col_base + c0;
      // thread t owns ONE column of the chunk (c = t mod cw) and every (CT / cw)-th row: its column's
      // constants stay in registers, and the active threads of a pass still cover rows_per_pass * cw
      // consecutive floats of a dense tensor (fully coalesced)
      const int rpp = CT / cw;                       // rows per pass (cw <= CCHUNK < CT)
      if (t < rpp * cw) {
        int m = t / cw;
        const int c = t - m * cw;
        const bgx_cdf_col cc = cols[c];
        // CU loads in flight per thread: one-at-a-time loads would leave the kernel bound by DRAM
        // latency (40 warps x 128 B per SM in flight), not by bandwidth.  Pointer stepping, no
        // per-element index arithmetic; the (< CU rows) remainder goes one row at a time.
        const float* pin = gin + m * si + c;
        float* pout = gout + m * so + c;
        float* pld = ld_s + (used + c) * CLD + m;
        const int istep = rpp * si, ostep = rpp * so;
        for (; m + (CU - 1) * rpp < nrow; m += CU * rpp) {
          float v[CU];
#pragma unroll
          for (int u = 0; u < CU; ++u) v[u] = __ldg(pin + u * istep);
#pragma unroll
          for (int u = 0; u < CU; ++u) {
            float y, ld;
            if (INVERSE) cdf_inverse(cc, a.clamp, v[u], y, ld);
            else cdf_forward(cc, a.clamp, v[u], y, ld);
            pout[u * ostep] = y;
            pld[u * rpp] = ld;
          }
          pin += CU * istep; pout += CU * ostep; pld += CU * rpp;
        }
        for (; m < nrow; m += rpp) {
          float y, ld;
          if (INVERSE) cdf_inverse(cc, a.clamp, __ldg(pin), y, ld);
          else cdf_forward(cc, a.clamp, __ldg(pin), y, ld);
          *pout = y;
          *pld = ld;
          pin += istep; pout += ostep; pld += rpp;
        }
      }
      used += cw;
    }
    col_base += W;
  }
  flush();
  if (t < nrow) a.dlogp_out[row0 + t] = (a.dlogp_in ? a.dlogp_in[row0 + t] : 0.f) + acc;
}

}  // namespace bgx

using namespace bgx;

extern "C" int bgx_cdf_col_init(int32_t kind, double a, double b, double lower, double upper, bgx_cdf_col* out) {
  return cdf_col_init_host(kind, a, b, lower, upper, out);
}

extern "C" int bgx_cdf_col_set_truncation(bgx_cdf_col* col, double cdf_lower, double cdf_upper) {
  if (!col || col->kind != BGX_DIST_TRUNCNORMAL) return BGX_ERR_INVALID;
  const double Z = cdf_upper - cdf_lower;
  if (!(Z > 0.0)) return BGX_ERR_INVALID;
  col->p[CP_CDF_LO] = (float)cdf_lower;
  col->p[CP_Z] = (float)Z;
  col->p[CP_LOGNORM] = (float)(::log(Z * (double)col->p[CP_SCALE]) + 0.91893853320467274178);
  return BGX_OK;
}

extern "C" int bgx_cdf_map(int64_t batch, int32_t n_seg, const bgx_seg* in, const bgx_seg* out,
                           const bgx_cdf_col* cols, float clamp_lo, float clamp_hi, float logdet_min, int flags,
                           const float* dlogp_in, float* dlogp_out, void* stream) {
  if (batch < 0 || n_seg < 1 || n_seg > BGX_MAX_SEGS || !in || !out || !cols || !dlogp_out) return BGX_ERR_INVALID;
  CdfArgs a{};
  a.B = batch;
  a.n_seg = n_seg;
  for (int s = 0; s < n_seg; ++s) {
    if (!in[s].ptr || !out[s].ptr || in[s].width < 1 || out[s].width != in[s].width || in[s].stride < in[s].width ||
        out[s].stride < out[s].width)
      return BGX_ERR_INVALID;
    if ((long long)CT * in[s].stride > 0x7fffffffLL || (long long)CT * out[s].stride > 0x7fffffffLL) return BGX_ERR_UNSUPPORTED;
    a.in[s] = in[s];
    a.out[s] = out[s];
  }
  a.cols = cols;
  a.clamp = {clamp_lo, clamp_hi, logdet_min};
  a.dlogp_in = dlogp_in;
  a.dlogp_out = dlogp_out;
  if (batch == 0) return BGX_OK;
  const long long grid = (batch + CT - 1) / CT;
  if (grid > 0x7fffffffLL) return BGX_ERR_UNSUPPORTED;
  if (flags & BGX_FLAG_INVERSE) cdf_map_kernel<true><<<(unsigned)grid, CT, 0, (cudaStream_t)stream>>>(a);
  else cdf_map_kernel<false><<<(unsigned)grid, CT, 0, (cudaStream_t)stream>>>(a);
  return post_launch();
}
