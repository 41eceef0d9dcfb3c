// SplitFlow / MergeFlow (bgflow/nn/flow/coupling.py:46-62,107-110) as ONE launch: column blocks of a
// [batch, width] tensor <-> n separate (dense) tensors.  The coupling kernels want dense halves
// (one bulk-TMA copy per tile needs 16-byte aligned, contiguous rows; a 33-column half of a
// 66-column tensor is neither), and torch needs one strided-copy kernel per half plus a cat.
// Pure HBM traffic: 8 B per element.  A CTA owns 64 rows; for every part its threads walk the
// [rows x w] block element-wise, so both sides see runs of w consecutive floats.
#include <algorithm>

#include "bgx_common.cuh"

namespace bgx {

constexpr int ST = 256;     // threads per CTA
constexpr int SROWS = 64;   // rows per CTA
constexpr int SU = 8;       // loads in flight per thread

struct SplitArgs {
  long long B;
  bgx_seg whole;
  int n_parts;
  bgx_seg parts[BGX_MAX_SEGS];
};

template <bool MERGE>
__global__ void __launch_bounds__(ST) split_merge_kernel(const SplitArgs a) {
  const int t = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * SROWS;
  const int nrow = (int)min((long long)SROWS, a.B - row0);
  int col0 = 0;
  for (int p = 0; p < a.n_parts; ++p) {
    const int w = a.parts[p].width;
    const int ws = a.whole.stride, ps = a.parts[p].stride;
    const float* whole = a.whole.ptr + row0 * ws + col0;
    const float* part = a.parts[p].ptr + row0 * ps;
    int m = t / w, c = t - m * w;
    const int dm = ST / w, dc = ST - dm * w;
    while (m < nrow) {          // SU loads in flight per thread before the first store
      int ms[SU], cs[SU];
      float v[SU];
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        ms[u] = m; cs[u] = c;
        m += dm; c += dc;
        if (c >= w) { c -= w; ++m; }
      }
#pragma unroll
      for (int u = 0; u < SU; ++u)
        if (ms[u] < nrow) v[u] = MERGE ? __ldg(part + ms[u] * ps + cs[u]) : __ldg(whole + ms[u] * ws + cs[u]);
#pragma unroll
      for (int u = 0; u < SU; ++u)
        if (ms[u] < nrow) {
          if (MERGE) const_cast<float*>(whole)[ms[u] * ws + cs[u]] = v[u];
          else const_cast<float*>(part)[ms[u] * ps + cs[u]] = v[u];
        }
    }
    col0 += w;
  }
}

}  // namespace bgx

using namespace bgx;

extern "C" int bgx_split_merge(int64_t batch, const bgx_seg* whole, int32_t n_parts, const bgx_seg* parts, int merge,
                               void* stream) {
  if (batch < 0 || !whole || !parts || n_parts < 1 || n_parts > BGX_MAX_SEGS || !whole->ptr) return BGX_ERR_INVALID;
  SplitArgs a{};
  a.B = batch;
  a.whole = *whole;
  a.n_parts = n_parts;
  int total = 0;
  for (int p = 0; p < n_parts; ++p) {
    if (!parts[p].ptr || parts[p].width < 1 || parts[p].stride < parts[p].width) return BGX_ERR_INVALID;
    a.parts[p] = parts[p];
    total += parts[p].width;
  }
  if (total != whole->width || whole->stride < whole->width) return BGX_ERR_INVALID;
  if ((long long)SROWS * whole->stride > 0x7fffffffLL) return BGX_ERR_UNSUPPORTED;
  if (batch == 0) return BGX_OK;
  const long long grid = (batch + SROWS - 1) / SROWS;
  if (grid > 0x7fffffffLL) return BGX_ERR_UNSUPPORTED;
  if (merge) split_merge_kernel<true><<<(unsigned)grid, ST, 0, (cudaStream_t)stream>>>(a);
  else split_merge_kernel<false><<<(unsigned)grid, ST, 0, (cudaStream_t)stream>>>(a);
  return post_launch();
}

// ---- exact two-term bf16 split of an fp32 tensor (training path: operands of the tensor-core backward GEMMs) ----
// hi = rn_bf16(x), lo = rn_bf16(x - hi): x = hi + lo up to 2^-17 |x|, so three bf16 GEMMs with fp32 accumulation
// (hi.hi + hi.lo + lo.hi) reproduce the fp32 product to ~2^-16 — the same scheme as the fused forward kernels.
namespace bgx {
__global__ void split_bf16_kernel(const float4* __restrict__ x, long long n4, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                  const float* __restrict__ x_tail, int tail, unsigned short* __restrict__ hi_tail,
                                  unsigned short* __restrict__ lo_tail) {
  auto split2 = [](float a, float b, unsigned int& h, unsigned int& l) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(h << 16), rb = b - __uint_as_float(h & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(rb), "f"(ra));
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = x[i];
    uint2 h, l;
    split2(v.x, v.y, h.x, l.x);
    split2(v.z, v.w, h.y, l.y);
    hi[i] = h;
    lo[i] = l;
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) {
    unsigned int h, l;
    split2(x_tail[threadIdx.x], 0.f, h, l);
    hi_tail[threadIdx.x] = (unsigned short)(h & 0xffffu);
    lo_tail[threadIdx.x] = (unsigned short)(l & 0xffffu);
  }
}
}  // namespace bgx

extern "C" int bgx_split_bf16(const float* x, int64_t n, void* hi, void* lo, void* stream) {
  if (n < 0 || (n > 0 && (!x || !hi || !lo))) return BGX_ERR_INVALID;
  if (n == 0) return BGX_OK;
  if (((uintptr_t)x & 15) || ((uintptr_t)hi & 7) || ((uintptr_t)lo & 7)) return BGX_ERR_INVALID;
  const long long n4 = n / 4;
  const int tail = (int)(n - 4 * n4);
  const unsigned grid = (unsigned)std::min<long long>(std::max<long long>((n4 + 255) / 256, 1), 148 * 16);
  bgx::split_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), n4, reinterpret_cast<uint2*>(hi), reinterpret_cast<uint2*>(lo), x + 4 * n4, tail,
      reinterpret_cast<unsigned short*>(hi) + 4 * n4, reinterpret_cast<unsigned short*>(lo) + 4 * n4);
  return bgx::post_launch();
}
