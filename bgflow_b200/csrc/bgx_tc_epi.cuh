// Helpers shared by the tensor-core coupling kernels (spline and affine): MUFU math, exact bf16
// operand splits, asynchronous tile copies.
#pragma once
#include "bgx_common.cuh"
#include "bgx_fastmath.cuh"
#include "bgx_tc.cuh"

namespace bgx {
using namespace tc;

// (MUFU special functions: bgx_fastmath.cuh)
template <int ACT>
__device__ __forceinline__ float act_fast(float x) {
  if (ACT == BGX_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == BGX_ACT_SILU) return x * rcp_fast(1.f + ex2_fast(-LOG2E * x));
  if (ACT == BGX_ACT_TANH) return 1.f - 2.f * rcp_fast(1.f + ex2_fast(2.f * LOG2E * x));
  return x;
}


// exact split of two adjacent-k fp32 values into packed bf16x2 terms (element k in the low half)
__device__ __forceinline__ void split_bf16(float x0, float x1, int nterms, uint32_t& t1, uint32_t& t2,
                                           uint32_t& t3) {
  t1 = pack_bf16x2(x0, x1);
  const float r0 = x0 - bf16_lo_to_f32(t1), r1 = x1 - bf16_hi_to_f32(t1);
  t2 = pack_bf16x2(r0, r1);
  t3 = 0;
  if (nterms == 3) t3 = pack_bf16x2(r0 - bf16_lo_to_f32(t2), r1 - bf16_hi_to_f32(t2));
}

// Two adjacent columns (k, k + 1) of a hidden layer at once: bias, activation and the exact two-term bf16 split in packed
// fp32 pairs (the accumulator registers of tcgen05.ld and the float4 bias loads are already adjacent pairs).
template <int ACT>
__device__ __forceinline__ F2 act_fast2(F2 x) {
  if (ACT == BGX_ACT_RELU) return f2(fmaxf(lo(x), 0.f), fmaxf(hi(x), 0.f));
  if (ACT == BGX_ACT_SILU)
    return mul2(x, map2(add2(bc2(1.f), map2(mul2(x, bc2(-LOG2E)), [](float v) { return ex2_fast(v); })),
                        [](float v) { return rcp_fast(v); }));
  if (ACT == BGX_ACT_TANH)
    return fma2(bc2(-2.f), map2(add2(bc2(1.f), map2(mul2(x, bc2(2.f * LOG2E)), [](float v) { return ex2_fast(v); })),
                                [](float v) { return rcp_fast(v); }), bc2(1.f));
  return x;
}
template <int ACT>
__device__ __forceinline__ void hidden_pair2(uint32_t v0, uint32_t v1, float b0, float b1, uint32_t& t1, uint32_t& t2) {
  const F2 h = act_fast2<ACT>(add2(f2(__uint_as_float(v0), __uint_as_float(v1)), f2(b0, b1)));
  t1 = pack_bf16x2(lo(h), hi(h));
  const F2 r = sub2(h, f2(bf16_lo_to_f32(t1), bf16_hi_to_f32(t1)));
  t2 = pack_bf16x2(lo(r), hi(r));
}

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_drain() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// shared -> global bulk copy (TMA engine); completion tracked with bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }


}  // namespace bgx
