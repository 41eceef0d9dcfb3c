// Shared pieces of the pair kernels (bgx_coupling_pair.cu: spline, bgx_coupling_pair_affine.cu: affine):
// geometry of the CTA, tensor-memory map, and the out-of-line layer-0 operand staging.
#pragma once
#include "bgx_coupling.cuh"
#include "bgx_tc.cuh"
#include "bgx_tc_epi.cuh"

namespace bgx {
using namespace tc;

constexpr int P_EPI_WARPS = 16;
constexpr int P_THREADS = (P_EPI_WARPS + 4) * 32;     // 640
constexpr int P_TM = 128;
constexpr int P_STAGES = 2;
constexpr uint32_t P_TILE_BYTES = 16384;              // one [128 x 64] bf16 k-tile of one term
constexpr uint32_t P_KT_BYTES = 2 * P_TILE_BYTES;     // both terms
constexpr uint32_t P_STAGE_BYTES = 2 * P_KT_BYTES;    // a unit has at most two k-tiles
constexpr int P_SLOT = 256, P_ACC = 0, P_A = 128, P_A_STRIDE = 64;
constexpr int P_NB = 8, P_PS = 3 * P_NB + 1, P_DPP = 5, P_BPAD = 28;
constexpr int P_CLUSTER_DEFAULT = 2; // CTAs per cluster sharing each weight stage by TMA multicast (1, 2 or 4)
constexpr int P_EPW_DEFAULT = 4;     // epilogue warps per quadrant of the spline pair kernel (4 or 6)

// Layer-0 operand of one slot: inputs 128 g + [32 j, 32 j + 32) of this thread's row -> two exact bf16 terms in
// tensor memory.  Out of line on purpose: three call sites, and the epilogue's hot loop has to stay inside
// the instruction cache (the first version of this kernel inlined it everywhere and lost 27 % of its warp
// samples to instruction fetch).  WrapPeriodic (periodic.py:30-37) through sinpi / cospi: no slow path.
// plain_cond: 0 = through the input map (WrapPeriodic), 1 = identity map, 2 = identity map and 16-byte aligned rows whose
// width is a multiple of 4 (four LDG.128 per half instead of sixteen scalar loads: the per-thread row reads touch one
// line per lane and instruction, so the number of instructions is what the L1 pays for)
static __device__ __noinline__ void pair_stage_x(const DevMlp& net, int plain_cond, const float* crow, bool live, int g, int j,
                                                 uint32_t a_col) {
  const int K0 = net.K[0];
  const int kg = 128 * g;                 // first input of the group
  const int kend = min(K0, kg + 128);     // one past the last real input of the group
  const float inv_pi_scale = net.pscale * 0.3183098861837907f;
  // this warp's 32 inputs as two 16-input halves (8 packed columns each); a half is written iff the MMAs of
  // the group read it (their k-steps cover inputs [kg, round_up(kend, 16)))
  for (int h = 0; h < 2; ++h) {
    const int b0 = kg + j * 32 + h * 16;
    if (b0 >= kend) break;
    float xv[16];
    if (plain_cond == 2) {
      const float4* c4 = reinterpret_cast<const float4*>(crow + b0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = (live && b0 + 4 * i < kend) ? __ldg(c4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        xv[4 * i] = v.x; xv[4 * i + 1] = v.y; xv[4 * i + 2] = v.z; xv[4 * i + 3] = v.w;
      }
    } else if (plain_cond) {
#pragma unroll
      for (int i = 0; i < 16; ++i) xv[i] = (live && b0 + i < kend) ? crow[b0 + i] : 0.f;
    } else {
#pragma unroll 1
      for (int i = 0; i < 16; ++i) {
        float v = 0.f;
        if (live && b0 + i < kend) {
          const int code = net.in_map[b0 + i];
          v = crow[code & 0xffffff];
          const int kind = code >> 24;
          if (kind) {
            const float t = (v - net.pleft) * inv_pi_scale;
            v = kind == 1 ? cospif(t) : sinpif(t);
          }
        }
        xv[i] = v;
      }
    }
    uint32_t t1[8], t2[8], t3[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_bf16(xv[2 * i], xv[2 * i + 1], 2, t1[i], t2[i], t3[i]);
    const uint32_t col = a_col + (uint32_t)((b0 - kg) / 2);
    tmem_st8(col, t1);
    tmem_st8(col + P_A_STRIDE, t2);
  }
  tmem_st_wait();
  tc_fence_before();
}

}  // namespace bgx
