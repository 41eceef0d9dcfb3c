// Tensor-core fused RQ-spline coupling block, TWO CTAs PER SM variant.
//
// Same math and building blocks as bgx_coupling_tc.cu (tcgen05.mma kind::f16 on exact bf16 operand
// splits, A operand + accumulator in tensor memory, bulk-TMA weight tiles, spline epilogue in
// registers), but each CTA takes only HALF of the SM's resources — 256 TMEM columns, <= 112 KB of
// shared memory, 320 threads — so that two CTAs are co-resident and the hardware interleaves them:
// while one CTA's epilogue warps work on an accumulator, the other CTA's MMAs own the tensor pipe.
// The single-CTA kernel serialises a tile's hidden-layer epilogues with its own MMAs because it
// needs all 512 TMEM columns; here the serialisation inside a CTA is hidden by its neighbour.
//
//   warps 0-7  epilogue (thread <-> sample row; two warps per TMEM lane quadrant)
//   warp  8    lane 0: weight producer (bulk TMA into a 2-slot ring) + tile I/O (one bulk copy per
//              tile for the conditioner input, the transformed input and the output)
//   warp  9    lane 0: tcgen05.mma issuer; the warp owns the 256-column TMEM allocation
//
// TMEM columns: [0,128) ACC, [128,192) A1, [192,256) A2   (bf16x3 only: two operand terms).
// Requires dense (single-tensor, contiguous, 16-B aligned) conditioner / input / output tensors;
// everything else runs the single-CTA kernel or the SIMT kernel.
#include <cstdlib>

#include "bgx_coupling.cuh"
#include "bgx_tc.cuh"
#include "bgx_tc_epi.cuh"
#include "bgx_spline_reg.cuh"
#include "bgx_spline_reg2.cuh"

namespace bgx {
using namespace tc;

constexpr int T2_THREADS = 320;
constexpr int T2_EPI_WARPS = 8;
constexpr int T2_TM = 128;
constexpr int T2_SLOTS = 2;
constexpr uint32_t T2_TILE_BYTES = 16384;
constexpr uint32_t T2_SLOT_BYTES = 2 * T2_TILE_BYTES;      // two bf16 terms of one 64-wide k-tile
constexpr int T2_ACC = 0, T2_A = 128, T2_A_STRIDE = 64;
constexpr int T2_NB = 8, T2_PS = 3 * T2_NB + 1, T2_DPP = 5;
constexpr int T2_BPAD = 28;           // last-layer bias: 25 parameters per dim padded to 7 float4
constexpr int T2_WIDE_MMA_DEFAULT = 1;   // BGX_T2_WIDE_MMA=0: lane-0-only issue loop (first version, A/B switch)
constexpr int T2_FAST_DEFAULT = 1;   // BGX_T2_FAST=0/1: see the kernel comment


struct T2Args {
  long long B;
  const float* cond;   // [B][K0raw] dense
  const float* tin;    // [B][D_t] dense
  float* tout;         // [B][D_t] dense
  int D_t, K0raw;
  DevMlp net;
  const uint16_t* wb[2][BGX_MAX_LAYERS];
  int ktiles[BGX_MAX_LAYERS];
  int npass, inverse;
  SplineK ck;
  int* oob;
  const float* dlogp_in;
  float* dlogp_out;
  int* status;
  long long ntiles;
  int bias_floats;
  int wide_mma;      // all lanes of the MMA warp run the issue loop, one elected lane issues
};

struct alignas(16) T2Smem {
  uint64_t full[T2_SLOTS];
  uint64_t x_ready;     // 8: layer-0 operand staged
  uint64_t a_ready;     // 8: hidden activations staged
  uint64_t acc_full;    // 1: a unit's accumulator is complete (also frees the unit's weight slots)
  uint64_t acc_empty;   // 8: chunk accumulator pulled into registers
  uint64_t y_full, c_full;   // 1 + tx
  uint64_t y_done, c_free;   // 8
  uint32_t tmem_base, pad;
  float dl_part[2][T2_TM];
};

// FAST: the last layer's bias in a 16-byte aligned [chunk][dim][28] layout (7 LDS.128 instead of 25
// loads per dim) and MUFU lg2 for the log-det.  (Tried and measured slower on the B200: evaluating two
// dims in one basic block for ILP; handing the accumulator back before evaluating — both cost more in
// registers, extra tcgen05.ld and instruction-cache misses than the shorter wait gains.)
// PACKED: the dims of a pass are evaluated two at a time in packed fp32 pairs (bgx_spline_reg2.cuh: FADD2 / FMUL2 /
// FFMA2, one issue slot for two dims) — the epilogue is issue-bound, so this is where the time goes.
template <bool INVERSE, int ACT, bool FAST, bool PACKED = false>
__global__ void __launch_bounds__(T2_THREADS, 2) spline_coupling_tc2_kernel(const T2Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array keeps the address space: LDS/STS instead of generic LD/ST)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = base;
  T2Smem* S = (T2Smem*)(base + T2_SLOTS * T2_SLOT_BYTES);
  float* bias_s = (float*)(S + 1);
  float* ybuf = bias_s + a.bias_floats;           // [128][D_t] dense
  float* cbuf = ybuf + T2_TM * a.D_t;             // [128][K0raw] dense (raw conditioner columns)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.net.n_layers;
  const int units_per_tile = (L - 1) + a.npass;
  const long long n_my = (a.ntiles > blockIdx.x) ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x == 0) {
    mbar_init(&S->full[0], 1);
    mbar_init(&S->full[1], 1);
    mbar_init(&S->x_ready, T2_EPI_WARPS);
    mbar_init(&S->a_ready, T2_EPI_WARPS);
    mbar_init(&S->acc_full, 1);
    mbar_init(&S->acc_empty, T2_EPI_WARPS);
    mbar_init(&S->y_full, 1);
    mbar_init(&S->c_full, 1);
    mbar_init(&S->y_done, T2_EPI_WARPS);
    mbar_init(&S->c_free, T2_EPI_WARPS);
    fence_mbar_init();
  }
  {
    int off = 0;
    for (int l = 0; l < L; ++l) {
      if (FAST && l == L - 1) break;
      for (int i = threadIdx.x; i < a.net.Np[l]; i += T2_THREADS) bias_s[off + i] = a.net.bias[l][i];
      off += a.net.Np[l];
    }
    if (FAST)      // last layer: [chunk][dim][28], zero padded
      for (int i = threadIdx.x; i < a.npass * T2_DPP * T2_BPAD; i += T2_THREADS) {
        const int c = i / (T2_DPP * T2_BPAD), r = i - c * (T2_DPP * T2_BPAD), d = r / T2_BPAD, k = r - d * T2_BPAD;
        bias_s[off + i] = (k < T2_PS) ? a.net.bias[L - 1][c * 128 + d * T2_PS + k] : 0.f;
      }
  }
  if (warp == 9) tmem_alloc<256>(&S->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S->tmem_base;

  if (warp == 8) {
    // ------------------------------------------------------------------ producer + tile I/O (one thread)
    if (lane == 0) {
      const uint32_t ybytes = (uint32_t)(T2_TM * a.D_t * 4), cbytes = (uint32_t)(T2_TM * a.K0raw * 4);
      auto tile_of = [&](long long it) { return blockIdx.x + it * (long long)gridDim.x; };
      auto rows_of = [&](long long it) { return (int)min((long long)T2_TM, a.B - tile_of(it) * T2_TM); };
      // I/O state machine (all tile indices are CTA-local): conditioner of tile k may be loaded once
      // tile k-1's staging released the buffer; y of tile k once tile k-1's output left the buffer.
      long long c_next = 0, y_store_next = 0, y_load_next = 0;
      uint32_t ph_cfree = 0, ph_ydone = 0;
      auto service_io = [&]() {
        if (c_next < n_my) {
          bool ok = (c_next == 0);
          if (!ok && mbar_try_wait(&S->c_free, ph_cfree)) { ph_cfree ^= 1; ok = true; }
          if (ok) {
            const uint32_t nb = (uint32_t)(rows_of(c_next) * a.K0raw * 4);
            mbar_expect_tx(&S->c_full, nb);
            bulk_g2s(cbuf, a.cond + tile_of(c_next) * T2_TM * (long long)a.K0raw, nb, &S->c_full);
            (void)cbytes;
            ++c_next;
          }
        }
        if (y_store_next < y_load_next && mbar_try_wait(&S->y_done, ph_ydone)) {
          ph_ydone ^= 1;
          const uint32_t nb = (uint32_t)(rows_of(y_store_next) * a.D_t * 4);
          bulk_s2g(a.tout + tile_of(y_store_next) * T2_TM * (long long)a.D_t, ybuf, nb);
          bulk_store_wait_read();
          ++y_store_next;
        }
        if (y_load_next < n_my && y_load_next == y_store_next) {
          const uint32_t nb = (uint32_t)(rows_of(y_load_next) * a.D_t * 4);
          mbar_expect_tx(&S->y_full, nb);
          bulk_g2s(ybuf, a.tin + tile_of(y_load_next) * T2_TM * (long long)a.D_t, nb, &S->y_full);
          (void)ybytes;
          ++y_load_next;
        }
      };
      int slot = 0;
      long long filled = 0, released = 0, events = 0;   // events = units observed complete
      uint32_t eph = 0;
      for (long long it = 0; it < n_my; ++it) {
        for (int l = 0; l < L; ++l) {
          const int nch = (l == L - 1) ? a.npass : 1;
          const int kt = a.ktiles[l];
          for (int c = 0; c < nch; ++c)
            for (int t = 0; t < kt; ++t) {
              while (filled - released >= T2_SLOTS) {
                service_io();
                if (mbar_try_wait(&S->acc_full, eph)) {
                  eph ^= 1;
                  const int u = (int)(events % units_per_tile);
                  released += a.ktiles[u < L - 1 ? u : L - 1];
                  ++events;
                } else if (a.status && *(volatile int*)a.status) {
                  break;                      // another role timed out: drain
                }
              }
              service_io();
              uint8_t* dst = ring + (size_t)slot * T2_SLOT_BYTES;
              mbar_expect_tx(&S->full[slot], T2_SLOT_BYTES);
              bulk_g2s(dst, a.wb[0][l] + ((long long)c * kt + t) * 8192, T2_TILE_BYTES, &S->full[slot]);
              bulk_g2s(dst + T2_TILE_BYTES, a.wb[1][l] + ((long long)c * kt + t) * 8192, T2_TILE_BYTES, &S->full[slot]);
              ++filled;
              slot ^= 1;
            }
        }
      }
      // drain the tile I/O of the last tiles
      for (uint32_t spin = 0; y_store_next < n_my && spin < (1u << 26); ++spin) {
        service_io();
        if (a.status && (spin & 0xfff) == 0xfff && *(volatile int*)a.status) break;
      }
      if (y_store_next < n_my && a.status) atomicExch(a.status, 1);
    }
    __syncwarp();
  } else if (warp == 9 && a.wide_mma) {
    // ------------------------------------------------------------------ MMA issuer, warp-wide (see bgx_tc.cuh)
    {
      const uint32_t idesc = idesc_bf16(128, 128);
      int slot = 0;
      uint32_t ph_full[2] = {0, 0};
      uint32_t ph_x = 0, ph_a = 0, ph_e = 0;
      bool first = true;
      for (long long it = 0; it < n_my; ++it) {
        for (int l = 0; l < L; ++l) {
          if (l == 0) { mbar_wait(&S->x_ready, ph_x, a.status); ph_x ^= 1; }
          else { mbar_wait(&S->a_ready, ph_a, a.status); ph_a ^= 1; }
          const bool last = (l == L - 1);
          const int nch = last ? a.npass : 1;
          const int kt = a.ktiles[l];
          const int ksteps_total = (a.net.K[l] + 15) / 16;
          for (int c = 0; c < nch; ++c) {
            // the single accumulator must have been drained: chunk c-1 (or the previous tile's last chunk)
            const bool need = last ? (c >= 1) : (l == 0 && !first);
            if (need) { mbar_wait(&S->acc_empty, ph_e, a.status); ph_e ^= 1; }
            tc_fence_after();
            uint32_t acc = 0;
            for (int t = 0; t < kt; ++t) {
              mbar_wait(&S->full[slot], ph_full[slot], a.status);
              ph_full[slot] ^= 1;
              const uint32_t b1 = smem_u32(ring + (size_t)slot * T2_SLOT_BYTES), b2 = b1 + T2_TILE_BYTES;
              slot ^= 1;
              tc_fence_after();
              const int nk = min(4, ksteps_total - t * 4);
              // descriptors of the k-steps differ only in the 16-byte-unit start address (+2 per step)
              const uint64_t d1 = smem_desc_sw128(b1), d2 = smem_desc_sw128(b2);
              const uint32_t a1 = tmem + T2_A + (uint32_t)(t * 32), a2 = a1 + T2_A_STRIDE;
              if (nk == 4) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  mma3_bf16x3_elect(tmem + T2_ACC, a1 + ks * 8, a2 + ks * 8, d1 + 2 * ks, d2 + 2 * ks, idesc,
                                    ks == 0 ? acc : 1u);
              } else {
                for (int ks = 0; ks < nk; ++ks)
                  mma3_bf16x3_elect(tmem + T2_ACC, a1 + ks * 8, a2 + ks * 8, d1 + 2 * ks, d2 + 2 * ks, idesc,
                                    ks == 0 ? acc : 1u);
              }
              acc = 1;
            }
            mma_commit_elect(&S->acc_full);
          }
        }
        first = false;
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(128, 128);
      int slot = 0;
      uint32_t ph_full[2] = {0, 0};
      uint32_t ph_x = 0, ph_a = 0, ph_e = 0;
      bool first = true;
      for (long long it = 0; it < n_my; ++it) {
        for (int l = 0; l < L; ++l) {
          if (l == 0) { mbar_wait(&S->x_ready, ph_x, a.status); ph_x ^= 1; }
          else { mbar_wait(&S->a_ready, ph_a, a.status); ph_a ^= 1; }
          const bool last = (l == L - 1);
          const int nch = last ? a.npass : 1;
          const int kt = a.ktiles[l];
          const int ksteps_total = (a.net.K[l] + 15) / 16;
          for (int c = 0; c < nch; ++c) {
            // the single accumulator must have been drained: chunk c-1 (or the previous tile's last chunk)
            const bool need = last ? (c >= 1) : (l == 0 && !first);
            if (need) { mbar_wait(&S->acc_empty, ph_e, a.status); ph_e ^= 1; }
            tc_fence_after();
            uint32_t acc = 0;
            for (int t = 0; t < kt; ++t) {
              mbar_wait(&S->full[slot], ph_full[slot], a.status);
              ph_full[slot] ^= 1;
              const uint32_t b1 = smem_u32(ring + (size_t)slot * T2_SLOT_BYTES), b2 = b1 + T2_TILE_BYTES;
              slot ^= 1;
              tc_fence_after();
              const int nk = min(4, ksteps_total - t * 4);
              for (int ks = 0; ks < nk; ++ks) {
                const uint32_t kcol = (uint32_t)(t * 32 + ks * 8);
                const uint32_t a1 = tmem + T2_A + kcol, a2 = a1 + T2_A_STRIDE;
                const uint64_t d1 = smem_desc_sw128(b1 + ks * 32), d2 = smem_desc_sw128(b2 + ks * 32);
                mma_bf16_ts(tmem + T2_ACC, a1, d2, idesc, acc);
                mma_bf16_ts(tmem + T2_ACC, a2, d1, idesc, 1);
                mma_bf16_ts(tmem + T2_ACC, a1, d1, idesc, 1);
                acc = 1;
              }
            }
            mma_commit(&S->acc_full);
          }
        }
        first = false;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (0..7)
    const int q = warp & 3, j = warp >> 2;            // quadrant, 0/1 within the quadrant
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t ph_acc = 0, ph_c = 0, ph_y = 0;
    int last_off = 0;
    for (int l = 0; l < L - 1; ++l) last_off += a.net.Np[l];
    const int K0 = a.net.K[0];

    auto cond_value = [&](const float* crow, int k) -> float {
      if (k >= K0) return 0.f;
      const int code = a.net.in_map[k];
      const float v = crow[code & 0xffffff];
      const int kind = code >> 24;
      if (kind == 0) return v;
      const float arg = (v - a.net.pleft) * a.net.pscale;
      return kind == 1 ? cosf(arg) : sinf(arg);
    };
    auto stage_x = [&]() {
      mbar_wait(&S->c_full, ph_c, a.status);
      ph_c ^= 1;
      const float* crow = cbuf + r_in_tile * a.K0raw;
      for (int b0 = j * 16; b0 < K0; b0 += 32) {
        uint32_t t1[8], t2[8], t3[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          split_bf16(cond_value(crow, b0 + 2 * i), cond_value(crow, b0 + 2 * i + 1), 2, t1[i], t2[i], t3[i]);
        const uint32_t col = tmem + lane_base + T2_A + b0 / 2;
        tmem_st8(col, t1);
        tmem_st8(col + T2_A_STRIDE, t2);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&S->x_ready);
        mbar_arrive(&S->c_free);
      }
    };

    for (long long it = 0; it < n_my; ++it) {
      const long long tile = blockIdx.x + it * (long long)gridDim.x;
      const long long row = tile * T2_TM + r_in_tile;
      float* yrow = ybuf + r_in_tile * a.D_t;
      if (it == 0) stage_x();
      // ---- hidden layers
      int boff = 0;
      for (int l = 0; l < L - 1; ++l) {
        mbar_wait(&S->acc_full, ph_acc, a.status);
        ph_acc ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int col = (j * 2 + h) * 32;
          uint32_t v[32];
          tmem_ld32(tmem + lane_base + T2_ACC + col, v);
          tmem_ld_wait();
          uint32_t t1[16], t2[16], t3[16];
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + boff + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i];
            if (PACKED) {
              hidden_pair2<ACT>(v[4 * i], v[4 * i + 1], bb.x, bb.y, t1[2 * i], t2[2 * i]);
              hidden_pair2<ACT>(v[4 * i + 2], v[4 * i + 3], bb.z, bb.w, t1[2 * i + 1], t2[2 * i + 1]);
            } else {
              const float h0 = act_fast<ACT>(__uint_as_float(v[4 * i]) + bb.x);
              const float h1 = act_fast<ACT>(__uint_as_float(v[4 * i + 1]) + bb.y);
              const float h2 = act_fast<ACT>(__uint_as_float(v[4 * i + 2]) + bb.z);
              const float h3 = act_fast<ACT>(__uint_as_float(v[4 * i + 3]) + bb.w);
              split_bf16(h0, h1, 2, t1[2 * i], t2[2 * i], t3[2 * i]);
              split_bf16(h2, h3, 2, t1[2 * i + 1], t2[2 * i + 1], t3[2 * i + 1]);
            }
          }
          const uint32_t acol = tmem + lane_base + T2_A + col / 2;
          tmem_st16(acol, t1);
          tmem_st16(acol + T2_A_STRIDE, t2);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->a_ready);
        boff += a.net.Np[l];
      }
      // ---- last layer: chunk c holds dims 5c..5c+4; this warp takes the dims d with d % 2 == j
      float ld = 0.f;
      int n_oob = 0;
      mbar_wait(&S->y_full, ph_y, a.status);
      ph_y ^= 1;
      for (int c = 0; c < a.npass; ++c) {
        mbar_wait(&S->acc_full, ph_acc, a.status);
        ph_acc ^= 1;
        tc_fence_after();
        const uint32_t acc_addr = tmem + lane_base + T2_ACC;
        const float* bl = bias_s + last_off + c * 128;
        const int i0 = (j + c) & 1;                           // (5c + i) % 2 == j  <=>  i % 2 == (j + c) % 2
        int n_mine = 0;
        for (int i = i0; i < T2_DPP; i += 2) n_mine += (5 * c + i < a.D_t) ? 1 : 0;
        const float* bv = bias_s + last_off + c * (T2_DPP * T2_BPAD);
        bool released = false;
        int m_first = 0;
        if (PACKED && FAST && n_mine >= 2) {
          // ---- dims i0 and i0 + 2 of this pass together, one per fp32 lane
          const int iA = i0, iB = i0 + 2;
          uint32_t va[25], vb[25];
          tmem_ld25(acc_addr + iA * T2_PS, va);
          tmem_ld25(acc_addr + iB * T2_PS, vb);
          tmem_ld_wait();
          if (n_mine == 2) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->acc_empty);
            released = true;
          }
          F2 p2[T2_PS];
          {
            const float4* bA = reinterpret_cast<const float4*>(bv + iA * T2_BPAD);
            const float4* bB = reinterpret_cast<const float4*>(bv + iB * T2_BPAD);
#pragma unroll
            for (int q = 0; q < 6; ++q) {
              const float4 x4 = bA[q], y4 = bB[q];
              p2[4 * q] = f2(__uint_as_float(va[4 * q]) + x4.x, __uint_as_float(vb[4 * q]) + y4.x);
              p2[4 * q + 1] = f2(__uint_as_float(va[4 * q + 1]) + x4.y, __uint_as_float(vb[4 * q + 1]) + y4.y);
              p2[4 * q + 2] = f2(__uint_as_float(va[4 * q + 2]) + x4.z, __uint_as_float(vb[4 * q + 2]) + y4.z);
              p2[4 * q + 3] = f2(__uint_as_float(va[4 * q + 3]) + x4.w, __uint_as_float(vb[4 * q + 3]) + y4.w);
            }
            p2[24] = f2(__uint_as_float(va[24]) + bv[iA * T2_BPAD + 24], __uint_as_float(vb[24]) + bv[iB * T2_BPAD + 24]);
          }
          float* ysA = yrow + 5 * c + iA;
          float xA = ysA[0], xB = ysA[2];
          n_oob += ((xA < a.ck.left || xA > a.ck.right) ? 1 : 0) + ((xB < a.ck.left || xB > a.ck.right) ? 1 : 0);
          xA = fminf(fmaxf(xA, a.ck.left), a.ck.right);
          xB = fminf(fmaxf(xB, a.ck.left), a.ck.right);
          F2 y2, l2;
          rqs_eval_reg2<!INVERSE>(p2, a.ck, f2(xA, xB), y2, l2);
          ysA[0] = lo(y2);
          ysA[2] = hi(y2);
          ld += lo(l2);
          ld += hi(l2);
          m_first = 2;
        }
        for (int m = m_first; m < n_mine; ++m) {
          const int i = i0 + 2 * m;
          uint32_t v[32];
          tmem_ld32(acc_addr + i * T2_PS, v);
          tmem_ld_wait();
          if (m == n_mine - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->acc_empty);
            released = true;
          }
          float p[T2_PS];
          if (FAST) {
            const float4* b4 = reinterpret_cast<const float4*>(bv + i * T2_BPAD);
#pragma unroll
            for (int q = 0; q < 6; ++q) {
              const float4 bb = b4[q];
              p[4 * q] = __uint_as_float(v[4 * q]) + bb.x;
              p[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + bb.y;
              p[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + bb.z;
              p[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + bb.w;
            }
            p[24] = __uint_as_float(v[24]) + bv[i * T2_BPAD + 24];
          } else {
#pragma unroll
            for (int k = 0; k < T2_PS; ++k) p[k] = __uint_as_float(v[k]) + bl[i * T2_PS + k];
          }
          float* ys = yrow + 5 * c + i;
          float x = *ys;
          n_oob += (x < a.ck.left || x > a.ck.right) ? 1 : 0;
          x = fminf(fmaxf(x, a.ck.left), a.ck.right);
          float y, lad;
          rqs_eval_reg<!INVERSE, FAST>(p, a.ck, x, y, lad);
          *ys = y;
          ld += lad;
        }
        if (!released) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S->acc_empty);
        }
        if (c == a.npass - 1 && it + 1 < n_my) stage_x();   // every MMA of this tile is complete
      }
      if (n_oob && a.oob) atomicAdd(a.oob, n_oob);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->y_done);
      S->dl_part[j][r_in_tile] = ld;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (j == 0 && row < a.B) {
        const float base_dl = a.dlogp_in ? a.dlogp_in[row] : 0.f;
        a.dlogp_out[row] = base_dl + (S->dl_part[0][r_in_tile] + S->dl_part[1][r_in_tile]);
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

// ------------------------------------------------------------------------------ host side

bool spline_tc2_eligible(const bgx_coupling_io* io, const bgx_packed_mlp* net, const bgx_spline_cfg* cfg, int flags) {
  if (!io || !net || !cfg || cfg->n_bins != T2_NB || (flags & BGX_FLAG_BF16X6)) return false;
  const char* e = getenv("BGX_TC_SINGLE_CTA");
  if (e && atoi(e)) return false;
  const int L = net->n_layers;
  if (L < 2 || L > 6 || net->K[0] > 128) return false;
  for (int l = 0; l + 1 < L; ++l)
    if (net->N[l] != 128) return false;
  for (int l = 0; l < L; ++l)
    if (!net->Wb[0][l] || !net->Wb[1][l]) return false;
  if (net->spline_dims_per_pass != T2_DPP || net->spline_stride != T2_PS) return false;
  if (io->n_cond != 1 || io->n_tr != 1) return false;
  if (io->batch % 4) return false;            // a partial last tile must still be a multiple of 16 bytes
  auto dense = [](const bgx_seg& s) { return s.stride == s.width && ((uintptr_t)s.ptr & 15) == 0; };
  if (!dense(io->cond[0]) || !dense(io->tr_in[0]) || !dense(io->tr_out[0])) return false;
  if (io->cond[0].width != net->raw_width) return false;
  size_t bias = 0;
  for (int l = 0; l < L; ++l) bias += (l == L - 1) ? (size_t)(net->Np[l] / 128) * T2_DPP * T2_BPAD : net->Np[l];
  const size_t need = 1024 + T2_SLOTS * T2_SLOT_BYTES + sizeof(T2Smem) +
                      4 * (bias + (size_t)T2_TM * (io->tr_in[0].width + io->cond[0].width)) + 64;
  return need <= 112 * 1024;
}

int spline_coupling_tc2(const bgx_coupling_io* io, const bgx_packed_mlp* net, const bgx_spline_cfg* cfg, int flags,
                        int* status, cudaStream_t st) {
  const int L = net->n_layers;
  const int d_t = io->tr_in[0].width;
  if (net->N[L - 1] != ceil_div(d_t, T2_DPP) * 128 || io->tr_out[0].width != d_t || !io->dlogp_out) return BGX_ERR_INVALID;
  if (io->batch == 0) return BGX_OK;
  T2Args a{};
  a.B = io->batch;
  a.cond = io->cond[0].ptr; a.tin = io->tr_in[0].ptr; a.tout = const_cast<float*>(io->tr_out[0].ptr);
  a.D_t = d_t; a.K0raw = io->cond[0].width;
  mlp_to_dev(net, a.net);
  int bias_floats = 0;
  for (int l = 0; l < L; ++l) {
    a.wb[0][l] = (const uint16_t*)net->Wb[0][l];
    a.wb[1][l] = (const uint16_t*)net->Wb[1][l];
    a.ktiles[l] = ceil_div(net->K[l], 64);
    bias_floats += (l == L - 1) ? (net->Np[l] / 128) * T2_DPP * T2_BPAD : net->Np[l];   // room for the padded layout
  }
  a.bias_floats = bias_floats;
  {
    static const int wide = [] { const char* e = getenv("BGX_T2_WIDE_MMA"); return e ? atoi(e) : T2_WIDE_MMA_DEFAULT; }();
    a.wide_mma = wide;
  }
  a.npass = net->N[L - 1] / 128;
  a.inverse = (flags & BGX_FLAG_INVERSE) ? 1 : 0;
  SplineParams sp;
  spline_params_from_cfg(cfg, sp);
  {
    const float wx = sp.right - sp.left, hy = sp.top - sp.bottom;
    a.ck.left = sp.left; a.ck.right = sp.right; a.ck.bottom = sp.bottom; a.ck.top = sp.top;
    a.ck.wscale = wx * (1.f - sp.min_w * T2_NB); a.ck.hscale = hy * (1.f - sp.min_h * T2_NB);
    a.ck.wstep = wx * sp.min_w; a.ck.hstep = hy * sp.min_h;
    a.ck.min_d = sp.min_d; a.ck.beta = sp.beta; a.ck.beta_l2e = sp.beta * LOG2E;
    a.ck.ln2_over_beta = LN2 * sp.inv_beta;
  }
  a.oob = sp.oob;
  a.dlogp_in = io->dlogp_in;
  a.dlogp_out = io->dlogp_out;
  a.status = status;
  a.ntiles = (a.B + T2_TM - 1) / T2_TM;
  const size_t smem = 1024 + T2_SLOTS * T2_SLOT_BYTES + sizeof(T2Smem) +
                      sizeof(float) * ((size_t)bias_floats + (size_t)T2_TM * (a.D_t + a.K0raw)) + 64;
  int sm_count = 0;
  int rc = device_sm_count(&sm_count);
  if (rc) return rc;
  using KernT = void (*)(const T2Args);
#define BGX_T2_ROW(INV, FAST, PK) \
  {spline_coupling_tc2_kernel<INV, 0, FAST, PK>, spline_coupling_tc2_kernel<INV, 1, FAST, PK>, \
   spline_coupling_tc2_kernel<INV, 2, FAST, PK>, spline_coupling_tc2_kernel<INV, 3, FAST, PK>}
  static const KernT kerns[3][2][4] = {{BGX_T2_ROW(false, false, false), BGX_T2_ROW(true, false, false)},
                                       {BGX_T2_ROW(false, true, false), BGX_T2_ROW(true, true, false)},
                                       {BGX_T2_ROW(false, true, true), BGX_T2_ROW(true, true, true)}};
#undef BGX_T2_ROW
  if (net->act < 0 || net->act > 3) return BGX_ERR_INVALID;
  // BGX_T2_FAST=0: first epilogue (A/B switch); BGX_T2_PACKED=0: scalar spline evaluation (A/B switch)
  static const int pair = [] {
    const char* e = getenv("BGX_T2_FAST");
    const int fast = e ? (atoi(e) ? 1 : 0) : T2_FAST_DEFAULT;
    const char* pk = getenv("BGX_T2_PACKED");
    const int packed = pk ? (atoi(pk) ? 1 : 0) : 0;   // measured on the B200: 0.898 ms packed vs 0.859 ms scalar per launch here
                                                      // (extra register moves + FFMA2 latency with only 8 warps per CTA); the
                                                      // pair kernel, which releases its accumulators early, keeps the packed form
    return fast ? (packed ? 2 : 1) : 0;
  }();
  KernT kern = kerns[pair][a.inverse][net->act];
  static size_t configured_all[BGX_MAX_DEVICES][3][2][4] = {};
  auto& configured = configured_all[device_slot()];
  if (smem > configured[pair][a.inverse][net->act]) {
    rc = check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
    rc = check(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (rc) return rc;
    configured[pair][a.inverse][net->act] = smem;
  }
  const unsigned grid = (unsigned)std::min<long long>(a.ntiles, 2LL * sm_count);
  kern<<<grid, T2_THREADS, smem, st>>>(a);
  return post_launch();
}

}  // namespace bgx
