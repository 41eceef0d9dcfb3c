// EXPERIMENTAL — NOT PART OF THE BUILD, NEVER RUN ON HARDWARE YET (see experimental/README.md).
//
// Tensor-core fused RQ-spline coupling block, two CTAs per SM, with the LAST layer issued in
// 64-column half-chunks that are DOUBLE-BUFFERED inside the CTA's 128 accumulator columns.
//
// Why: in spline_coupling_tc2_kernel a CTA owns ONE 128-column accumulator, so its eight epilogue
// warps idle while its own MMAs run and its MMA warp idles while they evaluate; only the sibling
// CTA fills the gaps (ncu source view: 24 % of all warp samples on the acc_full wait).  Here a
// half-chunk holds exactly TWO transformed dims (2 x 25 of 64 columns), the two warps of a TMEM
// lane quadrant take one dim each, pull it (32-column tcgen05.ld) and hand the half back BEFORE
// evaluating it, while the MMAs of the next half fill the other 64 columns: epilogue and tensor
// pipe of the same CTA overlap, and the dims are perfectly balanced between the two warps.
// Price: 14 of 64 columns are padding (17 half-chunks instead of 6.6 chunks for D_t = 33: +21 % MMA
// work in the last layer if an N = 64 MMA costs half an N = 128 one) and twice as many commits.
// The MMA issue loop is warp-wide (7.6 SASS instructions per MMA), which is what makes twice as
// many N = 64 instructions affordable.
//
// Needs the last layer packed as [chunk of 128][half of 64][dim of 25 | dim of 25 | 14 zero columns]
// (bgx_pack.cu: experimental/README.md shows the last_map change) and the bias of the last layer
// in [half][2][28] floats.
//
//   warps 0-7  epilogue (thread <-> sample row; warp j of quadrant q takes dim 2h + j of half h)
//   warp  8    lane 0: weight producer (bulk TMA, 2-slot ring: one slot = one k-tile of a hidden
//              layer, or BOTH k-tiles of a half-chunk) + tile I/O
//   warp  9    warp-wide tcgen05.mma issue; owns the 256-column TMEM allocation
//
// TMEM columns: [0,64) ACC half 0, [64,128) ACC half 1 (hidden layers use [0,128) as one N = 128
// accumulator), [128,192) A1, [192,256) A2.
#include <cstdlib>

#include "bgx_coupling.cuh"
#include "bgx_tc.cuh"
#include "bgx_tc_epi.cuh"
#include "bgx_spline_reg.cuh"

namespace bgx {
using namespace tc;

constexpr int T3_THREADS = 320;
constexpr int T3_EPI_WARPS = 8;
constexpr int T3_TM = 128;
constexpr int T3_SLOTS = 2;
constexpr uint32_t T3_TILE_BYTES = 16384;                  // [128 rows][64 k] bf16
constexpr uint32_t T3_HALF_BYTES = 8192;                   // [64 rows][64 k] bf16
constexpr uint32_t T3_SLOT_BYTES = 2 * T3_TILE_BYTES;
constexpr int T3_ACC = 0, T3_A = 128, T3_A_STRIDE = 64, T3_HALF = 64;
constexpr int T3_NB = 8, T3_PS = 3 * T3_NB + 1, T3_DPH = 2;
constexpr int T3_BPAD = 28;

struct T3Args {
  long long B;
  const float* cond;   // [B][K0raw] dense
  const float* tin;    // [B][D_t] dense
  float* tout;         // [B][D_t] dense
  int D_t, K0raw;
  DevMlp net;
  const uint16_t* wb[2][BGX_MAX_LAYERS];
  int ktiles[BGX_MAX_LAYERS];
  int nhalf, inverse;  // nhalf = ceil(D_t / 2) half-chunks of the last layer
  SplineK ck;
  int* oob;
  const float* dlogp_in;
  float* dlogp_out;
  int* status;
  long long ntiles;
  int bias_floats;
};

struct alignas(16) T3Smem {
  uint64_t full[T3_SLOTS];
  uint64_t x_ready;      // 8: layer-0 operand staged
  uint64_t a_ready;      // 8: hidden activations staged
  uint64_t hfull;        // 1: a hidden layer's (N = 128) accumulator is complete
  uint64_t afull[2];     // 1: half-chunk accumulator b is complete
  uint64_t aempty[2];    // 8: half-chunk accumulator b was pulled into registers
  uint64_t y_full, c_full;   // 1 + tx
  uint64_t y_done, c_free;   // 8
  uint32_t tmem_base, pad;
  float dl_part[2][T3_TM];
};

template <bool INVERSE, int ACT>
__global__ void __launch_bounds__(T3_THREADS, 2) spline_coupling_tc3_kernel(const T3Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = base;
  T3Smem* S = (T3Smem*)(base + T3_SLOTS * T3_SLOT_BYTES);
  float* bias_s = (float*)(S + 1);
  float* ybuf = bias_s + a.bias_floats;           // [128][D_t] dense
  float* cbuf = ybuf + T3_TM * a.D_t;             // [128][K0raw] dense (raw conditioner columns)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.net.n_layers;
  const int units_per_tile = (L - 1) + a.nhalf;
  const long long n_my = (a.ntiles > blockIdx.x) ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x == 0) {
    mbar_init(&S->full[0], 1);
    mbar_init(&S->full[1], 1);
    mbar_init(&S->x_ready, T3_EPI_WARPS);
    mbar_init(&S->a_ready, T3_EPI_WARPS);
    mbar_init(&S->hfull, 1);
    mbar_init(&S->afull[0], 1);
    mbar_init(&S->afull[1], 1);
    mbar_init(&S->aempty[0], T3_EPI_WARPS);
    mbar_init(&S->aempty[1], T3_EPI_WARPS);
    mbar_init(&S->y_full, 1);
    mbar_init(&S->c_full, 1);
    mbar_init(&S->y_done, T3_EPI_WARPS);
    mbar_init(&S->c_free, T3_EPI_WARPS);
    fence_mbar_init();
  }
  {
    int off = 0;
    for (int l = 0; l < L - 1; ++l) {
      for (int i = threadIdx.x; i < a.net.Np[l]; i += T3_THREADS) bias_s[off + i] = a.net.bias[l][i];
      off += a.net.Np[l];
    }
    // last layer: [half][dim 0 | dim 1][28], zero padded; packed column of (half h, dim slot d, k) =
    // (h / 2) * 128 + (h & 1) * 64 + d * 25 + k
    for (int i = threadIdx.x; i < a.nhalf * T3_DPH * T3_BPAD; i += T3_THREADS) {
      const int h = i / (T3_DPH * T3_BPAD), r = i - h * (T3_DPH * T3_BPAD), d = r / T3_BPAD, k = r - d * T3_BPAD;
      bias_s[off + i] = (k < T3_PS) ? a.net.bias[L - 1][(h >> 1) * 128 + (h & 1) * T3_HALF + d * T3_PS + k] : 0.f;
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
      auto tile_of = [&](long long it) { return blockIdx.x + it * (long long)gridDim.x; };
      auto rows_of = [&](long long it) { return (int)min((long long)T3_TM, a.B - tile_of(it) * T3_TM); };
      long long c_next = 0, y_store_next = 0, y_load_next = 0;
      uint32_t ph_cfree = 0, ph_ydone = 0;
      auto service_io = [&]() {
        if (c_next < n_my) {
          bool ok = (c_next == 0);
          if (!ok && mbar_try_wait(&S->c_free, ph_cfree)) { ph_cfree ^= 1; ok = true; }
          if (ok) {
            const uint32_t nb = (uint32_t)(rows_of(c_next) * a.K0raw * 4);
            mbar_expect_tx(&S->c_full, nb);
            bulk_g2s(cbuf, a.cond + tile_of(c_next) * T3_TM * (long long)a.K0raw, nb, &S->c_full);
            ++c_next;
          }
        }
        if (y_store_next < y_load_next && mbar_try_wait(&S->y_done, ph_ydone)) {
          ph_ydone ^= 1;
          const uint32_t nb = (uint32_t)(rows_of(y_store_next) * a.D_t * 4);
          bulk_s2g(a.tout + tile_of(y_store_next) * T3_TM * (long long)a.D_t, ybuf, nb);
          bulk_store_wait_read();
          ++y_store_next;
        }
        if (y_load_next < n_my && y_load_next == y_store_next) {
          const uint32_t nb = (uint32_t)(rows_of(y_load_next) * a.D_t * 4);
          mbar_expect_tx(&S->y_full, nb);
          bulk_g2s(ybuf, a.tin + tile_of(y_load_next) * T3_TM * (long long)a.D_t, nb, &S->y_full);
          ++y_load_next;
        }
      };
      // Slot accounting.  Units complete in issue order; unit u of a tile is hidden layer u (u < L - 1:
      // ktiles[u] slots, signalled by hfull) or half-chunk u - (L - 1) (one slot, signalled by afull[b]).
      int slot = 0;
      long long filled = 0, released = 0, events = 0;
      uint32_t ph_h = 0, ph_a[2] = {0, 0};
      auto wait_slot = [&]() {
        while (filled - released >= T3_SLOTS) {
          service_io();
          const int u = (int)(events % units_per_tile);
          bool done;
          if (u < L - 1) {
            done = mbar_try_wait(&S->hfull, ph_h);
            if (done) { ph_h ^= 1; released += a.ktiles[u]; }
          } else {
            const int b = (u - (L - 1)) & 1;
            done = mbar_try_wait(&S->afull[b], ph_a[b]);
            if (done) { ph_a[b] ^= 1; released += 1; }
          }
          if (done) ++events;
          else if (a.status && *(volatile int*)a.status) break;     // another role timed out: drain
        }
        service_io();
      };
      for (long long it = 0; it < n_my; ++it) {
        for (int l = 0; l < L - 1; ++l)
          for (int t = 0; t < a.ktiles[l]; ++t) {
            wait_slot();
            uint8_t* dst = ring + (size_t)slot * T3_SLOT_BYTES;
            mbar_expect_tx(&S->full[slot], T3_SLOT_BYTES);
            bulk_g2s(dst, a.wb[0][l] + (long long)t * 8192, T3_TILE_BYTES, &S->full[slot]);
            bulk_g2s(dst + T3_TILE_BYTES, a.wb[1][l] + (long long)t * 8192, T3_TILE_BYTES, &S->full[slot]);
            ++filled;
            slot ^= 1;
          }
        const int kt = a.ktiles[L - 1];       // == 2 (hidden width 128): both k-tiles of a half fit one slot
        for (int h = 0; h < a.nhalf; ++h) {
          wait_slot();
          uint8_t* dst = ring + (size_t)slot * T3_SLOT_BYTES;
          mbar_expect_tx(&S->full[slot], (uint32_t)(kt * 2) * T3_HALF_BYTES);
          for (int t = 0; t < kt; ++t) {
            // rows (h & 1) * 64 .. + 64 of the 128-row tile of chunk h / 2, k-tile t: 8 KB, contiguous
            const long long src = ((long long)(h >> 1) * kt + t) * 8192 + (long long)(h & 1) * 4096;
            bulk_g2s(dst + (2 * t) * T3_HALF_BYTES, a.wb[0][L - 1] + src, T3_HALF_BYTES, &S->full[slot]);
            bulk_g2s(dst + (2 * t + 1) * T3_HALF_BYTES, a.wb[1][L - 1] + src, T3_HALF_BYTES, &S->full[slot]);
          }
          ++filled;
          slot ^= 1;
        }
      }
      for (uint32_t spin = 0; y_store_next < n_my && spin < (1u << 26); ++spin) {
        service_io();
        if (a.status && (spin & 0xfff) == 0xfff && *(volatile int*)a.status) break;
      }
      if (y_store_next < n_my && a.status) atomicExch(a.status, 1);
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer, warp-wide
    const uint32_t idesc128 = idesc_bf16(128, 128), idesc64 = idesc_bf16(128, T3_HALF);
    int slot = 0;
    uint32_t ph_full[2] = {0, 0};
    uint32_t ph_x = 0, ph_a = 0, ph_e[2] = {0, 0};
    int outstanding[2] = {0, 0};      // half-chunks issued into buffer b whose aempty was not consumed yet
    auto drain = [&](int b) {         // buffer b may be overwritten once its last half was pulled
      if (outstanding[b]) {
        mbar_wait(&S->aempty[b], ph_e[b], a.status);
        ph_e[b] ^= 1;
        outstanding[b] = 0;
      }
    };
    for (long long it = 0; it < n_my; ++it) {
      for (int l = 0; l < L - 1; ++l) {
        if (l == 0) { mbar_wait(&S->x_ready, ph_x, a.status); ph_x ^= 1; }
        else { mbar_wait(&S->a_ready, ph_a, a.status); ph_a ^= 1; }
        if (l == 0) { drain(0); drain(1); }          // hidden layers accumulate into all 128 columns
        // (l > 0: the previous hidden accumulator was consumed before a_ready was signalled)
        tc_fence_after();
        const int kt = a.ktiles[l];
        const int ksteps_total = (a.net.K[l] + 15) / 16;
        uint32_t acc = 0;
        for (int t = 0; t < kt; ++t) {
          mbar_wait(&S->full[slot], ph_full[slot], a.status);
          ph_full[slot] ^= 1;
          const uint32_t b1 = smem_u32(ring + (size_t)slot * T3_SLOT_BYTES), b2 = b1 + T3_TILE_BYTES;
          slot ^= 1;
          tc_fence_after();
          const int nk = min(4, ksteps_total - t * 4);
          const uint64_t d1 = smem_desc_sw128(b1), d2 = smem_desc_sw128(b2);
          const uint32_t a1 = tmem + T3_A + (uint32_t)(t * 32), a2 = a1 + T3_A_STRIDE;
          for (int ks = 0; ks < nk; ++ks)
            mma3_bf16x3_elect(tmem + T3_ACC, a1 + ks * 8, a2 + ks * 8, d1 + 2 * ks, d2 + 2 * ks, idesc128,
                              ks == 0 ? acc : 1u);
          acc = 1;
        }
        mma_commit_elect(&S->hfull);
      }
      // ---- last layer: half-chunks, alternating accumulator halves
      mbar_wait(&S->a_ready, ph_a, a.status);
      ph_a ^= 1;
      const int kt = a.ktiles[L - 1];
      for (int h = 0; h < a.nhalf; ++h) {
        const int b = h & 1;
        drain(b);                       // half h - 2 (same buffer) pulled?  (h < 2: nothing outstanding)
        mbar_wait(&S->full[slot], ph_full[slot], a.status);
        ph_full[slot] ^= 1;
        const uint32_t sb = smem_u32(ring + (size_t)slot * T3_SLOT_BYTES);
        slot ^= 1;
        tc_fence_after();
        for (int t = 0; t < kt; ++t) {
          const uint64_t d1 = smem_desc_sw128(sb + (2 * t) * T3_HALF_BYTES);
          const uint64_t d2 = smem_desc_sw128(sb + (2 * t + 1) * T3_HALF_BYTES);
          const uint32_t a1 = tmem + T3_A + (uint32_t)(t * 32), a2 = a1 + T3_A_STRIDE;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)          // hidden width 128: 2 k-tiles x 4 k-steps
            mma3_bf16x3_elect(tmem + T3_ACC + b * T3_HALF, a1 + ks * 8, a2 + ks * 8, d1 + 2 * ks, d2 + 2 * ks, idesc64,
                              (t | ks) == 0 ? 0u : 1u);
        }
        mma_commit_elect(&S->afull[b]);
        outstanding[b] = 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (0..7)
    const int q = warp & 3, j = warp >> 2;            // quadrant, 0/1 within the quadrant
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t ph_h = 0, ph_f[2] = {0, 0}, ph_c = 0, ph_y = 0;
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
        const uint32_t col = tmem + lane_base + T3_A + b0 / 2;
        tmem_st8(col, t1);
        tmem_st8(col + T3_A_STRIDE, t2);
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
      const long long row = tile * T3_TM + r_in_tile;
      float* yrow = ybuf + r_in_tile * a.D_t;
      if (it == 0) stage_x();
      // ---- hidden layers (as in the tc2 kernel)
      int boff = 0;
      for (int l = 0; l < L - 1; ++l) {
        mbar_wait(&S->hfull, ph_h, a.status);
        ph_h ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int col = (j * 2 + h) * 32;
          uint32_t v[32];
          tmem_ld32(tmem + lane_base + T3_ACC + col, v);
          tmem_ld_wait();
          uint32_t t1[16], t2[16], t3[16];
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + boff + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i];
            const float h0 = act_fast<ACT>(__uint_as_float(v[4 * i]) + bb.x);
            const float h1 = act_fast<ACT>(__uint_as_float(v[4 * i + 1]) + bb.y);
            const float h2 = act_fast<ACT>(__uint_as_float(v[4 * i + 2]) + bb.z);
            const float h3 = act_fast<ACT>(__uint_as_float(v[4 * i + 3]) + bb.w);
            split_bf16(h0, h1, 2, t1[2 * i], t2[2 * i], t3[2 * i]);
            split_bf16(h2, h3, 2, t1[2 * i + 1], t2[2 * i + 1], t3[2 * i + 1]);
          }
          const uint32_t acol = tmem + lane_base + T3_A + col / 2;
          tmem_st16(acol, t1);
          tmem_st16(acol + T3_A_STRIDE, t2);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->a_ready);
        boff += a.net.Np[l];
      }
      // ---- last layer: half h holds dims 2h and 2h + 1; this warp takes dim 2h + j
      float ld = 0.f;
      int n_oob = 0;
      mbar_wait(&S->y_full, ph_y, a.status);
      ph_y ^= 1;
      for (int h = 0; h < a.nhalf; ++h) {
        const int b = h & 1;
        mbar_wait(&S->afull[b], ph_f[b], a.status);
        ph_f[b] ^= 1;
        tc_fence_after();
        const int d = 2 * h + j;
        const bool mine = d < a.D_t;
        uint32_t v[32];
        if (mine) {
          tmem_ld32(tmem + lane_base + T3_ACC + b * T3_HALF + j * T3_PS, v);   // columns j*25 .. +31 < 64
          tmem_ld_wait();
        }
        tc_fence_before();                        // hand the half back BEFORE evaluating it
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->aempty[b]);
        if (mine) {
          const float* bv = bias_s + last_off + (h * T3_DPH + j) * T3_BPAD;
          const float4* b4 = reinterpret_cast<const float4*>(bv);
          float p[T3_PS];
#pragma unroll
          for (int qq = 0; qq < 6; ++qq) {
            const float4 bb = b4[qq];
            p[4 * qq] = __uint_as_float(v[4 * qq]) + bb.x;
            p[4 * qq + 1] = __uint_as_float(v[4 * qq + 1]) + bb.y;
            p[4 * qq + 2] = __uint_as_float(v[4 * qq + 2]) + bb.z;
            p[4 * qq + 3] = __uint_as_float(v[4 * qq + 3]) + bb.w;
          }
          p[24] = __uint_as_float(v[24]) + bv[24];
          float* ys = yrow + d;
          float x = *ys;
          n_oob += (x < a.ck.left || x > a.ck.right) ? 1 : 0;
          x = fminf(fmaxf(x, a.ck.left), a.ck.right);
          float y, lad;
          rqs_eval_reg<!INVERSE, true>(p, a.ck, x, y, lad);
          *ys = y;
          ld += lad;
        }
        if (h == a.nhalf - 1 && it + 1 < n_my) stage_x();   // every MMA of this tile is complete
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

// `net` must have been packed with the half-chunk layout (4 dims per 128 columns as 2 + 2, see README)
int spline_coupling_tc3(const bgx_coupling_io* io, const bgx_packed_mlp* net, const bgx_spline_cfg* cfg, int flags,
                        int* status, cudaStream_t st) {
  const int L = net->n_layers;
  const int d_t = io->tr_in[0].width;
  if (net->N[L - 1] != ceil_div(d_t, 2 * T3_DPH) * 128 || io->tr_out[0].width != d_t || !io->dlogp_out) return BGX_ERR_INVALID;
  if (net->K[L - 1] != 128) return BGX_ERR_UNSUPPORTED;      // one slot holds exactly the two k-tiles of a half
  if (io->batch == 0) return BGX_OK;
  T3Args a{};
  a.B = io->batch;
  a.cond = io->cond[0].ptr; a.tin = io->tr_in[0].ptr; a.tout = const_cast<float*>(io->tr_out[0].ptr);
  a.D_t = d_t; a.K0raw = io->cond[0].width;
  mlp_to_dev(net, a.net);
  a.nhalf = ceil_div(d_t, T3_DPH);
  int bias_floats = 0;
  for (int l = 0; l < L; ++l) {
    a.wb[0][l] = (const uint16_t*)net->Wb[0][l];
    a.wb[1][l] = (const uint16_t*)net->Wb[1][l];
    a.ktiles[l] = ceil_div(net->K[l], 64);
    bias_floats += (l == L - 1) ? a.nhalf * T3_DPH * T3_BPAD : net->Np[l];
  }
  a.bias_floats = bias_floats;
  a.inverse = (flags & BGX_FLAG_INVERSE) ? 1 : 0;
  SplineParams sp;
  spline_params_from_cfg(cfg, sp);
  {
    const float wx = sp.right - sp.left, hy = sp.top - sp.bottom;
    a.ck.left = sp.left; a.ck.right = sp.right; a.ck.bottom = sp.bottom; a.ck.top = sp.top;
    a.ck.wscale = wx * (1.f - sp.min_w * T3_NB); a.ck.hscale = hy * (1.f - sp.min_h * T3_NB);
    a.ck.wstep = wx * sp.min_w; a.ck.hstep = hy * sp.min_h;
    a.ck.min_d = sp.min_d; a.ck.beta = sp.beta; a.ck.beta_l2e = sp.beta * LOG2E;
    a.ck.ln2_over_beta = LN2 * sp.inv_beta;
  }
  a.oob = sp.oob;
  a.dlogp_in = io->dlogp_in;
  a.dlogp_out = io->dlogp_out;
  a.status = status;
  a.ntiles = (a.B + T3_TM - 1) / T3_TM;
  const size_t smem = 1024 + T3_SLOTS * T3_SLOT_BYTES + sizeof(T3Smem) +
                      sizeof(float) * ((size_t)bias_floats + (size_t)T3_TM * (a.D_t + a.K0raw)) + 64;
  static int sm_count = 0;
  int rc;
  if (!sm_count) {
    int dev = 0;
    rc = check(cudaGetDevice(&dev));
    if (rc) return rc;
    rc = check(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (rc) return rc;
  }
  using KernT = void (*)(const T3Args);
  static const KernT kerns[2][4] = {
      {spline_coupling_tc3_kernel<false, 0>, spline_coupling_tc3_kernel<false, 1>, spline_coupling_tc3_kernel<false, 2>,
       spline_coupling_tc3_kernel<false, 3>},
      {spline_coupling_tc3_kernel<true, 0>, spline_coupling_tc3_kernel<true, 1>, spline_coupling_tc3_kernel<true, 2>,
       spline_coupling_tc3_kernel<true, 3>}};
  if (net->act < 0 || net->act > 3) return BGX_ERR_INVALID;
  KernT kern = kerns[a.inverse][net->act];
  static size_t configured[2][4] = {};
  if (smem > configured[a.inverse][net->act]) {
    rc = check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
    rc = check(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (rc) return rc;
    configured[a.inverse][net->act] = smem;
  }
  const unsigned grid = (unsigned)std::min<long long>(a.ntiles, 2LL * sm_count);
  kern<<<grid, T3_THREADS, smem, st>>>(a);
  return post_launch();
}

}  // namespace bgx
