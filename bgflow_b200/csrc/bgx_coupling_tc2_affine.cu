// Tensor-core fused AFFINE coupling block, TWO CTAs PER SM variant (see bgx_coupling_tc2.cu for the
// resource split and bgx_coupling_tc_affine.cu for the program).  With a single 128-column
// accumulator per CTA, mu (the shift net's output) cannot wait in tensor memory while the scale
// net runs: it is parked in the conditioner tile's shared-memory buffer, which is free once the
// scale net's first-layer operand has been staged (requires D_t <= conditioner width).
//
//   net 0: x -> hidden (ACC) -> mu (ACC, N = round16(D_t)) -> +bias -> shared memory
//   net 1: x -> hidden (ACC) -> s  (ACC)                    -> y' = y * exp(tanh(s) alpha) + mu
#include <cstdlib>
#include <type_traits>

#include "bgx_coupling.cuh"
#include "bgx_tc.cuh"
#include "bgx_tc_epi.cuh"

namespace bgx {
using namespace tc;

constexpr int A2_THREADS = 320;
constexpr int A2_EPI_WARPS = 8;
constexpr int A2_TM = 128;
constexpr int A2_SLOTS = 2;
constexpr uint32_t A2_TILE_BYTES = 16384;
constexpr uint32_t A2_SLOT_BYTES = 2 * A2_TILE_BYTES;
constexpr int A2_ACC = 0, A2_A = 128, A2_A_STRIDE = 64;

struct A2Args {
  long long B;
  const float* cond;
  const float* tin;
  float* tout;
  int D_t, K0raw, nfin;
  DevMlp net[2];
  const uint16_t* wb[2][2][BGX_MAX_LAYERS];   // [net][term][layer]
  int ktiles[BGX_MAX_LAYERS];
  int L, inverse;
  float alpha;
  const float* dlogp_in;
  float* dlogp_out;
  int* status;
  long long ntiles;
  int wide_mma;      // all lanes of the MMA warp run the issue loop, one elected lane issues
  int bias_floats;     // per net
};

struct alignas(16) A2Smem {
  uint64_t full[A2_SLOTS];
  uint64_t x_ready, a_ready, acc_full, acc_empty;
  uint64_t y_full, c_full, y_done, c_free;
  uint32_t tmem_base, pad;
  float dl_part[2][A2_TM];
};

template <bool INVERSE, int ACT>
__global__ void __launch_bounds__(A2_THREADS, 2) affine_coupling_tc2_kernel(const A2Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array keeps the address space: LDS/STS instead of generic LD/ST)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = base;
  A2Smem* S = (A2Smem*)(base + A2_SLOTS * A2_SLOT_BYTES);
  float* bias_s = (float*)(S + 1);                 // [2][bias_floats]
  float* ybuf = bias_s + 2 * a.bias_floats;        // [128][D_t]
  float* cbuf = ybuf + A2_TM * a.D_t;              // [128][K0raw]: raw conditioner, later mu
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.L;
  const int units_per_tile = 2 * L;
  const long long n_my = (a.ntiles > blockIdx.x) ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x == 0) {
    mbar_init(&S->full[0], 1);
    mbar_init(&S->full[1], 1);
    mbar_init(&S->x_ready, A2_EPI_WARPS);
    mbar_init(&S->a_ready, A2_EPI_WARPS);
    mbar_init(&S->acc_full, 1);
    mbar_init(&S->acc_empty, A2_EPI_WARPS);
    mbar_init(&S->y_full, 1);
    mbar_init(&S->c_full, 1);
    mbar_init(&S->y_done, A2_EPI_WARPS);
    mbar_init(&S->c_free, A2_EPI_WARPS);
    fence_mbar_init();
  }
  for (int n = 0; n < 2; ++n) {
    int off = 0;
    for (int l = 0; l < L; ++l) {
      for (int i = threadIdx.x; i < a.net[n].Np[l]; i += A2_THREADS) bias_s[n * a.bias_floats + off + i] = a.net[n].bias[l][i];
      off += a.net[n].Np[l];
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
      auto rows_of = [&](long long it) { return (int)min((long long)A2_TM, a.B - tile_of(it) * A2_TM); };
      long long c_next = 0, y_store_next = 0, y_load_next = 0;
      uint32_t ph_cfree = 0, ph_ydone = 0;
      auto service_io = [&]() {
        if (c_next < n_my) {
          bool ok = (c_next == 0);
          if (!ok && mbar_try_wait(&S->c_free, ph_cfree)) { ph_cfree ^= 1; ok = true; }
          if (ok) {
            const uint32_t nb = (uint32_t)(rows_of(c_next) * a.K0raw * 4);
            mbar_expect_tx(&S->c_full, nb);
            bulk_g2s(cbuf, a.cond + tile_of(c_next) * A2_TM * (long long)a.K0raw, nb, &S->c_full);
            ++c_next;
          }
        }
        if (y_store_next < y_load_next && mbar_try_wait(&S->y_done, ph_ydone)) {
          ph_ydone ^= 1;
          const uint32_t nb = (uint32_t)(rows_of(y_store_next) * a.D_t * 4);
          bulk_s2g(a.tout + tile_of(y_store_next) * A2_TM * (long long)a.D_t, ybuf, nb);
          bulk_store_wait_read();
          ++y_store_next;
        }
        if (y_load_next < n_my && y_load_next == y_store_next) {
          const uint32_t nb = (uint32_t)(rows_of(y_load_next) * a.D_t * 4);
          mbar_expect_tx(&S->y_full, nb);
          bulk_g2s(ybuf, a.tin + tile_of(y_load_next) * A2_TM * (long long)a.D_t, nb, &S->y_full);
          ++y_load_next;
        }
      };
      int slot = 0;
      long long filled = 0, released = 0, events = 0;
      uint32_t eph = 0;
      for (long long it = 0; it < n_my; ++it)
        for (int n = 0; n < 2; ++n)
          for (int l = 0; l < L; ++l)
            for (int t = 0; t < a.ktiles[l]; ++t) {
              while (filled - released >= A2_SLOTS) {
                service_io();
                if (mbar_try_wait(&S->acc_full, eph)) {
                  eph ^= 1;
                  released += a.ktiles[(int)(events % L)];
                  ++events;
                } else if (a.status && *(volatile int*)a.status) {
                  break;
                }
              }
              service_io();
              uint8_t* dst = ring + (size_t)slot * A2_SLOT_BYTES;
              mbar_expect_tx(&S->full[slot], A2_SLOT_BYTES);
              bulk_g2s(dst, a.wb[n][0][l] + (long long)t * 8192, A2_TILE_BYTES, &S->full[slot]);
              bulk_g2s(dst + A2_TILE_BYTES, a.wb[n][1][l] + (long long)t * 8192, A2_TILE_BYTES, &S->full[slot]);
              ++filled;
              slot ^= 1;
            }
      (void)units_per_tile;
      for (uint32_t spin = 0; y_store_next < n_my && spin < (1u << 26); ++spin) {
        service_io();
        if (a.status && (spin & 0xfff) == 0xfff && *(volatile int*)a.status) break;
      }
      if (y_store_next < n_my && a.status) atomicExch(a.status, 1);
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer.  WIDE: every lane runs the
    // (warp-uniform) loop and one elected lane issues (bgx_tc.cuh: mma3_bf16x3_elect); otherwise lane 0 alone.
    auto issue = [&](auto wide_tag) {
      constexpr bool WIDE = decltype(wide_tag)::value;
      if (!WIDE && lane != 0) return;
      auto mma3 = [&](uint32_t x1, uint32_t x2, uint64_t e1, uint64_t e2, uint32_t idesc, uint32_t acc) {
        if (WIDE) {
          mma3_bf16x3_elect(tmem + A2_ACC, x1, x2, e1, e2, idesc, acc);
        } else {
          mma_bf16_ts(tmem + A2_ACC, x1, e2, idesc, acc);
          mma_bf16_ts(tmem + A2_ACC, x2, e1, idesc, 1);
          mma_bf16_ts(tmem + A2_ACC, x1, e1, idesc, 1);
        }
      };
    {
      const uint32_t idesc_h = idesc_bf16(128, 128), idesc_f = idesc_bf16(128, a.nfin);
      int slot = 0;
      uint32_t ph_full[2] = {0, 0};
      uint32_t ph_x = 0, ph_a = 0, ph_e = 0;
      bool first = true;
      for (long long it = 0; it < n_my; ++it) {
        for (int n = 0; n < 2; ++n)
          for (int l = 0; l < L; ++l) {
            if (l == 0) { mbar_wait(&S->x_ready, ph_x, a.status); ph_x ^= 1; }
            else { mbar_wait(&S->a_ready, ph_a, a.status); ph_a ^= 1; }
            // accumulator drained?  net 0 / layer 0: the previous tile's s pull; net 1 / layer 0: this tile's mu pull
            if (l == 0 && !(n == 0 && first)) { mbar_wait(&S->acc_empty, ph_e, a.status); ph_e ^= 1; }
            tc_fence_after();
            const uint32_t idesc = (l == L - 1) ? idesc_f : idesc_h;
            const int ksteps_total = (a.net[n].K[l] + 15) / 16;
            uint32_t acc = 0;
            for (int t = 0; t < a.ktiles[l]; ++t) {
              mbar_wait(&S->full[slot], ph_full[slot], a.status);
              ph_full[slot] ^= 1;
              const uint32_t b1 = smem_u32(ring + (size_t)slot * A2_SLOT_BYTES), b2 = b1 + A2_TILE_BYTES;
              slot ^= 1;
              tc_fence_after();
              const int nk = min(4, ksteps_total - t * 4);
              const uint64_t d1 = smem_desc_sw128(b1), d2 = smem_desc_sw128(b2);   // +2 per k-step (16-byte units)
              const uint32_t a1 = tmem + A2_A + (uint32_t)(t * 32), a2 = a1 + A2_A_STRIDE;
              if (nk == 4) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  mma3(a1 + ks * 8, a2 + ks * 8, d1 + 2 * ks, d2 + 2 * ks, idesc, ks == 0 ? acc : 1u);
              } else {
                for (int ks = 0; ks < nk; ++ks)
                  mma3(a1 + ks * 8, a2 + ks * 8, d1 + 2 * ks, d2 + 2 * ks, idesc, ks == 0 ? acc : 1u);
              }
              acc = 1;
            }
            if (WIDE) mma_commit_elect(&S->acc_full);
            else mma_commit(&S->acc_full);
          }
        first = false;
      }
    }
    };
    if (a.wide_mma) issue(std::true_type{});
    else issue(std::false_type{});
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (0..7)
    const int q = warp & 3, j = warp >> 2;
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t ph_acc = 0, ph_c = 0, ph_y = 0;
    int fin_off = 0;
    for (int l = 0; l < L - 1; ++l) fin_off += a.net[0].Np[l];
    const int K0 = a.net[0].K[0];
    const int nblk = (a.D_t + 31) / 32;
    float* crow = cbuf + r_in_tile * a.K0raw;

    auto cond_value = [&](int k) -> float {
      if (k >= K0) return 0.f;
      const int code = a.net[0].in_map[k];
      const float v = crow[code & 0xffffff];
      const int kind = code >> 24;
      if (kind == 0) return v;
      const float arg = (v - a.net[0].pleft) * a.net[0].pscale;
      return kind == 1 ? cosf(arg) : sinf(arg);
    };
    auto stage_x = [&](bool fresh) {
      if (fresh) { mbar_wait(&S->c_full, ph_c, a.status); ph_c ^= 1; }
      for (int b0 = j * 16; b0 < K0; b0 += 32) {
        uint32_t t1[8], t2[8], t3[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_bf16(cond_value(b0 + 2 * i), cond_value(b0 + 2 * i + 1), 2, t1[i], t2[i], t3[i]);
        const uint32_t col = tmem + lane_base + A2_A + b0 / 2;
        tmem_st8(col, t1);
        tmem_st8(col + A2_A_STRIDE, t2);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->x_ready);
    };
    auto hidden_layers = [&](int n) {
      const float* bias_n = bias_s + n * a.bias_floats;
      int boff = 0;
      for (int l = 0; l < L - 1; ++l) {
        mbar_wait(&S->acc_full, ph_acc, a.status);
        ph_acc ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int col = (j * 2 + h) * 32;
          uint32_t v[32];
          tmem_ld32(tmem + lane_base + A2_ACC + col, v);
          tmem_ld_wait();
          uint32_t t1[16], t2[16], t3[16];
          const float4* b4 = reinterpret_cast<const float4*>(bias_n + boff + col);
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
          const uint32_t acol = tmem + lane_base + A2_A + col / 2;
          tmem_st16(acol, t1);
          tmem_st16(acol + A2_A_STRIDE, t2);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->a_ready);
        boff += a.net[n].Np[l];
      }
    };

    for (long long it = 0; it < n_my; ++it) {
      const long long tile = blockIdx.x + it * (long long)gridDim.x;
      const long long row = tile * A2_TM + r_in_tile;
      float* yrow = ybuf + r_in_tile * a.D_t;
      if (it == 0) stage_x(true);
      // ---- net 0 (shift)
      hidden_layers(0);
      mbar_wait(&S->acc_full, ph_acc, a.status);     // mu accumulator complete, net 0's activations dead
      ph_acc ^= 1;
      tc_fence_after();
      stage_x(false);                                 // net 1's first-layer operand from the same conditioner tile
      asm volatile("bar.sync 1, 256;" ::: "memory"); // every epilogue thread has read the raw tile
      {
        const float* bmu = bias_s + fin_off;
        for (int b = 0; b < nblk; ++b) {
          uint32_t vm[32];
          tmem_ld32(tmem + lane_base + A2_ACC + b * 32, vm);
          tmem_ld_wait();
          if (b == nblk - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->acc_empty);
          }
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            if (jj != j) continue;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = 2 * i + jj, d = b * 32 + c;
              if (d < a.D_t) crow[d] = __uint_as_float(vm[c]) + bmu[d];     // park mu in the conditioner buffer
            }
          }
        }
      }
      // ---- net 1 (scale)
      hidden_layers(1);
      mbar_wait(&S->acc_full, ph_acc, a.status);
      ph_acc ^= 1;
      tc_fence_after();
      mbar_wait(&S->y_full, ph_y, a.status);
      ph_y ^= 1;
      float ld = 0.f;
      {
        const float* bsc = bias_s + a.bias_floats + fin_off;
        for (int b = 0; b < nblk; ++b) {
          uint32_t vs[32];
          tmem_ld32(tmem + lane_base + A2_ACC + b * 32, vs);
          tmem_ld_wait();
          if (b == nblk - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->acc_empty);
          }
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            if (jj != j) continue;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = 2 * i + jj, d = b * 32 + c;
              if (d < a.D_t) {
                const float mu = crow[d];
                const float sv = __uint_as_float(vs[c]) + bsc[d];
                const float ls = (1.f - 2.f * rcp_fast(1.f + ex2_fast(2.f * LOG2E * sv))) * a.alpha;
                const float y = yrow[d];
                yrow[d] = INVERSE ? (y - mu) * ex2_fast(-LOG2E * ls) : fmaf(y, ex2_fast(LOG2E * ls), mu);
                ld += INVERSE ? -ls : ls;
              }
            }
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->c_free);        // mu consumed: the conditioner buffer may be refilled
      if (it + 1 < n_my) stage_x(true);               // next tile's first operand (all MMAs of this tile are done)
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->y_done);
      S->dl_part[j][r_in_tile] = ld;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (j == 0 && row < a.B) {
        const float base_dl = a.dlogp_in ? a.dlogp_in[row] : 0.f;
        a.dlogp_out[row] = base_dl + (S->dl_part[0][r_in_tile] + S->dl_part[1][r_in_tile]);
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");
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

bool affine_tc_eligible(const bgx_packed_mlp* shift, const bgx_packed_mlp* scale, int flags);

bool affine_tc2_eligible(const bgx_coupling_io* io, const bgx_packed_mlp* shift, const bgx_packed_mlp* scale, int flags) {
  if (!io || !affine_tc_eligible(shift, scale, flags) || (flags & BGX_FLAG_BF16X6)) return false;
  const char* e = getenv("BGX_TC_SINGLE_CTA");
  if (e && atoi(e)) return false;
  if (io->n_cond != 1 || io->n_tr != 1 || io->batch % 4) return false;
  auto dense = [](const bgx_seg& s) { return s.stride == s.width && ((uintptr_t)s.ptr & 15) == 0; };
  if (!dense(io->cond[0]) || !dense(io->tr_in[0]) || !dense(io->tr_out[0])) return false;
  if (io->cond[0].width != shift->raw_width || io->tr_in[0].width > io->cond[0].width) return false;
  size_t bias = 0;
  for (int l = 0; l < shift->n_layers; ++l) bias += shift->Np[l];
  const size_t need = 1024 + A2_SLOTS * A2_SLOT_BYTES + sizeof(A2Smem) +
                      4 * (2 * bias + (size_t)A2_TM * (io->tr_in[0].width + io->cond[0].width)) + 64;
  return need <= 112 * 1024;
}

int affine_coupling_tc2(const bgx_coupling_io* io, const bgx_packed_mlp* shift, const bgx_packed_mlp* scale,
                        float log_alpha, int flags, int* status, cudaStream_t st) {
  const int L = shift->n_layers;
  const int d_t = io->tr_in[0].width;
  if (shift->N[L - 1] != d_t || io->tr_out[0].width != d_t || !io->dlogp_out) return BGX_ERR_INVALID;
  if (io->batch == 0) return BGX_OK;
  A2Args a{};
  a.B = io->batch;
  a.cond = io->cond[0].ptr; a.tin = io->tr_in[0].ptr; a.tout = const_cast<float*>(io->tr_out[0].ptr);
  a.D_t = d_t; a.K0raw = io->cond[0].width; a.nfin = round_up(d_t, 16);
  mlp_to_dev(shift, a.net[0]);
  mlp_to_dev(scale, a.net[1]);
  a.L = L;
  int bias_floats = 0;
  for (int l = 0; l < L; ++l) {
    for (int t = 0; t < 2; ++t) {
      a.wb[0][t][l] = (const uint16_t*)shift->Wb[t][l];
      a.wb[1][t][l] = (const uint16_t*)scale->Wb[t][l];
    }
    a.ktiles[l] = ceil_div(shift->K[l], 64);
    bias_floats += shift->Np[l];
  }
  a.bias_floats = bias_floats;
  a.inverse = (flags & BGX_FLAG_INVERSE) ? 1 : 0;
  a.alpha = expf(log_alpha);
  a.dlogp_in = io->dlogp_in;
  a.dlogp_out = io->dlogp_out;
  a.status = status;
  a.ntiles = (a.B + A2_TM - 1) / A2_TM;
  {
    static const int wide = [] { const char* e = getenv("BGX_T2_WIDE_MMA"); return e ? atoi(e) : 1; }();
    a.wide_mma = wide;
  }
  const size_t smem = 1024 + A2_SLOTS * A2_SLOT_BYTES + sizeof(A2Smem) +
                      sizeof(float) * (2 * (size_t)bias_floats + (size_t)A2_TM * (a.D_t + a.K0raw)) + 64;
  int sm_count = 0;
  int rc = device_sm_count(&sm_count);
  if (rc) return rc;
  using KernT = void (*)(const A2Args);
  static const KernT kerns[2][4] = {
      {affine_coupling_tc2_kernel<false, 0>, affine_coupling_tc2_kernel<false, 1>, affine_coupling_tc2_kernel<false, 2>,
       affine_coupling_tc2_kernel<false, 3>},
      {affine_coupling_tc2_kernel<true, 0>, affine_coupling_tc2_kernel<true, 1>, affine_coupling_tc2_kernel<true, 2>,
       affine_coupling_tc2_kernel<true, 3>}};
  if (shift->act < 0 || shift->act > 3) return BGX_ERR_INVALID;
  KernT kern = kerns[a.inverse][shift->act];
  static size_t configured_all[BGX_MAX_DEVICES][2][4] = {};
  auto& configured = configured_all[device_slot()];
  if (smem > configured[a.inverse][shift->act]) {
    rc = check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
    rc = check(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (rc) return rc;
    configured[a.inverse][shift->act] = smem;
  }
  const unsigned grid = (unsigned)std::min<long long>(a.ntiles, 2LL * sm_count);
  kern<<<grid, A2_THREADS, smem, st>>>(a);
  return post_launch();
}

}  // namespace bgx
