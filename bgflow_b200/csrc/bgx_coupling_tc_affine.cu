// Tensor-core fused AFFINE (RealNVP) coupling block for sm_100a — same machinery as
// bgx_coupling_tc.cu (persistent CTA per SM, 128-sample tiles, bulk-TMA weight ring, tcgen05.mma
// kind::f16 with exact bf16 operand splits, A operand and accumulators in tensor memory, I/O
// warps staging the tile through shared memory), different program per tile:
//
//   net 0 (shift): x -> hidden layers (ACC0) -> mu   into ACC1   (one MMA pass, N = round16(D_t))
//   net 1 (scale): x -> hidden layers (ACC0) -> s    into ACC0
//   epilogue:      log_sigma = tanh(s) * exp(log_alpha);  y' = y * exp(log_sigma) + mu   (forward)
//                  y' = (y - mu) * exp(-log_sigma)  (inverse);  dlogp = +-sum(log_sigma)
//
// Replaces coupling.py:162-182 + transformer/affine.py:35-70 + dense.py:47-48 for blocks whose two
// conditioners are DenseNets with hidden width 128 (BASELINE config 2: 33-128-128-128-33 ReLU x2).
// volume-preserving / circular variants and other widths run the SIMT kernel.
#include <cstdlib>

#include "bgx_coupling.cuh"
#include "bgx_tc.cuh"
#include "bgx_tc_epi.cuh"

namespace bgx {
using namespace tc;

constexpr int AF_THREADS = 640;
constexpr int AF_EPI_WARPS = 16;
constexpr int AF_TM = 128;
constexpr int AF_MAX_SLOTS = 12;
constexpr uint32_t AF_TILE_BYTES = 16384;
constexpr int AF_ACC0 = 0, AF_ACC1 = 128, AF_A = 256, AF_A_STRIDE = 64;

struct AfArgs {
  long long B;
  Segs cond, tin, tout;
  int D_t, nfin;     // nfin = D_t rounded up to 16: N of the final MMA
  DevMlp net[2];
  const uint16_t* wb[2][3][BGX_MAX_LAYERS];
  int ktiles[BGX_MAX_LAYERS];
  int L, act;
  int nterms, inverse;
  float alpha;
  const float* dlogp_in;
  float* dlogp_out;
  int* status;
  long long ntiles;
  int bias_floats;   // per net
  int ldy, ldc, stages, y_dense, c_dense;
};

struct alignas(16) AfSmem {
  uint64_t full[AF_MAX_SLOTS];
  uint64_t x_ready, a_ready, acc_full_h;
  uint64_t fin_full[2];   // final layer of net 0 / net 1 complete
  uint64_t fin_empty;     // 16 arrivals: both final accumulators pulled into registers
  uint64_t y_full[2], y_done[2], c_full, c_free;
  uint32_t tmem_base, pad;
  float dl_part[4][AF_TM];
};

template <bool INVERSE, int ACT>
__global__ void __launch_bounds__(AF_THREADS, 1) affine_coupling_tc_kernel(const AfArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array keeps the address space: LDS/STS instead of generic LD/ST)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NT = a.nterms, NST = a.stages, L = a.L;
  uint8_t* ring = base;
  AfSmem* S = (AfSmem*)(base + (size_t)NST * NT * AF_TILE_BYTES);
  float* bias_s = (float*)(S + 1);                       // [2][bias_floats]
  float* ybuf = bias_s + 2 * a.bias_floats;              // 2 x [128][ldy]
  float* cbuf = ybuf + 2 * AF_TM * a.ldy;                // [128][ldc]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(&S->full[s], 1);
    mbar_init(&S->x_ready, AF_EPI_WARPS);
    mbar_init(&S->a_ready, AF_EPI_WARPS);
    mbar_init(&S->acc_full_h, 1);
    mbar_init(&S->fin_full[0], 1);
    mbar_init(&S->fin_full[1], 1);
    mbar_init(&S->fin_empty, AF_EPI_WARPS);
    mbar_init(&S->y_full[0], 2);
    mbar_init(&S->y_full[1], 2);
    mbar_init(&S->y_done[0], AF_EPI_WARPS);
    mbar_init(&S->y_done[1], AF_EPI_WARPS);
    mbar_init(&S->c_full, 2);
    mbar_init(&S->c_free, AF_EPI_WARPS);
    fence_mbar_init();
  }
  for (int n = 0; n < 2; ++n) {
    int off = 0;
    for (int l = 0; l < L; ++l) {
      for (int i = threadIdx.x; i < a.net[n].Np[l]; i += AF_THREADS) bias_s[n * a.bias_floats + off + i] = a.net[n].bias[l][i];
      off += a.net[n].Np[l];
    }
  }
  if (warp == 16) tmem_alloc<512>(&S->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S->tmem_base;

  if (warp == 18) {
    // ------------------------------------------------------------------ weight producer
    if (lane == 0) {
      int slot = 0;
      long long filled = 0, released = 0;
      int e_net = 0, e_l = 0;                 // event cursor (oldest unit not yet observed)
      uint32_t eph_h = 0, eph_f[2] = {0, 0};
      for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x)
        for (int n = 0; n < 2; ++n)
          for (int l = 0; l < L; ++l)
            for (int t = 0; t < a.ktiles[l]; ++t) {
              while (filled - released >= NST) {
                if (e_l == L - 1) { mbar_wait(&S->fin_full[e_net], eph_f[e_net], a.status); eph_f[e_net] ^= 1; }
                else { mbar_wait(&S->acc_full_h, eph_h, a.status); eph_h ^= 1; }
                released += a.ktiles[e_l];
                if (++e_l == L) { e_l = 0; e_net ^= 1; }
              }
              uint8_t* dst = ring + (size_t)slot * NT * AF_TILE_BYTES;
              mbar_expect_tx(&S->full[slot], (uint32_t)NT * AF_TILE_BYTES);
              for (int part = 0; part < NT; ++part)
                bulk_g2s(dst + part * AF_TILE_BYTES, a.wb[n][part][l] + (long long)t * 8192, AF_TILE_BYTES, &S->full[slot]);
              ++filled;
              if (++slot == NST) slot = 0;
            }
    }
    __syncwarp();
  } else if (warp == 19) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_h = idesc_bf16(128, 128), idesc_f = idesc_bf16(128, a.nfin);
      int stage = 0;
      uint32_t phase = 0, ph_x = 0, ph_a = 0, ph_e = 0;
      bool first = true;
      for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        for (int n = 0; n < 2; ++n)
          for (int l = 0; l < L; ++l) {
            if (l == 0) { mbar_wait(&S->x_ready, ph_x, a.status); ph_x ^= 1; }
            else { mbar_wait(&S->a_ready, ph_a, a.status); ph_a ^= 1; }
            const bool last = (l == L - 1);
            // both final accumulators of the previous tile must have been pulled before this tile
            // overwrites ACC0 (first hidden layer) / ACC1 (net 0's final layer)
            if (n == 0 && l == 0 && !first) { mbar_wait(&S->fin_empty, ph_e, a.status); ph_e ^= 1; }
            tc_fence_after();
            const uint32_t d_tmem = tmem + ((last && n == 0) ? AF_ACC1 : AF_ACC0);
            const uint32_t idesc = last ? idesc_f : idesc_h;
            const int ksteps_total = (a.net[n].K[l] + 15) / 16;
            uint32_t acc = 0;
            for (int t = 0; t < a.ktiles[l]; ++t) {
              mbar_wait(&S->full[stage], phase, a.status);
              const uint32_t sbase = smem_u32(ring + (size_t)stage * NT * AF_TILE_BYTES);
              if (++stage == NST) { stage = 0; phase ^= 1; }
              tc_fence_after();
              const int nk = min(4, ksteps_total - t * 4);
              const uint32_t b1 = sbase, b2 = sbase + AF_TILE_BYTES, b3 = sbase + 2 * AF_TILE_BYTES;
              for (int ks = 0; ks < nk; ++ks) {
                const uint32_t kcol = (uint32_t)(t * 32 + ks * 8);
                const uint32_t a1 = tmem + AF_A + kcol, a2 = a1 + AF_A_STRIDE, a3 = a2 + AF_A_STRIDE;
                const uint64_t d1 = smem_desc_sw128(b1 + ks * 32), d2 = smem_desc_sw128(b2 + ks * 32);
                if (NT == 3) {
                  const uint64_t d3 = smem_desc_sw128(b3 + ks * 32);
                  mma_bf16_ts(d_tmem, a1, d3, idesc, acc);
                  mma_bf16_ts(d_tmem, a3, d1, idesc, 1);
                  mma_bf16_ts(d_tmem, a2, d2, idesc, 1);
                  acc = 1;
                }
                mma_bf16_ts(d_tmem, a1, d2, idesc, acc);
                mma_bf16_ts(d_tmem, a2, d1, idesc, 1);
                mma_bf16_ts(d_tmem, a1, d1, idesc, 1);
                acc = 1;
              }
            }
            mma_commit(last ? &S->fin_full[n] : &S->acc_full_h);
          }
        first = false;
      }
    }
    __syncwarp();
  } else if (warp == 16 || warp == 17) {
    // ------------------------------------------------------------------ I/O warps
    const int t64 = threadIdx.x - 512;
    const int Dt = a.D_t, K0 = a.net[0].K[0];
    auto walk = [&](int W, auto&& body) {
      int r = t64 / W, d = t64 - r * W;
      const int dr = 64 / W, dd = 64 - dr * W;
      for (; r < AF_TM;) {
        body(r, d);
        r += dr; d += dd;
        if (d >= W) { d -= W; ++r; }
      }
    };
    auto full_rows = [&](long long tile) { return (tile + 1) * AF_TM <= a.B; };
    auto load_tile = [&](long long tile, int b) {
      float* Y = ybuf + b * AF_TM * a.ldy;
      if (a.y_dense && full_rows(tile)) {
        if (warp == 16 && lane == 0) {
          mbar_expect_tx(&S->y_full[b], (uint32_t)(AF_TM * Dt * 4));
          bulk_g2s(Y, a.tin.ptr[0] + tile * AF_TM * (long long)Dt, (uint32_t)(AF_TM * Dt * 4), &S->y_full[b]);
        } else if (warp == 17 && lane == 0) {
          mbar_arrive(&S->y_full[b]);
        }
        return;
      }
      walk(Dt, [&](int r, int d) {
        const long long row = tile * AF_TM + r;
        if (row < a.B) cp_async4(&Y[r * a.ldy + d], seg_addr(a.tin, row, d));
        else Y[r * a.ldy + d] = 0.f;
      });
      cp_async_drain();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->y_full[b]);
    };
    auto load_cond_tile = [&](long long tile) {
      if (a.c_dense && full_rows(tile)) {
        if (warp == 16 && lane == 0) {
          mbar_expect_tx(&S->c_full, (uint32_t)(AF_TM * K0 * 4));
          bulk_g2s(cbuf, a.cond.ptr[0] + tile * AF_TM * (long long)K0, (uint32_t)(AF_TM * K0 * 4), &S->c_full);
        } else if (warp == 17 && lane == 0) {
          mbar_arrive(&S->c_full);
        }
        return;
      }
      walk(K0, [&](int r, int k) {
        const long long row = tile * AF_TM + r;
        const int code = a.net[0].in_map[k];
        float* dst = &cbuf[r * a.ldc + k];
        if (row >= a.B) *dst = 0.f;
        else if ((code >> 24) == 0) cp_async4(dst, seg_addr(a.cond, row, code & 0xffffff));
        else *dst = load_cond(a.cond, a.net[0], a.B, row, k);
      });
      cp_async_drain();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->c_full);
    };
    auto store_tile = [&](long long tile, int b) {
      const float* Y = ybuf + b * AF_TM * a.ldy;
      if (a.y_dense && full_rows(tile)) {
        if (warp == 16 && lane == 0) {
          bulk_s2g(const_cast<float*>(a.tout.ptr[0]) + tile * AF_TM * (long long)Dt, Y, (uint32_t)(AF_TM * Dt * 4));
          bulk_store_wait_read();
        }
        return;
      }
      walk(Dt, [&](int r, int d) {
        const long long row = tile * AF_TM + r;
        if (row < a.B) *const_cast<float*>(seg_addr(a.tout, row, d)) = Y[r * a.ldy + d];
      });
    };
    long long n_my = (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    if (n_my > 0) load_cond_tile(blockIdx.x);
    if (n_my > 0) load_tile(blockIdx.x, 0);
    if (n_my > 1) load_tile(blockIdx.x + (long long)gridDim.x, 1);
    for (long long it = 0; it < n_my; ++it) {
      const int b = (int)(it & 1);
      const long long tile = blockIdx.x + it * gridDim.x;
      if (it + 1 < n_my) {
        mbar_wait(&S->c_free, (uint32_t)(it & 1), a.status);
        load_cond_tile(tile + gridDim.x);
      }
      mbar_wait(&S->y_done[b], (uint32_t)((it >> 1) & 1), a.status);
      store_tile(tile, b);
      if (it + 2 < n_my) {
        asm volatile("bar.sync 3, 64;" ::: "memory");
        load_tile(tile + 2LL * gridDim.x, b);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (0..15)
    const int q = warp & 3, j = warp >> 2;
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t ph_h = 0, ph_f0 = 0, ph_f1 = 0, ph_c = 0;
    const int K0 = a.net[0].K[0];
    int fin_off = 0;
    for (int l = 0; l < L - 1; ++l) fin_off += a.net[0].Np[l];

    // conditioner tile (shared memory) -> bf16 terms -> A operand of a net's first layer
    auto stage_x = [&](bool fresh, bool release) {
      if (fresh) { mbar_wait(&S->c_full, ph_c, a.status); ph_c ^= 1; }
      const float* crow = cbuf + r_in_tile * a.ldc;
      for (int b0 = j * 16; b0 < K0; b0 += 64) {
        uint32_t t1[8], t2[8], t3[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = b0 + 2 * i;
          split_bf16(k < K0 ? crow[k] : 0.f, k + 1 < K0 ? crow[k + 1] : 0.f, NT, t1[i], t2[i], t3[i]);
        }
        const uint32_t col = tmem + lane_base + AF_A + b0 / 2;
        tmem_st8(col, t1);
        tmem_st8(col + AF_A_STRIDE, t2);
        if (NT == 3) tmem_st8(col + 2 * AF_A_STRIDE, t3);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&S->x_ready);
        if (release) mbar_arrive(&S->c_free);
      }
    };

    bool first = true;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
      const long long row = tile * AF_TM + r_in_tile;
      const int yb = (int)(it & 1);
      float* yrow = ybuf + yb * AF_TM * a.ldy + r_in_tile * a.ldy;
      if (first) stage_x(true, false);
      first = false;
      for (int n = 0; n < 2; ++n) {
        const float* bias_n = bias_s + n * a.bias_floats;
        int boff = 0;
        for (int l = 0; l < L - 1; ++l) {
          mbar_wait(&S->acc_full_h, ph_h, a.status);
          ph_h ^= 1;
          tc_fence_after();
          {
            const int col = j * 32;
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + AF_ACC0 + col, v);
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
              split_bf16(h0, h1, NT, t1[2 * i], t2[2 * i], t3[2 * i]);
              split_bf16(h2, h3, NT, t1[2 * i + 1], t2[2 * i + 1], t3[2 * i + 1]);
            }
            const uint32_t acol = tmem + lane_base + AF_A + col / 2;
            tmem_st16(acol, t1);
            tmem_st16(acol + AF_A_STRIDE, t2);
            if (NT == 3) tmem_st16(acol + 2 * AF_A_STRIDE, t3);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S->a_ready);
          boff += a.net[n].Np[l];
        }
        if (n == 0) {
          // net 0's final MMAs are complete -> its activations are dead: stage x again for net 1
          mbar_wait(&S->fin_full[0], ph_f0, a.status);
          ph_f0 ^= 1;
          tc_fence_after();
          stage_x(false, true);
        }
      }
      // ---- both final layers done: mu in ACC1, s in ACC0
      mbar_wait(&S->fin_full[1], ph_f1, a.status);
      ph_f1 ^= 1;
      tc_fence_after();
      mbar_wait(&S->y_full[yb], (uint32_t)((it >> 1) & 1), a.status);
      float ld = 0.f;
      const float* bmu = bias_s + fin_off;
      const float* bsc = bias_s + a.bias_floats + fin_off;
      const int nblk = (a.D_t + 31) / 32;
      for (int b = 0; b < nblk; ++b) {
        uint32_t vm[32], vs[32];
        tmem_ld32(tmem + lane_base + AF_ACC1 + b * 32, vm);
        tmem_ld32(tmem + lane_base + AF_ACC0 + b * 32, vs);
        tmem_ld_wait();
        if (b == nblk - 1) {          // last pull: hand both accumulators back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S->fin_empty);
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {        // compile-time register indices: branch on the warp's j
          if (jj != j) continue;
#pragma unroll
          for (int i = 0; i < 8; ++i) {         // this warp's dims of the block: d = 32 b + 4 i + j
            const int c = 4 * i + jj, d = b * 32 + c;
            if (d < a.D_t) {
              const float mu = __uint_as_float(vm[c]) + bmu[d];
              const float sv = __uint_as_float(vs[c]) + bsc[d];
              const float ls = (1.f - 2.f * rcp_fast(1.f + ex2_fast(2.f * LOG2E * sv))) * a.alpha;   // tanh(s) * alpha
              const float y = yrow[d];
              yrow[d] = INVERSE ? (y - mu) * ex2_fast(-LOG2E * ls) : fmaf(y, ex2_fast(LOG2E * ls), mu);
              ld += INVERSE ? -ls : ls;
            }
          }
        }
      }
      // next tile's conditioner input (for net 0): every MMA of this tile has completed
      {
        const long long next = tile + gridDim.x;
        if (next < a.ntiles) stage_x(true, false);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->y_done[yb]);
      S->dl_part[j][r_in_tile] = ld;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (j == 0 && row < a.B) {
        const float base_dl = a.dlogp_in ? a.dlogp_in[row] : 0.f;
        a.dlogp_out[row] = base_dl + ((S->dl_part[0][r_in_tile] + S->dl_part[1][r_in_tile]) +
                                      (S->dl_part[2][r_in_tile] + S->dl_part[3][r_in_tile]));
      }
      asm volatile("bar.sync 2, 512;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------ host side

bool affine_tc_eligible(const bgx_packed_mlp* shift, const bgx_packed_mlp* scale, int flags) {
  if (!shift || !scale) return false;
  if (flags & (BGX_FLAG_PRESERVE_VOLUME | BGX_FLAG_CIRCULAR)) return false;
  const int L = shift->n_layers;
  if (L != scale->n_layers || L < 2 || L > 6 || shift->act != scale->act) return false;
  if (shift->K[0] != scale->K[0] || shift->K[0] > 128 || shift->raw_width != scale->raw_width) return false;
  if (shift->periodic_scale != scale->periodic_scale || shift->periodic_left != scale->periodic_left) return false;
  for (int l = 0; l + 1 < L; ++l)
    if (shift->N[l] != 128 || scale->N[l] != 128) return false;
  if (shift->N[L - 1] != scale->N[L - 1] || shift->N[L - 1] > 128) return false;
  for (int l = 0; l < L; ++l)
    for (int t = 0; t < 3; ++t)
      if (!shift->Wb[t][l] || !scale->Wb[t][l]) return false;
  return true;
}

int affine_coupling_tc(const bgx_coupling_io* io, const bgx_packed_mlp* shift, const bgx_packed_mlp* scale,
                       float log_alpha, int flags, int* status, cudaStream_t st) {
  CouplingArgs ca{};
  int d_c, d_t;
  int rc = coupling_fill_io(io, ca, d_c, d_t);
  if (rc) return rc;
  const int L = shift->n_layers;
  if (shift->raw_width != d_c || shift->N[L - 1] != d_t) return BGX_ERR_INVALID;
  if (ca.B == 0) return BGX_OK;
  AfArgs a{};
  a.B = ca.B;
  a.cond = ca.cond; a.tin = ca.tin; a.tout = ca.tout;
  a.D_t = d_t;
  a.nfin = round_up(d_t, 16);
  mlp_to_dev(shift, a.net[0]);
  mlp_to_dev(scale, a.net[1]);
  a.L = L;
  a.act = shift->act;
  int bias_floats = 0;
  for (int l = 0; l < L; ++l) {
    for (int t = 0; t < 3; ++t) {
      a.wb[0][t][l] = (const uint16_t*)shift->Wb[t][l];
      a.wb[1][t][l] = (const uint16_t*)scale->Wb[t][l];
    }
    a.ktiles[l] = ceil_div(shift->K[l], 64);
    bias_floats += shift->Np[l];
  }
  a.bias_floats = bias_floats;
  a.nterms = (flags & BGX_FLAG_BF16X6) ? 3 : 2;
  a.inverse = (flags & BGX_FLAG_INVERSE) ? 1 : 0;
  a.alpha = expf(log_alpha);
  a.dlogp_in = ca.dlogp_in;
  a.dlogp_out = ca.dlogp_out;
  a.status = status;
  a.ntiles = (a.B + AF_TM - 1) / AF_TM;
  auto dense16 = [](const Segs& sg) {
    return sg.n == 1 && sg.stride[0] == sg.width[0] && ((uintptr_t)sg.ptr[0] & 15) == 0;
  };
  a.y_dense = dense16(a.tin) && dense16(a.tout);
  a.c_dense = dense16(a.cond) && shift->periodic_scale == 0.f && shift->raw_width == shift->K[0];
  a.ldy = a.y_dense ? d_t : (d_t | 1);
  a.ldc = a.c_dense ? shift->K[0] : (shift->K[0] | 1);
  const size_t fixed = 1024 + sizeof(AfSmem) +
                       sizeof(float) * (2 * (size_t)bias_floats + 2 * AF_TM * a.ldy + AF_TM * a.ldc) + 64;
  const size_t slot_bytes = (size_t)a.nterms * AF_TILE_BYTES;
  a.stages = AF_MAX_SLOTS;
  while (a.stages > 3 && fixed + (size_t)a.stages * slot_bytes > 227 * 1024) a.stages -= 1;
  const size_t smem = fixed + (size_t)a.stages * slot_bytes;
  if (smem > 227 * 1024) return BGX_ERR_UNSUPPORTED;
  int sm_count = 0;
  rc = device_sm_count(&sm_count);
  if (rc) return rc;
  using KernT = void (*)(const AfArgs);
  static const KernT kerns[2][4] = {
      {affine_coupling_tc_kernel<false, 0>, affine_coupling_tc_kernel<false, 1>, affine_coupling_tc_kernel<false, 2>,
       affine_coupling_tc_kernel<false, 3>},
      {affine_coupling_tc_kernel<true, 0>, affine_coupling_tc_kernel<true, 1>, affine_coupling_tc_kernel<true, 2>,
       affine_coupling_tc_kernel<true, 3>}};
  if (a.act < 0 || a.act > 3) return BGX_ERR_INVALID;
  KernT kern = kerns[a.inverse][a.act];
  static size_t configured_all[BGX_MAX_DEVICES][2][4] = {};
  auto& configured = configured_all[device_slot()];
  if (smem > configured[a.inverse][a.act]) {
    rc = check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
    configured[a.inverse][a.act] = smem;
  }
  const unsigned grid = (unsigned)std::min<long long>(a.ntiles, sm_count);
  kern<<<grid, AF_THREADS, smem, st>>>(a);
  return post_launch();
}

}  // namespace bgx
