// Tensor-core fused RealNVP (affine) coupling block for WIDE shapes (BASELINE config 5: D = 384 / 3072):
// the same one-CTA-per-SM machine as bgx_coupling_pair.cu, but the two tensor-memory slots hold the two
// conditioner NETS of one 128-row tile instead of two tiles of one net:
//   slot 0 = shift net (mu), slot 1 = scale net (s); each unit (a 128-input group of layer 0, a hidden layer, a
//   128-column pass of the last layer) is issued for slot 0 and slot 1 back to back with that net's weights,
//   so the epilogue of one net overlaps the MMAs of the other; the last layer's epilogue pulls mu and s of the
//   same 32 columns, releases both accumulators and then evaluates
//       y' = y exp(tanh(s) alpha) + mu   (forward)      y' = (y - mu) exp(-tanh(s) alpha)   (inverse)
//   with dlogp = +- sum tanh(s) alpha (bgflow/nn/flow/transformer/affine.py:35-70) while the next pass runs.
// Transformed inputs / outputs are read and written in place in global memory (each thread owns 32
// consecutive columns of its row per pass: full 128-byte lines).
//
//   warps 0-15 epilogue (quadrant w % 4, column share w / 4), 16 weight producer, 17 idle, 18 MMA issuer,
//   19 log-det reducer.  TMEM: slot s at 256 s: [0,128) accumulator, [128,192) A term 1, [192,256) A term 2.
//
// Volume-preserving / circular / shift-only blocks are not handled here (bgx_api.cu sends them to the other kernels).
#include <cstdlib>

#include "bgx_pair.cuh"

namespace bgx {

struct AffPArgs {
  long long B;
  const float* cond;    // [B][K0raw] dense
  const float* tin;     // [B][D_t] dense
  float* tout;          // [B][D_t] dense
  int D_t, K0raw;
  DevMlp net[2];        // shift, scale (same layer structure)
  const uint16_t* wb[2][2][BGX_MAX_LAYERS];   // [net][term][layer]
  int ktiles[BGX_MAX_LAYERS];
  int npass, G;
  float alpha;
  const float* dlogp_in;
  float* dlogp_out;
  int* status;
  long long ntiles;
  int hid_bias_floats;  // per net
  int plain_cond, vec_ok;
};

struct alignas(16) AffPSmem {
  uint64_t w_full[P_STAGES], w_empty[P_STAGES];
  uint64_t a_ready[2], acc_full[2], acc_empty[2];
  uint64_t dl_ready, dl_free;
  uint32_t tmem_base, pad[3];
  float dl_part[4][P_TM];
};

template <bool INVERSE, int ACT>
__global__ void __launch_bounds__(P_THREADS, 1) affine_coupling_pair_kernel(const __grid_constant__ AffPArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = base;
  AffPSmem* S = (AffPSmem*)(base + P_STAGES * P_STAGE_BYTES);
  float* bias_h = (float*)(S + 1);          // [net][hidden layers x 128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.net[0].n_layers;
  const int G = a.G, P = a.npass;
  const long long n_my = (a.ntiles > blockIdx.x) ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < P_STAGES; ++i) {
      mbar_init(&S->w_full[i], 1);
      mbar_init(&S->w_empty[i], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&S->a_ready[s], P_EPI_WARPS);
      mbar_init(&S->acc_full[s], 1);
      mbar_init(&S->acc_empty[s], P_EPI_WARPS);
    }
    mbar_init(&S->dl_ready, P_EPI_WARPS);
    mbar_init(&S->dl_free, 1);
    fence_mbar_init();
  }
  for (int n = 0; n < 2; ++n) {
    int off = n * a.hid_bias_floats;
    for (int l = 0; l < L - 1; ++l) {
      for (int i = threadIdx.x; i < a.net[n].Np[l]; i += P_THREADS) bias_h[off + i] = a.net[n].bias[l][i];
      off += a.net[n].Np[l];
    }
  }
  if (warp == 18) tmem_alloc<512>(&S->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S->tmem_base;
  auto tile_of = [&](long long it) { return blockIdx.x + it * (long long)gridDim.x; };
  const int U = G + (L - 2) + P;       // units per net and tile

  if (warp == 16) {
    // ------------------------------------------------------------------ weight producer (one thread)
    if (lane == 0) {
      Phases ph;
      int stage = 0;
      long long nfill = 0;
      bool ok = true;
      for (long long it = 0; it < n_my && ok; ++it)
        for (int u = 0; u < U && ok; ++u) {
          const int l = u < G ? 0 : (u < G + L - 2 ? u - G + 1 : L - 1);
          const int c = u < G + L - 2 ? 0 : u - (G + L - 2);
          const int t0 = u < G ? 2 * u : 0;
          const int nt = u < G ? min(2, a.ktiles[0] - 2 * u) : a.ktiles[l];
          for (int n = 0; n < 2 && ok; ++n) {
            if (nfill >= P_STAGES) {
              ok = mbar_wait_sleep(&S->w_empty[stage], ph.get(4 + stage), a.status);
              ph.flip(4 + stage);
              if (!ok) break;
            }
            uint8_t* dst = ring + (size_t)stage * P_STAGE_BYTES;
            mbar_expect_tx(&S->w_full[stage], (uint32_t)nt * P_KT_BYTES);
            for (int t = 0; t < nt; ++t) {
              const long long src = ((long long)c * a.ktiles[l] + t0 + t) * 8192;
              bulk_g2s(dst + (size_t)t * P_KT_BYTES, a.wb[n][0][l] + src, P_TILE_BYTES, &S->w_full[stage]);
              bulk_g2s(dst + (size_t)t * P_KT_BYTES + P_TILE_BYTES, a.wb[n][1][l] + src, P_TILE_BYTES, &S->w_full[stage]);
            }
            ++nfill;
            stage ^= 1;
          }
        }
    }
    __syncwarp();
  } else if (warp == 17) {
    // (no tile I/O warp in this kernel: tiles are accessed in place)
  } else if (warp == 19) {
    // ------------------------------------------------------------------ log-det reducer
    uint32_t ph_r = 0;
    bool ok = true;
    for (long long it = 0; it < n_my && ok; ++it) {
      ok = mbar_wait_sleep(&S->dl_ready, ph_r, a.status);
      ph_r ^= 1;
      const float* dl = &S->dl_part[0][0];
#pragma unroll
      for (int r = lane; r < P_TM; r += 32) {
        const long long row = tile_of(it) * P_TM + r;
        if (row < a.B) {
          const float base_dl = a.dlogp_in ? a.dlogp_in[row] : 0.f;
          a.dlogp_out[row] = base_dl + ((dl[r] + dl[P_TM + r]) + (dl[2 * P_TM + r] + dl[3 * P_TM + r]));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->dl_free);
    }
  } else if (warp == 18) {
    // ------------------------------------------------------------------ MMA issuer (warp-wide, elected lane issues)
    const uint32_t idesc = idesc_bf16(128, 128);
    int stage = 0;
    Phases ph;
    bool ok = true;
    for (long long it = 0; it < n_my && ok; ++it) {
#pragma unroll 1
      for (int u = 0; u < U && ok; ++u) {
        const int K = u < G ? min(128, a.net[0].K[0] - 128 * u) : a.net[0].K[u < G + L - 2 ? u - G + 1 : L - 1];
        const bool wait_a = u <= G + L - 2;                          // every unit but the passes after the first
        const bool wait_e = (u == 0 && it > 0) || u > G + L - 2;     // a last-layer pass preceded on this accumulator
        const bool accum = u > 0 && u < G;
        const int ksteps = (K + 15) / 16;
#pragma unroll 1
        for (int s = 0; s < 2 && ok; ++s) {
          ok = mbar_wait(&S->w_full[stage], ph.get(0 + stage), a.status);
          ph.flip(0 + stage);
          if (wait_a && ok) { ok = mbar_wait(&S->a_ready[s], ph.get(2 + s), a.status); ph.flip(2 + s); }
          if (wait_e && ok) { ok = mbar_wait(&S->acc_empty[s], ph.get(4 + s), a.status); ph.flip(4 + s); }
          if (!ok) break;
          tc_fence_after();
          const uint32_t sb = smem_u32(ring + (size_t)stage * P_STAGE_BYTES);
          const uint32_t acc_addr = tmem + s * P_SLOT + P_ACC;
          uint32_t acc = accum ? 1u : 0u;
#pragma unroll 1
          for (int t = 0; t * 4 < ksteps; ++t) {
            const uint32_t b1 = sb + (uint32_t)t * P_KT_BYTES, b2 = b1 + P_TILE_BYTES;
            const uint64_t d1 = smem_desc_sw128(b1), d2 = smem_desc_sw128(b2);
            const uint32_t a1 = tmem + s * P_SLOT + P_A + (uint32_t)(t * 32), a2 = a1 + P_A_STRIDE;
            const int nk = min(4, ksteps - t * 4);
            if (nk == 4) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                mma3_bf16x3_elect(acc_addr, a1 + ks * 8, a2 + ks * 8, d1 + 2 * ks, d2 + 2 * ks, idesc, ks == 0 ? acc : 1u);
            } else {
#pragma unroll 1
              for (int ks = 0; ks < nk; ++ks)
                mma3_bf16x3_elect(acc_addr, a1 + ks * 8, a2 + ks * 8, d1 + 2 * ks, d2 + 2 * ks, idesc, ks == 0 ? acc : 1u);
            }
            acc = 1;
          }
          mma_commit_elect(&S->acc_full[s]);
          mma_commit_elect(&S->w_empty[stage]);
          stage ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (0..15)
    const int q = warp & 3, j = warp >> 2;
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    Phases ph;
    uint32_t ph_df = 0;

    auto stage_x = [&](long long it, int s, int g) {
      const long long row = tile_of(it) * P_TM + r_in_tile;
      pair_stage_x(a.net[s], a.plain_cond, a.cond + row * (long long)a.K0raw, row < a.B, g, j,
                   tmem + lane_base + s * P_SLOT + P_A);
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->a_ready[s]);
    };
    auto wait_acc = [&](int s) {
      mbar_wait_sleep(&S->acc_full[s], ph.get(6 + s), a.status);
      ph.flip(6 + s);
      tc_fence_after();
    };

    for (long long it = 0; it < n_my; ++it) {
      const long long row = tile_of(it) * P_TM + r_in_tile;
      const bool live = row < a.B;
      if (it == 0)
        for (int s = 0; s < 2; ++s) stage_x(0, s, 0);
      float ld = 0.f;
#pragma unroll 1
      for (int u = 0; u < (G - 1) + (L - 1); ++u) {
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
          wait_acc(s);
          if (u < G - 1) {
            stage_x(it, s, u + 1);
          } else {
            // ---- hidden layer l of net s: columns [32 j, 32 j + 32) -> bias, activation, bf16 split -> A operand
            const int l = u - (G - 1);
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + s * P_SLOT + P_ACC + j * 32, v);
            tmem_ld_wait();
            uint32_t t1[16], t2[16];
            const float4* b4 = reinterpret_cast<const float4*>(bias_h + s * a.hid_bias_floats + l * 128 + j * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = b4[i];
              hidden_pair2<ACT>(v[4 * i], v[4 * i + 1], bb.x, bb.y, t1[2 * i], t2[2 * i]);
              hidden_pair2<ACT>(v[4 * i + 2], v[4 * i + 3], bb.z, bb.w, t1[2 * i + 1], t2[2 * i + 1]);
            }
            const uint32_t acol = tmem + lane_base + s * P_SLOT + P_A + j * 16;
            tmem_st16(acol, t1);
            tmem_st16(acol + P_A_STRIDE, t2);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S->a_ready[s]);
          }
        }
      }
      // ---- last layer: pass c = columns [128 c, 128 c + 128); this warp owns [128 c + 32 j, + 32) of its rows
#pragma unroll 1
      for (int c = 0; c < P; ++c) {
        const int col0 = c * 128 + j * 32;
        const int ncol = min(32, a.D_t - col0);           // <= 0: nothing in this pass for this warp
        uint32_t vm[32], vs[32];
        wait_acc(0);
        tmem_ld32(tmem + lane_base + P_ACC + j * 32, vm);
        wait_acc(1);
        tmem_ld32(tmem + lane_base + P_SLOT + P_ACC + j * 32, vs);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&S->acc_empty[0]);
          mbar_arrive(&S->acc_empty[1]);
        }
        if (ncol > 0 && live) {
          const float* xin = a.tin + row * (long long)a.D_t + col0;
          float* xout = a.tout + row * (long long)a.D_t + col0;
          const float* bm = a.net[0].bias[L - 1] + col0;
          const float* bs = a.net[1].bias[L - 1] + col0;
          auto one = [&](float y, float mu, float sv) -> float {
            const float ls = (1.f - 2.f * rcp_fast(1.f + ex2_fast(2.f * LOG2E * sv))) * a.alpha;
            ld += INVERSE ? -ls : ls;
            return INVERSE ? (y - mu) * ex2_fast(-LOG2E * ls) : fmaf(y, ex2_fast(LOG2E * ls), mu);
          };
          if (ncol == 32 && a.vec_ok) {
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              const float4 y4 = *reinterpret_cast<const float4*>(xin + k);
              const float4 m4 = __ldg(reinterpret_cast<const float4*>(bm + k));
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(bs + k));
              float4 o;
              o.x = one(y4.x, __uint_as_float(vm[k]) + m4.x, __uint_as_float(vs[k]) + s4.x);
              o.y = one(y4.y, __uint_as_float(vm[k + 1]) + m4.y, __uint_as_float(vs[k + 1]) + s4.y);
              o.z = one(y4.z, __uint_as_float(vm[k + 2]) + m4.z, __uint_as_float(vs[k + 2]) + s4.z);
              o.w = one(y4.w, __uint_as_float(vm[k + 3]) + m4.w, __uint_as_float(vs[k + 3]) + s4.w);
              *reinterpret_cast<float4*>(xout + k) = o;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (k < ncol) xout[k] = one(xin[k], __uint_as_float(vm[k]) + __ldg(bm + k), __uint_as_float(vs[k]) + __ldg(bs + k));
          }
        }
      }
      // ---- end of the tile: log-det shares to the reducer, next tile's layer-0 operand for both nets
      if (it > 0) {
        mbar_wait_sleep(&S->dl_free, ph_df, a.status);
        ph_df ^= 1;
      }
      S->dl_part[j][r_in_tile] = ld;
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->dl_ready);
      if (it + 1 < n_my)
        for (int s = 0; s < 2; ++s) stage_x(it + 1, s, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 18) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------ host side

static size_t aff_pair_smem_bytes(const bgx_packed_mlp* net) {
  size_t hid = 0;
  for (int l = 0; l + 1 < net->n_layers; ++l) hid += net->Np[l];
  return 1024 + P_STAGES * P_STAGE_BYTES + sizeof(AffPSmem) + 4 * 2 * hid + 64;
}

bool affine_pair_eligible(const bgx_coupling_io* io, const bgx_packed_mlp* shift, const bgx_packed_mlp* scale, int flags) {
  if (!io || !shift || !scale) return false;
  if (flags & (BGX_FLAG_PRESERVE_VOLUME | BGX_FLAG_CIRCULAR | BGX_FLAG_BF16X6 | BGX_FLAG_NO_PAIR)) return false;
  const int L = shift->n_layers;
  if (L != scale->n_layers || L < 2 || L > 6 || shift->act != scale->act || shift->act < 0 || shift->act > 3) return false;
  if (shift->K[0] != scale->K[0] || shift->raw_width != scale->raw_width) return false;
  if (shift->periodic_scale != scale->periodic_scale || shift->periodic_left != scale->periodic_left) return false;
  for (int l = 0; l + 1 < L; ++l)
    if (shift->N[l] != 128 || scale->N[l] != 128) return false;
  if (shift->N[L - 1] != scale->N[L - 1]) return false;
  for (int l = 0; l < L; ++l)
    for (int t = 0; t < 2; ++t)
      if (!shift->Wb[t][l] || !scale->Wb[t][l]) return false;
  if (io->n_cond != 1 || io->n_tr != 1) return false;
  auto dense = [](const bgx_seg& s) { return s.stride == s.width; };
  if (!dense(io->cond[0]) || !dense(io->tr_in[0]) || !dense(io->tr_out[0])) return false;
  if (io->cond[0].width != shift->raw_width || io->tr_in[0].width != shift->N[L - 1]) return false;
  return aff_pair_smem_bytes(shift) <= 227 * 1024;
}

int affine_coupling_pair(const bgx_coupling_io* io, const bgx_packed_mlp* shift, const bgx_packed_mlp* scale,
                         float log_alpha, int flags, int* status, cudaStream_t st) {
  const int L = shift->n_layers;
  const int d_t = io->tr_in[0].width;
  if (io->tr_out[0].width != d_t || !io->dlogp_out) return BGX_ERR_INVALID;
  if (io->batch == 0) return BGX_OK;
  AffPArgs a{};
  a.B = io->batch;
  a.cond = io->cond[0].ptr; a.tin = io->tr_in[0].ptr; a.tout = const_cast<float*>(io->tr_out[0].ptr);
  a.D_t = d_t; a.K0raw = io->cond[0].width;
  mlp_to_dev(shift, a.net[0]);
  mlp_to_dev(scale, a.net[1]);
  int hid = 0;
  for (int l = 0; l < L; ++l) {
    for (int t = 0; t < 2; ++t) {
      a.wb[0][t][l] = (const uint16_t*)shift->Wb[t][l];
      a.wb[1][t][l] = (const uint16_t*)scale->Wb[t][l];
    }
    a.ktiles[l] = ceil_div(shift->K[l], 64);
    if (l < L - 1) hid += shift->Np[l];
  }
  a.hid_bias_floats = hid;
  a.npass = shift->Np[L - 1] / 128;
  a.G = ceil_div(shift->K[0], 128);
  a.alpha = expf(log_alpha);
  a.plain_cond = (shift->K[0] == shift->raw_width && shift->periodic_scale == 0.f) ? 1 : 0;
  if (a.plain_cond && a.K0raw % 4 == 0 && ((uintptr_t)a.cond & 15) == 0) a.plain_cond = 2;   // vector row reads
  a.vec_ok = (d_t % 4 == 0 && ((uintptr_t)a.tin & 15) == 0 && ((uintptr_t)a.tout & 15) == 0) ? 1 : 0;
  a.dlogp_in = io->dlogp_in;
  a.dlogp_out = io->dlogp_out;
  a.status = status;
  a.ntiles = (a.B + P_TM - 1) / P_TM;
  const size_t smem = aff_pair_smem_bytes(shift);
  int sm_count = 0;
  int rc = device_sm_count(&sm_count);
  if (rc) return rc;
  using KernT = void (*)(const AffPArgs);
  static const KernT kerns[2][4] = {
      {affine_coupling_pair_kernel<false, 0>, affine_coupling_pair_kernel<false, 1>,
       affine_coupling_pair_kernel<false, 2>, affine_coupling_pair_kernel<false, 3>},
      {affine_coupling_pair_kernel<true, 0>, affine_coupling_pair_kernel<true, 1>,
       affine_coupling_pair_kernel<true, 2>, affine_coupling_pair_kernel<true, 3>}};
  const int inv = (flags & BGX_FLAG_INVERSE) ? 1 : 0;
  KernT kern = kerns[inv][shift->act];
  static size_t configured_all[BGX_MAX_DEVICES][2][4] = {};
  auto& configured = configured_all[device_slot()];
  if (smem > configured[inv][shift->act]) {
    rc = check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
    configured[inv][shift->act] = smem;
  }
  const unsigned grid = (unsigned)std::min<long long>(a.ntiles, (long long)sm_count);
  kern<<<grid, P_THREADS, smem, st>>>(a);
  return post_launch();
}

}  // namespace bgx
