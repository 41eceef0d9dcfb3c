// Tensor-core fused RQ-spline coupling block for sm_100a (tcgen05 + TMEM + bulk TMA).
//
// One persistent CTA per SM walks over tiles of 128 samples.  Per tile the whole block runs
// on chip (replaces coupling.py:162-182 + spline.py:87-188 + dense.py:47-48 + nflows' RQS):
//
//   warp 0      bulk-TMA producer: streams pre-swizzled 16 KB weight tiles (hi and lo parts
//               of the 3xTF32 split) from L2 through an 8-stage shared-memory ring
//   warp 1      tcgen05.mma issuer (kind::tf32, M=128 N=128 K=8): A operand = activations in
//               TENSOR MEMORY (hi | lo halves), B operand = weight tiles in shared memory,
//               fp32 accumulators in tensor memory
//   warp 2      TMEM allocation (all 512 columns)
//   warps 4-7   epilogue group 0, warps 8-11 epilogue group 1 (thread <-> sample row):
//               hidden layers: tcgen05.ld accumulator -> bias + activation -> hi/lo split ->
//               tcgen05.st as the next layer's A operand (activations never leave the SM);
//               last layer: 128-column chunks (5 transformed dims x 25 spline parameters)
//               double-buffered in TMEM, groups alternate chunks: parameters go TMEM ->
//               registers -> softmax/cumsum/bin search/RQ evaluation/log-det in registers.
//
// TMEM columns: [0,128) ACC0, [128,256) ACC1, [256,384) A_hi, [384,512) A_lo.
// Numerics: 3xTF32 (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, fp32 accumulate) ~ fp32; BGX_FLAG_TF32X1
// keeps only the first product.
#include "bgx_coupling.cuh"
#include "bgx_tc.cuh"

namespace bgx {
using namespace tc;

constexpr int TC_THREADS = 384;
constexpr int TC_TM = 128;
constexpr int TC_STAGES = 8;
constexpr uint32_t TILE_BYTES = 16384;
constexpr int COL_ACC0 = 0, COL_ACC1 = 128, COL_AHI = 256, COL_ALO = 384;
constexpr int NB = 8;            // spline bins handled by this kernel
constexpr int PS = 3 * NB + 1;   // parameters per transformed dim
constexpr int DPP = 5;           // dims per 128-column chunk

struct TcArgs {
  long long B;
  Segs cond, tin, tout;
  int D_t;
  DevMlp net;
  const float* whi[BGX_MAX_LAYERS];
  const float* wlo[BGX_MAX_LAYERS];
  int ktiles[BGX_MAX_LAYERS];
  int npass;         // 128-column chunks of the last layer
  int x3;            // 1: 3xTF32, 0: 1xTF32
  int inverse;
  SplineParams sp;
  const float* dlogp_in;
  float* dlogp_out;
  int* status;       // device int: set to 1 when a barrier wait timed out
  long long ntiles;
};

struct TcSmem {
  uint64_t full[TC_STAGES];
  uint64_t empty[TC_STAGES];
  uint64_t x_ready;      // 4 arrivals: layer-0 operand staged in TMEM
  uint64_t a_ready;      // 8 arrivals: hidden activations staged in TMEM
  uint64_t acc_full_h;   // hidden-layer accumulator complete
  uint64_t acc_full[2];  // last-layer chunk accumulator complete
  uint64_t acc_empty[2]; // 4 arrivals: chunk accumulator drained into registers
  uint32_t tmem_base;
  uint32_t pad;
  float dl_part[TC_TM];
};

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}

// One (sample, dim) spline evaluation with the 25 parameters in registers (NB = 8 bins).
template <bool ROOT>
__device__ __forceinline__ void rqs_eval_reg(const float (&p)[PS], const SplineParams& sp, float x, float& y,
                                             float& lad) {
  float mw = p[0], mh = p[NB];
#pragma unroll
  for (int k = 1; k < NB; ++k) {
    mw = fmaxf(mw, p[k]);
    mh = fmaxf(mh, p[NB + k]);
  }
  float ew[NB], eh[NB];
  float sw = 0.f, sh = 0.f;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    ew[k] = expf(p[k] - mw);
    eh[k] = expf(p[NB + k] - mh);
    sw += ew[k];
    sh += eh[k];
  }
  const float cw_scale = (1.f - sp.min_w * NB) / sw;
  const float ch_scale = (1.f - sp.min_h * NB) / sh;
  const float wx = sp.right - sp.left, hy = sp.top - sp.bottom;
  float cumw = 0.f, cumh = 0.f;
  float kw_lo = sp.left, kh_lo = sp.bottom;
  float bw_lo = sp.left, bw_hi = sp.right, bh_lo = sp.bottom, bh_hi = sp.top;
  float s0 = p[2 * NB], s1 = p[2 * NB + 1];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    cumw += sp.min_w + cw_scale * ew[k];
    cumh += sp.min_h + ch_scale * eh[k];
    const float kw_hi = (k == NB - 1) ? sp.right : fmaf(wx, cumw, sp.left);
    const float kh_hi = (k == NB - 1) ? sp.top : fmaf(hy, cumh, sp.bottom);
    const float knot = ROOT ? kh_lo : kw_lo;
    if (k == 0 || x >= knot) {
      bw_lo = kw_lo; bw_hi = kw_hi; bh_lo = kh_lo; bh_hi = kh_hi;
      s0 = p[2 * NB + k];
      s1 = p[2 * NB + k + 1];
    }
    kw_lo = kw_hi;
    kh_lo = kh_hi;
  }
  const float w = bw_hi - bw_lo, h = bh_hi - bh_lo;
  const float delta = h / w;
  const float d0 = sp.min_d + softplus_beta(s0, sp.beta, sp.inv_beta);
  const float d1 = sp.min_d + softplus_beta(s1, sp.beta, sp.inv_beta);
  const float s = d0 + d1 - 2.f * delta;
  if (ROOT) {
    const float q = x - bh_lo;
    const float a = q * s + h * (delta - d0);
    const float b = h * d0 - q * s;
    const float c = -delta * q;
    const float disc = fmaxf(b * b - 4.f * a * c, 0.f);
    const float root = (2.f * c) / (-b - sqrtf(disc));
    y = fmaf(root, w, bw_lo);
    const float t1 = root * (1.f - root);
    const float den = delta + s * t1;
    const float omr = 1.f - root;
    const float num = delta * delta * (d1 * root * root + 2.f * delta * t1 + d0 * omr * omr);
    lad = -(logf(num) - 2.f * logf(den));
  } else {
    const float th = (x - bw_lo) / w;
    const float t1 = th * (1.f - th);
    const float den = delta + s * t1;
    y = bh_lo + h * (delta * th * th + d0 * t1) / den;
    const float omt = 1.f - th;
    const float num = delta * delta * (d1 * th * th + 2.f * delta * t1 + d0 * omt * omt);
    lad = logf(num) - 2.f * logf(den);
  }
}

// transform dim `d` of row `row` with the parameters v[0..24] (+ bias already added)
__device__ __forceinline__ float do_dim(const TcArgs& a, const float (&p)[PS], long long row, int d) {
  if (d >= a.D_t || row >= a.B) return 0.f;
  float x = __ldg(seg_addr(a.tin, row, d));
  if (x < a.sp.left || x > a.sp.right) {
    if (a.sp.oob) atomicAdd(a.sp.oob, 1);
    x = fminf(fmaxf(x, a.sp.left), a.sp.right);
  }
  float y, lad;
  if (a.inverse) rqs_eval_reg<false>(p, a.sp, x, y, lad);
  else rqs_eval_reg<true>(p, a.sp, x, y, lad);
  *const_cast<float*>(seg_addr(a.tout, row, d)) = y;
  return lad;
}

__global__ void __launch_bounds__(TC_THREADS, 1) spline_coupling_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = base;                                          // TC_STAGES x 16 KB
  TcSmem* S = (TcSmem*)(base + TC_STAGES * TILE_BYTES);
  float* bias_s = (float*)(S + 1);                               // all layers' biases, concatenated

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.net.n_layers;
  const int nparts = a.x3 ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&S->full[s], 1);
      mbar_init(&S->empty[s], 1);
    }
    mbar_init(&S->x_ready, 4);
    mbar_init(&S->a_ready, 8);
    mbar_init(&S->acc_full_h, 1);
    mbar_init(&S->acc_full[0], 1);
    mbar_init(&S->acc_full[1], 1);
    mbar_init(&S->acc_empty[0], 4);
    mbar_init(&S->acc_empty[1], 4);
    fence_mbar_init();
  }
  {  // biases -> shared memory (layer l at offset boff[l])
    int off = 0;
    for (int l = 0; l < L; ++l) {
      for (int i = threadIdx.x; i < a.net.Np[l]; i += TC_THREADS) bias_s[off + i] = a.net.bias[l][i];
      off += a.net.Np[l];
    }
  }
  if (warp == 2) tmem_alloc<512>(&S->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        for (int l = 0; l < L; ++l) {
          const int nch = (l == L - 1) ? a.npass : 1;
          const int kt = a.ktiles[l];
          for (int c = 0; c < nch; ++c)
            for (int t = 0; t < kt; ++t)
              for (int part = 0; part < nparts; ++part) {
                mbar_wait(&S->empty[stage], phase ^ 1, a.status);
                const float* src = (part == 0 ? a.whi[l] : a.wlo[l]) + ((long long)c * kt + t) * 4096;
                mbar_expect_tx(&S->full[stage], TILE_BYTES);
                bulk_g2s(ring + stage * TILE_BYTES, src, TILE_BYTES, &S->full[stage]);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
              }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(128, 128);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t ph_x = 0, ph_a = 0, ph_e0 = 0, ph_e1 = 0;
      bool first = true;
      for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        for (int l = 0; l < L; ++l) {
          if (l == 0) { mbar_wait(&S->x_ready, ph_x, a.status); ph_x ^= 1; }
          else { mbar_wait(&S->a_ready, ph_a, a.status); ph_a ^= 1; }
          const bool last = (l == L - 1);
          const int nch = last ? a.npass : 1;
          const int kt = a.ktiles[l];
          const int ksteps_total = (a.net.K[l] + 7) / 8;
          for (int c = 0; c < nch; ++c) {
            const int buf = last ? (c & 1) : 0;
            // accumulator free?  (hidden-layer reads are covered by a_ready / x ordering)
            if (buf == 0) {
              const bool need = last ? (c >= 2) : (l == 0 && !first);
              if (need) { mbar_wait(&S->acc_empty[0], ph_e0, a.status); ph_e0 ^= 1; }
            } else {
              const bool need = (c >= 3) || !first;
              if (need) { mbar_wait(&S->acc_empty[1], ph_e1, a.status); ph_e1 ^= 1; }
            }
            tc_fence_after();
            const uint32_t d_tmem = tmem + (buf ? COL_ACC1 : COL_ACC0);
            uint32_t acc = 0;
            for (int t = 0; t < kt; ++t) {
              const int s_hi = stage;
              mbar_wait(&S->full[s_hi], phase, a.status);
              if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
              int s_lo = -1;
              if (nparts == 2) {
                s_lo = stage;
                mbar_wait(&S->full[s_lo], phase, a.status);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
              }
              tc_fence_after();
              const int nk = min(4, ksteps_total - t * 4);
              const uint32_t b_hi = smem_u32(ring + s_hi * TILE_BYTES);
              const uint32_t b_lo = s_lo >= 0 ? smem_u32(ring + s_lo * TILE_BYTES) : 0;
              for (int ks = 0; ks < nk; ++ks) {
                const uint32_t kcol = (uint32_t)(t * 32 + ks * 8);
                const uint64_t dh = smem_desc_sw128(b_hi + ks * 32);
                mma_tf32_ts(d_tmem, tmem + COL_AHI + kcol, dh, idesc, acc);
                acc = 1;
                if (nparts == 2) {
                  mma_tf32_ts(d_tmem, tmem + COL_ALO + kcol, dh, idesc, 1);
                  mma_tf32_ts(d_tmem, tmem + COL_AHI + kcol, smem_desc_sw128(b_lo + ks * 32), idesc, 1);
                }
              }
              mma_commit(&S->empty[s_hi]);
              if (s_lo >= 0) mma_commit(&S->empty[s_lo]);
            }
            mma_commit(last ? &S->acc_full[buf] : &S->acc_full_h);
          }
        }
        first = false;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue groups
    const int g = (warp - 4) >> 2;            // 0 or 1
    const int q = warp & 3;                   // TMEM lane quadrant of this warp
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int gx = (a.npass - 1) & 1;         // the group that owns the last chunk stages x
    uint32_t ph_h = 0, ph_f = 0;
    const int last_off = [&] { int o = 0; for (int l = 0; l < L - 1; ++l) o += a.net.Np[l]; return o; }();

    auto stage_x = [&](long long tile) {
      const long long row = tile * TC_TM + r_in_tile;
      const int K0 = a.net.K[0];
      for (int b0 = 0; b0 < K0; b0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = load_cond(a.cond, a.net, a.B, row, b0 + j);
          split_tf32(v, hi[j], lo[j]);
        }
        tmem_st8(tmem + lane_base + COL_AHI + b0, hi);
        if (a.x3) tmem_st8(tmem + lane_base + COL_ALO + b0, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->x_ready);
    };

    bool first = true;
    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      const long long row = tile * TC_TM + r_in_tile;
      if (first && g == gx) stage_x(tile);
      first = false;
      // ---- hidden layers: ACC0 -> bias + activation -> A operand of the next layer
      int boff = 0;
      for (int l = 0; l < L - 1; ++l) {
        mbar_wait(&S->acc_full_h, ph_h, a.status);
        ph_h ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
          const int col = g * 64 + i * 32;
          uint32_t v[32], hi[32], lo[32];
          tmem_ld32(tmem + lane_base + COL_ACC0 + col, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float h = act_apply(__uint_as_float(v[j]) + bias_s[boff + col + j], a.net.act);
            split_tf32(h, hi[j], lo[j]);
          }
          tmem_st32(tmem + lane_base + COL_AHI + col, hi);
          if (a.x3) tmem_st32(tmem + lane_base + COL_ALO + col, lo);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->a_ready);
        boff += a.net.Np[l];
      }
      // ---- last layer: chunks c = g, g+2, ...  (5 dims x 25 parameters per 128 columns)
      float ld = 0.f;
      for (int c = g; c < a.npass; c += 2) {
        mbar_wait(&S->acc_full[g], ph_f, a.status);
        ph_f ^= 1;
        tc_fence_after();
        const uint32_t acc_addr = tmem + lane_base + (g ? COL_ACC1 : COL_ACC0);
        const float* bl = bias_s + last_off + c * 128;
        uint32_t A[32], Bv[32], C[32], D[32];
        tmem_ld32(acc_addr + 0, A);
        tmem_ld32(acc_addr + 32, Bv);
        tmem_ld_wait();
        float p[PS];
        const int d0 = c * DPP;
#pragma unroll
        for (int j = 0; j < 25; ++j) p[j] = __uint_as_float(A[j]) + bl[j];
        ld += do_dim(a, p, row, d0 + 0);
#pragma unroll
        for (int j = 0; j < 7; ++j) p[j] = __uint_as_float(A[25 + j]) + bl[25 + j];
#pragma unroll
        for (int j = 0; j < 18; ++j) p[7 + j] = __uint_as_float(Bv[j]) + bl[32 + j];
        ld += do_dim(a, p, row, d0 + 1);
        tmem_ld32(acc_addr + 64, C);
        tmem_ld32(acc_addr + 96, D);
        tmem_ld_wait();
        // accumulator is in registers: hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->acc_empty[g]);
        // the group that owns the tile's last chunk stages the next tile's conditioner input:
        // the commit behind acc_full of the last chunk covers every MMA of this tile, so the
        // A-operand columns are free
        if (g == gx && c + 2 >= a.npass) {
          const long long next = tile + gridDim.x;
          if (next < a.ntiles) stage_x(next);
        }
#pragma unroll
        for (int j = 0; j < 14; ++j) p[j] = __uint_as_float(Bv[18 + j]) + bl[50 + j];
#pragma unroll
        for (int j = 0; j < 11; ++j) p[14 + j] = __uint_as_float(C[j]) + bl[64 + j];
        ld += do_dim(a, p, row, d0 + 2);
#pragma unroll
        for (int j = 0; j < 21; ++j) p[j] = __uint_as_float(C[11 + j]) + bl[75 + j];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[21 + j] = __uint_as_float(D[j]) + bl[96 + j];
        ld += do_dim(a, p, row, d0 + 3);
#pragma unroll
        for (int j = 0; j < 25; ++j) p[j] = __uint_as_float(D[4 + j]) + bl[100 + j];
        ld += do_dim(a, p, row, d0 + 4);
      }
      // ---- per-sample log-det: group 1 hands its partial sum to group 0
      if (g == 1) S->dl_part[r_in_tile] = ld;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (g == 0 && row < a.B) {
        const float base_dl = a.dlogp_in ? a.dlogp_in[row] : 0.f;
        a.dlogp_out[row] = base_dl + ld + S->dl_part[r_in_tile];
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");   // dl_part may be overwritten by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------ host side

bool spline_tc_eligible(const bgx_packed_mlp* net, const bgx_spline_cfg* cfg, int d_c) {
  if (!net || !cfg || cfg->n_bins != NB) return false;
  const int L = net->n_layers;
  if (L < 2 || L > 6) return false;
  if (net->K[0] > 128) return false;
  for (int l = 0; l + 1 < L; ++l)
    if (net->N[l] != 128) return false;
  for (int l = 0; l < L; ++l)
    if (!net->Wk_hi[l] || !net->Wk_lo[l]) return false;
  if (net->spline_dims_per_pass != DPP || net->spline_stride != PS) return false;
  return true;
}

int spline_coupling_tc(const bgx_coupling_io* io, const bgx_packed_mlp* net, const bgx_spline_cfg* cfg,
                       int flags, int* status, cudaStream_t st) {
  CouplingArgs ca{};
  int d_c, d_t;
  int rc = coupling_fill_io(io, ca, d_c, d_t);
  if (rc) return rc;
  if (net->raw_width != d_c) return BGX_ERR_INVALID;
  if (net->N[net->n_layers - 1] != ceil_div(d_t, DPP) * 128) return BGX_ERR_INVALID;
  if (ca.B == 0) return BGX_OK;
  TcArgs a{};
  a.B = ca.B;
  a.cond = ca.cond; a.tin = ca.tin; a.tout = ca.tout;
  a.D_t = d_t;
  mlp_to_dev(net, a.net);
  int bias_floats = 0;
  for (int l = 0; l < net->n_layers; ++l) {
    a.whi[l] = net->Wk_hi[l];
    a.wlo[l] = net->Wk_lo[l];
    a.ktiles[l] = ceil_div(net->K[l], 32);
    bias_floats += net->Np[l];
  }
  a.npass = net->N[net->n_layers - 1] / 128;
  a.x3 = (flags & BGX_FLAG_TF32X1) ? 0 : 1;
  a.inverse = (flags & BGX_FLAG_INVERSE) ? 1 : 0;
  spline_params_from_cfg(cfg, a.sp);
  a.dlogp_in = ca.dlogp_in;
  a.dlogp_out = ca.dlogp_out;
  a.status = status;
  a.ntiles = (a.B + TC_TM - 1) / TC_TM;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    rc = check(cudaGetDevice(&dev));
    if (rc) return rc;
    rc = check(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (rc) return rc;
  }
  const size_t smem = 1024 + TC_STAGES * TILE_BYTES + sizeof(TcSmem) + sizeof(float) * bias_floats + 64;
  if (smem > 227 * 1024) return BGX_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > configured) {
    rc = check(cudaFuncSetAttribute(spline_coupling_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
    configured = smem;
  }
  const unsigned grid = (unsigned)std::min<long long>(a.ntiles, sm_count);
  spline_coupling_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a);
  return post_launch();
}

}  // namespace bgx
