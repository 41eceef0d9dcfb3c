// Tensor-core linear layer  Y[B, N] = X[B, K] . W^T + b  with fp32-class accuracy (training path).
//
// The backward of a conditioner (bgflow/nn/dense.py:47-48 differentiated; the reference leaves it to torch autograd =
// fp32 SIMT GEMMs on a GPU, 73 % of a KL training step: profiles/r2_train_profile_fp32.txt) needs the layer GEMMs
// again: the recompute z = h W^T + b and the input gradients dh = g W.  This kernel runs them on tcgen05 with the
// machinery of the pair kernels (bgx_pair.cuh): one persistent CTA per SM, two 128-row tiles ("slots") ping-ponging
// on the tensor pipe and sharing the weight ring, operands split exactly into two bf16 terms (x = x1 + x2,
// W = w1 + w2; products x1 w1 + x2 w1 + x1 w2, fp32 accumulation in tensor memory), the A operand staged by the
// epilogue warps straight from the fp32 rows in global memory, weight tiles by bulk TMA from the pre-swizzled layout of
// bgx_pack_mlp.  K is k-tiled in groups of 128 inputs accumulated in tensor memory (N <= 128), or N runs in 128-column
// passes over one staged operand (K <= 128): the two shapes the conditioner backward needs (dh2 = dP . W2: K = 825,
// N = 128;  P = h2 . W2^T + b2: K = 128, N = 825).  Anything else returns BGX_ERR_UNSUPPORTED (callers use torch).
//
//   warps 0-15 epilogue / operand staging (quadrant w % 4, column share w / 4), 16 weight producer, 17 idle,
//   18 MMA issuer, 19 idle.  TMEM: slot s at 256 s: [0,128) accumulator, [128,192) / [192,256) the two A terms.
#include <cstdlib>

#include "bgx_pair.cuh"

namespace bgx {

struct LinArgs {
  long long B;
  const float* x;     // [B][K], row stride ldx
  float* y;           // [B][N], row stride ldy
  long long ldx, ldy;
  int K, N;           // inputs read, outputs written (N may exceed the layer's width up to its padded width: zeros)
  DevMlp net;         // one layer: K[0], Np[0], bias[0]
  const uint16_t* wb[2];
  int ktiles;
  int G, P;           // 128-input groups, 128-column passes (one of them is 1)
  int* status;
  long long ntiles, npairs;
  int vec_ok, plain;
};

struct alignas(16) LinSmem {
  uint64_t w_full[P_STAGES], w_empty[P_STAGES];
  uint64_t a_ready[2], acc_full[2], acc_empty[2];
  uint32_t tmem_base, pad[3];
};

// A tcgen05.ld / .st thread owns one ROW of the tile (32 consecutive columns of it), so reading or writing global
// memory straight from those registers touches 32 different rows per instruction: half-used sectors and, measured,
// 6 B / cycle / SM.  Every warp therefore transposes its own [32 rows x 32 columns] block through a private patch of
// shared memory (row stride 36 floats: the float4 accesses of both views are bank-conflict free) and talks to global
// memory with lanes along the columns: 4 rows x 128 contiguous bytes per instruction.
constexpr int L_TW_STRIDE = 36;
constexpr int L_TW_FLOATS = 32 * L_TW_STRIDE;

// 16-byte asynchronous copy global -> shared (LDGSTS), zero-filled when !valid
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(valid ? 16 : 0)
               : "memory");
}

// Start fetching rows [row_base, + 32) x inputs [128 g + 32 j, + 32) of x into the warp's transpose patch (vector path
// only).  In the K-split mode a tile's groups are staged one after the other and each staging used to expose the full
// latency of its loads (ncu: 69 % of the cycles without an eligible warp); the patch is free as soon as the previous
// block has been read out of it, so the next block's copies fly while this one is split and stored to tensor memory
// and while the warp waits for the MMAs.
__device__ __forceinline__ bool lin_prefetch(const LinArgs& a, float* tw, long long row_base, int g, int j, int lane) {
  const int kg = 128 * g, kend = min(a.K, kg + 128);
  const int c0 = kg + j * 32;
  if (a.plain != 2 || c0 >= kend) return false;
  const int rr = lane >> 3, c4 = lane & 7;
  const int col = c0 + 4 * c4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = rr + 4 * i;
    const long long row = row_base + r;
    const bool valid = row < a.B && col < kend;
    cp_async16_zfill(tw + r * L_TW_STRIDE + 4 * c4, valid ? a.x + row * a.ldx + col : a.x, valid);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  return true;
}

// rows [row_base, + 32) x inputs [128 g + 32 j, + 32) of x -> this lane's row -> two bf16 terms -> tensor memory.
// `prefetched`: lin_prefetch already started the copies of this block.  Returns whether the NEXT block (has_next: rows
// next_row_base, group next_g) has been started.
__device__ __forceinline__ bool lin_stage(const LinArgs& a, float* tw, long long row_base, int g, int j, int lane,
                                          uint32_t a_col, bool prefetched, bool has_next, long long next_row_base,
                                          int next_g) {
  const int kg = 128 * g, kend = min(a.K, kg + 128);
  const int c0 = kg + j * 32;
  bool started = false;
  if (c0 < kend) {
    const int rr = lane >> 3, c4 = lane & 7;
    if (prefetched) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (a.plain == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rr + 4 * i;
        const long long row = row_base + r;
        const int col = c0 + 4 * c4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < a.B && col < kend) v = __ldg(reinterpret_cast<const float4*>(a.x + row * a.ldx + col));
        *reinterpret_cast<float4*>(tw + r * L_TW_STRIDE + 4 * c4) = v;
      }
    } else {
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const long long row = row_base + r;
        const int col = c0 + lane;
        tw[r * L_TW_STRIDE + lane] = (row < a.B && col < kend) ? __ldg(a.x + row * a.ldx + col) : 0.f;
      }
    }
    __syncwarp();
    float xv[32];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const float4 t = *reinterpret_cast<const float4*>(tw + lane * L_TW_STRIDE + 4 * m);
      xv[4 * m] = t.x; xv[4 * m + 1] = t.y; xv[4 * m + 2] = t.z; xv[4 * m + 3] = t.w;
    }
    __syncwarp();
    if (has_next) started = lin_prefetch(a, tw, next_row_base, next_g, j, lane);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (c0 + 16 * h < kend) {      // a half is written iff the group's k-steps read it
        uint32_t t1[8], t2[8], t3[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_bf16(xv[16 * h + 2 * i], xv[16 * h + 2 * i + 1], 2, t1[i], t2[i], t3[i]);
        const uint32_t col = a_col + (uint32_t)((j * 32 + 16 * h) / 2);
        tmem_st8(col, t1);
        tmem_st8(col + P_A_STRIDE, t2);
      }
    }
  }
  tmem_st_wait();
  tc_fence_before();
  return started;
}

__global__ void __launch_bounds__(P_THREADS, 1) linear_tc_kernel(const __grid_constant__ LinArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = base;
  LinSmem* S = (LinSmem*)(base + P_STAGES * P_STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = a.G, P = a.P;
  const int U = G + P - 1;                 // units per tile: groups accumulate into one pass, or passes over one group
  const long long n_my = (a.npairs > blockIdx.x) ? (a.npairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

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
    fence_mbar_init();
  }
  if (warp == 18) tmem_alloc<512>(&S->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S->tmem_base;
  auto tile_of = [&](long long it, int s) { return 2 * (blockIdx.x + it * (long long)gridDim.x) + s; };
  auto slots_of = [&](long long it) { return tile_of(it, 1) < a.ntiles ? 2 : 1; };

  if (warp == 16) {
    // ------------------------------------------------------------------ weight producer (one thread)
    if (lane == 0) {
      Phases ph;
      int stage = 0;
      long long nfill = 0;
      bool ok = true;
      for (long long it = 0; it < n_my && ok; ++it)
        for (int u = 0; u < U && ok; ++u) {
          const int g = P == 1 ? u : 0, c = P == 1 ? 0 : u;
          const int nt = min(2, a.ktiles - 2 * g);
          if (nfill >= P_STAGES) {
            ok = mbar_wait(&S->w_empty[stage], ph.get(4 + stage), a.status);
            ph.flip(4 + stage);
            if (!ok) break;
          }
          uint8_t* dst = ring + (size_t)stage * P_STAGE_BYTES;
          mbar_expect_tx(&S->w_full[stage], (uint32_t)nt * P_KT_BYTES);
          for (int t = 0; t < nt; ++t) {
            const long long src = ((long long)c * a.ktiles + 2 * g + t) * 8192;
            bulk_g2s(dst + (size_t)t * P_KT_BYTES, a.wb[0] + src, P_TILE_BYTES, &S->w_full[stage]);
            bulk_g2s(dst + (size_t)t * P_KT_BYTES + P_TILE_BYTES, a.wb[1] + src, P_TILE_BYTES, &S->w_full[stage]);
          }
          ++nfill;
          stage ^= 1;
        }
    }
    __syncwarp();
  } else if (warp == 18) {
    // ------------------------------------------------------------------ MMA issuer (warp-wide, elected lane issues)
    const uint32_t idesc = idesc_bf16(128, 128);
    int stage = 0;
    Phases ph;
    bool ok = true;
    for (long long it = 0; it < n_my && ok; ++it) {
      const int nslots = slots_of(it);
#pragma unroll 1
      for (int u = 0; u < U && ok; ++u) {
        const int g = P == 1 ? u : 0;
        const int Kg = min(128, a.K - 128 * g);
        const int ksteps = (Kg + 15) / 16;
        const bool wait_a = P == 1 || u == 0;                  // a freshly staged group (every group; or once per tile)
        const bool wait_e = P == 1 ? (u == 0 && it > 0) : (u > 0 || it > 0);   // the accumulator's previous pass was stored
        const bool accum = P == 1 && u > 0;
        ok = mbar_wait(&S->w_full[stage], ph.get(0 + stage), a.status);
        ph.flip(0 + stage);
        const uint32_t sb = smem_u32(ring + (size_t)stage * P_STAGE_BYTES);
#pragma unroll 1
        for (int s = 0; s < nslots && ok; ++s) {
          if (wait_a) { ok = mbar_wait(&S->a_ready[s], ph.get(2 + s), a.status); ph.flip(2 + s); }
          if (wait_e && ok) { ok = mbar_wait(&S->acc_empty[s], ph.get(4 + s), a.status); ph.flip(4 + s); }
          if (!ok) break;
          tc_fence_after();
          const uint32_t acc_addr = tmem + s * P_SLOT + P_ACC;
          uint32_t acc = accum ? 1u : 0u;
#pragma unroll 1
          for (int t = 0; t * 4 < ksteps; ++t) {
            const uint32_t b1 = sb + (uint32_t)t * P_KT_BYTES, b2 = b1 + P_TILE_BYTES;
            const uint64_t d1 = smem_desc_sw128(b1), d2 = smem_desc_sw128(b2);
            const uint32_t a1 = tmem + s * P_SLOT + P_A + (uint32_t)(t * 32), a2 = a1 + P_A_STRIDE;
            const int nk = min(4, ksteps - t * 4);
#pragma unroll 1
            for (int ks = 0; ks < nk; ++ks)
              mma3_bf16x3_elect(acc_addr, a1 + ks * 8, a2 + ks * 8, d1 + 2 * ks, d2 + 2 * ks, idesc, ks == 0 ? acc : 1u);
            acc = 1;
          }
          mma_commit_elect(&S->acc_full[s]);
        }
        mma_commit_elect(&S->w_empty[stage]);
        stage ^= 1;
      }
    }
    __syncwarp();
  } else if (warp < P_EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue / staging warps (0..15)
    const int q = warp & 3, j = warp >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    Phases ph;
    float* tw = reinterpret_cast<float*>(base + P_STAGES * P_STAGE_BYTES + 256) + warp * L_TW_FLOATS;
    bool pf = false;        // the transpose patch holds (or is receiving) the block of the next stage_x call
    auto stage_x = [&](long long it, int s, int g, bool allow_next = true) {
      // K-split mode: the next staging inside this tile is the other slot's same group, then slot 0's next group
      // (nothing is fetched ahead across a tile boundary: the output store uses the patch in between)
      bool has_next = false;
      int ns = 0, ng = 0;
      if (P == 1 && allow_next) {
        if (s + 1 < slots_of(it)) { has_next = true; ns = s + 1; ng = g; }
        else if (g + 1 < G) { has_next = true; ns = 0; ng = g + 1; }
      }
      pf = lin_stage(a, tw, tile_of(it, s) * P_TM + q * 32, g, j, lane, tmem + lane_base + s * P_SLOT + P_A, pf, has_next,
                     tile_of(it, ns) * P_TM + q * 32, ng);
      __syncwarp();
      if (lane == 0) mbar_arrive(&S->a_ready[s]);
    };
    for (long long it = 0; it < n_my; ++it) {
      const int nslots = slots_of(it);
      if (it == 0)
        for (int s = 0; s < nslots; ++s) stage_x(0, s, 0);
#pragma unroll 1
      for (int u = 0; u < U; ++u) {
#pragma unroll 1
        for (int s = 0; s < nslots; ++s) {
          mbar_wait(&S->acc_full[s], ph.get(6 + s), a.status);
          ph.flip(6 + s);
          tc_fence_after();
          if (P == 1 && u < G - 1) {            // group u consumed: stage the next one (the accumulator keeps summing)
            stage_x(it, s, u + 1);
            continue;
          }
          // ---- a finished pass: columns [128 c + 32 j, + 32) of this thread's row -> + bias -> global
          const int c = P == 1 ? 0 : u;
          uint32_t v[32];
          tmem_ld32(tmem + lane_base + s * P_SLOT + P_ACC + j * 32, v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S->acc_empty[s]);
          // transpose the warp's 32 x 32 block: lanes along the columns for the global stores
#pragma unroll
          for (int m = 0; m < 8; ++m)
            *reinterpret_cast<uint4*>(tw + lane * L_TW_STRIDE + 4 * m) = make_uint4(v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3]);
          __syncwarp();
          const long long row_base = tile_of(it, s) * P_TM + q * 32;
          const int col0 = c * 128 + j * 32;
          if (col0 < a.N) {
            if (a.vec_ok) {
              const int rr = lane >> 3, col = col0 + 4 * (lane & 7);
              if (col < a.N) {                 // N % 4 == 0: the whole float4 is inside
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.net.bias[0] + col));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int r = rr + 4 * i;
                  float4 o = *reinterpret_cast<const float4*>(tw + r * L_TW_STRIDE + 4 * (lane & 7));
                  o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
                  if (row_base + r < a.B) *reinterpret_cast<float4*>(a.y + (row_base + r) * a.ldy + col) = o;
                }
              }
            } else {
              const int col = col0 + lane;
              if (col < a.N) {
                const float bb = __ldg(a.net.bias[0] + col);
#pragma unroll 8
                for (int r = 0; r < 32; ++r)
                  if (row_base + r < a.B) a.y[(row_base + r) * a.ldy + col] = tw[r * L_TW_STRIDE + lane] + bb;
              }
            }
          }
          __syncwarp();
          // the tile's last unit: every MMA that reads its A operand is complete -> stage the next tile's first group
          // (slot 0's staging must not fetch ahead: slot 1's output store, next in this loop, uses the patch)
          if (u == U - 1 && it + 1 < n_my && tile_of(it + 1, s) < a.ntiles) stage_x(it + 1, s, 0, s == nslots - 1);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 18) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace bgx

using namespace bgx;

namespace bgx {
// Y[:, :n_out] = X[:, :K] . W^T + b for the single layer of `net`; x / y row strides ldx / ldy (floats).  n_out may
// exceed the layer's width N up to its padded width (the extra columns are written as zeros: packed rows and bias
// beyond N are zero) — the training drivers keep every width a multiple of 4 floats that way.
int linear_launch(int64_t batch, const float* x, int64_t ldx, const bgx_packed_mlp* net, float* y, int64_t ldy, int n_out,
                  int32_t* status, cudaStream_t stream) {
  if (batch < 0 || !net || net->n_layers != 1 || (batch > 0 && (!x || !y))) return BGX_ERR_INVALID;
  if (!net->Wb[0][0] || !net->Wb[1][0] || net->raw_width != net->K[0]) return BGX_ERR_INVALID;
  const int K = net->K[0], N = n_out;
  if (N < 1 || N > net->Np[0] || ldx < K || ldy < N) return BGX_ERR_INVALID;
  if (batch == 0) return BGX_OK;
  const int G = ceil_div(K, 128), P = ceil_div(N, 128);
  if (G > 1 && P > 1) return BGX_ERR_UNSUPPORTED;
  LinArgs a{};
  a.B = batch;
  a.x = x;
  a.y = y;
  a.ldx = ldx;
  a.ldy = ldy;
  a.K = K;
  a.N = N;
  mlp_to_dev(net, a.net);
  a.wb[0] = (const uint16_t*)net->Wb[0][0];
  a.wb[1] = (const uint16_t*)net->Wb[1][0];
  a.ktiles = ceil_div(K, 64);
  a.G = G;
  a.P = P;
  a.status = status;
  a.ntiles = (batch + P_TM - 1) / P_TM;
  a.npairs = (a.ntiles + 1) / 2;
  a.vec_ok = (N % 4 == 0 && ldy % 4 == 0 && (((uintptr_t)y | (uintptr_t)net->bias[0]) & 15) == 0) ? 1 : 0;
  // vector row reads: float4 loads may run past K inside the row's padding (ldx >= round_up(K, 4): zero weights there)
  a.plain = (ldx % 4 == 0 && ldx >= round_up(K, 4) && ((uintptr_t)x & 15) == 0) ? 2 : 1;
  const size_t smem = 1024 + P_STAGES * P_STAGE_BYTES + 256 + P_EPI_WARPS * L_TW_FLOATS * sizeof(float) + 64;
  static_assert(sizeof(LinSmem) <= 256, "barrier block");
  int sm_count = 0;
  int rc = device_sm_count(&sm_count);
  if (rc) return rc;
  static bool configured_all[BGX_MAX_DEVICES] = {};
  bool& configured = configured_all[device_slot()];
  if (!configured) {
    rc = check(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
    configured = true;
  }
  const unsigned grid = (unsigned)std::min<long long>(a.npairs, (long long)sm_count);
  linear_tc_kernel<<<grid, P_THREADS, smem, stream>>>(a);
  return post_launch();
}
}  // namespace bgx

// Y[B, N] = X[B, K] . W^T + b with W, b = the single layer of `net` (bgx_pack_mlp of a one-layer bgx_mlp, dims = {K, N}).
extern "C" int bgx_linear(int64_t batch, const float* x, const bgx_packed_mlp* net, float* y, int32_t* status,
                          void* stream) {
  if (!net || net->n_layers != 1) return BGX_ERR_INVALID;
  return bgx::linear_launch(batch, x, net->K[0], net, y, net->N[0], net->N[0], status, (cudaStream_t)stream);
}
