// Tensor-core fused RQ-spline coupling block, PAIR kernel: one persistent CTA per SM works on TWO
// 128-row tiles at a time ("slots") that ping-pong on the tensor pipe and share ONE weight ring.
//
// Why (measured on the two-CTAs-per-SM kernel, profiles/r1_spline_tc2_ncu_summary.txt + the ncu source
// page): that kernel streams every weight tile from L2 once per 128-row tile (600 KB per tile, ~22 B per
// cycle per SM = half of the L2 fabric's ~42 B/clk/SM share), a quarter of its executed instructions are
// barrier polling loops of the producer / I/O thread, and inside a CTA the epilogue and the MMAs of a
// tile are serialised.  Here
//   * both slots consume a weight stage before it is released: half the L2 -> SM weight traffic;
//   * the MMA issuer alternates slot 0 / slot 1 unit by unit, so the 16 epilogue warps work on one
//     slot's accumulator while the tensor pipe fills the other one (explicit ping-pong instead of
//     relying on a sibling CTA);
//   * every role blocks on exactly one mbarrier at a time (no polling state machines);
//   * the conditioner's first layer is k-tiled in groups of 128 inputs and the last layer runs any
//     number of 128-column passes, with the transformed tile either resident in shared memory
//     (narrow blocks: D_t + D_c <= ~80, bulk-TMA tile I/O) or read / written in place in global
//     memory (WIDE: BASELINE config 5, D = 384 / 3072).
//
//   warps 0-15  epilogue: warp w owns TMEM lane quadrant w % 4 (rows 32 (w%4) .. +31 of a tile) and the
//               column / dim share w / 4 of every unit
//   warp 16     lane 0: weight producer (bulk TMA, 2 stages x 64 KB = one unit's k-tiles x 2 bf16 terms)
//   warp 17     lane 0: tile I/O of the narrow mode (one bulk copy per tile and direction)
//   warp 18     tcgen05.mma issuer (warp-wide loop, one elected lane issues); owns the TMEM allocation
//
// TMEM (512 columns): slot s at 256 s: [0,128) accumulator, [128,192) A term 1, [192,256) A term 2.
// A "unit" is one accumulator's worth of MMAs: a 128-input group of layer 0, a hidden layer, or one
// 128-column pass of the last layer (5 dims x 25 spline parameters).
//
// Reference arithmetic: bgflow/nn/flow/coupling.py:162-182, bgflow/nn/dense.py:47-48,
// bgflow/nn/flow/transformer/spline.py:87-188 + nflows rational_quadratic_spline.
#include <cstdlib>

#include "bgx_coupling.cuh"
#include "bgx_tc.cuh"
#include "bgx_tc_epi.cuh"
#include "bgx_spline_reg.cuh"
#include "bgx_spline_reg2.cuh"
#include "bgx_pair.cuh"

namespace bgx {
using namespace tc;
unsigned long long* tc_get_trace(int* cap);

// Timing experiments (BGX_PAIR_DEBUG) and per-role wait accounting (bgx_debug_set_trace) are compiled in only with
// -DBGX_PAIR_INSTRUMENT=1: the extra predicates and clock reads cost 10 % of the kernel's time (measured).
#ifndef BGX_PAIR_INSTRUMENT
#define BGX_PAIR_INSTRUMENT 0
#endif
constexpr bool P_INSTR = BGX_PAIR_INSTRUMENT != 0;

struct PArgs {
  long long B;
  const float* cond;    // [B][K0raw] dense
  const float* tin;     // [B][D_t] dense
  float* tout;          // [B][D_t] dense
  int D_t, K0raw;
  int role_stride;      // role offset between the two slots (see the role comment in the epilogue)
  DevMlp net;
  const uint16_t* wb[2][BGX_MAX_LAYERS];
  int ktiles[BGX_MAX_LAYERS];
  int npass, G;         // last-layer passes, 128-input groups of layer 0
  SplineK ck;
  int* oob;
  const float* dlogp_in;
  float* dlogp_out;
  int* status;
  long long ntiles, npairs;
  int hid_bias_floats, last_bias_floats;
  const float* bias_last;     // [npass][5][28] (global)
  int plain_cond;             // conditioner input map is the identity (no WrapPeriodic)
  int cs;                     // CTAs per cluster (1, 2 or 4): every weight stage is fetched once per cluster (TMA multicast)
  unsigned long long* trace;  // bgx_debug_set_trace: CTA 0 writes per-role wait-time totals (cycles), >= 16 entries
  int debug;                  // BGX_PAIR_DEBUG (timing experiments only): 1 = no MMAs issued, 2 = no spline evaluation, 4 = no hidden-layer math
};

struct alignas(16) PSmem {
  uint64_t w_full[P_STAGES], w_empty[P_STAGES];
  uint64_t a_ready[2];     // 16: the slot's A operand is staged in TMEM
  uint64_t acc_full[2];    // 1 (tcgen05.commit): the slot's accumulator holds a finished unit
  uint64_t acc_empty[2];   // 16: last-layer pass pulled into registers
  uint64_t y_full[2], c_full[2];   // 1 + tx (narrow mode)
  uint64_t y_done[2], c_free[2];   // 16
  uint64_t dl_ready[2];    // 16: the four dim shares of every row's log-det are written
  uint64_t dl_free[2];     // 1: ... and summed (the buffer may be rewritten)
  uint32_t tmem_base, pad[3];
  float dl_part[2][12][P_TM];      // [slot][log-det share][row]: EPW 4: (pass mod 4, role) cells, EPW 6: one per warp
};

// EPW = epilogue warps per TMEM lane quadrant.  4: 16 epilogue warps at <= 96 registers, a quadrant's five dims of a pass
// go to its warps as {2 packed, 2 packed, 1, -}.  6: 24 epilogue warps at <= 72 registers, one dim per warp per pass
// {1, 1, 1, 1, 1, -}: a shorter critical path per pass and more warps per scheduler to hide the evaluation's latencies
// (the epilogue is bound by per-warp instruction latency, not by issue slots: profiles/README.md).
template <bool INVERSE, int ACT, bool WIDE, int EPW>
__global__ void __launch_bounds__((4 * EPW + 4) * 32, 1) spline_coupling_pair_kernel(const __grid_constant__ PArgs a) {
  constexpr int NW = 4 * EPW;                  // epilogue warps
  constexpr int NTHREADS = (NW + 4) * 32;
  constexpr int W_PROD = NW, W_IO = NW + 1, W_MMA = NW + 2, W_RED = NW + 3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = base;
  PSmem* S = (PSmem*)(base + P_STAGES * P_STAGE_BYTES);
  float* bias_h = (float*)(S + 1);                          // hidden layers' biases
  float* bias_l = bias_h + a.hid_bias_floats;               // narrow: last layer [npass][5][28]
  float* ybuf = bias_l + (WIDE ? 0 : a.last_bias_floats);   // narrow: [2][128][D_t]
  float* cbuf = ybuf + (WIDE ? 0 : 2 * P_TM * a.D_t);       // narrow: [2][128][K0raw]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.net.n_layers;
  const int G = a.G, P = a.npass;
  // Work distribution: cluster `cid` takes `cs` consecutive tile pairs per iteration, one per CTA; every CTA of a cluster
  // runs the same number of iterations (a CTA whose pair does not exist still takes part in the weight traffic)
  const int cs = a.cs;
  const uint32_t crank = cs > 1 ? cluster_ctarank() : 0;
  const long long ncl = gridDim.x / cs, cid = blockIdx.x / cs;
  const long long per_it = ncl * cs;
  const long long n_my = (a.npairs + per_it - 1) / per_it;
  const uint16_t cmask = (uint16_t)((1u << cs) - 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < P_STAGES; ++i) {
      mbar_init(&S->w_full[i], 1);
      mbar_init(&S->w_empty[i], (uint32_t)cs);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&S->a_ready[s], NW);
      mbar_init(&S->acc_full[s], 1);
      mbar_init(&S->acc_empty[s], NW);
      mbar_init(&S->y_full[s], 1);
      mbar_init(&S->c_full[s], 1);
      mbar_init(&S->y_done[s], NW);
      mbar_init(&S->c_free[s], NW);
      mbar_init(&S->dl_ready[s], NW);
      mbar_init(&S->dl_free[s], 1);
    }
    fence_mbar_init();
  }
  {
    int off = 0;
    for (int l = 0; l < L - 1; ++l) {
      for (int i = threadIdx.x; i < a.net.Np[l]; i += NTHREADS) bias_h[off + i] = a.net.bias[l][i];
      off += a.net.Np[l];
    }
    if (!WIDE)
      for (int i = threadIdx.x; i < a.last_bias_floats; i += NTHREADS) bias_l[i] = a.bias_last[i];
    // log-det cells of pass residues a narrow block never reaches (fewer than four passes) stay zero for good
    for (int i = threadIdx.x; i < 2 * 12 * P_TM; i += NTHREADS) (&S->dl_part[0][0][0])[i] = 0.f;
  }
  if (warp == W_MMA) tmem_alloc<512>(&S->tmem_base);
  tc_fence_before();
  if (cs > 1) cluster_sync_all();     // every CTA's barriers exist before a peer arrives on them / multicasts into them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S->tmem_base;

  // tile of slot s in CTA-local iteration `it`; a slot is active iff its tile exists
  auto tile_of = [&](long long it, int s) { return 2 * ((cid + it * ncl) * cs + crank) + s; };
  auto slots_of = [&](long long it) { return tile_of(it, 0) >= a.ntiles ? 0 : (tile_of(it, 1) < a.ntiles ? 2 : 1); };
  auto rows_of = [&](long long tile) { return (int)min((long long)P_TM, a.B - tile * P_TM); };

  if (warp == W_PROD) {
    // ------------------------------------------------------------------ weight producer (one thread)
    if (lane == 0) {
      Phases ph;
      int stage = 0;
      long long nfill = 0, t_prod = 0;
      const long long t_begin = P_INSTR ? clock64() : 0;
      auto fill = [&](int l, int c, int t0, int nt) -> bool {
        if (nfill >= P_STAGES) {
          const long long c0 = P_INSTR ? clock64() : 0;
          if (!mbar_wait_sleep(&S->w_empty[stage], ph.get(4 + stage), a.status)) return false;
          if (P_INSTR) t_prod += clock64() - c0;
          ph.flip(4 + stage);
        }
        uint8_t* dst = ring + (size_t)stage * P_STAGE_BYTES;
        mbar_expect_tx(&S->w_full[stage], (uint32_t)nt * P_KT_BYTES);
        const int kt = a.ktiles[l];
        for (int t = 0; t < nt; ++t) {
          const long long src = ((long long)c * kt + t0 + t) * 8192;
          for (int term = 0; term < 2; ++term) {
            uint8_t* d = dst + (size_t)t * P_KT_BYTES + (size_t)term * P_TILE_BYTES;
            if (cs == 1) bulk_g2s(d, a.wb[term][l] + src, P_TILE_BYTES, &S->w_full[stage]);
            else if ((uint32_t)((2 * t + term) % cs) == crank)     // this CTA's share, delivered to the whole cluster
              bulk_g2s_multicast(d, a.wb[term][l] + src, P_TILE_BYTES, &S->w_full[stage], cmask);
          }
        }
        ++nfill;
        stage ^= 1;
        return true;
      };
      bool ok = true;
      for (long long it = 0; it < n_my && ok; ++it) {
        for (int g = 0; g < G && ok; ++g) ok = fill(0, 0, 2 * g, min(2, a.ktiles[0] - 2 * g));
        for (int l = 1; l < L - 1 && ok; ++l) ok = fill(l, 0, 0, a.ktiles[l]);
        for (int c = 0; c < P && ok; ++c) ok = fill(L - 1, c, 0, a.ktiles[L - 1]);
      }
      // drain: the last stages' release arrivals (from every CTA of the cluster) have landed before this CTA may leave
      for (int k = 0; k < P_STAGES && k < nfill && ok; ++k) {
        ok = mbar_wait_sleep(&S->w_empty[stage], ph.get(4 + stage), a.status);
        ph.flip(4 + stage);
        stage ^= 1;
      }
      if (P_INSTR && a.trace && blockIdx.x == 0) {
        a.trace[12] = (unsigned long long)(clock64() - t_begin);  // producer: whole loop
        a.trace[13] = (unsigned long long)t_prod;                 //   waiting for a free stage
      }
    }
    __syncwarp();
  } else if (warp == W_IO) {
    // ------------------------------------------------------------------ tile I/O, narrow mode (one thread)
    if (!WIDE && lane == 0) {
      const uint32_t ystride = (uint32_t)(P_TM * a.D_t), cstride = (uint32_t)(P_TM * a.K0raw);
      auto load_c = [&](long long it, int s) {
        const long long tile = tile_of(it, s);
        if (tile >= a.ntiles) return;
        const uint32_t nb = (uint32_t)(rows_of(tile) * a.K0raw * 4);
        mbar_expect_tx(&S->c_full[s], nb);
        bulk_g2s(cbuf + s * cstride, a.cond + tile * P_TM * (long long)a.K0raw, nb, &S->c_full[s]);
      };
      auto load_y = [&](long long it, int s) {
        const long long tile = tile_of(it, s);
        if (tile >= a.ntiles) return;
        const uint32_t nb = (uint32_t)(rows_of(tile) * a.D_t * 4);
        mbar_expect_tx(&S->y_full[s], nb);
        bulk_g2s(ybuf + s * ystride, a.tin + tile * P_TM * (long long)a.D_t, nb, &S->y_full[s]);
      };
      Phases ph;
      if (n_my > 0)
        for (int s = 0; s < 2; ++s) { load_c(0, s); load_y(0, s); }
      bool ok = true;
      for (long long k = 0; k < n_my && ok; ++k) {
        // conditioner tile of iteration k+1: the buffer is free once iteration k's operand is staged
        for (int s = 0; s < 2 && ok; ++s) {
          if (tile_of(k, s) >= a.ntiles || k + 1 >= n_my) continue;
          ok = mbar_wait_sleep(&S->c_free[s], ph.get(14 + s), a.status);
          ph.flip(14 + s);
          load_c(k + 1, s);
        }
        // transformed tile: store iteration k's result, then fetch iteration k+1's input
        for (int s = 0; s < 2 && ok; ++s) {
          const long long tile = tile_of(k, s);
          if (tile >= a.ntiles) continue;
          ok = mbar_wait_sleep(&S->y_done[s], ph.get(16 + s), a.status);
          ph.flip(16 + s);
          if (!ok) break;
          bulk_s2g(a.tout + tile * P_TM * (long long)a.D_t, ybuf + s * ystride, (uint32_t)(rows_of(tile) * a.D_t * 4));
          bulk_store_wait_read();
          if (k + 1 < n_my) load_y(k + 1, s);
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else if (warp == W_RED) {
    // ------------------------------------------------------------------ log-det reducer: sums the four dim shares
    // of every row in a fixed order (deterministic) and adds the running dlogp; keeps that wait off the epilogue warps
    Phases ph;
    bool ok = true;
    for (long long it = 0; it < n_my && ok; ++it)
      for (int s = 0; s < 2 && ok; ++s) {
        const long long tile = tile_of(it, s);
        if (tile >= a.ntiles) continue;
        ok = mbar_wait_sleep(&S->dl_ready[s], ph.get(18 + s), a.status);
        ph.flip(18 + s);
        const float* dl = &S->dl_part[s][0][0];
#pragma unroll
        for (int r = lane; r < P_TM; r += 32) {
          const long long row = tile * P_TM + r;
          if (row < a.B) {
            const float base_dl = a.dlogp_in ? a.dlogp_in[row] : 0.f;
            float sum = 0.f;
#pragma unroll
            for (int cell = 0; cell < (EPW == 4 ? 12 : 6); ++cell) sum += dl[cell * P_TM + r];     // fixed order
            a.dlogp_out[row] = base_dl + sum;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->dl_free[s]);
      }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer (warp-wide, elected lane issues)
    const uint32_t idesc = idesc_bf16(128, 128);
    int stage = 0;
    Phases ph;
    bool ok = true;
    // one unit on both slots: `wait_a` = a freshly staged A operand is needed, `wait_e` = the accumulator must
    // have been pulled by the epilogue (a last-layer pass preceded), `accum` = keep the accumulator (layer-0 group > 0)
    const bool tr = P_INSTR && a.trace && blockIdx.x == 0;
    long long t_w = 0, t_a = 0, t_e = 0;
    const long long t_begin = P_INSTR ? clock64() : 0;
    auto unit = [&](int nslots, int K, bool wait_a, bool wait_e, bool accum) {
      if (!ok) return;
      long long c0 = tr ? clock64() : 0;
      ok = mbar_wait(&S->w_full[stage], ph.get(0 + stage), a.status);
      if (tr) t_w += clock64() - c0;
      ph.flip(0 + stage);
      const uint32_t sb = smem_u32(ring + (size_t)stage * P_STAGE_BYTES);
      const int ksteps = (K + 15) / 16;
#pragma unroll 1
      for (int s = 0; s < nslots && ok; ++s) {
        long long c1 = tr ? clock64() : 0;
        if (wait_a) { ok = mbar_wait(&S->a_ready[s], ph.get(2 + s), a.status); ph.flip(2 + s); }
        long long c2 = tr ? clock64() : 0;
        if (wait_e && ok) { ok = mbar_wait(&S->acc_empty[s], ph.get(4 + s), a.status); ph.flip(4 + s); }
        if (tr) { t_a += c2 - c1; t_e += clock64() - c2; }
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
          if (P_INSTR && (a.debug & 1)) {
          } else if (nk == 4) {
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
      }
      // both slots' MMAs on this stage are issued: release it (in every CTA of the cluster) when they finish
      if (cs == 1) mma_commit_elect(&S->w_empty[stage]);
      else mma_commit_elect_multicast(&S->w_empty[stage], cmask);
      stage ^= 1;
    };
    const int U = G + (L - 2) + P;
    for (long long it = 0; it < n_my && ok; ++it) {
      const int nslots = slots_of(it);
#pragma unroll 1
      for (int u = 0; u < U; ++u) {
        if (u < G) unit(nslots, min(128, a.net.K[0] - 128 * u), true, u == 0 && it > 0, u > 0);
        else if (u < G + L - 2) unit(nslots, a.net.K[u - G + 1], true, false, false);
        else unit(nslots, a.net.K[L - 1], u == G + L - 2, u > G + L - 2, false);
      }
    }
    if (tr && lane == 0) {
      a.trace[0] = (unsigned long long)(clock64() - t_begin);     // MMA warp: whole loop
      a.trace[1] = (unsigned long long)t_w;                       //   waiting for weights
      a.trace[2] = (unsigned long long)t_a;                       //   waiting for a staged A operand
      a.trace[3] = (unsigned long long)t_e;                       //   waiting for a pulled accumulator
      a.trace[4] = (unsigned long long)n_my;
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (0..15)
    const int q = warp & 3, j = warp >> 2;            // TMEM lane quadrant, column / dim share
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    Phases ph;
    bool ok = true;

    auto stage_x = [&](long long it, int s, int g) {
      const long long row = tile_of(it, s) * P_TM + r_in_tile;
      if (!WIDE && g == 0) {
        ok = ok && mbar_wait_sleep(&S->c_full[s], ph.get(8 + s), a.status);
        ph.flip(8 + s);
      }
      const float* crow = WIDE ? a.cond + row * (long long)a.K0raw : cbuf + (s * P_TM + r_in_tile) * a.K0raw;
      if (j < 4) pair_stage_x(a.net, a.plain_cond, crow, !WIDE || row < a.B, g, j, tmem + lane_base + s * P_SLOT + P_A);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&S->a_ready[s]);
        if (!WIDE && g == 0) mbar_arrive(&S->c_free[s]);
      }
    };
    const bool tr = P_INSTR && a.trace && blockIdx.x == 0 && warp == 0;
    long long t_acc = 0;
    const long long t_begin = P_INSTR ? clock64() : 0;
    auto wait_acc = [&](int s) {
      const long long c0 = tr ? clock64() : 0;
      ok = ok && mbar_wait_sleep(&S->acc_full[s], ph.get(6 + s), a.status);
      if (tr) t_acc += clock64() - c0;
      ph.flip(6 + s);
      tc_fence_after();
    };

    float ld[2][3];      // log-det shares of this thread: [slot][role] (EPW 6: [slot][0])
    int n_oob = 0;
    const int U = (G - 1) + (L - 1) + P;      // staging events of layer-0 groups 1.., hidden layers, last-layer passes
    for (long long it = 0; it < n_my; ++it) {
      const int nslots = slots_of(it);
      if (it == 0)
        for (int s = 0; s < nslots; ++s) stage_x(0, s, 0);
      for (int s = 0; s < 2; ++s) ld[s][0] = ld[s][1] = ld[s][2] = 0.f;
#pragma unroll 1
      for (int u = 0; u < U; ++u) {
#pragma unroll 1
        for (int s = 0; s < nslots; ++s) {
          if (u < G - 1) {
            // ---- layer 0, group u done (its commit frees the A operand): stage group u + 1
            wait_acc(s);
            stage_x(it, s, u + 1);
          } else if (u < G + L - 2) {
            // ---- hidden layer l: accumulator columns [32 j, 32 j + 32) -> bias, activation, exact bf16 split -> A
            const int l = u - (G - 1);
            wait_acc(s);
            if (j >= 4) {            // (EPW = 6) the hidden layers' 128 columns go to four warps per quadrant
              if (lane == 0) mbar_arrive(&S->a_ready[s]);
              continue;
            }
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + s * P_SLOT + P_ACC + j * 32, v);
            tmem_ld_wait();
            uint32_t t1[16], t2[16];
            const float4* b4 = reinterpret_cast<const float4*>(bias_h + l * 128 + j * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = b4[i];
              if (P_INSTR && (a.debug & 4)) {
                t1[2 * i] = v[4 * i]; t2[2 * i] = v[4 * i + 1]; t1[2 * i + 1] = v[4 * i + 2]; t2[2 * i + 1] = v[4 * i + 3];
                continue;
              }
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
          } else {
            // ---- last layer: pass c holds dims 5c .. 5c+4.  Per (pass, slot) the four warps of a quadrant take the
            // roles {dims 0,1 packed | dims 2,3 packed | dim 4 | nothing}, rotating with c so the load evens out;
            // a pair of dims is evaluated in packed fp32 lanes (bgx_spline_reg2.cuh)
            const int c = u - (G + L - 2);
            // The roles rotate with the pass AND the slot (a warp that idles or has one dim on slot 0's pass has a pair on slot 1's).  A row's
            // log-det must not depend on the slot it lands in, bit for bit (row independence is tested at full size), so the
            // shares are kept per (pass mod 4, role) cell — the same grouping and the same order of additions whichever warp
            // serves it — and the reducer adds the twelve cells in a fixed order.
            const int role = EPW == 4 ? ((j + c + a.role_stride * s) & 3) : ((j + c) % 6);
            const long long row = tile_of(it, s) * P_TM + r_in_tile;
            const bool live = row < a.B;
            if (!WIDE && c == 0) {
              ok = ok && mbar_wait_sleep(&S->y_full[s], ph.get(10 + s), a.status);
              ph.flip(10 + s);
            }
            const int iA = EPW == 4 ? (role == 2 ? 4 : 2 * role) : role;   // first dim of this warp inside the pass
            const int dA = P_DPP * c + iA;
            const bool hasA = (EPW == 4 ? role < 3 : role < 5) && dA < a.D_t;
            const bool hasB = EPW == 4 && role < 2 && dA + 1 < a.D_t;
            float* yrow = WIDE ? a.tout + row * (long long)a.D_t : ybuf + (s * P_TM + r_in_tile) * a.D_t;
            // the transformed inputs do not depend on the accumulator: fetch them before waiting for it
            float xA = 0.f, xB = 0.f;
            if (WIDE) {
              if (live && hasA) xA = a.tin[row * (long long)a.D_t + dA];
              if (live && hasB) xB = a.tin[row * (long long)a.D_t + dA + 1];
            } else {
              if (hasA) xA = yrow[dA];
              if (hasB) xB = yrow[dA + 1];
            }
            const float* bsrc = (WIDE ? a.bias_last : bias_l) + ((size_t)c * P_DPP + iA) * P_BPAD;
            // EPW 4: the log-det share of (pass, role) is accumulated straight in its canonical shared-memory cell
            // (first touch of a tile = plain store): no per-thread accumulators to carry through the unit loop
            float* cell = &S->dl_part[s][3 * (c & 3) + (role < 3 ? role : 0)][r_in_tile];
            if (EPW == 4 && c == 0 && it > 0) {      // the reducer has summed this slot's previous tile
              ok = ok && mbar_wait_sleep(&S->dl_free[s], ph.get(12 + s), a.status);
              ph.flip(12 + s);
            }
            wait_acc(s);
            const uint32_t acc_addr = tmem + lane_base + s * P_SLOT + P_ACC + iA * P_PS;
            auto release = [&]() {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&S->acc_empty[s]);
            };
            n_oob += (live && hasA && (xA < a.ck.left || xA > a.ck.right)) ? 1 : 0;
            n_oob += (live && hasB && (xB < a.ck.left || xB > a.ck.right)) ? 1 : 0;
            xA = fminf(fmaxf(xA, a.ck.left), a.ck.right);
            xB = fminf(fmaxf(xB, a.ck.left), a.ck.right);
            if (hasB) {
              uint32_t va[25], vb[25];
              tmem_ld25(acc_addr, va);
              tmem_ld25(acc_addr + P_PS, vb);
              tmem_ld_wait();
              release();
              F2 p2[P_PS];
              const float4* bA = reinterpret_cast<const float4*>(bsrc);
              const float4* bB = reinterpret_cast<const float4*>(bsrc + P_BPAD);
#pragma unroll
              for (int qq = 0; qq < 6; ++qq) {
                const float4 x4 = WIDE ? __ldg(bA + qq) : bA[qq], y4 = WIDE ? __ldg(bB + qq) : bB[qq];
                p2[4 * qq] = f2(__uint_as_float(va[4 * qq]) + x4.x, __uint_as_float(vb[4 * qq]) + y4.x);
                p2[4 * qq + 1] = f2(__uint_as_float(va[4 * qq + 1]) + x4.y, __uint_as_float(vb[4 * qq + 1]) + y4.y);
                p2[4 * qq + 2] = f2(__uint_as_float(va[4 * qq + 2]) + x4.z, __uint_as_float(vb[4 * qq + 2]) + y4.z);
                p2[4 * qq + 3] = f2(__uint_as_float(va[4 * qq + 3]) + x4.w, __uint_as_float(vb[4 * qq + 3]) + y4.w);
              }
              p2[24] = f2(__uint_as_float(va[24]) + (WIDE ? __ldg(bsrc + 24) : bsrc[24]),
                          __uint_as_float(vb[24]) + (WIDE ? __ldg(bsrc + P_BPAD + 24) : bsrc[P_BPAD + 24]));
              F2 y2 = f2(xA, xB), l2 = f2(0.f, 0.f);
              if (!(P_INSTR && (a.debug & 2))) rqs_eval_reg2<!INVERSE>(p2, a.ck, f2(xA, xB), y2, l2);
              if (!WIDE || live) {
                yrow[dA] = lo(y2);
                yrow[dA + 1] = hi(y2);
              }
              const float l_pair = lo(l2) + hi(l2);
              if (EPW == 6) ld[s][0] += l_pair;
              else *cell = c < 4 ? l_pair : *cell + l_pair;
            } else if (hasA) {
              uint32_t va[25];
              tmem_ld25(acc_addr, va);
              tmem_ld_wait();
              release();
              float pp[P_PS];
              const float4* bA = reinterpret_cast<const float4*>(bsrc);
#pragma unroll
              for (int qq = 0; qq < 6; ++qq) {
                const float4 x4 = WIDE ? __ldg(bA + qq) : bA[qq];
                pp[4 * qq] = __uint_as_float(va[4 * qq]) + x4.x;
                pp[4 * qq + 1] = __uint_as_float(va[4 * qq + 1]) + x4.y;
                pp[4 * qq + 2] = __uint_as_float(va[4 * qq + 2]) + x4.z;
                pp[4 * qq + 3] = __uint_as_float(va[4 * qq + 3]) + x4.w;
              }
              pp[24] = __uint_as_float(va[24]) + (WIDE ? __ldg(bsrc + 24) : bsrc[24]);
              float y = xA, lad = 0.f;
              if (!(P_INSTR && (a.debug & 2))) rqs_eval_reg<!INVERSE, true>(pp, a.ck, xA, y, lad);
              if (!WIDE || live) yrow[dA] = y;
              if (EPW == 6) ld[s][0] += lad;      // (roles 0 / 1 land here on a last pass with an odd dim count)
              else *cell = c < 4 ? lad : *cell + lad;
            } else {
              release();
              if (EPW == 4 && role < 3 && c < 4) *cell = 0.f;      // a role without dims on this pass: clear the cell
            }
            if (c == P - 1) {
              // ---- end of this slot's tile: hand the output tile to the I/O thread and the log-det shares to the
              // reducer, then stage the next tile's layer-0 operand (every MMA of this tile is complete)
              if (!WIDE) {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&S->y_done[s]);
              }
              if (EPW == 6) {
                if (it > 0) {
                  ok = ok && mbar_wait_sleep(&S->dl_free[s], ph.get(12 + s), a.status);
                  ph.flip(12 + s);
                }
                S->dl_part[s][j][r_in_tile] = ld[s][0];
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(&S->dl_ready[s]);
              if (it + 1 < n_my && tile_of(it + 1, s) < a.ntiles) stage_x(it + 1, s, 0);
            }
          }
        }
      }
    }
    if (n_oob && a.oob) atomicAdd(a.oob, n_oob);
    if (tr && lane == 0) {
      a.trace[8] = (unsigned long long)(clock64() - t_begin);     // epilogue warp 0: whole loop
      a.trace[9] = (unsigned long long)t_acc;                     //   waiting for accumulators
    }
  }
  tc_fence_before();
  if (cs > 1) cluster_sync_all();     // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  else __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------ host side

static size_t pair_smem_bytes(const bgx_packed_mlp* net, int d_t, int k0raw, bool wide) {
  const int L = net->n_layers;
  size_t hid = 0;
  for (int l = 0; l + 1 < L; ++l) hid += net->Np[l];
  const size_t last = wide ? 0 : (size_t)(net->Np[L - 1] / 128) * P_DPP * P_BPAD;
  const size_t tiles = wide ? 0 : (size_t)2 * P_TM * (d_t + k0raw);
  return 1024 + P_STAGES * P_STAGE_BYTES + sizeof(PSmem) + 4 * (hid + last + tiles) + 64;
}

static int pair_mode() {   // BGX_PAIR: 0 = off, 1 = on (default), 2 = wide mode even where the narrow one fits
  static const int m = [] { const char* e = getenv("BGX_PAIR"); return e ? atoi(e) : 1; }();
  return m;
}

// 0 = not eligible, 1 = narrow (tiles in shared memory), 2 = wide (in-place global access)
int spline_pair_eligible(const bgx_coupling_io* io, const bgx_packed_mlp* net, const bgx_spline_cfg* cfg, int flags) {
  if (!io || !net || !cfg || cfg->n_bins != P_NB || (flags & BGX_FLAG_BF16X6) || !pair_mode()) return 0;
  const int L = net->n_layers;
  if (L < 2 || L > 6) return 0;
  for (int l = 0; l + 1 < L; ++l)
    if (net->N[l] != 128) return 0;
  for (int l = 0; l < L; ++l)
    if (!net->Wb[0][l] || !net->Wb[1][l]) return 0;
  if (net->spline_dims_per_pass != P_DPP || net->spline_stride != P_PS || net->spline_bias_pad != P_BPAD ||
      !net->spline_bias)
    return 0;
  if (io->n_cond != 1 || io->n_tr != 1) return 0;
  auto dense = [](const bgx_seg& s) { return s.stride == s.width; };
  if (!dense(io->cond[0]) || !dense(io->tr_in[0]) || !dense(io->tr_out[0])) return 0;
  if (io->cond[0].width != net->raw_width) return 0;
  auto al16 = [](const bgx_seg& s) { return ((uintptr_t)s.ptr & 15) == 0; };
  const bool narrow_ok = net->K[0] <= 128 && io->batch % 4 == 0 && al16(io->cond[0]) && al16(io->tr_in[0]) &&
                         al16(io->tr_out[0]) &&
                         pair_smem_bytes(net, io->tr_in[0].width, io->cond[0].width, false) <= 227 * 1024;
  if (narrow_ok && pair_mode() != 2 && !(flags & BGX_FLAG_FORCE_WIDE)) return 1;
  return pair_smem_bytes(net, io->tr_in[0].width, io->cond[0].width, true) <= 227 * 1024 ? 2 : 0;
}

int spline_coupling_pair(const bgx_coupling_io* io, const bgx_packed_mlp* net, const bgx_spline_cfg* cfg, int flags,
                         int mode, int* status, cudaStream_t st) {
  const int L = net->n_layers;
  const int d_t = io->tr_in[0].width;
  if (net->N[L - 1] != ceil_div(d_t, P_DPP) * 128 || io->tr_out[0].width != d_t || !io->dlogp_out) return BGX_ERR_INVALID;
  if (net->act < 0 || net->act > 3) return BGX_ERR_INVALID;
  if (io->batch == 0) return BGX_OK;
  const bool wide = mode == 2;
  PArgs a{};
  a.B = io->batch;
  a.cond = io->cond[0].ptr; a.tin = io->tr_in[0].ptr; a.tout = const_cast<float*>(io->tr_out[0].ptr);
  a.D_t = d_t; a.K0raw = io->cond[0].width;
  mlp_to_dev(net, a.net);
  int hid = 0;
  for (int l = 0; l < L; ++l) {
    a.wb[0][l] = (const uint16_t*)net->Wb[0][l];
    a.wb[1][l] = (const uint16_t*)net->Wb[1][l];
    a.ktiles[l] = ceil_div(net->K[l], 64);
    if (l < L - 1) hid += net->Np[l];
  }
  a.hid_bias_floats = hid;
  a.npass = net->N[L - 1] / 128;
  a.last_bias_floats = a.npass * P_DPP * P_BPAD;
  a.bias_last = net->spline_bias;
  a.G = ceil_div(net->K[0], 128);
  a.plain_cond = (net->K[0] == net->raw_width && net->periodic_scale == 0.f) ? 1 : 0;
  if (a.plain_cond && wide && a.K0raw % 4 == 0 && ((uintptr_t)a.cond & 15) == 0) a.plain_cond = 2;   // vector row reads
  {
    int cap = 0;
    unsigned long long* tb = tc_get_trace(&cap);
    a.trace = cap >= 16 ? tb : nullptr;
  }
  {
    static const int dbg = [] { const char* e = getenv("BGX_PAIR_DEBUG"); return e ? atoi(e) : 0; }();
    a.debug = dbg;
  }
  SplineParams sp;
  spline_params_from_cfg(cfg, sp);
  {
    const float wx = sp.right - sp.left, hy = sp.top - sp.bottom;
    a.ck.left = sp.left; a.ck.right = sp.right; a.ck.bottom = sp.bottom; a.ck.top = sp.top;
    a.ck.wscale = wx * (1.f - sp.min_w * P_NB); a.ck.hscale = hy * (1.f - sp.min_h * P_NB);
    a.ck.wstep = wx * sp.min_w; a.ck.hstep = hy * sp.min_h;
    a.ck.min_d = sp.min_d; a.ck.beta = sp.beta; a.ck.beta_l2e = sp.beta * LOG2E;
    a.ck.ln2_over_beta = LN2 * sp.inv_beta;
  }
  a.oob = sp.oob;
  a.dlogp_in = io->dlogp_in;
  a.dlogp_out = io->dlogp_out;
  a.status = status;
  // roles {dims 0,1 | dims 2,3 | dim 4 | idle} cost {2, 2, 1, 0} evaluations: with an offset of 2 between the slots a
  // warp's two concurrent events cost 2+1, 2+0, 1+2, 0+2 (max 3) instead of 2+2, 2+1, 1+0, 0+2 (max 4) with offset 1
  static const int role_stride = [] { const char* e = getenv("BGX_PAIR_ROLE_STRIDE"); return e ? atoi(e) : 2; }();
  a.role_stride = role_stride;
  a.ntiles = (a.B + P_TM - 1) / P_TM;
  a.npairs = (a.ntiles + 1) / 2;
  const size_t smem = pair_smem_bytes(net, d_t, a.K0raw, wide);
  int sm_count = 0;
  int rc = device_sm_count(&sm_count);
  if (rc) return rc;
  using KernT = void (*)(const PArgs);
#define BGX_P_ROW(INV, W, E) \
  {spline_coupling_pair_kernel<INV, 0, W, E>, spline_coupling_pair_kernel<INV, 1, W, E>, \
   spline_coupling_pair_kernel<INV, 2, W, E>, spline_coupling_pair_kernel<INV, 3, W, E>}
  static const KernT kerns[2][2][2][4] = {
      {{BGX_P_ROW(false, false, 4), BGX_P_ROW(true, false, 4)}, {BGX_P_ROW(false, true, 4), BGX_P_ROW(true, true, 4)}},
      {{BGX_P_ROW(false, false, 6), BGX_P_ROW(true, false, 6)}, {BGX_P_ROW(false, true, 6), BGX_P_ROW(true, true, 6)}}};
#undef BGX_P_ROW
  // epilogue warps per quadrant: BGX_PAIR_EPW=4 (16 warps, packed pairs of dims) or 6 (24 warps, one dim per warp and pass)
  static const int epw6 = [] { const char* e = getenv("BGX_PAIR_EPW"); return (e ? atoi(e) : P_EPW_DEFAULT) == 6 ? 1 : 0; }();
  const int inv = (flags & BGX_FLAG_INVERSE) ? 1 : 0;
  KernT kern = kerns[epw6][wide ? 1 : 0][inv][net->act];
  static size_t configured_all[BGX_MAX_DEVICES][2][2][2][4] = {};
  auto& configured = configured_all[device_slot()];
  if (smem > configured[epw6][wide ? 1 : 0][inv][net->act]) {
    rc = check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
    configured[epw6][wide ? 1 : 0][inv][net->act] = smem;
  }
  // CTAs per cluster (BGX_PAIR_CLUSTER = 1, 2 or 4): the CTAs of a cluster fetch every weight stage ONCE (each issues its
  // share of the bulk copies with .multicast::cluster), which divides the L2 -> SM weight traffic by the cluster size
  static const int cs_env = [] { const char* e = getenv("BGX_PAIR_CLUSTER"); return e ? atoi(e) : P_CLUSTER_DEFAULT; }();
  const int cs = (cs_env == 2 || cs_env == 4) ? cs_env : 1;
  a.cs = cs;
  cudaLaunchConfig_t lc = {};
  lc.blockDim = dim3((epw6 ? 28 : 20) * 32, 1, 1);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  lc.attrs = attr;
  lc.numAttrs = cs > 1 ? 1 : 0;
  long long ncl = sm_count / cs;
  if (cs > 1) {
    static int max_clusters[BGX_MAX_DEVICES][2][2][2][2][4] = {};     // [device][cs == 4][epw6][wide][inv][act]: co-resident clusters
    int& mc = max_clusters[device_slot()][cs == 4][epw6][wide ? 1 : 0][inv][net->act];
    if (!mc) {
      lc.gridDim = dim3((unsigned)(ncl * cs), 1, 1);
      rc = check(cudaOccupancyMaxActiveClusters(&mc, kern, &lc));
      if (rc) return rc;
      if (mc < 1) return BGX_ERR_UNSUPPORTED;
    }
    ncl = std::min<long long>(ncl, mc);
  }
  ncl = std::min<long long>(ncl, (a.npairs + cs - 1) / cs);
  lc.gridDim = dim3((unsigned)(ncl * cs), 1, 1);
  rc = check(cudaLaunchKernelEx(&lc, kern, a));
  if (rc) return rc;
  return post_launch();
}

}  // namespace bgx
