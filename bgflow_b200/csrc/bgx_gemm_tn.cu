// Batch-reduced weight gradient  dW[N, K] = G[B, N]^T . H[B, K]  and  db[N] = sum_b G[b, :]  on tcgen05, fp32-class
// accuracy (training path).
//
// The weight gradients of a conditioner layer (bgflow/nn/dense.py:47-48 differentiated; the reference leaves them to
// torch autograd = fp32 SIMT GEMMs, after the bgx_linear change the largest item of a KL training step) reduce over
// the BATCH: M = N out-features, N = K in-features (<= 128), reduction length B = 65536+.  Both operands are "batch-
// minor" in memory (row b of G / H holds all features of sample b), the transpose of what a K-major UMMA operand
// wants, and they are fp32.  So the CTA's 16 converter warps do the layout change themselves: they read a
// [64 samples x 128 features] fp32 slab with full 32-byte sectors, split every value exactly into two bf16 terms
// (x = x1 + x2) and store the terms TRANSPOSED into two [128 features x 64 samples] K-major SWIZZLE_128B tiles in
// shared memory (the lane mapping below makes those stores bank-conflict free), fence them to the async proxy and
// hand them to the MMA warp, which issues  g1 h2 + g2 h1 + g1 h1  per 16-sample k-step into fp32 accumulators in
// tensor memory (SS mode: both operands from shared memory).  Nothing is written back to HBM but the results: the
// kernel reads G and H exactly once (H once per group of four 128-feature tiles of G).
//
//   grid = (slices, ceil(ceil(N / 128) / 4)): a CTA owns a contiguous range of 64-sample batch tiles and up to four
//   128-feature tiles of G (4 x 128 accumulator columns = all of tensor memory); it writes its partial sums to
//   part_w[slice][n][k] / part_b[slice][n]; the caller sums over slices (fixed order: deterministic gradients).
//
//   warps 0-15 converters (loads of the next slab in flight while one is converted) / epilogue, warp 16 MMA issuer.  Shared memory: H tiles 2 x 32 KB, G tiles 4 x 32 KB.
#include <algorithm>

#include "bgx_pair.cuh"

namespace bgx {

constexpr int T_CONV_WARPS = 16;
constexpr int T_THREADS = (T_CONV_WARPS + 1) * 32;
constexpr int T_BUF = 32768;          // one operand tile: two bf16 terms x 16 KB
constexpr int T_HBUFS = 2, T_GBUFS = 4;
constexpr int T_NT = 4;               // 128-feature tiles of G per CTA

struct TnArgs {
  long long B;
  const float* g;
  long long ldg;
  int N;
  const float* h;
  long long ldh;
  int K;
  float* part_w;     // [slices][ntp * 128][128]
  float* part_b;     // [slices][ntp * 128] or null
  int slices, nt_total;
  long long bt_total;
  int* status;
};

struct alignas(16) TnSmem {
  uint64_t h_full[T_HBUFS], h_empty[T_HBUFS], g_full[T_GBUFS], g_empty[T_GBUFS], done;
  uint32_t tmem_base, pad[3];
};

// D[tmem] (+)= A[smem] . B[smem]^T for the three products of one k-step (a1 b2, a2 b1, a1 b1), elected lane issues
__device__ __forceinline__ void mma3_bf16x3_ss_elect(uint32_t d_tmem, uint64_t a1, uint64_t a2, uint64_t b1, uint64_t b2,
                                                     uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %6, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %4, %5, p;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %5, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %5, 1;\n\t}" ::"r"(d_tmem),
      "l"(a1), "l"(a2), "l"(b1), "l"(b2), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One warp's share of a [64 samples x 128 features] slab: 32 consecutive features (group fg = warp & 3, lane = feature)
// of two sample octets (bo = warp >> 2 and bo + 4).  A load instruction reads ONE row segment of 128 contiguous bytes
// (a full line per request: the kernel lives on how many bytes it keeps in flight), and the 8 samples of an octet a
// thread collects for its feature are exactly one 16-byte chunk of the swizzled K-major row, written with one
// st.shared.v4 per term — conflict-free, since the 8 lanes of a quarter warp are 8 consecutive features, whose chunk
// index (octet ^ (feature & 7)) differs.  Loads and stores are separate functions so that the loads of the NEXT slab
// are in flight while this one is converted.
__device__ __forceinline__ void load_slab(float (&v)[16], const float* __restrict__ src, long long ld, int ncols, int col,
                                          long long row0, long long B, int wq) {
  const bool cok = col < ncols;
  if (cok && row0 + 64 <= B) {
    // whole slab inside the batch (all but the last batch tile): one base pointer, no per-load guards (120 -> 94 us)
    const float* p = src + (row0 + 8 * wq) * ld + col;
    const long long ld4 = ld * 4;          // row stride in bytes
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      // walk the 8 rows of an octet with one 64-bit add per load (the guarded form below costs ~12 integer
      // instructions per load, index arithmetic from the row number ~5)
      const char* q = reinterpret_cast<const char*>(p) + 32 * u * ld4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[8 * u + i] = __ldg(reinterpret_cast<const float*>(q));
        q += ld4;
      }
    }
    return;
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const long long c0 = row0 + 8 * (wq + 4 * u);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[8 * u + i] = (cok && c0 + i < B) ? __ldg(src + (c0 + i) * ld + col) : 0.f;
  }
}
// r = feature row inside the tile; returns this thread's sum of the values it converted (for db)
__device__ __forceinline__ float store_slab(const float (&v)[16], uint8_t* buf, int r, int wq) {
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int bo = wq + 4 * u;
    uint32_t t1[4], t2[4], t3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      split_bf16(v[8 * u + 2 * i], v[8 * u + 2 * i + 1], 2, t1[i], t2[i], t3);
      s += v[8 * u + 2 * i] + v[8 * u + 2 * i + 1];
    }
    const uint32_t off = (uint32_t)r * 128u + ((uint32_t)(bo ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(buf + off) = make_uint4(t1[0], t1[1], t1[2], t1[3]);
    *reinterpret_cast<uint4*>(buf + 16384 + off) = make_uint4(t2[0], t2[1], t2[2], t2[3]);
  }
  return s;
}

__global__ void __launch_bounds__(T_THREADS, 1) gemm_tn_kernel(const __grid_constant__ TnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* hbuf = base;
  uint8_t* gbuf = base + T_HBUFS * T_BUF;
  TnSmem* S = (TnSmem*)(base + (T_HBUFS + T_GBUFS) * T_BUF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x;
  const int nt0 = blockIdx.y * T_NT, ntl = min(T_NT, a.nt_total - nt0);
  const int ntp = a.nt_total;
  const long long bt_lo = a.bt_total * slice / a.slices, bt_hi = a.bt_total * (slice + 1) / a.slices;
  const long long nbt = bt_hi - bt_lo;

  if (threadIdx.x == 0) {
    for (int i = 0; i < T_HBUFS; ++i) {
      mbar_init(&S->h_full[i], T_CONV_WARPS);
      mbar_init(&S->h_empty[i], 1);
    }
    for (int i = 0; i < T_GBUFS; ++i) {
      mbar_init(&S->g_full[i], T_CONV_WARPS);
      mbar_init(&S->g_empty[i], 1);
    }
    mbar_init(&S->done, 1);
    fence_mbar_init();
  }
  // feature rows of the H tiles beyond K are never written by the converters: zero them once
  for (int i = threadIdx.x; i < T_HBUFS * T_BUF / 16; i += T_THREADS) reinterpret_cast<uint4*>(hbuf)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  if (warp == T_CONV_WARPS) tmem_alloc<512>(&S->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S->tmem_base;

  if (warp == T_CONV_WARPS) {
    // ------------------------------------------------------------------ MMA issuer (warp-wide, elected lane issues)
    const uint32_t idesc = idesc_bf16(128, 128);
    uint32_t ph_hf = 0, ph_gf = 0;      // phase bits, one per buffer
    long long gcons = 0;
    bool ok = true;
#pragma unroll 1
    for (long long i = 0; i < nbt && ok; ++i) {
      const int hb = (int)(i & 1);
      ok = mbar_wait(&S->h_full[hb], (ph_hf >> hb) & 1u, a.status);
      ph_hf ^= 1u << hb;
      const uint32_t ha = smem_u32(hbuf + hb * T_BUF);
      const uint64_t b1 = smem_desc_sw128(ha), b2 = smem_desc_sw128(ha + 16384);
#pragma unroll 1
      for (int t = 0; t < ntl && ok; ++t) {
        const int gi = (int)(gcons & 3);
        ok = mbar_wait(&S->g_full[gi], (ph_gf >> gi) & 1u, a.status);
        ph_gf ^= 1u << gi;
        if (!ok) break;
        tc_fence_after();
        const uint32_t ga = smem_u32(gbuf + gi * T_BUF);
        const uint64_t a1 = smem_desc_sw128(ga), a2 = smem_desc_sw128(ga + 16384);
#pragma unroll 1
        for (int ks = 0; ks < 4; ++ks)
          mma3_bf16x3_ss_elect(tmem + (uint32_t)t * 128u, a1 + 2 * ks, a2 + 2 * ks, b1 + 2 * ks, b2 + 2 * ks, idesc,
                               (i > 0 || ks > 0) ? 1u : 0u);
        mma_commit_elect(&S->g_empty[gi]);
        ++gcons;
      }
      mma_commit_elect(&S->h_empty[hb]);
    }
    mma_commit_elect(&S->done);
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ converters (warps 0..15), then epilogue
    const int fg = warp & 3, wq = warp >> 2;
    const int r = 32 * fg + lane;
    float bsum[T_NT] = {0.f, 0.f, 0.f, 0.f};
    uint32_t ph_he = 0, ph_ge = 0;
    bool ok = true;
    // slabs in consumption order: per batch tile the H slab (kind 0), then the G slabs of this CTA (kind 1 .. ntl)
    const int per = 1 + ntl;
    const long long total = nbt * per;
    const bool h_active = 32 * fg < a.K;      // feature groups of H that carry data (the rest stays zero)
    auto issue_loads = [&](float (&v)[16], long long i, int kind) {
      const long long row0 = (bt_lo + i) * 64;
      if (kind == 0) {
        if (h_active) load_slab(v, a.h, a.ldh, a.K, r, row0, a.B, wq);
      } else {
        load_slab(v, a.g, a.ldg, a.N, (nt0 + kind - 1) * 128 + r, row0, a.B, wq);
      }
    };
    // (A third register buffer — loads two slabs ahead — and cheaper load addressing were both measured: no change,
    // 94-96 us at N = 825, B = 65536.  The SS-mode MMAs are not the limit either: tools/mma_rate.py runs them at 64.1
    // cycles per 128x128x16, the same as TS mode.  What is left is the converters' own instruction stream — ncu: ~300
    // warp instructions per warp and slab at an IPC of 2, issue slots 51 % busy, against 768 cycles of MMA work.)
    float va[16], vb[16];
    if (total > 0) issue_loads(va, 0, 0);
    long long i = 0;
    int kind = 0;
    auto step = [&](float (&cur)[16], float (&nxt)[16]) {
      // loads of the next slab first, then wait for this slab's buffer and convert
      int nk = kind + 1;
      long long ni = i;
      if (nk == per) { nk = 0; ++ni; }
      if (ni < nbt) issue_loads(nxt, ni, nk);
      if (kind == 0) {
        const int hb = (int)(i & 1);
        if (i >= T_HBUFS) {
          ok = mbar_wait(&S->h_empty[hb], (ph_he >> hb) & 1u, a.status);
          ph_he ^= 1u << hb;
        }
        if (ok) {
          if (h_active) store_slab(cur, hbuf + hb * T_BUF, r, wq);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S->h_full[hb]);
        }
      } else {
        const long long gfill = i * ntl + (kind - 1);
        const int gi = (int)(gfill & 3);
        if (gfill >= T_GBUFS) {
          ok = mbar_wait(&S->g_empty[gi], (ph_ge >> gi) & 1u, a.status);
          ph_ge ^= 1u << gi;
        }
        if (ok) {
          const float sum = store_slab(cur, gbuf + gi * T_BUF, r, wq);
#pragma unroll
          for (int t = 0; t < T_NT; ++t) bsum[t] += (kind - 1 == t) ? sum : 0.f;
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S->g_full[gi]);
        }
      }
      kind = nk;
      i = ni;
    };
#pragma unroll 1
    while (i < nbt && ok) {
      step(va, vb);
      if (i < nbt && ok) step(vb, va);
    }
    // ---- epilogue: the accumulators (row = out-feature, column = in-feature) and the column sums of G
    if (ok && nbt > 0) ok = mbar_wait(&S->done, 0, a.status);
    tc_fence_after();
    const int q = warp & 3, j = warp >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
#pragma unroll
    for (int t = 0; t < T_NT; ++t) {
      if (t < ntl) {
        uint32_t v[32];
        if (nbt > 0) {
          tmem_ld32(tmem + lane_base + (uint32_t)t * 128u + (uint32_t)j * 32u, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = 0;
        }
        const long long row = (long long)(nt0 + t) * 128 + q * 32 + lane;
        float* out = a.part_w + (((long long)slice * ntp * 128 + row) * 128 + j * 32);
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          *reinterpret_cast<uint4*>(out + k) = make_uint4(v[k], v[k + 1], v[k + 2], v[k + 3]);
      }
    }
    tc_fence_before();
    if (a.part_b) {
      // the four warps wq = 0..3 of a feature group hold partial sums of the same 32 features: add them in a fixed
      // order through shared memory (the G ring is idle: every MMA has completed)
      float* red = reinterpret_cast<float*>(gbuf);
#pragma unroll
      for (int t = 0; t < T_NT; ++t) red[(wq * T_NT + t) * 128 + r] = bsum[t];
      asm volatile("bar.sync 1, %0;" ::"n"(T_CONV_WARPS * 32) : "memory");
      for (int e = threadIdx.x; e < ntl * 128; e += T_CONV_WARPS * 32) {
        const int t = e >> 7, f = e & 127;
        const float sum = ((red[(0 * T_NT + t) * 128 + f] + red[(1 * T_NT + t) * 128 + f]) + red[(2 * T_NT + t) * 128 + f]) +
                          red[(3 * T_NT + t) * 128 + f];
        a.part_b[(long long)slice * ntp * 128 + (nt0 + t) * 128 + f] = sum;
      }
    }
  }
  __syncthreads();
  if (warp == T_CONV_WARPS) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

static int tn_sm_count(int* out) { return device_sm_count(out); }

}  // namespace bgx

using namespace bgx;

// number of batch slices bgx_gemm_tn should be called with (one CTA per SM over slices x feature-tile groups)
extern "C" int bgx_gemm_tn_slices(int64_t batch, int n) {
  if (batch <= 0 || n <= 0) return 0;
  int sms = 0;
  if (tn_sm_count(&sms)) return 0;
  const int groups = ceil_div(ceil_div(n, 128), T_NT);
  const long long bt = (batch + 63) / 64;
  return (int)std::max<long long>(1, std::min<long long>(bt, sms / groups));
}

extern "C" int bgx_gemm_tn(int64_t batch, const float* g, int64_t ldg, int n, const float* h, int64_t ldh, int k,
                           int slices, float* part_w, float* part_b, int32_t* status, void* stream) {
  if (batch <= 0 || n <= 0 || k <= 0 || !g || !h || !part_w || ldg < n || ldh < k) return BGX_ERR_INVALID;
  if (k > 128 || ldg > (1 << 24) || ldh > (1 << 24)) return BGX_ERR_UNSUPPORTED;
  TnArgs a{};
  a.B = batch;
  a.g = g; a.ldg = ldg; a.N = n;
  a.h = h; a.ldh = ldh; a.K = k;
  a.part_w = part_w; a.part_b = part_b;
  a.nt_total = ceil_div(n, 128);
  a.bt_total = (batch + 63) / 64;
  if (slices < 1 || slices > a.bt_total) return BGX_ERR_INVALID;
  a.slices = slices;
  a.status = status;
  const size_t smem = 1024 + (T_HBUFS + T_GBUFS) * T_BUF + sizeof(TnSmem) + 64;
  static bool configured_all[BGX_MAX_DEVICES] = {};
  bool& configured = configured_all[device_slot()];
  if (!configured) {
    int rc = check(cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc) return rc;
    configured = true;
  }
  dim3 grid((unsigned)slices, (unsigned)ceil_div(a.nt_total, T_NT));
  gemm_tn_kernel<<<grid, T_THREADS, smem, (cudaStream_t)stream>>>(a);
  return post_launch();
}
