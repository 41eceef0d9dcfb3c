// Global Z-matrix <-> Cartesian transform, one thread per sample, closed-form log-dets.
//
// bgx_ic_to_xyz   : GlobalInternalCoordinateTransformation._inverse  (ic.py:678-716,
//                   ReferenceSystemTransformation._inverse ic.py:209-265 + init_ics2xyz
//                   ic_helper.py:480-575, RelativeInternalCoordinateTransformation._inverse
//                   ic.py:435-513 + ic2xyz_deriv ic_helper.py:372-452)
// bgx_ic_from_xyz : GlobalInternalCoordinateTransformation._forward  (ic.py:633-676, 162-206,
//                   386-433; dist/angle/torsion_deriv ic_helper.py:148-293; init_xyz2ics :578-680)
//
// The reference forms a 3x3 Jacobian per placed atom and a 9x9 autograd Jacobian for the
// reference frame; both determinants have closed forms:
//   placed atom: 2 ln b + ln sin a ;  frame: 2 ln d01 + 2 ln d12 + ln sin a012.
//
// The same two kernels also run the RELATIVE transform (RelativeInternalCoordinateTransformation,
// ic.py:268-513: no reference frame, the fixed atoms' Cartesian coordinates are an input / output
// block) and the MIXED transform (MixedCoordinateTransformation, ic.py:719-884: the fixed block is
// whitened by a static PCA, pca.py:37-107), and can apply the IC-domain CDF maps of the builder
// (cdf.py:29-46, bgx_cdf_math.cuh) to every internal coordinate on the fly.
//
// Layout: the [tile x 3N] coordinate block of a CTA is contiguous in global memory; it is
// staged through shared memory as P[c][t] (leading dim BT+1: conflict-free both for the
// per-thread phase (fixed c, consecutive t) and for the coalesced copy phase).
#include <cstdlib>
#include <map>

#include "bgx_ic.cuh"
#include "bgx_cdf_math.cuh"
#include "bgx_tc.cuh"

namespace bgx {

enum { IC_GLOBAL = 0, IC_RELATIVE = 1 };
// Tile I/O of the kernels.  IO_GLOBAL: every thread works on its row in global memory (molecules whose tile does not
// fit shared memory).  IO_STAGED: element-wise staging through a transposed shared-memory tile (any batch, any
// alignment; the tail rows of a batch).  IO_BULK: whole row-major tiles move with ONE bulk-TMA copy per tensor
// (cp.async.bulk, SASS UBLKCP) and the per-thread phase reads / writes its row of the tile in place — the staging
// walks were 60 % of the instructions and two thirds of the time of these kernels (ncu source page, profiles/).
enum { IO_GLOBAL = 0, IO_STAGED = 1, IO_BULK = 2 };

using tc::bulk_g2s;
using tc::fence_async_smem;
using tc::fence_mbar_init;
using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;

__device__ __forceinline__ void ic_bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(tc::smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void ic_bulk_commit_wait_read() {
  asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
}

struct IcArgs {
  long long B;
  int kind;               // IC_GLOBAL: 3 seed atoms + reference frame; IC_RELATIVE: fixed-atom block
  int n_atoms, n_rel;
  int nb, na, nt;         // columns of bonds / angles / torsions (global: N-1, N-2, N-3; relative: n_rel each)
  int off_b, off_a;       // column of z-matrix row 0 in bonds / angles (global: 2, 1; relative: 0, 0)
  int s0, s1, s2;
  // relative / mixed transforms
  int n_fixed, keep, fixed_w;          // fixed_w = keep ? keep : 3 * n_fixed (columns of the fixed tensor)
  const int* fixed;                    // [n_fixed] atom ids
  const float *mean, *blacken, *whiten;
  float ld_whiten;
  const float* fixed_in;
  float* fixed_out;
  // optional CDF maps of the IC columns (bonds | angles | torsions order)
  const bgx_cdf_col* marg;
  CdfClamp clamp;
  const int* rel;    // [n_rel][4]
  const int* order;  // [n_rel]
  const int* slot_of_col;  // [3N-6]: smem slot (3*atom + {0:bond,1:angle,2:torsion}) of every IC column,
                           //          columns ordered bonds | angles | torsions
  int normalize;
  float eps, eps2, cmin, cmax;
  const float *bonds, *angles, *torsions, *x0, *R;
  int x0_stride, r_stride;
  float* xyz;
  float *o_bonds, *o_angles, *o_torsions, *o_x0, *o_R;
  const float* xyz_in;
  const float* dlogp_in;
  float* dlogp_out;
};

// ---------------------------------------------------------------- IC -> Cartesian
template <int MODE>
__global__ void __launch_bounds__(BT) ic_to_xyz_kernel(const IcArgs a) {
  constexpr bool SMEM = MODE == IO_STAGED, BULK = MODE == IO_BULK;
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bulk_bar;
  const int t = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * BT;
  const long long row = row0 + t;
  const bool live = row < a.B;
  const int N = a.n_atoms, nb = a.nb, na = a.na, nt = a.nt;
  const int nrow = (int)min((long long)BT, a.B - row0);
  const bool global = a.kind == IC_GLOBAL;
  PosStore<SMEM> pos;
  if (SMEM) {
    // Every placed atom owns exactly three internal coordinates and three Cartesian coordinates:
    // its (bond, angle, torsion) are staged in the very shared-memory slots its (x, y, z) will
    // overwrite.  Coalesced loads of the three [tile x n] blocks, scattered by column -> slot.
    pos.base = sm + t; pos.stride_c = LDT;
    const float* gb = a.bonds + row0 * nb;
    const float* ga = a.angles + row0 * na;
    const float* gt = a.torsions + row0 * nt;
    if (global) {
      walk_block_ld(t, nb, nrow, [&](int m, int c) { return __ldg(gb + m * nb + c); },
                    [&](int m, int c, float v) { sm[a.slot_of_col[c] * LDT + m] = v; });
      walk_block_ld(t, na, nrow, [&](int m, int c) { return __ldg(ga + m * na + c); },
                    [&](int m, int c, float v) { sm[a.slot_of_col[nb + c] * LDT + m] = v; });
      walk_block_ld(t, nt, nrow, [&](int m, int c) { return __ldg(gt + m * nt + c); },
                    [&](int m, int c, float v) { sm[a.slot_of_col[nb + na + c] * LDT + m] = v; });
    } else {
      // column r of every IC tensor belongs to the atom of z-matrix row r
      walk_block_ld(t, nb, nrow, [&](int m, int c) { return __ldg(gb + m * nb + c); },
                    [&](int m, int c, float v) { sm[(3 * a.rel[4 * c]) * LDT + m] = v; });
      walk_block_ld(t, na, nrow, [&](int m, int c) { return __ldg(ga + m * na + c); },
                    [&](int m, int c, float v) { sm[(3 * a.rel[4 * c] + 1) * LDT + m] = v; });
      walk_block_ld(t, nt, nrow, [&](int m, int c) { return __ldg(gt + m * nt + c); },
                    [&](int m, int c, float v) { sm[(3 * a.rel[4 * c] + 2) * LDT + m] = v; });
      const int wf = a.fixed_w;
      const float* gf = a.fixed_in + row0 * wf;
      if (a.keep == 0)     // Cartesian block of the fixed atoms goes straight to their position slots
        walk_block_ld(t, wf, nrow, [&](int m, int c) { return __ldg(gf + m * wf + c); },
                      [&](int m, int c, float v) {
                        const int i = c / 3;
                        sm[(3 * a.fixed[i] + (c - 3 * i)) * LDT + m] = v;
                      });
      else                 // whitened block: parked behind the positions, un-whitened per thread below
        walk_block_ld(t, wf, nrow, [&](int m, int c) { return __ldg(gf + m * wf + c); },
                      [&](int m, int c, float v) { sm[(3 * N + c) * LDT + m] = v; });
    }
    __syncthreads();
  } else if (BULK) {
    // row-major tiles [BT x width]: bonds | angles | torsions (| fixed block) | output coordinates
    float* raw_b = sm;
    float* raw_a = raw_b + BT * nb;
    float* raw_t = raw_a + BT * na;
    float* raw_f = raw_t + BT * nt;
    float* out = raw_f + (global ? 0 : BT * a.fixed_w);
    if (t == 0) {
      mbar_init(&bulk_bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (t == 0) {
      mbar_expect_tx(&bulk_bar, (uint32_t)(BT * 4 * (nb + na + nt + (global ? 0 : a.fixed_w))));
      bulk_g2s(raw_b, a.bonds + row0 * nb, (uint32_t)(BT * nb * 4), &bulk_bar);
      bulk_g2s(raw_a, a.angles + row0 * na, (uint32_t)(BT * na * 4), &bulk_bar);
      bulk_g2s(raw_t, a.torsions + row0 * nt, (uint32_t)(BT * nt * 4), &bulk_bar);
      if (!global) bulk_g2s(raw_f, a.fixed_in + row0 * a.fixed_w, (uint32_t)(BT * a.fixed_w * 4), &bulk_bar);
    }
    mbar_wait(&bulk_bar, 0, nullptr);
    pos.base = out + t * (3 * N); pos.stride_c = 1;
  } else {
    pos.base = a.xyz + row * (long long)(3 * N); pos.stride_c = 1;
  }

  if (live) {
    const float* bo = BULK ? sm + t * nb : a.bonds + row * nb;
    const float* an = BULK ? sm + BT * nb + t * na : a.angles + row * na;
    const float* to = BULK ? sm + BT * (nb + na) + t * nt : a.torsions + row * nt;
    const float* fx = BULK ? sm + BT * (nb + na + nt) + t * a.fixed_w : a.fixed_in + row * (long long)a.fixed_w;
    float dl = 0.f;
    if (global) {
    const float* x0p = a.x0 + row * a.x0_stride;
    const float* rp = a.R + row * a.r_stride;
    float d01, d12, a012;
    if (SMEM) {
      d01 = pos.base[(3 * a.s1) * LDT]; d12 = pos.base[(3 * a.s2) * LDT]; a012 = pos.base[(3 * a.s2 + 1) * LDT];
    } else {
      d01 = bo[0]; d12 = bo[1]; a012 = an[0];
    }
    float alpha = rp[0], beta = rp[1], gamma = rp[2];
    if (a.marg) {
      float ld;
      cdf_inverse(a.marg[0], a.clamp, d01, d01, ld); dl += ld;
      cdf_inverse(a.marg[1], a.clamp, d12, d12, ld); dl += ld;
      cdf_inverse(a.marg[nb], a.clamp, a012, a012, ld); dl += ld;
    }
    if (a.normalize) {
      a012 *= PI_F;
      alpha = alpha * TWO_PI_F - PI_F;
      gamma = gamma * TWO_PI_F - PI_F;
      dl += (float)((double)na * 1.1447298858494002 + (double)(nt + 2) * 1.8378770664093453);
    }
    float sa, ca, sb, cb, sg, cg, s012, c012;
    __sincosf(alpha, &sa, &ca);                 // arguments in [-pi, pi]: MUFU error ~4e-7
    cb = beta;                                  // theta = acos(beta): cos = beta, sin = sqrt(1 - beta^2)
    sb = sqrt_approx(fmaxf(1.f - beta * beta, 0.f));
    __sincosf(gamma, &sg, &cg);
    __sincosf(a012, &s012, &c012);
    // R = Rz(alpha) Rx(theta) Rz(gamma)   (ic_helper.py:344-368)
    const float m00 = ca, m01 = -sa * cb, m02 = sa * sb;
    const float m10 = sa, m11 = ca * cb, m12 = -ca * sb;
    const float m20 = 0.f, m21 = sb, m22 = cb;
    const float r00 = m00 * cg + m01 * sg, r01 = -m00 * sg + m01 * cg, r02 = m02;
    const float r10 = m10 * cg + m11 * sg, r11 = -m10 * sg + m11 * cg, r12 = m12;
    const float r20 = m20 * cg + m21 * sg, r21 = -m20 * sg + m21 * cg, r22 = m22;
    V3 o = {x0p[0], x0p[1], x0p[2]};
    // p1 = (0,0,d01); p2 = (d12 sin a, [4.4e-8 d12 sin a: the reference's float32 pi/2], d01 - d12 cos a)
    const float px = d12 * s012, py = px * 4.371139e-8f, pz = d01 - d12 * c012;
    V3 x1 = {o.x + r02 * d01, o.y + r12 * d01, o.z + r22 * d01};
    V3 x2 = {o.x + r00 * px + r01 * py + r02 * pz, o.y + r10 * px + r11 * py + r12 * pz,
             o.z + r20 * px + r21 * py + r22 * pz};
    pos.set(a.s0, o);
    pos.set(a.s1, x1);
    pos.set(a.s2, x2);
    // log-det terms are accumulated as a product where that cannot overflow: one log per 4 atoms
    dl += 2.f * __logf(d01 * d12) + __logf(s012);
    } else {
      if (a.normalize) dl = (float)((double)na * 1.1447298858494002 + (double)nt * 1.8378770664093453);
      const int nf3 = 3 * a.n_fixed;
      if (a.keep > 0) {
        // x_fixed = z_fixed . Tblacken + mean   (pca.py:93-99), log-det -jacobian_xz
        const float* zf = fx;
        for (int i = 0; i < a.n_fixed; ++i) {
          float acc[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[c] = __ldg(a.mean + 3 * i + c);
          for (int k = 0; k < a.keep; ++k) {
            const float zk = SMEM ? pos.base[(3 * N + k) * LDT] : zf[k];
            const float* tb = a.blacken + (long long)k * nf3 + 3 * i;
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[c] = fmaf(zk, __ldg(tb + c), acc[c]);
          }
          pos.set(a.fixed[i], {acc[0], acc[1], acc[2]});
        }
        dl -= a.ld_whiten;
      } else if (!SMEM) {
        const float* xf = fx;
        for (int i = 0; i < a.n_fixed; ++i) pos.set(a.fixed[i], {xf[3 * i], xf[3 * i + 1], xf[3 * i + 2]});
      }
    }
    for (int q = 0; q < a.n_rel; ++q) {
      const int r = a.order[q];
      const int4 z = *reinterpret_cast<const int4*>(a.rel + 4 * r);
      float d, ang, tor;
      if (SMEM) {
        const float* ic = pos.base + (3 * z.x) * LDT;
        d = ic[0]; ang = ic[LDT]; tor = ic[2 * LDT];
      } else {
        d = bo[a.off_b + r]; ang = an[a.off_a + r]; tor = to[r];
      }
      if (a.marg) {
        float l0, l1, l2;
        cdf_inverse(a.marg[a.off_b + r], a.clamp, d, d, l0);
        cdf_inverse(a.marg[nb + a.off_a + r], a.clamp, ang, ang, l1);
        cdf_inverse(a.marg[nb + na + r], a.clamp, tor, tor, l2);
        dl += l0 + l1 + l2;
      }
      if (a.normalize) {
        ang *= PI_F;
        tor = tor * TWO_PI_F - PI_F;
      }
      const V3 p1 = pos.get(z.y), p2 = pos.get(z.z), p3 = pos.get(z.w);
      V3 v1 = p1 - p2, v2 = p1 - p3;
      V3 n = cross(v1, v2);
      V3 nn = cross(v1, n);
      n = n * inv_norm_c(n, a.eps2);
      nn = nn * inv_norm_c(nn, a.eps2);
      float st, ct, sA, cA;
      __sincosf(tor, &st, &ct);
      __sincosf(ang, &sA, &cA);
      V3 v3 = n * (-st) + nn * ct;
      v3 = v3 * inv_norm_c(v3, a.eps2);
      v1 = v1 * inv_norm_c(v1, a.eps2);
      pos.set(z.x, p1 + v3 * (d * sA) - v1 * (d * cA));
      dl += __logf(d * d * sA);
    }
    a.dlogp_out[row] = (a.dlogp_in ? a.dlogp_in[row] : 0.f) + dl;
  }
  if (SMEM) {
    __syncthreads();
    const int W = 3 * N;
    float* g = a.xyz + row0 * W;
    walk_block(t, W, nrow, [&](int m, int c) { g[m * W + c] = sm[c * LDT + m]; });
  }
  if (BULK) {
    fence_async_smem();          // the tile was written through the generic proxy, the bulk store reads it through the async one
    __syncthreads();
    if (t == 0) {
      const int W = 3 * N;
      ic_bulk_s2g(a.xyz + row0 * W, sm + BT * (nb + na + nt + (global ? 0 : a.fixed_w)), (uint32_t)(BT * W * 4));
      ic_bulk_commit_wait_read();      // shared memory must outlive the copy's reads
    }
  }
}

// ---------------------------------------------------------------- Cartesian -> IC
// All z-matrix rows of a sample are independent, so FOUR threads share a sample (rows r = g, g+4, ...;
// thread g = 3 also does the reference frame): 32 samples per 128-thread CTA, 16.6 KB of shared
// memory per CTA (13 CTAs / 52 warps per SM) instead of one thread walking 19 rows.
constexpr int FS = 32;        // samples per CTA (from_xyz)
constexpr int LDF = FS + 1;

template <int MODE>
__global__ void __launch_bounds__(BT) ic_from_xyz_kernel(const IcArgs a) {
  constexpr bool SMEM = MODE == IO_STAGED, BULK = MODE == IO_BULK, QUAD = MODE != IO_GLOBAL;
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bulk_bar;
  const int t = threadIdx.x;
  const int N = a.n_atoms, nb = a.nb, na = a.na, nt = a.nt;
  const int W = 3 * N;
  const bool global = a.kind == IC_GLOBAL;
  const int spb = QUAD ? FS : BT;                    // samples per CTA
  const long long row0 = (long long)blockIdx.x * spb;
  const int s_loc = QUAD ? (t >> 2) : t, g = QUAD ? (t & 3) : 0, ng = QUAD ? 4 : 1;
  const long long row = row0 + s_loc;
  const bool live = row < a.B;
  const int nrow = (int)min((long long)spb, a.B - row0);
  float* Qs = sm + W * LDF + s_loc;     // staged outputs: column c of bonds|angles|torsions at Qs[c * LDF]
  PosStore<SMEM> pos;
  if (SMEM) {
    const float* gx = a.xyz_in + row0 * W;
    walk_block_ld(t, W, nrow, [&](int m, int c) { return __ldg(gx + m * W + c); },
                  [&](int m, int c, float v) { sm[c * LDF + m] = v; });
    __syncthreads();
    pos.base = sm + s_loc;
    pos.stride_c = LDF;
  } else if (BULK) {
    // row-major tiles [FS x width]: coordinates in | bonds | angles | torsions (| fixed block) out
    if (t == 0) {
      mbar_init(&bulk_bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (t == 0) {
      mbar_expect_tx(&bulk_bar, (uint32_t)(FS * W * 4));
      bulk_g2s(sm, a.xyz_in + row0 * W, (uint32_t)(FS * W * 4), &bulk_bar);
    }
    mbar_wait(&bulk_bar, 0, nullptr);
    pos.base = sm + s_loc * W;
    pos.stride_c = 1;
  } else {
    pos.base = const_cast<float*>(a.xyz_in) + row * (long long)(3 * N);
    pos.stride_c = 1;
  }
  float dl = 0.f;
  float* const qb = sm + FS * W;          // (BULK) output tiles
  if (live) {
    float* bo = BULK ? qb + s_loc * nb : a.o_bonds + row * nb;
    float* an = BULK ? qb + FS * nb + s_loc * na : a.o_angles + row * na;
    float* to = BULK ? qb + FS * (nb + na) + s_loc * nt : a.o_torsions + row * nt;
    for (int r = g; r < a.n_rel; r += ng) {
      const int4 z = *reinterpret_cast<const int4*>(a.rel + 4 * r);
      const V3 xi = pos.get(z.x), xj = pos.get(z.y), xk = pos.get(z.z), xl = pos.get(z.w);
      const V3 r12 = xi - xj, r32 = xk - xj;
      const float i12 = inv_norm_c(r12, a.eps2), i32 = inv_norm_c(r32, a.eps2);
      const float n12 = fmaxf(dot(r12, r12) * i12, a.eps);        // |r12| clamped like the reference
      float c = dot(r12, r32) * (i12 * i32);
      c = fminf(fmaxf(c, a.cmin), a.cmax);
      const float sin2 = 1.f - c * c;
      const float sin1 = sqrt_approx(sin2);
      float ang = atan2_fast(sin1, c);                 // acos(c), c in [-1 + eps, 1 - eps]
      // torsion (ic_helper.py:220-278): b0 = xi - xj, b1 = xk - xj, b2 = xl - xk
      const V3 b2 = xl - xk;
      const V3 u = r32 * i32;
      const V3 v = r12 - u * dot(r12, u);
      const V3 w = b2 - u * dot(b2, u);
      float tor = atan2_fast(dot(cross(u, v), w), dot(v, w));
      if (a.normalize) {
        ang *= (1.f / PI_F);
        tor = (tor + PI_F) * (1.f / TWO_PI_F);
      }
      dl -= __logf(n12 * n12 * sin1);     // 2 ln b + ln sin a
      float ob = n12;
      if (a.marg) {
        float l0, l1, l2;
        cdf_forward(a.marg[a.off_b + r], a.clamp, ob, ob, l0);
        cdf_forward(a.marg[nb + a.off_a + r], a.clamp, ang, ang, l1);
        cdf_forward(a.marg[nb + na + r], a.clamp, tor, tor, l2);
        dl += l0 + l1 + l2;
      }
      if (SMEM) {
        Qs[(a.off_b + r) * LDF] = ob; Qs[(nb + a.off_a + r) * LDF] = ang; Qs[(nb + na + r) * LDF] = tor;
      } else {
        bo[a.off_b + r] = ob; an[a.off_a + r] = ang; to[r] = tor;
      }
    }
    if (!global) {
      // fixed block (ic.py:419) and, for the mixed transform, its whitening (pca.py:83-91): the
      // threads of a sample share the atoms / the whitened components
      float* fo = BULK ? qb + FS * (nb + na + nt) + s_loc * a.fixed_w : a.fixed_out + row * a.fixed_w;
      float* Qf = Qs + (nb + na + nt) * LDF;
      if (a.keep == 0) {
        for (int i = g; i < a.n_fixed; i += ng) {
          const V3 p = pos.get(a.fixed[i]);
          if (SMEM) { Qf[(3 * i) * LDF] = p.x; Qf[(3 * i + 1) * LDF] = p.y; Qf[(3 * i + 2) * LDF] = p.z; }
          else { fo[3 * i] = p.x; fo[3 * i + 1] = p.y; fo[3 * i + 2] = p.z; }
        }
      } else {
        for (int k = g; k < a.keep; k += ng) {
          float acc = 0.f;
          for (int i = 0; i < a.n_fixed; ++i) {
            const V3 p = pos.get(a.fixed[i]);
            const float* tw = a.whiten + (long long)(3 * i) * a.keep + k;
            acc = fmaf(p.x - __ldg(a.mean + 3 * i), __ldg(tw), acc);
            acc = fmaf(p.y - __ldg(a.mean + 3 * i + 1), __ldg(tw + a.keep), acc);
            acc = fmaf(p.z - __ldg(a.mean + 3 * i + 2), __ldg(tw + 2 * a.keep), acc);
          }
          if (SMEM) Qf[k * LDF] = acc;
          else fo[k] = acc;
        }
        if (g == ng - 1) dl += a.ld_whiten;
      }
      if (g == ng - 1 && a.normalize) dl -= (float)((double)na * 1.1447298858494002 + (double)nt * 1.8378770664093453);
    }
    if (g == ng - 1 && global) {
      const V3 p0 = pos.get(a.s0), p1 = pos.get(a.s1), p2 = pos.get(a.s2);
      const V3 e01 = p1 - p0, e12 = p2 - p1;
      const float d01 = norm_c(e01, a.eps), d12 = norm_c(e12, a.eps);
      // angle at p1 between p0 and p2
      const V3 ra = p0 - p1;
      float c = dot(ra * (1.f / norm_c(ra, a.eps)), e12 * (1.f / d12));
      c = fminf(fmaxf(c, a.cmin), a.cmax);
      const float sin012 = sqrt_approx(1.f - c * c);
      float a012 = atan2_fast(sin012, c);
      // tripod (ic_helper.py:114-138): e1 = (p1-p0)/|.|, e2 = ((p2-p0) x e1)/|.|, e3 = e2 x e1
      const V3 e1 = e01 * (1.f / d01);
      V3 e2 = cross(p2 - p0, e1);
      e2 = e2 * (1.f / norm_c(e2, a.eps));
      const V3 e3 = cross(e2, e1);
      // basis (X, Y, Z) = (-e3, -e2, e1); euler (ic_helper.py:330-341)
      float alpha = atan2_fast(e1.x, -e1.y);
      const float beta = e1.z;
      float gamma = atan2_fast(-e3.z, -e2.z);
      dl -= __logf(d01 * d01 * d12 * d12 * sin012);
      if (a.normalize) {
        a012 *= (1.f / PI_F);
        alpha = (alpha + PI_F) * (1.f / TWO_PI_F);
        gamma = (gamma + PI_F) * (1.f / TWO_PI_F);
        dl -= (float)((double)na * 1.1447298858494002 + (double)(nt + 2) * 1.8378770664093453);
      }
      float o01 = d01, o12 = d12;
      if (a.marg) {
        float l0, l1, l2;
        cdf_forward(a.marg[0], a.clamp, o01, o01, l0);
        cdf_forward(a.marg[1], a.clamp, o12, o12, l1);
        cdf_forward(a.marg[nb], a.clamp, a012, a012, l2);
        dl += l0 + l1 + l2;
      }
      if (SMEM) {
        Qs[0] = o01; Qs[LDF] = o12; Qs[nb * LDF] = a012;
      } else {
        bo[0] = o01; bo[1] = o12; an[0] = a012;
      }
      a.o_x0[row * 3 + 0] = p0.x;
      a.o_x0[row * 3 + 1] = p0.y;
      a.o_x0[row * 3 + 2] = p0.z;
      a.o_R[row * 3 + 0] = alpha;
      a.o_R[row * 3 + 1] = beta;
      a.o_R[row * 3 + 2] = gamma;
    }
  }
  if (QUAD) {   // the four threads of a sample are adjacent lanes
    dl += __shfl_xor_sync(0xffffffffu, dl, 1);
    dl += __shfl_xor_sync(0xffffffffu, dl, 2);
  }
  if (live && g == 0) a.dlogp_out[row] = (a.dlogp_in ? a.dlogp_in[row] : 0.f) + dl;
  if (SMEM) {
    __syncthreads();
    const float* Qb = sm + W * LDF;
    float* gb = a.o_bonds + row0 * nb;
    float* ga = a.o_angles + row0 * na;
    float* gt = a.o_torsions + row0 * nt;
    walk_block_f(t, nb, nrow, [&](int m, int c) { gb[m * nb + c] = Qb[c * LDF + m]; });
    walk_block_f(t, na, nrow, [&](int m, int c) { ga[m * na + c] = Qb[(nb + c) * LDF + m]; });
    walk_block_f(t, nt, nrow, [&](int m, int c) { gt[m * nt + c] = Qb[(nb + na + c) * LDF + m]; });
    if (!global) {
      const int wf = a.fixed_w;
      float* gf = a.fixed_out + row0 * wf;
      walk_block_f(t, wf, nrow, [&](int m, int c) { gf[m * wf + c] = Qb[(nb + na + nt + c) * LDF + m]; });
    }
  }
  if (BULK) {
    fence_async_smem();
    __syncthreads();
    if (t == 0) {
      ic_bulk_s2g(a.o_bonds + row0 * nb, qb, (uint32_t)(FS * nb * 4));
      ic_bulk_s2g(a.o_angles + row0 * na, qb + FS * nb, (uint32_t)(FS * na * 4));
      ic_bulk_s2g(a.o_torsions + row0 * nt, qb + FS * (nb + na), (uint32_t)(FS * nt * 4));
      if (!global) ic_bulk_s2g(a.fixed_out + row0 * a.fixed_w, qb + FS * (nb + na + nt), (uint32_t)(FS * a.fixed_w * 4));
      ic_bulk_commit_wait_read();
    }
  }
}

static int fill_plan(const bgx_zplan* plan, long long batch, IcArgs& a) {
  if (!plan || plan->n_atoms < 4 || plan->n_rel != plan->n_atoms - 3 || !plan->rel || !plan->order || batch < 0)
    return BGX_ERR_INVALID;
  a.B = batch;
  a.kind = IC_GLOBAL;
  a.n_atoms = plan->n_atoms;
  a.n_rel = plan->n_rel;
  a.nb = plan->n_atoms - 1; a.na = plan->n_atoms - 2; a.nt = plan->n_atoms - 3;
  a.off_b = 2; a.off_a = 1;
  a.s0 = plan->seeds[0]; a.s1 = plan->seeds[1]; a.s2 = plan->seeds[2];
  a.rel = plan->rel;
  a.order = plan->order;
  a.normalize = plan->normalize_angles;
  a.eps = plan->eps;
  a.eps2 = plan->eps * plan->eps;
  a.slot_of_col = plan->slot_of_col;
  a.cmin = (float)(-1.0 + (double)plan->eps);
  a.cmax = (float)(1.0 - (double)plan->eps);
  return BGX_OK;
}

static int fill_relplan(const bgx_relplan* plan, long long batch, IcArgs& a) {
  if (!plan || plan->n_fixed < 3 || plan->n_rel < 1 || plan->n_atoms != plan->n_fixed + plan->n_rel || !plan->fixed ||
      !plan->rel || !plan->order || batch < 0 || plan->keepdims < 0 || plan->keepdims > 3 * plan->n_fixed)
    return BGX_ERR_INVALID;
  if (plan->keepdims > 0 && (!plan->mean || !plan->blacken || !plan->whiten)) return BGX_ERR_INVALID;
  a.B = batch;
  a.kind = IC_RELATIVE;
  a.n_atoms = plan->n_atoms;
  a.n_rel = plan->n_rel;
  a.nb = a.na = a.nt = plan->n_rel;
  a.off_b = a.off_a = 0;
  a.n_fixed = plan->n_fixed;
  a.keep = plan->keepdims;
  a.fixed_w = plan->keepdims ? plan->keepdims : 3 * plan->n_fixed;
  a.fixed = plan->fixed;
  a.mean = plan->mean; a.blacken = plan->blacken; a.whiten = plan->whiten;
  a.ld_whiten = plan->log_det_whiten;
  a.rel = plan->rel;
  a.order = plan->order;
  a.normalize = plan->normalize_angles;
  a.eps = plan->eps;
  a.eps2 = plan->eps * plan->eps;
  a.slot_of_col = nullptr;
  a.cmin = (float)(-1.0 + (double)plan->eps);
  a.cmax = (float)(1.0 - (double)plan->eps);
  return BGX_OK;
}

static bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// raise a kernel's dynamic shared-memory limit once per (kernel, size) instead of on every launch
template <typename K>
static int ensure_smem(K kern, size_t bytes) {
  static std::map<std::pair<int, const void*>, size_t> seen;      // per (device, kernel)
  if (bytes <= 48 * 1024) return BGX_OK;
  size_t& cur = seen[std::make_pair(device_slot(), (const void*)kern)];
  if (bytes <= cur) return BGX_OK;
  int rc = check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (!rc) cur = bytes;
  return rc;
}

// advance every per-sample pointer of `a` by `rows` samples (the tail launch after the bulk-tile launch)
static IcArgs skip_rows(const IcArgs& a, long long rows, bool to_xyz) {
  IcArgs r = a;
  r.B = a.B - rows;
  const long long W = 3LL * a.n_atoms;
  if (to_xyz) {
    r.bonds += rows * a.nb; r.angles += rows * a.na; r.torsions += rows * a.nt;
    if (r.x0) r.x0 += rows * a.x0_stride;
    if (r.R) r.R += rows * a.r_stride;
    if (r.fixed_in) r.fixed_in += rows * a.fixed_w;
    r.xyz += rows * W;
  } else {
    r.xyz_in += rows * W;
    r.o_bonds += rows * a.nb; r.o_angles += rows * a.na; r.o_torsions += rows * a.nt;
    if (r.o_x0) r.o_x0 += rows * 3;
    if (r.o_R) r.o_R += rows * 3;
    if (r.fixed_out) r.fixed_out += rows * a.fixed_w;
  }
  if (r.dlogp_in) r.dlogp_in += rows;
  r.dlogp_out += rows;
  return r;
}

static int ic_io_mode() {   // BGX_IC_BULK=0: element-wise staging everywhere (A/B switch)
  static const int m = [] { const char* e = getenv("BGX_IC_BULK"); return e ? atoi(e) : 1; }();
  return m;
}

// kb / ksm / kgl: the kernel with bulk-TMA tiles, with element-wise staging, on global memory
template <typename K>
static int launch_ic(K kb, K ksm, K kgl, const IcArgs& a, int extra_cols, int samples_per_cta, bool to_xyz,
                     cudaStream_t st) {
  if (a.B == 0) return BGX_OK;
  long long grid = (a.B + BT - 1) / BT;
  if (grid > 0x7fffffffLL) return BGX_ERR_UNSUPPORTED;
  size_t sb = sizeof(float) * (size_t)(3 * a.n_atoms + extra_cols) * (samples_per_cta + 1);
  if (a.kind == IC_GLOBAL && !a.slot_of_col) sb = (size_t)1 << 30;   // plans without a slot map: global-memory path
  if (sb > 200 * 1024) {
    kgl<<<(unsigned)grid, BT, 0, st>>>(a);
    return post_launch();
  }
  // ---- full tiles through bulk-TMA copies (row-major tiles: inputs + outputs of a tile, no padding)
  const int io_w = a.nb + a.na + a.nt + (a.kind == IC_GLOBAL ? 0 : a.fixed_w) + 3 * a.n_atoms;
  const size_t sb_bulk = sizeof(float) * (size_t)io_w * samples_per_cta;
  const bool aligned = to_xyz ? (al16(a.bonds) && al16(a.angles) && al16(a.torsions) && al16(a.xyz) &&
                                 (a.kind == IC_GLOBAL || al16(a.fixed_in)))
                              : (al16(a.xyz_in) && al16(a.o_bonds) && al16(a.o_angles) && al16(a.o_torsions) &&
                                 (a.kind == IC_GLOBAL || al16(a.fixed_out)));
  long long done = 0;
  if (ic_io_mode() && aligned && sb_bulk <= 200 * 1024 && a.B >= samples_per_cta) {
    const long long tiles = a.B / samples_per_cta;
    if (tiles > 0x7fffffffLL) return BGX_ERR_UNSUPPORTED;
    int rc = ensure_smem(kb, sb_bulk);
    if (rc) return rc;
    IcArgs full = a;
    full.B = tiles * samples_per_cta;
    kb<<<(unsigned)tiles, BT, sb_bulk, st>>>(full);
    rc = post_launch();
    if (rc) return rc;
    done = full.B;
    if (done == a.B) return BGX_OK;
  }
  // ---- the remaining rows (or everything) through the element-wise staged kernel
  const IcArgs rest = done ? skip_rows(a, done, to_xyz) : a;
  int rc = ensure_smem(ksm, sb);
  if (rc) return rc;
  const long long grid_s = (rest.B + samples_per_cta - 1) / samples_per_cta;
  if (grid_s > 0x7fffffffLL) return BGX_ERR_UNSUPPORTED;
  ksm<<<(unsigned)grid_s, BT, sb, st>>>(rest);
  return post_launch();
}

}  // namespace bgx

using namespace bgx;

extern "C" int bgx_ic_to_xyz(const bgx_zplan* plan, int64_t batch, const float* bonds, const float* angles,
                             const float* torsions, const float* x0, int32_t x0_stride, const float* R,
                             int32_t r_stride, float* xyz, const float* dlogp_in, float* dlogp_out,
                             void* stream) {
  IcArgs a{};
  int rc = fill_plan(plan, batch, a);
  if (rc) return rc;
  if (!bonds || !angles || !torsions || !x0 || !R || !xyz || !dlogp_out) return BGX_ERR_INVALID;
  a.bonds = bonds; a.angles = angles; a.torsions = torsions;
  a.x0 = x0; a.R = R; a.x0_stride = x0_stride; a.r_stride = r_stride;
  a.xyz = xyz; a.dlogp_in = dlogp_in; a.dlogp_out = dlogp_out;
  return launch_ic(ic_to_xyz_kernel<IO_BULK>, ic_to_xyz_kernel<IO_STAGED>, ic_to_xyz_kernel<IO_GLOBAL>, a, 0, BT, true,
                   (cudaStream_t)stream);
}

extern "C" int bgx_ic_to_xyz_mapped(const bgx_zplan* plan, const bgx_cdf_col* marginals, float clamp_lo,
                                    float clamp_hi, float logdet_min, int64_t batch, const float* bonds,
                                    const float* angles, const float* torsions, const float* x0, int32_t x0_stride,
                                    const float* R, int32_t r_stride, float* xyz, const float* dlogp_in,
                                    float* dlogp_out, void* stream) {
  IcArgs a{};
  int rc = fill_plan(plan, batch, a);
  if (rc) return rc;
  if (!marginals || !bonds || !angles || !torsions || !x0 || !R || !xyz || !dlogp_out) return BGX_ERR_INVALID;
  a.marg = marginals;
  a.clamp = {clamp_lo, clamp_hi, logdet_min};
  a.bonds = bonds; a.angles = angles; a.torsions = torsions;
  a.x0 = x0; a.R = R; a.x0_stride = x0_stride; a.r_stride = r_stride;
  a.xyz = xyz; a.dlogp_in = dlogp_in; a.dlogp_out = dlogp_out;
  return launch_ic(ic_to_xyz_kernel<IO_BULK>, ic_to_xyz_kernel<IO_STAGED>, ic_to_xyz_kernel<IO_GLOBAL>, a, 0, BT, true,
                   (cudaStream_t)stream);
}

extern "C" int bgx_relic_to_xyz(const bgx_relplan* plan, int64_t batch, const float* bonds, const float* angles,
                                const float* torsions, const float* fixed, float* xyz, const float* dlogp_in,
                                float* dlogp_out, void* stream) {
  IcArgs a{};
  int rc = fill_relplan(plan, batch, a);
  if (rc) return rc;
  if (!bonds || !angles || !torsions || !fixed || !xyz || !dlogp_out) return BGX_ERR_INVALID;
  a.bonds = bonds; a.angles = angles; a.torsions = torsions; a.fixed_in = fixed;
  a.xyz = xyz; a.dlogp_in = dlogp_in; a.dlogp_out = dlogp_out;
  return launch_ic(ic_to_xyz_kernel<IO_BULK>, ic_to_xyz_kernel<IO_STAGED>, ic_to_xyz_kernel<IO_GLOBAL>, a, a.keep, BT, true,
                   (cudaStream_t)stream);
}

extern "C" int bgx_ic_from_xyz(const bgx_zplan* plan, int64_t batch, const float* xyz, float* bonds,
                               float* angles, float* torsions, float* x0, float* R, const float* dlogp_in,
                               float* dlogp_out, void* stream) {
  IcArgs a{};
  int rc = fill_plan(plan, batch, a);
  if (rc) return rc;
  if (!xyz || !bonds || !angles || !torsions || !x0 || !R || !dlogp_out) return BGX_ERR_INVALID;
  a.xyz_in = xyz; a.o_bonds = bonds; a.o_angles = angles; a.o_torsions = torsions;
  a.o_x0 = x0; a.o_R = R; a.dlogp_in = dlogp_in; a.dlogp_out = dlogp_out;
  return launch_ic(ic_from_xyz_kernel<IO_BULK>, ic_from_xyz_kernel<IO_STAGED>, ic_from_xyz_kernel<IO_GLOBAL>, a,
                   3 * a.n_atoms - 6, FS, false, (cudaStream_t)stream);
}

extern "C" int bgx_ic_from_xyz_mapped(const bgx_zplan* plan, const bgx_cdf_col* marginals, float clamp_lo,
                                      float clamp_hi, float logdet_min, int64_t batch, const float* xyz, float* bonds,
                                      float* angles, float* torsions, float* x0, float* R, const float* dlogp_in,
                                      float* dlogp_out, void* stream) {
  IcArgs a{};
  int rc = fill_plan(plan, batch, a);
  if (rc) return rc;
  if (!marginals || !xyz || !bonds || !angles || !torsions || !x0 || !R || !dlogp_out) return BGX_ERR_INVALID;
  a.marg = marginals;
  a.clamp = {clamp_lo, clamp_hi, logdet_min};
  a.xyz_in = xyz; a.o_bonds = bonds; a.o_angles = angles; a.o_torsions = torsions;
  a.o_x0 = x0; a.o_R = R; a.dlogp_in = dlogp_in; a.dlogp_out = dlogp_out;
  return launch_ic(ic_from_xyz_kernel<IO_BULK>, ic_from_xyz_kernel<IO_STAGED>, ic_from_xyz_kernel<IO_GLOBAL>, a,
                   3 * a.n_atoms - 6, FS, false, (cudaStream_t)stream);
}

extern "C" int bgx_relic_from_xyz(const bgx_relplan* plan, int64_t batch, const float* xyz, float* bonds,
                                  float* angles, float* torsions, float* fixed, const float* dlogp_in,
                                  float* dlogp_out, void* stream) {
  IcArgs a{};
  int rc = fill_relplan(plan, batch, a);
  if (rc) return rc;
  if (!xyz || !bonds || !angles || !torsions || !fixed || !dlogp_out) return BGX_ERR_INVALID;
  a.xyz_in = xyz; a.o_bonds = bonds; a.o_angles = angles; a.o_torsions = torsions; a.fixed_out = fixed;
  a.dlogp_in = dlogp_in; a.dlogp_out = dlogp_out;
  return launch_ic(ic_from_xyz_kernel<IO_BULK>, ic_from_xyz_kernel<IO_STAGED>, ic_from_xyz_kernel<IO_GLOBAL>, a,
                   3 * a.n_rel + a.fixed_w, FS, false, (cudaStream_t)stream);
}
