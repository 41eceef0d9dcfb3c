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
// Layout: the [tile x 3N] coordinate block of a CTA is contiguous in global memory; it is
// staged through shared memory as P[c][t] (leading dim BT+1: conflict-free both for the
// per-thread phase (fixed c, consecutive t) and for the coalesced copy phase).
#include "bgx_common.cuh"

namespace bgx {

constexpr int BT = 128;       // samples (threads) per CTA
constexpr int LDT = BT + 1;
constexpr float PI_F = 3.14159265358979323846f;
constexpr float TWO_PI_F = 6.28318530717958647692f;

struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float norm_c(V3 a, float eps) { return fmaxf(sqrtf(dot(a, a)), eps); }
// 1 / max(|a|, eps) with one MUFU.RSQ (eps^2 = 1e-14 is representable)
__device__ __forceinline__ float inv_norm_c(V3 a, float eps2) { return rsqrtf(fmaxf(dot(a, a), eps2)); }

struct IcArgs {
  long long B;
  int n_atoms, n_rel;
  int s0, s1, s2;
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

template <bool SMEM>
struct PosStore {
  float* base;
  long long stride_c;  // distance between consecutive coordinates
  __device__ __forceinline__ V3 get(int atom) const {
    const float* p = base + (long long)(3 * atom) * stride_c;
    return {p[0], p[stride_c], p[2 * stride_c]};
  }
  __device__ __forceinline__ void set(int atom, V3 v) const {
    float* p = base + (long long)(3 * atom) * stride_c;
    p[0] = v.x;
    p[stride_c] = v.y;
    p[2 * stride_c] = v.z;
  }
};

// ---------------------------------------------------------------- IC -> Cartesian
// division-free walk of thread t over the elements e = t, t+BT, ... of a [rows x W] block
template <typename F>
__device__ __forceinline__ void walk_block(int t, int W, int rows, F&& body) {
  int m = t / W, c = t - m * W;
  const int dm = BT / W, dc = BT - dm * W;
  while (m < rows) {
    body(m, c);
    m += dm; c += dc;
    if (c >= W) { c -= W; ++m; }
  }
}

template <bool SMEM>
__global__ void __launch_bounds__(BT) ic_to_xyz_kernel(const IcArgs a) {
  extern __shared__ float sm[];
  const int t = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * BT;
  const long long row = row0 + t;
  const bool live = row < a.B;
  const int N = a.n_atoms, nb = N - 1, na = N - 2, nt = N - 3;
  const int nrow = (int)min((long long)BT, a.B - row0);
  PosStore<SMEM> pos;
  if (SMEM) {
    // Every placed atom owns exactly three internal coordinates and three Cartesian coordinates:
    // its (bond, angle, torsion) are staged in the very shared-memory slots its (x, y, z) will
    // overwrite.  Coalesced loads of the three [tile x n] blocks, scattered by column -> slot.
    pos.base = sm + t; pos.stride_c = LDT;
    const float* gb = a.bonds + row0 * nb;
    const float* ga = a.angles + row0 * na;
    const float* gt = a.torsions + row0 * nt;
    walk_block(t, nb, nrow, [&](int m, int c) { sm[a.slot_of_col[c] * LDT + m] = __ldg(gb + m * nb + c); });
    walk_block(t, na, nrow, [&](int m, int c) { sm[a.slot_of_col[nb + c] * LDT + m] = __ldg(ga + m * na + c); });
    walk_block(t, nt, nrow, [&](int m, int c) { sm[a.slot_of_col[nb + na + c] * LDT + m] = __ldg(gt + m * nt + c); });
    __syncthreads();
  } else {
    pos.base = a.xyz + row * (long long)(3 * N); pos.stride_c = 1;
  }

  if (live) {
    const float* bo = a.bonds + row * nb;
    const float* an = a.angles + row * na;
    const float* to = a.torsions + row * nt;
    const float* x0p = a.x0 + row * a.x0_stride;
    const float* rp = a.R + row * a.r_stride;
    float d01, d12, a012;
    if (SMEM) {
      d01 = pos.base[(3 * a.s1) * LDT]; d12 = pos.base[(3 * a.s2) * LDT]; a012 = pos.base[(3 * a.s2 + 1) * LDT];
    } else {
      d01 = bo[0]; d12 = bo[1]; a012 = an[0];
    }
    float alpha = rp[0], beta = rp[1], gamma = rp[2];
    float dl = 0.f;
    if (a.normalize) {
      a012 *= PI_F;
      alpha = alpha * TWO_PI_F - PI_F;
      gamma = gamma * TWO_PI_F - PI_F;
      dl = (float)((double)na * 1.1447298858494002 + (double)(nt + 2) * 1.8378770664093453);
    }
    float sa, ca, sb, cb, sg, cg, s012, c012;
    sincosf(alpha, &sa, &ca);
    sincosf(acosf(beta), &sb, &cb);
    sincosf(gamma, &sg, &cg);
    sincosf(a012, &s012, &c012);
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
    for (int q = 0; q < a.n_rel; ++q) {
      const int r = a.order[q];
      const int4 z = *reinterpret_cast<const int4*>(a.rel + 4 * r);
      float d, ang, tor;
      if (SMEM) {
        const float* ic = pos.base + (3 * z.x) * LDT;
        d = ic[0]; ang = ic[LDT]; tor = ic[2 * LDT];
      } else {
        d = bo[2 + r]; ang = an[1 + r]; tor = to[r];
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
}

// ---------------------------------------------------------------- Cartesian -> IC
// All z-matrix rows of a sample are independent, so FOUR threads share a sample (rows r = g, g+4, ...;
// thread g = 3 also does the reference frame): 32 samples per 128-thread CTA, 16.6 KB of shared
// memory per CTA (13 CTAs / 52 warps per SM) instead of one thread walking 19 rows.
constexpr int FS = 32;        // samples per CTA (from_xyz)
constexpr int LDF = FS + 1;

template <typename F>
__device__ __forceinline__ void walk_block_f(int t, int W, int rows, F&& body) {
  int m = t / W, c = t - m * W;
  const int dm = BT / W, dc = BT - dm * W;
  while (m < rows) {
    body(m, c);
    m += dm; c += dc;
    if (c >= W) { c -= W; ++m; }
  }
}

template <bool SMEM>
__global__ void __launch_bounds__(BT) ic_from_xyz_kernel(const IcArgs a) {
  extern __shared__ float sm[];
  const int t = threadIdx.x;
  const int N = a.n_atoms, nb = N - 1, na = N - 2, nt = N - 3;
  const int W = 3 * N;
  const int spb = SMEM ? FS : BT;                    // samples per CTA
  const long long row0 = (long long)blockIdx.x * spb;
  const int s_loc = SMEM ? (t >> 2) : t, g = SMEM ? (t & 3) : 0, ng = SMEM ? 4 : 1;
  const long long row = row0 + s_loc;
  const bool live = row < a.B;
  const int nrow = (int)min((long long)spb, a.B - row0);
  float* Qs = sm + W * LDF + s_loc;     // staged outputs: column c of bonds|angles|torsions at Qs[c * LDF]
  PosStore<SMEM> pos;
  if (SMEM) {
    const float* gx = a.xyz_in + row0 * W;
    walk_block_f(t, W, nrow, [&](int m, int c) { sm[c * LDF + m] = __ldg(gx + m * W + c); });
    __syncthreads();
    pos.base = sm + s_loc;
    pos.stride_c = LDF;
  } else {
    pos.base = const_cast<float*>(a.xyz_in) + row * (long long)(3 * N);
    pos.stride_c = 1;
  }
  float dl = 0.f;
  if (live) {
    float* bo = a.o_bonds + row * nb;
    float* an = a.o_angles + row * na;
    float* to = a.o_torsions + row * nt;
    for (int r = g; r < a.n_rel; r += ng) {
      const int4 z = *reinterpret_cast<const int4*>(a.rel + 4 * r);
      const V3 xi = pos.get(z.x), xj = pos.get(z.y), xk = pos.get(z.z), xl = pos.get(z.w);
      const V3 r12 = xi - xj, r32 = xk - xj;
      const float i12 = inv_norm_c(r12, a.eps2), i32 = inv_norm_c(r32, a.eps2);
      const float n12 = fmaxf(dot(r12, r12) * i12, a.eps);        // |r12| clamped like the reference
      float c = dot(r12, r32) * (i12 * i32);
      c = fminf(fmaxf(c, a.cmin), a.cmax);
      float ang = acosf(c);
      const float sin2 = 1.f - c * c;
      // torsion (ic_helper.py:220-278): b0 = xi - xj, b1 = xk - xj, b2 = xl - xk
      const V3 b2 = xl - xk;
      const V3 u = r32 * i32;
      const V3 v = r12 - u * dot(r12, u);
      const V3 w = b2 - u * dot(b2, u);
      float tor = atan2f(dot(cross(u, v), w), dot(v, w));
      if (a.normalize) {
        ang *= (1.f / PI_F);
        tor = (tor + PI_F) * (1.f / TWO_PI_F);
      }
      if (SMEM) {
        Qs[(2 + r) * LDF] = n12; Qs[(nb + 1 + r) * LDF] = ang; Qs[(nb + na + r) * LDF] = tor;
      } else {
        bo[2 + r] = n12; an[1 + r] = ang; to[r] = tor;
      }
      dl -= __logf(n12 * n12 * sqrtf(sin2));     // 2 ln b + ln sin a
    }
    if (g == ng - 1) {
      const V3 p0 = pos.get(a.s0), p1 = pos.get(a.s1), p2 = pos.get(a.s2);
      const V3 e01 = p1 - p0, e12 = p2 - p1;
      const float d01 = norm_c(e01, a.eps), d12 = norm_c(e12, a.eps);
      // angle at p1 between p0 and p2
      const V3 ra = p0 - p1;
      float c = dot(ra * (1.f / norm_c(ra, a.eps)), e12 * (1.f / d12));
      c = fminf(fmaxf(c, a.cmin), a.cmax);
      float a012 = acosf(c);
      const float sin012 = sqrtf(1.f - c * c);
      // tripod (ic_helper.py:114-138): e1 = (p1-p0)/|.|, e2 = ((p2-p0) x e1)/|.|, e3 = e2 x e1
      const V3 e1 = e01 * (1.f / d01);
      V3 e2 = cross(p2 - p0, e1);
      e2 = e2 * (1.f / norm_c(e2, a.eps));
      const V3 e3 = cross(e2, e1);
      // basis (X, Y, Z) = (-e3, -e2, e1); euler (ic_helper.py:330-341)
      float alpha = atan2f(e1.x, -e1.y);
      const float beta = e1.z;
      float gamma = atan2f(-e3.z, -e2.z);
      dl -= 2.f * logf(d01) + 2.f * logf(d12) + logf(sin012);
      if (a.normalize) {
        a012 *= (1.f / PI_F);
        alpha = (alpha + PI_F) * (1.f / TWO_PI_F);
        gamma = (gamma + PI_F) * (1.f / TWO_PI_F);
        dl -= (float)((double)na * 1.1447298858494002 + (double)(nt + 2) * 1.8378770664093453);
      }
      if (SMEM) {
        Qs[0] = d01; Qs[LDF] = d12; Qs[nb * LDF] = a012;
      } else {
        bo[0] = d01; bo[1] = d12; an[0] = a012;
      }
      a.o_x0[row * 3 + 0] = p0.x;
      a.o_x0[row * 3 + 1] = p0.y;
      a.o_x0[row * 3 + 2] = p0.z;
      a.o_R[row * 3 + 0] = alpha;
      a.o_R[row * 3 + 1] = beta;
      a.o_R[row * 3 + 2] = gamma;
    }
  }
  if (SMEM) {   // the four threads of a sample are adjacent lanes
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
  }
}

static int fill_plan(const bgx_zplan* plan, long long batch, IcArgs& a) {
  if (!plan || plan->n_atoms < 4 || plan->n_rel != plan->n_atoms - 3 || !plan->rel || !plan->order || batch < 0)
    return BGX_ERR_INVALID;
  a.B = batch;
  a.n_atoms = plan->n_atoms;
  a.n_rel = plan->n_rel;
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

template <typename KS, typename KG>
static int launch_ic(KS ksm, KG kgl, const IcArgs& a, int extra_cols, int samples_per_cta, cudaStream_t st) {
  if (a.B == 0) return BGX_OK;
  long long grid = (a.B + BT - 1) / BT;
  if (grid > 0x7fffffffLL) return BGX_ERR_UNSUPPORTED;
  size_t sb = sizeof(float) * (size_t)(3 * a.n_atoms + extra_cols) * (samples_per_cta + 1);
  if (!a.slot_of_col) sb = (size_t)1 << 30;      // plans without a slot map use the global-memory path
  if (sb <= 200 * 1024) {
    if (sb > 48 * 1024) {  // (both kernels share this instantiation: no static cache here)
      int rc = check(cudaFuncSetAttribute(ksm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
      if (rc) return rc;
    }
    const long long grid_s = (a.B + samples_per_cta - 1) / samples_per_cta;
    if (grid_s > 0x7fffffffLL) return BGX_ERR_UNSUPPORTED;
    ksm<<<(unsigned)grid_s, BT, sb, st>>>(a);
  } else {
    kgl<<<(unsigned)grid, BT, 0, st>>>(a);
  }
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
  return launch_ic(ic_to_xyz_kernel<true>, ic_to_xyz_kernel<false>, a, 0, BT, (cudaStream_t)stream);
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
  return launch_ic(ic_from_xyz_kernel<true>, ic_from_xyz_kernel<false>, a, 3 * a.n_atoms - 6, FS, (cudaStream_t)stream);
}
