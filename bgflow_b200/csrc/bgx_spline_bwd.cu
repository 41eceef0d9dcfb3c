// Backward of the rational-quadratic spline transform (training path).
//
// One thread per (sample, transformed dim): re-evaluates the spline from the conditioner output
// P (reference column layout [W: D x K | H: D x K | S: D x K | S_last: n_noncircular],
// spline.py:113-125) and applies the chain rule by hand:
//
//   (g_out, g_dlogp) -> scalar partials wrt (x, knot_b, width_b, height knot_b, height_b, d_b, d_{b+1})
//                    -> knots (cumulative sums)  -> softmax            -> dW, dH
//                    -> derivatives (softplus)                          -> dS, dS_last
//
// The root branch (bgflow forward) is differentiated implicitly: theta solves
// ch + h N(theta)/den(theta) = x, so d theta/d p = -(dG/dp)/(dG/d theta).  HBM-bound: reads and
// writes one [B, 3KD+n] row block each.  The conditioner's own backward (dense GEMMs) stays with
// the caller (bgflow_b200/autograd.py).  Checked against autograd of the oracle in fp64
// (tests/test_gpu_autograd.py).
#include "bgx_common.cuh"
#include "bgx_fastmath.cuh"

namespace bgx {

struct SplineBwdArgs {
  long long B;
  int D, K;
  const float* P;
  long long p_stride;
  const float* y;
  const float* g_out;
  const float* g_dl;       // [B] or null
  const int* end_col;      // [D]: column of P holding the last slope of dim d
  float* dP;
  float* dy;
  float left, right, bottom, top, min_w, min_h, min_d, beta;
  int root;
  int zero_from;      // >= 0: columns [zero_from, p_stride) of dP are written as zeros (row padding that later GEMMs read)
};

// VEC (K == MAXK == 8, rows of P / dP 16-byte aligned): a dim's 8 widths / heights / slopes are 32 contiguous bytes,
// read and written as two float4 each, so every lane moves whole sectors (the scalar form touches a sector eight
// times with 4-byte accesses; measured 1.65 TB/s on the 2 x 217 MB of a B = 65536, 33-dim block).
template <int MAXK, bool VEC>
__global__ void __launch_bounds__(256) spline_backward_kernel(const SplineBwdArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B * a.D) return;
  const long long r = idx / a.D;
  const int d = (int)(idx - r * a.D);
  const int K = VEC ? 8 : a.K, KD = K * a.D;
  const float* Pr = a.P + r * a.p_stride;
  float* dPr = a.dP + r * a.p_stride;
  const int ecol = a.end_col[d];

  // VEC is the training drivers' path: the MUFU forms of the fused forward kernels (ex2 / lg2 / rcp / sqrt approx,
  // ~1e-6 relative) instead of expf / logf / IEEE division — the kernel issued ~990 instructions per thread, 56 % of
  // the issue slots, and a third of them were the software expansions of those functions
  constexpr bool FAST = VEC;
  auto ex = [](float v) { return FAST ? ex2_fast(v * 1.4426950408889634f) : expf(v); };
  auto rc = [](float v) { return FAST ? rcp_fast(v) : 1.f / v; };
  auto sq = [](float v) { return FAST ? sqrt_fast(v) : sqrtf(v); };
  auto dv = [](float n, float d) { return FAST ? n * rcp_fast(d) : n / d; };      // (the precise path keeps IEEE division)
  auto splus = [&](float x, float beta, float inv_beta) {
    if (!FAST) return softplus_beta(x, beta, inv_beta);
    const float bx = beta * x;
    return bx > 20.f ? x : lg2_fast(1.f + ex2_fast(bx * 1.4426950408889634f)) * 0.6931471805599453f * inv_beta;
  };
  float sw[MAXK], sh[MAXK], us[MAXK + 1], cw[MAXK + 1], ch[MAXK + 1];
  float mw = -INFINITY, mh = -INFINITY;
  if (VEC) {
    auto load8 = [](const float* p, float* o) {
      const float4 lo = __ldg(reinterpret_cast<const float4*>(p)), hi = __ldg(reinterpret_cast<const float4*>(p) + 1);
      o[0] = lo.x; o[1] = lo.y; o[2] = lo.z; o[3] = lo.w; o[4] = hi.x; o[5] = hi.y; o[6] = hi.z; o[7] = hi.w;
    };
    load8(Pr + d * 8, sw);
    load8(Pr + KD + d * 8, sh);
    load8(Pr + 2 * KD + d * 8, us);
  } else {
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < K) {
        sw[k] = Pr[d * K + k];
        sh[k] = Pr[KD + d * K + k];
        us[k] = Pr[2 * KD + d * K + k];
      }
  }
#pragma unroll
  for (int k = 0; k < MAXK; ++k)
    if (k < K) {
      mw = fmaxf(mw, sw[k]);
      mh = fmaxf(mh, sh[k]);
    }
  const float us_end = Pr[ecol];
  float zw = 0.f, zh = 0.f;
#pragma unroll
  for (int k = 0; k < MAXK; ++k)
    if (k < K) {
      sw[k] = ex(sw[k] - mw);
      sh[k] = ex(sh[k] - mh);
      zw += sw[k];
      zh += sh[k];
    }
  const float izw = rc(zw), izh = rc(zh);
  const float fw = 1.f - a.min_w * K, fh = 1.f - a.min_h * K;
  const float span_w = a.right - a.left, span_h = a.top - a.bottom;
  float accw = 0.f, acch = 0.f;
  cw[0] = a.left;
  ch[0] = a.bottom;
#pragma unroll
  for (int k = 0; k < MAXK; ++k)
    if (k < K) {
      sw[k] *= izw;
      sh[k] *= izh;
      accw += a.min_w + fw * sw[k];
      acch += a.min_h + fh * sh[k];
      cw[k + 1] = (k == K - 1) ? a.right : a.left + span_w * accw;
      ch[k + 1] = (k == K - 1) ? a.top : a.bottom + span_h * acch;
    }

  const float yin = a.y[idx];
  const float x = fminf(fmaxf(yin, a.left), a.right);
  // bin: number of knots <= x, minus one (last knot nudged by 1e-6 like nflows' searchsorted)
  int b = -1;
#pragma unroll
  for (int k = 0; k <= MAXK; ++k)
    if (k <= K) {
      const float loc = (a.root ? ch[k] : cw[k]) + (k == K ? 1e-6f : 0.f);
      b += (x >= loc) ? 1 : 0;
    }
  b = min(max(b, 0), K - 1);
  float in_cw = 0.f, in_cw1 = 0.f, in_ch = 0.f, in_ch1 = 0.f, u0 = 0.f, u1 = us_end;
#pragma unroll
  for (int k = 0; k < MAXK; ++k)
    if (k < K) {
      if (k == b) {
        in_cw = cw[k];
        in_cw1 = cw[k + 1];
        in_ch = ch[k];
        in_ch1 = ch[k + 1];
        u0 = us[k];
      }
      if (k == b + 1) u1 = us[k];
    }
  const float in_w = in_cw1 - in_cw, in_h = in_ch1 - in_ch;
  const float inv_beta = rc(a.beta);
  const float d0 = a.min_d + splus(u0, a.beta, inv_beta);
  const float d1 = a.min_d + splus(u1, a.beta, inv_beta);
  const float iw = rc(in_w);
  const float delta = dv(in_h, in_w);
  const float ss = d0 + d1 - 2.f * delta;
  float th;
  if (a.root) {
    const float q = x - in_ch;
    const float qa = q * ss + in_h * (delta - d0);
    const float qb = in_h * d0 - q * ss;
    const float qc = -delta * q;
    th = dv(2.f * qc, -qb - sq(fmaxf(qb * qb - 4.f * qa * qc, 0.f)));
  } else {
    th = dv(x - in_cw, in_w);
  }
  const float om = 1.f - th;
  const float t1 = th * om;
  const float den = delta + ss * t1;
  const float N = delta * th * th + d0 * t1;
  const float A = d1 * th * th + 2.f * delta * t1 + d0 * om * om;
  const float iA = rc(A), iden = rc(den);
  const float go = a.g_out[idx];
  const float gL = (a.g_dl ? a.g_dl[r] : 0.f) * (a.root ? -1.f : 1.f);
  const float den_th = ss * (1.f - 2.f * th);
  const float L_th = (2.f * d1 * th + 2.f * delta * (1.f - 2.f * th) - 2.f * d0 * om) * iA - 2.f * den_th * iden;
  const float L_de = dv(2.f, delta) + 2.f * t1 * iA - 2.f * (1.f - 2.f * t1) * iden;
  const float L_d0 = om * om * iA - 2.f * t1 * iden;
  const float L_d1 = th * th * iA - 2.f * t1 * iden;
  const float N_th = 2.f * delta * th + d0 * (1.f - 2.f * th);
  const float hid2 = in_h * iden * iden;
  const float G_th = hid2 * (N_th * den - N * den_th);
  const float G_de = hid2 * (th * th * den - N * (1.f - 2.f * t1));
  const float G_d0 = hid2 * (t1 * den - N * t1);
  const float G_d1 = hid2 * (-N * t1);
  const float G_h = N * iden;
  float g_x, g_cw, g_w, g_ch, g_h, Gde, Gd0, Gd1;
  if (!a.root) {
    const float Gth = go * G_th + gL * L_th;
    Gde = go * G_de + gL * L_de;
    Gd0 = go * G_d0 + gL * L_d0;
    Gd1 = go * G_d1 + gL * L_d1;
    g_x = Gth * iw;
    g_cw = -g_x;
    g_w = -g_x * th;
    g_h = go * G_h;
    g_ch = go;
  } else {
    const float rr = dv(go * in_w + gL * L_th, G_th);
    g_x = rr;
    g_ch = -rr;
    g_h = -rr * G_h;
    Gde = gL * L_de - rr * G_de;
    Gd0 = gL * L_d0 - rr * G_d0;
    Gd1 = gL * L_d1 - rr * G_d1;
    g_w = go * th;
    g_cw = go;
  }
  g_h += Gde * iw;
  g_w -= Gde * delta * iw;

  // knots -> softmax.  knot b gets (g_c - g_width), knot b+1 gets g_width; only interior knots move.
  const float kw_b = (g_cw - g_w) * span_w * fw, kw_b1 = (b + 1 <= K - 1) ? g_w * span_w * fw : 0.f;
  const float kh_b = (g_ch - g_h) * span_h * fh, kh_b1 = (b + 1 <= K - 1) ? g_h * span_h * fh : 0.f;
  float dotw = 0.f, doth = 0.f;
#pragma unroll
  for (int k = 0; k < MAXK; ++k)
    if (k < K) {
      const float gw = (k < b ? kw_b : 0.f) + (k <= b ? kw_b1 : 0.f);
      const float gh = (k < b ? kh_b : 0.f) + (k <= b ? kh_b1 : 0.f);
      dotw += gw * sw[k];
      doth += gh * sh[k];
    }
  const float sg0 = dv(Gd0, 1.f + ex(-a.beta * u0));
  const float sg1 = dv(Gd1, 1.f + ex(-a.beta * u1));
  const bool end_hit = (b + 1 == K);
  const bool circular = ecol < 3 * KD;
  float ow[MAXK], oh[MAXK], os[MAXK];
#pragma unroll
  for (int k = 0; k < MAXK; ++k)
    if (k < K) {
      const float gw = (k < b ? kw_b : 0.f) + (k <= b ? kw_b1 : 0.f);
      const float gh = (k < b ? kh_b : 0.f) + (k <= b ? kh_b1 : 0.f);
      ow[k] = sw[k] * (gw - dotw);
      oh[k] = sh[k] * (gh - doth);
      float gs = (k == b ? sg0 : 0.f) + (k == b + 1 ? sg1 : 0.f);
      if (k == 0 && circular && end_hit) gs += sg1;
      os[k] = gs;
    }
  if (VEC) {
    auto store8 = [](float* p, const float* o) {
      reinterpret_cast<float4*>(p)[0] = make_float4(o[0], o[1], o[2], o[3]);
      reinterpret_cast<float4*>(p)[1] = make_float4(o[4], o[5], o[6], o[7]);
    };
    store8(dPr + d * 8, ow);
    store8(dPr + KD + d * 8, oh);
    store8(dPr + 2 * KD + d * 8, os);
  } else {
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < K) {
        dPr[d * K + k] = ow[k];
        dPr[KD + d * K + k] = oh[k];
        dPr[2 * KD + d * K + k] = os[k];
      }
  }
  if (!circular) dPr[ecol] = end_hit ? sg1 : 0.f;
  if (a.zero_from >= 0 && d == 0)
    for (long long c = a.zero_from; c < a.p_stride; ++c) dPr[c] = 0.f;
  a.dy[idx] = (yin >= a.left && yin <= a.right) ? g_x : 0.f;
}

}  // namespace bgx

namespace bgx {
int spline_backward_launch(int64_t batch, int32_t d_t, const float* params, int64_t params_stride, const float* y,
                           const float* g_out, const float* g_dlogp, const int32_t* end_slope_col, const bgx_spline_cfg* cfg,
                           int flags, float* d_params, float* d_y, int zero_from, void* stream);
}
extern "C" int bgx_spline_backward(int64_t batch, int32_t d_t, const float* params, int64_t params_stride,
                                   const float* y, const float* g_out, const float* g_dlogp,
                                   const int32_t* end_slope_col, const bgx_spline_cfg* cfg, int flags,
                                   float* d_params, float* d_y, void* stream) {
  return bgx::spline_backward_launch(batch, d_t, params, params_stride, y, g_out, g_dlogp, end_slope_col, cfg, flags,
                                     d_params, d_y, -1, stream);
}
int bgx::spline_backward_launch(int64_t batch, int32_t d_t, const float* params, int64_t params_stride, const float* y,
                                const float* g_out, const float* g_dlogp, const int32_t* end_slope_col,
                                const bgx_spline_cfg* cfg, int flags, float* d_params, float* d_y, int zero_from,
                                void* stream) {
  using namespace bgx;
  if (!cfg || batch < 0 || d_t <= 0 || cfg->n_bins < 1 || cfg->n_bins > 48) return BGX_ERR_INVALID;
  if (batch == 0) return BGX_OK;
  if (!params || !y || !g_out || !end_slope_col || !d_params || !d_y) return BGX_ERR_INVALID;
  if (params_stride < 3LL * cfg->n_bins * d_t) return BGX_ERR_INVALID;
  SplineBwdArgs a{};
  a.B = batch; a.D = d_t; a.K = cfg->n_bins;
  a.P = params; a.p_stride = params_stride; a.y = y; a.g_out = g_out; a.g_dl = g_dlogp;
  a.end_col = end_slope_col; a.dP = d_params; a.dy = d_y;
  a.left = cfg->left; a.right = cfg->right; a.bottom = cfg->bottom; a.top = cfg->top;
  a.min_w = cfg->min_bin_width; a.min_h = cfg->min_bin_height; a.min_d = cfg->min_derivative;
  a.beta = cfg->identity_init ? logf(2.f) / (1.f - cfg->min_derivative) : 1.f;
  a.root = (flags & BGX_FLAG_INVERSE) ? 0 : 1;
  a.zero_from = zero_from;
  const long long n = batch * (long long)d_t;
  const unsigned grid = (unsigned)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = a.K == 8 && params_stride % 4 == 0 && (((uintptr_t)params | (uintptr_t)d_params) & 15) == 0;
  if (vec) spline_backward_kernel<8, true><<<grid, 256, 0, st>>>(a);
  else if (a.K <= 8) spline_backward_kernel<8, false><<<grid, 256, 0, st>>>(a);
  else if (a.K <= 16) spline_backward_kernel<16, false><<<grid, 256, 0, st>>>(a);
  else spline_backward_kernel<48, false><<<grid, 256, 0, st>>>(a);
  return post_launch();
}
