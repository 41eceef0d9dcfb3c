// Two spline evaluations per thread with Blackwell's packed fp32 instructions.
//
// sm_100 adds add / mul / fma .f32x2 (SASS FADD2 / FMUL2 / FFMA2): one issue slot, two fp32 lanes held in an
// aligned register pair.  tools/f32x2_probe.cu measures them at the SAME lane throughput as the scalar forms
// (124 lane-ops per cycle per SM either way), i.e. the FMA pipe is no faster, but a packed instruction takes ONE
// issue slot for two results.  The fused coupling kernels are issue-bound in their epilogue (ncu: issue slots
// 54-59 % busy with the FMA and ALU pipes at ~25 % each, profiles/), so evaluating the rational-quadratic spline of
// TWO transformed dims of a sample at once — lane 0 = dim A, lane 1 = dim B — removes about 45 % of the epilogue's
// instructions: every add / mul / fma of rqs_eval_reg is issued once for both dims; MUFU, min / max, compares and
// selects stay scalar (no packed forms exist).
//
// Same algorithm and the same operation order per lane as rqs_eval_reg<ROOT, true> (bgx_spline_reg.cuh), so both give
// bit-identical results for a dim (packed fma.rn / add.rn / mul.rn round like the scalar instructions);
// tests/native/ checks that on the host build and tests/test_gpu_pair.py on the device.
#pragma once
#include <stdint.h>

#include "bgx_spline_reg.cuh"

namespace bgx {

// p[k] = (parameter k of dim A, parameter k of dim B), bias already added; x = (input A, input B), clamped to the domain
template <bool ROOT>
BGX_HD void rqs_eval_reg2(const F2 (&p)[PS], const SplineK& c, F2 x, F2& y, F2& lad) {
  // ---- softmax numerators of widths and heights
  float mwa = fmaxf(fmaxf(fmaxf(lo(p[0]), lo(p[1])), fmaxf(lo(p[2]), lo(p[3]))), fmaxf(fmaxf(lo(p[4]), lo(p[5])), fmaxf(lo(p[6]), lo(p[7]))));
  float mwb = fmaxf(fmaxf(fmaxf(hi(p[0]), hi(p[1])), fmaxf(hi(p[2]), hi(p[3]))), fmaxf(fmaxf(hi(p[4]), hi(p[5])), fmaxf(hi(p[6]), hi(p[7]))));
  float mha = fmaxf(fmaxf(fmaxf(lo(p[8]), lo(p[9])), fmaxf(lo(p[10]), lo(p[11]))), fmaxf(fmaxf(lo(p[12]), lo(p[13])), fmaxf(lo(p[14]), lo(p[15]))));
  float mhb = fmaxf(fmaxf(fmaxf(hi(p[8]), hi(p[9])), fmaxf(hi(p[10]), hi(p[11]))), fmaxf(fmaxf(hi(p[12]), hi(p[13])), fmaxf(hi(p[14]), hi(p[15]))));
  const F2 l2e = bc2(LOG2E);
  const F2 nmw = mul2(f2(-mwa, -mwb), l2e), nmh = mul2(f2(-mha, -mhb), l2e);
  F2 pw[NB], ph[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    pw[k] = map2(fma2(p[k], l2e, nmw), [](float v) { return ex2_fast(v); });
    ph[k] = map2(fma2(p[NB + k], l2e, nmh), [](float v) { return ex2_fast(v); });
  }
#pragma unroll
  for (int k = 1; k < NB; ++k) {
    pw[k] = add2(pw[k], pw[k - 1]);
    ph[k] = add2(ph[k], ph[k - 1]);
  }
  const F2 aw = mul2(bc2(c.wscale), map2(pw[NB - 1], [](float v) { return rcp_fast(v); }));
  const F2 ah = mul2(bc2(c.hscale), map2(ph[NB - 1], [](float v) { return rcp_fast(v); }));
  // ---- knots 1..NB-1 (knot 0 = left / bottom, knot NB = right / top exactly)
  F2 kw[NB + 1], kh[NB + 1];
  kw[0] = bc2(c.left); kh[0] = bc2(c.bottom); kw[NB] = bc2(c.right); kh[NB] = bc2(c.top);
#pragma unroll
  for (int k = 1; k < NB; ++k) {
    kw[k] = fma2(aw, pw[k - 1], bc2(fmaf(c.wstep, (float)k, c.left)));
    kh[k] = fma2(ah, ph[k - 1], bc2(fmaf(c.hstep, (float)k, c.bottom)));
  }
  // ---- binary search over the 8 bins, per lane: the largest k with x >= knot[k]
  const F2* sl = p + 2 * NB;
  const float xa = lo(x), xb = hi(x);
  const bool a4 = xa >= (ROOT ? lo(kh[4]) : lo(kw[4])), b4 = xb >= (ROOT ? hi(kh[4]) : hi(kw[4]));
  F2 qw[5], qh[5], qs[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    qw[q] = sel2(a4, b4, kw[4 + q], kw[q]);
    qh[q] = sel2(a4, b4, kh[4 + q], kh[q]);
    qs[q] = sel2(a4, b4, sl[4 + q], sl[q]);
  }
  const bool a2 = xa >= (ROOT ? lo(qh[2]) : lo(qw[2])), b2 = xb >= (ROOT ? hi(qh[2]) : hi(qw[2]));
  F2 bw[3], bh[3], bs[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    bw[q] = sel2(a2, b2, qw[2 + q], qw[q]);
    bh[q] = sel2(a2, b2, qh[2 + q], qh[q]);
    bs[q] = sel2(a2, b2, qs[2 + q], qs[q]);
  }
  const bool a1 = xa >= (ROOT ? lo(bh[1]) : lo(bw[1])), b1 = xb >= (ROOT ? hi(bh[1]) : hi(bw[1]));
  const F2 w_lo = sel2(a1, b1, bw[1], bw[0]), w_hi = sel2(a1, b1, bw[2], bw[1]);
  const F2 h_lo = sel2(a1, b1, bh[1], bh[0]), h_hi = sel2(a1, b1, bh[2], bh[1]);
  const F2 s0 = sel2(a1, b1, bs[1], bs[0]), s1 = sel2(a1, b1, bs[2], bs[1]);
  // ---- the bin's rational-quadratic segment
  const F2 w = sub2(w_hi, w_lo), h = sub2(h_hi, h_lo);
  const F2 rw = map2(w, [](float v) { return rcp_fast(v); });
  const F2 delta = mul2(h, rw);
  auto softplus2 = [&](F2 s) {
    // as softplus_fast: max(lg2(1 + 2^min(s beta log2e, 64)) ln2 / beta, s)
    const F2 t = map2(mul2(s, bc2(c.beta_l2e)), [](float v) { return ex2_fast(fminf(v, 64.f)); });
    const F2 v = mul2(map2(add2(bc2(1.f), t), [](float u) { return lg2_fast(u); }), bc2(c.ln2_over_beta));
    return f2(fmaxf(lo(v), lo(s)), fmaxf(hi(v), hi(s)));
  };
  const F2 d0 = add2(bc2(c.min_d), softplus2(s0));
  const F2 d1 = add2(bc2(c.min_d), softplus2(s1));
  const F2 s = sub2(add2(d0, d1), mul2(bc2(2.f), delta));
  F2 th;
  if (ROOT) {
    const F2 q = sub2(x, h_lo);
    const F2 qs2 = mul2(q, s);
    const F2 a = fma2(h, sub2(delta, d0), qs2);
    const F2 b = fma2(h, d0, neg2(qs2));
    const F2 cc = mul2(neg2(delta), q);
    const F2 disc0 = fma2(b, b, mul2(mul2(bc2(-4.f), a), cc));
    const F2 sq = map2(disc0, [](float v) { return sqrt_fast(fmaxf(v, 0.f)); });
    const F2 den = map2(sub2(neg2(b), sq), [](float v) { return rcp_fast(v); });
    th = mul2(mul2(bc2(2.f), cc), den);
    y = fma2(th, w, w_lo);
  } else {
    th = mul2(sub2(x, w_lo), rw);
  }
  const F2 omt = sub2(bc2(1.f), th);
  const F2 t1 = mul2(th, omt);
  const F2 den = fma2(s, t1, delta);
  const F2 rden = map2(den, [](float v) { return rcp_fast(v); });
  if (!ROOT) y = fma2(mul2(h, fma2(mul2(delta, th), th, mul2(d0, t1))), rden, h_lo);
  const F2 num = mul2(mul2(delta, delta), fma2(mul2(d1, th), th, fma2(mul2(bc2(2.f), delta), t1, mul2(mul2(d0, omt), omt))));
  const F2 l = mul2(bc2(LN2), map2(mul2(mul2(num, rden), rden), [](float v) { return lg2_fast(v); }));
  lad = ROOT ? neg2(l) : l;
}

}  // namespace bgx
