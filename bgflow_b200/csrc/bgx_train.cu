// Training path of a conditioner (bgflow/nn/dense.py:10-48 under torch autograd in the reference): the whole
// recompute and the whole backward of one DenseNet as ONE host call each.
//
// A KL training step of the 8-block Ala2 stack is ~30 GEMM-shaped kernels and ~25 elementwise ones per coupling
// block.  Issued from Python (one torch op / ctypes call each, 20-30 us of interpreter time apiece) the step is
// CPU-bound: 12.2 ms wall for 7.2 ms of GPU work (profiles/r2_train_profile_tcgen05_python.txt).  The drivers below
// issue the same work from C++ — bgx_linear (recompute, input gradients), bgx_gemm_tn (weight / bias gradients), and
// three small kernels of this file (activation, activation gradient, the fixed-order sum of bgx_gemm_tn's per-slice
// partials) — a few microseconds per launch, no allocation, nothing returned to the interpreter in between.
// (Folding the activation and its slope into bgx_linear's store path was measured and dropped: the precise expf / tanhf
// forms run on the critical path of that kernel's tile pipeline, +0.8 ms per step against 0.37 ms of HBM-bound
// elementwise kernels saved.)
//
// Layout: every activation / gradient buffer of layer i is [batch, pad4(dims[i + 1])] (row stride = a multiple of 4
// floats: 16-byte aligned rows for the vector paths of bgx_linear and bgx_spline_backward); pad columns are written
// as zeros by the kernels (packed weights and biases are zero beyond the true width).
#include <algorithm>

#include "bgx_common.cuh"

namespace bgx {
long long pack_linear_floats(int K, int N);
int pack_linear(const float* W, long long rs, long long cs, const float* bias, int K, int N, float* dst, bgx_packed_mlp* out,
                cudaStream_t st);
int linear_launch(int64_t batch, const float* x, int64_t ldx, const bgx_packed_mlp* net, float* y, int64_t ldy, int n_out,
                  int32_t* status, cudaStream_t stream);

__device__ __forceinline__ float act_value(float z, int act) {
  if (act == BGX_ACT_RELU) return fmaxf(z, 0.f);
  if (act == BGX_ACT_SILU) return z / (1.f + expf(-z));
  if (act == BGX_ACT_TANH) return tanhf(z);
  return z;
}
__device__ __forceinline__ float act_slope(float z, int act) {
  if (act == BGX_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == BGX_ACT_SILU) {
    const float s = 1.f / (1.f + expf(-z));
    return s * (1.f + z * (1.f - s));
  }
  if (act == BGX_ACT_TANH) {
    const float t = tanhf(z);
    return 1.f - t * t;
  }
  return 1.f;
}
// h = act(z)
__global__ void __launch_bounds__(256) act_forward_kernel(const float4* __restrict__ z, float4* __restrict__ h, long long n4,
                                                          int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = z[i];
  h[i] = make_float4(act_value(v.x, act), act_value(v.y, act), act_value(v.z, act), act_value(v.w, act));
}
// g *= act'(z)
__global__ void __launch_bounds__(256) act_backward_kernel(const float4* __restrict__ z, float4* __restrict__ g, long long n4,
                                                           int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = z[i];
  float4 o = g[i];
  o.x *= act_slope(v.x, act); o.y *= act_slope(v.y, act); o.z *= act_slope(v.z, act); o.w *= act_slope(v.w, act);
  g[i] = o;
}
// dW[n][c0 + c] = sum_s part_w[s][n][c] (c < kc), db[n] = sum_s part_b[s][n]: slices added in index order
__global__ void __launch_bounds__(256) sum_slices_kernel(const float* __restrict__ part_w, const float* __restrict__ part_b,
                                                         int slices, int rows, int n, int kc, float* __restrict__ dW,
                                                         long long ldw, float* __restrict__ db) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nw = (long long)n * 128;
  if (idx < nw) {
    const int row = (int)(idx >> 7), col = (int)(idx & 127);
    if (col < kc) {
      float s = 0.f;
      for (int i = 0; i < slices; ++i) s += part_w[((long long)i * rows + row) * 128 + col];
      dW[row * ldw + col] = s;
    }
  } else if (db && idx < nw + n) {
    const int row = (int)(idx - nw);
    float s = 0.f;
    for (int i = 0; i < slices; ++i) s += part_b[(long long)i * rows + row];
    db[row] = s;
  }
}

// Elementwise part of the RealNVP block's backward (affine.py:35-70 differentiated):  ls = tanh(s) exp(log_alpha),
//   forward  y' = y exp(ls) + mu,      dlogp = +sum ls:   dy = g e^{ls},   dmu = g,     dls = dy y + g_dl
//   inverse  y' = (y - mu) exp(-ls),   dlogp = -sum ls:   dy = g e^{-ls},  dmu = -dy,   dls = -dy (y - mu) - g_dl
//   ds = dls exp(log_alpha) (1 - tanh^2 s),   d log_alpha = sum dls ls  (per-block partials, added in index order)
// mu / s / dmu / ds are [B][wp] (wp = pad4(d_t); pad columns of the outputs are written as zeros), y / g / dy [B][d_t].
__global__ void __launch_bounds__(256) affine_backward_kernel(long long B, int d_t, int wp, const float* __restrict__ mu,
                                                              const float* __restrict__ sc, const float* __restrict__ y,
                                                              const float* __restrict__ g_out, const float* __restrict__ g_dl,
                                                              const float* __restrict__ log_alpha, int inverse,
                                                              float* __restrict__ d_y, float* __restrict__ d_mu,
                                                              float* __restrict__ d_s, float* __restrict__ part_alpha) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (idx < B * wp) {
    const long long row = idx / wp;
    const int col = (int)(idx - row * wp);
    float o_mu = 0.f, o_s = 0.f;
    if (col < d_t) {
      const float alpha = expf(__ldg(log_alpha));
      const float th = tanhf(sc[idx]);
      const float ls = th * alpha;
      const float g = g_out[row * d_t + col], yy = y[row * d_t + col];
      const float gd = g_dl ? g_dl[row] : 0.f;
      float dy, dls;
      if (!inverse) {
        dy = g * expf(ls);
        o_mu = g;
        dls = dy * yy + gd;
      } else {
        dy = g * expf(-ls);
        o_mu = -dy;
        dls = -dy * (yy - mu[idx]) - gd;
      }
      d_y[row * d_t + col] = dy;
      o_s = dls * alpha * (1.f - th * th);
      contrib = dls * ls;
    }
    d_mu[idx] = o_mu;
    d_s[idx] = o_s;
  }
  // block sum in a fixed order: lanes by shuffle tree, then the 8 warps in index order
  __shared__ float wsum[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += wsum[w];
    part_alpha[blockIdx.x] = t;
  }
}
// out[0] = sum of part[0..n) in index order (one block); dst[i] += src[i]
__global__ void __launch_bounds__(256) sum_vector_kernel(const float* __restrict__ part, long long n, float* __restrict__ out) {
  __shared__ float acc[256];
  float t = 0.f;
  for (long long i = threadIdx.x; i < n; i += 256) t += part[i];
  acc[threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 256; ++i) tot += acc[i];
    out[0] = tot;
  }
}
__global__ void __launch_bounds__(256) add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

static inline int pad4(int n) { return (n + 3) / 4 * 4; }
static inline int width_in(const bgx_train_mlp* net, int i) { return i == 0 ? net->dims[0] : pad4(net->dims[i]); }

static int check_net(const bgx_train_mlp* net) {
  if (!net || net->n_layers < 1 || net->n_layers > BGX_MAX_LAYERS) return BGX_ERR_INVALID;
  for (int i = 0; i <= net->n_layers; ++i)
    if (net->dims[i] < 1) return BGX_ERR_INVALID;
  return BGX_OK;
}

}  // namespace bgx

using namespace bgx;

extern "C" int bgx_gemm_tn_slices(int64_t batch, int n);
extern "C" int bgx_gemm_tn(int64_t batch, const float* g, int64_t ldg, int n, const float* h, int64_t ldh, int k,
                           int slices, float* part_w, float* part_b, int32_t* status, void* stream);

// Supported shapes: every layer has <= 128 inputs or <= 128 outputs (bgx_linear).
extern "C" int bgx_train_pack(const bgx_mlp* src, const int32_t* act, float* dst, int64_t dst_floats, bgx_train_mlp* out,
                              void* stream) {
  if (!src || !out || src->n_layers < 1 || src->n_layers > BGX_MAX_LAYERS || src->n_periodic > 0) return BGX_ERR_INVALID;
  const int L = src->n_layers;
  bgx_train_mlp tn{};
  tn.n_layers = L;
  long long off = 0;
  for (int i = 0; i <= L; ++i) {
    if (src->dims[i] < 1) return BGX_ERR_INVALID;
    tn.dims[i] = src->dims[i];
  }
  for (int i = 0; i < L; ++i) {
    tn.act[i] = (i + 1 < L) ? (act ? act[i] : src->act) : BGX_ACT_NONE;
    if (tn.dims[i] > 128 && pad4(tn.dims[i + 1]) > 128) return BGX_ERR_UNSUPPORTED;
    off += pack_linear_floats(tn.dims[i], tn.dims[i + 1]) + pack_linear_floats(tn.dims[i + 1], tn.dims[i]);
  }
  tn.total_floats = off;
  if (!dst) {
    *out = tn;
    return BGX_OK;
  }
  if (dst_floats < off) return BGX_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  off = 0;
  for (int i = 0; i < L; ++i) {
    if (!src->W[i] || !src->b[i]) return BGX_ERR_INVALID;
    const int k = tn.dims[i], n = tn.dims[i + 1];
    int rc = pack_linear(src->W[i], k, 1, src->b[i], k, n, dst + off, &tn.fwd[i], st);             // y = x W^T + b
    if (rc) return rc;
    off += pack_linear_floats(k, n);
    rc = pack_linear(src->W[i], 1, k, nullptr, n, k, dst + off, &tn.bwd[i], st);                   // dx = g W
    if (rc) return rc;
    off += pack_linear_floats(n, k);
  }
  *out = tn;
  return BGX_OK;
}

extern "C" int64_t bgx_mlp_train_part_floats(int64_t batch, const bgx_train_mlp* net) {
  if (check_net(net) || batch <= 0) return 0;
  int64_t need = 0;
  for (int i = 0; i < net->n_layers; ++i) {
    const int n = net->dims[i + 1];
    const int64_t rows = (int64_t)ceil_div(n, 128) * 128;
    need = std::max<int64_t>(need, (int64_t)bgx_gemm_tn_slices(batch, n) * rows * 129);
  }
  return need;
}

extern "C" int bgx_mlp_forward_train(int64_t batch, const bgx_train_mlp* net, const float* x, const bgx_train_buffers* buf,
                                     int32_t* status, void* stream) {
  int rc = check_net(net);
  if (rc) return rc;
  if (batch < 0 || !buf || (batch > 0 && !x)) return BGX_ERR_INVALID;
  if (batch == 0) return BGX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int L = net->n_layers;
  const float* hin = x;
  for (int i = 0; i < L; ++i) {
    const int wout = pad4(net->dims[i + 1]);
    if (!buf->z[i] || (i + 1 < L && !buf->h[i])) return BGX_ERR_INVALID;
    rc = linear_launch(batch, hin, width_in(net, i), &net->fwd[i], buf->z[i], wout, wout, status, st);
    if (rc) return rc;
    if (i + 1 < L) {
      const long long n4 = batch * (long long)wout / 4;
      act_forward_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(buf->z[i]),
                                                                      reinterpret_cast<float4*>(buf->h[i]), n4, net->act[i]);
      rc = post_launch();
      if (rc) return rc;
      hin = buf->h[i];
    }
  }
  return BGX_OK;
}

extern "C" int bgx_mlp_backward(int64_t batch, const bgx_train_mlp* net, const float* x, const bgx_train_buffers* buf,
                                const float* d_out, float* d_x, float* const* d_w, float* const* d_b, int32_t* status,
                                void* stream) {
  int rc = check_net(net);
  if (rc) return rc;
  if (batch <= 0 || !buf || !x || !d_out || !d_w || !d_b || !buf->part) return BGX_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int L = net->n_layers;
  const float* g = d_out;
  for (int i = L - 1; i >= 0; --i) {
    const int k = net->dims[i], n = net->dims[i + 1];
    const int ldg = pad4(n), ldh = width_in(net, i);
    const float* hin = i == 0 ? x : buf->h[i - 1];
    if (!d_w[i] || !d_b[i]) return BGX_ERR_INVALID;
    // ---- dW = g^T h, db = sum_b g: per-slice partials on the tensor cores, then their fixed-order sum
    const int slices = bgx_gemm_tn_slices(batch, n);
    const int rows = ceil_div(n, 128) * 128;
    float* part_w = buf->part;
    float* part_b = buf->part + (long long)slices * rows * 128;
    for (int c0 = 0; c0 < k; c0 += 128) {
      const int kc = std::min(128, k - c0);
      rc = bgx_gemm_tn(batch, g, ldg, n, hin + c0, ldh, kc, slices, part_w, c0 == 0 ? part_b : nullptr, status, stream);
      if (rc) return rc;
      const long long total = (long long)n * 128 + (c0 == 0 ? n : 0);
      sum_slices_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part_w, part_b, slices, rows, n, kc, d_w[i] + c0, k,
                                                                       c0 == 0 ? d_b[i] : nullptr);
      rc = post_launch();
      if (rc) return rc;
    }
    // ---- dh = g W (a layer with weight W^T), then through the activation of the layer below
    if (i > 0) {
      if (!buf->g[i - 1] || !buf->z[i - 1]) return BGX_ERR_INVALID;
      rc = linear_launch(batch, g, ldg, &net->bwd[i], buf->g[i - 1], ldh, ldh, status, st);
      if (rc) return rc;
      const long long n4 = batch * (long long)ldh / 4;
      act_backward_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(buf->z[i - 1]),
                                                                       reinterpret_cast<float4*>(buf->g[i - 1]), n4,
                                                                       net->act[i - 1]);
      rc = post_launch();
      if (rc) return rc;
      g = buf->g[i - 1];
    } else if (d_x) {
      rc = linear_launch(batch, g, ldg, &net->bwd[0], d_x, k, k, status, st);
      if (rc) return rc;
    }
  }
  return BGX_OK;
}

namespace bgx {
int spline_backward_launch(int64_t batch, int32_t d_t, const float* params, int64_t params_stride, const float* y,
                           const float* g_out, const float* g_dlogp, const int32_t* end_slope_col, const bgx_spline_cfg* cfg,
                           int flags, float* d_params, float* d_y, int zero_from, void* stream);
}

// The whole backward of one spline coupling block (coupling.py:161-180 + spline.py:87-188 + dense.py:47-48 under
// autograd in the reference) in ONE host call: conditioner recompute -> spline chain rule (dP, dy) -> conditioner
// backward (d cond, every dW, db).  `d_p` is scratch [batch, pad4(dims[n_layers])].
extern "C" int bgx_spline_coupling_backward(int64_t batch, const bgx_train_mlp* net, const float* cond, int32_t d_t,
                                            const float* y, const float* g_out, const float* g_dlogp,
                                            const int32_t* end_slope_col, const bgx_spline_cfg* cfg, int flags,
                                            const bgx_train_buffers* buf, float* d_p, float* d_cond, float* d_y,
                                            float* const* d_w, float* const* d_b, int32_t* status, void* stream) {
  int rc = check_net(net);
  if (rc) return rc;
  if (batch <= 0 || !buf || !cond || !y || !g_out || !d_p || !d_y || !cfg) return BGX_ERR_INVALID;
  const int L = net->n_layers, n_out = net->dims[L], ldp = pad4(n_out);
  rc = bgx_mlp_forward_train(batch, net, cond, buf, status, stream);
  if (rc) return rc;
  // pad columns of dP are operands of the backward GEMMs: the transform kernel writes them as zeros (a 2-D memset of
  // three columns x 65,536 rows took 38 us)
  rc = spline_backward_launch(batch, d_t, buf->z[L - 1], ldp, y, g_out, g_dlogp, end_slope_col, cfg, flags, d_p, d_y,
                              ldp > n_out ? n_out : -1, stream);
  if (rc) return rc;
  return bgx_mlp_backward(batch, net, cond, buf, d_p, d_cond, d_w, d_b, status, stream);
}

// scratch floats of bgx_affine_coupling_backward
extern "C" int64_t bgx_affine_backward_scratch_floats(int64_t batch, int32_t d_t, int32_t cond_width) {
  if (batch <= 0 || d_t <= 0 || cond_width <= 0) return 0;
  const int64_t wp = pad4(d_t);
  const int64_t blocks = (batch * wp + 255) / 256;
  return 2 * batch * wp + (batch * cond_width + 3) / 4 * 4 + (blocks + 3) / 4 * 4;
}

// The whole backward of one RealNVP coupling block (coupling.py:161-180 + affine.py:35-70 + dense.py:47-48 under
// autograd in the reference) in ONE host call: both conditioners recomputed, the elementwise chain rule, both
// conditioner backwards, d cond = their sum, d log_alpha.  Plain blocks only (shift and scale nets, no volume
// preservation, no circular wrap).
extern "C" int bgx_affine_coupling_backward(int64_t batch, const bgx_train_mlp* shift, const bgx_train_mlp* scale,
                                            const float* log_alpha, const float* cond, int32_t d_t, const float* y,
                                            const float* g_out, const float* g_dlogp, int flags,
                                            const bgx_train_buffers* buf_shift, const bgx_train_buffers* buf_scale,
                                            float* scratch, float* d_cond, float* d_y, float* const* d_w_shift,
                                            float* const* d_b_shift, float* const* d_w_scale, float* const* d_b_scale,
                                            float* d_log_alpha, int32_t* status, void* stream) {
  int rc = check_net(shift);
  if (rc) return rc;
  rc = check_net(scale);
  if (rc) return rc;
  if (batch <= 0 || !log_alpha || !cond || !y || !g_out || !buf_shift || !buf_scale || !scratch || !d_cond || !d_y ||
      !d_log_alpha)
    return BGX_ERR_INVALID;
  const int Ls = shift->n_layers, Lc = scale->n_layers;
  if (shift->dims[Ls] != d_t || scale->dims[Lc] != d_t || shift->dims[0] != scale->dims[0]) return BGX_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int wp = pad4(d_t), k0 = shift->dims[0];
  const long long nel = batch * (long long)wp;
  const long long blocks = (nel + 255) / 256;
  float* d_mu = scratch;
  float* d_s = d_mu + nel;
  float* dx_s = d_s + nel;
  float* part = dx_s + (batch * (long long)k0 + 3) / 4 * 4;
  rc = bgx_mlp_forward_train(batch, shift, cond, buf_shift, status, stream);
  if (rc) return rc;
  rc = bgx_mlp_forward_train(batch, scale, cond, buf_scale, status, stream);
  if (rc) return rc;
  affine_backward_kernel<<<(unsigned)blocks, 256, 0, st>>>(batch, d_t, wp, buf_shift->z[Ls - 1], buf_scale->z[Lc - 1], y, g_out,
                                                         g_dlogp, log_alpha, (flags & BGX_FLAG_INVERSE) ? 1 : 0, d_y, d_mu,
                                                         d_s, part);
  rc = post_launch();
  if (rc) return rc;
  sum_vector_kernel<<<1, 256, 0, st>>>(part, blocks, d_log_alpha);
  rc = post_launch();
  if (rc) return rc;
  rc = bgx_mlp_backward(batch, shift, cond, buf_shift, d_mu, d_cond, d_w_shift, d_b_shift, status, stream);
  if (rc) return rc;
  rc = bgx_mlp_backward(batch, scale, cond, buf_scale, d_s, dx_s, d_w_scale, d_b_scale, status, stream);
  if (rc) return rc;
  const long long nx = batch * (long long)k0;
  add_inplace_kernel<<<(unsigned)((nx + 255) / 256), 256, 0, st>>>(d_cond, dx_s, nx);
  return post_launch();
}
