// Generic fused coupling block, fp32 SIMT (any widths; the shape-general engine).
//
// One CTA owns a tile of TM samples and runs the whole block for them:
//   conditioner MLP(s) (register-tiled SGEMM, activations resident in shared memory,
//   weights streamed from L2 with cp.async) -> affine / rational-quadratic-spline transform
//   -> per-sample log|det J|, without writing any intermediate to global memory.
// Replaces bgflow/nn/flow/coupling.py:162-182 + transformer/affine.py:35-70 or
// transformer/spline.py:87-188 + dense.py:47-48 + periodic.py:30-37 (one kernel per block).
//
// The tensor-core (tcgen05) kernel in bgx_coupling_tc.cu covers the headline shapes; this
// kernel is the fallback for every other shape and the numerical cross-check of that one.
#include "bgx_coupling.cuh"

namespace bgx {

constexpr int TM = 64;    // samples per CTA
constexpr int NC = 128;   // output columns per pass
constexpr int KC = 16;    // k-chunk
constexpr int NT = 256;   // threads
constexpr int LDA = TM + 4;
constexpr int LDP = NC + 1;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

struct Smem {
  float* act[2];
  float* ws[2];
  float* a0[2];
  float* P;
  float* rowsum;
  float* rowaux;
};

// weight chunk [KC][NC] at (k0, n0) of Wt[Kp][Np]
__device__ __forceinline__ void load_w_chunk(float* dst, const float* __restrict__ Wt, int Np, int k0,
                                             int n0) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < (KC * NC / 4) / NT; ++i) {
    int e = tid + i * NT;          // float4 index
    int k = e / (NC / 4), c = (e % (NC / 4)) * 4;
    cp_async16(dst + k * NC + c, Wt + (long long)(k0 + k) * Np + n0 + c);
  }
}

__device__ __forceinline__ float load_input(const CouplingArgs& a, const DevMlp& net, long long row0, int m,
                                            int k) {
  return load_cond(a.cond, net, a.B, row0 + m, k);
}

// acc[4][8] (rows ty*4+i, cols n0 + {tx*4+j, 64+tx*4+j}) = A . Wt[:, n0:n0+128]
//   layer 0: A streamed from global through sm.a0 ; other layers: A = resident activations
__device__ __forceinline__ void gemm_pass(const CouplingArgs& a, const DevMlp& net, int layer, int n0,
                                          const float* Ares, const Smem& sm, long long row0,
                                          float (&acc)[4][8]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int Kp = net.Kp[layer], Np = net.Np[layer];
  const float* Wt = net.Wt[layer];
  const int nchunk = Kp / KC;
  const bool stream_a = (layer == 0);
  float areg[4];
  __syncthreads();  // previous users of ws / a0 / act are done
  load_w_chunk(sm.ws[0], Wt, Np, 0, n0);
  cp_async_commit();
  if (stream_a) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * NT, k = e % KC, m = e / KC;
      sm.a0[0][k * LDA + m] = load_input(a, net, row0, m, k);
    }
  }
  for (int c = 0; c < nchunk; ++c) {
    cp_async_wait_all();
    __syncthreads();
    if (c + 1 < nchunk) {
      load_w_chunk(sm.ws[(c + 1) & 1], Wt, Np, (c + 1) * KC, n0);
      cp_async_commit();
      if (stream_a) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int e = tid + i * NT, k = e % KC, m = e / KC;
          areg[i] = load_input(a, net, row0, m, (c + 1) * KC + k);
        }
      }
    }
    const float* W = sm.ws[c & 1];
    const float* A = stream_a ? sm.a0[c & 1] : Ares + (long long)c * KC * LDA;
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(A + kk * LDA + ty * 4);
      float4 b0 = *reinterpret_cast<const float4*>(W + kk * NC + tx * 4);
      float4 b1 = *reinterpret_cast<const float4*>(W + kk * NC + 64 + tx * 4);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    if (stream_a && c + 1 < nchunk) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int e = tid + i * NT, k = e % KC, m = e / KC;
        sm.a0[(c + 1) & 1][k * LDA + m] = areg[i];
      }
    }
  }
}

__device__ __forceinline__ int col_of(int n0, int tx, int j) {
  return n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
}

// run all hidden layers of `net`; returns the buffer index that holds the last hidden activations
__device__ int run_hidden(const CouplingArgs& a, const DevMlp& net, const Smem& sm, long long row0) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][8];
  int cur = 1;  // layer l writes act[l & 1]
  for (int l = 0; l < net.n_layers - 1; ++l) {
    const int out = l & 1;
    const float* Ares = sm.act[cur];
    for (int n0 = 0; n0 < net.Np[l]; n0 += NC) {
      gemm_pass(a, net, l, n0, Ares, sm, row0, acc);
      float* O = sm.act[out];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int col = col_of(n0, tx, j);
        float bj = __ldg(net.bias[l] + col);
        float4 v;
        v.x = act_apply(acc[0][j] + bj, net.act);
        v.y = act_apply(acc[1][j] + bj, net.act);
        v.z = act_apply(acc[2][j] + bj, net.act);
        v.w = act_apply(acc[3][j] + bj, net.act);
        if (col >= net.N[l]) v = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(O + col * LDA + ty * 4) = v;
      }
    }
    cur = out;
  }
  return cur;
}

// write the pass result (+bias) as P[m][c], c in [0,128)
__device__ __forceinline__ void stage_pass(const DevMlp& net, int layer, int n0, const Smem& sm,
                                           const float (&acc)[4][8]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int c = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
    float bj = __ldg(net.bias[layer] + n0 + c);
#pragma unroll
    for (int i = 0; i < 4; ++i) sm.P[(ty * 4 + i) * LDP + c] = acc[i][j] + bj;
  }
}

template <bool SPLINE, bool INVERSE>
__global__ void __launch_bounds__(NT, 2) coupling_simt_kernel(const CouplingArgs a) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem sm;
  {
    float* p = smem_raw;
    sm.act[0] = p; p += a.hb * LDA;
    sm.act[1] = p; p += a.hb * LDA;
    sm.ws[0] = p; p += KC * NC;
    sm.ws[1] = p; p += KC * NC;
    sm.a0[0] = p; p += KC * LDA;
    sm.a0[1] = p; p += KC * LDA;
    sm.P = nullptr;  // aliased onto the activation buffer that is free during the last layer
    sm.rowsum = p; p += TM;
    sm.rowaux = p; p += TM;
  }
  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * TM;
  if (tid < TM) {
    sm.rowsum[tid] = 0.f;
    sm.rowaux[tid] = 0.f;
  }
  float acc[4][8];

  if (SPLINE) {
    const DevMlp& net = a.net0;
    const int L = net.n_layers - 1;
    int cur = run_hidden(a, net, sm, row0);
    sm.P = sm.act[cur ^ 1];
    const int my_row = tid & (TM - 1);
    const long long grow = row0 + my_row;
    float ld_acc = 0.f;
    int npass = net.Np[L] / NC;
    for (int p = 0; p < npass; ++p) {
      gemm_pass(a, net, L, p * NC, sm.act[cur], sm, row0, acc);
      stage_pass(net, L, p * NC, sm, acc);
      __syncthreads();
      for (int it = tid; it < TM * a.dpp; it += NT) {
        int slot = it / TM;
        int d = p * a.dpp + slot;
        if (d < a.D_t && grow < a.B) {
          float x = __ldg(seg_addr(a.tin, grow, d));
          if (x < a.sp.left || x > a.sp.right) {
            if (a.sp.oob) atomicAdd(a.sp.oob, 1);
            x = fminf(fmaxf(x, a.sp.left), a.sp.right);
          }
          float y, lad;
          const float* prm = sm.P + my_row * LDP + slot * a.pstride;
          if (INVERSE) rqs_eval<false>(prm, 1, a.sp, x, y, lad);
          else rqs_eval<true>(prm, 1, a.sp, x, y, lad);
          *const_cast<float*>(seg_addr(a.tout, grow, d)) = y;
          ld_acc += lad;
        }
      }
    }
    atomicAdd(&sm.rowsum[my_row], ld_acc);
    __syncthreads();
  } else {
    const float sign = INVERSE ? -1.f : 1.f;
    const bool circular = a.flags & BGX_FLAG_CIRCULAR;
    const bool pv = (a.flags & BGX_FLAG_PRESERVE_VOLUME) && a.has1;
    // ---- shift net: mu -> tr_out (global scratch, re-read below by the same CTA)
    if (a.has0) {
      const DevMlp& net = a.net0;
      const int L = net.n_layers - 1;
      int cur = run_hidden(a, net, sm, row0);
      sm.P = sm.act[cur ^ 1];
      for (int n0 = 0; n0 < net.Np[L]; n0 += NC) {
        gemm_pass(a, net, L, n0, sm.act[cur], sm, row0, acc);
        stage_pass(net, L, n0, sm, acc);
        __syncthreads();
        for (int it = tid; it < TM * NC; it += NT) {
          int c = it % NC, m = it / NC;
          int col = n0 + c;
          long long grow = row0 + m;
          if (col < a.D_t && grow < a.B) {
            float mu = sm.P[m * LDP + c];
            float* o = const_cast<float*>(seg_addr(a.tout, grow, col));
            if (a.has1) {
              *o = mu;
            } else {
              float y = __ldg(seg_addr(a.tin, grow, col));
              float r = INVERSE ? (y - mu) : (y + mu);
              if (circular) r = r - floorf(r);
              *o = r;
            }
          }
        }
      }
    }
    if (a.has1) {
      const DevMlp& net = a.net1;
      const int L = net.n_layers - 1;
      int cur = run_hidden(a, net, sm, row0);
      sm.P = sm.act[cur ^ 1];
      float mean_ls = 0.f;
      for (int phase = pv ? 0 : 1; phase < 2; ++phase) {
        for (int n0 = 0; n0 < net.Np[L]; n0 += NC) {
          gemm_pass(a, net, L, n0, sm.act[cur], sm, row0, acc);
          stage_pass(net, L, n0, sm, acc);
          __syncthreads();
          for (int it = tid; it < TM * NC; it += NT) {
            int c = it % NC, m = it / NC;   // a warp = 32 consecutive columns of one row
            int col = n0 + c;
            long long grow = row0 + m;
            float ls = 0.f;
            bool ok = col < a.D_t && grow < a.B;
            if (ok) ls = tanhf(sm.P[m * LDP + c]) * a.alpha;
            if (phase == 0) {
              float s = warp_sum(ls);
              if ((tid & 31) == 0) atomicAdd(&sm.rowaux[m], s);
            } else {
              if (pv) mean_ls = sm.rowaux[m] / (float)a.D_t;
              ls -= mean_ls;
              if (ok) {
                float y = __ldg(seg_addr(a.tin, grow, col));
                float* o = const_cast<float*>(seg_addr(a.tout, grow, col));
                float mu = a.has0 ? *o : 0.f;
                float r = INVERSE ? expf(-ls) * (y - mu) : fmaf(expf(ls), y, mu);
                *o = r;
              } else {
                ls = 0.f;
              }
              float s = warp_sum(ls);
              if ((tid & 31) == 0) atomicAdd(&sm.rowsum[m], sign * s);
            }
          }
        }
        __syncthreads();
      }
    } else if (!a.has0) {
      // identity transformer (both nets absent): copy
      for (int it = tid; it < TM * a.D_t; it += NT) {
        int col = it % a.D_t, m = it / a.D_t;
        long long grow = row0 + m;
        if (grow < a.B) {
          float r = __ldg(seg_addr(a.tin, grow, col));
          if (circular) r = r - floorf(r);
          *const_cast<float*>(seg_addr(a.tout, grow, col)) = r;
        }
      }
    }
    __syncthreads();
  }
  if (tid < TM && row0 + tid < a.B) {
    float base = a.dlogp_in ? a.dlogp_in[row0 + tid] : 0.f;
    a.dlogp_out[row0 + tid] = base + sm.rowsum[tid];
  }
}

// ------------------------------------------------------------------------------ host side

void mlp_to_dev(const bgx_packed_mlp* p, DevMlp& d) {
  d.n_layers = p->n_layers;
  d.act = p->act;
  for (int i = 0; i < p->n_layers; ++i) {
    d.K[i] = p->K[i]; d.N[i] = p->N[i]; d.Kp[i] = p->Kp[i]; d.Np[i] = p->Np[i];
    d.Wt[i] = p->Wt[i]; d.bias[i] = p->bias[i];
  }
  d.in_map = p->in_map;
  d.pscale = p->periodic_scale;
  d.pleft = p->periodic_left;
}

int coupling_fill_io(const bgx_coupling_io* io, CouplingArgs& a, int& d_c, int& d_t) {
  if (!io || io->batch < 0 || io->n_cond < 0 || io->n_cond > BGX_MAX_SEGS || io->n_tr < 1 ||
      io->n_tr > BGX_MAX_SEGS || !io->dlogp_out)
    return BGX_ERR_INVALID;
  a.B = io->batch;
  d_c = d_t = 0;
  a.cond.n = io->n_cond; a.tin.n = io->n_tr; a.tout.n = io->n_tr;
  for (int i = 0; i < io->n_cond; ++i) {
    a.cond.ptr[i] = io->cond[i].ptr; a.cond.width[i] = io->cond[i].width; a.cond.stride[i] = io->cond[i].stride;
    d_c += io->cond[i].width;
  }
  for (int i = 0; i < io->n_tr; ++i) {
    if (io->tr_in[i].width != io->tr_out[i].width) return BGX_ERR_INVALID;
    a.tin.ptr[i] = io->tr_in[i].ptr; a.tin.width[i] = io->tr_in[i].width; a.tin.stride[i] = io->tr_in[i].stride;
    a.tout.ptr[i] = io->tr_out[i].ptr; a.tout.width[i] = io->tr_out[i].width; a.tout.stride[i] = io->tr_out[i].stride;
    d_t += io->tr_in[i].width;
  }
  a.D_t = d_t;
  a.dlogp_in = io->dlogp_in;
  a.dlogp_out = io->dlogp_out;
  return BGX_OK;
}

void spline_params_from_cfg(const bgx_spline_cfg* cfg, SplineParams& sp) {
  sp.K = cfg->n_bins;
  sp.left = cfg->left; sp.right = cfg->right; sp.bottom = cfg->bottom; sp.top = cfg->top;
  sp.min_w = cfg->min_bin_width; sp.min_h = cfg->min_bin_height; sp.min_d = cfg->min_derivative;
  sp.beta = cfg->identity_init ? (float)(0.6931471805599453 / (1.0 - (double)cfg->min_derivative)) : 1.f;
  sp.inv_beta = 1.f / sp.beta;
  sp.oob = cfg->oob_counter;
}

static int hidden_rows(const bgx_packed_mlp* p) {
  int hb = NC;
  for (int l = 0; l + 1 < p->n_layers; ++l) hb = max(hb, p->Np[l]);
  return hb;
}

static size_t smem_bytes(int hb) {
  static_assert(NC * LDA >= TM * LDP, "P must fit in one activation buffer");
  return sizeof(float) * (size_t)(2 * hb * LDA + 2 * KC * NC + 2 * KC * LDA + 2 * TM);
}

template <bool SPLINE, bool INVERSE>
static int launch(const CouplingArgs& a, cudaStream_t st) {
  if (a.B == 0) return BGX_OK;
  size_t sb = smem_bytes(a.hb);
  if (sb > 227 * 1024) return BGX_ERR_UNSUPPORTED;
  auto kern = coupling_simt_kernel<SPLINE, INVERSE>;
  static size_t configured_all[BGX_MAX_DEVICES] = {};
  size_t& configured = configured_all[device_slot()];
  if (sb > configured) {
    int rc = check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
    if (rc) return rc;
    configured = sb;
  }
  long long grid = (a.B + TM - 1) / TM;
  if (grid > 0x7fffffffLL) return BGX_ERR_UNSUPPORTED;
  kern<<<(unsigned)grid, NT, sb, st>>>(a);
  return post_launch();
}

int affine_coupling_simt(const bgx_coupling_io* io, const bgx_packed_mlp* shift, const bgx_packed_mlp* scale,
                         float log_alpha, int flags, cudaStream_t st) {
  CouplingArgs a{};
  int d_c, d_t;
  int rc = coupling_fill_io(io, a, d_c, d_t);
  if (rc) return rc;
  a.has0 = shift != nullptr;
  a.has1 = scale != nullptr;
  a.hb = NC;
  for (const bgx_packed_mlp* p : {shift, scale}) {
    if (!p) continue;
    if (p->n_layers < 1 || p->raw_width != d_c || p->N[p->n_layers - 1] != d_t) return BGX_ERR_INVALID;
    a.hb = max(a.hb, hidden_rows(p));
  }
  if ((flags & BGX_FLAG_CIRCULAR) && scale) return BGX_ERR_INVALID;  // affine.py:26-27
  if (shift) mlp_to_dev(shift, a.net0);
  if (scale) mlp_to_dev(scale, a.net1);
  a.alpha = expf(log_alpha);
  a.flags = flags;
  return (flags & BGX_FLAG_INVERSE) ? launch<false, true>(a, st) : launch<false, false>(a, st);
}

int spline_coupling_simt(const bgx_coupling_io* io, const bgx_packed_mlp* net, const bgx_spline_cfg* cfg,
                         int flags, cudaStream_t st) {
  CouplingArgs a{};
  int d_c, d_t;
  int rc = coupling_fill_io(io, a, d_c, d_t);
  if (rc) return rc;
  if (!net || !cfg || net->n_layers < 1 || net->raw_width != d_c) return BGX_ERR_INVALID;
  const int K = cfg->n_bins;
  if (K < 1 || net->spline_stride != 3 * K + 1 || net->spline_dims_per_pass < 1) return BGX_ERR_INVALID;
  if (net->N[net->n_layers - 1] != ceil_div(d_t, net->spline_dims_per_pass) * NC) return BGX_ERR_INVALID;
  if (cfg->min_bin_width * K > 1.f || cfg->min_bin_height * K > 1.f) return BGX_ERR_INVALID;
  mlp_to_dev(net, a.net0);
  a.has0 = 1;
  a.hb = hidden_rows(net);
  a.dpp = net->spline_dims_per_pass;
  a.pstride = net->spline_stride;
  spline_params_from_cfg(cfg, a.sp);
  a.flags = flags;
  return (flags & BGX_FLAG_INVERSE) ? launch<true, true>(a, st) : launch<true, false>(a, st);
}

}  // namespace bgx
