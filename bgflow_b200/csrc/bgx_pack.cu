// Parameter re-layout for the kernels (no reference counterpart: the reference keeps
// nn.Linear weights [out,in] and lets ATen pick a GEMM; spline.py:113-125 then slices the
// conditioner output by column blocks).  Here the slicing is folded into the weight layout once.
#include <mutex>
#include <set>
#include <vector>

#include "bgx_common.cuh"

namespace bgx {

thread_local cudaError_t g_last_error = cudaSuccess;
long long g_launches = 0;

// dst[k][n] = W[row(n)][k]   (K-padded rows / unmapped columns are zero)
__global__ void pack_transpose_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                      const int* __restrict__ row_map, int K, int N, int Kp, int Np,
                                      float* __restrict__ Wt, float* __restrict__ bias) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)Kp * Np;
  if (idx < total) {
    int n = (int)(idx % Np), k = (int)(idx / Np);
    int r = row_map ? row_map[n] : (n < N ? n : -1);
    Wt[idx] = (r >= 0 && k < K) ? W[(long long)r * K + k] : 0.f;
  }
  if (idx < Np) {
    int n = (int)idx;
    int r = row_map ? row_map[n] : (n < N ? n : -1);
    bias[n] = r >= 0 ? b[r] : 0.f;
  }
}

// Tensor-core layout: W is split exactly into three bf16 terms, W = b1 + b2 + b3 (each the
// round-to-nearest bf16 of the remaining residual: 24 mantissa bits in total).  Per term, per
// 128-row chunk of W and per 64-wide k-tile one 16 KB tile [128 rows][64 k] bf16, K-major, rows
// of 128 B with the 16-B chunks XOR-swizzled by (row & 7) (the UMMA SWIZZLE_128B canonical
// layout), so that ONE 1-D bulk copy lands a ready B-operand tile in shared memory.
__device__ __forceinline__ unsigned short f32_to_bf16_rn(float f) {
  unsigned int u = __float_as_uint(f);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (unsigned short)(u >> 16);
}
__global__ void pack_tc_tiles_kernel(const float* __restrict__ W, const int* __restrict__ row_map, int K,
                                     int N, int ktiles, int Np, unsigned short* __restrict__ t1,
                                     unsigned short* __restrict__ t2, unsigned short* __restrict__ t3) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)Np * ktiles * 64;
  if (idx >= total) return;
  int k = (int)(idx % (ktiles * 64)), n = (int)(idx / (ktiles * 64));
  int r = row_map ? row_map[n] : (n < N ? n : -1);
  float w = (r >= 0 && k < K) ? W[(long long)r * K + k] : 0.f;
  unsigned short b1 = f32_to_bf16_rn(w);
  float r1 = w - __uint_as_float((unsigned int)b1 << 16);
  unsigned short b2 = f32_to_bf16_rn(r1);
  float r2 = r1 - __uint_as_float((unsigned int)b2 << 16);
  unsigned short b3 = f32_to_bf16_rn(r2);
  int chunk = n >> 7, nr = n & 127, t = k >> 6, kk = k & 63;
  long long off = ((long long)chunk * ktiles + t) * 8192 + (nr * 64 + ((((kk >> 3) ^ (nr & 7)) << 3) | (kk & 7)));
  t1[off] = b1;
  t2[off] = b2;
  t3[off] = b3;
}

// The same tiles (two terms) from a strided source: element (n, k) of the layer's weight is W[n * rs + k * cs], so
// that the TRANSPOSED layer (dx = g W: a layer with weight W^T) is packed straight from nn.Linear.weight
// (rs = 1, cs = K of the source).  Rows n >= N and inputs k >= K are zero.
__global__ void pack_tc_tiles_strided_kernel(const float* __restrict__ W, long long rs, long long cs, int K, int N,
                                             int ktiles, int Np, unsigned short* __restrict__ t1,
                                             unsigned short* __restrict__ t2) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)Np * ktiles * 64;
  if (idx >= total) return;
  int k = (int)(idx % (ktiles * 64)), n = (int)(idx / (ktiles * 64));
  float w = (n < N && k < K) ? W[n * rs + k * cs] : 0.f;
  unsigned short b1 = f32_to_bf16_rn(w);
  float r1 = w - __uint_as_float((unsigned int)b1 << 16);
  unsigned short b2 = f32_to_bf16_rn(r1);
  int chunk = n >> 7, nr = n & 127, t = k >> 6, kk = k & 63;
  long long off = ((long long)chunk * ktiles + t) * 8192 + (nr * 64 + ((((kk >> 3) ^ (nr & 7)) << 3) | (kk & 7)));
  t1[off] = b1;
  t2[off] = b2;
}
__global__ void pack_bias_kernel(const float* __restrict__ b, int N, int Np, float* __restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < Np) out[n] = (b && n < N) ? b[n] : 0.f;
}

// One plain linear layer for the tensor-core kernels only (bgx_linear / the training drivers): two bf16 terms + bias.
long long pack_linear_floats(int K, int N) {
  const long long kp64 = round_up(K, 64), Np = round_up(N, 128);
  return ((kp64 * Np + Np) + 255) / 256 * 256;      // 2 terms x kp64 x Np bf16 = kp64 x Np floats, then the bias
}
int pack_linear(const float* W, long long rs, long long cs, const float* bias, int K, int N, float* dst, bgx_packed_mlp* out,
                cudaStream_t st) {
  if (!W || !dst || K < 1 || N < 1 || ((uintptr_t)dst & 255) != 0) return BGX_ERR_INVALID;
  bgx_packed_mlp pk{};
  pk.n_layers = 1;
  pk.act = BGX_ACT_NONE;
  pk.K[0] = K; pk.N[0] = N;
  pk.Kp[0] = round_up(K, 16); pk.Np[0] = round_up(N, 128);
  pk.raw_width = K;
  const int kp64 = round_up(K, 64), Np = pk.Np[0];
  unsigned short* t1 = reinterpret_cast<unsigned short*>(dst);
  unsigned short* t2 = t1 + (long long)kp64 * Np;
  float* bb = dst + (long long)kp64 * Np;
  const long long total = (long long)kp64 * Np;
  pack_tc_tiles_strided_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(W, rs, cs, K, N, kp64 / 64, Np, t1, t2);
  int rc = post_launch();
  if (rc) return rc;
  pack_bias_kernel<<<(Np + 255) / 256, 256, 0, st>>>(bias, N, Np, bb);
  rc = post_launch();
  if (rc) return rc;
  pk.Wb[0][0] = t1;
  pk.Wb[1][0] = t2;
  pk.bias[0] = bb;
  pk.total_floats = pack_linear_floats(K, N);
  *out = pk;
  return BGX_OK;
}

// last-layer bias of a spline net once more, 16-byte aligned per dim: [pass][dim][pad]
__global__ void pack_bias_pad_kernel(const float* __restrict__ bias, int npass, int dpp, int ps, int pad,
                                     float* __restrict__ out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npass * dpp * pad) return;
  int c = idx / (dpp * pad), r = idx - c * (dpp * pad), d = r / pad, k = r - d * pad;
  out[idx] = (k < ps) ? bias[c * 128 + d * ps + k] : 0.f;
}

static int spline_dims_per_pass(int n_bins) { return 128 / (3 * n_bins + 1); }

}  // namespace bgx

using namespace bgx;

// index maps interned by content: the returned pointer is valid for the life of the process (std::set nodes never move)
static const int* stable_host(const std::vector<int>& v) {
  static std::mutex mu;
  static std::set<std::vector<int>> pool;
  std::lock_guard<std::mutex> lock(mu);
  return pool.insert(v).first->data();
}

extern "C" int bgx_pack_mlp(const bgx_mlp* src, const bgx_spline_layout* spline, float* dst, int64_t dst_floats,
                            bgx_packed_mlp* out, void* stream) {
  if (!src || !out || src->n_layers < 1 || src->n_layers > BGX_MAX_LAYERS) return BGX_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int L = src->n_layers;
  bgx_packed_mlp pk{};
  pk.n_layers = L;
  pk.act = src->act;
  const bool is_spline = spline && spline->d_t > 0;
  int n_periodic = src->n_periodic > 0 ? src->n_periodic : 0;
  pk.raw_width = n_periodic ? src->raw_width : src->dims[0];
  if (src->dims[0] != pk.raw_width + n_periodic) return BGX_ERR_INVALID;
  if (n_periodic && !src->periodic_idx) return BGX_ERR_INVALID;
  for (int l = 0; l < L; ++l) {
    if (src->dims[l] < 1 || src->dims[l + 1] < 1) return BGX_ERR_INVALID;
    pk.K[l] = src->dims[l];
    pk.N[l] = src->dims[l + 1];
    pk.Kp[l] = round_up(pk.K[l], 16);
    pk.Np[l] = round_up(pk.N[l], 128);
  }
  std::vector<int> last_map;  // packed column -> source row of the last Linear (or -1)
  if (is_spline) {
    const int K = spline->n_bins, dt = spline->d_t;
    if (K < 1 || 3 * K + 1 > 128) return BGX_ERR_UNSUPPORTED;
    int n_nc = 0;
    std::vector<int> nc_rank(dt, -1);
    for (int d = 0; d < dt; ++d) {
      bool circ = spline->is_circular && spline->is_circular[d];
      if (!circ) nc_rank[d] = n_nc++;
    }
    // spline.py:112-117: the conditioner width must be 3*K*d_t + n_noncircular
    if (src->dims[L] != 3 * K * dt + n_nc) return BGX_ERR_INVALID;
    const int dpp = spline_dims_per_pass(K), ps = 3 * K + 1;
    pk.spline_dims_per_pass = dpp;
    pk.spline_stride = ps;
    pk.N[L - 1] = ceil_div(dt, dpp) * 128;  // packed width
    pk.Np[L - 1] = pk.N[L - 1];
    last_map.assign(pk.Np[L - 1], -1);
    for (int n = 0; n < pk.Np[L - 1]; ++n) {
      int p = n / 128, c = n % 128, slot = c / ps, j = c % ps, d = p * dpp + slot;
      if (slot >= dpp || d >= dt) continue;
      int r;
      if (j < K) r = d * K + j;
      else if (j < 2 * K) r = K * dt + d * K + (j - K);
      else if (j < 3 * K) r = 2 * K * dt + d * K + (j - 2 * K);
      else r = nc_rank[d] >= 0 ? 3 * K * dt + nc_rank[d] : 2 * K * dt + d * K;  // periodic: knot K := knot 0
      last_map[n] = r;
    }
  }
  // layout: [Wt_l | bias_l]* , [Wb1_l | Wb2_l | Wb3_l]* (16 KB swizzled bf16 tiles), in_map, last_map
  int64_t off = 0;
  int64_t o_wt[BGX_MAX_LAYERS], o_b[BGX_MAX_LAYERS], o_t[3][BGX_MAX_LAYERS];
  int kp64[BGX_MAX_LAYERS];
  for (int l = 0; l < L; ++l) {
    o_wt[l] = off; off += (int64_t)pk.Kp[l] * pk.Np[l];
    o_b[l] = off; off += pk.Np[l];
  }
  off = (off + 255) / 256 * 256;  // tiles are bulk-copied: keep them 1 KB aligned
  for (int l = 0; l < L; ++l) {
    kp64[l] = round_up(pk.K[l], 64);
    for (int term = 0; term < 3; ++term) {      // bf16: two elements per float slot
      o_t[term][l] = off;
      off += (int64_t)kp64[l] * pk.Np[l] / 2;
    }
  }
  int64_t o_bpad = off;
  if (is_spline) {
    pk.spline_bias_pad = round_up(pk.spline_stride, 4);
    off += (int64_t)(pk.Np[L - 1] / 128) * pk.spline_dims_per_pass * pk.spline_bias_pad;
    off = (off + 3) / 4 * 4;
  }
  int64_t o_inmap = off; off += round_up(pk.Kp[0], 4);
  int64_t o_lastmap = off; off += (int64_t)last_map.size();
  off = (off + 3) / 4 * 4;
  pk.total_floats = off;
  if (!dst) {
    *out = pk;
    return BGX_OK;
  }
  if (dst_floats < off) return BGX_ERR_WORKSPACE;
  if (((uintptr_t)dst & 255) != 0) return BGX_ERR_INVALID;  // torch allocations are 512-B aligned

  // conditioner input map (periodic.py:30-37): cos block, sin block, then the other columns
  std::vector<int> in_map(pk.Kp[0], 0);
  {
    std::vector<char> is_p(pk.raw_width, 0);
    for (int i = 0; i < n_periodic; ++i) {
      int c = src->periodic_idx[i];
      if (c < 0 || c >= pk.raw_width) return BGX_ERR_INVALID;
      is_p[c] = 1;
      in_map[i] = c | (1 << 24);
      in_map[n_periodic + i] = c | (2 << 24);
    }
    int k = 2 * n_periodic;
    for (int c = 0; c < pk.raw_width; ++c)
      if (!is_p[c]) {
        if (k >= pk.K[0]) return BGX_ERR_INVALID;
        in_map[k++] = c;
      }
    if (k != pk.K[0]) return BGX_ERR_INVALID;
  }
  int* d_inmap = reinterpret_cast<int*>(dst + o_inmap);
  int* d_lastmap = last_map.empty() ? nullptr : reinterpret_cast<int*>(dst + o_lastmap);
  // The host side of these copies must outlive the call: inside a CUDA-graph capture (a whole training step,
  // distributed.GraphedStep) the copy becomes a graph node that reads its HOST source again at every replay.
  // stable_host() interns the map by content (a model has a handful of distinct layouts), so the pointer stays valid.
  const int* h_inmap = stable_host(in_map);
  int rc = check(cudaMemcpyAsync(d_inmap, h_inmap, in_map.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  if (rc) return rc;
  if (d_lastmap) {
    const int* h_lastmap = stable_host(last_map);
    rc = check(cudaMemcpyAsync(d_lastmap, h_lastmap, last_map.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    if (rc) return rc;
  }
  pk.in_map = d_inmap;
  if (n_periodic) {
    pk.periodic_left = src->periodic_left;
    pk.periodic_scale = (float)(6.283185307179586 / ((double)src->periodic_right - (double)src->periodic_left));
  }
  for (int l = 0; l < L; ++l) {
    if (!src->W[l] || !src->b[l]) return BGX_ERR_INVALID;
    const int* rmap = (l == L - 1) ? d_lastmap : nullptr;
    const int n_true = (l == L - 1 && is_spline) ? pk.Np[l] : pk.N[l];
    float* wt = dst + o_wt[l];
    float* bb = dst + o_b[l];
    long long total = (long long)pk.Kp[l] * pk.Np[l];
    pack_transpose_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        src->W[l], src->b[l], rmap, pk.K[l], n_true, pk.Kp[l], pk.Np[l], wt, bb);
    rc = post_launch();
    if (rc) return rc;
    long long total2 = (long long)kp64[l] * pk.Np[l];
    pack_tc_tiles_kernel<<<(unsigned)((total2 + 255) / 256), 256, 0, st>>>(
        src->W[l], rmap, pk.K[l], n_true, kp64[l] / 64, pk.Np[l], (unsigned short*)(dst + o_t[0][l]),
        (unsigned short*)(dst + o_t[1][l]), (unsigned short*)(dst + o_t[2][l]));
    rc = post_launch();
    if (rc) return rc;
    pk.Wt[l] = wt;
    pk.bias[l] = bb;
    for (int term = 0; term < 3; ++term) pk.Wb[term][l] = dst + o_t[term][l];
    if (l == L - 1 && is_spline) {
      const int npass = pk.Np[l] / 128, n = npass * pk.spline_dims_per_pass * pk.spline_bias_pad;
      pack_bias_pad_kernel<<<(n + 255) / 256, 256, 0, st>>>(bb, npass, pk.spline_dims_per_pass, pk.spline_stride,
                                                          pk.spline_bias_pad, dst + o_bpad);
      rc = post_launch();
      if (rc) return rc;
      pk.spline_bias = dst + o_bpad;
    }
  }
  *out = pk;
  return BGX_OK;
}

extern "C" const char* bgx_version(void) { return "bgflow_b200 0.1 (sm_100a)"; }
extern "C" const char* bgx_last_cuda_error(void) { return cudaGetErrorString(g_last_error); }
extern "C" int64_t bgx_launch_count(void) { return g_launches; }
