/*
 * bgflow_b200 — C ABI of the B200-native coupling-flow engine.
 *
 * Drop-in boundary for the hot path of noegroup/bgflow (reference paths are relative to the
 * reference checkout):
 *
 *   bgx_affine_coupling   replaces  CouplingFlow._forward/_inverse (bgflow/nn/flow/coupling.py:162-182)
 *                                   + AffineTransformer (bgflow/nn/flow/transformer/affine.py:35-70)
 *                                   + DenseNet.forward x2 (bgflow/nn/dense.py:47-48)
 *   bgx_spline_coupling   replaces  CouplingFlow (coupling.py:162-182)
 *                                   + ConditionalSplineTransformer._compute_params/_forward/_inverse
 *                                     (bgflow/nn/flow/transformer/spline.py:87-188)
 *                                   + nflows.transforms.splines.rational_quadratic_spline (3rd party)
 *                                   + WrapPeriodic.forward (bgflow/nn/periodic.py:30-37)
 *   bgx_ic_to_xyz         replaces  GlobalInternalCoordinateTransformation._inverse
 *                                   (bgflow/nn/flow/crd_transform/ic.py:678-716, 209-265, 435-513,
 *                                    ic_helper.py:372-452, 480-575)
 *   bgx_ic_from_xyz       replaces  GlobalInternalCoordinateTransformation._forward
 *                                   (ic.py:633-676, 162-206, 386-433, ic_helper.py:148-293, 578-680)
 *   bgx_cdf_map           replaces  CDFTransform._forward/_inverse (bgflow/nn/flow/cdf.py:29-46) over
 *                                   TruncatedNormalDistribution / Normal / Uniform marginals
 *   bgx_ic_to_xyz_mapped  replaces  the builder tail: icdf maps + InverseFlow(GlobalIC)
 *   bgx_ic_from_xyz_mapped          (factory/generator_builder.py:408-459) in one kernel per direction
 *   bgx_relic_to_xyz      replaces  RelativeInternalCoordinateTransformation._inverse (ic.py:435-513) and
 *                                   MixedCoordinateTransformation._inverse (ic.py:862-884, pca.py:83-107)
 *   bgx_relic_from_xyz    replaces  their _forward (ic.py:386-433, 836-860)
 *   bgx_split_merge       replaces  SplitFlow._forward / MergeFlow by sizes (coupling.py:46-62, 107-110)
 *   bgx_pack_mlp          (no reference counterpart) re-lays nn.Linear weights for the kernels
 *
 * Conventions: plain pointers and sizes only, no torch types.  All device pointers are fp32,
 * row-major.  Every function returns BGX_OK (0) or a negative error code, never throws, never
 * allocates device memory (callers pass outputs and workspaces) and never synchronises the
 * host with the device.
 * `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 * dlogp follows bgflow: log|det J| of the direction evaluated, one float per sample.
 */
#ifndef BGFLOW_B200_H
#define BGFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGX_OK 0
#define BGX_ERR_INVALID (-1)      /* bad argument / inconsistent shapes */
#define BGX_ERR_UNSUPPORTED (-2)  /* shape outside what the kernels support */
#define BGX_ERR_WORKSPACE (-3)    /* workspace too small */
#define BGX_ERR_CUDA (-4)         /* a CUDA runtime call failed: see bgx_last_cuda_error */

#define BGX_MAX_LAYERS 8
#define BGX_MAX_SEGS 8

#define BGX_ACT_NONE 0
#define BGX_ACT_RELU 1
#define BGX_ACT_SILU 2
#define BGX_ACT_TANH 3

/* One DenseNet (bgflow/nn/dense.py:10-45): W[i] is nn.Linear.weight [dims[i+1], dims[i]]
 * row-major, b[i] is [dims[i+1]]; `act` follows every layer but the last.  If n_periodic > 0 the
 * net is wrapped in WrapPeriodic (periodic.py:30-37): the raw conditioner vector of width
 * `raw_width` is mapped to [cos(2pi(x_c-l)/(r-l)) for c in periodic_idx,
 * sin(...) likewise, x_others in original order], and dims[0] == raw_width + n_periodic. */
typedef struct bgx_mlp {
  int32_t n_layers;
  int32_t act;
  int32_t dims[BGX_MAX_LAYERS + 1];
  const float* W[BGX_MAX_LAYERS];    /* device */
  const float* b[BGX_MAX_LAYERS];    /* device */
  int32_t raw_width;
  int32_t n_periodic;
  const int32_t* periodic_idx;       /* HOST array, n_periodic entries (may be NULL if 0) */
  float periodic_left, periodic_right;
} bgx_mlp;

/* Re-laid parameters for the kernels.  Opaque to callers except `total_floats`. */
typedef struct bgx_packed_mlp {
  int32_t n_layers, act;
  int32_t K[BGX_MAX_LAYERS], N[BGX_MAX_LAYERS];    /* true in / out widths */
  int32_t Kp[BGX_MAX_LAYERS], Np[BGX_MAX_LAYERS];  /* padded */
  const float* Wt[BGX_MAX_LAYERS];                 /* device, [Kp][Np] (SIMT layout) */
  const float* bias[BGX_MAX_LAYERS];               /* device, [Np] */
  const int32_t* in_map;                           /* device, Kp[0] entries: raw col | kind<<24 */
  float periodic_scale, periodic_left;             /* arg = (x - left) * scale */
  int32_t raw_width;
  int32_t spline_dims_per_pass, spline_stride;     /* last-layer column layout (spline nets) */
  int64_t total_floats;
  /* tensor-core layout (tcgen05 path): W = b1 + b2 + b3 exactly, three bf16 terms; per term
   * 16 KB SWIZZLE_128B tiles [Np/128][ceil(K/64)][128 x 64] bf16 (device pointers) */
  const void* Wb[3][BGX_MAX_LAYERS];
  /* spline nets: the last layer's bias once more as [Np/128][dims_per_pass][spline_bias_pad] with
   * spline_bias_pad = round_up(spline_stride, 4) (zero padded: 16-byte vector loads per dim) */
  const float* spline_bias;
  int32_t spline_bias_pad, reserved_;
} bgx_packed_mlp;

/* Last-layer re-layout request for a spline conditioner: the raw column layout of
 * spline.py:113-125 ([W: d*K+b | H | S | S_last over non-circular dims]) becomes dim-major
 * [W(K) H(K) S(K) S_end(1)] per transformed dim; circular dims get S_end := S[d][0]. */
typedef struct bgx_spline_layout {
  int32_t d_t;                 /* transformed width (0 = not a spline net) */
  int32_t n_bins;
  const uint8_t* is_circular;  /* HOST, d_t entries, NULL = none circular */
} bgx_spline_layout;

/* Pack `src` into `dst` (device, `dst_floats` floats).  With dst == NULL only the required
 * size is computed (out->total_floats).  Launches small gather kernels on `stream` and does
 * one small async host->device copy of index maps taken from `dst`'s tail. */
int bgx_pack_mlp(const bgx_mlp* src, const bgx_spline_layout* spline, float* dst,
                 int64_t dst_floats, bgx_packed_mlp* out, void* stream);

/* A column slice of the flow state: tensor `ptr` with `width` columns, row stride `stride`
 * floats.  CouplingFlow concatenates several along the last dim (coupling.py:163-165). */
typedef struct bgx_seg {
  const float* ptr;
  int32_t width;
  int32_t stride;
} bgx_seg;

typedef struct bgx_coupling_io {
  int64_t batch;
  int32_t n_cond;
  bgx_seg cond[BGX_MAX_SEGS];     /* conditioner inputs, concatenated in order */
  int32_t n_tr;
  bgx_seg tr_in[BGX_MAX_SEGS];    /* transformed inputs */
  bgx_seg tr_out[BGX_MAX_SEGS];   /* outputs, same widths (ptr is written) */
  const float* dlogp_in;          /* [batch] running log-det to add to, or NULL */
  float* dlogp_out;               /* [batch] */
} bgx_coupling_io;

#define BGX_FLAG_INVERSE 1          /* evaluate bgflow's _inverse direction */
#define BGX_FLAG_PRESERVE_VOLUME 2  /* affine.py:44-45 */
#define BGX_FLAG_CIRCULAR 4         /* affine.py:56-57: y %= 1 (shift-only) */
#define BGX_FLAG_BF16X6 8           /* tensor-core path: 6 bf16 products (fp32-equivalent); default 3 (~2^-16) */
#define BGX_FLAG_FORCE_SIMT 16      /* always use the generic fp32 SIMT kernel */
#define BGX_FLAG_NO_PAIR 32         /* spline: never the pair kernel */
#define BGX_FLAG_FORCE_WIDE 64      /* spline pair kernel: in-place global tile access even where the tile fits shared memory */
#define BGX_FLAG_PREFER_PAIR 128    /* affine: the pair kernel even where the two-CTAs-per-SM kernel applies (default: the
                                       latter for narrow dense blocks, the pair kernel for wide ones).  Spline blocks use
                                       the pair kernel wherever it is eligible unless BGX_FLAG_NO_PAIR is set. */

/* y' = y * exp(ls) + mu   (forward)   |   y' = (y - mu) * exp(-ls)   (inverse)
 * mu = shift(cond), ls = tanh(scale(cond)) * exp(log_alpha);  dlogp = +-sum(ls).
 * `shift` / `scale` may be NULL (zeros). */
int bgx_affine_coupling(const bgx_coupling_io* io, const bgx_packed_mlp* shift,
                        const bgx_packed_mlp* scale, float log_alpha, int flags, void* stream);

typedef struct bgx_spline_cfg {
  int32_t n_bins;
  float left, right, bottom, top;
  float min_bin_width, min_bin_height, min_derivative;
  int32_t identity_init;   /* softplus beta = ln2/(1-min_derivative) (nflows PR #65) */
  int32_t* oob_counter;    /* device int, incremented per out-of-domain input (or NULL) */
  int32_t* status;         /* device int, set to 1 if the kernel's internal pipeline timed out
                              (results invalid; never happens in a correct build) (or NULL) */
} bgx_spline_cfg;

/* bgflow forward  = nflows inverse=True  (quadratic-root branch, the sampling direction);
 * bgflow inverse  = nflows inverse=False (direct rational-quadratic evaluation).
 * Inputs outside [left, right] are clamped (== spline.py:145-155) and counted. */
int bgx_spline_coupling(const bgx_coupling_io* io, const bgx_packed_mlp* params_net,
                        const bgx_spline_cfg* cfg, int flags, void* stream);

/* Backward of the spline TRANSFORM (training): the chain rule of spline.py:128-188 + nflows
 * rational_quadratic_spline written out by hand, one thread per (sample, dim).
 *   params        conditioner output [batch, >= 3*K*d_t + n_noncircular] in the reference's column
 *                 layout (spline.py:113-125), row stride `params_stride` floats
 *   y, g_out      [batch, d_t] dense: transformed input, gradient wrt the transformed output
 *   g_dlogp       [batch] gradient wrt dlogp (or NULL = 0)
 *   end_slope_col device, d_t entries: column of `params` that holds dim d's LAST slope — its own
 *                 first slope column for circular dims, its S_last column otherwise (spline.py:123-125)
 *   d_params      [batch, params_stride] gradient wrt `params` (every column the transform reads)
 *   d_y           [batch, d_t] gradient wrt y (0 where y was clamped)
 * flags: BGX_FLAG_INVERSE selects the direct branch (bgflow inverse).  The conditioner's own
 * backward (dense GEMMs) stays with the caller (torch autograd in bgflow_b200/autograd.py). */
int bgx_spline_backward(int64_t batch, int32_t d_t, const float* params, int64_t params_stride,
                        const float* y, const float* g_out, const float* g_dlogp,
                        const int32_t* end_slope_col, const bgx_spline_cfg* cfg, int flags,
                        float* d_params, float* d_y, void* stream);

/* ---- internal coordinates ------------------------------------------------------------- */

/* Device-side plan of a global z-matrix (ic.py:25-97 staging done on the host). */
typedef struct bgx_zplan {
  int32_t n_atoms;
  int32_t seeds[3];
  int32_t n_rel;               /* n_atoms - 3 */
  const int32_t* rel;          /* device, [n_rel][4] rows (i,j,k,l) in tensor column order */
  const int32_t* order;        /* device, [n_rel] placement order (row ids) */
  const int32_t* slot_of_col;  /* device, [3*n_atoms-6]: for the columns of bonds | angles | torsions the
                                  shared-memory slot 3*atom + {0 bond, 1 angle, 2 torsion} (or NULL) */
  int32_t normalize_angles;
  float eps;
} bgx_zplan;

/* x0: [batch,3] (x0_stride floats between rows; 0 broadcasts one origin), R: [batch,3] likewise */
int bgx_ic_to_xyz(const bgx_zplan* plan, int64_t batch, const float* bonds, const float* angles,
                  const float* torsions, const float* x0, int32_t x0_stride, const float* R,
                  int32_t r_stride, float* xyz, const float* dlogp_in, float* dlogp_out,
                  void* stream);

int bgx_ic_from_xyz(const bgx_zplan* plan, int64_t batch, const float* xyz, float* bonds,
                    float* angles, float* torsions, float* x0, float* R, const float* dlogp_in,
                    float* dlogp_out, void* stream);

/* ---- IC-domain maps ------------------------------------------------------- */

/* Column-wise CDF maps: CDFTransform._forward/_inverse (bgflow/nn/flow/cdf.py:29-46) over the
 * marginals the builder installs with add_map_to_ic_domains (factory/generator_builder.py:443-459,
 * factory/icmarginals.py:39-77): TruncatedNormalDistribution (bgflow/distribution/normal.py:95-227)
 * for bonds / angles, (Sloppy)Uniform for torsions, torch.distributions.Normal for fixed /
 * augmented fields.  One bgx_cdf_col per tensor column. */
#define BGX_DIST_NONE 0          /* identity column, log-det 0 */
#define BGX_DIST_NORMAL 1        /* Normal(loc, scale) */
#define BGX_DIST_TRUNCNORMAL 2   /* Normal(mu, sigma) truncated to [lower, upper] */
#define BGX_DIST_UNIFORM 3       /* Uniform(low, high) */

typedef struct bgx_cdf_col {
  int32_t kind;
  float p[7];   /* filled by bgx_cdf_col_init: loc|low, scale|high-low, Phi(alpha), Z, log-normaliser,
                   1/scale, high */
} bgx_cdf_col;

/* Host helper: derive the per-column constants in double precision.
 *   NORMAL       a = loc, b = scale                     (lower / upper ignored)
 *   TRUNCNORMAL  a = mu,  b = sigma, lower, upper       (+-INFINITY allowed)
 *   UNIFORM      a = low, b = high
 * Returns BGX_ERR_INVALID for scale <= 0, upper <= lower or an unknown kind. */
int bgx_cdf_col_init(int32_t kind, double a, double b, double lower, double upper, bgx_cdf_col* out);

/* TruncatedNormalDistribution freezes Phi(alpha), Phi(beta) in registered buffers at construction
 * (normal.py:148-151) and evaluates cdf / icdf / log_prob with THOSE, even after mu / sigma were trained or
 * loaded from a checkpoint: override the constants bgx_cdf_col_init derived from mu / sigma. */
int bgx_cdf_col_set_truncation(bgx_cdf_col* col, double cdf_lower, double cdf_upper);

/* Map `n_seg` tensors of one flow state in ONE launch.  Segment i has in[i].width columns whose
 * bgx_cdf_col entries follow each other in `cols` (device array, sum of widths entries);
 * out[i] has the same width (out[i].ptr is written; may alias in[i].ptr).
 *   flags & BGX_FLAG_INVERSE : u -> x = icdf(clamp(u, clamp_lo, clamp_hi)), log-det = max(-log_prob(x), logdet_min)
 *                              (CDFTransform._inverse = what InverseFlow(CDFTransform) runs when sampling)
 *   otherwise                : x -> u = clamp(cdf(x), clamp_lo, clamp_hi), log-det = max(log_prob(x), logdet_min)
 * The reference's eps = 1e-7 is clamp_lo = (float)eps, clamp_hi = (float)(1 - eps),
 * logdet_min = -1/eps; eps = None is (0, 1, -INFINITY).
 * dlogp_out[row] = (dlogp_in ? dlogp_in[row] : 0) + sum over all columns of the log-dets. */
int bgx_cdf_map(int64_t batch, int32_t n_seg, const bgx_seg* in, const bgx_seg* out,
                const bgx_cdf_col* cols, float clamp_lo, float clamp_hi, float logdet_min, int flags,
                const float* dlogp_in, float* dlogp_out, void* stream);

/* bgx_ic_to_xyz with the icdf maps of the IC fields applied on the fly (the tail of every builder
 * stack: WrapFlow(InverseFlow(CDFTransform(marginal))) per field followed by
 * InverseFlow(GlobalInternalCoordinateTransformation), generator_builder.py:408-459): bonds, angles
 * and torsions arrive in [0,1], `marginals` (device, 3N-6 entries in bonds | angles | torsions column
 * order) maps them to their IC domains, then the atoms are placed; dlogp gets both log-dets. */
int bgx_ic_to_xyz_mapped(const bgx_zplan* plan, const bgx_cdf_col* marginals, float clamp_lo,
                         float clamp_hi, float logdet_min, int64_t batch, const float* bonds,
                         const float* angles, const float* torsions, const float* x0,
                         int32_t x0_stride, const float* R, int32_t r_stride, float* xyz,
                         const float* dlogp_in, float* dlogp_out, void* stream);

/* The reverse tail (energy / NLL direction): Cartesian -> ICs -> cdf of every IC column. */
int bgx_ic_from_xyz_mapped(const bgx_zplan* plan, const bgx_cdf_col* marginals, float clamp_lo,
                           float clamp_hi, float logdet_min, int64_t batch, const float* xyz,
                           float* bonds, float* angles, float* torsions, float* x0, float* R,
                           const float* dlogp_in, float* dlogp_out, void* stream);

/* ---- relative / mixed internal coordinates ------------------------------ */

/* RelativeInternalCoordinateTransformation (bgflow/nn/flow/crd_transform/ic.py:268-513): the
 * atoms in `fixed` keep their Cartesian coordinates, every other atom i has a z-matrix row
 * (i, j, k, l).  With keepdims > 0 the plan is a MixedCoordinateTransformation (ic.py:719-884):
 * the fixed block is additionally whitened by a static PCA (WhitenFlow, crd_transform/pca.py:37-107):
 *   forward  z_fixed = (x_fixed - mean) . whiten      dlogp += log_det_whiten
 *   inverse  x_fixed = z_fixed . blacken + mean       dlogp -= log_det_whiten */
typedef struct bgx_relplan {
  int32_t n_atoms, n_fixed, n_rel;   /* n_rel = n_atoms - n_fixed */
  const int32_t* fixed;              /* device [n_fixed] atom ids, in the column order of x_fixed */
  const int32_t* rel;                /* device [n_rel][4] rows (i,j,k,l) = column order of the IC tensors */
  const int32_t* order;              /* device [n_rel] placement order (row ids): j,k,l before i */
  int32_t normalize_angles;
  float eps;
  int32_t keepdims;                  /* 0 = relative transform; > 0 = mixed (whitened fixed block) */
  const float* mean;                 /* device [3 n_fixed] */
  const float* blacken;              /* device [keepdims][3 n_fixed] */
  const float* whiten;               /* device [3 n_fixed][keepdims] */
  float log_det_whiten;              /* WhitenFlow.jacobian_xz = -sum(log std) */
} bgx_relplan;

/* (bonds, angles, torsions [batch, n_rel]; fixed [batch, keepdims or 3 n_fixed]) -> xyz [batch, 3 n_atoms] */
int bgx_relic_to_xyz(const bgx_relplan* plan, int64_t batch, const float* bonds, const float* angles,
                     const float* torsions, const float* fixed, float* xyz, const float* dlogp_in,
                     float* dlogp_out, void* stream);

int bgx_relic_from_xyz(const bgx_relplan* plan, int64_t batch, const float* xyz, float* bonds,
                       float* angles, float* torsions, float* fixed, const float* dlogp_in,
                       float* dlogp_out, void* stream);

/* SplitFlow._forward / MergeFlow (bgflow/nn/flow/coupling.py:46-62, 107-110) by sizes along the last
 * dim as one launch: column blocks of `whole` ([batch, width], row stride whole->stride) <-> `parts`
 * (widths must add up to whole->width).  merge == 0 writes the parts, merge != 0 writes `whole`. */
int bgx_split_merge(int64_t batch, const bgx_seg* whole, int32_t n_parts, const bgx_seg* parts, int merge,
                    void* stream);

/* Training path: exact two-term bf16 split of an fp32 tensor, hi = rn(x), lo = rn(x - hi) (x = hi + lo to 2^-17 |x|).
 * The backward GEMMs of the conditioner (bgflow_b200/_mlp_grad.py; the reference differentiates dense.py:47-48 with
 * autograd) run as three bf16 tensor-core products hi.hi + hi.lo + lo.hi with fp32 accumulation, like the forward
 * kernels.  x: n floats, 16-byte aligned; hi / lo: n bf16 values each, 8-byte aligned. */
int bgx_split_bf16(const float* x, int64_t n, void* hi, void* lo, void* stream);

/* Training path: Y[batch, N] = X[batch, K] . W^T + b on tcgen05 with exact two-term bf16 operand splits and fp32
 * accumulation (fp32-class accuracy), for the layer GEMMs of the conditioner's recompute and input-gradient backward
 * (dense.py:47-48 differentiated; the reference uses torch autograd).  `layer` = bgx_pack_mlp of a ONE-layer bgx_mlp
 * {dims = {K, N}, W = [N, K], b = [N]}.  Supported: K <= 128 with any N, or N <= 128 with any K (else
 * BGX_ERR_UNSUPPORTED).  x, y dense row-major; `status` (or NULL) as in the bgx_spline_cfg struct. */
int bgx_linear(int64_t batch, const float* x, const bgx_packed_mlp* layer, float* y, int32_t* status, void* stream);

/* Weight gradient of a layer on the tensor cores (same operand splits): the batch-reduced products
 *   dW[n, k] = sum_b g[b, n] h[b, k]   and   db[n] = sum_b g[b, n]        (k <= 128, any n; else BGX_ERR_UNSUPPORTED)
 * g = [batch, n] with row stride ldg, h = [batch, k] with row stride ldh (floats).  The batch is cut into `slices`
 * (= bgx_gemm_tn_slices(batch, n)) contiguous ranges; every range writes its own partial sums
 *   part_w[slices][128 ceil(n / 128)][128]   and   part_b[slices][128 ceil(n / 128)]   (part_b may be NULL),
 * which the caller adds up over the first index (fixed order = reproducible gradients); rows >= n and columns >= k
 * of the partials are zero. */
int bgx_gemm_tn_slices(int64_t batch, int n);
int bgx_gemm_tn(int64_t batch, const float* g, int64_t ldg, int n, const float* h, int64_t ldh, int k, int slices,
                float* part_w, float* part_b, int32_t* status, void* stream);

/* ---- training path: a conditioner's recompute and backward as one call each ------------------------
 * (dense.py:10-48 under torch autograd in the reference.)  bgx_train_pack lays every layer out twice for bgx_linear's
 * kernel: as y = x W^T + b (fwd) and transposed, dx = g W (bwd), straight from nn.Linear.weight.  Supported: every
 * layer has <= 128 inputs or <= 128 outputs (else BGX_ERR_UNSUPPORTED); no WrapPeriodic (n_periodic == 0).
 * `act` = n_layers - 1 activation codes (between the layers) or NULL for src->act everywhere. */
typedef struct bgx_train_mlp {
  int32_t n_layers;
  int32_t dims[BGX_MAX_LAYERS + 1];       /* true widths */
  int32_t act[BGX_MAX_LAYERS];            /* activation after layer i (BGX_ACT_NONE after the last) */
  bgx_packed_mlp fwd[BGX_MAX_LAYERS];
  bgx_packed_mlp bwd[BGX_MAX_LAYERS];
  int64_t total_floats;
} bgx_train_mlp;
/* dst == NULL: size query (out->total_floats). */
int bgx_train_pack(const bgx_mlp* src, const int32_t* act, float* dst, int64_t dst_floats, bgx_train_mlp* out,
                   void* stream);

/* Caller-owned device buffers of one recompute + backward.  pad4(n) = n rounded up to a multiple of 4. */
typedef struct bgx_train_buffers {
  float* z[BGX_MAX_LAYERS];   /* [batch, pad4(dims[i + 1])] pre-activations of layer i; z[n_layers - 1] = the net's output */
  float* h[BGX_MAX_LAYERS];   /* [batch, pad4(dims[i + 1])] act(z[i]), i < n_layers - 1 */
  float* g[BGX_MAX_LAYERS];   /* [batch, pad4(dims[i + 1])] gradient wrt z[i], i < n_layers - 1 (scratch) */
  float* part;                /* bgx_mlp_train_part_floats(batch, net) floats (scratch of the weight gradients) */
} bgx_train_buffers;
int64_t bgx_mlp_train_part_floats(int64_t batch, const bgx_train_mlp* net);

/* Recompute: fills buf->z[*] and buf->h[*] from x [batch, dims[0]] (dense).  Pad columns come out as zeros. */
int bgx_mlp_forward_train(int64_t batch, const bgx_train_mlp* net, const float* x, const bgx_train_buffers* buf,
                          int32_t* status, void* stream);
/* Backward from d_out [batch, pad4(dims[n_layers])] (pad columns must be zero) with the buffers of the recompute:
 * d_w[i] [dims[i + 1], dims[i]] and d_b[i] [dims[i + 1]] (dense, overwritten), d_x [batch, dims[0]] or NULL. */
int bgx_mlp_backward(int64_t batch, const bgx_train_mlp* net, const float* x, const bgx_train_buffers* buf,
                     const float* d_out, float* d_x, float* const* d_w, float* const* d_b, int32_t* status, void* stream);

/* The whole backward of one RQ-spline coupling block in one call (coupling.py:161-180, spline.py:87-188 and
 * dense.py:47-48 under torch autograd in the reference): conditioner recompute, the transform's chain rule
 * (bgx_spline_backward) and the conditioner's backward.  cond [batch, dims[0]] and y [batch, d_t] (the block's
 * transformed INPUT) dense; g_out [batch, d_t], g_dlogp [batch] or NULL the upstream gradients; flags / cfg /
 * end_slope_col as for bgx_spline_backward.  Outputs: d_cond [batch, dims[0]] (or NULL), d_y [batch, d_t], d_w[i],
 * d_b[i] as for bgx_mlp_backward.  d_p: scratch, batch * pad4(dims[n_layers]) floats. */
int bgx_spline_coupling_backward(int64_t batch, const bgx_train_mlp* net, const float* cond, int32_t d_t,
                                 const float* y, const float* g_out, const float* g_dlogp,
                                 const int32_t* end_slope_col, const bgx_spline_cfg* cfg, int flags,
                                 const bgx_train_buffers* buf, float* d_p, float* d_cond, float* d_y,
                                 float* const* d_w, float* const* d_b, int32_t* status, void* stream);

/* The whole backward of one RealNVP coupling block in one call (coupling.py:161-180, affine.py:35-70 and dense.py:47-48
 * under torch autograd in the reference): both conditioners recomputed, the transform's chain rule, both conditioner
 * backwards.  Plain blocks only (shift and scale nets with dims[0] equal and dims[n_layers] == d_t; no volume
 * preservation, no circular wrap).  log_alpha: the device scalar of AffineTransformer._log_alpha; cond, y, g_out,
 * g_dlogp, flags as for bgx_spline_coupling_backward; scratch: bgx_affine_backward_scratch_floats(...) floats.
 * Outputs: d_cond [batch, dims[0]], d_y [batch, d_t], the weight / bias gradients of both nets, d_log_alpha[1]. */
int64_t bgx_affine_backward_scratch_floats(int64_t batch, int32_t d_t, int32_t cond_width);
int bgx_affine_coupling_backward(int64_t batch, const bgx_train_mlp* shift, const bgx_train_mlp* scale,
                                 const float* log_alpha, const float* cond, int32_t d_t, const float* y,
                                 const float* g_out, const float* g_dlogp, int flags,
                                 const bgx_train_buffers* buf_shift, const bgx_train_buffers* buf_scale,
                                 float* scratch, float* d_cond, float* d_y, float* const* d_w_shift,
                                 float* const* d_b_shift, float* const* d_w_scale, float* const* d_b_scale,
                                 float* d_log_alpha, int32_t* status, void* stream);

/* ---- misc -------------------------------------------------------------------------------- */

/* Self-test of the tcgen05 / TMEM / bulk-TMA building blocks: out[128][128] = A[128][K] . W[128][K]^T
 * on one CTA (mode 0: A from shared memory, mode 1: A from tensor memory).  K in {32,64,96,128};
 * scratch >= 128*K floats; *status (device int) becomes 1 if a barrier wait timed out. */
int bgx_tc_selftest(int mode, const float* A, const float* W, int K, float* scratch, float* out,
                    int* status, void* stream);

/* Device int the tensor-core kernels set to 1 if their internal barrier pipeline ever times out
 * (results invalid; a bug, never an input property).  Used when a call carries no status pointer. */
int bgx_set_status_buffer(int32_t* device_int);

/* Debug: subsequent tensor-core launches record a timeline of CTA 0 into `device_buffer`
 * ([0] = event count, then (clock64, code) pairs; 1 + 2*capacity uint64).  NULL disables. */
int bgx_debug_set_trace(uint64_t* device_buffer, int capacity);

/* Which kernel family served the coupling calls so far (tests assert the tensor-core path ran). */
#define BGX_KERNEL_SPLINE_PAIR 0        /* spline_coupling_pair_kernel, tiles in shared memory */
#define BGX_KERNEL_SPLINE_PAIR_WIDE 1   /* spline_coupling_pair_kernel, wide (config 5) mode */
#define BGX_KERNEL_SPLINE_TC2 2
#define BGX_KERNEL_SPLINE_TC 3
#define BGX_KERNEL_SPLINE_SIMT 4
#define BGX_KERNEL_AFFINE_TC2 5
#define BGX_KERNEL_AFFINE_TC 6
#define BGX_KERNEL_AFFINE_SIMT 7
#define BGX_KERNEL_AFFINE_PAIR 8
#define BGX_KERNEL_AFFINE_PAIR_WIDE 9
#define BGX_KERNEL_IDS 16
int64_t bgx_kernel_count(int kernel_id);

const char* bgx_version(void);
const char* bgx_last_cuda_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t bgx_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BGFLOW_B200_H */
