/*
 * nerf_b200.h -- C ABI of the B200-native NeRF volume-rendering path.
 *
 * The reference (JulianKnodt/nerf_atlas) has no FFI: its surface for this path is
 * the Python duck-type  model(rays[B,H,W,6]) -> rgb[B,H,W,3]  (reference
 * src/nerf.py:326-361, called from runner.py:490-509).  This header is the
 * boundary a maintainer binds that surface to (see INTEGRATION.md for the ctypes
 * stub).  Conventions:
 *   - every export returns int: 0 = ok, >0 = cudaError_t, <0 = NF_E_* below;
 *     nf_last_error() gives the message (thread local).  Nothing throws or exits.
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the
 *     name ends in _host; tensors are contiguous fp32 row-major.
 *   - the library owns no parameters: nf_pack_weights() snapshots the caller's
 *     live parameter tensors into a caller-allocated `packed` blob.
 *   - every launch goes to the caller's stream; no global mutable state.
 */
#ifndef NERF_B200_H
#define NERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NF_ABI_VERSION 7

/* error codes (negative; positive values are cudaError_t) */
#define NF_E_BADARG    (-1)
#define NF_E_UNSUPPORTED (-2)
#define NF_E_SMALLBUF  (-3)

/* activation that PRECEDES each hidden Linear (reference src/neural_blocks.py:290-296) */
enum nf_act { NF_ACT_NONE = 0, NF_ACT_LEAKY = 1 /* LeakyReLU(0.01) */, NF_ACT_SIN = 2, NF_ACT_RELU = 3 };
/* input encoder of the density MLP (reference src/neural_blocks.py:36-55,92-193) */
enum nf_enc { NF_ENC_NONE = 0, NF_ENC_HASH = 1, NF_ENC_FOURIER = 2 };
/* raw density -> sigma (reference src/nerf.py:60-65) */
enum nf_density_act {
  NF_DENS_SOFTPLUS_M1 = 0, /* softplus(x-1) */
  NF_DENS_RELU = 1,
  NF_DENS_LAPLACE = 2      /* VolSDF: relu(laplace_cdf(-sdf, beta) / beta), beta = the learned `scale` (reference
                              src/nerf.py:1000-1003, src/utils.py:50-58) */
};
/* feature activation = the sigmoid family of reference src/utils.py:484-518 */
enum nf_feat_act { NF_FEAT_NORMAL = 0, NF_FEAT_THIN = 1, NF_FEAT_TANH = 2, NF_FEAT_CYCLIC = 3, NF_FEAT_UPSHIFTED = 4,
                   NF_FEAT_FAT = 5, NF_FEAT_LEAKY_RELU = 6, NF_FEAT_RELU = 7, NF_FEAT_SIN = 8,
                   NF_FEAT_UPSHIFTED_SOFTPLUS = 9, NF_FEAT_UPSHIFTED_RELU = 10,
                   NF_FEAT_SOFTMAX = 11 /* nn.Softmax(dim=-1) over the three colour channels (utils.py:507) */ };
/* background (reference src/nerf.py:96-109) */
enum nf_bg { NF_BG_BLACK = 0, NF_BG_WHITE = 1,
             NF_BG_RANDOM = 2 /* random_color (nerf.py:100-103): rand * (1 - sum_{t<T-1} w_t), one draw per ray, passed in (nf_render_aux.bg_rand) */ };
/* model family */
enum nf_kind {
  NF_KIND_PLAIN = 0, /* density MLP -> [raw density | sdf, intermediate(I)] -> View head -> composite:
                        PlainNeRF + View (reference src/nerf.py:310-361, src/refl.py:190-207) and the volume branch of
                        VolSDF (reference src/nerf.py:981-1013, src/sdf.py:109-112,250-287) */
  NF_KIND_TINY  = 1, /* TinyNeRF (intended semantics): reference src/nerf.py:278-305 */
  NF_KIND_DYN   = 2  /* DynamicNeRF, direct deformation MLP over a canonical PlainNeRF (reference src/nerf.py:1209-1303):
                        delta_estim([p, t]) -> (dp[1], rigidity[3]); p' = p + dp * sigmoid(rigidity / 2); then NF_KIND_PLAIN on p' */
};
/* Mip-NeRF integrated positional encoding appended to BOTH MLP inputs (reference src/nerf.py:255-261,340-358;
 * src/utils.py:39-48,60-140): 96 features per sample, E[sin/cos(2^k x)], k = 0..15, of the Gaussian of the ray segment
 * [t_i, t_i+1].  x0 becomes [p, enc(p), ipe(96)] and the View head's [p, elaz, ipe(96), intermediate(I)]. */
enum nf_mip {
  NF_MIP_NONE = 0,
  NF_MIP_CYLINDER = 1,     /* CylinderGaussian as intended: per-sample variance, last segment capped to t[T-1]+(t[T-1]-t[T-2]) */
  NF_MIP_CONE = 2,         /* ConicGaussian as intended (same layout and cap; the reference's own cone run is all-NaN) */
  NF_MIP_CYLINDER_REF = 3  /* CylinderGaussian exactly as the reference computes it, bug for bug: last segment ends at 1e10
                              and the variance of feature c of sample (t, ray) is taken from the element with the same flat
                              index of the [xyz, ray, k, t] covariance array (utils.py:73 moves the wrong axis, 42-44
                              reinterpret) -- needs every ray of the crop: nf_mip_args.rays_all / radius_all */
};
/* RGB head */
enum nf_refl {
  NF_REFL_VIEW = 0,        /* refl.View (reference src/refl.py:190-207): x0 = [p, elaz(view), latent] */
  NF_REFL_POSITIONAL = 1,  /* refl.Positional (reference src/refl.py:230-245; makefile:12 `--refl-kind pos`): view independent,
                              x0 = [p, p, hash'(p), latent] with the head's OWN HashEncoder, 5 layers, LeakyReLU.  Both pipelines
                              (tensor pipeline: x0 is 112 columns wide, two tiles in flight over one shared x0 buffer) */
  NF_REFL_POSLINVIEW = 2   /* refl.PosLinearView (reference src/refl.py:248-290; `--refl-kind pos-linear-view`): rgb = (sigmoid(view) / 2
                              + 1/2) * pos, with  [pos(3), im(I2)] = act(pos_mlp([p, p, hash'(p), latent]))  (own HashEncoder, 2 layers,
                              hidden 256, LeakyReLU; the feature activation on all 3 + I2 outputs) and  view = view_mlp([p,
                              normalize(r_d), latent, im])  (2 layers, hidden 128, sin).  `refl` describes pos_mlp, `refl_view`
                              view_mlp.  NF_PREC_FP32 only; not with NF_KIND_DYN, Mip or the softmax activation. */
};
/* arithmetic of the MLP contractions */
enum nf_precision {
  NF_PREC_FP32 = 0, /* CUDA-core fp32 FMA; the exact mode (matches the reference's SGEMM to ~1e-6) */
  NF_PREC_FP16_TC = 1 /* tcgen05 tensor cores, fp16 operands, fp32 accumulate in TMEM */
};

/* One SkipConnMLP (reference src/neural_blocks.py:204-296). */
typedef struct nf_mlp_desc {
  int32_t in_dims;   /* dim_p = width of x0 = [p, enc(p), latent] */
  int32_t hidden;    /* hidden_size (256 in every reference instance on this path) */
  int32_t n_layers;  /* len(self.layers) */
  int32_t out_dims;  /* out.out_features */
  int32_t skip;      /* x0 is re-concatenated before hidden layer i iff i%skip==0 && i!=n_layers-1 */
  int32_t act;       /* enum nf_act */
} nf_mlp_desc;

typedef struct nf_model_desc {
  int32_t struct_bytes;      /* = sizeof(nf_model_desc); ABI guard */
  int32_t kind;              /* enum nf_kind */
  nf_mlp_desc density;       /* PlainNeRF.first / TinyNeRF.estim */
  nf_mlp_desc refl;          /* View.mlp (ignored for NF_KIND_TINY) */
  int32_t intermediate;      /* I = intermediate_size */
  int32_t enc;               /* enum nf_enc, encoder of `density` */
  int32_t hash_levels;       /* 8 */
  int32_t hash_table_size;   /* 65536 (power of two) */
  int32_t hash_feat;         /* 4 */
  uint32_t hash_primes[3];   /* 1, 2654435761, 805459861 */
  float hash_res[16];        /* fp32(N_l) per level, exactly as `x * N_l` rounds it */
  int32_t density_act;       /* enum nf_density_act */
  int32_t feat_act;          /* enum nf_feat_act */
  int32_t bg;                /* enum nf_bg */
  int32_t fourier_freqs;     /* NF_ENC_FOURIER: columns of the basis [3, freqs] (x0 = [p, sin(pB), cos(pB)]) */
  nf_mlp_desc deform;        /* NF_KIND_DYN: DynamicNeRF.delta_estim (direct: in 4 = xyz,t, out 4; spline: in 38, out 1+3n) */
  int32_t mip;               /* enum nf_mip (both pipelines; tensor pipeline: x0 is 144 / 176 columns wide, shared-x0 mode) */
  int32_t deform_enc;        /* NF_KIND_DYN: encoder of `deform`: NF_ENC_NONE (direct) or NF_ENC_HASH (spline; its own tables) */
  int32_t refl_kind;         /* enum nf_refl */
  int32_t spline_points;     /* NF_KIND_DYN: 0 = direct_predict (nerf.py:1261-1266); n in 2..8 = spline_interpolate with n
                                Bezier control points (nerf.py:1267-1278; de_casteljau 1173-1178, cubic_bezier 1201-1206) */
  nf_mlp_desc refl_view;     /* NF_REFL_POSLINVIEW: PosLinearView.view (hidden 128: evaluated as 256 with zero-padded weights, exact) */
} nf_model_desc;

/* Per-call inputs of the Mip encoder (nf_model_desc.mip != NF_MIP_NONE); all device pointers. */
typedef struct nf_mip_args {
  const float* radius;       /* [R] pixel radius of every ray of this call (nf_ray_radii) */
  const float* rays_all;     /* NF_MIP_CYLINDER_REF: rays of the WHOLE crop [R_all,6] (== rays when not sharded) */
  const float* radius_all;   /* NF_MIP_CYLINDER_REF: [R_all] */
  int64_t n_rays_all;        /* R_all = B*H*W of the crop */
  int64_t ray_base;          /* index of this call's first ray within the crop */
} nf_mip_args;

/* ---- library ----------------------------------------------------------- */
int nf_version(void);
const char* nf_last_error(void);
/* How the library was built: the default build reads NO environment variable and contains only the product kernels;
 * NF_BUILD_EXPERIMENTS builds additionally compile the superseded pipelines and the NF_TC_* timing switches (A/B timing,
 * profiles/perf_variants.py) -- never ship one. */
#define NF_BUILD_EXPERIMENTS 1
#define NF_BUILD_STATS 2
#define NF_BUILD_TRACE 4
int nf_build_flags(void);

/* NULL if nf_render_forward(..., NF_PREC_FP16_TC) can run `desc` on the tcgen05 pipeline, else the reason (a static string, or
 * nf_last_error() text for an invalid descriptor).  Host-only: no CUDA call, usable without a GPU.  There is no silent
 * fallback to NF_PREC_FP32: an unsupported descriptor makes nf_render_forward fail with NF_E_UNSUPPORTED. */
const char* nf_tensor_pipeline_support(const nf_model_desc* desc);

/* ---- parameters -------------------------------------------------------- */
/* Number of parameter pointers nf_pack_weights expects for `desc`, in this order:
 *   density MLP: init.weight, init.bias, layers[0].weight, layers[0].bias, ..., out.weight, out.bias
 *   refl MLP   : same order                                   (NF_KIND_PLAIN, NF_KIND_DYN; NF_REFL_POSLINVIEW: PosLinearView.pos)
 *   deform MLP : same order                                   (NF_KIND_DYN only)
 *   view MLP   : same order                                   (NF_REFL_POSLINVIEW only: PosLinearView.view)
 *   hash tables: embs[0].weight ... embs[levels-1].weight      (NF_ENC_HASH only)
 *   deform hash: delta_estim.enc.embs[0..levels-1].weight      (NF_KIND_DYN with deform_enc == NF_ENC_HASH only)
 *   refl hash  : refl.mlp.enc.embs[0..levels-1].weight         (NF_REFL_POSITIONAL; NF_REFL_POSLINVIEW: refl.pos.enc.embs)
 *   fourier    : enc.basis [3, freqs]                          (NF_ENC_FOURIER only)
 *   beta       : VolSDF.scale (scalar)                         (NF_DENS_LAPLACE only)
 * Each is the live fp32 nn.Parameter storage ([out,in] row-major for weights). */
int nf_param_count(const nf_model_desc* desc);
/* Bytes of the packed blob for `desc` (all precisions). */
int64_t nf_packed_bytes(const nf_model_desc* desc);
/* Snapshot the parameters into `packed` (device, >= nf_packed_bytes, 1024-byte aligned):
 * fp32 k-major copies for NF_PREC_FP32, fp16 UMMA-canonical weight chunks for
 * NF_PREC_FP16_TC, biases, and the hash tables. Replaces: the implicit read of
 * nn.Parameter by F.linear / nn.Embedding (reference src/neural_blocks.py:166,289-296). */
int nf_pack_weights(const nf_model_desc* desc, const float* const* params_host, int32_t n_params,
                    void* packed, int64_t packed_bytes, void* stream);

/* ---- the hot path ------------------------------------------------------ */
/* rays[R,6] -> rgb[R,out].  Replaces PlainNeRF.forward / TinyNeRF.forward
 * (reference src/nerf.py:326-361, 292-305):  sample positions (nerf.py:50-55),
 * encode, both MLPs, alpha_from_density (nerf.py:60-73), volumetric_integrate
 * (nerf.py:79-80) and the sky term.
 *   ts             sample distances; ts_ray_stride == 0: one ts[T] shared by all rays
 *                  (the reference's layout); == T: per-ray ts[R,T] (coarse+fine pass)
 *   density_noise  nullable [R,T]; added to the raw density (nerf.py:347-348, already scaled)
 *   ray_time       NF_KIND_DYN: time of every ray, [R] (the reference broadcasts times[B] over the view's pixels,
 *                  nerf.py:1301); NULL otherwise
 *   mip            nf_model_desc.mip != NF_MIP_NONE: radii (and, for NF_MIP_CYLINDER_REF, the whole crop); needs a shared
 *                  ts[T] (ts_ray_stride == 0, as in the reference); NULL otherwise
 *   alpha_out, weights_out  nullable [R,T] (ray-major; the reference's self.alpha/self.weights
 *                  are the [T,R] transposes)
 */
int nf_render_forward(const nf_model_desc* desc, const void* packed,
                      const float* rays, int64_t n_rays,
                      const float* ts, int32_t T, int64_t ts_ray_stride,
                      const float* density_noise, const float* ray_time, const nf_mip_args* mip,
                      float* rgb_out, float* alpha_out, float* weights_out,
                      int32_t precision, void* stream);

/* ---- stages of the path, exported for parity tests and micro-benchmarks -- */
/* Camera rays of a crop: runner.render's pixel grid + NeRFCamera.sample_positions with_noise=False
 * (reference runner.py:490-505, src/cameras.py:45-66):  pixel (row top+h, column left+w) of view b ->
 *   d = [(u - size/2) / focal, -(v - size/2) / focal, -1],  u = column, v = row
 *   r_d[c] = (d0*R[c][0] + d1*R[c][1]) + d2*R[c][2]  (R = cam_to_world[b,:3,:3]; NOT normalised),  r_o = cam_to_world[b,:3,3]
 * cam_to_world is a DEVICE pointer to [B,3,4] fp32; rays_out[B,H,W,6].  Bit-exact against the reference camera.
 * scalar_div_as_reciprocal = 0: IEEE division by fp32(focal) (what torch does on the CPU); 1: multiply by 1.0f/focal (what
 * torch's CUDA kernel does when dividing by a Python scalar) -- pick the device the reference ran on. */
int nf_generate_rays(const float* cam_to_world, int64_t B, float focal, int32_t size, int32_t top, int32_t left,
                     int32_t H, int32_t W, int32_t scalar_div_as_reciprocal, float* rays_out, void* stream);
/* Camera rays of a DTU / IDR camera (BASELINE config 4): runner.render's pixel grid + DTUCamera.sample_positions (reference
 * src/cameras.py:159-174,189-223): pixel (u = column, v = row) scaled by (1600, 1200) / size, lifted through the intrinsics
 * (fx, skew, cx, fy, cy), transformed by pose[b] (4x4, camera to world), r_d = normalize(world - r_o): UNIT-norm directions.
 * pose[B,4,4] and intrinsic[B,intr_rows,intr_cols] (>= 3x3) are DEVICE pointers.  Matches the reference camera to 1e-6 (its
 * torch.bmm fixes no summation order, so the last bit is not pinned). */
int nf_generate_rays_dtu(const float* pose, const float* intrinsic, int32_t intr_rows, int32_t intr_cols, int64_t B, int32_t size,
                         int32_t top, int32_t left, int32_t H, int32_t W, float* rays_out, void* stream);
/* radii_x (reference src/utils.py:77-81) on a crop of rays[B,H,W,6] (H >= 3): radius_out[B,H,W] = |r_d[h] - r_d[h+1]| * 2/sqrt(12),
 * the last row repeating difference H-3 exactly like the reference's `dx[:, -2:-1, :]`. */
int nf_ray_radii(const float* rays, int64_t B, int32_t H, int32_t W, float* radius_out, void* stream);
/* pts[R,T,3] = r_o + ts*r_d, rounded product then rounded add (reference src/nerf.py:54). */
int nf_sample_points(const float* rays, int64_t n_rays, const float* ts, int32_t T, int64_t ts_ray_stride,
                     float* pts_out, void* stream);
/* HashEncoder.forward (reference src/neural_blocks.py:139-193) on pts[N,3]:
 * feats_out[N, levels*feat]; idx_out nullable uint16 [levels, 8, N] table rows per corner. */
int nf_hash_encode(const nf_model_desc* desc, const void* packed, const float* pts, int64_t n,
                   float* feats_out, uint16_t* idx_out, void* stream);
/* alpha_from_density + volumetric_integrate (+sky) (reference src/nerf.py:60-80) on
 * sigma_raw[R,T], feats[R,T,3] (already activated). HBM-bound stand-alone form of the fused tail.
 * `packed` is only read for NF_DENS_LAPLACE (beta); it may be NULL otherwise. */
int nf_composite(const nf_model_desc* desc, const void* packed, const float* sigma_raw, const float* feats,
                 const float* rays, int64_t n_rays, const float* ts, int32_t T, int64_t ts_ray_stride,
                 float* rgb_out, float* alpha_out, float* weights_out, void* stream);
/* volumetric_integrate (reference src/nerf.py:79-80) of per-sample values other than the colours, over the weights a render kept:
 *   out[r, c] = sum_t weights[r, t] * vals[r * vals_ray_stride + t * channels + c]
 * depth (runner.py:511-514,894-897): vals = ts, channels 1, vals_ray_stride 0 (shared ts[T]) or T' (per-ray ts); flow / rigidity
 * maps (runner.py:521-531,909-914): vals = the rigid_dp / rigidity side channels of nf_render_aux ([R,T,3] / [R,T,1] ray-major),
 * vals_ray_stride = T * channels.  channels is 1 or 3.  HBM-bound: 4 T (1 + channels) bytes read per ray. */
int nf_integrate(const float* weights, const float* vals, int64_t n_rays, int32_t T, int32_t channels, int64_t vals_ray_stride,
                 float* out, void* stream);
/* Hierarchical resampling for the coarse+fine configuration: restatement of the reference's dead
 * sample_pdf (reference src/nerf.py:1745-1779, called from CoarseFineNeRF.from_pts nerf.py:573-577):
 * bins = mid-points of ts_coarse[T]; pdf from weights[r, 1:T-1] + 1e-5; inverse-CDF at u[R,Nf] in [0,1)
 * (searchsorted right, linear interpolation); the Nf new positions are merged with ts_coarse and
 * written sorted: ts_out[R, T+Nf] (the per-ray ts of the fine pass, ts_ray_stride = T+Nf). */
int nf_sample_pdf(const float* ts_coarse, int32_t T, const float* weights, int64_t n_rays,
                  const float* u, int32_t n_fine, float* ts_out, void* stream);
/* One SkipConnMLP.forward (reference src/neural_blocks.py:279-296) on assembled inputs
 * x0[N,in_dims] -> out[N,out_dims]; which = 0 density MLP (or the SDF network), 1 refl MLP, 2 deformation MLP (NF_KIND_DYN).
 * NF_PREC_FP16_TC: MLPs whose x0 fits 80 columns (the single-CTA tcgen05 kernel of nf_tc.cu). */
int nf_mlp_forward(const nf_model_desc* desc, const void* packed, int32_t which,
                   const float* x0, int64_t n, float* out, int32_t precision, void* stream);

/* ---- training: forward with an activation stash, backward of the whole path ------------------------------------------
 * What the reference gets from PyTorch autograd (loss.backward(), runner.py:820) for model(rays): gradients of every
 * parameter of the path given d loss / d rgb.  NF_PREC_FP16_TC only (tcgen05 forward AND backward): PlainNeRF (hash-encoded density
 * MLP) with the View or the Positional head (refl.py:230-245; T % 32 == 0), TinyNeRF (one MLP), and VolSDF's volume branch with the SIREN SDF + View (src/nerf.py:981-1013) including the learned beta (`scale`).
 *   1. nf_train_layout_of()    -> workspace size (total_bytes) and where everything lives in it
 *   2. nf_render_forward_aux() with aux.train_ws: the normal forward, plus the stash (per 128-sample tile and Linear: the
 *      input operand as the MMA consumed it, fp16; cos of the pre-activations for sin MLPs; raw density / colours per sample)
 *   3. nf_render_backward()    -> gradients in the reference's parameter layout
 * The layout is public so that parity tests can read every intermediate of the backward out of the workspace. */
#define NF_TRAIN_LIN_MAX 24
typedef struct nf_train_lin {  /* one Linear, in execution order (density MLP, then the RGB head) */
  int32_t m, j, n, n_pad, k0_pad, k_hidden, act, x0_raw;   /* (mlp, index), out features (padded to 16), x0 / hidden input columns */
  int64_t a_off, a_tile;       /* input operand stash: fp16 [tile][(k0_pad + k_hidden)/8][128][8], K order [x0 | hidden]; bytes per tile */
  int64_t c_off;               /* cos(pre-activation) of the hidden input, fp16 [tile][32][128][8]; -1 unless the MLP is sin-activated */
  int64_t g_off, g_tile;       /* d loss / d output of the Linear (loss-scaled), fp16 [tile][n_pad/8][128][8] */
  int64_t dw_off, db_off;      /* fp32 dWt[n_pad][k0_pad + k_hidden], db[n_pad] (tensor column orders, loss-scaled) */
} nf_train_lin;
typedef struct nf_train_layout {
  int32_t n_lin, T, rpu, tpr;          /* a work unit = rpu rays = tpr tiles of 128 samples (sample-stream tiling) */
  int64_t n_rays, n_tiles;
  int64_t sigma_off, rgbraw_off;       /* fp32 [R,T], [R,T,3]: raw density (noise included), raw colours */
  int64_t dsigma_off, drgbraw_off;     /* fp32 gradients of those (written by the backward) */
  int64_t dx0_off;                     /* fp32 [n_tiles*128][32]: gradient of the density MLP's hash features (Positional head: followed
                                          by a second block of the same shape, the head's own encoder's) */
  int64_t scale_off;                   /* float[4]: loss scale S, 1/S, max|g| (bits), - */
  int64_t dw_begin, dw_end;
  int64_t total_bytes;                 /* = the workspace size */
  int64_t bgrand_off;                  /* NF_BG_RANDOM: fp32 [R], the forward's background draws (nf_render_aux.bg_rand) kept for the backward; else -1 */
  nf_train_lin lin[NF_TRAIN_LIN_MAX];
} nf_train_layout;
int nf_train_layout_of(const nf_model_desc* desc, int64_t n_rays, int32_t T, nf_train_layout* out);

/* Optional inputs / outputs of the forward (all nullable; struct_bytes = sizeof(nf_render_aux)). */
typedef struct nf_render_aux {
  int32_t struct_bytes, reserved;
  void* train_ws;            /* training forward: workspace of nf_train_layout_of(desc, n_rays, T).total_bytes bytes, 1024-aligned */
  int64_t train_ws_bytes;
  const float* pts;          /* from_pts (reference src/nerf.py:340-361; called by DynamicNeRF 1303, render_keyframes 1317): explicit
                                sample positions [R,T,3] (ray-major) instead of r_o + ts r_d; ts still gives the segment lengths,
                                rays the origin / view direction.  Not with Mip, NF_KIND_DYN or a training workspace. */
  const float* bg_rand;      /* NF_BG_RANDOM: uniform draws [R] (the reference's rand_like(summed), nerf.py:101-102) */
  /* NF_KIND_DYN side channels the runner's regularisers and visualisations read after forward (runner.py:523-531,694-700,769,
   * 777-781), all [R,T,C] ray-major (the reference keeps [T,B,H,W,C]): */
  float* pts_out;            /* C = 3: the UNdeformed sample positions (model.pts) */
  float* dp_out;             /* direct: C = 1, spline: C = 3 (model.dp; the reference's direct split names the 1-channel output dp) */
  float* rigid_dp_out;       /* C = 3 (model.rigid_dp = dp * rigidity) */
  float* rigidity_out;       /* direct: C = 3, spline: C = 1 (model.rigidity = sigmoid(raw / 2)) */
} nf_render_aux;
/* nf_render_forward with the optional extras of nf_render_aux (aux == NULL: identical to nf_render_forward). */
int nf_render_forward_aux(const nf_model_desc* desc, const void* packed,
                          const float* rays, int64_t n_rays,
                          const float* ts, int32_t T, int64_t ts_ray_stride,
                          const float* density_noise, const float* ray_time, const nf_mip_args* mip,
                          float* rgb_out, float* alpha_out, float* weights_out,
                          const nf_render_aux* aux, int32_t precision, void* stream);
/* Backward of the render that filled `train_ws` (same desc, packed, rays, ts): d_rgb[R,3] -> one gradient per parameter, in
 * the order of nf_pack_weights (grads_host[i] == NULL skips parameter i; every other gradient buffer is OVERWRITTEN, shapes as
 * the parameters; VolSDF: the last one is the scalar d loss / d beta).  Replaces loss.backward() through PlainNeRF.forward /
 * VolSDF.forward (reference runner.py:820, src/nerf.py:326-361, 981-1013). */
int nf_render_backward(const nf_model_desc* desc, const void* packed, void* train_ws, int64_t train_ws_bytes,
                       const float* rays, int64_t n_rays, const float* ts, int32_t T, int64_t ts_ray_stride,
                       const float* d_rgb, float* const* grads_host, int32_t n_grads, void* stream);

/* ---- the SDF surface side --------------------------------------------------------------------------
 * `desc` = a VolSDF-style descriptor (describe_volsdf): density MLP = the SDF network (SIREN: NF_ENC_NONE, or the Fourier-
 * encoded MLP), refl = the View head.  All kernels of a call are stream-ordered, with no host round trip: the set of active rays
 * is compacted on the device and the MLP kernels (CUDA-core fp32, or tcgen05 for the SIREN network) read its size from there. */
int64_t nf_sdf_workspace_bytes(const nf_model_desc* desc, int64_t n_rays);
/* sphere_march (reference src/march.py:27-47): t = near; `iters` times, for the rays still active: d = sdf(o + t dir)
 * [max(d, |p| - bound_rad) with bound_rad > 0: UnitSphere, src/sdf.py:66-83]; hit |= d < eps && t <= far; t += d; a ray leaves
 * the active set once hit or t > far.  Outputs (nullable): pts_out[R,3] = o + t dir, hit_out[R] (0/1 bytes), t_out[R]. */
int nf_sphere_march(const nf_model_desc* desc, const void* packed, const float* rays, int64_t n_rays, float t_near, float t_far,
                    int32_t iters, float eps, float bound_rad, int32_t precision, float* pts_out, uint8_t* hit_out, float* t_out,
                    void* workspace, int64_t workspace_bytes, void* stream);
/* SDF.forward in eval mode (reference src/sdf.py:137-156): sphere_march, latent = sdf_net(pts[hit])[1:], rgb[hit] =
 * act(View([pts, elaz(r_d), latent])), rgb[~hit] = 0.  rgb_out[R,3]; the other outputs as above (nullable). */
int nf_sdf_render(const nf_model_desc* desc, const void* packed, const float* rays, int64_t n_rays, float t_near, float t_far,
                  int32_t iters, float eps, float bound_rad, int32_t precision, float* rgb_out, uint8_t* hit_out, float* t_out,
                  float* pts_out, void* workspace, int64_t workspace_bytes, void* stream);
/* march.bisect behind SDF.forward (reference src/march.py:63-75, `--sdf-isect-kind bisect`): throughput_with_sign_change over
 * iters + 1 equidistant samples of EVERY ray (march.py:78-110, bug for bug: the first sample at r_o + t_near added to every coordinate,
 * the bracket indices scaled by the step without the near offset; `jitter` in [0, 1) = its random.random() draw, march.py:86), then
 * min(32, iters) bisection steps between the samples around the first sign change (march.py:147-180, eps 1e-6).
 * Outputs (nullable): pts_out[R,3] (the bisection's point), hit_out[R] = tput < 0, tput_out[R] (the SDF at the sample where it is
 * smallest), best_pos_out[R,3] (that sample); rgb_out[R,3] non-null: also SDF.forward's shading of the hits (sdf.py:143-153). */
int nf_sdf_bisect(const nf_model_desc* desc, const void* packed, const float* rays, int64_t n_rays, float t_near, float t_far,
                  int32_t iters, float jitter, float bound_rad, int32_t precision, float* pts_out, uint8_t* hit_out, float* tput_out,
                  float* best_pos_out, float* rgb_out, void* workspace, int64_t workspace_bytes, void* stream);
/* SDFModel.normals (reference src/sdf.py:43-49; SDF.normals / intersect_w_n, sdf.py:112,114-125): the gradient with respect to the
 * point of the SUM OF ALL outputs of the SDF network -- utils.autograd (utils.py:266-277) back-propagates ones over every channel,
 * the latent included -- by forward-mode differentiation on the fp32 CUDA-core pipeline (value + three tangents per point through
 * the same Linears; act' at the value's pre-activation).  pts[N,3] -> normals_out[N,3]; values_out[N, 1 + I] (nullable) = the
 * network's outputs at the points.  bound_rad > 0: UnitSphere (sdf.py:66-83), output 0 = max(inner, |p| - rad). */
int nf_sdf_normals(const nf_model_desc* desc, const void* packed, const float* pts, int64_t n, float bound_rad,
                   float* normals_out, float* values_out, void* stream);

/* ---- backward of the non-GEMM stages (first blocks of the training half; the reference differentiates these ops through
 *      PyTorch autograd, runner.py:820) --------------------------------------------------------------------------- */
/* Backward of nf_composite: d_rgb[R,3] -> d_sigma_raw_out[R,T], d_feats_out[R,T,3] (same inputs as the forward; T <= 2048).
 * The gradient with respect to VolSDF's beta is not produced by this stand-alone stage (nf_render_backward produces it). */
int nf_composite_backward(const nf_model_desc* desc, const void* packed, const float* sigma_raw, const float* feats,
                          const float* rays, int64_t n_rays, const float* ts, int32_t T, int64_t ts_ray_stride,
                          const float* d_rgb, float* d_sigma_raw_out, float* d_feats_out, void* stream);
/* Backward of nf_hash_encode with respect to the tables: d_feats[N, levels*feat] is scatter-ADDED (trilinear weights, float4
 * atomics) into d_tables[levels][table][feat], which the caller zeroes (or accumulates into across micro-batches). */
int nf_hash_encode_backward(const nf_model_desc* desc, const float* pts, int64_t n, const float* d_feats, float* d_tables,
                            void* stream);

/* One torch.optim.Adam step on one fp32 tensor of n elements, in place (the reference's optimiser, runner.py:448-458: Adam,
 * eps 1e-7, L2 weight_decay added to the gradient; the cosine learning-rate schedule of runner.py:1289 stays on the host and
 * arrives as `lr`).  step = 1 for the first update.  All pointers 16-byte aligned device pointers. */
int nf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                 float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream);

/* The same step for n_tensors tensors that share the hyper-parameters and the step count (one parameter group of
 * runner.load_optim's Adam, runner.py:448-458) in ONE launch: params / grads / exp_avg / exp_avg_sq are HOST arrays of device
 * pointers (16-byte aligned), numel a host array of element counts (0 = skip). */
int nf_adam_step_multi(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                       float* const* exp_avg_sq, const int64_t* numel, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int32_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H */
