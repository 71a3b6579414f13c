// nf_common.cuh -- layout plan of the packed parameter blob + device helpers shared by
// the fp32 and tcgen05 pipelines.  Reference file:line citations are to JulianKnodt/nerf_atlas.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/nerf_b200.h"

#define NF_MAX_LIN 10          // init + up to 8 hidden + out
#define NF_HIDDEN 256
#define NF_TC_ROWS 128         // samples per tensor-core tile (= UMMA M)
#define NF_TC_CHUNK_K 64       // K columns per streamed weight chunk (4 UMMA K-steps; 32 KB at N=256 -- see profiles/microbench)
#define NF_MIP_FEATS 96        // Mip IPE: 2 (sin, cos) x 16 degrees x 3 axes (reference src/utils.py:104-111, src/nerf.py:255)

// ---- plan ---------------------------------------------------------------------
struct NfLinPlan {
  int32_t k_hidden;   // inputs fed from the previous layer's (activated) output
  int32_t k_x0;       // inputs fed from x0 (raw for `init`, activated for skip layers)
  int32_t k0_pad;     // k_x0 rounded up to 16 (tensor path)
  int32_t n;          // out features
  int32_t n_pad;      // rounded up to 16
  int32_t x0_raw;     // 1: x0 is consumed without activation (the `init` Linear)
  int32_t is_out;     // 1: the `out` Linear (result is NOT activated)
  int32_t n_chunks;   // tensor path: number of streamed weight chunks
  int64_t wt_off;     // byte offset: fp32 Wt[k_hidden+k_x0][n_pad], k order = reference column order
  int64_t b_off;      // byte offset: fp32 bias[n_pad]
  int64_t w16_off;    // byte offset: fp16 UMMA-canonical image [K_tc/8][n_pad][8], K order = [x0 (k0_pad) | hidden]
  int64_t b16_off;    // byte offset: fp32 bias[n_pad] in tensor-path column order
  int64_t w16h_off;   // byte offset: the same image split for a CTA pair, plus one K-step (16 rows) carrying the bias as fp16 hi / lo
                      // halves (rows K_tc, K_tc + 1; the rest zero): [rank 0..1][(K_tc + 16)/8][n_pad/2][8]
  int64_t w16t_off;   // byte offset: the TRANSPOSED fp16 images of the backward (dX = dZ W): [n_pad/8][k0_pad][8] (x0 part, if any)
                      // followed by [n_pad/8][256][8] (hidden part, if any): the reduction dimension is the Linear's OUTPUT
};
struct NfMlpPlan {
  int32_t n_lin, in_dims, k0_pad, act, out_dims, hidden_ref;   // hidden_ref: the reference's hidden_size (< 256: zero-padded to 256)
  NfLinPlan lin[NF_MAX_LIN];
};
struct NfPlan {
  int32_t kind, n_mlps, intermediate, enc;
  int32_t hash_levels, hash_mask, density_act, feat_act;
  int32_t bg, pad_;
  uint32_t hash_primes[3]; uint32_t pad2_;
  float hash_res[16];
  int64_t hash_off;     // byte offset: fp32 [levels][table][4]
  int64_t fourier_off;  // byte offset: fp32 basis [3][freqs]
  int64_t scale_off;    // byte offset: fp32 beta (VolSDF.scale)
  int32_t fourier_freqs, mip;
  int32_t deform_enc, spline_points;
  int32_t refl_kind, pad4_;
  int64_t hash3_off;    // byte offset: fp32 [levels][table][4], tables of the Positional head's own HashEncoder
  int64_t hash2_off;    // byte offset: fp32 [levels][table][4], tables of the spline deformation MLP's own HashEncoder
  int64_t total_bytes;
  NfMlpPlan mlp[3];     // [0] density, [1] refl, [2] deformation (NF_KIND_DYN; executed first) or PosLinearView.view (NF_REFL_POSLINVIEW)
};

__host__ __device__ inline int nf_round_up(int x, int m) { return (x + m - 1) / m * m; }

// Builds the plan; returns 0 or an NF_E_* code with *why set.
static inline int nf_build_plan(const nf_model_desc* d, NfPlan* p, const char** why) {
  *why = "";
  if (!d || d->struct_bytes != (int32_t)sizeof(nf_model_desc)) { *why = "bad nf_model_desc (struct_bytes)"; return NF_E_BADARG; }
  if (d->kind != NF_KIND_PLAIN && d->kind != NF_KIND_TINY && d->kind != NF_KIND_DYN) { *why = "unsupported model kind"; return NF_E_UNSUPPORTED; }
  *p = NfPlan{};
  p->kind = d->kind; p->n_mlps = d->kind == NF_KIND_DYN ? 3 : d->kind == NF_KIND_PLAIN ? 2 : 1;
  p->intermediate = d->intermediate; p->enc = d->enc;
  p->density_act = d->density_act; p->feat_act = d->feat_act; p->bg = d->bg;
  if (d->mip < NF_MIP_NONE || d->mip > NF_MIP_CYLINDER_REF) { *why = "unknown mip kind"; return NF_E_BADARG; }
  if (d->mip != NF_MIP_NONE && d->kind == NF_KIND_TINY) { *why = "mip: not for NF_KIND_TINY"; return NF_E_UNSUPPORTED; }
  p->mip = d->mip;
  if (d->refl_kind < NF_REFL_VIEW || d->refl_kind > NF_REFL_POSLINVIEW) { *why = "unknown refl kind"; return NF_E_BADARG; }
  if (d->refl_kind != NF_REFL_VIEW && (d->kind == NF_KIND_TINY || d->enc != NF_ENC_HASH)) {
    *why = "the Positional / PosLinearView heads need a hash-encoded PlainNeRF / DynamicNeRF"; return NF_E_UNSUPPORTED; }
  if (d->refl_kind == NF_REFL_POSLINVIEW && (d->kind != NF_KIND_PLAIN || d->mip != NF_MIP_NONE || d->feat_act == NF_FEAT_SOFTMAX)) {
    *why = "PosLinearView: PlainNeRF without Mip, any feature activation but softmax"; return NF_E_UNSUPPORTED; }
  if (d->refl_kind == NF_REFL_POSLINVIEW) p->n_mlps = 3;
  p->refl_kind = d->refl_kind;
  if (d->kind == NF_KIND_DYN) {
    p->deform_enc = d->deform_enc; p->spline_points = d->spline_points;
    if (d->spline_points == 0 ? d->deform_enc != NF_ENC_NONE : (d->deform_enc != NF_ENC_HASH || d->enc != NF_ENC_HASH)) {
      *why = "dyn: direct deformation takes no encoder, the spline variant takes the hash encoder"; return NF_E_BADARG; }
    if (d->spline_points != 0 && (d->spline_points < 2 || d->spline_points > 8)) { *why = "dyn: spline_points must be 0 or 2..8"; return NF_E_UNSUPPORTED; }
  }
  if (d->enc == NF_ENC_HASH) {
    if (d->hash_feat != 4 || d->hash_levels < 1 || d->hash_levels > 16 ||
        (d->hash_table_size & (d->hash_table_size - 1)) != 0 || d->hash_table_size < 2) {
      *why = "hash encoder: need feat==4, 1..16 levels, power-of-two table"; return NF_E_UNSUPPORTED; }
    p->hash_levels = d->hash_levels; p->hash_mask = d->hash_table_size - 1;
    for (int i = 0; i < 3; ++i) p->hash_primes[i] = d->hash_primes[i];
    for (int i = 0; i < 16; ++i) p->hash_res[i] = d->hash_res[i];
  } else if (d->enc == NF_ENC_FOURIER) {
    if (d->fourier_freqs < 1 || d->fourier_freqs > 128 || (d->fourier_freqs & 3)) { *why = "fourier encoder: need freqs in 4..128, multiple of 4"; return NF_E_UNSUPPORTED; }
    p->fourier_freqs = d->fourier_freqs;
  } else if (d->enc != NF_ENC_NONE) { *why = "unsupported encoder"; return NF_E_UNSUPPORTED; }
  if (d->density_act < 0 || d->density_act > NF_DENS_LAPLACE) { *why = "unknown density activation"; return NF_E_BADARG; }
  if (d->feat_act < 0 || d->feat_act > NF_FEAT_SOFTMAX) { *why = "unknown feature activation (enum nf_feat_act)"; return NF_E_BADARG; }
  if (d->bg < 0 || d->bg > NF_BG_RANDOM) { *why = "unknown background (enum nf_bg)"; return NF_E_BADARG; }
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
  if (d->enc == NF_ENC_HASH) p->hash_off = take((int64_t)d->hash_levels * d->hash_table_size * 4 * sizeof(float));
  if (d->enc == NF_ENC_FOURIER) p->fourier_off = take((int64_t)3 * d->fourier_freqs * sizeof(float));
  if (d->kind == NF_KIND_DYN && d->deform_enc == NF_ENC_HASH) p->hash2_off = take((int64_t)d->hash_levels * d->hash_table_size * 4 * sizeof(float));
  if (d->refl_kind != NF_REFL_VIEW) p->hash3_off = take((int64_t)d->hash_levels * d->hash_table_size * 4 * sizeof(float));
  p->scale_off = take(sizeof(float));
  for (int m = 0; m < p->n_mlps; ++m) {
    const nf_mlp_desc& md = m == 0 ? d->density : m == 1 ? d->refl : d->refl_kind == NF_REFL_POSLINVIEW ? d->refl_view : d->deform;
    NfMlpPlan& mp = p->mlp[m];
    if (md.hidden != NF_HIDDEN && !(d->refl_kind == NF_REFL_POSLINVIEW && m == 2 && md.hidden >= 16 && md.hidden < NF_HIDDEN && (md.hidden & 15) == 0)) {
      *why = "hidden_size must be 256 (PosLinearView.view: a multiple of 16 below 256, zero-padded)"; return NF_E_UNSUPPORTED; }
    mp.hidden_ref = md.hidden;
    if (md.n_layers < 1 || md.n_layers + 2 > NF_MAX_LIN) { *why = "unsupported number of layers"; return NF_E_UNSUPPORTED; }
    if (md.in_dims < 1 || md.in_dims > 272 || md.out_dims < 1 || md.out_dims > 256 || md.skip < 1) { *why = "unsupported MLP dims"; return NF_E_UNSUPPORTED; }
    mp.n_lin = md.n_layers + 2; mp.in_dims = md.in_dims; mp.k0_pad = nf_round_up(md.in_dims, 16);
    mp.act = md.act; mp.out_dims = md.out_dims;
    for (int j = 0; j < mp.n_lin; ++j) {
      NfLinPlan& L = mp.lin[j];
      const bool is_init = j == 0, is_out = j == mp.n_lin - 1;
      const int i = j - 1;  // index into SkipConnMLP.layers
      const bool skip = !is_init && !is_out && i != md.n_layers - 1 && (i % md.skip) == 0;
      L.k_hidden = is_init ? 0 : NF_HIDDEN;
      L.k_x0 = (is_init || skip) ? md.in_dims : 0;
      L.k0_pad = L.k_x0 ? mp.k0_pad : 0;
      L.n = is_out ? md.out_dims : NF_HIDDEN; L.n_pad = nf_round_up(L.n, 16);
      L.x0_raw = is_init; L.is_out = is_out;
      const int k_tc = L.k0_pad + L.k_hidden;
      L.n_chunks = (k_tc + NF_TC_CHUNK_K - 1) / NF_TC_CHUNK_K;
      L.wt_off = take((int64_t)(L.k_hidden + L.k_x0) * L.n_pad * sizeof(float));
      L.b_off = take((int64_t)L.n_pad * sizeof(float));
      L.w16_off = take((int64_t)k_tc * L.n_pad * sizeof(__half));
      L.b16_off = take((int64_t)L.n_pad * sizeof(float));
      L.w16h_off = take((int64_t)(k_tc + 16) * L.n_pad * sizeof(__half));
      L.w16t_off = take((int64_t)k_tc * L.n_pad * sizeof(__half));
    }
  }
  if (d->kind == NF_KIND_DYN) {
    if (d->spline_points == 0 && (d->deform.in_dims != 4 || d->deform.out_dims != 4)) { *why = "dyn: direct deformation MLP must map 4 -> 4"; return NF_E_BADARG; }
    if (d->spline_points != 0 && (d->deform.in_dims != 6 + d->hash_levels * 4 || d->deform.out_dims != 1 + 3 * d->spline_points)) {
      *why = "dyn: spline deformation MLP must map [p, hash(p)] -> 1 + 3n"; return NF_E_BADARG; }
  }
  if (d->kind == NF_KIND_PLAIN || d->kind == NF_KIND_DYN) {
    const int ml = d->mip != NF_MIP_NONE ? NF_MIP_FEATS : 0;
    if (d->density.out_dims != 1 + d->intermediate) { *why = "density MLP out must be 1+intermediate"; return NF_E_BADARG; }
    const int rin = (d->refl_kind != NF_REFL_VIEW ? 6 + d->hash_levels * 4 : 5) + ml + d->intermediate;
    if (d->refl_kind == NF_REFL_POSLINVIEW) {
      const int im = d->refl.out_dims - 3;                   // PosLinearView's own intermediate (refl.py:249,253)
      if (d->refl.in_dims != rin || im < 1 || d->refl_view.in_dims != 6 + d->intermediate + im || d->refl_view.out_dims != 1) {
        *why = "PosLinearView: pos MLP [p, p, hash'(p), latent] -> 3 + im, view MLP [p, dir, latent, im] -> 1"; return NF_E_BADARG; }
    } else
    if (d->refl.in_dims != rin || d->refl.out_dims != 3) { *why = "refl MLP must map [p, elaz | p, hash(p)] (+96 mip) + intermediate -> 3"; return NF_E_BADARG; }
    const int want = (d->enc == NF_ENC_HASH ? 6 + d->hash_levels * 4 : d->enc == NF_ENC_FOURIER ? 3 + 2 * d->fourier_freqs : 3) + ml;
    if (d->density.in_dims != want) { *why = "density MLP in_dims does not match the encoder"; return NF_E_BADARG; }
  } else {
    if (d->density.in_dims != 3 || d->density.out_dims != 4 || d->enc != NF_ENC_NONE) { *why = "tiny: density MLP must map 3 -> 4 without encoder"; return NF_E_BADARG; }
  }
  p->total_bytes = off;
  return 0;
}

// ---- tensor-path column orders ---------------------------------------------------------------
// The tensor path permutes x0 columns (and the density MLP's output columns) so that every
// producer writes whole 16-byte groups of 8 halves:
//   density x0 (hash): reference [p(3), p(3), feats(4L)]      -> [feats(4L), p, p]
//   refl x0          : reference [p(3), elaz(2), inter(I)]    -> [inter(I), p, elaz]
//   density out      : reference [sigma, inter(I)]            -> [inter(I), sigma]
// With a Mip latent (96 columns) and/or the Positional head the wide parts keep 8-column alignment:
//   density x0 + mip : reference [p, p, feats(4L), mip(96)]            -> [feats(4L), p, p, pad(2), mip(96)]
//   View x0 + mip    : reference [p, elaz, mip(96), inter(I)]          -> [inter(I), p, elaz, pad(3), mip(96)]
//   Positional x0    : reference [p, p, feats'(4L), (mip), inter(I)]   -> [inter(I), feats'(4L), p, p, pad(2), (mip)]
__host__ __device__ inline int nf_mip_col(const NfPlan& p, int m) {       // first tensor-order column of the mip latent in MLP m's x0
  const int nfe = p.hash_levels * 4;
  if (m == 0) return nf_round_up(6 + nfe, 8);
  return p.refl_kind == NF_REFL_POSITIONAL ? p.intermediate + nfe + 8 : p.intermediate + 8;
}
__host__ __device__ inline int nf_x0_perm(const NfPlan& p, int m, int k_ref) {
  const int ml = p.mip != NF_MIP_NONE ? NF_MIP_FEATS : 0;
  const int nfe = p.hash_levels * 4;
  if (m == 2 && p.deform_enc == NF_ENC_HASH) return k_ref < 6 ? nfe + k_ref : k_ref - 6;
  if (p.kind == NF_KIND_PLAIN || p.kind == NF_KIND_DYN) {
    if (m == 0 && p.enc == NF_ENC_HASH) {
      if (k_ref >= 6 + nfe) return nf_mip_col(p, 0) + (k_ref - 6 - nfe);
      return k_ref < 6 ? nfe + k_ref : k_ref - 6;
    }
    if (m == 1 && p.refl_kind == NF_REFL_POSITIONAL) {
      if (k_ref < 6) return p.intermediate + nfe + k_ref;
      if (k_ref < 6 + nfe) return p.intermediate + (k_ref - 6);
      if (k_ref < 6 + nfe + ml) return nf_mip_col(p, 1) + (k_ref - 6 - nfe);
      return k_ref - (6 + nfe + ml);
    }
    if (m == 1) {
      if (k_ref < 5) return p.intermediate + k_ref;
      if (k_ref < 5 + ml) return nf_mip_col(p, 1) + (k_ref - 5);
      return k_ref - 5 - ml;
    }
  }
  return k_ref;
}
__host__ __device__ inline int nf_out_perm(const NfPlan& p, int m, int n_ref) {
  if ((p.kind == NF_KIND_PLAIN || p.kind == NF_KIND_DYN) && m == 0) return n_ref == 0 ? p.intermediate : n_ref - 1;
  return n_ref;
}

// ---- training workspace (nf_render_forward_aux with a train workspace -> nf_render_backward) --------------------------
// The tensor pipeline's training forward stashes, per 128-sample tile and per Linear (execution order), the INPUT operand of
// the Linear exactly as the MMA consumed it (fp16, UMMA canonical K-major image [K/8][128][8], K order = [x0 | hidden]) and,
// for sin-activated MLPs, the cosine of the pre-activation behind every hidden input.  The backward adds the gradient with
// respect to every Linear's output (G, same image, loss-scaled fp16).  dW = G^T A then reads both stashes as MN-major operands.
#define NF_TRAIN_MAX_LIN NF_TRAIN_LIN_MAX
typedef nf_train_lin NfTrainLin;          // the layout is public (include/nerf_b200.h): parity tests decode the stash
typedef nf_train_layout NfTrainPlan;

#ifdef __CUDACC__
// ---- activations --------------------------------------------------------------
__device__ __forceinline__ float nf_apply_act(float x, int act) {
  switch (act) {
    case NF_ACT_LEAKY: return x > 0.f ? x : 0.01f * x;
    case NF_ACT_SIN:   return sinf(x);
    case NF_ACT_RELU:  return fmaxf(x, 0.f);
    default:           return x;
  }
}
// d act / d x at the pre-activation x (torch's conventions at the kinks: LeakyReLU' = ReLU' = the negative-side slope at x <= 0)
__device__ __forceinline__ float nf_act_grad(float x, int act) {
  switch (act) {
    case NF_ACT_LEAKY: return x > 0.f ? 1.f : 0.01f;
    case NF_ACT_SIN:   return cosf(x);
    case NF_ACT_RELU:  return x > 0.f ? 1.f : 0.f;
    default:           return 1.f;
  }
}
__device__ __forceinline__ float nf_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float nf_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }
// sigmoid family, reference src/utils.py:484-518
__device__ __forceinline__ float nf_feat_act_fn(float v, int kind) {
  switch (kind) {
    case NF_FEAT_NORMAL:    return nf_sigmoid(v);
    case NF_FEAT_THIN:      return nf_sigmoid(v) * (1.f - 2e-2f) + 1e-2f + 1e-2f;  // fat(v,-eps)+eps
    case NF_FEAT_TANH:      return tanhf(v);
    case NF_FEAT_CYCLIC:    return (sinf(v / 5.f) + 1.f) / 2.f * (1.f - 2e-2f) + 1e-2f;
    case NF_FEAT_UPSHIFTED: return nf_sigmoid(v) + 1e-2f;
    case NF_FEAT_FAT:       return nf_sigmoid(v) * (1.f + 2e-2f) - 1e-2f;
    case NF_FEAT_LEAKY_RELU:return v > 0.f ? v : 0.01f * v;
    case NF_FEAT_RELU:      return fmaxf(v, 0.f);
    case NF_FEAT_SIN:       return sinf(v);
    case NF_FEAT_UPSHIFTED_SOFTPLUS: return nf_softplus(v) + 1e-2f;
    case NF_FEAT_UPSHIFTED_RELU:     return fmaxf(v, 0.f) + 1e-2f;
    default: return v;      // unreachable: nf_build_plan rejects unknown kinds; NF_FEAT_SOFTMAX goes through nf_feat_act3
  }
}
// the three colour channels of one sample; nn.Softmax(dim=-1) (reference src/utils.py:507) couples them: exp(v - max) / sum in
// torch's order (max-subtracted exponentials, one division per channel)
__device__ __forceinline__ void nf_feat_act3(float& r, float& g, float& b, int kind) {
  if (kind == NF_FEAT_SOFTMAX) {
    const float m = fmaxf(r, fmaxf(g, b));
    const float er = expf(r - m), eg = expf(g - m), eb = expf(b - m);
    const float s = er + eg + eb;
    r = er / s; g = eg / s; b = eb / s;
    return;
  }
  r = nf_feat_act_fn(r, kind); g = nf_feat_act_fn(g, kind); b = nf_feat_act_fn(b, kind);
}
// raw density -> sigma, reference src/nerf.py:64-65
// `beta` is only used by NF_DENS_LAPLACE (VolSDF): sigma = relu(Psi_beta(-sdf) / beta), reference src/nerf.py:1000-1003,
// src/utils.py:50-58 (the raw MLP output is the signed distance).
__device__ __forceinline__ float nf_density_act_fn(float d, int kind, float beta = 1.f) {
  if (kind == NF_DENS_RELU) return fmaxf(d, 0.f);
  if (kind == NF_DENS_LAPLACE) {
    const float sc = (-d) / beta;
    const float cdf = sc <= 0.f ? expf(fminf(sc, 0.f)) / 2.f : 1.f - expf(-fmaxf(sc, 0.f)) / 2.f;
    return fmaxf(1.f / beta * cdf, 0.f);
  }
  return nf_softplus(d - 1.f);
}

// ---- sample position, reference src/nerf.py:54: rounded product, then rounded add (no FMA) ----
__device__ __forceinline__ float nf_pt(float o, float t, float d) { return __fadd_rn(o, __fmul_rn(t, d)); }

// ---- view direction -> (elev, azim), reference src/utils.py:247-254 ------------------------
__device__ __forceinline__ void nf_elaz(float dx, float dy, float dz, float& elev, float& azim) {
  const float nrm = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);  // F.normalize eps
  const float lim = 1.f - 1e-6f;
  const float x = fminf(fmaxf(dx / nrm, -lim), lim), y = fminf(fmaxf(dy / nrm, -lim), lim),
              z = fminf(fmaxf(dz / nrm, -lim), lim);
  elev = acosf(z); azim = atan2f(y, x);
}

// ---- hash grid, reference src/neural_blocks.py:139-193 -------------------------------------
// One level for one point. Corner order = (bx,by,bz) with z the fastest bit (lines 155-165).
// idx8 (nullable) receives the 8 table rows. Products are rounded separately (no FMA) and the
// 8 corner terms are summed in corner order, like the reference's stack(...).sum(dim=0).
__device__ __forceinline__ float4 nf_hash_level(const float4* __restrict__ table, float px, float py, float pz,
                                                float res, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t mask,
                                                uint32_t* idx8) {
  const float vx = __fmul_rn(px, res), vy = __fmul_rn(py, res), vz = __fmul_rn(pz, res);
  const float fx = floorf(vx), fy = floorf(vy), fz = floorf(vz);
  const uint32_t ix = (uint32_t)(int32_t)fx, iy = (uint32_t)(int32_t)fy, iz = (uint32_t)(int32_t)fz;
  const float wx = __fsub_rn(vx, fx), wy = __fsub_rn(vy, fy), wz = __fsub_rn(vz, fz);
  const float ux = __fsub_rn(1.f, wx), uy = __fsub_rn(1.f, wy), uz = __fsub_rn(1.f, wz);
  const uint32_t hx0 = ix * p0, hx1 = (ix + 1u) * p0, hy0 = iy * p1, hy1 = (iy + 1u) * p1,
                 hz0 = iz * p2, hz1 = (iz + 1u) * p2;
  uint32_t id[8];
  float w[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int bx = (c >> 2) & 1, by = (c >> 1) & 1, bz = c & 1;
    id[c] = ((bx ? hx1 : hx0) ^ (by ? hy1 : hy0) ^ (bz ? hz1 : hz0)) & mask;
    w[c] = __fmul_rn(__fmul_rn(bx ? wx : ux, by ? wy : uy), bz ? wz : uz);
  }
  float4 e[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) e[c] = __ldg(table + id[c]);
  float4 s;
  s.x = __fmul_rn(e[0].x, w[0]); s.y = __fmul_rn(e[0].y, w[0]); s.z = __fmul_rn(e[0].z, w[0]); s.w = __fmul_rn(e[0].w, w[0]);
#pragma unroll
  for (int c = 1; c < 8; ++c) {
    s.x = __fadd_rn(s.x, __fmul_rn(e[c].x, w[c])); s.y = __fadd_rn(s.y, __fmul_rn(e[c].y, w[c]));
    s.z = __fadd_rn(s.z, __fmul_rn(e[c].z, w[c])); s.w = __fadd_rn(s.w, __fmul_rn(e[c].w, w[c]));
  }
  if (idx8) {
#pragma unroll
    for (int c = 0; c < 8; ++c) idx8[c] = id[c];
  }
  return s;
}

// ---- Mip IPE, reference src/nerf.py:257-261, src/utils.py:22-27,39-48,60-101 ---------------------------------
struct NfMipIn {
  int mode;                    // enum nf_mip
  const float* ts; int T;      // shared ts[T]
  const float* rays; const float* radius;            // this call's rays [R,6] / radii [R]
  const float* rays_all; const float* radius_all;    // NF_MIP_CYLINDER_REF: the whole crop
  long long n_rays_all, ray_base;
};
// segment [t0, t1] of sample t: the reference appends 1e10 (nerf.py:258); the intended encoder repeats the last spacing
__device__ __forceinline__ void nf_mip_segment(const NfMipIn& m, int t, float& t0, float& t1) {
  t0 = __ldg(m.ts + t);
  if (t + 1 < m.T) t1 = __ldg(m.ts + t + 1);
  else if (m.mode == NF_MIP_CYLINDER_REF) t1 = 1e10f;
  else t1 = m.T > 1 ? __fadd_rn(t0, __fsub_rn(t0, __ldg(m.ts + t - 1))) : t0;
}
// (t_mean, t_var, r_var) of a segment: cylinder utils.py:95-101, cone utils.py:83-93 (operation order as written there)
__device__ __forceinline__ void nf_mip_moments(int mode, float t0, float t1, float rad, float& t_mean, float& t_var, float& r_var) {
  if (mode == NF_MIP_CONE) {
    const float mu = __fdiv_rn(__fadd_rn(t1, t0), 2.f), hw = __fdiv_rn(__fsub_rn(t1, t0), 2.f);
    const float mu2 = __fmul_rn(mu, mu), hw2 = __fmul_rn(hw, hw), hw4 = __fmul_rn(hw2, hw2);
    const float den = __fadd_rn(__fmul_rn(3.f, mu2), hw2);
    t_mean = __fadd_rn(mu, __fdiv_rn(__fmul_rn(__fmul_rn(2.f, mu), hw2), den));
    t_var = __fsub_rn(__fdiv_rn(hw, 3.f), __fmul_rn(4.f / 15.f, __fdiv_rn(__fmul_rn(hw4, __fsub_rn(__fmul_rn(12.f, mu2), hw2)), __fmul_rn(den, den))));
    r_var = __fmul_rn(__fmul_rn(rad, rad), __fsub_rn(__fadd_rn(__fdiv_rn(mu2, 4.f), __fmul_rn(5.f / 12.f, hw2)), __fdiv_rn(__fmul_rn(4.f / 15.f, hw4), den)));
  } else {
    t_mean = __fdiv_rn(__fadd_rn(t1, t0), 2.f);
    const float d = __fsub_rn(t1, t0);
    t_var = __fdiv_rn(__fmul_rn(d, d), 12.f);
    r_var = __fdiv_rn(__fmul_rn(rad, rad), 4.f);
  }
}
// diagonal covariance entry `x` of the lifted Gaussian (lift_gaussian, utils.py:60-73)
__device__ __forceinline__ float nf_mip_cov(const float* __restrict__ ray6, int x, float t_var, float r_var) {
  const float dx = __ldg(ray6 + 3), dy = __ldg(ray6 + 4), dz = __ldg(ray6 + 5);
  const float magn = fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)), 1e-10f);
  const float d = x == 0 ? dx : x == 1 ? dy : dz;
  const float od = __fmul_rn(d, d);
  return __fadd_rn(__fmul_rn(t_var, od), __fmul_rn(r_var, __fsub_rn(1.f, __fdiv_rn(od, magn))));
}
// features cc and cc + 48 (cc in 0..47) of sample t of local ray `ray`: [ (k, xyz) sin | (k, xyz) sin(. + pi/2) ]; they share
// the mean, the variance and the exponential
// FP16_OUT: the caller rounds the features to fp16 (tensor pipeline).  A pair whose damping factor e is below 2^-25 rounds to
// zero whatever the sines are (|e sin| < half the smallest fp16 subnormal), so the two sines -- for the top octaves 2^14 x,
// 2^15 x the slow Payne-Hanek path of sinf -- are skipped: bit-identical fp16 features, about a third fewer pairs at T = 128.
// sin(y) for |y| < 1e6 at MUFU cost (tensor pipeline, FP16_OUT): y / 2pi as a two-float product (hi + lo carries ~48 bits), the
// integer part of hi removed exactly, sin(2 pi f) of the remainder on the MUFU unit.  Error vs the exact sine < 1e-6 (reduction
// 1.9e-7 measured over 2 M arguments up to 2e5, MUFU.SIN ~5e-7) -- invisible after the fp16 rounding of the feature (ulp 4.9e-4) --
// against ~250 instructions for the two sinf of a pair, whose top octaves take the Payne-Hanek path.
__device__ __forceinline__ float nf_sin_reduced(float y) {
  if (!(fabsf(y) < 1.0e6f)) return sinf(y);          // the reference layout's 1e10 last segment (and NaN / inf): the exact path
  const float hi = __fmul_rn(y, 0.15915493667125702f);
  const float lo = fmaf(y, 6.4206382432985265e-09f, fmaf(y, 0.15915493667125702f, -hi));
  const float k = __fadd_rn(__fadd_rn(hi, 12582912.f), -12582912.f);
  return __sinf(__fmul_rn(6.2831853071795865f, __fadd_rn(__fsub_rn(hi, k), lo)));
}
template <bool FP16_OUT = false>
__device__ __forceinline__ void nf_mip_feature_pair(const NfMipIn& m, long long ray, int t, int cc, float& f_sin, float& f_cos) {
  const int k = cc / 3, x = cc - 3 * k;
  float t0, t1, t_mean, t_var, r_var;
  nf_mip_segment(m, t, t0, t1);
  nf_mip_moments(m.mode, t0, t1, __ldg(m.radius + ray), t_mean, t_var, r_var);
  const float* r6 = m.rays + ray * 6;
  const float y = __fmul_rn(__fadd_rn(__fmul_rn(__ldg(r6 + 3 + x), t_mean), __ldg(r6 + x)), (float)(1 << k));
  float cov; int kv = k;
  if (m.mode == NF_MIP_CYLINDER_REF) {
    // the reference's layout: same flat index in [xyz, ray, k, t] as (t, ray, c) has in [t, ray, 48]
    long long flat = ((long long)t * m.n_rays_all + (m.ray_base + ray)) * 48 + cc;
    const int tv = (int)(flat % m.T); flat /= m.T;
    kv = (int)(flat & 15); flat >>= 4;
    const long long rv = flat % m.n_rays_all; const int xv = (int)(flat / m.n_rays_all);
    float a0, a1, tm, tvv, rvv;
    nf_mip_segment(m, tv, a0, a1);
    nf_mip_moments(m.mode, a0, a1, __ldg(m.radius_all + rv), tm, tvv, rvv);
    cov = nf_mip_cov(m.rays_all + rv * 6, xv, tvv, rvv);
  } else {
    cov = nf_mip_cov(r6, x, t_var, r_var);
  }
  if (FP16_OUT) {
    const float e = __expf(__fmul_rn(-0.5f, __fmul_rn(cov, (float)(1u << (2 * kv)))));
    if (e < 2.9e-8f) { f_sin = 0.f; f_cos = 0.f; return; }
    f_sin = __fmul_rn(e, nf_sin_reduced(y));
    f_cos = __fmul_rn(e, nf_sin_reduced(__fadd_rn(y, 1.5707963267948966f)));
    return;
  }
  const float e = expf(__fmul_rn(-0.5f, __fmul_rn(cov, (float)(1u << (2 * kv)))));
  f_sin = __fmul_rn(e, sinf(y));
  f_cos = __fmul_rn(e, sinf(__fadd_rn(y, 1.5707963267948966f)));
}
// Per-row invariants of the intended encoders (NF_MIP_CYLINDER / NF_MIP_CONE): the segment's moments and the three diagonal
// covariance entries do not depend on the feature index, so a thread that produces many features of one row computes them once.
// Same operations in the same order as nf_mip_feature_pair: bit-identical features.
struct NfMipRow { float o[3], d[3], cov[3], t_mean; };
__device__ __forceinline__ void nf_mip_row(const NfMipIn& m, long long ray, int t, NfMipRow& r) {
  float t0, t1, t_var, r_var;
  nf_mip_segment(m, t, t0, t1);
  nf_mip_moments(m.mode, t0, t1, __ldg(m.radius + ray), r.t_mean, t_var, r_var);
  const float* r6 = m.rays + ray * 6;
#pragma unroll
  for (int x = 0; x < 3; ++x) { r.o[x] = __ldg(r6 + x); r.d[x] = __ldg(r6 + 3 + x); r.cov[x] = nf_mip_cov(r6, x, t_var, r_var); }
}
template <bool FP16_OUT>
__device__ __forceinline__ void nf_mip_pair_of_row(const NfMipRow& r, int cc, float& f_sin, float& f_cos) {
  const int k = cc / 3, x = cc - 3 * k;
  const float dx = x == 0 ? r.d[0] : x == 1 ? r.d[1] : r.d[2], ox = x == 0 ? r.o[0] : x == 1 ? r.o[1] : r.o[2], cv = x == 0 ? r.cov[0] : x == 1 ? r.cov[1] : r.cov[2];
  const float y = __fmul_rn(__fadd_rn(__fmul_rn(dx, r.t_mean), ox), (float)(1 << k));
  if (FP16_OUT) {
    const float e = __expf(__fmul_rn(-0.5f, __fmul_rn(cv, (float)(1u << (2 * k)))));
    if (e < 2.9e-8f) { f_sin = 0.f; f_cos = 0.f; return; }
    f_sin = __fmul_rn(e, nf_sin_reduced(y));
    f_cos = __fmul_rn(e, nf_sin_reduced(__fadd_rn(y, 1.5707963267948966f)));
    return;
  }
  const float e = expf(__fmul_rn(-0.5f, __fmul_rn(cv, (float)(1u << (2 * k)))));
  f_sin = __fmul_rn(e, sinf(y));
  f_cos = __fmul_rn(e, sinf(__fadd_rn(y, 1.5707963267948966f)));
}
__device__ __forceinline__ float nf_mip_feature(const NfMipIn& m, long long ray, int t, int c) {
  float a, b;
  nf_mip_feature_pair(m, ray, t, c >= 48 ? c - 48 : c, a, b);
  return c >= 48 ? b : a;
}

// ---- Bezier spline of the deformation, reference src/nerf.py:1173-1178 (de_casteljau), 1201-1206 (cubic_bezier) ----
// ps[i] = control point i (one coordinate), n points, parameter t
__device__ __forceinline__ float nf_bezier(const float* ps, int n, float t) {
  const float m1t = __fsub_rn(1.f, t);
  if (n == 4) {
    const float m2 = __fmul_rn(m1t, m1t), t2 = __fmul_rn(t, t);
    const float k0 = __fmul_rn(m2, m1t), k1 = __fmul_rn(__fmul_rn(3.f, m2), t), k2 = __fmul_rn(__fmul_rn(3.f, t2), m1t), k3 = __fmul_rn(t2, t);
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(k0, ps[0]), __fmul_rn(k1, ps[1])), __fmul_rn(k2, ps[2])), __fmul_rn(k3, ps[3]));
  }
  float b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = i < n ? ps[i] : 0.f;
  // fully unrolled with predicates: every index is a compile-time constant, so b[] stays in registers
#pragma unroll
  for (int i = 1; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8 - i; ++j)
      if (i < n && j < n - i) b[j] = __fadd_rn(__fmul_rn(b[j], m1t), __fmul_rn(b[j + 1], t));
  return b[0];
}

// ---- compositing step, reference src/nerf.py:60-73 -----------------------------------------
// delta for sample t of a ray: clamp(ts[t+1]-ts[t], 1e-5) (1e10 for the last) times |r_d|.
__device__ __forceinline__ float nf_delta(const float* __restrict__ ts_ray, int t, int T, float rd_norm) {
  const float d = (t == T - 1) ? 1e10f : fmaxf(__fsub_rn(ts_ray[t + 1], ts_ray[t]), 1e-5f);
  return __fmul_rn(d, rd_norm);
}
__device__ __forceinline__ float nf_alpha(float sigma_raw, float delta, int density_act, float beta = 1.f) {
  const float s = nf_density_act_fn(sigma_raw, density_act, beta);
  return 1.f - expf(-s * delta);
}

// ---- tile <-> (ray, t) map -------------------------------------------------------------------
// A tile is `rows` consecutive samples. T <= rows: rpt = rows/T whole rays per tile.
// T > rows: a ray spans tpr = ceil(T/rows) consecutive tiles of the same CTA (carry in smem).
struct NfTileMap {
  int T, rows, rpt, tpr;
  __host__ __device__ NfTileMap(int T_, int rows_) : T(T_), rows(rows_) {
    if (T <= rows) { rpt = rows / T; tpr = 1; } else { rpt = 1; tpr = (T + rows - 1) / rows; }
  }
  // number of work units (a unit = one tile of rpt rays, or one whole ray of tpr tiles)
  __host__ __device__ long long units(long long n_rays) const { return T <= rows ? (n_rays + rpt - 1) / rpt : n_rays; }
  // row r of sub-tile s of unit u -> ray, t; returns validity
  __device__ __forceinline__ bool locate(long long u, int s, int r, long long n_rays, long long& ray, int& t) const {
    if (T <= rows) { const int rl = r / T; ray = u * rpt + rl; t = r - rl * T; return rl < rpt && ray < n_rays; }
    ray = u; t = s * rows + r; return t < T && ray < n_rays;
  }
};
// ---- tile <-> (ray, t) map of this kernel: the SAMPLE STREAM of a unit is cut into 128-row tiles ----------------------
// A unit = rpu whole rays = tpu whole tiles (rpu * Tp == tpu * 128, Tp = per-ray stride in the stream).  Tp = T packs rays back
// to back, so a tile may hold the tail of one ray and the head of the next (T = 192: 2 rays in 3 tiles instead of 4; T = 160:
// 4 rays in 5 tiles instead of 8).  Otherwise (T not a multiple of 32) Tp pads every ray to whole tiles / a divisor of 128 as
// the other kernels do.  All carries stay inside a unit, which one slot walks tile by tile.
struct NfStreamMap {
  int T, Tp, tpr, rpu;        // tpr = tiles per unit (the name the schedule code uses), rpu = rays per unit
  __host__ __device__ static int gcd(int a, int b) { while (b) { const int r = a % b; a = b; b = r; } return a; }
  __host__ __device__ NfStreamMap(int T_, int rows) : T(T_) {
    Tp = T_;
    // packing needs warp-aligned rays (T % 32 == 0): the in-warp scan/reduction trees then see every ray at the same lanes, so
    // a ray's rounding does not depend on its position in the unit (a sharded render must equal the whole bit for bit)
    if ((T_ & 31) != 0 || T_ / gcd(T_, rows) > 64) Tp = T_ <= rows ? rows / (rows / T_) : (T_ + rows - 1) / rows * rows;
    const int g = gcd(Tp, rows);
    tpr = Tp / g; rpu = rows / g;
  }
  __host__ __device__ long long units(long long n_rays) const { return (n_rays + rpu - 1) / rpu; }
  __device__ __forceinline__ bool locate(long long u, int sub, int r, long long n_rays, long long& ray, int& t) const {
    const int q = sub * NF_TC_ROWS + r, rl = q / Tp;
    t = q - rl * Tp; ray = u * rpu + rl;
    return t < T && ray < n_rays;
  }
};


// ---- training workspace layout (host) ------------------------------------------------------------------------------
static inline int nf_build_train_plan(const NfPlan& p, int64_t n_rays, int T, NfTrainPlan* tp) {
  *tp = NfTrainPlan{};
  const NfStreamMap map(T, NF_TC_ROWS);
  tp->T = T; tp->rpu = map.rpu; tp->tpr = map.tpr; tp->n_rays = n_rays;
  tp->n_tiles = map.units(n_rays) * map.tpr;
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
  tp->scale_off = take(16);
  tp->sigma_off = take(n_rays * T * 4); tp->rgbraw_off = take(n_rays * T * 12);
  tp->dsigma_off = take(n_rays * T * 4); tp->drgbraw_off = take(n_rays * T * 12);
  // fp32 gradient of the density MLP's hash features [n_tiles * 128][32]; Positional head: followed by the head's own encoder's
  tp->dx0_off = take(tp->n_tiles * NF_TC_ROWS * 32 * 4 * (p.refl_kind == NF_REFL_POSITIONAL ? 2 : 1));
  tp->bgrand_off = p.bg == NF_BG_RANDOM ? take(n_rays * 4) : -1;
  int nl = 0;
  for (int mi = 0; mi < p.n_mlps; ++mi) {
    const int m = p.kind == NF_KIND_DYN ? (mi + 2) % 3 : mi;          // execution order (as build_prog3)
    for (int j = 0; j < p.mlp[m].n_lin; ++j, ++nl) {
      if (nl >= NF_TRAIN_MAX_LIN) return NF_E_UNSUPPORTED;
      const NfLinPlan& L = p.mlp[m].lin[j];
      NfTrainLin& R = tp->lin[nl];
      R.m = m; R.j = j; R.n = L.n; R.n_pad = L.n_pad; R.k0_pad = L.k0_pad; R.k_hidden = L.k_hidden; R.act = p.mlp[m].act; R.x0_raw = L.x0_raw;
      R.a_tile = (int64_t)(L.k0_pad + L.k_hidden) * 256; R.g_tile = (int64_t)L.n_pad * 256;
    }
  }
  tp->n_lin = nl;
  tp->dw_begin = off;
  for (int i = 0; i < nl; ++i) {
    NfTrainLin& R = tp->lin[i];
    R.dw_off = take((int64_t)R.n_pad * (R.k0_pad + R.k_hidden) * 4); R.db_off = take((int64_t)R.n_pad * 4);
  }
  tp->dw_end = off;
  for (int i = 0; i < nl; ++i) {
    NfTrainLin& R = tp->lin[i];
    R.a_off = take(tp->n_tiles * R.a_tile);
    R.c_off = (R.act == NF_ACT_SIN && R.k_hidden) ? take(tp->n_tiles * 65536) : -1;
    R.g_off = take(tp->n_tiles * R.g_tile);
  }
  tp->total_bytes = off;
  return 0;
}
#endif  // __CUDACC__
