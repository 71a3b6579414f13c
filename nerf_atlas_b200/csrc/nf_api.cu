// nf_api.cu -- the C ABI declared in include/nerf_b200.h.  Argument checking, plan building and
// kernel launches only; no torch types, no global mutable state (errors are thread local).
#include <string>
#include <cstdio>
#include <cstdlib>
#include "nf_common.cuh"
#include "nf_kernels.h"

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char* where) {
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
  return (int)e;
}
int plan_of(const nf_model_desc* d, NfPlan* p) {
  const char* why = "";
  const int rc = nf_build_plan(d, p, &why);
  if (rc) g_err = std::string("nf_model_desc: ") + why;
  return rc;
}
int check_ts(int32_t T, int64_t ts_stride) {
  if (T < 1) return fail(NF_E_BADARG, "T must be >= 1");
  if (ts_stride != 0 && ts_stride != T) return fail(NF_E_BADARG, "ts_ray_stride must be 0 (shared ts[T]) or T (per-ray ts[R,T])");
  return 0;
}
}  // namespace

extern "C" {

int nf_version(void) { return NF_ABI_VERSION; }
int nf_build_flags(void) {
  int f = 0;
#ifdef NF_EXPERIMENTS
  f |= NF_BUILD_EXPERIMENTS;
#endif
#ifdef NF_TC_STATS
  f |= NF_BUILD_STATS;
#endif
#ifdef NF_TC_TRACE
  f |= NF_BUILD_TRACE;
#endif
  return f;
}
const char* nf_last_error(void) { return g_err.c_str(); }

const char* nf_tensor_pipeline_support(const nf_model_desc* desc) {
  NfPlan p;
  if (plan_of(desc, &p)) return g_err.c_str();
  return nf_tc3_unsupported(p);          // the staggered pipeline is the superset of what the older tensor pipelines run
}

int nf_param_count(const nf_model_desc* desc) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  int n = 0;
  for (int m = 0; m < p.n_mlps; ++m) n += 2 * p.mlp[m].n_lin;
  if (p.enc == NF_ENC_HASH) n += p.hash_levels;
  if (p.kind == NF_KIND_DYN && p.deform_enc == NF_ENC_HASH) n += p.hash_levels;
  if (p.refl_kind != NF_REFL_VIEW) n += p.hash_levels;
  if (p.enc == NF_ENC_FOURIER) n += 1;
  if (p.density_act == NF_DENS_LAPLACE) n += 1;
  return n;
}

int64_t nf_packed_bytes(const nf_model_desc* desc) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  return p.total_bytes;
}

int nf_pack_weights(const nf_model_desc* desc, const float* const* params, int32_t n_params,
                    void* packed, int64_t packed_bytes, void* stream) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  if (!params || !packed) return fail(NF_E_BADARG, "nf_pack_weights: null pointer");
  if (n_params != nf_param_count(desc)) return fail(NF_E_BADARG, "nf_pack_weights: wrong number of parameter pointers");
  if (packed_bytes < p.total_bytes) return fail(NF_E_SMALLBUF, "nf_pack_weights: packed buffer too small");
  if (((uintptr_t)packed & 1023) != 0) return fail(NF_E_BADARG, "nf_pack_weights: packed must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* base = (uint8_t*)packed;
  int pi = 0;
  for (int m = 0; m < p.n_mlps; ++m)
    for (int j = 0; j < p.mlp[m].n_lin; ++j) { if (!params[pi] || !params[pi + 1]) return fail(NF_E_BADARG, "nf_pack_weights: null parameter"); pi += 2; }
  {
    cudaError_t e = nf_launch_pack_all(p, params, packed, st);          // one memset + one launch for every Linear's images
    if (e != cudaSuccess) return cuda_fail(e, "pack Linears");
  }
  {
    // the embedding tables (one parameter per level; up to three encoders): one launch
    const size_t per = (size_t)(p.hash_mask + 1) * 4 * sizeof(float);
    const float* src[48]; float* dst[48]; int nt = 0;
    const int64_t offs[3] = {p.hash_off, p.hash2_off, p.hash3_off};
    const bool on[3] = {p.enc == NF_ENC_HASH, p.kind == NF_KIND_DYN && p.deform_enc == NF_ENC_HASH, p.refl_kind != NF_REFL_VIEW};
    for (int set = 0; set < 3; ++set) {
      if (!on[set]) continue;
      for (int l = 0; l < p.hash_levels; ++l) {
        const float* t = params[pi++];
        if (!t) return fail(NF_E_BADARG, "nf_pack_weights: null hash table");
        src[nt] = t; dst[nt] = (float*)(base + offs[set] + l * per); ++nt;
      }
    }
    cudaError_t e = nf_launch_copy_tables(src, dst, nt, per, st);
    if (e != cudaSuccess) return cuda_fail(e, "pack hash tables");
  }
  if (p.enc == NF_ENC_FOURIER) {
    const float* b = params[pi++];
    if (!b) return fail(NF_E_BADARG, "nf_pack_weights: null fourier basis");
    cudaError_t e = cudaMemcpyAsync(base + p.fourier_off, b, (size_t)3 * p.fourier_freqs * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "pack fourier basis");
  }
  if (p.density_act == NF_DENS_LAPLACE) {
    const float* b = params[pi++];
    if (!b) return fail(NF_E_BADARG, "nf_pack_weights: null beta");
    cudaError_t e = cudaMemcpyAsync(base + p.scale_off, b, sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "pack beta");
  }
  return 0;
}

int nf_render_forward(const nf_model_desc* desc, const void* packed, const float* rays, int64_t n_rays,
                      const float* ts, int32_t T, int64_t ts_ray_stride, const float* density_noise, const float* ray_time,
                      const nf_mip_args* mip, float* rgb_out, float* alpha_out, float* weights_out, int32_t precision, void* stream) {
  return nf_render_forward_aux(desc, packed, rays, n_rays, ts, T, ts_ray_stride, density_noise, ray_time, mip, rgb_out, alpha_out,
                               weights_out, nullptr, precision, stream);
}

int nf_train_layout_of(const nf_model_desc* desc, int64_t n_rays, int32_t T, nf_train_layout* out) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  if (!out || n_rays < 0 || T < 1) return fail(NF_E_BADARG, "nf_train_layout_of: bad argument");
  if (const char* why = nf_train_unsupported(p)) return fail(NF_E_UNSUPPORTED, why);
  if (p.refl_kind == NF_REFL_POSITIONAL && (T & 31)) return fail(NF_E_UNSUPPORTED, "training the Positional head: T must be a multiple of 32 (warp-aligned rays)");
  if (int rc = nf_build_train_plan(p, n_rays, T, out)) return fail(rc, "nf_train_layout_of: too many Linear layers");
  return 0;
}

int nf_render_backward(const nf_model_desc* desc, const void* packed, void* train_ws, int64_t train_ws_bytes,
                       const float* rays, int64_t n_rays, const float* ts, int32_t T, int64_t ts_ray_stride,
                       const float* d_rgb, float* const* grads, int32_t n_grads, void* stream) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  if (const char* why = nf_train_unsupported(p)) return fail(NF_E_UNSUPPORTED, why);
  if (n_rays < 0) return fail(NF_E_BADARG, "n_rays < 0");
  if (int rc = check_ts(T, ts_ray_stride)) return rc;
  if (!packed || !train_ws || !rays || !ts || !d_rgb || !grads) return fail(NF_E_BADARG, "nf_render_backward: null pointer");
  if (n_grads != nf_param_count(desc)) return fail(NF_E_BADARG, "nf_render_backward: wrong number of gradient pointers");
  if (((uintptr_t)train_ws & 1023) != 0) return fail(NF_E_BADARG, "nf_render_backward: train_ws must be 1024-byte aligned");
  NfTrainPlan tp;
  if (int rc = nf_build_train_plan(p, n_rays, T, &tp)) return fail(rc, "nf_render_backward: too many Linear layers");
  if (train_ws_bytes < tp.total_bytes) return fail(NF_E_SMALLBUF, "nf_render_backward: workspace too small");
  if (T > 2048) return fail(NF_E_UNSUPPORTED, "nf_render_backward: T <= 2048");
  if (n_rays == 0) return 0;
  cudaError_t e = nf_launch_render_backward(p, tp, packed, train_ws, rays, ts, ts_ray_stride, d_rgb, grads, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_render_backward");
}

int nf_render_forward_aux(const nf_model_desc* desc, const void* packed, const float* rays, int64_t n_rays,
                          const float* ts, int32_t T, int64_t ts_ray_stride, const float* density_noise, const float* ray_time,
                          const nf_mip_args* mip, float* rgb_out, float* alpha_out, float* weights_out, const nf_render_aux* aux,
                          int32_t precision, void* stream) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  if (aux && aux->struct_bytes != (int32_t)sizeof(nf_render_aux)) return fail(NF_E_BADARG, "bad nf_render_aux (struct_bytes)");
  NfTrainPlan tp; const NfTrainPlan* train = nullptr;
  if (aux && aux->train_ws) {
    if (precision != NF_PREC_FP16_TC) return fail(NF_E_UNSUPPORTED, "training forward: NF_PREC_FP16_TC only");
    if (const char* why = nf_train_unsupported(p)) return fail(NF_E_UNSUPPORTED, why);
    if (n_rays < 0 || T < 1) return fail(NF_E_BADARG, "bad n_rays / T");
    if (int rc = nf_build_train_plan(p, n_rays, T, &tp)) return fail(rc, "training forward: too many Linear layers");
    if (aux->train_ws_bytes < tp.total_bytes) return fail(NF_E_SMALLBUF, "training forward: workspace too small (nf_train_layout_of().total_bytes)");
    if (((uintptr_t)aux->train_ws & 1023) != 0) return fail(NF_E_BADARG, "training forward: train_ws must be 1024-byte aligned");
    train = &tp;
  }
  if (n_rays < 0) return fail(NF_E_BADARG, "n_rays < 0");
  if (n_rays == 0) return 0;
  if (!packed || !rays || !ts || !rgb_out) return fail(NF_E_BADARG, "nf_render_forward: null pointer");
  if (p.bg == NF_BG_RANDOM && !(aux && aux->bg_rand)) return fail(NF_E_BADARG, "NF_BG_RANDOM needs nf_render_aux.bg_rand (one uniform draw per ray)");
  const bool side = aux && (aux->pts_out || aux->dp_out || aux->rigid_dp_out || aux->rigidity_out);
  if (side && p.kind != NF_KIND_DYN && (aux->dp_out || aux->rigid_dp_out || aux->rigidity_out)) return fail(NF_E_BADARG, "dp / rigid_dp / rigidity outputs exist for NF_KIND_DYN only");
  if (aux && aux->pts && p.kind == NF_KIND_DYN) return fail(NF_E_UNSUPPORTED, "from_pts (explicit sample positions) is not built for NF_KIND_DYN");
  if (train && aux->pts) return fail(NF_E_UNSUPPORTED, "training forward: explicit pts are not built");
  if (int rc = check_ts(T, ts_ray_stride)) return rc;
  if (p.kind == NF_KIND_DYN && !ray_time) return fail(NF_E_BADARG, "nf_render_forward: ray_time is required for NF_KIND_DYN");
  if (p.mip != NF_MIP_NONE) {
    if (!mip || !mip->radius) return fail(NF_E_BADARG, "nf_render_forward: nf_mip_args with radii is required when desc.mip is set");
    if (ts_ray_stride != 0) return fail(NF_E_UNSUPPORTED, "nf_render_forward: the Mip encoder needs a shared ts[T]");
    if (p.mip == NF_MIP_CYLINDER_REF && (!mip->rays_all || !mip->radius_all || mip->ray_base < 0 || mip->ray_base + n_rays > mip->n_rays_all))
      return fail(NF_E_BADARG, "nf_render_forward: NF_MIP_CYLINDER_REF needs the whole crop (rays_all, radius_all, n_rays_all, ray_base)");
  }
  cudaError_t e;
  if (precision == NF_PREC_FP32)
    e = nf_launch_render_fp32(p, packed, rays, n_rays, ts, T, ts_ray_stride, density_noise, ray_time, mip, rgb_out, alpha_out, weights_out, (cudaStream_t)stream, aux);
  else if (precision == NF_PREC_FP16_TC) {
    // The staggered paired pipeline (nf_tc3.cu) is the product path; the single-CTA pipeline (nf_tc.cu) takes the few
    // descriptors it cannot (odd hash levels, intermediate % 16 != 0).  Only NF_EXPERIMENTS builds (A/B timing,
    // profiles/perf_variants.py) read NF_TC_PIPE: 3 = staggered, 1 = single CTA.
    int pipe = 3;
#ifdef NF_EXPERIMENTS
    if (const char* env = getenv("NF_TC_PIPE")) pipe = atoi(env);
    if (const char* env = getenv("NF_TC_PAIRED")) if (env[0] == '0') pipe = 1;
#endif
    if (p.kind == NF_KIND_DYN || p.mip != NF_MIP_NONE || p.refl_kind != NF_REFL_VIEW || p.enc == NF_ENC_FOURIER) {
      // only the staggered pipeline runs the three-MLP chain and the wide-x0 (single-tile) mode
      if (const char* why = nf_tc3_unsupported(p)) return fail(NF_E_UNSUPPORTED, why);
      pipe = 3;
    }
    if (pipe == 3 && nf_tc3_unsupported(p)) pipe = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (pipe != 3) pipe = 1;
    if (train && pipe != 3) return fail(NF_E_UNSUPPORTED, "training forward: the staggered pipeline only");
    const bool want_aux = aux && (aux->pts || aux->bg_rand || side);
    if (want_aux && pipe != 3) return fail(NF_E_UNSUPPORTED, "explicit pts / random background / side channels: staggered tensor pipeline or NF_PREC_FP32 only");
    if (want_aux && (p.mip != NF_MIP_NONE || p.refl_kind != NF_REFL_VIEW || p.enc == NF_ENC_FOURIER) && aux->pts)
      return fail(NF_E_UNSUPPORTED, "explicit pts with a wide-x0 model (Mip, Positional, Fourier SDF): NF_PREC_FP32 only");
    if (aux && aux->pts_out && p.kind != NF_KIND_DYN) return fail(NF_E_UNSUPPORTED, "pts_out on the tensor pipeline: NF_KIND_DYN only (use nf_sample_points)");
    e = pipe == 3 ? nf_launch_render_tc3(p, packed, rays, n_rays, ts, T, ts_ray_stride, density_noise, ray_time, mip, rgb_out, alpha_out, weights_out, st, train, train ? aux->train_ws : nullptr, aux)
                  : nf_launch_render_tc(p, packed, rays, n_rays, ts, T, ts_ray_stride, density_noise, rgb_out, alpha_out, weights_out, st);
  }
  else return fail(NF_E_BADARG, "unknown precision");
  if (e != cudaSuccess) return cuda_fail(e, "nf_render_forward");
  return 0;
}

int nf_generate_rays(const float* cam_to_world, int64_t B, float focal, int32_t size, int32_t top, int32_t left,
                     int32_t H, int32_t W, int32_t scalar_div_as_reciprocal, float* rays_out, void* stream) {
  if (B < 0 || H < 0 || W < 0 || size <= 0 || !(focal > 0.f)) return fail(NF_E_BADARG, "nf_generate_rays: bad size / focal");
  if (B == 0 || H == 0 || W == 0) return 0;
  if (!cam_to_world || !rays_out) return fail(NF_E_BADARG, "nf_generate_rays: null pointer");
  cudaError_t e = nf_launch_generate_rays(cam_to_world, B, focal, size, top, left, H, W, scalar_div_as_reciprocal ? 1 : 0, rays_out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_generate_rays");
}

int nf_generate_rays_dtu(const float* pose, const float* intrinsic, int32_t intr_rows, int32_t intr_cols, int64_t B, int32_t size,
                         int32_t top, int32_t left, int32_t H, int32_t W, float* rays_out, void* stream) {
  if (B < 0 || H < 0 || W < 0 || size <= 0 || intr_rows < 3 || intr_cols < 3) return fail(NF_E_BADARG, "nf_generate_rays_dtu: bad size / intrinsic shape");
  if (B == 0 || H == 0 || W == 0) return 0;
  if (!pose || !intrinsic || !rays_out) return fail(NF_E_BADARG, "nf_generate_rays_dtu: null pointer");
  cudaError_t e = nf_launch_generate_rays_dtu(pose, intrinsic, intr_rows, intr_cols, B, size, top, left, H, W, rays_out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_generate_rays_dtu");
}

int nf_ray_radii(const float* rays, int64_t B, int32_t H, int32_t W, float* radius_out, void* stream) {
  if (B < 0 || H < 0 || W < 0) return fail(NF_E_BADARG, "nf_ray_radii: negative size");
  if (B == 0 || W == 0 || H == 0) return 0;
  if (H < 3) return fail(NF_E_UNSUPPORTED, "nf_ray_radii: needs H >= 3 (the reference's radii_x indexes row H-3)");
  if (!rays || !radius_out) return fail(NF_E_BADARG, "nf_ray_radii: null pointer");
  cudaError_t e = nf_launch_ray_radii(rays, B, H, W, radius_out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_ray_radii");
}

int nf_sample_points(const float* rays, int64_t n_rays, const float* ts, int32_t T, int64_t ts_ray_stride, float* pts_out, void* stream) {
  if (n_rays < 0) return fail(NF_E_BADARG, "n_rays < 0");
  if (n_rays == 0) return 0;
  if (!rays || !ts || !pts_out) return fail(NF_E_BADARG, "nf_sample_points: null pointer");
  if (int rc = check_ts(T, ts_ray_stride)) return rc;
  cudaError_t e = nf_launch_sample_points(rays, n_rays, ts, T, ts_ray_stride, pts_out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_sample_points");
}

int nf_hash_encode(const nf_model_desc* desc, const void* packed, const float* pts, int64_t n, float* feats_out, uint16_t* idx_out, void* stream) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  if (p.enc != NF_ENC_HASH) return fail(NF_E_BADARG, "nf_hash_encode: model has no hash encoder");
  if (n < 0) return fail(NF_E_BADARG, "n < 0");
  if (n == 0) return 0;
  if (!packed || !pts || !feats_out) return fail(NF_E_BADARG, "nf_hash_encode: null pointer");
  cudaError_t e = nf_launch_hash_encode(p, packed, pts, n, feats_out, idx_out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_hash_encode");
}

int nf_composite(const nf_model_desc* desc, const void* packed, const float* sigma_raw, const float* feats, const float* rays, int64_t n_rays,
                 const float* ts, int32_t T, int64_t ts_ray_stride, float* rgb_out, float* alpha_out, float* weights_out, void* stream) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  if (n_rays < 0) return fail(NF_E_BADARG, "n_rays < 0");
  if (n_rays == 0) return 0;
  if (!sigma_raw || !feats || !rays || !ts || !rgb_out) return fail(NF_E_BADARG, "nf_composite: null pointer");
  if (p.bg == NF_BG_RANDOM) return fail(NF_E_UNSUPPORTED, "nf_composite: NF_BG_RANDOM exists on the fused path only (nf_render_forward_aux)");
  if (p.density_act == NF_DENS_LAPLACE && !packed) return fail(NF_E_BADARG, "nf_composite: packed (beta) required for the Laplace density");
  if (int rc = check_ts(T, ts_ray_stride)) return rc;
  cudaError_t e = nf_launch_composite(p, packed, sigma_raw, feats, rays, n_rays, ts, T, ts_ray_stride, rgb_out, alpha_out, weights_out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_composite");
}

int nf_integrate(const float* weights, const float* vals, int64_t n_rays, int32_t T, int32_t channels, int64_t vals_ray_stride,
                 float* out, void* stream) {
  if (n_rays < 0) return fail(NF_E_BADARG, "n_rays < 0");
  if (n_rays == 0) return 0;
  if (!weights || !vals || !out) return fail(NF_E_BADARG, "nf_integrate: null pointer");
  if (T < 1) return fail(NF_E_BADARG, "nf_integrate: T >= 1");
  if (channels != 1 && channels != 3) return fail(NF_E_UNSUPPORTED, "nf_integrate: 1 or 3 channels (depth / rigidity, flow)");
  if (vals_ray_stride != 0 && vals_ray_stride < (int64_t)T * channels) return fail(NF_E_BADARG, "nf_integrate: vals_ray_stride is 0 (values shared by all rays) or >= T * channels");
  cudaError_t e = nf_launch_integrate(weights, vals, n_rays, T, channels, vals_ray_stride, out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_integrate");
}

int nf_sample_pdf(const float* ts_coarse, int32_t T, const float* weights, int64_t n_rays, const float* u, int32_t n_fine,
                  float* ts_out, void* stream) {
  if (n_rays < 0) return fail(NF_E_BADARG, "n_rays < 0");
  if (n_rays == 0) return 0;
  if (!ts_coarse || !weights || !u || !ts_out) return fail(NF_E_BADARG, "nf_sample_pdf: null pointer");
  if (T < 3 || T > 256 || n_fine < 1 || n_fine > 256) return fail(NF_E_UNSUPPORTED, "nf_sample_pdf: need 3 <= T <= 256 and 1 <= n_fine <= 256");
  cudaError_t e = nf_launch_sample_pdf(ts_coarse, T, weights, n_rays, u, n_fine, ts_out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_sample_pdf");
}

int nf_mlp_forward(const nf_model_desc* desc, const void* packed, int32_t which, const float* x0, int64_t n, float* out,
                   int32_t precision, void* stream) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  if (which < 0 || which >= p.n_mlps) return fail(NF_E_BADARG, "nf_mlp_forward: no such MLP");
  if (n < 0) return fail(NF_E_BADARG, "n < 0");
  if (n == 0) return 0;
  if (!packed || !x0 || !out) return fail(NF_E_BADARG, "nf_mlp_forward: null pointer");
  cudaError_t e;
  if (precision == NF_PREC_FP32) e = nf_launch_mlp_fp32(p, which, packed, x0, n, out, (cudaStream_t)stream);
  else if (precision == NF_PREC_FP16_TC) e = nf_launch_mlp_tc(p, which, packed, x0, n, out, (cudaStream_t)stream);
  else return fail(NF_E_BADARG, "unknown precision");
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_mlp_forward");
}

namespace {
int sdf_check(const nf_model_desc* desc, NfPlan* p, const void* packed, const float* rays, int64_t n_rays, int32_t iters, int32_t precision,
              void* workspace, int64_t workspace_bytes) {
  if (int rc = plan_of(desc, p)) return rc;
  if (p->kind != NF_KIND_PLAIN || p->mip != NF_MIP_NONE || p->refl_kind != NF_REFL_VIEW || (p->enc != NF_ENC_NONE && p->enc != NF_ENC_FOURIER))
    return fail(NF_E_UNSUPPORTED, "SDF surface side: a VolSDF-style descriptor (SIREN or Fourier-encoded SDF network + View head)");
  if (n_rays < 0 || iters < 0 || iters > 1000) return fail(NF_E_BADARG, "SDF surface side: n_rays >= 0, 0 <= iters <= 1000");
  if (n_rays >= (1LL << 31)) return fail(NF_E_UNSUPPORTED, "SDF surface side: fewer than 2^31 rays per call");
  if (precision != NF_PREC_FP32 && precision != NF_PREC_FP16_TC) return fail(NF_E_BADARG, "unknown precision");
  if (precision == NF_PREC_FP16_TC && p->enc == NF_ENC_FOURIER)
    return fail(NF_E_UNSUPPORTED, "SDF surface side: the Fourier-encoded SDF network runs on NF_PREC_FP32 only (x0 is 259 wide)");
  if (n_rays == 0) return 0;
  if (!packed || !rays || !workspace) return fail(NF_E_BADARG, "SDF surface side: null pointer");
  if (((uintptr_t)workspace & 255) != 0) return fail(NF_E_BADARG, "SDF surface side: workspace must be 256-byte aligned");
  if (workspace_bytes < nf_sdf_workspace_bytes_of(*p, n_rays)) return fail(NF_E_SMALLBUF, "SDF surface side: workspace too small (nf_sdf_workspace_bytes)");
  return 0;
}
}  // namespace

int64_t nf_sdf_workspace_bytes(const nf_model_desc* desc, int64_t n_rays) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  if (n_rays < 0) return fail(NF_E_BADARG, "n_rays < 0");
  return nf_sdf_workspace_bytes_of(p, n_rays);
}

int nf_sphere_march(const nf_model_desc* desc, const void* packed, const float* rays, int64_t n_rays, float t_near, float t_far, int32_t iters,
                    float eps, float bound_rad, int32_t precision, float* pts_out, uint8_t* hit_out, float* t_out, void* workspace,
                    int64_t workspace_bytes, void* stream) {
  NfPlan p; if (int rc = sdf_check(desc, &p, packed, rays, n_rays, iters, precision, workspace, workspace_bytes)) return rc;
  if (n_rays == 0) return 0;
  cudaError_t e = nf_launch_sphere_march(p, packed, rays, n_rays, t_near, t_far, iters, eps, bound_rad, precision, pts_out, hit_out, t_out, workspace, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_sphere_march");
}

int nf_sdf_render(const nf_model_desc* desc, const void* packed, const float* rays, int64_t n_rays, float t_near, float t_far, int32_t iters,
                  float eps, float bound_rad, int32_t precision, float* rgb_out, uint8_t* hit_out, float* t_out, float* pts_out, void* workspace,
                  int64_t workspace_bytes, void* stream) {
  NfPlan p; if (int rc = sdf_check(desc, &p, packed, rays, n_rays, iters, precision, workspace, workspace_bytes)) return rc;
  if (n_rays == 0) return 0;
  if (!rgb_out) return fail(NF_E_BADARG, "nf_sdf_render: null rgb_out");
  cudaError_t e = nf_launch_sdf_render(p, packed, rays, n_rays, t_near, t_far, iters, eps, bound_rad, precision, rgb_out, hit_out, t_out, pts_out, workspace, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_sdf_render");
}

int nf_sdf_bisect(const nf_model_desc* desc, const void* packed, const float* rays, int64_t n_rays, float t_near, float t_far, int32_t iters,
                  float jitter, float bound_rad, int32_t precision, float* pts_out, uint8_t* hit_out, float* tput_out, float* best_pos_out,
                  float* rgb_out, void* workspace, int64_t workspace_bytes, void* stream) {
  NfPlan p; if (int rc = sdf_check(desc, &p, packed, rays, n_rays, iters, precision, workspace, workspace_bytes)) return rc;
  if (n_rays == 0) return 0;
  if (iters < 1) return fail(NF_E_BADARG, "nf_sdf_bisect: iters >= 1");
  if (!(jitter >= 0.f && jitter < 1.f)) return fail(NF_E_BADARG, "nf_sdf_bisect: jitter in [0, 1) (the reference's random.random() draw)");
  cudaError_t e = nf_launch_sdf_bisect(p, packed, rays, n_rays, t_near, t_far, iters, jitter, bound_rad, precision, pts_out, hit_out, tput_out, best_pos_out,
                                       rgb_out, workspace, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_sdf_bisect");
}

int nf_sdf_normals(const nf_model_desc* desc, const void* packed, const float* pts, int64_t n, float bound_rad, float* normals_out,
                   float* values_out, void* stream) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  if (p.kind != NF_KIND_PLAIN || p.mip != NF_MIP_NONE || (p.enc != NF_ENC_NONE && p.enc != NF_ENC_FOURIER))
    return fail(NF_E_UNSUPPORTED, "nf_sdf_normals: a VolSDF-style descriptor (SIREN or Fourier-encoded SDF network)");
  if (p.mlp[0].hidden_ref != NF_HIDDEN) return fail(NF_E_UNSUPPORTED, "nf_sdf_normals: hidden_size 256");
  if (n < 0) return fail(NF_E_BADARG, "nf_sdf_normals: n >= 0");
  if (n == 0) return 0;
  if (!packed || !pts || !normals_out) return fail(NF_E_BADARG, "nf_sdf_normals: null pointer");
  cudaError_t e = nf_launch_sdf_normals(p, packed, pts, n, bound_rad, normals_out, values_out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_sdf_normals");
}

int nf_composite_backward(const nf_model_desc* desc, const void* packed, const float* sigma_raw, const float* feats,
                          const float* rays, int64_t n_rays, const float* ts, int32_t T, int64_t ts_ray_stride,
                          const float* d_rgb, float* d_sigma_raw_out, float* d_feats_out, void* stream) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  if (n_rays < 0) return fail(NF_E_BADARG, "n_rays < 0");
  if (n_rays == 0) return 0;
  if (!sigma_raw || !feats || !rays || !ts || !d_rgb || !d_sigma_raw_out || !d_feats_out) return fail(NF_E_BADARG, "nf_composite_backward: null pointer");
  if (p.bg == NF_BG_RANDOM) return fail(NF_E_UNSUPPORTED, "nf_composite_backward: NF_BG_RANDOM is not built");
  if (p.density_act == NF_DENS_LAPLACE && !packed) return fail(NF_E_BADARG, "nf_composite_backward: packed (beta) required for the Laplace density");
  if (int rc = check_ts(T, ts_ray_stride)) return rc;
  if (T > 2048) return fail(NF_E_UNSUPPORTED, "nf_composite_backward: T <= 2048");
  cudaError_t e = nf_launch_composite_bwd(p, packed, sigma_raw, feats, rays, n_rays, ts, T, ts_ray_stride, d_rgb, d_sigma_raw_out, d_feats_out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_composite_backward");
}

int nf_hash_encode_backward(const nf_model_desc* desc, const float* pts, int64_t n, const float* d_feats, float* d_tables, void* stream) {
  NfPlan p; if (int rc = plan_of(desc, &p)) return rc;
  if (p.enc != NF_ENC_HASH) return fail(NF_E_BADARG, "nf_hash_encode_backward: model has no hash encoder");
  if (n < 0) return fail(NF_E_BADARG, "n < 0");
  if (n == 0) return 0;
  if (!pts || !d_feats || !d_tables) return fail(NF_E_BADARG, "nf_hash_encode_backward: null pointer");
  if (((uintptr_t)d_tables & 15) || ((uintptr_t)d_feats & 15)) return fail(NF_E_BADARG, "nf_hash_encode_backward: d_tables / d_feats must be 16-byte aligned");
  cudaError_t e = nf_launch_hash_encode_bwd(p, pts, n, d_feats, d_tables, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_hash_encode_backward");
}

int nf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                 float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream) {
  if (n < 0 || step < 1) return fail(NF_E_BADARG, "nf_adam_step: n >= 0 and step >= 1");
  if (n == 0) return 0;
  if (!param || !grad || !exp_avg || !exp_avg_sq) return fail(NF_E_BADARG, "nf_adam_step: null pointer");
  if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0) return fail(NF_E_BADARG, "nf_adam_step: pointers must be 16-byte aligned");
  cudaError_t e = nf_launch_adam_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_adam_step");
}

int nf_adam_step_multi(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                       const int64_t* numel, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream) {
  if (n_tensors < 0 || step < 1) return fail(NF_E_BADARG, "nf_adam_step_multi: n_tensors >= 0 and step >= 1");
  if (n_tensors == 0) return 0;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !numel) return fail(NF_E_BADARG, "nf_adam_step_multi: null pointer");
  for (int t = 0; t < n_tensors; ++t) {
    if (numel[t] < 0) return fail(NF_E_BADARG, "nf_adam_step_multi: numel < 0");
    if (numel[t] == 0) continue;
    if (!params[t] || !grads[t] || !exp_avg[t] || !exp_avg_sq[t]) return fail(NF_E_BADARG, "nf_adam_step_multi: null tensor pointer");
    if ((((uintptr_t)params[t] | (uintptr_t)grads[t] | (uintptr_t)exp_avg[t] | (uintptr_t)exp_avg_sq[t]) & 15) != 0)
      return fail(NF_E_BADARG, "nf_adam_step_multi: pointers must be 16-byte aligned");
  }
  cudaError_t e = nf_launch_adam_multi(n_tensors, params, grads, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps, weight_decay, step, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "nf_adam_step_multi");
}

}  // extern "C"
