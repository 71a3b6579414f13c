// nf_kernels.h -- launchers implemented in nf_fp32.cu / nf_tc.cu, called from nf_api.cu.
#pragma once
#include "nf_common.cuh"

// W[n_ref][kh_ref + kx] (nn.Linear) -> Wt[kh_pad + kx][n_pad] (hidden rows kh_ref..kh_pad-1 and columns n_ref..n_pad-1 zero), bias[n_pad]
cudaError_t nf_launch_render_fp32(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, const float* ts,
                                  int T, int64_t ts_stride, const float* noise, const float* ray_time, const nf_mip_args* mip,
                                  float* rgb, float* alpha, float* weights, cudaStream_t st, const nf_render_aux* aux = nullptr);
cudaError_t nf_launch_generate_rays(const float* c2w, int64_t B, float focal, int size, int top, int left, int H, int W, int recip,
                                    float* out, cudaStream_t st);
cudaError_t nf_launch_generate_rays_dtu(const float* pose, const float* intr, int ir, int ic, int64_t B, int size, int top, int left, int H, int W,
                                        float* out, cudaStream_t st);
cudaError_t nf_launch_ray_radii(const float* rays, int64_t B, int H, int W, float* out, cudaStream_t st);
cudaError_t nf_launch_render_tc(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, const float* ts,
                                int T, int64_t ts_stride, const float* noise, float* rgb, float* alpha, float* weights,
                                cudaStream_t st);
// n_dev (nullable): the row count lives in device memory and n is only the capacity that sizes the grid
cudaError_t nf_launch_mlp_fp32(const NfPlan& plan, int which, const void* packed, const float* x0, int64_t n, float* out, cudaStream_t st,
                               const long long* n_dev = nullptr);
cudaError_t nf_launch_mlp_tc(const NfPlan& plan, int which, const void* packed, const float* x0, int64_t n, float* out, cudaStream_t st,
                             const long long* n_dev = nullptr);
// SDF surface side (nf_march.cu): sphere tracing + shading of the hit points
cudaError_t nf_launch_sphere_march(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, float near, float far, int iters,
                                   float eps, float bound_rad, int precision, float* pts_out, uint8_t* hit_out, float* t_out, void* ws, cudaStream_t st);
cudaError_t nf_launch_sdf_normals(const NfPlan& plan, const void* packed, const float* pts, int64_t n, float bound_rad, float* normals, float* values,
                                  cudaStream_t st);
cudaError_t nf_launch_sdf_bisect(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, float near, float far, int iters, float jitter,
                                 float bound_rad, int precision, float* pts_out, uint8_t* hit_out, float* tput_out, float* best_out, float* rgb_out,
                                 void* ws, cudaStream_t st);
cudaError_t nf_launch_sdf_render(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, float near, float far, int iters,
                                 float eps, float bound_rad, int precision, float* rgb_out, uint8_t* hit_out, float* t_out, float* pts_out,
                                 void* ws, cudaStream_t st);
int64_t nf_sdf_workspace_bytes_of(const NfPlan& plan, int64_t n_rays);
cudaError_t nf_launch_sample_points(const float* rays, int64_t n_rays, const float* ts, int T, int64_t ts_stride, float* pts, cudaStream_t st);
cudaError_t nf_launch_hash_encode(const NfPlan& plan, const void* packed, const float* pts, int64_t n, float* feats, uint16_t* idx, cudaStream_t st);
cudaError_t nf_launch_composite(const NfPlan& plan, const void* packed, const float* sigma_raw, const float* feats, const float* rays, int64_t n_rays,
                                const float* ts, int T, int64_t ts_stride, float* rgb, float* alpha, float* weights, cudaStream_t st);
cudaError_t nf_launch_sample_pdf(const float* ts, int T, const float* weights, int64_t n_rays, const float* u, int nf, float* out, cudaStream_t st);
// staggered paired pipeline (slot 1 half a round behind slot 0, biases in shared memory), nf_tc3.cu
const char* nf_tc3_unsupported(const NfPlan& plan);
cudaError_t nf_launch_render_tc3(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, const float* ts,
                                 int T, int64_t ts_stride, const float* noise, const float* ray_time, const nf_mip_args* mip,
                                 float* rgb, float* alpha, float* weights, cudaStream_t st, const NfTrainPlan* train = nullptr, void* ws = nullptr,
                                 const nf_render_aux* aux = nullptr);
// training (nf_tc3.cu TRAIN instantiation + nf_train.cu): transposed weight images, the backward of the MLP chain on tcgen05
const char* nf_train_unsupported(const NfPlan& plan);
cudaError_t nf_launch_pack_all(const NfPlan& plan, const float* const* params, void* packed, cudaStream_t st);      // every Linear's images, one launch
cudaError_t nf_launch_copy_tables(const float* const* src, float* const* dst, int n, size_t bytes, cudaStream_t st);
cudaError_t nf_launch_render_backward(const NfPlan& plan, const NfTrainPlan& tp, const void* packed, void* ws, const float* rays,
                                      const float* ts, int64_t ts_stride, const float* d_rgb, float* const* grads, cudaStream_t st);
cudaError_t nf_launch_integrate(const float* weights, const float* vals, int64_t n_rays, int T, int C, int64_t vals_ray_stride, float* out, cudaStream_t st);
// backward of the non-GEMM stages (nf_bwd.cu)
cudaError_t nf_launch_composite_bwd(const NfPlan& plan, const void* packed, const float* sigma_raw, const float* feats, const float* rays,
                                    int64_t n_rays, const float* ts, int T, int64_t ts_stride, const float* d_rgb, float* d_sigma,
                                    float* d_feats, cudaStream_t st, int feat_act = -1, float* d_beta = nullptr,    // d_beta (NF_DENS_LAPLACE, nullable): ACCUMULATED into
                                    const float* bg_rand = nullptr);                                          // NF_BG_RANDOM: the forward's draws [R]
cudaError_t nf_launch_hash_encode_bwd(const NfPlan& plan, const float* pts, int64_t n, const float* d_feats, float* d_tables, cudaStream_t st);
cudaError_t nf_launch_adam_multi(int n_tensors, float* const* p, const float* const* g, float* const* m, float* const* v, const int64_t* numel,
                                 float lr, float beta1, float beta2, float eps, float wd, int step, cudaStream_t st);
cudaError_t nf_launch_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                                float wd, int step, cudaStream_t st);
