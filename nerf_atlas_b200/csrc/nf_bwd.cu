// nf_bwd.cu -- backward of the two non-GEMM stages of the render path: the first building blocks of the training half
// (the reference differentiates the same ops implicitly through PyTorch autograd, runner.py:820).
//   k_composite_bwd   : d rgb[R,3] -> d sigma_raw[R,T], d feats[R,T,3]        (alpha_from_density + volumetric_integrate + sky,
//                       reference src/nerf.py:22-27,60-80,96-103)
//   k_hash_encode_bwd : d feats[N, 4L] -> d tables[L][table][4] (scatter-add)  (HashEncoder, reference src/neural_blocks.py:139-193)
// Both are HBM/L2-bound stage kernels; the fused backward of the MLP chain is not built yet.
#include <cmath>
#include "nf_common.cuh"
#include "nf_kernels.h"

namespace {

constexpr int BWD_MAX_CHUNKS = 64;     // T <= 2048

// d sigma / d raw density (reference src/nerf.py:64-65; VolSDF: src/utils.py:50-58, src/nerf.py:1000-1003)
__device__ __forceinline__ float dens_act_grad(float d, int kind, float beta) {
  if (kind == NF_DENS_RELU) return d > 0.f ? 1.f : 0.f;
  if (kind == NF_DENS_LAPLACE) {
    const float sc = (-d) / beta;
    const float cdf = sc <= 0.f ? expf(fminf(sc, 0.f)) / 2.f : 1.f - expf(-fmaxf(sc, 0.f)) / 2.f;
    if (!(1.f / beta * cdf > 0.f)) return 0.f;                       // relu
    return -(expf(-fabsf(sc)) / 2.f) / (beta * beta);
  }
  const float x = d - 1.f;                                          // F.softplus(x), threshold 20
  return x > 20.f ? 1.f : 1.f / (1.f + expf(-x));
}

// d sigma / d beta of VolSDF's density (reference src/nerf.py:1000-1003: density = 1 / beta * laplace_cdf(-sdf, beta), then relu):
//   sc = -sdf / beta;  d sigma / d beta = -cdf(sc) / beta^2 - pdf(sc) sc / beta^2,  pdf(sc) = exp(-|sc|) / 2
__device__ __forceinline__ float dens_beta_grad(float d, float beta) {
  const float sc = (-d) / beta;
  const float cdf = sc <= 0.f ? expf(fminf(sc, 0.f)) / 2.f : 1.f - expf(-fmaxf(sc, 0.f)) / 2.f;
  if (!(1.f / beta * cdf > 0.f)) return 0.f;                         // relu
  return -(cdf + (expf(-fabsf(sc)) / 2.f) * sc) / (beta * beta);
}

// d act(v) / d v for the sigmoid family (reference src/utils.py:484-518); `d` = gradient with respect to the activated colours
__device__ __forceinline__ void feat_act3_bwd(float r, float g, float b, int kind, float& dr, float& dg, float& db) {
  if (kind == NF_FEAT_SOFTMAX) {
    float yr = r, yg = g, yb = b; nf_feat_act3(yr, yg, yb, kind);
    const float dot = dr * yr + dg * yg + db * yb;
    dr = yr * (dr - dot); dg = yg * (dg - dot); db = yb * (db - dot);
    return;
  }
  auto one = [kind](float v) -> float {
    switch (kind) {
      case NF_FEAT_NORMAL: case NF_FEAT_UPSHIFTED: { const float s = nf_sigmoid(v); return s * (1.f - s); }
      case NF_FEAT_THIN: { const float s = nf_sigmoid(v); return s * (1.f - s) * (1.f - 2e-2f); }
      case NF_FEAT_FAT:  { const float s = nf_sigmoid(v); return s * (1.f - s) * (1.f + 2e-2f); }
      case NF_FEAT_TANH: { const float t = tanhf(v); return 1.f - t * t; }
      case NF_FEAT_CYCLIC: return cosf(v / 5.f) / 5.f / 2.f * (1.f - 2e-2f);
      case NF_FEAT_LEAKY_RELU: return v > 0.f ? 1.f : 0.01f;
      case NF_FEAT_RELU: case NF_FEAT_UPSHIFTED_RELU: return v > 0.f ? 1.f : 0.f;
      case NF_FEAT_SIN: return cosf(v);
      case NF_FEAT_UPSHIFTED_SOFTPLUS: return v > 20.f ? 1.f : nf_sigmoid(v);
      default: return 1.f;
    }
  };
  dr *= one(r); dg *= one(g); db *= one(b);
}

// Warp per ray.  Pass 1 (forward order): the transmittance entering every 32-sample chunk.  Pass 2 (reverse order):
//   dL/dw_t = A_t = g . f_t - [white bg, t < T-1] (g_r + g_g + g_b)
//   dL/dalpha_t = T_t A_t - (sum_{u>t} w_u A_u) / (1 - alpha_t + 1e-10)
//   d sigma_raw_t = dL/dalpha_t * delta_t (1 - alpha_t) * dens_act'(sigma_raw_t);   d f_t = w_t g
__global__ void k_composite_bwd(int density_act, const float* __restrict__ beta_ptr, int bg, const float* __restrict__ sigma_raw,
                                const float* __restrict__ feats, const float* __restrict__ rays, long long n_rays,
                                const float* __restrict__ ts, int T, long long ts_stride, const float* __restrict__ d_rgb,
                                float* __restrict__ d_sigma, float* __restrict__ d_feats, int feat_act, float* __restrict__ d_beta,
                                const float* __restrict__ bg_rand) {
  // feat_act >= 0: `feats` are the RAW colours (the training stash); the activation is applied here and d_feats is the
  // gradient with respect to the raw values.  feat_act < 0: `feats` are already activated (the stand-alone stage).
  __shared__ float s_carry[8][BWD_MAX_CHUNKS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float beta = beta_ptr ? __ldg(beta_ptr) : 1.f;
  float* carry_in = s_carry[wib];
  const int nchunk = (T + 31) >> 5;
  float acc_beta = 0.f;                                 // d_beta (nullable): sum over this thread's samples of dL/dsigma * d sigma / d beta
  for (long long ray = warp; ray < n_rays; ray += nwarps) {
    const float* r = rays + ray * 6;
    const float dx = __ldg(r + 3), dy = __ldg(r + 4), dz = __ldg(r + 5);
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float* tsr = ts + ray * ts_stride;
    const float gr = __ldg(d_rgb + ray * 3), gg = __ldg(d_rgb + ray * 3 + 1), gb = __ldg(d_rgb + ray * 3 + 2);
    // sky = c (1 - sum_{t<T-1} w_t) added to every channel: c = 1 (white) or the ray's draw (random_color, nerf.py:100-103)
    const float gsky = bg == NF_BG_WHITE ? gr + gg + gb : (bg == NF_BG_RANDOM && bg_rand) ? __ldg(bg_rand + ray) * (gr + gg + gb) : 0.f;
    // pass 1
    float carry = 1.f;
    for (int c = 0; c < nchunk; ++c) {
      const int t = c * 32 + lane;
      float p = 1.f;
      if (t < T) p = (1.f - nf_alpha(__ldg(sigma_raw + ray * T + t), nf_delta(tsr, t, T, nrm), density_act, beta)) + 1e-10f;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, d);
      if (lane == 0) carry_in[c] = carry;
      carry *= p;
    }
    __syncwarp();
    // pass 2
    float suffix = 0.f;                                  // sum of w_u A_u over all later chunks
    for (int c = nchunk - 1; c >= 0; --c) {
      const int t = c * 32 + lane;
      float al = 0.f, delta = 0.f, sr = 0.f, fr = 0.f, fg = 0.f, fb = 0.f;
      if (t < T) {
        sr = __ldg(sigma_raw + ray * T + t);
        delta = nf_delta(tsr, t, T, nrm);
        al = nf_alpha(sr, delta, density_act, beta);
        const float* f = feats + (ray * T + t) * 3;
        fr = __ldg(f); fg = __ldg(f + 1); fb = __ldg(f + 2);
      }
      const float rr = fr, rg = fg, rb = fb;
      if (feat_act >= 0) nf_feat_act3(fr, fg, fb, feat_act);
      float p = t < T ? (1.f - al) + 1e-10f : 1.f;
      const float om = p;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const float q = __shfl_up_sync(0xffffffffu, p, d); if (lane >= d) p *= q; }
      float excl = __shfl_up_sync(0xffffffffu, p, 1); if (lane == 0) excl = 1.f;
      const float trans = carry_in[c] * excl;
      const float w = al * trans;
      const float A = (gr * fr + gg * fg + gb * fb) - ((t < T - 1) ? gsky : 0.f);
      float wa = t < T ? w * A : 0.f;                    // inclusive suffix sum over the chunk, then made exclusive
      float sfx = wa;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const float q = __shfl_down_sync(0xffffffffu, sfx, d); if (lane + d < 32) sfx += q; }
      const float later = (sfx - wa) + suffix;
      if (t < T) {
        const float dalpha = trans * A - later / om;
        const float dsig = dalpha * delta * (1.f - al);              // dL / d sigma
        d_sigma[ray * T + t] = dsig * dens_act_grad(sr, density_act, beta);
        if (d_beta) acc_beta = fmaf(dsig, dens_beta_grad(sr, beta), acc_beta);
        float* df = d_feats + (ray * T + t) * 3;
        float d0 = w * gr, d1 = w * gg, d2 = w * gb;
        if (feat_act >= 0) feat_act3_bwd(rr, rg, rb, feat_act, d0, d1, d2);
        df[0] = d0; df[1] = d1; df[2] = d2;
      }
      suffix += __shfl_sync(0xffffffffu, sfx, 0);
    }
    __syncwarp();
  }
  if (d_beta) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc_beta += __shfl_xor_sync(0xffffffffu, acc_beta, d);
    if (lane == 0 && acc_beta != 0.f) atomicAdd(d_beta, acc_beta);
  }
}

// One thread per (point, level): recompute the 8 corner rows and trilinear weights, scatter-add w_c * d feat into the table.
__global__ void k_hash_encode_bwd(const __grid_constant__ NfPlan plan, const float* __restrict__ pts, long long n,
                                  const float* __restrict__ d_feats, float* __restrict__ d_tables) {
  const int L = plan.hash_levels;
  const long long total = n * L;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / L; const int lvl = (int)(i - p * L);
    const float px = __ldg(pts + p * 3), py = __ldg(pts + p * 3 + 1), pz = __ldg(pts + p * 3 + 2);
    const float res = plan.hash_res[lvl];
    const float vx = __fmul_rn(px, res), vy = __fmul_rn(py, res), vz = __fmul_rn(pz, res);
    const float fx = floorf(vx), fy = floorf(vy), fz = floorf(vz);
    const uint32_t ix = (uint32_t)(int32_t)fx, iy = (uint32_t)(int32_t)fy, iz = (uint32_t)(int32_t)fz;
    const float wx = vx - fx, wy = vy - fy, wz = vz - fz;
    const float4 g = __ldg(reinterpret_cast<const float4*>(d_feats) + i);
    float4* table = reinterpret_cast<float4*>(d_tables) + (size_t)lvl * (plan.hash_mask + 1);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int bx = (c >> 2) & 1, by = (c >> 1) & 1, bz = c & 1;
      const uint32_t id = (((ix + bx) * plan.hash_primes[0]) ^ ((iy + by) * plan.hash_primes[1]) ^ ((iz + bz) * plan.hash_primes[2])) & plan.hash_mask;
      const float w = (bx ? wx : 1.f - wx) * (by ? wy : 1.f - wy) * (bz ? wz : 1.f - wz);
      atomicAdd(table + id, make_float4(w * g.x, w * g.y, w * g.z, w * g.w));
    }
  }
}

// torch.optim.Adam.step for one tensor (the reference's optimiser: runner.py:448-458, eps 1e-7, L2 weight decay added to the
// gradient): exp_avg.lerp_(g, 1-b1); exp_avg_sq = b2*exp_avg_sq + (1-b2) g*g; p -= (lr / bc1) * exp_avg / (sqrt(exp_avg_sq)/sqrt(bc2) + eps).
// HBM-bound: 16 B read + 12 B written per element, float4-vectorised.
__global__ void k_adam_step(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                            float step_size, float beta1, float beta2, float eps, float wd, float inv_bc2_sqrt) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4 + (n & 3); i += (long long)gridDim.x * blockDim.x) {
    if (i < n4) {
      float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
      const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
      float* pp = &P.x; float* mm = &M.x; float* vv = &V.x; const float* gg = &G.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gr = gg[k] + wd * pp[k];
        mm[k] = mm[k] + (1.f - beta1) * (gr - mm[k]);
        vv[k] = beta2 * vv[k] + (1.f - beta2) * gr * gr;
        pp[k] = pp[k] - step_size * (mm[k] / (sqrtf(vv[k]) * inv_bc2_sqrt + eps));
      }
      reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
    } else {
      const long long j = n4 * 4 + (i - n4);
      const float gr = g[j] + wd * p[j];
      m[j] = m[j] + (1.f - beta1) * (gr - m[j]);
      v[j] = beta2 * v[j] + (1.f - beta2) * gr * gr;
      p[j] = p[j] - step_size * (m[j] / (sqrtf(v[j]) * inv_bc2_sqrt + eps));
    }
  }
}

// the same update for up to ADAM_MAX tensors in ONE launch (blockIdx.y = tensor; the reference model has 32 parameter tensors, from
// an 8 MB embedding table to a 12-byte bias: 32 launches of ~4 us each were 0.12 ms of a 4.6 ms optimiser step)
constexpr int ADAM_MAX = 64;
struct AdamArgs { float* p[ADAM_MAX]; const float* g[ADAM_MAX]; float* m[ADAM_MAX]; float* v[ADAM_MAX]; long long n[ADAM_MAX]; };
__global__ void k_adam_multi(const __grid_constant__ AdamArgs a, float step_size, float beta1, float beta2, float eps, float wd, float inv_bc2_sqrt) {
  const int t = blockIdx.y;
  float* __restrict__ p = a.p[t]; const float* __restrict__ g = a.g[t]; float* __restrict__ m = a.m[t]; float* __restrict__ v = a.v[t];
  const long long n = a.n[t], n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4 + (n & 3); i += (long long)gridDim.x * blockDim.x) {
    if (i < n4) {
      float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
      const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
      float* pp = &P.x; float* mm = &M.x; float* vv = &V.x; const float* gg = &G.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gr = gg[k] + wd * pp[k];
        mm[k] = mm[k] + (1.f - beta1) * (gr - mm[k]);
        vv[k] = beta2 * vv[k] + (1.f - beta2) * gr * gr;
        pp[k] = pp[k] - step_size * (mm[k] / (sqrtf(vv[k]) * inv_bc2_sqrt + eps));
      }
      reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
    } else {
      const long long j = n4 * 4 + (i - n4);
      const float gr = g[j] + wd * p[j];
      m[j] = m[j] + (1.f - beta1) * (gr - m[j]);
      v[j] = beta2 * v[j] + (1.f - beta2) * gr * gr;
      p[j] = p[j] - step_size * (m[j] / (sqrtf(v[j]) * inv_bc2_sqrt + eps));
    }
  }
}

int bwd_num_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

}  // namespace

cudaError_t nf_launch_composite_bwd(const NfPlan& plan, const void* packed, const float* sigma_raw, const float* feats, const float* rays,
                                    int64_t n_rays, const float* ts, int T, int64_t ts_stride, const float* d_rgb, float* d_sigma,
                                    float* d_feats, cudaStream_t st, int feat_act, float* d_beta, const float* bg_rand) {
  if (n_rays == 0) return cudaSuccess;
  if (plan.density_act != NF_DENS_LAPLACE) d_beta = nullptr;
  if (T > 32 * BWD_MAX_CHUNKS) return cudaErrorInvalidValue;
  const long long want = (n_rays * 32 + 255) / 256;
  const int grid = (int)(want < (long long)bwd_num_sms() * 8 ? want : (long long)bwd_num_sms() * 8);
  const float* beta = (plan.density_act == NF_DENS_LAPLACE && packed) ? reinterpret_cast<const float*>((const uint8_t*)packed + plan.scale_off) : nullptr;
  k_composite_bwd<<<grid, 256, 0, st>>>(plan.density_act, beta, plan.bg, sigma_raw, feats, rays, n_rays, ts, T, ts_stride, d_rgb, d_sigma, d_feats, feat_act, d_beta, bg_rand);
  return cudaGetLastError();
}

cudaError_t nf_launch_hash_encode_bwd(const NfPlan& plan, const float* pts, int64_t n, const float* d_feats, float* d_tables, cudaStream_t st) {
  const long long total = n * plan.hash_levels;
  if (total == 0) return cudaSuccess;
  const long long want = (total + 255) / 256;
  const int grid = (int)(want < (long long)bwd_num_sms() * 16 ? want : (long long)bwd_num_sms() * 16);
  k_hash_encode_bwd<<<grid, 256, 0, st>>>(plan, pts, n, d_feats, d_tables);
  return cudaGetLastError();
}

cudaError_t nf_launch_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                                float wd, int step, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  const long long want = ((n >> 2) + 3 + 255) / 256;
  const int grid = (int)(want < (long long)bwd_num_sms() * 16 ? want : (long long)bwd_num_sms() * 16);
  k_adam_step<<<grid, 256, 0, st>>>(p, g, m, v, n, (float)(lr / bc1), beta1, beta2, eps, wd, (float)(1.0 / sqrt(bc2)));
  return cudaGetLastError();
}

cudaError_t nf_launch_adam_multi(int n_tensors, float* const* p, const float* const* g, float* const* m, float* const* v, const int64_t* numel,
                                 float lr, float beta1, float beta2, float eps, float wd, int step, cudaStream_t st) {
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  for (int t0 = 0; t0 < n_tensors; t0 += ADAM_MAX) {
    AdamArgs a{}; int cnt = 0; long long max_n = 0;
    for (int t = t0; t < n_tensors && cnt < ADAM_MAX; ++t) {
      if (numel[t] == 0) continue;
      a.p[cnt] = p[t]; a.g[cnt] = g[t]; a.m[cnt] = m[t]; a.v[cnt] = v[t]; a.n[cnt] = numel[t];
      max_n = numel[t] > max_n ? numel[t] : max_n; ++cnt;
    }
    if (cnt == 0) continue;
    const long long want = ((max_n >> 2) + 3 + 255) / 256;
    const int gx = (int)(want < (long long)bwd_num_sms() * 4 ? want : (long long)bwd_num_sms() * 4);
    k_adam_multi<<<dim3(gx, cnt), 256, 0, st>>>(a, (float)(lr / bc1), beta1, beta2, eps, wd, (float)(1.0 / sqrt(bc2)));
  }
  return cudaGetLastError();
}
