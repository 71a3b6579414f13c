// nf_march.cu -- the SDF surface side: sphere tracing of an SDF network and shading of the hit points.
//
//   sphere_march   reference src/march.py:27-47:   t = near;  repeat `iters` times over the rays still active:
//                                                  d = sdf(o + t dir);  hit |= d < eps && t <= far;  t += d;
//                                                  a ray leaves the active set once it has hit or t > far
//   SDF.forward    reference src/sdf.py:137-156:   pts, hit = march(...);  latent = sdf_net(pts[hit])[1:];
//                                                  rgb[hit] = refl(x = pts[hit], view = r_d[hit], latent);  rgb[~hit] = 0
//
// The per-iteration work is one SDF-network evaluation of the ACTIVE rays only (the reference indexes with `rem` the same way).
// It runs as a stream-ordered loop with no host round trip: the active list is compacted on the device and every kernel --
// including the MLP kernels of nf_fp32.cu / nf_tc.cu (tcgen05) -- reads the current count from device memory.
#include "nf_common.cuh"
#include "nf_kernels.h"

namespace {

int mr_num_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

struct MarchWs {          // byte offsets into the caller's workspace
  int64_t t, hit, idx_a, idx_b, x0, out, x1, out1, counts, total;
};
MarchWs march_ws(const NfPlan& p, int64_t R) {
  MarchWs w{};
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  w.counts = take(1024 * 8);                                  // one counter per iteration (+ the hit count), zeroed per call
  w.t = take(R * 4); w.hit = take(R);
  w.idx_a = take(R * 4); w.idx_b = take(R * 4);
  w.x0 = take(R * (int64_t)p.mlp[0].in_dims * 4); w.out = take(R * (int64_t)p.mlp[0].out_dims * 4);
  w.x1 = take(R * (int64_t)p.mlp[1].in_dims * 4); w.out1 = take(R * (int64_t)p.mlp[1].out_dims * 4);
  w.total = off;
  return w;
}

__global__ void k_march_init(long long n, float near, float* __restrict__ t, uint8_t* __restrict__ hit, int* __restrict__ idx, long long* __restrict__ counts) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) { t[i] = near; hit[i] = 0; idx[i] = (int)i; }
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = n;
}

// x0 of the SDF network for the active rays: [p] (SIREN) or [p, sin(p B), cos(p B)] (Fourier-encoded MLP, sdf.py:250-258)
__global__ void k_march_pts(const __grid_constant__ NfPlan plan, const uint8_t* __restrict__ packed, const float* __restrict__ rays,
                            const float* __restrict__ t, const int* __restrict__ idx, const long long* __restrict__ n_dev, float* __restrict__ x0) {
  const long long n = *n_dev;
  const int in_dims = plan.mlp[0].in_dims;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ray = idx[i];
    const float* r = rays + (long long)ray * 6;
    const float tt = t[ray];
    const float px = nf_pt(__ldg(r), tt, __ldg(r + 3)), py = nf_pt(__ldg(r + 1), tt, __ldg(r + 4)), pz = nf_pt(__ldg(r + 2), tt, __ldg(r + 5));
    float* o = x0 + i * in_dims;
    o[0] = px; o[1] = py; o[2] = pz;
    if (plan.enc == NF_ENC_FOURIER) {
      const float* B = reinterpret_cast<const float*>(packed + plan.fourier_off);
      const int F = plan.fourier_freqs;
      for (int f = 0; f < F; ++f) {
        const float m = fmaf(pz, __ldg(B + 2 * F + f), fmaf(py, __ldg(B + F + f), __fmul_rn(px, __ldg(B + f))));
        o[3 + f] = sinf(m); o[3 + F + f] = cosf(m);
      }
    }
  }
}

// march.py:39-45 for the active rays, and the compaction of the survivors
__global__ void k_march_update(const float* __restrict__ out, int out_dims, const float* __restrict__ x0, int in_dims, const int* __restrict__ idx,
                               const long long* __restrict__ n_dev, float eps, float far, float bound_rad, float* __restrict__ t,
                               uint8_t* __restrict__ hit, int* __restrict__ idx_next, long long* __restrict__ n_next) {
  const long long n = *n_dev;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ray = idx[i];
    float dist = out[i * out_dims];
    if (bound_rad > 0.f) {                              // UnitSphere (sdf.py:66-83): max(inner, |p| - rad)
      const float* p = x0 + i * in_dims;
      dist = fmaxf(dist, sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) - bound_rad);
    }
    float cd = t[ray];
    const bool h = hit[ray] != 0 || (dist < eps && cd <= far);
    cd = __fadd_rn(cd, dist);
    t[ray] = cd; hit[ray] = h ? 1 : 0;
    if (!(h || cd > far)) idx_next[atomicAdd(reinterpret_cast<unsigned long long*>(n_next), 1ull)] = ray;
  }
}

__global__ void k_march_finish(const float* __restrict__ rays, long long n, const float* __restrict__ t, const uint8_t* __restrict__ hit,
                               float* __restrict__ pts_out, uint8_t* __restrict__ hit_out, float* __restrict__ t_out,
                               int* __restrict__ idx_hit, long long* __restrict__ n_hit) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float* r = rays + i * 6;
    const float tt = t[i];
    if (pts_out) { pts_out[i * 3] = nf_pt(__ldg(r), tt, __ldg(r + 3)); pts_out[i * 3 + 1] = nf_pt(__ldg(r + 1), tt, __ldg(r + 4)); pts_out[i * 3 + 2] = nf_pt(__ldg(r + 2), tt, __ldg(r + 5)); }
    if (hit_out) hit_out[i] = hit[i];
    if (t_out) t_out[i] = tt;
    if (idx_hit && hit[i]) idx_hit[atomicAdd(reinterpret_cast<unsigned long long*>(n_hit), 1ull)] = (int)i;
  }
}

// x0 of the View head for the hit rays: [p, elaz(view), latent] (refl.py:205-207; the latent is the SDF network's output 1..I)
__global__ void k_shade_x0(const __grid_constant__ NfPlan plan, const float* __restrict__ rays, const float* __restrict__ sdf_x0, const float* __restrict__ sdf_out,
                           const int* __restrict__ idx, const long long* __restrict__ n_dev, float* __restrict__ x1) {
  const long long n = *n_dev;
  const int in0 = plan.mlp[0].in_dims, od = plan.mlp[0].out_dims, in1 = plan.mlp[1].in_dims, I = plan.intermediate;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ray = idx[i];
    const float* r = rays + (long long)ray * 6;
    float el, az;
    nf_elaz(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5), el, az);
    float* o = x1 + i * in1;
    const float* p = sdf_x0 + i * in0;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = el; o[4] = az;
    for (int k = 0; k < I; ++k) o[5 + k] = sdf_out[i * od + 1 + k];
  }
}

__global__ void k_shade_scatter(const float* __restrict__ out1, int od1, int feat_act, const int* __restrict__ idx, const long long* __restrict__ n_dev,
                                float* __restrict__ rgb) {
  const long long n = *n_dev;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float r = out1[i * od1], g = out1[i * od1 + 1], b = out1[i * od1 + 2];
    nf_feat_act3(r, g, b, feat_act);
    float* o = rgb + (long long)idx[i] * 3;
    o[0] = r; o[1] = g; o[2] = b;
  }
}

cudaError_t mlp(const NfPlan& plan, int which, const void* packed, const float* x0, int64_t cap, float* out, int precision, const long long* n_dev, cudaStream_t st) {
  return precision == NF_PREC_FP32 ? nf_launch_mlp_fp32(plan, which, packed, x0, cap, out, st, n_dev)
                                   : nf_launch_mlp_tc(plan, which, packed, x0, cap, out, st, n_dev);
}

}  // namespace

int64_t nf_sdf_workspace_bytes_of(const NfPlan& plan, int64_t n_rays) { return march_ws(plan, n_rays).total; }

// Leaves, in the workspace: t[R], hit[R]; and (for nf_launch_sdf_render) the compacted hit list in idx_a with its count in counts[iters + 1].
static cudaError_t march_core(const NfPlan& plan, const void* packed, const float* rays, int64_t R, float near, float far, int iters, float eps,
                              float bound_rad, int precision, float* pts_out, uint8_t* hit_out, float* t_out, uint8_t* ws, bool want_hit_list, cudaStream_t st) {
  if (iters < 0 || iters > 1000 || R >= (1LL << 31)) return cudaErrorInvalidValue;
  const MarchWs w = march_ws(plan, R);
  float* t = (float*)(ws + w.t); uint8_t* hit = ws + w.hit;
  int* idx[2] = {(int*)(ws + w.idx_a), (int*)(ws + w.idx_b)};
  float* x0 = (float*)(ws + w.x0); float* out = (float*)(ws + w.out);
  long long* counts = (long long*)(ws + w.counts);
  cudaError_t e = cudaMemsetAsync(counts, 0, 1024 * 8, st);
  if (e != cudaSuccess) return e;
  const int sms = mr_num_sms();
  const long long want = (R + 255) / 256;
  const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
  k_march_init<<<grid, 256, 0, st>>>(R, near, t, hit, idx[0], counts);
  for (int i = 0; i < iters; ++i) {
    const int* cur = idx[i & 1]; int* nxt = idx[(i + 1) & 1];
    k_march_pts<<<grid, 256, 0, st>>>(plan, (const uint8_t*)packed, rays, t, cur, counts + i, x0);
    if ((e = mlp(plan, 0, packed, x0, R, out, precision, counts + i, st)) != cudaSuccess) return e;
    k_march_update<<<grid, 256, 0, st>>>(out, plan.mlp[0].out_dims, x0, plan.mlp[0].in_dims, cur, counts + i, eps, far, bound_rad, t, hit, nxt, counts + i + 1);
  }
  k_march_finish<<<grid, 256, 0, st>>>(rays, R, t, hit, pts_out, hit_out, t_out, want_hit_list ? idx[0] : nullptr, counts + iters + 1);
  return cudaGetLastError();
}

cudaError_t nf_launch_sphere_march(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, float near, float far, int iters,
                                   float eps, float bound_rad, int precision, float* pts_out, uint8_t* hit_out, float* t_out, void* ws, cudaStream_t st) {
  if (n_rays == 0) return cudaSuccess;
  return march_core(plan, packed, rays, n_rays, near, far, iters, eps, bound_rad, precision, pts_out, hit_out, t_out, (uint8_t*)ws, false, st);
}

cudaError_t nf_launch_sdf_render(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, float near, float far, int iters,
                                 float eps, float bound_rad, int precision, float* rgb_out, uint8_t* hit_out, float* t_out, float* pts_out,
                                 void* ws_, cudaStream_t st) {
  if (n_rays == 0) return cudaSuccess;
  uint8_t* ws = (uint8_t*)ws_;
  const MarchWs w = march_ws(plan, n_rays);
  cudaError_t e = march_core(plan, packed, rays, n_rays, near, far, iters, eps, bound_rad, precision, pts_out, hit_out, t_out, ws, true, st);
  if (e != cudaSuccess) return e;
  const int* idx_hit = (const int*)(ws + w.idx_a);
  const long long* n_hit = (const long long*)(ws + w.counts) + iters + 1;
  float* x0 = (float*)(ws + w.x0); float* out = (float*)(ws + w.out);
  float* x1 = (float*)(ws + w.x1); float* out1 = (float*)(ws + w.out1);
  const int sms = mr_num_sms();
  const long long want = (n_rays + 255) / 256;
  const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
  if ((e = cudaMemsetAsync(rgb_out, 0, (size_t)n_rays * 3 * sizeof(float), st)) != cudaSuccess) return e;
  // latent = sdf_net(pts[hit])[1:]  (sdf.py:143), then the View head on [pts[hit], elaz(r_d[hit]), latent]
  k_march_pts<<<grid, 256, 0, st>>>(plan, (const uint8_t*)packed, rays, (const float*)(ws + w.t), idx_hit, n_hit, x0);
  if ((e = mlp(plan, 0, packed, x0, n_rays, out, precision, n_hit, st)) != cudaSuccess) return e;
  k_shade_x0<<<grid, 256, 0, st>>>(plan, rays, x0, out, idx_hit, n_hit, x1);
  if ((e = mlp(plan, 1, packed, x1, n_rays, out1, precision, n_hit, st)) != cudaSuccess) return e;
  k_shade_scatter<<<grid, 256, 0, st>>>(out1, plan.mlp[1].out_dims, plan.feat_act, idx_hit, n_hit, rgb_out);
  return cudaGetLastError();
}
