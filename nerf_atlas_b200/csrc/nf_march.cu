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
  int64_t cmin, idxs, lastp, firstn, low, high, slow, shigh, z, todo;      // bisect: throughput_with_sign_change / bisection state
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
  w.cmin = take(R * 4); w.idxs = take(R * 4); w.lastp = take(R * 4); w.firstn = take(R * 4);
  w.low = take(R * 4); w.high = take(R * 4); w.slow = take(R * 4); w.shigh = take(R * 4); w.z = take(R * 4); w.todo = take(R);
  w.total = off;
  return w;
}

__global__ void k_march_init(long long n, float near, float* __restrict__ t, uint8_t* __restrict__ hit, int* __restrict__ idx, long long* __restrict__ counts) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) { t[i] = near; hit[i] = 0; idx[i] = (int)i; }
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = n;
}

// x0 of the SDF network at one position: [p] (SIREN) or [p, sin(p B), cos(p B)] (Fourier-encoded MLP, sdf.py:250-258)
__device__ __forceinline__ void sdf_x0(const NfPlan& plan, const uint8_t* __restrict__ packed, float px, float py, float pz, float* __restrict__ o) {
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
// the network's first output at row i, intersected with the bounding sphere (UnitSphere, sdf.py:66-83: max(inner, |p| - rad))
__device__ __forceinline__ float sdf_val(const float* __restrict__ out, int out_dims, const float* __restrict__ x0, int in_dims, long long i, float bound_rad) {
  float d = out[i * out_dims];
  if (bound_rad > 0.f) { const float* p = x0 + i * in_dims; d = fmaxf(d, sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) - bound_rad); }
  return d;
}
// x0 of the SDF network for the active rays
__global__ void k_march_pts(const __grid_constant__ NfPlan plan, const uint8_t* __restrict__ packed, const float* __restrict__ rays,
                            const float* __restrict__ t, const int* __restrict__ idx, const long long* __restrict__ n_dev, float* __restrict__ x0) {
  const long long n = *n_dev;
  const int in_dims = plan.mlp[0].in_dims;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ray = idx[i];
    const float* r = rays + (long long)ray * 6;
    const float tt = t[ray];
    const float px = nf_pt(__ldg(r), tt, __ldg(r + 3)), py = nf_pt(__ldg(r + 1), tt, __ldg(r + 4)), pz = nf_pt(__ldg(r + 2), tt, __ldg(r + 5));
    sdf_x0(plan, packed, px, py, pz, x0 + i * in_dims);
  }
}

// ---- march.bisect (reference src/march.py:63-110,147-180): every kernel runs over ALL rays, as the reference does ----
// x0 at o + t dir for one t shared by all rays (march.py:93: python double t, cast to fp32 by the multiply); mode 1: o + t added to
// every coordinate (the reference's first sample, march.py:90: `r_o + near`)
__global__ void k_ray_pts_t(const __grid_constant__ NfPlan plan, const uint8_t* __restrict__ packed, const float* __restrict__ rays, long long n,
                            float t, int mode, float* __restrict__ x0) {
  const int in_dims = plan.mlp[0].in_dims;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float* r = rays + i * 6;
    float px, py, pz;
    if (mode) { px = __fadd_rn(__ldg(r), t); py = __fadd_rn(__ldg(r + 1), t); pz = __fadd_rn(__ldg(r + 2), t); }
    else { px = nf_pt(__ldg(r), t, __ldg(r + 3)); py = nf_pt(__ldg(r + 1), t, __ldg(r + 4)); pz = nf_pt(__ldg(r + 2), t, __ldg(r + 5)); }
    sdf_x0(plan, packed, px, py, pz, x0 + i * in_dims);
  }
}
// x0 at o + t[i] dir (per-ray distances: the bracket ends and the midpoints of the bisection); pts_out (nullable) receives the points
__global__ void k_ray_pts_arr(const __grid_constant__ NfPlan plan, const uint8_t* __restrict__ packed, const float* __restrict__ rays, long long n,
                              const float* __restrict__ t, float* __restrict__ x0, float* __restrict__ pts_out) {
  const int in_dims = plan.mlp[0].in_dims;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float* r = rays + i * 6;
    const float tt = t[i];
    const float px = nf_pt(__ldg(r), tt, __ldg(r + 3)), py = nf_pt(__ldg(r + 1), tt, __ldg(r + 4)), pz = nf_pt(__ldg(r + 2), tt, __ldg(r + 5));
    if (x0) sdf_x0(plan, packed, px, py, pz, x0 + i * in_dims);
    if (pts_out) { pts_out[i * 3] = px; pts_out[i * 3 + 1] = py; pts_out[i * 3 + 2] = pz; }
  }
}
// throughput_with_sign_change, one sample (march.py:94-99); it < 0: the first sample (march.py:90-92)
__global__ void k_tput_update(const float* __restrict__ out, int out_dims, const float* __restrict__ x0, int in_dims, float bound_rad, long long n, int it,
                              float* __restrict__ cmin, int* __restrict__ idxs, int* __restrict__ lastp, int* __restrict__ firstn) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float sd = sdf_val(out, out_dims, x0, in_dims, i, bound_rad);
    if (it < 0) { cmin[i] = sd; idxs[i] = 0; lastp[i] = -1; firstn[i] = -1; continue; }
    const float cm = cmin[i];
    if (sd < cm) idxs[i] = it + 1;
    cmin[i] = (sd != sd || cm != cm) ? __int_as_float(0x7fc00000) : fminf(cm, sd);           // torch.minimum propagates NaN
    if (firstn[i] == -1 && sd < 0.f) { lastp[i] = it; firstn[i] = it + 1; }
  }
}
// march.py:100-106: best_pos = o + (near + idx * step) dir; the bracket [last_pos, first_neg] * step (no near offset, -1 -> -step)
__global__ void k_tput_best(const __grid_constant__ NfPlan plan, const uint8_t* __restrict__ packed, const float* __restrict__ rays, long long n,
                            float near, float step, const int* __restrict__ idxs, const int* __restrict__ lastp, const int* __restrict__ firstn,
                            float* __restrict__ x0, float* __restrict__ best_out, float* __restrict__ low, float* __restrict__ high) {
  const int in_dims = plan.mlp[0].in_dims;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float* r = rays + i * 6;
    const float tb = __fadd_rn(near, __fmul_rn((float)idxs[i], step));
    const float px = nf_pt(__ldg(r), tb, __ldg(r + 3)), py = nf_pt(__ldg(r + 1), tb, __ldg(r + 4)), pz = nf_pt(__ldg(r + 2), tb, __ldg(r + 5));
    sdf_x0(plan, packed, px, py, pz, x0 + i * in_dims);
    if (best_out) { best_out[i * 3] = px; best_out[i * 3 + 1] = py; best_out[i * 3 + 2] = pz; }
    low[i] = __fmul_rn((float)lastp[i], step); high[i] = __fmul_rn((float)firstn[i], step);
  }
}
// copies the SDF value of every row (which = 0: throughput -> tput_out and hit = tput < 0; 1: sdf_low; 2: sdf_high + the bisection's
// initial state, march.py:160-162)
__global__ void k_bis_store(const float* __restrict__ out, int out_dims, const float* __restrict__ x0, int in_dims, float bound_rad, long long n, int which,
                            float eps, float* __restrict__ tput_out, uint8_t* __restrict__ hit, float* __restrict__ slow, float* __restrict__ shigh,
                            const float* __restrict__ low, const float* __restrict__ high, float* __restrict__ z, uint8_t* __restrict__ todo) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float sd = sdf_val(out, out_dims, x0, in_dims, i, bound_rad);
    if (which == 0) { if (tput_out) tput_out[i] = sd; hit[i] = sd < 0.f ? 1 : 0; }
    else if (which == 1) slow[i] = sd;
    else {
      shigh[i] = sd;
      const float lo = low[i], hi = high[i];
      todo[i] = (__fsub_rn(hi, lo) > eps && slow[i] > 0.f && sd < 0.f && hi > lo) ? 1 : 0;
      z[i] = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
    }
  }
}
// one bisection step (march.py:165-179)
__global__ void k_bis_update(const float* __restrict__ out, int out_dims, const float* __restrict__ x0, int in_dims, float bound_rad, long long n, float eps,
                             float* __restrict__ low, float* __restrict__ high, float* __restrict__ slow, float* __restrict__ shigh,
                             float* __restrict__ z, uint8_t* __restrict__ todo) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (!todo[i]) continue;                                   // z = (low + high) / 2 of an untouched ray does not change
    const float sm = sdf_val(out, out_dims, x0, in_dims, i, bound_rad);
    float lo = low[i], hi = high[i], sl = slow[i], sh = shigh[i];
    const float zp = z[i];
    if (sm > 0.f) { lo = zp; sl = sm; }
    if (sm < 0.f) { hi = zp; sh = sm; }
    low[i] = lo; high[i] = hi; slow[i] = sl; shigh[i] = sh;
    z[i] = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
    todo[i] = (__fsub_rn(hi, lo) > eps && sl > 0.f && sh < 0.f && hi > lo) ? 1 : 0;
  }
}
// the hit list of an explicit hit mask, and x0 of the SDF network at explicit points for a compacted list
__global__ void k_compact_hits(const uint8_t* __restrict__ hit, long long n, int* __restrict__ idx_hit, long long* __restrict__ n_hit) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (hit[i]) idx_hit[atomicAdd(reinterpret_cast<unsigned long long*>(n_hit), 1ull)] = (int)i;
}
__global__ void k_pts_x0(const __grid_constant__ NfPlan plan, const uint8_t* __restrict__ packed, const float* __restrict__ pts, const int* __restrict__ idx,
                         const long long* __restrict__ n_dev, float* __restrict__ x0) {
  const long long n = *n_dev;
  const int in_dims = plan.mlp[0].in_dims;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float* p = pts + (long long)idx[i] * 3;
    sdf_x0(plan, packed, __ldg(p), __ldg(p + 1), __ldg(p + 2), x0 + i * in_dims);
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

// rgb[hit] = act(View([pts[hit], elaz(r_d[hit]), sdf_net(pts[hit])[1:]])), rgb[~hit] = 0 (sdf.py:143-153).  The hit list is idx_hit[0, *n_hit);
// the hit points are o + t dir (t_arr, sphere march) or explicit (pts, bisect).
static cudaError_t shade_core(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, const float* t_arr, const float* pts,
                              const int* idx_hit, const long long* n_hit, int precision, float* rgb_out, uint8_t* ws, cudaStream_t st) {
  const MarchWs w = march_ws(plan, n_rays);
  float* x0 = (float*)(ws + w.x0); float* out = (float*)(ws + w.out);
  float* x1 = (float*)(ws + w.x1); float* out1 = (float*)(ws + w.out1);
  const int sms = mr_num_sms();
  const long long want = (n_rays + 255) / 256;
  const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
  cudaError_t e;
  if ((e = cudaMemsetAsync(rgb_out, 0, (size_t)n_rays * 3 * sizeof(float), st)) != cudaSuccess) return e;
  if (pts) k_pts_x0<<<grid, 256, 0, st>>>(plan, (const uint8_t*)packed, pts, idx_hit, n_hit, x0);
  else k_march_pts<<<grid, 256, 0, st>>>(plan, (const uint8_t*)packed, rays, t_arr, idx_hit, n_hit, x0);
  if ((e = mlp(plan, 0, packed, x0, n_rays, out, precision, n_hit, st)) != cudaSuccess) return e;
  k_shade_x0<<<grid, 256, 0, st>>>(plan, rays, x0, out, idx_hit, n_hit, x1);
  if ((e = mlp(plan, 1, packed, x1, n_rays, out1, precision, n_hit, st)) != cudaSuccess) return e;
  k_shade_scatter<<<grid, 256, 0, st>>>(out1, plan.mlp[1].out_dims, plan.feat_act, idx_hit, n_hit, rgb_out);
  return cudaGetLastError();
}

cudaError_t nf_launch_sdf_render(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, float near, float far, int iters,
                                 float eps, float bound_rad, int precision, float* rgb_out, uint8_t* hit_out, float* t_out, float* pts_out,
                                 void* ws_, cudaStream_t st) {
  if (n_rays == 0) return cudaSuccess;
  uint8_t* ws = (uint8_t*)ws_;
  const MarchWs w = march_ws(plan, n_rays);
  cudaError_t e = march_core(plan, packed, rays, n_rays, near, far, iters, eps, bound_rad, precision, pts_out, hit_out, t_out, ws, true, st);
  if (e != cudaSuccess) return e;
  return shade_core(plan, packed, rays, n_rays, (const float*)(ws + w.t), nullptr, (const int*)(ws + w.idx_a), (const long long*)(ws + w.counts) + iters + 1,
                    precision, rgb_out, ws, st);
}

// march.bisect (reference src/march.py:63-75): throughput_with_sign_change over `iters` + 1 samples (march.py:78-110; `jitter` = its
// random.random() draw), then min(32, iters) bisection steps between the samples around the first sign change (march.py:147-180).
// Outputs (nullable except hit): pts_out[R,3], hit_out[R] = tput < 0, tput_out[R], best_out[R,3] (the sample with the smallest SDF).
// rgb_out != nullptr: also SDF.forward's shading of the hits (sdf.py:143-153).
cudaError_t nf_launch_sdf_bisect(const NfPlan& plan, const void* packed, const float* rays, int64_t R, float near, float far, int iters, float jitter,
                                 float bound_rad, int precision, float* pts_out, uint8_t* hit_out, float* tput_out, float* best_out, float* rgb_out,
                                 void* ws_, cudaStream_t st) {
  if (R == 0) return cudaSuccess;
  if (iters < 1 || iters > 1000 || R >= (1LL << 31)) return cudaErrorInvalidValue;
  uint8_t* ws = (uint8_t*)ws_;
  const MarchWs w = march_ws(plan, R);
  float* x0 = (float*)(ws + w.x0); float* out = (float*)(ws + w.out);
  float* cmin = (float*)(ws + w.cmin); int* idxs = (int*)(ws + w.idxs); int* lastp = (int*)(ws + w.lastp); int* firstn = (int*)(ws + w.firstn);
  float* low = (float*)(ws + w.low); float* high = (float*)(ws + w.high); float* slow = (float*)(ws + w.slow); float* shigh = (float*)(ws + w.shigh);
  float* z = (float*)(ws + w.z); uint8_t* todo = ws + w.todo; uint8_t* hit = ws + w.hit;
  long long* counts = (long long*)(ws + w.counts);
  const int od = plan.mlp[0].out_dims, id = plan.mlp[0].in_dims;
  const int sms = mr_num_sms();
  const long long want = (R + 255) / 256;
  const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
  cudaError_t e;
  // python arithmetic of march.py:86-87,93 in double; tensors see the fp32 casts
  const double max_t = (double)far - (double)near + (double)jitter * (2.0 / iters), step = max_t / iters;
  k_ray_pts_t<<<grid, 256, 0, st>>>(plan, (const uint8_t*)packed, rays, R, near, 1, x0);
  if ((e = mlp(plan, 0, packed, x0, R, out, precision, nullptr, st)) != cudaSuccess) return e;
  k_tput_update<<<grid, 256, 0, st>>>(out, od, x0, id, bound_rad, R, -1, cmin, idxs, lastp, firstn);
  for (int i = 0; i < iters; ++i) {
    k_ray_pts_t<<<grid, 256, 0, st>>>(plan, (const uint8_t*)packed, rays, R, (float)((double)near + step * (i + 1)), 0, x0);
    if ((e = mlp(plan, 0, packed, x0, R, out, precision, nullptr, st)) != cudaSuccess) return e;
    k_tput_update<<<grid, 256, 0, st>>>(out, od, x0, id, bound_rad, R, i, cmin, idxs, lastp, firstn);
  }
  k_tput_best<<<grid, 256, 0, st>>>(plan, (const uint8_t*)packed, rays, R, near, (float)step, idxs, lastp, firstn, x0, best_out, low, high);
  if ((e = mlp(plan, 0, packed, x0, R, out, precision, nullptr, st)) != cudaSuccess) return e;
  k_bis_store<<<grid, 256, 0, st>>>(out, od, x0, id, bound_rad, R, 0, 0.f, tput_out, hit, slow, shigh, low, high, z, todo);
  // bisection(near = last_pos, far = first_neg, iters = min(32, iters), eps = 1e-6)
  const float beps = 1e-6f;
  k_ray_pts_arr<<<grid, 256, 0, st>>>(plan, (const uint8_t*)packed, rays, R, low, x0, nullptr);
  if ((e = mlp(plan, 0, packed, x0, R, out, precision, nullptr, st)) != cudaSuccess) return e;
  k_bis_store<<<grid, 256, 0, st>>>(out, od, x0, id, bound_rad, R, 1, beps, nullptr, hit, slow, shigh, low, high, z, todo);
  k_ray_pts_arr<<<grid, 256, 0, st>>>(plan, (const uint8_t*)packed, rays, R, high, x0, nullptr);
  if ((e = mlp(plan, 0, packed, x0, R, out, precision, nullptr, st)) != cudaSuccess) return e;
  k_bis_store<<<grid, 256, 0, st>>>(out, od, x0, id, bound_rad, R, 2, beps, nullptr, hit, slow, shigh, low, high, z, todo);
  const int bit = iters < 32 ? iters : 32;
  for (int i = 0; i < bit; ++i) {
    k_ray_pts_arr<<<grid, 256, 0, st>>>(plan, (const uint8_t*)packed, rays, R, z, x0, nullptr);
    if ((e = mlp(plan, 0, packed, x0, R, out, precision, nullptr, st)) != cudaSuccess) return e;
    k_bis_update<<<grid, 256, 0, st>>>(out, od, x0, id, bound_rad, R, beps, low, high, slow, shigh, z, todo);
  }
  // pts = o + z dir; always into the workspace's x1-free area?  no: pts_out may be null, the shading needs the points -> use t = z
  k_ray_pts_arr<<<grid, 256, 0, st>>>(plan, (const uint8_t*)packed, rays, R, z, nullptr, pts_out);
  if (hit_out && (e = cudaMemcpyAsync(hit_out, hit, (size_t)R, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
  if (rgb_out) {
    if ((e = cudaMemsetAsync(counts, 0, 8, st)) != cudaSuccess) return e;
    k_compact_hits<<<grid, 256, 0, st>>>(hit, R, (int*)(ws + w.idx_a), counts);
    return shade_core(plan, packed, rays, R, z, nullptr, (const int*)(ws + w.idx_a), counts, precision, rgb_out, ws, st);   // hit points = o + z dir
  }
  return cudaGetLastError();
}
