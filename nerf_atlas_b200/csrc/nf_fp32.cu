// nf_fp32.cu -- the exact (fp32 CUDA-core) form of the fused render pipeline, plus the
// stand-alone stage kernels (sample points, hash encode, composite) used by parity tests
// and HBM-bound micro-benchmarks.  One CTA owns a tile of 64 consecutive samples and keeps
// every activation in shared memory: rays in, RGB (+ optional alpha/weights) out.
#include "nf_common.cuh"
#include "nf_kernels.h"

namespace {

constexpr int ROWS = 64;          // samples per tile
constexpr int THREADS = 256;      // 16 row-groups (4 rows) x 16 col-groups (16 cols)
constexpr int X0_MAX = 272;

struct Fp32Smem {
  float H[2][NF_HIDDEN * ROWS];   // [k][row], ping-pong between layers
  float X0[X0_MAX * ROWS];        // [k][row], raw inputs of the current MLP
  float P[3 * ROWS];              // sample positions
  float M[NF_MIP_FEATS * ROWS];   // Mip IPE latent of the tile (appended to both MLP inputs)
  float sig[ROWS];                // raw density
  float carry[8];                 // T > ROWS: transmittance, rgb, sum w (before last) carried across sub-tiles
  long long ray[ROWS];
  int t[ROWS];
  int valid[ROWS];
};

// One Linear of SkipConnMLP (reference src/neural_blocks.py:289-296): Hout = [act?](W [act(Hin), act?(X0)] + b)
// JAC (forward-mode derivative, nf_sdf_normals): a thread's four rows are ONE point -- row 0 its value, rows 1-3 the tangents
// d/dp_x, d/dp_y, d/dp_z.  The same FMAs propagate all four (W is linear); the bias goes to the value row only and the activation
// becomes act(z) for the value, act'(z) * tangent for the other three.
template <bool JAC = false>
__device__ __forceinline__ void linear_fp32(const NfLinPlan& L, int act, const uint8_t* __restrict__ packed,
                                            const float* __restrict__ Hin, const float* __restrict__ X0,
                                            float* __restrict__ Hout) {
  const int tid = threadIdx.x, cg = tid & 15, rg = tid >> 4;
  const int n0 = cg * 16, r0 = rg * 4;
  if (n0 < L.n_pad) {
    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
    const float* __restrict__ Wt = reinterpret_cast<const float*>(packed + L.wt_off) + n0;
    const int np = L.n_pad;
    auto fma_row = [&](const float4 a, const float* __restrict__ wrow) {
      float w[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(wrow) + q);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(av[i], w[j], acc[i][j]);
    };
#pragma unroll 4
    for (int k = 0; k < L.k_hidden; ++k)
      fma_row(*reinterpret_cast<const float4*>(Hin + k * ROWS + r0), Wt + (size_t)k * np);
    const float* __restrict__ Wx = Wt + (size_t)L.k_hidden * np;
    for (int k = 0; k < L.k_x0; ++k) {
      float4 a = *reinterpret_cast<const float4*>(X0 + k * ROWS + r0);
      if (!L.x0_raw) {
        if (JAC) { const float d = nf_act_grad(a.x, act); a.x = nf_apply_act(a.x, act); a.y *= d; a.z *= d; a.w *= d; }
        else { a.x = nf_apply_act(a.x, act); a.y = nf_apply_act(a.y, act); a.z = nf_apply_act(a.z, act); a.w = nf_apply_act(a.w, act); }
      }
      fma_row(a, Wx + (size_t)k * np);
    }
    const float* __restrict__ b = reinterpret_cast<const float*>(packed + L.b_off) + n0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float bj = __ldg(b + j);
      float4 o;
      if (JAC) {
        const float z = acc[0][j] + bj, d = L.is_out ? 1.f : nf_act_grad(z, act);
        o.x = L.is_out ? z : nf_apply_act(z, act); o.y = acc[1][j] * d; o.z = acc[2][j] * d; o.w = acc[3][j] * d;
      } else {
      o.x = acc[0][j] + bj; o.y = acc[1][j] + bj; o.z = acc[2][j] + bj; o.w = acc[3][j] + bj;
      if (!L.is_out) { o.x = nf_apply_act(o.x, act); o.y = nf_apply_act(o.y, act); o.z = nf_apply_act(o.z, act); o.w = nf_apply_act(o.w, act); }
      }
      *reinterpret_cast<float4*>(Hout + (n0 + j) * ROWS + r0) = o;
    }
  }
  __syncthreads();
}

// Runs every Linear of one MLP; returns the H buffer index holding out[n][row].
template <bool JAC = false>
__device__ __forceinline__ int mlp_fp32(const NfMlpPlan& M, const uint8_t* __restrict__ packed, Fp32Smem& s) {
  int cur = 0;
  for (int j = 0; j < M.n_lin; ++j) {
    linear_fp32<JAC>(M.lin[j], M.act, packed, s.H[cur], s.X0, s.H[cur ^ 1]);
    cur ^= 1;
  }
  return cur;
}

struct RenderArgs {
  const uint8_t* packed;
  const float* rays; long long n_rays;
  const float* ts; int T; long long ts_stride;
  const float* noise; const float* ray_time;
  float* rgb_out; float* alpha_out; float* weights_out;
  NfMipIn mip;
  const float* pts; const float* bg_rand;                                    // nf_render_aux: from_pts, random background
  float* pts_out; float* dp_out; float* rigid_dp_out; float* rigidity_out;   // nf_render_aux: DynamicNeRF side channels
};

__global__ void __launch_bounds__(THREADS, 1)
k_render_fp32(const __grid_constant__ NfPlan plan, const RenderArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Fp32Smem& s = *reinterpret_cast<Fp32Smem*>(smem_raw);
  const NfTileMap map(a.T, ROWS);
  const long long units = map.units(a.n_rays);
  const int tid = threadIdx.x;
  const int out_ch = 3;
  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    for (int sub = 0; sub < map.tpr; ++sub) {
      // ---- stage 0: sample positions (reference nerf.py:50-55)
      {
        const int row = tid % ROWS, part = tid / ROWS;
        if (part == 0) {
          long long ray; int t;
          const bool ok = map.locate(u, sub, row, a.n_rays, ray, t);
          float px = 0.f, py = 0.f, pz = 0.f;
          if (ok) {
            const float* r = a.rays + ray * 6;
            const float tt = __ldg(a.ts + ray * a.ts_stride + t);
            px = nf_pt(__ldg(r + 0), tt, __ldg(r + 3)); py = nf_pt(__ldg(r + 1), tt, __ldg(r + 4)); pz = nf_pt(__ldg(r + 2), tt, __ldg(r + 5));
            if (a.pts) { const float* pp = a.pts + (ray * a.T + t) * 3; px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2); }   // from_pts
            if (a.pts_out) { float* o = a.pts_out + (ray * a.T + t) * 3; o[0] = px; o[1] = py; o[2] = pz; }
          }
          s.ray[row] = ray; s.t[row] = t; s.valid[row] = ok;
          s.P[row] = px; s.P[ROWS + row] = py; s.P[2 * ROWS + row] = pz;
        }
      }
      __syncthreads();
      if (plan.kind == NF_KIND_DYN) {
        // ---- stage 0b: deformation (reference nerf.py:1261-1278,1292-1303)
        //   direct: delta_estim([p, t]) -> (dp[1], rigidity[3]);                       p' = p + dp * sigmoid(rigidity / 2)
        //   spline: delta_estim([p, hash(p)]) -> (rigidity[1], n control points[3]);    p' = p + Bezier(points, t) * sigmoid(rigidity / 2)
        if (plan.spline_points == 0) {
          if (tid < ROWS) {
            const int row = tid;
            s.X0[0 * ROWS + row] = s.P[row]; s.X0[1 * ROWS + row] = s.P[ROWS + row]; s.X0[2 * ROWS + row] = s.P[2 * ROWS + row];
            s.X0[3 * ROWS + row] = s.valid[row] ? __ldg(a.ray_time + s.ray[row]) : 0.f;
          }
        } else {
          const int row = tid % ROWS, part = tid / ROWS;
          const float px = s.P[row], py = s.P[ROWS + row], pz = s.P[2 * ROWS + row];
          if (part == 0) {
            s.X0[0 * ROWS + row] = px; s.X0[1 * ROWS + row] = py; s.X0[2 * ROWS + row] = pz;
            s.X0[3 * ROWS + row] = px; s.X0[4 * ROWS + row] = py; s.X0[5 * ROWS + row] = pz;
          }
          const float4* tables = reinterpret_cast<const float4*>(a.packed + plan.hash2_off);
          for (int lvl = part; lvl < plan.hash_levels; lvl += THREADS / ROWS) {
            const float4 f = nf_hash_level(tables + (size_t)lvl * (plan.hash_mask + 1), px, py, pz, plan.hash_res[lvl],
                                           plan.hash_primes[0], plan.hash_primes[1], plan.hash_primes[2], plan.hash_mask, nullptr);
            float* x = s.X0 + (6 + lvl * 4) * ROWS + row;
            x[0] = f.x; x[ROWS] = f.y; x[2 * ROWS] = f.z; x[3 * ROWS] = f.w;
          }
        }
        __syncthreads();
        const int ob = mlp_fp32(plan.mlp[2], a.packed, s);
        if (tid < ROWS) {
          const float* O = s.H[ob]; const int row = tid;
          const long long sidx = s.valid[row] ? s.ray[row] * a.T + s.t[row] : -1;
          if (plan.spline_points == 0) {
            const float dp = O[row];
            const float r0 = nf_sigmoid(O[1 * ROWS + row] / 2.f), r1 = nf_sigmoid(O[2 * ROWS + row] / 2.f), r2 = nf_sigmoid(O[3 * ROWS + row] / 2.f);
            s.P[row] += dp * r0; s.P[ROWS + row] += dp * r1; s.P[2 * ROWS + row] += dp * r2;
            if (sidx >= 0) {
              if (a.dp_out) a.dp_out[sidx] = dp;
              if (a.rigidity_out) { float* o = a.rigidity_out + sidx * 3; o[0] = r0; o[1] = r1; o[2] = r2; }
              if (a.rigid_dp_out) { float* o = a.rigid_dp_out + sidx * 3; o[0] = dp * r0; o[1] = dp * r1; o[2] = dp * r2; }
            }
          } else {
            const int n = plan.spline_points;
            const float rig = nf_sigmoid(O[row] / 2.f);
            const float tt = s.valid[row] ? __ldg(a.ray_time + s.ray[row]) : 0.f;
#pragma unroll
            for (int x = 0; x < 3; ++x) {
              float ps[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) ps[i] = i < n ? O[(1 + 3 * i + x) * ROWS + row] : 0.f;
              const float d = nf_bezier(ps, n, tt);
              s.P[x * ROWS + row] += d * rig;
              if (sidx >= 0) {
                if (a.dp_out) a.dp_out[sidx * 3 + x] = d;
                if (a.rigid_dp_out) a.rigid_dp_out[sidx * 3 + x] = d * rig;
              }
            }
            if (sidx >= 0 && a.rigidity_out) a.rigidity_out[sidx] = rig;
          }
        }
        __syncthreads();
      }
      // ---- stage 0c: encode -> X0 (reference neural_blocks.py:139-193 / 36-55)
      {
        const int row = tid % ROWS, part = tid / ROWS;  // 4 threads per sample, levels / frequencies interleaved
        const float px = s.P[row], py = s.P[ROWS + row], pz = s.P[2 * ROWS + row];
        if (part == 0) {
          s.X0[0 * ROWS + row] = px; s.X0[1 * ROWS + row] = py; s.X0[2 * ROWS + row] = pz;
          if (plan.enc == NF_ENC_HASH) { s.X0[3 * ROWS + row] = px; s.X0[4 * ROWS + row] = py; s.X0[5 * ROWS + row] = pz; }
        }
        if (plan.enc == NF_ENC_FOURIER) {
          // x0 = [p, sin(p B), cos(p B)], B = basis[3][F] (reference src/neural_blocks.py:36-55, src/utils.py:14-17)
          const float* B = reinterpret_cast<const float*>(a.packed + plan.fourier_off);
          const int F = plan.fourier_freqs;
          for (int f = part; f < F; f += THREADS / ROWS) {
            const float m = fmaf(pz, __ldg(B + 2 * F + f), fmaf(py, __ldg(B + F + f), __fmul_rn(px, __ldg(B + f))));
            s.X0[(3 + f) * ROWS + row] = sinf(m); s.X0[(3 + F + f) * ROWS + row] = cosf(m);
          }
        }
        if (plan.enc == NF_ENC_HASH) {
          const float4* tables = reinterpret_cast<const float4*>(a.packed + plan.hash_off);
          for (int lvl = part; lvl < plan.hash_levels; lvl += THREADS / ROWS) {
            const float4 f = nf_hash_level(tables + (size_t)lvl * (plan.hash_mask + 1), px, py, pz, plan.hash_res[lvl],
                                           plan.hash_primes[0], plan.hash_primes[1], plan.hash_primes[2], plan.hash_mask, nullptr);
            float* x = s.X0 + (6 + lvl * 4) * ROWS + row;
            x[0] = f.x; x[ROWS] = f.y; x[2 * ROWS] = f.z; x[3 * ROWS] = f.w;
          }
        }
        if (plan.mip != NF_MIP_NONE) {
          // Mip IPE of the UNdeformed ray segment (nerf.py:340: mip_encoding(r_o, r_d, ts)), kept in s.M for the View head
          const int base = plan.mlp[0].in_dims - NF_MIP_FEATS;
          for (int cc = part; cc < NF_MIP_FEATS / 2; cc += THREADS / ROWS) {
            float fs = 0.f, fc = 0.f;
            if (s.valid[row]) nf_mip_feature_pair(a.mip, s.ray[row], s.t[row], cc, fs, fc);
            s.M[cc * ROWS + row] = fs; s.X0[(base + cc) * ROWS + row] = fs;
            s.M[(cc + 48) * ROWS + row] = fc; s.X0[(base + cc + 48) * ROWS + row] = fc;
          }
        }
      }
      __syncthreads();
      // ---- stage 1: density MLP
      int ob = mlp_fp32(plan.mlp[0], a.packed, s);
      const float* rgb_raw;
      if (plan.kind == NF_KIND_PLAIN || plan.kind == NF_KIND_DYN) {
        // ---- glue: sigma_raw, x0 of the RGB head (nerf.py:344-358): View = [pts, elaz(view), (mip), intermediate] (refl.py:205-207),
        //      Positional = [pts, pts, hash'(pts), (mip), intermediate] with the head's own tables (refl.py:230-245)
        const float* O = s.H[ob];
        const int ml = plan.mip != NF_MIP_NONE ? NF_MIP_FEATS : 0;
        const bool pos_head = plan.refl_kind != NF_REFL_VIEW;          // Positional, and PosLinearView's `pos` MLP: the same x0
        const int base = pos_head ? 6 + plan.hash_levels * 4 : 5;
        if (tid < ROWS) {
          const int row = tid;
          s.sig[row] = O[row];
          s.X0[0 * ROWS + row] = s.P[row]; s.X0[1 * ROWS + row] = s.P[ROWS + row]; s.X0[2 * ROWS + row] = s.P[2 * ROWS + row];
          if (pos_head) {
            s.X0[3 * ROWS + row] = s.P[row]; s.X0[4 * ROWS + row] = s.P[ROWS + row]; s.X0[5 * ROWS + row] = s.P[2 * ROWS + row];
          } else {
            float el = 0.f, az = 0.f;
            if (s.valid[row]) { const float* r = a.rays + s.ray[row] * 6; nf_elaz(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5), el, az); }
            s.X0[3 * ROWS + row] = el; s.X0[4 * ROWS + row] = az;
          }
        }
        if (pos_head) {
          const int row = tid % ROWS, part = tid / ROWS;
          const float4* tables = reinterpret_cast<const float4*>(a.packed + plan.hash3_off);
          for (int lvl = part; lvl < plan.hash_levels; lvl += THREADS / ROWS) {
            const float4 f = nf_hash_level(tables + (size_t)lvl * (plan.hash_mask + 1), s.P[row], s.P[ROWS + row], s.P[2 * ROWS + row], plan.hash_res[lvl],
                                           plan.hash_primes[0], plan.hash_primes[1], plan.hash_primes[2], plan.hash_mask, nullptr);
            float* x = s.X0 + (6 + lvl * 4) * ROWS + row;
            x[0] = f.x; x[ROWS] = f.y; x[2 * ROWS] = f.z; x[3 * ROWS] = f.w;
          }
        }
        for (int i = tid; i < ml * ROWS; i += THREADS) s.X0[base * ROWS + i] = s.M[i];
        for (int i = tid; i < plan.intermediate * ROWS; i += THREADS) s.X0[(base + ml) * ROWS + i] = O[ROWS + i];
        __syncthreads();
        // ---- stage 2: View MLP
        ob = mlp_fp32(plan.mlp[1], a.packed, s);
        rgb_raw = s.H[ob];
        if (plan.refl_kind == NF_REFL_POSLINVIEW) {
          // PosLinearView (refl.py:281-290): [pos(3), im] = act(pos_mlp(...)); linear = sigmoid(view_mlp([p, dir, latent, im])) / 2 + 1/2;
          // rgb = linear * pos.  The density MLP's intermediate (the latent) still sits in X0 behind [p, p, hash'(p)].
          const int I = plan.intermediate, im = plan.mlp[1].out_dims - 3, lat0 = 6 + plan.hash_levels * 4;
          float* T = s.H[ob ^ 1];                                   // free buffer: [latent (I) | act(pos) (3 + im)]
          const float* O2 = s.H[ob];
          for (int i = tid; i < I * ROWS; i += THREADS) T[i] = s.X0[lat0 * ROWS + i];
          for (int i = tid; i < (3 + im) * ROWS; i += THREADS) T[I * ROWS + i] = nf_feat_act_fn(O2[i], plan.feat_act);
          __syncthreads();
          if (tid < ROWS) {
            const int row = tid;
            float dx = 0.f, dy = 0.f, dz = 0.f;
            if (s.valid[row]) {
              const float* r = a.rays + s.ray[row] * 6;
              dx = __ldg(r + 3); dy = __ldg(r + 4); dz = __ldg(r + 5);
              const float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), 1e-12f);   // F.normalize
              dx = __fdiv_rn(dx, nrm); dy = __fdiv_rn(dy, nrm); dz = __fdiv_rn(dz, nrm);
            }
            s.X0[0 * ROWS + row] = s.P[row]; s.X0[1 * ROWS + row] = s.P[ROWS + row]; s.X0[2 * ROWS + row] = s.P[2 * ROWS + row];
            s.X0[3 * ROWS + row] = dx; s.X0[4 * ROWS + row] = dy; s.X0[5 * ROWS + row] = dz;
          }
          for (int i = tid; i < I * ROWS; i += THREADS) s.X0[6 * ROWS + i] = T[i];
          for (int i = tid; i < im * ROWS; i += THREADS) s.X0[(6 + I) * ROWS + i] = T[(I + 3) * ROWS + i];
          __syncthreads();
          // the three colours leave T before the view MLP reuses the H buffers: park them behind x0
          float* C = s.X0 + (6 + I + im) * ROWS;
          for (int i = tid; i < 3 * ROWS; i += THREADS) C[i] = T[I * ROWS + i];
          __syncthreads();
          const int ov = mlp_fp32(plan.mlp[2], a.packed, s);
          float* Rg = s.H[ov ^ 1];
          if (tid < ROWS) {
            const float lin = nf_sigmoid(s.H[ov][tid]) / 2.f + 0.5f;
            Rg[tid] = __fmul_rn(lin, C[tid]); Rg[ROWS + tid] = __fmul_rn(lin, C[ROWS + tid]); Rg[2 * ROWS + tid] = __fmul_rn(lin, C[2 * ROWS + tid]);
          }
          __syncthreads();
          rgb_raw = Rg;
        }
      } else {
        const float* O = s.H[ob];
        if (tid < ROWS) s.sig[tid] = O[tid];
        rgb_raw = O + ROWS;
        __syncthreads();
      }
      // ---- stage 3: alpha_from_density + volumetric_integrate (nerf.py:60-80), sequential per ray
      const int nseg = a.T <= ROWS ? map.rpt : 1;
      if (tid < nseg) {
        const int row0 = a.T <= ROWS ? tid * a.T : 0;
        if (s.valid[row0]) {
          const long long ray = s.ray[row0];
          const float* r = a.rays + ray * 6;
          const float dx = __ldg(r + 3), dy = __ldg(r + 4), dz = __ldg(r + 5);
          const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
          const float* tsr = a.ts + ray * a.ts_stride;
          const float beta = plan.density_act == NF_DENS_LAPLACE ? __ldg(reinterpret_cast<const float*>(a.packed + plan.scale_off)) : 1.f;
          float trans = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, wsum = 0.f;
          if (sub > 0) { trans = s.carry[0]; cr = s.carry[1]; cg = s.carry[2]; cb = s.carry[3]; wsum = s.carry[4]; }
          const int nrow = a.T <= ROWS ? a.T : min(ROWS, a.T - sub * ROWS);
          for (int i = 0; i < nrow; ++i) {
            const int row = row0 + i, t = s.t[row];
            float sr = s.sig[row];
            if (a.noise) sr += __ldg(a.noise + ray * a.T + t);
            const float al = nf_alpha(sr, nf_delta(tsr, t, a.T, nrm), plan.density_act, beta);
            const float w = al * trans;
            trans *= (1.f - al) + 1e-10f;
            float fr = rgb_raw[0 * ROWS + row], fg = rgb_raw[1 * ROWS + row], fb = rgb_raw[2 * ROWS + row];
            if (plan.refl_kind != NF_REFL_POSLINVIEW) nf_feat_act3(fr, fg, fb, plan.feat_act);      // PosLinearView's colours are final
            cr += w * fr; cg += w * fg; cb += w * fb;
            if (t < a.T - 1) wsum += w;
            if (a.alpha_out) a.alpha_out[ray * a.T + t] = al;
            if (a.weights_out) a.weights_out[ray * a.T + t] = w;
          }
          if (sub == map.tpr - 1) {
            float skyv = plan.bg == NF_BG_WHITE ? 1.f - wsum : 0.f;
            if (plan.bg == NF_BG_RANDOM && a.bg_rand) skyv = __ldg(a.bg_rand + ray) * (1.f - wsum);      // random_color (nerf.py:100-103)
            a.rgb_out[ray * out_ch + 0] = cr + skyv; a.rgb_out[ray * out_ch + 1] = cg + skyv; a.rgb_out[ray * out_ch + 2] = cb + skyv;
          } else { s.carry[0] = trans; s.carry[1] = cr; s.carry[2] = cg; s.carry[3] = cb; s.carry[4] = wsum; }
        }
      }
      __syncthreads();
    }
  }
}

// Stand-alone SkipConnMLP on assembled inputs x0[N,in] -> out[N,out].
__global__ void __launch_bounds__(THREADS, 1)
k_mlp_fp32(const __grid_constant__ NfPlan plan, int which, const uint8_t* __restrict__ packed,
           const float* __restrict__ x0, long long n, float* __restrict__ out, const long long* __restrict__ n_dev) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Fp32Smem& s = *reinterpret_cast<Fp32Smem*>(smem_raw);
  const NfMlpPlan& M = plan.mlp[which];
  if (n_dev) n = *n_dev;                                 // row count from device memory (stream-ordered loops, e.g. the sphere march)
  const long long tiles = (n + ROWS - 1) / ROWS;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    for (int i = threadIdx.x; i < M.in_dims * ROWS; i += THREADS) {
      const int row = i / M.in_dims, k = i - row * M.in_dims;
      const long long g = tile * ROWS + row;
      s.X0[k * ROWS + row] = g < n ? __ldg(x0 + g * M.in_dims + k) : 0.f;
    }
    __syncthreads();
    const int ob = mlp_fp32(M, packed, s);
    for (int i = threadIdx.x; i < M.out_dims * ROWS; i += THREADS) {
      const int row = i / M.out_dims, c = i - row * M.out_dims;
      const long long g = tile * ROWS + row;
      if (g < n) out[g * M.out_dims + c] = s.H[ob][c * ROWS + row];
    }
    __syncthreads();
  }
}

// SDFModel.normals (reference src/sdf.py:43-49): d (sum of ALL outputs of the SDF network) / d p -- utils.autograd (utils.py:266-277)
// back-propagates ones over every output channel, the latent included -- by forward-mode differentiation: 16 points per tile, four
// rows each (value + three tangents).  bound_rad > 0: UnitSphere (sdf.py:66-83), output 0 = max(inner, |p| - rad).
__global__ void __launch_bounds__(THREADS, 1)
k_sdf_normals_fp32(const __grid_constant__ NfPlan plan, const uint8_t* __restrict__ packed, const float* __restrict__ pts, long long n, float bound_rad,
                   float* __restrict__ normals, float* __restrict__ values) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Fp32Smem& s = *reinterpret_cast<Fp32Smem*>(smem_raw);
  const NfMlpPlan& M = plan.mlp[0];
  constexpr int PPT = ROWS / 4;                          // points per tile
  const long long tiles = (n + PPT - 1) / PPT;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    // x0 rows: [p] (SIREN) or [p, sin(p B), cos(p B)] (Fourier MLP) and their derivatives with respect to p
    for (int i = threadIdx.x; i < M.in_dims * PPT; i += THREADS) {
      const int pt = i / M.in_dims, k = i - pt * M.in_dims;
      const long long g = tile * PPT + pt;
      float v = 0.f, t3[3] = {0.f, 0.f, 0.f};
      if (g < n) {
        const float px = __ldg(pts + g * 3), py = __ldg(pts + g * 3 + 1), pz = __ldg(pts + g * 3 + 2);
        if (k < 3) { v = k == 0 ? px : k == 1 ? py : pz; t3[k] = 1.f; }
        else {
          const float* B = reinterpret_cast<const float*>(packed + plan.fourier_off);
          const int F = plan.fourier_freqs, f = (k - 3) % F;
          const bool is_cos = k - 3 >= F;
          const float bx = __ldg(B + f), by = __ldg(B + F + f), bz = __ldg(B + 2 * F + f);
          const float m = fmaf(pz, bz, fmaf(py, by, __fmul_rn(px, bx)));
          const float sn = sinf(m), cs = cosf(m);
          v = is_cos ? cs : sn;
          const float d = is_cos ? -sn : cs;
          t3[0] = d * bx; t3[1] = d * by; t3[2] = d * bz;
        }
      }
      float* col = s.X0 + k * ROWS + pt * 4;
      col[0] = v; col[1] = t3[0]; col[2] = t3[1]; col[3] = t3[2];
    }
    __syncthreads();
    const int ob = mlp_fp32<true>(M, packed, s);
    for (int i = threadIdx.x; i < PPT * 3; i += THREADS) {
      const int pt = i / 3, c = i - pt * 3;
      const long long g = tile * PPT + pt;
      if (g >= n) continue;
      float acc = 0.f;
      for (int k = 0; k < M.out_dims; ++k) {
        float t = s.H[ob][k * ROWS + pt * 4 + 1 + c];
        if (k == 0 && bound_rad > 0.f) {
          const float px = __ldg(pts + g * 3), py = __ldg(pts + g * 3 + 1), pz = __ldg(pts + g * 3 + 2);
          const float r = sqrtf(px * px + py * py + pz * pz);
          // torch.maximum's gradient: the larger operand takes it (ties: split evenly)
          const float inner = s.H[ob][pt * 4], sph = r - bound_rad;
          const float ds = (c == 0 ? px : c == 1 ? py : pz) / r;
          t = inner > sph ? t : inner < sph ? ds : 0.5f * (t + ds);
        }
        acc += t;
      }
      normals[g * 3 + c] = acc;
    }
    if (values) {
      for (int i = threadIdx.x; i < PPT * M.out_dims; i += THREADS) {
        const int pt = i / M.out_dims, c = i - pt * M.out_dims;
        const long long g = tile * PPT + pt;
        if (g >= n) continue;
        float v = s.H[ob][c * ROWS + pt * 4];
        if (c == 0 && bound_rad > 0.f) {
          const float px = __ldg(pts + g * 3), py = __ldg(pts + g * 3 + 1), pz = __ldg(pts + g * 3 + 2);
          v = fmaxf(v, sqrtf(px * px + py * py + pz * pz) - bound_rad);
        }
        values[g * M.out_dims + c] = v;
      }
    }
    __syncthreads();
  }
}

// ---- stage kernels ---------------------------------------------------------------------------
__global__ void k_sample_points(const float* __restrict__ rays, long long n_rays, const float* __restrict__ ts,
                                int T, long long ts_stride, float* __restrict__ pts) {
  const long long total = n_rays * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long ray = i / T; const int t = (int)(i - ray * T);
    const float* r = rays + ray * 6;
    const float tt = __ldg(ts + ray * ts_stride + t);
    pts[i * 3 + 0] = nf_pt(__ldg(r + 0), tt, __ldg(r + 3));
    pts[i * 3 + 1] = nf_pt(__ldg(r + 1), tt, __ldg(r + 4));
    pts[i * 3 + 2] = nf_pt(__ldg(r + 2), tt, __ldg(r + 5));
  }
}

// T % 4 == 0: one thread per 4 consecutive samples of a ray = 48 contiguous, 16-byte aligned output bytes (three float4 stores)
__global__ void k_sample_points4(const float* __restrict__ rays, long long n_rays, const float* __restrict__ ts,
                                 int T, long long ts_stride, float* __restrict__ pts) {
  const int q = T >> 2;
  const long long total = n_rays * q;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long ray = i / q; const int t0 = (int)(i - ray * q) * 4;
    const float* r = rays + ray * 6;
    const float ox = __ldg(r), oy = __ldg(r + 1), oz = __ldg(r + 2), dx = __ldg(r + 3), dy = __ldg(r + 4), dz = __ldg(r + 5);
    const float* tp = ts + ray * ts_stride + t0;
    float o[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float tt = __ldg(tp + k);
      o[3 * k] = nf_pt(ox, tt, dx); o[3 * k + 1] = nf_pt(oy, tt, dy); o[3 * k + 2] = nf_pt(oz, tt, dz);
    }
    float4* dst = reinterpret_cast<float4*>(pts + (ray * T + t0) * 3);
    dst[0] = make_float4(o[0], o[1], o[2], o[3]); dst[1] = make_float4(o[4], o[5], o[6], o[7]); dst[2] = make_float4(o[8], o[9], o[10], o[11]);
  }
}

__global__ void k_hash_encode(const __grid_constant__ NfPlan plan, const uint8_t* __restrict__ packed,
                              const float* __restrict__ pts, long long n, float* __restrict__ feats,
                              uint16_t* __restrict__ idx_out) {
  const float4* tables = reinterpret_cast<const float4*>(packed + plan.hash_off);
  const int L = plan.hash_levels;
  const long long total = n * L;  // one thread per (point, level); level fastest -> coalesced 16 B stores
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / L; const int lvl = (int)(i - p * L);
    uint32_t id[8];
    const float4 f = nf_hash_level(tables + (size_t)lvl * (plan.hash_mask + 1), __ldg(pts + p * 3), __ldg(pts + p * 3 + 1),
                                   __ldg(pts + p * 3 + 2), plan.hash_res[lvl], plan.hash_primes[0], plan.hash_primes[1],
                                   plan.hash_primes[2], plan.hash_mask, id);
    reinterpret_cast<float4*>(feats)[i] = f;
    if (idx_out) for (int c = 0; c < 8; ++c) idx_out[((size_t)lvl * 8 + c) * n + p] = (uint16_t)id[c];
  }
}

// Warp per ray; lanes own consecutive samples in chunks of 32; running transmittance is a
// warp-shuffle multiplicative scan (the stand-alone form of the fused tail). HBM-bound:
// reads 16 B/sample (sigma + rgb), writes 8 B/sample when alpha/weights are requested.
__global__ void k_composite(int density_act, const float* __restrict__ beta_ptr, int bg, const float* __restrict__ sigma_raw,
                            const float* __restrict__ feats, const float* __restrict__ rays, long long n_rays,
                            const float* __restrict__ ts, int T, long long ts_stride,
                            float* __restrict__ rgb_out, float* __restrict__ alpha_out, float* __restrict__ weights_out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float beta = beta_ptr ? __ldg(beta_ptr) : 1.f;
  for (long long ray = warp; ray < n_rays; ray += nwarps) {
    const float* r = rays + ray * 6;
    const float dx = __ldg(r + 3), dy = __ldg(r + 4), dz = __ldg(r + 5);
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float* tsr = ts + ray * ts_stride;
    float carry = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, wsum = 0.f;
    for (int base0 = 0; base0 < T; base0 += 128) {
      // all loads of up to four 32-sample chunks are issued before the first (dependent) scan: the kernel is latency-bound otherwise
      float sg[4], fr[4], fg[4], fb[4], d0[4], d1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t = base0 + q * 32 + lane;
        sg[q] = fr[q] = fg[q] = fb[q] = d0[q] = d1[q] = 0.f;
        if (t < T) {
          sg[q] = __ldg(sigma_raw + ray * T + t);
          const float* f = feats + (ray * T + t) * 3;
          fr[q] = __ldg(f); fg[q] = __ldg(f + 1); fb[q] = __ldg(f + 2);
          d0[q] = __ldg(tsr + t); d1[q] = t + 1 < T ? __ldg(tsr + t + 1) : 0.f;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t = base0 + q * 32 + lane;
        if (base0 + q * 32 >= T) break;
        float al = 0.f;
        if (t < T) {
          const float dl = (t == T - 1) ? 1e10f : fmaxf(__fsub_rn(d1[q], d0[q]), 1e-5f);
          al = nf_alpha(sg[q], __fmul_rn(dl, nrm), density_act, beta);
        }
        float p = t < T ? (1.f - al) + 1e-10f : 1.f;   // inclusive product scan of (1 - alpha + 1e-10)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const float qq = __shfl_up_sync(0xffffffffu, p, d); if (lane >= d) p *= qq; }
        float excl = __shfl_up_sync(0xffffffffu, p, 1); if (lane == 0) excl = 1.f;
        const float w = al * (carry * excl);
        carry *= __shfl_sync(0xffffffffu, p, 31);
        if (t < T) {
          if (alpha_out) alpha_out[ray * T + t] = al;
          if (weights_out) weights_out[ray * T + t] = w;
          cr += w * fr[q]; cg += w * fg[q]; cb += w * fb[q];
          if (t < T - 1) wsum += w;
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      cr += __shfl_xor_sync(0xffffffffu, cr, d); cg += __shfl_xor_sync(0xffffffffu, cg, d);
      cb += __shfl_xor_sync(0xffffffffu, cb, d); wsum += __shfl_xor_sync(0xffffffffu, wsum, d);
    }
    if (lane == 0) {
      const float skyv = bg == NF_BG_WHITE ? 1.f - wsum : 0.f;
      rgb_out[ray * 3 + 0] = cr + skyv; rgb_out[ray * 3 + 1] = cg + skyv; rgb_out[ray * 3 + 2] = cb + skyv;
    }
  }
}

// Inverse-CDF resampling + sorted merge, one warp per ray (T, Nf <= 256).  reference src/nerf.py:1745-1779 (restated).
constexpr int PDF_MAX = 256;
__global__ void k_sample_pdf(const float* __restrict__ ts, int T, const float* __restrict__ weights, long long n_rays,
                             const float* __restrict__ u, int nf, float* __restrict__ out) {
  __shared__ float s_cdf[8][PDF_MAX];
  __shared__ float s_new[8][PDF_MAX];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float* cdf = s_cdf[wib]; float* nw = s_new[wib];
  const int nb = T - 1;                                   // bins (mid-points) == cdf entries
  for (long long ray = warp; ray < n_rays; ray += nwarps) {
    const float* w = weights + ray * T + 1;               // weights[1:-1]: T-2 values
    // sum of (w + 1e-5), then pdf and its running sum; sequential adds like torch.cumsum
    float part = 0.f;
    for (int i = lane; i < T - 2; i += 32) part += __ldg(w + i) + 1e-5f;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) {
      float c = 0.f; cdf[0] = 0.f;
      for (int i = 0; i < T - 2; ++i) { c += (__ldg(w + i) + 1e-5f) / part; cdf[i + 1] = c; }
    }
    __syncwarp();
    for (int j = lane; j < nf; j += 32) {
      const float uu = __ldg(u + ray * nf + j);
      int lo = 0, hi = nb;                                // searchsorted(cdf, u, right=True): first index with cdf > u
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] > uu) hi = mid; else lo = mid + 1; }
      const int below = max(lo - 1, 0), above = min(lo, nb - 1);
      const float cb = cdf[below], ca = cdf[above];
      float denom = ca - cb; if (denom < 1e-5f) denom = 1.f;
      const float t = (uu - cb) / denom;
      const float bb = 0.5f * (__ldg(ts + below) + __ldg(ts + below + 1)), ba = 0.5f * (__ldg(ts + above) + __ldg(ts + above + 1));
      nw[j] = __fadd_rn(bb, __fmul_rn(t, ba - bb));
    }
    __syncwarp();
    // rank-based merge of {coarse ts (sorted), new samples (unsorted)} into the sorted output; ties: coarse first,
    // then new samples by index
    float* o = out + ray * (T + nf);
    for (int j = lane; j < nf; j += 32) {
      const float v = nw[j];
      int lo = 0, hi = T;                                 // coarse ts <= v
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(ts + mid) <= v) lo = mid + 1; else hi = mid; }
      int r = lo;
      for (int k = 0; k < nf; ++k) { const float x = nw[k]; r += (x < v) || (x == v && k < j); }
      o[r] = v;
    }
    for (int i = lane; i < T; i += 32) {
      const float v = __ldg(ts + i);
      int r = i;
      for (int k = 0; k < nf; ++k) r += nw[k] < v;
      o[r] = v;
    }
    __syncwarp();
  }
}

// runner.render's pixel grid + NeRFCamera.sample_positions (reference runner.py:490-505, src/cameras.py:45-66): one thread per ray
__global__ void k_generate_rays(const float* __restrict__ c2w, long long B, float focal, float half, int top, int left, int H, int W,
                                int recip, float* __restrict__ out) {
  const long long total = B * H * W;
  const float inv = __fdiv_rn(1.f, focal);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W); const int h = (int)((i / W) % H); const long long b = i / ((long long)W * H);
    const float u = (float)(left + w), v = (float)(top + h);
    const float a0 = __fsub_rn(u, half), a1 = __fsub_rn(v, half);
    const float d0 = recip ? __fmul_rn(a0, inv) : __fdiv_rn(a0, focal);
    const float d1 = -(recip ? __fmul_rn(a1, inv) : __fdiv_rn(a1, focal));
    const float* M = c2w + b * 12;
    float* o = out + i * 6;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = __ldg(M + c * 4 + 3);
      o[3 + c] = __fadd_rn(__fadd_rn(__fmul_rn(d0, __ldg(M + c * 4)), __fmul_rn(d1, __ldg(M + c * 4 + 1))), __fmul_rn(-1.f, __ldg(M + c * 4 + 2)));
    }
  }
}

// DTUCamera.sample_positions (reference src/cameras.py:159-174 lift, 189-223): pixel (u, v) scaled to the 1600 x 1200 original,
// lifted through the intrinsics, moved to world space by the pose (the reference's bmm, K = 4), direction = normalised difference
// to the camera centre -- UNIT-norm r_d, unlike NeRFCamera.  pose [B,4,4] (rows 0..2 used), intrinsic [B,ir,ic] (ir, ic >= 3).
__global__ void k_generate_rays_dtu(const float* __restrict__ pose, const float* __restrict__ intr, int ir, int ic, long long B, float sx, float sy,
                                    int top, int left, int H, int W, float* __restrict__ out) {
  const long long total = B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W); const int h = (int)((i / W) % H); const long long b = i / ((long long)W * H);
    const float u = __fmul_rn((float)(left + w), sx), v = __fmul_rn((float)(top + h), sy);
    const float* K = intr + b * ir * ic;
    const float fx = __ldg(K), sk = __ldg(K + 1), cx = __ldg(K + 2), fy = __ldg(K + ic + 1), cy = __ldg(K + ic + 2);
    // x_lift = (x - cx + cy*sk/fy - sk*y/fy) / fx * z,  y_lift = (y - cy) / fy * z,  z = 1  (operation order as written in lift())
    const float xl = __fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(u, cx), __fdiv_rn(__fmul_rn(cy, sk), fy)), __fdiv_rn(__fmul_rn(sk, v), fy)), fx);
    const float yl = __fdiv_rn(__fsub_rn(v, cy), fy);
    const float* P = pose + b * 16;
    float d[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = __ldg(P + c * 4 + 3);
      float acc = __fmul_rn(__ldg(P + c * 4), xl);
      acc = fmaf(__ldg(P + c * 4 + 1), yl, acc); acc = fmaf(__ldg(P + c * 4 + 2), 1.f, acc); acc = fmaf(t, 1.f, acc);
      d[c] = __fsub_rn(acc, t);
    }
    const float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]))), 1e-12f);
    float* o = out + i * 6;
#pragma unroll
    for (int c = 0; c < 3; ++c) { o[c] = __ldg(P + c * 4 + 3); o[3 + c] = __fdiv_rn(d[c], nrm); }
  }
}

// radii_x, reference src/utils.py:77-81: one thread per ray of the [B,H,W] crop
__global__ void k_ray_radii(const float* __restrict__ rays, long long B, int H, int W, float* __restrict__ out) {
  const long long total = B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W); const int h = (int)((i / W) % H); const long long b = i / ((long long)W * H);
    const int hd = h < H - 1 ? h : H - 3;              // cat([dx, dx[:, -2:-1, :]]): the last row repeats difference H-3
    const float* r0 = rays + ((b * H + hd) * W + w) * 6 + 3;
    const float* r1 = r0 + (long long)W * 6;
    const float dx = __fsub_rn(__ldg(r0), __ldg(r1)), dy = __fsub_rn(__ldg(r0 + 1), __ldg(r1 + 1)), dz = __fsub_rn(__ldg(r0 + 2), __ldg(r1 + 2));
    const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    out[i] = __fdiv_rn(__fmul_rn(d, 2.f), 3.4641016151377544f);
  }
}

int num_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

}  // namespace

// ---- launchers (called from nf_api.cu) -------------------------------------------------------
cudaError_t nf_launch_generate_rays(const float* c2w, int64_t B, float focal, int size, int top, int left, int H, int W, int recip,
                                    float* out, cudaStream_t st) {
  const long long total = B * H * W;
  if (total == 0) return cudaSuccess;
  const long long want = (total + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  k_generate_rays<<<grid, 256, 0, st>>>(c2w, B, focal, (float)size * 0.5f, top, left, H, W, recip, out);
  return cudaGetLastError();
}

cudaError_t nf_launch_generate_rays_dtu(const float* pose, const float* intr, int ir, int ic, int64_t B, int size, int top, int left, int H, int W,
                                        float* out, cudaStream_t st) {
  const long long total = B * H * W;
  if (total == 0) return cudaSuccess;
  const long long want = (total + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  k_generate_rays_dtu<<<grid, 256, 0, st>>>(pose, intr, ir, ic, B, 1600.f / (float)size, 1200.f / (float)size, top, left, H, W, out);
  return cudaGetLastError();
}

cudaError_t nf_launch_ray_radii(const float* rays, int64_t B, int H, int W, float* out, cudaStream_t st) {
  const long long total = B * H * W;
  if (total == 0) return cudaSuccess;
  const long long want = (total + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  k_ray_radii<<<grid, 256, 0, st>>>(rays, B, H, W, out);
  return cudaGetLastError();
}

cudaError_t nf_launch_render_fp32(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, const float* ts,
                                  int T, int64_t ts_stride, const float* noise, const float* ray_time, const nf_mip_args* mip,
                                  float* rgb, float* alpha, float* weights, cudaStream_t st, const nf_render_aux* aux) {
  static_assert(sizeof(Fp32Smem) <= 227 * 1024, "fp32 pipeline smem");
  cudaError_t e = cudaFuncSetAttribute(k_render_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Fp32Smem));
  if (e != cudaSuccess) return e;
  const NfTileMap map(T, ROWS);
  const long long units = map.units(n_rays);
  if (units == 0) return cudaSuccess;
  const int grid = (int)(units < num_sms() ? units : num_sms());
  RenderArgs a{(const uint8_t*)packed, rays, n_rays, ts, T, ts_stride, noise, ray_time, rgb, alpha, weights, NfMipIn{}};
  if (plan.mip != NF_MIP_NONE) {
    if (!mip || !mip->radius || ts_stride != 0) return cudaErrorInvalidValue;
    a.mip = NfMipIn{plan.mip, ts, T, rays, mip->radius, mip->rays_all, mip->radius_all, (long long)mip->n_rays_all, (long long)mip->ray_base};
    if (plan.mip == NF_MIP_CYLINDER_REF && (!mip->rays_all || !mip->radius_all || mip->ray_base < 0 || mip->ray_base + n_rays > mip->n_rays_all))
      return cudaErrorInvalidValue;
  }
  if (aux) {
    a.pts = aux->pts; a.bg_rand = aux->bg_rand;
    a.pts_out = aux->pts_out; a.dp_out = aux->dp_out; a.rigid_dp_out = aux->rigid_dp_out; a.rigidity_out = aux->rigidity_out;
  }
  if (plan.bg == NF_BG_RANDOM && !a.bg_rand) return cudaErrorInvalidValue;
  k_render_fp32<<<grid, THREADS, sizeof(Fp32Smem), st>>>(plan, a);
  return cudaGetLastError();
}

cudaError_t nf_launch_mlp_fp32(const NfPlan& plan, int which, const void* packed, const float* x0, int64_t n, float* out, cudaStream_t st,
                               const long long* n_dev) {
  cudaError_t e = cudaFuncSetAttribute(k_mlp_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Fp32Smem));
  if (e != cudaSuccess) return e;
  const long long tiles = (n + ROWS - 1) / ROWS;
  if (tiles == 0) return cudaSuccess;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  k_mlp_fp32<<<grid, THREADS, sizeof(Fp32Smem), st>>>(plan, which, (const uint8_t*)packed, x0, n, out, n_dev);
  return cudaGetLastError();
}

cudaError_t nf_launch_sdf_normals(const NfPlan& plan, const void* packed, const float* pts, int64_t n, float bound_rad, float* normals, float* values,
                                  cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_sdf_normals_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Fp32Smem));
  if (e != cudaSuccess) return e;
  const long long tiles = (n + ROWS / 4 - 1) / (ROWS / 4);
  if (tiles == 0) return cudaSuccess;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  k_sdf_normals_fp32<<<grid, THREADS, sizeof(Fp32Smem), st>>>(plan, (const uint8_t*)packed, pts, n, bound_rad, normals, values);
  return cudaGetLastError();
}

cudaError_t nf_launch_sample_points(const float* rays, int64_t n_rays, const float* ts, int T, int64_t ts_stride, float* pts, cudaStream_t st) {
  const long long total = n_rays * T;
  if (total == 0) return cudaSuccess;
  const long long want = (total + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  if ((T & 3) == 0 && ((uintptr_t)pts & 15) == 0) {
    const long long want4 = (total / 4 + 255) / 256;
    const int grid4 = (int)(want4 < (long long)num_sms() * 16 ? want4 : (long long)num_sms() * 16);
    k_sample_points4<<<grid4, 256, 0, st>>>(rays, n_rays, ts, T, ts_stride, pts);
  } else {
    k_sample_points<<<grid, 256, 0, st>>>(rays, n_rays, ts, T, ts_stride, pts);
  }
  return cudaGetLastError();
}

cudaError_t nf_launch_hash_encode(const NfPlan& plan, const void* packed, const float* pts, int64_t n, float* feats, uint16_t* idx, cudaStream_t st) {
  const long long total = n * plan.hash_levels;
  if (total == 0) return cudaSuccess;
  const long long want = (total + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  k_hash_encode<<<grid, 256, 0, st>>>(plan, (const uint8_t*)packed, pts, n, feats, idx);
  return cudaGetLastError();
}

cudaError_t nf_launch_sample_pdf(const float* ts, int T, const float* weights, int64_t n_rays, const float* u, int nf, float* out, cudaStream_t st) {
  if (n_rays == 0) return cudaSuccess;
  if (T < 3 || T > PDF_MAX || nf < 1 || nf > PDF_MAX) return cudaErrorInvalidValue;
  const long long want = (n_rays + 7) / 8;
  const int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  k_sample_pdf<<<grid, 256, 0, st>>>(ts, T, weights, n_rays, u, nf, out);
  return cudaGetLastError();
}

// volumetric_integrate (reference src/nerf.py:79-80) of per-sample values other than the colours: out[r, c] = sum_t w[r, t] v[r, t, c]
// (depth: v = ts; flow / rigidity maps: v = rigid_dp / rigidity, runner.py:511-531,894-916).  A warp per ray, lanes over (t, c) pairs
// in memory order (coalesced), one shuffle reduction per channel.  HBM-bound: 4 T (1 + C) bytes read per ray.
template <int C>
__global__ void k_integrate(const float* __restrict__ w, const float* __restrict__ v, long long n_rays, int T, long long v_stride,
                            float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long ray = warp; ray < n_rays; ray += nwarps) {
    const float* wr = w + ray * T; const float* vr = v + ray * v_stride;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float wt = __ldg(wr + t);
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fmaf(wt, __ldg(vr + (long long)t * C + c), acc[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], d);
      if (lane == 0) out[ray * C + c] = acc[c];
    }
  }
}
cudaError_t nf_launch_integrate(const float* weights, const float* vals, int64_t n_rays, int T, int C, int64_t vals_ray_stride, float* out, cudaStream_t st) {
  if (n_rays == 0) return cudaSuccess;
  const long long want = (n_rays * 32 + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  if (C == 1) k_integrate<1><<<grid, 256, 0, st>>>(weights, vals, n_rays, T, vals_ray_stride, out);
  else if (C == 3) k_integrate<3><<<grid, 256, 0, st>>>(weights, vals, n_rays, T, vals_ray_stride, out);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t nf_launch_composite(const NfPlan& plan, const void* packed, const float* sigma_raw, const float* feats, const float* rays, int64_t n_rays,
                                const float* ts, int T, int64_t ts_stride, float* rgb, float* alpha, float* weights, cudaStream_t st) {
  if (n_rays == 0) return cudaSuccess;
  const long long want = (n_rays * 32 + 255) / 256;
  const int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  const float* beta = (plan.density_act == NF_DENS_LAPLACE && packed) ? reinterpret_cast<const float*>((const uint8_t*)packed + plan.scale_off) : nullptr;
  k_composite<<<grid, 256, 0, st>>>(plan.density_act, beta, plan.bg, sigma_raw, feats, rays, n_rays, ts, T, ts_stride, rgb, alpha, weights);
  return cudaGetLastError();
}
