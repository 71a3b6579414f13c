// nf_train.cu -- backward of the fused render path: what the reference gets from loss.backward()
// (runner.py:820) through PyTorch autograd, as hand-written sm_100a kernels over the activation stash that the training
// forward (k_render_tc3<..., TRAIN>) leaves in the workspace.
//
//   d rgb[R,3] --k_composite_bwd--> d sigma_raw[R,T], d rgb_raw[R,T,3]      (nf_bwd.cu; the feature activation's derivative folded in)
//              --k_grad_scale-----> loss scale S (power of two): the per-sample gradients are tiny (~1e-7), far below fp16's range
//              --k_bwd_chain------> per 128-sample tile, the Linears walked in REVERSE: G_{l-1} = (G_l W_l) * act'(z_{l-1})
//                                   tcgen05 (cta_group::1, M = 128): A = G_l (fp16, K-major canonical image in smem, written by
//                                   the epilogue warps), B = the TRANSPOSED weight image (w16t), D in TMEM; act' from the stash
//                                   (LeakyReLU: sign of the stashed activation; sin: the stashed cosine).  Every G_l goes to the
//                                   workspace (fp16); the gradient of the hash features leaves as fp32.
//              --k_bwd_dw---------> dW_l = sum over tiles G_l^T A_l: tcgen05 with BOTH operands MN-major straight from the two
//                                   stashes (the K-major [K/8][128][8] image of a tile IS the MN-major canonical layout of its
//                                   transpose: LBO = 128 B along the samples, SBO = 2048 B along the features), split-K over
//                                   tiles across the CTAs, fp32 atomics into dWt; db_l = G_l^T 1 rides along as a 16-column MMA
//                                   against a block of ones.
//              --k_unpack_grads---> gradients in the reference's parameter layout ([out,in] row-major, reference column orders)
//              --k_hash_bwd_tiles-> scatter-add of d feats into the embedding tables (float4 atomics)
#include <cstdio>
#include <cstddef>
#include <cmath>
#include "nf_common.cuh"
#include "nf_kernels.h"
#include "nf_tc_ptx.cuh"

namespace {
using namespace nf_ptx;

int tr_num_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// =====================================================================================================================
// transposed weight images
// =====================================================================================================================
// (layout: x0 part [n_pad/8][k0_pad][8] then hidden part [n_pad/8][256][8], fp16; reduction dimension = the Linear's OUTPUT feature in
// tensor order, rows = its INPUT feature in tensor order -- written by k_pack_all)
// Everything nf_pack_weights writes for the MLPs, ONE launch (blockIdx.y = Linear): the fp32 Wt[k][n_pad] + bias of the CUDA-core
// pipeline, the fp16 UMMA-canonical images (whole and CTA-pair split, the latter with the bias K-step), the bias in tensor order
// and the transposed images of the backward.  The caller has zeroed the region (one memset): padding rows / columns stay zero.
// (Round 1-2 packed with 3 launches + 4 memsets per Linear: ~90 stream operations per optimiser step.)
constexpr int PACK_MAX = 3 * NF_MAX_LIN;
struct PackArgs { const float* W[PACK_MAX]; const float* b[PACK_MAX]; int32_t mj[PACK_MAX]; };
__global__ void k_pack_all(const __grid_constant__ NfPlan plan, const __grid_constant__ PackArgs a, uint8_t* __restrict__ packed) {
  const int li = blockIdx.y, m = a.mj[li] >> 4, j = a.mj[li] & 15;
  const NfLinPlan& L = plan.mlp[m].lin[j];
  const float* __restrict__ W = a.W[li]; const float* __restrict__ b = a.b[li];
  // a narrower reference hidden size (PosLinearView.view: 128) is zero-padded to 256: exact, the extra units stay act(0) = 0
  const int href = plan.mlp[m].hidden_ref, kh_ref = L.k_hidden ? href : 0, n_src = L.is_out ? L.n : href;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  {
    float* __restrict__ Wt = reinterpret_cast<float*>(packed + L.wt_off); float* __restrict__ bp = reinterpret_cast<float*>(packed + L.b_off);
    const int kh_pad = L.k_hidden, kx = L.k_x0, n_pad = L.n_pad, total = (kh_pad + kx) * n_pad, k_src = kh_ref + kx;
    for (int i = tid; i < total; i += nth) {
      const int kk = i / n_pad, nn = i - kk * n_pad;
      const int kr = kk < kh_pad ? (kk < kh_ref ? kk : -1) : kh_ref + (kk - kh_pad);     // padded hidden rows have no source column
      Wt[i] = (nn < n_src && kr >= 0) ? W[(size_t)nn * k_src + kr] : 0.f;
    }
    for (int i = tid; i < n_pad; i += nth) bp[i] = i < n_src ? b[i] : 0.f;
  }
  if (href != NF_HIDDEN) return;                     // the tensor pipeline does not take such a model (nf_tensor_pipeline_support)
  const int k_ref_total = L.k_hidden + L.k_x0, k_total = L.k0_pad + L.k_hidden, nh = L.n_pad >> 1;
  __half* __restrict__ img = reinterpret_cast<__half*>(packed + L.w16_off);
  __half* __restrict__ imgh = reinterpret_cast<__half*>(packed + L.w16h_off);
  __half* __restrict__ img_x = reinterpret_cast<__half*>(packed + L.w16t_off);
  __half* __restrict__ img_h = img_x + (size_t)L.n_pad * L.k0_pad;
  float* __restrict__ b16 = reinterpret_cast<float*>(packed + L.b16_off);
  const int total = L.n * k_ref_total;
  for (int i = tid; i < total; i += nth) {
    const int n_ref = i / k_ref_total, k_ref = i - n_ref * k_ref_total;
    const int kx_tc = k_ref < L.k_hidden ? 0 : nf_x0_perm(plan, m, k_ref - L.k_hidden);
    const int k_tc = k_ref < L.k_hidden ? L.k0_pad + k_ref : kx_tc;
    const int n_tc = L.is_out ? nf_out_perm(plan, m, n_ref) : n_ref;
    const __half h = __float2half_rn(W[i]);
    img[(size_t)(k_tc >> 3) * (L.n_pad * 8) + n_tc * 8 + (k_tc & 7)] = h;
    const int rank = n_tc / nh, nl = n_tc - rank * nh;                               // CTA-pair split along N
    imgh[(size_t)rank * ((k_total + 16) >> 3) * (nh * 8) + (size_t)(k_tc >> 3) * (nh * 8) + nl * 8 + (k_tc & 7)] = h;
    if (k_ref < L.k_hidden) img_h[(size_t)(n_tc >> 3) * (NF_HIDDEN * 8) + k_ref * 8 + (n_tc & 7)] = h;
    else img_x[(size_t)(n_tc >> 3) * (L.k0_pad * 8) + kx_tc * 8 + (n_tc & 7)] = h;
  }
  for (int n_ref = tid; n_ref < L.n; n_ref += nth) {
    const int n_tc = L.is_out ? nf_out_perm(plan, m, n_ref) : n_ref;
    b16[n_tc] = b[n_ref];
    // the bias K-step of the pair images: rows k_total (fp16 hi) and k_total + 1 (fp16 lo) against a [1, 1, 0, ...] operand
    const int rank = n_tc / nh, nl = n_tc - rank * nh;
    const __half hi = __float2half_rn(b[n_ref]);
    const __half lo = __float2half_rn(b[n_ref] - __half2float(hi));
    __half* row = imgh + (size_t)rank * ((k_total + 16) >> 3) * (nh * 8) + (size_t)(k_total >> 3) * (nh * 8) + nl * 8;
    row[0] = hi; row[1] = lo;
  }
}
// the embedding tables (one nn.Parameter per level) -> the blob, one launch (blockIdx.y = table)
struct CopyArgs { const float4* src[48]; float4* dst[48]; };
__global__ void k_copy_tables(const __grid_constant__ CopyArgs a, long long n4) {
  const float4* __restrict__ s = a.src[blockIdx.y]; float4* __restrict__ d = a.dst[blockIdx.y];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) d[i] = __ldg(s + i);
}

#ifndef NF_BWD_SIDE_STREAM
#define NF_BWD_SIDE_STREAM 1
#endif
// =====================================================================================================================
// loss scale
// =====================================================================================================================
__global__ void k_grad_absmax(const float* __restrict__ a, long long na, const float* __restrict__ b, long long nb, unsigned* __restrict__ out) {
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < na + nb; i += (long long)gridDim.x * blockDim.x) {
    const float v = fabsf(i < na ? a[i] : b[i - na]);
    if (v == v && v < 3e38f) m = fmaxf(m, v);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
// scale[0] = S = 2^k with max|g| * S in [64, 128), scale[1] = 1/S, scale[2] = max|g|
__global__ void k_grad_scale(float* scale) {
  const float m = __uint_as_float(reinterpret_cast<unsigned*>(scale)[2]);
  int e = 0;
  float S = 1.f;
  if (m > 0.f) { frexpf(m, &e); S = ldexpf(1.f, 7 - e); }         // m = f * 2^e, f in [0.5, 1)
  scale[0] = S; scale[1] = 1.f / S;
}

// =====================================================================================================================
// k_bwd_chain: dX through the Linears in reverse
// =====================================================================================================================
constexpr int BW_NST = 3;
constexpr int BW_STAGE = 32 * 1024;
constexpr int BW_EPIW = 16;
constexpr int BW_THREADS = 32 * (BW_EPIW + 3);           // + weight producer, stash producer, MMA issuer
constexpr int BW_MAX_JOBS = 2 * NF_TRAIN_MAX_LIN;
constexpr uint32_t COL_MAIN = 0, COL_XA = 256, COL_XR = 384;   // TMEM: dA (256), d act(x0) from the skip Linears, d x0 from `init`

struct BwSmem {
  uint8_t H[ROWS * 256 * 2];       // G_l: the A operand (K-major canonical image)
  uint8_t S[ROWS * 256 * 2];       // the stash tile act' needs
  uint8_t W[BW_NST * BW_STAGE];
  unsigned long long w_full[BW_NST], w_empty[BW_NST], s_full, s_empty, acc_full, a_ready;
  uint32_t tmem_base; uint32_t pad_;
};
static_assert(sizeof(BwSmem) <= 227 * 1024, "backward chain smem");

struct __align__(16) BwJob {          // D[128 x n] (+)= G[128 x 16 k_steps] * Wt
  uint32_t k_steps, spc, step_bytes, w_off;
  uint32_t idesc, bhi, bstep4, d_col;
  uint32_t acc0, pad0_, pad1_, pad2_;
};
enum { PH_DY = 0, PH_HIDDEN = 1, PH_MID = 2, PH_END = 3 };
struct __align__(16) BwPhase {
  int32_t kind, act, job0, n_jobs;                 // the jobs run AFTER this phase's epilogue
  uint32_t s_off256, s_tile256, s_bytes, has_xa;   // stash tile this phase's act' needs
  uint32_t g_off256, g_tile256, g_cols, k0_pad;    // where the G this phase produces goes (g_cols = its width)
  int32_t inter, hcols2, tiny, pad2_;              // hcols2 (PH_MID): x0 columns [inter, inter + hcols2) are the head's own hash features; tiny (PH_DY): one MLP, outputs [sigma, rgb]
};
struct __align__(16) BwProg { int32_t n_phases, n_jobs, pad0_, pad1_; BwPhase ph[NF_TRAIN_MAX_LIN + 1]; BwJob job[BW_MAX_JOBS]; };

struct BwArgs {
  const uint8_t* packed; uint8_t* ws;
  long long n_rays, n_tiles; int T, rpu, tpr;
  const float* d_sigma; const float* d_rgbraw; float* dx0_out; const float* scale;
  float* dx0b_out;      // Positional head: gradient of its own encoder's features [n_tiles * 128][32]
};

__device__ __forceinline__ void stg_v4(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float h_lo(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }
__device__ __forceinline__ float h_hi(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
// d act / d z from what the forward stashed: LeakyReLU -> the activated value's sign; sin -> the stashed cosine
__device__ __forceinline__ float act_grad(float d, float s, int act) { return act == NF_ACT_SIN ? d * s : (s > 0.f ? d : 0.01f * d); }

__global__ void __launch_bounds__(BW_THREADS, 1)
k_bwd_chain(const __grid_constant__ BwProg prog, const BwArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  BwSmem& s = *reinterpret_cast<BwSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long my_tiles = a.n_tiles > blockIdx.x ? (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int np = prog.n_phases;                      // n_lin + 1

  if (threadIdx.x == 0) {
    for (int i = 0; i < BW_NST; ++i) { mbar_init(smem_u32(&s.w_full[i]), 1); mbar_init(smem_u32(&s.w_empty[i]), 1); }
    mbar_init(smem_u32(&s.s_full), 1); mbar_init(smem_u32(&s.s_empty), BW_EPIW);
    mbar_init(smem_u32(&s.acc_full), 1); mbar_init(smem_u32(&s.a_ready), BW_EPIW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == BW_EPIW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;

  if (warp == BW_EPIW) {
    // ================= weight producer: every job's transposed image, chunk by chunk, through the ring =================
    if (elect_one()) {
      uint32_t stage = 0, use = 0;
      for (long long it = 0; it < my_tiles; ++it)
        for (int p = 0; p + 1 < np; ++p)
          for (int jb = prog.ph[p].job0; jb < prog.ph[p].job0 + prog.ph[p].n_jobs; ++jb) {
            const BwJob& J = prog.job[jb];
            const uint8_t* src = a.packed + J.w_off;
            for (uint32_t k0 = 0; k0 < J.k_steps; k0 += J.spc) {
              const uint32_t bytes = (J.k_steps - k0 < J.spc ? J.k_steps - k0 : J.spc) * J.step_bytes;
              mbar_wait(smem_u32(&s.w_empty[stage]), ((use / BW_NST) & 1u) ^ 1u);
              mbar_expect_tx(smem_u32(&s.w_full[stage]), bytes);
              bulk_g2s(smem_u32(s.W + stage * BW_STAGE), src + (size_t)k0 * J.step_bytes, bytes, smem_u32(&s.w_full[stage]));
              ++use; if (++stage == BW_NST) stage = 0;
            }
          }
    }
  } else if (warp == BW_EPIW + 1) {
    // ================= stash producer: the tile act' of the next epilogue needs =================
    if (elect_one()) {
      uint32_t use = 0;
      for (long long it = 0; it < my_tiles; ++it) {
        const long long g = blockIdx.x + it * gridDim.x;
        for (int p = 1; p < np; ++p) {
          const BwPhase& P = prog.ph[p];
          mbar_wait(smem_u32(&s.s_empty), (use & 1u) ^ 1u);
          mbar_expect_tx(smem_u32(&s.s_full), P.s_bytes);
          bulk_g2s(smem_u32(s.S), a.ws + ((size_t)P.s_off256 + (size_t)g * P.s_tile256) * 256, P.s_bytes, smem_u32(&s.s_full));
          ++use;
        }
      }
    }
  } else if (warp == BW_EPIW + 2) {
    // ================= MMA issuer =================
    if (elect_one()) {
      uint32_t stage = 0, use = 0, a_par = 0;
      const uint32_t h4 = (smem_u32(s.H) >> 4) | ((uint32_t)(KG_BYTES >> 4) << 16);
      const uint32_t w4 = smem_u32(s.W) >> 4;
      const uint32_t kstep4 = (uint32_t)(2 * KG_BYTES) >> 4;
      for (long long it = 0; it < my_tiles; ++it)
        for (int p = 0; p + 1 < np; ++p) {
          mbar_wait(smem_u32(&s.a_ready), a_par); a_par ^= 1u;
          tc_fence_after();
          for (int jb = prog.ph[p].job0; jb < prog.ph[p].job0 + prog.ph[p].n_jobs; ++jb) {
            const BwJob& J = prog.job[jb];
            for (uint32_t k0 = 0; k0 < J.k_steps; k0 += J.spc) {
              const uint32_t nst = J.k_steps - k0 < J.spc ? J.k_steps - k0 : J.spc;
              mbar_wait(smem_u32(&s.w_full[stage]), (use / BW_NST) & 1u);
              tc_fence_after();
              const uint32_t b4 = (w4 + stage * (uint32_t)(BW_STAGE >> 4)) | J.bhi;
              for (uint32_t i = 0; i < nst; ++i)
                umma_f16(tmem + J.d_col, umma_desc_lo(h4 + (k0 + i) * kstep4), umma_desc_lo(b4 + i * J.bstep4), J.idesc,
                         (k0 + i > 0 || !J.acc0) ? 1u : 0u);
              umma_commit(smem_u32(&s.w_empty[stage]));
              ++use; if (++stage == BW_NST) stage = 0;
            }
          }
          umma_commit(smem_u32(&s.acc_full));
        }
    }
  } else {
    // ================= epilogue warps: lane quarter q = warp % 4, column quarter cq = warp / 4 =================
    const int q = warp & 3, cq = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    const NfStreamMap map(a.T, ROWS);
    const float S = __ldg(a.scale), invS = __ldg(a.scale + 1);
    uint32_t acc_par = 0, s_par = 0;
    for (long long it = 0; it < my_tiles; ++it) {
      const long long g = blockIdx.x + it * gridDim.x;
      const long long u = g / a.tpr; const int sub = (int)(g - u * a.tpr);
      long long ray; int t;
      const bool valid = map.locate(u, sub, row, a.n_rays, ray, t);
      for (int p = 0; p < np; ++p) {
        const BwPhase& P = prog.ph[p];
        if (p > 0) {
          mbar_wait_suspend(smem_u32(&s.acc_full), acc_par); acc_par ^= 1u;
          mbar_wait_suspend(smem_u32(&s.s_full), s_par); s_par ^= 1u;
          tc_fence_after();
        }
        uint8_t* gG = a.ws + ((size_t)P.g_off256 + (size_t)g * P.g_tile256) * 256;     // the G this phase produces (if any)
        if (P.kind == PH_DY) {
          // G of the path's last Linear: [d rgb_raw (3) * S, 0 ...], 16 columns (TinyNeRF's single MLP: [d sigma_raw, d rgb_raw] * S)
          if (cq == 0) {
            float r = 0.f, gg = 0.f, b = 0.f, ds = 0.f;
            if (valid) { const float* d = a.d_rgbraw + (ray * a.T + t) * 3; r = __ldg(d) * S; gg = __ldg(d + 1) * S; b = __ldg(d + 2) * S; }
            if (valid && P.tiny) ds = __ldg(a.d_sigma + ray * a.T + t) * S;
            const uint32_t o0 = P.tiny ? pack_h2(ds, r) : pack_h2(r, gg), o1 = P.tiny ? pack_h2(gg, b) : pack_h2(b, 0.f);
            st_v4(s.H + row * 16, o0, o1, 0, 0); st_v4(s.H + KG_BYTES + row * 16, 0, 0, 0, 0);
            stg_v4(gG + row * 16, o0, o1, 0, 0); stg_v4(gG + KG_BYTES + row * 16, 0, 0, 0, 0);
          }
        } else if (P.kind == PH_HIDDEN) {
          // G_{l-1} = dA * act'(z_{l-1}): 256 columns, 4 units of 16 per warp
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            const int un = cq * 4 + i;
            uint32_t v[16];
            tmem_ld16(t_lane + COL_MAIN + un * 16, v);
            const uint4 s0 = *reinterpret_cast<const uint4*>(s.S + (un * 2) * KG_BYTES + row * 16);
            const uint4 s1 = *reinterpret_cast<const uint4*>(s.S + (un * 2 + 1) * KG_BYTES + row * 16);
            tmem_ld_wait(); reg_fence16(v);
            const uint32_t sw[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            uint32_t o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              o[k] = pack_h2(act_grad(__uint_as_float(v[2 * k]), h_lo(sw[k]), P.act), act_grad(__uint_as_float(v[2 * k + 1]), h_hi(sw[k]), P.act));
            uint8_t* dst = s.H + (un * 2) * KG_BYTES + row * 16;
            st_v4(dst, o[0], o[1], o[2], o[3]); st_v4(dst + KG_BYTES, o[4], o[5], o[6], o[7]);
            uint8_t* gd = gG + (un * 2) * KG_BYTES + row * 16;
            stg_v4(gd, o[0], o[1], o[2], o[3]); stg_v4(gd + KG_BYTES, o[4], o[5], o[6], o[7]);
          }
        } else {
          // MLP boundary: d x0 = XR (from `init`, raw x0) + XA (from the skip Linears, act(x0)) * act'(x0 raw)
          const int units = (int)P.k0_pad >> 4;
          for (int un = cq; un < units; un += 4) {
            uint32_t vr[16], va[16];
            tmem_ld16(t_lane + COL_XR + un * 16, vr);
            if (P.has_xa) tmem_ld16(t_lane + COL_XA + un * 16, va);
            const uint4 s0 = *reinterpret_cast<const uint4*>(s.S + (un * 2) * KG_BYTES + row * 16);
            const uint4 s1 = *reinterpret_cast<const uint4*>(s.S + (un * 2 + 1) * KG_BYTES + row * 16);
            tmem_ld_wait(); reg_fence16(vr); if (P.has_xa) reg_fence16(va);
            const uint32_t sw[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            float d[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              d[k] = __uint_as_float(vr[k]);
              if (P.has_xa) {
                const float x = (k & 1) ? h_hi(sw[k >> 1]) : h_lo(sw[k >> 1]);
                d[k] += act_grad(__uint_as_float(va[k]), P.act == NF_ACT_SIN ? __cosf(x) : x, P.act);
              }
            }
            if (P.kind == PH_END) {
              // the first MLP's x0 (tensor order [hash feats(4L), p, p, pad]): the feature columns leave as fp32, unscaled
              if (un < 2) {
                float4* o = reinterpret_cast<float4*>(a.dx0_out + ((size_t)g * ROWS + row) * 32 + un * 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = make_float4(d[4 * k] * invS, d[4 * k + 1] * invS, d[4 * k + 2] * invS, d[4 * k + 3] * invS);
              }
            } else if (P.hcols2 && un * 16 >= P.inter && un * 16 < P.inter + P.hcols2) {
              // the Positional head's x0 (tensor order [inter(I), hash'(4L), p, p, pad]): its own encoder's feature columns leave as fp32, unscaled
              float4* o = reinterpret_cast<float4*>(a.dx0b_out + ((size_t)g * ROWS + row) * 32 + (un * 16 - P.inter));
#pragma unroll
              for (int k = 0; k < 4; ++k) o[k] = make_float4(d[4 * k] * invS, d[4 * k + 1] * invS, d[4 * k + 2] * invS, d[4 * k + 3] * invS);
            } else if (un * 16 < P.inter) {
              // columns [0, I) of the RGB head's x0 are the density MLP's `intermediate` outputs (tensor order [inter(I), sigma])
              uint32_t o[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] = pack_h2(d[2 * k], d[2 * k + 1]);
              uint8_t* dst = s.H + (un * 2) * KG_BYTES + row * 16;
              st_v4(dst, o[0], o[1], o[2], o[3]); st_v4(dst + KG_BYTES, o[4], o[5], o[6], o[7]);
              uint8_t* gd = gG + (un * 2) * KG_BYTES + row * 16;
              stg_v4(gd, o[0], o[1], o[2], o[3]); stg_v4(gd + KG_BYTES, o[4], o[5], o[6], o[7]);
            }
          }
          if (P.kind == PH_MID && cq == 3) {
            // column I = the raw density (gradient from the composite), the padding columns are zero
            for (int un = P.inter >> 4; un < ((int)P.g_cols >> 4); ++un) {
              const float ds = (un == (P.inter >> 4) && valid) ? __ldg(a.d_sigma + ray * a.T + t) * S : 0.f;
              const uint32_t o0 = pack_h2(ds, 0.f);
              uint8_t* dst = s.H + (un * 2) * KG_BYTES + row * 16;
              st_v4(dst, o0, 0, 0, 0); st_v4(dst + KG_BYTES, 0, 0, 0, 0);
              uint8_t* gd = gG + (un * 2) * KG_BYTES + row * 16;
              stg_v4(gd, o0, 0, 0, 0); stg_v4(gd + KG_BYTES, 0, 0, 0, 0);
            }
          }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (p > 0) mbar_arrive(smem_u32(&s.s_empty));
          if (p + 1 < np) mbar_arrive(smem_u32(&s.a_ready));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == BW_EPIW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// host: the reverse program
bool build_bw_prog(const NfPlan& plan, const NfTrainPlan& tp, BwProg* P) {
  *P = BwProg{};
  const int n = tp.n_lin;
  int nj = 0;
  P->n_phases = n + 1;
  // phase p >= 1 consumes Linear li = n - p; the jobs after phase p are those of Linear n - 1 - p
  for (int p = 0; p <= n; ++p) {
    BwPhase& Ph = P->ph[p];
    Ph.job0 = nj;
    if (p < n) {
      const int li = n - 1 - p;
      const NfTrainLin& L = tp.lin[li];
      const NfLinPlan& LP = plan.mlp[L.m].lin[L.j];
      const uint32_t ks = (uint32_t)L.n_pad >> 4;
      // is there a later (in forward order) skip Linear in the same MLP?  then XA already holds a partial sum
      bool xa_started = false;
      for (int k = li + 1; k < n && tp.lin[k].m == L.m; ++k) if (tp.lin[k].k0_pad && !tp.lin[k].x0_raw) xa_started = true;
      auto add = [&](uint32_t ncols, uint32_t w_off, uint32_t d_col, bool acc0) {
        BwJob& J = P->job[nj++];
        J.k_steps = ks; J.step_bytes = ncols * 32u; J.spc = (uint32_t)BW_STAGE / J.step_bytes; if (J.spc > ks) J.spc = ks;
        J.w_off = w_off; J.d_col = d_col; J.acc0 = acc0 ? 1u : 0u;
        J.idesc = (1u << 4) | ((ncols >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
        const uint32_t b_lbo = ncols * 16u;
        J.bhi = (b_lbo >> 4) << 16; J.bstep4 = (2u * b_lbo) >> 4;
      };
      if (LP.w16t_off + (int64_t)L.n_pad * (L.k0_pad + L.k_hidden) * 2 >= (1LL << 32)) return false;
      if (L.k_hidden) add(NF_HIDDEN, (uint32_t)(LP.w16t_off + (int64_t)L.n_pad * L.k0_pad * 2), COL_MAIN, true);
      if (L.k0_pad) add((uint32_t)L.k0_pad, (uint32_t)LP.w16t_off, L.x0_raw ? COL_XR : COL_XA, L.x0_raw ? true : !xa_started);
      if (L.k0_pad > 128 || nj > BW_MAX_JOBS) return false;
    }
    Ph.n_jobs = nj - Ph.job0;
    if (p == 0) {
      Ph.kind = PH_DY; Ph.tiny = plan.kind == NF_KIND_TINY ? 1 : 0;
      const NfTrainLin& L = tp.lin[n - 1];
      Ph.g_off256 = (uint32_t)(L.g_off >> 8); Ph.g_tile256 = (uint32_t)(L.g_tile >> 8); Ph.g_cols = (uint32_t)L.n_pad;
      if (L.n_pad != 16) return false;
    } else {
      const int li = n - p;
      const NfTrainLin& L = tp.lin[li];
      Ph.act = L.act;
      if (!L.x0_raw) {
        Ph.kind = PH_HIDDEN;
        const int64_t src = L.act == NF_ACT_SIN ? L.c_off : L.a_off + (int64_t)L.k0_pad * 256;
        Ph.s_off256 = (uint32_t)(src >> 8); Ph.s_tile256 = L.act == NF_ACT_SIN ? 256u : (uint32_t)(L.a_tile >> 8); Ph.s_bytes = 65536;
      } else {
        Ph.kind = li == 0 ? PH_END : PH_MID;
        Ph.s_off256 = (uint32_t)(L.a_off >> 8); Ph.s_tile256 = (uint32_t)(L.a_tile >> 8); Ph.s_bytes = (uint32_t)L.k0_pad * 256u;
        Ph.k0_pad = (uint32_t)L.k0_pad; Ph.inter = plan.intermediate;
        if (Ph.kind == PH_MID && plan.refl_kind == NF_REFL_POSITIONAL) Ph.hcols2 = plan.hash_levels * 4;
        for (int k = li + 1; k < n && tp.lin[k].m == L.m; ++k) if (tp.lin[k].k0_pad) Ph.has_xa = 1;
      }
      if (li > 0) {
        const NfTrainLin& Lp = tp.lin[li - 1];
        Ph.g_off256 = (uint32_t)(Lp.g_off >> 8); Ph.g_tile256 = (uint32_t)(Lp.g_tile >> 8); Ph.g_cols = (uint32_t)Lp.n_pad;
        if (Ph.kind == PH_MID && (Lp.n_pad < plan.intermediate + 1 || (plan.intermediate & 15))) return false;
      }
      if ((L.a_off >> 8) >= (1LL << 32) || (L.g_off >> 8) >= (1LL << 32) || (L.c_off >> 8) >= (1LL << 32)) return false;
    }
  }
  P->n_jobs = nj;
  return true;
}

// =====================================================================================================================
// k_bwd_dw: dWt_l += G_l^T A_l over tiles, MN-major operands straight from the stashes, CTA pairs (cta_group::2)
// =====================================================================================================================
// A 256 x 256 fp32 dW block is exactly the TMEM of one SM, so a single CTA would have to cut it in two and read one operand
// twice.  A CTA pair holds it whole (M = 256 across the pair: rank r accumulates rows 128 r ..), and each CTA stages only ITS
// half of both operands -- G features [128 r, 128 r + 128) and A-stash columns [N/2 r, N/2 r + N/2) -- so every stash byte is
// read from HBM exactly once.
constexpr int DW_STAGE_G = 32 * 1024, DW_STAGE_B = 32 * 1024, DW_NST = 3;
constexpr int DW_THREADS = 32 * (DW_NST + 5);      // one loader warp per ring stage, issuer (leader CTA), 4 writers
constexpr int DW_MAX_ITEMS = 64, DW_MAX_PAIRS = 80;

struct DwSmem {
  uint8_t G[DW_NST][DW_STAGE_G];
  uint8_t B[DW_NST][DW_STAGE_B];
  uint8_t ones[ROWS * 16];                         // [1 K-group of 8 columns][128 samples][8] halves of 1.0: this CTA's half of the N = 16 bias operand
  unsigned long long land[DW_NST], full[DW_NST], empty[DW_NST], acc_full, d_free;
  uint32_t tmem_base; uint32_t pad_;
};
static_assert(sizeof(DwSmem) <= 227 * 1024, "dW smem");
struct __align__(16) DwItem {     // one [n_pad rows] x [n columns] block of a Linear's dWt
  uint32_t g_off256, g_tile256, b_off256, b_tile256;
  uint32_t g_sub[2], g_bytes[2];                         // per CTA rank: byte offset within a G tile, bytes (0: nothing for this rank)
  uint32_t b_sub[2], b_bytes[2];                         // per CTA rank: byte offset within an A-stash tile, bytes
  uint32_t n, bias, rows, ld;                            // MMA N; 1 = db rides along; valid rows (n_pad); leading dimension of dWt
  uint32_t out_off4, db_off4, pad0_, pad1_;              // float offsets (from ws) of dWt[0][col0] and db[0]
};
struct __align__(16) DwProg {
  int32_t n_items, n_pairs, pad0_, pad1_;
  DwItem item[DW_MAX_ITEMS];
  int32_t pair_item[DW_MAX_PAIRS + 1]; int32_t pair_tile[DW_MAX_PAIRS + 1];   // pair c runs (item, tile) from [c] up to [c+1]
};
struct DwArgs { uint8_t* ws; long long n_tiles; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DW_THREADS, 1)
k_bwd_dw(const __grid_constant__ DwProg prog, const DwArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  DwSmem& s = *reinterpret_cast<DwSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  for (int i = threadIdx.x; i < ROWS * 4; i += DW_THREADS) reinterpret_cast<uint32_t*>(s.ones)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < DW_NST; ++i) { mbar_init(smem_u32(&s.land[i]), 1); mbar_init(smem_u32(&s.full[i]), 2); mbar_init(smem_u32(&s.empty[i]), 1); }
    mbar_init(smem_u32(&s.acc_full), 1); mbar_init(smem_u32(&s.d_free), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;
  const int it_b = prog.pair_item[pair], it_e = prog.pair_item[pair + 1];
  const long long t_b = prog.pair_tile[pair], t_e = prog.pair_tile[pair + 1];

  if (warp < DW_NST) {
    // ================= loaders (both CTAs), one warp per ring stage: this rank's halves of G and of the A stash =================
    // (a loader waits for its own copy to land before it tells the leader, so one thread per stage keeps DW_NST copies in flight)
    if (elect_one()) {
      const uint32_t mine = (uint32_t)warp;
      uint32_t stage = 0, use = 0;
      for (int it = it_b; it <= it_e && it < prog.n_items; ++it) {
        const DwItem& I = prog.item[it];
        const long long t0 = it == it_b ? t_b : 0, t1 = it == it_e ? t_e : a.n_tiles;
        const uint32_t gb = I.g_bytes[crank], bb = I.b_bytes[crank];
        for (long long t = t0; t < t1; ++t) {
          if (stage == mine) {
            const uint32_t par = (use / DW_NST) & 1u;
            mbar_wait(smem_u32(&s.empty[stage]), par ^ 1u);
            mbar_expect_tx(smem_u32(&s.land[stage]), gb + bb);
            if (gb) bulk_g2s(smem_u32(s.G[stage]), a.ws + ((size_t)I.g_off256 + (size_t)t * I.g_tile256) * 256 + I.g_sub[crank], gb, smem_u32(&s.land[stage]));
            if (bb) bulk_g2s(smem_u32(s.B[stage]), a.ws + ((size_t)I.b_off256 + (size_t)t * I.b_tile256) * 256 + I.b_sub[crank], bb, smem_u32(&s.land[stage]));
            mbar_wait(smem_u32(&s.land[stage]), par);                         // landed in THIS CTA ...
            mbar_arrive_cluster_relaxed(leader_addr(smem_u32(&s.full[stage])));   // ... tell the leader's MMA thread
          }
          ++use; if (++stage == DW_NST) stage = 0;
        }
      }
    }
  } else if (warp == DW_NST) {
    // ================= MMA issuer (leader CTA) =================
    if (crank == 0 && elect_one()) {
      uint32_t stage = 0, use = 0, free_par = 0; bool first_run = true;
      // MN-major canonical layout of a [K/8][128][8] image read "transposed": cores of 8 samples x 8 features, 128 B each;
      // along the samples (the MMA's K) cores are 128 B apart (LBO), along the features (M / N) 2048 B apart (SBO)
      const uint64_t hi = ((uint64_t)(0x4000u | ((uint32_t)KG_BYTES >> 4)) << 32) | ((uint64_t)(128u >> 4) << 16);
      for (int it = it_b; it <= it_e && it < prog.n_items; ++it) {
        const DwItem& I = prog.item[it];
        const long long t0 = it == it_b ? t_b : 0, t1 = it == it_e ? t_e : a.n_tiles;
        if (t0 >= t1) continue;
        if (!first_run) { mbar_wait(smem_u32(&s.d_free), free_par); free_par ^= 1u; tc_fence_after(); }
        first_run = false;
        const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((I.n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint32_t idesc1 = (1u << 4) | (1u << 15) | (1u << 16) | ((16u >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        for (long long t = t0; t < t1; ++t) {
          mbar_wait(smem_u32(&s.full[stage]), (use / DW_NST) & 1u);
          tc_fence_after();
          const uint32_t g4 = smem_u32(s.G[stage]) >> 4, b4 = smem_u32(s.B[stage]) >> 4, o4 = smem_u32(s.ones) >> 4;
#pragma unroll 1
          for (uint32_t k = 0; k < 8; ++k) {                     // 16 samples per MMA
            const uint32_t acc = (t > t0 || k > 0) ? 1u : 0u;
            umma2_f16(tmem, hi | (uint64_t)(g4 + k * 16u), hi | (uint64_t)(b4 + k * 16u), idesc, acc);
            if (I.bias) umma2_f16(tmem + I.n, hi | (uint64_t)(g4 + k * 16u), hi | (uint64_t)(o4 + k * 16u), idesc1, acc);
          }
          umma2_commit_mc(smem_u32(&s.empty[stage]));
          ++use; if (++stage == DW_NST) stage = 0;
        }
        umma2_commit_mc(smem_u32(&s.acc_full));
      }
    }
  } else {
    // ================= writers (both CTAs): TMEM -> fp32 atomics into dWt / db; rank r owns rows 128 r .. =================
    const int q = warp & 3;
    const int row = (int)crank * ROWS + q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    float* wsf = reinterpret_cast<float*>(a.ws);
    uint32_t par = 0;
    for (int it = it_b; it <= it_e && it < prog.n_items; ++it) {
      const DwItem& I = prog.item[it];
      const long long t0 = it == it_b ? t_b : 0, t1 = it == it_e ? t_e : a.n_tiles;
      if (t0 >= t1) continue;
      mbar_wait_suspend(smem_u32(&s.acc_full), par); par ^= 1u;
      tc_fence_after();
      const bool rok = (uint32_t)row < I.rows;
      for (uint32_t c0 = 0; c0 < I.n; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_lane + c0, v); tmem_ld_wait(); reg_fence16(v);
        if (rok) {
          float* dst = wsf + I.out_off4 + (size_t)row * I.ld + c0;
#pragma unroll
          for (int k = 0; k < 16; ++k) atomicAdd(dst + k, __uint_as_float(v[k]));
        }
      }
      if (I.bias) {
        uint32_t v[16];
        tmem_ld16(t_lane + I.n, v); tmem_ld_wait(); reg_fence16(v);
        if (rok) atomicAdd(wsf + I.db_off4 + row, __uint_as_float(v[0]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(leader_addr(smem_u32(&s.d_free)));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}
static_assert(((DW_NST + 1) & 3) == 0, "the four writer warps must cover the four TMEM lane quarters (warp % 4)");

bool build_dw_prog(const NfTrainPlan& tp, int n_pairs, DwProg* P) {
  *P = DwProg{};
  int ni = 0;
  long long cost[DW_MAX_ITEMS];
  for (int li = 0; li < tp.n_lin; ++li) {
    const NfTrainLin& L = tp.lin[li];
    const int kt = L.k0_pad + L.k_hidden;
    bool bias_done = false;
    // column blocks: the x0 part (<= 256 columns), the hidden part (256)
    const int blocks[2][2] = {{0, L.k0_pad}, {L.k0_pad, L.k_hidden}};
    for (int b = 0; b < 2; ++b) {
      const int c0 = blocks[b][0], nc = blocks[b][1];
      if (!nc) continue;
      if (ni >= DW_MAX_ITEMS || nc > 256 || (nc & 15) || L.n_pad > 256) return false;
      DwItem& I = P->item[ni];
      I.g_off256 = (uint32_t)(L.g_off >> 8); I.g_tile256 = (uint32_t)(L.g_tile >> 8);
      I.b_off256 = (uint32_t)(L.a_off >> 8); I.b_tile256 = (uint32_t)(L.a_tile >> 8);
      for (int r = 0; r < 2; ++r) {
        const int rows = L.n_pad - 128 * r < 0 ? 0 : (L.n_pad - 128 * r < 128 ? L.n_pad - 128 * r : 128);
        I.g_sub[r] = (uint32_t)(128 * r) * 256u; I.g_bytes[r] = (uint32_t)rows * 256u;
        I.b_sub[r] = (uint32_t)(c0 + r * (nc / 2)) * 256u; I.b_bytes[r] = (uint32_t)(nc / 2) * 256u;
      }
      I.n = (uint32_t)nc; I.bias = bias_done ? 0u : 1u; bias_done = true; I.rows = (uint32_t)L.n_pad; I.ld = (uint32_t)kt;
      const int64_t o = L.dw_off / 4 + c0, ob = L.db_off / 4;
      if (o >= (1LL << 32) || ob >= (1LL << 32) || (L.a_off >> 8) >= (1LL << 32) || (L.g_off >> 8) >= (1LL << 32)) return false;
      I.out_off4 = (uint32_t)o; I.db_off4 = (uint32_t)ob;
      cost[ni] = (long long)(L.n_pad + nc) + 32;           // bytes / 256 per tile (+ a fixed per-tile overhead)
      ++ni;
    }
  }
  P->n_items = ni;
  if (n_pairs > DW_MAX_PAIRS) n_pairs = DW_MAX_PAIRS;
  long long total = 0;
  for (int i = 0; i < ni; ++i) total += cost[i] * tp.n_tiles;
  if (total == 0) { P->n_pairs = 0; return true; }
  if ((long long)n_pairs > (long long)ni * tp.n_tiles) n_pairs = (int)((long long)ni * tp.n_tiles);
  P->n_pairs = n_pairs;
  // pair c starts at the first (item, tile) whose preceding cost reaches c * total / n_pairs
  int it = 0; long long acc = 0;                       // acc = cost of all items before `it`
  for (int c = 0; c <= n_pairs; ++c) {
    const long long want = c == n_pairs ? total : (total * c) / n_pairs;
    while (it < ni && acc + cost[it] * tp.n_tiles <= want) { acc += cost[it] * tp.n_tiles; ++it; }
    long long tile = it < ni ? (want - acc + cost[it] - 1) / cost[it] : 0;
    if (it < ni && tile >= tp.n_tiles) { acc += cost[it] * tp.n_tiles; ++it; tile = 0; }
    if (tile >= (1LL << 31)) return false;
    P->pair_item[c] = it; P->pair_tile[c] = (int32_t)tile;
  }
  return true;
}

// =====================================================================================================================
// gradients in the reference's parameter layout; hash-table scatter
// =====================================================================================================================
struct UnpackArgs { const float* dwt[PACK_MAX]; const float* db[PACK_MAX]; float* gW[PACK_MAX]; float* gb[PACK_MAX]; int32_t mj[PACK_MAX]; };
__global__ void k_unpack_grads(const __grid_constant__ NfPlan plan, const __grid_constant__ UnpackArgs a, const float* __restrict__ scale) {
  const int li = blockIdx.y, m = a.mj[li] >> 4, j = a.mj[li] & 15;
  const float* __restrict__ dwt = a.dwt[li]; const float* __restrict__ db = a.db[li];
  float* __restrict__ gW = a.gW[li]; float* __restrict__ gb = a.gb[li];
  const NfLinPlan& L = plan.mlp[m].lin[j];
  const int k_ref_total = L.k_hidden + L.k_x0, ld = L.k0_pad + L.k_hidden;
  const float invS = __ldg(scale + 1);
  const int total = L.n * k_ref_total;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total + L.n; i += gridDim.x * blockDim.x) {
    if (i < total) {
      const int n_ref = i / k_ref_total, k_ref = i - n_ref * k_ref_total;
      const int k_tc = k_ref < L.k_hidden ? L.k0_pad + k_ref : nf_x0_perm(plan, m, k_ref - L.k_hidden);
      const int n_tc = L.is_out ? nf_out_perm(plan, m, n_ref) : n_ref;
      if (gW) gW[i] = dwt[(size_t)n_tc * ld + k_tc] * invS;
    } else if (gb) {
      const int n_ref = i - total;
      gb[n_ref] = db[L.is_out ? nf_out_perm(plan, m, n_ref) : n_ref] * invS;
    }
  }
}

struct HashGradPtrs { float* t[16]; };
// one thread per (tile row, level): recompute the corners and trilinear weights of the sample, scatter-add w_c * d feat
__global__ void k_hash_bwd_tiles(const __grid_constant__ NfPlan plan, const float* __restrict__ rays, long long n_rays, const float* __restrict__ ts,
                                 int T, long long ts_stride, long long n_tiles, int tpr, const float* __restrict__ dx0, HashGradPtrs out) {
  const int L = plan.hash_levels;
  const NfStreamMap map(T, NF_TC_ROWS);
  const long long total = n_tiles * NF_TC_ROWS * L;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long rr = i / L; const int lvl = (int)(i - rr * L);
    const long long g = rr / NF_TC_ROWS; const int row = (int)(rr - g * NF_TC_ROWS);
    const long long u = g / tpr; const int sub = (int)(g - u * tpr);
    long long ray; int t;
    if (!map.locate(u, sub, row, n_rays, ray, t)) continue;
    if (!out.t[lvl]) continue;
    const float* r6 = rays + ray * 6;
    const float tt = __ldg(ts + ray * ts_stride + t);
    const float px = nf_pt(__ldg(r6), tt, __ldg(r6 + 3)), py = nf_pt(__ldg(r6 + 1), tt, __ldg(r6 + 4)), pz = nf_pt(__ldg(r6 + 2), tt, __ldg(r6 + 5));
    const float res = plan.hash_res[lvl];
    const float vx = __fmul_rn(px, res), vy = __fmul_rn(py, res), vz = __fmul_rn(pz, res);
    const float fx = floorf(vx), fy = floorf(vy), fz = floorf(vz);
    const uint32_t ix = (uint32_t)(int32_t)fx, iy = (uint32_t)(int32_t)fy, iz = (uint32_t)(int32_t)fz;
    const float wx = vx - fx, wy = vy - fy, wz = vz - fz;
    const float4 gq = __ldg(reinterpret_cast<const float4*>(dx0 + rr * 32) + lvl);
    float4* table = reinterpret_cast<float4*>(out.t[lvl]);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int bx = (c >> 2) & 1, by = (c >> 1) & 1, bz = c & 1;
      const uint32_t id = (((ix + bx) * plan.hash_primes[0]) ^ ((iy + by) * plan.hash_primes[1]) ^ ((iz + bz) * plan.hash_primes[2])) & plan.hash_mask;
      const float w = (bx ? wx : 1.f - wx) * (by ? wy : 1.f - wy) * (bz ? wz : 1.f - wz);
      atomicAdd(table + id, make_float4(w * gq.x, w * gq.y, w * gq.z, w * gq.w));
    }
  }
}

}  // namespace

// params: W, b per Linear in the order of nf_pack_weights (MLP by MLP); one memset over the blob's MLP region + one launch
cudaError_t nf_launch_pack_all(const NfPlan& plan, const float* const* params, void* packed, cudaStream_t st) {
  PackArgs a{}; int n = 0, max_total = 0;
  for (int m = 0; m < plan.n_mlps; ++m)
    for (int j = 0; j < plan.mlp[m].n_lin; ++j, ++n) {
      if (n >= PACK_MAX) return cudaErrorInvalidValue;
      a.W[n] = params[2 * n]; a.b[n] = params[2 * n + 1]; a.mj[n] = m * 16 + j;
      const NfLinPlan& L = plan.mlp[m].lin[j];
      const int total = (L.k_hidden + L.k_x0) * L.n_pad;
      max_total = total > max_total ? total : max_total;
    }
  const int64_t lo = plan.mlp[0].lin[0].wt_off;                 // the Linears' images are the tail of the blob (nf_build_plan)
  cudaError_t e = cudaMemsetAsync((uint8_t*)packed + lo, 0, (size_t)(plan.total_bytes - lo), st);
  if (e != cudaSuccess) return e;
  const int gx = (max_total + 255) / 256;
  k_pack_all<<<dim3(gx < 96 ? gx : 96, n), 256, 0, st>>>(plan, a, (uint8_t*)packed);
  return cudaGetLastError();
}
// n tables of `bytes` each (a multiple of 16): src[i] -> dst[i]
cudaError_t nf_launch_copy_tables(const float* const* src, float* const* dst, int n, size_t bytes, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  if (n > 48 || (bytes & 15)) return cudaErrorInvalidValue;
  CopyArgs a{};
  for (int i = 0; i < n; ++i) { a.src[i] = (const float4*)src[i]; a.dst[i] = (float4*)dst[i]; }
  const long long n4 = (long long)(bytes >> 4);
  const long long want = (n4 + 255) / 256;
  k_copy_tables<<<dim3((unsigned)(want < 64 ? want : 64), n), 256, 0, st>>>(a, n4);
  return cudaGetLastError();
}

// grads: one pointer per parameter in nf_pack_weights order (nullable entries are skipped): W, b per Linear (MLP order), then
// the hash tables.  Every non-null gradient is OVERWRITTEN.
// The hash-table scatter (atomic-bound, ~0.3 ms, little HBM traffic) and the dW pass (HBM-bound, ~1.1 ms) both depend only on the
// chain kernel: the scatter runs on a side stream forked from / joined to the caller's stream (events; legal under graph capture),
// next to the dW pass.  One side stream and two events per host thread and device, created on first use.
struct BwSide { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static BwSide* bw_side() {
  thread_local BwSide side[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  BwSide& b = side[dev];
  if (!b.s) {
    if (cudaStreamCreateWithFlags(&b.s, cudaStreamNonBlocking) != cudaSuccess) { b.s = nullptr; return nullptr; }
    if (cudaEventCreateWithFlags(&b.fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&b.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &b;
}

cudaError_t nf_launch_render_backward(const NfPlan& plan, const NfTrainPlan& tp, const void* packed, void* ws_, const float* rays,
                                      const float* ts, int64_t ts_stride, const float* d_rgb, float* const* grads, cudaStream_t st) {
  if (nf_train_unsupported(plan)) return cudaErrorNotSupported;
  uint8_t* ws = (uint8_t*)ws_;
  if (tp.n_tiles == 0) return cudaSuccess;
  float* scale = (float*)(ws + tp.scale_off);
  float* sigma = (float*)(ws + tp.sigma_off); float* rgbraw = (float*)(ws + tp.rgbraw_off);
  float* dsigma = (float*)(ws + tp.dsigma_off); float* drgbraw = (float*)(ws + tp.drgbraw_off);
  float* dx0 = (float*)(ws + tp.dx0_off);
  cudaError_t e;
  if ((e = cudaMemsetAsync(ws + tp.dw_begin, 0, (size_t)(tp.dw_end - tp.dw_begin), st)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(scale, 0, 16, st)) != cudaSuccess) return e;
  // 1. composite backward with the feature activation's derivative folded in (VolSDF: + the gradient of the learned beta, the last
  //    parameter of nf_pack_weights' order)
  float* d_beta = nullptr;
  if (plan.density_act == NF_DENS_LAPLACE) {
    int pi_b = 0;
    for (int m = 0; m < plan.n_mlps; ++m) pi_b += 2 * plan.mlp[m].n_lin;
    if (plan.enc == NF_ENC_HASH) pi_b += plan.hash_levels;
    if (plan.enc == NF_ENC_FOURIER) pi_b += 1;
    d_beta = grads[pi_b];
    if (d_beta && (e = cudaMemsetAsync(d_beta, 0, sizeof(float), st)) != cudaSuccess) return e;
  }
  if ((e = nf_launch_composite_bwd(plan, packed, sigma, rgbraw, rays, tp.n_rays, ts, tp.T, ts_stride, d_rgb, dsigma, drgbraw, st, plan.feat_act, d_beta,
                                   plan.bg == NF_BG_RANDOM ? (const float*)(ws + tp.bgrand_off) : nullptr)) != cudaSuccess) return e;
  // 2. loss scale
  const long long ns = tp.n_rays * tp.T;
  const int sms = tr_num_sms();
  k_grad_absmax<<<sms * 4, 256, 0, st>>>(dsigma, ns, drgbraw, ns * 3, reinterpret_cast<unsigned*>(scale) + 2);
  k_grad_scale<<<1, 1, 0, st>>>(scale);
  // 3. dX through the chain
  {
    BwProg prog;
    if (!build_bw_prog(plan, tp, &prog)) return cudaErrorNotSupported;
    BwArgs a{};
    a.packed = (const uint8_t*)packed; a.ws = ws; a.n_rays = tp.n_rays; a.n_tiles = tp.n_tiles; a.T = tp.T; a.rpu = tp.rpu; a.tpr = tp.tpr;
    a.d_sigma = dsigma; a.d_rgbraw = drgbraw; a.dx0_out = dx0; a.scale = scale;
    a.dx0b_out = dx0 + (size_t)tp.n_tiles * NF_TC_ROWS * 32;
    if ((e = cudaFuncSetAttribute(k_bwd_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwSmem))) != cudaSuccess) return e;
    const int grid = (int)(tp.n_tiles < sms ? tp.n_tiles : sms);
    k_bwd_chain<<<grid, BW_THREADS, sizeof(BwSmem), st>>>(prog, a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  // 3b. fork: the hash-table scatter next to the dW pass
  BwSide* side = NF_BWD_SIDE_STREAM ? bw_side() : nullptr;
  cudaStream_t hs = st;
  if (side && plan.enc == NF_ENC_HASH) {
    if ((e = cudaEventRecord(side->fork, st)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(side->s, side->fork, 0)) != cudaSuccess) return e;
    hs = side->s;
  }
  if (plan.enc == NF_ENC_HASH) {
    int pi_h = 0;
    for (int m = 0; m < plan.n_mlps; ++m) pi_h += 2 * plan.mlp[m].n_lin;
    HashGradPtrs hp{};
    bool any = false;
    const size_t per = (size_t)(plan.hash_mask + 1) * 4 * sizeof(float);
    for (int l = 0; l < plan.hash_levels; ++l) {
      hp.t[l] = grads[pi_h++];
      if (hp.t[l]) { any = true; if ((e = cudaMemsetAsync(hp.t[l], 0, per, hs)) != cudaSuccess) return e; }
    }
    if (any) {
      const long long total = tp.n_tiles * NF_TC_ROWS * plan.hash_levels;
      const long long want = (total + 255) / 256;
      const int grid = (int)(want < (long long)sms * 16 ? want : (long long)sms * 16);
      k_hash_bwd_tiles<<<grid, 256, 0, hs>>>(plan, rays, tp.n_rays, ts, tp.T, ts_stride, tp.n_tiles, tp.tpr, dx0, hp);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (plan.refl_kind == NF_REFL_POSITIONAL) {
      // the head's own encoder (refl.py:233-237): same sample positions and level resolutions, its own tables (the next `levels` parameters)
      HashGradPtrs hp2{};
      bool any2 = false;
      for (int l = 0; l < plan.hash_levels; ++l) {
        hp2.t[l] = grads[pi_h++];
        if (hp2.t[l]) { any2 = true; if ((e = cudaMemsetAsync(hp2.t[l], 0, per, hs)) != cudaSuccess) return e; }
      }
      if (any2) {
        const long long total = tp.n_tiles * NF_TC_ROWS * plan.hash_levels;
        const long long want = (total + 255) / 256;
        const int grid = (int)(want < (long long)sms * 16 ? want : (long long)sms * 16);
        k_hash_bwd_tiles<<<grid, 256, 0, hs>>>(plan, rays, tp.n_rays, ts, tp.T, ts_stride, tp.n_tiles, tp.tpr, dx0 + (size_t)tp.n_tiles * NF_TC_ROWS * 32, hp2);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
      }
    }
    if (hs != st && (e = cudaEventRecord(side->join, hs)) != cudaSuccess) return e;
  }
  // 4. dW / db
  {
    DwProg prog;
    if (!build_dw_prog(tp, sms / 2, &prog)) return cudaErrorNotSupported;
    DwArgs a{}; a.ws = ws; a.n_tiles = tp.n_tiles;
    if ((e = cudaFuncSetAttribute(k_bwd_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DwSmem))) != cudaSuccess) return e;
    if (prog.n_pairs > 0) k_bwd_dw<<<2 * prog.n_pairs, DW_THREADS, sizeof(DwSmem), st>>>(prog, a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  // 5. reference layout
  int pi = 0;
  {
    UnpackArgs ua{}; int nu = 0, max_total = 0;
    for (int m = 0; m < plan.n_mlps; ++m)
      for (int j = 0; j < plan.mlp[m].n_lin; ++j) {
        float* gW = grads[pi++]; float* gb = grads[pi++];
        if (!gW && !gb) continue;
        int li = -1;
        for (int k = 0; k < tp.n_lin; ++k) if (tp.lin[k].m == m && tp.lin[k].j == j) li = k;
        const NfLinPlan& L = plan.mlp[m].lin[j];
        const int total = L.n * (L.k_hidden + L.k_x0) + L.n;
        max_total = total > max_total ? total : max_total;
        ua.dwt[nu] = (const float*)(ws + tp.lin[li].dw_off); ua.db[nu] = (const float*)(ws + tp.lin[li].db_off);
        ua.gW[nu] = gW; ua.gb[nu] = gb; ua.mj[nu] = m * 16 + j; ++nu;
      }
    if (nu > 0) k_unpack_grads<<<dim3((max_total + 255) / 256 < 64 ? (max_total + 255) / 256 : 64, nu), 256, 0, st>>>(plan, ua, scale);
  }
  // join
  if (hs != st && (e = cudaStreamWaitEvent(st, side->join, 0)) != cudaSuccess) return e;
  return cudaGetLastError();
}
