// nf_tc2.cu -- the paired (cta_group::2) form of the tensor-core render pipeline: two tiles in flight per CTA.
//
// Why (profiles/r01_*): with ONE 128-sample tile per CTA the MMA issue, the epilogue and the
// barrier handshakes of a layer serialise (MMA(j+1) needs EPI(j), EPI(j+1) needs MMA(j+1)).  A second tile in flight
// hides all of that, but two tiles of activations (2 x 64 KB) only fit beside the weight ring if the ring is halved --
// which is what a CTA pair gives: tcgen05.mma.cta_group::2 computes [256 x N] per instruction (128 rows per SM) while
// each CTA stages only HALF of every weight chunk (N/2 rows, 16 KB per 64 K-columns).
//
// Cluster of 2 CTAs, 640 threads each:
//   warps 0-15   encode + epilogue: all sixteen work on whichever slot is in its epilogue (lane quarter = warp % 4,
//                column quarter = warp / 4), so an epilogue takes half as long as with one 8-warp group per slot
//   warps 16,18,19  weight producers (one ring stage each): bulk-copy this CTA's half chunk, wait for it to land, then
//                arrive on the LEADER's w_ready[stage] (count 2) through the shared::cluster window
//   warp 17      MMA issuer -- only in the leader CTA (rank 0), one thread, M = 256 instructions for both SMs
// (the role warps have the HIGHEST warp ids: the SM's schedulers favour higher ids, and the single MMA-issuing thread
//  must never queue behind the sixteen busy epilogue warps)
// Per slot a single 256-column TMEM accumulator: per layer the slot alternates MMA -> epilogue, and the two slots are
// out of phase, so while slot A is in its epilogue the tensor pipe runs slot B's layer.
// Barriers: w_land/w_empty/acc_full are CTA-local (w_empty, acc_full signalled by multicast tcgen05.commit from the
// leader); w_ready and a_ready[slot] live in the leader and are arrived on remotely by the peer.
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include "nf_common.cuh"
#include "nf_kernels.h"
#include "nf_tc_ptx.cuh"

namespace {
using namespace nf_ptx;

constexpr int X0K = 80;
constexpr int STAGES = 3;
constexpr int SPC = NF_TC_CHUNK_K / 16;                     // UMMA K-steps per weight chunk (4)
constexpr int STAGE_BYTES = NF_TC_CHUNK_K * 128 * 2;       // 16 KB: 64 K-columns x 128 N (this CTA's half) x fp16
constexpr int THREADS = 640;
constexpr int EPI_THREADS = 512;                             // 16 encode/epilogue warps
constexpr int MAX_ENT = 128;                                // ring entries per round (every Linear's chunks, twice)
constexpr int MAX_LIN2 = 16;
constexpr uint32_t F_NSTEP = 7, F_FIRST = 8, F_LAST = 16, F_SLOT = 32, F_WAIT_A = 64;

struct Tc2Smem {
  uint8_t H[2][ROWS * 256 * 2];
  uint8_t X0[2][ROWS * X0K * 2];
  uint8_t W[STAGES][STAGE_BYTES];
  float sig[2][ROWS];
  long long ray[2][ROWS];
  int t[2][ROWS];
  int valid[2][ROWS];
  float wrgb[2][ROWS * 4];
  float warp_agg[2][4]; int warp_cont[2][4]; float warp_sum[2][4][4]; float carry[2][8];
  unsigned long long w_land[STAGES], w_empty[STAGES], w_ready[STAGES], acc_full[2], a_ready[2];
  uint32_t tmem_base; int n_chunks;
  uint2 chunks[MAX_ENT / 2];                                // one round's distinct weight chunks: {byte offset, bytes}
};
static_assert(sizeof(Tc2Smem) <= 227 * 1024, "paired tensor pipeline smem");

// Host-built program (kernel parameter => uniform constant loads in the issuing thread).
struct __align__(16) Tc2Lin { uint32_t k0_steps, h_steps, idesc, bstep4, bhi, pad0_, pad1_, pad2_; };   // one Linear
struct __align__(16) Tc2Prog {
  int32_t n_lin, n_ent, pad0_, pad1_;
  Tc2Lin lin[MAX_LIN2];
  uint8_t ent_chunk[MAX_ENT];     // producers: distinct weight-chunk index of every ring entry of a round
};

struct Tc2Args {
  const uint8_t* packed;
  const float* rays; long long n_rays;
  const float* ts; int T; long long ts_stride;
  const float* noise;
  float* rgb_out; float* alpha_out; float* weights_out;
  long long* trace;   // NF_TC_TRACE builds: [role][512] x {tag, clock64} for one round of cluster 0
  int debug;
};

#ifdef NF_TC_TRACE
#define NF_TRACE2(role, tag) do { if ((NF_TC_TRACE != 2 || (role) >= 2) && tr_on && tr_n < 512) { a.trace[((role) * 512 + tr_n) * 2] = (tag); a.trace[((role) * 512 + tr_n) * 2 + 1] = clock64(); ++tr_n; } } while (0)
#else
#define NF_TRACE2(role, tag) do { } while (0)
#endif

// ---- epilogue of a hidden Linear: H <- fp16(act(acc + bias)) ----------------------------------------------
// Bias of this warp's first 16 columns, fetched BEFORE the acc_full wait (it does not depend on the accumulator): with a
// 227 KB shared-memory carve-out the L1 is nearly gone, so every bias load is a ~300-cycle L2 hit that must be hidden.
struct Bias16 { float4 b[4]; };
__device__ __forceinline__ Bias16 bias_prefetch(const float* __restrict__ bias, int col) {
  Bias16 r; const float4* b4 = reinterpret_cast<const float4*>(bias + col);
#pragma unroll
  for (int i = 0; i < 4; ++i) r.b[i] = __ldg(b4 + i);
  return r;
}
template <int ACT>
__device__ __forceinline__ void epi_hidden2(uint8_t* __restrict__ H, uint32_t t_acc, const float* __restrict__ bias, int cq, int row, Bias16 bcur) {
  // this warp's 64-column quarter = 4 units of 16 columns; unit u+1's TMEM load and bias loads are in flight while unit u
  // is converted and stored
  uint32_t v[2][16];
  tmem_ld16(t_acc + cq * 64, v[0]);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int col = cq * 64 + u * 16;
    Bias16 bnext = bcur;
    if (u < 3) bnext = bias_prefetch(bias, col + 16);
    tmem_ld_wait();
    reg_fence16(v[u & 1]);
    if (u < 3) tmem_ld16(t_acc + col + 16, v[(u + 1) & 1]);
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b = bcur.b[i];
      o[2 * i]     = act_pack_t<ACT>(__uint_as_float(v[u & 1][4 * i]) + b.x, __uint_as_float(v[u & 1][4 * i + 1]) + b.y);
      o[2 * i + 1] = act_pack_t<ACT>(__uint_as_float(v[u & 1][4 * i + 2]) + b.z, __uint_as_float(v[u & 1][4 * i + 3]) + b.w);
    }
    uint8_t* dst = H + (col >> 3) * KG_BYTES + row * 16;
    st_v4(dst, o[0], o[1], o[2], o[3]); st_v4(dst + KG_BYTES, o[4], o[5], o[6], o[7]);
    bcur = bnext;
  }
}

// x0 raw -> act(x0), in place (the `init` Linear consumed the raw form; the skip Linear wants the activated one)
__device__ __forceinline__ void x0_activate(uint8_t* X0, int k0_pad, int act, int g_tid) {
  const int n16 = (k0_pad >> 3) * ROWS;             // 16-byte groups
  for (int i = g_tid; i < n16; i += EPI_THREADS) {
    uint4 q = *reinterpret_cast<uint4*>(X0 + i * 16);
    uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
      w[k] = pack_h2(tc_act(f.x, act), tc_act(f.y, act));
    }
    *reinterpret_cast<uint4*>(X0 + i * 16) = q;
  }
}

// ---- composite of one tile by the 4 column-half-0 warps of a slot's group (thread = row); reference nerf.py:60-80 ----
__device__ __forceinline__ void composite_tile2(Tc2Smem& s, int slot, const NfPlan& plan, const Tc2Args& a, const NfTileMap& map,
                                                int sub, int row, int lane, int q, float cr, float cg, float cb) {
  const bool valid = s.valid[slot][row] != 0;
  const int t = valid ? s.t[slot][row] : 0;
  const long long ray = s.ray[slot][row];
  float al = 0.f;
  if (valid) {
    float sr = s.sig[slot][row];
    if (a.noise) sr += __ldg(a.noise + ray * a.T + t);
    const float* rr = a.rays + ray * 6;
    const float dx = __ldg(rr + 3), dy = __ldg(rr + 4), dz = __ldg(rr + 5);
    const float beta = plan.density_act == NF_DENS_LAPLACE ? __ldg(reinterpret_cast<const float*>(a.packed + plan.scale_off)) : 1.f;
    al = nf_alpha(sr, nf_delta(a.ts + ray * a.ts_stride, t, a.T, sqrtf(dx * dx + dy * dy + dz * dz)), plan.density_act, beta);
  }
  float incl = valid ? (1.f - al) + 1e-10f : 1.f;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d && t >= d) incl *= o;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0 || t == 0) excl = 1.f;
  if (lane == 31) { s.warp_agg[slot][q] = incl; s.warp_cont[slot][q] = t > 31; }
  named_bar(3 + slot, 128);
  float c = 1.f; bool open = true;
  for (int v = q - 1; v >= 0 && open; --v) { c *= s.warp_agg[slot][v]; open = s.warp_cont[slot][v] != 0; }
  if (open && sub > 0) c *= s.carry[slot][0];
  const float trans = excl * (t > lane ? c : 1.f);
  const float w = al * trans;
  if (valid) {
    if (a.alpha_out) a.alpha_out[ray * a.T + t] = al;
    if (a.weights_out) a.weights_out[ray * a.T + t] = w;
  }
  const float wr = w * cr, wg = w * cg, wb = w * cb, wl = (valid && t < a.T - 1) ? w : 0.f;
  const int row_thread = q * 32 + lane;
  float* carry = s.carry[slot];
  const bool fast = (a.T & 31) == 0;
  if (fast) {
    float x0 = wr, x1 = wg, x2 = wb, x3 = wl;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      x0 += __shfl_xor_sync(0xffffffffu, x0, d); x1 += __shfl_xor_sync(0xffffffffu, x1, d);
      x2 += __shfl_xor_sync(0xffffffffu, x2, d); x3 += __shfl_xor_sync(0xffffffffu, x3, d);
    }
    if (lane == 0) { float* ws = s.warp_sum[slot][q]; ws[0] = x0; ws[1] = x1; ws[2] = x2; ws[3] = x3; }
  } else {
    float* wq = s.wrgb[slot] + row_thread * 4;
    wq[0] = wr; wq[1] = wg; wq[2] = wb; wq[3] = wl;
  }
  named_bar(3 + slot, 128);
  const int nseg = a.T <= ROWS ? map.rpt : 1;
  const int row0 = a.T <= ROWS ? row_thread * a.T : 0;
  if (row_thread < nseg && s.valid[slot][row0]) {
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    if (sub > 0) { o0 = carry[1]; o1 = carry[2]; o2 = carry[3]; o3 = carry[4]; }
    if (fast) {
      const int wpr = a.T <= ROWS ? a.T / 32 : 4;
      for (int v = 0; v < wpr; ++v) { const float* ws = s.warp_sum[slot][row_thread * wpr + v]; o0 += ws[0]; o1 += ws[1]; o2 += ws[2]; o3 += ws[3]; }
    } else {
      const int nrow = a.T <= ROWS ? a.T : min(ROWS, a.T - sub * ROWS);
      for (int i = 0; i < nrow; ++i) { const float* x = s.wrgb[slot] + (row0 + i) * 4; o0 += x[0]; o1 += x[1]; o2 += x[2]; o3 += x[3]; }
    }
    const long long r = s.ray[slot][row0];
    if (sub == map.tpr - 1) {
      const float skyv = plan.bg == NF_BG_WHITE ? 1.f - o3 : 0.f;
      a.rgb_out[r * 3 + 0] = o0 + skyv; a.rgb_out[r * 3 + 1] = o1 + skyv; a.rgb_out[r * 3 + 2] = o2 + skyv;
    } else {
      carry[1] = o0; carry[2] = o1; carry[3] = o2; carry[4] = o3;
      carry[0] = (sub > 0 ? carry[0] : 1.f) * s.warp_agg[slot][0] * s.warp_agg[slot][1] * s.warp_agg[slot][2] * s.warp_agg[slot][3];
    }
  }
}

// =====================================================================================================
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_render_tc2(const __grid_constant__ NfPlan plan, const __grid_constant__ Tc2Prog prog, const Tc2Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Tc2Smem& s = *reinterpret_cast<Tc2Smem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const NfTileMap map(a.T, ROWS);
  const long long units = map.units(a.n_rays);
  const long long trips = (units + 2LL * gridDim.x - 1) / (2LL * gridDim.x);     // every CTA, every slot: same trip count
  const long long passes = trips * map.tpr;
  const int lin_base1 = plan.mlp[0].n_lin;
  int tr_n = 0; bool tr_on = false; (void)tr_n; (void)tr_on;

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(smem_u32(&s.w_land[i]), 1); mbar_init(smem_u32(&s.w_empty[i]), 1); mbar_init(smem_u32(&s.w_ready[i]), 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s.acc_full[i]), 1); mbar_init(smem_u32(&s.a_ready[i]), 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 18) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp == 19 && lane == 0) {
    // this CTA's half of every weight chunk of one tile: {byte offset into packed, bytes}
    int nc = 0;
    for (int m = 0; m < plan.n_mlps; ++m)
      for (int j = 0; j < plan.mlp[m].n_lin; ++j) {
        const NfLinPlan& L = plan.mlp[m].lin[j];
        const int steps = (L.k0_pad + L.k_hidden) >> 4, nh = L.n_pad >> 1;
        const int64_t half_bytes = (int64_t)(L.k0_pad + L.k_hidden + 16) * nh * 2;     // the pair images end in a bias K-step (nf_tc3.cu); unused here
        for (int c = 0; c < L.n_chunks; ++c, ++nc)
          s.chunks[nc] = make_uint2((uint32_t)(L.w16h_off + crank * half_bytes + (int64_t)c * (2 * SPC) * nh * 16),
                                    (uint32_t)min(SPC, steps - SPC * c) * 2u * (uint32_t)nh * 16u);
      }
    s.n_chunks = nc;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (s.tmem_base != 0) __trap();

  if (warp == 16 || warp == 18 || warp == 19) {
    // ================= weight producers (both CTAs): one ring stage each =================
    if (lane == 0 && !(a.debug & 64)) {
      const int p = warp == 16 ? 0 : warp - 17;
      const int nc = s.n_chunks;
      // the round's chunk sequence is: for each Linear { its chunks for slot 0, the same chunks again for slot 1 }
      const long long total = passes * prog.n_ent;
      const uint32_t ready_leader = leader_addr(smem_u32(&s.w_ready[p]));
      (void)nc;
      for (long long g = p; g < total; g += STAGES) {
        const int e = (int)(g % prog.n_ent);
        const uint2 ch = s.chunks[prog.ent_chunk[e]];
        const uint32_t par = (uint32_t)((g / STAGES) & 1);
        tr_on = (a.debug & 4) && blockIdx.x == 0 && p == 0 && g / prog.n_ent == 2;
        NF_TRACE2(0, e * 4 + 0);
        mbar_wait(smem_u32(&s.w_empty[p]), par ^ 1u);
        NF_TRACE2(0, e * 4 + 1);
        mbar_expect_tx(smem_u32(&s.w_land[p]), ch.y);
        bulk_g2s(smem_u32(s.W[p]), a.packed + ch.x, ch.y, smem_u32(&s.w_land[p]));
        mbar_wait(smem_u32(&s.w_land[p]), par);                      // landed in THIS CTA ...
        NF_TRACE2(0, e * 4 + 2);
        mbar_arrive_cluster_relaxed(ready_leader);                   // ... tell the leader's MMA thread
      }
    }
  } else if (warp == 17) {
    // ================= MMA issuer (leader CTA only): one thread, tight nested loops =================
    // Per Linear and slot: wait a_ready, then k0 steps on X0[slot] and 16 steps on H[slot]; every 4th step starts a new ring
    // stage (wait w_ready) and every chunk end releases it (multicast commit).  All operands are uniform arithmetic on
    // kernel parameters and loop counters, so ptxas keeps them in uniform registers (~15 instructions per MMA).
    if (crank == 0 && lane == 0) {
      uint32_t stage = 0, phase = 0, a_par = 0;
      const uint32_t base4 = smem_u32(smem_raw) >> 4;
      const uint32_t w4 = base4 + (uint32_t)(offsetof(Tc2Smem, W) >> 4);
      const uint32_t bar_wready = smem_u32(&s.w_ready[0]), bar_wempty = smem_u32(&s.w_empty[0]);
      const uint32_t bar_a = smem_u32(&s.a_ready[0]), bar_acc = smem_u32(&s.acc_full[0]);
      const uint32_t a_lbo = (uint32_t)(KG_BYTES >> 4) << 16, kstep4 = (uint32_t)(2 * KG_BYTES) >> 4;
      const bool no_w = (a.debug & 64) != 0, no_mma = (a.debug & 2) != 0, alone = (a.debug & 128) != 0;
      for (long long pass = 0; pass < passes; ++pass)
        for (int li = 0; li < prog.n_lin; ++li) {
          const uint4 r0 = *reinterpret_cast<const uint4*>(&prog.lin[li].k0_steps);
          const uint32_t bhi = prog.lin[li].bhi;
          const uint32_t k0s = r0.x, total = r0.x + r0.y, idesc = r0.z, bstep4 = r0.w;
#pragma unroll 1
          for (uint32_t slot = 0; slot < 2; ++slot) {
            if (!alone) { mbar_wait(bar_a + slot * 8u, (a_par >> slot) & 1u); a_par ^= 1u << slot; }
            tc_fence_after();
            const uint32_t d_tmem = slot * 256u;
            const uint32_t x4 = (base4 + (uint32_t)(offsetof(Tc2Smem, X0) >> 4) + slot * (uint32_t)(sizeof(s.X0[0]) >> 4)) | a_lbo;
            const uint32_t h4 = (base4 + (uint32_t)(offsetof(Tc2Smem, H) >> 4) + slot * (uint32_t)(sizeof(s.H[0]) >> 4)) | a_lbo;
            const uint32_t n_chunks = (total + 3u) >> 2;
#pragma unroll 1
            for (uint32_t c = 0; c < n_chunks; ++c) {
              const uint32_t gs0 = c << 2;
              if (!no_w) { mbar_wait(bar_wready + stage * 8u, phase); tc_fence_after(); }
              const uint32_t b4 = (w4 + stage * (uint32_t)(STAGE_BYTES >> 4)) | bhi;
              if (gs0 + 4u <= total && (gs0 + 4u <= k0s || gs0 >= k0s)) {
                // fast path: a full chunk fed from one buffer -> four back-to-back MMAs, operands differ by constants
                const uint32_t a4 = gs0 < k0s ? x4 + gs0 * kstep4 : h4 + (gs0 - k0s) * kstep4;
                if (!no_mma) {
                  umma2_f16(d_tmem, umma_desc_lo(a4), umma_desc_lo(b4), idesc, gs0 > 0 ? 1u : 0u);
                  umma2_f16(d_tmem, umma_desc_lo(a4 + kstep4), umma_desc_lo(b4 + bstep4), idesc, 1u);
                  umma2_f16(d_tmem, umma_desc_lo(a4 + 2u * kstep4), umma_desc_lo(b4 + 2u * bstep4), idesc, 1u);
                  umma2_f16(d_tmem, umma_desc_lo(a4 + 3u * kstep4), umma_desc_lo(b4 + 3u * bstep4), idesc, 1u);
                }
              } else {
                const uint32_t nst = total - gs0 < 4u ? total - gs0 : 4u;
                for (uint32_t i = 0; i < nst; ++i) {
                  const uint32_t gs = gs0 + i;
                  const uint32_t a4 = gs < k0s ? x4 + gs * kstep4 : h4 + (gs - k0s) * kstep4;
                  if (!no_mma) umma2_f16(d_tmem, umma_desc_lo(a4), umma_desc_lo(b4 + i * bstep4), idesc, gs > 0 ? 1u : 0u);
                }
              }
              if (!no_w) umma2_commit_mc(bar_wempty + stage * 8u);
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            if (!alone) umma2_commit_mc(bar_acc + slot * 8u);
          }
        }
    }
  } else {
    // ================= encode + epilogue: ALL 16 warps work on whichever slot is in its epilogue =================
    // TMEM lane quarter q = warp % 4 (rows 32q..32q+31), column quarter cq = warp / 4 (64 accumulator columns).
    const int q = warp & 3, cq = warp >> 2;
    const int e_tid = warp * 32 + lane;
    const int row = q * 32 + lane;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    uint32_t acc_par = 0;
    for (long long pass = 0; pass < ((a.debug & 128) ? 0 : passes); ++pass) {
      const long long trip = pass / map.tpr; const int sub = (int)(pass - trip * map.tpr);
      // ---------- stage the density MLP's raw x0 for both slots ----------
      if (!(a.debug & 1)) {
        const int r = e_tid & (ROWS - 1), part = e_tid >> 7;            // 4 threads per row, hash levels interleaved
#pragma unroll 1
        for (int slot = 0; slot < 2; ++slot) {
          uint8_t* X0 = s.X0[slot];
          const long long u = (trip * gridDim.x + blockIdx.x) * 2 + slot;
          long long ray; int t;
          const bool ok = map.locate(u, sub, r, a.n_rays, ray, t);
          float px = 0.f, py = 0.f, pz = 0.f;
          if (ok) {
            const float* rr = a.rays + ray * 6;
            const float tt = __ldg(a.ts + ray * a.ts_stride + t);
            px = nf_pt(__ldg(rr + 0), tt, __ldg(rr + 3)); py = nf_pt(__ldg(rr + 1), tt, __ldg(rr + 4)); pz = nf_pt(__ldg(rr + 2), tt, __ldg(rr + 5));
          }
          int kg = 0;
          if (plan.enc == NF_ENC_HASH) {
            const float4* tables = reinterpret_cast<const float4*>(a.packed + plan.hash_off);
            for (int lvl = part; lvl < plan.hash_levels; lvl += 4) {
              const float4 f = nf_hash_level(tables + (size_t)lvl * (plan.hash_mask + 1), px, py, pz, plan.hash_res[lvl],
                                             plan.hash_primes[0], plan.hash_primes[1], plan.hash_primes[2], plan.hash_mask, nullptr);
              *reinterpret_cast<uint2*>(X0 + (lvl >> 1) * KG_BYTES + r * 16 + (lvl & 1) * 8) = make_uint2(pack_h2(f.x, f.y), pack_h2(f.z, f.w));
            }
            kg = plan.hash_levels >> 1;
          }
          if (part == 0) {
            if (plan.enc == NF_ENC_HASH) st_v4(X0 + kg * KG_BYTES + r * 16, pack_h2(px, py), pack_h2(pz, px), pack_h2(py, pz), 0);
            else st_v4(X0 + kg * KG_BYTES + r * 16, pack_h2(px, py), pack_h2(pz, 0.f), 0, 0);
            for (int g = kg + 1; g < (plan.mlp[0].k0_pad >> 3); ++g) st_v4(X0 + g * KG_BYTES + r * 16, 0, 0, 0, 0);
            s.ray[slot][r] = ray; s.t[slot][r] = t; s.valid[slot][r] = ok ? 1 : 0;
          }
        }
      }
      fence_proxy_async();
      named_bar(1, EPI_THREADS);
      if (lane == 0) { mbar_arrive_cluster_relaxed(leader_addr(smem_u32(&s.a_ready[0]))); mbar_arrive_cluster_relaxed(leader_addr(smem_u32(&s.a_ready[1]))); }

      // ---------- the MLPs: slot 0's epilogue while the tensor pipe runs slot 1's Linear, and vice versa ----------
      for (int m = 0; m < plan.n_mlps; ++m) {
        const NfMlpPlan& M = plan.mlp[m];
        const int act = M.act;
        for (int j = 0; j < M.n_lin; ++j) {
          const NfLinPlan& L = M.lin[j];
          const float* bias = reinterpret_cast<const float*>(a.packed + L.b16_off);
#pragma unroll 1
          for (int slot = 0; slot < 2; ++slot) {
            uint8_t* H = s.H[slot]; uint8_t* X0 = s.X0[slot];
            const uint32_t t_acc = t_lane + (uint32_t)slot * 256u;
            const uint32_t a_ready_leader = leader_addr(smem_u32(&s.a_ready[slot]));
            const Bias16 b0 = bias_prefetch(bias, L.is_out ? 0 : cq * 64);    // in flight across the acc_full wait
            tr_on = (a.debug & 4) && blockIdx.x == 0 && pass == 2 && warp == 0 && lane == 0;
            NF_TRACE2(2 + slot, (m * 16 + j) * 4 + 0);
            if (a.debug & 256) mbar_wait_backoff(smem_u32(&s.acc_full[slot]), (acc_par >> slot) & 1u);
            else mbar_wait_suspend(smem_u32(&s.acc_full[slot]), (acc_par >> slot) & 1u);
            acc_par ^= 1u << slot;
            NF_TRACE2(2 + slot, (m * 16 + j) * 4 + 1);
            tc_fence_after();
            if (a.debug & 1) {
              if (!(m == plan.n_mlps - 1 && L.is_out)) { __syncwarp(); if (lane == 0) mbar_arrive_cluster_relaxed(a_ready_leader); }
            } else if (!L.is_out) {
              if (j == 0) x0_activate(X0, M.k0_pad, act, e_tid);       // init consumed raw x0; the skip Linear wants act(x0)
              if (act == NF_ACT_SIN) epi_hidden2<NF_ACT_SIN>(H, t_acc, bias, cq, row, b0);
              else if (act == NF_ACT_LEAKY) epi_hidden2<NF_ACT_LEAKY>(H, t_acc, bias, cq, row, b0);
              else if (act == NF_ACT_RELU) epi_hidden2<NF_ACT_RELU>(H, t_acc, bias, cq, row, b0);
              else epi_hidden2<NF_ACT_NONE>(H, t_acc, bias, cq, row, b0);
              NF_TRACE2(2 + slot, (m * 16 + j) * 4 + 2);
              tc_fence_before();
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster_relaxed(a_ready_leader);
              NF_TRACE2(2 + slot, (m * 16 + j) * 4 + 3);
            } else if (plan.kind == NF_KIND_PLAIN && m == 0) {
              // density MLP out (tensor order [inter(I), sigma]) -> raw x0 of the View head + raw density
              const int iu = plan.intermediate >> 4;
              for (int un = cq; un <= iu; un += 4) {
                uint32_t v[16];
                tmem_ld16(t_acc + un * 16, v); tmem_ld_wait(); reg_fence16(v);
                if (un < iu) {
                  uint32_t o[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    o[i] = pack_h2(__uint_as_float(v[2 * i]) + __ldg(bias + un * 16 + 2 * i), __uint_as_float(v[2 * i + 1]) + __ldg(bias + un * 16 + 2 * i + 1));
                  uint8_t* d0 = X0 + (un * 2) * KG_BYTES + row * 16;
                  st_v4(d0, o[0], o[1], o[2], o[3]); st_v4(d0 + KG_BYTES, o[4], o[5], o[6], o[7]);
                } else {
                  s.sig[slot][row] = __uint_as_float(v[0]) + __ldg(bias + plan.intermediate);
                  float px = 0.f, py = 0.f, pz = 0.f, el = 0.f, az = 0.f;
                  if (s.valid[slot][row]) {
                    const long long ray = s.ray[slot][row];
                    const float* rr = a.rays + ray * 6;
                    const float tt = __ldg(a.ts + ray * a.ts_stride + s.t[slot][row]);
                    const float dx = __ldg(rr + 3), dy = __ldg(rr + 4), dz = __ldg(rr + 5);
                    px = nf_pt(__ldg(rr + 0), tt, dx); py = nf_pt(__ldg(rr + 1), tt, dy); pz = nf_pt(__ldg(rr + 2), tt, dz);
                    nf_elaz(dx, dy, dz, el, az);
                  }
                  uint8_t* d0 = X0 + (iu * 2) * KG_BYTES + row * 16;
                  st_v4(d0, pack_h2(px, py), pack_h2(pz, el), pack_h2(az, 0.f), 0);
                  st_v4(d0 + KG_BYTES, 0, 0, 0, 0);
                }
              }
              tc_fence_before();
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster_relaxed(a_ready_leader);
            } else {
              // final Linear of the path -> colours (and raw density for TinyNeRF) -> composite by the cq == 0 warps
              if (cq == 0) {
                uint32_t v[16];
                tmem_ld16(t_acc, v); tmem_ld_wait(); reg_fence16(v);
                tc_fence_before();
                float cr, cg, cb;
                if (plan.kind == NF_KIND_TINY) {
                  s.sig[slot][row] = __uint_as_float(v[0]) + __ldg(bias);
                  cr = __uint_as_float(v[1]) + __ldg(bias + 1); cg = __uint_as_float(v[2]) + __ldg(bias + 2); cb = __uint_as_float(v[3]) + __ldg(bias + 3);
                } else {
                  cr = __uint_as_float(v[0]) + __ldg(bias); cg = __uint_as_float(v[1]) + __ldg(bias + 1); cb = __uint_as_float(v[2]) + __ldg(bias + 2);
                }
                nf_feat_act3(cr, cg, cb, plan.feat_act);
                composite_tile2(s, slot, plan, a, map, sub, row, lane, q, cr, cg, cb);
              }
            }
          }
        }
      }
      named_bar(1, EPI_THREADS);   // both slots' tile-private smem is free for the next tiles
    }
    (void)lin_base1;
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 18) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(0), "r"(512) : "memory");
  }
}

// Per-round program: one record per Linear (issuer) and the ring-entry -> weight-chunk map (producers): for every Linear
// its chunks for slot 0 and then the same chunks again for slot 1.
void build_prog2(const NfPlan& plan, Tc2Prog* P) {
  *P = Tc2Prog{};
  int nl = 0, ne = 0, nc_base = 0;
  for (int m = 0; m < plan.n_mlps; ++m)
    for (int j = 0; j < plan.mlp[m].n_lin; ++j, ++nl) {
      const NfLinPlan& L = plan.mlp[m].lin[j];
      const uint32_t nh = (uint32_t)L.n_pad >> 1, b_lbo = nh * 16u;
      Tc2Lin& R = P->lin[nl];
      R.k0_steps = (uint32_t)L.k0_pad >> 4; R.h_steps = (uint32_t)L.k_hidden >> 4;
      R.idesc = (1u << 4) | ((uint32_t)(L.n_pad >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // M = 256 across the pair
      R.bstep4 = (2u * b_lbo) >> 4; R.bhi = (b_lbo >> 4) << 16;
      for (int slot = 0; slot < 2; ++slot)
        for (int c = 0; c < L.n_chunks; ++c) P->ent_chunk[ne++] = (uint8_t)(nc_base + c);
      nc_base += L.n_chunks;
    }
  P->n_lin = nl; P->n_ent = ne;
}

}  // namespace

// nullptr if the paired pipeline can run this model, else the reason.
const char* nf_tc2_unsupported(const NfPlan& p) {
  if (p.kind == NF_KIND_DYN) return "NF_KIND_DYN runs on the fp32 pipeline only in this build";
  int chunks = 0, nlin = 0;
  for (int m = 0; m < p.n_mlps; ++m) {
    nlin += p.mlp[m].n_lin;
    if (p.mlp[m].k0_pad > X0K) return "x0 wider than 80 columns";
    for (int j = 0; j < p.mlp[m].n_lin; ++j) chunks += p.mlp[m].lin[j].n_chunks;
  }
  if (2 * chunks > MAX_ENT) return "too many weight chunks";
  if (nlin > MAX_LIN2) return "more than 16 Linear layers";
  if (p.enc == NF_ENC_FOURIER) return "Fourier-encoded density MLP (x0 is 259 wide) runs on the fp32 pipeline only";
  if (p.enc == NF_ENC_HASH && (p.hash_levels & 1)) return "odd number of hash levels";
  if (p.kind == NF_KIND_PLAIN && (p.intermediate & 15)) return "intermediate_size not a multiple of 16";
  return nullptr;
}

cudaError_t nf_launch_render_tc2(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, const float* ts,
                                 int T, int64_t ts_stride, const float* noise, float* rgb, float* alpha, float* weights,
                                 cudaStream_t st) {
  if (nf_tc2_unsupported(plan)) return cudaErrorNotSupported;
  Tc2Args a{};
  a.packed = (const uint8_t*)packed; a.rays = rays; a.n_rays = n_rays; a.ts = ts; a.T = T; a.ts_stride = ts_stride;
  a.noise = noise; a.rgb_out = rgb; a.alpha_out = alpha; a.weights_out = weights;
#ifdef NF_EXPERIMENTS
  if (const char* dbg = getenv("NF_TC_DEBUG")) a.debug = atoi(dbg);
#endif
  cudaError_t e = cudaFuncSetAttribute(k_render_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tc2Smem));
  if (e != cudaSuccess) return e;
  const NfTileMap map(T, ROWS);
  const long long units = map.units(n_rays);
  if (units == 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long want = (units + 3) / 4 * 2;                      // 2 CTAs x 2 slots per cluster
  const int grid = (int)(want < (sms / 2) * 2 ? want : (sms / 2) * 2);
  Tc2Prog prog;
  build_prog2(plan, &prog);
#ifdef NF_TC_TRACE
  if (a.debug & 4) {
    static long long* d_trace = nullptr;
    if (!d_trace) cudaMalloc(&d_trace, 4 * 512 * 2 * sizeof(long long));
    cudaMemset(d_trace, 0, 4 * 512 * 2 * sizeof(long long));
    a.trace = d_trace;
    k_render_tc2<<<grid, THREADS, sizeof(Tc2Smem), st>>>(plan, prog, a);
    cudaStreamSynchronize(st);
    static long long h[4 * 512 * 2];
    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
    long long t0 = -1;
    for (int i = 0; i < 4 * 512; ++i) if (h[2 * i + 1] && (t0 < 0 || h[2 * i + 1] < t0)) t0 = h[2 * i + 1];
    static int printed = 0;
    if (t0 >= 0 && printed++ < 1)
      for (int r = 0; r < 4; ++r) for (int i = 0; i < 512; ++i) if (h[2 * (r * 512 + i) + 1])
        printf("TRACE role=%d tag=%lld t=%lld\n", r, h[2 * (r * 512 + i)], h[2 * (r * 512 + i) + 1] - t0);
    return cudaGetLastError();
  }
#endif
  k_render_tc2<<<grid, THREADS, sizeof(Tc2Smem), st>>>(plan, prog, a);
  return cudaGetLastError();
}
