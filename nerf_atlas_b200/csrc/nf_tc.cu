// nf_tc.cu -- the Blackwell (sm_100a) tensor-core form of the fused render pipeline.
//
// One persistent CTA per SM owns tiles of 128 consecutive samples (UMMA M = 128).  Per tile:
//   sample positions -> hash-grid encode -> SkipConnMLP (density) -> SkipConnMLP (View) ->
//   alpha composite, with nothing but rays in and RGB (+ optional alpha/weights) out.
// Every Linear is a tcgen05.mma (kind::f16, fp16 operands, fp32 accumulators in TMEM):
//   A = activations [128 x K] in shared memory, UMMA-canonical K-major no-swizzle layout
//       [K/8][128 rows][8 halves] (16-byte cores; written conflict-free by the epilogue warps),
//   B = weights [N x K], pre-packed on the device by nf_pack_weights into the same canonical
//       layout and streamed from L2 by the bulk-copy engine (cp.async.bulk = TMA 1-D) through
//       a 3-stage mbarrier ring of 32 KB chunks (the bulk-copy engine has a ~430-cycle fixed cost per copy and
//       ~2 copies in flight per SM, so bandwidth scales with copy size: profiles/microbench/copy_bw.cu),
//   D = [128 x N] fp32 in TMEM, two 256-column accumulators used alternately by consecutive
//       layers so that layer j+1's MMAs start on the K-chunks of layer j's output as soon as the
//       epilogue has written them (chunk-granular h_ready barriers).
// Warp roles: 0..7 = encode/epilogue warps (TMEM lane quarter = warp % 4, column half = warp / 4); 8, 10, 11 = weight
// producers (one ring stage each; 10 also allocates TMEM), 9 = MMA issuer (one thread).  The role warps carry the highest
// warp ids because the SM schedulers favour higher ids: the lone issuing thread must not queue behind the epilogue warps.
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include "nf_common.cuh"
#include "nf_kernels.h"
#include "nf_tc_ptx.cuh"

namespace {
using namespace nf_ptx;

constexpr int X0K = 80;                   // max padded x0 width on this path
constexpr int STAGES = 3;
constexpr int SPC = NF_TC_CHUNK_K / 16;                 // UMMA K-steps per weight chunk
constexpr int STAGE_BYTES = NF_TC_CHUNK_K * 256 * 2;   // 32 KB: 64 K-columns x 256 N x fp16
constexpr int MAX_LIN_TOTAL = 12;
constexpr int THREADS = 384;
constexpr int EPI_THREADS = 256;
constexpr int MAX_CHUNKS = 56;
// flags of one weight chunk of the per-tile MMA program: bits 0-2 = UMMA steps in the chunk (1..4)
constexpr uint32_t F_NSTEP = 7, F_FIRST = 8, F_LAST = 16, F_BUF = 32, F_WAIT_X0 = 64, F_WAIT_H = 0xF00;

struct TcSmem {
  uint8_t H[ROWS * 256 * 2];
  uint8_t X0raw[ROWS * X0K * 2];
  uint8_t X0act[ROWS * X0K * 2];
  uint8_t W[STAGES][STAGE_BYTES];
  float bias[MAX_LIN_TOTAL * 256];
  float P[3 * ROWS];
  float sig[ROWS], delta[ROWS], el[ROWS], az[ROWS];
  float wrgb[ROWS * 4];
  long long ray[ROWS];
  int t[ROWS];
  int valid[ROWS];
  float warp_agg[4]; int warp_cont[4]; float warp_sum[4][4]; float carry[8];
  unsigned long long w_full[STAGES], w_empty[STAGES], acc_full[2], h_ready[4], x0_ready;
  uint32_t tmem_base;
  int n_chunks, pad_;
  uint2 chunks[MAX_CHUNKS];     // per-tile weight chunk list: {byte offset into packed, bytes}
};
static_assert(sizeof(TcSmem) <= 227 * 1024, "tensor pipeline smem");

// Per-tile MMA program, built on the host and passed as a kernel parameter so that the issuing thread reads it
// through uniform constant loads (LDCU) and every descriptor lives in uniform registers.
struct __align__(16) TcProgEnt { uint32_t a4[4]; uint32_t bstep4, idesc, flags, bhi; };   // a4 = (A smem offset >> 4) | LBO field
struct __align__(16) TcProg { int32_t n_chunks, odd_lin, pad0_, pad1_; TcProgEnt e[MAX_CHUNKS]; };

struct TcArgs {
  const uint8_t* packed;
  const float* rays; long long n_rays;
  const float* ts; int T; long long ts_stride;
  const float* noise;
  float* rgb_out; float* alpha_out; float* weights_out;
  // MLP-only mode
  int mlp_only; int which; const float* x0; long long n; float* out;
  const long long* n_dev;   // MLP-only mode: if set, the number of rows is read from device memory (stream-ordered loops such as the sphere march: no host round trip)
  long long* trace;   // debug & 4: clock64 timeline of one tile of block 0: [role][512] x {tag, clock}
  int debug;   // timing experiments only (NF_TC_DEBUG): 1 = epilogue does no work, 2 = no MMA issued
};

#ifdef NF_TC_TRACE   // compile-time only: the trace stores push the MMA issuer's descriptors out of uniform registers
#define NF_TRACE(role, tag) do { if ((NF_TC_TRACE != 2 || (role) == 2) && tr_on && tr_n[role] < 512) { tr[(role * 512 + tr_n[role]) * 2] = (tag); tr[(role * 512 + tr_n[role]) * 2 + 1] = clock64(); ++tr_n[role]; } } while (0)
#else
#define NF_TRACE(role, tag) do { (void)tr_on; } while (0)
#endif

// ---- the sequence of tiles a CTA walks (identical in every warp role) ---------------------------
struct TileIter {
  long long units, trips; int tpr;   // every CTA makes `trips` passes; passes with u >= units are all-masked tiles
  __device__ TileIter(const TcArgs& a, const NfTileMap& map) {
    if (a.mlp_only) { units = ((a.n_dev ? *a.n_dev : a.n) + ROWS - 1) / ROWS; tpr = 1; } else { units = map.units(a.n_rays); tpr = map.tpr; }
    trips = (units + gridDim.x - 1) / gridDim.x;
  }
};

// ---- epilogue of a hidden Linear: H <- fp16(act(acc + bias)), in place, 64-column chunks ----------
// TMEM loads of chunk c+1 are in flight while chunk c is converted and stored.
template <int ACT>
__device__ __forceinline__ void epi_hidden(TcSmem& s, uint32_t t_acc, const float* __restrict__ bias, int half, int row, int lane,
                                           int trace_tag = -1, bool tr_on_ = false, long long* tr_ = nullptr, int* tr_n_ = nullptr) {
  uint32_t v[2][32];
  tmem_ld16(t_acc + half * 32, v[0]); tmem_ld16(t_acc + half * 32 + 16, v[0] + 16);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_ld_wait();
    reg_fence16(v[c & 1]); reg_fence16(v[c & 1] + 16);
    if (c < 3) { tmem_ld16(t_acc + (c + 1) * 64 + half * 32, v[(c + 1) & 1]); tmem_ld16(t_acc + (c + 1) * 64 + half * 32 + 16, v[(c + 1) & 1] + 16); }
    const int col = c * 64 + half * 32;
    const float4* b4 = reinterpret_cast<const float4*>(bias + col);
    uint32_t o[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = b4[i];
      o[2 * i]     = pack_h2(tc_act_t<ACT>(__uint_as_float(v[c & 1][4 * i]) + b.x), tc_act_t<ACT>(__uint_as_float(v[c & 1][4 * i + 1]) + b.y));
      o[2 * i + 1] = pack_h2(tc_act_t<ACT>(__uint_as_float(v[c & 1][4 * i + 2]) + b.z), tc_act_t<ACT>(__uint_as_float(v[c & 1][4 * i + 3]) + b.w));
    }
    uint8_t* dst = s.H + (col >> 3) * KG_BYTES + row * 16;
#pragma unroll
    for (int g = 0; g < 4; ++g) st_v4(dst + g * KG_BYTES, o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&s.h_ready[c]));
    if (trace_tag >= 0 && tr_on_ && lane == 0 && *tr_n_ < 512) { tr_[(2 * 512 + *tr_n_) * 2] = 1000 + trace_tag * 4 + c; tr_[(2 * 512 + *tr_n_) * 2 + 1] = clock64(); ++*tr_n_; }
  }
}

// ---- composite of one tile by the 4 column-half-0 epilogue warps (thread = row) ----------------
// reference src/nerf.py:60-80. Running transmittance = segmented warp-shuffle product scan,
// stitched across the 4 warps (and across sub-tiles of a ray longer than 128 samples) in smem.
__device__ __forceinline__ void composite_tile(TcSmem& s, const NfPlan& plan, const TcArgs& a, const NfTileMap& map,
                                               int sub, int row, int lane, int q, float cr, float cg, float cb) {
  const bool valid = s.valid[row] != 0;
  const int t = valid ? s.t[row] : 0;
  const long long ray = s.ray[row];
  float al = 0.f;
  if (valid) {
    float sr = s.sig[row];
    if (a.noise) sr += __ldg(a.noise + ray * a.T + t);
    const float beta = plan.density_act == NF_DENS_LAPLACE ? __ldg(reinterpret_cast<const float*>(a.packed + plan.scale_off)) : 1.f;
    al = nf_alpha(sr, s.delta[row], plan.density_act, beta);
  }
  float incl = valid ? (1.f - al) + 1e-10f : 1.f;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d && t >= d) incl *= o;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0 || t == 0) excl = 1.f;
  if (lane == 31) { s.warp_agg[q] = incl; s.warp_cont[q] = t > 31; }
  named_bar(2, 128);
  float c = 1.f; bool open = true;
  for (int v = q - 1; v >= 0 && open; --v) { c *= s.warp_agg[v]; open = s.warp_cont[v] != 0; }
  if (open && sub > 0) c *= s.carry[0];
  const float trans = excl * (t > lane ? c : 1.f);
  const float w = al * trans;
  if (valid) {
    if (a.alpha_out) a.alpha_out[ray * a.T + t] = al;
    if (a.weights_out) a.weights_out[ray * a.T + t] = w;
  }
  const float wr = w * cr, wg = w * cg, wb = w * cb, wl = (valid && t < a.T - 1) ? w : 0.f;
  const int row_thread = q * 32 + lane;   // 0..127
  if ((a.T & 31) == 0) {
    float x0 = wr, x1 = wg, x2 = wb, x3 = wl;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      x0 += __shfl_xor_sync(0xffffffffu, x0, d); x1 += __shfl_xor_sync(0xffffffffu, x1, d);
      x2 += __shfl_xor_sync(0xffffffffu, x2, d); x3 += __shfl_xor_sync(0xffffffffu, x3, d);
    }
    if (lane == 0) { s.warp_sum[q][0] = x0; s.warp_sum[q][1] = x1; s.warp_sum[q][2] = x2; s.warp_sum[q][3] = x3; }
    named_bar(2, 128);
    const int wpr = a.T <= ROWS ? a.T / 32 : 4;          // warps per ray (segment) inside this tile
    const int nseg = a.T <= ROWS ? map.rpt : 1;
    if (row_thread < nseg && s.valid[row_thread * (a.T <= ROWS ? a.T : 0)]) {
      float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
      if (sub > 0) { o0 = s.carry[1]; o1 = s.carry[2]; o2 = s.carry[3]; o3 = s.carry[4]; }
      for (int v = 0; v < wpr; ++v) {
        const float* ws = s.warp_sum[row_thread * wpr + v];
        o0 += ws[0]; o1 += ws[1]; o2 += ws[2]; o3 += ws[3];
      }
      const long long r = s.ray[row_thread * (a.T <= ROWS ? a.T : 0)];
      if (sub == map.tpr - 1) {
        const float skyv = plan.bg == NF_BG_WHITE ? 1.f - o3 : 0.f;
        a.rgb_out[r * 3 + 0] = o0 + skyv; a.rgb_out[r * 3 + 1] = o1 + skyv; a.rgb_out[r * 3 + 2] = o2 + skyv;
      } else {
        s.carry[1] = o0; s.carry[2] = o1; s.carry[3] = o2; s.carry[4] = o3;
        s.carry[0] = (sub > 0 ? s.carry[0] : 1.f) * s.warp_agg[0] * s.warp_agg[1] * s.warp_agg[2] * s.warp_agg[3];
      }
    }
  } else {
    float* wq = s.wrgb + row_thread * 4;
    wq[0] = wr; wq[1] = wg; wq[2] = wb; wq[3] = wl;
    named_bar(2, 128);
    const int nseg = a.T <= ROWS ? map.rpt : 1;
    const int row0 = a.T <= ROWS ? row_thread * a.T : 0;
    if (row_thread < nseg && s.valid[row0]) {
      float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
      if (sub > 0) { o0 = s.carry[1]; o1 = s.carry[2]; o2 = s.carry[3]; o3 = s.carry[4]; }
      const int nrow = a.T <= ROWS ? a.T : min(ROWS, a.T - sub * ROWS);
      for (int i = 0; i < nrow; ++i) { const float* x = s.wrgb + (row0 + i) * 4; o0 += x[0]; o1 += x[1]; o2 += x[2]; o3 += x[3]; }
      const long long r = s.ray[row0];
      if (sub == map.tpr - 1) {
        const float skyv = plan.bg == NF_BG_WHITE ? 1.f - o3 : 0.f;
        a.rgb_out[r * 3 + 0] = o0 + skyv; a.rgb_out[r * 3 + 1] = o1 + skyv; a.rgb_out[r * 3 + 2] = o2 + skyv;
      } else {
        s.carry[1] = o0; s.carry[2] = o1; s.carry[3] = o2; s.carry[4] = o3;
        s.carry[0] = (sub > 0 ? s.carry[0] : 1.f) * s.warp_agg[0] * s.warp_agg[1] * s.warp_agg[2] * s.warp_agg[3];
      }
    }
  }
}

// =================================================================================================
__global__ void __launch_bounds__(THREADS, 1)
k_render_tc(const __grid_constant__ NfPlan plan, const __grid_constant__ TcProg prog, const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TcSmem& s = *reinterpret_cast<TcSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const NfTileMap map(a.mlp_only ? ROWS : a.T, ROWS);
  const TileIter it(a, map);
  const int m_begin = a.mlp_only ? a.which : 0, m_end = a.mlp_only ? a.which + 1 : plan.n_mlps;
  const int lin_base1 = a.mlp_only ? 0 : plan.mlp[0].n_lin;   // bias slot of the refl MLP's first Linear (mlp-only runs: the one MLP at 0)

  long long* tr = a.trace; int tr_n[3] = {0, 0, 0}; bool tr_on = false; (void)tr; (void)tr_n;
  // ---- one-time setup ----
  for (int i = threadIdx.x; i < MAX_LIN_TOTAL * 256; i += THREADS) {
    const int g = i >> 8, n = i & 255;
    const int m = a.mlp_only ? a.which : (g >= lin_base1 ? 1 : 0), j = g - ((!a.mlp_only && m) ? lin_base1 : 0);
    float v = 0.f;
    if (m < plan.n_mlps && j < plan.mlp[m].n_lin && n < plan.mlp[m].lin[j].n_pad)
      v = __ldg(reinterpret_cast<const float*>(a.packed + plan.mlp[m].lin[j].b16_off) + n);
    s.bias[i] = v;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(smem_u32(&s.w_full[i]), 1); mbar_init(smem_u32(&s.w_empty[i]), 1); }
    mbar_init(smem_u32(&s.acc_full[0]), 1); mbar_init(smem_u32(&s.acc_full[1]), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&s.h_ready[i]), 8);
    mbar_init(smem_u32(&s.x0_ready), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 10) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 11 && lane == 0) {
    // weight chunk list for the producers (identical for every tile): {byte offset into packed, bytes}
    int nc = 0;
    for (int m = m_begin; m < m_end; ++m)
      for (int j = 0; j < plan.mlp[m].n_lin; ++j) {
        const NfLinPlan& L = plan.mlp[m].lin[j];
        const int steps = (L.k0_pad + L.k_hidden) >> 4;
        for (int c = 0; c < L.n_chunks; ++c, ++nc)
          s.chunks[nc] = make_uint2((uint32_t)(L.w16_off + (int64_t)c * (2 * SPC) * L.n_pad * 16),
                                    (uint32_t)min(SPC, steps - SPC * c) * 2u * (uint32_t)L.n_pad * 16u);
      }
    s.n_chunks = nc;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;
  if (tmem_base != 0) { if (threadIdx.x == 0) printf("nf_tc: unexpected TMEM base %u\n", tmem_base); __trap(); }

  if (warp == 8 || warp == 10 || warp == 11) {
    // ================= weight producers: three single-thread issuers, one per ring stage =================
    // cp.async.bulk costs the ISSUING warp ~450 cycles per copy (about 2 copies in flight per warp, any size;
    // profiles/microbench/copy_bw.cu), so one issuer tops out at ~75 B/cycle with 32 KB chunks; issuers scale.
    if (lane == 0 && !(a.debug & 64)) {
      const int p = warp == 8 ? 0 : warp - 9;           // producer index == ring stage it owns
      const int nc = s.n_chunks;
      const long long total = it.trips * it.tpr * nc;
      int c = p % nc;
      for (long long g = p; g < total; g += STAGES) {
        const uint2 ch = s.chunks[c];
        tr_on = (a.debug & 4) && blockIdx.x == 0 && p == 0 && g / nc == 2;
        mbar_wait(smem_u32(&s.w_empty[p]), (uint32_t)(((g / STAGES) & 1) ^ 1));
        NF_TRACE(0, c);
        mbar_expect_tx(smem_u32(&s.w_full[p]), ch.y);
        bulk_g2s(smem_u32(s.W[p]), a.packed + ch.x, ch.y, smem_u32(&s.w_full[p]));
        c += STAGES; while (c >= nc) c -= nc;
      }
    }
  } else if (warp == 9) {
    // ================= MMA issuer: ONE thread walks the chunk program and issues every tcgen05.mma =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t h_par = 0, x0_par = 0, tile_par = 0;   // expected parities of h_ready[0..3], x0_ready; accumulator flip
      const uint32_t base4 = smem_u32(smem_raw) >> 4;
      const uint32_t w4 = base4 + (uint32_t)(offsetof(TcSmem, W) >> 4);
      const uint32_t bar_wfull = smem_u32(&s.w_full[0]), bar_wempty = smem_u32(&s.w_empty[0]);
      const uint32_t bar_h = smem_u32(&s.h_ready[0]), bar_x0 = smem_u32(&s.x0_ready), bar_acc = smem_u32(&s.acc_full[0]);
      for (long long trip = 0; trip < it.trips * it.tpr; ++trip) {
        tr_on = (a.debug & 4) && blockIdx.x == 0 && trip == 2;
        for (int c = 0; c < prog.n_chunks; ++c) {
          // the whole entry with two 128-bit uniform loads, before anything depends on it
          const uint4 ea = *reinterpret_cast<const uint4*>(prog.e[c].a4);
          const uint4 eb = *reinterpret_cast<const uint4*>(&prog.e[c].bstep4);
          const uint32_t f = eb.z;
          NF_TRACE(1, c * 4 + 0);
          // One polling loop for everything this chunk needs (weights landed, x0 staged, H chunk(s) written): the
          // test_wait latencies (~150 cycles each) overlap instead of adding up.
          {
            const uint32_t hm = (f >> 8) & 15u;
            const uint32_t hc0 = hm ? (uint32_t)__ffs((int)hm) - 1u : 0u;          // first needed H chunk
            const uint32_t hm2 = hm & (hm - 1u);                                    // a second one (misaligned skip layers)
            const uint32_t hc1 = hm2 ? (uint32_t)__ffs((int)hm2) - 1u : hc0;
            uint32_t spins = 0;
            while (true) {
              bool ok = (a.debug & 64) ? true : mbar_test_wait(bar_wfull + stage * 8u, phase);
              if (hm) ok &= mbar_test_wait(bar_h + hc0 * 8u, (h_par >> hc0) & 1u);
              if (hm2) ok &= mbar_test_wait(bar_h + hc1 * 8u, (h_par >> hc1) & 1u);
              if (f & F_WAIT_X0) ok &= mbar_test_wait(bar_x0, x0_par);
              if (ok) break;
              if (++spins > (1u << 26)) __trap();
            }
            h_par ^= hm;
            if (f & F_WAIT_X0) x0_par ^= 1u;
          }
          if (!(a.debug & 8)) tc_fence_after();
          NF_TRACE(1, c * 4 + 1);
          const uint32_t buf = ((f >> 5) & 1u) ^ tile_par;
          const uint32_t d_tmem = buf * 256u;                       // TMEM base is 0 (checked at setup)
          const uint32_t b0 = (w4 + stage * (STAGE_BYTES >> 4)) | eb.w;
          const uint32_t nst = f & F_NSTEP;
          // debug & 32 (timing experiment): issue half-width MMAs (N/2) -- results are wrong, execution time halves
          const uint32_t idesc = (a.debug & 32) ? ((eb.y & ~(0x3Fu << 17)) | ((((eb.y >> 17) & 0x3Fu) >> 1) << 17)) : eb.y;
          if (!(a.debug & 2)) {
            const uint32_t a0 = base4 + ea.x, a1 = base4 + ea.y, a2 = base4 + ea.z, a3 = base4 + ea.w;
            const uint32_t b1 = b0 + eb.x, b2 = b1 + eb.x, b3 = b2 + eb.x;
            umma_f16(d_tmem, umma_desc_lo(a0), umma_desc_lo(b0), idesc, (f & F_FIRST) ? 0u : 1u);
            if (nst > 1) umma_f16(d_tmem, umma_desc_lo(a1), umma_desc_lo(b1), idesc, 1u);
            if (nst > 2) umma_f16(d_tmem, umma_desc_lo(a2), umma_desc_lo(b2), idesc, 1u);
            if (nst > 3) umma_f16(d_tmem, umma_desc_lo(a3), umma_desc_lo(b3), idesc, 1u);
          }
          NF_TRACE(1, c * 4 + 2);
          if (!(a.debug & 64)) umma_commit(bar_wempty + stage * 8u);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          if (f & F_LAST) umma_commit(bar_acc + buf * 8u);
          NF_TRACE(1, c * 4 + 3);
        }
        tile_par ^= (uint32_t)prog.odd_lin;
      }
    }
  } else if (warp < 8) {
    // ================= encode + epilogue warps =================
    const int ew = warp, q = warp & 3, half = ew >> 2;
    const int e_tid = ew * 32 + lane;
    const int row = q * 32 + lane;                       // TMEM lane == tile row owned in epilogues
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_par = 0, lin_count = 0;
    for (long long trip = 0, u = blockIdx.x; trip < it.trips; ++trip, u += gridDim.x)
      for (int sub = 0; sub < it.tpr; ++sub) {
        // ---------- stage inputs of the first MLP ----------
        if (a.debug & 1) {
        } else if (a.mlp_only) {
          const NfMlpPlan& M = plan.mlp[a.which];
          for (int i = e_tid; i < (M.k0_pad >> 3) * ROWS; i += EPI_THREADS) {
            st_v4(s.X0raw + i * 16, 0, 0, 0, 0); st_v4(s.X0act + i * 16, 0, 0, 0, 0);
          }
          named_bar(1, EPI_THREADS);
          for (int i = e_tid; i < M.in_dims * ROWS; i += EPI_THREADS) {
            const int r = i / M.in_dims, k = i - r * M.in_dims;
            const long long g = u * ROWS + r;
            const float v = g < (a.n_dev ? *a.n_dev : a.n) ? __ldg(a.x0 + g * M.in_dims + k) : 0.f;
            const int kt = nf_x0_perm(plan, a.which, k);
            const int off = (kt >> 3) * KG_BYTES + r * 16 + (kt & 7) * 2;
            *reinterpret_cast<__half*>(s.X0raw + off) = __float2half_rn(v);
            *reinterpret_cast<__half*>(s.X0act + off) = __float2half_rn(tc_act(v, M.act));
          }
        } else {
          const int r = e_tid & (ROWS - 1), part = e_tid >> 7;
          long long ray; int t;
          const bool ok = map.locate(u, sub, r, a.n_rays, ray, t);
          float px = 0.f, py = 0.f, pz = 0.f, dx = 0.f, dy = 0.f, dz = 1.f;
          if (ok) {
            const float* rr = a.rays + ray * 6;
            const float tt = __ldg(a.ts + ray * a.ts_stride + t);
            dx = __ldg(rr + 3); dy = __ldg(rr + 4); dz = __ldg(rr + 5);
            px = nf_pt(__ldg(rr + 0), tt, dx); py = nf_pt(__ldg(rr + 1), tt, dy); pz = nf_pt(__ldg(rr + 2), tt, dz);
          }
          const int act0 = plan.mlp[0].act;
          int kg = 0;   // first K-group after the encoder features
          if (plan.enc == NF_ENC_HASH) {
            const float4* tables = reinterpret_cast<const float4*>(a.packed + plan.hash_off);
            for (int lvl = part; lvl < plan.hash_levels; lvl += 2) {
              const float4 f = nf_hash_level(tables + (size_t)lvl * (plan.hash_mask + 1), px, py, pz, plan.hash_res[lvl],
                                             plan.hash_primes[0], plan.hash_primes[1], plan.hash_primes[2], plan.hash_mask, nullptr);
              const int off = (lvl >> 1) * KG_BYTES + r * 16 + (lvl & 1) * 8;
              *reinterpret_cast<uint2*>(s.X0raw + off) = make_uint2(pack_h2(f.x, f.y), pack_h2(f.z, f.w));
              *reinterpret_cast<uint2*>(s.X0act + off) = make_uint2(pack_h2(tc_act(f.x, act0), tc_act(f.y, act0)),
                                                                    pack_h2(tc_act(f.z, act0), tc_act(f.w, act0)));
            }
            kg = plan.hash_levels >> 1;
          }
          if (part == 0) {
            const float ax = tc_act(px, act0), ay = tc_act(py, act0), az_ = tc_act(pz, act0);
            if (plan.enc == NF_ENC_HASH) {   // [p, p, 0, 0]
              st_v4(s.X0raw + kg * KG_BYTES + r * 16, pack_h2(px, py), pack_h2(pz, px), pack_h2(py, pz), 0);
              st_v4(s.X0act + kg * KG_BYTES + r * 16, pack_h2(ax, ay), pack_h2(az_, ax), pack_h2(ay, az_), 0);
            } else {                          // [p, 0...]
              st_v4(s.X0raw + kg * KG_BYTES + r * 16, pack_h2(px, py), pack_h2(pz, 0.f), 0, 0);
              st_v4(s.X0act + kg * KG_BYTES + r * 16, pack_h2(ax, ay), pack_h2(az_, 0.f), 0, 0);
            }
            for (int g = kg + 1; g < (plan.mlp[0].k0_pad >> 3); ++g) {
              st_v4(s.X0raw + g * KG_BYTES + r * 16, 0, 0, 0, 0); st_v4(s.X0act + g * KG_BYTES + r * 16, 0, 0, 0, 0);
            }
            s.P[r] = px; s.P[ROWS + r] = py; s.P[2 * ROWS + r] = pz;
            s.ray[r] = ray; s.t[r] = t; s.valid[r] = ok ? 1 : 0;
          } else {
            float el = 0.f, az = 0.f, dl = 0.f;
            if (ok) {
              nf_elaz(dx, dy, dz, el, az);
              dl = nf_delta(a.ts + ray * a.ts_stride, t, a.T, sqrtf(dx * dx + dy * dy + dz * dz));
            }
            s.el[r] = el; s.az[r] = az; s.delta[r] = dl;
          }
        }
        fence_proxy_async();
        named_bar(1, EPI_THREADS);
        if (lane == 0) mbar_arrive(smem_u32(&s.x0_ready));

        // ---------- the MLPs ----------
        for (int m = m_begin; m < m_end; ++m) {
          const NfMlpPlan& M = plan.mlp[m];
          const int act = M.act;
          for (int j = 0; j < M.n_lin; ++j, ++lin_count) {
            const NfLinPlan& L = M.lin[j];
            const uint32_t buf = lin_count & 1;
            const float* bias = s.bias + (((!a.mlp_only && m) ? lin_base1 : 0) + j) * 256;
            tr_on = (a.debug & 4) && blockIdx.x == 0 && trip == 2 && warp == 0 && lane == 0;
            NF_TRACE(2, (m * 16 + j) * 4 + 0);
            mbar_wait_backoff(smem_u32(&s.acc_full[buf]), (acc_par >> buf) & 1u);
            NF_TRACE(2, (m * 16 + j) * 4 + 1);
            acc_par ^= 1u << buf;
            tc_fence_after();
            const uint32_t t_acc = t_lane + buf * 256;
            if (a.debug & 1) {
              if (!L.is_out) { for (int c = 0; c < 4; ++c) { __syncwarp(); if (lane == 0) mbar_arrive(smem_u32(&s.h_ready[c])); } }
              else if (!a.mlp_only && plan.kind == NF_KIND_PLAIN && m == 0) { __syncwarp(); if (lane == 0) mbar_arrive(smem_u32(&s.x0_ready)); }
            } else if (!L.is_out) {
              // hidden Linear: H <- act(acc + bias) as fp16, chunk by chunk (64 columns = one h_ready)
#ifdef NF_TC_TRACE
              if (act == NF_ACT_SIN) epi_hidden<NF_ACT_SIN>(s, t_acc, bias, half, row, lane, m * 16 + j, tr_on, tr, &tr_n[2]);
              else if (act == NF_ACT_LEAKY) epi_hidden<NF_ACT_LEAKY>(s, t_acc, bias, half, row, lane, m * 16 + j, tr_on, tr, &tr_n[2]);
#else
              if (act == NF_ACT_SIN) epi_hidden<NF_ACT_SIN>(s, t_acc, bias, half, row, lane);
              else if (act == NF_ACT_LEAKY) epi_hidden<NF_ACT_LEAKY>(s, t_acc, bias, half, row, lane);
#endif
              else if (act == NF_ACT_RELU) epi_hidden<NF_ACT_RELU>(s, t_acc, bias, half, row, lane);
              else epi_hidden<NF_ACT_NONE>(s, t_acc, bias, half, row, lane);
            } else if (a.mlp_only) {
              // dump raw outputs in reference column order
              const long long g = u * ROWS + row;
              for (int un = half; un < (L.n_pad >> 4); un += 2) {
                uint32_t v[16];
                tmem_ld16(t_acc + un * 16, v); tmem_ld_wait(); reg_fence16(v);
                if (g < (a.n_dev ? *a.n_dev : a.n)) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const int nt = un * 16 + i;
                    int nr = nt;
                    if (plan.kind == NF_KIND_PLAIN && m == 0) nr = nt == plan.intermediate ? 0 : nt + 1;
                    if (nr < L.n) a.out[g * L.n + nr] = __uint_as_float(v[i]) + bias[nt];
                  }
                }
              }
              tc_fence_before();
            } else if (plan.kind == NF_KIND_PLAIN && m == 0) {
              // density MLP out (tensor order [inter(I), sigma]): x0 of the View head + raw density
              const int act1 = plan.mlp[1].act;
              const int iu = plan.intermediate >> 4;       // 16-column units of intermediate
              for (int un = half; un <= iu; un += 2) {
                uint32_t v[16];
                tmem_ld16(t_acc + un * 16, v); tmem_ld_wait(); reg_fence16(v);
                if (un < iu) {
                  uint32_t o[8], oa[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float x0 = __uint_as_float(v[2 * i]) + bias[un * 16 + 2 * i];
                    const float x1 = __uint_as_float(v[2 * i + 1]) + bias[un * 16 + 2 * i + 1];
                    o[i] = pack_h2(x0, x1); oa[i] = pack_h2(tc_act(x0, act1), tc_act(x1, act1));
                  }
                  uint8_t* d0 = s.X0raw + (un * 2) * KG_BYTES + row * 16; uint8_t* d1 = s.X0act + (un * 2) * KG_BYTES + row * 16;
                  st_v4(d0, o[0], o[1], o[2], o[3]); st_v4(d0 + KG_BYTES, o[4], o[5], o[6], o[7]);
                  st_v4(d1, oa[0], oa[1], oa[2], oa[3]); st_v4(d1 + KG_BYTES, oa[4], oa[5], oa[6], oa[7]);
                } else {
                  s.sig[row] = __uint_as_float(v[0]) + bias[plan.intermediate];
                  const float px = s.P[row], py = s.P[ROWS + row], pz = s.P[2 * ROWS + row], el = s.el[row], az = s.az[row];
                  uint8_t* d0 = s.X0raw + (iu * 2) * KG_BYTES + row * 16; uint8_t* d1 = s.X0act + (iu * 2) * KG_BYTES + row * 16;
                  st_v4(d0, pack_h2(px, py), pack_h2(pz, el), pack_h2(az, 0.f), 0);
                  st_v4(d1, pack_h2(tc_act(px, act1), tc_act(py, act1)), pack_h2(tc_act(pz, act1), tc_act(el, act1)),
                        pack_h2(tc_act(az, act1), 0.f), 0);
                  st_v4(d0 + KG_BYTES, 0, 0, 0, 0); st_v4(d1 + KG_BYTES, 0, 0, 0, 0);
                }
              }
              tc_fence_before();
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(smem_u32(&s.x0_ready));
            } else {
              // final Linear of the path -> colours (and raw density for TinyNeRF) -> composite
              if (half == 0) {
                uint32_t v[16];
                tmem_ld16(t_acc, v); tmem_ld_wait(); reg_fence16(v);
                tc_fence_before();
                float cr, cg, cb;
                if (plan.kind == NF_KIND_TINY) {
                  s.sig[row] = __uint_as_float(v[0]) + bias[0];
                  cr = __uint_as_float(v[1]) + bias[1]; cg = __uint_as_float(v[2]) + bias[2]; cb = __uint_as_float(v[3]) + bias[3];
                } else {
                  cr = __uint_as_float(v[0]) + bias[0]; cg = __uint_as_float(v[1]) + bias[1]; cb = __uint_as_float(v[2]) + bias[2];
                }
                nf_feat_act3(cr, cg, cb, plan.feat_act);
                composite_tile(s, plan, a, map, sub, row, lane, q, cr, cg, cb);
              }
            }
          }
        }
        named_bar(1, EPI_THREADS);   // tile-private smem (sig/delta/ray/...) is free for the next tile
      }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int tc_num_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// nullptr if the tensor path can run this model, else the reason.
const char* tc_unsupported(const NfPlan& p, int only_mlp = -1) {
  if (p.refl_kind == NF_REFL_POSLINVIEW) return "PosLinearView runs on the fp32 pipeline only";
  if (p.kind == NF_KIND_DYN && only_mlp < 0) return "NF_KIND_DYN runs on the fp32 pipeline only in this build";
  int total = 0;
  for (int m = 0; m < p.n_mlps; ++m) {
    if (only_mlp >= 0 && m != only_mlp) continue;
    total += p.mlp[m].n_lin; if (p.mlp[m].k0_pad > X0K) return "x0 wider than 80 columns";
  }
  if (total > MAX_LIN_TOTAL) return "more than 12 Linear layers";
  if (p.enc == NF_ENC_FOURIER && only_mlp != 1) return "Fourier-encoded density MLP runs on the fp32 pipeline only";
  if (p.enc == NF_ENC_HASH && (p.hash_levels & 1)) return "odd number of hash levels";
  if (p.kind == NF_KIND_PLAIN && (p.intermediate & 15)) return "intermediate_size not a multiple of 16";
  return nullptr;
}

// Builds the per-tile MMA program (one entry per weight chunk) for MLPs [m_begin, m_end).
void build_prog(const NfPlan& plan, int m_begin, int m_end, TcProg* P) {
  *P = TcProg{};
  const uint32_t Hs = (uint32_t)offsetof(TcSmem, H), Xr = (uint32_t)offsetof(TcSmem, X0raw), Xa = (uint32_t)offsetof(TcSmem, X0act);
  int nc = 0, lin = 0;
  for (int m = m_begin; m < m_end; ++m) {
    bool x0_first = true;
    for (int j = 0; j < plan.mlp[m].n_lin; ++j, ++lin) {
      const NfLinPlan& L = plan.mlp[m].lin[j];
      const int steps = (L.k0_pad + L.k_hidden) >> 4;
      const uint32_t b_lbo = (uint32_t)L.n_pad * 16u;
      const uint32_t idesc = (1u << 4) | ((uint32_t)(L.n_pad >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
      uint32_t hmask = 0;
      for (int c = 0; c < L.n_chunks; ++c, ++nc) {
        const int nst = steps - SPC * c < SPC ? steps - SPC * c : SPC;
        TcProgEnt& e = P->e[nc];
        uint32_t f = (uint32_t)nst | (c == 0 ? F_FIRST : 0u) | (c == L.n_chunks - 1 ? F_LAST : 0u) | ((lin & 1) ? F_BUF : 0u);
        for (int q4 = 0; q4 < nst; ++q4) {
          const int k = (c * SPC + q4) << 4;
          uint32_t off;
          if (k < L.k0_pad) {
            if (x0_first) { f |= F_WAIT_X0; x0_first = false; }
            off = (L.x0_raw ? Xr : Xa) + (uint32_t)(k >> 3) * KG_BYTES;
          } else {
            const int kh = k - L.k0_pad, hc = kh >> 6;
            if (!(hmask & (1u << hc))) { f |= 0x100u << hc; hmask |= 1u << hc; }
            off = Hs + (uint32_t)(kh >> 3) * KG_BYTES;
          }
          e.a4[q4] = (off >> 4) | ((uint32_t)(KG_BYTES >> 4) << 16);
        }
        e.bstep4 = (2u * b_lbo) >> 4; e.idesc = idesc; e.flags = f; e.bhi = (b_lbo >> 4) << 16;
      }
    }
  }
  P->n_chunks = nc; P->odd_lin = lin & 1;
}

cudaError_t launch_tc(const NfPlan& plan, TcArgs a, long long units, cudaStream_t st) {
#ifdef NF_EXPERIMENTS
  if (const char* dbg = getenv("NF_TC_DEBUG")) a.debug = atoi(dbg);       // timing experiments only: some bits change results
#endif
  if (tc_unsupported(plan, a.mlp_only ? a.which : -1)) return cudaErrorNotSupported;
  cudaError_t e = cudaFuncSetAttribute(k_render_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem));
  if (e != cudaSuccess) return e;
  if (units == 0) return cudaSuccess;
  const int sms = tc_num_sms();
  const int grid = (int)(units < sms ? units : sms);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = sizeof(TcSmem); cfg.stream = st;
  cfg.attrs = nullptr; cfg.numAttrs = 0;
  TcProg prog;
  build_prog(plan, a.mlp_only ? a.which : 0, a.mlp_only ? a.which + 1 : plan.n_mlps, &prog);
  if (a.debug & 4) {
    static long long* d_trace = nullptr;
    if (!d_trace) cudaMalloc(&d_trace, 3 * 512 * 2 * sizeof(long long));
    cudaMemset(d_trace, 0, 3 * 512 * 2 * sizeof(long long));
    a.trace = d_trace;
    cudaError_t e2 = cudaLaunchKernelEx(&cfg, k_render_tc, plan, prog, a);
    cudaStreamSynchronize(st);
    static long long h[3 * 512 * 2];
    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
    long long t0 = -1;
    for (int i = 0; i < 3 * 512; ++i) if (h[2 * i + 1] && (t0 < 0 || h[2 * i + 1] < t0)) t0 = h[2 * i + 1];
    static int printed = 0;
    if (t0 >= 0 && printed++ < 1)
      for (int r = 0; r < 3; ++r) for (int i = 0; i < 512; ++i) if (h[2 * (r * 512 + i) + 1])
        printf("TRACE role=%d tag=%lld t=%lld\n", r, h[2 * (r * 512 + i)], h[2 * (r * 512 + i) + 1] - t0);
    return e2;
  }
  return cudaLaunchKernelEx(&cfg, k_render_tc, plan, prog, a);
}

}  // namespace

cudaError_t nf_launch_render_tc(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, const float* ts,
                                int T, int64_t ts_stride, const float* noise, float* rgb, float* alpha, float* weights,
                                cudaStream_t st) {
  TcArgs a{};
  a.packed = (const uint8_t*)packed; a.rays = rays; a.n_rays = n_rays; a.ts = ts; a.T = T; a.ts_stride = ts_stride;
  a.noise = noise; a.rgb_out = rgb; a.alpha_out = alpha; a.weights_out = weights; a.mlp_only = 0;
  const NfTileMap map(T, ROWS);
  return launch_tc(plan, a, map.units(n_rays), st);
}

cudaError_t nf_launch_mlp_tc(const NfPlan& plan, int which, const void* packed, const float* x0, int64_t n, float* out, cudaStream_t st,
                             const long long* n_dev) {
  TcArgs a{};
  a.packed = (const uint8_t*)packed; a.mlp_only = 1; a.which = which; a.x0 = x0; a.n = n; a.out = out; a.T = ROWS; a.n_dev = n_dev;
  return launch_tc(plan, a, (n + ROWS - 1) / ROWS, st);      // with n_dev, n is the capacity (an upper bound: it sizes the grid)
}
