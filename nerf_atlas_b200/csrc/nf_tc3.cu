// nf_tc3.cu -- the STAGGERED form of the paired (cta_group::2) tensor-core render pipeline.
//
// Data path (that of the lockstep paired kernel of round 1, nf_tc2.cu, since removed): CTA pair, two 128-sample tiles -- "slots" -- in flight per CTA, M = 256 MMAs issued by one
// thread of the leader, fp16 operands in the UMMA canonical no-swizzle K-major layout, 3-stage bulk-copy weight ring, the
// epilogue warps write the next layer's A operand in place).  What changes is the schedule (profiles/r01_trace_paired_*):
//
//  * In the lockstep kernel both slots walk the Linears in lockstep, so at every tile boundary BOTH slots sit in "composite the old tile,
//    hash-encode the new one" while the tensor pipe has nothing to do: 22 % of a round.  Here slot 1 runs half a round
//    (n_lin / 2 Linears) behind slot 0: while one slot is at its tile boundary the other is in the middle of its MLPs and
//    keeps the tensor pipe fed, and the LeakyReLU (density MLP) and sin (View head) epilogues alternate instead of bunching.
//  * The tile boundary itself is one phase: the four cq == 0 warps composite the finished tile while the other twelve
//    already gather the next tile's hash features.
//  * Biases come from shared memory (per slot, double-buffered, fetched one phase ahead) instead of L2: with a 227 KB
//    carve-out there is no L1 left, and the ~300-700 cycle bias loads were the epilogue's largest stall (stall_long_sb).
//  * No per-row bookkeeping in shared memory: (ray, t) of a row is recomputed from the tile index.
//
// A slot's life is a cycle of n = n_lin phases; phase j >= 1 = epilogue of Linear j-1, phase 0 = composite of the previous
// tile (result of Linear n-1) + encode of the next one.  After phase j the issuer runs Linear j for that slot.
//
// Wide x0 (the Mip latent: 144 / 176 columns; the Positional head: 112): the two 80-column x0 buffers are too small, so the
// kernel runs ONE tile in flight per CTA ("single" mode) and keeps x0 in the idle slot's 64 KB activation buffer (the Fourier-
// encoded SDF MLP's 272 columns spill over into the two unused 20 KB x0 buffers that follow it in shared memory).
//
// Warp roles: 0-15 encode/epilogue (TMEM lane quarter q = warp % 4, column quarter cq = warp / 4), 16/18/19
// weight producers (one ring stage each), 17 the MMA issuer (leader CTA only).
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include "nf_common.cuh"
#include "nf_kernels.h"
#include "nf_tc_ptx.cuh"

namespace {
using namespace nf_ptx;

// Biases ride in the MMA: every Linear's weight image carries one extra K-step (16 rows: fp16 hi and lo halves of the fp32 bias,
// then zeros) that is multiplied with a constant [1, 1, 0, ...] block, so the accumulator already holds W x + b and no epilogue
// adds (or loads) a bias.  Only the first 8-row K-group of that step is non-zero, so only IT is streamed: it rides at the end of the
// Linear's last weight chunk (2 KB at N = 256) and the bias MMA reads it with a B descriptor whose LBO is 0 (both K-groups = the
// same core matrices) against an A block [1, 1, 0, ... | 0 ...] of two 128-byte core matrices read with SBO = 0 (all 16 row groups
// = the same core matrix).  A hidden Linear is then 4 ring stages instead of 4 + a one-step stage of its own.
constexpr int X0K = 80;
constexpr int BIAS_PIECE = 2048;                            // the bias K-group of the widest half image (128 rows x 16 B)
constexpr int RING_BYTES = 3 * (4 * 4096 + BIAS_PIECE);     // weight ring per CTA: NST stages of SPCT K-steps (4 KB each at N = 256) + a bias piece
constexpr int MAX_LIN3 = 24;
constexpr int MAX_STAGES3 = 6;

// X0C = x0 columns per slot (80; 112 for the Positional head in boundary-warp mode, whose ring then has two stages), RINGB = ring bytes
constexpr int X0K_POS = 112;
constexpr int RING_BYTES_POS = 2 * (4 * 4096 + BIAS_PIECE);
template <int X0C, int RINGB>
struct Tc3SmemT {
  uint8_t H[2][ROWS * 256 * 2];
  uint8_t X0[2][ROWS * X0C * 2];
  uint8_t W[RINGB];
  uint8_t ones[256];                                         // A operand of the bias K-step: core matrix [8 rows][1, 1, 0, ...] + a zero core matrix (read with SBO = 0)
  float sig[2][2][ROWS];                                     // raw density per row: [slot][tile parity (boundary-warp mode; else 0)]
  float warp_agg[2][4]; int warp_cont[2][4]; float warp_sum[2][4][4]; float carry[2][8];
  unsigned long long w_land[MAX_STAGES3], w_empty[MAX_STAGES3], w_ready[MAX_STAGES3], acc_full[2], a_ready[4];   // a_ready[slot]: x0 + hidden columns 0-127 of the next Linear's operand are written; a_ready[2 + slot]: columns 128-255 too
  unsigned long long bnd_full[2], x0_free[2];   // boundary-warp mode: the path's last Linear is complete / X0[slot] is no longer read
  unsigned long long scr_ready[2][2], mip_land[2], col_read[2];   // scr_ready[slot][tile parity]: the producer runs one tile ahead, two barriers keep its phases from aliasing   // WB: a tile's Mip block is in the scratch / has landed in x0 / the colours are out of TMEM
  uint32_t tmem_base; int pad_;
  int4 lin[MAX_LIN3][2];        // per Linear, for the epilogue warps: {n_pad, bias byte offset, act, flags}, {k0_pad, -, -, -}
                                // flags: 1 = `out` Linear, 2 = `init` Linear, bits 2-3 = what the epilogue of an `out` does:
                                // 1 density-out -> View x0, 2 deformation-out -> deform + encode, 3 the path's last Linear
};
using Tc3Smem = Tc3SmemT<X0K, RING_BYTES>;
static_assert(sizeof(Tc3Smem) <= 227 * 1024 && sizeof(Tc3SmemT<X0K_POS, RING_BYTES_POS>) <= 227 * 1024, "staggered tensor pipeline smem");
static_assert(offsetof(Tc3Smem, X0) == offsetof(Tc3Smem, H) + sizeof(Tc3Smem::H), "single mode: wide x0 runs from H[1] on into X0[]");
static_assert(offsetof(Tc3Smem, W) == offsetof(Tc3Smem, X0) + sizeof(Tc3Smem::X0), "shared wide x0 runs from X0[0] on into the first 12 KB of W");
constexpr int SHARED_X0_COLS = (2 * ROWS * X0K * 2 + 12 * 1024) / (ROWS * 2);          // 208 columns

// Host-built program (kernel parameter => uniform constant loads in the issuing thread).  One record per Linear:
// MMA shape/steps (issuer), this Linear's weight image (producers), (mlp, layer) (epilogue).
struct __align__(16) Tc3Lin {
  uint32_t k0_steps, h_steps, idesc, bstep4;
  uint32_t bhi, mj, w_off, half_bytes;       // mj = m * 16 + j; w_off = byte offset of rank 0's half image, half_bytes = its size
  uint32_t step_bytes, pad0_, pad1_, pad2_;  // bytes of one K-step (16 K-columns) of a half image
};
// single: 0 = two tiles in flight, one 80-column x0 buffer per slot; 1 = ONE tile in flight, x0 (up to 256 + 144 columns) in slot
// 1's H buffer; 2 = two tiles in flight SHARING one wide x0 buffer (up to 208 columns: the two X0 buffers + the first 12 KB of
// the ring region; the ring shrinks to 3 x 12 KB).  Sharing works because x0 is only live while a slot is in the first two
// Linears of an MLP: with slot 1 three Linears behind slot 0 the two live windows never meet (checked on the host).
struct __align__(16) Tc3Prog { int32_t n_lin, lag, single, x0_last; Tc3Lin lin[MAX_LIN3]; };   // x0_last: the last Linear of the chain that reads X0

// Training forward (TRAIN instantiation): where every Linear's input operand (and, for sin MLPs, the cosine of its
// pre-activation) of every tile goes (NfTrainPlan, nf_common.cuh).  Offsets in 256-byte units from Tc3Args::ws.
struct __align__(16) Tc3TrainLin { uint32_t a_off256, c_off256, a_tile256, k0; int32_t skip0, skip1, pad0_, pad1_; };   // k0 = columns of THIS Linear's x0 part; skip0/1: for an `init` Linear, the Linears of its MLP that consume act(x0) (-1: none)
struct __align__(16) Tc3Train { Tc3TrainLin lin[MAX_LIN3]; long long n_tiles; float* sigma_out; float* rgbraw_out; };

struct Tc3Args {
  const uint8_t* packed;
  const float* rays; long long n_rays;
  const float* ts; int T; long long ts_stride;
  const float* noise; const float* ray_time;
  float* rgb_out; float* alpha_out; float* weights_out;
  int debug;          // NF_TC_DEBUG (timing experiments): 256 = poll acc_full with backoff, 512 = try_wait with a short suspend hint
  NfMipIn mip;        // Mip encoder inputs (plan.mip != NF_MIP_NONE)
  uint8_t* ws;        // TRAIN: the training workspace (activation stash)
  const float* pts;   // AUX: explicit sample positions [R,T,3] (from_pts, reference nerf.py:340-361) instead of r_o + ts r_d
  const float* bg_rand;   // AUX / DYN: NF_BG_RANDOM draws [R]
  float* pts_out; float* dp_out; float* rigid_dp_out; float* rigidity_out;   // DYN side channels (runner.py:694-700,769,777-781)
  uint8_t* scratch;   // WB (Mip, boundary warps): per CTA [slot][tile parity] 24 KB images of the 96 Mip features of a tile
  long long* stats;   // NF_TC_STATS builds: time-in-state counters of CTA 0 (issuer, producer 0, epilogue warps 0 and 15)
};

#ifdef NF_TC_STATS
#define ST_DECL long long st_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long st_t = clock64(); const long long st_begin = st_t; (void)st_begin
#define ST_MARK() (st_t = clock64())
#define ST_ADD(i) do { const long long st_now = clock64(); st_acc[i] += st_now - st_t; st_t = st_now; } while (0)
#define ST_INC(i) (++st_acc[i])
#define ST_FLUSH(base, cond) do { if ((cond) && a.stats) { a.stats[(base)] = clock64() - st_begin; for (int st_i = 0; st_i < 7; ++st_i) a.stats[(base) + 1 + st_i] = st_acc[st_i]; } } while (0)
#else
#define ST_DECL do { } while (0)
#define ST_MARK() do { } while (0)
#define ST_ADD(i) do { } while (0)
#define ST_INC(i) do { } while (0)
#define ST_FLUSH(base, cond) do { } while (0)
#endif

#ifndef NF_ACC_POLL
#define NF_ACC_POLL 0
#endif
// acc_full wait of the 16 epilogue warps.  Default: hardware-suspended try_wait (no issue slots burnt).
__device__ __forceinline__ void wait_acc(uint32_t bar, uint32_t parity, int debug) {
  if (debug & 256) { mbar_wait_backoff(bar, parity); return; }
  if (debug & 512) {
    uint32_t spins = 0, ok = 0;
    while (true) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(ok) : "r"(bar), "r"(parity), "r"(64u) : "memory");
      if (ok) break;
      if (++spins > (1u << 24)) mbar_timeout(bar);
    }
    return;
  }
#if NF_ACC_POLL == 1
  mbar_wait(bar, parity);                                  // experiment: pure test_wait polling by all 16 warps
#elif NF_ACC_POLL >= 2
  for (int i = 0; i < NF_ACC_POLL; ++i) if (mbar_test_wait(bar, parity)) return;     // experiment: poll a few times, then suspend
  mbar_wait_suspend(bar, parity);
#else
  mbar_wait_suspend(bar, parity);
#endif
}

// sin on the FMA pipe: the MUFU unit retires 16 sines per cycle per SM, so a 128 x 256 sin epilogue cannot finish in less than
// 2048 cycles -- as long as the layer's MMAs.  A share of every 16 columns (NF_SIN_POLY_PAIRS of the 8 column pairs) takes
// this route instead: revolutions t = x / 2pi, r = t - rint(t) in [-0.5, 0.5] (magic-number rounding), then an odd degree-9
// minimax polynomial for sin(2 pi r): max error 6.3e-6, far inside the fp16 rounding of the result (2.4e-4).
// MEASURED (profiles/r01_sin_poly_sweep.txt): 0 pairs 113.8 ms/frame, 2 pairs 115.0, 3 pairs 116.2, 4 pairs 119.7 -- the sin
// epilogue is not MUFU-throughput-bound but issue/latency-bound, so the extra FMA-pipe instructions only cost.  Default 0.
// boundary-warp mode (four more warps own the tile boundary); A/B switches of this round's measurements
#ifndef NF_WIDE_CHUNKS
#define NF_WIDE_CHUNKS 1      // narrow Linears: 8 / 16 K-steps per ring stage
#endif
#ifndef NF_BW
#define NF_BW 1
#endif
#ifndef NF_WB_NST2
#define NF_WB_NST2 0
#endif
#ifndef NF_WB
#define NF_WB 1               // the Mip encoder's shared-wide-x0 schedule with boundary warps (else all work on the 16 epilogue warps)
#endif
#ifndef NF_POS_BW
#define NF_POS_BW 1           // the Positional head on the boundary-warp kernel (else the shared-wide-x0 schedule)
#endif
#ifndef NF_ISSUE_STRAIGHT
#define NF_ISSUE_STRAIGHT 1   // straight-line issuer code for the 16-step Linears
#endif
#ifndef NF_REG_E
#define NF_REG_E 88           // setmaxnreg: epilogue warps; NF_REG_W4: the producer / issuer warp group
#define NF_REG_W4 40
#endif
// the boundary warps also write the View head's [p, elaz] x0 tail (else the density-out epilogue does)
#ifndef NF_BW_PRETAIL
#define NF_BW_PRETAIL 1
#endif
#ifndef NF_SIN_POLY_PAIRS
#define NF_SIN_POLY_PAIRS 0
#endif
__device__ __forceinline__ float sin_poly(float x) {
  const float t = x * 0.15915494309189535f;
  const float k = (t + 12582912.f) - 12582912.f;
  const float r = t - k, r2 = r * r;
  float p = fmaf(32.78138732910156f, r2, -74.47799682617188f);
  p = fmaf(p, r2, 81.36681365966797f);
  p = fmaf(p, r2, -41.331214904785156f);
  p = fmaf(p, r2, 6.283055782318115f);
  return p * r;
}

// ---- epilogue of a hidden Linear: H <- fp16(act(acc + bias)), bias from shared memory ---------------------
__device__ __forceinline__ void st_global_v4(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");   // streaming: the stash is read once, much later
}
// gA (TRAIN, nullable): global image of the same tile -- the activated output is ALSO stashed there (the backward's dW operand
// and LeakyReLU' mask); gC (TRAIN, sin MLPs): cos of the pre-activation (the backward's sin').
template <int ACT, int NCQ, bool TRAIN = false>
__device__ __forceinline__ void epi_hidden3(uint8_t* __restrict__ H, uint32_t t_acc, const float* __restrict__ bias_s, int cq, int row,
                                            uint8_t* __restrict__ gA = nullptr, uint8_t* __restrict__ gC = nullptr) {
  // the 16 units of 16 columns are dealt round-robin to the NCQ warps of a lane quarter (unit = cq, cq + NCQ, ...); the next
  // unit's TMEM load is in flight while this one is converted
  uint32_t v[2][16];
  tmem_ld16(t_acc + cq * 16, v[0]);
  constexpr int MAXU = (16 + NCQ - 1) / NCQ;
#pragma unroll
  for (int u = 0; u < MAXU; ++u) {
    const int un = cq + u * NCQ;
    if (un >= 16) break;
    const int col = un * 16;
    tmem_ld_wait();
    reg_fence16(v[u & 1]);
    if (un + NCQ < 16) tmem_ld16(t_acc + col + NCQ * 16, v[(u + 1) & 1]);
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float x0 = __uint_as_float(v[u & 1][4 * i]), x1 = __uint_as_float(v[u & 1][4 * i + 1]);
      const float x2 = __uint_as_float(v[u & 1][4 * i + 2]), x3 = __uint_as_float(v[u & 1][4 * i + 3]);
      if (ACT == NF_ACT_SIN && 2 * i < NF_SIN_POLY_PAIRS) o[2 * i] = pack_h2(sin_poly(x0), sin_poly(x1));
      else o[2 * i] = act_pack_t<ACT>(x0, x1);
      if (ACT == NF_ACT_SIN && 2 * i + 1 < NF_SIN_POLY_PAIRS) o[2 * i + 1] = pack_h2(sin_poly(x2), sin_poly(x3));
      else o[2 * i + 1] = act_pack_t<ACT>(x2, x3);
    }
    uint8_t* dst = H + (col >> 3) * KG_BYTES + row * 16;
    st_v4(dst, o[0], o[1], o[2], o[3]); st_v4(dst + KG_BYTES, o[4], o[5], o[6], o[7]);
    if (TRAIN && gA) {
      uint8_t* g = gA + (col >> 3) * KG_BYTES + row * 16;
      st_global_v4(g, o[0], o[1], o[2], o[3]); st_global_v4(g + KG_BYTES, o[4], o[5], o[6], o[7]);
      if (ACT == NF_ACT_SIN && gC) {
        uint32_t c[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          c[2 * i] = pack_h2(__cosf(__uint_as_float(v[u & 1][4 * i]) + b.x), __cosf(__uint_as_float(v[u & 1][4 * i + 1]) + b.y));
          c[2 * i + 1] = pack_h2(__cosf(__uint_as_float(v[u & 1][4 * i + 2]) + b.z), __cosf(__uint_as_float(v[u & 1][4 * i + 3]) + b.w));
        }
        uint8_t* gc = gC + (col >> 3) * KG_BYTES + row * 16;
        st_global_v4(gc, c[0], c[1], c[2], c[3]); st_global_v4(gc + KG_BYTES, c[4], c[5], c[6], c[7]);
      }
    }
  }
}

// sin epilogue, software-pipelined by hand across the 16-column units: the 16 MUFU.SINs of unit u are ISSUED first, then the
// warp waits for unit u+1's TMEM load and does its bias add + range scaling while the MUFU results arrive, and only then packs
// and stores unit u.  (ptxas cannot do this itself: the TMEM wait is an ordering point.)
__device__ __forceinline__ float mufu_sin(float t) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t)); return y; }
template <int NCQ>
__device__ __forceinline__ void epi_hidden3_sin_pipelined(uint8_t* __restrict__ H, uint32_t t_acc, const float* __restrict__ bias_s, int cq, int row) {
  uint32_t v[16];
  float x[16], y[16];
  tmem_ld16(t_acc + cq * 16, v);
  tmem_ld_wait(); reg_fence16(v);
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[i]);
  constexpr int MAXU = (16 + NCQ - 1) / NCQ;
#pragma unroll
  for (int u = 0; u < MAXU; ++u) {
    const int un = cq + u * NCQ;
    if (un >= 16) break;
    const int col = un * 16;
    const bool more = un + NCQ < 16;
    if (more) tmem_ld16(t_acc + col + NCQ * 16, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = mufu_sin(x[i]);                // issue the unit's sines ...
    if (more) {
      tmem_ld_wait(); reg_fence16(v);                                      // ... and prepare the next unit while they execute
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[i]);
    }
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = pack_h2(y[2 * i], y[2 * i + 1]);
    uint8_t* dst = H + (col >> 3) * KG_BYTES + row * 16;
    st_v4(dst, o[0], o[1], o[2], o[3]); st_v4(dst + KG_BYTES, o[4], o[5], o[6], o[7]);
  }
}

// Split hand-over (NCQ == 4): the warp first reads ALL 64 of its accumulator columns into registers -- from then on the
// accumulator may be overwritten -- converts and stores its two units of columns 0-127, hands that half to the issuer
// (a_ready[slot]: the next Linear's first eight hidden K-steps start), and then converts the two units of columns 128-255 while
// those MMAs run (the caller arrives on a_ready[2 + slot] at the end of the phase).  With every load issued up front there is
// no TMEM wait between units, so the sines of different units overlap without hand pipelining.
#ifndef NF_SIN_POLY_SPLIT
#define NF_SIN_POLY_SPLIT 0
#endif
#ifndef NF_SPLIT_HANDOFF
#define NF_SPLIT_HANDOFF 1
#endif
template <int ACT, bool TRAIN>
__device__ __forceinline__ void epi_hidden_split(uint8_t* __restrict__ H, uint32_t t_acc, int cq, int row, int lane, uint32_t a_lo_leader,
                                                 uint8_t* __restrict__ gA = nullptr, uint8_t* __restrict__ gC = nullptr) {
  uint32_t v[4][16];
#pragma unroll
  for (int u = 0; u < 4; ++u) tmem_ld16(t_acc + (cq + 4 * u) * 16, v[u]);
  tmem_ld_wait();
#pragma unroll
  for (int u = 0; u < 4; ++u) reg_fence16(v[u]);
  tc_fence_before();
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int col = (cq + 4 * u) * 16;
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // NF_SIN_POLY_SPLIT of every 8 column pairs take the FMA-pipe polynomial instead of MUFU.SIN (the sin epilogue's floor is the
      // MUFU rate: 16 per cycle per SM = 2048 cycles per 128 x 256 tile)
      if (ACT == NF_ACT_SIN && i < NF_SIN_POLY_SPLIT) o[i] = pack_h2(sin_poly(__uint_as_float(v[u][2 * i])), sin_poly(__uint_as_float(v[u][2 * i + 1])));
      else o[i] = act_pack_t<ACT>(__uint_as_float(v[u][2 * i]), __uint_as_float(v[u][2 * i + 1]));
    }
    uint8_t* dst = H + (col >> 3) * KG_BYTES + row * 16;
    st_v4(dst, o[0], o[1], o[2], o[3]); st_v4(dst + KG_BYTES, o[4], o[5], o[6], o[7]);
    if (TRAIN && gA) {
      uint8_t* g = gA + (col >> 3) * KG_BYTES + row * 16;
      st_global_v4(g, o[0], o[1], o[2], o[3]); st_global_v4(g + KG_BYTES, o[4], o[5], o[6], o[7]);
      if (ACT == NF_ACT_SIN && gC) {
        uint32_t c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = pack_h2(__cosf(__uint_as_float(v[u][2 * i])), __cosf(__uint_as_float(v[u][2 * i + 1])));
        uint8_t* gc = gC + (col >> 3) * KG_BYTES + row * 16;
        st_global_v4(gc, c[0], c[1], c[2], c[3]); st_global_v4(gc + KG_BYTES, c[4], c[5], c[6], c[7]);
      }
    }
    if (u == 1) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(a_lo_leader);
    }
  }
}

// x0 raw -> act(x0), in place (the `init` Linear consumed the raw form; the skip Linear wants the activated one)
// gRaw / gAct0 / gAct1 (TRAIN, nullable): stash of the raw x0 (the `init` Linear's operand) and of act(x0) (the skip Linears')
__device__ __forceinline__ void x0_activate3(uint8_t* X0, int k0_pad, int act, int g_tid, int n_threads,
                                             uint8_t* gRaw = nullptr, uint8_t* gAct0 = nullptr, uint8_t* gAct1 = nullptr) {
  const int n16 = (k0_pad >> 3) * ROWS;             // 16-byte groups
  for (int i = g_tid; i < n16; i += n_threads) {
    uint4 q = *reinterpret_cast<uint4*>(X0 + i * 16);
    if (gRaw) st_global_v4(gRaw + i * 16, q.x, q.y, q.z, q.w);
    uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
      w[k] = pack_h2(tc_act(f.x, act), tc_act(f.y, act));
    }
    *reinterpret_cast<uint4*>(X0 + i * 16) = q;
    if (gAct0) st_global_v4(gAct0 + i * 16, q.x, q.y, q.z, q.w);
    if (gAct1) st_global_v4(gAct1 + i * 16, q.x, q.y, q.z, q.w);
  }
}

// this thread's share (levels first, first+stride, ...) of the hash features of one row -> x0 columns [4*lvl, 4*lvl+4)
__device__ __forceinline__ void hash_x0(uint8_t* X0, const float4* tables, const NfPlan& plan, float px, float py, float pz,
                                        int row, int first, int stride) {
  if (first < 0) return;
  for (int lvl = first; lvl < plan.hash_levels; lvl += stride) {
    const float4 f = nf_hash_level(tables + (size_t)lvl * (plan.hash_mask + 1), px, py, pz, plan.hash_res[lvl],
                                   plan.hash_primes[0], plan.hash_primes[1], plan.hash_primes[2], plan.hash_mask, nullptr);
    *reinterpret_cast<uint2*>(X0 + (lvl >> 1) * KG_BYTES + row * 16 + (lvl & 1) * 8) = make_uint2(pack_h2(f.x, f.y), pack_h2(f.z, f.w));
  }
}
// the [p, p] tail of a hash-encoded x0 (tensor order [feats, p, p]) and the zero padding up to k0_pad; the Mip latent's
// column groups [mip0, mip0 + 12) are written by the Mip writers and skipped here
__device__ __forceinline__ void hash_x0_tail(uint8_t* X0, const NfPlan& plan, int k0_pad, float px, float py, float pz, int row, int mip_col = -1) {
  const int kg = plan.hash_levels >> 1;
  st_v4(X0 + kg * KG_BYTES + row * 16, pack_h2(px, py), pack_h2(pz, px), pack_h2(py, pz), 0);
  const int m0 = mip_col >= 0 ? mip_col >> 3 : 1 << 20;
  for (int g = kg + 1; g < (k0_pad >> 3); ++g)
    if (g < m0 || g >= m0 + NF_MIP_FEATS / 8) st_v4(X0 + g * KG_BYTES + row * 16, 0, 0, 0, 0);
}
// this thread's share (features first, first + stride, ...) of the Mip latent of one row -> x0 columns [col0, col0 + 96)
#ifndef NF_MIP_INLINE
#define NF_MIP_INLINE 0
#endif
#if NF_MIP_INLINE
__device__ __forceinline__
#else
__device__ __noinline__          // a call keeps the Mip row state (NfMipRow) out of the main loop's register allocation
#endif
void mip_x0(uint8_t* X0, const NfMipIn& mip, int col0, bool ok, long long ray, int t, int row, int first, int stride) {
  NfMipRow R;
  const bool per_row = mip.mode != NF_MIP_CYLINDER_REF;               // the reference layout gathers its variance per feature
  if (ok && per_row && first < NF_MIP_FEATS / 2) nf_mip_row(mip, ray, t, R);
  for (int cc = first; cc < NF_MIP_FEATS / 2; cc += stride) {            // (sin, cos) pairs share mean, variance and exponential
    float fs = 0.f, fc = 0.f;
    if (ok) { if (per_row) nf_mip_pair_of_row<true>(R, cc, fs, fc); else nf_mip_feature_pair<true>(mip, ray, t, cc, fs, fc); }
    const int c0 = col0 + cc, c1 = c0 + NF_MIP_FEATS / 2;
    *reinterpret_cast<__half*>(X0 + (c0 >> 3) * KG_BYTES + row * 16 + (c0 & 7) * 2) = __float2half_rn(fs);
    *reinterpret_cast<__half*>(X0 + (c1 >> 3) * KG_BYTES + row * 16 + (c1 & 7) * 2) = __float2half_rn(fc);
  }
}

// work unit of (pass, slot) for this CTA, and the sub-tile within a ray (T > 128)
__device__ __forceinline__ void unit_of(int pass, int slot, int tpr, int nslot, long long& u, int& sub) {
  const int trip = tpr == 1 ? pass : pass / tpr;
  sub = pass - trip * tpr;
  u = ((long long)trip * gridDim.x + blockIdx.x) * nslot + slot;     // nslot = tiles in flight per CTA (2; 1 in single mode)
}

// ---- composite of one tile by four warps (thread = row); reference nerf.py:60-80 ----
// A tile holds up to 128 / Tp + 2 ray segments; the first may continue a ray from the previous tile of the unit (carry in), the
// last may be unfinished (carry out).  Products and sums are LEFT folds in sample order -- 32-sample chunk by chunk, across
// tile boundaries -- so a ray's result does not depend on where in a unit it happens to sit.
template <class Smem>
__device__ __forceinline__ void composite_tile3(Smem& s, int slot, const NfPlan& plan, const Tc3Args& a, const NfStreamMap& map,
                                                long long u, int sub, int row, int lane, int q, float cr, float cg, float cb,
                                                const float* sig, float* sigma_out = nullptr, bool may_use_H = true) {
  long long ray; int t;
  const bool valid = map.locate(u, sub, row, a.n_rays, ray, t);       // t is the position within the (padded) ray even when invalid
  float al = 0.f;
  if (valid) {
    float sr = sig[row];
    if (a.noise) sr += __ldg(a.noise + ray * a.T + t);
    if (sigma_out) sigma_out[ray * a.T + t] = sr;                     // TRAIN: the raw density the composite consumed (noise included)
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
  if (lane == 31) s.warp_agg[slot][q] = incl;             // product over the last ray segment of this warp
  float* carry = s.carry[slot];
  const float carry_T = carry[0];                          // read before anyone overwrites it (the writes come after the 2nd barrier)
  named_bar(3 + slot, 128);
  // transmittance entering this warp for a ray that began before it: left fold from the ray's first chunk in this tile
  float c = 1.f;
  if (t > lane) {
    const int first = row - t;                             // row of the ray's sample 0 (negative: it began in an earlier tile)
    c = first < 0 ? carry_T : 1.f;
    for (int v = first < 0 ? 0 : first >> 5; v < q; ++v) c *= s.warp_agg[slot][v];
  }
  const float trans = excl * c;
  const float w = al * trans;
  if (valid) {
    if (a.alpha_out) a.alpha_out[ray * a.T + t] = al;
    if (a.weights_out) a.weights_out[ray * a.T + t] = w;
  }
  const float wr = w * cr, wg = w * cg, wb = w * cb, wl = (valid && t < a.T - 1) ? w : 0.f;
  float* wrgb = reinterpret_cast<float*>(s.H[slot]);          // H[slot] is dead between the last Linear's MMA and the next tile
  const bool fast = (map.Tp & 31) == 0 || !may_use_H;         // no warp straddles two rays: per-warp sums suffice (boundary-warp mode: always, H is live)
  if (fast) {
    float x0 = wr, x1 = wg, x2 = wb, x3 = wl;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      x0 += __shfl_xor_sync(0xffffffffu, x0, d); x1 += __shfl_xor_sync(0xffffffffu, x1, d);
      x2 += __shfl_xor_sync(0xffffffffu, x2, d); x3 += __shfl_xor_sync(0xffffffffu, x3, d);
    }
    if (lane == 0) { float* ws = s.warp_sum[slot][q]; ws[0] = x0; ws[1] = x1; ws[2] = x2; ws[3] = x3; }
  } else {
    float* wq = wrgb + row * 4;
    wq[0] = wr; wq[1] = wg; wq[2] = wb; wq[3] = wl;
  }
  // thread i owns ray segment i of this tile
  const int q0 = sub * ROWS;
  const int rl_first = q0 / map.Tp;
  const int rl = rl_first + row;
  const int seg_begin = rl * map.Tp - q0, seg_end = seg_begin + map.Tp;      // rows of the segment, before clipping to the tile
  const bool owner = seg_begin < ROWS && rl < map.rpu;
  const bool cont = owner && seg_begin < 0, ends = owner && seg_end <= ROWS;
  float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f, pT = 1.f;
  if (cont) { pT = carry_T; o0 = carry[1]; o1 = carry[2]; o2 = carry[3]; o3 = carry[4]; }
  named_bar(3 + slot, 128);
  if (owner) {
    const int rb = seg_begin < 0 ? 0 : seg_begin, re = seg_end > ROWS ? ROWS : seg_end;
    if (fast) {
      for (int v = rb >> 5; v < (re >> 5); ++v) { const float* ws = s.warp_sum[slot][v]; o0 += ws[0]; o1 += ws[1]; o2 += ws[2]; o3 += ws[3]; }
    } else {
      for (int i = rb; i < re; ++i) { const float* x = wrgb + i * 4; o0 += x[0]; o1 += x[1]; o2 += x[2]; o3 += x[3]; }
    }
    const long long r = u * map.rpu + rl;
    if (ends) {
      if (r < a.n_rays) {
        float skyv = plan.bg == NF_BG_WHITE ? 1.f - o3 : 0.f;
        if (plan.bg == NF_BG_RANDOM && a.bg_rand) skyv = __ldg(a.bg_rand + r) * (1.f - o3);      // random_color (nerf.py:100-103)
        a.rgb_out[r * 3 + 0] = o0 + skyv; a.rgb_out[r * 3 + 1] = o1 + skyv; a.rgb_out[r * 3 + 2] = o2 + skyv;
      }
    } else {
      // the unfinished last segment: carry the sums and the transmittance (left fold over its chunks; it starts warp-aligned or
      // at row 0, and in the non-fast case its first warp's aggregate is that of the warp's last segment = this ray)
      for (int v = rb >> 5; v < 4; ++v) pT *= s.warp_agg[slot][v];
      carry[0] = pT; carry[1] = o0; carry[2] = o1; carry[3] = o2; carry[4] = o3;
    }
  }
}

// =====================================================================================================
// NST ring stages of SPCT K-steps (+ a bias piece) each (3 x 18 KB); NCQ epilogue warps per TMEM lane quarter.  Warps
// 0..4*NCQ-1 encode/epilogue, then NST weight producers (one stage each; the first also owns the TMEM allocation), then the
// MMA issuer (highest warp id).
// WIDE: the single-tile wide-x0 mode (Mip latent, Positional head) is compiled in.  The common path uses the WIDE = false
// instantiation: with the wide code merely branched around, it ran 3.9 % slower (measured by bisection on one box).
// DYN: the DynamicNeRF chain (deformation-out epilogue, Bezier, re-encode) is compiled in; same reason.
// TRAIN: the training forward -- every Linear's input operand (+ cos for sin MLPs) and the raw density / colours of every
// sample are stashed in the training workspace for nf_render_backward (nf_train.cu).
// AUX: explicit sample positions (from_pts) and the random background are compiled in (kept out of the common instantiation for
// the same reason as WIDE / DYN).
// WIDE: 0 = no wide-x0 code, 1 = wide x0 without the Mip encoder (Positional head, Fourier SDF), 2 = with it.
// BW ("boundary warps"): four more warps (one per TMEM lane quarter) own the tile boundary -- they encode the next tile's x0 as
// soon as the chain's last x0 consumer is complete (x0_free), read the finished tile's colours when its last Linear is complete
// (bnd_full), hand the slot back to the issuer and only then composite -- so the 16 epilogue warps never leave the MLP phases
// and the ~7 K-cycle boundary is off both the slot's critical path and the epilogue warps' time.  Needs T % 32 == 0, WIDE == 0.
// X0C = 112 (with BW, WIDE = 1, NST = 2): the Positional head (reference src/refl.py:230-245, the makefile's first target) with ITS OWN
// x0 buffer per slot -- [inter(64) | hash'(32) | p, p | pad] -- instead of the shared-wide-x0 schedule: the boundary warps gather the
// head's hash features into columns 64.. at encode time (the density MLP's x0 ends at column 48), so the epilogue warps see the
// same phases as the View head's.  The 32 KB come out of the ring: two 18 KB stages.
template <int NST, int SPCT, int NCQ, int WIDE, bool DYN, bool TRAIN = false, bool AUX = false, bool BW = false, int X0C = X0K>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32 * (4 * NCQ + 3 + 1 + (BW ? 4 : 0)), 1)
k_render_tc3(const __grid_constant__ NfPlan plan, const __grid_constant__ Tc3Prog prog, const Tc3Args a, const __grid_constant__ Tc3Train tr) {
  static_assert(!BW || NCQ == 4, "boundary warps: warp groups {0-15, 16-19, 20-23}");
  // WB: the Mip encoder's shared-wide-x0 schedule with boundary warps.  The x0 buffer is shared by the two slots, so the epilogue warps
  // keep the boundary phase (hash gathers, p) -- but the boundary warps take the colour read + composite, and they compute every tile's
  // 96 Mip features ONE TILE AHEAD into an L2-resident scratch image (the byte image of the 12 K-groups), from where one bulk copy
  // brings them into the x0 buffer at the boundary phase (density x0) and again at the density-out phase (View x0).
  constexpr bool WB = BW && WIDE == 2;
  constexpr uint32_t MIP_BLOCK = NF_MIP_FEATS / 8 * KG_BYTES;      // 24 KB
  static_assert(X0C == X0K || (X0C == X0K_POS && BW && WIDE == 1 && NST == 2), "the 112-column instantiation is the Positional head's");
  constexpr int RINGB = X0C == X0K ? RING_BYTES : RING_BYTES_POS;
  using Smem = Tc3SmemT<X0C, RINGB>;
  constexpr int STAGE_BYTES = SPCT * 4096 + BIAS_PIECE;
  static_assert(NST * STAGE_BYTES <= RINGB && NST <= 3, "ring geometry");
  constexpr int RING_OFF = RINGB - NST * STAGE_BYTES;     // a smaller ring sits at the END of the region (shared wide x0 in front)
  constexpr int EPIW = 4 * NCQ, EPI_THREADS = 32 * EPIW;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const NfStreamMap map(a.T, ROWS);
  const long long units = map.units(a.n_rays);
  const bool single = WIDE && prog.single == 1;    // one tile in flight: slot 1 never runs
  const bool shared_x0 = WIDE && prog.single == 2; // two tiles in flight, one wide x0 buffer
  const int nslot = single ? 1 : 2;
  const int trips = (int)((units + (long long)nslot * gridDim.x - 1) / ((long long)nslot * gridDim.x));   // every CTA, every slot: same trip count
  const int passes = trips * map.tpr;
  const int n = prog.n_lin, lag = prog.lag;
  const int nsteps = passes * n;                   // MMA steps (Linears) per slot

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(smem_u32(&s.w_land[i]), 1); mbar_init(smem_u32(&s.w_empty[i]), 1); mbar_init(smem_u32(&s.w_ready[i]), 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s.acc_full[i]), 1); mbar_init(smem_u32(&s.a_ready[i]), 2 * EPIW); mbar_init(smem_u32(&s.a_ready[2 + i]), 2 * EPIW); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s.bnd_full[i]), 1); mbar_init(smem_u32(&s.x0_free[i]), EPIW); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s.scr_ready[i][0]), 4); mbar_init(smem_u32(&s.scr_ready[i][1]), 4); mbar_init(smem_u32(&s.mip_land[i]), 1); mbar_init(smem_u32(&s.col_read[i]), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EPIW) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 16) *reinterpret_cast<uint4*>(s.ones + threadIdx.x * 16) = threadIdx.x < 8 ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  if (threadIdx.x < prog.n_lin) {
    // the epilogue's per-Linear facts, copied once from the kernel parameters (dependent indexed constant loads cost
    // ~250 cycles each when they miss the constant cache: ~600 cycles per phase before this table existed)
    const uint32_t mj = prog.lin[threadIdx.x].mj;
    const NfMlpPlan& M = plan.mlp[mj >> 4];
    const NfLinPlan& L = M.lin[mj & 15];
    const int role = !L.is_out ? 0 : (int)threadIdx.x == prog.n_lin - 1 ? 3 : (mj >> 4) == 2 ? 2 : 1;
    s.lin[threadIdx.x][0] = make_int4(L.n_pad, (int)L.b16_off, M.act, (L.is_out ? 1 : 0) | ((mj & 15) == 0 ? 2 : 0) | (role << 2));
    s.lin[threadIdx.x][1] = make_int4(M.k0_pad, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (s.tmem_base != 0) __trap();

  // ---- the MMA issuer (one thread of the leader CTA) ----
  // A tcgen05.mma costs the issuing thread >= 76 cycles and a commit + barrier probe ~190 (profiles/r01_mma_issue_microbench.csv):
  // >= 500 cycles per 4-MMA ring stage against 512 cycles of execution before any bookkeeping, and a lone warp retires about one
  // instruction per 4-5 cycles -- the instruction count per chunk decides whether the tensor pipe is fed (round 2: ~90 -> ~45
  // instructions per chunk took the frame from 103 to 96 ms).  Hence: chunks that read ONE activation buffer, straight-line code
  // for the 16-step Linears, the probe of the NEXT ring stage issued before the current chunk's MMAs (its ~150-cycle latency is
  // hidden).  (One issuing thread per slot, the two sharing the ring, was built and trapped on the hardware; removed.)
  auto run_issuer = [&]() {
    constexpr uint32_t mask = 3u;
    uint32_t stage = 0, phase = 0, a_par = 0;
    const uint32_t base4 = smem_u32(smem_raw) >> 4;
    const uint32_t w4 = base4 + (uint32_t)((offsetof(Smem, W) + RING_OFF) >> 4);
    const uint32_t bar_wready = smem_u32(&s.w_ready[0]), bar_wempty = smem_u32(&s.w_empty[0]);
    const uint32_t bar_a = smem_u32(&s.a_ready[0]), bar_acc = smem_u32(&s.acc_full[0]), bar_bnd = smem_u32(&s.bnd_full[0]);
    const uint32_t a_lbo = (uint32_t)(KG_BYTES >> 4) << 16, kstep4 = (uint32_t)(2 * KG_BYTES) >> 4;
    const uint64_t ones_desc = ((uint64_t)0x4000u << 32) | (base4 + (uint32_t)(offsetof(Smem, ones) >> 4)) | (8u << 16);   // SBO = 0, LBO = 128 B
    int li0 = 0, li1 = 0;
    bool w_ok = false;                               // ring stage `stage` is known to be full
    ST_DECL;
    for (int k = 0; k < nsteps + lag; ++k) {
#pragma unroll 1
      for (uint32_t slot = 0; slot < 2; ++slot) {
        const int kl = k - (slot ? lag : 0);
        if (kl < 0 || kl >= nsteps || (slot && single)) continue;
        const int li = slot ? li1 : li0;
        const int ln = li + 1 == n ? 0 : li + 1;
        if (slot) li1 = ln; else li0 = ln;
        const uint4 r0 = *reinterpret_cast<const uint4*>(&prog.lin[li].k0_steps);
        const uint32_t k0s = r0.x, hs = r0.y, idesc = r0.z, bstep4 = r0.w;
        const uint32_t bhi = prog.lin[li].bhi;
        ST_ADD(2);
        mbar_wait(bar_a + slot * 8u, (a_par >> slot) & 1u); a_par ^= 1u << slot;
        ST_ADD(0);
        tc_fence_after();
        const uint32_t d_tmem = slot * 256u;
        const uint32_t x4 = (single ? base4 + (uint32_t)((offsetof(Smem, H) + sizeof(s.H[0])) >> 4)
                                    : base4 + (uint32_t)(offsetof(Smem, X0) >> 4) + (shared_x0 ? 0u : slot * (uint32_t)(sizeof(s.X0[0]) >> 4))) | a_lbo;
        const uint32_t h4 = (base4 + (uint32_t)(offsetof(Smem, H) >> 4) + slot * (uint32_t)(sizeof(s.H[0]) >> 4)) | a_lbo;
        // one ring stage: nst K-steps read from ONE activation buffer; `last`: the Linear's last chunk, followed by the bias
        // K-step (its one non-zero K-group follows the chunk's data steps; LBO = 0 in the B descriptor)
        auto chunk = [&](const uint32_t a4, const uint32_t nst, const uint32_t acc0, const bool last) {
          if (!w_ok) { ST_ADD(2); mbar_wait(bar_wready + stage * 8u, phase); ST_ADD(1); ST_INC(3); }
          const uint32_t nstage = stage + 1 == NST ? 0u : stage + 1u, nphase = stage + 1 == NST ? phase ^ 1u : phase;
          w_ok = (last && mask != 3u) ? false : mbar_test_wait(bar_wready + nstage * 8u, nphase);       // consumed by the next chunk
          tc_fence_after();
          const uint32_t wb = w4 + stage * (uint32_t)(STAGE_BYTES >> 4), b4 = wb | bhi;
          if (nst == (uint32_t)SPCT) {
            umma2_f16(d_tmem, umma_desc_lo(a4), umma_desc_lo(b4), idesc, acc0);
#pragma unroll
            for (uint32_t i = 1; i < SPCT; ++i) umma2_f16(d_tmem, umma_desc_lo(a4 + i * kstep4), umma_desc_lo(b4 + i * bstep4), idesc, 1u);
          } else {
            for (uint32_t i = 0; i < nst; ++i) umma2_f16(d_tmem, umma_desc_lo(a4 + i * kstep4), umma_desc_lo(b4 + i * bstep4), idesc, i ? 1u : acc0);
          }
          if (last) umma2_f16(d_tmem, ones_desc, umma_desc_lo(wb + nst * bstep4), idesc, 1u);
          umma2_commit_mc(bar_wempty + stage * 8u);
          stage = nstage; phase = nphase;
        };
        const uint32_t sbytes = bstep4 << 4;                                   // bytes of one K-step of this CTA's half image
        const uint32_t spc = NF_WIDE_CHUNKS && sbytes * 16u <= (uint32_t)SPCT * 4096u ? 16u : NF_WIDE_CHUNKS && sbytes * 8u <= (uint32_t)SPCT * 4096u ? 8u : (uint32_t)SPCT;   // as the producers
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < k0s; c0 += spc)
          chunk(x4 + c0 * kstep4, k0s - c0 < spc ? k0s - c0 : spc, c0 ? 1u : 0u, hs == 0 && c0 + spc >= k0s);
        // the second half of the hidden operand (columns 128-255) is handed over separately: the first eight hidden K-steps run
        // while the epilogue warps still convert the second half (they read the whole accumulator into registers first)
        bool hi_ok = false;
        if (NF_ISSUE_STRAIGHT && SPCT == 4 && hs == 16u && bstep4 == 256u) {
          // the common shape (256 hidden inputs, 256 outputs: 10 of Plain+View's 12 Linears) as straight-line code: four full
          // chunks, every descriptor a constant offset from two bases -- the single issuing thread retires ~1 instruction per
          // 4-5 cycles, so instructions per chunk are what bounds the MMA issue rate
          auto chunk4 = [&](const uint32_t a4, const uint32_t acc0, const bool last) {
            if (!w_ok) { ST_ADD(2); mbar_wait(bar_wready + stage * 8u, phase); ST_ADD(1); ST_INC(3); }
            const uint32_t nstage = stage + 1 == NST ? 0u : stage + 1u, nphase = stage + 1 == NST ? phase ^ 1u : phase;
            w_ok = (last && mask != 3u) ? false : mbar_test_wait(bar_wready + nstage * 8u, nphase);
            tc_fence_after();
            const uint32_t wb = w4 + stage * (uint32_t)(STAGE_BYTES >> 4), b4 = wb | (128u << 16);     // LBO = 2048 B (128 rows x 16 B)
            umma2_f16(d_tmem, umma_desc_lo(a4), umma_desc_lo(b4), idesc, acc0);
#pragma unroll
            for (uint32_t i = 1; i < 4; ++i) umma2_f16(d_tmem, umma_desc_lo(a4 + i * kstep4), umma_desc_lo(b4 + i * 256u), idesc, 1u);
            if (last) umma2_f16(d_tmem, ones_desc, umma_desc_lo(wb + 4u * 256u), idesc, 1u);
            umma2_commit_mc(bar_wempty + stage * 8u);
            stage = nstage; phase = nphase;
          };
          chunk4(h4, k0s ? 1u : 0u, false);
          chunk4(h4 + 4u * kstep4, 1u, false);
          ST_ADD(2);
          mbar_wait(bar_a + (2u + slot) * 8u, (a_par >> (2u + slot)) & 1u); a_par ^= 1u << (2u + slot);
          ST_ADD(0);
          tc_fence_after();
          chunk4(h4 + 8u * kstep4, 1u, false);
          chunk4(h4 + 12u * kstep4, 1u, true);
        } else if (NF_ISSUE_STRAIGHT && SPCT == 4 && hs == 16u && spc == 4u) {
          // the same for the narrower Linears (density-out, the head's last Linear): their MMAs are issue-bound anyway
          chunk(h4, 4u, k0s ? 1u : 0u, false);
          chunk(h4 + 4u * kstep4, 4u, 1u, false);
          ST_ADD(2);
          mbar_wait(bar_a + (2u + slot) * 8u, (a_par >> (2u + slot)) & 1u); a_par ^= 1u << (2u + slot);
          ST_ADD(0);
          tc_fence_after();
          chunk(h4 + 8u * kstep4, 4u, 1u, false);
          chunk(h4 + 12u * kstep4, 4u, 1u, true);
        } else if (NF_ISSUE_STRAIGHT && SPCT == 3 && hs == 16u && spc == 3u) {
          // the shared-wide-x0 ring (3 x 14 KB): 3 + 3 + 3 + 3 + 3 + 1 K-steps
          chunk(h4, 3u, k0s ? 1u : 0u, false);
          chunk(h4 + 3u * kstep4, 3u, 1u, false);
          ST_ADD(2);
          mbar_wait(bar_a + (2u + slot) * 8u, (a_par >> (2u + slot)) & 1u); a_par ^= 1u << (2u + slot);
          ST_ADD(0);
          tc_fence_after();
          chunk(h4 + 6u * kstep4, 3u, 1u, false);
          chunk(h4 + 9u * kstep4, 3u, 1u, false);
          chunk(h4 + 12u * kstep4, 3u, 1u, false);
          chunk(h4 + 15u * kstep4, 1u, 1u, true);
        } else
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < hs || !hi_ok; c0 += spc) {
          if (!hi_ok && (c0 + spc > 8u || c0 >= hs)) {
            ST_ADD(2);
            mbar_wait(bar_a + (2u + slot) * 8u, (a_par >> (2u + slot)) & 1u); a_par ^= 1u << (2u + slot); hi_ok = true;
            ST_ADD(0);
            tc_fence_after();
          }
          if (c0 < hs) chunk(h4 + c0 * kstep4, hs - c0 < spc ? hs - c0 : spc, (k0s || c0) ? 1u : 0u, c0 + spc >= hs);
        }
        umma2_commit_mc((BW && li + 1 == n ? bar_bnd : bar_acc) + slot * 8u);      // BW: the path's last Linear reports to the boundary warps
      }
    }
    ST_FLUSH(0, blockIdx.x == 0 && (mask & 1u));
  };

  if (warp >= EPIW && warp <= EPIW + 3) {
  // BW: 24 warps start at 80 registers (768 x 80 = the CTA's pool); the producer / issuer warp group gives 32 of them to the epilogue warp groups
  if (BW) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(NF_REG_W4));          // register budget: see the epilogue branch
  if (warp < EPIW + NST) {
    // ================= weight producers (both CTAs): one ring stage each =================
    // The ring carries, step by step, slot 0's Linear then slot 1's (half a round behind); entry g goes to stage g % NST.
    if (elect_one()) {
      const int p = warp - EPIW;
      const uint32_t ready_leader = leader_addr(smem_u32(&s.w_ready[p]));
      const uint32_t bar_empty = smem_u32(&s.w_empty[p]), bar_land = smem_u32(&s.w_land[p]), dst = smem_u32(s.W + RING_OFF + p * STAGE_BYTES);
      uint32_t rs = 0, use = 0; int li0 = 0, li1 = 0;            // rs: ring stage of the current Linear's first chunk
      ST_DECL;
      for (int k = 0; k < nsteps + lag; ++k) {
#pragma unroll 1
        for (int slot = 0; slot < 2; ++slot) {
          const int kl = k - (slot ? lag : 0);
          if (kl < 0 || kl >= nsteps || (slot && single)) continue;
          const int li = slot ? li1 : li0;
          // chunks never straddle the x0 part and the hidden part of a Linear (the issuer's chunks then read ONE buffer each):
          // n0 = ceil(k0 / SPCT) chunks of x0 K-steps, then n1 = ceil(h / SPCT) chunks of hidden K-steps; the Linear's last chunk
          // also carries the bias K-group (it follows the data steps in the image).  This thread's chunks of the Linear are
          // first, first + NST, ... (no walk over the other stages' chunks: the bookkeeping between two copies sits between
          // "stage empty" and the next copy's issue whenever the issuer is ahead)
          const uint32_t k0s = prog.lin[li].k0_steps, hs = prog.lin[li].h_steps, sb = prog.lin[li].step_bytes;
          const uint8_t* src = a.packed + prog.lin[li].w_off + (size_t)crank * prog.lin[li].half_bytes;
          // K-steps per chunk: SPCT at 4 KB per step (N = 256); a narrow Linear (density-out: 1.25 KB per step, the head's last Linear:
          // 256 B) packs 8 or 16 steps into a stage -- fewer commits and probes for the issuer, whose MMAs are issue-bound there anyway
          const uint32_t spc = NF_WIDE_CHUNKS && sb * 16u <= (uint32_t)SPCT * 4096u ? 16u : NF_WIDE_CHUNKS && sb * 8u <= (uint32_t)SPCT * 4096u ? 8u : (uint32_t)SPCT;
          const uint32_t n0 = (k0s + spc - 1) / spc, nch = n0 + (hs + spc - 1) / spc;
#pragma unroll 1
          for (uint32_t c = (uint32_t)p >= rs ? (uint32_t)p - rs : (uint32_t)p + NST - rs; c < nch; c += NST) {
            const bool hid = c >= n0;
            const uint32_t st0 = (hid ? c - n0 : c) * spc, steps = hid ? hs : k0s, base = hid ? k0s : 0u;
            const uint32_t nst = steps - st0 < spc ? steps - st0 : spc;
            const uint32_t bytes = nst * sb + (c + 1 == nch ? sb >> 1 : 0u);
            const uint32_t par = use & 1u; ++use;
            ST_ADD(2);
            mbar_wait(bar_empty, par ^ 1u);
            ST_ADD(0);
            mbar_expect_tx(bar_land, bytes);
            bulk_g2s(dst, src + (size_t)(base + st0) * sb, bytes, bar_land);
            mbar_wait(bar_land, par);                                    // landed in THIS CTA ...
            ST_ADD(1);
            mbar_arrive_cluster_relaxed(ready_leader);                   // ... tell the leader's MMA thread
            ST_INC(3);
          }
          rs = (rs + nch) % NST;
          const int ln = li + 1 == n ? 0 : li + 1;
          if (slot) li1 = ln; else li0 = ln;
        }
      }
      ST_FLUSH(8, blockIdx.x == 0 && p == 0);
    }
  } else if (warp == EPIW + NST) {
    // ================= MMA issuer (leader CTA only) =================
    if (crank == 0 && elect_one()) run_issuer();
  }                                          // (with NST = 2 the group's fourth warp idles)
  } else if (!BW || warp < EPIW) {
    // ================= encode + epilogue: all 16 warps serve the two slots alternately =================
    // The pool is what the CTA was launched with: 24 warps x 80 >= 16 x 88 (epilogue) + 4 x 40 (producers, issuer) + 4 x 80 (boundary).
    if (BW) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(NF_REG_E));
    // TMEM lane quarter q = warp % 4 (rows 32q..32q+31), column quarter cq = warp / 4 (64 accumulator columns).
    const int q = warp & 3, cq = warp >> 2;
    const int e_tid = warp * 32 + lane;
    const int row = q * 32 + lane;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    uint32_t acc_par = 0;
    int j0 = 0, j1 = 0, P0 = 0, P1 = 0;                 // per slot: phase within the cycle, pass (tile) index
    ST_DECL;
    for (int k = 0; k <= nsteps + lag; ++k) {
#pragma unroll 1
      for (int slot = 0; slot < 2; ++slot) {
        const int kl = k - (slot ? lag : 0);
        if (kl < 0 || kl > nsteps || (slot && single)) continue;
        const int j = slot ? j1 : j0, P = slot ? P1 : P0;
        if (BW && !WB && j == 0) { if (slot) j1 = 1; else j0 = 1; continue; }      // the tile boundary belongs to the boundary warps
        uint8_t* H = s.H[slot]; uint8_t* X0 = single ? s.H[1] : shared_x0 ? s.X0[0] : s.X0[slot];
        const uint32_t t_acc = t_lane + (uint32_t)slot * 256u;
        const uint32_t a_ready_leader = leader_addr(smem_u32(&s.a_ready[slot])), a_hi_leader = leader_addr(smem_u32(&s.a_ready[2 + slot]));
        const bool has_next = kl < nsteps;              // Linear j of this slot runs after this phase
        const int comp_cq = (NCQ / 2) * slot, tail_cq = comp_cq + 1;
        // bias of Linear j (consumed by this slot's NEXT phase): in flight across the acc_full wait
        const float* bias = nullptr; (void)bias;
        if (kl > 0 && !(WB && j == 0)) {              // (WB: the last Linear reports to the boundary warps)
          ST_ADD(5);
          wait_acc(smem_u32(&s.acc_full[slot]), (acc_par >> slot) & 1u, a.debug);
          ST_ADD(0);
          acc_par ^= 1u << slot;
          tc_fence_after();
        }

        if (j == 0) {
          // ---------- tile boundary: composite of pass P-1 (cq == 0 warps) while the others start encoding pass P ----------
          // Duties that only one warp per lane quarter can do rotate with the slot, so that no column quarter becomes the
          // critical path (measured: the composite is ~4.9 K cycles, the View-x0 unit ~1.5 K, the x0 tail ~1.4 K):
          //   composite: cq == 2 * slot;  x0 tail / View-x0 unit: cq == 2 * slot + 1
          const bool comp = P >= 1 && !WB;            // WB: the boundary warps composite
          if (WB && has_next && e_tid == 0) {
            // the tile's Mip block: scratch -> the density x0's columns [mip0, mip0 + 96) (12 contiguous K-groups)
            mbar_wait(smem_u32(&s.scr_ready[slot][P & 1]), (uint32_t)(P >> 1) & 1u);
            mbar_expect_tx(smem_u32(&s.mip_land[slot]), MIP_BLOCK);
            bulk_g2s(smem_u32(X0 + (nf_mip_col(plan, 0) >> 3) * KG_BYTES), a.scratch + ((size_t)(blockIdx.x * 2 + slot) * 2 + (P & 1)) * MIP_BLOCK, MIP_BLOCK,
                     smem_u32(&s.mip_land[slot]));
          }
          if (comp && cq == comp_cq) {
            long long u; int sub; unit_of(P - 1, slot, map.tpr, nslot, u, sub);
            uint32_t v[16];
            tmem_ld16(t_acc, v); tmem_ld_wait(); reg_fence16(v);
            tc_fence_before();
            float cr, cg, cb;
            if (plan.kind == NF_KIND_TINY) {
              s.sig[slot][0][row] = __uint_as_float(v[0]);
              cr = __uint_as_float(v[1]); cg = __uint_as_float(v[2]); cb = __uint_as_float(v[3]);
            } else {
              cr = __uint_as_float(v[0]); cg = __uint_as_float(v[1]); cb = __uint_as_float(v[2]);
            }
            if (TRAIN) {
              long long ray; int t;
              if (map.locate(u, sub, row, a.n_rays, ray, t)) { float* o = tr.rgbraw_out + (ray * a.T + t) * 3; o[0] = cr; o[1] = cg; o[2] = cb; }
            }
            nf_feat_act3(cr, cg, cb, plan.feat_act);
            ST_ADD(6);
            composite_tile3(s, slot, plan, a, map, u, sub, row, lane, q, cr, cg, cb, s.sig[slot][0], TRAIN ? tr.sigma_out : nullptr);
            ST_ADD(5);
          }
          if (has_next) {
            long long u; int sub; unit_of(P, slot, map.tpr, nslot, u, sub);
            long long ray; int t;
            const bool ok = map.locate(u, sub, row, a.n_rays, ray, t);
            float px = 0.f, py = 0.f, pz = 0.f;
            if (ok) {
              const float* rr = a.rays + ray * 6;
              const float tt = __ldg(a.ts + ray * a.ts_stride + t);
              px = nf_pt(__ldg(rr + 0), tt, __ldg(rr + 3)); py = nf_pt(__ldg(rr + 1), tt, __ldg(rr + 4)); pz = nf_pt(__ldg(rr + 2), tt, __ldg(rr + 5));
              if (AUX && a.pts) { const float* pp = a.pts + (ray * a.T + t) * 3; px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2); }
            }
            // x0 of the FIRST MLP of the path: the density MLP, or (NF_KIND_DYN) the deformation MLP.  4 threads per row share
            // the hash levels; when the cq == 0 warps are compositing, the other three take them all
            const bool dyn = DYN && plan.kind == NF_KIND_DYN;
            const int first_m = dyn ? 2 : 0;
            const bool hashed = dyn ? plan.deform_enc == NF_ENC_HASH : plan.enc == NF_ENC_HASH;
            if (hashed)
              hash_x0(X0, reinterpret_cast<const float4*>(a.packed + (dyn ? plan.hash2_off : plan.hash_off)), plan, px, py, pz, row,
                      comp ? (cq == comp_cq ? -1 : ((cq - comp_cq - 1 + NCQ) % NCQ)) : cq, comp ? NCQ - 1 : NCQ);
            const int mip0 = (WIDE == 2 && plan.mip != NF_MIP_NONE && !dyn) ? nf_mip_col(plan, 0) : -1;
            if (mip0 >= 0 && !WB) mip_x0(X0, a.mip, mip0, ok, ray, t, row, comp ? (cq == comp_cq ? NF_MIP_FEATS : ((cq - comp_cq - 1 + NCQ) % NCQ)) : cq, comp ? NCQ - 1 : NCQ);
            const bool fourier = WIDE && !dyn && plan.enc == NF_ENC_FOURIER;
            if (fourier) {
              // x0 = [p, sin(p B), cos(p B)], B = basis[3][F] (reference src/neural_blocks.py:36-55, src/utils.py:14-17); reference
              // column order, element-wise half stores (columns are not 8-aligned); the frequencies are shared like the hash levels
              const float* Bm = reinterpret_cast<const float*>(a.packed + plan.fourier_off);
              const int F = plan.fourier_freqs;
              const int first = comp ? (cq == comp_cq ? F : ((cq - comp_cq - 1 + NCQ) % NCQ)) : cq, stride = comp ? NCQ - 1 : NCQ;
              for (int f = first; f < F; f += stride) {
                const float m = fmaf(pz, __ldg(Bm + 2 * F + f), fmaf(py, __ldg(Bm + F + f), __fmul_rn(px, __ldg(Bm + f))));
                const int c0 = 3 + f, c1 = 3 + F + f;
                *reinterpret_cast<__half*>(X0 + (c0 >> 3) * KG_BYTES + row * 16 + (c0 & 7) * 2) = __float2half_rn(ok ? sinf(m) : 0.f);
                *reinterpret_cast<__half*>(X0 + (c1 >> 3) * KG_BYTES + row * 16 + (c1 & 7) * 2) = __float2half_rn(ok ? cosf(m) : 0.f);
              }
              if (cq == tail_cq) {
                const float pv[3] = {px, py, pz};
                for (int c = 0; c < 3; ++c) *reinterpret_cast<__half*>(X0 + row * 16 + c * 2) = __float2half_rn(pv[c]);
                for (int c = 3 + 2 * F; c < plan.mlp[0].k0_pad; ++c) *reinterpret_cast<__half*>(X0 + (c >> 3) * KG_BYTES + row * 16 + (c & 7) * 2) = __float2half_rn(0.f);
              }
            } else if (cq == tail_cq) {
              if (hashed) hash_x0_tail(X0, plan, plan.mlp[first_m].k0_pad, px, py, pz, row, mip0);
              else {
                const float tt = (dyn && ok) ? __ldg(a.ray_time + ray) : 0.f;        // direct deformation: x0 = [p, t]
                st_v4(X0 + row * 16, pack_h2(px, py), pack_h2(pz, tt), 0, 0);
                for (int g = 1; g < (plan.mlp[first_m].k0_pad >> 3); ++g) st_v4(X0 + g * KG_BYTES + row * 16, 0, 0, 0, 0);
              }
            }
            if (WB) {
              mbar_wait_suspend(smem_u32(&s.mip_land[slot]), 0u);                                  // the Mip block has landed in x0 ...
              if (P >= 1) mbar_wait_suspend(smem_u32(&s.col_read[slot]), (uint32_t)(P - 1) & 1u);  // ... and the previous tile's colours are out of TMEM
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive_cluster_relaxed(a_ready_leader); mbar_arrive_cluster_relaxed(a_hi_leader); }
          }
          ST_ADD(1);
        } else {
          // ---------- epilogue of Linear j-1 ----------
          if (BW && j - 1 == prog.x0_last && lane == 0) mbar_arrive(smem_u32(&s.x0_free[slot]));   // X0[slot] may take the next tile's x0
          if (!BW && j == n - 2 && cq == 1 && P + 1 < passes) {
            // the next tile's rays are a first touch (HBM, ~2 K cycles): pull them into L2 two phases before phase 0 needs them
            long long u; int sub; unit_of(P + 1, slot, map.tpr, nslot, u, sub);
            long long ray; int t;
            if (map.locate(u, sub, row, a.n_rays, ray, t)) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(a.rays + ray * 6));
              if (a.ts_stride) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ts + ray * a.ts_stride + t));
              if (a.noise) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.noise + ray * a.T + t));
            }
          }
          const int4 Lc = s.lin[j - 1][0];
          const int act = Lc.z;
          const bool is_out = (Lc.w & 1) != 0;
          bool lo_done = false;                 // this phase has already handed over the first half of the operand
          if (!is_out) {
            if (TRAIN) {
              // this phase writes the hidden input of Linear j (and, after an `init`, activates x0): stash all of it
              long long u; int sub; unit_of(P, slot, map.tpr, nslot, u, sub);
              const long long g = u * map.tpr + sub;
              const bool st_ok = g < tr.n_tiles;
              const Tc3TrainLin& Tn = tr.lin[j];
              uint8_t* tileA = a.ws + ((size_t)Tn.a_off256 + (size_t)g * Tn.a_tile256) * 256;
              uint8_t* gA = st_ok ? tileA + (size_t)Tn.k0 * 256 : nullptr;
              uint8_t* gC = (st_ok && Tn.c_off256) ? a.ws + ((size_t)Tn.c_off256 + (size_t)g * 256) * 256 : nullptr;
              if (Lc.w & 2) {
                const Tc3TrainLin& Ti = tr.lin[j - 1];
                uint8_t* gRaw = st_ok ? a.ws + ((size_t)Ti.a_off256 + (size_t)g * Ti.a_tile256) * 256 : nullptr;
                uint8_t* g0 = (st_ok && Ti.skip0 >= 0) ? a.ws + ((size_t)tr.lin[Ti.skip0].a_off256 + (size_t)g * tr.lin[Ti.skip0].a_tile256) * 256 : nullptr;
                uint8_t* g1 = (st_ok && Ti.skip1 >= 0) ? a.ws + ((size_t)tr.lin[Ti.skip1].a_off256 + (size_t)g * tr.lin[Ti.skip1].a_tile256) * 256 : nullptr;
                x0_activate3(X0, s.lin[j - 1][1].x, act, e_tid, EPI_THREADS, gRaw, g0, g1);
              }
              if (NF_SPLIT_HANDOFF && NCQ == 4 && act == NF_ACT_SIN) { epi_hidden_split<NF_ACT_SIN, true>(H, t_acc, cq, row, lane, a_ready_leader, gA, gC); lo_done = true; }
              else if (NF_SPLIT_HANDOFF && NCQ == 4 && act == NF_ACT_LEAKY) { epi_hidden_split<NF_ACT_LEAKY, true>(H, t_acc, cq, row, lane, a_ready_leader, gA, nullptr); lo_done = true; }
              else if (act == NF_ACT_SIN) epi_hidden3<NF_ACT_SIN, NCQ, true>(H, t_acc, bias, cq, row, gA, gC);
              else if (act == NF_ACT_LEAKY) epi_hidden3<NF_ACT_LEAKY, NCQ, true>(H, t_acc, bias, cq, row, gA, nullptr);
              else if (act == NF_ACT_RELU) epi_hidden3<NF_ACT_RELU, NCQ, true>(H, t_acc, bias, cq, row, gA, nullptr);
              else epi_hidden3<NF_ACT_NONE, NCQ, true>(H, t_acc, bias, cq, row, gA, nullptr);
            } else {
            if (Lc.w & 2) x0_activate3(X0, s.lin[j - 1][1].x, act, e_tid, EPI_THREADS);       // init consumed raw x0; the skip Linear wants act(x0)
            if (NF_SPLIT_HANDOFF && NCQ == 4 && act == NF_ACT_SIN) { epi_hidden_split<NF_ACT_SIN, false>(H, t_acc, cq, row, lane, a_ready_leader); lo_done = true; }
            else if (NF_SPLIT_HANDOFF && NCQ == 4 && act == NF_ACT_LEAKY) { epi_hidden_split<NF_ACT_LEAKY, false>(H, t_acc, cq, row, lane, a_ready_leader); lo_done = true; }
            else if (act == NF_ACT_SIN) { if (a.debug & 2048) epi_hidden3<NF_ACT_SIN, NCQ>(H, t_acc, bias, cq, row); else epi_hidden3_sin_pipelined<NCQ>(H, t_acc, bias, cq, row); }
            else if (act == NF_ACT_LEAKY) epi_hidden3<NF_ACT_LEAKY, NCQ>(H, t_acc, bias, cq, row);
            else if (act == NF_ACT_RELU) epi_hidden3<NF_ACT_RELU, NCQ>(H, t_acc, bias, cq, row);
            else epi_hidden3<NF_ACT_NONE, NCQ>(H, t_acc, bias, cq, row);
            }
          } else if (DYN && ((Lc.w >> 2) & 3) == 2) {
            // deformation MLP out (reference nerf.py:1261-1278): every thread deforms its row's sample, then takes its share of
            // the density MLP's hash levels at the DEFORMED position (the canonical NeRF sees pts + rigid_dp, nerf.py:1303)
            uint32_t v[32];
            tmem_ld16(t_acc, v);
            if (Lc.x > 16) tmem_ld16(t_acc + 16, v + 16);
            tmem_ld_wait(); reg_fence16(v); reg_fence16(v + 16);
            long long u; int sub; unit_of(P, slot, map.tpr, nslot, u, sub);
            long long ray; int t;
            float px = 0.f, py = 0.f, pz = 0.f;
            if (map.locate(u, sub, row, a.n_rays, ray, t)) {
              const float* rr = a.rays + ray * 6;
              const float tt = __ldg(a.ts + ray * a.ts_stride + t);
              px = nf_pt(__ldg(rr + 0), tt, __ldg(rr + 3)); py = nf_pt(__ldg(rr + 1), tt, __ldg(rr + 4)); pz = nf_pt(__ldg(rr + 2), tt, __ldg(rr + 5));
              const int nsp = plan.spline_points;
              const bool side = cq == tail_cq;                      // one warp per lane quarter writes the side channels
              const long long sidx = ray * a.T + t;
              if (side && a.pts_out) { float* o = a.pts_out + sidx * 3; o[0] = px; o[1] = py; o[2] = pz; }
              if (nsp == 0) {
                const float dp = __uint_as_float(v[0]);
                const float r0 = nf_sigmoid(__uint_as_float(v[1]) / 2.f), r1 = nf_sigmoid(__uint_as_float(v[2]) / 2.f),
                            r2 = nf_sigmoid(__uint_as_float(v[3]) / 2.f);
                px += dp * r0; py += dp * r1; pz += dp * r2;
                if (side) {
                  // the reference's split names the 1-channel output dp and the 3-channel one rigidity (nerf.py:1231,1261-1266)
                  if (a.dp_out) a.dp_out[sidx] = dp;
                  if (a.rigidity_out) { float* o = a.rigidity_out + sidx * 3; o[0] = r0; o[1] = r1; o[2] = r2; }
                  if (a.rigid_dp_out) { float* o = a.rigid_dp_out + sidx * 3; o[0] = dp * r0; o[1] = dp * r1; o[2] = dp * r2; }
                }
              } else {
                const float rig = nf_sigmoid(__uint_as_float(v[0]) / 2.f);
                const float time = __ldg(a.ray_time + ray);
                float d[3];
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                  float ps[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) ps[i] = i < nsp ? __uint_as_float(v[1 + 3 * i + x]) : 0.f;
                  d[x] = nf_bezier(ps, nsp, time);
                }
                if (side) {
                  if (a.dp_out) { float* o = a.dp_out + sidx * 3; o[0] = d[0]; o[1] = d[1]; o[2] = d[2]; }
                  if (a.rigidity_out) a.rigidity_out[sidx] = rig;
                  if (a.rigid_dp_out) { float* o = a.rigid_dp_out + sidx * 3; o[0] = d[0] * rig; o[1] = d[1] * rig; o[2] = d[2] * rig; }
                }
                px += d[0] * rig; py += d[1] * rig; pz += d[2] * rig;
              }
            }
            // the deformed point, for the View head's x0: parked as three floats in the last K-group of X0 (columns 72-79), which
            // neither the density MLP's x0 (k0_pad <= 64, checked on the host) nor anything else touches until the density-out epilogue
            if (cq == tail_cq) *reinterpret_cast<float4*>(X0 + (X0K / 8 - 1) * KG_BYTES + row * 16) = make_float4(px, py, pz, 0.f);
            hash_x0(X0, reinterpret_cast<const float4*>(a.packed + plan.hash_off), plan, px, py, pz, row, cq, NCQ);
            if (cq == tail_cq) hash_x0_tail(X0, plan, plan.mlp[0].k0_pad, px, py, pz, row);
          } else {
            // density MLP out (tensor order [inter(I), sigma]) -> raw x0 of the RGB head + raw density
            const int iu = plan.intermediate >> 4;
            const bool pos_head = WIDE && plan.refl_kind == NF_REFL_POSITIONAL;
            const bool pre_tail = BW && NF_BW_PRETAIL && !DYN && plan.kind == NF_KIND_PLAIN && plan.refl_kind == NF_REFL_VIEW && plan.mlp[0].k0_pad <= plan.intermediate;   // the boundary warps wrote [p, elaz]
            const bool pos_pre = BW && X0C == X0K_POS && pos_head;       // ... or the Positional head's [hash'(p), p, p]
            const int mip1 = (WIDE == 2 && plan.mip != NF_MIP_NONE) ? nf_mip_col(plan, 1) : -1;
            if (WB && mip1 >= 0 && e_tid == 0) {
              // the same Mip block again, now into the View x0's columns [mip1, mip1 + 96)
              mbar_expect_tx(smem_u32(&s.mip_land[slot]), MIP_BLOCK);
              bulk_g2s(smem_u32(X0 + (mip1 >> 3) * KG_BYTES), a.scratch + ((size_t)(blockIdx.x * 2 + slot) * 2 + (P & 1)) * MIP_BLOCK, MIP_BLOCK,
                       smem_u32(&s.mip_land[slot]));
            }
            if ((pos_head && !pos_pre) || mip1 >= 0) {
              // wide RGB-head inputs (single mode): every thread takes a share of its row's Positional hash features and Mip latent
              long long u; int sub; unit_of(P, slot, map.tpr, nslot, u, sub);
              long long ray; int t;
              const bool ok = map.locate(u, sub, row, a.n_rays, ray, t);
              if (pos_head) {
                float px = 0.f, py = 0.f, pz = 0.f;
                if (ok) {
                  const float* rr = a.rays + ray * 6;
                  const float tt = __ldg(a.ts + ray * a.ts_stride + t);
                  px = nf_pt(__ldg(rr + 0), tt, __ldg(rr + 3)); py = nf_pt(__ldg(rr + 1), tt, __ldg(rr + 4)); pz = nf_pt(__ldg(rr + 2), tt, __ldg(rr + 5));
                }
                uint8_t* Xh = X0 + (plan.intermediate >> 3) * KG_BYTES;                     // columns [I, I + 4L): the head's own hash features
                hash_x0(Xh, reinterpret_cast<const float4*>(a.packed + plan.hash3_off), plan, px, py, pz, row, cq, NCQ);
                if (cq == tail_cq) hash_x0_tail(Xh, plan, plan.mlp[1].k0_pad - plan.intermediate, px, py, pz, row, mip1 >= 0 ? mip1 - plan.intermediate : -1);
              }
              if (mip1 >= 0 && !WB) mip_x0(X0, a.mip, mip1, ok, ray, t, row, cq, NCQ);
            }
            for (int un0 = cq; un0 < iu + NCQ; un0 += NCQ) {
              // units 0..iu-1 (intermediate columns) are dealt round-robin; the last unit (sigma + View x0 tail) goes to tail_cq
              int un = un0;
              if (un0 >= iu) { if (cq != tail_cq) break; un = iu; }
              uint32_t v[16];
              tmem_ld16(t_acc + un * 16, v); tmem_ld_wait(); reg_fence16(v);
              if (un < iu) {
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  o[i] = pack_h2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                uint8_t* d0 = X0 + (un * 2) * KG_BYTES + row * 16;
                st_v4(d0, o[0], o[1], o[2], o[3]); st_v4(d0 + KG_BYTES, o[4], o[5], o[6], o[7]);
              } else {
                s.sig[slot][BW ? (P & 1) : 0][row] = __uint_as_float(v[0]);
                if (!pos_head && !pre_tail) {
                  long long u; int sub; unit_of(P, slot, map.tpr, nslot, u, sub);
                  long long ray; int t;
                  float px = 0.f, py = 0.f, pz = 0.f, el = 0.f, az = 0.f;
                  if (map.locate(u, sub, row, a.n_rays, ray, t)) {
                    const float* rr = a.rays + ray * 6;
                    const float tt = __ldg(a.ts + ray * a.ts_stride + t);
                    const float dx = __ldg(rr + 3), dy = __ldg(rr + 4), dz = __ldg(rr + 5);
                    if (DYN && plan.kind == NF_KIND_DYN) { const float4 pp = *reinterpret_cast<const float4*>(X0 + (X0K / 8 - 1) * KG_BYTES + row * 16); px = pp.x; py = pp.y; pz = pp.z; }
                    else { px = nf_pt(__ldg(rr + 0), tt, dx); py = nf_pt(__ldg(rr + 1), tt, dy); pz = nf_pt(__ldg(rr + 2), tt, dz); }
                    if (AUX && a.pts) { const float* pp = a.pts + (ray * a.T + t) * 3; px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2); }
                    nf_elaz(dx, dy, dz, el, az);
                  }
                  uint8_t* d0 = X0 + (iu * 2) * KG_BYTES + row * 16;
                  st_v4(d0, pack_h2(px, py), pack_h2(pz, el), pack_h2(az, 0.f), 0);
                  // columns I+8 .. : the Mip latent (written above), then zero padding up to k0_pad
                  const int m0 = mip1 >= 0 ? mip1 >> 3 : 1 << 20;
                  for (int g = iu * 2 + 1; g < (plan.mlp[1].k0_pad >> 3); ++g)
                    if (g < m0 || g >= m0 + NF_MIP_FEATS / 8) st_v4(X0 + g * KG_BYTES + row * 16, 0, 0, 0, 0);
                }
              }
            }
          }
          if (WB && is_out && ((Lc.w >> 2) & 3) == 1 && plan.mip != NF_MIP_NONE) mbar_wait_suspend(smem_u32(&s.mip_land[slot]), 1u);   // density-out: the View x0's Mip block has landed
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) { if (!lo_done) mbar_arrive_cluster_relaxed(a_ready_leader); mbar_arrive_cluster_relaxed(a_hi_leader); }
          if (is_out) ST_ADD(4); else if (act == NF_ACT_SIN) ST_ADD(3); else ST_ADD(2);
        }
        int jn = j + 1, Pn = P;
        if (jn == n) { jn = 0; ++Pn; }
        if (slot) { j1 = jn; P1 = Pn; } else { j0 = jn; P0 = Pn; }
      }
    }
    ST_FLUSH(16, blockIdx.x == 0 && warp == 0 && lane == 0);
    ST_FLUSH(24, blockIdx.x == 0 && warp == EPIW - 1 && lane == 0);
    ST_FLUSH(32, blockIdx.x == 1 && warp == 0 && lane == 0);
  } else {
    // ================= boundary warps (BW): one per TMEM lane quarter, thread = row =================
    // Per slot and tile boundary (tile P-1 -> tile P): [wait x0_free] encode tile P's x0 (+ the View head's [p, elaz] tail, whose
    // columns the density MLP does not touch) -> [wait bnd_full] read tile P-1's colours out of TMEM -> hand the slot to the
    // issuer -> composite tile P-1.  The two slots' boundaries alternate in the global step order (slot 1 lags by `lag`
    // Linears), so one set of warps serves both in turn; every wait is for an event that depends only on earlier boundaries.
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const bool dyn = DYN && plan.kind == NF_KIND_DYN;
    const int first_m = dyn ? 2 : 0;
    const bool hashed = dyn ? plan.deform_enc == NF_ENC_HASH : plan.enc == NF_ENC_HASH;
    const bool pre_tail = NF_BW_PRETAIL && !DYN && plan.kind == NF_KIND_PLAIN && plan.refl_kind == NF_REFL_VIEW && plan.mlp[0].k0_pad <= plan.intermediate;
    const bool pos_pre = X0C == X0K_POS && WIDE && plan.refl_kind == NF_REFL_POSITIONAL;
    ST_DECL;
    if (WB) {
      // ---- Mip with a shared x0 buffer: colour read + composite, and the NEXT tile's Mip block into the scratch ----
      auto mip_block = [&](int slot, int P) {
        long long u; int sub; unit_of(P, slot, map.tpr, 2, u, sub);
        long long ray; int t;
        const bool ok = map.locate(u, sub, row, a.n_rays, ray, t);
        uint8_t* scr = a.scratch + ((size_t)(blockIdx.x * 2 + slot) * 2 + (P & 1)) * MIP_BLOCK + row * 16;
        NfMipRow R;
        const bool per_row = a.mip.mode != NF_MIP_CYLINDER_REF;
        if (ok && per_row) nf_mip_row(a.mip, ray, t, R);
#pragma unroll 1
        for (int g = 0; g < NF_MIP_FEATS / 16; ++g) {                       // pairs 8g .. 8g+7: sines -> K-group g, cosines -> K-group g + 6
          float fs[8], fc[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            fs[k] = 0.f; fc[k] = 0.f;
            if (ok) { if (per_row) nf_mip_pair_of_row<true>(R, 8 * g + k, fs[k], fc[k]); else nf_mip_feature_pair<true>(a.mip, ray, t, 8 * g + k, fs[k], fc[k]); }
          }
          *reinterpret_cast<uint4*>(scr + g * KG_BYTES) = make_uint4(pack_h2(fs[0], fs[1]), pack_h2(fs[2], fs[3]), pack_h2(fs[4], fs[5]), pack_h2(fs[6], fs[7]));
          *reinterpret_cast<uint4*>(scr + (g + NF_MIP_FEATS / 16) * KG_BYTES) = make_uint4(pack_h2(fc[0], fc[1]), pack_h2(fc[2], fc[3]), pack_h2(fc[4], fc[5]), pack_h2(fc[6], fc[7]));
        }
        asm volatile("fence.proxy.async;" ::: "memory");                  // generic-proxy global writes -> the bulk copy (async proxy) that reads them
        __threadfence();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s.scr_ready[slot][P & 1]));
      };
      if (passes > 0) { mip_block(0, 0); mip_block(1, 0); }
      for (int P = 0; P <= passes; ++P) {
#pragma unroll 1
        for (int slot = 0; slot < 2; ++slot) {
          if (P >= 1) {
            mbar_wait_suspend(smem_u32(&s.bnd_full[slot]), (uint32_t)(P - 1) & 1u);
            tc_fence_after();
            uint32_t v[16];
            tmem_ld16(t_lane + (uint32_t)slot * 256u, v); tmem_ld_wait(); reg_fence16(v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s.col_read[slot]));
            float cr = __uint_as_float(v[0]), cg = __uint_as_float(v[1]), cb = __uint_as_float(v[2]);
            long long u; int sub; unit_of(P - 1, slot, map.tpr, 2, u, sub);
            nf_feat_act3(cr, cg, cb, plan.feat_act);
            composite_tile3(s, slot, plan, a, map, u, sub, row, lane, q, cr, cg, cb, s.sig[slot][(P - 1) & 1], nullptr, false);
          }
          if (P + 1 < passes) mip_block(slot, P + 1);
        }
      }
    } else
    for (int P = 0; P <= passes; ++P) {
#pragma unroll 1
      for (int slot = 0; slot < 2; ++slot) {
        const bool comp = P >= 1, has_next = P < passes;
        uint8_t* X0 = s.X0[slot];
        if (has_next) {
          long long u; int sub; unit_of(P, slot, map.tpr, 2, u, sub);
          long long ray; int t;
          const bool ok = map.locate(u, sub, row, a.n_rays, ray, t);
          float px = 0.f, py = 0.f, pz = 0.f, el = 0.f, az = 0.f, tt_dyn = 0.f;
          if (ok) {
            const float* rr = a.rays + ray * 6;
            const float tt = __ldg(a.ts + ray * a.ts_stride + t);
            const float dx = __ldg(rr + 3), dy = __ldg(rr + 4), dz = __ldg(rr + 5);
            px = nf_pt(__ldg(rr + 0), tt, dx); py = nf_pt(__ldg(rr + 1), tt, dy); pz = nf_pt(__ldg(rr + 2), tt, dz);
            if (AUX && a.pts) { const float* pp = a.pts + (ray * a.T + t) * 3; px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2); }
            if (pre_tail) nf_elaz(dx, dy, dz, el, az);
            if (dyn && !hashed) tt_dyn = __ldg(a.ray_time + ray);
          }
          ST_ADD(1);
          if (P >= 1) mbar_wait_suspend(smem_u32(&s.x0_free[slot]), (uint32_t)(P - 1) & 1u);
          ST_ADD(0);
          if (hashed) {
            hash_x0(X0, reinterpret_cast<const float4*>(a.packed + (dyn ? plan.hash2_off : plan.hash_off)), plan, px, py, pz, row, 0, 1);
            hash_x0_tail(X0, plan, plan.mlp[first_m].k0_pad, px, py, pz, row);
          } else {
            st_v4(X0 + row * 16, pack_h2(px, py), pack_h2(pz, tt_dyn), 0, 0);                 // direct deformation: x0 = [p, t]; else [p]
            for (int g = 1; g < (plan.mlp[first_m].k0_pad >> 3); ++g) st_v4(X0 + g * KG_BYTES + row * 16, 0, 0, 0, 0);
          }
          if (pre_tail) {
            const int g0 = plan.intermediate >> 3;
            st_v4(X0 + g0 * KG_BYTES + row * 16, pack_h2(px, py), pack_h2(pz, el), pack_h2(az, 0.f), 0);
            for (int g = g0 + 1; g < (plan.mlp[1].k0_pad >> 3); ++g) st_v4(X0 + g * KG_BYTES + row * 16, 0, 0, 0, 0);
          }
          if (WIDE && pos_pre) {
            // the Positional head's own encoder (refl.py:233-237): columns [I, I + 4L) hash'(p), then [p, p] and the zero padding
            uint8_t* Xh = X0 + (plan.intermediate >> 3) * KG_BYTES;
            hash_x0(Xh, reinterpret_cast<const float4*>(a.packed + plan.hash3_off), plan, px, py, pz, row, 0, 1);
            hash_x0_tail(Xh, plan, plan.mlp[1].k0_pad - plan.intermediate, px, py, pz, row, -1);
          }
          fence_proxy_async();
          ST_ADD(1);
        }
        float cr = 0.f, cg = 0.f, cb = 0.f;
        if (comp) {
          mbar_wait_suspend(smem_u32(&s.bnd_full[slot]), (uint32_t)(P - 1) & 1u);
          ST_ADD(2);
          tc_fence_after();
          uint32_t v[16];
          tmem_ld16(t_lane + (uint32_t)slot * 256u, v); tmem_ld_wait(); reg_fence16(v);
          tc_fence_before();
          if (plan.kind == NF_KIND_TINY) {
            s.sig[slot][(P - 1) & 1][row] = __uint_as_float(v[0]);
            cr = __uint_as_float(v[1]); cg = __uint_as_float(v[2]); cb = __uint_as_float(v[3]);
          } else {
            cr = __uint_as_float(v[0]); cg = __uint_as_float(v[1]); cb = __uint_as_float(v[2]);
          }
        }
        if (has_next) {
          __syncwarp();
          if (lane < 4) {                                                                           // 4 warps x 4 lanes = the 16 arrivals of a phase
            mbar_arrive_cluster_relaxed(leader_addr(smem_u32(&s.a_ready[slot]))); mbar_arrive_cluster_relaxed(leader_addr(smem_u32(&s.a_ready[2 + slot])));
          }
        }
        ST_ADD(3);
        if (comp) {
          long long u; int sub; unit_of(P - 1, slot, map.tpr, 2, u, sub);
          if (TRAIN) {
            long long ray; int t;
            if (map.locate(u, sub, row, a.n_rays, ray, t)) { float* o = tr.rgbraw_out + (ray * a.T + t) * 3; o[0] = cr; o[1] = cg; o[2] = cb; }
          }
          nf_feat_act3(cr, cg, cb, plan.feat_act);
          composite_tile3(s, slot, plan, a, map, u, sub, row, lane, q, cr, cg, cb, s.sig[slot][(P - 1) & 1], TRAIN ? tr.sigma_out : nullptr, false);
          ST_ADD(4);
        }
      }
    }
    ST_FLUSH(40, blockIdx.x == 0 && q == 0 && lane == 0);
  }
  // ---- teardown ----
#ifdef NF_TC_STATS
  // (the epilogue warps' counters live in their branch scope; they are flushed there)
#endif
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == EPIW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(0), "r"(512) : "memory");
  }
}

bool build_prog3(const NfPlan& plan, Tc3Prog* P, int x0_cap = X0K) {
  *P = Tc3Prog{};
  int nl = 0;
  for (int mi = 0; mi < plan.n_mlps; ++mi) {
    const int m = plan.kind == NF_KIND_DYN ? (mi + 2) % 3 : mi;          // execution order: deformation, density, View
    for (int j = 0; j < plan.mlp[m].n_lin; ++j, ++nl) {
      const NfLinPlan& L = plan.mlp[m].lin[j];
      const uint32_t nh = (uint32_t)L.n_pad >> 1, b_lbo = nh * 16u;
      Tc3Lin& R = P->lin[nl];
      R.k0_steps = (uint32_t)L.k0_pad >> 4; R.h_steps = (uint32_t)L.k_hidden >> 4;
      R.idesc = (1u << 4) | ((uint32_t)(L.n_pad >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // M = 256 across the pair
      R.bstep4 = (2u * b_lbo) >> 4; R.bhi = (b_lbo >> 4) << 16;
      R.mj = (uint32_t)(m * 16 + j);
      const int64_t half_bytes = (int64_t)(L.k0_pad + L.k_hidden + 16) * nh * 2;     // + the bias K-step
      if (L.w16h_off + 2 * half_bytes >= (1LL << 32)) return false;
      R.w_off = (uint32_t)L.w16h_off; R.half_bytes = (uint32_t)half_bytes; R.step_bytes = 2u * b_lbo;
    }
  }
  P->n_lin = nl; P->lag = nl / 2;
#ifdef NF_LAG
  P->lag = NF_LAG < nl ? NF_LAG : nl / 2;       // experiment: how far slot 1 runs behind slot 0
#endif
  P->x0_last = 0;
  for (int i = 0; i < nl; ++i) if (P->lin[i].k0_steps) P->x0_last = i;
  int kmax = 0;
  for (int m = 0; m < plan.n_mlps; ++m) kmax = plan.mlp[m].k0_pad > kmax ? plan.mlp[m].k0_pad : kmax;
  if (kmax > x0_cap) {
    P->single = 1;
    if (kmax <= SHARED_X0_COLS && plan.kind == NF_KIND_PLAIN) {
      // shared wide x0: writers = the phases that produce / activate an MLP's x0, consumers = the Linears that read it.  Slot 1
      // runs `lag` phases behind slot 0 and is served second within a step: when slot 0 writes at phase s, slot 1's Linear
      // s - lag - 1 may still be in flight; when slot 1 writes at phase s, slot 0's Linear s + lag may be.
      bool writer[MAX_LIN3] = {}, consumer[MAX_LIN3] = {};
      int li = 0;
      for (int mi = 0; mi < plan.n_mlps; ++mi)
        for (int j = 0; j < plan.mlp[mi].n_lin; ++j, ++li)
          if (plan.mlp[mi].lin[j].k0_pad) { consumer[li] = true; writer[li] = true; if (j == 0) writer[(li + 1) % nl] = true; }
      const int lag = 3;
      bool ok = nl > 2 * lag + 2;
      for (int sph = 0; sph < nl && ok; ++sph)
        if (writer[sph] && (consumer[((sph - lag - 1) % nl + nl) % nl] || consumer[(sph + lag) % nl] || consumer[((sph - lag) % nl + nl) % nl])) ok = false;
      if (ok) { P->single = 2; P->lag = lag; }
    }
  }
  return true;
}

}  // namespace

// nullptr if the staggered paired pipeline can run this model, else the reason.
const char* nf_tc3_unsupported(const NfPlan& p) {
  if (p.refl_kind == NF_REFL_POSLINVIEW) return "PosLinearView runs on the fp32 pipeline only (three-MLP head, hidden 128)";
  const bool wide = p.mip != NF_MIP_NONE || p.refl_kind != NF_REFL_VIEW;
  if (wide && p.kind != NF_KIND_PLAIN) return "Mip / Positional on the tensor pipeline: PlainNeRF only (DynamicNeRF runs them on the fp32 pipeline)";
  if (wide && p.enc != NF_ENC_HASH) return "Mip / Positional on the tensor pipeline need the hash-encoded density MLP";
  const bool fourier = p.enc == NF_ENC_FOURIER && p.kind == NF_KIND_PLAIN;        // the Fourier-encoded SDF MLP of VolSDF: x0 259 -> 272
  if (p.kind == NF_KIND_DYN && p.enc != NF_ENC_HASH) return "NF_KIND_DYN: the canonical NeRF must be hash-encoded";
  if (p.kind == NF_KIND_DYN && p.mlp[0].k0_pad > 64) return "NF_KIND_DYN: density x0 wider than 64 columns (the deformed point is parked in X0 columns 72-79)";
  if (p.kind == NF_KIND_DYN && p.mlp[2].lin[p.mlp[2].n_lin - 1].n_pad > 32) return "deformation MLP with more than 32 outputs";
  int nlin = 0;
  for (int m = 0; m < p.n_mlps; ++m) {
    nlin += p.mlp[m].n_lin;
    if (p.mlp[m].k0_pad > (fourier ? 400 : 256)) return "x0 too wide for the single-tile mode";
    if (p.mlp[m].k0_pad > X0K && !wide && !fourier) return "x0 wider than 80 columns";
    for (int j = 0; j < p.mlp[m].n_lin; ++j) {
      // only the density MLP of a two-MLP model may end in a non-final `out` Linear
      if (p.mlp[m].lin[j].is_out && p.kind == NF_KIND_TINY && m != 0) return "unsupported MLP chain";
    }
  }
  if (nlin > MAX_LIN3 || nlin < 2) return "unsupported number of Linear layers";
  if (p.enc == NF_ENC_FOURIER && !fourier) return "Fourier-encoded density MLP: PlainNeRF / VolSDF kind only";
  if (p.enc == NF_ENC_HASH && (p.hash_levels & 1)) return "odd number of hash levels";
  if (p.kind == NF_KIND_PLAIN && (p.intermediate & 15)) return "intermediate_size not a multiple of 16";
  return nullptr;
}

// nullptr if the training forward/backward (activation stash + nf_train.cu) can run this model, else the reason.
const char* nf_train_unsupported(const NfPlan& p) {
  if (const char* why = nf_tc3_unsupported(p)) return why;
  if (p.kind == NF_KIND_DYN) return "training: DynamicNeRF needs the gradient with respect to the sample position (not built)";
  if (p.mip != NF_MIP_NONE || p.enc == NF_ENC_FOURIER) return "training: wide-x0 models (Mip, Fourier SDF) are not built";
  if (p.refl_kind != NF_REFL_VIEW && p.refl_kind != NF_REFL_POSITIONAL) return "training: View and Positional heads only";
  // the Positional head (refl.py:230-245): on its boundary-warp instantiation (112-column x0 per slot; T % 32 == 0, checked at launch)
  if (p.refl_kind == NF_REFL_POSITIONAL && (p.enc != NF_ENC_HASH || p.hash_levels * 4 != 32 || p.mlp[0].k0_pad > p.intermediate || p.mlp[1].k0_pad > X0K_POS))
    return "training: Positional head needs the hash-encoded density MLP (8 levels x 4) and an x0 of at most 112 columns";
  // (VolSDF's SIREN SDF, x0 = [p], trains here: weights and beta; its eikonal regulariser (runner.py:736) needs d sdf / d p: nf_sdf_normals)
  for (int m = 0; m < p.n_mlps; ++m) if (p.mlp[m].act != NF_ACT_LEAKY && p.mlp[m].act != NF_ACT_SIN) return "training: LeakyReLU / sin MLPs only";
  return nullptr;
}

cudaError_t nf_launch_render_tc3(const NfPlan& plan, const void* packed, const float* rays, int64_t n_rays, const float* ts,
                                 int T, int64_t ts_stride, const float* noise, const float* ray_time, const nf_mip_args* mip,
                                 float* rgb, float* alpha, float* weights, cudaStream_t st, const NfTrainPlan* tp, void* ws,
                                 const nf_render_aux* aux) {
  if (nf_tc3_unsupported(plan)) return cudaErrorNotSupported;
  if (tp && nf_train_unsupported(plan)) return cudaErrorNotSupported;
  Tc3Args a{};
  a.ws = (uint8_t*)ws;
  if (aux) {
    a.pts = aux->pts; a.bg_rand = aux->bg_rand;
    a.pts_out = aux->pts_out; a.dp_out = aux->dp_out; a.rigid_dp_out = aux->rigid_dp_out; a.rigidity_out = aux->rigidity_out;
  }
  const bool want_aux = a.pts != nullptr;      // explicit positions need the AUX instantiation; every instantiation reads bg_rand (composite_tile3)
  if (plan.bg == NF_BG_RANDOM && !a.bg_rand) return cudaErrorInvalidValue;
  a.packed = (const uint8_t*)packed; a.rays = rays; a.n_rays = n_rays; a.ts = ts; a.T = T; a.ts_stride = ts_stride;
  a.noise = noise; a.ray_time = ray_time; a.rgb_out = rgb; a.alpha_out = alpha; a.weights_out = weights;
  if (plan.kind == NF_KIND_DYN && !ray_time) return cudaErrorInvalidValue;
  if (plan.mip != NF_MIP_NONE) {
    if (!mip || !mip->radius || ts_stride != 0) return cudaErrorInvalidValue;
    a.mip = NfMipIn{plan.mip, ts, T, rays, mip->radius, mip->rays_all, mip->radius_all, (long long)mip->n_rays_all, (long long)mip->ray_base};
  }
  // The shipped library reads no environment variable: the timing-experiment hooks below (some change results: bit 2 of
  // NF_TC_DEBUG skips every MMA) exist only in NF_EXPERIMENTS builds (`NF_EXPERIMENTS=1 python -m nerf_atlas_b200.build`).
  int ring = 3, epiw = 16;
#ifdef NF_EXPERIMENTS
  if (const char* dbg = getenv("NF_TC_DEBUG")) a.debug = atoi(dbg);
  // NF_TC_EPIW selects the number of epilogue warps: 16 (default) or 24 (6 per TMEM lane quarter; 72 registers per thread).
  if (const char* r = getenv("NF_TC_EPIW")) epiw = atoi(r);
  if (epiw != 24) epiw = 16;
#endif
  // the Positional head with warp-aligned rays: per-slot 112-column x0, boundary warps (k_render_tc3<2, 4, 4, 1, ..., BW, 112>)
  const bool pos_bw = NF_BW && NF_POS_BW && plan.refl_kind == NF_REFL_POSITIONAL && plan.mip == NF_MIP_NONE && plan.kind == NF_KIND_PLAIN && (T & 31) == 0 &&
                      plan.mlp[0].k0_pad <= plan.intermediate && plan.mlp[1].k0_pad <= X0K_POS && !(aux && aux->pts);
  Tc3Prog prog;
  if (!build_prog3(plan, &prog, pos_bw ? X0K_POS : X0K)) return cudaErrorNotSupported;
  const bool wide = prog.single != 0 || plan.mip != NF_MIP_NONE || plan.refl_kind != NF_REFL_VIEW;
  if (wide) { ring = 3; epiw = 16; }
  const bool wide_shared = wide && prog.single == 2;     // two tiles in flight over one wide x0 buffer, 3 x 12 KB ring
  const bool mipk = plan.mip != NF_MIP_NONE;
  const bool dynk = plan.kind == NF_KIND_DYN;
  if (dynk) { ring = 3; epiw = 16; }
  const bool train = tp != nullptr;
  if (train) { ring = 3; epiw = 16; }
  if (train && wide && !pos_bw) return cudaErrorNotSupported;      // the Positional head trains on its boundary-warp instantiation (T % 32 == 0)
  if (want_aux) {
    // from_pts / random background: the AUX instantiation of the plain two-tile kernel, or the DynamicNeRF one (background only)
    if (wide || train || (dynk && a.pts)) return cudaErrorNotSupported;
    ring = 3; epiw = 16;
  }
  const bool auxk = want_aux && !dynk;
  Tc3Train tr{};
  if (train) {
    tr.n_tiles = tp->n_tiles;
    tr.sigma_out = (float*)((uint8_t*)ws + tp->sigma_off); tr.rgbraw_out = (float*)((uint8_t*)ws + tp->rgbraw_off);
    for (int i = 0; i < tp->n_lin; ++i) {
      const NfTrainLin& L = tp->lin[i];
      Tc3TrainLin& R = tr.lin[i];
      R.a_off256 = (uint32_t)(L.a_off >> 8); R.c_off256 = L.c_off < 0 ? 0u : (uint32_t)(L.c_off >> 8);
      R.a_tile256 = (uint32_t)(L.a_tile >> 8); R.k0 = (uint32_t)L.k0_pad; R.skip0 = R.skip1 = -1;
      if ((L.a_off >> 8) >= (1LL << 32) || (L.c_off >> 8) >= (1LL << 32)) return cudaErrorNotSupported;
    }
    for (int i = 0; i < tp->n_lin; ++i) {
      if (tp->lin[i].j != 0) continue;                       // an `init`: which later Linears of its MLP re-concatenate act(x0)?
      for (int k = i + 1; k < tp->n_lin && tp->lin[k].m == tp->lin[i].m; ++k)
        if (tp->lin[k].k0_pad) { if (tr.lin[i].skip0 < 0) tr.lin[i].skip0 = k; else if (tr.lin[i].skip1 < 0) tr.lin[i].skip1 = k; else return cudaErrorNotSupported; }
    }
  }
  // boundary-warp mode (four more warps own the tile boundary): the plain two-tile instantiations, warp-aligned rays
  // WB: the Mip encoder's shared-wide-x0 schedule with boundary warps (colour read + composite + the Mip features via an L2 scratch)
  // (not for the reference's bug-compatible layout: its per-feature crop-wide gathers want all 16 epilogue warps, 540 vs 625 ms)
  const bool wb = NF_BW && NF_WB && wide_shared && mipk && plan.mip != NF_MIP_CYLINDER_REF && (T & 31) == 0 && !want_aux;
  const bool bw = NF_BW && (!wide || pos_bw || wb) && (T & 31) == 0 && ring == 3 && epiw == 16;
  const int threads = 32 * (epiw + ring + 1 + (bw ? 4 : 0));
  const NfStreamMap map(T, ROWS);
  const long long units = map.units(n_rays);
  if (units == 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long nslot = prog.single == 1 ? 1 : 2;
  long long want = (units + 2 * nslot - 1) / (2 * nslot) * 2;                      // 2 CTAs x nslot tiles per cluster
  const int grid = (int)(want < (sms / 2) * 2 ? want : (sms / 2) * 2);
  const long long trips = (units + nslot * grid - 1) / (nslot * grid);
  if (trips * map.tpr * prog.n_lin + prog.lag >= (1LL << 30)) return cudaErrorNotSupported;   // 32-bit step counters in the kernel
#ifdef NF_TC_STATS
  static long long* d_stats = nullptr;
  if (!d_stats) cudaMalloc(&d_stats, 64 * sizeof(long long));
  cudaMemsetAsync(d_stats, 0, 64 * sizeof(long long), st);
  a.stats = d_stats;
#endif
  cudaError_t e = cudaSuccess;
  const size_t smem_bytes = pos_bw ? sizeof(Tc3SmemT<X0K_POS, RING_BYTES_POS>) : sizeof(Tc3Smem);
  auto go = [&](auto kern) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e == cudaSuccess) kern<<<grid, threads, smem_bytes, st>>>(plan, prog, a, tr);
  };
  uint8_t* scratch = nullptr;
  if (wb) {
    // 2 slots x 2 tile parities x 24 KB per CTA, stream-ordered allocation (no state kept between calls)
    if ((e = cudaMallocAsync((void**)&scratch, (size_t)grid * 4 * (NF_MIP_FEATS / 8) * KG_BYTES, st)) != cudaSuccess) return e;
    a.scratch = scratch;
  }
#if NF_WB_NST2
  if (wb) go(k_render_tc3<2, 4, 4, 2, false, false, false, true>);       // two 18 KB stages (4 K-steps): 4 chunks per Linear, straight-line issuer
#else
  if (wb) go(k_render_tc3<3, 3, 4, 2, false, false, false, true>);       // three 14 KB stages (3 K-steps)
#endif
  else if (pos_bw && train) go(k_render_tc3<2, 4, 4, 1, false, true, false, true, X0K_POS>);
  else if (pos_bw) go(k_render_tc3<2, 4, 4, 1, false, false, false, true, X0K_POS>);
  else if (bw) {
    if (train) go(k_render_tc3<3, 4, 4, 0, false, true, false, true>);
    else if (auxk) go(k_render_tc3<3, 4, 4, 0, false, false, true, true>);
    else if (dynk) go(k_render_tc3<3, 4, 4, 0, true, false, false, true>);
    else go(k_render_tc3<3, 4, 4, 0, false, false, false, true>);
  }
  else if (train) go(k_render_tc3<3, 4, 4, 0, false, true>);
  else if (auxk) go(k_render_tc3<3, 4, 4, 0, false, false, true>);
  else if (wide_shared && mipk) go(k_render_tc3<3, 3, 4, 2, false>);
  else if (wide_shared) go(k_render_tc3<3, 3, 4, 1, false>);
  else if (wide) go(k_render_tc3<3, 4, 4, 2, false>);
  else if (dynk) go(k_render_tc3<3, 4, 4, 0, true>);
#ifdef NF_EXPERIMENTS
  else if (epiw == 24) go(k_render_tc3<3, 4, 6, 0, false>);
#endif
  else go(k_render_tc3<3, 4, 4, 0, false>);
  if (scratch) { const cudaError_t ef = cudaFreeAsync(scratch, st); if (e == cudaSuccess) e = ef; }
  if (e == cudaSuccess && train && plan.bg == NF_BG_RANDOM)       // the backward's sky term needs the same draws
    e = cudaMemcpyAsync((uint8_t*)ws + tp->bgrand_off, a.bg_rand, (size_t)n_rays * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return e;
#ifdef NF_TC_STATS
  if (getenv("NF_TC_STATS_PRINT")) {
    cudaStreamSynchronize(st);
    long long h[64];
    cudaMemcpy(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[6] = {"issuer  [total, wait_a, wait_w, issue+other, n_w_waits]", "producer0 [total, wait_empty, copy, other, n]",
                            "epi w0  [total, wait_acc, phase0-encode, leaky, sin, dens_out, other+composite, colour-read]", "epi w15 [same]", "epi w0 of the peer CTA [same]",
                            "boundary warp 0 [total, wait x0_free, encode, wait bnd_full, colour read + arrive, composite]"};
    for (int r = 0; r < 6; ++r) {
      printf("STATS %s:", names[r]);
      for (int i = 0; i < 8; ++i) printf(" %lld", h[r * 8 + i]);
      printf("\n");
    }
    fflush(stdout);
  }
#endif
  return cudaGetLastError();
}
