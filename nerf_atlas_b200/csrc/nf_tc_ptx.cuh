// nf_tc_ptx.cuh -- PTX wrappers shared by the tcgen05 pipelines (nf_tc.cu, nf_tc2.cu): mbarrier, bulk copy (TMA 1-D),
// tcgen05 mma/commit/ld/fences, UMMA descriptors.  Measured behaviour the wrappers encode is documented in DESIGN.md section 5.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include "nf_common.cuh"

namespace nf_ptx {
constexpr int ROWS = NF_TC_ROWS;          // 128 = UMMA M per CTA
constexpr int KG_BYTES = ROWS * 16;       // one 8-column K-group of an A operand: 128 rows x 16 B

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU.
// A protocol bug must surface as a trapped launch (cudaErrorLaunchFailure), never as a hung GPU.  No function call
// here: a call in the wait loop makes ptxas drop the MMA issuer's descriptors out of uniform registers.
__device__ __forceinline__ void mbar_timeout(uint32_t) { __trap(); }
// Non-blocking probe. Measured on B200 (profiles/trace_*.txt): a failed mbarrier.try_wait suspends the thread for a
// ~440-cycle quantum and is NOT woken early by async-proxy completions (TMA complete_tx, tcgen05.commit), so every
// wait on the critical path polls with test_wait instead.
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_test_wait(bar, parity)) { if (++spins > (1u << 26)) mbar_timeout(bar); }
}
// two barriers polled together: the two test_wait latencies (~150 cycles each) overlap
__device__ __forceinline__ void mbar_wait2(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b) {
  uint32_t spins = 0;
  while (true) {
    const bool oa = mbar_test_wait(bar_a, par_a), ob = mbar_test_wait(bar_b, par_b);
    if (oa && ob) break;
    if (++spins > (1u << 26)) mbar_timeout(oa ? bar_b : bar_a);
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred P1;\n elect.sync _|P1, 0xffffffff;\n selp.u32 %0, 1, 0, P1;\n}" : "=r"(pred));
  return pred != 0;
}
// Long waits (epilogue warps waiting for a whole layer of MMAs): try_wait suspends the warp in hardware (no issue slots
// burnt, ~440-cycle wake-up quantum), which beats polling when 8-16 warps wait at once (ncu: polling loops were ~60 % of
// all executed instructions).
__device__ __forceinline__ void mbar_wait_suspend(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 22)) mbar_timeout(bar); }
}
// for the 8 epilogue warps (they share schedulers with the MMA issuer): poll, but yield between polls
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_test_wait(bar, parity)) { __nanosleep(32); if (++spins > (1u << 24)) mbar_timeout(bar); }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// K-major, no-swizzle UMMA shared-memory descriptors: 8x16-byte core matrices; LBO (bits 16-29) = byte distance
// between the two K-adjacent cores of one K=16 step, SBO (bits 32-45) = byte distance between 8-row groups.
// high word shared by every operand here: SBO = 128 B, descriptor version 1 (bit 46)
__device__ __forceinline__ uint64_t umma_desc_lo(uint32_t lo) { return ((uint64_t)0x4008u << 32) | lo; }
// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=n.
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ties 16 registers to the preceding tcgen05.wait::ld so no consumer is scheduled above it
__device__ __forceinline__ void reg_fence16(uint32_t* v) {
  asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                    "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory");
}
// ---- CTA-pair (cta_group::2) helpers, shared by nf_tc2.cu / nf_tc3.cu -------------------------------
// shared::cluster address of the same smem object in the pair's leader (rank 0): bit 24 of the window address is the rank
__device__ __forceinline__ uint32_t leader_addr(uint32_t a) { return a & 0xFEFFFFFFu; }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// After fence.proxy.async the payload (this CTA's smem, read by this SM's tensor core through the async proxy) is already
// ordered; the cross-CTA arrive then only has to be delivered, not to release memory at cluster scope.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <int ACT> __device__ __forceinline__ float tc_act_t(float x) {
  if (ACT == NF_ACT_LEAKY) return fmaxf(x, 0.01f * x);
  if (ACT == NF_ACT_SIN) return __sinf(x);
  if (ACT == NF_ACT_RELU) return fmaxf(x, 0.f);
  return x;
}

__device__ __forceinline__ float tc_act(float x, int act) {
  switch (act) {
    case NF_ACT_LEAKY: return fmaxf(x, 0.01f * x);
    case NF_ACT_SIN:   return __sinf(x);
    case NF_ACT_RELU:  return fmaxf(x, 0.f);
    default:           return x;
  }
}
// act(x0), act(x1) -> packed fp16x2.  LeakyReLU/ReLU run as ONE packed half2 op pair after the conversion (the result is
// rounded to fp16 anyway); sin stays in fp32 (MUFU) and is rounded afterwards.
template <int ACT> __device__ __forceinline__ uint32_t act_pack_t(float x0, float x1) {
  if (ACT == NF_ACT_LEAKY) {
    __half2 h = __floats2half2_rn(x0, x1);
    h = __hmax2(h, __hmul2(h, __float2half2_rn(0.01f)));
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  if (ACT == NF_ACT_RELU) {
    __half2 h = __floats2half2_rn(x0, x1);
    h = __hmax2(h, __float2half2_rn(0.f));
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __half2 h = __floats2half2_rn(tc_act_t<ACT>(x0), tc_act_t<ACT>(x1));
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void st_v4(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}


}  // namespace nf_ptx
