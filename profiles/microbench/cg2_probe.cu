// Probe: every cta_group::2 mechanism the paired pipeline needs, in one tiny GEMM.
//   D[256 x 256] = A[256 x 64] * B[256 x 64]^T, fp16 in / fp32 out, one cluster of 2 CTAs per GEMM:
//   CTA r holds A rows [128r, 128r+128) and B rows (N) [128r, 128r+128), both in UMMA-canonical no-swizzle K-major smem.
//   paired TMEM alloc, remote mbarrier arrive (peer -> leader), tcgen05.mma.cta_group::2 (M=256) by ONE leader thread,
//   tcgen05.commit multicast to both CTAs, tcgen05.ld of each CTA's 128 rows.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cg2_probe cg2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t par) {
  uint32_t ok; asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(bar), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
constexpr int K = 64, KG = K / 8;
struct Smem { uint8_t A[128 * K * 2]; uint8_t B[128 * K * 2]; unsigned long long ab_ready, done; uint32_t tmem; };
extern __shared__ __align__(1024) uint8_t smem_raw[];
__global__ void __cluster_dims__(2, 1, 1) k(const __half* A, const __half* B, float* D) {
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const uint32_t r = ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x >> 1;
  const __half* Ag = A + (size_t)pair * 256 * K; const __half* Bg = B + (size_t)pair * 256 * K; float* Dg = D + (size_t)pair * 256 * 256;
  // canonical layout [kg][row][8 halves]
  for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
    const int row = i / K, k = i % K;
    const int off = (k >> 3) * 2048 + row * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(s.A + off) = Ag[(size_t)(r * 128 + row) * K + k];
    *reinterpret_cast<__half*>(s.B + off) = Bg[(size_t)(r * 128 + row) * K + k];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&s.ab_ready)), "r"(2));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&s.done)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(&s.tmem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s.tmem;
  // both CTAs tell the LEADER (rank 0) that their operands are in place: remote arrive through the shared::cluster window
  if (threadIdx.x == 0) {
    const uint32_t leader_bar = (su32(&s.ab_ready) & 0x00FFFFFFu) | (su32(&s.ab_ready) & 0xFE000000u);   // clear bit 24 = rank bit
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_bar) : "memory");
  }
  if (r == 0 && threadIdx.x == 32) {
    while (!test_wait(su32(&s.ab_ready), 0)) {}
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t a4 = su32(s.A) >> 4, b4 = su32(s.B) >> 4;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t ad = ((uint64_t)0x4008u << 32) | (((a4 + ks * 256) & 0x3FFF) | (128u << 16));
      const uint64_t bd = ((uint64_t)0x4008u << 32) | (((b4 + ks * 256) & 0x3FFF) | (128u << 16));
      asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(su32(&s.done)), "h"((uint16_t)3) : "memory");
  }
  // every CTA drains its own 128 rows
  while (!test_wait(su32(&s.done), 0)) {}
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = (warp & 3) * 32 + lane;
  for (int c0 = 0; c0 < 256; c0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) Dg[(size_t)(r * 128 + row) * 256 + c0 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}
int main() {
  const int pairs = 4;
  std::vector<__half> hA(pairs * 256 * K), hB(pairs * 256 * K);
  srand(1);
  for (auto& x : hA) x = __float2half((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : hB) x = __float2half((rand() % 2001 - 1000) / 1000.f);
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, pairs * 256 * 256 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, pairs * 256 * 256 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  k<<<pairs * 2, 128, sizeof(Smem)>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> hD(pairs * 256 * 256);
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int p = 0; p < pairs; ++p) for (int m = 0; m < 256; ++m) for (int n = 0; n < 256; ++n) {
    double acc = 0; for (int kk = 0; kk < K; ++kk) acc += (double)__half2float(hA[(size_t)(p * 256 + m) * K + kk]) * __half2float(hB[(size_t)(p * 256 + n) * K + kk]);
    const double err = fabs(acc - hD[(size_t)(p * 256 + m) * 256 + n]);
    if (!(err < 1e-2)) { if (bad < 8) printf("mismatch pair %d m %d n %d: ref %f got %f\n", p, m, n, acc, hD[(size_t)(p * 256 + m) * 256 + n]); ++bad; }
    if (err > maxerr) maxerr = err;
  }
  printf("cg2_probe: max err %.3e, mismatches %d of %d -> %s\n", maxerr, bad, pairs * 65536, bad ? "FAIL" : "PASS");
  return bad != 0;
}
