// Microbenchmark: round-trip latency of the MMA-issuer <-> epilogue handshake on B200.
//   issuer (1 thread):   signal X            -> poll Y (count = 8 warps)
//   8 epilogue warps:    poll X -> [fences]  -> lane 0 arrives on Y
// variants: X signalled by tcgen05.commit (UTCBAR) or by a plain mbarrier.arrive; epilogue fences on/off;
//           epilogue poll = test_wait spin / test_wait + nanosleep(32) / try_wait.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o handshake handshake.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t par) {
  uint32_t ok; asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(bar), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par) {
  uint32_t ok; asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(bar), "r"(par) : "memory"); return ok; }
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
// mode bits: 1 = X via tcgen05.commit, 2 = epilogue fences (tcgen05 fences + fence.proxy.async), 4 = nanosleep(32) backoff, 8 = try_wait
__global__ void k(int iters, int mode, long long* out) {
  __shared__ unsigned long long X, Y; __shared__ uint32_t tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&X))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(su32(&Y))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 8) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(&tmem)), "r"(32) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  __syncthreads();
  if (warp == 9 && lane == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (mode & 1) commit(su32(&X)); else arrive(su32(&X));
      while (!test_wait(su32(&Y), i & 1)) {}
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    out[blockIdx.x] = clock64() - t0;
  } else if (warp < 8) {
    for (int i = 0; i < iters; ++i) {
      if (mode & 8) { while (!try_wait(su32(&X), i & 1)) {} }
      else if (mode & 4) { while (!test_wait(su32(&X), i & 1)) { __nanosleep(32); } }
      else { while (!test_wait(su32(&X), i & 1)) {} }
      if (mode & 2) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
      __syncwarp();
      if (lane == 0) arrive(su32(&Y));
    }
  }
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}
int main() {
  long long* out; cudaMallocManaged(&out, 148 * sizeof(long long));
  printf("mode(1=commit,2=fences,4=nanosleep32,8=try_wait),cycles_per_round_trip\n");
  for (int mode : {0, 1, 2, 3, 4, 5, 7, 8, 9, 11}) {
    const int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) { k<<<148, 320>>>(iters, mode, out); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    double a = 0; for (int i = 0; i < 148; ++i) a += out[i];
    printf("%d,%.1f\n", mode, a / 148 / iters);
  }
  return 0;
}
