// Microbenchmark: TMEM -> register read bandwidth per SM (tcgen05.ld 32x32b.x16 / .x32) vs number of reading warps.
// The epilogue of the render pipeline must read 128 lanes x 256 columns x 4 B = 128 KB of fp32 accumulators per tile-layer.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_read tmem_read.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int X32>
__global__ void k(int iters, int nwarps, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(&tmem)), "r"(512) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t acc = 0;
  const long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {       // 64 columns per warp per iteration
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(base + (uint32_t)(u * 16) % 64) : "memory");
        if (X32 == 0) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= v[0] ^ v[5] ^ v[15];
      }
      if (X32 == 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x % 32 == 0) { out[blockIdx.x * 32 + warp] = t1 - t0; sink[blockIdx.x * 32 + warp] = acc; }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
  long long* out; uint32_t* sink; cudaMallocManaged(&out, 148 * 32 * sizeof(long long)); cudaMallocManaged(&sink, 148 * 32 * 4);
  printf("wait_mode(0=after every ld,1=after 4 lds),reading_warps,bytes_per_cycle_per_SM\n");
  const int iters = 2000;
  for (int mode : {0, 1}) for (int nw : {4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) { if (mode) k<1><<<148, 512>>>(iters, nw, out, sink); else k<0><<<148, 512>>>(iters, nw, out, sink); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    double mx = 0; for (int b = 0; b < 148; ++b) for (int w = 0; w < nw; ++w) mx += out[b * 32 + w];
    mx /= (148.0 * nw);
    const double bytes = (double)nw * iters * 4 * 16 * 32 * 4;   // warps x iters x 4 lds x 16 cols x 32 lanes x 4 B
    printf("%d,%d,%.1f\n", mode, nw, bytes / mx);
  }
  return 0;
}
