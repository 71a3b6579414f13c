// Microbenchmark: per-SM L2->smem ingest rate on B200 for (a) cp.async.bulk (1-D TMA) and (b) LDGSTS (cp.async 16 B/lane),
// as a function of chunk size and chunks in flight, with every SM reading the SAME buffer (the weight-streaming pattern)
// or distinct buffers.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o copy_bw copy_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#ifdef USE_TRY_WAIT
#define WAITOP "try_wait"
#else
#define WAITOP "test_wait"   // a failed try_wait sleeps a ~430-cycle quantum and hides the real engine rate
#endif
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par) {
  uint32_t ok; asm volatile("{ .reg .pred p; mbarrier." WAITOP ".parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(bar), "r"(par) : "memory"); return ok; }
extern __shared__ __align__(1024) uint8_t smem[];
// mode 0: cp.async.bulk, one thread; mode 1: LDGSTS by `nw` warps
__global__ void k_bulk(const uint8_t* src, size_t per_sm_stride, int region_bytes, int chunk, int stages, int iters, long long* cycles) {
  uint64_t* bars = (uint64_t*)smem; uint8_t* buf = smem + 1024;
  const uint8_t* my = src + (size_t)blockIdx.x * per_sm_stride;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int nchunk = region_bytes / chunk;
    long long t0 = clock64();
    int issued = 0, done = 0; uint32_t par = 0; int st_i = 0, st_w = 0;
    while (done < iters) {
      while (issued < iters && issued - done < stages) {
        const uint32_t bar = su32(&bars[st_i]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(su32(buf + (size_t)st_i * chunk)), "l"(my + (size_t)(issued % nchunk) * chunk), "r"(chunk), "r"(bar) : "memory");
        ++issued; if (++st_i == stages) st_i = 0;
      }
      while (!try_wait(su32(&bars[st_w]), par)) {}
      ++done; if (++st_w == stages) { st_w = 0; par ^= 1; }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}
__global__ void k_ldgsts(const uint8_t* src, size_t per_sm_stride, int region_bytes, int chunk, int stages, int iters, long long* cycles) {
  uint8_t* buf = smem + 1024;
  const uint8_t* my = src + (size_t)blockIdx.x * per_sm_stride;
  const int nchunk = region_bytes / chunk;
  long long t0 = clock64();
  // all threads cooperate on every chunk; `stages` commit groups in flight
  for (int it = 0; it < iters + stages; ++it) {
    if (it < iters) {
      const uint8_t* g = my + (size_t)(it % nchunk) * chunk; uint8_t* d = buf + (size_t)(it % stages) * chunk;
      for (int o = threadIdx.x * 16; o < chunk; o += blockDim.x * 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(su32(d + o)), "l"(g + o) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(7) : "memory");   // placeholder, real wait below
    if (it >= stages - 1) { /* oldest group done when at most stages-1 pending */ }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
// NW independent rings, each driven by lane 0 of its own warp: is the ~2-in-flight limit per SM or per issuing warp?
__global__ void k_bulk_multi(const uint8_t* src, int region_bytes, int chunk, int stages, int iters, int nw, long long* cycles) {
  uint64_t* bars = (uint64_t*)smem; uint8_t* buf = smem + 1024;
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < nw) {
    uint64_t* mybars = bars + w * 8; uint8_t* mybuf = buf + (size_t)w * stages * chunk;
    for (int i = 0; i < stages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&mybars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int nchunk = region_bytes / chunk;
    long long t0 = clock64();
    int issued = 0, done = 0; uint32_t par = 0; int st_i = 0, st_w = 0;
    while (done < iters) {
      while (issued < iters && issued - done < stages) {
        const uint32_t bar = su32(&mybars[st_i]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(su32(mybuf + (size_t)st_i * chunk)), "l"(src + (size_t)((issued * nw + w) % nchunk) * chunk), "r"(chunk), "r"(bar) : "memory");
        ++issued; if (++st_i == stages) st_i = 0;
      }
      while (!try_wait(su32(&mybars[st_w]), par)) {}
      ++done; if (++st_w == stages) { st_w = 0; par ^= 1; }
    }
    if (w == 0) cycles[blockIdx.x] = clock64() - t0;
  }
}
template <int S> __global__ void k_ldgsts_s(const uint8_t* src, size_t per_sm_stride, int region_bytes, int chunk, int iters, long long* cycles) {
  uint8_t* buf = smem + 1024;
  const uint8_t* my = src + (size_t)blockIdx.x * per_sm_stride;
  const int nchunk = region_bytes / chunk;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint8_t* g = my + (size_t)(it % nchunk) * chunk; uint8_t* d = buf + (size_t)(it % S) * chunk;
    for (int o = threadIdx.x * 16; o < chunk; o += blockDim.x * 16)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(su32(d + o)), "l"(g + o) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
int main() {
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int region = 512 * 1024;
  uint8_t* src; cudaMalloc(&src, (size_t)region * sms); cudaMemset(src, 1, (size_t)region * sms);
  long long* cyc; cudaMallocManaged(&cyc, sizeof(long long) * sms);
  const int smem_bytes = 200 * 1024;
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(k_ldgsts_s<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(k_ldgsts_s<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(k_ldgsts_s<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  printf("mode,shared_src,chunk,stages,B_per_cycle_per_SM,agg_TBps_at_1.9GHz\n");
  for (int shared_src = 1; shared_src >= (getenv("ONLY_MULTI") ? 2 : 0); --shared_src)
    for (int chunk : {2048, 4096, 8192, 16384, 32768})
      for (int stages : {1, 2, 4, 6}) {
        if ((size_t)chunk * stages > 190 * 1024) continue;
        const int iters = 4000;
        for (int rep = 0; rep < 2; ++rep) { k_bulk<<<sms, 128, smem_bytes>>>(src, shared_src ? 0 : region, region, chunk, stages, iters, cyc); cudaDeviceSynchronize(); }
        cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        double mx = 0; for (int i = 0; i < sms; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
        const double bpc = (double)chunk * iters / mx;
        printf("bulk,%d,%d,%d,%.1f,%.2f\n", shared_src, chunk, stages, bpc, bpc * sms * 1.9e9 / 1e12);
      }
  cudaFuncSetAttribute(k_bulk_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int nw : {1, 2, 4})
    for (int chunk : {4096, 8192, 16384})
      for (int stages : {2, 3}) {
        if ((size_t)chunk * stages * nw > 190 * 1024) continue;
        const int iters = 4000;
        for (int rep = 0; rep < 2; ++rep) { k_bulk_multi<<<sms, 128, smem_bytes>>>(src, region, chunk, stages, iters, nw, cyc); cudaDeviceSynchronize(); }
        double mx = 0; for (int i = 0; i < sms; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
        const double bpc = (double)chunk * iters * nw / mx;
        printf("bulk_multi_nw%d,1,%d,%d,%.1f,%.2f\n", nw, chunk, stages, bpc, bpc * sms * 1.9e9 / 1e12);
      }
  if (getenv("ONLY_MULTI")) return 0;
  for (int shared_src = 1; shared_src >= 0; --shared_src)
    for (int threads : {128, 256, 512})
      for (int chunk : {8192, 16384}) {
        const int iters = 4000;
        for (int S : {2, 4, 8}) {
          for (int rep = 0; rep < 2; ++rep) {
            if (S == 2) k_ldgsts_s<2><<<sms, threads, smem_bytes>>>(src, shared_src ? 0 : region, region, chunk, iters, cyc);
            if (S == 4) k_ldgsts_s<4><<<sms, threads, smem_bytes>>>(src, shared_src ? 0 : region, region, chunk, iters, cyc);
            if (S == 8) k_ldgsts_s<8><<<sms, threads, smem_bytes>>>(src, shared_src ? 0 : region, region, chunk, iters, cyc);
            cudaDeviceSynchronize();
          }
          double mx = 0; for (int i = 0; i < sms; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
          const double bpc = (double)chunk * iters / mx;
          printf("ldgsts_t%d,%d,%d,%d,%.1f,%.2f\n", threads, shared_src, chunk, S, bpc, bpc * sms * 1.9e9 / 1e12);
        }
      }
  return 0;
}
