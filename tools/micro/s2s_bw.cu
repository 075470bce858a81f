// Microbenchmark 5: shared -> shared bulk copies inside one CTA (cp.async.bulk.shared::cluster.shared::cta with the
// CTA's own shared window as destination): bytes per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "mil_tc_ptx.cuh"

__device__ __forceinline__ void bulk_s2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  uint32_t d32 = smem_u32(dst), b32 = smem_u32(bar), dc, bc;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(dc) : "r"(d32));
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(bc) : "r"(b32));
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dc),
               "r"(smem_u32(src)), "r"(bytes), "r"(bc)
               : "memory");
}

__global__ void k(int copy_bytes, int per_stage, int iters, int shift, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x < 32) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        mbar_expect_tx(&bar, (uint32_t)copy_bytes * per_stage);
        for (int c = 0; c < per_stage; ++c)
          bulk_s2s(smem + 100 * 1024 + (size_t)c * copy_bytes, smem + shift + (size_t)c * copy_bytes, copy_bytes, &bar);
      }
      __syncwarp();
      mbar_wait(&bar, it & 1);
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = reinterpret_cast<uint32_t*>(smem)[100 * 256 + 5]; }
  }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { int cb, ps, shift; } cs[] = {{2048, 1, 0}, {2048, 6, 16}, {2048, 12, 16}, {4096, 6, 16}, {4096, 12, 16}, {16384, 4, 16}};
  for (auto& c : cs) {
    const int iters = 2000;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k, c.cb, c.ps, iters, c.shift, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * c.cb * c.ps;
    printf("s2s copy %6d B x %2d (src shifted %2d B), one batch in flight: %6.1f B/clk/SM, %7.1f cycles per batch  check %lld  %s\n",
           c.cb, c.ps, c.shift, bytes / h[0], (double)h[0] / iters, h[1], cudaGetErrorString(e));
  }
  return 0;
}
