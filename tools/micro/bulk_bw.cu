// Microbenchmark 4: global -> shared bulk-copy (cp.async.bulk, 1-D) throughput per SM as a function of copy size,
// copies in flight and footprint (L2-resident vs DRAM).  One CTA per SM, one elected lane issues.
#include <cstdio>
#include <cuda_runtime.h>
#include "mil_tc_ptx.cuh"

__global__ void k(const unsigned char* src, long long footprint, int copy_bytes, int per_stage, int depth, int iters,
                  long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x < 32) {
    const long long stage_bytes = (long long)copy_bytes * per_stage;
    const long long cta_span = footprint / gridDim.x / stage_bytes * stage_bytes;
    const unsigned char* base = src + (long long)blockIdx.x * cta_span;
    long long off = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters + depth; ++it) {
      const int st = it % depth;
      if (it >= depth) mbar_wait(&full[st], ((it / depth) - 1) & 1);
      if (it < iters) {
        if (elect_one()) {
          mbar_expect_tx(&full[st], (uint32_t)stage_bytes);
          for (int c = 0; c < per_stage; ++c)
            bulk_g2s(smem + (size_t)st * stage_bytes + (size_t)c * copy_bytes, base + off + (long long)c * copy_bytes,
                     copy_bytes, &full[st]);
        }
        __syncwarp();
        off += stage_bytes;
        if (off + stage_bytes > cta_span) off = 0;
      }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
}

int main() {
  const long long N = 2LL << 30;
  unsigned char* src; cudaMalloc(&src, N); cudaMemset(src, 1, N);
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { int cb, ps, depth; } cs[] = {{2048, 1, 8}, {2048, 4, 4}, {2048, 16, 3}, {4096, 8, 3}, {16384, 2, 3}, {16384, 4, 3},
                                        {2048, 18, 4}, {2048, 32, 3}, {32768, 2, 3}};
  for (long long fp : {32LL << 20, 2LL << 30})
    for (auto& c : cs) {
      const int iters = 2000;
      k<<<148, 128, 200 * 1024>>>(src, fp, c.cb, c.ps, c.depth, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)iters * c.cb * c.ps;
      printf("footprint %5lld MB  copy %6d B x %2d per stage, %d stages in flight: %6.1f B/clk/SM  (%5.2f TB/s at 1.9 GHz x 148)  %s\n",
             fp >> 20, c.cb, c.ps, c.depth, bytes / h, bytes / h * 148 * 1.9e9 / 1e12, cudaGetErrorString(e));
    }
  return 0;
}
