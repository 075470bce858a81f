// Microbenchmark 3: the per-tile MMA patterns of conv_tc (14 x M128 N32 K-major into one of 4 accumulators, commit)
// and wgrad_tc (3 x 8 x M64 N72 MN-major into 3 accumulators + 8 x M64 N16, commit), operands static in shared
// memory, no loads, no epilogue: what the tensor pipe alone takes per tile.
#include <cstdio>
#include <cuda_runtime.h>
#include "mil_tc_ptx.cuh"

__global__ void k(int pattern, int tiles, int nbias, int n_main, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar, sink;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&sink, 1 << 20); fence_barrier_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x < 32) {
    const uint32_t a0 = smem_u32(smem) + 1024, b0 = smem_u32(smem) + 96 * 1024;
    const uint32_t t = tbase;
    long long t0 = clock64();
    if (pattern == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_main >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t ad = make_desc(a0, 4096, 128), bd = make_desc(b0, 512, 128);
      for (int it = 0; it < tiles; ++it) {
        const uint32_t d = t + (it & 3) * 32;
        if (elect_one()) {
          umma_bf16(d, ad, bd, idesc, 0u);
#pragma unroll
          for (int j = 1; j < 14; ++j) umma_bf16(d, ad + j * 59, bd + j * 64, idesc, 1u);
          umma_commit(&sink);
        }
        __syncwarp();
      }
    } else {
      const uint32_t ib = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 4) << 24);
      const uint32_t idesc = ib | ((uint32_t)(n_main >> 3) << 17), idesc_b = ib | ((uint32_t)(16 >> 3) << 17);
      const uint64_t ad = make_desc(a0, 128, 2048), bd = make_desc(b0, 128, 3904);
      for (int it = 0; it < tiles; ++it) {
        if (elect_one()) {
          for (int tl = 0; tl < 3; ++tl) {
#pragma unroll
            for (int j = 0; j < 8; ++j) umma_bf16(t + tl * 80, ad + j * 16, bd + tl * 58 + j * 16, idesc, 1u);
          }
          if (nbias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) umma_bf16(t + 256, ad + j * 16, bd, idesc_b, 1u);
          }
          umma_commit(&sink);
        }
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { __syncwarp(); tmem_dealloc(tbase, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int T = 400;
  struct { int p, nb, n; const char* nm; } cs[] = {
      {0, 0, 32, "conv  14 x M128 N32, 4 accumulators"}, {0, 0, 16, "conv  14 x M128 N16"}, {0, 0, 64, "conv  14 x M128 N64"},
      {1, 1, 72, "wgrad 24 x M64 N72 + 8 x M64 N16"}, {1, 0, 72, "wgrad 24 x M64 N72"}, {1, 0, 80, "wgrad 24 x M64 N80"},
      {1, 0, 48, "wgrad 24 x M64 N48"}};
  for (auto& c : cs) {
    k<<<148, 128, 200 * 1024>>>(c.p, T, c.nb, c.n, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-40s : %7.1f cycles/tile  (%s)\n", c.nm, (double)h / T, cudaGetErrorString(e));
  }
  return 0;
}
