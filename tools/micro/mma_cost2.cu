// Microbenchmark 2: tensor-pipe cost of tcgen05.mma (cta_group::1, kind::f16, bf16, K = 16) as a function of M, N and
// operand major-ness, issued the way the real kernels do it (whole warp in uniform control flow, one elected lane,
// unrolled).  One CTA per SM, operands static in shared memory.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I <csrc> mma_cost2.cu -o mma_cost2 && ./mma_cost2
#include <cstdio>
#include <cuda_runtime.h>
#include "mil_tc_ptx.cuh"

struct Cfg { int m, n, mn_major, nrep; };

__global__ void k(Cfg c, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x < 32) {
    const uint32_t mj = c.mn_major ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (mj << 15) | (mj << 16) |
                           ((uint32_t)(c.n >> 3) << 17) | ((uint32_t)(c.m >> 4) << 24);
    const uint32_t a0 = smem_u32(smem) + 1024, b0 = smem_u32(smem) + 96 * 1024;
    // K-major: LBO = next 8 K (plane stride 4096), SBO = next 8 rows (128);  MN-major: LBO = next 8 K rows (128),
    // SBO = next 8 MN (plane stride 4096)
    const uint64_t ad = c.mn_major ? make_desc(a0, 128, 4096) : make_desc(a0, 4096, 128);
    const uint64_t bd = c.mn_major ? make_desc(b0, 128, 4096) : make_desc(b0, 4096, 128);
    const uint32_t t = tbase;
    long long t0 = clock64();
    for (int r = 0; r < c.nrep; r += 8) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 8; ++j) umma_bf16(t, ad + j * 16, bd + j * 16, idesc, 1u);
      }
      __syncwarp();
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
  const int R = 4000;
  const int ns[] = {16, 24, 32, 48, 64, 72, 80, 96, 128, 160, 192, 240, 256};
  for (int mn = 0; mn < 2; ++mn)
    for (int m : {64, 128})
      for (int n : ns) {
        if (m == 128 && n % 16) continue;
        Cfg c{m, n, mn, R};
        k<<<148, 128, 200 * 1024>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%s M%-3d N%-3d : %6.1f cycles/MMA   (M+N)/4 = %5.1f   M*N/256 = %5.1f  (%s)\n", mn ? "MN-major" : "K-major ",
               m, n, (double)h / R, (m + n) / 4.0, m * n / 256.0, cudaGetErrorString(e));
      }
  return 0;
}
