// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16) for the operand layouts the conv / wgrad
// kernels use.  One CTA per SM, operands static in shared memory, NREP back-to-back MMAs, clock64 around commit.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I <csrc> mma_cost.cu -o mma_cost && ./mma_cost
#include <cstdio>
#include <cuda_runtime.h>
#include "mil_tc_ptx.cuh"

struct Cfg { int m, n, amajor, bmajor, a_shift16, a_sbo, a_lbo, b_sbo, b_lbo, nrep; const char* name; int commit_every, vary, alt_d, zero_first; };

__global__ void k(Cfg c, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1 << 20); fence_barrier_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.amajor << 15) | ((uint32_t)c.bmajor << 16) |
                           ((uint32_t)(c.n >> 3) << 17) | ((uint32_t)(c.m >> 4) << 24);
    const uint32_t a0 = smem_u32(smem) + 1024 + c.a_shift16 * 16, b0 = smem_u32(smem) + 128 * 1024;
    const uint64_t ad = make_desc(a0, c.a_lbo, c.a_sbo), bd = make_desc(b0, c.b_lbo, c.b_sbo);
    long long t0 = clock64();
    for (int r = 0; r < c.nrep; ++r) {
      const int g = c.commit_every ? r / c.commit_every : 0, j = c.commit_every ? r % c.commit_every : r;
      const uint64_t a = c.vary ? ad + (uint64_t)((j * 37) % 61) : ad;
      const uint64_t b = c.vary ? bd + (uint64_t)(j * 64) : bd;
      umma_bf16(tbase + (c.alt_d ? (g & 1) * 64 : 0), a, b, idesc, c.zero_first ? (j > 0) : (r > 0));
      if (c.commit_every && j == c.commit_every - 1) { umma_commit(&bar2); if (c.commit_every < 0) umma_commit(&bar2); }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { __syncwarp(); tmem_dealloc(tbase, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int R = 2000;
  Cfg cfgs[] = {
    {128, 32, 0, 0, 0, 128, 4096, 128, 512, R, "N32 plain", 0, 0, 0, 0},
    {128, 32, 0, 0, 0, 128, 4096, 128, 512, R, "N32 commit/14", 14, 0, 0, 0},
    {128, 32, 0, 0, 0, 128, 4096, 128, 512, R, "N32 commit/14 + zero-first", 14, 0, 0, 1},
    {128, 32, 0, 0, 0, 128, 4096, 128, 512, R, "N32 commit/14 + zero-first + alt D", 14, 0, 1, 1},
    {128, 32, 0, 0, 0, 128, 4096, 128, 512, R, "N32 commit/14 + vary desc", 14, 1, 0, 0},
    {128, 32, 0, 0, 0, 128, 4096, 128, 512, R, "N32 all", 14, 1, 1, 1},
    {128, 32, 0, 0, 0, 128, 4096, 128, 512, R, "N32 commit/1", 1, 0, 0, 0},
    {128, 32, 0, 0, 0, 128, 4096, 128, 512, R, "N32 commit/4", 4, 0, 0, 0},
  };
  for (auto& c : cfgs) {
    for (int grid : {1, 148}) {
      k<<<grid, 128, 200 * 1024>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("%-44s grid %3d : %7.1f cycles/MMA  (%s)\n", c.name, grid, (double)h / c.nrep, cudaGetErrorString(e));
    }
  }
  return 0;
}
