// Tail of the extractor: global average pool over the layer-4 map, flatten, fc 80 -> 80 without bias
// (reference gbm/model.py:31-32,58-60) and its backward.  fp32 throughout (hard part 4).
#include <algorithm>

#include "mil_common.cuh"

#define TAIL_C 80

// One block per tile, one warp per 8-channel chunk: the lanes stride over the tile's pixels with one chunk load each
// (a chunk plane of a tile is contiguous: coalesced), butterfly-reduce the eight channel sums, then every warp
// computes eight rows of the fc with coalesced weight reads.
#define TAIL_THREADS (TAIL_C / 8 * 32)
template <typename T>
__global__ void __launch_bounds__(TAIL_THREADS)
tail_fwd_kernel(const T* __restrict__ y4, MilPF8 g4, const float* __restrict__ wfc, float* __restrict__ avg,
                float* __restrict__ H) {
  __shared__ float s_avg[TAIL_C];
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const T* base = y4 + mil_pf8_off(g4, warp, (long long)n * g4.P);
  for (int p = lane; p < (int)g4.P; p += 32) {
    const int yy = p / g4.wp, xx = p - yy * g4.wp;
    if (yy < g4.h && xx < g4.w) {
      float v[8];
      mil_load8(base + (size_t)p * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a[j] += __shfl_xor_sync(0xFFFFFFFFu, a[j], off);
  if (lane == 0) {
    const float inv = 1.f / (float)(g4.h * g4.w);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_avg[warp * 8 + j] = a[j] * inv;
      avg[(size_t)n * TAIL_C + warp * 8 + j] = a[j] * inv;
    }
  }
  __syncthreads();
  // H[n][c] = sum_k Wfc[c][k] avg[k]: warp w owns outputs 8w .. 8w+7, the lanes split k
  float av[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) av[i] = lane + 32 * i < TAIL_C ? s_avg[lane + 32 * i] : 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int c = warp * 8 + r;
    float o = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (lane + 32 * i < TAIL_C) o = fmaf(__ldg(wfc + c * TAIL_C + lane + 32 * i), av[i], o);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) o += __shfl_xor_sync(0xFFFFFFFFu, o, off);
    if (lane == 0) H[(size_t)n * TAIL_C + c] = o;
  }
}

int mil_launch_tail_fwd(int dtype, const void* y4, const MilPF8& g4, const float* wfc, float* avg, float* H,
                        cudaStream_t s) {
  MIL_REQUIRE(g4.c == TAIL_C, "tail_fwd: expected 80 channels");
  if (dtype == MIL_BF16)
    tail_fwd_kernel<<<g4.n, TAIL_THREADS, 0, s>>>((const __nv_bfloat16*)y4, g4, wfc, avg, H);
  else
    tail_fwd_kernel<<<g4.n, TAIL_THREADS, 0, s>>>((const float*)y4, g4, wfc, avg, H);
  MIL_LAUNCH_OK();
  return 0;
}

// backward:  davg[n][c] = sum_o dH[n][o] Wfc[o][c];  dz4(n,y,x)[c] = davg[n][c]/(h*w) * lrelu'(y4(n,y,x)[c])
//            dWfc[o][c] += sum_n dH[n][o] avg[n][c]
template <typename T>
__global__ void __launch_bounds__(TAIL_THREADS)
tail_bwd_dz_kernel(const T* __restrict__ y4, MilPF8 g4, const float* __restrict__ wfc, const float* __restrict__ dH,
                   T* __restrict__ dz4) {
  __shared__ float s_dh[TAIL_C], s_d[TAIL_C];
  const int n = blockIdx.x;
  const int c = threadIdx.x;
  if (c < TAIL_C) s_dh[c] = dH[(size_t)n * TAIL_C + c];
  __syncthreads();
  if (c < TAIL_C) {
    float d = 0.f;
    for (int o = 0; o < TAIL_C; ++o) d = fmaf(s_dh[o], __ldg(wfc + o * TAIL_C + c), d);
    s_d[c] = d * (1.f / (float)(g4.h * g4.w));
  }
  __syncthreads();
  // warp = chunk, lanes stride over the tile's pixels (pad pixels are written as zeros): coalesced chunk loads / stores
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float d8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) d8[j] = s_d[warp * 8 + j];
  const long long o0 = mil_pf8_off(g4, warp, (long long)n * g4.P);
  for (int p = lane; p < (int)g4.P; p += 32) {
    const int yy = p / g4.wp, xx = p - yy * g4.wp;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (yy < g4.h && xx < g4.w) {
      float a[8];
      mil_load8(y4 + o0 + (size_t)p * 8, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = d8[j] * mil_lrelu_grad(a[j]);
    }
    mil_store8(dz4 + o0 + (size_t)p * 8, v);
  }
}

// dWfc[o][c] = sum_n dH[n][o] avg[n][c]: every block takes a range of tiles, stages 32 rows of dH and avg at a time in
// shared memory, and thread (to, tc) of its 16 x 16 threads accumulates the 5 x 5 outputs o = 5 to + i, c = 5 tc + j
// (10 shared-memory reads per 25 FMAs); one partial record per block, fixed-order reduction afterwards.
#define TAIL_WG_BLOCKS 128
#define TAIL_WG_ROWS 32
__global__ void __launch_bounds__(256)
tail_bwd_w_kernel(const float* __restrict__ avg, const float* __restrict__ dH, int n_tiles,
                  float* __restrict__ partial) {
  __shared__ float s_dh[TAIL_WG_ROWS][TAIL_C], s_av[TAIL_WG_ROWS][TAIL_C];
  const int per = (int)mil_cdiv(n_tiles, (int)gridDim.x);
  const int n0 = blockIdx.x * per, n1 = min(n0 + per, n_tiles);
  const int to = threadIdx.x >> 4, tc = threadIdx.x & 15;
  float acc[5][5];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) acc[i][j] = 0.f;
  for (int nb = n0; nb < n1; nb += TAIL_WG_ROWS) {
    const int rows = min(TAIL_WG_ROWS, n1 - nb);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * TAIL_C; i += blockDim.x) {
      (&s_dh[0][0])[i] = dH[(size_t)nb * TAIL_C + i];
      (&s_av[0][0])[i] = avg[(size_t)nb * TAIL_C + i];
    }
    __syncthreads();
    for (int r = 0; r < rows; ++r) {
      float d[5], a[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) { d[i] = s_dh[r][5 * to + i]; a[i] = s_av[r][5 * tc + i]; }
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = fmaf(d[i], a[j], acc[i][j]);
    }
  }
  float* rec = partial + (size_t)blockIdx.x * TAIL_C * TAIL_C;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) rec[(5 * to + i) * TAIL_C + 5 * tc + j] = acc[i][j];
}

size_t mil_tail_bwd_partial_floats() { return (size_t)TAIL_WG_BLOCKS * TAIL_C * TAIL_C; }

int mil_launch_tail_bwd(int dtype, const void* y4, const MilPF8& g4, const float* wfc, const float* avg,
                        const float* dH, void* dz4, float* partial, float* dwfc, cudaStream_t s) {
  if (dtype == MIL_BF16)
    tail_bwd_dz_kernel<<<g4.n, TAIL_THREADS, 0, s>>>((const __nv_bfloat16*)y4, g4, wfc, dH, (__nv_bfloat16*)dz4);
  else
    tail_bwd_dz_kernel<<<g4.n, TAIL_THREADS, 0, s>>>((const float*)y4, g4, wfc, dH, (float*)dz4);
  MIL_LAUNCH_OK();
  tail_bwd_w_kernel<<<TAIL_WG_BLOCKS, 256, 0, s>>>(avg, dH, g4.n, partial);
  MIL_LAUNCH_OK();
  return mil_launch_reduce_partials(partial, TAIL_WG_BLOCKS, TAIL_C * TAIL_C, dwfc, TAIL_C * TAIL_C, s);
}
