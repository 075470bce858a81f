// Tail of the extractor: global average pool over the layer-4 map, flatten, fc 80 -> 80 without bias
// (reference gbm/model.py:31-32,58-60) and its backward.  fp32 throughout (hard part 4).
#include <algorithm>

#include "mil_common.cuh"

#define TAIL_C 80

template <typename T>
__global__ void __launch_bounds__(96)
tail_fwd_kernel(const T* __restrict__ y4, MilPF8 g4, const float* __restrict__ wfc, float* __restrict__ avg,
                float* __restrict__ H) {
  __shared__ float s_avg[TAIL_C];
  const int n = blockIdx.x;
  const int c = threadIdx.x;
  if (c < TAIL_C) {
    float a = 0.f;
    for (int y = 0; y < g4.h; ++y)
      for (int x = 0; x < g4.w; ++x) {
        const long long q = (long long)n * g4.P + (long long)y * g4.wp + x;
        a += mil_to_float(y4[mil_pf8_off(g4, c >> 3, q) + (c & 7)]);
      }
    a *= 1.f / (float)(g4.h * g4.w);
    s_avg[c] = a;
    avg[(size_t)n * TAIL_C + c] = a;
  }
  __syncthreads();
  if (c < TAIL_C) {
    float o = 0.f;
    for (int k = 0; k < TAIL_C; ++k) o = fmaf(wfc[c * TAIL_C + k], s_avg[k], o);
    H[(size_t)n * TAIL_C + c] = o;
  }
}

int mil_launch_tail_fwd(int dtype, const void* y4, const MilPF8& g4, const float* wfc, float* avg, float* H,
                        cudaStream_t s) {
  MIL_REQUIRE(g4.c == TAIL_C, "tail_fwd: expected 80 channels");
  if (dtype == MIL_BF16)
    tail_fwd_kernel<<<g4.n, 96, 0, s>>>((const __nv_bfloat16*)y4, g4, wfc, avg, H);
  else
    tail_fwd_kernel<<<g4.n, 96, 0, s>>>((const float*)y4, g4, wfc, avg, H);
  MIL_LAUNCH_OK();
  return 0;
}

// backward:  davg[n][c] = sum_o dH[n][o] Wfc[o][c];  dz4(n,y,x)[c] = davg[n][c]/(h*w) * lrelu'(y4(n,y,x)[c])
//            dWfc[o][c] += sum_n dH[n][o] avg[n][c]
template <typename T>
__global__ void __launch_bounds__(96)
tail_bwd_dz_kernel(const T* __restrict__ y4, MilPF8 g4, const float* __restrict__ wfc, const float* __restrict__ dH,
                   T* __restrict__ dz4) {
  __shared__ float s_dh[TAIL_C];
  const int n = blockIdx.x;
  const int c = threadIdx.x;
  if (c < TAIL_C) s_dh[c] = dH[(size_t)n * TAIL_C + c];
  __syncthreads();
  if (c >= TAIL_C) return;
  float d = 0.f;
  for (int o = 0; o < TAIL_C; ++o) d = fmaf(s_dh[o], wfc[o * TAIL_C + c], d);
  d *= 1.f / (float)(g4.h * g4.w);
  for (int y = 0; y < g4.hp; ++y)
    for (int x = 0; x < g4.wp; ++x) {
      const long long q = (long long)n * g4.P + (long long)y * g4.wp + x;
      const long long o = mil_pf8_off(g4, c >> 3, q) + (c & 7);
      float v = 0.f;
      if (y < g4.h && x < g4.w) v = d * mil_lrelu_grad(mil_to_float(y4[o]));
      mil_from_float(dz4 + o, v);
    }
}

#define TAIL_WG_BLOCKS 64
__global__ void __launch_bounds__(256)
tail_bwd_w_kernel(const float* __restrict__ avg, const float* __restrict__ dH, int n_tiles,
                  float* __restrict__ partial) {
  // thread owns 25 consecutive outputs of the 80x80 gradient; block owns a range of tiles
  const int per = (int)mil_cdiv(n_tiles, (int)gridDim.x);
  const int n0 = blockIdx.x * per, n1 = min(n0 + per, n_tiles);
  float acc[25];
#pragma unroll
  for (int k = 0; k < 25; ++k) acc[k] = 0.f;
  const int base = threadIdx.x * 25;  // o = base/80, c = base%80 ... (25 divides 80? no -> general indexing)
  for (int n = n0; n < n1; ++n) {
#pragma unroll
    for (int k = 0; k < 25; ++k) {
      const int i = base + k;
      acc[k] = fmaf(dH[(size_t)n * TAIL_C + i / TAIL_C], avg[(size_t)n * TAIL_C + i % TAIL_C], acc[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 25; ++k) partial[(size_t)blockIdx.x * TAIL_C * TAIL_C + base + k] = acc[k];
}

size_t mil_tail_bwd_partial_floats() { return (size_t)TAIL_WG_BLOCKS * TAIL_C * TAIL_C; }

int mil_launch_tail_bwd(int dtype, const void* y4, const MilPF8& g4, const float* wfc, const float* avg,
                        const float* dH, void* dz4, float* partial, float* dwfc, cudaStream_t s) {
  if (dtype == MIL_BF16)
    tail_bwd_dz_kernel<<<g4.n, 96, 0, s>>>((const __nv_bfloat16*)y4, g4, wfc, dH, (__nv_bfloat16*)dz4);
  else
    tail_bwd_dz_kernel<<<g4.n, 96, 0, s>>>((const float*)y4, g4, wfc, dH, (float*)dz4);
  MIL_LAUNCH_OK();
  tail_bwd_w_kernel<<<TAIL_WG_BLOCKS, 256, 0, s>>>(avg, dH, g4.n, partial);
  MIL_LAUNCH_OK();
  return mil_launch_reduce_partials(partial, TAIL_WG_BLOCKS, TAIL_C * TAIL_C, dwfc, TAIL_C * TAIL_C, s);
}
