// Stem of the extractor: conv 7x7 / stride 2 / pad 3, 3 -> 20 channels, + bias, LeakyReLU(0.1), max-pool
// 3x3 / stride 2 / pad 1 -- fused: the 112x112x20 conv map (the largest tensor of the network) never goes to
// HBM.  Reference: gbm/model.py:24-26,51-53.  Input is the caller's fp32 NCHW bag (optionally gathered
// through the train-mode subsample index list, gbm/model.py:193-194); output is the PF8 pooled map plus a
// 1-byte argmax (0..8, window scan order, first maximum wins like ATen's max_pool2d) per pooled element.
//
// Backward (input is detached -> only d/dW and d/db, gbm/model.py:194,196): the pooled gradient is routed
// through the saved argmax to ONE conv-output position per (pooled pixel, channel); the weight gradient is
// the sum of (gradient x the 3x7x7 input patch of that position).
#include <algorithm>

#include "mil_common.cuh"

#define STEM_CO 20
#define STEM_K 147   // 3*7*7
#define STEM_TP 8    // pooled tile side per block
#define STEM_TC 17   // conv tile side  = 2*TP+1
#define STEM_TI 39   // input tile side = 2*(TC-1)+7
#define STEM_TIW 40  // padded row length in shared memory

// bf16 mode: these kernels are the cross-check of the tensor-core stem (mil_stem_tc.cu), so they round where it rounds
// -- the input tile, the weights and the conv map go through bf16 -- and differ from it by summation order only
template <typename T> __device__ __forceinline__ float stem_round(float v) { return v; }
template <> __device__ __forceinline__ float stem_round<__nv_bfloat16>(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

template <typename T>
__global__ void __launch_bounds__(256)
stem_fwd_kernel(const float* __restrict__ x, const int* __restrict__ idx, int side, int hc, const float* __restrict__ w,
                const float* __restrict__ b, T* __restrict__ pooled, MilPF8 gp, uint8_t* __restrict__ argmax) {
  extern __shared__ float sm[];
  float* s_in = sm;                               // [3][TI][TIW]
  float* s_w = s_in + 3 * STEM_TI * STEM_TIW;     // [147][20]
  float* s_cv = s_w + STEM_K * STEM_CO;           // [TC*TC][21]
  const int n = blockIdx.z;
  const int src_n = idx ? idx[n] : n;
  const int py0 = blockIdx.y * STEM_TP, px0 = blockIdx.x * STEM_TP;
  const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;  // conv-output origin of the tile
  const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;  // input origin of the tile
  const float* xin = x + (size_t)src_n * 3 * side * side;
  for (int i = threadIdx.x; i < 3 * STEM_TI * STEM_TIW; i += blockDim.x) {
    const int xx = i % STEM_TIW, yy = (i / STEM_TIW) % STEM_TI, c = i / (STEM_TIW * STEM_TI);
    const int gy = iy0 + yy, gx = ix0 + xx;
    float v = 0.f;
    if (xx < STEM_TI && gy >= 0 && gy < side && gx >= 0 && gx < side) v = xin[((size_t)c * side + gy) * side + gx];
    s_in[i] = stem_round<T>(v);
  }
  // weights: PyTorch [20][3][7][7] -> s_w[k][co]
  for (int i = threadIdx.x; i < STEM_K * STEM_CO; i += blockDim.x) {
    const int co = i % STEM_CO, k = i / STEM_CO;
    s_w[i] = stem_round<T>(w[co * STEM_K + k]);
  }
  __syncthreads();
  // conv tile: item = (channel group of 4, conv pixel); lanes of a warp share the channel group
  for (int it = threadIdx.x; it < 5 * STEM_TC * STEM_TC; it += blockDim.x) {
    const int p = it % (STEM_TC * STEM_TC), cg = it / (STEM_TC * STEM_TC);
    const int cy = p / STEM_TC, cx = p % STEM_TC;
    const int gy = cy0 + cy, gx = cx0 + cx;
    float a0, a1, a2, a3;
    if (gy < 0 || gy >= hc || gx < 0 || gx >= hc) {
      a0 = a1 = a2 = a3 = -INFINITY;  // max-pool padding
    } else {
      a0 = b[cg * 4 + 0]; a1 = b[cg * 4 + 1]; a2 = b[cg * 4 + 2]; a3 = b[cg * 4 + 3];
      for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
          const float* row = s_in + (c * STEM_TI + 2 * cy + ky) * STEM_TIW + 2 * cx;
          const float* wr = s_w + ((c * 7 + ky) * 7) * STEM_CO + cg * 4;
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) {
            const float xv = row[kx];
            const float4 wv = *reinterpret_cast<const float4*>(wr + kx * STEM_CO);
            a0 = fmaf(xv, wv.x, a0); a1 = fmaf(xv, wv.y, a1); a2 = fmaf(xv, wv.z, a2); a3 = fmaf(xv, wv.w, a3);
          }
        }
      }
      a0 = stem_round<T>(mil_lrelu(a0)); a1 = stem_round<T>(mil_lrelu(a1));
      a2 = stem_round<T>(mil_lrelu(a2)); a3 = stem_round<T>(mil_lrelu(a3));
    }
    float* o = s_cv + p * 21 + cg * 4;
    o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
  }
  __syncthreads();
  // pooled tile: item = (pooled pixel, 8-channel chunk)
  for (int it = threadIdx.x; it < STEM_TP * STEM_TP * 3; it += blockDim.x) {
    const int cb = it % 3, pp = it / 3;
    const int ly = pp / STEM_TP, lx = pp % STEM_TP;
    const int py = py0 + ly, px = px0 + lx;
    if (py > gp.h || px > gp.w) continue;
    float v[8];
    const long long q = (long long)n * gp.P + (long long)py * gp.wp + px;
    if (py == gp.h || px == gp.w) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cb * 8 + j;
        float best = 0.f;
        int am = 0;
        if (c < STEM_CO) {
          best = -INFINITY;
#pragma unroll
          for (int wy = 0; wy < 3; ++wy)
#pragma unroll
            for (int wx = 0; wx < 3; ++wx) {
              const float cv = s_cv[((2 * ly + wy) * STEM_TC + 2 * lx + wx) * 21 + c];
              if (cv > best) { best = cv; am = wy * 3 + wx; }
            }
          argmax[((size_t)n * gp.h * gp.w + (size_t)py * gp.w + px) * STEM_CO + c] = (uint8_t)am;
        }
        v[j] = best;
      }
    }
    mil_store8(pooled + mil_pf8_off(gp, cb, q), v);
  }
}

int mil_launch_stem_fwd(int dtype, const float* x, const int* idx, int n, int side, const float* w, const float* b,
                        void* pooled, const MilPF8& gp, uint8_t* argmax, cudaStream_t s) {
  const MilGeom g = mil_geom(side);
  MIL_REQUIRE(gp.h == g.h[0] && gp.w == g.h[0] && gp.c == STEM_CO && gp.n == n, "stem_fwd: geometry mismatch");
  const size_t smem = (3 * STEM_TI * STEM_TIW + STEM_K * STEM_CO + STEM_TC * STEM_TC * 21) * sizeof(float);
  dim3 grid((unsigned)mil_cdiv(gp.wp, STEM_TP), (unsigned)mil_cdiv(gp.hp, STEM_TP), (unsigned)n);
  MIL_REQUIRE(n <= 65535, "stem_fwd: at most 65535 tiles per launch (got %d)", n);
  if (dtype == MIL_BF16) {
    MIL_SET_SMEM((stem_fwd_kernel<__nv_bfloat16>), (int)smem);
    stem_fwd_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(x, idx, side, g.hc, w, b, (__nv_bfloat16*)pooled, gp, argmax);
  } else {
    MIL_SET_SMEM((stem_fwd_kernel<float>), (int)smem);
    stem_fwd_kernel<float><<<grid, 256, smem, s>>>(x, idx, side, g.hc, w, b, (float*)pooled, gp, argmax);
  }
  MIL_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// backward: dW[co][c][ky][kx] += sum_{n,py,px} g(n,py,px)[co] * x[n][c][2*yc+ky-3][2*xc+kx-3],
//           (yc,xc) = (2py-1+am/3, 2px-1+am%3) the arg-max conv position;  db[co] += sum g.
// g is the gradient w.r.t. the PRE-activation conv output at the arg-max position, i.e. the pooled
// gradient already multiplied by lrelu'(pooled) (the producer's DGRAD epilogue does that; the sign of
// the pooled value equals the sign of the pre-activation because the slope is positive).
// Persistent blocks: each walks (tile, pooled row) work items, accumulates its (co,ky,kx) x 3 channel
// outputs in registers, and writes ONE partial record; a fixed-order reduction finishes the sum.
// ---------------------------------------------------------------------------------------------------
#define STEM_BWD_THREADS 256
#define STEM_BWD_ITEMS 4  // ceil(20*49 / 256)
#define STEM_BWD_BLOCKS (148 * 2)
template <typename T>
__global__ void __launch_bounds__(STEM_BWD_THREADS)
stem_bwd_kernel(const float* __restrict__ x, const int* __restrict__ idx, int n_tiles, int side,
                const T* __restrict__ g, MilPF8 gp, const uint8_t* __restrict__ argmax, float* __restrict__ partial) {
  extern __shared__ float sm[];
  float* s_g = sm;                                    // [w][20]
  uint8_t* s_am = reinterpret_cast<uint8_t*>(s_g + gp.w * STEM_CO);  // [w][20]
  int it_co[STEM_BWD_ITEMS], it_ky[STEM_BWD_ITEMS], it_kx[STEM_BWD_ITEMS];
  float acc[STEM_BWD_ITEMS][3];
#pragma unroll
  for (int k = 0; k < STEM_BWD_ITEMS; ++k) {
    const int it = threadIdx.x + k * STEM_BWD_THREADS;
    it_co[k] = it < STEM_CO * 49 ? it / 49 : -1;
    it_ky[k] = (it % 49) / 7;
    it_kx[k] = it % 7;
    acc[k][0] = acc[k][1] = acc[k][2] = 0.f;
  }
  float dbacc = 0.f;  // threads 0..19: bias gradient of channel threadIdx.x
  const long long n_work = (long long)n_tiles * gp.h;
  for (long long wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
    const int n = (int)(wk / gp.h), py = (int)(wk % gp.h);
    const int src_n = idx ? idx[n] : n;
    const float* xin = x + (size_t)src_n * 3 * side * side;
    __syncthreads();
    for (int i = threadIdx.x; i < gp.w * STEM_CO; i += blockDim.x) {
      const int c = i % STEM_CO, px = i / STEM_CO;
      const long long q = (long long)n * gp.P + (long long)py * gp.wp + px;
      s_g[i] = mil_to_float(g[mil_pf8_off(gp, c >> 3, q) + (c & 7)]);
      s_am[i] = argmax[((size_t)n * gp.h * gp.w + (size_t)py * gp.w + px) * STEM_CO + c];
    }
    __syncthreads();
    if (threadIdx.x < STEM_CO) {
      for (int px = 0; px < gp.w; ++px) dbacc += s_g[px * STEM_CO + threadIdx.x];
    }
#pragma unroll
    for (int k = 0; k < STEM_BWD_ITEMS; ++k) {
      if (it_co[k] < 0) continue;
      for (int px = 0; px < gp.w; ++px) {
        const float gv = s_g[px * STEM_CO + it_co[k]];
        const int am = s_am[px * STEM_CO + it_co[k]];
        const int yc = 2 * py - 1 + am / 3, xc = 2 * px - 1 + am % 3;
        const int iy = 2 * yc + it_ky[k] - 3, ix = 2 * xc + it_kx[k] - 3;
        if (iy >= 0 && iy < side && ix >= 0 && ix < side) {
          const float* p = xin + (size_t)iy * side + ix;
          acc[k][0] = fmaf(gv, stem_round<T>(p[0]), acc[k][0]);
          acc[k][1] = fmaf(gv, stem_round<T>(p[(size_t)side * side]), acc[k][1]);
          acc[k][2] = fmaf(gv, stem_round<T>(p[(size_t)2 * side * side]), acc[k][2]);
        }
      }
    }
  }
  // record layout = PyTorch parameter layout [20][3][7][7] followed by [20] bias sums
  float* rec = partial + (size_t)blockIdx.x * (STEM_CO * STEM_K + STEM_CO);
#pragma unroll
  for (int k = 0; k < STEM_BWD_ITEMS; ++k) {
    if (it_co[k] < 0) continue;
    for (int c = 0; c < 3; ++c) rec[it_co[k] * STEM_K + c * 49 + it_ky[k] * 7 + it_kx[k]] = acc[k][c];
  }
  if (threadIdx.x < STEM_CO) rec[STEM_CO * STEM_K + threadIdx.x] = dbacc;
}

size_t mil_stem_bwd_partial_floats() { return (size_t)STEM_BWD_BLOCKS * (STEM_CO * STEM_K + STEM_CO); }

int mil_launch_stem_bwd(int dtype, const float* x, const int* idx, int n, int side, const void* g, const MilPF8& gp,
                        const uint8_t* argmax, float* partial, float* dw, float* db, cudaStream_t s) {
  const size_t smem = (size_t)gp.w * STEM_CO * sizeof(float) + mil_rup((size_t)gp.w * STEM_CO, 16);
  if (dtype == MIL_BF16)
    stem_bwd_kernel<__nv_bfloat16><<<STEM_BWD_BLOCKS, STEM_BWD_THREADS, smem, s>>>(x, idx, n, side, (const __nv_bfloat16*)g, gp, argmax, partial);
  else
    stem_bwd_kernel<float><<<STEM_BWD_BLOCKS, STEM_BWD_THREADS, smem, s>>>(x, idx, n, side, (const float*)g, gp, argmax, partial);
  MIL_LAUNCH_OK();
  const long long rec = STEM_CO * STEM_K + STEM_CO;
  MIL_TRY(mil_launch_reduce_partials(partial, STEM_BWD_BLOCKS, rec, dw, STEM_CO * STEM_K, s));
  return mil_launch_reduce_partials(partial + STEM_CO * STEM_K, STEM_BWD_BLOCKS, rec, db, STEM_CO, s);
}
