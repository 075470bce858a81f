// Stem on the tensor cores (bf16 mode).  Reference: gbm/model.py:24-26,51-53 -- conv 7x7 / stride 2 / pad 3,
// 3 -> 20 channels, + bias, LeakyReLU(0.1), max-pool 3x3 / stride 2 / pad 1; backward = weight / bias gradient
// only (the bag is detached, gbm/model.py:194,196).
//
// A 7x7 stride-2 convolution over 3 channels is a 4x4 stride-1 convolution over the 12 channels of the
// space-to-depth image  xs[(c,py,px)][Y][X] = x[c][2Y+py][2X+px]  with window offsets a,b in {-2..1}
// (ky = 2a+py+3, kx = 2b+px+3; the 15 % of taps that fall outside 0..6 carry zero weights).  In PF8 with a
// 2-pixel shared halo that is exactly the shape conv_tc / wgrad_tc handle: 16 taps x 2 chunks = 16 MMAs (K=16).
//
//   forward : s2d (fp32 NCHW -> bf16 PF8, fused with the train-mode tile gather)  ->  conv_tc (+bias, LeakyReLU)
//             ->  pool (3x3/2 max, first maximum wins like ATen, 1-byte arg-max per pooled element)
//   backward: scatter (pooled gradient -> dense conv-resolution gradient through the arg-max, as a gather so that
//             it is deterministic)  ->  wgrad_tc (16 taps)  ->  fixed-order reduction into the [20][3][7][7] layout
#include <algorithm>

#include "mil_common.cuh"
#include "mil_conv_tc.cuh"

#define STC_CO 20
#define STC_CI 12

// ---- 1. space-to-depth --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stem_s2d_kernel(const float* __restrict__ x, const int* __restrict__ idx, int side, __nv_bfloat16* __restrict__ xs,
                MilPF8 g) {
  const long long total = 2 * g.Q;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(i / g.Q);
    const long long q = i - (long long)cb * g.Q;
    const int n = (int)(q / g.P);
    const int r = (int)(q - (long long)n * g.P);
    const int Y = r / g.wp, X = r - Y * g.wp;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (Y < g.h && X < g.w) {
      const int src_n = idx ? idx[n] : n;
      const float* xin = x + (size_t)src_n * 3 * side * side;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cc = cb * 8 + j;  // (c, py, px)
        if (cc < STC_CI) {
          const int c = cc >> 2, py = (cc >> 1) & 1, px = cc & 1;
          const int iy = 2 * Y + py, ix = 2 * X + px;
          if (iy < side && ix < side) v[j] = xin[((size_t)c * side + iy) * side + ix];
        }
      }
    }
    mil_store8(xs + mil_pf8_off(g, cb, q), v);
  }
}

// ---- weights: w[20][3][7][7] -> wp[tap(a,b)][kin_pad 16][nout_pad 24] (fp32; mil_launch_pack_tc finishes) -----
__global__ void stem_pack_w_kernel(const float* __restrict__ w, float* __restrict__ wp) {
  const int total = 16 * 16 * 24;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % 24, cc = (i / 24) % 16, t = i / (24 * 16);
    float v = 0.f;
    if (co < STC_CO && cc < STC_CI) {
      const int c = cc >> 2, py = (cc >> 1) & 1, px = cc & 1;
      const int a = t / 4 - 2, b = t % 4 - 2;
      const int ky = 2 * a + py + 3, kx = 2 * b + px + 3;
      if (ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = w[((co * 3 + c) * 7 + ky) * 7 + kx];
    }
    wp[i] = v;
  }
}

// ---- 3. max-pool 3x3 / stride 2 / pad 1 over the (already activated) conv map ---------------------------------
__global__ void __launch_bounds__(256)
stem_pool_kernel(const __nv_bfloat16* __restrict__ cv, MilPF8 gc, __nv_bfloat16* __restrict__ pooled, MilPF8 gp,
                 uint8_t* __restrict__ argmax) {
  const long long total = (long long)gp.cb * gp.Q;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(i / gp.Q);
    const long long q = i - (long long)cb * gp.Q;
    const int n = (int)(q / gp.P);
    const int r = (int)(q - (long long)n * gp.P);
    const int py = r / gp.wp, px = r - py * gp.wp;
    float best[8];
    int am[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = 0.f; am[j] = 0; }
    if (py < gp.h && px < gp.w) {
#pragma unroll
      for (int j = 0; j < 8; ++j) best[j] = -INFINITY;
#pragma unroll
      for (int wy = 0; wy < 3; ++wy) {
#pragma unroll
        for (int wx = 0; wx < 3; ++wx) {
          const int cy = 2 * py - 1 + wy, cx = 2 * px - 1 + wx;
          if (cy >= 0 && cy < gc.h && cx >= 0 && cx < gc.w) {
            float v[8];
            mil_load8(cv + mil_pf8_off(gc, cb, (long long)n * gc.P + (long long)cy * gc.wp + cx), v);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (v[j] > best[j]) { best[j] = v[j]; am[j] = wy * 3 + wx; }
          }
        }
      }
      uint8_t* ap = argmax + ((size_t)n * gp.h * gp.w + (size_t)py * gp.w + px) * STC_CO + cb * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (cb * 8 + j < STC_CO) ap[j] = (uint8_t)am[j];
        else best[j] = 0.f;
      }
    }
    mil_store8(pooled + mil_pf8_off(gp, cb, q), best);
  }
}

// ---- 4. backward scatter as a gather: dY(Y,X)[c] = sum of g(py,px)[c] over the pooled windows whose arg-max is (Y,X)
__global__ void __launch_bounds__(256)
stem_unpool_kernel(const __nv_bfloat16* __restrict__ g, MilPF8 gp, const uint8_t* __restrict__ argmax,
                   __nv_bfloat16* __restrict__ dy, MilPF8 gc) {
  const long long total = (long long)gc.cb * gc.Q;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(i / gc.Q);
    const long long q = i - (long long)cb * gc.Q;
    const int n = (int)(q / gc.P);
    const int r = (int)(q - (long long)n * gc.P);
    const int Y = r / gc.wp, X = r - Y * gc.wp;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (Y < gc.h && X < gc.w) {
      // pooled windows containing conv row Y: py with 2py-1 <= Y <= 2py+1
      const int py0 = Y >> 1, py1 = (Y + 1) >> 1, px0 = X >> 1, px1 = (X + 1) >> 1;
      for (int py = py0; py <= py1; ++py) {
        if (py >= gp.h) continue;
        const int wy = Y - (2 * py - 1);
        for (int px = px0; px <= px1; ++px) {
          if (px >= gp.w) continue;
          const int wx = X - (2 * px - 1);
          const int want = wy * 3 + wx;
          const uint8_t* ap = argmax + ((size_t)n * gp.h * gp.w + (size_t)py * gp.w + px) * STC_CO + cb * 8;
          float gv[8];
          mil_load8(g + mil_pf8_off(gp, cb, (long long)n * gp.P + (long long)py * gp.wp + px), gv);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (cb * 8 + j < STC_CO && ap[j] == want) acc[j] += gv[j];
        }
      }
    }
    mil_store8(dy + mil_pf8_off(gc, cb, q), acc);
  }
}

// ---- 5. reduction of the wgrad_tc partial records into the PyTorch layout ------------------------------------
// record: [tap 16][kin_pad 16][cout_pad 24] + [24] bias sums
__global__ void stem_reduce_kernel(const float* __restrict__ partial, int nblk, long long stride,
                                   float* __restrict__ dw, float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < STC_CO * 147) {
    const int kx = i % 7, ky = (i / 7) % 7, c = (i / 49) % 3, co = i / 147;
    const int py = (ky + 1) & 1, px = (kx + 1) & 1;           // ky - 3 = 2a + py
    const int a = (ky - 3 - py) / 2, b = (kx - 3 - px) / 2;   // exact: ky-3-py is even
    const int t = (a + 2) * 4 + (b + 2), cc = c * 4 + py * 2 + px;
    const size_t src = ((size_t)t * 16 + cc) * 24 + co;
    float acc = 0.f;
    for (int k = 0; k < nblk; ++k) acc += partial[(size_t)k * stride + src];
    dw[i] += acc;
  } else if (i < STC_CO * 147 + STC_CO) {
    const int co = i - STC_CO * 147;
    const size_t src = (size_t)16 * 16 * 24 + co;
    float acc = 0.f;
    for (int k = 0; k < nblk; ++k) acc += partial[(size_t)k * stride + src];
    db[co] += acc;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
MilPF8 mil_stem_tc_geom_in(int n, int side) {
  const int hc = (side - 1) / 2 + 1;
  return mil_pf8p(n, STC_CI, hc, hc, 2);
}
MilPF8 mil_stem_tc_geom_conv(int n, int side) {
  const int hc = (side - 1) / 2 + 1;
  return mil_pf8p(n, STC_CO, hc, hc, 2);
}
size_t mil_stem_tc_wpack_floats() { return 16 * 16 * 24; }
size_t mil_stem_tc_wtc_bytes() {
  MilTcShape sh;
  mil_tc_shape(STC_CI, STC_CO, 7, &sh);
  return mil_tc_wpack_bytes(sh);
}
size_t mil_stem_tc_partial_floats(int n, int side) {
  return mil_wgrad_tc_partial_floats(mil_stem_tc_geom_in(n, side), mil_stem_tc_geom_conv(n, side), 7);
}

static int grid_for(long long work) { return (int)std::max<long long>(1, std::min<long long>(mil_cdiv(work, 256), 148 * 16)); }

int mil_launch_stem_tc_fwd(const float* x, const int* idx, int n, int side, const float* w, const float* b, void* xs,
                           void* convout, float* wp, void* wtc, void* pooled, const MilPF8& gp, uint8_t* argmax,
                           cudaStream_t s) {
  const MilPF8 gi = mil_stem_tc_geom_in(n, side), gc = mil_stem_tc_geom_conv(n, side);
  MIL_REQUIRE(gp.h == (gc.h - 1) / 2 + 1 && gp.c == STC_CO && gp.n == n, "stem_tc_fwd: geometry mismatch");
  stem_s2d_kernel<<<grid_for(2 * gi.Q), 256, 0, s>>>(x, idx, side, (__nv_bfloat16*)xs, gi);
  MIL_LAUNCH_OK();
  stem_pack_w_kernel<<<24, 256, 0, s>>>(w, wp);
  MIL_LAUNCH_OK();
  MilTcShape sh;
  MIL_TRY(mil_tc_shape(STC_CI, STC_CO, 7, &sh));
  MIL_TRY(mil_launch_pack_tc(wp, wtc, sh, s));
  MIL_TRY(mil_launch_conv_tc(0, xs, gi, wtc, sh, b, nullptr, nullptr, convout, gc, MIL_EPI_FWD, 0, s));
  stem_pool_kernel<<<grid_for((long long)gp.cb * gp.Q), 256, 0, s>>>((const __nv_bfloat16*)convout, gc,
                                                                      (__nv_bfloat16*)pooled, gp, argmax);
  MIL_LAUNCH_OK();
  return 0;
}

int mil_launch_stem_tc_bwd(const void* xs, int n, int side, const void* g, const MilPF8& gp, const uint8_t* argmax,
                           void* dy, float* partial, float* dw, float* db, cudaStream_t s) {
  const MilPF8 gi = mil_stem_tc_geom_in(n, side), gc = mil_stem_tc_geom_conv(n, side);
  stem_unpool_kernel<<<grid_for((long long)gc.cb * gc.Q), 256, 0, s>>>((const __nv_bfloat16*)g, gp, argmax,
                                                                        (__nv_bfloat16*)dy, gc);
  MIL_LAUNCH_OK();
  int ctas;
  long long rec;
  MIL_TRY(mil_launch_wgrad_tc_partials(xs, gi, dy, gc, partial, 7, &ctas, &rec, s));
  stem_reduce_kernel<<<(int)mil_cdiv(STC_CO * 147 + STC_CO, 128), 128, 0, s>>>(partial, ctas, rec, dw, db);
  MIL_LAUNCH_OK();
  return 0;
}
