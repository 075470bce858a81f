// CUDA-core (FFMA) convolution kernels on the PF8 layout: forward, data-gradient and weight-gradient
// for 3x3 / 1x1, stride 1 / 2.  They are (a) the whole path in the fp32 check mode, (b) the kernels the
// bf16 mode uses for the shapes the tcgen05 implicit-GEMM kernels (mil_conv_tc.cu) do not cover, and
// (c) the on-device cross-check for those kernels.
//
// Reference semantics: nnBlocks.py:157-189 (BasicResBlock: conv3x3(+bias) -> LeakyReLU(0.1) ->
// conv3x3(+bias) -> += identity / 1x1-stride-2 projection -> LeakyReLU(0.1)); backward is autograd of
// the same (gbm/classify_combined.py:447).
#include <algorithm>

#include "mil_common.cuh"

// ---------------------------------------------------------------------------------------------------
// forward / dgrad.  One thread = one output flat pixel x one 8-channel output chunk (blockIdx.y).
//   normal     : out(n,y,x)[co] = sum_{dy,dx,ci} in(n, y*S+dy-pad, x*S+dx-pad)[ci] * w[co][ci][dy][dx]
//   transposed : out(n,y,x)[ci] = sum_{dy,dx,co : (y+pad-dy)%S==0, (x+pad-dx)%S==0}
//                                 in(n, (y+pad-dy)/S, (x+pad-dx)/S)[co] * w[co][ci][dy][dx]
// The zero halo of PF8 makes every access in-bounds and every out-of-image tap read exact zeros.
// ---------------------------------------------------------------------------------------------------
template <typename T, int KS, int STRIDE, bool TRANSPOSED>
__global__ void __launch_bounds__(128)
conv_direct_kernel(const T* __restrict__ x, MilPF8 gi, const float* __restrict__ wp, const float* __restrict__ bias,
                   const T* res, const T* __restrict__ act, T* out, MilPF8 go, int epi) {
  extern __shared__ float ws[];  // [KS*KS][cin_pad][8] slice of the packed weights for this output chunk
  constexpr int TAPS = KS * KS;
  constexpr int PAD = KS / 2;
  const int cinp = gi.cb * 8, coutp = go.cb * 8;
  const int co8 = blockIdx.y;
  for (int i = threadIdx.x; i < TAPS * cinp * 8; i += blockDim.x) {
    const int j = i & 7, a = i >> 3;  // a = tap*cinp + ci
    ws[i] = wp[(size_t)a * coutp + co8 * 8 + j];
  }
  __syncthreads();
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= go.Q) return;
  const int n = (int)(q / go.P);
  const int r = (int)(q % go.P);
  const int y = r / go.wp, xo = r % go.wp;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const bool is_pad = (y == go.h) || (xo == go.w);
  if (!is_pad) {
#pragma unroll
    for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
      for (int dx = 0; dx < KS; ++dx) {
        int iy, ix;
        if (!TRANSPOSED) {
          iy = y * STRIDE + dy - PAD;
          ix = xo * STRIDE + dx - PAD;
        } else {
          const int ty = y + PAD - dy, tx = xo + PAD - dx;
          if (STRIDE == 2) {
            if ((ty & 1) || (tx & 1)) continue;
            iy = ty >> 1;
            ix = tx >> 1;
            if (iy > gi.h || ix > gi.w) continue;  // cannot happen for the geometries used; keeps reads in-bounds
          } else {
            iy = ty;
            ix = tx;
          }
        }
        const long long qi = (long long)n * gi.P + (long long)iy * gi.wp + ix;
        const float* wt = ws + (dy * KS + dx) * cinp * 8;
        for (int cbi = 0; cbi < gi.cb; ++cbi) {
          float v[8];
          mil_load8(x + mil_pf8_off(gi, cbi, qi), v);
#pragma unroll
          for (int ci = 0; ci < 8; ++ci) {
            const float4 w0 = *reinterpret_cast<const float4*>(wt + (cbi * 8 + ci) * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(wt + (cbi * 8 + ci) * 8 + 4);
            acc[0] = fmaf(v[ci], w0.x, acc[0]); acc[1] = fmaf(v[ci], w0.y, acc[1]);
            acc[2] = fmaf(v[ci], w0.z, acc[2]); acc[3] = fmaf(v[ci], w0.w, acc[3]);
            acc[4] = fmaf(v[ci], w1.x, acc[4]); acc[5] = fmaf(v[ci], w1.y, acc[5]);
            acc[6] = fmaf(v[ci], w1.z, acc[6]); acc[7] = fmaf(v[ci], w1.w, acc[7]);
          }
        }
      }
    }
    const long long o = mil_pf8_off(go, co8, q);
    float rv[8], av[8];
    if (res != nullptr) {
      mil_load8(res + o, rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += rv[j];
    }
    if (bias != nullptr && epi != MIL_EPI_DGRAD) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (co8 * 8 + j < go.c) acc[j] += bias[co8 * 8 + j];
    }
    if (epi == MIL_EPI_FWD) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = mil_lrelu(acc[j]);
    } else if (epi == MIL_EPI_DGRAD) {
      mil_load8(act + o, av);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= mil_lrelu_grad(av[j]);
    }
  }
  mil_store8(out + mil_pf8_off(go, co8, q), acc);
}

template <typename T>
static int launch_conv_direct_t(int transposed, const T* x, const MilPF8& gi, const float* wp, const float* bias,
                                const T* res, const T* act, T* out, const MilPF8& go, int ks, int stride, int epi,
                                cudaStream_t s) {
  dim3 grid((unsigned)mil_cdiv(go.Q, 128), (unsigned)go.cb);
  const size_t smem = (size_t)ks * ks * gi.cb * 8 * 8 * sizeof(float);
#define MIL_LAUNCH(KS, ST, TR)                                                                                 \
  conv_direct_kernel<T, KS, ST, TR><<<grid, 128, smem, s>>>(x, gi, wp, bias, res, act, out, go, epi)
  if (ks == 3 && stride == 1 && !transposed) MIL_LAUNCH(3, 1, false);
  else if (ks == 3 && stride == 1 && transposed) MIL_LAUNCH(3, 1, true);
  else if (ks == 3 && stride == 2 && !transposed) MIL_LAUNCH(3, 2, false);
  else if (ks == 3 && stride == 2 && transposed) MIL_LAUNCH(3, 2, true);
  else if (ks == 1 && stride == 2 && !transposed) MIL_LAUNCH(1, 2, false);
  else if (ks == 1 && stride == 2 && transposed) MIL_LAUNCH(1, 2, true);
  else if (ks == 1 && stride == 1 && !transposed) MIL_LAUNCH(1, 1, false);
  else if (ks == 1 && stride == 1 && transposed) MIL_LAUNCH(1, 1, true);
  else {
    mil_set_error("conv_direct: unsupported ks=%d stride=%d", ks, stride);
    return 2;
  }
#undef MIL_LAUNCH
  MIL_LAUNCH_OK();
  return 0;
}

int mil_launch_conv_direct(int dtype, int transposed, const void* x, const MilPF8& gi, const float* wp,
                           const float* bias, const void* res, const void* act, void* out, const MilPF8& go, int ks,
                           int stride, int epi, cudaStream_t s) {
  MIL_REQUIRE(gi.n == go.n, "conv_direct: batch mismatch %d vs %d", gi.n, go.n);
  MIL_REQUIRE(epi != MIL_EPI_DGRAD || act != nullptr, "conv_direct: DGRAD epilogue needs the activation tensor");
  if (dtype == MIL_BF16)
    return launch_conv_direct_t<__nv_bfloat16>(transposed, (const __nv_bfloat16*)x, gi, wp, bias,
                                               (const __nv_bfloat16*)res, (const __nv_bfloat16*)act,
                                               (__nv_bfloat16*)out, go, ks, stride, epi, s);
  return launch_conv_direct_t<float>(transposed, (const float*)x, gi, wp, bias, (const float*)res,
                                     (const float*)act, (float*)out, go, ks, stride, epi, s);
}

// ---------------------------------------------------------------------------------------------------
// weight gradient (+ bias gradient), split over flat-pixel ranges:
//   dw[co][ci][dy][dx] = sum_{n,y,x} dz(n,y,x)[co] * in(n, y*S+dy-pad, x*S+dx-pad)[ci] ;  db[co] = sum dz
// Block (range r, output chunk co8): thread owns (tap, ci) pairs, 8 accumulators (the 8 co of the chunk)
// each.  Partials [r][tap][cin_pad][cout_pad] (+[cout_pad] bias sums) are summed in fixed order by
// mil_launch_reduce_conv_w -> deterministic.
// ---------------------------------------------------------------------------------------------------
#define MIL_WG_THREADS 256
#define MIL_WG_MAXPAIR 3  // ceil(9*80 / 256)
template <typename T, int KS, int STRIDE>
__global__ void __launch_bounds__(MIL_WG_THREADS)
wgrad_direct_kernel(const T* __restrict__ x, MilPF8 gi, const T* __restrict__ dz, MilPF8 go,
                    float* __restrict__ partial, long long stride_rec, long long px_per_block) {
  constexpr int TAPS = KS * KS;
  constexpr int PAD = KS / 2;
  const int cinp = gi.cb * 8, coutp = go.cb * 8;
  const int co8 = blockIdx.y;
  const int npair = TAPS * cinp;
  int p_shift[MIL_WG_MAXPAIR];  // tap shift in input flat pixels (for STRIDE 1) or dy,dx packed (for STRIDE 2)
  int p_ci[MIL_WG_MAXPAIR];
  float acc[MIL_WG_MAXPAIR][8];
  float dbacc[8];
#pragma unroll
  for (int k = 0; k < MIL_WG_MAXPAIR; ++k) {
    const int pr = threadIdx.x + k * MIL_WG_THREADS;
    const int tap = pr < npair ? pr / cinp : 0;
    p_ci[k] = pr < npair ? pr % cinp : -1;
    const int dy = tap / KS, dx = tap % KS;
    p_shift[k] = (dy - PAD) * gi.wp + (dx - PAD);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) dbacc[j] = 0.f;

  const long long q0 = blockIdx.x * px_per_block;
  const long long q1 = min(q0 + px_per_block, go.Q);
  if (q0 < q1) {
    int n = (int)(q0 / go.P);
    int r = (int)(q0 % go.P);
    int y = r / go.wp, xo = r % go.wp;
    for (long long q = q0; q < q1; ++q) {
      if (y < go.h && xo < go.w) {
        float d[8];
        mil_load8(dz + mil_pf8_off(go, co8, q), d);
        // input flat pixel of tap (pad,pad) i.e. the window centre / origin
        const long long qc = (long long)n * gi.P + (long long)(y * STRIDE) * gi.wp + xo * STRIDE;
#pragma unroll
        for (int k = 0; k < MIL_WG_MAXPAIR; ++k) {
          if (p_ci[k] >= 0) {
            const long long qi = qc + p_shift[k];
            const float xv = mil_to_float(x[mil_pf8_off(gi, p_ci[k] >> 3, qi) + (p_ci[k] & 7)]);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[k][j] = fmaf(xv, d[j], acc[k][j]);
          }
        }
        if (threadIdx.x == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dbacc[j] += d[j];
        }
      }
      if (++xo == go.wp) {
        xo = 0;
        if (++y == go.hp) { y = 0; ++n; }
      }
    }
  }
  float* rec = partial + (size_t)blockIdx.x * stride_rec;
#pragma unroll
  for (int k = 0; k < MIL_WG_MAXPAIR; ++k) {
    const int pr = threadIdx.x + k * MIL_WG_THREADS;
    if (pr < npair) {
#pragma unroll
      for (int j = 0; j < 8; ++j) rec[(size_t)pr * coutp + co8 * 8 + j] = acc[k][j];
    }
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) rec[(size_t)npair * coutp + co8 * 8 + j] = dbacc[j];
  }
}

int mil_wgrad_direct_blocks(const MilPF8& go) {
  // enough ranges to fill 148 SMs a few times over, but at least ~512 pixels of work per block
  long long b = std::min<long long>(148 * 4, std::max<long long>(1, go.Q / 512));
  return (int)b;
}
size_t mil_wgrad_direct_partial_floats(const MilPF8& gi, const MilPF8& go, int ks) {
  const size_t rec = (size_t)ks * ks * gi.cb * 8 * go.cb * 8 + go.cb * 8;
  return rec * (size_t)mil_wgrad_direct_blocks(go);
}

template <typename T>
static int launch_wgrad_direct_t(const T* x, const MilPF8& gi, const T* dz, const MilPF8& go, float* partial,
                                 float* dw, float* db, int ks, int stride, cudaStream_t s) {
  const int nblk = mil_wgrad_direct_blocks(go);
  const long long ppb = mil_cdiv(go.Q, nblk);
  const long long rec = (long long)ks * ks * gi.cb * 8 * go.cb * 8 + go.cb * 8;
  dim3 grid((unsigned)nblk, (unsigned)go.cb);
  if (ks == 3 && stride == 1) wgrad_direct_kernel<T, 3, 1><<<grid, MIL_WG_THREADS, 0, s>>>(x, gi, dz, go, partial, rec, ppb);
  else if (ks == 3 && stride == 2) wgrad_direct_kernel<T, 3, 2><<<grid, MIL_WG_THREADS, 0, s>>>(x, gi, dz, go, partial, rec, ppb);
  else if (ks == 1 && stride == 2) wgrad_direct_kernel<T, 1, 2><<<grid, MIL_WG_THREADS, 0, s>>>(x, gi, dz, go, partial, rec, ppb);
  else if (ks == 1 && stride == 1) wgrad_direct_kernel<T, 1, 1><<<grid, MIL_WG_THREADS, 0, s>>>(x, gi, dz, go, partial, rec, ppb);
  else {
    mil_set_error("wgrad_direct: unsupported ks=%d stride=%d", ks, stride);
    return 2;
  }
  MIL_LAUNCH_OK();
  return mil_launch_reduce_conv_w(partial, nblk, rec, dw, db, go.c, gi.c, ks, s);
}

int mil_launch_wgrad_direct(int dtype, const void* x, const MilPF8& gi, const void* dz, const MilPF8& go,
                            float* partial, float* dw, float* db, int ks, int stride, cudaStream_t s) {
  MIL_REQUIRE(gi.n == go.n, "wgrad_direct: batch mismatch %d vs %d", gi.n, go.n);
  MIL_REQUIRE(ks * ks * gi.cb * 8 <= MIL_WG_MAXPAIR * MIL_WG_THREADS, "wgrad_direct: cin %d too wide", gi.c);
  if (dtype == MIL_BF16)
    return launch_wgrad_direct_t<__nv_bfloat16>((const __nv_bfloat16*)x, gi, (const __nv_bfloat16*)dz, go, partial,
                                                dw, db, ks, stride, s);
  return launch_wgrad_direct_t<float>((const float*)x, gi, (const float*)dz, go, partial, dw, db, ks, stride, s);
}
