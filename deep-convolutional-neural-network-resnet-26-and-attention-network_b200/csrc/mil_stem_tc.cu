// Stem on the tensor cores (bf16 mode).  Reference: gbm/model.py:24-26,51-53 -- conv 7x7 / stride 2 / pad 3,
// 3 -> 20 channels, + bias, LeakyReLU(0.1), max-pool 3x3 / stride 2 / pad 1; backward = weight / bias gradient
// only (the bag is detached, gbm/model.py:194,196).
//
// Space-to-depth by FOUR turns the stem into an ordinary 3x3 / stride-1 convolution at the POOLED resolution:
//   input   xs[(c, ry, rx)][Y][X] = x[c][4Y+ry][4X+rx]                       3*16 = 48 channels
//   output  cv[(co, a, b)][Y][X]  = conv1(x)[co][2Y+a][2X+b]                 20*4 = 80 channels (4 phases)
//   weights W4[(dy,dx)][(c,ry,rx)][(co,a,b)] = w[co][c][4dy+ry-2a+3][4dx+rx-2b+3]   (zero outside 0..6)
// because conv1 row 2Y+a reads input rows 4Y + (2a+ky-3) and 2a+ky-3 = 4dy+ry has a unique (dy in -1..1, ry in 0..3).
// So the stem runs on conv_tc / wgrad_tc exactly like a layer (48 -> 80 channels, N = 80 per MMA instead of the
// 32 a stride-2 formulation gives), the window halo is one row of 57 pixels instead of two rows of 114, and the
// max-pool becomes a reduction over (neighbour pixel, phase) pairs.
//
//   forward : s2d4 (fp32 NCHW -> bf16 PF8, fused with the train-mode tile gather)  ->  conv_tc (+bias, LeakyReLU)
//             ->  pool (first maximum in window scan order wins like ATen; 1-byte arg-max per pooled element)
//   backward: unpool (pooled gradient -> the 4-phase conv gradient through the arg-max, as a gather => deterministic)
//             ->  wgrad_tc (9 taps, 48 x 80)  ->  fixed-order reduction + fold back into the [20][3][7][7] layout
#include <algorithm>
#include <cstdlib>

#include "mil_common.cuh"
#include "mil_conv_tc.cuh"
#include "mil_stem_unpool.cuh"

#define STC_CO 20
#define STC_CI 48   // 3 * 4 * 4
#define STC_CO4 80  // 20 * 2 * 2

// ---- 1. space-to-depth by 4 ---------------------------------------------------------------------------------
// TIn = float: the reference's normalised fp32 tiles;  TIn = uint8_t: raw 8-bit tiles, normalised on the fly exactly
// like the reference's transform ToTensor() + Normalize(0.5, 0.5) (RoiBuilder.py:199-202): (u/255 - 0.5) / 0.5
__device__ __forceinline__ float stem_norm(float v) { return v; }
__device__ __forceinline__ float stem_norm(uint8_t v) { return ((float)v / 255.0f - 0.5f) / 0.5f; }

template <typename TIn> struct StemIsU8 { static constexpr bool value = false; };
template <> struct StemIsU8<uint8_t> { static constexpr bool value = true; };

template <typename TIn>
__global__ void __launch_bounds__(256)
stem_s2d4_kernel(const TIn* __restrict__ x, const int* __restrict__ idx, int side, __nv_bfloat16* __restrict__ xs,
                 MilPF8 g) {
  // one thread per (pixel, chunk k of 6): chunk k = channels (c, ry, rx) with c = k/2, ry = 2(k&1) + {0,1}, rx = 0..3
  // -> two 4-element row segments in (one vector load each when aligned), ONE 16-byte chunk out
  // 8-bit tiles: the normalised value is rounded to bf16 on its way out, and for all 256 byte values
  // bf16(fma(u, 2/255, -1)) == bf16(((float)u / 255 - 0.5) / 0.5) (checked exhaustively, tests/test_gpu_parity.py) -- one FMA
  // instead of an IEEE division per element (with eight divisions per thread the kernel was bound by its instruction
  // stream and SLOWER than the fp32 form, 908 vs 623 us at 4096 tiles, although it reads a quarter of the bytes)
  // grid = (image, pixel block, chunk): no 64-bit division per thread (with a flat index the kernel executed ~175
  // instructions per thread, two thirds of them index arithmetic, and was bound by its instruction stream:
  // smsp__issue_active 65-68 %, DRAM 33 % for 8-bit tiles)
  const bool vec = (side & 3) == 0;
  const int n = blockIdx.x, k = blockIdx.z;
  for (int r = blockIdx.y * blockDim.x + threadIdx.x; r < (int)g.P; r += gridDim.y * blockDim.x) {
    const long long q = (long long)n * g.P + r;
    const int Y = r / g.wp, X = r - Y * g.wp;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (StemIsU8<TIn>::value) {
      uint32_t w[4] = {0u, 0u, 0u, 0u};  // the chunk as packed bf16 pairs: word h * 2 + (rx >> 1)
      if (Y < g.h && X < g.w) {
        const int c = k >> 1;
        const int src_n = idx ? idx[n] : n;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int iy = 4 * Y + 2 * (k & 1) + h;
          if (iy < side) {
            const uint8_t* row = reinterpret_cast<const uint8_t*>(x) + (((size_t)src_n * 3 + c) * side + iy) * side + 4 * X;
            uint32_t f = 0u;
            int valid = 4;
            if (vec) f = *reinterpret_cast<const uint32_t*>(row);
            else {
              valid = min(4, side - 4 * X);
              for (int rx = 0; rx < valid; ++rx) f |= (uint32_t)row[rx] << (8 * rx);
            }
            float e[4];
#pragma unroll
            for (int rx = 0; rx < 4; ++rx)
              e[rx] = rx < valid ? fmaf((float)((f >> (8 * rx)) & 0xFFu), 2.0f / 255.0f, -1.0f) : 0.f;
            const __nv_bfloat162 p0 = __floats2bfloat162_rn(e[0], e[1]), p1 = __floats2bfloat162_rn(e[2], e[3]);
            w[h * 2] = *reinterpret_cast<const uint32_t*>(&p0);
            w[h * 2 + 1] = *reinterpret_cast<const uint32_t*>(&p1);
          }
        }
      }
      *reinterpret_cast<uint4*>(xs + mil_pf8_off(g, k, q)) = make_uint4(w[0], w[1], w[2], w[3]);
      continue;
    }
    if (Y < g.h && X < g.w) {
      const int c = k >> 1;
      const int src_n = idx ? idx[n] : n;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int iy = 4 * Y + 2 * (k & 1) + h;
        if (iy < side) {
          const TIn* row = x + (((size_t)src_n * 3 + c) * side + iy) * side + 4 * X;
          if (vec) {
            struct alignas(4 * sizeof(TIn)) Quad { TIn e[4]; };
            const Quad f = *reinterpret_cast<const Quad*>(row);
#pragma unroll
            for (int rx = 0; rx < 4; ++rx) v[h * 4 + rx] = stem_norm(f.e[rx]);
          } else {
#pragma unroll
            for (int rx = 0; rx < 4; ++rx)
              if (4 * X + rx < side) v[h * 4 + rx] = stem_norm(row[rx]);
          }
        }
      }
    }
    // channel cc = (c*4 + ry)*4 + rx  ->  chunk cc/8 = k, lane (ry&1)*4 + rx
    mil_store8(xs + mil_pf8_off(g, k, q), v);
  }
}

// ---- weights: w[20][3][7][7] -> wp[tap 9][kin 48][nout 80] (fp32; mil_launch_pack_tc finishes), bias4[80] ------
__global__ void stem_pack_w4_kernel(const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ wp,
                                    float* __restrict__ bias4) {
  const int total = 9 * STC_CI * STC_CO4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co4 = i % STC_CO4, cc = (i / STC_CO4) % STC_CI, t = i / (STC_CO4 * STC_CI);
    const int co = co4 >> 2, a = (co4 >> 1) & 1, bb = co4 & 1;
    const int c = cc >> 4, ry = (cc >> 2) & 3, rx = cc & 3;
    const int dy = t / 3 - 1, dx = t % 3 - 1;
    const int ky = 4 * dy + ry - 2 * a + 3, kx = 4 * dx + rx - 2 * bb + 3;
    float v = 0.f;
    if (ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = w[((co * 3 + c) * 7 + ky) * 7 + kx];
    wp[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < STC_CO4) bias4[threadIdx.x] = b[threadIdx.x >> 2];
}

// ---- 3. max-pool 3x3 / stride 2 / pad 1 of the conv1 map held as 4 phases ----------------------------------------
// pooled (py,px) covers conv rows 2py-1, 2py, 2py+1 = (py-1, a=1), (py, a=0), (py, a=1); same for columns: the
// nine window positions live in FOUR neighbouring pixels of the phase map.  cv channel = co*4 + a*2 + b -> chunk
// co/2; thread = (pooled pixel, pair of output channels) loads those four 16-byte chunks once.
// arg-max records: mil_stem_unpool.cuh (uint2 per pooled chunk and flat pixel).
__device__ __forceinline__ void unpack8h(const uint4& r, float v[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

// One pooled pixel, the nine window positions in ATen's scan order (first maximum wins).  The arguments are the
// 16-byte chunks of the four neighbouring phase-map pixels, already set to -inf where a neighbour lies outside the conv
// map; per channel the words are  w0 = phases (a=0: b=0 | b=1), w1 = phases (a=1: b=0 | b=1).
// Both channels of a pair at once on packed bf16x2 words (comparisons of bf16 values are exact, so this is the same
// arithmetic as pool9, at half the instructions): lane 0 = channel 2cp, lane 1 = channel 2cp+1.  Returns the pooled
// pair (already in output layout) and the two window positions as (am0 | am1 << 8).
__device__ __forceinline__ uint32_t pool9_pair(const uint4& UL, const uint4& U, const uint4& L, const uint4& S,
                                               uint32_t& am_pair) {
  auto lo = [](uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x5410); };  // (a.lo, b.lo)
  auto hi = [](uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7632); };  // (a.hi, b.hi)
  auto as2 = [](uint32_t w) { return *reinterpret_cast<const __nv_bfloat162*>(&w); };
  uint32_t best = hi(UL.y, UL.w), idx = 0;  // window (0,0)
  auto step = [&](uint32_t cand, uint32_t pos) {
    const uint32_t m = __hgt2_mask(as2(cand), as2(best));  // 0xFFFF per lane where cand > best (first maximum wins)
    const __nv_bfloat162 mx = __hmax2(as2(cand), as2(best));
    best = *reinterpret_cast<const uint32_t*>(&mx);
    idx = (idx & ~m) | ((pos | (pos << 16)) & m);
  };
  step(lo(U.y, U.w), 1);   // (py-1, px) phase (1,0) -> (0,1)
  step(hi(U.y, U.w), 2);   //            phase (1,1) -> (0,2)
  step(hi(L.x, L.z), 3);   // (py, px-1) phase (0,1) -> (1,0)
  step(lo(S.x, S.z), 4);   // (py, px)   phase (0,0) -> (1,1)
  step(hi(S.x, S.z), 5);   //            phase (0,1) -> (1,2)
  step(hi(L.y, L.w), 6);   // (py, px-1) phase (1,1) -> (2,0)
  step(lo(S.y, S.w), 7);   // (py, px)   phase (1,0) -> (2,1)
  step(hi(S.y, S.w), 8);   //            phase (1,1) -> (2,2)
  am_pair = (idx | (idx >> 8)) & 0xFFFFu;
  return best;
}

__global__ void __launch_bounds__(256)
stem_pool4_kernel(const __nv_bfloat16* __restrict__ cv, MilPF8 gc, int hc, __nv_bfloat16* __restrict__ pooled,
                  MilPF8 gp, uint2* __restrict__ argmax) {
  // one thread per (pooled pixel, pooled chunk pc of 3) = four channel pairs cp = 4pc .. 4pc+3 (cp < 10): every
  // pooled chunk leaves as ONE 16-byte store.  The kernel is bound by its instruction stream, not by HBM, so the
  // common case (even conv size: phases a = 1 / b = 1 always inside the map) is branch-free: neighbours outside the
  // map (py = 0 / px = 0) are replaced by -inf word-wise.
  // grid = (image, pixel block, pooled chunk): no 64-bit index arithmetic per thread
  const uint32_t NINF2 = 0xFF80FF80u;
  const int n = blockIdx.x, pc = blockIdx.z;
  const int r = blockIdx.y * blockDim.x + threadIdx.x;
  if (r < (int)gp.P) {
    const long long q = (long long)n * gp.P + r;
    const int py = r / gp.wp, px = r - py * gp.wp;
    float out[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t outw[4] = {0u, 0u, 0u, 0u};  // the same chunk as packed bf16x2 words (fast path)
    bool packed = false;
    if (py < gp.h && px < gp.w) {
      const long long qc = (long long)n * gc.P + (long long)py * gc.wp + px;
      const int ncp = min(4, (gp.c >> 1) - 4 * pc);  // 20 channels: pairs 8, 9 only in the last chunk (20..23 are padding)
      const bool up_ok = py > 0, left_ok = px > 0;
      const bool all_phases = 2 * py + 1 < hc && 2 * px + 1 < hc;
      const uint4* ps = reinterpret_cast<const uint4*>(cv + mil_pf8_off(gc, pc * 4, qc));  // chunk cp: + j * gc.PS
      uint32_t amw[4] = {MIL_AM_CODE, MIL_AM_CODE, MIL_AM_CODE, MIL_AM_CODE};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j >= ncp) break;
        const uint4* pj = ps + (size_t)j * gc.PS;
        // the zero halo of PF8 makes (py-1, px-1) readable everywhere; validity is decided on conv coordinates
        uint4 S = __ldg(pj), L = __ldg(pj - 1), U = __ldg(pj - gc.wp), UL = __ldg(pj - gc.wp - 1);
        float best[2];
        int am[2];
        if (all_phases) {
          if (!left_ok) { L.x = L.y = L.z = L.w = NINF2; }
          if (!up_ok) { U.y = U.w = NINF2; }
          if (!(up_ok && left_ok)) { UL.y = UL.w = NINF2; }
          uint32_t amp;
          outw[j] = pool9_pair(UL, U, L, S, amp);
          amw[j] = amp | MIL_AM_CODE;
          packed = true;
          continue;
        } else {  // odd conv size, last row / column: test every window position
          float nb[2][2][8];  // [dy: py-1, py][dx: px-1, px][8 lanes = (co sub-index)*4 + a*2 + b]
          unpack8h(UL, nb[0][0]); unpack8h(U, nb[0][1]); unpack8h(L, nb[1][0]); unpack8h(S, nb[1][1]);
          best[0] = best[1] = -INFINITY;
          am[0] = am[1] = 0;
#pragma unroll
          for (int wy = 0; wy < 3; ++wy) {
            const int cy = 2 * py - 1 + wy;
            const bool oky = cy >= 0 && cy < hc;
            const int dy = wy == 0 ? 0 : 1, a = wy == 1 ? 0 : 1;
#pragma unroll
            for (int wx = 0; wx < 3; ++wx) {
              const int cx = 2 * px - 1 + wx;
              const int dx = wx == 0 ? 0 : 1, b = wx == 1 ? 0 : 1;
              if (oky && cx >= 0 && cx < hc) {
                const float v0 = nb[dy][dx][a * 2 + b], v1 = nb[dy][dx][4 + a * 2 + b];
                if (v0 > best[0]) { best[0] = v0; am[0] = wy * 3 + wx; }
                if (v1 > best[1]) { best[1] = v1; am[1] = wy * 3 + wx; }
              }
            }
          }
        }
        amw[j] = (uint32_t)(am[0] | (am[1] << 8)) | MIL_AM_CODE;
        // pooled channel co = 2cp + {0,1} -> chunk co/8 = pc, lane co%8 = 2j + {0,1}   (j is compile-time here)
        out[2 * j] = best[0];
        out[2 * j + 1] = best[1];
      }
      if (argmax != nullptr)
        argmax[(long long)pc * gp.PS + gp.G + q] = make_uint2(amw[0] | (amw[1] << 16), amw[2] | (amw[3] << 16));
    }
    if (packed) *reinterpret_cast<uint4*>(pooled + mil_pf8_off(gp, pc, q)) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
    else mil_store8(pooled + mil_pf8_off(gp, pc, q), out);
  }
}

// ---- 4. backward: dY4(Y,X)[(co,a,b)] = sum of g(py,px)[co] over the pooled windows whose arg-max is conv (2Y+a, 2X+b)
// Un-fused form (cross-check of mil_stem_wgrad.cu, which builds the same chunks in shared memory): the windows that
// can point into pixel (Y,X) of the phase map are the pooled pixels (Y,X), (Y,X+1), (Y+1,X), (Y+1,X+1); the routing
// arithmetic is mil_unpool_pair, shared with the fused kernel.
__global__ void __launch_bounds__(256)
stem_unpool4_kernel(const __nv_bfloat16* __restrict__ g, MilPF8 gp, const uint2* __restrict__ argmax,
                    __nv_bfloat16* __restrict__ dy, MilPF8 gc) {
  // one thread per (phase-map pixel, pooled chunk pc of 3): the (up to) four channel pairs cp = 4pc .. 4pc+3, each one
  // 16-byte chunk of dY4.  grid = (image, pixel block, pooled chunk): no 64-bit index arithmetic per thread
  const int n = blockIdx.x, pc = blockIdx.z;
  const int r = blockIdx.y * blockDim.x + threadIdx.x;
  if (r < (int)gc.P) {
    const long long q = (long long)n * gc.P + r;
    const int Y = r / gc.wp, X = r - Y * gc.wp;
    const bool inside = Y < gc.h && X < gc.w;
    const int ncp = min(4, (gp.c >> 1) - 4 * pc);
    const int wp = gp.wp;
    uint4 out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = make_uint4(0, 0, 0, 0);
    if (inside) {
      const uint4* gpl = reinterpret_cast<const uint4*>(g) + (long long)pc * gp.PS + gp.G;
      const uint2* apl = argmax + (long long)pc * gp.PS + gp.G;
      const uint4 G00 = __ldg(gpl + q), G01 = __ldg(gpl + q + 1), G10 = __ldg(gpl + q + wp), G11 = __ldg(gpl + q + wp + 1);
      const uint2 A00 = __ldg(apl + q), A01 = __ldg(apl + q + 1), A10 = __ldg(apl + q + wp), A11 = __ldg(apl + q + wp + 1);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < ncp)
          out[k] = mil_unpool_pair(mil_word(G00, k), mil_word(G01, k), mil_word(G10, k), mil_word(G11, k),
                                   mil_am_lanes(A00, k), mil_am_lanes(A01, k), mil_am_lanes(A10, k), mil_am_lanes(A11, k));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < ncp) *reinterpret_cast<uint4*>(dy + mil_pf8_off(gc, pc * 4 + k, q)) = out[k];
  }
}

// ---- 5. reduction of the wgrad_tc partial records + fold back into the PyTorch layout --------------------------
// record: [tap 9][kin 48][cout 80] + [80] bias sums
#define SR4_PARTS 8
__global__ void __launch_bounds__(32 * SR4_PARTS)
stem_reduce4_kernel(const float* __restrict__ partial, int nblk, long long stride, float* __restrict__ dw,
                    float* __restrict__ db) {
  // threadIdx.x = output element, threadIdx.y takes every SR4_PARTS-th partial record; fixed-order sum of the parts
  __shared__ float part[SR4_PARTS][32];
  const int i = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (i < STC_CO * 147) {
    const int kx = i % 7, ky = (i / 7) % 7, c = (i / 49) % 3, co = i / 147;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int uy = 2 * a + ky - 3;          // = 4 dy + ry
      const int dy = (uy + 4) / 4 - 1, ry = uy - 4 * dy;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ux = 2 * b + kx - 3;
        const int dx = (ux + 4) / 4 - 1, rx = ux - 4 * dx;
        const int t = (dy + 1) * 3 + (dx + 1), cc = (c * 4 + ry) * 4 + rx, co4 = co * 4 + a * 2 + b;
        const size_t src = ((size_t)t * STC_CI + cc) * STC_CO4 + co4;
        for (int k = threadIdx.y; k < nblk; k += SR4_PARTS) acc += partial[(size_t)k * stride + src];
      }
    }
  } else if (i < STC_CO * 147 + STC_CO) {
    const int co = i - STC_CO * 147;
    for (int ph = 0; ph < 4; ++ph) {
      const size_t src = (size_t)9 * STC_CI * STC_CO4 + co * 4 + ph;
      for (int k = threadIdx.y; k < nblk; k += SR4_PARTS) acc += partial[(size_t)k * stride + src];
    }
  }
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y != 0) return;
#pragma unroll
  for (int k = 1; k < SR4_PARTS; ++k) acc += part[k][threadIdx.x];
  if (i < STC_CO * 147) dw[i] += acc;
  else if (i < STC_CO * 147 + STC_CO) db[i - STC_CO * 147] += acc;
}

// ---- host side ---------------------------------------------------------------------------------------------
static int stem_h0(int side) { return (((side - 1) / 2 + 1) - 1) / 2 + 1; }
MilPF8 mil_stem_tc_geom_in(int n, int side) { return mil_pf8(n, STC_CI, stem_h0(side), stem_h0(side)); }
MilPF8 mil_stem_tc_geom_conv(int n, int side) { return mil_pf8(n, STC_CO4, stem_h0(side), stem_h0(side)); }
size_t mil_stem_tc_wpack_floats() { return (size_t)9 * STC_CI * STC_CO4 + STC_CO4; }
size_t mil_stem_tc_wtc_bytes() {
  MilTcShape sh;
  mil_tc_shape(STC_CI, STC_CO4, 3, &sh);
  return mil_tc_wpack_bytes(sh);
}
size_t mil_stem_tc_partial_floats(int n, int side) {
  return std::max(mil_wgrad_tc_partial_floats(mil_stem_tc_geom_in(n, side), mil_stem_tc_geom_conv(n, side), 3),
                  mil_stem_wgrad_partial_floats());
}
size_t mil_stem_tc_argmax_bytes(const MilPF8& gp) { return (size_t)gp.cb * gp.PS * sizeof(uint2); }

// true when the forward pass runs the fused conv + pool kernel (which also writes the pooled map's sign mask);
// MIL_B200_STEM_UNFUSED=1 forces the two-kernel path
bool mil_stem_tc_fused_pool(const MilPF8& gp, int side) {
  return mil_opt(MIL_OPT_STEM_UNFUSED) == 0 && mil_stem_conv_pool_supported(gp, (side - 1) / 2 + 1);
}

int mil_launch_stem_tc_fwd(const void* x, int x_u8, const int* idx, int n, int side, const float* w, const float* b, void* xs,
                           void* convout, float* wp, void* wtc, void* pooled, const MilPF8& gp, uint8_t* argmax8,
                           cudaStream_t s, void* pooled_mask) {
  uint2* argmax = reinterpret_cast<uint2*>(argmax8);  // arg-max records [3][gp.PS] (mil_stem_unpool.cuh)
  const MilPF8 gi = mil_stem_tc_geom_in(n, side), gc = mil_stem_tc_geom_conv(n, side);
  const int hc = (side - 1) / 2 + 1;
  MIL_REQUIRE(gp.h == gi.h && gp.w == gi.w && gp.c == STC_CO && gp.n == n, "stem_tc_fwd: geometry mismatch");
  if (x_u8)
    stem_s2d4_kernel<uint8_t><<<dim3((unsigned)gi.n, (unsigned)std::min<long long>(mil_cdiv(gi.P, 256), 65535), 6), 256, 0, s>>>((const uint8_t*)x, idx, side, (__nv_bfloat16*)xs, gi);
  else
    stem_s2d4_kernel<float><<<dim3((unsigned)gi.n, (unsigned)std::min<long long>(mil_cdiv(gi.P, 256), 65535), 6), 256, 0, s>>>((const float*)x, idx, side, (__nv_bfloat16*)xs, gi);
  MIL_LAUNCH_OK();
  float* bias4 = wp + (size_t)9 * STC_CI * STC_CO4;
  stem_pack_w4_kernel<<<64, 256, 0, s>>>(w, b, wp, bias4);
  MIL_LAUNCH_OK();
  MilTcShape sh;
  MIL_TRY(mil_tc_shape(STC_CI, STC_CO4, 3, &sh));
  MIL_TRY(mil_launch_pack_tc(wp, wtc, sh, s));
  if (mil_stem_tc_fused_pool(gp, side))  // conv + pool in one kernel: the conv map never reaches HBM
    return mil_launch_stem_conv_pool(xs, gi, wtc, bias4, pooled, gp, argmax, pooled_mask, hc, s);
  MIL_TRY(mil_launch_conv_tc(0, xs, gi, wtc, sh, bias4, nullptr, nullptr, convout, gc, MIL_EPI_FWD, 0, s));
  stem_pool4_kernel<<<dim3(gp.n, (unsigned)mil_cdiv(gp.P, 256), gp.cb), 256, 0, s>>>((const __nv_bfloat16*)convout, gc, hc, (__nv_bfloat16*)pooled,
                                                        gp, argmax);
  MIL_LAUNCH_OK();
  return 0;
}

int mil_launch_stem_tc_bwd(const void* xs, int n, int side, const void* g, const MilPF8& gp, const uint8_t* argmax8,
                           void* dy, float* partial, float* dw, float* db, cudaStream_t s) {
  const uint2* argmax = reinterpret_cast<const uint2*>(argmax8);
  const MilPF8 gi = mil_stem_tc_geom_in(n, side), gc = mil_stem_tc_geom_conv(n, side);
  int ctas;
  long long rec;
  if (mil_opt(MIL_OPT_STEM_UNFUSED) == 0 && mil_stem_wgrad_supported(gp)) {
    // the un-pooled gradient (80 channels at the phase-map resolution, the largest tensor of the backward pass) is
    // built inside the weight-gradient kernel, tile by tile, straight into its A-operand planes
    MIL_TRY(mil_launch_stem_wgrad(xs, gi, g, gp, argmax, partial, &ctas, &rec, s));
  } else {
    stem_unpool4_kernel<<<dim3(gc.n, (unsigned)mil_cdiv(gc.P, 256), gp.cb), 256, 0, s>>>((const __nv_bfloat16*)g, gp, argmax,
                                                                                   (__nv_bfloat16*)dy, gc);
    MIL_LAUNCH_OK();
    MIL_TRY(mil_launch_wgrad_tc_partials(xs, gi, dy, gc, partial, 3, &ctas, &rec, s));
  }
  stem_reduce4_kernel<<<(int)mil_cdiv(STC_CO * 147 + STC_CO, 32), dim3(32, SR4_PARTS), 0, s>>>(partial, ctas, rec, dw, db);
  MIL_LAUNCH_OK();
  return 0;
}

// ---- the stem's layout / pool kernels on their own (the wide parameterisation, mil_wide_net.cu: any even channel
// count C, conv map = 4 * C channels (co, a, b) at the pooled resolution) ------------------------------------------
int mil_launch_stem_s2d4(const void* x, int x_u8, const int* idx, int side, void* xs, const MilPF8& gi, cudaStream_t s) {
  if (x_u8)
    stem_s2d4_kernel<uint8_t><<<dim3((unsigned)gi.n, (unsigned)std::min<long long>(mil_cdiv(gi.P, 256), 65535), 6), 256, 0, s>>>((const uint8_t*)x, idx, side, (__nv_bfloat16*)xs, gi);
  else
    stem_s2d4_kernel<float><<<dim3((unsigned)gi.n, (unsigned)std::min<long long>(mil_cdiv(gi.P, 256), 65535), 6), 256, 0, s>>>((const float*)x, idx, side, (__nv_bfloat16*)xs, gi);
  MIL_LAUNCH_OK();
  return 0;
}
int mil_launch_stem_pool4(const void* cv, const MilPF8& gc, int hc, void* pooled, const MilPF8& gp, void* argmax,
                          cudaStream_t s) {
  MIL_REQUIRE(gc.c == 4 * gp.c && (gp.c & 1) == 0 && gc.wp == gp.wp && gc.n == gp.n,
              "stem_pool4: geometry mismatch");
  stem_pool4_kernel<<<dim3(gp.n, (unsigned)mil_cdiv(gp.P, 256), gp.cb), 256, 0, s>>>((const __nv_bfloat16*)cv, gc, hc,
                                                                                     (__nv_bfloat16*)pooled, gp, (uint2*)argmax);
  MIL_LAUNCH_OK();
  return 0;
}
int mil_launch_stem_unpool4(const void* g, const MilPF8& gp, const void* argmax, void* dy, const MilPF8& gc, cudaStream_t s) {
  MIL_REQUIRE(gc.c == 4 * gp.c && (gp.c & 1) == 0 && gc.wp == gp.wp && gc.n == gp.n, "stem_unpool4: geometry mismatch");
  stem_unpool4_kernel<<<dim3(gc.n, (unsigned)mil_cdiv(gc.P, 256), gp.cb), 256, 0, s>>>((const __nv_bfloat16*)g, gp,
                                                                                       (const uint2*)argmax, (__nv_bfloat16*)dy, gc);
  MIL_LAUNCH_OK();
  return 0;
}
