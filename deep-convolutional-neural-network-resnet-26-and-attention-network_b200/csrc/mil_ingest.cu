// Tile ingest on the GPU: the per-tile finalisation the reference runs on the CPU with
// PIL for every tile of every bag (RoiBuilder.py:193-210, 215-238) --
//     Pad(100) -> RandomCrop(roi) -> Resize(S) -> RandomHorizontalFlip -> RandomVerticalFlip     (training)
//     Resize(S)                                                                                   (validation)
// -- from the cached 8-bit HWC tiles [T, R, R, 3] straight to the 8-bit NCHW bag [T, 3, S, S] the stem's 8-bit load
// consumes (ToTensor + Normalize(.5, .5) are fused into that load: mil_extractor_forward_u8).
//
// The arithmetic is Pillow's antialiased bilinear resampling of 8-bit images, reproduced bit for bit
// (libImaging/Resample.c, 8bpc path; the CPU restatement the tests compare against is pinned to Pillow itself):
//   horizontal pass over the input rows, then vertical pass, each  out = clip((2^21 + sum_k pixel_k * coef_k) >> 22)
// with the fixed-point triangle weights the HOST computes in double precision exactly as precompute_coeffs /
// normalize_coeffs_8bpc do (ingest.py) -- square tiles, one table serves both passes.  Pad + crop are a source offset
// with zero fill, the flips an output index mapping.
//
// One CTA = (tile, band of BH output rows): it runs the horizontal pass for the input rows the band's vertical windows
// touch (staged RB rows at a time in shared memory with 32-bit loads; consecutive lanes take consecutive ROWS, the
// row pitch is an odd number of words: conflict-free, and a half-warp shares one coefficient window), keeps the 8-bit
// intermediate rows in shared memory, runs the vertical pass and writes the band.  Integer / byte work: bound by the
// shared-memory byte loads and IMADs (39 per output pixel and pass), far below the 12 s per 2 500-tile slide the
// reference's DataLoader workers spend on it.
#include <algorithm>

#include "mil_common.cuh"

#define ING_THREADS 1024  // one CTA per SM (shared memory): its 32 warps hide the byte-load -> IMAD latencies
#define ING_PRECISION_BITS 22

__global__ void __launch_bounds__(ING_THREADS)
ingest_kernel(const uint8_t* __restrict__ rois, int R, const int* __restrict__ crops, int pad,
              const uint8_t* __restrict__ flips, int S, const int* __restrict__ bounds, const int* __restrict__ coef,
              int ksize, uint8_t* __restrict__ out, int BH, int RB, int pitch, int rows_max) {
  extern __shared__ __align__(16) unsigned char sm[];
  const size_t stage_bytes = max((size_t)RB * pitch, ((size_t)3 * BH * S + 15) / 16 * 16);
  uint8_t* rowbuf = sm;                                   // [RB][pitch]   staged input rows (later: the band's output)
  uint8_t* tmp = sm + stage_bytes;                        // [rows_max][S * 3]  horizontal-pass output
  int* s_bounds = reinterpret_cast<int*>(tmp + ((size_t)rows_max * S * 3 + 15) / 16 * 16);   // [S][2]
  int* s_coef = s_bounds + 2 * S;                                                            // [S][ksize]
  __shared__ int s_rowsrc[64];
  for (int i = threadIdx.x; i < 2 * S; i += ING_THREADS) s_bounds[i] = bounds[i];
  for (int i = threadIdx.x; i < S * ksize; i += ING_THREADS) s_coef[i] = coef[i];
  __syncthreads();
  const int t = blockIdx.y, y0 = blockIdx.x * BH, y1 = min(S, y0 + BH);
  const int top = crops ? crops[2 * t] - pad : 0, left = crops ? crops[2 * t + 1] - pad : 0;
  const int in_lo = bounds[2 * y0], in_hi = bounds[2 * (y1 - 1)] + bounds[2 * (y1 - 1) + 1];
  const int nrows = in_hi - in_lo;
  const uint8_t* tile = rois + (size_t)t * R * R * 3;
  const int row_bytes = R * 3;
  const bool words = (row_bytes & 3) == 0 && ((size_t)rois & 3) == 0;

  // ---- horizontal pass, RB input rows at a time ----
  for (int r0 = 0; r0 < nrows; r0 += RB) {
    const int nr = min(RB, nrows - r0);
    if (threadIdx.x < RB) {
      const int srow = in_lo + r0 + threadIdx.x + top;    // row of the cached tile (outside: the zero padding)
      s_rowsrc[threadIdx.x] = ((int)threadIdx.x < nr && srow >= 0 && srow < R) ? srow : -1;
    }
    __syncthreads();
    if (words) {
      const int nw = row_bytes >> 2;
      for (int i = threadIdx.x; i < nr * nw; i += ING_THREADS) {
        const int k = i / nw, w = i - k * nw, srow = s_rowsrc[k];
        if (srow >= 0)
          reinterpret_cast<uint32_t*>(rowbuf + (size_t)k * pitch)[w] =
              __ldg(reinterpret_cast<const uint32_t*>(tile + (size_t)srow * row_bytes) + w);
      }
    } else {
      for (int i = threadIdx.x; i < nr * row_bytes; i += ING_THREADS) {
        const int k = i / row_bytes, b = i - k * row_bytes, srow = s_rowsrc[k];
        if (srow >= 0) rowbuf[(size_t)k * pitch + b] = __ldg(tile + (size_t)srow * row_bytes + b);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RB * S; i += ING_THREADS) {
      const int ox = i / RB, k = i - ox * RB;             // consecutive lanes: consecutive rows, one window
      if (k >= nr) continue;
      int a0 = 1 << (ING_PRECISION_BITS - 1), a1 = a0, a2 = a0;
      if (s_rowsrc[k] >= 0) {
        const int xmin = s_bounds[2 * ox], cnt = s_bounds[2 * ox + 1];
        const uint8_t* row = rowbuf + (size_t)k * pitch;
        const int* cf = s_coef + ox * ksize;
        for (int j = 0; j < cnt; ++j) {
          const int sx = xmin + j + left;
          if ((unsigned)sx < (unsigned)R) {
            const int c = cf[j];
            const uint8_t* px = row + sx * 3;
            a0 += px[0] * c; a1 += px[1] * c; a2 += px[2] * c;
          }
        }
      }
      uint8_t* o = tmp + (size_t)(r0 + k) * S * 3 + ox * 3;
      o[0] = (uint8_t)min(max(a0 >> ING_PRECISION_BITS, 0), 255);
      o[1] = (uint8_t)min(max(a1 >> ING_PRECISION_BITS, 0), 255);
      o[2] = (uint8_t)min(max(a2 >> ING_PRECISION_BITS, 0), 255);
    }
    __syncthreads();
  }

  // ---- vertical pass: the band's output, channel-planar in shared memory [3][BH][S] ----
  uint8_t* band = rowbuf;
  const int nout = (y1 - y0) * S * 3;
  for (int i = threadIdx.x; i < nout; i += ING_THREADS) {
    const int xc = i % (S * 3), oy = y0 + i / (S * 3);
    const int ymin = s_bounds[2 * oy], cnt = s_bounds[2 * oy + 1];
    const int* cf = s_coef + oy * ksize;
    int a = 1 << (ING_PRECISION_BITS - 1);
    const uint8_t* col = tmp + (size_t)(ymin - in_lo) * S * 3 + xc;
    for (int j = 0; j < cnt; ++j) a += col[(size_t)j * S * 3] * cf[j];
    const int ox = xc / 3, c = xc - ox * 3;
    band[((size_t)c * BH + (oy - y0)) * S + ox] = (uint8_t)min(max(a >> ING_PRECISION_BITS, 0), 255);
  }
  __syncthreads();

  // ---- write the band (flips = output index mapping) ----
  const int fl = flips ? flips[t] : 0;
  const bool hf = fl & 1, vf = fl & 2;
  uint8_t* ot = out + (size_t)t * 3 * S * S;
  if (!hf && (S & 3) == 0 && ((size_t)out & 3) == 0) {
    const int sw = S >> 2;
    for (int i = threadIdx.x; i < 3 * (y1 - y0) * sw; i += ING_THREADS) {
      const int w = i % sw, ry = (i / sw) % (y1 - y0), c = i / (sw * (y1 - y0));
      const int oy = y0 + ry, dy = vf ? S - 1 - oy : oy;
      reinterpret_cast<uint32_t*>(ot + ((size_t)c * S + dy) * S)[w] =
          reinterpret_cast<const uint32_t*>(band + ((size_t)c * BH + ry) * S)[w];
    }
  } else {
    for (int i = threadIdx.x; i < 3 * (y1 - y0) * S; i += ING_THREADS) {
      const int ox = i % S, ry = (i / S) % (y1 - y0), c = i / (S * (y1 - y0));
      const int oy = y0 + ry, dy = vf ? S - 1 - oy : oy, dx = hf ? S - 1 - ox : ox;
      ot[((size_t)c * S + dy) * S + dx] = band[((size_t)c * BH + ry) * S + ox];
    }
  }
}

// bounds_host: the host copy of `bounds` (the launcher sizes the band's shared memory from it)
int mil_launch_ingest_u8(const uint8_t* rois, int T, int R, const int* crops, int pad, const uint8_t* flips, int S,
                         const int* bounds, const int* coef, int ksize, const int* bounds_host, uint8_t* out,
                         cudaStream_t s) {
  MIL_REQUIRE(T >= 1 && R >= 1 && S >= 1 && ksize >= 1, "ingest: bad sizes (T %d, R %d, S %d)", T, R, S);
  int pitch = (R * 3 + 3) / 4 * 4;
  if (((pitch >> 2) & 1) == 0) pitch += 4;  // odd number of words: consecutive rows fall into different banks
  int BH = 16, RB = 16, rows_max = 0;
  size_t smem = 0;
  for (;;) {
    rows_max = 0;
    for (int y0 = 0; y0 < S; y0 += BH) {
      const int y1 = std::min(S, y0 + BH);
      rows_max = std::max(rows_max, bounds_host[2 * (y1 - 1)] + bounds_host[2 * (y1 - 1) + 1] - bounds_host[2 * y0]);
    }
    // the staging rows double as the band's output buffer [3][BH][S]
    const size_t stage = std::max((size_t)RB * pitch, ((size_t)3 * BH * S + 15) / 16 * 16);
    smem = stage + ((size_t)rows_max * S * 3 + 15) / 16 * 16 + (size_t)S * (2 + ksize) * sizeof(int);
    if (smem <= 200 * 1024) break;
    if (RB > 1 && (size_t)RB * pitch >= (size_t)rows_max * S * 3) RB /= 2;
    else if (BH > 1) BH /= 2;
    else if (RB > 1) RB /= 2;
    else break;
  }
  MIL_REQUIRE(smem <= 200 * 1024, "ingest: a %d-pixel tile resized to %d needs %zu bytes of shared memory", R, S, smem);
  MIL_SET_SMEM(ingest_kernel, smem);
  ingest_kernel<<<dim3((unsigned)mil_cdiv(S, BH), (unsigned)T), ING_THREADS, smem, s>>>(
      rois, R, crops, pad, flips, S, bounds, coef, ksize, out, BH, RB, pitch, rows_max);
  MIL_LAUNCH_OK();
  return 0;
}
