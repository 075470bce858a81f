// Whole-network drivers of the WIDE extractor parameterisation: the reference's alt_resnet.py (torchvision's ResNet
// with the BatchNorm layers stripped) as a tile feature extractor in front of the same MIL head.
//   alt_resnet.py:81-91   conv1 7x7 / stride 2 / pad 3 (3 -> 64, no bias), ReLU, MaxPool 3x3 / stride 2 / pad 1,
//                         layer1..4 of BasicBlocks (widths 64 / 128 / 256 / 512), AdaptiveAvgPool, fc WITH bias
//   alt_resnet.py:35-67   BasicBlock: conv3x3 (stride) -> ReLU -> conv3x3 -> += identity | conv1x1(stride) -> ReLU
//   alt_resnet.py:107-123 _make_layer: a 1x1 / stride-2 projection (no bias) wherever the shape changes
// Parameter order = the state dict of `Attention` with `cnn = DataParallel(alt_resnet.ResNet(BasicBlock, layers,
// num_classes = 80))`: weight_mask, cnn.module.conv1.weight, cnn.module.layerL.B.{conv1,conv2}.weight
// [, downsample.0.weight], cnn.module.fc.{weight,bias}, then the head's ten tensors (gbm/model.py:137-159).
// bf16 activations, fp32 accumulation; everything is enqueued on the caller's stream, nothing is allocated here.
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/mil_b200.h"
#include "mil_common.cuh"
#include "mil_conv_tc.cuh"
#include "mil_extractor.cuh"
#include "mil_wide.cuh"
#include "mil_wide_net.cuh"

// ---------------------------------------------------------------------------------------------------
// parameter table
// ---------------------------------------------------------------------------------------------------
int mil_wide_check_desc(const MilWideDesc& d) {
  MIL_REQUIRE(d.stem == d.widths[0], "wide extractor: layer1 keeps the stem's width (alt_resnet.py:81,87), got %d / %d",
              d.stem, d.widths[0]);
  MIL_REQUIRE(d.stem % 32 == 0 && d.stem >= 32 && d.stem <= 128, "wide extractor: stem width %d (need 32, 64, 96 or 128)", d.stem);
  for (int l = 0; l < 4; ++l) {
    MIL_REQUIRE(d.layers[l] >= 1 && d.layers[l] <= 8, "wide extractor: %d blocks in layer %d", d.layers[l], l + 1);
    MIL_REQUIRE(d.widths[l] % 64 == 0 && d.widths[l] >= 64 && d.widths[l] <= 512, "wide extractor: width %d of layer %d",
                d.widths[l], l + 1);
    MIL_REQUIRE(d.widths[l] < 128 || d.widths[l] % 128 == 0, "wide extractor: width %d of layer %d (multiples of 128 above 64)",
                d.widths[l], l + 1);
  }
  MIL_REQUIRE(d.features == MIL_FEATURES, "wide extractor: the MIL head takes %d features (gbm/model.py:120)", MIL_FEATURES);
  MIL_REQUIRE(d.slope >= 0.f && d.slope < 1.f, "wide extractor: activation slope %g", d.slope);
  return 0;
}

std::vector<MilParamInfo> mil_wide_param_table(const MilWideDesc& d) {
  std::vector<MilParamInfo> t;
  long long off = 0;
  auto add = [&](const std::string& name, std::initializer_list<long long> shape) {
    MilParamInfo p;
    p.name = name;
    p.ndim = (int)shape.size();
    p.numel = 1;
    int i = 0;
    for (long long v : shape) { p.shape[i++] = v; p.numel *= v; }
    for (; i < 4; ++i) p.shape[i] = 1;
    p.offset = off;
    off += p.numel;
    t.push_back(p);
  };
  add("weight_mask", {3});
  add("cnn.module.conv1.weight", {d.stem, 3, 7, 7});
  int inpl = d.stem;
  for (int l = 0; l < 4; ++l) {
    const int w = d.widths[l];
    for (int b = 0; b < d.layers[l]; ++b) {
      const int cin = b == 0 ? inpl : w;
      const std::string p = "cnn.module.layer" + std::to_string(l + 1) + "." + std::to_string(b);
      add(p + ".conv1.weight", {w, cin, 3, 3});
      add(p + ".conv2.weight", {w, w, 3, 3});
      if (b == 0 && (l > 0 || cin != w)) add(p + ".downsample.0.weight", {w, cin, 1, 1});
    }
    inpl = w;
  }
  add("cnn.module.fc.weight", {d.features, d.widths[3]});
  add("cnn.module.fc.bias", {d.features});
  add("context.bn.weight", {80});
  add("context.bn.bias", {80});
  add("attention.lin1.weight", {40, 80});
  add("attention.lin1.bias", {40});
  add("attention.lin2.weight", {3, 40});
  add("attention.lin2.bias", {3});
  add("buffer.lin1.weight", {40, 80});
  add("buffer.lin1.bias", {40});
  add("buffer.classifier.weight", {1, 40});
  add("buffer.classifier.bias", {1});
  return t;
}

static int pindex(const std::vector<MilParamInfo>& t, const std::string& name) {
  for (size_t i = 0; i < t.size(); ++i)
    if (t[i].name == name) return (int)i;
  return -1;
}

// ---------------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------------
static size_t walign(size_t v) { return (v + 255) / 256 * 256; }

static size_t wgrad_partial_floats(const MilPF8& gx, const MilPF8& gz, int ks) {
  return mil_wide_wgrad_partial_floats(gx, gz, ks);
}

int mil_wide_make_plan(const MilWideDesc& d, int n, int side, MilWidePlan* plan) {
  MIL_TRY(mil_wide_check_desc(d));
  MIL_REQUIRE(n >= 1 && side >= 16, "wide extractor: bad shape n=%d side=%d", n, side);
  MilWidePlan& pl = *plan;
  pl.d = d; pl.n = n; pl.side = side;
  pl.params = mil_wide_param_table(d);
  pl.hc = (side - 1) / 2 + 1;
  pl.h[0] = (pl.hc - 1) / 2 + 1;
  for (int l = 1; l < 4; ++l) pl.h[l] = (pl.h[l - 1] - 1) / 2 + 1;
  for (int l = 0; l < 4; ++l) pl.g[l] = mil_pf8(n, d.widths[l], pl.h[l], pl.h[l]);
  pl.gxs = mil_pf8(n, 48, pl.h[0], pl.h[0]);
  pl.gcv = mil_pf8(n, 4 * d.stem, pl.h[0], pl.h[0]);
  pl.convs.clear();
  size_t wofs = 0;
  auto add_conv = [&](int l, int b, int which, int cin, int cout, int ks, int stride, const std::string& nm) -> int {
    MilWideConv c;
    c.layer = l; c.block = b; c.which = which; c.cin = cin; c.cout = cout; c.ks = ks; c.stride = stride;
    c.p_w = pindex(pl.params, nm);
    MIL_REQUIRE(c.p_w >= 0, "wide extractor: internal error, no parameter %s", nm.c_str());
    MilWideShape sf, st;
    MIL_TRY(mil_wide_shape(ks == 3 && stride == 2 ? 1 : 0, 0, cout, cin, ks, &sf));
    MIL_TRY(mil_wide_shape(0, 1, cout, cin, ks, &st));
    c.wf_off = wofs; wofs += walign(mil_wide_wpack_bytes(sf));
    c.wt_off = wofs; wofs += walign(mil_wide_wpack_bytes(st));
    if (ks == 3 && stride == 2) {
      MilWideShape sp;
      MIL_TRY(mil_wide_shape(7, 0, cout, cin, 3, &sp));
      pl.s2_wt_off[l][0] = wofs; wofs += walign(mil_wide_wpack_bytes(sp));
    }
    pl.convs.push_back(c);
    return 0;
  };
  int inpl = d.stem;
  pl.first_conv.assign(4, std::vector<int>());
  for (int l = 0; l < 4; ++l) {
    const int w = d.widths[l];
    for (int b = 0; b < d.layers[l]; ++b) {
      const int cin = b == 0 ? inpl : w;
      const bool down = b == 0 && l > 0;
      const std::string p = "cnn.module.layer" + std::to_string(l + 1) + "." + std::to_string(b);
      pl.first_conv[l].push_back((int)pl.convs.size());
      MIL_TRY(add_conv(l, b, 0, cin, w, 3, down ? 2 : 1, p + ".conv1.weight"));
      MIL_TRY(add_conv(l, b, 1, w, w, 3, 1, p + ".conv2.weight"));
      if (down) MIL_TRY(add_conv(l, b, 2, cin, w, 1, 2, p + ".downsample.0.weight"));
    }
    inpl = w;
  }
  {
    MilWideShape ss;
    MIL_TRY(mil_wide_shape(2, 0, d.stem, 3, 7, &ss));
    pl.stem_w_off = wofs; wofs += walign(mil_wide_wpack_bytes(ss));
  }
  pl.wpack_bytes = wofs;

  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = walign(off + bytes); return o; };
  pl.off_xs = take(mil_pf8_bytes(pl.gxs, MIL_BF16));
  pl.off_pooled = take(mil_pf8_bytes(pl.g[0], MIL_BF16));
  pl.off_argmax = take(mil_stem_tc_argmax_bytes(pl.g[0]));
  pl.off_h.assign(4, std::vector<size_t>());
  pl.off_y.assign(4, std::vector<size_t>());
  for (int l = 0; l < 4; ++l)
    for (int b = 0; b < d.layers[l]; ++b) {
      pl.off_h[l].push_back(take(mil_pf8_bytes(pl.g[l], MIL_BF16)));
      pl.off_y[l].push_back(take(mil_pf8_bytes(pl.g[l], MIL_BF16)));
    }
  for (int l = 0; l < 4; ++l) pl.off_xs2[l] = 0;
  for (int l = 1; l < 4; ++l) pl.off_xs2[l] = take(mil_pf8_bytes(mil_split2_geom(n, d.widths[l - 1], pl.h[l]), MIL_BF16));
  pl.off_avg = take((size_t)n * d.widths[3] * sizeof(float));
  // scratch shared by the forward pass (the stem's four-phase conv map) and the backward pass (three rotating gradient
  // maps, the zero-stuffed gradient of a stride-2 block, the projection branch's gradient at both resolutions, the
  // stem's four-phase gradient map)
  size_t gmax = 0, upmax = 0, tmax = 0;
  for (int l = 0; l < 4; ++l) gmax = std::max(gmax, mil_pf8_bytes(pl.g[l], MIL_BF16));
  for (int l = 1; l < 4; ++l) {
    upmax = std::max(upmax, mil_pf8_bytes(mil_split2_geom(n, d.widths[l - 1], pl.h[l]), MIL_BF16));   // phase-split gradient
    tmax = std::max(tmax, mil_pf8_bytes(mil_pf8(n, d.widths[l - 1], pl.h[l], pl.h[l]), MIL_BF16));
  }
  const size_t cvb = mil_pf8_bytes(pl.gcv, MIL_BF16);
  pl.off_cv = take(std::max(cvb, walign(upmax) + walign(tmax)));
  pl.off_up = pl.off_cv;
  pl.off_tsub = pl.off_cv + walign(upmax);
  for (int i = 0; i < 3; ++i) pl.off_grad[i] = take(gmax);
  pl.off_tfull = 0;
  pl.off_wpack = take(pl.wpack_bytes);
  size_t pf = (size_t)64 * d.features * d.widths[3] + 64 * d.features;  // tail
  pf = std::max(pf, wgrad_partial_floats(pl.gxs, pl.gcv, 7));
  for (const auto& c : pl.convs) {
    const MilPF8& go = pl.g[c.layer];
    if (c.stride == 2 && c.ks == 3) {
      pf = std::max(pf, mil_wide_wgrad_partial_floats(mil_split2_geom(n, c.cin, go.h), go, 3, 1));
    } else {
      pf = std::max(pf, wgrad_partial_floats(mil_pf8(n, c.cin, go.h, go.w), go, c.ks));
    }
  }
  pl.partial_floats = pf;
  pl.off_partial = take(pf * sizeof(float));
  pl.total_bytes = off;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// tail: global average pool + fc with bias (alt_resnet.py:89-90,134-136) and its backward, fp32
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
wide_avg_kernel(const __nv_bfloat16* __restrict__ y, MilPF8 g, float* __restrict__ avg) {
  // one thread per (tile, chunk): sums the h x w pixels of its eight channels
  const long long total = (long long)g.n * g.cb;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(i % g.cb), n = (int)(i / g.cb);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int yy = 0; yy < g.h; ++yy)
      for (int xx = 0; xx < g.w; ++xx) {
        float v[8];
        mil_load8(y + mil_pf8_off(g, cb, (long long)n * g.P + (long long)yy * g.wp + xx), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += v[j];
      }
    const float inv = 1.f / (float)(g.h * g.w);
#pragma unroll
    for (int j = 0; j < 8; ++j) avg[(size_t)n * g.c + cb * 8 + j] = a[j] * inv;
  }
}
// H[n][f] = b[f] + sum_c avg[n][c] W[f][c]: one block per tile, one warp per output feature (strided), lanes over c
__global__ void __launch_bounds__(256)
wide_fc_kernel(const float* __restrict__ avg, const float* __restrict__ w, const float* __restrict__ b, int C, int Fo,
               float* __restrict__ H) {
  extern __shared__ float s_avg[];
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_avg[c] = avg[(size_t)n * C + c];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int f = warp; f < Fo; f += blockDim.x >> 5) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(w[(size_t)f * C + c], s_avg[c], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) H[(size_t)n * Fo + f] = a + b[f];
  }
}
// dz(n, y, x)[c] = (sum_f dH[n][f] W[f][c]) / (h w) * act'(y(n, y, x)[c]); pad pixels zero
__global__ void __launch_bounds__(256)
wide_tail_dz_kernel(const __nv_bfloat16* __restrict__ y, MilPF8 g, const float* __restrict__ w, const float* __restrict__ dH,
                    int Fo, float slope, __nv_bfloat16* __restrict__ dz) {
  // one block per tile: (1) d[c] = sum_f dH[n][f] W[f][c], threads along c (coalesced weight rows); (2) the tile's
  // (chunk, pixel) items with the pixel index fastest, so that a warp loads / stores 512 contiguous bytes of a chunk plane
  extern __shared__ float s_mem[];
  float* s_dh = s_mem;       // [Fo]
  float* s_d = s_mem + Fo;   // [g.cb * 8] (channels past g.c: zero)
  const int n = blockIdx.x;
  for (int f = threadIdx.x; f < Fo; f += blockDim.x) s_dh[f] = dH[(size_t)n * Fo + f];
  __syncthreads();
  const float inv = 1.f / (float)(g.h * g.w);
  for (int c = threadIdx.x; c < g.cb * 8; c += blockDim.x) {
    float d = 0.f;
    if (c < g.c)
      for (int f = 0; f < Fo; ++f) d = fmaf(s_dh[f], __ldg(w + (size_t)f * g.c + c), d);
    s_d[c] = d * inv;
  }
  __syncthreads();
  const int P = (int)g.P, total = g.cb * P;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int cb = i / P, p = i - cb * P;
    const int yy = p / g.wp, xx = p - yy * g.wp;
    const long long o = mil_pf8_off(g, cb, (long long)n * g.P + p);
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (yy < g.h && xx < g.w) {
      float a[8];
      mil_load8(y + o, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = s_d[cb * 8 + j] * (a[j] > 0.f ? 1.f : slope);
    }
    mil_store8(dz + o, v);
  }
}
// partial[blk][f][c] = sum over the block's tiles of dH[n][f] avg[n][c]  (+ [f] bias sums); fixed-order reduction after
#define WIDE_TAIL_BLOCKS 64
__global__ void __launch_bounds__(256)
wide_tail_w_kernel(const float* __restrict__ avg, const float* __restrict__ dH, int n_tiles, int C, int Fo,
                   float* __restrict__ partial) {
  const int per = (int)mil_cdiv(n_tiles, (int)gridDim.y);
  const int n0 = blockIdx.y * per, n1 = min(n0 + per, n_tiles);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (f, c), c fastest; the last Fo entries are the bias sums
  const int total = Fo * C + Fo;
  if (i >= total) return;
  float acc = 0.f;
  if (i < Fo * C) {
    const int f = i / C, c = i - f * C;
    for (int n = n0; n < n1; ++n) acc = fmaf(dH[(size_t)n * Fo + f], avg[(size_t)n * C + c], acc);
  } else {
    const int f = i - Fo * C;
    for (int n = n0; n < n1; ++n) acc += dH[(size_t)n * Fo + f];
  }
  partial[(size_t)blockIdx.y * total + i] = acc;
}

// ---------------------------------------------------------------------------------------------------
// phase merge (inverse of split2_kernel): out(n, 2Y + a, 2X + b) = in plane (2a + b) * cb + c at (Y, X)
// grid = (image, pixel block of the HALF-resolution map incl. its pad row / column, output chunk): a thread writes a
// 2x2 block of the full-resolution map, zeros on its pad row / column
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge2_kernel(const uint4* __restrict__ in, MilPF8 gin, uint4* __restrict__ out,
                                                     MilPF8 gout) {
  const int n = blockIdx.x, c = blockIdx.z;
  const int r = blockIdx.y * blockDim.x + threadIdx.x;
  if (r >= (int)gin.P) return;
  const int Y = r / gin.wp, X = r - Y * gin.wp;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  const bool inside = Y < gin.h && X < gin.w;
  const size_t i0 = gin.G + (size_t)n * gin.P + r;
  uint4* po = out + (size_t)c * gout.PS + gout.G + (size_t)n * gout.P;
#pragma unroll
  for (int ph = 0; ph < 4; ++ph) {
    const int y = 2 * Y + (ph >> 1), x = 2 * X + (ph & 1);
    if (y >= gout.hp || x >= gout.wp) continue;
    uint4 v = zero;
    if (inside && y < gout.h && x < gout.w) v = in[(size_t)(ph * gout.cb + c) * gin.PS + i0];
    po[(size_t)y * gout.wp + x] = v;
  }
}
int mil_launch_merge2(const void* in, const MilPF8& gin, void* out, const MilPF8& gout, cudaStream_t s) {
  MIL_REQUIRE(gin.n == gout.n && gin.cb == 4 * gout.cb && gin.h == (gout.h - 1) / 2 + 1 && gin.w == (gout.w - 1) / 2 + 1,
              "merge2: geometry mismatch");
  merge2_kernel<<<dim3(gout.n, (unsigned)mil_cdiv(gin.P, 256), gout.cb), 256, 0, s>>>((const uint4*)in, gin, (uint4*)out, gout);
  MIL_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// forward / backward
// ---------------------------------------------------------------------------------------------------
static inline char* wsp(void* ws, size_t off) { return reinterpret_cast<char*>(ws) + off; }

static int zg(void* buf, const MilPF8& g, cudaStream_t s) { return mil_zero_guards(MIL_BF16, buf, g, s); }

static int pack_all(const void* const* params, const MilWidePlan& pl, void* ws, bool transposed, cudaStream_t s) {
  char* area = wsp(ws, pl.off_wpack);
  for (const auto& c : pl.convs) {
    MilWideShape sh;
    if (transposed && c.ks == 3 && c.stride == 2) {
      MIL_TRY(mil_wide_shape(7, 0, c.cout, c.cin, 3, &sh));
      MIL_TRY(mil_launch_wide_pack((const float*)params[c.p_w], area + pl.s2_wt_off[c.layer][0], sh, s));
      continue;
    }
    if (transposed) MIL_TRY(mil_wide_shape(0, 1, c.cout, c.cin, c.ks, &sh));
    else MIL_TRY(mil_wide_shape(c.ks == 3 && c.stride == 2 ? 1 : 0, 0, c.cout, c.cin, c.ks, &sh));
    MIL_TRY(mil_launch_wide_pack((const float*)params[c.p_w], area + (transposed ? c.wt_off : c.wf_off), sh, s));
  }
  if (!transposed) {
    MilWideShape ss;
    MIL_TRY(mil_wide_shape(2, 0, pl.d.stem, 3, 7, &ss));
    MIL_TRY(mil_launch_wide_pack((const float*)params[pindex(pl.params, "cnn.module.conv1.weight")], area + pl.stem_w_off, ss, s));
  }
  return 0;
}

int mil_wide_forward_impl(const void* const* params, const void* bag, int bag_u8, const int* idx, const MilWidePlan& pl,
                          void* ws, float* H, cudaStream_t s) {
  const MilPdlScope pdl_scope((long long)pl.n * pl.side * pl.side);
  const MilWideDesc& d = pl.d;
  const float slope = d.slope;
  char* wpk = wsp(ws, pl.off_wpack);
  MIL_TRY(pack_all(params, pl, ws, false, s));
  // ---- stem: space-to-depth -> conv (3x3 taps, 48 -> 4 C) + ReLU -> max-pool over (neighbour, phase) pairs ----
  void* xs = wsp(ws, pl.off_xs);
  void* cv = wsp(ws, pl.off_cv);
  void* pooled = wsp(ws, pl.off_pooled);
  MIL_TRY(zg(xs, pl.gxs, s));
  MIL_TRY(zg(cv, pl.gcv, s));
  MIL_TRY(zg(pooled, pl.g[0], s));
  MIL_TRY(mil_launch_stem_s2d4(bag, bag_u8, idx, pl.side, xs, pl.gxs, s));
  {
    MilWideShape ss;
    MIL_TRY(mil_wide_shape(2, 0, d.stem, 3, 7, &ss));
    MIL_TRY(mil_launch_wide_conv(xs, pl.gxs, wpk + pl.stem_w_off, ss, nullptr, nullptr, nullptr, cv, pl.gcv, MIL_EPI_FWD,
                                 slope, 0, s));
  }
  MIL_TRY(mil_launch_stem_pool4(cv, pl.gcv, pl.hc, pooled, pl.g[0], wsp(ws, pl.off_argmax), s));
  // ---- residual blocks ----
  const void* X = pooled;
  MilPF8 gx = pl.g[0];
  for (int l = 0; l < 4; ++l) {
    const MilPF8& go = pl.g[l];
    for (int b = 0; b < d.layers[l]; ++b) {
      const int ci = pl.first_conv[l][b];
      const MilWideConv& c1 = pl.convs[ci];
      const MilWideConv& c2 = pl.convs[ci + 1];
      void* h = wsp(ws, pl.off_h[l][b]);
      void* y = wsp(ws, pl.off_y[l][b]);
      MIL_TRY(zg(h, go, s));
      MIL_TRY(zg(y, go, s));
      const void* res = X;
      MilWideShape sh;
      if (c1.stride == 2) {
        // the 3x3 / stride-2 convolution reads the four parity phases of the block input at the OUTPUT resolution; the
        // 1x1 / stride-2 projection is a plain 1x1 convolution of phase (0, 0) = the even positions
        const MilWideConv& cd = pl.convs[ci + 2];
        const MilPF8 gs = mil_split2_geom(pl.n, gx.c, go.h);
        void* xs2 = wsp(ws, pl.off_xs2[l]);
        MIL_TRY(zg(xs2, gs, s));
        MIL_TRY(mil_launch_split2(X, gx, xs2, gs, s));
        MIL_TRY(mil_wide_shape(1, 0, c1.cout, c1.cin, 3, &sh));
        MIL_TRY(mil_launch_wide_conv(xs2, gs, wpk + c1.wf_off, sh, nullptr, nullptr, nullptr, h, go, MIL_EPI_FWD, slope, 0, s));
        MIL_TRY(mil_wide_shape(0, 0, cd.cout, cd.cin, 1, &sh));
        MIL_TRY(mil_launch_wide_conv(xs2, mil_split2_phase0(gs, gx.c), wpk + cd.wf_off, sh, nullptr, nullptr, nullptr, y, go,
                                     MIL_EPI_PLAIN, slope, 0, s));
        res = y;
      } else {
        MIL_TRY(mil_wide_shape(0, 0, c1.cout, c1.cin, 3, &sh));
        MIL_TRY(mil_launch_wide_conv(X, gx, wpk + c1.wf_off, sh, nullptr, nullptr, nullptr, h, go, MIL_EPI_FWD, slope, 0, s));
      }
      MIL_TRY(mil_wide_shape(0, 0, c2.cout, c2.cin, 3, &sh));
      MIL_TRY(mil_launch_wide_conv(h, go, wpk + c2.wf_off, sh, nullptr, res, nullptr, y, go, MIL_EPI_FWD, slope, 0, s));
      X = y;
      gx = go;
    }
  }
  // ---- tail ----
  const int C = d.widths[3], Fo = d.features;
  float* avg = (float*)wsp(ws, pl.off_avg);
  wide_avg_kernel<<<(int)std::min<long long>(mil_cdiv((long long)pl.n * gx.cb, 256), 148 * 8), 256, 0, s>>>(
      (const __nv_bfloat16*)X, gx, avg);
  MIL_LAUNCH_OK();
  wide_fc_kernel<<<pl.n, 256, C * sizeof(float), s>>>(avg, (const float*)params[pindex(pl.params, "cnn.module.fc.weight")],
                                                      (const float*)params[pindex(pl.params, "cnn.module.fc.bias")], C, Fo, H);
  MIL_LAUNCH_OK();
  return 0;
}

static int wide_wgrad(const void* x, const MilPF8& gx, const void* dz, const MilPF8& gz, float* partial, float* dw, int ks,
                      cudaStream_t s) {
  return mil_launch_wide_wgrad(x, gx, dz, gz, partial, dw, nullptr, ks, s);
}

int mil_wide_backward_impl(const void* const* params, const MilWidePlan& pl, void* ws, const float* dH, float* grads,
                           cudaStream_t s) {
  const MilPdlScope pdl_scope((long long)pl.n * pl.side * pl.side);
  const MilWideDesc& d = pl.d;
  const float slope = d.slope;
  char* wpk = wsp(ws, pl.off_wpack);
  float* partial = (float*)wsp(ws, pl.off_partial);
  auto gptr = [&](int p) { return grads + pl.params[p].offset; };
  MIL_TRY(pack_all(params, pl, ws, true, s));
  void* gb[3] = {wsp(ws, pl.off_grad[0]), wsp(ws, pl.off_grad[1]), wsp(ws, pl.off_grad[2])};
  void* dz = gb[0];
  void* dpre = gb[1];
  void* dnew = gb[2];
  // ---- tail ----
  {
    const int C = d.widths[3], Fo = d.features;
    const int p_fw = pindex(pl.params, "cnn.module.fc.weight"), p_fb = pindex(pl.params, "cnn.module.fc.bias");
    const float* avg = (const float*)wsp(ws, pl.off_avg);
    const MilPF8& g3 = pl.g[3];
    wide_tail_dz_kernel<<<pl.n, 256, (Fo + g3.cb * 8) * sizeof(float), s>>>((const __nv_bfloat16*)wsp(ws, pl.off_y[3].back()), g3,
                                                             (const float*)params[p_fw], dH, Fo, slope, (__nv_bfloat16*)dz);
    MIL_LAUNCH_OK();
    const int total = Fo * C + Fo;
    const int blocks = (int)std::min<long long>(WIDE_TAIL_BLOCKS, pl.n);
    wide_tail_w_kernel<<<dim3((unsigned)mil_cdiv(total, 256), blocks), 256, 0, s>>>(avg, dH, pl.n, C, Fo, partial);
    MIL_LAUNCH_OK();
    // fc.weight and fc.bias are adjacent in the flat buffer: one reduction covers both
    MIL_REQUIRE(pl.params[p_fb].offset == pl.params[p_fw].offset + (long long)Fo * C, "wide extractor: fc parameters not adjacent");
    MIL_TRY(mil_launch_reduce_partials(partial, blocks, total, gptr(p_fw), total, s));
  }
  for (int l = 3; l >= 0; --l) {
    const MilPF8& go = pl.g[l];
    for (int i = 0; i < 3; ++i) MIL_TRY(zg(gb[i], go, s));  // guards only: the pixels of dz are live
    for (int b = d.layers[l] - 1; b >= 0; --b) {
      const bool down = b == 0 && l > 0;
      const MilPF8& gi = down ? pl.g[l - 1] : pl.g[l];
      const void* xin = b > 0 ? wsp(ws, pl.off_y[l][b - 1]) : (l > 0 ? wsp(ws, pl.off_y[l - 1].back()) : wsp(ws, pl.off_pooled));
      const void* h = wsp(ws, pl.off_h[l][b]);
      const int ci = pl.first_conv[l][b];
      const MilWideConv& c1 = pl.convs[ci];
      const MilWideConv& c2 = pl.convs[ci + 1];
      MilWideShape sh;
      // conv2: weight gradient, then the data gradient through conv2 and the first activation
      MIL_TRY(wide_wgrad(h, go, dz, go, partial, gptr(c2.p_w), 3, s));
      MIL_TRY(mil_wide_shape(0, 1, c2.cout, c2.cin, 3, &sh));
      MIL_TRY(mil_launch_wide_conv(dz, go, wpk + c2.wt_off, sh, nullptr, nullptr, h, dpre, go, MIL_EPI_DGRAD, slope, 0, s));
      if (!down) {
        MIL_TRY(wide_wgrad(xin, gi, dpre, go, partial, gptr(c1.p_w), 3, s));
        MIL_TRY(mil_wide_shape(0, 1, c1.cout, c1.cin, 3, &sh));
        MIL_TRY(mil_launch_wide_conv(dpre, go, wpk + c1.wt_off, sh, nullptr, dz, xin, dnew, gi, MIL_EPI_DGRAD, slope, 0, s));
      } else {
        // stride-2 block: every gradient is computed at the OUTPUT resolution.
        //   conv1 wgrad : the nine taps grouped by the parity phase of the saved split input they read
        //   projection  : 1x1 weight gradient against phase (0, 0); data gradient into t_sub (= phase (0, 0) only)
        //   conv1 dgrad : one launch per INPUT parity phase (1 / 2 / 2 / 4 taps of the output gradient), activation mask
        //                 from the same phase of the split input, written into a phase-split gradient map, which is then
        //                 interleaved back to the full resolution
        const MilWideConv& cd = pl.convs[ci + 2];
        const MilPF8 gts = mil_pf8(pl.n, gi.c, go.h, go.w);    // one phase / the projection's data gradient
        const MilPF8 gs = mil_split2_geom(pl.n, gi.c, go.h);
        const char* xs2 = wsp(ws, pl.off_xs2[l]);
        char* dsplit = wsp(ws, pl.off_up);
        void* t_sub = wsp(ws, pl.off_tsub);
        MIL_TRY(zg(t_sub, gts, s));
        MIL_TRY(mil_launch_wide_wgrad(xs2, gs, dpre, go, partial, gptr(c1.p_w), nullptr, 3, s, 1));
        MIL_TRY(wide_wgrad(xs2, mil_split2_phase0(gs, gi.c), dz, go, partial, gptr(cd.p_w), 1, s));
        MIL_TRY(mil_wide_shape(0, 1, cd.cout, cd.cin, 1, &sh));
        MIL_TRY(mil_launch_wide_conv(dz, go, wpk + cd.wt_off, sh, nullptr, nullptr, nullptr, t_sub, gts, MIL_EPI_PLAIN, slope, 0, s));
        // all four phases in one launch: output channels (phase, ci) = the planes of the phase-split map; the projection's
        // gradient t_sub is the residual of phase (0, 0) = the first cb planes
        MIL_TRY(mil_wide_shape(7, 0, c1.cout, c1.cin, 3, &sh));
        MIL_TRY(mil_launch_wide_conv(dpre, go, wpk + pl.s2_wt_off[l][0], sh, nullptr, t_sub, xs2, dsplit, gs, MIL_EPI_DGRAD,
                                     slope, 0, s, gts.cb));
        MIL_TRY(mil_launch_merge2(dsplit, gs, dnew, gi, s));
      }
      std::swap(dz, dnew);
    }
  }
  // ---- stem: un-pool the gradient of the pooled map into the four-phase conv gradient, then the weight gradient ----
  void* dy4 = wsp(ws, pl.off_cv);
  MIL_TRY(zg(dy4, pl.gcv, s));
  MIL_TRY(mil_launch_stem_unpool4(dz, pl.g[0], wsp(ws, pl.off_argmax), dy4, pl.gcv, s));
  return wide_wgrad(wsp(ws, pl.off_xs), pl.gxs, dy4, pl.gcv, partial, gptr(pindex(pl.params, "cnn.module.conv1.weight")), 7, s);
}
