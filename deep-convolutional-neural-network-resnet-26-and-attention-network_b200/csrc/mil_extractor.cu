// Whole-network drivers of the ResNet-26 tile feature extractor (reference gbm/model.py:14-61,
// nnBlocks.py:157-189): parameter table, workspace plan, forward and backward launch sequences.
// Everything is enqueued on the caller's stream; nothing is allocated here (the caller owns the workspace).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "mil_extractor.cuh"

// ---------------------------------------------------------------------------------------------------
// parameter table (reference state-dict order)
// ---------------------------------------------------------------------------------------------------
static std::vector<MilParamInfo> build_param_table() {
  std::vector<MilParamInfo> t;
  long long off = 0;
  auto add = [&](const std::string& name, std::initializer_list<long long> shape) {
    MilParamInfo p;
    p.name = name;
    p.ndim = (int)shape.size();
    p.numel = 1;
    int i = 0;
    for (long long d : shape) { p.shape[i++] = d; p.numel *= d; }
    for (; i < 4; ++i) p.shape[i] = 1;
    p.offset = off;
    off += p.numel;
    t.push_back(p);
  };
  add("weight_mask", {3});
  add("cnn.module.conv1.weight", {20, 3, 7, 7});
  add("cnn.module.conv1.bias", {20});
  int inpl = 20;
  for (int l = 0; l < 4; ++l) {
    const int w = kMilWidths[l];
    for (int b = 0; b < 3; ++b) {
      const int cin = (b == 0) ? inpl : w;
      const std::string p = "cnn.module.layer" + std::to_string(l + 1) + "." + std::to_string(b);
      add(p + ".conv1.weight", {w, cin, 3, 3});
      add(p + ".conv1.bias", {w});
      add(p + ".conv2.weight", {w, w, 3, 3});
      add(p + ".conv2.bias", {w});
      if (b == 0 && l > 0) add(p + ".downsample.0.weight", {w, cin, 1, 1});
    }
    inpl = w;
  }
  add("cnn.module.fc.weight", {80, 80});
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
const std::vector<MilParamInfo>& mil_param_table() {
  static const std::vector<MilParamInfo> t = build_param_table();
  return t;
}
int mil_param_index(const char* name) {
  const auto& t = mil_param_table();
  for (size_t i = 0; i < t.size(); ++i)
    if (t[i].name == name) return (int)i;
  return -1;
}

// ---------------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

MilPF8 mil_xs2_geom(const MilPlan& pl, int l) {
  return pl.s2_split[l] ? mil_split2_geom(pl.n, kMilWidths[l - 1], pl.geo.h[l])
                        : mil_pf8(pl.n, kMilWidths[l - 1], pl.geo.h[l], pl.geo.h[l]);
}

int mil_make_plan(int n, int side, int dtype, MilPlan* plan, bool infer) {
  MIL_REQUIRE(n >= 1, "extractor: need at least one tile (got %d)", n);
  MIL_REQUIRE(side >= 8, "extractor: tile side %d too small", side);
  MIL_REQUIRE(dtype == MIL_F32 || dtype == MIL_BF16, "extractor: unknown dtype %d", dtype);
  MilPlan& pl = *plan;
  pl.n = n; pl.side = side; pl.dtype = dtype; pl.infer = infer;
  pl.geo = mil_geom(side);
  for (int l = 0; l < 4; ++l) pl.g[l] = mil_pf8(n, kMilWidths[l], pl.geo.h[l], pl.geo.h[l]);
  pl.convs.clear();
  size_t wofs = 0, tcofs = 0;
  int inpl = 20;
  for (int l = 0; l < 4; ++l) {
    const int w = kMilWidths[l];
    for (int b = 0; b < 3; ++b) {
      const int cin = (b == 0) ? inpl : w;
      const std::string p = "cnn.module.layer" + std::to_string(l + 1) + "." + std::to_string(b);
      for (int which = 0; which < 3; ++which) {
        if (which == 2 && !(b == 0 && l > 0)) continue;
        MilConvDesc c;
        c.layer = l; c.block = b; c.which = which;
        c.cin = (which == 1) ? w : cin;
        c.cout = w;
        c.ks = (which == 2) ? 1 : 3;
        c.stride = (which != 1 && b == 0 && l > 0) ? 2 : 1;
        const char* nm = which == 0 ? ".conv1" : (which == 1 ? ".conv2" : ".downsample.0");
        c.p_w = mil_param_index((p + nm + ".weight").c_str());
        c.p_b = which == 2 ? -1 : mil_param_index((p + nm + ".bias").c_str());
        const size_t sz = (size_t)c.ks * c.ks * ((c.cin + 7) / 8 * 8) * ((c.cout + 7) / 8 * 8);
        c.wp_off = wofs; wofs += sz;
        c.wpt_off = wofs; wofs += sz;
        c.tc = mil_tc_enabled() && mil_tc_supported(dtype, c.ks, c.stride, c.cin, c.cout);
        c.wtc_off = c.wtct_off = 0;
        c.fold = false;
        c.wtc_fold_off = 0;
        if (c.tc) {
          MilTcShape sf, sb;
          MIL_TRY(mil_tc_shape(c.cin, c.cout, c.ks, &sf));
          MIL_TRY(mil_tc_shape(c.cout, c.cin, c.ks, &sb));
          c.wtc_off = tcofs; tcofs += align_up(mil_tc_wpack_bytes(sf), 256);
          c.wtct_off = tcofs; tcofs += align_up(mil_tc_wpack_bytes(sb), 256);
          c.fold = false;
          c.wtc_fold_off = 0;
          if (which == 1 && b == 0 && l > 0) {
            MilTcShape sf2;
            MIL_TRY(mil_tc_shape_fold(c.cin, inpl, c.cout, &sf2));
            if (mil_conv_tc_fits(sf2, pl.g[l].wp)) {
              c.fold = true;
              c.wtc_fold_off = tcofs; tcofs += align_up(mil_tc_wpack_bytes(sf2), 256);
            }
          }
          for (int a = 0; a < 2; ++a) c.wtct_s2_off[a] = 0;
          if (c.ks == 3 && c.stride == 2 && 2 * ((c.cin + 7) / 8) <= 10)
            for (int a = 0; a < 2; ++a) {
              MilTcShape sp;
              MIL_TRY(mil_tc_shape_s2_dgrad(c.cout, c.cin, a, &sp));
              c.wtct_s2_off[a] = tcofs; tcofs += align_up(mil_tc_wpack_bytes(sp), 256);
            }
        }
        pl.convs.push_back(c);
      }
    }
    inpl = w;
  }
  pl.wpack_floats = wofs;
  pl.wtc_bytes = tcofs;

  // which stride-2 blocks run in the phase-split form (decides the geometry of their saved input copies, mil_xs2_geom)
  for (int l = 0; l < 4; ++l) { pl.off_xs2[l] = 0; pl.s2_split[l] = false; }
  if (dtype == MIL_BF16 && mil_tc_enabled())
    for (int l = 1; l < 4; ++l) {
      MilTcShape sh;
      MIL_TRY(mil_tc_shape_s2(kMilWidths[l - 1], kMilWidths[l], &sh));
      pl.s2_split[l] = mil_conv_tc_fits(sh, pl.g[l].wp) && 2 * ((kMilWidths[l - 1] + 7) / 8) <= 10;  // dgrad: 2 * cb output chunks
    }
  bool fold_l[4] = {false, false, false, false};
  for (const auto& c : pl.convs) if (c.fold) fold_l[c.layer] = true;

  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  pl.off_pooled = take(mil_pf8_bytes(pl.g[0], dtype));
  // max-pool arg-max: 1 byte per pooled element (CUDA-core stem) or the records of mil_stem_unpool.cuh (tensor-core stem)
  pl.off_argmax = take(infer && dtype == MIL_BF16 && mil_tc_enabled()
                           ? 256
                           : std::max((size_t)n * pl.geo.h[0] * pl.geo.h[0] * 20, mil_stem_tc_argmax_bytes(pl.g[0])));
  if (!infer) {
    for (int l = 0; l < 4; ++l)
      for (int b = 0; b < 3; ++b) {
        if (b == 0 && fold_l[l]) {
          // folded projection: conv2 reads [h | even positions of the block input] as ONE map -- the saved input copy
          // (phase (0,0) first) follows h without a gap (same geometry => same plane stride)
          const size_t hb = mil_pf8_bytes(pl.g[l], dtype);
          pl.off_h[l * 3] = take(hb + mil_pf8_bytes(mil_xs2_geom(pl, l), dtype));
          pl.off_xs2[l] = pl.off_h[l * 3] + hb;
        } else {
          pl.off_h[l * 3 + b] = take(mil_pf8_bytes(pl.g[l], dtype));
        }
        pl.off_y[l * 3 + b] = take(mil_pf8_bytes(pl.g[l], dtype));
      }
  } else {
    // forward only (attention-map extraction, gbm/classify_combined.py:221-298): nothing is kept for a backward pass.
    // The pooled map and two more buffers of the largest (layer-1) size rotate: block input -> h -> y -> next input
    const size_t big = mil_pf8_bytes(pl.g[0], dtype);
    const size_t rot[3] = {pl.off_pooled, take(big), take(big)};
    int cur = 0;  // buffer holding the block input
    for (int l = 0; l < 4; ++l)
      for (int b = 0; b < 3; ++b) {
        pl.off_h[l * 3 + b] = rot[(cur + 1) % 3];
        pl.off_y[l * 3 + b] = rot[(cur + 2) % 3];
        cur = (cur + 2) % 3;
      }
    // folded projections: [h | saved input copy] of a stride-2 block must be ONE run of planes -- an area of its own,
    // shared by the three blocks (sized for the largest)
    size_t fold_bytes = 0;
    for (int l = 1; l < 4; ++l)
      if (fold_l[l]) fold_bytes = std::max(fold_bytes, mil_pf8_bytes(pl.g[l], dtype) + mil_pf8_bytes(mil_xs2_geom(pl, l), dtype));
    if (fold_bytes > 0) {
      const size_t area = take(fold_bytes);
      for (int l = 1; l < 4; ++l)
        if (fold_l[l]) {
          pl.off_h[l * 3] = area;
          pl.off_xs2[l] = area + mil_pf8_bytes(pl.g[l], dtype);
        }
    }
  }
  // 1-bit sign masks of every saved activation map: what the data-gradient epilogues read in place of the map
  pl.masks = dtype == MIL_BF16 && mil_tc_enabled() && !infer;
  for (int i = 0; i < 12; ++i) pl.off_mh[i] = pl.off_my[i] = 0;
  if (pl.masks)
    for (int l = 0; l < 4; ++l)
      for (int b = 0; b < 3; ++b) {
        pl.off_mh[l * 3 + b] = take(mil_sign_mask_bytes(pl.g[l]));
        pl.off_my[l * 3 + b] = take(mil_sign_mask_bytes(pl.g[l]));
      }
  pl.off_mpool = 0;
  pl.pool_mask = pl.masks && mil_stem_tc_fused_pool(pl.g[0], side);
  if (pl.pool_mask) pl.off_mpool = take(mil_sign_mask_bytes(pl.g[0]));
  pl.off_avg = take((size_t)n * 80 * sizeof(float));
  pl.grad_bytes = 0;
  for (int l = 0; l < 4; ++l) pl.grad_bytes = std::max(pl.grad_bytes, mil_pf8_bytes(pl.g[l], dtype));
  if (infer) pl.grad_bytes = 0;
  for (int i = 0; i < 3; ++i) pl.off_grad[i] = take(pl.grad_bytes);
  // zero-stuffed output gradients of the three stride-2 blocks (tcgen05 path only): Cout channels at the
  // block's INPUT resolution
  pl.up_bytes = 0;
  if (dtype == MIL_BF16 && mil_tc_enabled() && !infer)
    for (int l = 1; l < 4; ++l)
      pl.up_bytes = std::max(pl.up_bytes, mil_pf8_bytes(mil_pf8(n, kMilWidths[l], pl.geo.h[l - 1], pl.geo.h[l - 1]), dtype));
  for (int i = 0; i < 2; ++i) pl.off_up[i] = take(pl.up_bytes);
  // inputs of the three stride-2 blocks split into their four (row, column) parity phases at the OUTPUT resolution
  // (what the stride-2 convolutions read; phase (0,0) = the first planes = the 1x1 projection's input), kept for
  // the backward pass
  // (a layer whose split form does not fit the convolution kernel's shared memory -- 32 input planes on the way into
  // layer 4 -- keeps the full-resolution evaluation and stores only the even positions here)
  if (dtype == MIL_BF16 && mil_tc_enabled())
    for (int l = 1; l < 4; ++l) {
      if (fold_l[l]) continue;  // allocated behind h above
      pl.off_xs2[l] = (infer && l > 1) ? pl.off_xs2[1] : take(mil_pf8_bytes(mil_xs2_geom(pl, l), dtype));  // layer 2's is the largest
    }
  pl.stem_tc = (dtype == MIL_BF16) && mil_tc_enabled();
  pl.off_xs = pl.off_cv = pl.off_stem_wp = pl.off_stem_wtc = 0;
  if (pl.stem_tc) {
    pl.off_xs = take(mil_pf8_bytes(mil_stem_tc_geom_in(n, side), dtype));
    // the 80-channel conv map exists in HBM only on the un-fused stem path (odd conv size, MIL_B200_STEM_UNFUSED=1)
    if (!mil_stem_tc_fused_pool(pl.g[0], side)) pl.off_cv = take(mil_pf8_bytes(mil_stem_tc_geom_conv(n, side), dtype));
    pl.off_stem_wp = take(mil_stem_tc_wpack_floats() * sizeof(float));
    pl.off_stem_wtc = take(mil_stem_tc_wtc_bytes());
  }
  pl.off_wpack = take(pl.wpack_floats * sizeof(float));
  pl.off_wtc = take(pl.wtc_bytes + 256);
  size_t pf = std::max(mil_stem_bwd_partial_floats(), mil_tail_bwd_partial_floats());
  size_t layer_sum[4] = {0, 0, 0, 0};
  for (const auto& c : pl.convs) {
    const MilPF8& go = pl.g[c.layer];
    const MilPF8 gi = (c.stride == 2) ? pl.g[c.layer - 1] : mil_pf8(n, c.cin, go.h, go.w);
    size_t need = mil_wgrad_direct_partial_floats(gi, go, c.ks);
    if (mil_tc_enabled() && mil_wgrad_tc_supported(dtype, c.ks, 1, c.cin, c.cout))
      need = std::max(need, mil_wgrad_tc_partial_floats(gi, mil_pf8(n, c.cout, gi.h, gi.w), c.ks));
    pf = std::max(pf, need);
    layer_sum[c.layer] += (size_t)mil_rup((long long)need, 64);
  }
  if (pl.stem_tc) pf = std::max(pf, mil_stem_tc_partial_floats(n, side));
  if (infer) pf = 64;
  pl.partial_floats = pf;
  // room for the partial records of a whole layer's weight gradients, written one after the other: their reductions run
  // as ONE launch per layer (mil_reduce_batch_*; a layer that does not fit is flushed in two launches)
  size_t arena = (size_t)mil_rup((long long)pf, 64);
  for (int l = 0; l < 4; ++l) arena = std::max(arena, layer_sum[l] + (size_t)mil_rup((long long)pf, 64));
  pl.partial_arena_floats = infer ? pf : arena;
  pl.off_partial = take(pl.partial_arena_floats * sizeof(float));
  pl.total_bytes = off;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// table-driven helpers: pack all conv weights in one launch, zero all guards in one launch
// ---------------------------------------------------------------------------------------------------
#define MIL_MAX_PACK 54
struct PackTable {
  const float* src[MIL_MAX_PACK];
  float* dst[MIL_MAX_PACK];
  short cout[MIL_MAX_PACK], cin[MIL_MAX_PACK];
  unsigned char ks[MIL_MAX_PACK], transposed[MIL_MAX_PACK];
  int count;
  int round_bf16;  // bf16 mode: the CUDA-core kernels multiply with bf16-rounded weights, like the tensor-core path
};
__global__ void pack_all_kernel(PackTable t) {
  const int e = blockIdx.x;
  const int cout = t.cout[e], cin = t.cin[e], ks = t.ks[e], tr = t.transposed[e];
  const int cip = (cin + 7) / 8 * 8, cop = (cout + 7) / 8 * 8, taps = ks * ks;
  const int A = tr ? cop : cip, B = tr ? cip : cop;
  const int total = taps * A * B;
  const float* w = t.src[e];
  float* wp = t.dst[e];
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
    const int b = i % B, a = (i / B) % A, tp = i / (A * B);
    const int co = tr ? a : b, ci = tr ? b : a;
    float v = (co < cout && ci < cin) ? w[((size_t)co * cin + ci) * taps + tp] : 0.f;
    if (t.round_bf16) v = __bfloat162float(__float2bfloat16_rn(v));
    wp[i] = v;
  }
}

#define MIL_MAX_GUARD 32
struct GuardTable {
  void* base[MIL_MAX_GUARD];
  long long PS[MIL_MAX_GUARD], G[MIL_MAX_GUARD], Q[MIL_MAX_GUARD];
  int cb[MIL_MAX_GUARD];
  int count, esize;
};
__global__ void zero_guards_kernel(GuardTable t) {
  const int e = blockIdx.x;
  const long long PS = t.PS[e], G = t.G[e], Q = t.Q[e];
  const long long unit = 8 * t.esize / 16;  // 16-byte words per pixel chunk
  uint4* base = reinterpret_cast<uint4*>(t.base[e]);
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int cb = blockIdx.y; cb < t.cb[e]; cb += gridDim.y) {
    uint4* plane = base + (size_t)cb * PS * unit;
    for (long long i = threadIdx.x; i < G * unit; i += blockDim.x) plane[i] = z;
    for (long long i = (G + Q) * unit + threadIdx.x; i < PS * unit; i += blockDim.x) plane[i] = z;
  }
}
static int launch_guards(GuardTable& t, cudaStream_t s) {
  if (t.count == 0) return 0;
  zero_guards_kernel<<<dim3(t.count, 10), 128, 0, s>>>(t);
  MIL_LAUNCH_OK();
  return 0;
}
static void guard_add(GuardTable& t, void* buf, const MilPF8& g) {
  t.base[t.count] = buf; t.PS[t.count] = g.PS; t.G[t.count] = g.G; t.Q[t.count] = g.Q; t.cb[t.count] = g.cb;
  ++t.count;
}
// zero row / zero column of every image plus the guards: what a kernel that stores only the real pixels leaves undefined
// (the stride-2 data gradient writes pixel pairs of the full-resolution map) -- instead of clearing the whole map
__global__ void zero_pads_kernel(uint4* __restrict__ base, MilPF8 g, int unit) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  const int c = blockIdx.y;
  uint4* plane = base + (size_t)c * g.PS * unit;
  if (blockIdx.x == gridDim.x - 1) {  // guards
    for (long long i = threadIdx.x; i < g.G * unit; i += blockDim.x) plane[i] = z;
    for (long long i = (g.G + g.Q) * unit + threadIdx.x; i < g.PS * unit; i += blockDim.x) plane[i] = z;
    return;
  }
  const int per = g.hp - 1 + g.wp;  // pad pixels per image: last column of rows 0..hp-2, then the whole last row
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)g.n * per;
       i += (long long)(gridDim.x - 1) * blockDim.x) {
    const int n = (int)(i / per), k = (int)(i - (long long)n * per);
    const long long q = (long long)n * g.P + (k < g.hp - 1 ? (long long)k * g.wp + g.wp - 1 : (long long)(g.hp - 1) * g.wp + (k - (g.hp - 1)));
    for (int u = 0; u < unit; ++u) plane[(g.G + q) * unit + u] = z;
  }
}
int mil_zero_pads(int dtype, void* buf, const MilPF8& g, cudaStream_t s) {
  const int unit = (int)(8 * mil_esize(dtype) / 16);
  const long long items = (long long)g.n * (g.hp - 1 + g.wp);
  const int blocks = (int)std::min<long long>(mil_cdiv(items, 256), 592) + 1;
  zero_pads_kernel<<<dim3(blocks, g.cb), 256, 0, s>>>(reinterpret_cast<uint4*>(buf), g, unit);
  MIL_LAUNCH_OK();
  return 0;
}

int mil_zero_guards(int dtype, void* buf, const MilPF8& g, cudaStream_t s) {
  GuardTable t;
  t.count = 0;
  t.esize = (int)mil_esize(dtype);
  guard_add(t, buf, g);
  return launch_guards(t, s);
}

// ---------------------------------------------------------------------------------------------------
// conv dispatch
// ---------------------------------------------------------------------------------------------------
int mil_wgrad_dispatch(int dtype, const void* x, const MilPF8& gi, const void* dz, const MilPF8& go, float* partial,
                       float* dw, float* db, int ks, int stride, cudaStream_t s) {
  if (mil_tc_enabled() && mil_wgrad_tc_supported(dtype, ks, stride, gi.c, go.c))
    return mil_launch_wgrad_tc(x, gi, dz, go, partial, dw, db, ks, s);
  return mil_launch_wgrad_direct(dtype, x, gi, dz, go, partial, dw, db, ks, stride, s);
}

// runtime switches: initial value from the environment, changed through mil_set_option (tests flip them to run the
// same bag through the tcgen05 and the CUDA-core kernels, the fused and the un-fused stem)
static const char* const kOptNames[MIL_OPT_COUNT] = {"disable_tc", "stem_unfused", "no_pdl"};
static const char* const kOptEnv[MIL_OPT_COUNT] = {"MIL_B200_DISABLE_TC", "MIL_B200_STEM_UNFUSED", "MIL_B200_NO_PDL"};
static std::atomic<int>* opt_slots() {
  static std::atomic<int> v[MIL_OPT_COUNT];
  static const bool init = [] {
    for (int i = 0; i < MIL_OPT_COUNT; ++i) {
      const char* e = getenv(kOptEnv[i]);
      v[i].store(e != nullptr ? atoi(e) : 0);
    }
    return true;
  }();
  (void)init;
  return v;
}
int mil_opt(int id) { return opt_slots()[id].load(std::memory_order_relaxed); }
int mil_opt_set(const char* name, int value) {
  for (int i = 0; i < MIL_OPT_COUNT; ++i)
    if (name != nullptr && strcmp(name, kOptNames[i]) == 0) {
      opt_slots()[i].store(value);
      return 0;
    }
  mil_set_error("mil_set_option: unknown option '%s'", name ? name : "(null)");
  return 2;
}
int mil_opt_get(const char* name, int* value) {
  for (int i = 0; i < MIL_OPT_COUNT; ++i)
    if (name != nullptr && strcmp(name, kOptNames[i]) == 0) {
      *value = mil_opt(i);
      return 0;
    }
  mil_set_error("mil_get_option: unknown option '%s'", name ? name : "(null)");
  return 2;
}

bool mil_tc_enabled() { return mil_opt(MIL_OPT_DISABLE_TC) == 0; }
static thread_local bool g_pdl_allowed = true;
bool mil_pdl_allowed() { return g_pdl_allowed; }
void mil_pdl_allow(bool on) { g_pdl_allowed = on; }

int mil_conv_dispatch(int dtype, int transposed, const void* x, const MilPF8& gi, const float* wp, const void* wtc,
                      const float* bias, const void* res, const void* act, void* out, const MilPF8& go, int ks,
                      int stride, int epi, cudaStream_t s, const void* mask_in, void* mask_out) {
  if (wtc != nullptr && mil_tc_supported(dtype, ks, stride, gi.c, go.c) && !(stride == 2 && transposed)) {
    MilTcShape sh;
    MIL_TRY(mil_tc_shape(gi.c, go.c, ks, &sh));
    return mil_launch_conv_tc(transposed, x, gi, wtc, sh, bias, res, act, out, go, epi, stride == 2, s, nullptr, -1,
                              mask_in, mask_out);
  }
  return mil_launch_conv_direct(dtype, transposed, x, gi, wp, bias, res, act, out, go, ks, stride, epi, s);
}

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
static inline char* wsp(void* ws, size_t off) { return reinterpret_cast<char*>(ws) + off; }

static int pack_weights(const void* const* params, const MilPlan& pl, void* ws, bool transposed, cudaStream_t s) {
  PackTable t;
  t.count = 0;
  t.round_bf16 = pl.dtype == MIL_BF16 ? 1 : 0;
  float* area = reinterpret_cast<float*>(wsp(ws, pl.off_wpack));
  for (const auto& c : pl.convs) {
    if (c.tc) continue;  // the tensor-core convolutions pack straight into their operand blocks below
    const int e = t.count++;
    t.src[e] = reinterpret_cast<const float*>(params[c.p_w]);
    t.dst[e] = area + (transposed ? c.wpt_off : c.wp_off);
    t.cout[e] = (short)c.cout; t.cin[e] = (short)c.cin; t.ks[e] = (unsigned char)c.ks;
    t.transposed[e] = transposed ? 1 : 0;
  }
  if (t.count > 0) {
    pack_all_kernel<<<dim3(t.count, 8), 256, 0, s>>>(t);
    MIL_LAUNCH_OK();
  }
  char* tca = wsp(ws, pl.off_wtc);
  std::vector<MilTcPackJob> jobs;
  for (const auto& c : pl.convs) {
    if (!c.tc) continue;
    const float* w = reinterpret_cast<const float*>(params[c.p_w]);
    if (c.ks == 3 && c.stride == 2 && pl.s2_split[c.layer]) {  // phase-split forms (mil_tc_shape_s2 / mil_tc_shape_s2_dgrad)
      if (!transposed) jobs.push_back({w, tca + c.wtc_off, c.cout, c.cin, 3, 0, 1});
      else {
        for (int a = 0; a < 2; ++a) jobs.push_back({w, tca + c.wtct_s2_off[a], c.cout, c.cin, 3, 1, 2 + a});
      }
      continue;
    }
    jobs.push_back({w, tca + (transposed ? c.wtct_off : c.wtc_off), c.cout, c.cin, c.ks, transposed ? 1 : 0, 0});
    if (c.fold && !transposed) {
      // conv2 + the block's 1x1 projection as one operand (the projection is the next descriptor of the block)
      const MilConvDesc& cd = *(&c + 1);
      MilTcPackJob j = {w, tca + c.wtc_fold_off, c.cout, c.cin, 3, 0, 0};
      j.w2 = reinterpret_cast<const float*>(params[cd.p_w]);
      j.cin2 = cd.cin;
      jobs.push_back(j);
    }
  }
  if (!jobs.empty()) MIL_TRY(mil_launch_pack_tc_table(jobs.data(), (int)jobs.size(), s));
  return 0;
}

int mil_extractor_forward_impl(const void* const* params, const void* bag, int bag_u8, const int* idx,
                               const MilPlan& pl, void* ws, float* H, cudaStream_t s) {
  MIL_REQUIRE(!bag_u8 || pl.stem_tc, "extractor: 8-bit tiles need the bf16 tensor-core stem (dtype bf16)");
  const MilPdlScope pdl_scope((long long)pl.n * pl.side * pl.side);
  const int dt = pl.dtype;
  // guards of every saved activation buffer (cheap, makes the workspace self-initialising)
  {
    GuardTable t;
    t.count = 0;
    t.esize = (int)mil_esize(dt);
    guard_add(t, wsp(ws, pl.off_pooled), pl.g[0]);
    if (pl.stem_tc) {
      guard_add(t, wsp(ws, pl.off_xs), mil_stem_tc_geom_in(pl.n, pl.side));
      if (!mil_stem_tc_fused_pool(pl.g[0], pl.side)) guard_add(t, wsp(ws, pl.off_cv), mil_stem_tc_geom_conv(pl.n, pl.side));
    }
    if (!pl.infer) {
      for (int l = 0; l < 4; ++l)
        for (int b = 0; b < 3; ++b) {
          guard_add(t, wsp(ws, pl.off_h[l * 3 + b]), pl.g[l]);
          guard_add(t, wsp(ws, pl.off_y[l * 3 + b]), pl.g[l]);
        }
      if (pl.dtype == MIL_BF16 && mil_tc_enabled())
        for (int l = 1; l < 4; ++l)
          guard_add(t, wsp(ws, pl.off_xs2[l]), mil_xs2_geom(pl, l));
    } else {  // rotating buffers: layer 1's geometry now, re-done at every layer boundary below
      guard_add(t, wsp(ws, pl.off_h[0]), pl.g[0]);
      guard_add(t, wsp(ws, pl.off_y[0]), pl.g[0]);
    }
    MIL_TRY(launch_guards(t, s));
  }
  MIL_TRY(pack_weights(params, pl, ws, false, s));
  const float* wpack = reinterpret_cast<const float*>(wsp(ws, pl.off_wpack));
  auto TCW = [&](const MilConvDesc& c, bool tr) -> const void* {
    return c.tc ? wsp(ws, pl.off_wtc) + (tr ? c.wtct_off : c.wtc_off) : nullptr;
  };
  const int p_c1w = mil_param_index("cnn.module.conv1.weight"), p_c1b = mil_param_index("cnn.module.conv1.bias");
  if (pl.stem_tc)
    MIL_TRY(mil_launch_stem_tc_fwd(bag, bag_u8, idx, pl.n, pl.side, (const float*)params[p_c1w], (const float*)params[p_c1b],
                                   wsp(ws, pl.off_xs), wsp(ws, pl.off_cv), (float*)wsp(ws, pl.off_stem_wp),
                                   wsp(ws, pl.off_stem_wtc), wsp(ws, pl.off_pooled), pl.g[0],
                                   pl.infer ? nullptr : (uint8_t*)wsp(ws, pl.off_argmax), s,
                                   pl.pool_mask ? wsp(ws, pl.off_mpool) : nullptr));
  else
    MIL_TRY(mil_launch_stem_fwd(dt, (const float*)bag, idx, pl.n, pl.side, (const float*)params[p_c1w],
                                (const float*)params[p_c1b], wsp(ws, pl.off_pooled), pl.g[0],
                                (uint8_t*)wsp(ws, pl.off_argmax), s));
  const void* X = wsp(ws, pl.off_pooled);
  MilPF8 gx = pl.g[0];
  size_t ci = 0;
  for (int l = 0; l < 4; ++l) {
    if (pl.infer && l > 0) {
      // forward-only plan: the block's outputs land in buffers that held maps of the previous layer's geometry
      GuardTable t;
      t.count = 0;
      t.esize = (int)mil_esize(dt);
      guard_add(t, wsp(ws, pl.off_h[l * 3]), pl.g[l]);
      guard_add(t, wsp(ws, pl.off_y[l * 3]), pl.g[l]);
      // block 1 writes its y into the second free rotating buffer: block 0's h unless that lives behind a folded projection's
      // input (then nothing has given this buffer the layer's geometry yet)
      if (pl.off_y[l * 3 + 1] != pl.off_h[l * 3]) guard_add(t, wsp(ws, pl.off_y[l * 3 + 1]), pl.g[l]);
      if (pl.dtype == MIL_BF16 && mil_tc_enabled()) guard_add(t, wsp(ws, pl.off_xs2[l]), mil_xs2_geom(pl, l));
      MIL_TRY(launch_guards(t, s));
    }
    for (int b = 0; b < 3; ++b) {
      if (pl.infer && l > 0 && b == 1) {
        // the third rotating buffer still has the previous layer's geometry (it held this layer's block-0 input)
        MIL_TRY(mil_zero_guards(dt, wsp(ws, pl.off_h[l * 3 + 1]), pl.g[l], s));
      }
      const MilPF8& go = pl.g[l];
      void* h = wsp(ws, pl.off_h[l * 3 + b]);
      void* y = wsp(ws, pl.off_y[l * 3 + b]);
      void* mh = pl.masks ? wsp(ws, pl.off_mh[l * 3 + b]) : nullptr;  // sign masks of h and y for the backward pass
      void* my = pl.masks ? wsp(ws, pl.off_my[l * 3 + b]) : nullptr;
      const MilConvDesc& c1 = pl.convs[ci++];
      const MilConvDesc& c2 = pl.convs[ci++];
      if (c1.stride == 2 && c1.tc && !pl.s2_split[l]) {
        if ((gx.h & 1) || (gx.w & 1)) {
          // the subsampled store of the full-resolution evaluation reaches the zero row / column of the
          // half-resolution map only when the input size is even: clear the map first
          MIL_CHECK_CUDA(cudaMemsetAsync(h, 0, mil_pf8_bytes(go, dt), s));
        }
        MIL_TRY(mil_launch_subsample2(X, gx, wsp(ws, pl.off_xs2[l]), mil_xs2_geom(pl, l), s));
        MIL_TRY(mil_conv_dispatch(dt, 0, X, gx, wpack + c1.wp_off, TCW(c1, false), (const float*)params[c1.p_b], nullptr, nullptr,
                                  h, go, 3, c1.stride, MIL_EPI_FWD, s, nullptr, mh));
      } else if (c1.stride == 2 && c1.tc) {
        // stride-2 block on the tensor cores: split the input into its four parity phases once; the 3x3 / stride-2
        // convolution is then a 2x2-window convolution over 4x the channels at the OUTPUT resolution
        const MilPF8 gs = mil_split2_geom(pl.n, gx.c, go.h);
        void* xs2 = wsp(ws, pl.off_xs2[l]);
        MIL_TRY(mil_launch_split2(X, gx, xs2, gs, s));
        MilTcShape sh;
        MIL_TRY(mil_tc_shape_s2(c1.cin, c1.cout, &sh));
        MIL_TRY(mil_launch_conv_tc(0, xs2, gs, TCW(c1, false), sh, (const float*)params[c1.p_b], nullptr, nullptr, h, go,
                                   MIL_EPI_FWD, 0, s, nullptr, -1, nullptr, mh));
      } else
        MIL_TRY(mil_conv_dispatch(dt, 0, X, gx, wpack + c1.wp_off, TCW(c1, false), (const float*)params[c1.p_b], nullptr, nullptr,
                                  h, go, 3, c1.stride, MIL_EPI_FWD, s, nullptr, c1.tc ? mh : nullptr));
      const void* res = X;
      if (b == 0 && l > 0 && c2.fold) {
        // the 1x1 / stride-2 projection shortcut as extra K-groups of conv2: the kernel reads h and the
        // even positions of the block input (the first planes of the saved copy, which follows h in memory) as one map;
        // no projection launch, no residual map
        const MilConvDesc& cd = pl.convs[ci++];
        MilTcShape shf;
        MIL_TRY(mil_tc_shape_fold(c2.cin, cd.cin, c2.cout, &shf));
        MilPF8 gf = go;
        gf.cb = shf.cbin;
        gf.c = shf.cbin * 8;
        MIL_REQUIRE(pl.off_xs2[l] == pl.off_h[l * 3] + mil_pf8_bytes(go, dt) && mil_xs2_geom(pl, l).PS == go.PS,
                    "extractor: folded projection needs its input behind h");
        MIL_TRY(mil_launch_conv_tc(0, h, gf, wsp(ws, pl.off_wtc) + c2.wtc_fold_off, shf, (const float*)params[c2.p_b], nullptr,
                                   nullptr, y, go, MIL_EPI_FWD, 0, s, nullptr, -1, nullptr, my));
        X = y;
        gx = go;
        continue;
      }
      if (b == 0 && l > 0) {
        const MilConvDesc& cd = pl.convs[ci++];
        // projection shortcut (1x1 / stride 2, no bias) written into y, then consumed in place as the residual
        if (cd.tc) {
          // tensor-core path: phase (0,0) of the split input = the even positions; a plain 1x1 conv on it
          const MilPF8 gxs = mil_split2_phase0(mil_xs2_geom(pl, l), gx.c);
          void* xsub = wsp(ws, pl.off_xs2[l]);
          MIL_TRY(mil_conv_dispatch(dt, 0, xsub, gxs, wpack + cd.wp_off, TCW(cd, false), nullptr, nullptr, nullptr, y, go, 1,
                                    1, MIL_EPI_PLAIN, s));
        } else
          MIL_TRY(mil_conv_dispatch(dt, 0, X, gx, wpack + cd.wp_off, nullptr, nullptr, nullptr, nullptr, y, go, 1, 2,
                                    MIL_EPI_PLAIN, s));
        res = y;
      }
      MIL_TRY(mil_conv_dispatch(dt, 0, h, go, wpack + c2.wp_off, TCW(c2, false), (const float*)params[c2.p_b], res, nullptr, y, go,
                                3, 1, MIL_EPI_FWD, s, nullptr, c2.tc ? my : nullptr));
      X = y;
      gx = go;
    }
  }
  const int p_fc = mil_param_index("cnn.module.fc.weight");
  return mil_launch_tail_fwd(dt, X, pl.g[3], (const float*)params[p_fc], (float*)wsp(ws, pl.off_avg), H, s);
}

// ---------------------------------------------------------------------------------------------------
// backward (weight gradients are ACCUMULATED into `grads`, the flat state-dict-ordered buffer)
// ---------------------------------------------------------------------------------------------------
// test / debugging aid: dump one intermediate gradient of the next backward pass as fp32 NCHW
static struct { int layer, block, which; float* dst; } g_dump = {-1, -1, 0, nullptr};
void mil_debug_request_dump(int layer, int block, int which, float* dst) {
  g_dump.layer = layer; g_dump.block = block; g_dump.which = which; g_dump.dst = dst;
}

int mil_extractor_backward_impl(const void* const* params, const float* bag, const int* idx, const MilPlan& pl,
                                void* ws, const float* dH, float* grads, cudaStream_t s,
                                const cudaEvent_t* layer_events) {
  const MilPdlScope pdl_scope((long long)pl.n * pl.side * pl.side);
  const int dt = pl.dtype;
  const auto& pt = mil_param_table();
  MIL_TRY(pack_weights(params, pl, ws, true, s));
  const float* wpack = reinterpret_cast<const float*>(wsp(ws, pl.off_wpack));
  auto TCW = [&](const MilConvDesc& c, bool tr) -> const void* {
    return c.tc ? wsp(ws, pl.off_wtc) + (tr ? c.wtct_off : c.wtc_off) : nullptr;
  };
  float* partial = reinterpret_cast<float*>(wsp(ws, pl.off_partial));
  void* gb[3] = {wsp(ws, pl.off_grad[0]), wsp(ws, pl.off_grad[1]), wsp(ws, pl.off_grad[2])};
  auto gptr = [&](int p) { return grads + pt[p].offset; };

  const int p_fc = mil_param_index("cnn.module.fc.weight");
  void* dz = gb[0];
  void* dpre = gb[1];
  void* dnew = gb[2];

  {
    GuardTable t;
    t.count = 0;
    t.esize = (int)mil_esize(dt);
    for (int i = 0; i < 3; ++i) guard_add(t, gb[i], pl.g[3]);
    MIL_TRY(launch_guards(t, s));
  }
  MIL_TRY(mil_launch_tail_bwd(dt, wsp(ws, pl.off_y[11]), pl.g[3], (const float*)params[p_fc],
                              (const float*)wsp(ws, pl.off_avg), dH, dz, partial, gptr(p_fc), s));
  // the reductions of a layer's weight gradients are recorded and run as one launch at the end of the layer
  struct BatchScope {
    ~BatchScope() { mil_reduce_batch_begin(nullptr, 0, 0); }  // closes the batch on every exit path
  } batch_scope;
  mil_reduce_batch_begin(partial, pl.partial_arena_floats, pl.partial_floats);
#define MIL_PBUF(var)                                         \
  float* var = mil_reduce_batch_cursor(partial, s);           \
  if (var == nullptr) return 1
  // index of the first conv descriptor of (l, b)
  auto conv_base = [&](int l, int b) {
    size_t i = 0;
    for (; i < pl.convs.size(); ++i)
      if (pl.convs[i].layer == l && pl.convs[i].block == b) break;
    return i;
  };
  for (int l = 3; l >= 0; --l) {
    for (int b = 2; b >= 0; --b) {
      const MilPF8& go = pl.g[l];
      const bool down = (b == 0 && l > 0);
      const void* xin = (b > 0) ? wsp(ws, pl.off_y[l * 3 + b - 1])
                                : (l > 0 ? wsp(ws, pl.off_y[(l - 1) * 3 + 2]) : wsp(ws, pl.off_pooled));
      const MilPF8& gi = down ? pl.g[l - 1] : pl.g[l];
      const void* h = wsp(ws, pl.off_h[l * 3 + b]);
      // sign masks: of h (this block's first activation) and of the block input (the previous block's output; the
      // pooled stem output has none -> the activation itself is read)
      const void* mh = pl.masks ? wsp(ws, pl.off_mh[l * 3 + b]) : nullptr;
      const void* mx = !pl.masks ? nullptr
                       : (b > 0 ? wsp(ws, pl.off_my[l * 3 + b - 1])
                                : (l > 0 ? wsp(ws, pl.off_my[(l - 1) * 3 + 2])
                                         : (pl.pool_mask ? wsp(ws, pl.off_mpool) : nullptr)));
      const size_t cb = conv_base(l, b);
      const MilConvDesc& c1 = pl.convs[cb];
      const MilConvDesc& c2 = pl.convs[cb + 1];
      // conv2: weight gradient, then data gradient through conv2 and the first LeakyReLU
      {
        MIL_PBUF(pb);
        MIL_TRY(mil_wgrad_dispatch(dt, h, go, dz, go, pb, gptr(c2.p_w), gptr(c2.p_b), 3, 1, s));
      }
      MIL_TRY(mil_conv_dispatch(dt, 1, dz, go, wpack + c2.wpt_off, TCW(c2, true), nullptr, nullptr, h, dpre, go, 3, 1,
                                MIL_EPI_DGRAD, s, c2.tc ? mh : nullptr));
      // which 1: gradient w.r.t. the pre-activation of this block's first conv (geometry go)
      if (g_dump.dst != nullptr && g_dump.layer == l && g_dump.block == b && g_dump.which == 1)
        MIL_TRY(mil_launch_from_pf8(dt, dpre, g_dump.dst, go.n, go.c, go.h, go.w, s));
      // conv1: weight gradient (the stride-2 blocks do it inside their own branch below)
      if (!down) {
        MIL_PBUF(pb);
        MIL_TRY(mil_wgrad_dispatch(dt, xin, gi, dpre, go, pb, gptr(c1.p_w), gptr(c1.p_b), 3, 1, s));
      }
      if (down && c1.tc && pl.s2_split[l]) {
        // stride-2 block, phase-split form: every gradient of the block is computed at the OUTPUT resolution.
        //   conv1 wgrad : nine single-tap MMAs over the saved phase-split input
        //   projection  : 1x1 wgrad against phase (0,0), data gradient into t_sub (half resolution)
        //   conv1 dgrad : one launch per input ROW parity (2 / 4 taps, both column parities as output chunks), stored
        //                 straight into the full-resolution map as pixel pairs; the even-even pixels add t_sub; the
        //                 LeakyReLU' mask is read from the block input
        const MilConvDesc& cd = pl.convs[cb + 2];
        const MilPF8 gs = mil_xs2_geom(pl, l);
        const MilPF8 gxs = mil_split2_phase0(gs, gi.c);
        const MilPF8 gts = mil_pf8(pl.n, gi.c, go.h, go.w);
        void* xs2 = wsp(ws, pl.off_xs2[l]);
        void* t_sub = wsp(ws, pl.off_up[1]);
        MIL_TRY(mil_zero_guards(dt, t_sub, gts, s));
        MIL_TRY(mil_zero_pads(dt, dnew, gi, s));  // pads + guards of the full-resolution map (its pixels are all written below)
        {
          MIL_PBUF(pb);
          MIL_TRY(mil_launch_wgrad_tc_s2(xs2, gs, dpre, go, pb, gptr(c1.p_w), gptr(c1.p_b), c1.cin, s));
        }
        {
          MIL_PBUF(pb);
          MIL_TRY(mil_wgrad_dispatch(dt, xs2, gxs, dz, go, pb, gptr(cd.p_w), nullptr, 1, 1, s));
        }
        MIL_TRY(mil_conv_dispatch(dt, 1, dz, go, wpack + cd.wpt_off, TCW(cd, true), nullptr, nullptr, nullptr, t_sub, gts,
                                  1, 1, MIL_EPI_PLAIN, s));
        for (int a = 0; a < 2; ++a) {
          MilTcShape sh;
          MIL_TRY(mil_tc_shape_s2_dgrad(c1.cout, c1.cin, a, &sh));
          MIL_TRY(mil_launch_conv_tc(1, dpre, go, wsp(ws, pl.off_wtc) + c1.wtct_s2_off[a], sh, nullptr,
                                     a == 0 ? t_sub : nullptr, xin, dnew, gi, MIL_EPI_DGRAD, 0, s,
                                     a == 0 ? &gts : nullptr, a, mx));
        }
        {
          GuardTable t;
          t.count = 0;
          t.esize = (int)mil_esize(dt);
          guard_add(t, dz, gi);
          guard_add(t, dpre, gi);
          MIL_TRY(launch_guards(t, s));
        }
      } else if (down && c1.tc) {
        // stride-2 block on the tensor-core kernels.  The 3x3 conv: zero-stuff its output gradient to the input
        // resolution, after which its gradients are stride-1 problems.  The 1x1 projection: everything stays at
        // the OUTPUT resolution (weight gradient against the even-position input saved by the forward pass, data
        // gradient into t_sub), and the full-resolution dgrad epilogue adds t_sub at the even positions.
        const MilConvDesc& cd = pl.convs[cb + 2];
        const MilPF8 gu = mil_pf8(pl.n, go.c, gi.h, gi.w);
        const MilPF8 gxs = mil_split2_phase0(mil_xs2_geom(pl, l), gi.c);
        const MilPF8 gts = mil_pf8(pl.n, gi.c, go.h, go.w);  // geometry of t_sub (a tensor of its own)
        void* up_pre = wsp(ws, pl.off_up[0]);
        void* t_sub = wsp(ws, pl.off_up[1]);
        {
          GuardTable t;
          t.count = 0;
          t.esize = (int)mil_esize(dt);
          guard_add(t, up_pre, gu);
          guard_add(t, t_sub, gts);
          guard_add(t, dnew, gi);
          MIL_TRY(launch_guards(t, s));
        }
        MIL_TRY(mil_launch_upsample2(dpre, go, up_pre, gu, s));
        {
          MIL_PBUF(pb);
          MIL_TRY(mil_wgrad_dispatch(dt, xin, gi, up_pre, gu, pb, gptr(c1.p_w), gptr(c1.p_b), 3, 1, s));
        }
        {
          MIL_PBUF(pb);
          MIL_TRY(mil_wgrad_dispatch(dt, wsp(ws, pl.off_xs2[l]), gxs, dz, go, pb, gptr(cd.p_w), nullptr, 1, 1, s));
        }
        MIL_TRY(mil_conv_dispatch(dt, 1, dz, go, wpack + cd.wpt_off, TCW(cd, true), nullptr, nullptr, nullptr, t_sub, gts,
                                  1, 1, MIL_EPI_PLAIN, s));
        {
          MilTcShape sh;
          MIL_TRY(mil_tc_shape(gu.c, gi.c, 3, &sh));
          MIL_TRY(mil_launch_conv_tc(1, up_pre, gu, TCW(c1, true), sh, nullptr, t_sub, xin, dnew, gi, MIL_EPI_DGRAD, 0, s,
                                     &gts, -1, mx));
        }
        {
          GuardTable t;
          t.count = 0;
          t.esize = (int)mil_esize(dt);
          guard_add(t, dz, gi);
          guard_add(t, dpre, gi);
          MIL_TRY(launch_guards(t, s));
        }
      } else if (down) {
        // CUDA-core path (fp32 check mode): strided weight gradients and transposed stride-2 convolutions
        const MilConvDesc& cd = pl.convs[cb + 2];
        {
          MIL_PBUF(pb);
          MIL_TRY(mil_wgrad_dispatch(dt, xin, gi, dpre, go, pb, gptr(c1.p_w), gptr(c1.p_b), 3, c1.stride, s));
        }
        {
          MIL_PBUF(pb);
          MIL_TRY(mil_wgrad_dispatch(dt, xin, gi, dz, go, pb, gptr(cd.p_w), nullptr, 1, 2, s));
        }
        {
          GuardTable t;
          t.count = 0;
          t.esize = (int)mil_esize(dt);
          guard_add(t, dnew, gi);
          MIL_TRY(launch_guards(t, s));
        }
        MIL_TRY(mil_conv_dispatch(dt, 1, dz, go, wpack + cd.wpt_off, nullptr, nullptr, nullptr, nullptr, dnew, gi, 1, 2,
                                  MIL_EPI_PLAIN, s));
        MIL_TRY(mil_conv_dispatch(dt, 1, dpre, go, wpack + c1.wpt_off, nullptr, nullptr, dnew, xin, dnew, gi, 3, 2,
                                  MIL_EPI_DGRAD, s));
        {
          GuardTable t;
          t.count = 0;
          t.esize = (int)mil_esize(dt);
          guard_add(t, dz, gi);
          guard_add(t, dpre, gi);
          MIL_TRY(launch_guards(t, s));
        }
      } else {
        MIL_TRY(mil_conv_dispatch(dt, 1, dpre, go, wpack + c1.wpt_off, TCW(c1, true), nullptr, dz, xin, dnew, gi, 3, 1,
                                  MIL_EPI_DGRAD, s, c1.tc ? mx : nullptr));
      }
      // which 0: gradient w.r.t. the pre-activation feeding this block's input (geometry gi)
      if (g_dump.dst != nullptr && g_dump.layer == l && g_dump.block == b && g_dump.which == 0)
        MIL_TRY(mil_launch_from_pf8(dt, dnew, g_dump.dst, gi.n, gi.c, gi.h, gi.w, s));
      std::swap(dz, dnew);
    }
    MIL_TRY(mil_reduce_batch_flush(s));
    // every gradient of layer l+1 (and, for l = 3, of fc and the head) is final: let the caller start reducing it
    if (layer_events != nullptr && layer_events[l] != nullptr) MIL_CHECK_CUDA(cudaEventRecord(layer_events[l], s));
  }

#undef MIL_PBUF
  MIL_TRY(mil_reduce_batch_end(s));
  const int p_c1w = mil_param_index("cnn.module.conv1.weight"), p_c1b = mil_param_index("cnn.module.conv1.bias");
  if (pl.stem_tc)  // the conv-map buffer of the forward pass is free by now: it takes the dense conv-resolution gradient
    return mil_launch_stem_tc_bwd(wsp(ws, pl.off_xs), pl.n, pl.side, dz, pl.g[0],
                                  (const uint8_t*)wsp(ws, pl.off_argmax), wsp(ws, pl.off_cv), partial, gptr(p_c1w),
                                  gptr(p_c1b), s);
  return mil_launch_stem_bwd(dt, bag, idx, pl.n, pl.side, dz, pl.g[0], (const uint8_t*)wsp(ws, pl.off_argmax),
                             partial, gptr(p_c1w), gptr(p_c1b), s);
}
