// tcgen05 weight gradient for the WIDE extractor parameterisation (128 .. 512 channels; alt_resnet.py:24-67 has no
// convolution bias, so there is no bias-gradient column here) and for its 7x7 / stride-2 stem in space-to-depth form.
// Reference semantics: autograd of alt_resnet.py:52-67 / :125-139 (gbm/classify_combined.py:447).
//
//   dW[tap][ci][co] = sum over flat pixels q of  x[q + shift_tap][ci] * dz[q][co]
//
// GEMM view:  D[co (M = 128)][ci (N)] += A[co][k] * B[ci][k],  k = flat pixel -- both operands exactly as PF8 stores
// them (MN-major core matrices of 8 pixels x 16 B), a tap = a start-address shift of the x window (mil_wgrad_tc.cu).
// A 3x3 layer of C >= 128 channels needs 9 * C * C accumulators, far more than the 128 x 512 TMEM holds, so the work
// is cut into KINDS = (128 output channels, up to 128 input channels, one ROW of taps): 3 * N <= 384 TMEM columns,
// every MMA is M = 128 x N = 128 (the tensor pipe's full rate), and the CTAs of a kind split the pixels (split-K) and
// write one partial record each; a fixed-order reduction sums them into the PyTorch layout (deterministic).
// The stem (48 input channels) keeps all nine taps in one kind (9 * 48 = 432 columns).
#include <algorithm>

#include "mil_common.cuh"
#include "mil_conv_tc.cuh"
#include "mil_tc_ptx.cuh"
#include "mil_wide.cuh"

#define WW_TK 128
#define WW_THREADS 192  // warp 0 producer, warp 1 MMA issuer, warps 2..5 epilogue
#define WW_MAX_STAGES 4
#define WW_A_PLANE (WW_TK * 16)

struct WwSmemHeader {
  uint64_t full[WW_MAX_STAGES], empty[WW_MAX_STAGES], done;
  uint32_t tmem_base;
};

struct WwParams {
  int n_cot, n_cit, n_tg;  // kinds = output-channel tiles x input-channel tiles x tap groups (blockIdx.y)
  int cit;                 // input channels per tile (N of the MMAs, a multiple of 16)
  int tg_ntaps[3], tg_tap0[3];  // taps of each group: tg_tap0 .. tg_tap0 + tg_ntaps - 1 of the shift table
  int shift[9];            // pixel shift of every tap
  int n_stages;
  long long rec_floats;    // floats per partial record
};

__global__ void __launch_bounds__(WW_THREADS, 1)
wide_wgrad_kernel(const __nv_bfloat16* __restrict__ x, MilPF8 gx, const __nv_bfloat16* __restrict__ dz, MilPF8 gz,
                  float* __restrict__ partial, const __grid_constant__ WwParams wp_) {
  const WwParams& pr = wp_;
  extern __shared__ __align__(128) unsigned char smem[];
  WwSmemHeader* hd = reinterpret_cast<WwSmemHeader*>(smem);
  unsigned char* stage0 = smem + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // kind -> (output tile, input tile, tap group)
  const int kind = blockIdx.y;
  const int tg = kind % pr.n_tg, cit_i = (kind / pr.n_tg) % pr.n_cit, cot_i = kind / (pr.n_tg * pr.n_cit);
  const int ntaps = pr.tg_ntaps[tg], tap0 = pr.tg_tap0[tg];
  int smin = pr.shift[tap0], smax = pr.shift[tap0];
  for (int t = 1; t < ntaps; ++t) { smin = min(smin, pr.shift[tap0 + t]); smax = max(smax, pr.shift[tap0 + t]); }
  const int span = WW_TK + (smax - smin);
  const int ncic = pr.cit >> 3;
  // output-channel tile: 16 chunk planes of dz, fewer in the last tile of a 64-channel layer (the MMA still reads 128
  // rows: whatever the missing planes' slots hold lands in accumulator rows nobody reads)
  const int ncoc = min(16, gz.cb - cot_i * 16);
  const uint32_t a_bytes = 16 * WW_A_PLANE, b_plane = (uint32_t)span * 16;
  const uint32_t stage_bytes = a_bytes + (uint32_t)ncic * b_plane;
  const uint32_t load_bytes = (uint32_t)ncoc * WW_A_PLANE + (uint32_t)ncic * b_plane;
  const long long n_tiles = mil_cdiv(gz.Q, WW_TK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < pr.n_stages; ++s) { mbar_init(&hd->full[s], 1); mbar_init(&hd->empty[s], 1); }
    mbar_init(&hd->done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&hd->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hd->tmem_base;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t s0 = smem_u32(stage0), full0 = smem_u32(&hd->full[0]);
    const char* srca = reinterpret_cast<const char*>(dz) + ((long long)cot_i * 16 * gz.PS + gz.G + (long long)blockIdx.x * WW_TK) * 16;
    const char* srcb = reinterpret_cast<const char*>(x) + ((long long)cit_i * ncic * gx.PS + gx.G + smin + (long long)blockIdx.x * WW_TK) * 16;
    const long long stra = gz.PS * 16, strb = gx.PS * 16, tstride = (long long)gridDim.x * WW_TK * 16;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&hd->empty[stage], phase ^ 1);
      if (elect_one()) {
        const uint32_t bar = full0 + (uint32_t)stage * 8;
        mbar_expect_tx_u32(bar, load_bytes);
        uint32_t dst = s0 + (uint32_t)stage * stage_bytes;
        const char* sp = srca;
        for (int c = 0; c < ncoc; ++c, dst += WW_A_PLANE, sp += stra) bulk_g2s_u32(dst, sp, WW_A_PLANE, bar);
        dst = s0 + (uint32_t)stage * stage_bytes + a_bytes;
        sp = srcb;
        for (int c = 0; c < ncic; ++c, dst += b_plane, sp += strb) bulk_g2s_u32(dst, sp, b_plane, bar);
      }
      __syncwarp();
      srca += tstride; srcb += tstride;
      if (++stage == pr.n_stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // D = f32, A = B = bf16, both MN-major (bits 15, 16), M = 128, N = cit
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24) |
                           ((uint32_t)(pr.cit >> 3) << 17);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&hd->full[stage], phase);
      tc_fence_after();
      const uint32_t a_base = smem_u32(stage0 + (size_t)stage * stage_bytes);
      const uint32_t b_base = a_base + a_bytes;
      const uint64_t ad0 = make_desc(a_base, 128, WW_A_PLANE);
      const uint32_t acc0 = first ? 0u : 1u;
      if (elect_one()) {
        for (int tl = 0; tl < ntaps; ++tl) {
          const uint64_t bd0 = make_desc(b_base + (uint32_t)(pr.shift[tap0 + tl] - smin) * 16, 128, b_plane);
          const uint32_t d = tmem_base + (uint32_t)(tl * pr.cit);
          umma_bf16(d, ad0, bd0, idesc, acc0);
#pragma unroll
          for (int kk = 1; kk < WW_TK / 16; ++kk) umma_bf16(d, ad0 + kk * 16, bd0 + kk * 16, idesc, 1u);
        }
        umma_commit(&hd->empty[stage]);
      }
      __syncwarp();
      first = false;
      if (++stage == pr.n_stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(&hd->done);
    __syncwarp();
  } else {
    // epilogue: TMEM lane = output channel of the tile, columns = (local tap, input channel of the tile)
    const int quarter = warp & 3;
    const int co = quarter * 32 + lane;
    mbar_wait(&hd->done, 0);
    tc_fence_after();
    float* rec = partial + ((size_t)kind * gridDim.x + blockIdx.x) * pr.rec_floats;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool any = blockIdx.x < n_tiles;  // a CTA without a tile has accumulated nothing: its record is zero
    for (int tl = 0; tl < ntaps; ++tl)
      for (int c = 0; c < ncic; ++c) {
        float v[8];
        tmem_ld8(taddr + (uint32_t)(tl * pr.cit + c * 8), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) rec[((size_t)tl * pr.cit + c * 8 + j) * 128 + co] = any ? v[j] : 0.f;
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- reduction of the partial records into the PyTorch layout (+=) -----------------------------------------------
// one thread per (local tap, input channel, output channel) of a kind, fixed-order sum over the kind's split-K records
struct WwReduceParams {
  WwParams p;
  int nsplit;        // records per kind
  int cout, cin, ks; // the PyTorch weight [cout][cin][ks][ks];  ks = 7: stem, folded back from the space-to-depth form
};
__global__ void __launch_bounds__(256)
wide_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, const __grid_constant__ WwReduceParams rp) {
  const WwParams& pr = rp.p;
  const int kind = blockIdx.y;
  const int tg = kind % pr.n_tg, cit_i = (kind / pr.n_tg) % pr.n_cit, cot_i = kind / (pr.n_tg * pr.n_cit);
  const int ntaps = pr.tg_ntaps[tg], tap0 = pr.tg_tap0[tg];
  const int per = ntaps * pr.cit * 128;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per) return;
  const int col = i & 127, cl = (i >> 7) % pr.cit, tl = (i >> 7) / pr.cit;
  const float* src = partial + (size_t)kind * rp.nsplit * pr.rec_floats + i;
  float acc = 0.f;
  for (int k = 0; k < rp.nsplit; ++k) acc += src[(size_t)k * pr.rec_floats];
  const int tap = tap0 + tl;
  const int kc = cit_i * pr.cit + cl;  // kernel-side input channel
  const int nc = cot_i * 128 + col;    // kernel-side output channel
  if (nc < rp.cout && kc < rp.cin) dw[((size_t)nc * rp.cin + kc) * (rp.ks * rp.ks) + tap] += acc;
}

// stem: the records hold the space-to-depth form [tap (dy, dx)][(c, ry, rx)][(co, a, b)]; weight element
// w[co][c][ky][kx] collects the four conv phases (a, b) with 2a + ky - 3 = 4 dy + ry (mil_stem_tc.cu).  One thread per
// weight element, fixed summation order.
__global__ void __launch_bounds__(256)
wide_stem_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, const __grid_constant__ WwReduceParams rp) {
  const WwParams& pr = rp.p;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rp.cout * 147) return;
  const int kx = i % 7, ky = (i / 7) % 7, c = (i / 49) % 3, co = i / 147;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int uy = 2 * a + ky - 3;
    const int dy = (uy + 4) / 4 - 1, ry = uy - 4 * dy;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ux = 2 * b + kx - 3;
      const int dx = (ux + 4) / 4 - 1, rx = ux - 4 * dx;
      const int t = (dy + 1) * 3 + (dx + 1), cc = (c * 4 + ry) * 4 + rx, co4 = co * 4 + a * 2 + b;
      const int kind = co4 >> 7, col = co4 & 127;  // one input tile, one tap group: kind = output tile
      const float* src = partial + (size_t)kind * rp.nsplit * pr.rec_floats + ((size_t)t * 48 + cc) * 128 + col;
      for (int k = 0; k < rp.nsplit; ++k) acc += src[(size_t)k * pr.rec_floats];
    }
  }
  dw[i] += acc;
}

// ---- host side ---------------------------------------------------------------------------------------------
static int ww_sm_count() {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      n_sm = 148;
  }
  return n_sm;
}

struct WwConfig {
  WwParams p;
  int kinds, nsplit;
  size_t smem;
};

static int ww_config(const MilPF8& gx, const MilPF8& gz, int ks, WwConfig* out) {
  WwConfig& c = *out;
  c = WwConfig{};
  const int cin = gx.cb * 8, cout = gz.cb * 8;
  MIL_REQUIRE(ks == 1 || ks == 3 || ks == 7, "wide_wgrad: unsupported window %d", ks);
  MIL_REQUIRE(cout % 64 == 0, "wide_wgrad: %d output channels (need a multiple of 64)", cout);
  const int r = ks == 1 ? 0 : 1;  // ks = 7: the stem's space-to-depth form has 3x3 taps
  int nt = 0;
  for (int a = -r; a <= r; ++a)
    for (int b = -r; b <= r; ++b) c.p.shift[nt++] = a * gx.wp + b;
  if (ks == 7) {
    MIL_REQUIRE(cin == 48, "wide_wgrad: the stem form reads the 48-channel space-to-depth input");
    c.p.cit = 48; c.p.n_cit = 1; c.p.n_tg = 1;
    c.p.tg_ntaps[0] = 9; c.p.tg_tap0[0] = 0;
  } else {
    c.p.cit = cin >= 128 ? 128 : cin;
    MIL_REQUIRE(cin % c.p.cit == 0 && c.p.cit % 16 == 0, "wide_wgrad: %d input channels", cin);
    c.p.n_cit = cin / c.p.cit;
    c.p.n_tg = ks == 3 ? 3 : 1;
    for (int g = 0; g < c.p.n_tg; ++g) { c.p.tg_ntaps[g] = ks == 3 ? 3 : 1; c.p.tg_tap0[g] = 3 * g; }
  }
  c.p.n_cot = (cout + 127) / 128;
  c.kinds = c.p.n_cot * c.p.n_cit * c.p.n_tg;
  c.p.rec_floats = (long long)c.p.tg_ntaps[0] * c.p.cit * 128;
  const long long n_tiles = mil_cdiv(gz.Q, WW_TK);
  c.nsplit = (int)std::max<long long>(1, std::min<long long>(n_tiles, ww_sm_count() / c.kinds));
  int span_max = 0;
  for (int g = 0; g < c.p.n_tg; ++g) {
    int lo = c.p.shift[c.p.tg_tap0[g]], hi = lo;
    for (int t = 1; t < c.p.tg_ntaps[g]; ++t) {
      lo = std::min(lo, c.p.shift[c.p.tg_tap0[g] + t]);
      hi = std::max(hi, c.p.shift[c.p.tg_tap0[g] + t]);
    }
    span_max = std::max(span_max, WW_TK + hi - lo);
  }
  const size_t stage = (size_t)16 * WW_A_PLANE + (size_t)(c.p.cit / 8) * span_max * 16;
  c.p.n_stages = WW_MAX_STAGES;
  while (c.p.n_stages > 1 && 128 + c.p.n_stages * stage > 224 * 1024) --c.p.n_stages;
  c.smem = 128 + c.p.n_stages * stage;
  MIL_REQUIRE(c.smem <= 227 * 1024, "wide_wgrad: row length %d needs %zu bytes of shared memory", gx.wp, c.smem);
  return 0;
}

size_t mil_wide_wgrad_partial_floats(const MilPF8& gx, const MilPF8& gz, int ks) {
  WwConfig c;
  if (ww_config(gx, gz, ks, &c) != 0) return 0;
  return (size_t)c.kinds * c.nsplit * c.p.rec_floats;
}

int mil_launch_wide_wgrad(const void* x, const MilPF8& gx, const void* dz, const MilPF8& gz, float* partial, float* dw,
                          float* db, int ks, cudaStream_t s) {
  MIL_REQUIRE(gx.n == gz.n && gx.h == gz.h && gx.w == gz.w && gx.wp == gz.wp && gx.hp == gz.hp, "wide_wgrad: geometry mismatch");
  MIL_REQUIRE(db == nullptr, "wide_wgrad: the wide parameterisation has no convolution bias (alt_resnet.py:24-32)");
  WwConfig c;
  MIL_TRY(ww_config(gx, gz, ks, &c));
  MIL_REQUIRE(gx.wp + 1 <= gx.G, "wide_wgrad: the window reaches %d pixels back but the map's guard is %lld", gx.wp + 1, gx.G);
  MIL_SET_SMEM(wide_wgrad_kernel, c.smem);
  wide_wgrad_kernel<<<dim3(c.nsplit, c.kinds), WW_THREADS, c.smem, s>>>((const __nv_bfloat16*)x, gx, (const __nv_bfloat16*)dz,
                                                                        gz, partial, c.p);
  MIL_LAUNCH_OK();
  WwReduceParams rp;
  rp.p = c.p;
  rp.nsplit = c.nsplit;
  rp.ks = ks;
  if (ks == 7) { rp.cout = gz.cb * 8 / 4; rp.cin = 3; }
  else { rp.cout = gz.cb * 8; rp.cin = gx.cb * 8; }
  if (ks == 7) {
    wide_stem_reduce_kernel<<<(unsigned)mil_cdiv(rp.cout * 147, 256), 256, 0, s>>>(partial, dw, rp);
  } else {
    const int per = c.p.tg_ntaps[0] * c.p.cit * 128;
    wide_wgrad_reduce_kernel<<<dim3((unsigned)mil_cdiv(per, 256), c.kinds), 256, 0, s>>>(partial, dw, rp);
  }
  MIL_LAUNCH_OK();
  return 0;
}
