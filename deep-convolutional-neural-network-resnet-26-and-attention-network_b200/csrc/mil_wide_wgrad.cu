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

struct WwGroup {
  int ntaps;           // B-side taps = MMAs per K-step
  int plane0;          // first x chunk plane of the group (a parity phase of the split input; 0 otherwise)
  int bshift[9];       // pixel shift of each tap's x window
  int nslots;          // A-side slots = shifted copies of the dz tile on the M axis (2 for 64-channel layers: two tap rows)
  int ashift[2];       // pixel shift of each slot's dz window
  short tapid[2][9];   // weight tap ky * ks + kx of (slot, tap)
};
struct WwParams {
  int n_cot, n_cit, n_tg;  // kinds = output-channel tiles x input-channel tiles x tap groups (blockIdx.y)
  int cit;                 // input channels per tile (N of the MMAs, a multiple of 16)
  int a_chunks;            // dz chunk planes per A slot: 16 (128 output channels), 8 when two slots share the M axis
  int n_stages;
  long long rec_floats;    // floats per partial record
  WwGroup g[4];
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
  const WwGroup& gr = pr.g[tg];
  const int ntaps = gr.ntaps;
  int smin = gr.bshift[0], smax = gr.bshift[0];
  for (int t = 1; t < ntaps; ++t) { smin = min(smin, gr.bshift[t]); smax = max(smax, gr.bshift[t]); }
  const int span = WW_TK + (smax - smin);
  const int ncic = pr.cit >> 3;
  // A side: nslots shifted copies of the output-channel tile's planes (fewer planes in the last tile of a layer whose
  // width is not a multiple of 128; the MMA still reads 128 rows: whatever the missing planes' slots hold lands in
  // accumulator rows nobody reads)
  const int ncoc = min(pr.a_chunks, gz.cb - cot_i * pr.a_chunks);
  const uint32_t a_bytes = 16 * WW_A_PLANE, b_plane = (uint32_t)span * 16;
  const uint32_t stage_bytes = a_bytes + (uint32_t)ncic * b_plane;
  const uint32_t load_bytes = (uint32_t)(gr.nslots * ncoc) * WW_A_PLANE + (uint32_t)ncic * b_plane;
  const long long n_tiles = mil_cdiv(gz.Q, WW_TK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < pr.n_stages; ++s) { mbar_init(&hd->full[s], 1); mbar_init(&hd->empty[s], 1); }
    mbar_init(&hd->done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&hd->tmem_base, 512);
  mil_pdl_wait();   // everything above overlaps the previous kernel's tail; global memory only from here on (PDL, mil_common.cuh)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hd->tmem_base;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t s0 = smem_u32(stage0), full0 = smem_u32(&hd->full[0]);
    const char* srca = reinterpret_cast<const char*>(dz) + ((long long)cot_i * pr.a_chunks * gz.PS + gz.G + (long long)blockIdx.x * WW_TK) * 16;
    const char* srcb = reinterpret_cast<const char*>(x) +
                       ((long long)(gr.plane0 + cit_i * ncic) * gx.PS + gx.G + smin + (long long)blockIdx.x * WW_TK) * 16;
    const long long stra = gz.PS * 16, strb = gx.PS * 16, tstride = (long long)gridDim.x * WW_TK * 16;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&hd->empty[stage], phase ^ 1);
      if (elect_one()) {
        const uint32_t bar = full0 + (uint32_t)stage * 8;
        mbar_expect_tx_u32(bar, load_bytes);
        const uint32_t sb = s0 + (uint32_t)stage * stage_bytes;
        for (int sl = 0; sl < gr.nslots; ++sl) {
          uint32_t dst = sb + (uint32_t)(sl * pr.a_chunks) * WW_A_PLANE;
          const char* sp = srca + (long long)gr.ashift[sl] * 16;
          for (int c = 0; c < ncoc; ++c, dst += WW_A_PLANE, sp += stra) bulk_g2s_u32(dst, sp, WW_A_PLANE, bar);
        }
        uint32_t dst = sb + a_bytes;
        const char* sp = srcb;
        for (int c = 0; c < ncic; ++c, dst += b_plane, sp += strb) bulk_g2s_u32(dst, sp, b_plane, bar);
      }
      __syncwarp();
      srca += tstride; srcb += tstride;
      if (++stage == pr.n_stages) { stage = 0; phase ^= 1; }
    }
    mil_pdl_trigger();  // all of this CTA's loads are issued: the next kernel may start its prologue (PDL, mil_common.cuh)
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
          const uint64_t bd0 = make_desc(b_base + (uint32_t)(gr.bshift[tl] - smin) * 16, 128, b_plane);
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
    // epilogue: TMEM lane = accumulator row (slot, output channel of the tile), columns = (local tap, input channel)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
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
        for (int j = 0; j < 8; ++j) rec[((size_t)tl * pr.cit + c * 8 + j) * 128 + row] = any ? v[j] : 0.f;
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
// Block = (32 output channels, 8 input channels: one per warp) of one (output tile, input tile): it walks over ALL tap groups / slots / taps
// of that pair -- a warp reads 32 consecutive accumulator rows (= output channels) of one (tap, input channel) from every
// split-K record (coalesced, all loads of the fixed-order sum in flight) -- parks the sums in shared memory as
// [output channel][input channel][tap] and adds each output channel's 8 * ks * ks contiguous floats to dw in one coalesced run
// (the first version wrote one float per thread at a stride of cin * ks * ks: 123 us for 38 MB on the 512-channel layers,
// profiles/r2_ncu_wide_small_kernels.txt).
#define WWR_T 32   // output channels per block
#define WWR_K 8    // input channels per block = warps
__global__ void __launch_bounds__(256)
wide_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, const __grid_constant__ WwReduceParams rp) {
  __shared__ float tile[WWR_T * (WWR_K * 9 + 1)];
  const WwParams& pr = rp.p;
  const int kk = rp.ks * rp.ks, pitch = WWR_K * kk + 1;
  const int rows_per_slot = pr.a_chunks * 8;
  const int nct = rows_per_slot / WWR_T;                 // output-channel tiles per slot
  const int j = blockIdx.x % nct, c0 = (blockIdx.x / nct) * WWR_K;
  const int cit_i = blockIdx.y % pr.n_cit, cot_i = blockIdx.y / pr.n_cit;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < WWR_T * pitch; i += blockDim.x) tile[i] = 0.f;
  __syncthreads();
  for (int tg = 0; tg < pr.n_tg; ++tg) {
    const WwGroup& gr = pr.g[tg];
    const int kind = (cot_i * pr.n_cit + cit_i) * pr.n_tg + tg;
    const float* base = partial + (size_t)kind * rp.nsplit * pr.rec_floats;
    for (int slot = 0; slot < gr.nslots; ++slot) {
      const int row = slot * rows_per_slot + j * WWR_T + lane;
      for (int tl = 0; tl < gr.ntaps; ++tl) {
        const int tap = gr.tapid[slot][tl];
        if (tap < 0) continue;
        const int cl = c0 + warp;
        if (cl < pr.cit) {
          const float* src = base + (((size_t)tl * pr.cit + cl) << 7) + row;
          float acc = 0.f;
          int k = 0;
          for (; k + 4 <= rp.nsplit; k += 4) {
            const float v0 = src[(size_t)k * pr.rec_floats], v1 = src[(size_t)(k + 1) * pr.rec_floats],
                        v2 = src[(size_t)(k + 2) * pr.rec_floats], v3 = src[(size_t)(k + 3) * pr.rec_floats];
            acc += v0; acc += v1; acc += v2; acc += v3;
          }
          for (; k < rp.nsplit; ++k) acc += src[(size_t)k * pr.rec_floats];
          tile[lane * pitch + (cl - c0) * kk + tap] += acc;
        }
      }
    }
  }
  __syncthreads();
  const int kc0 = cit_i * pr.cit + c0;
  const int ncols = min(WWR_K, min(pr.cit - c0, rp.cin - kc0)) * kk;  // this block's contiguous floats per output channel
  for (int r = warp; r < WWR_T; r += 8) {
    const int nc = cot_i * rows_per_slot + j * WWR_T + r;
    if (nc >= rp.cout) break;
    float* dst = dw + ((size_t)nc * rp.cin + kc0) * kk;
    for (int i = lane; i < ncols; i += 32) dst[i] += tile[r * pitch + i];
  }
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

// s2 = 1: x is the PHASE-SPLIT input of a 3x3 / stride-2 convolution (mil_launch_split2: 4 * cb planes at the output
// resolution) and dz its output gradient -- the nine taps grouped by the parity phase they read (4 + 2 + 2 + 1)
static int ww_config(const MilPF8& gx, const MilPF8& gz, int ks, int s2, WwConfig* out) {
  WwConfig& c = *out;
  c = WwConfig{};
  const int cin = s2 ? gx.cb * 2 : gx.cb * 8, cout = gz.cb * 8;
  MIL_REQUIRE(ks == 1 || ks == 3 || ks == 7, "wide_wgrad: unsupported window %d", ks);
  MIL_REQUIRE(!s2 || ks == 3, "wide_wgrad: the phase-split form is the 3x3 / stride-2 convolution");
  MIL_REQUIRE(cout % 64 == 0, "wide_wgrad: %d output channels (need a multiple of 64)", cout);
  const int wp = gx.wp;
  for (int g = 0; g < 4; ++g)
    for (int sl = 0; sl < 2; ++sl)
      for (int t = 0; t < 9; ++t) c.p.g[g].tapid[sl][t] = -1;
  c.p.a_chunks = 16;
  if (ks == 7) {
    MIL_REQUIRE(cin == 48, "wide_wgrad: the stem form reads the 48-channel space-to-depth input");
    c.p.cit = 48; c.p.n_cit = 1; c.p.n_tg = 1;
    WwGroup& g = c.p.g[0];
    g.ntaps = 9; g.plane0 = 0; g.nslots = 1; g.ashift[0] = 0;
    for (int t = 0; t < 9; ++t) { g.bshift[t] = (t / 3 - 1) * wp + (t % 3 - 1); g.tapid[0][t] = (short)t; }
  } else {
    c.p.cit = cin >= 128 ? 128 : cin;
    MIL_REQUIRE(cin % c.p.cit == 0 && c.p.cit % 16 == 0, "wide_wgrad: %d input channels", cin);
    c.p.n_cit = cin / c.p.cit;
    if (s2) {
      // phase (a, b): tap (dy, dx) in {-1, 0}^2 carries w[2 dy + a + 1][2 dx + b + 1]; dy = -1 needs a = 1
      c.p.n_tg = 4;
      for (int ph = 0; ph < 4; ++ph) {
        const int a = ph >> 1, b = ph & 1;
        WwGroup& g = c.p.g[ph];
        g.plane0 = ph * (cin / 8); g.nslots = 1; g.ashift[0] = 0; g.ntaps = 0;
        for (int dy = -1; dy <= 0; ++dy)
          for (int dx = -1; dx <= 0; ++dx) {
            if ((dy == -1 && a == 0) || (dx == -1 && b == 0)) continue;
            g.bshift[g.ntaps] = dy * wp + dx;
            g.tapid[0][g.ntaps] = (short)((2 * dy + a + 1) * 3 + (2 * dx + b + 1));
            ++g.ntaps;
          }
      }
    } else if (ks == 3 && cout == 64) {
      // 64 output channels fill half of the M = 128 rows: two tap ROWS share the M axis (slot = the dz tile shifted by
      // -dy * wp), the three dx taps are the MMAs -- dW[dy][dx] = sum_q' x[q' + dx] dz[q' - dy wp]
      c.p.a_chunks = 8;
      c.p.n_tg = 2;
      for (int gi = 0; gi < 2; ++gi) {
        WwGroup& g = c.p.g[gi];
        g.ntaps = 3; g.plane0 = 0;
        g.nslots = gi == 0 ? 2 : 1;
        const int dys[2] = {gi == 0 ? -1 : 1, 0};
        for (int sl = 0; sl < g.nslots; ++sl) {
          g.ashift[sl] = -dys[sl] * wp;
          for (int t = 0; t < 3; ++t) g.tapid[sl][t] = (short)((dys[sl] + 1) * 3 + t);
        }
        for (int t = 0; t < 3; ++t) g.bshift[t] = t - 1;
      }
    } else {
      c.p.n_tg = ks == 3 ? 3 : 1;
      for (int gi = 0; gi < c.p.n_tg; ++gi) {
        WwGroup& g = c.p.g[gi];
        g.ntaps = ks == 3 ? 3 : 1; g.plane0 = 0; g.nslots = 1; g.ashift[0] = 0;
        for (int t = 0; t < g.ntaps; ++t) {
          g.bshift[t] = ks == 3 ? (gi - 1) * wp + (t - 1) : 0;
          g.tapid[0][t] = (short)(ks == 3 ? gi * 3 + t : 0);
        }
      }
    }
  }
  c.p.n_cot = (cout + c.p.a_chunks * 8 - 1) / (c.p.a_chunks * 8);
  if (c.p.a_chunks == 8) c.p.n_cot = 1;
  c.kinds = c.p.n_cot * c.p.n_cit * c.p.n_tg;
  int max_taps = 0, span_max = 0;
  for (int g = 0; g < c.p.n_tg; ++g) {
    const WwGroup& gr = c.p.g[g];
    max_taps = std::max(max_taps, gr.ntaps);
    int lo = gr.bshift[0], hi = lo;
    for (int t = 1; t < gr.ntaps; ++t) { lo = std::min(lo, gr.bshift[t]); hi = std::max(hi, gr.bshift[t]); }
    span_max = std::max(span_max, WW_TK + hi - lo);
  }
  MIL_REQUIRE(max_taps * c.p.cit <= 512, "wide_wgrad: %d taps x %d input channels exceed the 512 TMEM columns", max_taps, c.p.cit);
  c.p.rec_floats = (long long)max_taps * c.p.cit * 128;
  const long long n_tiles = mil_cdiv(gz.Q, WW_TK);
  c.nsplit = (int)std::max<long long>(1, std::min<long long>(n_tiles, ww_sm_count() / c.kinds));
  const size_t stage = (size_t)16 * WW_A_PLANE + (size_t)(c.p.cit / 8) * span_max * 16;
  c.p.n_stages = WW_MAX_STAGES;
  while (c.p.n_stages > 1 && 128 + c.p.n_stages * stage > 224 * 1024) --c.p.n_stages;
  c.smem = 128 + c.p.n_stages * stage;
  MIL_REQUIRE(c.smem <= 227 * 1024, "wide_wgrad: row length %d needs %zu bytes of shared memory", gx.wp, c.smem);
  return 0;
}

size_t mil_wide_wgrad_partial_floats(const MilPF8& gx, const MilPF8& gz, int ks, int s2) {
  WwConfig c;
  if (ww_config(gx, gz, ks, s2, &c) != 0) return 0;
  return (size_t)c.kinds * c.nsplit * c.p.rec_floats;
}

int mil_launch_wide_wgrad(const void* x, const MilPF8& gx, const void* dz, const MilPF8& gz, float* partial, float* dw,
                          float* db, int ks, cudaStream_t s, int s2) {
  MIL_REQUIRE(gx.n == gz.n && gx.h == gz.h && gx.w == gz.w && gx.wp == gz.wp && gx.hp == gz.hp, "wide_wgrad: geometry mismatch");
  MIL_REQUIRE(db == nullptr, "wide_wgrad: the wide parameterisation has no convolution bias (alt_resnet.py:24-32)");
  WwConfig c;
  MIL_TRY(ww_config(gx, gz, ks, s2, &c));
  MIL_REQUIRE(gx.wp + 1 <= gx.G, "wide_wgrad: the window reaches %d pixels back but the map's guard is %lld", gx.wp + 1, gx.G);
  MIL_SET_SMEM(wide_wgrad_kernel, c.smem);
  MIL_LAUNCH_PDL(wide_wgrad_kernel, dim3(c.nsplit, c.kinds), WW_THREADS, c.smem, s, (const __nv_bfloat16*)x, gx, (const __nv_bfloat16*)dz,
                                                                        gz, partial, c.p);
  WwReduceParams rp;
  rp.p = c.p;
  rp.nsplit = c.nsplit;
  rp.ks = ks;
  if (ks == 7) { rp.cout = gz.cb * 8 / 4; rp.cin = 3; }
  else { rp.cout = gz.cb * 8; rp.cin = s2 ? gx.cb * 2 : gx.cb * 8; }
  if (ks == 7) {
    wide_stem_reduce_kernel<<<(unsigned)mil_cdiv(rp.cout * 147, 256), 256, 0, s>>>(partial, dw, rp);
  } else {
    MIL_REQUIRE(c.p.a_chunks * 8 % WWR_T == 0 && ks * ks <= 9, "wide_wgrad: reduction tile does not divide the slot (%d rows)", c.p.a_chunks * 8);
    const int nct = c.p.a_chunks * 8 / WWR_T;
    wide_wgrad_reduce_kernel<<<dim3((unsigned)(nct * mil_cdiv(c.p.cit, WWR_K)), c.p.n_cot * c.p.n_cit), 256, 0, s>>>(partial, dw, rp);
  }
  MIL_LAUNCH_OK();
  return 0;
}
