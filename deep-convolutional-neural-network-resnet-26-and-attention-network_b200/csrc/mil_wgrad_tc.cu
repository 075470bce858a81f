// tcgen05 weight gradient of the 3x3 / 1x1 stride-1 convolutions on the PF8 layout (bf16 operands, fp32
// accumulation in TMEM), with the bias gradient as one extra accumulator column block.
// Reference semantics: autograd of nnBlocks.py:178-181 (gbm/classify_combined.py:447).
//
//   dW[t][ci][co] = sum over flat pixels q of  x[q + shift_t][ci] * dz[q][co]         db[co] = sum_q dz[q][co]
//
// GEMM view per tap t:  D_t[co][ci] += A[co][k] * B_t[ci][k],  k = flat pixel.  Both operands are used exactly
// as PF8 stores them -- [pixel][8 channels], i.e. "MN-major" core matrices of 8 pixels x 16 B -- so a 128-pixel
// K-tile needs one bulk-TMA copy per channel chunk and NO transposition: A = the dz planes, B_t = the x planes
// read from the start address (halo + shift_t) pixels into the span, like the forward kernel does.
// The nine D_t (and the bias block, B = a constant all-ones block) stay in TMEM for the whole kernel: every CTA
// accumulates its share of the pixel tiles (split-K over CTAs), then writes ONE partial record; the fixed-order
// reduction mil_launch_reduce_conv_w sums the records -> deterministic.  TMEM holds 512 columns, so layers
// with 9 * Cin_pad > 496 split the taps over two CTA groups (blockIdx.y).
#include <algorithm>

#include "mil_common.cuh"
#include "mil_conv_tc.cuh"
#include "mil_tc_ptx.cuh"

#define WG_TK 128
#define WG_STAGES 3
#define WG_THREADS 192  // warp 0 producer, warp 1 MMA issuer, warps 2..5 epilogue
#define WG_A_PLANE (WG_TK * 16)
#define WG_SLACK (40 * 1024)

struct WgSmemHeader {
  uint64_t full[WG_STAGES], empty[WG_STAGES], done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __nv_bfloat16* __restrict__ x, MilPF8 gx, const __nv_bfloat16* __restrict__ dz, MilPF8 gz,
                float* __restrict__ partial, long long rec_stride, MilTcShape sh, int halo, int taps_per_group,
                int npad, int mma_m) {
  extern __shared__ __align__(128) unsigned char smem[];
  WgSmemHeader* hd = reinterpret_cast<WgSmemHeader*>(smem);
  unsigned char* ones = smem + 128;            // 512 B of bf16 1.0: the B operand of the bias-gradient MMA
  unsigned char* stage0 = smem + 128 + 512;
  const int ntaps = sh.ntaps;
  const int span = WG_TK + 2 * halo;
  const uint32_t b_plane = (uint32_t)span * 16;
  const uint32_t a_bytes = (uint32_t)gz.cb * WG_A_PLANE;
  const uint32_t stage_bytes = a_bytes + (uint32_t)gx.cb * b_plane;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap_lo = blockIdx.y * taps_per_group;
  const int tap_hi = min(ntaps, tap_lo + taps_per_group);
  const int ntl = tap_hi - tap_lo;
  const bool with_bias = (blockIdx.y == gridDim.y - 1);
  const long long n_tiles = mil_cdiv(gz.Q, WG_TK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&hd->full[s], 1); mbar_init(&hd->empty[s], 1); }
    mbar_init(&hd->done, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x)
    reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;  // two bf16 1.0
  fence_proxy_async();
  if (warp == 1) tmem_alloc(&hd->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hd->tmem_base;

  if (warp == 0) {
    // producer: whole warp in uniform control flow, one elected lane issues the bulk copies
    int stage = 0;
    uint32_t phase = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&hd->empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&hd->full[stage], stage_bytes);
        const long long q0 = t * WG_TK;
        unsigned char* dst = stage0 + (size_t)stage * stage_bytes;
        for (int c = 0; c < gz.cb; ++c)
          bulk_g2s(dst + (size_t)c * WG_A_PLANE, dz + mil_pf8_off(gz, c, q0), WG_A_PLANE, &hd->full[stage]);
        for (int c = 0; c < gx.cb; ++c)
          bulk_g2s(dst + a_bytes + (size_t)c * b_plane, x + mil_pf8_off(gx, c, q0 - halo), b_plane, &hd->full[stage]);
      }
      __syncwarp();
      if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: uniform control flow (descriptors stay in uniform registers), one elected lane issues.
    // M = 64 when the output channels fit (the accumulator rows then sit in 16 lanes of every lane quarter).
    // D = f32, A = B = bf16, both MN-major (bits 15, 16)
    const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(mma_m >> 4) << 24);
    const uint32_t idesc = idesc_base | ((uint32_t)(npad >> 3) << 17);
    const uint32_t idesc_b = idesc_base | ((uint32_t)(16 >> 3) << 17);
    const uint64_t ones_desc = make_desc(smem_u32(ones), 128, 256);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&hd->full[stage], phase);
      tc_fence_after();
      const uint32_t a_base = smem_u32(stage0 + (size_t)stage * stage_bytes);
      const uint32_t b_base = a_base + a_bytes;
      // descriptors of K-step kk = descriptor of K-step 0 + kk * 256 B (start-address field, 16-byte units)
      const uint64_t ad0 = make_desc(a_base, 128, WG_A_PLANE);
      const uint32_t acc0 = first ? 0u : 1u;
      if (elect_one()) {
        for (int tl = 0; tl < ntl; ++tl) {
          const int tap = tap_lo + tl;
          const int s = sh.t_dy[tap] * gx.wp + sh.t_dx[tap];
          const uint64_t bd0 = make_desc(b_base + (uint32_t)(halo + s) * 16, 128, b_plane);
          const uint32_t d = tmem_base + tl * npad;
          umma_bf16(d, ad0, bd0, idesc, acc0);
#pragma unroll
          for (int kk = 1; kk < WG_TK / 16; ++kk) umma_bf16(d, ad0 + kk * 16, bd0 + kk * 16, idesc, 1u);
        }
        if (with_bias) {
          const uint32_t d = tmem_base + ntl * npad;
          umma_bf16(d, ad0, ones_desc, idesc_b, acc0);
#pragma unroll
          for (int kk = 1; kk < WG_TK / 16; ++kk) umma_bf16(d, ad0 + kk * 16, ones_desc, idesc_b, 1u);
        }
        umma_commit(&hd->empty[stage]);
      }
      __syncwarp();
      first = false;
      if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(&hd->done);
    __syncwarp();
  } else {
    // epilogue: TMEM lane = output channel co, columns = (local tap, ci)
    const int quarter = warp & 3;
    // accumulator row -> TMEM lane: M = 128: row i in lane i;  M = 64: row i in lane 32*(i/16) + i%16 (each
    // lane quarter holds 16 rows)
    const int co = mma_m == 128 ? quarter * 32 + lane : (lane < 16 ? quarter * 16 + lane : 1 << 20);
    const int coutp = gz.cb * 8, cinp = gx.cb * 8;
    mbar_wait(&hd->done, 0);
    tc_fence_after();
    if ((mma_m == 128 ? quarter * 32 : quarter * 16) < coutp) {  // warp-uniform
      float* rec = partial + (size_t)blockIdx.x * rec_stride;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      for (int tl = 0; tl < ntl; ++tl) {
        for (int c = 0; c < gx.cb; ++c) {
          float v[8];
          tmem_ld8(taddr + tl * npad + c * 8, v);
          tmem_ld_wait();
          if (co < coutp) {
#pragma unroll
            for (int j = 0; j < 8; ++j) rec[((size_t)(tap_lo + tl) * cinp + c * 8 + j) * coutp + co] = v[j];
          }
        }
      }
      if (with_bias) {
        float v[8];
        tmem_ld8(taddr + ntl * npad, v);
        tmem_ld_wait();
        if (co < coutp) rec[(size_t)ntaps * cinp * coutp + co] = v[0];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
static int wg_sm_count() {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      n_sm = 148;
  }
  return n_sm;
}

static void wg_config(const MilPF8& gx, const MilPF8& gz, int ks, int* npad, int* groups, int* tpg, int* ctas) {
  *npad = (gx.c + 15) / 16 * 16;
  const int ntaps = ks == 7 ? 16 : ks * ks;
  *groups = (ntaps * *npad + 16 <= 512) ? 1 : 2;
  *tpg = (ntaps + *groups - 1) / *groups;
  const long long n_tiles = mil_cdiv(gz.Q, WG_TK);
  *ctas = (int)std::max<long long>(1, std::min<long long>(n_tiles, wg_sm_count() / *groups));
}

bool mil_wgrad_tc_supported(int dtype, int ks, int stride, int cin, int cout) {
  return dtype == MIL_BF16 && (ks == 3 || ks == 1) && stride == 1 && cin <= 80 && cout <= 80;
}

size_t mil_wgrad_tc_partial_floats(const MilPF8& gx, const MilPF8& gz, int ks) {
  int npad, groups, tpg, ctas;
  wg_config(gx, gz, ks, &npad, &groups, &tpg, &ctas);
  const size_t ntaps = ks == 7 ? 16 : (size_t)ks * ks;
  return (size_t)ctas * (ntaps * gx.cb * 8 * gz.cb * 8 + gz.cb * 8);
}

// accumulate-only part: writes `*ctas_out` partial records [tap][cin_pad][cout_pad] (+[cout_pad] bias sums)
int mil_launch_wgrad_tc_partials(const void* x, const MilPF8& gx, const void* dz, const MilPF8& gz, float* partial,
                                 int ks, int* ctas_out, long long* rec_out, cudaStream_t s) {
  MIL_REQUIRE(gx.n == gz.n && gx.h == gz.h && gx.w == gz.w && gx.wp == gz.wp && gx.hp == gz.hp,
              "wgrad_tc: geometry mismatch");
  int npad, groups, tpg, ctas;
  wg_config(gx, gz, ks, &npad, &groups, &tpg, &ctas);
  MilTcShape sh;
  MIL_TRY(mil_tc_shape(gx.c, gz.c, ks, &sh));
  const int halo = mil_tc_halo(sh, gx.wp);
  MIL_REQUIRE(halo <= gx.G, "wgrad_tc: the window reaches %d pixels back but the map's guard is %lld", halo, gx.G);
  const size_t stage = (size_t)gz.cb * WG_A_PLANE + (size_t)gx.cb * (WG_TK + 2 * halo) * 16;
  const size_t smem = 128 + 512 + WG_STAGES * stage + WG_SLACK;
  MIL_REQUIRE(smem <= 227 * 1024, "wgrad_tc: tile width %d needs %zu bytes of shared memory", gx.w, smem);
  MIL_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long rec = (long long)sh.ntaps * gx.cb * 8 * gz.cb * 8 + gz.cb * 8;
  wgrad_tc_kernel<<<dim3(ctas, groups), WG_THREADS, smem, s>>>((const __nv_bfloat16*)x, gx, (const __nv_bfloat16*)dz,
                                                             gz, partial, rec, sh, halo, tpg, npad,
                                                             gz.cb * 8 <= 64 ? 64 : 128);
  MIL_LAUNCH_OK();
  *ctas_out = ctas;
  *rec_out = rec;
  return 0;
}

int mil_launch_wgrad_tc(const void* x, const MilPF8& gx, const void* dz, const MilPF8& gz, float* partial, float* dw,
                        float* db, int ks, cudaStream_t s) {
  int ctas;
  long long rec;
  MIL_TRY(mil_launch_wgrad_tc_partials(x, gx, dz, gz, partial, ks, &ctas, &rec, s));
  return mil_launch_reduce_conv_w(partial, ctas, rec, dw, db, gz.c, gx.c, ks, s);
}
