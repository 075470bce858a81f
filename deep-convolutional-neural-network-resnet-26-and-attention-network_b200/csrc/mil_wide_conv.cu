// tcgen05 implicit-GEMM convolution for the WIDE extractor parameterisation (64 .. 512 channels, alt_resnet.py:24-67):
// forward and data gradient of conv3x3 / conv1x1 (no bias in alt_resnet; an optional bias is supported), the stride-2
// 3x3 forward on the phase-split input, and the 7x7 / stride-2 stem in space-to-depth form, with the fused epilogues
// of nnBlocks.py:175-189 / alt_resnet.py:52-67 (residual add, ReLU / LeakyReLU, activation-derivative mask).
//
// GEMM view of one unit = (super-tile of TM x 128 flat pixels, tile of NT output channels):
//   D_m[128][NT] += A_{m,step,tap}[128][16] * B_{step,tap}[16][NT]        m < TM, step = (group, chunk pair), tap
// A is a start-address shift into the bulk-copied span of two chunk planes (PF8: a tap is a constant pixel shift),
// B is the weight block of (N-tile, step, tap).  A ring stage holds ONE K-step: the two input planes of the whole
// super-tile (+ halo) and the weight slab of all its taps; TM M-tiles reuse every slab, so a CTA pulls
// (2 * (TM*128 + 2 halo) * 16 + taps * NT * 32) bytes per TM * taps MMAs -- 26 B/clk at TM = 4, NT = 128, nine taps --
// while each MMA (M = 128, N = 128, K = 16) runs at the tensor pipe's full rate (profiles/r1_mma_cost.txt:
// cost = max((128 + N) / 4, N / 2) = 64 cycles).  Accumulators: TM * NT TMEM columns per set.
//
// Warp roles (one persistent CTA per SM, units strided over CTAs; consecutive units share the super-tile => its planes
// hit L2):  warp 0 producer (bulk-TMA), warp 1 MMA issuer (uniform control flow, one elected lane), warps 2..9
// epilogue: two warps per TMEM lane quarter, each taking half of the N-tile's chunks.
#include <algorithm>

#include "mil_common.cuh"
#include "mil_tc_ptx.cuh"
#include "mil_wide.cuh"

#define WIDE_MAX_STAGES 6
#define WIDE_THREADS 320
#define WIDE_MAX_SETS 4

struct WideSmemHeader {
  uint64_t full[WIDE_MAX_STAGES], empty[WIDE_MAX_STAGES], acc_full[WIDE_MAX_SETS], acc_empty[WIDE_MAX_SETS];
  uint32_t tmem_base;
  alignas(16) float bias[512];
};

struct WideKParams {
  int ngroups, npairs;
  int gplane0[WIDE_MAX_GROUPS];  // first chunk plane of the group inside x
  int gntaps[WIDE_MAX_GROUPS];
  int gwoff[WIDE_MAX_GROUPS];    // byte offset of the group's first slab inside an N-tile's weights
  int shift[WIDE_MAX_GROUPS][WIDE_MAX_TAPS];  // pixel shift of each tap (sign already chosen: forward +, data gradient -)
  int halo, tm, nt, n_ntiles, nsets, n_stages;
  uint32_t a_plane, b_tap, stage_bytes, wtile_bytes;
  int epi, has_bias;
  float slope;
  // per_tile = 1 (single group): N-tile nti issues only the taps tile_tap[nti][0 .. tile_ntaps[nti]) -- the weight blocks
  // of the others are all zero for that tile (the phase-wise stride-2 data gradient)
  int per_tile;
  int tile_ntaps[8];
  int tile_shift[8][4];   // pixel shift of the tile's t-th tap
  int tile_blk[8][4];     // which tap block of the slab it multiplies with
  int res_chunks;         // the residual exists for the first res_chunks output chunks only (all: go.cb)
};

__device__ __forceinline__ uint4 wide_ld16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void wide_unpack8(const uint4& r, float v[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

__global__ void __launch_bounds__(WIDE_THREADS, 1)
wide_conv_kernel(const __nv_bfloat16* __restrict__ x, MilPF8 gx, const __nv_bfloat16* __restrict__ wpk,
                 const float* __restrict__ bias, const __nv_bfloat16* res, const __nv_bfloat16* __restrict__ act,
                 __nv_bfloat16* out, MilPF8 go, const __grid_constant__ WideKParams kp) {
  extern __shared__ __align__(128) unsigned char smem[];
  WideSmemHeader* hd = reinterpret_cast<WideSmemHeader*>(smem);
  const uint32_t hdr_bytes = (uint32_t)((sizeof(WideSmemHeader) + 127) / 128 * 128);
  unsigned char* stage0 = smem + hdr_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = mil_cdiv(gx.Q, 128);
  const long long n_super = mil_cdiv(n_tiles, kp.tm);
  const long long n_units = n_super * kp.n_ntiles;
  const uint32_t set_cols = (uint32_t)(kp.tm * kp.nt);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kp.n_stages; ++s) { mbar_init(&hd->full[s], 1); mbar_init(&hd->empty[s], 1); }
    for (int a = 0; a < kp.nsets; ++a) { mbar_init(&hd->acc_full[a], 1); mbar_init(&hd->acc_empty[a], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&hd->tmem_base, 512);
  mil_pdl_wait();   // everything above overlaps the previous kernel's tail; global memory only from here on (PDL, mil_common.cuh)
  for (int i = threadIdx.x; i < 512; i += blockDim.x) hd->bias[i] = (kp.has_bias && i < go.c) ? bias[i] : 0.f;  // (<= 512 channels with a bias)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hd->tmem_base;

  if (warp == 0) {
    // ===================== producer =====================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t s0 = smem_u32(stage0), full0 = smem_u32(&hd->full[0]);
    const char* xb = reinterpret_cast<const char*>(x);
    const char* wb = reinterpret_cast<const char*>(wpk);
    const long long plane_bytes = gx.PS * 16;
    for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int nti = (int)(u % kp.n_ntiles);
      const long long sup = u / kp.n_ntiles;
      const long long start_px = gx.G + sup * kp.tm * 128 - kp.halo;  // >= 0: halo <= G
      // the last super-tile may reach past the plane's tail guard: copy what exists (rows past Q are never stored)
      const long long avail = (gx.PS - start_px) * 16;
      const uint32_t abytes = (uint32_t)(avail < (long long)kp.a_plane ? avail : (long long)kp.a_plane);
      const char* wt = wb + (size_t)nti * kp.wtile_bytes;
      for (int g = 0; g < kp.ngroups; ++g) {
        const uint32_t slab = (uint32_t)kp.gntaps[g] * kp.b_tap;
        const char* src = xb + ((long long)kp.gplane0[g] * gx.PS + start_px) * 16;
        const char* ws = wt + kp.gwoff[g];
        for (int p = 0; p < kp.npairs; ++p) {
          mbar_wait(&hd->empty[stage], phase ^ 1);
          if (elect_one()) {
            const uint32_t bar = full0 + (uint32_t)stage * 8;
            const uint32_t dst = s0 + (uint32_t)stage * kp.stage_bytes;
            mbar_expect_tx_u32(bar, 2 * abytes + slab);
            bulk_g2s_u32(dst, src, abytes, bar);
            bulk_g2s_u32(dst + kp.a_plane, src + plane_bytes, abytes, bar);
            bulk_g2s_u32(dst + 2 * kp.a_plane, ws, slab, bar);
          }
          __syncwarp();
          src += 2 * plane_bytes;
          ws += slab;
          if (++stage == kp.n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    mil_pdl_trigger();  // all of this CTA's loads are issued: the next kernel may start its prologue (PDL, mil_common.cuh)
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D = f32, A = B = bf16, both K-major, N = nt, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kp.nt >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a_hi = make_desc_bits(0, kp.a_plane, 128);          // K halves = the two chunk planes
    const uint64_t b_hi = make_desc_bits(0, (uint32_t)kp.nt * 16, 128);  // K halves = the two [nt][8] blocks of a tap
    int stage = 0, set = 0;
    uint32_t phase = 0, set_par = 0;
    const uint32_t s0 = smem_u32(stage0);
    for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
      const long long sup = u / kp.n_ntiles;
      const long long left = n_tiles - sup * kp.tm;
      const int tm_eff = (int)(left < kp.tm ? left : kp.tm);
      mbar_wait(&hd->acc_empty[set], set_par ^ 1);
      tc_fence_after();
      const int nti_m = (int)(u % kp.n_ntiles) & 7;
      const int tile_nt = kp.tile_ntaps[nti_m];
      const uint32_t d0 = tmem_base + (uint32_t)set * set_cols;
      bool first = true;
      for (int g = 0; g < kp.ngroups; ++g) {
        const int ntaps = kp.gntaps[g];
        for (int p = 0; p < kp.npairs; ++p) {
          mbar_wait(&hd->full[stage], phase);
          tc_fence_after();
          const uint32_t sb = s0 + (uint32_t)stage * kp.stage_bytes;
          if (elect_one()) {
            const uint32_t bb = (sb + 2 * kp.a_plane) >> 4;
            for (int m = 0; m < tm_eff; ++m) {
              const uint32_t ab = (sb + (uint32_t)(m * 128 + kp.halo) * 16) >> 4;
              const uint32_t d = d0 + (uint32_t)(m * kp.nt);
              if (!kp.per_tile) {
                for (int t = 0; t < ntaps; ++t)
                  umma_bf16(d, a_hi + (uint64_t)(ab + kp.shift[g][t]), b_hi + (uint64_t)(bb + (uint32_t)t * (kp.b_tap >> 4)), idesc,
                            (first && t == 0) ? 0u : 1u);
              } else {
                for (int t = 0; t < tile_nt; ++t)
                  umma_bf16(d, a_hi + (uint64_t)(ab + kp.tile_shift[nti_m][t]),
                            b_hi + (uint64_t)(bb + (uint32_t)kp.tile_blk[nti_m][t] * (kp.b_tap >> 4)), idesc, (first && t == 0) ? 0u : 1u);
              }
            }
            umma_commit(&hd->empty[stage]);
          }
          __syncwarp();
          first = false;
          if (++stage == kp.n_stages) { stage = 0; phase ^= 1; }
        }
      }
      if (elect_one()) umma_commit(&hd->acc_full[set]);
      __syncwarp();
      if (++set == kp.nsets) { set = 0; set_par ^= 1; }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // which half of the N-tile's chunks
    const int nc = kp.nt >> 3;
    const int c_lo = half * (nc >> 1), c_hi = c_lo + (nc >> 1);
    const int P = (int)go.P;
    const float slope = kp.slope;
    int set = 0;
    uint32_t set_par = 0;
    for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int nti = (int)(u % kp.n_ntiles);
      const long long sup = u / kp.n_ntiles;
      const long long left = n_tiles - sup * kp.tm;
      const int tm_eff = (int)(left < kp.tm ? left : kp.tm);
      mbar_wait(&hd->acc_full[set], set_par);
      tc_fence_after();
      for (int m = 0; m < tm_eff; ++m) {
        const long long q = (sup * kp.tm + m) * 128 + quarter * 32 + lane;
        const bool in_range = q < go.Q;
        bool live = false;
        if (in_range) {
          const int r = (int)(q % P);
          const int y = r / go.wp, xo = r - y * go.wp;
          live = y < go.h && xo < go.w;
        }
        const uint32_t taddr = tmem_base + (uint32_t)set * set_cols + (uint32_t)(m * kp.nt) + ((uint32_t)(quarter * 32) << 16);
        const long long e0 = ((long long)(nti * nc) * go.PS + go.G + q) * 8;  // element offset of chunk 0 of this N-tile
        for (int cb = c_lo; cb < c_hi; cb += 4) {
          float acc[4][8];
#pragma unroll
          for (int k = 0; k < 4; ++k) tmem_ld8(taddr + (uint32_t)(cb + k) * 8, acc[k]);
          uint4 rr[4], aa[4];
          if (live) {
            if (res != nullptr) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                rr[k] = (nti * nc + cb + k < kp.res_chunks) ? wide_ld16(res + e0 + (long long)(cb + k) * go.PS * 8) : make_uint4(0, 0, 0, 0);
            }
            if (kp.epi == MIL_EPI_DGRAD) {
#pragma unroll
              for (int k = 0; k < 4; ++k) aa[k] = wide_ld16(act + e0 + (long long)(cb + k) * go.PS * 8);
            }
          }
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float* v = acc[k];
              if (res != nullptr) {
                float rv[8];
                wide_unpack8(rr[k], rv);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += rv[j];
              }
              if (kp.has_bias) {
                const float* bp = &hd->bias[(nti * nc + cb + k) * 8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += bp[j];
              }
              if (kp.epi == MIL_EPI_FWD) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : slope * v[j];
              } else if (kp.epi == MIL_EPI_DGRAD) {
                float av[8];
                wide_unpack8(aa[k], av);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (!(av[j] > 0.f)) v[j] *= slope;
              }
              mil_store8(out + e0 + (long long)(cb + k) * go.PS * 8, v);
            }
          } else if (in_range) {  // pad pixel: keep the zero row / column zero (PF8 invariant)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<uint4*>(out + e0 + (long long)(cb + k) * go.PS * 8) = make_uint4(0, 0, 0, 0);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hd->acc_empty[set]);
      if (++set == kp.nsets) { set = 0; set_par ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- weight pre-pack ------------------------------------------------------------------------------------------
struct WidePackParams {
  MilWideShape sh;
  int gblk0[WIDE_MAX_GROUPS];  // first block of each group inside an N-tile
};
__global__ void wide_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wpk,
                                 const __grid_constant__ WidePackParams pp) {
  const MilWideShape& sh = pp.sh;
  const int nt = sh.nt;
  const long long per_block = (long long)nt * 16;  // bf16 elements of one (tap, chunk pair) block
  const long long total = (long long)sh.n_ntiles * sh.blocks_per_ntile * per_block;
  const int taps2 = sh.ks * sh.ks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 7);
    const int n = (int)((i >> 3) % nt);
    const int half = (int)((i / (8 * nt)) & 1);
    const long long blk = i / per_block;
    const int nti = (int)(blk / sh.blocks_per_ntile);
    int b = (int)(blk - (long long)nti * sh.blocks_per_ntile);
    int g = 0;
    while (g + 1 < sh.ngroups && b >= pp.gblk0[g + 1]) ++g;
    b -= pp.gblk0[g];
    const int p = b / sh.gntaps[g], t = b - p * sh.gntaps[g];
    const int kc = (2 * p + half) * 8 + e;  // kernel-side input channel inside the group
    const int nc = nti * nt + n;            // kernel-side output channel
    const int dy = sh.tdy[g][t], dx = sh.tdx[g][t];
    float v = 0.f;
    if (sh.mode == 0) {
      const int r = sh.ks / 2;
      const int tap = (dy + r) * sh.ks + (dx + r);
      const int co = sh.transposed ? kc : nc, ci = sh.transposed ? nc : kc;
      if (co < sh.wcout && ci < sh.wcin) v = w[((size_t)co * sh.wcin + ci) * taps2 + tap];
    } else if (sh.mode == 1) {
      const int ky = 2 * dy + sh.ga[g] + 1, kx = 2 * dx + sh.gb[g] + 1;
      if (nc < sh.wcout && kc < sh.wcin) v = w[((size_t)nc * sh.wcin + kc) * 9 + ky * 3 + kx];
    } else if (sh.mode == 7) {
      // stride-2 data gradient, all four input phases as output channels (phase, ci): tap (dY, dX) in {0, 1}^2 carries
      // w[a + 1 - 2 dY][b + 1 - 2 dX]^T where dY <= a and dX <= b, zero otherwise
      const int ph = nc / sh.wcin, ci = nc - ph * sh.wcin;
      const int a = ph >> 1, bq = ph & 1;
      if (ph < 4 && dy <= a && dx <= bq && kc < sh.wcout)
        v = w[((size_t)kc * sh.wcin + ci) * 9 + (a + 1 - 2 * dy) * 3 + (bq + 1 - 2 * dx)];
    } else if (sh.mode >= 3) {
      // stride-2 data gradient, output phase (a, b): tap (dY, dX) in {0, 1}^2 carries w[a + 1 - 2 dY][b + 1 - 2 dX]^T
      const int ky = sh.ga[g] + 1 - 2 * dy, kx = sh.gb[g] + 1 - 2 * dx;
      if (kc < sh.wcout && nc < sh.wcin) v = w[((size_t)kc * sh.wcin + nc) * 9 + ky * 3 + kx];
    } else {
      // stem: kernel input (c, ry, rx), output (co, a, b):  w[co][c][4 dy + ry - 2a + 3][4 dx + rx - 2b + 3]
      const int c = kc >> 4, ry = (kc >> 2) & 3, rx = kc & 3;
      const int co = nc >> 2, a = (nc >> 1) & 1, bb = nc & 1;
      const int ky = 4 * dy + ry - 2 * a + 3, kx = 4 * dx + rx - 2 * bb + 3;
      if (co < sh.wcout && c < 3 && ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = w[(((size_t)co * 3 + c) * 7 + ky) * 7 + kx];
    }
    wpk[i] = __float2bfloat16_rn(v);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
int mil_wide_shape(int mode, int transposed, int wcout, int wcin, int ks, MilWideShape* out) {
  MilWideShape& sh = *out;
  sh = MilWideShape{};
  sh.mode = mode; sh.transposed = transposed; sh.wcout = wcout; sh.wcin = wcin; sh.ks = ks;
  if (mode == 0) {
    MIL_REQUIRE(ks == 1 || ks == 3, "wide_conv: unsupported window %d", ks);
    sh.kin = transposed ? wcout : wcin;
    sh.nout = transposed ? wcin : wcout;
    sh.ngroups = 1;
    const int r = ks / 2;
    int t = 0;
    for (int a = -r; a <= r; ++a)
      for (int b = -r; b <= r; ++b) { sh.tdy[0][t] = (signed char)a; sh.tdx[0][t] = (signed char)b; ++t; }
    sh.gntaps[0] = t;
  } else if (mode == 1) {
    MIL_REQUIRE(ks == 3 && !transposed, "wide_conv: the phase-split form is the 3x3 / stride-2 forward");
    sh.kin = wcin;
    sh.nout = wcout;
    sh.ngroups = 4;
    for (int g = 0; g < 4; ++g) {
      const int a = g >> 1, b = g & 1;
      sh.ga[g] = a; sh.gb[g] = b;
      int t = 0;
      // tap (dy, dx) in {-1, 0}^2 of phase (a, b) carries w[2 dy + a + 1][2 dx + b + 1]: dy = -1 needs a = 1
      for (int dy = -1; dy <= 0; ++dy)
        for (int dx = -1; dx <= 0; ++dx) {
          if ((dy == -1 && a == 0) || (dx == -1 && b == 0)) continue;
          sh.tdy[g][t] = (signed char)dy; sh.tdx[g][t] = (signed char)dx; ++t;
        }
      sh.gntaps[g] = t;
    }
  } else if (mode == 2) {
    MIL_REQUIRE(ks == 7 && wcin == 3 && !transposed, "wide_conv: the stem form is conv 7x7 / stride 2 on 3 channels");
    sh.kin = 48;
    sh.nout = 4 * wcout;
    sh.ngroups = 1;
    int t = 0;
    for (int a = -1; a <= 1; ++a)
      for (int b = -1; b <= 1; ++b) { sh.tdy[0][t] = (signed char)a; sh.tdx[0][t] = (signed char)b; ++t; }
    sh.gntaps[0] = t;
  } else if (mode >= 3 && mode <= 6) {
    // data gradient of the 3x3 / stride-2 convolution for the input pixels of parity phase (a, b):
    //   dx[2Y + a][2X + b] = sum over ky = a + 1 (mod 2), kx = b + 1 (mod 2) of w[ky][kx]^T dz[Y + dY][X + dX],
    //   dY = (a + 1 - ky) / 2, dX = (b + 1 - kx) / 2   -- 1 / 2 / 2 / 4 taps at the OUTPUT resolution, no zero-stuffing
    MIL_REQUIRE(ks == 3 && !transposed, "wide_conv: the phase-wise form is the 3x3 / stride-2 data gradient");
    const int a = (mode - 3) >> 1, b = (mode - 3) & 1;
    sh.kin = wcout;
    sh.nout = wcin;
    sh.ngroups = 1;
    sh.ga[0] = a; sh.gb[0] = b;
    int t = 0;
    for (int dY = 0; dY <= a; ++dY)
      for (int dX = 0; dX <= b; ++dX) { sh.tdy[0][t] = (signed char)dY; sh.tdx[0][t] = (signed char)dX; ++t; }
    sh.gntaps[0] = t;
  } else if (mode == 7) {
    // the same data gradient with the four input phases as OUTPUT CHANNELS (phase, ci): a 2x2-window convolution of the
    // output gradient whose weight blocks are zero where a phase does not use a tap (the kernel skips those MMAs)
    MIL_REQUIRE(ks == 3 && !transposed, "wide_conv: the phase-wise form is the 3x3 / stride-2 data gradient");
    sh.kin = wcout;
    sh.nout = 4 * wcin;
    sh.ngroups = 1;
    int t = 0;
    for (int dY = 0; dY <= 1; ++dY)
      for (int dX = 0; dX <= 1; ++dX) { sh.tdy[0][t] = (signed char)dY; sh.tdx[0][t] = (signed char)dX; ++t; }
    sh.gntaps[0] = t;
  } else {
    MIL_REQUIRE(false, "wide_conv: unknown mode %d", mode);
  }
  MIL_REQUIRE(sh.kin % 16 == 0, "wide_conv: %d input channels (need a multiple of 16)", sh.kin);
  MIL_REQUIRE(sh.nout % 64 == 0 && sh.nout <= 1024, "wide_conv: %d output channels (need a multiple of 64, at most 1024)", sh.nout);
  sh.npairs = sh.kin / 16;
  sh.nt = sh.nout % 128 == 0 ? 128 : 64;
  sh.n_ntiles = sh.nout / sh.nt;
  sh.blocks_per_ntile = 0;
  for (int g = 0; g < sh.ngroups; ++g) sh.blocks_per_ntile += sh.npairs * sh.gntaps[g];
  return 0;
}

size_t mil_wide_wpack_bytes(const MilWideShape& sh) {
  return (size_t)sh.n_ntiles * sh.blocks_per_ntile * sh.nt * 32;
}

int mil_launch_wide_pack(const float* w, void* wpk, const MilWideShape& sh, cudaStream_t s) {
  WidePackParams pp;
  pp.sh = sh;
  int b = 0;
  for (int g = 0; g < WIDE_MAX_GROUPS; ++g) {
    pp.gblk0[g] = b;
    if (g < sh.ngroups) b += sh.npairs * sh.gntaps[g];
  }
  const long long total = (long long)mil_wide_wpack_bytes(sh) / 2;
  const int blocks = (int)std::min<long long>(mil_cdiv(total, 256), 148 * 8);
  wide_pack_kernel<<<blocks, 256, 0, s>>>(w, (__nv_bfloat16*)wpk, pp);
  MIL_LAUNCH_OK();
  return 0;
}

int mil_launch_wide_conv(const void* x, const MilPF8& gx, const void* wpk, const MilWideShape& sh, const float* bias,
                         const void* res, const void* act, void* out, const MilPF8& go, int epi, float slope, int tm,
                         cudaStream_t s, int res_chunks) {
  MIL_REQUIRE(gx.n == go.n && gx.h == go.h && gx.w == go.w && gx.wp == go.wp && gx.hp == go.hp && gx.Q == go.Q,
              "wide_conv: input and output maps must share their pixel geometry");
  MIL_REQUIRE(gx.cb * 8 == sh.kin * sh.ngroups, "wide_conv: x has %d chunk planes, the packed weights expect %d", gx.cb,
              sh.kin * sh.ngroups / 8);
  MIL_REQUIRE(go.cb * 8 == sh.nout, "wide_conv: out has %d channels, the packed weights produce %d", go.cb * 8, sh.nout);
  MIL_REQUIRE(epi != MIL_EPI_DGRAD || act != nullptr, "wide_conv: the DGRAD epilogue needs the activation map");
  WideKParams kp{};
  kp.ngroups = sh.ngroups; kp.npairs = sh.npairs;
  int halo = 0, max_taps = 0, woff = 0;
  kp.nt = sh.nt; kp.n_ntiles = sh.n_ntiles;
  kp.b_tap = (uint32_t)sh.nt * 32;
  for (int g = 0; g < sh.ngroups; ++g) {
    kp.gplane0[g] = g * (sh.kin / 8);
    kp.gntaps[g] = sh.gntaps[g];
    kp.gwoff[g] = woff;
    woff += sh.npairs * sh.gntaps[g] * (int)kp.b_tap;
    max_taps = std::max(max_taps, sh.gntaps[g]);
    for (int t = 0; t < sh.gntaps[g]; ++t) {
      int sft = sh.tdy[g][t] * gx.wp + sh.tdx[g][t];  // forward reads x(q + s); the data gradient reads dz(q - s)
      if (sh.transposed) sft = -sft;
      kp.shift[g][t] = sft;
      halo = std::max(halo, sft < 0 ? -sft : sft);
    }
  }
  kp.wtile_bytes = (uint32_t)woff;
  kp.halo = halo;
  MIL_REQUIRE(halo <= gx.G, "wide_conv: the window reaches %d pixels back but the map's guard is %lld", halo, gx.G);
  const long long n_tiles = mil_cdiv(gx.Q, 128);
  const size_t hdr = (sizeof(WideSmemHeader) + 127) / 128 * 128;
  const size_t budget = 227 * 1024 - hdr;
  auto stage_of = [&](int t) { return (size_t)2 * (t * 128 + 2 * halo) * 16 + (size_t)max_taps * kp.b_tap; };
  // default: two accumulator sets (the epilogue of a unit overlaps the next unit's MMAs; measured on the layer shapes,
  // profiles/r2_wide_kernel_timings.txt): 4 M-tiles of 64 channels, 2 of 128
  if (tm <= 0) tm = 512 / (2 * sh.nt);
  tm = (int)std::min<long long>(tm, n_tiles);
  while (tm * sh.nt > 512) --tm;
  while (tm > 1 && stage_of(tm) * 2 > budget) --tm;
  MIL_REQUIRE(stage_of(tm) * 2 <= budget, "wide_conv: a ring stage of %zu bytes does not fit twice (row length %d)", stage_of(tm), gx.wp);
  kp.tm = tm;
  kp.a_plane = (uint32_t)(tm * 128 + 2 * halo) * 16;
  kp.stage_bytes = (uint32_t)stage_of(tm);
  kp.n_stages = (int)std::min<size_t>(WIDE_MAX_STAGES, budget / kp.stage_bytes);
  kp.nsets = std::min(WIDE_MAX_SETS, 512 / (tm * sh.nt));
  kp.per_tile = 0;
  if (sh.mode == 7) {
    MIL_REQUIRE(sh.n_ntiles <= 8, "wide_conv: %d output tiles in the phase-wise data gradient", sh.n_ntiles);
    kp.per_tile = 1;
    // N-tile -> the phases whose channels it holds -> the taps any of them uses (tap 2 dY + dX; phase (a, b): dY <= a, dX <= b)
    for (int nti = 0; nti < sh.n_ntiles; ++nti) {
      const int c0 = nti * sh.nt, c1 = c0 + sh.nt;
      int cnt = 0;
      for (int t = 0; t < 4; ++t) {
        bool used = false;
        for (int ph = 0; ph < 4; ++ph)
          if (c0 < (ph + 1) * sh.wcin && c1 > ph * sh.wcin && (t >> 1) <= (ph >> 1) && (t & 1) <= (ph & 1)) used = true;
        if (used) {
          kp.tile_shift[nti][cnt] = kp.shift[0][t];
          kp.tile_blk[nti][cnt] = t;
          ++cnt;
        }
      }
      kp.tile_ntaps[nti] = cnt;
    }
  }
  kp.res_chunks = res_chunks > 0 ? res_chunks : go.cb;
  kp.epi = epi;
  kp.has_bias = (bias != nullptr && epi != MIL_EPI_DGRAD) ? 1 : 0;
  kp.slope = slope;
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    MIL_CHECK_CUDA(cudaGetDevice(&dev));
    MIL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long n_units = mil_cdiv(n_tiles, tm) * sh.n_ntiles;
  const int grid = (int)std::min<long long>(n_units, n_sm);
  const size_t smem = hdr + (size_t)kp.n_stages * kp.stage_bytes;
  MIL_SET_SMEM(wide_conv_kernel, smem);
  MIL_LAUNCH_PDL(wide_conv_kernel, grid, WIDE_THREADS, smem, s, (const __nv_bfloat16*)x, gx, (const __nv_bfloat16*)wpk, bias,
                                                    (const __nv_bfloat16*)res, (const __nv_bfloat16*)act,
                                                    (__nv_bfloat16*)out, go, kp);
  return 0;
}
