// tcgen05 weight gradient of the 3x3 / 1x1 stride-1 convolutions on the PF8 layout (bf16 operands, fp32
// accumulation in TMEM), with the bias gradient as one extra accumulator column block.
// Reference semantics: autograd of nnBlocks.py:178-181 (gbm/classify_combined.py:447).
//
//   dW[t][ci][co] = sum over flat pixels q of  x[q + shift_t][ci] * dz[q][co]         db[co] = sum_q dz[q][co]
//
// GEMM view:  D[co][n] += A[co][k] * B[n][k],  k = flat pixel.  Both operands are used exactly as PF8 stores them
// -- [pixel][8 channels], i.e. "MN-major" core matrices of 8 pixels x 16 B -- so a 128-pixel K-tile needs only
// bulk-TMA copies and NO transposition: A = the dz planes, B = x planes read from a shifted start address.
//
// Three kernels share this file:
//   wgrad_sq_kernel  layers 1-2 (3 or 5 chunks on both sides): ONE M = 128 MMA per K-step yields all nine taps and
//                    the bias gradient; shifted operand planes are built in shared memory (see its own header)
//   wgrad_tc_kernel  everything else (the stem's own weight gradient lives in mil_stem_wgrad.cu).  A thin MMA is bound by its operand reads (profiles/r1_mma_cost2.txt), so for
//                    the 3x3 layers the three dx taps are CONCATENATED along N: the producer loads three copies of
//                    every x plane, shifted by -1 / 0 / +1 pixel, as consecutive shared-memory planes; N-group
//                    (chunk, dx) then sits at a uniform stride and ONE MMA per (dy, K-step) covers N = 3*Cin columns
//                    (+8 for the bias gradient when M = 64).  1x1 windows and the stride-2 3x3 convolution on its
//                    phase-split input (explicit MMA list, WgCombo) use one MMA per tap.
//
// The tap accumulators (and the bias block, B = a constant all-ones block) stay in TMEM for the whole kernel: every
// CTA accumulates its share of the pixel tiles (split-K over CTAs), then writes ONE partial record
// [tap][cin_pad][cout_pad] + [cout_pad]; a fixed-order reduction sums the records -> deterministic.  TMEM holds 512
// columns, so wide layers split the dy taps over two CTA groups (blockIdx.y).
#include <algorithm>
#include <cstdlib>

#include "mil_common.cuh"
#include "mil_conv_tc.cuh"
#include "mil_tc_ptx.cuh"

#define WG_TK 128
#define WG_MAX_STAGES 3
#define WG_THREADS 192  // warp 0 producer, warp 1 MMA issuer, warps 2..5 epilogue
#define WG_A_PLANE (WG_TK * 16)
#define WG_SLACK (8 * 1024)

struct WgSmemHeader {
  uint64_t full[WG_MAX_STAGES], empty[WG_MAX_STAGES], done;
  uint32_t tmem_base;
};
// Explicit MMA list for the stride-2 3x3 convolution on its phase-split input (mil_tc_shape_s2).  The nine taps fall
// into FOUR groups by their pixel shift (0, -1, -wp, -wp-1); the taps of a group differ only in the parity phase they
// read, and the phase planes sit at a uniform stride -- so ONE MMA per group covers all its taps: it multiplies dz with
// the `nplanes[i]` planes starting at plane0[i] (N = 8 * nplanes, padded to 16), read `shift[i]` pixels away, into the
// TMEM columns from col0[i].  Column block p of group i is plane plane0[i] + p = (phase, chunk); rtap[i][phase] is the
// tap ky * 3 + kx it belongs to (0xFF: a phase that lies between two wanted ones -- computed, not stored).  A thin M = 64
// MMA costs max(28, N / 2) cycles, so four wide MMAs (140 cycles per K-step on layer 2) replace nine thin ones (252).
// n = 0: the taps come from the window description `sh`.
struct WgCombo {
  int n, cb;
  short shift[4], col0[4], npad[4];
  unsigned char plane0[4], nplanes[4], rtap[4][4];
};

// dxcat = 1: 3x3 window, taps grouped by dy, N = (dx, chunk) planes;  dxcat = 0: one tap per MMA (1x1 window, or the
// explicit list `cmb`)
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __nv_bfloat16* __restrict__ x, MilPF8 gx, const __nv_bfloat16* __restrict__ dz, MilPF8 gz,
                float* __restrict__ partial, long long rec_stride, MilTcShape sh, int halo, int taps_per_group,
                int npad, int mma_m, int dxcat, int fold, int n_stages, WgCombo cmb, int lane_split) {
  extern __shared__ __align__(128) unsigned char smem[];
  WgSmemHeader* hd = reinterpret_cast<WgSmemHeader*>(smem);
  unsigned char* ones = smem + 128;            // 512 B of bf16 1.0: the B operand of the bias-gradient MMA
  unsigned char* stage0 = smem + 128 + 512;
  const int ntaps = sh.ntaps;                  // taps of the record layout (9 or 1)
  const int ngrp_taps = cmb.n ? cmb.n : (dxcat ? 3 : ntaps);  // MMA "taps": dy rows when dxcat
  const int nbp = dxcat ? 3 * gx.cb : gx.cb;   // B planes per stage
  const int span = dxcat ? WG_TK + 2 * gx.wp : WG_TK + 2 * halo;
  const int back = dxcat ? gx.wp : halo;       // pixels before q0 held by a plane (of the dx = 0 copy)
  const uint32_t b_plane = (uint32_t)span * 16;
  const uint32_t a_bytes = (uint32_t)gz.cb * WG_A_PLANE;
  // (building the dx = -1 / +1 copies in shared memory instead of fetching them, as wgrad_sq_kernel does, was
  // measured SLOWER for these wide layers: their MMAs already saturate the shared-memory bandwidth)
  const uint32_t b_pitch = b_plane;
  const uint32_t load_bytes = a_bytes + (uint32_t)nbp * b_plane;
  // fold = 1: one more B plane per stage, constant 1.0, so the bias gradient is one more N-group of the tap MMAs
  // (a thin M = 64 MMA costs the tensor pipe 28 cycles whatever its N: profiles/r1_mma_cost.txt)
  const uint32_t stage_bytes = a_bytes + (uint32_t)(nbp + (fold ? 1 : 0)) * b_pitch;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap_lo = blockIdx.y * taps_per_group;
  const int tap_hi = min(ngrp_taps, tap_lo + taps_per_group);
  const int ntl = tap_hi - tap_lo;
  const bool with_bias = (blockIdx.y == gridDim.y - 1);
  const long long n_tiles = mil_cdiv(gz.Q, WG_TK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&hd->full[s], 1);
      mbar_init(&hd->empty[s], 1);
    }
    mbar_init(&hd->done, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x)
    reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;  // two bf16 1.0
  if (fold)
    for (int st = 0; st < n_stages; ++st) {
      uint32_t* op = reinterpret_cast<uint32_t*>(stage0 + (size_t)st * stage_bytes + a_bytes + (size_t)nbp * b_pitch);
      for (int i = threadIdx.x; i < (int)(b_plane / 4); i += blockDim.x) op[i] = 0x3F803F80u;
    }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(&hd->tmem_base, 512);
  mil_pdl_wait();   // everything above overlaps the previous kernel's tail; global memory only from here on (PDL, mil_common.cuh)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hd->tmem_base;

  if (warp == 0) {
    // producer: whole warp in uniform control flow, one elected lane issues the bulk copies
    // (running pointers advanced by constants: one warp issues up to 38 copies per tile)
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t s0 = smem_u32(stage0), full0 = smem_u32(&hd->full[0]);
    const uint32_t cba = (uint32_t)gz.cb, cbb = (uint32_t)gx.cb;
    const char* srca = reinterpret_cast<const char*>(dz) + (gz.G + (long long)blockIdx.x * WG_TK) * 16;
    const char* srcb = reinterpret_cast<const char*>(x) + (gx.G - back - (dxcat ? 1 : 0) + (long long)blockIdx.x * WG_TK) * 16;
    const long long stra = gz.PS * 16, strb = gx.PS * 16, tstride = (long long)gridDim.x * WG_TK * 16;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&hd->empty[stage], phase ^ 1);
      if (elect_one()) {
        const uint32_t bar = full0 + (uint32_t)stage * 8;
        mbar_expect_tx_u32(bar, load_bytes);
        uint32_t dst = s0 + (uint32_t)stage * stage_bytes;
        const char* sp = srca;
        for (uint32_t c = 0; c < cba; ++c, dst += WG_A_PLANE, sp += stra) bulk_g2s_u32(dst, sp, WG_A_PLANE, bar);
        sp = srcb;
        if (dxcat) {  // three shifted fetches of every plane, slots (chunk, dx)
          for (uint32_t c = 0; c < cbb; ++c, sp += strb) {
            bulk_g2s_u32(dst, sp, b_plane, bar);
            bulk_g2s_u32(dst + b_plane, sp + 16, b_plane, bar);
            bulk_g2s_u32(dst + 2 * b_plane, sp + 32, b_plane, bar);
            dst += 3 * b_plane;
          }
        } else {
          for (uint32_t c = 0; c < cbb; ++c, dst += b_plane, sp += strb) bulk_g2s_u32(dst, sp, b_plane, bar);
        }
      }
      __syncwarp();
      srca += tstride; srcb += tstride;
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    mil_pdl_trigger();  // all of this CTA's loads are issued: the next kernel may start its prologue (PDL, mil_common.cuh)
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
        if (cmb.n) {
          for (int gi = 0; gi < cmb.n; ++gi) {
            const uint64_t bd0 = make_desc(b_base + cmb.plane0[gi] * b_pitch + (uint32_t)(back + cmb.shift[gi]) * 16, 128, b_pitch);
            const uint32_t idesc_g = idesc_base | ((uint32_t)(cmb.npad[gi] >> 3) << 17);
            const uint32_t d = tmem_base + cmb.col0[gi];
            umma_bf16(d, ad0, bd0, idesc_g, acc0);
#pragma unroll
            for (int kk = 1; kk < WG_TK / 16; ++kk) umma_bf16(d, ad0 + kk * 16, bd0 + kk * 16, idesc_g, 1u);
          }
        }
        for (int tl = 0; tl < (cmb.n ? 0 : ntl); ++tl) {
          const int tap = tap_lo + tl;
          // pixel offset of this tap's window start inside a B plane
          const int s = dxcat ? tap * gx.wp : back + sh.t_dy[tap] * gx.wp + sh.t_dx[tap];
          const uint32_t p0 = 0u;
          const uint64_t bd0 = make_desc(b_base + p0 + (uint32_t)s * 16, 128, b_pitch);
          // M = 64 accumulators occupy lanes 0-15 of every lane quarter only: lane_split = h > 0 parks the taps from
          // h on in lanes 16-31 of the same columns (TMEM holds twice as many M = 64 accumulators that way)
          const uint32_t d = (lane_split && tl >= lane_split) ? tmem_base + (16u << 16) + (tl - lane_split) * npad
                                                              : tmem_base + tl * npad;
          umma_bf16(d, ad0, bd0, idesc, acc0);
#pragma unroll
          for (int kk = 1; kk < WG_TK / 16; ++kk) umma_bf16(d, ad0 + kk * 16, bd0 + kk * 16, idesc, 1u);
        }
        if (with_bias && !fold) {
          const uint32_t d = tmem_base + (cmb.n ? cmb.col0[cmb.n - 1] + cmb.npad[cmb.n - 1] : ntl * npad);
          umma_bf16(d, ad0, ones_desc, idesc_b, acc0);
#pragma unroll
          for (int kk = 1; kk < WG_TK / 16; ++kk) umma_bf16(d, ad0 + kk * 16, ones_desc, idesc_b, 1u);
        }
        umma_commit(&hd->empty[stage]);
      }
      __syncwarp();
      first = false;
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(&hd->done);
    __syncwarp();
  } else {
    // epilogue: TMEM lane = output channel co, columns = (local tap, B plane, 8 ci)
    const int quarter = warp & 3;
    // accumulator row -> TMEM lane: M = 128: row i in lane i;  M = 64: row i in lane 32*(i/16) + i%16 (each
    // lane quarter holds 16 rows)
    const bool upper = lane >= 16;  // M = 64: second accumulator set (lane_split) or unused lanes
    const int co = mma_m == 128 ? quarter * 32 + lane
                                : ((!upper || lane_split) ? quarter * 16 + (lane & 15) : 1 << 20);
    const int ncp = cmb.n ? cmb.cb : nbp;  // column blocks (planes) of one tap accumulator
    const int coutp = gz.cb * 8, cinp = (cmb.n ? cmb.cb : gx.cb) * 8;
    mbar_wait(&hd->done, 0);
    tc_fence_after();
    if ((mma_m == 128 ? quarter * 32 : quarter * 16) < coutp) {  // warp-uniform
      float* rec = partial + (size_t)blockIdx.x * rec_stride;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      if (cmb.n) {
        for (int gi = 0; gi < cmb.n; ++gi)
          for (int p = 0; p < cmb.nplanes[gi]; ++p) {
            float v[8];
            tmem_ld8(taddr + cmb.col0[gi] + p * 8, v);
            tmem_ld_wait();
            const int plane = cmb.plane0[gi] + p, ph = plane / cmb.cb, c = plane - ph * cmb.cb;
            const int rtap = cmb.rtap[gi][ph];
            if (co < coutp && rtap != 0xFF) {
#pragma unroll
              for (int j = 0; j < 8; ++j) rec[((size_t)rtap * cinp + c * 8 + j) * coutp + co] = v[j];
            }
          }
      }
      const int ncol_taps = cmb.n ? 0 : (lane_split ? lane_split : ntl);  // accumulator column blocks to walk
      for (int tc = 0; tc < ncol_taps; ++tc) {
        // the tap this LANE finds in column block tc
        const int tl = (lane_split && upper) ? lane_split + tc : tc;
        for (int p = 0; p < ncp; ++p) {
          float v[8];
          tmem_ld8(taddr + tc * npad + p * 8, v);
          tmem_ld_wait();
          // record tap and input chunk of this column block
          const int rtap = dxcat ? (tap_lo + tl) * 3 + p % 3 : tap_lo + tl;
          const int c = dxcat ? p / 3 : p;
          if (co < coutp && tl < ntl) {
#pragma unroll
            for (int j = 0; j < 8; ++j) rec[((size_t)rtap * cinp + c * 8 + j) * coutp + co] = v[j];
          }
        }
      }
      if (with_bias) {
        float v[8];
        tmem_ld8(taddr + (cmb.n ? cmb.col0[cmb.n - 1] + cmb.npad[cmb.n - 1] : (fold ? nbp * 8 : ntl * npad)), v);
        tmem_ld_wait();
        if (co < coutp && !(lane_split && upper)) rec[(size_t)ntaps * cinp * coutp + co] = v[0];
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

// ---- "square" form for the thin layers (3 or 5 chunks on both sides) ----------------------------------------
// A thin MMA is bound by its operand reads, not by its FLOPs (profiles/r1_mma_cost.txt: M = 128 costs
// max((128 + N) / 4, N / 2) cycles, M = 64 costs max(28, N / 2)), so the cheapest form has the FEWEST MMAs per K-step:
// put the three dy shifts on the dz side and the three dx shifts on the x side,
//   dW[dy][dx][ci][co] = sum_q' x[q' + dx][ci] * dz[q' - dy*wp][co],
// and ONE M = 128 MMA per K-step produces all nine taps: rows = (co chunk, dy), columns = (ci chunk, dx) + one
// constant-one plane for the bias gradient (layer 1: 52 cycles per 16 pixels instead of 3 x 40).
//
// What bounds this kernel is the L2 -> shared-memory fill (~24 B/clk/SM measured, profiles/r1_bulk_copy.txt), so the
// shifted operand planes are NOT fetched three times.  One bulk copy per chunk brings the tile with its halo
// ([q0 - wp, q0 + TK + wp) of dz, [q0 - 1, q0 + TK + 1) of x), placed so that the UNSHIFTED window already sits in its
// final plane slot; the four otherwise idle epilogue warps then build the two shifted planes of every chunk with one
// 16-byte shared-memory copy per thread and plane.  Plane slots are pitch = 2048 + halo bytes apart (the halo of the
// in-place plane spills into its neighbours' padding), which the MMA descriptor's group stride absorbs.
#define WGS_MAX_STAGES 4
struct WgsSmemHeader {
  uint64_t full[WGS_MAX_STAGES], ready[WGS_MAX_STAGES], empty[WGS_MAX_STAGES], done;
  uint32_t tmem_base;
};

// CBA / CBB: chunk counts of dz / x as compile-time facts (3 / 5 = layers 1 / 2; 0 = read them from the geometry) so that a
// copy thread issues ALL of a tile's shared-memory loads before its first store.  Every ring stage has its OWN group of
// four copy warps (group g serves stage g, i.e. the CTA's tiles g, g + n_stages, ...; a stage that changed hands
// between groups would let a group wait on a phase parity that is still one revolution behind): the copy phase is a
// chain of LDS -> STS latencies (profiles/r2_ncu_conv_wgrad_stalls.txt: the four warps of the first version were busy
// 95 % of the time, the MMA warp waited on them a third of its time).
#define WGS_THREADS(n_stages) (64 + 128 * (n_stages))  // warp 0 producer, warp 1 MMA issuer, 4 copy warps per stage (group 0: epilogue)
template <int CBA, int CBB>
__global__ void __launch_bounds__(WGS_THREADS(WGS_MAX_STAGES), 1)
wgrad_sq_kernel(const __nv_bfloat16* __restrict__ x, MilPF8 gx, const __nv_bfloat16* __restrict__ dz, MilPF8 gz,
                float* __restrict__ partial, long long rec_stride, int npad, int n_stages) {
  extern __shared__ __align__(128) unsigned char smem[];
  WgsSmemHeader* hd = reinterpret_cast<WgsSmemHeader*>(smem);
  unsigned char* stage0 = smem + 128;
  const int na = 3 * gz.cb, nb = 3 * gx.cb;  // A plane slots (co chunk, dy), B plane slots (ci chunk, dx) + ones
  const uint32_t halo_a = (uint32_t)gz.wp * 16, halo_b = 16;
  const uint32_t pitch_a = WG_A_PLANE + halo_a, pitch_b = WG_A_PLANE + halo_b;
  const uint32_t b_bytes = (uint32_t)(nb + 1) * pitch_b;
  const uint32_t stage_bytes = b_bytes + (uint32_t)na * pitch_a;  // [B slots | ones][A slots]
  const uint32_t load_bytes = (uint32_t)gz.cb * (WG_A_PLANE + 2 * halo_a) + (uint32_t)gx.cb * (WG_A_PLANE + 2 * halo_b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = mil_cdiv(gz.Q, WG_TK);
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < npad) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&hd->full[s], 1);
      mbar_init(&hd->ready[s], 4);
      mbar_init(&hd->empty[s], 1);
    }
    mbar_init(&hd->done, 1);
    fence_barrier_init();
  }
  for (int st = 0; st < n_stages; ++st) {
    uint32_t* op = reinterpret_cast<uint32_t*>(stage0 + (size_t)st * stage_bytes + (size_t)nb * pitch_b);
    for (int i = threadIdx.x; i < WG_A_PLANE / 4; i += blockDim.x) op[i] = 0x3F803F80u;  // two bf16 1.0
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(&hd->tmem_base, tmem_cols);
  mil_pdl_wait();   // everything above overlaps the previous kernel's tail; global memory only from here on (PDL, mil_common.cuh)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hd->tmem_base;

  if (warp == 0) {
    // running pointers advanced by constants (the producer is one serial instruction stream: see conv_tc_kernel)
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t s0 = smem_u32(stage0), full0 = smem_u32(&hd->full[0]);
    const uint32_t cba = (uint32_t)gz.cb, cbb = (uint32_t)gx.cb;
    const uint32_t sz_a = WG_A_PLANE + 2 * halo_a, sz_b = WG_A_PLANE + 2 * halo_b;
    const char* srca = reinterpret_cast<const char*>(dz) + (gz.G - gz.wp + (long long)blockIdx.x * WG_TK) * 16;
    const char* srcb = reinterpret_cast<const char*>(x) + (gx.G - 1 + (long long)blockIdx.x * WG_TK) * 16;
    const long long stra = gz.PS * 16, strb = gx.PS * 16, tstride = (long long)gridDim.x * WG_TK * 16;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&hd->empty[stage], phase ^ 1);
      if (elect_one()) {
        const uint32_t bar = full0 + (uint32_t)stage * 8;
        mbar_expect_tx_u32(bar, load_bytes);
        const uint32_t bdst = s0 + (uint32_t)stage * stage_bytes;
        uint32_t da = bdst + b_bytes + pitch_a - halo_a;  // the unshifted window [q0, q0 + TK) lands in slot (c, dy = 0)
        const char* sp = srca;
        for (uint32_t c = 0; c < cba; ++c, da += 3 * pitch_a, sp += stra) bulk_g2s_u32(da, sp, sz_a, bar);
        uint32_t db = bdst + pitch_b - halo_b;
        sp = srcb;
        for (uint32_t c = 0; c < cbb; ++c, db += 3 * pitch_b, sp += strb) bulk_g2s_u32(db, sp, sz_b, bar);
      }
      __syncwarp();
      srca += tstride; srcb += tstride;
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    mil_pdl_trigger();  // all of this CTA's loads are issued: the next kernel may start its prologue (PDL, mil_common.cuh)
  } else if (warp == 1) {
    // D = f32, A = B = bf16, both MN-major, M = 128 (the M-groups past 3 * cb read whatever follows in shared
    // memory -- the next stage or the tail pad: values landing in accumulator rows nobody reads)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24) |
                           ((uint32_t)(npad >> 3) << 17);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&hd->ready[stage], phase);
      tc_fence_after();
      const uint32_t b_base = smem_u32(stage0 + (size_t)stage * stage_bytes);
      const uint64_t ad0 = make_desc(b_base + b_bytes, 128, pitch_a);
      const uint64_t bd0 = make_desc(b_base, 128, pitch_b);
      const uint32_t acc0 = first ? 0u : 1u;
      if (elect_one()) {
        umma_bf16(tmem_base, ad0, bd0, idesc, acc0);
#pragma unroll
        for (int kk = 1; kk < WG_TK / 16; ++kk) umma_bf16(tmem_base, ad0 + kk * 16, bd0 + kk * 16, idesc, 1u);
        umma_commit(&hd->empty[stage]);
      }
      __syncwarp();
      first = false;
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(&hd->done);
    __syncwarp();
  } else {
    // ---- main loop: build the shifted planes (thread i of the four warps moves pixel i of every plane) ----
    const int cg = (warp - 2) >> 2;                   // copy group = ring stage: the CTA's tiles cg, cg + n_stages, ...
    {
      const int tid = (threadIdx.x - 64) & 127;       // 0..127 = pixel of the K-tile
      const int stage = cg;
      uint32_t phase = 0;
      constexpr int MA = CBA ? CBA : 5, MB = CBB ? CBB : 5;
      const int cba = CBA ? CBA : gz.cb, cbb = CBB ? CBB : gx.cb;
      for (long long t = blockIdx.x + (long long)cg * gridDim.x; t < n_tiles; t += (long long)n_stages * gridDim.x) {
        mbar_wait(&hd->full[stage], phase);
        unsigned char* bdst = stage0 + (size_t)stage * stage_bytes;
        unsigned char* amid = bdst + b_bytes + pitch_a + (size_t)tid * 16;  // dz[q0 + tid] of chunk 0
        unsigned char* bmid = bdst + pitch_b + (size_t)tid * 16;            // x[q0 + tid] of chunk 0
        uint4 dn[MA], up[MA], lf[MB], rt[MB];
#pragma unroll
        for (int c = 0; c < MA; ++c)
          if (c < cba) {
            dn[c] = *reinterpret_cast<const uint4*>(amid + (size_t)c * 3 * pitch_a + halo_a);  // dz[q + wp]: dy = -1
            up[c] = *reinterpret_cast<const uint4*>(amid + (size_t)c * 3 * pitch_a - halo_a);  // dz[q - wp]: dy = +1
          }
#pragma unroll
        for (int c = 0; c < MB; ++c)
          if (c < cbb) {
            lf[c] = *reinterpret_cast<const uint4*>(bmid + (size_t)c * 3 * pitch_b - 16);      // x[q - 1]: dx = -1
            rt[c] = *reinterpret_cast<const uint4*>(bmid + (size_t)c * 3 * pitch_b + 16);      // x[q + 1]: dx = +1
          }
#pragma unroll
        for (int c = 0; c < MA; ++c)
          if (c < cba) {
            *reinterpret_cast<uint4*>(amid + (size_t)c * 3 * pitch_a - pitch_a) = dn[c];
            *reinterpret_cast<uint4*>(amid + (size_t)c * 3 * pitch_a + pitch_a) = up[c];
          }
#pragma unroll
        for (int c = 0; c < MB; ++c)
          if (c < cbb) {
            *reinterpret_cast<uint4*>(bmid + (size_t)c * 3 * pitch_b - pitch_b) = lf[c];
            *reinterpret_cast<uint4*>(bmid + (size_t)c * 3 * pitch_b + pitch_b) = rt[c];
          }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&hd->ready[stage]);
        phase ^= 1;
      }
    }
    // ---- epilogue: accumulator row (= TMEM lane) (c, e, j) -> output channel c * 8 + j, tap row dy = e - 1;
    // column block p = (cc, d) -> input chunk cc, tap column dx = d - 1
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int g = row >> 3, c = g / 3, e = g - c * 3;
    const int co = c * 8 + (row & 7);
    const int coutp = gz.cb * 8, cinp = gx.cb * 8;
    if (cg == 0) mbar_wait(&hd->done, 0);  // (the second copy group has no epilogue share)
    tc_fence_after();
    if (cg == 0 && quarter * 32 < na * 8) {  // warp-uniform
      float* rec = partial + (size_t)blockIdx.x * rec_stride;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      for (int p = 0; p < nb; ++p) {
        float v[8];
        tmem_ld8(taddr + p * 8, v);
        tmem_ld_wait();
        const int cc = p / 3, d = p - cc * 3;
        if (g < na) {
#pragma unroll
          for (int j = 0; j < 8; ++j) rec[((size_t)(e * 3 + d) * cinp + cc * 8 + j) * coutp + co] = v[j];
        }
      }
      float v[8];
      tmem_ld8(taddr + nb * 8, v);
      tmem_ld_wait();
      if (g < na && e == 1) rec[(size_t)9 * cinp * coutp + co] = v[0];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, tmem_cols);
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

struct WgConfig {
  int sq, dxcat, fold, npad, groups, tpg, ctas, mma_m, n_stages, lane_split;
  size_t smem;
};

static WgConfig wg_config(const MilPF8& gx, const MilPF8& gz, int ks) {
  WgConfig c;
  const long long n_tiles_sq = mil_cdiv(gz.Q, WG_TK);
  c.sq = (ks == 3 && 3 * gz.cb <= 16 && ((3 * gx.cb + 1) & 1) == 0 && (3 * gx.cb + 1) * 8 <= 256) ? 1 : 0;
  if (c.sq) {
    c.dxcat = c.fold = c.lane_split = 0;
    c.mma_m = 128;
    c.npad = (3 * gx.cb + 1) * 8;
    c.groups = 1;
    c.tpg = 9;
    c.ctas = (int)std::max<long long>(1, std::min<long long>(n_tiles_sq, wg_sm_count()));
    // stage = [B slots + ones][A slots]; the M = 128 MMA reads 16 A slots whatever 3 * cb is: tail pad for the last
    const size_t pitch_a = WG_A_PLANE + (size_t)gz.wp * 16, pitch_b = WG_A_PLANE + 16;
    const size_t stage = (size_t)(3 * gx.cb + 1) * pitch_b + (size_t)3 * gz.cb * pitch_a;
    const size_t tail = (size_t)(16 - 3 * gz.cb) * pitch_a + 128;
    c.n_stages = WGS_MAX_STAGES;
    while (c.n_stages > 1 && 128 + c.n_stages * stage + tail > 220 * 1024) --c.n_stages;
    c.smem = 128 + c.n_stages * stage + tail;
    return c;
  }
  c.dxcat = ks == 3 ? 1 : 0;
  c.mma_m = gz.cb * 8 <= 64 ? 64 : 128;
  c.fold = (c.dxcat && c.mma_m == 64) ? 1 : 0;
  c.npad = c.dxcat ? (3 * gx.cb + c.fold) * 8 : (gx.c + 15) / 16 * 16;
  const int gtaps = c.dxcat ? 3 : ks * ks;
  c.groups = (gtaps * c.npad + (c.fold ? 0 : 16) <= 512) ? 1 : 2;
  c.lane_split = 0;
  if (c.groups == 2 && c.mma_m == 64 && c.fold && ((gtaps + 1) / 2) * c.npad <= 512) {
    // M = 64 accumulators use half of the TMEM lanes: the second half of the taps goes to lanes 16-31 of the same
    // columns, and ONE CTA group covers all taps (no second group re-fetching every tile)
    c.groups = 1;
    c.lane_split = (gtaps + 1) / 2;
  }
  c.tpg = (gtaps + c.groups - 1) / c.groups;
  const long long n_tiles = mil_cdiv(gz.Q, WG_TK);
  c.ctas = (int)std::max<long long>(1, std::min<long long>(n_tiles, wg_sm_count() / c.groups));
  const int halo = ks == 3 ? gx.wp + 1 : 0;
  const size_t span = c.dxcat ? WG_TK + 2 * (size_t)gx.wp : WG_TK + 2 * (size_t)halo;
  const size_t stage = (size_t)gz.cb * WG_A_PLANE + (size_t)((c.dxcat ? 3 * gx.cb : gx.cb) + c.fold) * span * 16;
  c.n_stages = WG_MAX_STAGES;
  while (c.n_stages > 1 && 128 + 512 + c.n_stages * stage + WG_SLACK > 220 * 1024) --c.n_stages;
  c.smem = 128 + 512 + c.n_stages * stage + WG_SLACK;
  return c;
}

bool mil_wgrad_tc_supported(int dtype, int ks, int stride, int cin, int cout) {
  return dtype == MIL_BF16 && (ks == 3 || ks == 1) && stride == 1 && cin <= 80 && cout <= 80;
}

size_t mil_wgrad_tc_partial_floats(const MilPF8& gx, const MilPF8& gz, int ks) {
  const WgConfig c = wg_config(gx, gz, ks);
  return (size_t)c.ctas * ((size_t)ks * ks * gx.cb * 8 * gz.cb * 8 + gz.cb * 8);
}

// accumulate-only part: writes `*ctas_out` partial records [tap][cin_pad][cout_pad] (+[cout_pad] bias sums)
int mil_launch_wgrad_tc_partials(const void* x, const MilPF8& gx, const void* dz, const MilPF8& gz, float* partial,
                                 int ks, int* ctas_out, long long* rec_out, cudaStream_t s) {
  MIL_REQUIRE(gx.n == gz.n && gx.h == gz.h && gx.w == gz.w && gx.wp == gz.wp && gx.hp == gz.hp,
              "wgrad_tc: geometry mismatch");
  MIL_REQUIRE(ks == 3 || ks == 1, "wgrad_tc: unsupported window %d", ks);
  const WgConfig c = wg_config(gx, gz, ks);
  MilTcShape sh;
  MIL_TRY(mil_tc_shape(gx.c, gz.c, ks, &sh));
  const int halo = mil_tc_halo(sh, gx.wp);
  MIL_REQUIRE(halo <= gx.G, "wgrad_tc: the window reaches %d pixels back but the map's guard is %lld", halo, gx.G);
  const long long rec_sq = (long long)sh.ntaps * gx.cb * 8 * gz.cb * 8 + gz.cb * 8;
  if (c.sq) {
    MIL_REQUIRE(halo <= gz.G, "wgrad_tc: the window reaches %d pixels back but the gradient map's guard is %lld", halo, gz.G);
#define MIL_WGS_LAUNCH(CBA, CBB)                                                                                       \
  do {                                                                                                               \
    MIL_SET_SMEM((wgrad_sq_kernel<CBA, CBB>), (int)c.smem);                                                          \
    MIL_CHECK_CUDA(mil_launch_pdl((wgrad_sq_kernel<CBA, CBB>), dim3(c.ctas), dim3(WGS_THREADS(c.n_stages)), c.smem, s,   \
                                  (const __nv_bfloat16*)x, gx, (const __nv_bfloat16*)dz, gz, partial, rec_sq, c.npad,  \
                                  c.n_stages));                                                                      \
  } while (0)
    if (gz.cb == 3 && gx.cb == 3) MIL_WGS_LAUNCH(3, 3);
    else if (gz.cb == 5 && gx.cb == 5) MIL_WGS_LAUNCH(5, 5);
    else MIL_WGS_LAUNCH(0, 0);
#undef MIL_WGS_LAUNCH
    MIL_LAUNCH_OK();
    *ctas_out = c.ctas;
    *rec_out = rec_sq;
    return 0;
  }
  MIL_REQUIRE(c.mma_m == 64 || c.npad % 16 == 0, "wgrad_tc: N = %d is not a multiple of 16 (M = 128)", c.npad);
  MIL_REQUIRE(c.npad <= 256, "wgrad_tc: N = %d too wide", c.npad);
  MIL_REQUIRE(c.smem <= 227 * 1024, "wgrad_tc: tile width %d needs %zu bytes of shared memory", gx.w, c.smem);
  MIL_SET_SMEM((wgrad_tc_kernel), (int)c.smem);
  const long long rec = (long long)sh.ntaps * gx.cb * 8 * gz.cb * 8 + gz.cb * 8;
  MIL_LAUNCH_PDL(wgrad_tc_kernel, dim3(c.ctas, c.groups), WG_THREADS, c.smem, s, 
      (const __nv_bfloat16*)x, gx, (const __nv_bfloat16*)dz, gz, partial, rec, sh, halo, c.tpg, c.npad, c.mma_m, c.dxcat,
      c.fold, c.n_stages, WgCombo{}, c.lane_split);
  *ctas_out = c.ctas;
  *rec_out = rec;
  return 0;
}

// weight + bias gradient of a 3x3 / stride-2 convolution from its phase-split input (mil_launch_split2) and the
// output gradient, both at the OUTPUT resolution: nine single-tap MMAs per K-step over a 2x2 window
int mil_launch_wgrad_tc_s2(const void* xs2, const MilPF8& gs, const void* dz, const MilPF8& gz, float* partial, float* dw,
                           float* db, int cin, cudaStream_t s) {
  const int cb = (cin + 7) / 8;
  MIL_REQUIRE(gs.n == gz.n && gs.h == gz.h && gs.w == gz.w && gs.wp == gz.wp && gs.cb == 4 * cb,
              "wgrad_tc_s2: geometry mismatch");
  WgCombo cmb;
  cmb.n = 4;
  cmb.cb = cb;
  int col = 0;
  for (int gi = 0; gi < 4; ++gi) {
    // group (dy, dx): w[ky][kx] multiplies phase (a, b) at pixel offset (dy, dx) with ky = 2 dy + a + 1, kx = 2 dx + b + 1,
    // i.e. dy = -1 needs a = 1 (ky = 0), dy = 0 takes a = 0 (ky = 1) and a = 1 (ky = 2); the same for the columns
    const int dy = gi >> 1 ? -1 : 0, dx = gi & 1 ? -1 : 0;
    int first = 4, last = -1;
    for (int ph = 0; ph < 4; ++ph) {
      const int a = ph >> 1, b = ph & 1;
      const bool ok = (dy == 0 || a == 1) && (dx == 0 || b == 1);
      const int ky = dy == -1 ? 0 : a + 1, kx = dx == -1 ? 0 : b + 1;
      cmb.rtap[gi][ph] = ok ? (unsigned char)(ky * 3 + kx) : 0xFF;
      if (ok) { first = std::min(first, ph); last = std::max(last, ph); }
    }
    cmb.shift[gi] = (short)(dy * gs.wp + dx);
    cmb.plane0[gi] = (unsigned char)(first * cb);
    cmb.nplanes[gi] = (unsigned char)((last - first + 1) * cb);
    cmb.npad[gi] = (short)((cmb.nplanes[gi] * 8 + 15) / 16 * 16);
    cmb.col0[gi] = (short)col;
    col += cmb.npad[gi];
  }
  MIL_REQUIRE(col + 16 <= 512 && cmb.npad[0] <= 256, "wgrad_tc_s2: %d input channels need %d TMEM columns", cin, col + 16);
  MilTcShape sh;
  MIL_TRY(mil_tc_shape(cin, gz.c, 3, &sh));  // record layout: nine taps
  const int halo = gs.wp + 1;
  MIL_REQUIRE(halo <= gs.G, "wgrad_tc_s2: the window reaches %d pixels back but the map's guard is %lld", halo, gs.G);
  const int mma_m = gz.cb * 8 <= 64 ? 64 : 128;
  const int npad = (cb * 8 + 15) / 16 * 16;
  const int groups = 1;
  const int tpg = 4;
  const long long n_tiles = mil_cdiv(gz.Q, WG_TK);
  const int ctas = (int)std::max<long long>(1, std::min<long long>(n_tiles, wg_sm_count() / groups));
  const size_t span = WG_TK + 2 * (size_t)halo;
  const size_t stage = (size_t)gz.cb * WG_A_PLANE + (size_t)gs.cb * span * 16;
  int n_stages = WG_MAX_STAGES;
  while (n_stages > 1 && 128 + 512 + n_stages * stage + WG_SLACK > 220 * 1024) --n_stages;
  const size_t smem = 128 + 512 + n_stages * stage + WG_SLACK;
  MIL_REQUIRE(smem <= 227 * 1024, "wgrad_tc_s2: tile width %d needs %zu bytes of shared memory", gs.w, smem);
  MIL_SET_SMEM((wgrad_tc_kernel), (int)smem);
  const long long rec = (long long)9 * cb * 8 * gz.cb * 8 + gz.cb * 8;
  MIL_LAUNCH_PDL(wgrad_tc_kernel, dim3(ctas, groups), WG_THREADS, smem, s, (const __nv_bfloat16*)xs2, gs, (const __nv_bfloat16*)dz, gz,
                                                              partial, rec, sh, halo, tpg, npad, mma_m, 0, 0, n_stages, cmb, 0);
  return mil_launch_reduce_conv_w(partial, ctas, rec, dw, db, gz.c, cin, 3, s);
}

int mil_launch_wgrad_tc(const void* x, const MilPF8& gx, const void* dz, const MilPF8& gz, float* partial, float* dw,
                        float* db, int ks, cudaStream_t s) {
  int ctas;
  long long rec;
  MIL_TRY(mil_launch_wgrad_tc_partials(x, gx, dz, gz, partial, ks, &ctas, &rec, s));
  return mil_launch_reduce_conv_w(partial, ctas, rec, dw, db, gz.c, gx.c, ks, s);
}
