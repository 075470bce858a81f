// tcgen05 implicit-GEMM convolution on the PF8 layout -- forward AND data gradient of the 3x3 / 1x1 layers, the
// stride-2 3x3 convolutions in their phase-split forms (mil_tc_shape_s2 / mil_tc_shape_s2_dgrad; the block into
// layer 4 still uses full-resolution evaluation + subsampled store) and the stem's 7x7 / stride-2 convolution in
// space-to-depth form.  bf16 operands, fp32 accumulation in TMEM, fused epilogue (bias, residual, LeakyReLU /
// LeakyReLU').
// Reference semantics: nnBlocks.py:178-187 (conv3x3 + bias -> [+identity] -> LeakyReLU(0.1)), gbm/model.py:24-25,51-52
// (conv1 + LeakyReLU) and their autograd.
//
// GEMM view:  D[128 flat pixels][Cout] = sum over K-groups g = (tap t, 8-channel chunk c) of
//             A_g[128][8] * B_g[8][Cout],   A_g[i][:] = x[chunk c][q0 + i + shift_t][0..7]
// Because PF8 turns a tap into a constant shift of the flat pixel index, ALL taps read the same shared-memory
// copy of the input: per tile one bulk-TMA copy per channel chunk brings in the 128 + 2*halo pixels around the
// tile ("span"), 16 bytes per pixel, and a tap is just a different START ADDRESS of the (un-swizzled, K-major)
// UMMA descriptor.  8 consecutive pixels x 16 B = one 128-byte core matrix (SBO = 128 B); the two 8-channel
// halves of a K=16 MMA are two (tap, chunk) groups, LBO = their address distance.  Weights sit in shared
// memory for the whole kernel in the matching core-matrix layout.
//
// Warp roles (one persistent CTA per SM, tiles strided over CTAs):
//   warp 0      : producer -- bulk-TMA (cp.async.bulk) of the span planes into a ring of up to 8 stages (mbarrier tx),
//                 plus an L2 prefetch of the residual / activation rows the tile's epilogue will read
//   warps 1..2  : MMA issuers, alternate tiles -- whole warp in uniform control flow, one elected lane issues the
//                 tcgen05.mma chain of a tile (descriptor templates from kernel parameters + the stage base) into one
//                 of NG TMEM accumulator stages; ONE tcgen05.commit per tile releases the smem stage and publishes
//                 the accumulator
//   warps 3..   : epilogue, NG groups of 4 warps (NG = 4 up to 5 output chunks, 2 above).  Per tile a thread (=
//                 pixel) first issues its residual / activation loads, THEN waits for the accumulator, pulls its
//                 WHOLE row out of TMEM, hands the accumulator stage back, and only then does the arithmetic and the
//                 16-byte coalesced stores; pad pixels are written as zeros (PF8 invariant).  The epilogue is
//                 compiled per (chunk count, kind): conv_tc_kernel<MAXCB, NG, MODE>.
#include <algorithm>

#include "mil_common.cuh"
#include "mil_conv_tc.cuh"
#include "mil_tc_ptx.cuh"

#define TC_M 128
#define TC_MAX_STAGES 8  // the ring depth is chosen per launch: as many stages as shared memory holds
#define TC_MAX_ACC 4     // accumulator stages = epilogue groups (4 for <= 40 output channels, 2 above)

// ---- weight pre-pack: fp32 wp[tap][kin_pad][nout_pad8] -> bf16 B operand blocks --------------------------
// B block of MMA j: [half h][n (npad rows)][8 k-elements], group (tap,chunk) = order[2j+h]  (0xFF = dummy)
__global__ void pack_tc_kernel(const float* __restrict__ wp, __nv_bfloat16* __restrict__ wtc, MilTcShape sh) {
  const int total = sh.nmma * 2 * sh.npad * 8;
  const int kinp = sh.cbin * 8, noutp = sh.cbout * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7, n = (i >> 3) % sh.npad, g = (i >> 3) / sh.npad;  // g = 2j+h
    const int tap = sh.g_tap[g], chunk = sh.g_chunk[g];
    float v = 0.f;
    if (tap != 0xFF && n < noutp) v = wp[((size_t)tap * kinp + chunk * 8 + e) * noutp + n];
    wtc[i] = __float2bfloat16_rn(v);
  }
}

// All of a pass's convolutions in ONE launch, straight from the PyTorch layout w[cout][cin][tap] (no fp32 staging):
// blockIdx.x = table entry.  Forward: K = cin, N = cout;  data gradient (transposed): K = cout, N = cin.
#define MIL_TC_PACK_MAX 28
struct TcPackEntry {
  const float* w;
  __nv_bfloat16* wtc;
  short cout, cin;
  unsigned char taps, transposed, nmma, npad;
  unsigned char s2cb;       // != 0: phase-split stride-2 forward, chunk planes = phase * s2cb + chunk
  unsigned char pair_cbh;   // != 0: stride-2 data gradient of row parity pair_a, output rows n = (b, ci)
  unsigned char pair_a;
  unsigned char fold_cbm;   // != 0: folded form, K-groups of chunks >= fold_cbm read the projection weight w2 [cout][cin2]
  short cin2;
  const float* w2;
  signed char t_dy[MIL_TC_MAX_TAPS], t_dx[MIL_TC_MAX_TAPS];
  unsigned char g_tap[2 * MIL_TC_MAX_MMA], g_chunk[2 * MIL_TC_MAX_MMA];
};
struct TcPackTable {
  TcPackEntry e[MIL_TC_PACK_MAX];
};
__global__ void pack_tc_table_kernel(const __grid_constant__ TcPackTable t) {
  const TcPackEntry& p = t.e[blockIdx.x];
  const int total = p.nmma * 2 * p.npad * 8;
  const int kin = p.transposed ? p.cout : p.cin, nout = p.transposed ? p.cin : p.cout;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
    const int e = i & 7, n = (i >> 3) % p.npad, g = (i >> 3) / p.npad;  // g = 2j+h
    const int tap = p.g_tap[g];
    int k = p.g_chunk[g] * 8 + e, stap = tap == 0xFF ? 0 : tap, nn = n;
    bool ok = true;
    if (p.pair_cbh != 0 && tap != 0xFF) {
      const int b = n / (p.pair_cbh * 8), dY = -p.t_dy[tap], dX = -p.t_dx[tap];
      nn = n - b * p.pair_cbh * 8;
      const int ky = p.pair_a ? (dY ? 0 : 2) : 1;
      const int kx = b ? (dX ? 0 : 2) : 1;
      ok = b < 2 && !(b == 0 && dX != 0);
      stap = ky * 3 + kx;
    }
    if (p.s2cb != 0 && tap != 0xFF) {
      const int plane = p.g_chunk[g], phase = plane / p.s2cb;
      k = (plane - phase * p.s2cb) * 8 + e;
      const int ky = 2 * ((tap >> 1) - 1) + (phase >> 1) + 1, kx = 2 * ((tap & 1) - 1) + (phase & 1) + 1;
      stap = ky * 3 + kx;
    }
    float v = 0.f;
    if (p.fold_cbm != 0 && tap != 0xFF && p.g_chunk[g] >= p.fold_cbm) {
      const int k2 = (p.g_chunk[g] - p.fold_cbm) * 8 + e;
      if (nn < nout && k2 < p.cin2) v = p.w2[(size_t)nn * p.cin2 + k2];
    } else if (tap != 0xFF && ok && nn < nout && k < kin) {
      const int co = p.transposed ? k : nn, ci = p.transposed ? nn : k;
      v = p.w[((size_t)co * p.cin + ci) * p.taps + stap];
    }
    p.wtc[i] = __float2bfloat16_rn(v);
  }
}

// ---- the kernel ----------------------------------------------------------------------------------------
struct TcSmemHeader {
  uint64_t full[TC_MAX_STAGES], empty[TC_MAX_STAGES], acc_empty[TC_MAX_ACC], b_full;
  uint32_t tmem_base;
  alignas(16) float bias[96];
};


__device__ __forceinline__ uint4 ld_nc16(const __nv_bfloat16* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void unpack8(const uint4& r, float v[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

// MAXCB: upper bound of the output chunks (sizes the per-thread register arrays exactly); NG: accumulator
// stages = epilogue groups of 4 warps.  The thin layers are bound by the epilogue's instruction stream, not by
// the tensor pipe, so they get 4 groups; the wide layers (more registers per thread) get 2.
enum { TCM_GENERIC = 0, TCM_PLAIN, TCM_FWD, TCM_FWD_RES, TCM_DGRAD, TCM_DGRAD_RES };

// Column split: with more than five output chunks a pixel's accumulator row goes to TWO epilogue threads (chunks
// [0, 4) and [4, MAXCB); mask words are 4 chunks wide, so each half owns whole words).  The wide layers were bound by
// the epilogue's instruction stream -- ~1000 dependent instructions per pixel on two warps per scheduler
// (profiles/r2_ncu_conv_wide.txt: the epilogue warps waited for an accumulator 9 % of their time while the tensor pipe
// sat at 50 %); twice the warps at half the registers each hide those latencies.
template <int MAXCB, int MODE>
struct TcSplit {
  static constexpr int value = (MAXCB > 5 && MODE != TCM_GENERIC) ? 2 : 1;
  static constexpr int at = value == 2 ? 4 : MAXCB;  // first chunk of the second half
};
template <int C>
struct TcInt { static constexpr int value = C; };

template <int MAXCB, int NG, int MODE>
__global__ void __launch_bounds__(96 + NG * 128 * TcSplit<MAXCB, MODE>::value, 1)
conv_tc_kernel(const __nv_bfloat16* __restrict__ x, MilPF8 gx, const __nv_bfloat16* __restrict__ wtc,
               const float* __restrict__ bias, const __nv_bfloat16* res, const __nv_bfloat16* act,
               __nv_bfloat16* out, MilPF8 go, MilTcShape sh, const __grid_constant__ TcIssue iss, int epi,
               int sub, int halo, int n_stages, MilPF8 gr, int res_half, int up, int cbh,
               const uint32_t* __restrict__ mask_in, uint32_t* __restrict__ mask_out) {
  extern __shared__ __align__(128) unsigned char smem[];
  TcSmemHeader* hd = reinterpret_cast<TcSmemHeader*>(smem);
  const uint32_t hdr_bytes = (uint32_t)((sizeof(TcSmemHeader) + 127) / 128 * 128);
  unsigned char* bsm = smem + hdr_bytes;                      // B operand blocks
  const uint32_t b_bytes = (uint32_t)sh.nmma * 2 * sh.npad * 16;
  const int span = TC_M + 2 * halo;
  const uint32_t plane = (uint32_t)span * 16;                 // one channel-chunk plane of a stage
  const uint32_t stage_bytes = plane * (sh.cbin + 1);         // + one all-zero plane (odd K-group count)
  unsigned char* asm0 = bsm + ((b_bytes + 127) / 128 * 128);  // A stages
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = mil_cdiv(gx.Q, TC_M);  // tiles run over the INPUT resolution (== output unless sub)
  const uint32_t acc_stride = (uint32_t)((sh.npad + 31) / 32 * 32);
  const bool pf_rows = !sub && !up && !res_half;  // residual / activation rows share the tile's flat pixel range
  uint32_t tmem_cols = 32;
  while (tmem_cols < acc_stride * NG) tmem_cols <<= 1;

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) { mbar_init(&hd->full[s], 1); mbar_init(&hd->empty[s], 1); }
    for (int a = 0; a < NG; ++a) mbar_init(&hd->acc_empty[a], 4 * TcSplit<MAXCB, MODE>::value);
    mbar_init(&hd->b_full, 1);
    fence_barrier_init();
  }
  for (int s = 0; s < n_stages; ++s) {  // zero plane of every stage
    uint4* zp = reinterpret_cast<uint4*>(asm0 + (size_t)s * stage_bytes + (size_t)sh.cbin * plane);
    for (int i = threadIdx.x; i < span; i += blockDim.x) zp[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();  // generic-proxy writes (zero planes) -> visible to the tensor-core (async) proxy
  if (warp == 1) tmem_alloc(&hd->tmem_base, tmem_cols);
  // everything above overlaps the previous kernel's tail; global memory only from here on (mil_common.cuh, PDL)
  mil_pdl_wait();
  for (int i = threadIdx.x; i < 96; i += blockDim.x)
    hd->bias[i] = (bias != nullptr && i < go.c) ? bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hd->tmem_base;

  if (warp == 0) {
    // ===================== producer (whole warp in uniform control flow, one elected lane issues) ==========
    if (elect_one()) {
      mbar_expect_tx(&hd->b_full, b_bytes);
      bulk_g2s(bsm, wtc, b_bytes, &hd->b_full);
    }
    __syncwarp();
    // The producer's instruction stream is on the kernel's critical path: ONE warp issues every tile's copies, and with
    // the addresses recomputed per copy (64-bit multiplies, generic -> shared conversions, one bulk prefetch per
    // residual row) it needed ~1000 cycles per tile on layer 1 -- more than the tensor pipe (620) or HBM
    // (profiles/r2_ncu_conv_wgrad_stalls.txt: the producer never waited for a free stage).  So: every address is a
    // running pointer advanced by a constant, and the L2 prefetch of the epilogue's residual / activation rows is one
    // 128-byte line per LANE (a 2 KB row = 16 lanes) instead of a bulk prefetch per row issued by the elected lane.
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t cbin = (uint32_t)sh.cbin, tx_bytes = plane * cbin;
    const uint32_t a0 = smem_u32(asm0), full0 = smem_u32(&hd->full[0]);
    const char* src = reinterpret_cast<const char*>(x) + (gx.G - halo + (long long)blockIdx.x * TC_M) * 16;  // chunk 0
    const long long cstride = gx.PS * 16, tstride = (long long)gridDim.x * TC_M * 16;
    // prefetch: lane -> (row of pass k = chunk 2k + lane / 16, line lane % 16)
    const bool pf_res = pf_rows && res != nullptr, pf_act = pf_rows && epi == MIL_EPI_DGRAD && mask_in == nullptr;
    const int npass = (pf_res || pf_act) ? (sh.cbout + 1) / 2 : 0;
    const long long pf_lane = ((long long)(lane >> 4) * go.PS + go.G + (long long)blockIdx.x * TC_M) * 16 + (lane & 15) * 128;
    const long long pf_pass = 2 * go.PS * 16;
    const char* pres = reinterpret_cast<const char*>(res) + pf_lane;
    const char* pact = reinterpret_cast<const char*>(act) + pf_lane;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&hd->empty[stage], phase ^ 1);
      if (elect_one()) {
        const uint32_t bar = full0 + (uint32_t)stage * 8;
        mbar_expect_tx_u32(bar, tx_bytes);
        uint32_t dst = a0 + (uint32_t)stage * stage_bytes;
        const char* sp = src;
        for (uint32_t c = 0; c < cbin; ++c, dst += plane, sp += cstride) bulk_g2s_u32(dst, sp, plane, bar);
      }
      __syncwarp();
      // the epilogue of this tile reads its residual / activation rows with ordinary loads a few tiles from now and
      // waits out their full latency (only NG tiles are in flight): have L2 fetch those rows already
      for (int k = 0; k < npass; ++k) {
        if (2 * k + (lane >> 4) < sh.cbout) {
          if (pf_res) prefetch_l2_line(pres + k * pf_pass);
          if (pf_act) prefetch_l2_line(pact + k * pf_pass);
        }
      }
      src += tstride; pres += tstride; pact += tstride;
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    mil_pdl_trigger();  // all of this CTA's loads are issued: the next kernel may start its prologue (PDL, mil_common.cuh)
  } else if (warp <= 2) {
    // ===================== MMA issuers: two warps take alternate tiles (uniform control flow, one elected lane) ====
    // The per-tile fixed cost of an issuing warp (two mbarrier waits, fence, elect, commit) is ~1/3 of its loop on
    // the thin layers; with two warps it overlaps the other warp's MMA issue.  ONE commit per tile: `empty[stage]`
    // tells the producer that the stage is free AND the epilogue that the accumulator is complete.
    // instruction descriptor: D = f32, A = B = bf16, both K-major, N = npad, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(sh.npad >> 3) << 17) |
                           ((uint32_t)(TC_M >> 4) << 24);
    mbar_wait(&hd->b_full, 0);
    const int nmma = sh.nmma;
    const uint64_t b_add = (uint64_t)(smem_u32(bsm) >> 4);
    // CTA-local tile counter it = warp - 1, +2 per iteration, kept as ring positions + phase bits: the issuing warps'
    // per-tile bookkeeping is on the kernel's critical path (profiles/r2_ncu_conv_wgrad_stalls.txt: with 64-bit
    // divisions here the two warps needed ~1700 cycles per tile each while the tensor pipe needs 620)
    int stage = warp - 1, acc = warp - 1;        // n_stages and NG are even: a step of 2 wraps at most once
    uint32_t full_par = 0, acc_par = 1;
    const uint32_t a_base = smem_u32(asm0) >> 4, a_step = stage_bytes >> 4;
    for (long long t = blockIdx.x + (long long)(warp - 1) * gridDim.x; t < n_tiles; t += 2LL * gridDim.x) {
      mbar_wait(&hd->acc_empty[acc], acc_par);
      mbar_wait(&hd->full[stage], full_par);
      tc_fence_after();
      // the start-address field sits in the low 14 bits of the descriptor: adding a base cannot carry
      const uint64_t a_add = (uint64_t)(a_base + (uint32_t)stage * a_step);
      const uint32_t d = tmem_base + acc * acc_stride;
      if (elect_one()) {
#pragma unroll 4
        for (int j = 0; j < nmma; ++j) umma_bf16(d, iss.a_desc[j] + a_add, iss.b_desc[j] + b_add, idesc, j > 0);
        umma_commit(&hd->empty[stage]);
      }
      __syncwarp();
      stage += 2;
      if (stage >= n_stages) { stage -= n_stages; full_par ^= 1; }
      acc += 2;
      if (acc >= NG) { acc -= NG; acc_par ^= 1; }
    }
  } else {
    // ===================== epilogue: group eg serves accumulator stage eg =====================
    // MODE != TCM_GENERIC: chunk count, epilogue kind and residual are compile-time facts -- the per-chunk branches
    // fold away (the generic form executes ~540 instructions per pixel on layer 1, the specialised one about a third)
    constexpr bool SPEC = MODE != TCM_GENERIC;
    constexpr int CSPLIT = TcSplit<MAXCB, MODE>::value;
    const int eg = ((warp - 3) >> 2) / CSPLIT;   // epilogue group = accumulator stage
    const int csub = ((warp - 3) >> 2) % CSPLIT;  // which half of the chunk range this warp takes (warp-uniform)
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int cbout = SPEC ? MAXCB : sh.cbout;
    const bool has_res = SPEC ? (MODE == TCM_FWD_RES || MODE == TCM_DGRAD_RES) : res != nullptr;
    const bool has_act = SPEC ? (MODE == TCM_DGRAD || MODE == TCM_DGRAD_RES) : epi == MIL_EPI_DGRAD;
    const bool do_bias = SPEC ? (MODE == TCM_FWD || MODE == TCM_FWD_RES) : epi != MIL_EPI_DGRAD;
    const bool do_lrelu = SPEC ? (MODE == TCM_FWD || MODE == TCM_FWD_RES) : epi == MIL_EPI_FWD;
    int estage = eg;        // ring position + phase bit of this group's tiles (CTA-local tile counter eg, +NG per
    uint32_t epar = 0;      // iteration; n_stages is a multiple of NG)
    // This thread's pixel as (image n, in-plane offset r), advanced incrementally from tile to tile: the only
    // 64-bit divisions of the kernel happen here, once (the epilogue's instruction stream is what bounds the
    // small-channel layers, not the tensor pipe).
    const long long step_q = (long long)NG * gridDim.x * TC_M;  // this group takes every NG-th tile of the CTA
    const int step_n = (int)(step_q / gx.P), step_r = (int)(step_q % gx.P);
    const long long q_first = ((long long)blockIdx.x + (long long)eg * gridDim.x) * TC_M + row;
    int n = (int)(q_first / gx.P), r = (int)(q_first % gx.P);
    const float inv_wp = 1.0f / (float)gx.wp;
    const int P = (int)gx.P;
    const long long ostride = go.PS * 8, rstride = (res_half ? gr.PS : go.PS) * 8;  // elements between chunks
    // element offset of output chunk c: plane c, or in the stride-2 data-gradient mode plane c % cbh, pixel + c / cbh
#define KOFF(c) (cbh == 0 ? (long long)(c) * ostride : (long long)((c) >= cbh ? (c) - cbh : (c)) * ostride + ((c) >= cbh ? 8 : 0))
    // the whole per-tile loop, compiled per chunk range [C0, C1) of the accumulator row (all chunks are pulled from TMEM
    // at once, so that the accumulator stage is handed back to the MMA warps before any arithmetic)
    auto run = [&](auto c0_tag, auto c1_tag) {
    constexpr int C0 = decltype(c0_tag)::value, C1 = decltype(c1_tag)::value, NC = C1 - C0;
    constexpr int W0 = C0 / 4, W1 = (C1 + 3) / 4, MW = W1 - W0;  // sign-mask words of this range
    for (long long t = blockIdx.x + (long long)eg * gridDim.x; t < n_tiles; t += (long long)NG * gridDim.x) {
      // flat index q = n * P + r at the input resolution; qo = where the pixel is stored.  sub: the stride-2
      // convolutions are evaluated at full resolution and only the even (y, x) positions are kept.
      bool in_range = n < gx.n;
      bool is_pad = true, res_ok = true, second_ok = true;
      long long qo = (long long)n * P + r, qres = qo;
      if (in_range) {
        int y = __float2int_rz(((float)r + 0.5f) * inv_wp);  // r / wp for r < 2^22 (exact after the fix-up)
        if (y * gx.wp > r) --y;
        else if ((y + 1) * gx.wp <= r) ++y;
        const int xo = r - y * gx.wp;
        if (up) {
          // data gradient of a stride-2 convolution, input rows of parity a = up - 1: tiles run over the
          // half-resolution gradient; output chunks [0, cbh) belong at (2y + a, 2x), chunks [cbh, 2 cbh) at
          // (2y + a, 2x + 1) of the full-resolution map `go` (pads are not touched); the residual (if any) is a
          // half-resolution map of its own and feeds the first pixel only
          const int yf = 2 * y + (up - 1), xf = 2 * xo;
          in_range = y < gx.h && xo < gx.w && yf < go.h && xf < go.w;
          second_ok = xf + 1 < go.w;
          is_pad = false;
          qo = (long long)n * go.P + (long long)yf * go.wp + xf;
          qres = (long long)n * gr.P + (long long)y * gr.wp + xo;
        } else if (!sub) {
          is_pad = (y >= gx.h) || (xo >= gx.w);
          if (res_half) {  // the residual lives at HALF resolution and only feeds the even (y, x) positions
            res_ok = !(y & 1) && !(xo & 1);
            qres = (long long)n * gr.P + (long long)(y >> 1) * gr.wp + (xo >> 1);
          }
        } else {
          const int yh = y >> 1, xh = xo >> 1;
          in_range = !(y & 1) && !(xo & 1) && yh <= go.h && xh <= go.w;
          is_pad = (yh == go.h) || (xh == go.w);
          qo = (long long)n * go.P + (long long)yh * go.wp + xh;
          qres = qo;
        }
      }
      n += step_n;
      r += step_r;
      if (r >= P) { r -= P; ++n; }
      const bool live = in_range && !is_pad;
      __nv_bfloat16* po = out + (go.G + qo) * 8;  // chunk c: po + c * ostride
      // 1. residual / activation loads go out BEFORE we wait for the tensor core (the producer has already asked L2
      // for these rows).  Issuing them after the accumulator drain instead -- fewer live registers on the wide
      // layers -- was measured slower, even against a few spilled registers.
      uint4 rres[NC], ract[NC];
      uint32_t rmask[MW], wmask[MW];
#pragma unroll
      for (int w = 0; w < MW; ++w) rmask[w] = wmask[w] = 0;
      if (live) {
        if (has_res) {
          const __nv_bfloat16* pr = res + ((res_half ? gr.G : go.G) + qres) * 8;
#pragma unroll
          for (int c = C0; c < C1; ++c)
            if (c < cbout) rres[c - C0] = (res_ok && (cbh == 0 || c < cbh)) ? ld_nc16(pr + c * rstride) : make_uint4(0, 0, 0, 0);
        }
        if (has_act && mask_in == nullptr) {
          const __nv_bfloat16* pa = act + (go.G + qo) * 8;
#pragma unroll
          for (int c = C0; c < C1; ++c)
            if (c < cbout && (c < cbh || cbh == 0 || second_ok)) ract[c - C0] = ld_nc16(pa + KOFF(c));
        }
        if (has_act && mask_in != nullptr) {
          // sign bits of the activation (written by the forward epilogue): 4 bytes per pixel and 4 chunks instead of
          // 16 bytes per pixel and chunk.  Stride-2 data gradient: chunks [cbh, 2 cbh) belong to the next pixel.
          const uint32_t* pm = mask_in + go.G + qo;
#pragma unroll
          for (int w = W0; w < W1; ++w) {
            if (cbh == 0) {
              if (w * 4 < cbout) rmask[w - W0] = __ldg(pm + (long long)w * go.PS);
            } else {
              // word w covers kernel chunks 4w .. 4w+3 = (pixel b, map chunk c) pairs; gather their bytes
              uint32_t v = 0;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int k = w * 4 + e;
                if (k < cbout) {
                  const int b = k >= cbh ? 1 : 0, c = k - b * cbh;
                  if (b == 0 || second_ok) {
                    const uint32_t mwrd = __ldg(pm + b + (long long)(c >> 2) * go.PS);
                    v |= ((mwrd >> ((c & 3) * 8)) & 0xFFu) << (e * 8);
                  }
                }
              }
              rmask[w - W0] = v;
            }
          }
        }
      }
      mbar_wait(&hd->empty[estage], epar);
      tc_fence_after();
      estage += NG;
      if (estage >= n_stages) { estage -= n_stages; epar ^= 1; }
      const uint32_t taddr = tmem_base + eg * acc_stride + ((uint32_t)(quarter * 32) << 16);
      {
        // 2. pull this thread's share of the accumulator row out of TMEM
        float acc[NC][8];
#pragma unroll
        for (int k = 0; k < NC; ++k)
          if (C0 + k < cbout) tmem_ld8(taddr + (C0 + k) * 8, acc[k]);
        tmem_ld_wait();
        // 3. the accumulator stage is free again: the MMA warp can start the tile after next
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hd->acc_empty[eg]);
        // 4. arithmetic + stores: one divergent branch per batch, straight-line code inside
        if (live) {
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            const int c = C0 + k;
            if (c < cbout) {
              float* v = acc[k];
              if (has_res) {
                float rv[8];
                unpack8(rres[k], rv);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += rv[j];
              }
              if (do_bias) {
                const float4 b0 = *reinterpret_cast<const float4*>(&hd->bias[c * 8]);
                const float4 b1 = *reinterpret_cast<const float4*>(&hd->bias[c * 8 + 4]);
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += bv[j];
              }
              if (do_lrelu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], MIL_SLOPE * v[j]);  // == x > 0 ? x : slope * x
              } else if (has_act && mask_in != nullptr) {
                const uint32_t bits = rmask[(c >> 2) - W0] >> ((c & 3) * 8);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (!((bits >> j) & 1u)) v[j] *= MIL_SLOPE;
              } else if (has_act) {
                float av[8];
                unpack8(ract[k], av);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (!(av[j] > 0.f)) v[j] *= MIL_SLOPE;
              }
              if (do_lrelu && mask_out != nullptr) {
                // pack here (not in mil_store8) and derive the eight "value > 0" bits from the packed words: one packed
                // compare per word (0xFFFF per positive half; an activation of exactly 0 -- every pixel of an all-zero
                // tile at zero bias, RoiBuilder.py:234-236 -- is NOT positive, LeakyReLU'(0) = slope as in ATen), the
                // top bytes of the four masks gathered by one PRMT per pair, compressed by one multiply.
                uint4 pk;
                __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int i = 0; i < 4; ++i) hp[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                const uint32_t bits = mil_positive_bits(pk);
                wmask[(c >> 2) - W0] |= bits << ((c & 3) * 8);
                *reinterpret_cast<uint4*>(po + KOFF(c)) = pk;
              } else if (c < cbh || cbh == 0 || second_ok) {
                mil_store8(po + KOFF(c), v);
              }
            }
          }
          if (do_lrelu && mask_out != nullptr) {
            uint32_t* pm = mask_out + go.G + qo;
#pragma unroll
            for (int w = W0; w < W1; ++w)
              if (w * 4 < cbout) pm[(long long)w * go.PS] = wmask[w - W0];
          }
        } else if (in_range) {  // pad pixel of the output map: keep the zero row / column zero
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            const int c = C0 + k;
            if (c < cbout) *reinterpret_cast<uint4*>(po + KOFF(c)) = make_uint4(0, 0, 0, 0);
          }
        }
      }
    }
    };  // run
    if constexpr (CSPLIT == 1) {
      run(TcInt<0>{}, TcInt<MAXCB>{});
    } else {
      if (csub == 0) run(TcInt<0>{}, TcInt<TcSplit<MAXCB, MODE>::at>{});
      else run(TcInt<TcSplit<MAXCB, MODE>::at>{}, TcInt<MAXCB>{});
    }
  }
#undef KOFF
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
// stride 2 is supported for the FORWARD direction only (full-resolution evaluation + subsampled store); the
// stride-2 data gradient is the stride-1 data gradient of the zero-stuffed output gradient (mil_launch_upsample2)
bool mil_tc_supported(int dtype, int ks, int stride, int cin, int cout) {
  return dtype == MIL_BF16 && (ks == 3 || ks == 1) && (stride == 1 || stride == 2) && cin >= 9 && cin <= 80 &&
         cout <= 80;
}

int mil_tc_shape(int cin, int cout, int ks, MilTcShape* out) {
  MilTcShape& sh = *out;
  MIL_REQUIRE(ks == 1 || ks == 3 || ks == 7, "conv_tc: unsupported window %d", ks);
  sh.ks = ks;
  sh.ntaps = 0;
  if (ks == 7) {  // stem in space-to-depth form: 4x4 window, offsets -2..1 (mil_stem_tc.cu)
    for (int a = -2; a <= 1; ++a)
      for (int b = -2; b <= 1; ++b) { sh.t_dy[sh.ntaps] = (signed char)a; sh.t_dx[sh.ntaps] = (signed char)b; ++sh.ntaps; }
  } else {
    const int r = ks / 2;
    for (int a = -r; a <= r; ++a)
      for (int b = -r; b <= r; ++b) { sh.t_dy[sh.ntaps] = (signed char)a; sh.t_dx[sh.ntaps] = (signed char)b; ++sh.ntaps; }
  }
  sh.cbin = (cin + 7) / 8;
  sh.cbout = (cout + 7) / 8;
  sh.npad = (cout + 15) / 16 * 16;
  const int ng = sh.ntaps * sh.cbin;
  sh.nmma = (ng + 1) / 2;
  MIL_REQUIRE(sh.nmma <= MIL_TC_MAX_MMA, "conv_tc: too many K groups (%d)", ng);
  for (int g = 0; g < 2 * sh.nmma; ++g) {
    sh.g_tap[g] = g < ng ? (unsigned char)(g / sh.cbin) : 0xFF;
    sh.g_chunk[g] = g < ng ? (unsigned char)(g % sh.cbin) : 0xFF;
  }
  // a pair that straddles two taps would have its second half at a LOWER shared-memory address than its first
  // (chunk cbin-1 -> chunk 0): swap the halves so that the descriptor's (unsigned) LBO stays positive
  for (int j = 0; j < sh.nmma; ++j) {
    const int a = 2 * j, b = 2 * j + 1;
    if (sh.g_tap[b] != 0xFF && sh.g_tap[a] != sh.g_tap[b]) {
      std::swap(sh.g_tap[a], sh.g_tap[b]);
      std::swap(sh.g_chunk[a], sh.g_chunk[b]);
    }
  }
  return 0;
}

int mil_tc_shape_fold(int cmain, int cproj, int cout, MilTcShape* out) {
  MIL_TRY(mil_tc_shape(cmain, cout, 3, out));
  MilTcShape& sh = *out;
  const int cbm = sh.cbin, cbp = (cproj + 7) / 8;
  int ng = 9 * cbm;
  // undo the pair ordering of mil_tc_shape, append the projection's groups (centre tap = 4), then order every pair by
  // shared-memory address again: (chunk plane, shift) ascending, so that the descriptor's unsigned LBO stays positive
  for (int g = 0; g < ng; ++g) { sh.g_tap[g] = (unsigned char)(g / cbm); sh.g_chunk[g] = (unsigned char)(g % cbm); }
  for (int c = 0; c < cbp; ++c, ++ng) {
    MIL_REQUIRE(ng < 2 * MIL_TC_MAX_MMA, "conv_tc: too many K groups (folded projection)");
    sh.g_tap[ng] = 4;
    sh.g_chunk[ng] = (unsigned char)(cbm + c);
  }
  sh.cbin = cbm + cbp;
  sh.nmma = (ng + 1) / 2;
  MIL_REQUIRE(sh.nmma <= MIL_TC_MAX_MMA, "conv_tc: too many K groups (%d)", ng);
  for (int g = ng; g < 2 * MIL_TC_MAX_MMA; ++g) sh.g_tap[g] = sh.g_chunk[g] = 0xFF;
  for (int j = 0; j < sh.nmma; ++j) {
    const int a = 2 * j, b = 2 * j + 1;
    if (sh.g_tap[b] == 0xFF) continue;
    const bool swap = sh.g_chunk[a] > sh.g_chunk[b] || (sh.g_chunk[a] == sh.g_chunk[b] && sh.g_tap[a] > sh.g_tap[b]);
    if (swap) {
      std::swap(sh.g_tap[a], sh.g_tap[b]);
      std::swap(sh.g_chunk[a], sh.g_chunk[b]);
    }
  }
  return 0;
}

int mil_tc_shape_s2(int cin, int cout, MilTcShape* out) {
  MilTcShape& sh = *out;
  const int cb = (cin + 7) / 8;
  sh.ks = 2;
  sh.ntaps = 4;
  for (int t = 0; t < 4; ++t) { sh.t_dy[t] = (signed char)((t >> 1) - 1); sh.t_dx[t] = (signed char)((t & 1) - 1); }
  sh.cbin = 4 * cb;
  sh.cbout = (cout + 7) / 8;
  sh.npad = (cout + 15) / 16 * 16;
  // K-groups sorted by (plane, shift): any two consecutive groups then sit at ascending shared-memory offsets
  int ng = 0;
  for (int phase = 0; phase < 4; ++phase)
    for (int c = 0; c < cb; ++c)
      for (int t = 0; t < 4; ++t) {
        const int dy = (t >> 1) - 1, dx = (t & 1) - 1;
        if ((dy == -1 && !(phase >> 1)) || (dx == -1 && !(phase & 1))) continue;  // that row / column is not read
        MIL_REQUIRE(ng < 2 * MIL_TC_MAX_MMA, "conv_tc: too many K groups");
        sh.g_tap[ng] = (unsigned char)t;
        sh.g_chunk[ng] = (unsigned char)(phase * cb + c);
        ++ng;
      }
  sh.nmma = (ng + 1) / 2;
  MIL_REQUIRE(sh.nmma <= MIL_TC_MAX_MMA, "conv_tc: too many K groups (%d)", ng);
  for (int g = ng; g < 2 * MIL_TC_MAX_MMA; ++g) sh.g_tap[g] = sh.g_chunk[g] = 0xFF;
  return 0;
}

int mil_tc_shape_s2_dgrad(int cout_conv, int cin_conv, int a, MilTcShape* out) {
  // dx[2Y+a][2X+b] = sum over (ky, kx) with ky = a + 1 (mod 2), kx = b + 1 (mod 2) of
  //   W[ky][kx]^T dz[Y + dY][X + dX],  dY = 1 iff (a == 1 and ky == 0), dX = 1 iff (b == 1 and kx == 0).
  // The kernel's data-gradient convention is out(q) = sum_t W_t^T in(q - shift_t): t_dy = -dY, t_dx = -dX.
  // Taps in ascending (dY, dX) so that the K-group pairs ascend in shared memory.
  MilTcShape& sh = *out;
  sh.ks = 2;
  sh.ntaps = 0;
  for (int dY = 0; dY <= a; ++dY)
    for (int dX = 0; dX <= 1; ++dX) {
      sh.t_dy[sh.ntaps] = (signed char)-dY;
      sh.t_dx[sh.ntaps] = (signed char)-dX;
      ++sh.ntaps;
    }
  for (int t = sh.ntaps; t < MIL_TC_MAX_TAPS; ++t) sh.t_dy[t] = sh.t_dx[t] = 0;
  const int cbh = (cin_conv + 7) / 8;
  sh.cbin = (cout_conv + 7) / 8;  // the kernel reads dz (the conv's output channels) ...
  sh.cbout = 2 * cbh;             // ... and produces (column parity, the conv's input channels)
  sh.npad = (2 * cbh * 8 + 15) / 16 * 16;
  int ng = 0;
  for (int c = 0; c < sh.cbin; ++c)
    for (int t = 0; t < sh.ntaps; ++t) { sh.g_tap[ng] = (unsigned char)t; sh.g_chunk[ng] = (unsigned char)c; ++ng; }
  sh.nmma = (ng + 1) / 2;
  MIL_REQUIRE(sh.nmma <= MIL_TC_MAX_MMA && sh.cbout <= 10, "conv_tc: stride-2 data gradient too wide (%d groups)", ng);
  for (int g = ng; g < 2 * MIL_TC_MAX_MMA; ++g) sh.g_tap[g] = sh.g_chunk[g] = 0xFF;
  return 0;
}

size_t mil_tc_wpack_bytes(const MilTcShape& sh) { return (size_t)sh.nmma * 2 * sh.npad * 16; }

int mil_launch_pack_tc(const float* wp, void* wtc, const MilTcShape& sh, cudaStream_t s) {
  const int total = sh.nmma * 2 * sh.npad * 8;
  pack_tc_kernel<<<(int)mil_cdiv(total, 256), 256, 0, s>>>(wp, (__nv_bfloat16*)wtc, sh);
  MIL_LAUNCH_OK();
  return 0;
}

int mil_launch_pack_tc_table(const MilTcPackJob* jobs, int count, cudaStream_t s) {
  for (int base = 0; base < count; base += MIL_TC_PACK_MAX) {
    TcPackTable t;
    const int m = std::min(MIL_TC_PACK_MAX, count - base);
    for (int i = 0; i < m; ++i) {
      const MilTcPackJob& j = jobs[base + i];
      MilTcShape sh;
      if (j.w2 != nullptr) MIL_TRY(mil_tc_shape_fold(j.cin, j.cin2, j.cout, &sh));
      else if (j.s2 == 1) MIL_TRY(mil_tc_shape_s2(j.cin, j.cout, &sh));
      else if (j.s2 >= 2) MIL_TRY(mil_tc_shape_s2_dgrad(j.cout, j.cin, j.s2 - 2, &sh));
      else MIL_TRY(mil_tc_shape(j.transposed ? j.cout : j.cin, j.transposed ? j.cin : j.cout, j.ks, &sh));
      TcPackEntry& e = t.e[i];
      e.s2cb = j.s2 == 1 ? (unsigned char)((j.cin + 7) / 8) : 0;
      e.pair_cbh = j.s2 >= 2 ? (unsigned char)((j.cin + 7) / 8) : 0;
      e.pair_a = j.s2 >= 2 ? (unsigned char)(j.s2 - 2) : 0;
      for (int q = 0; q < MIL_TC_MAX_TAPS; ++q) { e.t_dy[q] = sh.t_dy[q]; e.t_dx[q] = sh.t_dx[q]; }
      e.w = j.w;
      e.w2 = j.w2;
      e.cin2 = (short)j.cin2;
      e.fold_cbm = j.w2 != nullptr ? (unsigned char)((j.cin + 7) / 8) : 0;
      e.wtc = (__nv_bfloat16*)j.wtc;
      e.cout = (short)j.cout; e.cin = (short)j.cin;
      e.taps = (unsigned char)(j.ks * j.ks); e.transposed = j.transposed ? 1 : 0;
      e.nmma = (unsigned char)sh.nmma; e.npad = (unsigned char)sh.npad;
      for (int g = 0; g < 2 * MIL_TC_MAX_MMA; ++g) { e.g_tap[g] = sh.g_tap[g]; e.g_chunk[g] = sh.g_chunk[g]; }
    }
    pack_tc_table_kernel<<<dim3(m, 8), 256, 0, s>>>(t);
    MIL_LAUNCH_OK();
  }
  return 0;
}

static size_t tc_smem_bytes(int halo, const MilTcShape& sh, int n_stages) {
  const size_t hdr = (sizeof(TcSmemHeader) + 127) / 128 * 128;
  const size_t b = ((size_t)sh.nmma * 2 * sh.npad * 16 + 127) / 128 * 128;
  const size_t span = TC_M + 2 * (size_t)halo;
  return hdr + b + (size_t)n_stages * span * 16 * (sh.cbin + 1);
}

bool mil_conv_tc_fits(const MilTcShape& sh, int wp) {
  const int ng = sh.cbout <= 5 ? 4 : 2;
  return tc_smem_bytes(mil_tc_halo(sh, wp), sh, ng) <= 227 * 1024;
}

int mil_tc_build_issue(const MilTcShape& sh, int wp, int halo, int transposed, TcIssue* out) {
  TcIssue& iss = *out;
  const int plane = (TC_M + 2 * halo) * 16;
  for (int j = 0; j < sh.nmma; ++j) {
    int off[2];
    for (int h = 0; h < 2; ++h) {
      const int tap = sh.g_tap[2 * j + h], chunk = sh.g_chunk[2 * j + h];
      if (tap == 0xFF) {
        off[h] = -1;
      } else {
        int sft = sh.t_dy[tap] * wp + sh.t_dx[tap];  // forward reads x(q + s); the data gradient reads dz(q - s)
        if (transposed) sft = -sft;
        off[h] = chunk * plane + (halo + sft) * 16;
      }
    }
    if (off[1] < 0) off[1] = sh.cbin * plane + halo * 16;  // dummy half -> the all-zero plane
    MIL_REQUIRE(off[1] > off[0], "conv_tc: internal error, K-halves out of order");
    iss.a_desc[j] = make_desc_bits((uint32_t)off[0], (uint32_t)(off[1] - off[0]), 128);
    iss.b_desc[j] = make_desc_bits((uint32_t)j * 2 * sh.npad * 16, (uint32_t)sh.npad * 16, 128);
  }
  for (int j = sh.nmma; j < MIL_TC_MAX_MMA; ++j) iss.a_desc[j] = iss.b_desc[j] = 0;
  return 0;
}

int mil_launch_conv_tc(int transposed, const void* x, const MilPF8& gx, const void* wtc, const MilTcShape& sh,
                       const float* bias, const void* res, const void* act, void* out, const MilPF8& go, int epi,
                       int sub, cudaStream_t s, const MilPF8* gres_half, int up_row, const void* mask_in,
                       void* mask_out) {
  if (up_row >= 0)
    MIL_REQUIRE(transposed && !sub && gx.n == go.n && gx.h == (go.h - 1) / 2 + 1 && gx.w == (go.w - 1) / 2 + 1 &&
                    (res == nullptr || gres_half != nullptr) && sh.cbout == 2 * go.cb,
                "conv_tc: stride-2 data-gradient geometry mismatch");
  if (gres_half != nullptr)
    MIL_REQUIRE(!sub && res != nullptr && gres_half->cb == go.cb && gres_half->h == (go.h - 1) / 2 + 1 &&
                    gres_half->w == (go.w - 1) / 2 + 1,
                "conv_tc: half-resolution residual geometry mismatch");
  if (sub)
    MIL_REQUIRE(gx.n == go.n && go.h == (gx.h - 1) / 2 + 1 && go.w == (gx.w - 1) / 2 + 1 && !transposed,
                "conv_tc: stride-2 geometry mismatch");
  else if (up_row < 0)
    MIL_REQUIRE(gx.n == go.n && gx.h == go.h && gx.w == go.w && gx.wp == go.wp && gx.hp == go.hp,
                "conv_tc: geometry mismatch");
  MIL_REQUIRE(gx.cb == sh.cbin && (up_row >= 0 || go.cb == sh.cbout), "conv_tc: channel chunks do not match the packed weights");
  MIL_REQUIRE(epi != MIL_EPI_DGRAD || act != nullptr, "conv_tc: DGRAD epilogue needs the activation tensor");
  const int halo = mil_tc_halo(sh, gx.wp);
  MIL_REQUIRE(halo <= gx.G, "conv_tc: the window reaches %d pixels back but the map's guard is %lld", halo, gx.G);
  // ring depth: as deep as shared memory allows, and a MULTIPLE of the accumulator ring (the one `empty[stage]`
  // barrier per tile is waited on by the producer and by the epilogue group of that accumulator: with
  // n_stages % NG == 0 the next completion of a stage's barrier needs this epilogue group to have moved on)
  const int ng = sh.cbout <= 5 ? 4 : 2;
  int n_stages = TC_MAX_STAGES;
  while (n_stages > ng && tc_smem_bytes(halo, sh, n_stages) > 200 * 1024) n_stages -= ng;
  const size_t smem = tc_smem_bytes(halo, sh, n_stages);
  MIL_REQUIRE(smem <= 227 * 1024, "conv_tc: tile width %d needs %zu bytes of shared memory", gx.w, smem);
  // the second K-half must sit above the first one (see mil_tc_shape); holds whenever a plane is larger than
  // twice the largest tap shift, i.e. always for cbin >= 2; cbin == 1 would need taps in ascending shift order
  MIL_REQUIRE(sh.cbin >= 2, "conv_tc: needs at least 9 input channels");
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    MIL_CHECK_CUDA(cudaGetDevice(&dev));
    MIL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  TcIssue iss;
  MIL_TRY(mil_tc_build_issue(sh, gx.wp, halo, transposed, &iss));
  const long long n_tiles = mil_cdiv(gx.Q, TC_M);
  const int grid = (int)std::min<long long>(n_tiles, n_sm);
#define MIL_TC_LAUNCH1(MAXCB, NG, MODE)                                                                           \
  do {                                                                                                            \
    MIL_SET_SMEM((conv_tc_kernel<MAXCB, NG, MODE>), smem);                                                        \
    MIL_CHECK_CUDA(mil_launch_pdl((conv_tc_kernel<MAXCB, NG, MODE>), dim3(grid),                                  \
                                  dim3(96 + NG * 128 * TcSplit<MAXCB, MODE>::value), smem, s,                     \
        (const __nv_bfloat16*)x, gx, (const __nv_bfloat16*)wtc, bias, (const __nv_bfloat16*)res,                  \
        (const __nv_bfloat16*)act, (__nv_bfloat16*)out, go, sh, iss, epi, sub, halo, n_stages,                    \
        gres_half ? *gres_half : go, gres_half ? 1 : 0, up_row + 1, up_row >= 0 ? go.cb : 0,                      \
        (const uint32_t*)mask_in, (uint32_t*)mask_out));                                                          \
  } while (0)
  // the network's layers (3 / 5 / 8 / 10 output chunks, five epilogue kinds) run specialised instantiations
#define MIL_TC_LAUNCH(MAXCB, NG)                                                                                  \
  do {                                                                                                            \
    const int mode = sh.cbout != MAXCB ? TCM_GENERIC                                                              \
                     : (epi == MIL_EPI_PLAIN && bias == nullptr && res == nullptr) ? TCM_PLAIN                    \
                     : epi == MIL_EPI_FWD ? (res ? TCM_FWD_RES : TCM_FWD)                                          \
                     : epi == MIL_EPI_DGRAD ? (res ? TCM_DGRAD_RES : TCM_DGRAD) : TCM_GENERIC;                     \
    switch (mode) {                                                                                               \
      case TCM_PLAIN: MIL_TC_LAUNCH1(MAXCB, NG, TCM_PLAIN); break;                                                \
      case TCM_FWD: MIL_TC_LAUNCH1(MAXCB, NG, TCM_FWD); break;                                                    \
      case TCM_FWD_RES: MIL_TC_LAUNCH1(MAXCB, NG, TCM_FWD_RES); break;                                            \
      case TCM_DGRAD: MIL_TC_LAUNCH1(MAXCB, NG, TCM_DGRAD); break;                                                \
      case TCM_DGRAD_RES: MIL_TC_LAUNCH1(MAXCB, NG, TCM_DGRAD_RES); break;                                        \
      default: MIL_TC_LAUNCH1(MAXCB, NG, TCM_GENERIC); break;                                                     \
    }                                                                                                             \
  } while (0)
  if (sh.cbout <= 3) MIL_TC_LAUNCH(3, 4);
  else if (sh.cbout <= 5) MIL_TC_LAUNCH(5, 4);
  else if (sh.cbout == 6) MIL_TC_LAUNCH(6, 2);
  else if (sh.cbout <= 8) MIL_TC_LAUNCH(8, 2);
  else MIL_TC_LAUNCH(10, 2);
#undef MIL_TC_LAUNCH
#undef MIL_TC_LAUNCH1
  MIL_LAUNCH_OK();
  return 0;
}
