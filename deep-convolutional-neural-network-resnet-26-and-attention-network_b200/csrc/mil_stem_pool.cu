// Stem convolution with the max-pool FUSED into its epilogue (bf16 mode, even conv-map size).
// Reference: gbm/model.py:24-26,51-53 -- conv 7x7 / stride 2 + bias, LeakyReLU(0.1), max-pool 3x3 / stride 2 / pad 1.
//
// In the space-to-depth-by-4 form (mil_stem_tc.cu) the conv map and the pooled map share ONE pixel geometry: conv
// pixel q holds the four conv phases under pooled pixel q, and pooled pixel q is the maximum over (neighbour, phase)
// pairs of the conv pixels q, q-1, q-wp, q-wp-1.  The un-fused path writes the 80-channel conv map to HBM (520 KB per
// tile, the largest tensor of the whole network) only for the pool kernel to read it back.  Here the conv map never
// leaves the SM:
//   * every CTA takes a CONTIGUOUS range of 128-pixel tiles (plus one warm-up tile in front of it, whose conv output
//     is computed but whose pooled pixels belong to the previous CTA),
//   * the epilogue of tile t packs its 80 conv values per pixel to bf16 and parks them in a two-tile ring in shared
//     memory; a pooled pixel then finds its three other neighbours in the ring (this tile or the previous one),
//   * the two epilogue groups (alternate tiles) hand the ring slots to each other through two mbarrier pairs:
//       cv_full[s]  "tile in slot s is written"      (waited by the pool phase of the NEXT tile, other group)
//       cv_free[s]  "the next tile has read slot s"  (waited before slot s is overwritten two tiles later)
// Outputs: pooled map (PF8, 20 channels), arg-max (same layout as stem_pool4_kernel), sign mask of the pooled map.
// Pipeline roles as in conv_tc_kernel: warp 0 bulk-TMA producer, warps 1-2 MMA issuers, 2 x 4 epilogue warps.
#include <algorithm>

#include "mil_common.cuh"
#include "mil_conv_tc.cuh"
#include "mil_stem_unpool.cuh"
#include "mil_tc_ptx.cuh"

#define SP_M 128
#define SP_NG 2
#define SP_CB 10  // conv output chunks: 80 = 20 channels x 4 phases; chunk cp = channel pair (2cp, 2cp+1)
#define SP_MAX_STAGES 4
#define SP_RING_BYTES (2 * SP_CB * SP_M * 16)

struct SpSmemHeader {
  uint64_t full[SP_MAX_STAGES], empty[SP_MAX_STAGES], acc_empty[SP_NG], b_full, cv_full[2], cv_free[2];
  uint32_t tmem_base;
  alignas(16) float bias[80];
};

// both channels of a pair at once on packed bf16x2 words, window positions in ATen's scan order (first maximum wins):
// same routine as in mil_stem_tc.cu (kept local: both kernels inline it)
__device__ __forceinline__ uint32_t sp_pool9_pair(const uint4& UL, const uint4& U, const uint4& L, const uint4& S,
                                                  uint32_t& am_pair) {
  auto lo = [](uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x5410); };
  auto hi = [](uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7632); };
  auto as2 = [](uint32_t w) { return *reinterpret_cast<const __nv_bfloat162*>(&w); };
  uint32_t best = hi(UL.y, UL.w), idx = 0;
  auto step = [&](uint32_t cand, uint32_t pos) {
    const uint32_t m = __hgt2_mask(as2(cand), as2(best));
    const __nv_bfloat162 mx = __hmax2(as2(cand), as2(best));
    best = *reinterpret_cast<const uint32_t*>(&mx);
    idx = (idx & ~m) | ((pos | (pos << 16)) & m);
  };
  step(lo(U.y, U.w), 1);
  step(hi(U.y, U.w), 2);
  step(hi(L.x, L.z), 3);
  step(lo(S.x, S.z), 4);
  step(hi(S.x, S.z), 5);
  step(hi(L.y, L.w), 6);
  step(lo(S.y, S.w), 7);
  step(hi(S.y, S.w), 8);
  am_pair = (idx | (idx >> 8)) & 0xFFFFu;
  return best;
}

__global__ void __launch_bounds__(96 + SP_NG * 128, 1)
stem_conv_pool_kernel(const __nv_bfloat16* __restrict__ x, MilPF8 gx, const __nv_bfloat16* __restrict__ wtc,
                      const float* __restrict__ bias4, __nv_bfloat16* __restrict__ pooled, MilPF8 gp,
                      uint2* __restrict__ argmax, uint32_t* __restrict__ mask_out, MilTcShape sh,
                      const __grid_constant__ TcIssue iss, int halo, int n_stages, long long chunk) {
  extern __shared__ __align__(128) unsigned char smem[];
  SpSmemHeader* hd = reinterpret_cast<SpSmemHeader*>(smem);
  const uint32_t hdr_bytes = (uint32_t)((sizeof(SpSmemHeader) + 127) / 128 * 128);
  unsigned char* bsm = smem + hdr_bytes;
  const uint32_t b_bytes = (uint32_t)sh.nmma * 2 * sh.npad * 16;
  const int span = SP_M + 2 * halo;
  const uint32_t plane = (uint32_t)span * 16;
  const uint32_t stage_bytes = plane * (sh.cbin + 1);  // + one all-zero plane (odd K-group count)
  unsigned char* asm0 = bsm + ((b_bytes + 127) / 128 * 128);
  uint4* ring = reinterpret_cast<uint4*>(asm0 + (size_t)n_stages * stage_bytes);  // [slot 2][chunk 10][pixel 128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = mil_cdiv(gx.Q, SP_M);
  const uint32_t acc_stride = (uint32_t)((sh.npad + 31) / 32 * 32);
  // this CTA's tiles: [t_lo, t_hi) produce pooled pixels; one warm-up tile in front (its conv output only)
  const long long t_lo = (long long)blockIdx.x * chunk, t_hi = min(n_tiles, t_lo + chunk);
  const bool warm = blockIdx.x > 0 && t_lo < t_hi;
  const long long t_begin = t_lo - (warm ? 1 : 0);
  const long long nloc = t_hi > t_begin ? t_hi - t_begin : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) { mbar_init(&hd->full[s], 1); mbar_init(&hd->empty[s], 1); }
    for (int a = 0; a < SP_NG; ++a) mbar_init(&hd->acc_empty[a], 4);
    for (int s = 0; s < 2; ++s) { mbar_init(&hd->cv_full[s], 128); mbar_init(&hd->cv_free[s], 128); }
    mbar_init(&hd->b_full, 1);
    fence_barrier_init();
  }
  for (int s = 0; s < n_stages; ++s) {
    uint4* zp = reinterpret_cast<uint4*>(asm0 + (size_t)s * stage_bytes + (size_t)sh.cbin * plane);
    for (int i = threadIdx.x; i < span; i += blockDim.x) zp[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(&hd->tmem_base, 256);
  mil_pdl_wait();   // everything above overlaps the previous kernel's tail; global memory only from here on (PDL, mil_common.cuh)
  for (int i = threadIdx.x; i < 80; i += blockDim.x) hd->bias[i] = bias4[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hd->tmem_base;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&hd->b_full, b_bytes);
      bulk_g2s(bsm, wtc, b_bytes, &hd->b_full);
    }
    __syncwarp();
    // running pointers advanced by constants: the producer's instruction stream is on the critical path (conv_tc_kernel)
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t cbin = (uint32_t)sh.cbin, tx_bytes = plane * cbin;
    const uint32_t a0 = smem_u32(asm0), full0 = smem_u32(&hd->full[0]);
    const char* src = reinterpret_cast<const char*>(x) + (gx.G - halo + t_begin * SP_M) * 16;  // chunk 0
    const long long cstride = gx.PS * 16;
    for (long long it = 0; it < nloc; ++it) {
      mbar_wait(&hd->empty[stage], phase ^ 1);
      if (elect_one()) {
        const uint32_t bar = full0 + (uint32_t)stage * 8;
        mbar_expect_tx_u32(bar, tx_bytes);
        uint32_t dst = a0 + (uint32_t)stage * stage_bytes;
        const char* sp = src;
        for (uint32_t c = 0; c < cbin; ++c, dst += plane, sp += cstride) bulk_g2s_u32(dst, sp, plane, bar);
      }
      __syncwarp();
      src += SP_M * 16;
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    mil_pdl_trigger();  // all of this CTA's loads are issued: the next kernel may start its prologue (PDL, mil_common.cuh)
  } else if (warp <= 2) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(sh.npad >> 3) << 17) |
                           ((uint32_t)(SP_M >> 4) << 24);
    mbar_wait(&hd->b_full, 0);
    const int nmma = sh.nmma;
    const uint64_t b_add = (uint64_t)(smem_u32(bsm) >> 4);
    // ring positions + phase bits instead of 64-bit divisions (the issuing warps' bookkeeping is on the critical path,
    // see conv_tc_kernel); n_stages is a multiple of SP_NG, both are even
    int stage = warp - 1, acc = warp - 1;
    uint32_t full_par = 0, acc_par = 1;
    const uint32_t a_base = smem_u32(asm0) >> 4, a_step = stage_bytes >> 4;
    for (long long it = warp - 1; it < nloc; it += 2) {
      mbar_wait(&hd->acc_empty[acc], acc_par);
      mbar_wait(&hd->full[stage], full_par);
      tc_fence_after();
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
      if (acc >= SP_NG) { acc -= SP_NG; acc_par ^= 1; }
    }
  } else {
    // ===================== epilogue + pool: group eg takes the local tiles eg, eg + 2, ... =====================
    const int eg = (warp - 3) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // pixel of the tile = TMEM lane
    const int wp = (int)gp.wp;
    const uint32_t NINF2 = 0xFF80FF80u;
    int estage = eg;
    uint32_t epar = 0;
    for (long long it = eg; it < nloc; it += SP_NG) {
      const long long q = (t_begin + it) * SP_M + row;
      const int n = (int)(q / gp.P);
      const int r = (int)(q - (long long)n * gp.P);
      const int y = r / wp, xo = r - y * wp;
      const bool in_range = n < gp.n;
      const bool live = in_range && y < gp.h && xo < gp.w;
      const int slot = (int)(it & 1);
      // 1. accumulator row out of TMEM, stage handed back at once
      mbar_wait(&hd->empty[estage], epar);
      tc_fence_after();
      estage += SP_NG;
      if (estage >= n_stages) { estage -= n_stages; epar ^= 1; }
      const uint32_t taddr = tmem_base + eg * acc_stride + ((uint32_t)(quarter * 32) << 16);
      float acc[SP_CB][8];
#pragma unroll
      for (int c = 0; c < SP_CB; ++c) tmem_ld8(taddr + c * 8, acc[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hd->acc_empty[eg]);
      // 2. bias + LeakyReLU, packed to bf16: chunk cp = (channel 2cp: phases 00 01 10 11 | channel 2cp+1: ...)
      uint4 pk[SP_CB];
#pragma unroll
      for (int c = 0; c < SP_CB; ++c) {
        const float4 b0 = *reinterpret_cast<const float4*>(&hd->bias[c * 8]);
        const float4 b1 = *reinterpret_cast<const float4*>(&hd->bias[c * 8 + 4]);
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk[c]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v0 = acc[c][2 * i] + bv[2 * i], v1 = acc[c][2 * i + 1] + bv[2 * i + 1];
          v0 = fmaxf(v0, MIL_SLOPE * v0);
          v1 = fmaxf(v1, MIL_SLOPE * v1);
          hp[i] = __floats2bfloat162_rn(v0, v1);
        }
      }
      // 3. park the tile in the ring (slot `slot` is free once the pool phase of tile it-1 has read tile it-2)
      if (it >= 2) mbar_wait(&hd->cv_free[slot], (uint32_t)(((it >> 1) - 1) & 1));
      uint4* mine = ring + (size_t)slot * SP_CB * SP_M;
#pragma unroll
      for (int c = 0; c < SP_CB; ++c) mine[c * SP_M + row] = pk[c];
      mbar_arrive(&hd->cv_full[slot]);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");  // this group's own writes
      if (warm && it == 0) continue;  // warm-up tile: its pooled pixels belong to the previous CTA
      // 4. pool: neighbours q-1, q-wp, q-wp-1 from the ring (this tile, or the previous one in the other slot)
      if (it >= 1) mbar_wait(&hd->cv_full[slot ^ 1], (uint32_t)(((it - 1) >> 1) & 1));
      if (live) {
        const uint4* prev = ring + (size_t)(slot ^ 1) * SP_CB * SP_M;
        const int jl = row - 1, ju = row - wp, jul = row - wp - 1;
        const uint4* pl = jl >= 0 ? mine + jl : prev + (SP_M + jl);
        const uint4* pu = ju >= 0 ? mine + ju : prev + (SP_M + ju);
        const uint4* pul = jul >= 0 ? mine + jul : prev + (SP_M + jul);
        const bool up_ok = y > 0, left_ok = xo > 0;
        uint32_t outw[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) outw[i] = 0u;
        uint32_t amw[12];  // arg-max codes of the channel pairs (mil_stem_unpool.cuh)
#pragma unroll
        for (int i = 0; i < 12; ++i) amw[i] = MIL_AM_CODE;
#pragma unroll
        for (int cp = 0; cp < SP_CB; ++cp) {
          uint4 L = pl[cp * SP_M], U = pu[cp * SP_M], UL = pul[cp * SP_M];
          if (!left_ok) { L.x = L.y = L.z = L.w = NINF2; }
          if (!up_ok) { U.y = U.w = NINF2; }
          if (!(up_ok && left_ok)) { UL.y = UL.w = NINF2; }
          uint32_t amp;
          outw[cp] = sp_pool9_pair(UL, U, L, pk[cp], amp);
          amw[cp] = amp | MIL_AM_CODE;
        }
        // pooled channel 2cp + {0,1} -> chunk cp / 4, word cp % 4 (channels 20..23 stay zero)
        const long long qo = (long long)n * gp.P + r;
        if (argmax != nullptr)  // (forward-only runs keep no arg-max records)
#pragma unroll
          for (int pc = 0; pc < 3; ++pc)
            argmax[(long long)pc * gp.PS + gp.G + qo] =
                make_uint2(amw[4 * pc] | (amw[4 * pc + 1] << 16), amw[4 * pc + 2] | (amw[4 * pc + 3] << 16));
#pragma unroll
        for (int pc = 0; pc < 3; ++pc)
          *reinterpret_cast<uint4*>(pooled + mil_pf8_off(gp, pc, qo)) =
              make_uint4(outw[4 * pc], outw[4 * pc + 1], outw[4 * pc + 2], outw[4 * pc + 3]);
        if (mask_out != nullptr) {
          // sign mask of the pooled map (see conv_tc_kernel): byte = chunk, bit = channel
          uint32_t m = 0;
#pragma unroll
          for (int pc = 0; pc < 3; ++pc) {
            m |= mil_positive_bits(make_uint4(outw[4 * pc], outw[4 * pc + 1], outw[4 * pc + 2], outw[4 * pc + 3])) << (8 * pc);
          }
          mask_out[gp.G + qo] = m;
        }
      } else if (in_range) {  // pad pixel of the pooled map
        const long long qo = (long long)n * gp.P + r;
#pragma unroll
        for (int pc = 0; pc < 3; ++pc)
          *reinterpret_cast<uint4*>(pooled + mil_pf8_off(gp, pc, qo)) = make_uint4(0, 0, 0, 0);
      }
      if (it >= 1) mbar_arrive(&hd->cv_free[slot ^ 1]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

static size_t sp_smem_bytes(int halo, const MilTcShape& sh, int n_stages) {
  const size_t hdr = (sizeof(SpSmemHeader) + 127) / 128 * 128;
  const size_t b = ((size_t)sh.nmma * 2 * sh.npad * 16 + 127) / 128 * 128;
  return hdr + b + (size_t)n_stages * (SP_M + 2 * (size_t)halo) * 16 * (sh.cbin + 1) + SP_RING_BYTES;
}

bool mil_stem_conv_pool_supported(const MilPF8& gp, int hc) {
  MilTcShape sh;
  if (mil_tc_shape(48, 80, 3, &sh) != 0) return false;
  return (hc & 1) == 0 && gp.wp + 1 <= SP_M && sp_smem_bytes((int)gp.wp + 1, sh, SP_NG) <= 227 * 1024;
}

int mil_launch_stem_conv_pool(const void* xs, const MilPF8& gi, const void* wtc, const float* bias4, void* pooled,
                              const MilPF8& gp, void* argmax, void* mask_out, int hc, cudaStream_t s) {
  MIL_REQUIRE(mil_stem_conv_pool_supported(gp, hc), "stem_conv_pool: unsupported geometry (conv size %d, row %lld)", hc, gp.wp);
  MIL_REQUIRE(gi.n == gp.n && gi.h == gp.h && gi.w == gp.w && gi.wp == gp.wp && gi.cb == 6 && gp.cb == 3,
              "stem_conv_pool: geometry mismatch");
  MilTcShape sh;
  MIL_TRY(mil_tc_shape(48, 80, 3, &sh));
  const int halo = mil_tc_halo(sh, gi.wp);
  MIL_REQUIRE(halo <= gi.G, "stem_conv_pool: window reaches %d pixels back but the guard is %lld", halo, gi.G);
  int n_stages = SP_MAX_STAGES;
  while (n_stages > SP_NG && sp_smem_bytes(halo, sh, n_stages) > 224 * 1024) n_stages -= SP_NG;
  const size_t smem = sp_smem_bytes(halo, sh, n_stages);
  TcIssue iss;
  MIL_TRY(mil_tc_build_issue(sh, gi.wp, halo, 0, &iss));
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    MIL_CHECK_CUDA(cudaGetDevice(&dev));
    MIL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long n_tiles = mil_cdiv(gi.Q, SP_M);
  const int grid = (int)std::min<long long>(n_tiles, n_sm);
  const long long chunk = mil_cdiv(n_tiles, grid);
  MIL_SET_SMEM(stem_conv_pool_kernel, smem);
  MIL_LAUNCH_PDL(stem_conv_pool_kernel, grid, 96 + SP_NG * 128, smem, s, (const __nv_bfloat16*)xs, gi, (const __nv_bfloat16*)wtc, bias4,
                                                          (__nv_bfloat16*)pooled, gp, (uint2*)argmax, (uint32_t*)mask_out, sh, iss,
                                                          halo, n_stages, chunk);
  return 0;
}
