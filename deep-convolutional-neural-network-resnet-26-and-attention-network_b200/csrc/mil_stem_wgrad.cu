// Stem backward on the tensor cores: weight + bias gradient of conv1 (gbm/model.py:24,51; the bag is detached, so this
// is the stem's whole backward pass, gbm/model.py:194,196) straight from the POOLED gradient -- the un-pooled gradient
// (80 channels at the phase-map resolution, the largest tensor of the backward pass) is built tile by tile inside the
// kernel and never touches HBM.
//
// In the space-to-depth-by-4 form (mil_stem_tc.cu) the weight gradient is that of a 3x3 / stride-1 convolution 48 -> 80:
//   dW4[dy][dx][ci][co] = sum over flat pixels q of  xs[q + dy * wp][ci] * dY4[q - dx][co]
// "Square" GEMM form (as wgrad_sq_kernel): the three dx shifts live on the dY4 side -- A slots (co chunk c, dx),
// 30 x 8 rows -> TWO M = 128 MMAs per 16-pixel K-step (15 slots each; 94 % of the rows are real) -- and the three dy
// shifts on the xs side -- B slots (ci chunk, dy) + a constant-one slot for the bias gradient, N = 160.  Both operands
// are MN-major core matrices of 8 pixels x 16 B: exactly how PF8 stores a pixel chunk.
//   * B: bulk-TMA, three row-shifted fetches of every xs chunk plane;
//   * A: 15 builder warps turn pooled gradient + arg-max records into dY4 chunks (mil_stem_unpool.cuh: ~35
//     instructions per chunk on packed bf16x2 lanes) for the tile's pixels +- ONE pixel (the dx shifts are on this
//     side so that the halo is a pixel, not a row), and store every chunk into its dx = -1 / 0 / +1 slots.
//     No index arithmetic: a pad pixel of the phase map comes out as zero by itself -- its own pooled gradient is a
//     zero pad pixel (PF8 invariant), and the neighbouring windows that could reach it would have to hold their
//     maximum in window row / column 0 of pooled row / column 0, which is the pool's -inf padding.
// Accumulators stay in TMEM for the whole kernel (2 x 160 columns); every CTA writes ONE partial record
// [tap 9][cin 48][cout 80] + [80], summed in fixed order by stem_reduce4_kernel -> deterministic.
//
// Warp roles: 0 = bulk-TMA producer, 1 = MMA issuer, 2..5 = epilogue, 6..20 = builders (5 warps per pooled chunk).
#include <algorithm>

#include "mil_common.cuh"
#include "mil_conv_tc.cuh"
#include "mil_stem_unpool.cuh"
#include "mil_tc_ptx.cuh"

#define SW_TK 128                   // pixels per K-tile
#define SW_PLANE (SW_TK * 16)       // bytes of one slot
#define SW_CA 10                    // dY4 chunks (80 channels = 20 x 4 phases)
#define SW_CB 6                     // xs chunks (48 channels = 3 x 4 x 4)
#define SW_A_SLOTS (3 * SW_CA)      // (chunk, dx)
#define SW_B_SLOTS (3 * SW_CB + 2)  // (chunk, dy) + ones + one zero slot: N = 160
#define SW_N (SW_B_SLOTS * 8)
#define SW_STAGE ((SW_A_SLOTS + SW_B_SLOTS) * SW_PLANE)
#define SW_STAGES 2
#define SW_BUILD_WARPS 15
#define SW_SPAN (SW_TK + 2)            // pixels a tile's builders produce: [q0 - 1, q0 + TK + 1)
#define SW_THREADS (192 + 32 * SW_BUILD_WARPS)

struct SwSmemHeader {
  uint64_t full[SW_STAGES], empty[SW_STAGES], a_ready[SW_STAGES], done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(SW_THREADS, 1)
stem_wgrad_kernel(const __nv_bfloat16* __restrict__ xs, MilPF8 gx, const __nv_bfloat16* __restrict__ g, MilPF8 gp,
                  const uint2* __restrict__ am, float* __restrict__ partial, long long rec_stride) {
  extern __shared__ __align__(128) unsigned char smem[];
  SwSmemHeader* hd = reinterpret_cast<SwSmemHeader*>(smem);
  unsigned char* stage0 = smem + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = mil_cdiv(gp.Q, SW_TK);
  const int wp = gp.wp;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SW_STAGES; ++s) {
      mbar_init(&hd->full[s], 1);
      mbar_init(&hd->empty[s], 1);
      mbar_init(&hd->a_ready[s], SW_BUILD_WARPS);
    }
    mbar_init(&hd->done, 1);
    fence_barrier_init();
  }
  for (int st = 0; st < SW_STAGES; ++st) {  // constant slots of every stage: bf16 1.0, then zeros
    uint32_t* op = reinterpret_cast<uint32_t*>(stage0 + (size_t)st * SW_STAGE + (size_t)(SW_A_SLOTS + 3 * SW_CB) * SW_PLANE);
    for (int i = threadIdx.x; i < SW_PLANE / 4; i += blockDim.x) {
      op[i] = 0x3F803F80u;
      op[i + SW_PLANE / 4] = 0u;
    }
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(&hd->tmem_base, 512);
  mil_pdl_wait();   // everything above overlaps the previous kernel's tail; global memory only from here on (PDL, mil_common.cuh)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hd->tmem_base;

  if (warp == 0) {
    // ---- producer: xs planes, slots (c, dy) = the plane read from pixel q0 + dy * wp ----
    // (running pointers advanced by constants: the 18 copies per tile are one serial instruction stream)
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t b0 = smem_u32(stage0) + SW_A_SLOTS * SW_PLANE, full0 = smem_u32(&hd->full[0]);
    const char* src = reinterpret_cast<const char*>(xs) + (gx.G - wp + (long long)blockIdx.x * SW_TK) * 16;  // (c = 0, dy = -1)
    const long long cstride = gx.PS * 16, tstride = (long long)gridDim.x * SW_TK * 16, rstride = (long long)wp * 16;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&hd->empty[stage], phase ^ 1);
      if (elect_one()) {
        const uint32_t bar = full0 + (uint32_t)stage * 8;
        mbar_expect_tx_u32(bar, 3 * SW_CB * SW_PLANE);
        uint32_t dst = b0 + (uint32_t)stage * SW_STAGE;
        const char* sp = src;
#pragma unroll
        for (int c = 0; c < SW_CB; ++c, sp += cstride) {
          bulk_g2s_u32(dst, sp, SW_PLANE, bar);
          bulk_g2s_u32(dst + SW_PLANE, sp + rstride, SW_PLANE, bar);
          bulk_g2s_u32(dst + 2 * SW_PLANE, sp + 2 * rstride, SW_PLANE, bar);
          dst += 3 * SW_PLANE;
        }
      }
      __syncwarp();
      src += tstride;
      if (++stage == SW_STAGES) { stage = 0; phase ^= 1; }
    }
    mil_pdl_trigger();  // all of this CTA's loads are issued: the next kernel may start its prologue (PDL, mil_common.cuh)
  } else if (warp == 1) {
    // ---- MMA issuer: D = f32, A = B = bf16, both MN-major, M = 128, N = 160.  Half h multiplies A slots 15 h ..
    // 15 h + 15 (its 16th slot is the other half's first one, or the first B slot: rows nobody reads) ----
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24) |
                           ((uint32_t)(SW_N >> 3) << 17);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&hd->full[stage], phase);
      mbar_wait(&hd->a_ready[stage], phase);
      tc_fence_after();
      const uint32_t a_base = smem_u32(stage0 + (size_t)stage * SW_STAGE);
      const uint64_t ad0 = make_desc(a_base, 128, SW_PLANE);
      const uint64_t ad1 = make_desc(a_base + 15 * SW_PLANE, 128, SW_PLANE);
      const uint64_t bd0 = make_desc(a_base + SW_A_SLOTS * SW_PLANE, 128, SW_PLANE);
      const uint32_t acc0 = first ? 0u : 1u;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < SW_TK / 16; ++kk) {
          umma_bf16(tmem_base, ad0 + kk * 16, bd0 + kk * 16, idesc, kk > 0 ? 1u : acc0);
          umma_bf16(tmem_base + SW_N, ad1 + kk * 16, bd0 + kk * 16, idesc, kk > 0 ? 1u : acc0);
        }
        umma_commit(&hd->empty[stage]);
      }
      __syncwarp();
      first = false;
      if (++stage == SW_STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(&hd->done);
    __syncwarp();
  } else if (warp >= 6) {
    // ---- builders: thread = (pixel j of the span [q0 - 1, q0 + TK + 1), pooled chunk pc); pc is warp-uniform ----
    const int bw = warp - 6;
    const int pc = bw / 5;                     // 5 warps (160 threads >= 130 span pixels) per pooled chunk
    const int j = (bw - pc * 5) * 32 + lane;
    const int npair = pc == 2 ? 2 : 4;         // channel pairs 8, 9 only in the last pooled chunk
    const uint4* gpl = reinterpret_cast<const uint4*>(g) + (long long)pc * gp.PS + gp.G;  // pooled-gradient chunk plane
    const uint2* apl = am + (long long)pc * gp.PS + gp.G;
    // chunk c = 4 pc + k goes to slot (c, d): d = 1 holds dY4[q0 + i], d = 0 dY4[q0 + i + 1], d = 2 dY4[q0 + i - 1]
    const int i0 = j - 2, i1 = j - 1, i2 = j;
    const bool st0 = i0 >= 0 && i0 < SW_TK, st1 = i1 >= 0 && i1 < SW_TK, st2 = i2 < SW_TK;
    int stage = 0;
    uint32_t phase = 0;
    // The pooled gradient and the arg-max records of tile t + gridDim.x are fetched BEFORE tile t is worked on: the
    // builders are the kernel's critical path, and without the prefetch they spent half of their time waiting for
    // these eight loads (profiles/r2_ncu_stem_stalls.txt).
    uint4 G00, G01, G10, G11;
    uint2 A00, A01, A10, A11;
    auto fetch = [&](long long t, bool& ok) {
      const long long q = t * SW_TK - 1 + j;
      ok = t < n_tiles && j < SW_SPAN && q < gp.Q;  // (q = -1 reads the zero lead guard; beyond the bag: zeros)
      if (ok) {
        G00 = __ldg(gpl + q); G01 = __ldg(gpl + q + 1); G10 = __ldg(gpl + q + wp); G11 = __ldg(gpl + q + wp + 1);
        A00 = __ldg(apl + q); A01 = __ldg(apl + q + 1); A10 = __ldg(apl + q + wp); A11 = __ldg(apl + q + wp + 1);
      }
    };
    bool have;
    fetch(blockIdx.x, have);
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      uint4 out[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) out[k] = make_uint4(0, 0, 0, 0);
      const uint4 g00 = G00, g01 = G01, g10 = G10, g11 = G11;
      const uint2 a00 = A00, a01 = A01, a10 = A10, a11 = A11;
      const bool cur = have;
      fetch(t + gridDim.x, have);
      if (cur) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < npair)
            out[k] = mil_unpool_pair(mil_word(g00, k), mil_word(g01, k), mil_word(g10, k), mil_word(g11, k),
                                     mil_am_lanes(a00, k), mil_am_lanes(a01, k), mil_am_lanes(a10, k), mil_am_lanes(a11, k));
      }
      mbar_wait(&hd->empty[stage], phase ^ 1);  // (the loads and the arithmetic above overlap this wait)
      uint4* abase = reinterpret_cast<uint4*>(stage0 + (size_t)stage * SW_STAGE);
      if (j < SW_SPAN) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < npair) {
            uint4* slot = abase + (size_t)((4 * pc + k) * 3) * SW_TK;
            if (st0) slot[i0] = out[k];
            if (st1) slot[SW_TK + i1] = out[k];
            if (st2) slot[2 * SW_TK + i2] = out[k];
          }
        }
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&hd->a_ready[stage]);
      if (++stage == SW_STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    // ---- epilogue: TMEM lane = accumulator row (slot, channel), column block p = (ci chunk, dx) ----
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    mbar_wait(&hd->done, 0);
    tc_fence_after();
    float* rec = partial + (size_t)blockIdx.x * rec_stride;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int half = 0; half < 2; ++half) {
      const int gl = row >> 3;             // slot within the half; 15 = rows nobody reads
      const int s = half * 15 + gl;
      const int c = s / 3, d = s - c * 3;   // A slot = (co chunk, dx + 1)
      const int co = c * 8 + (row & 7);
      const bool ok = gl < 15;
      for (int p = 0; p <= 3 * SW_CB; ++p) {
        float v[8];
        tmem_ld8(taddr + half * SW_N + p * 8, v);
        tmem_ld_wait();
        if (!ok) continue;
        if (p < 3 * SW_CB) {
          const int cc = p / 3, e = p - cc * 3;  // B slot = (ci chunk, dy + 1)
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) rec[((size_t)(e * 3 + d) * (SW_CB * 8) + cc * 8 + jj) * (SW_CA * 8) + co] = v[jj];
        } else if (d == 1) {
          rec[(size_t)9 * (SW_CB * 8) * (SW_CA * 8) + co] = v[0];
        }
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

int mil_stem_wgrad_ctas() {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      n_sm = 148;
  }
  return n_sm;
}
size_t mil_stem_wgrad_partial_floats() {
  return (size_t)mil_stem_wgrad_ctas() * ((size_t)9 * SW_CB * 8 * SW_CA * 8 + SW_CA * 8);
}
bool mil_stem_wgrad_supported(const MilPF8& gp) { return gp.wp + 1 <= gp.G; }

int mil_launch_stem_wgrad(const void* xs, const MilPF8& gx, const void* g, const MilPF8& gp, const void* am,
                          float* partial, int* ctas_out, long long* rec_out, cudaStream_t s) {
  MIL_REQUIRE(gx.n == gp.n && gx.h == gp.h && gx.w == gp.w && gx.wp == gp.wp && gx.cb == SW_CB && gp.cb == 3,
              "stem_wgrad: geometry mismatch");
  MIL_REQUIRE(mil_stem_wgrad_supported(gp), "stem_wgrad: padded row of %d pixels not supported", gp.wp);
  const long long n_tiles = mil_cdiv(gp.Q, SW_TK);
  const int ctas = (int)std::max<long long>(1, std::min<long long>(n_tiles, mil_stem_wgrad_ctas()));
  const size_t smem = 128 + (size_t)SW_STAGES * SW_STAGE + SW_PLANE;  // + one slot read past the last stage's A half
  const long long rec = (long long)9 * SW_CB * 8 * SW_CA * 8 + SW_CA * 8;
  MIL_SET_SMEM(stem_wgrad_kernel, smem);
  MIL_LAUNCH_PDL(stem_wgrad_kernel, ctas, SW_THREADS, smem, s, (const __nv_bfloat16*)xs, gx, (const __nv_bfloat16*)g, gp,
                                                   (const uint2*)am, partial, rec);
  *ctas_out = ctas;
  *rec_out = rec;
  return 0;
}
