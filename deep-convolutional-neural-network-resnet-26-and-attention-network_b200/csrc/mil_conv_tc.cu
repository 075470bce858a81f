// tcgen05 implicit-GEMM 3x3 / stride-1 convolution on the PF8 layout (forward AND data gradient), bf16
// operands, fp32 accumulation in TMEM, fused epilogue (bias, residual, LeakyReLU / LeakyReLU').
// Reference semantics: nnBlocks.py:178-187 (conv3x3 + bias -> [+identity] -> LeakyReLU(0.1)) and its autograd.
//
// GEMM view:  D[128 flat pixels][Cout] = sum over K-groups g = (tap t, 8-channel chunk c) of
//             A_g[128][8] * B_g[8][Cout],   A_g[i][:] = x[chunk c][q0 + i + shift_t][0..7]
// Because PF8 turns a tap into a constant shift of the flat pixel index, ALL nine taps read the same
// shared-memory copy of the input: per tile one bulk-TMA copy per channel chunk brings in the
// 128 + 2*(W+2) pixels around the tile ("span"), 16 bytes per pixel, and a tap is just a different START
// ADDRESS of the (un-swizzled, K-major) UMMA descriptor.  8 consecutive pixels x 16 B = one 128-byte core
// matrix (SBO = 128 B); the two 8-channel halves of a K=16 MMA are two (tap, chunk) groups, LBO = their
// address distance.  Weights sit in shared memory for the whole kernel in the matching core-matrix layout.
//
// Warp roles (one persistent CTA per SM, tiles strided over CTAs):
//   warp 0      : producer -- bulk-TMA (cp.async.bulk) of the span planes into a 3-stage ring, mbarrier tx
//   warp 1      : MMA issuer -- one elected lane issues the tcgen05.mma chain of a tile into one of two
//                 TMEM accumulator stages, tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2..9  : epilogue, two groups of 4 warps (one per accumulator stage): prefetch residual / activation
//                 chunks, tcgen05.ld the accumulator row (thread = pixel), bias / residual / LeakyReLU,
//                 16-byte coalesced stores; pad pixels are written as zeros (PF8 invariant).
#include <algorithm>

#include "mil_common.cuh"
#include "mil_conv_tc.cuh"
#include "mil_tc_ptx.cuh"

#define TC_M 128
#define TC_STAGES 3
#define TC_ACC 2
#define TC_THREADS 320  // 10 warps

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

// ---- the kernel ----------------------------------------------------------------------------------------
struct TcSmemHeader {
  uint64_t full[TC_STAGES], empty[TC_STAGES], acc_full[TC_ACC], acc_empty[TC_ACC], b_full;
  uint32_t tmem_base;
  uint32_t a_off[MIL_TC_MAX_MMA];  // byte offset of the first K-half inside an A stage
  uint32_t a_lbo[MIL_TC_MAX_MMA];  // byte distance to the second K-half
  float bias[96];
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __nv_bfloat16* __restrict__ x, MilPF8 gx, const __nv_bfloat16* __restrict__ wtc,
               const float* __restrict__ bias, const __nv_bfloat16* res, const __nv_bfloat16* __restrict__ act,
               __nv_bfloat16* out, MilPF8 go, MilTcShape sh, int epi, int transposed, int sub) {
  extern __shared__ __align__(128) unsigned char smem[];
  TcSmemHeader* hd = reinterpret_cast<TcSmemHeader*>(smem);
  const uint32_t hdr_bytes = (uint32_t)((sizeof(TcSmemHeader) + 127) / 128 * 128);
  unsigned char* bsm = smem + hdr_bytes;                      // B operand blocks
  const uint32_t b_bytes = (uint32_t)sh.nmma * 2 * sh.npad * 16;
  const int halo = sh.ks == 3 ? gx.wp + 1 : 0;
  const int span = TC_M + 2 * halo;
  const uint32_t plane = (uint32_t)span * 16;                 // one channel-chunk plane of a stage
  const uint32_t stage_bytes = plane * (sh.cbin + 1);         // + one all-zero plane (odd K-group count)
  unsigned char* asm0 = bsm + ((b_bytes + 127) / 128 * 128);  // A stages
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = mil_cdiv(gx.Q, TC_M);  // tiles run over the INPUT resolution (== output unless sub)
  const uint32_t acc_stride = (uint32_t)((sh.npad + 31) / 32 * 32);
  uint32_t tmem_cols = 32;
  while (tmem_cols < acc_stride * TC_ACC) tmem_cols <<= 1;

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&hd->full[s], 1); mbar_init(&hd->empty[s], 1); }
    for (int a = 0; a < TC_ACC; ++a) { mbar_init(&hd->acc_full[a], 1); mbar_init(&hd->acc_empty[a], 4); }
    mbar_init(&hd->b_full, 1);
    fence_barrier_init();
  }
  for (int j = threadIdx.x; j < sh.nmma; j += blockDim.x) {
    // shift of tap (dy,dx): forward reads x(q + (dy-1)*wp + dx-1); the data gradient reads dz(q + (1-dy)*wp + 1-dx)
    int off[2];
    for (int h = 0; h < 2; ++h) {
      const int tap = sh.g_tap[2 * j + h], chunk = sh.g_chunk[2 * j + h];
      if (tap == 0xFF) {
        off[h] = -1;
      } else {
        const int dy = tap / 3, dx = tap % 3;
        int s = transposed ? ((1 - dy) * gx.wp + (1 - dx)) : ((dy - 1) * gx.wp + (dx - 1));
        if (sh.ks == 1) s = 0;
        off[h] = chunk * (int)plane + (halo + s) * 16;
      }
    }
    if (off[1] < 0) off[1] = sh.cbin * (int)plane + halo * 16;  // dummy half -> the all-zero plane
    hd->a_off[j] = (uint32_t)off[0];
    hd->a_lbo[j] = (uint32_t)(off[1] - off[0]);                 // host guarantees off[1] > off[0]
  }
  for (int i = threadIdx.x; i < 96; i += blockDim.x)
    hd->bias[i] = (bias != nullptr && i < go.c) ? bias[i] : 0.f;
  // zero plane of every stage
  for (int s = 0; s < TC_STAGES; ++s) {
    uint4* zp = reinterpret_cast<uint4*>(asm0 + (size_t)s * stage_bytes + (size_t)sh.cbin * plane);
    for (int i = threadIdx.x; i < span; i += blockDim.x) zp[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();  // generic-proxy writes (zero planes) -> visible to the tensor-core (async) proxy
  if (warp == 1) tmem_alloc(&hd->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hd->tmem_base;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      mbar_expect_tx(&hd->b_full, b_bytes);
      bulk_g2s(bsm, wtc, b_bytes, &hd->b_full);
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        mbar_wait(&hd->empty[stage], phase ^ 1);
        mbar_expect_tx(&hd->full[stage], plane * sh.cbin);
        const long long q0 = t * TC_M;
        unsigned char* dst = asm0 + (size_t)stage * stage_bytes;
        for (int c = 0; c < sh.cbin; ++c)
          bulk_g2s(dst + (size_t)c * plane, x + mil_pf8_off(gx, c, q0 - halo), plane, &hd->full[stage]);
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D = f32, A = B = bf16, both K-major, N = npad, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(sh.npad >> 3) << 17) |
                             ((uint32_t)(TC_M >> 4) << 24);
      mbar_wait(&hd->b_full, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t b_base = smem_u32(bsm);
      const uint32_t b_blk = (uint32_t)sh.npad * 16;  // one K-half of B
      for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        mbar_wait(&hd->acc_empty[acc], acc_phase ^ 1);
        mbar_wait(&hd->full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(asm0 + (size_t)stage * stage_bytes);
        const uint32_t d = tmem_base + acc * acc_stride;
        for (int j = 0; j < sh.nmma; ++j) {
          const uint64_t ad = make_desc(a_base + hd->a_off[j], hd->a_lbo[j], 128);
          const uint64_t bd = make_desc(b_base + (uint32_t)j * 2 * b_blk, b_blk, 128);
          umma_bf16(d, ad, bd, idesc, j > 0);
        }
        umma_commit(&hd->empty[stage]);   // smem stage reusable once these MMAs have read it
        umma_commit(&hd->acc_full[acc]);  // accumulator complete
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        if (++acc == TC_ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: group eg serves accumulator stage eg =====================
    const int eg = (warp - 2) >> 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    uint32_t acc_phase = 0;
    long long it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      if ((it & 1) != eg) continue;
      // this thread's pixel: flat index q at the input resolution; qo = where it is stored.  sub: the stride-2
      // convolutions are evaluated at full resolution and only the even (y, x) positions are kept.
      const long long q = t * TC_M + row;
      long long qo = q;
      bool in_range = q < gx.Q;
      bool is_pad = true;
      if (in_range) {
        const int n = (int)(q / gx.P);
        const int r = (int)(q - (long long)n * gx.P);
        const int y = r / gx.wp, xo = r - y * gx.wp;
        if (!sub) {
          is_pad = (y == gx.h) || (xo == gx.w);
        } else {
          const int yh = y >> 1, xh = xo >> 1;
          in_range = !(y & 1) && !(xo & 1) && yh <= go.h && xh <= go.w;
          is_pad = (yh == go.h) || (xh == go.w);
          qo = (long long)n * go.P + (long long)yh * go.wp + xh;
        }
      }
      mbar_wait(&hd->acc_full[eg], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + eg * acc_stride + ((uint32_t)(quarter * 32) << 16);
      for (int c = 0; c < sh.cbout; ++c) {
        float v[8];
        tmem_ld8(taddr + c * 8, v);
        tmem_ld_wait();
        if (in_range) {
          const long long o = mil_pf8_off(go, c, qo);
          if (is_pad) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
          } else {
            if (res != nullptr) {
              float rv[8];
              mil_load8(res + o, rv);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += rv[j];
            }
            if (epi != MIL_EPI_DGRAD) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += hd->bias[c * 8 + j];
            }
            if (epi == MIL_EPI_FWD) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = mil_lrelu(v[j]);
            } else if (epi == MIL_EPI_DGRAD) {
              float av[8];
              mil_load8(act + o, av);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] *= mil_lrelu_grad(av[j]);
            }
          }
          mil_store8(out + o, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hd->acc_empty[eg]);
      acc_phase ^= 1;
    }
  }
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
  sh.ks = ks;
  sh.cbin = (cin + 7) / 8;
  sh.cbout = (cout + 7) / 8;
  sh.npad = (cout + 15) / 16 * 16;
  const int ng = ks * ks * sh.cbin;
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

size_t mil_tc_wpack_bytes(const MilTcShape& sh) { return (size_t)sh.nmma * 2 * sh.npad * 16; }

int mil_launch_pack_tc(const float* wp, void* wtc, const MilTcShape& sh, cudaStream_t s) {
  const int total = sh.nmma * 2 * sh.npad * 8;
  pack_tc_kernel<<<(int)mil_cdiv(total, 256), 256, 0, s>>>(wp, (__nv_bfloat16*)wtc, sh);
  MIL_LAUNCH_OK();
  return 0;
}

static size_t tc_smem_bytes(const MilPF8& gx, const MilTcShape& sh) {
  const size_t hdr = (sizeof(TcSmemHeader) + 127) / 128 * 128;
  const size_t b = ((size_t)sh.nmma * 2 * sh.npad * 16 + 127) / 128 * 128;
  const size_t span = TC_M + 2 * (sh.ks == 3 ? gx.wp + 1 : 0);
  return hdr + b + (size_t)TC_STAGES * span * 16 * (sh.cbin + 1);
}

int mil_launch_conv_tc(int transposed, const void* x, const MilPF8& gx, const void* wtc, const MilTcShape& sh,
                       const float* bias, const void* res, const void* act, void* out, const MilPF8& go, int epi,
                       int sub, cudaStream_t s) {
  if (sub)
    MIL_REQUIRE(gx.n == go.n && go.h == (gx.h - 1) / 2 + 1 && go.w == (gx.w - 1) / 2 + 1 && !transposed,
                "conv_tc: stride-2 geometry mismatch");
  else
    MIL_REQUIRE(gx.n == go.n && gx.h == go.h && gx.w == go.w, "conv_tc: geometry mismatch");
  MIL_REQUIRE(gx.cb == sh.cbin && go.cb == sh.cbout, "conv_tc: channel chunks do not match the packed weights");
  MIL_REQUIRE(epi != MIL_EPI_DGRAD || act != nullptr, "conv_tc: DGRAD epilogue needs the activation tensor");
  const size_t smem = tc_smem_bytes(gx, sh);
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
  MIL_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_tiles = mil_cdiv(gx.Q, TC_M);
  const int grid = (int)std::min<long long>(n_tiles, n_sm);
  conv_tc_kernel<<<grid, TC_THREADS, smem, s>>>((const __nv_bfloat16*)x, gx, (const __nv_bfloat16*)wtc, bias,
                                                (const __nv_bfloat16*)res, (const __nv_bfloat16*)act,
                                                (__nv_bfloat16*)out, go, sh, epi, transposed, sub);
  MIL_LAUNCH_OK();
  return 0;
}
