// The attention-MIL head (reference gbm/model.py:200-246, ContextLayer :89-111, nnBlocks.py:47-138), forward
// and backward, as HBM-bound kernels over the bag's feature matrix H [n,80] (fp32).
//
// The head couples the tiles of a bag only through three bag-wide sums, so every kernel works on the
// LOCAL shard of the bag and the caller all-reduces a handful of doubles between the phases
// (SURVEY.md section 8e):
//   phase 1  head_stats    : sum_n x, sum_n x^2 per feature                      -> stats[160]   (AR-1)
//   phase 2  head_scores   : BN (bag statistics), attention MLP, instance MLP, softplus, mask-mix
//                            -> raw[n,3], g[n,3], b[n] and sum g, sum g*b, sum raw, Gram(raw)  -> sums[16] (AR-2)
//   phase 3  head_finalize : A = g / sum g, wROIs, Mterm, softmax, loss, metrics, dLoss/dM
//   phase 4  head_bwd_a    : per-tile backward down to dHz (BN output grad) and the instance-path dH,
//                            weight-gradient partials, sum dHz, sum dHz*xhat          -> bnsums[160] (AR-3)
//   phase 5  head_bwd_b    : BatchNorm1d backward with the bag-wide sums -> dH [n,80]
// All bag-wide sums are accumulated in double, in a fixed order (bit-reproducible).
//
// Work distribution (round 2): the per-tile kernels used to give ONE thread a whole tile (6400 dependent FMAs, 64 blocks of
// 64 threads on a 148-SM part: 46 + 99 us at 4096 tiles, the same at 512).  Now eight threads share a tile -- each owns
// five hidden units of both MLPs, the 3 + 1 outputs are finished with warp shuffles -- a group of 16 tiles is a block
// of 128 threads (256 blocks at 4096 tiles), H rows are read with 128-bit loads, and the backward kernel is persistent
// (<= 2 blocks per SM, parameter-gradient partials kept in registers across groups: one record per block).
#include <algorithm>

#include "mil_common.cuh"
#include "mil_head.cuh"

#define HL 80
#define HD 40
#define HK 3
#define HT 16          // tiles per group: a block of HTHREADS threads works on HT tiles at a time, HSUB threads per tile
#define HSUB 8         // threads per tile: thread `sub` owns the hidden units sub, sub + 8, ..., sub + 32 of both MLPs
#define HJ (HD / HSUB) // hidden units per thread (5)
#define HTHREADS (HT * HSUB)
#define HROW (HL + 1)  // padded smem row
#define BN_EPS 1e-5

// ---------------------------------------------------------------------------------------------------
// phase 1
// ---------------------------------------------------------------------------------------------------
#define STATS_BLOCKS 148
#define STATS_ROWS 16  // rows in flight per block: 16 x 20 threads, one float4 of a row each
__global__ void __launch_bounds__(STATS_ROWS * (HL / 4))
head_stats_kernel(const float* __restrict__ H, int n, double* __restrict__ part) {
  __shared__ double s1[STATS_ROWS][HL], s2[STATS_ROWS][HL];
  const int f4 = threadIdx.x % (HL / 4), r = threadIdx.x / (HL / 4);
  const int per = (int)mil_cdiv(n, (int)gridDim.x);
  const int n0 = blockIdx.x * per, n1 = min(n0 + per, n);
  double a[4] = {0.0, 0.0, 0.0, 0.0}, b[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = n0 + r; i < n1; i += STATS_ROWS) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(H + (size_t)i * HL) + f4);
    const double x0 = v.x, x1 = v.y, x2 = v.z, x3 = v.w;
    a[0] += x0; a[1] += x1; a[2] += x2; a[3] += x3;
    b[0] += x0 * x0; b[1] += x1 * x1; b[2] += x2 * x2; b[3] += x3 * x3;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { s1[r][f4 * 4 + k] = a[k]; s2[r][f4 * 4 + k] = b[k]; }
  __syncthreads();
  if (threadIdx.x < HL) {
    double x = 0.0, y = 0.0;
    for (int k = 0; k < STATS_ROWS; ++k) { x += s1[k][threadIdx.x]; y += s2[k][threadIdx.x]; }
    part[(size_t)blockIdx.x * 2 * HL + threadIdx.x] = x;
    part[(size_t)blockIdx.x * 2 * HL + HL + threadIdx.x] = y;
  }
}
// out[i] = sum over the nblk records of part[b][i]: one block per output, strided partial sums + a fixed tree
// (the shape of the tree depends only on nblk: bit-reproducible)
#define RD_THREADS 128
__global__ void __launch_bounds__(RD_THREADS)
reduce_double_kernel(const double* __restrict__ part, int nblk, int count, double* __restrict__ out) {
  __shared__ double sh[RD_THREADS];
  const int i = blockIdx.x;
  double a = 0.0;
  for (int b = threadIdx.x; b < nblk; b += RD_THREADS) a += part[(size_t)b * count + i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = RD_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[i] = sh[0];
}

int mil_launch_head_stats(const float* H, int n, double* part_ws, double* stats, cudaStream_t s) {
  head_stats_kernel<<<STATS_BLOCKS, STATS_ROWS * (HL / 4), 0, s>>>(H, n, part_ws);
  MIL_LAUNCH_OK();
  reduce_double_kernel<<<2 * HL, RD_THREADS, 0, s>>>(part_ws, STATS_BLOCKS, 2 * HL, stats);
  MIL_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// shared device helpers
// ---------------------------------------------------------------------------------------------------
struct HeadSmem {
  float w1t[HL][HD];   // attention.lin1 weight, transposed [f][d]
  float v1t[HL][HD];   // buffer.lin1 weight, transposed
  float hz[HT][HROW];  // BN output per tile   (later reused for dHz)
  float hm[HT][HROW];  // instance-path input  (later reused for the instance-path dH)
  float scale[HL], shift[HL], mean[HL], rstd[HL];
  float b1[HD], c1[HD], v2[HD];
  float w2[HK][HD];
  float b2[HK], sneg[HK], spos[HK];
  float c2;
};

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// parameters + BN scale / shift of the bag (once per block)
__device__ void head_load_params(HeadSmem& S, const MilHeadParams& P, long long n_global,
                                 const double* __restrict__ stats) {
  for (int i = threadIdx.x; i < HL * HD; i += blockDim.x) {
    const int d = i / HL, f = i % HL;  // coalesced read of W[d][f]
    S.w1t[f][d] = P.att_w1[i];
    S.v1t[f][d] = P.buf_w1[i];
  }
  for (int i = threadIdx.x; i < HD; i += blockDim.x) {
    S.b1[i] = P.att_b1[i];
    S.c1[i] = P.buf_b1[i];
    S.v2[i] = P.buf_w2[i];
    for (int k = 0; k < HK; ++k) S.w2[k][i] = P.att_w2[k * HD + i];
  }
  if (threadIdx.x < HK) {
    S.b2[threadIdx.x] = P.att_b2[threadIdx.x];
    S.sneg[threadIdx.x] = sigmoid_f(-10.f * P.weight_mask[threadIdx.x]);
    S.spos[threadIdx.x] = sigmoid_f(10.f * P.weight_mask[threadIdx.x]);
  }
  if (threadIdx.x == 0) S.c2 = P.buf_b2[0];
  for (int f = threadIdx.x; f < HL; f += blockDim.x) {
    const double m = stats[f] / (double)n_global;
    double var = stats[HL + f] / (double)n_global - m * m;  // biased variance (BatchNorm1d batch statistics)
    if (var < 0.0) var = 0.0;
    const double rs = 1.0 / sqrt(var + BN_EPS);
    S.mean[f] = (float)m;
    S.rstd[f] = (float)rs;
    S.scale[f] = (float)(rs * (double)P.bn_w[f]);
    S.shift[f] = (float)((double)P.bn_b[f] - m * rs * (double)P.bn_w[f]);
  }
  __syncthreads();
}
// the group's H rows (128-bit loads) -> hz (BN output) / hm (instance-path input, LeakyReLU + dropout)
__device__ void head_load_rows(HeadSmem& S, const float* __restrict__ H, const float* __restrict__ drop, int n, int i0) {
  for (int i = threadIdx.x; i < HT * (HL / 4); i += blockDim.x) {
    const int t = i / (HL / 4), f = (i % (HL / 4)) * 4;
    float4 h = make_float4(0.f, 0.f, 0.f, 0.f), dm = make_float4(1.f, 1.f, 1.f, 1.f);
    const bool ok = i0 + t < n;
    if (ok) {
      h = __ldg(reinterpret_cast<const float4*>(H + (size_t)(i0 + t) * HL + f));
      if (drop != nullptr) dm = __ldg(reinterpret_cast<const float4*>(drop + (size_t)(i0 + t) * HL + f));
    }
    const float hv[4] = {h.x, h.y, h.z, h.w}, dv[4] = {dm.x, dm.y, dm.z, dm.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float hz = 0.f, hm = 0.f;
      if (ok) {
        hz = fmaf(hv[k], S.scale[f + k], S.shift[f + k]);
        hm = mil_lrelu(hv[k]);
        if (drop != nullptr) hm *= dv[k] * (1.f / 0.75f);
      }
      S.hz[t][f + k] = hz;
      S.hm[t][f + k] = hm;
    }
  }
  __syncthreads();
}

// forward of both MLPs for tile t, split over its HSUB threads: thread `sub` gets a[j] = tanh(att pre-activation) and
// u[j] = buffer pre-activation (before lrelu) of its hidden units d = sub + HSUB * j; raw[] / bval are complete in EVERY
// thread of the tile (xor-shuffle all-reduce over the 8 lanes, fixed order)
__device__ __forceinline__ void head_tile_forward(const HeadSmem& S, int t, int sub, float a[HJ], float u[HJ],
                                                  float raw[HK], float& bval) {
#pragma unroll
  for (int j = 0; j < HJ; ++j) { a[j] = S.b1[sub + HSUB * j]; u[j] = S.c1[sub + HSUB * j]; }
#pragma unroll 4
  for (int f = 0; f < HL; ++f) {
    const float hz = S.hz[t][f], hm = S.hm[t][f];
#pragma unroll
    for (int j = 0; j < HJ; ++j) {
      a[j] = fmaf(S.w1t[f][sub + HSUB * j], hz, a[j]);
      u[j] = fmaf(S.v1t[f][sub + HSUB * j], hm, u[j]);
    }
  }
#pragma unroll
  for (int k = 0; k < HK; ++k) raw[k] = 0.f;
  bval = 0.f;
#pragma unroll
  for (int j = 0; j < HJ; ++j) {
    const int d = sub + HSUB * j;
    a[j] = tanhf(a[j]);
#pragma unroll
    for (int k = 0; k < HK; ++k) raw[k] = fmaf(S.w2[k][d], a[j], raw[k]);
    bval = fmaf(S.v2[d], mil_lrelu(u[j]), bval);
  }
#pragma unroll
  for (int o = 1; o < HSUB; o <<= 1) {
#pragma unroll
    for (int k = 0; k < HK; ++k) raw[k] += __shfl_xor_sync(0xffffffffu, raw[k], o);
    bval += __shfl_xor_sync(0xffffffffu, bval, o);
  }
#pragma unroll
  for (int k = 0; k < HK; ++k) raw[k] += S.b2[k];
  bval += S.c2;
}

// block-wide sum of `cnt` doubles per thread (cnt <= 16), result in out[] of thread 0 .. written to dst
template <int CNT>
__device__ void block_sum_doubles(double v[CNT], double* s_red /*[blockDim.x/32][CNT]*/, double* dst) {
#pragma unroll
  for (int k = 0; k < CNT; ++k) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    v[k] = x;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int k = 0; k < CNT; ++k) s_red[warp * CNT + k] = v[k];
  __syncthreads();
  if (threadIdx.x < CNT) {
    double x = 0.0;
    for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) x += s_red[wi * CNT + threadIdx.x];
    dst[threadIdx.x] = x;
  }
}

// ---------------------------------------------------------------------------------------------------
// phase 2
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HTHREADS)
head_scores_kernel(MilHeadParams P, const float* __restrict__ H, const float* __restrict__ drop, int n,
                   long long n_global, const double* __restrict__ stats, float* __restrict__ raw_o,
                   float* __restrict__ g_o, float* __restrict__ b_o, double* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HeadSmem& S = *reinterpret_cast<HeadSmem*>(smem_raw);
  __shared__ double s_red[(HTHREADS / 32) * MIL_HEAD_NSUMS];
  const int i0 = blockIdx.x * HT;
  head_load_params(S, P, n_global, stats);
  head_load_rows(S, H, drop, n, i0);
  const int t = threadIdx.x / HSUB, sub = threadIdx.x % HSUB, i = i0 + t;
  double v[MIL_HEAD_NSUMS];
#pragma unroll
  for (int k = 0; k < MIL_HEAD_NSUMS; ++k) v[k] = 0.0;
  float a[HJ], u[HJ], raw[HK], bval;
  head_tile_forward(S, t, sub, a, u, raw, bval);  // (whole warps: the shuffles need every lane)
  if (i < n && sub == 0) {
    float g[HK];
#pragma unroll
    for (int k = 0; k < HK; ++k) {
      g[k] = fmaf(S.sneg[k], softplus_f(raw[k]), S.spos[k]);
      raw_o[(size_t)i * HK + k] = raw[k];
      g_o[(size_t)i * HK + k] = g[k];
      v[k] = (double)g[k];
      v[3 + k] = (double)g[k] * (double)bval;
      v[6 + k] = (double)raw[k];
    }
    b_o[i] = bval;
    v[9] = (double)raw[0] * raw[0];  v[10] = (double)raw[0] * raw[1]; v[11] = (double)raw[0] * raw[2];
    v[12] = (double)raw[1] * raw[1]; v[13] = (double)raw[1] * raw[2]; v[14] = (double)raw[2] * raw[2];
  }
  block_sum_doubles<MIL_HEAD_NSUMS>(v, s_red, part + (size_t)blockIdx.x * MIL_HEAD_NSUMS);
}

int mil_launch_head_scores(const MilHeadParams& P, const float* H, const float* drop, int n, long long n_global,
                           const double* stats, float* raw, float* g, float* b, double* part_ws, double* sums,
                           cudaStream_t s) {
  const int nblk = (int)mil_cdiv(n, HT);
  MIL_SET_SMEM((head_scores_kernel), (int)sizeof(HeadSmem));
  head_scores_kernel<<<nblk, HTHREADS, sizeof(HeadSmem), s>>>(P, H, drop, n, n_global, stats, raw, g, b, part_ws);
  MIL_LAUNCH_OK();
  reduce_double_kernel<<<MIL_HEAD_NSUMS, RD_THREADS, 0, s>>>(part_ws, nblk, MIL_HEAD_NSUMS, sums);
  MIL_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// phase 3: scalars (every block recomputes them from the 16 sums; block 0 stores them) + A, wROIs
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_finalize_kernel(const double* __restrict__ sums, const double* __restrict__ stats, long long n_global,
                     const long long* __restrict__ Y, const float* __restrict__ class_w, int n,
                     const float* __restrict__ g, const float* __restrict__ b, float* __restrict__ A,
                     float* __restrict__ wroi, float* __restrict__ scal) {
  __shared__ float s_inv[HK];
  if (threadIdx.x == 0) {
    double S[HK], M[HK];
    for (int k = 0; k < HK; ++k) {
      S[k] = fmax(sums[k], 1e-12);  // F.normalize(p=1, dim=0): g / max(sum |g|, eps), g > 0
      M[k] = sums[3 + k] / S[k];
      s_inv[k] = (float)(1.0 / S[k]);
    }
    if (blockIdx.x == 0) {
      // softmax / log-softmax of the 3 logits (gbm/model.py:233-235)
      const double mx = fmax(M[0], fmax(M[1], M[2]));
      double e[HK], se = 0.0;
      for (int k = 0; k < HK; ++k) { e[k] = exp(M[k] - mx); se += e[k]; }
      double p[HK], logp[HK];
      int yhat = 0;
      for (int k = 0; k < HK; ++k) {
        p[k] = e[k] / se;
        logp[k] = (M[k] - mx) - log(se);
      }
      // argmax on the fp32 probabilities, first maximum wins (torch.argmax)
      for (int k = 1; k < HK; ++k)
        if ((float)p[k] > (float)p[yhat]) yhat = k;
      const long long y = Y[0];
      // label-smoothed, class-weighted cross entropy with probabilities (nnBlocks.py:71-85,121-133)
      double loss = 0.0, tw = 0.0, t[HK], cw[HK];
      for (int k = 0; k < HK; ++k) {
        t[k] = (k == y) ? 0.75 : 0.125;
        cw[k] = class_w ? (double)class_w[k] : 1.0;
        loss -= cw[k] * t[k] * logp[k];
        tw += cw[k] * t[k];
      }
      for (int k = 0; k < HK; ++k) {
        scal[MIL_SC_M + k] = (float)M[k];
        scal[MIL_SC_P + k] = (float)p[k];
        scal[MIL_SC_DM + k] = (float)(tw * p[k] - cw[k] * t[k]);
        scal[MIL_SC_S + k] = (float)S[k];
      }
      scal[MIL_SC_LOSS] = (float)loss;
      double mu = 0.0;
      for (int k = 0; k < HK; ++k) {
        const double m = sums[6 + k] / (double)n_global;
        mu += m * m;
      }
      scal[MIL_SC_AMU] = (float)(0.5 * mu);
      const double n0 = fmax(sqrt(sums[9]), 1e-12), n1 = fmax(sqrt(sums[12]), 1e-12),
                   n2 = fmax(sqrt(sums[14]), 1e-12);
      scal[MIL_SC_AVAR] = (float)(2.0 * (sums[10] / (n0 * n1) + sums[11] / (n0 * n2) + sums[13] / (n1 * n2)) / 9.0);
      double hh = 0.0;
      for (int f = 0; f < HL; ++f) hh += stats[HL + f];
      scal[MIL_SC_KLD] = (float)(0.5 * hh / ((double)n_global * HL));
      scal[MIL_SC_YHAT] = (float)yhat;
      scal[MIL_SC_ERR] = (yhat == (int)y) ? 0.f : 1.f;
    }
  }
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float bv = b[i];
#pragma unroll
    for (int k = 0; k < HK; ++k) {
      const float a = g[(size_t)i * HK + k] * s_inv[k];
      A[(size_t)k * n + i] = a;
      wroi[(size_t)k * n + i] = a * bv;
    }
  }
}

int mil_launch_head_finalize(const double* sums, const double* stats, long long n_global, const long long* Y,
                             const float* class_w, int n, const float* g, const float* b, float* A, float* wroi,
                             float* scal, cudaStream_t s) {
  const int nblk = (int)std::max<long long>(1, std::min<long long>(mil_cdiv(n, 256), 148 * 4));
  head_finalize_kernel<<<nblk, 256, 0, s>>>(sums, stats, n_global, Y, class_w, n, g, b, A, wroi, scal);
  MIL_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// phase 4: per-tile backward.  Record of one block's parameter-gradient partials (floats):
// ---------------------------------------------------------------------------------------------------
#define HB_W1 0                       // attention.lin1.weight [40][80]
#define HB_B1 (HB_W1 + HD * HL)       // attention.lin1.bias   [40]
#define HB_W2 (HB_B1 + HD)            // attention.lin2.weight [3][40]
#define HB_B2 (HB_W2 + HK * HD)       // attention.lin2.bias   [3]
#define HB_V1 (HB_B2 + HK)            // buffer.lin1.weight    [40][80]
#define HB_C1 (HB_V1 + HD * HL)       // buffer.lin1.bias      [40]
#define HB_V2 (HB_C1 + HD)            // buffer.classifier.weight [40]
#define HB_C2 (HB_V2 + HD)            // buffer.classifier.bias   [1]
#define HB_WM (HB_C2 + 1)             // weight_mask [3]
#define HB_BNW (HB_WM + HK)           // context.bn.weight [80]  (local sum dHz*xhat)
#define HB_BNB (HB_BNW + HL)          // context.bn.bias   [80]  (local sum dHz)
#define HB_REC (HB_BNB + HL)
#define HB_MAX_BLOCKS 296             // persistent grid: at most two blocks per SM
#define HB_OUT (HD * HL / HTHREADS)   // lin1 weight-gradient elements per thread and matrix (25)

__global__ void __launch_bounds__(HTHREADS)
head_bwd_a_kernel(MilHeadParams P, const float* __restrict__ H, const float* __restrict__ drop, int n,
                  long long n_global, const double* __restrict__ stats, const float* __restrict__ raw_i,
                  const float* __restrict__ g_i, const float* __restrict__ b_i, const float* __restrict__ scal,
                  const float* __restrict__ gloss, float* __restrict__ dHz_o, float* __restrict__ dHi_o,
                  float* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HeadSmem& S = *reinterpret_cast<HeadSmem*>(smem_raw);
  float(*s_da)[HD + 1] = reinterpret_cast<float(*)[HD + 1]>(smem_raw + sizeof(HeadSmem));  // [HT][41]
  float(*s_dp)[HD + 1] = s_da + HT;                                                        // [HT][41]
  float(*s_t)[HD + 1] = s_dp + HT;                                                         // tanh output  [HT][41]
  float(*s_u)[HD + 1] = s_t + HT;                                                          // lrelu output [HT][41]
  float(*s_small)[8] = reinterpret_cast<float(*)[8]>(s_u + HT);                            // [HT][8]: draw[3], db, dwm[3]
  const int t = threadIdx.x / HSUB, sub = threadIdx.x % HSUB;
  const float gl = gloss ? gloss[0] : 1.f;
  head_load_params(S, P, n_global, stats);
  float dMs[HK], Sk[HK], Mk[HK];
#pragma unroll
  for (int k = 0; k < HK; ++k) { dMs[k] = gl * scal[MIL_SC_DM + k]; Sk[k] = scal[MIL_SC_S + k]; Mk[k] = scal[MIL_SC_M + k]; }

  // parameter-gradient partials of THIS BLOCK, kept in registers across its groups of tiles:
  //   every thread: lin1 weight elements o = threadIdx.x + HTHREADS * k  (d = o / 80, f = o % 80) of both MLPs
  //   threads < 40: hidden unit d = threadIdx.x -> b1, c1, v2, w2[3];  threads < 7: b2[3], c2, weight_mask[3]
  //   threads < 80: feature f = threadIdx.x -> the local BatchNorm sums
  float gw1[HB_OUT], gv1[HB_OUT];
#pragma unroll
  for (int k = 0; k < HB_OUT; ++k) gw1[k] = gv1[k] = 0.f;
  float gb1 = 0.f, gc1 = 0.f, gv2 = 0.f, gw2[HK] = {0.f, 0.f, 0.f}, gsm = 0.f, gbnw = 0.f, gbnb = 0.f;

  const int ngroups = (int)mil_cdiv(n, HT);
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int i0 = grp * HT, i = i0 + t;
    head_load_rows(S, H, drop, n, i0);
    float a[HJ], u[HJ], raw[HK], bval;
    head_tile_forward(S, t, sub, a, u, raw, bval);  // a = tanh(.), u = pre-activation of the instance MLP
    float da[HJ], dp[HJ];
    float draw[HK] = {0.f, 0.f, 0.f}, dwm[HK] = {0.f, 0.f, 0.f}, db = 0.f;
    const bool ok = i < n;
    if (ok) {
#pragma unroll
      for (int k = 0; k < HK; ++k) {
        const float gk = g_i[(size_t)i * HK + k];
        db = fmaf(dMs[k], gk / Sk[k], db);                   // dL/db_n = sum_k dM_k A_kn
        const float dg = dMs[k] * (bval - Mk[k]) / Sk[k];    // L1-normalise backward (needs only S_k, M_k)
        const float sg = raw[k] > 20.f ? 1.f : sigmoid_f(raw[k]);  // softplus'
        draw[k] = dg * S.sneg[k] * sg;
        const float sn = S.sneg[k], sp = S.spos[k];
        dwm[k] = dg * (-10.f * sn * (1.f - sn) * softplus_f(raw[k]) + 10.f * sp * (1.f - sp));
      }
    }
#pragma unroll
    for (int j = 0; j < HJ; ++j) {
      const int d = sub + HSUB * j;
      float dt = 0.f;
#pragma unroll
      for (int k = 0; k < HK; ++k) dt = fmaf(S.w2[k][d], draw[k], dt);
      da[j] = ok ? dt * (1.f - a[j] * a[j]) : 0.f;
      dp[j] = ok ? db * S.v2[d] * mil_lrelu_grad(u[j]) : 0.f;
      s_da[t][d] = da[j];
      s_dp[t][d] = dp[j];
      s_t[t][d] = ok ? a[j] : 0.f;
      s_u[t][d] = ok ? mil_lrelu(u[j]) : 0.f;
    }
    if (sub == 0) {
      s_small[t][0] = draw[0]; s_small[t][1] = draw[1]; s_small[t][2] = draw[2]; s_small[t][3] = db;
      s_small[t][4] = dwm[0]; s_small[t][5] = dwm[1]; s_small[t][6] = dwm[2]; s_small[t][7] = 0.f;
    }
    __syncthreads();

    // ---- parameter-gradient partials (outer products over the group's HT tiles) ----
#pragma unroll
    for (int k = 0; k < HB_OUT; ++k) {
      const int o = threadIdx.x + HTHREADS * k, d = o / HL, f = o % HL;
      float w1 = gw1[k], v1 = gv1[k];
#pragma unroll
      for (int tt = 0; tt < HT; ++tt) {
        w1 = fmaf(s_da[tt][d], S.hz[tt][f], w1);
        v1 = fmaf(s_dp[tt][d], S.hm[tt][f], v1);
      }
      gw1[k] = w1;
      gv1[k] = v1;
    }
    if (threadIdx.x < HD) {
      const int d = threadIdx.x;
#pragma unroll
      for (int tt = 0; tt < HT; ++tt) {
        gb1 += s_da[tt][d];
        gc1 += s_dp[tt][d];
        gv2 = fmaf(s_small[tt][3], s_u[tt][d], gv2);
#pragma unroll
        for (int k = 0; k < HK; ++k) gw2[k] = fmaf(s_small[tt][k], s_t[tt][d], gw2[k]);
      }
    }
    if (threadIdx.x < 7)
      for (int tt = 0; tt < HT; ++tt) gsm += s_small[tt][threadIdx.x];
    __syncthreads();  // everyone is done reading hz / hm as forward activations

    // ---- input gradients: dHz = W1^T da (kept for phase 5), instance path dH = lrelu'(H) * mask/0.75 * V1^T dp ----
    // thread (t, sub) takes the features f = sub + HSUB * m of its tile
#pragma unroll 2
    for (int m = 0; m < HL / HSUB; ++m) {
      const int f = sub + HSUB * m;
      float dz = 0.f, dm = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < HD / 4; ++d4) {
        const float4 w = *reinterpret_cast<const float4*>(&S.w1t[f][d4 * 4]);
        const float4 v = *reinterpret_cast<const float4*>(&S.v1t[f][d4 * 4]);
        dz = fmaf(w.x, s_da[t][d4 * 4 + 0], dz); dz = fmaf(w.y, s_da[t][d4 * 4 + 1], dz);
        dz = fmaf(w.z, s_da[t][d4 * 4 + 2], dz); dz = fmaf(w.w, s_da[t][d4 * 4 + 3], dz);
        dm = fmaf(v.x, s_dp[t][d4 * 4 + 0], dm); dm = fmaf(v.y, s_dp[t][d4 * 4 + 1], dm);
        dm = fmaf(v.z, s_dp[t][d4 * 4 + 2], dm); dm = fmaf(v.w, s_dp[t][d4 * 4 + 3], dm);
      }
      S.hz[t][f] = dz;  // own tile only: no hazard with the other tiles' threads
      S.hm[t][f] = dm;
    }
    __syncthreads();

    // xhat is recovered from H (xhat = (hz - beta) / gamma is ill-defined for gamma = 0)
    for (int idx = threadIdx.x; idx < HT * (HL / 4); idx += HTHREADS) {
      const int tt = idx / (HL / 4), f = (idx % (HL / 4)) * 4;
      if (i0 + tt < n) {
        const float4 h = __ldg(reinterpret_cast<const float4*>(H + (size_t)(i0 + tt) * HL + f));
        float4 dm = make_float4(1.f, 1.f, 1.f, 1.f);
        if (drop != nullptr) dm = __ldg(reinterpret_cast<const float4*>(drop + (size_t)(i0 + tt) * HL + f));
        const float hv[4] = {h.x, h.y, h.z, h.w}, dv[4] = {dm.x, dm.y, dm.z, dm.w};
        float oz[4], oi[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float dmi = S.hm[tt][f + k] * mil_lrelu_grad(hv[k]);
          if (drop != nullptr) dmi *= dv[k] * (1.f / 0.75f);
          oz[k] = S.hz[tt][f + k];
          oi[k] = dmi;
        }
        *reinterpret_cast<float4*>(dHz_o + (size_t)(i0 + tt) * HL + f) = make_float4(oz[0], oz[1], oz[2], oz[3]);
        *reinterpret_cast<float4*>(dHi_o + (size_t)(i0 + tt) * HL + f) = make_float4(oi[0], oi[1], oi[2], oi[3]);
      }
    }
    // local BN sums: sum_t dHz, sum_t dHz * xhat  (per feature; fixed order over the group's tiles)
    if (threadIdx.x < HL) {
      const int f = threadIdx.x;
      for (int tt = 0; tt < HT; ++tt) {
        if (i0 + tt < n) {
          const float h = __ldg(H + (size_t)(i0 + tt) * HL + f);
          const float xh = (h - S.mean[f]) * S.rstd[f];
          gbnb += S.hz[tt][f];
          gbnw = fmaf(S.hz[tt][f], xh, gbnw);
        }
      }
    }
    __syncthreads();  // hz / hm are rewritten by the next group's rows
  }

  // ---- this block's record ----
  float* rec = part + (size_t)blockIdx.x * HB_REC;
#pragma unroll
  for (int k = 0; k < HB_OUT; ++k) {
    const int o = threadIdx.x + HTHREADS * k;
    rec[HB_W1 + o] = gw1[k];
    rec[HB_V1 + o] = gv1[k];
  }
  if (threadIdx.x < HD) {
    const int d = threadIdx.x;
    rec[HB_B1 + d] = gb1;
    rec[HB_C1 + d] = gc1;
    rec[HB_V2 + d] = gv2;
#pragma unroll
    for (int k = 0; k < HK; ++k) rec[HB_W2 + k * HD + d] = gw2[k];
  }
  if (threadIdx.x < 3) rec[HB_B2 + threadIdx.x] = gsm;
  else if (threadIdx.x == 3) rec[HB_C2] = gsm;
  else if (threadIdx.x < 7) rec[HB_WM + threadIdx.x - 4] = gsm;
  if (threadIdx.x < HL) {
    rec[HB_BNW + threadIdx.x] = gbnw;
    rec[HB_BNB + threadIdx.x] = gbnb;
  }
}

// fixed-order reduction of the per-block records: parameter gradients (+=, fp32) and BN sums (double).
// threadIdx.x walks the record (coalesced), threadIdx.y takes every HBR_PARTS-th record, four loads in flight
#define HBR_PARTS 8
__global__ void __launch_bounds__(32 * HBR_PARTS)
head_bwd_reduce_kernel(const float* __restrict__ part, int nblk, MilHeadGrads G, double* __restrict__ bnsums) {
  __shared__ double sh[HBR_PARTS][32];
  const int o = blockIdx.x * 32 + threadIdx.x;
  double acc = 0.0;
  if (o < HB_REC) {
    int b = threadIdx.y;
    for (; b + 3 * HBR_PARTS < nblk; b += 4 * HBR_PARTS) {
      const float v0 = part[(size_t)b * HB_REC + o], v1 = part[(size_t)(b + HBR_PARTS) * HB_REC + o];
      const float v2 = part[(size_t)(b + 2 * HBR_PARTS) * HB_REC + o], v3 = part[(size_t)(b + 3 * HBR_PARTS) * HB_REC + o];
      acc += (double)v0; acc += (double)v1; acc += (double)v2; acc += (double)v3;
    }
    for (; b < nblk; b += HBR_PARTS) acc += (double)part[(size_t)b * HB_REC + o];
  }
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y != 0 || o >= HB_REC) return;
#pragma unroll
  for (int k = 1; k < HBR_PARTS; ++k) acc += sh[k][threadIdx.x];
  const float v = (float)acc;
  if (o < HB_B1) G.att_w1[o - HB_W1] += v;
  else if (o < HB_W2) G.att_b1[o - HB_B1] += v;
  else if (o < HB_B2) G.att_w2[o - HB_W2] += v;
  else if (o < HB_V1) G.att_b2[o - HB_B2] += v;
  else if (o < HB_C1) G.buf_w1[o - HB_V1] += v;
  else if (o < HB_V2) G.buf_b1[o - HB_C1] += v;
  else if (o < HB_C2) G.buf_w2[o - HB_V2] += v;
  else if (o < HB_WM) G.buf_b2[0] += v;
  else if (o < HB_BNW) G.weight_mask[o - HB_WM] += v;
  else if (o < HB_BNB) {
    G.bn_w[o - HB_BNW] += v;
    bnsums[HL + (o - HB_BNW)] = acc;  // sum dHz * xhat
  } else {
    G.bn_b[o - HB_BNB] += v;
    bnsums[o - HB_BNB] = acc;         // sum dHz
  }
}

static int head_bwd_blocks(int n) { return (int)std::min<long long>(mil_cdiv(n, HT), HB_MAX_BLOCKS); }
size_t mil_head_bwd_partial_floats(int n) { return (size_t)head_bwd_blocks(n) * HB_REC; }

int mil_launch_head_bwd_a(const MilHeadParams& P, const MilHeadGrads& G, const float* H, const float* drop, int n,
                          long long n_global, const double* stats, const float* raw, const float* g, const float* b,
                          const float* scal, const float* gloss, float* dHz, float* dHi, float* part_ws,
                          double* bnsums, cudaStream_t s) {
  const int nblk = head_bwd_blocks(n);
  const size_t smem = sizeof(HeadSmem) + (size_t)4 * HT * (HD + 1) * sizeof(float) + (size_t)HT * 8 * sizeof(float);
  MIL_SET_SMEM((head_bwd_a_kernel), (int)smem);
  head_bwd_a_kernel<<<nblk, HTHREADS, smem, s>>>(P, H, drop, n, n_global, stats, raw, g, b, scal, gloss, dHz, dHi, part_ws);
  MIL_LAUNCH_OK();
  head_bwd_reduce_kernel<<<(int)mil_cdiv(HB_REC, 32), dim3(32, HBR_PARTS), 0, s>>>(part_ws, nblk, G, bnsums);
  MIL_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// phase 5: BatchNorm1d backward with bag-wide sums + the instance-path gradient
//   dH = gamma*rstd * (dHz - mean_n dHz - xhat * mean_n(dHz*xhat)) + dHi
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_bwd_b_kernel(const float* __restrict__ bn_w, const float* __restrict__ H, int n, long long n_global,
                  const double* __restrict__ stats, const double* __restrict__ bnsums,
                  const float* __restrict__ dHz, const float* __restrict__ dHi, float* __restrict__ dH) {
  __shared__ float s_mean[HL], s_rstd[HL], s_k[HL], s_m1[HL], s_m2[HL];
  for (int f = threadIdx.x; f < HL; f += blockDim.x) {
    const double m = stats[f] / (double)n_global;
    double var = stats[HL + f] / (double)n_global - m * m;
    if (var < 0.0) var = 0.0;
    const double rs = 1.0 / sqrt(var + BN_EPS);
    s_mean[f] = (float)m;
    s_rstd[f] = (float)rs;
    s_k[f] = (float)(rs * (double)bn_w[f]);
    s_m1[f] = (float)(bnsums[f] / (double)n_global);
    s_m2[f] = (float)(bnsums[HL + f] / (double)n_global);
  }
  __syncthreads();
  const long long total = (long long)n * HL;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % HL);
    const float xh = (H[i] - s_mean[f]) * s_rstd[f];
    dH[i] = fmaf(s_k[f], dHz[i] - s_m1[f] - xh * s_m2[f], dHi[i]);
  }
}

int mil_launch_head_bwd_b(const float* bn_w, const float* H, int n, long long n_global, const double* stats,
                          const double* bnsums, const float* dHz, const float* dHi, float* dH, cudaStream_t s) {
  const int nblk = (int)std::max<long long>(1, std::min<long long>(mil_cdiv((long long)n * HL, 256), 148 * 8));
  head_bwd_b_kernel<<<nblk, 256, 0, s>>>(bn_w, H, n, n_global, stats, bnsums, dHz, dHi, dH);
  MIL_LAUNCH_OK();
  return 0;
}

size_t mil_head_part_doubles(int n) {
  return std::max<size_t>((size_t)STATS_BLOCKS * 2 * HL, (size_t)mil_cdiv(n, HT) * MIL_HEAD_NSUMS);
}
