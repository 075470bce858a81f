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
#include <algorithm>

#include "mil_common.cuh"
#include "mil_head.cuh"

#define HL 80
#define HD 40
#define HK 3
#define HT 64          // tiles per block (one thread per tile)
#define HROW (HL + 1)  // padded smem row
#define BN_EPS 1e-5

// ---------------------------------------------------------------------------------------------------
// phase 1
// ---------------------------------------------------------------------------------------------------
#define STATS_BLOCKS 128
__global__ void __launch_bounds__(320)
head_stats_kernel(const float* __restrict__ H, int n, double* __restrict__ part) {
  __shared__ double s1[4][HL], s2[4][HL];
  const int f = threadIdx.x % HL, r = threadIdx.x / HL;
  const int per = (int)mil_cdiv(n, (int)gridDim.x);
  const int n0 = blockIdx.x * per, n1 = min(n0 + per, n);
  double a = 0.0, b = 0.0;
  for (int i = n0 + r; i < n1; i += 4) {
    const double v = (double)H[(size_t)i * HL + f];
    a += v;
    b += v * v;
  }
  s1[r][f] = a;
  s2[r][f] = b;
  __syncthreads();
  if (r == 0) {
    part[(size_t)blockIdx.x * 2 * HL + f] = s1[0][f] + s1[1][f] + s1[2][f] + s1[3][f];
    part[(size_t)blockIdx.x * 2 * HL + HL + f] = s2[0][f] + s2[1][f] + s2[2][f] + s2[3][f];
  }
}
__global__ void reduce_double_kernel(const double* __restrict__ part, int nblk, int count, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double a = 0.0;
  for (int b = 0; b < nblk; ++b) a += part[(size_t)b * count + i];
  out[i] = a;
}

int mil_launch_head_stats(const float* H, int n, double* part_ws, double* stats, cudaStream_t s) {
  head_stats_kernel<<<STATS_BLOCKS, 320, 0, s>>>(H, n, part_ws);
  MIL_LAUNCH_OK();
  reduce_double_kernel<<<2, 128, 0, s>>>(part_ws, STATS_BLOCKS, 2 * HL, stats);
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

// loads parameters + BN scale/shift and the block's H tile; fills hz / hm
__device__ void head_load_tile(HeadSmem& S, const MilHeadParams& P, const float* __restrict__ H,
                               const float* __restrict__ drop, int n, long long n_global,
                               const double* __restrict__ stats, int i0) {
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
  for (int i = threadIdx.x; i < HT * HL; i += blockDim.x) {
    const int t = i / HL, f = i % HL;
    float hz = 0.f, hm = 0.f;
    if (i0 + t < n) {
      const float h = H[(size_t)(i0 + t) * HL + f];
      hz = fmaf(h, S.scale[f], S.shift[f]);
      hm = mil_lrelu(h);
      if (drop != nullptr) hm *= drop[(size_t)(i0 + t) * HL + f] * (1.f / 0.75f);
    }
    S.hz[t][f] = hz;
    S.hm[t][f] = hm;
  }
  __syncthreads();
}

// per-tile forward of both MLPs; a[] = tanh(att pre-activation), u[] = buffer pre-activation (before lrelu)
__device__ __forceinline__ void head_tile_forward(const HeadSmem& S, int t, float a[HD], float u[HD], float raw[HK],
                                                  float& bval) {
#pragma unroll
  for (int d = 0; d < HD; ++d) { a[d] = S.b1[d]; u[d] = S.c1[d]; }
  for (int f = 0; f < HL; ++f) {
    const float hz = S.hz[t][f], hm = S.hm[t][f];
#pragma unroll
    for (int d4 = 0; d4 < HD / 4; ++d4) {
      const float4 w = *reinterpret_cast<const float4*>(&S.w1t[f][d4 * 4]);
      const float4 v = *reinterpret_cast<const float4*>(&S.v1t[f][d4 * 4]);
      a[d4 * 4 + 0] = fmaf(w.x, hz, a[d4 * 4 + 0]); a[d4 * 4 + 1] = fmaf(w.y, hz, a[d4 * 4 + 1]);
      a[d4 * 4 + 2] = fmaf(w.z, hz, a[d4 * 4 + 2]); a[d4 * 4 + 3] = fmaf(w.w, hz, a[d4 * 4 + 3]);
      u[d4 * 4 + 0] = fmaf(v.x, hm, u[d4 * 4 + 0]); u[d4 * 4 + 1] = fmaf(v.y, hm, u[d4 * 4 + 1]);
      u[d4 * 4 + 2] = fmaf(v.z, hm, u[d4 * 4 + 2]); u[d4 * 4 + 3] = fmaf(v.w, hm, u[d4 * 4 + 3]);
    }
  }
#pragma unroll
  for (int k = 0; k < HK; ++k) raw[k] = S.b2[k];
  bval = S.c2;
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    a[d] = tanhf(a[d]);
#pragma unroll
    for (int k = 0; k < HK; ++k) raw[k] = fmaf(S.w2[k][d], a[d], raw[k]);
    bval = fmaf(S.v2[d], mil_lrelu(u[d]), bval);
  }
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
__global__ void __launch_bounds__(HT)
head_scores_kernel(MilHeadParams P, const float* __restrict__ H, const float* __restrict__ drop, int n,
                   long long n_global, const double* __restrict__ stats, float* __restrict__ raw_o,
                   float* __restrict__ g_o, float* __restrict__ b_o, double* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HeadSmem& S = *reinterpret_cast<HeadSmem*>(smem_raw);
  __shared__ double s_red[(HT / 32) * MIL_HEAD_NSUMS];
  const int i0 = blockIdx.x * HT;
  head_load_tile(S, P, H, drop, n, n_global, stats, i0);
  const int t = threadIdx.x, i = i0 + t;
  double v[MIL_HEAD_NSUMS];
#pragma unroll
  for (int k = 0; k < MIL_HEAD_NSUMS; ++k) v[k] = 0.0;
  if (i < n) {
    float a[HD], u[HD], raw[HK], bval;
    head_tile_forward(S, t, a, u, raw, bval);
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
  head_scores_kernel<<<nblk, HT, sizeof(HeadSmem), s>>>(P, H, drop, n, n_global, stats, raw, g, b, part_ws);
  MIL_LAUNCH_OK();
  reduce_double_kernel<<<1, 32, 0, s>>>(part_ws, nblk, MIL_HEAD_NSUMS, sums);
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

__global__ void __launch_bounds__(HT)
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
  const int i0 = blockIdx.x * HT;
  head_load_tile(S, P, H, drop, n, n_global, stats, i0);
  const int t = threadIdx.x, i = i0 + t;
  float* rec = part + (size_t)blockIdx.x * HB_REC;
  const float gl = gloss ? gloss[0] : 1.f;

  float a[HD], u[HD], raw[HK], bval;
  float da[HD], dp[HD];
  float draw[HK] = {0.f, 0.f, 0.f}, dwm[HK] = {0.f, 0.f, 0.f}, db = 0.f;
  if (i < n) {
    head_tile_forward(S, t, a, u, raw, bval);  // a = tanh(.), u = pre-activation of the instance MLP
#pragma unroll
    for (int k = 0; k < HK; ++k) {
      const float dM = gl * scal[MIL_SC_DM + k], Sk = scal[MIL_SC_S + k], Mk = scal[MIL_SC_M + k];
      const float gk = g_i[(size_t)i * HK + k];
      db = fmaf(dM, gk / Sk, db);                     // dL/db_n = sum_k dM_k A_kn
      const float dg = dM * (bval - Mk) / Sk;         // L1-normalise backward (needs only S_k, M_k)
      const float sg = raw[k] > 20.f ? 1.f : sigmoid_f(raw[k]);  // softplus'
      draw[k] = dg * S.sneg[k] * sg;
      const float sn = S.sneg[k], sp = S.spos[k];
      dwm[k] = dg * (-10.f * sn * (1.f - sn) * softplus_f(raw[k]) + 10.f * sp * (1.f - sp));
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      float dt = 0.f;
#pragma unroll
      for (int k = 0; k < HK; ++k) dt = fmaf(S.w2[k][d], draw[k], dt);
      da[d] = dt * (1.f - a[d] * a[d]);
      dp[d] = db * S.v2[d] * mil_lrelu_grad(u[d]);
    }
  } else {
#pragma unroll
    for (int d = 0; d < HD; ++d) { a[d] = 0.f; u[d] = 0.f; da[d] = 0.f; dp[d] = 0.f; }
  }
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    s_da[t][d] = da[d];
    s_dp[t][d] = dp[d];
    s_t[t][d] = a[d];
    s_u[t][d] = (i < n) ? mil_lrelu(u[d]) : 0.f;
  }
  s_small[t][0] = draw[0]; s_small[t][1] = draw[1]; s_small[t][2] = draw[2]; s_small[t][3] = db;
  s_small[t][4] = dwm[0]; s_small[t][5] = dwm[1]; s_small[t][6] = dwm[2]; s_small[t][7] = 0.f;
  __syncthreads();

  // ---- parameter-gradient partials of this block (outer products over its HT tiles) ----
  for (int o = threadIdx.x; o < HD * HL; o += blockDim.x) {
    const int d = o / HL, f = o % HL;
    float w1 = 0.f, v1 = 0.f;
    for (int tt = 0; tt < HT; ++tt) {
      w1 = fmaf(s_da[tt][d], S.hz[tt][f], w1);
      v1 = fmaf(s_dp[tt][d], S.hm[tt][f], v1);
    }
    rec[HB_W1 + o] = w1;
    rec[HB_V1 + o] = v1;
  }
  for (int d = threadIdx.x; d < HD; d += blockDim.x) {
    float sb1 = 0.f, sc1 = 0.f, sv2 = 0.f, w2k[HK] = {0.f, 0.f, 0.f};
    for (int tt = 0; tt < HT; ++tt) {
      sb1 += s_da[tt][d];
      sc1 += s_dp[tt][d];
      sv2 = fmaf(s_small[tt][3], s_u[tt][d], sv2);
#pragma unroll
      for (int k = 0; k < HK; ++k) w2k[k] = fmaf(s_small[tt][k], s_t[tt][d], w2k[k]);
    }
    rec[HB_B1 + d] = sb1;
    rec[HB_C1 + d] = sc1;
    rec[HB_V2 + d] = sv2;
#pragma unroll
    for (int k = 0; k < HK; ++k) rec[HB_W2 + k * HD + d] = w2k[k];
  }
  if (threadIdx.x < 7) {
    float sacc = 0.f;
    for (int tt = 0; tt < HT; ++tt) sacc += s_small[tt][threadIdx.x];
    if (threadIdx.x < 3) rec[HB_B2 + threadIdx.x] = sacc;
    else if (threadIdx.x == 3) rec[HB_C2] = sacc;
    else rec[HB_WM + threadIdx.x - 4] = sacc;
  }
  __syncthreads();  // everyone is done reading hz / hm as forward activations

  // ---- input gradients: dHz = W1^T da  (kept for phase 5), instance path dH = lrelu'(H) * mask/0.75 * V1^T dp ----
  // xhat is recovered from hz: xhat = (hz - beta) / gamma is ill-defined for gamma = 0 -> recompute from H instead.

  for (int f = 0; f < HL; ++f) {
    float dz = 0.f, dm = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < HD / 4; ++d4) {
      const float4 w = *reinterpret_cast<const float4*>(&S.w1t[f][d4 * 4]);
      const float4 v = *reinterpret_cast<const float4*>(&S.v1t[f][d4 * 4]);
      dz = fmaf(w.x, da[d4 * 4 + 0], dz); dz = fmaf(w.y, da[d4 * 4 + 1], dz);
      dz = fmaf(w.z, da[d4 * 4 + 2], dz); dz = fmaf(w.w, da[d4 * 4 + 3], dz);
      dm = fmaf(v.x, dp[d4 * 4 + 0], dm); dm = fmaf(v.y, dp[d4 * 4 + 1], dm);
      dm = fmaf(v.z, dp[d4 * 4 + 2], dm); dm = fmaf(v.w, dp[d4 * 4 + 3], dm);
    }
    S.hz[t][f] = dz;  // own row only: no hazard with other threads
    S.hm[t][f] = dm;
  }
  __syncthreads();

  for (int idx = threadIdx.x; idx < HT * HL; idx += blockDim.x) {
    const int tt = idx / HL, f = idx % HL;
    if (i0 + tt < n) {
      const float h = H[(size_t)(i0 + tt) * HL + f];
      float dmi = S.hm[tt][f] * mil_lrelu_grad(h);
      if (drop != nullptr) dmi *= drop[(size_t)(i0 + tt) * HL + f] * (1.f / 0.75f);
      dHz_o[(size_t)(i0 + tt) * HL + f] = S.hz[tt][f];
      dHi_o[(size_t)(i0 + tt) * HL + f] = dmi;
    }
  }
  // local BN sums: sum_t dHz, sum_t dHz * xhat  (per feature; fixed order over the block's tiles)
  for (int f = threadIdx.x; f < HL; f += blockDim.x) {
    float sb = 0.f, sw = 0.f;
    for (int tt = 0; tt < HT; ++tt) {
      if (i0 + tt < n) {
        const float h = H[(size_t)(i0 + tt) * HL + f];
        const float xh = (h - S.mean[f]) * S.rstd[f];
        sb += S.hz[tt][f];
        sw = fmaf(S.hz[tt][f], xh, sw);
      }
    }
    rec[HB_BNW + f] = sw;
    rec[HB_BNB + f] = sb;
  }
}

// fixed-order reduction of the per-block records: parameter gradients (+=, fp32) and BN sums (double)
__global__ void head_bwd_reduce_kernel(const float* __restrict__ part, int nblk, MilHeadGrads G,
                                       double* __restrict__ bnsums) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= HB_REC) return;
  double acc = 0.0;
  for (int b = 0; b < nblk; ++b) acc += (double)part[(size_t)b * HB_REC + o];
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

size_t mil_head_bwd_partial_floats(int n) { return (size_t)mil_cdiv(n, HT) * HB_REC; }

int mil_launch_head_bwd_a(const MilHeadParams& P, const MilHeadGrads& G, const float* H, const float* drop, int n,
                          long long n_global, const double* stats, const float* raw, const float* g, const float* b,
                          const float* scal, const float* gloss, float* dHz, float* dHi, float* part_ws,
                          double* bnsums, cudaStream_t s) {
  const int nblk = (int)mil_cdiv(n, HT);
  const size_t smem = sizeof(HeadSmem) + (size_t)4 * HT * (HD + 1) * sizeof(float) + (size_t)HT * 8 * sizeof(float);
  MIL_SET_SMEM((head_bwd_a_kernel), (int)smem);
  head_bwd_a_kernel<<<nblk, HT, smem, s>>>(P, H, drop, n, n_global, stats, raw, g, b, scal, gloss, dHz, dHi, part_ws);
  MIL_LAUNCH_OK();
  head_bwd_reduce_kernel<<<(int)mil_cdiv(HB_REC, 128), 128, 0, s>>>(part_ws, nblk, G, bnsums);
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
