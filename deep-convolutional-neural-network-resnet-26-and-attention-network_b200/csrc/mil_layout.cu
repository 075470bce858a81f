// Layout helpers: weight packing, NCHW <-> PF8 conversion (used by the layer-wise parity tests), and the
// deterministic second-stage reductions of split-K weight-gradient partials.
#include <algorithm>

#include "mil_common.cuh"
#include "mil_conv_tc.cuh"

#include <map>
#include <mutex>
#include <utility>

int mil_raise_max_dynamic_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> high;
  int dev = 0;
  MIL_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  int& h = high[{func, dev}];
  if (bytes > h) {
    MIL_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    h = bytes;
  }
  return 0;
}

// ---- conv weight packing ---------------------------------------------------------------------------
// PyTorch layout w[cout][cin][ks][ks] (reference nnBlocks.py:160-168 nn.Conv2d) ->
//   normal     : wp[tap][cin_pad ][cout_pad]   (forward conv: contraction over cin)
//   transposed : wp[tap][cout_pad][cin_pad ]   (dgrad: contraction over cout)
// zero padded to multiples of 8 channels.
__global__ void pack_conv_w_kernel(const float* __restrict__ w, float* __restrict__ wp, int cout, int cin, int ks,
                                   int transposed) {
  const int cip = (cin + 7) / 8 * 8, cop = (cout + 7) / 8 * 8;
  const int taps = ks * ks;
  const int A = transposed ? cop : cip, B = transposed ? cip : cop;
  const int total = taps * A * B;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int b = i % B, a = (i / B) % A, t = i / (A * B);
    int co = transposed ? a : b, ci = transposed ? b : a;
    float v = 0.f;
    if (co < cout && ci < cin) v = w[((size_t)co * cin + ci) * taps + t];
    wp[i] = v;
  }
}
int mil_launch_pack_conv_w(const float* w, float* wp, int cout, int cin, int ks, int transposed, cudaStream_t s) {
  const int cip = (cin + 7) / 8 * 8, cop = (cout + 7) / 8 * 8;
  const int total = ks * ks * cip * cop;
  pack_conv_w_kernel<<<(int)mil_cdiv(total, 256), 256, 0, s>>>(w, wp, cout, cin, ks, transposed);
  MIL_LAUNCH_OK();
  return 0;
}

// ---- NCHW fp32 <-> PF8 -------------------------------------------------------------------------------
template <typename T>
__global__ void to_pf8_kernel(const float* __restrict__ src, T* __restrict__ dst, MilPF8 g) {
  // one thread per (chunk, flat pixel): writes 8 channels (zeros on pad pixels / pad channels)
  const long long total = (long long)g.cb * g.Q;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(i / g.Q);
    const long long q = i % g.Q;
    const int n = (int)(q / g.P);
    const int r = (int)(q % g.P);
    const int y = r / g.wp, x = r % g.wp;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cb * 8 + j;
      v[j] = (y < g.h && x < g.w && c < g.c) ? src[(((size_t)n * g.c + c) * g.h + y) * g.w + x] : 0.f;
    }
    mil_store8(dst + mil_pf8_off(g, cb, q), v);
  }
}
template <typename T>
__global__ void from_pf8_kernel(const T* __restrict__ src, float* __restrict__ dst, MilPF8 g) {
  const long long total = (long long)g.n * g.c * g.h * g.w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % g.w);
    const int y = (int)((i / g.w) % g.h);
    const int c = (int)((i / ((long long)g.w * g.h)) % g.c);
    const int n = (int)(i / ((long long)g.w * g.h * g.c));
    const long long q = (long long)n * g.P + (long long)y * g.wp + x;
    dst[i] = mil_to_float(src[mil_pf8_off(g, c >> 3, q) + (c & 7)]);
  }
}
int mil_launch_to_pf8(int dtype, const float* nchw, void* pf8, int n, int c, int h, int w, cudaStream_t s) {
  MilPF8 g = mil_pf8(n, c, h, w);
  const int blocks = (int)std::min<long long>(mil_cdiv((long long)g.cb * g.Q, 256), 148 * 32);
  if (dtype == MIL_BF16)
    to_pf8_kernel<<<blocks, 256, 0, s>>>(nchw, (__nv_bfloat16*)pf8, g);
  else
    to_pf8_kernel<<<blocks, 256, 0, s>>>(nchw, (float*)pf8, g);
  MIL_LAUNCH_OK();
  return 0;
}
int mil_launch_from_pf8(int dtype, const void* pf8, float* nchw, int n, int c, int h, int w, cudaStream_t s) {
  MilPF8 g = mil_pf8(n, c, h, w);
  const int blocks = (int)std::min<long long>(mil_cdiv((long long)n * c * h * w, 256), 148 * 32);
  if (dtype == MIL_BF16)
    from_pf8_kernel<<<blocks, 256, 0, s>>>((const __nv_bfloat16*)pf8, nchw, g);
  else
    from_pf8_kernel<<<blocks, 256, 0, s>>>((const float*)pf8, nchw, g);
  MIL_LAUNCH_OK();
  return 0;
}

// ---- deterministic second-stage reductions ----------------------------------------------------------
// out[i] += sum_b partial[b*stride + i]   (fixed summation order: bit-reproducible run to run)
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nblk, long long stride,
                                       float* __restrict__ out, long long count) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= count) return;
  float acc = 0.f;
  for (int b = 0; b < nblk; ++b) acc += partial[(size_t)b * stride + i];
  out[i] += acc;
}
int mil_launch_reduce_partials(const float* partial, int nblk, long long stride, float* out, long long count,
                               cudaStream_t s) {
  reduce_partials_kernel<<<(int)mil_cdiv(count, 128), 128, 0, s>>>(partial, nblk, stride, out, count);
  MIL_LAUNCH_OK();
  return 0;
}

// partial[b][tap][cin_pad][cout_pad] (+ [cout_pad] bias sums at the end of each record) ->
//   dw[cout][cin][tap] += ... ,  db[cout] += ...      (PyTorch parameter layout)
#define RCW_PARTS 8
__device__ __forceinline__ void reduce_conv_w_body(const float* __restrict__ partial, int nblk, long long stride,
                                                   float* __restrict__ dw, float* __restrict__ db, int cout, int cin, int ks) {
  // threadIdx.x walks the RECORD layout (coalesced reads of every partial record), threadIdx.y takes every
  // RCW_PARTS-th record; the parts are then summed in a fixed order (bit-reproducible) and scattered into the
  // parameter layout
  __shared__ float part[RCW_PARTS][32];
  const int cip = (cin + 7) / 8 * 8, cop = (cout + 7) / 8 * 8;
  const int taps = ks * ks;
  const int nrec = taps * cip * cop;
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int total = nrec + (db != nullptr ? cop : 0);
  float acc = 0.f;
  if (i < total) {
    // four records in flight per thread (the loop used to be a chain of dependent-latency global loads: ~7 us per launch,
    // 27 launches per step); the order of the additions is unchanged
    int b = threadIdx.y;
    for (; b + 3 * RCW_PARTS < nblk; b += 4 * RCW_PARTS) {
      const float v0 = partial[(size_t)b * stride + i], v1 = partial[(size_t)(b + RCW_PARTS) * stride + i];
      const float v2 = partial[(size_t)(b + 2 * RCW_PARTS) * stride + i], v3 = partial[(size_t)(b + 3 * RCW_PARTS) * stride + i];
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; b < nblk; b += RCW_PARTS) acc += partial[(size_t)b * stride + i];
  }
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y != 0 || i >= total) return;
#pragma unroll
  for (int k = 1; k < RCW_PARTS; ++k) acc += part[k][threadIdx.x];
  if (i < nrec) {
    const int co = i % cop, ci = (i / cop) % cip, t = i / (cop * cip);
    if (co < cout && ci < cin) dw[((size_t)co * cin + ci) * taps + t] += acc;
  } else {
    const int co = i - nrec;
    if (co < cout) db[co] += acc;
  }
}
__global__ void __launch_bounds__(32 * RCW_PARTS)
reduce_conv_w_kernel(const float* __restrict__ partial, int nblk, long long stride, float* __restrict__ dw,
                     float* __restrict__ db, int cout, int cin, int ks) {
  reduce_conv_w_body(partial, nblk, stride, dw, db, cout, cin, ks);
}
// ---- deferred reductions -------------------------------------------------------------------------------
// A backward pass makes 27 of these reductions, each a ~7 us latency-bound launch whatever the bag size (7 % of the step
// at the reference's live bag shape of ~500 tiles through the CNN).  While a batch is open on the calling thread,
// mil_launch_reduce_conv_w only RECORDS the job -- the weight-gradient kernels of a layer write their partial records one
// after the other into an arena instead of sharing one buffer -- and mil_reduce_batch_flush sums all recorded jobs in ONE
// launch (blockIdx.y = job).  Per output element the summation order is the same as in the immediate form.
#define MIL_RB_MAX 8
struct ReduceJob {
  const float* partial;
  float *dw, *db;
  long long stride;
  int nblk, cout, cin, ks;
};
struct ReduceTable {
  ReduceJob job[MIL_RB_MAX];
};
__global__ void __launch_bounds__(32 * RCW_PARTS)
reduce_conv_w_table_kernel(const __grid_constant__ ReduceTable t) {
  const ReduceJob& j = t.job[blockIdx.y];
  const int cop = (j.cout + 7) / 8 * 8, cip = (j.cin + 7) / 8 * 8;
  const int total = j.ks * j.ks * cip * cop + (j.db != nullptr ? cop : 0);
  if (blockIdx.x * 32 >= total) return;  // block-uniform: the grid is sized for the largest job
  reduce_conv_w_body(j.partial, j.nblk, j.stride, j.dw, j.db, j.cout, j.cin, j.ks);
}

namespace {
struct ReduceBatch {
  bool open = false;
  float* arena = nullptr;
  size_t arena_floats = 0, max_need = 0, used = 0;
  int count = 0, max_total = 0;
  ReduceTable table;
};
thread_local ReduceBatch g_rb;
}  // namespace

void mil_reduce_batch_begin(float* arena, size_t arena_floats, size_t max_need) {
  g_rb.open = arena != nullptr && arena_floats >= 2 * max_need;  // too small an arena: stay in the immediate form
  g_rb.arena = arena;
  g_rb.arena_floats = arena_floats;
  g_rb.max_need = max_need;
  g_rb.used = 0;
  g_rb.count = 0;
  g_rb.max_total = 0;
}
int mil_reduce_batch_flush(cudaStream_t s) {
  if (g_rb.count > 0) {
    reduce_conv_w_table_kernel<<<dim3((unsigned)mil_cdiv(g_rb.max_total, 32), g_rb.count), dim3(32, RCW_PARTS), 0, s>>>(g_rb.table);
    MIL_LAUNCH_OK();
  }
  g_rb.used = 0;
  g_rb.count = 0;
  g_rb.max_total = 0;
  return 0;
}
int mil_reduce_batch_end(cudaStream_t s) {
  const int rc = mil_reduce_batch_flush(s);
  g_rb.open = false;
  return rc;
}
float* mil_reduce_batch_cursor(float* fallback, cudaStream_t s) {
  if (!g_rb.open) return fallback;
  if (g_rb.arena_floats - g_rb.used < g_rb.max_need && mil_reduce_batch_flush(s) != 0) return nullptr;
  return g_rb.arena + g_rb.used;
}

int mil_launch_reduce_conv_w(const float* partial, int nblk, long long stride, float* dw, float* db, int cout,
                             int cin, int ks, cudaStream_t s) {
  const int cop = (cout + 7) / 8 * 8;
  const int total = ks * ks * ((cin + 7) / 8 * 8) * cop + cop;
  if (g_rb.open && partial == g_rb.arena + g_rb.used) {
    ReduceJob& j = g_rb.table.job[g_rb.count++];
    j.partial = partial; j.dw = dw; j.db = db; j.stride = stride; j.nblk = nblk; j.cout = cout; j.cin = cin; j.ks = ks;
    g_rb.max_total = std::max(g_rb.max_total, total);
    g_rb.used += (size_t)mil_rup((long long)nblk * stride, 64);
    if (g_rb.count == MIL_RB_MAX) return mil_reduce_batch_flush(s);
    return 0;
  }
  reduce_conv_w_kernel<<<(int)mil_cdiv(total, 32), dim3(32, RCW_PARTS), 0, s>>>(partial, nblk, stride, dw, db, cout, cin,
                                                                               ks);
  MIL_LAUNCH_OK();
  return 0;
}

// ---- Adam over the flat parameter buffer (reference: optim.Adam, gbm/classify_combined.py:519) ----
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long count, float step_size, float beta1, float beta2,
                                 float bc2_sqrt, float eps, float weight_decay) {
  const float w1 = 1.f - beta1, w2 = 1.f - beta2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i];
    float gi = g[i];
    if (weight_decay != 0.f) gi += weight_decay * pi;
    const float mi = m[i] + w1 * (gi - m[i]);
    const float vi = beta2 * v[i] + w2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}
// the same step with the six scalars read from DEVICE memory: hyper = {step_size, beta1, beta2, bc2_sqrt, eps,
// weight_decay}.  A captured CUDA graph replays this launch while the host refreshes `hyper` between replays
// (the bias corrections change with every step).
__global__ void adam_step_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                     float* __restrict__ v, long long count, const float* __restrict__ hyper) {
  const float step_size = hyper[0], beta1 = hyper[1], beta2 = hyper[2], bc2_sqrt = hyper[3], eps = hyper[4],
              weight_decay = hyper[5];
  const float w1 = 1.f - beta1, w2 = 1.f - beta2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i];
    float gi = g[i];
    if (weight_decay != 0.f) gi += weight_decay * pi;
    const float mi = m[i] + w1 * (gi - m[i]);
    const float vi = beta2 * v[i] + w2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}
int mil_launch_adam_step_dev(float* p, const float* g, float* m, float* v, long long count, const float* hyper,
                             cudaStream_t s) {
  const int blocks = (int)std::min<long long>(mil_cdiv(count, 256), 148 * 8);
  adam_step_dev_kernel<<<blocks, 256, 0, s>>>(p, g, m, v, count, hyper);
  MIL_LAUNCH_OK();
  return 0;
}
int mil_launch_adam_step(float* p, const float* g, float* m, float* v, long long count, float step_size, float beta1,
                         float beta2, float bc2_sqrt, float eps, float weight_decay, cudaStream_t s) {
  const int blocks = (int)std::min<long long>(mil_cdiv(count, 256), 148 * 8);
  adam_step_kernel<<<blocks, 256, 0, s>>>(p, g, m, v, count, step_size, beta1, beta2, bc2_sqrt, eps, weight_decay);
  MIL_LAUNCH_OK();
  return 0;
}

// ---- min-max scaling of an attention map (reference gbm/classify.py:209) ----------------------
__device__ __forceinline__ int mm_key(float f) {  // order-preserving float -> int
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float mm_val(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }
__global__ void minmax_init_kernel(int* mm) { mm[0] = 0x7FFFFFFF; mm[1] = (int)0x80000000; }
__global__ void minmax_reduce_kernel(const float* __restrict__ in, long long count, int* mm) {
  int lo = 0x7FFFFFFF, hi = (int)0x80000000;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = mm_key(in[i]);
    lo = min(lo, k);
    hi = max(hi, k);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
}
__global__ void minmax_scale_kernel(const float* __restrict__ in, float* __restrict__ out, long long count, int* mm) {
  const float lo = mm_val(mm[0]), hi = mm_val(mm[1]);
  const float inv = hi > lo ? 1.f / (hi - lo) : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = (in[i] - lo) * inv;
}
// (min, max) handed back as floats once every block of the scaling pass has read the integer keys
__global__ void minmax_finish_kernel(int* mm) {
  const float lo = mm_val(mm[0]), hi = mm_val(mm[1]);
  reinterpret_cast<float*>(mm)[0] = lo;
  reinterpret_cast<float*>(mm)[1] = hi;
}
int mil_launch_minmax_normalize(const float* in, float* out, long long count, float* minmax, cudaStream_t s) {
  int* mm = reinterpret_cast<int*>(minmax);
  const int blocks = (int)std::max<long long>(1, std::min<long long>(mil_cdiv(count, 256), 148 * 4));
  minmax_init_kernel<<<1, 1, 0, s>>>(mm);
  MIL_LAUNCH_OK();
  minmax_reduce_kernel<<<blocks, 256, 0, s>>>(in, count, mm);
  MIL_LAUNCH_OK();
  minmax_scale_kernel<<<blocks, 256, 0, s>>>(in, out, count, mm);
  MIL_LAUNCH_OK();
  minmax_finish_kernel<<<1, 1, 0, s>>>(mm);
  MIL_LAUNCH_OK();
  return 0;
}

// ---- zero-stuffing (bf16): out(n, 2y, 2x) = in(n, y, x), zero elsewhere ------------------------------------
// Turns the stride-2 convolutions' data / weight gradients into stride-1 problems the tcgen05 kernels handle:
// conv_transpose_s2(dz) == conv_transpose_s1(zero_stuff(dz)),  wgrad_s2(x, dz) == wgrad_s1(x, zero_stuff(dz)).
__global__ void upsample2_kernel(const uint4* __restrict__ in, MilPF8 gin, uint4* __restrict__ out, MilPF8 gout) {
  const long long total = (long long)gout.cb * gout.Q;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(i / gout.Q);
    const long long q = i - (long long)cb * gout.Q;
    const int n = (int)(q / gout.P);
    const int r = (int)(q - (long long)n * gout.P);
    const int y = r / gout.wp, x = r - y * gout.wp;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y < gout.h && x < gout.w && !(y & 1) && !(x & 1))
      v = in[(size_t)cb * gin.PS + gin.G + (size_t)n * gin.P + (size_t)(y >> 1) * gin.wp + (x >> 1)];
    out[(size_t)cb * gout.PS + gout.G + q] = v;
  }
}
// ---- phase split (bf16): out plane (phase * cb + c), pixel (Y, X) = in plane c at (2Y + a, 2X + b), phase = 2a + b ----
// grid = (image, pixel block of the OUTPUT map, input chunk): a thread reads one 2x2 block, writes four planes
__global__ void __launch_bounds__(256) split2_kernel(const uint4* __restrict__ in, MilPF8 gin, uint4* __restrict__ out,
                                                     MilPF8 gout) {
  const int n = blockIdx.x, c = blockIdx.z;
  const int r = blockIdx.y * blockDim.x + threadIdx.x;
  if (r >= (int)gout.P) return;
  const int Y = r / gout.wp, X = r - Y * gout.wp;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  uint4 v[4] = {zero, zero, zero, zero};
  if (Y < gout.h && X < gout.w) {
    const uint4* p = in + (size_t)c * gin.PS + gin.G + (size_t)n * gin.P + (size_t)(2 * Y) * gin.wp + 2 * X;
    const bool a1 = 2 * Y + 1 < gin.h, b1 = 2 * X + 1 < gin.w;
    v[0] = p[0];
    if (b1) v[1] = p[1];
    if (a1) v[2] = p[gin.wp];
    if (a1 && b1) v[3] = p[gin.wp + 1];
  }
  const size_t o = gout.G + (size_t)n * gout.P + r;
#pragma unroll
  for (int ph = 0; ph < 4; ++ph) out[(size_t)(ph * gin.cb + c) * gout.PS + o] = v[ph];
}
int mil_launch_split2(const void* in, const MilPF8& gin, void* out, const MilPF8& gout, cudaStream_t s) {
  MIL_REQUIRE(gin.n == gout.n && gout.cb == 4 * gin.cb && gout.h == (gin.h - 1) / 2 + 1 && gout.w == (gin.w - 1) / 2 + 1,
              "split2: geometry mismatch");
  split2_kernel<<<dim3(gin.n, (unsigned)mil_cdiv(gout.P, 256), gin.cb), 256, 0, s>>>((const uint4*)in, gin, (uint4*)out,
                                                                                    gout);
  MIL_LAUNCH_OK();
  return 0;
}

// ---- even-position subsampling (bf16): out(n, y, x) = in(n, 2y, 2x) -------------------------------------------
__global__ void subsample2_kernel(const uint4* __restrict__ in, MilPF8 gin, uint4* __restrict__ out, MilPF8 gout) {
  const long long total = (long long)gout.cb * gout.Q;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(i / gout.Q);
    const long long q = i - (long long)cb * gout.Q;
    const int n = (int)(q / gout.P);
    const int r = (int)(q - (long long)n * gout.P);
    const int y = r / gout.wp, x = r - y * gout.wp;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y < gout.h && x < gout.w)
      v = in[(size_t)cb * gin.PS + gin.G + (size_t)n * gin.P + (size_t)(2 * y) * gin.wp + 2 * x];
    out[(size_t)cb * gout.PS + gout.G + q] = v;
  }
}
int mil_launch_subsample2(const void* in, const MilPF8& gin, void* out, const MilPF8& gout, cudaStream_t s) {
  MIL_REQUIRE(gin.n == gout.n && gin.cb == gout.cb && gout.h == (gin.h - 1) / 2 + 1 && gout.w == (gin.w - 1) / 2 + 1,
              "subsample2: geometry mismatch");
  const int blocks = (int)std::min<long long>(mil_cdiv((long long)gout.cb * gout.Q, 256), 148 * 16);
  subsample2_kernel<<<blocks, 256, 0, s>>>((const uint4*)in, gin, (uint4*)out, gout);
  MIL_LAUNCH_OK();
  return 0;
}

int mil_launch_upsample2(const void* in, const MilPF8& gin, void* out, const MilPF8& gout, cudaStream_t s) {
  MIL_REQUIRE(gin.n == gout.n && gin.cb == gout.cb && gin.h == (gout.h - 1) / 2 + 1 && gin.w == (gout.w - 1) / 2 + 1,
              "upsample2: geometry mismatch");
  const int blocks = (int)std::min<long long>(mil_cdiv((long long)gout.cb * gout.Q, 256), 148 * 16);
  upsample2_kernel<<<blocks, 256, 0, s>>>((const uint4*)in, gin, (uint4*)out, gout);
  MIL_LAUNCH_OK();
  return 0;
}
