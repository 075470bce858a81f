// Shared definitions for the B200 (sm_100a) ResNet-26 + attention-MIL kernels.
//
// Activation layout in HBM -- "PF8" (padded-flat, 8-channel chunks):
//
//     [channel chunk cb][ lead guard G | image 0 | image 1 | ... | image n-1 | tail guard GT ][8 channels]
//
// Every image is stored as (H+1) x (W+1) pixels: one zero column on the right of every row and one
// zero row below every image.  With that single shared halo, "pixel (y-1, x-1)" of a 3x3 window is
// simply "flat pixel index - (W+1) - 1" -- the left neighbour of column 0 is the zero column of the
// row above, the row above row 0 is the zero row of the previous image (or the lead guard).  So a
// convolution tap is a CONSTANT shift of the flat pixel index, no bounds checks anywhere, and an
// implicit-GEMM M-tile is any run of 128 consecutive flat pixels (tiles may straddle images).
// A pixel's 8-channel chunk is one 16-byte unit in bf16: the unit a bulk-TMA copy moves, one row of
// a tcgen05 "core matrix" (8 pixels x 16 B = 128 contiguous bytes), and one vector load/store per
// thread.  Channel counts 20/40/60/80 are stored as 24/40/64/80; pad channels, pad pixels and
// guards are always exactly zero (every kernel that writes a PF8 tensor writes zeros there).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define MIL_SLOPE 0.1f

enum MilDtype { MIL_F32 = 0, MIL_BF16 = 1 };
enum MilEpilogue {
  MIL_EPI_FWD = 0,    // out = lrelu(acc + bias [+ res])
  MIL_EPI_DGRAD = 1,  // out = (acc [+ res]) * lrelu'(act)
  MIL_EPI_PLAIN = 2   // out = acc [+ bias] [+ res]
};

void mil_set_error(const char* fmt, ...);
#define MIL_CHECK_CUDA(expr)                                                               \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      mil_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)
// one per kernel launch: counts the launch (mil_kernel_launch_count) and surfaces launch errors
void mil_count_launch();
#define MIL_LAUNCH_OK()                 \
  do {                                  \
    mil_count_launch();                 \
    MIL_CHECK_CUDA(cudaGetLastError()); \
  } while (0)
#define MIL_REQUIRE(cond, ...)    \
  do {                            \
    if (!(cond)) {                \
      mil_set_error(__VA_ARGS__); \
      return 2;                   \
    }                             \
  } while (0)
#define MIL_TRY(expr)         \
  do {                        \
    int _rc = (expr);         \
    if (_rc != 0) return _rc; \
  } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// A step is ~140 dependent launches; at the reference's live bag shape (~500 tiles through the CNN) every one of them
// costs ~6 us of launch latency + prologue (barrier init, TMEM allocation, shared-memory set-up) + drain against 5-40 us of
// work.  Kernels launched through MIL_LAUNCH_PDL may START while their predecessor in the stream is still running: the
// prologue overlaps the predecessor's tail, and mil_pdl_wait() blocks until the predecessor has completed and its memory
// is visible.  Rules that keep this equivalent to stream order:
//   * a kernel touches NO global memory before mil_pdl_wait() (shared memory, TMEM, barriers, kernel parameters only);
//   * it calls mil_pdl_trigger() only AFTER its own wait, so a dependent can overlap its immediate predecessor only -- by
//     then everything older has completed -- and LATE: when its producer warp has issued the CTA's last loads.  (Triggering
//     at the start parks the dependent's CTAs next to the running kernel for its whole duration wherever both fit on an
//     SM: measured 2.5 % slower on 4096-tile bags);
//   * only the persistent tcgen05 kernels are launched this way: small many-CTA kernels (layout passes, reductions) waiting
//     next to a running convolution slowed it down more than their launch latency is worth;
//   * kernels launched the ordinary way (torch's, NCCL's, ours) remain fully ordered; for them both calls are no-ops.
__device__ __forceinline__ void mil_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void mil_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- runtime switches (mil_set_option / mil_get_option of the C ABI; the environment variables of the same
// meaning -- MIL_B200_DISABLE_TC, MIL_B200_STEM_UNFUSED, ... -- only give the initial values) -------------------
enum MilOpt {
  MIL_OPT_DISABLE_TC = 0,    // 1: CUDA-core kernels only (cross-check of the tcgen05 path, same rounding points)
  MIL_OPT_STEM_UNFUSED = 1,  // 1: stem as separate conv / pool / unpool / wgrad kernels (cross-check of the fused ones)
  MIL_OPT_NO_PDL = 2,        // 1: plain stream-ordered launches (cross-check of the programmatic dependent launches below)
  MIL_OPT_COUNT
};
int mil_opt(int id);
int mil_opt_set(const char* name, int value);  // 0 = ok
int mil_opt_get(const char* name, int* value);

// The extractor drivers switch PDL off for large bags (measured: -7 % step time at ~500 tiles through the CNN, neutral
// around 2500, +1.5 % at 4096 tiles x 224^2, tools/pdl_ab.py): thread-local, on by default (layer-level entry points).
bool mil_pdl_allowed();
void mil_pdl_allow(bool on);
struct MilPdlScope {
  bool prev;
  // pixels = tiles x side^2 of the bag this call works on
  explicit MilPdlScope(long long pixels) : prev(mil_pdl_allowed()) { mil_pdl_allow(pixels <= 1536LL * 224 * 224); }
  ~MilPdlScope() { mil_pdl_allow(prev); }
};

template <typename... KArgs, typename... Args>
static inline cudaError_t mil_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                         Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = (mil_opt(MIL_OPT_NO_PDL) || !mil_pdl_allowed()) ? 0 : 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// launch + count + error check, in place of kernel<<<grid, block, smem, s>>>(args...); MIL_LAUNCH_OK();
#define MIL_LAUNCH_PDL(kernel, grid, block, smem, s, ...)                                   \
  do {                                                                                      \
    MIL_CHECK_CUDA(mil_launch_pdl(kernel, dim3(grid), dim3(block), smem, s, __VA_ARGS__)); \
    mil_count_launch();                                                                     \
  } while (0)

__host__ __device__ static inline long long mil_cdiv(long long a, long long b) { return (a + b - 1) / b; }
__host__ __device__ static inline long long mil_rup(long long a, long long b) { return mil_cdiv(a, b) * b; }
static inline size_t mil_esize(int dtype) { return dtype == MIL_BF16 ? 2 : 4; }

__host__ __device__ __forceinline__ float mil_lrelu(float v) { return v > 0.f ? v : MIL_SLOPE * v; }
__host__ __device__ __forceinline__ float mil_lrelu_grad(float act) { return act > 0.f ? 1.f : MIL_SLOPE; }

// ---- PF8 tensor geometry -----------------------------------------------------------------------
struct MilPF8 {
  int n, c, cb;      // images, channels, 8-channel chunks
  int h, w, hp, wp;  // spatial size and padded spatial size (h+pad, w+pad); pad = 1 unless stated
  long long P;       // pixels per image plane = hp*wp
  long long Q;       // flat pixels = n*P
  long long G, GT;   // lead / tail guard, in pixels
  long long PS;      // chunk-plane stride in pixels = G + Q + GT
};
#define MIL_TILE_M 128  // flat pixels per implicit-GEMM tile
// pad = width of the shared zero halo (1 for the 3x3 layers; 2 for the stem's space-to-depth 4x4 window)
static inline MilPF8 mil_pf8p(int n, int c, int h, int w, int pad) {
  MilPF8 t;
  t.n = n; t.c = c; t.cb = (c + 7) / 8;
  t.h = h; t.w = w; t.hp = h + pad; t.wp = w + pad;
  t.P = (long long)t.hp * t.wp;
  t.Q = (long long)n * t.P;
  t.G = mil_rup((long long)pad * (t.wp + 1), 8);
  t.GT = mil_rup(MIL_TILE_M + (long long)pad * (t.wp + 1), 8) + 8;
  t.PS = mil_rup(t.G + t.Q + t.GT, 8);
  return t;
}
static inline MilPF8 mil_pf8(int n, int c, int h, int w) { return mil_pf8p(n, c, h, w, 1); }
static inline size_t mil_pf8_bytes(const MilPF8& t, int dtype) {
  return (size_t)t.cb * (size_t)t.PS * 8 * mil_esize(dtype);
}
// element offset of (chunk cb, flat pixel q, lane 0)
__host__ __device__ __forceinline__ long long mil_pf8_off(const MilPF8& t, int cb, long long q) {
  return ((long long)cb * t.PS + t.G + q) * 8;
}

// ---- 8-channel chunk load/store ----------------------------------------------------------------
__device__ __forceinline__ void mil_load8(const float* p, float v[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void mil_load8(const __nv_bfloat16* p, float v[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void mil_store8(float* p, const float v[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void mil_store8(__nv_bfloat16* p, const float v[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}
// bit j = (bf16 value j of the chunk > 0): the sign-mask byte of one 8-channel chunk
__device__ __forceinline__ uint32_t mil_positive_bits(const uint4& pk) {
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
  const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&pk);
  const uint32_t m0 = __hgt2_mask(hp[0], z), m1 = __hgt2_mask(hp[1], z), m2 = __hgt2_mask(hp[2], z),
                 m3 = __hgt2_mask(hp[3], z);
  const uint32_t t0 = __byte_perm(m0, m1, 0x7531) & 0x80808080u;
  const uint32_t t1 = __byte_perm(m2, m3, 0x7531) & 0x80808080u;
  return ((t0 * 0x00204081u) >> 28) | (((t1 * 0x00204081u) >> 28) << 4);
}
__device__ __forceinline__ float mil_to_float(float v) { return v; }
__device__ __forceinline__ float mil_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void mil_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void mil_from_float(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- geometry of the extractor (reference gbm/model.py:24-32) ------------------------------------
struct MilGeom {
  int side;  // input tile side S
  int hc;    // conv1 output side  (7x7, stride 2, pad 3)
  int h[4];  // h[0] = after maxpool (= layer1 side), h[1..3] = layer2..4 sides
};
static inline MilGeom mil_geom(int side) {
  MilGeom g;
  g.side = side;
  g.hc = (side - 1) / 2 + 1;
  g.h[0] = (g.hc - 1) / 2 + 1;
  for (int i = 1; i < 4; ++i) g.h[i] = (g.h[i - 1] - 1) / 2 + 1;
  return g;
}
static const int kMilWidths[4] = {20, 40, 60, 80};

// ---- launchers implemented in the .cu files (host side, C++ linkage) -----------------------------
// mil_layout.cu
int mil_launch_pack_conv_w(const float* w, float* wp, int cout, int cin, int ks, int transposed, cudaStream_t s);
int mil_launch_to_pf8(int dtype, const float* nchw, void* pf8, int n, int c, int h, int w, cudaStream_t s);
int mil_launch_from_pf8(int dtype, const void* pf8, float* nchw, int n, int c, int h, int w, cudaStream_t s);
// cudaFuncAttributeMaxDynamicSharedMemorySize is a high-water mark: raise it only when a launch needs more than any
// earlier one did (the call costs microseconds, and a training step makes ~140 launches).  Keyed by (kernel, device).
int mil_raise_max_dynamic_smem(const void* func, int bytes);
#define MIL_SET_SMEM(kernel, bytes) MIL_TRY(mil_raise_max_dynamic_smem(reinterpret_cast<const void*>(kernel), (int)(bytes)))

int mil_launch_minmax_normalize(const float* in, float* out, long long count, float* minmax, cudaStream_t s);
int mil_launch_adam_step(float* p, const float* g, float* m, float* v, long long count, float step_size, float beta1,
                         float beta2, float bc2_sqrt, float eps, float weight_decay, cudaStream_t s);
int mil_launch_adam_step_dev(float* p, const float* g, float* m, float* v, long long count, const float* hyper,
                             cudaStream_t s);
// mil_ingest.cu
int mil_launch_ingest_u8(const uint8_t* rois, int T, int R, const int* crops, int pad, const uint8_t* flips, int S,
                         const int* bounds, const int* coef, int ksize, const int* bounds_host, uint8_t* out,
                         cudaStream_t s);
int mil_launch_reduce_partials(const float* partial, int nblk, long long stride, float* out, long long count,
                               cudaStream_t s);
int mil_launch_reduce_conv_w(const float* partial, int nblk, long long stride, float* dw, float* db, int cout,
                             int cin, int ks, cudaStream_t s);
// deferred form (mil_layout.cu): between begin and end, a weight gradient that put its partial records at
// mil_reduce_batch_cursor() has its reduction recorded instead of launched; flush sums all recorded jobs in one launch.
// max_need = the largest partial-record area one weight gradient may write (floats).
void mil_reduce_batch_begin(float* arena, size_t arena_floats, size_t max_need);
float* mil_reduce_batch_cursor(float* fallback, cudaStream_t s);  // nullptr: a flush failed (mil_last_error)
int mil_reduce_batch_flush(cudaStream_t s);
int mil_reduce_batch_end(cudaStream_t s);
// mil_conv_direct.cu
int mil_launch_conv_direct(int dtype, int transposed, const void* x, const MilPF8& gi, const float* wp,
                           const float* bias, const void* res, const void* act, void* out, const MilPF8& go, int ks,
                           int stride, int epi, cudaStream_t s);
int mil_wgrad_direct_blocks(const MilPF8& go);
size_t mil_wgrad_direct_partial_floats(const MilPF8& gi, const MilPF8& go, int ks);
int mil_launch_wgrad_direct(int dtype, const void* x, const MilPF8& gi, const void* dz, const MilPF8& go,
                            float* partial, float* dw, float* db, int ks, int stride, cudaStream_t s);
// mil_stem.cu
int mil_launch_stem_fwd(int dtype, const float* x, const int* idx, int n, int side, const float* w, const float* b,
                        void* pooled, const MilPF8& gp, uint8_t* argmax, cudaStream_t s);
size_t mil_stem_bwd_partial_floats();
int mil_launch_stem_bwd(int dtype, const float* x, const int* idx, int n, int side, const void* g, const MilPF8& gp,
                        const uint8_t* argmax, float* partial, float* dw, float* db, cudaStream_t s);
// mil_tail.cu
int mil_launch_tail_fwd(int dtype, const void* y4, const MilPF8& g4, const float* wfc, float* avg, float* H,
                        cudaStream_t s);
size_t mil_tail_bwd_partial_floats();
int mil_launch_tail_bwd(int dtype, const void* y4, const MilPF8& g4, const float* wfc, const float* avg,
                        const float* dH, void* dz4, float* partial, float* dwfc, cudaStream_t s);
