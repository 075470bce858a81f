// extern "C" entry points of libmil_b200.so (declared in include/mil_b200.h).  Thin: argument checks,
// plan construction, launch sequences.  Errors never cross the boundary as exceptions.
#include <atomic>
#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <exception>

#include "../../include/mil_b200.h"
#include "mil_extractor.cuh"
#include "mil_head.cuh"
#include "mil_wide.cuh"
#include "mil_wide_net.cuh"

static thread_local char g_err[1024] = "";
void mil_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void mil_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

#define MIL_API_BEGIN try {
#define MIL_API_END                                        \
  }                                                        \
  catch (const std::exception& e) {                        \
    mil_set_error("internal C++ exception: %s", e.what()); \
    return 3;                                              \
  }                                                        \
  catch (...) {                                            \
    mil_set_error("internal C++ exception");               \
    return 3;                                              \
  }

static int require_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    mil_set_error("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    return 4;
  }
  static thread_local int checked_dev = -1;
  if (checked_dev != dev) {
    cudaDeviceProp p;
    MIL_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
    MIL_REQUIRE(p.major == 10, "device %d is sm_%d%d; libmil_b200 is built for sm_100a only", dev, p.major, p.minor);
    checked_dev = dev;
  }
  return 0;
}

static MilHeadParams head_params(const void* const* params) {
  MilHeadParams P;
  auto f = [&](const char* n) { return reinterpret_cast<const float*>(params[mil_param_index(n)]); };
  P.weight_mask = f("weight_mask");
  P.bn_w = f("context.bn.weight");
  P.bn_b = f("context.bn.bias");
  P.att_w1 = f("attention.lin1.weight");
  P.att_b1 = f("attention.lin1.bias");
  P.att_w2 = f("attention.lin2.weight");
  P.att_b2 = f("attention.lin2.bias");
  P.buf_w1 = f("buffer.lin1.weight");
  P.buf_b1 = f("buffer.lin1.bias");
  P.buf_w2 = f("buffer.classifier.weight");
  P.buf_b2 = f("buffer.classifier.bias");
  return P;
}
static MilHeadGrads head_grads(float* grads) {
  MilHeadGrads G;
  const auto& t = mil_param_table();
  auto f = [&](const char* n) { return grads + t[mil_param_index(n)].offset; };
  G.weight_mask = f("weight_mask");
  G.bn_w = f("context.bn.weight");
  G.bn_b = f("context.bn.bias");
  G.att_w1 = f("attention.lin1.weight");
  G.att_b1 = f("attention.lin1.bias");
  G.att_w2 = f("attention.lin2.weight");
  G.att_b2 = f("attention.lin2.bias");
  G.buf_w1 = f("buffer.lin1.weight");
  G.buf_b1 = f("buffer.lin1.bias");
  G.buf_w2 = f("buffer.classifier.weight");
  G.buf_b2 = f("buffer.classifier.bias");
  return G;
}

extern "C" {

int mil_abi_version(void) { return MIL_ABI_VERSION; }
const char* mil_last_error(void) { return g_err; }
long long mil_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int mil_param_count(void) { return (int)mil_param_table().size(); }
const char* mil_param_name(int i) {
  const auto& t = mil_param_table();
  return (i >= 0 && i < (int)t.size()) ? t[i].name.c_str() : nullptr;
}
int mil_param_shape(int i, int* ndim, long long shape4[4]) {
  const auto& t = mil_param_table();
  MIL_REQUIRE(i >= 0 && i < (int)t.size(), "mil_param_shape: index %d out of range", i);
  *ndim = t[i].ndim;
  for (int k = 0; k < 4; ++k) shape4[k] = t[i].shape[k];
  return 0;
}
long long mil_param_offset(int i) {
  const auto& t = mil_param_table();
  return (i >= 0 && i < (int)t.size()) ? t[i].offset : -1;
}
long long mil_param_total(void) {
  const auto& t = mil_param_table();
  return t.back().offset + t.back().numel;
}

int mil_set_option(const char* name, int value) { return mil_opt_set(name, value); }
int mil_get_option(const char* name, int* value) {
  MIL_REQUIRE(value != nullptr, "mil_get_option: null pointer argument");
  return mil_opt_get(name, value);
}

// ---- extractor ---------------------------------------------------------------------------------------
size_t mil_extractor_workspace_bytes(int n_tiles, int side, int dtype) {
  try {
    MilPlan pl;
    if (mil_make_plan(n_tiles, side, dtype, &pl) != 0) return 0;
    return pl.total_bytes;
  } catch (...) {
    mil_set_error("internal C++ exception");
    return 0;
  }
}

size_t mil_extractor_infer_workspace_bytes(int n_tiles, int side, int dtype) {
  try {
    MilPlan pl;
    if (mil_make_plan(n_tiles, side, dtype, &pl, true) != 0) return 0;
    return pl.total_bytes;
  } catch (...) {
    mil_set_error("internal C++ exception");
    return 0;
  }
}

int mil_extractor_infer(const void* const* params, const void* bag, int bag_is_u8, const int32_t* idx, int n_tiles,
                        int side, int dtype, void* ws, size_t ws_bytes, float* H, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(params && bag && ws && H, "mil_extractor_infer: null pointer argument");
  MilPlan pl;
  MIL_TRY(mil_make_plan(n_tiles, side, dtype, &pl, true));
  MIL_REQUIRE(ws_bytes >= pl.total_bytes, "mil_extractor_infer: workspace too small (%zu < %zu)", ws_bytes,
              pl.total_bytes);
  return mil_extractor_forward_impl(params, bag, bag_is_u8 ? 1 : 0, idx, pl, ws, H, (cudaStream_t)stream);
  MIL_API_END
}

int mil_extractor_read_activation(int n_tiles, int side, int dtype, const void* ws, int which, float* nchw,
                                  void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MilPlan pl;
  MIL_TRY(mil_make_plan(n_tiles, side, dtype, &pl));
  MIL_REQUIRE(which >= -1 && which < 24, "mil_extractor_read_activation: which=%d out of range", which);
  const char* base = reinterpret_cast<const char*>(ws);
  if (which < 0) {
    const MilPF8& g = pl.g[0];
    return mil_launch_from_pf8(dtype, base + pl.off_pooled, nchw, g.n, g.c, g.h, g.w, (cudaStream_t)stream);
  }
  const int lb = which / 2;
  const MilPF8& g = pl.g[lb / 3];
  const size_t off = (which & 1) ? pl.off_y[lb] : pl.off_h[lb];
  return mil_launch_from_pf8(dtype, base + off, nchw, g.n, g.c, g.h, g.w, (cudaStream_t)stream);
  MIL_API_END
}

int mil_debug_dump_gradient(int layer, int block, int which, float* nchw) {
  mil_debug_request_dump(layer, block, which, nchw);
  return 0;
}

int mil_extractor_forward(const void* const* params, const float* bag, const int32_t* idx, int n_tiles, int side,
                          int dtype, void* ws, size_t ws_bytes, float* H, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(params && bag && ws && H, "mil_extractor_forward: null pointer argument");
  MilPlan pl;
  MIL_TRY(mil_make_plan(n_tiles, side, dtype, &pl));
  MIL_REQUIRE(ws_bytes >= pl.total_bytes, "mil_extractor_forward: workspace too small (%zu < %zu)", ws_bytes,
              pl.total_bytes);
  return mil_extractor_forward_impl(params, bag, 0, idx, pl, ws, H, (cudaStream_t)stream);
  MIL_API_END
}

int mil_extractor_forward_u8(const void* const* params, const uint8_t* bag, const int32_t* idx, int n_tiles, int side,
                             int dtype, void* ws, size_t ws_bytes, float* H, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(params && bag && ws && H, "mil_extractor_forward_u8: null pointer argument");
  MilPlan pl;
  MIL_TRY(mil_make_plan(n_tiles, side, dtype, &pl));
  MIL_REQUIRE(ws_bytes >= pl.total_bytes, "mil_extractor_forward_u8: workspace too small (%zu < %zu)", ws_bytes,
              pl.total_bytes);
  return mil_extractor_forward_impl(params, bag, 1, idx, pl, ws, H, (cudaStream_t)stream);
  MIL_API_END
}

int mil_extractor_backward_staged(const void* const* params, const float* bag, const int32_t* idx, int n_tiles,
                                  int side, int dtype, void* ws, size_t ws_bytes, const float* dH, float* grads_flat,
                                  void* const* layer_events_host, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(params && bag && ws && dH && grads_flat, "mil_extractor_backward_staged: null pointer argument");
  MilPlan pl;
  MIL_TRY(mil_make_plan(n_tiles, side, dtype, &pl));
  MIL_REQUIRE(ws_bytes >= pl.total_bytes, "mil_extractor_backward_staged: workspace too small (%zu < %zu)", ws_bytes,
              pl.total_bytes);
  return mil_extractor_backward_impl(params, bag, idx, pl, ws, dH, grads_flat, (cudaStream_t)stream,
                                     reinterpret_cast<const cudaEvent_t*>(layer_events_host));
  MIL_API_END
}

int mil_extractor_backward(const void* const* params, const float* bag, const int32_t* idx, int n_tiles, int side,
                           int dtype, void* ws, size_t ws_bytes, const float* dH, float* grads_flat, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(params && bag && ws && dH && grads_flat, "mil_extractor_backward: null pointer argument");
  MilPlan pl;
  MIL_TRY(mil_make_plan(n_tiles, side, dtype, &pl));
  MIL_REQUIRE(ws_bytes >= pl.total_bytes, "mil_extractor_backward: workspace too small (%zu < %zu)", ws_bytes,
              pl.total_bytes);
  return mil_extractor_backward_impl(params, bag, idx, pl, ws, dH, grads_flat, (cudaStream_t)stream);
  MIL_API_END
}

// ---- head ----------------------------------------------------------------------------------------------
size_t mil_head_workspace_bytes(int n) {
  if (n < 1) n = 1;
  return mil_head_part_doubles(n) * sizeof(double) + mil_head_bwd_partial_floats(n) * sizeof(float) + 256;
}
static int head_check(int n, long long n_global, size_t ws_bytes) {
  MIL_TRY(require_device());
  MIL_REQUIRE(n >= 0 && n_global >= n, "head: bad tile counts n=%d n_global=%lld", n, n_global);
  // nn.BatchNorm1d(track_running_stats=False) refuses a single row in train AND eval (reference behaviour;
  // the Python binding raises ValueError before reaching this point)
  MIL_REQUIRE(n_global >= 2, "Expected more than 1 value per channel when training, got input size (%lld, 80)",
              n_global);
  MIL_REQUIRE(ws_bytes >= mil_head_workspace_bytes(n), "head: workspace too small (%zu < %zu)", ws_bytes,
              mil_head_workspace_bytes(n));
  return 0;
}

int mil_head_stats(const float* H, int n, void* ws, size_t ws_bytes, double* stats, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(H && ws && stats, "mil_head_stats: null pointer argument");
  MIL_REQUIRE(ws_bytes >= mil_head_workspace_bytes(n), "mil_head_stats: workspace too small");
  return mil_launch_head_stats(H, n, (double*)ws, stats, (cudaStream_t)stream);
  MIL_API_END
}

int mil_head_scores(const void* const* params, const float* H, const float* drop, int n, long long n_global,
                    const double* stats, float* raw, float* g, float* b, void* ws, size_t ws_bytes, double* sums,
                    void* stream) {
  MIL_API_BEGIN
  MIL_TRY(head_check(n, n_global, ws_bytes));
  MIL_REQUIRE(params && H && stats && raw && g && b && ws && sums, "mil_head_scores: null pointer argument");
  return mil_launch_head_scores(head_params(params), H, drop, n, n_global, stats, raw, g, b, (double*)ws, sums,
                                (cudaStream_t)stream);
  MIL_API_END
}

int mil_head_finalize(const double* sums, const double* stats, long long n_global, const long long* Y,
                      const float* class_w, int n, const float* g, const float* b, float* A, float* wroi,
                      float* scal, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(sums && stats && Y && g && b && A && wroi && scal, "mil_head_finalize: null pointer argument");
  return mil_launch_head_finalize(sums, stats, n_global, Y, class_w, n, g, b, A, wroi, scal, (cudaStream_t)stream);
  MIL_API_END
}

int mil_head_backward_a(const void* const* params, const float* H, const float* drop, int n, long long n_global,
                        const double* stats, const float* raw, const float* g, const float* b, const float* scal,
                        const float* gloss, float* dHz, float* dHi, float* grads_flat, void* ws, size_t ws_bytes,
                        double* bnsums, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(head_check(n, n_global, ws_bytes));
  MIL_REQUIRE(params && H && stats && raw && g && b && scal && dHz && dHi && grads_flat && ws && bnsums,
              "mil_head_backward_a: null pointer argument");
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + mil_head_part_doubles(n) * sizeof(double));
  return mil_launch_head_bwd_a(head_params(params), head_grads(grads_flat), H, drop, n, n_global, stats, raw, g, b,
                               scal, gloss, dHz, dHi, part, bnsums, (cudaStream_t)stream);
  MIL_API_END
}

int mil_head_backward_b(const void* const* params, const float* H, int n, long long n_global, const double* stats,
                        const double* bnsums, const float* dHz, const float* dHi, float* dH, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(params && H && stats && bnsums && dHz && dHi && dH, "mil_head_backward_b: null pointer argument");
  return mil_launch_head_bwd_b(head_params(params).bn_w, H, n, n_global, stats, bnsums, dHz, dHi, dH,
                               (cudaStream_t)stream);
  MIL_API_END
}

// ---- layer-level operators -----------------------------------------------------------------------------
size_t mil_pf8_bytes(int n, int c, int h, int w, int dtype) { return mil_pf8_bytes(mil_pf8(n, c, h, w), dtype); }

int mil_to_pf8(int dtype, const float* nchw, void* pf8, int n, int c, int h, int w, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  return mil_launch_to_pf8(dtype, nchw, pf8, n, c, h, w, (cudaStream_t)stream);
  MIL_API_END
}
int mil_from_pf8(int dtype, const void* pf8, float* nchw, int n, int c, int h, int w, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  return mil_launch_from_pf8(dtype, pf8, nchw, n, c, h, w, (cudaStream_t)stream);
  MIL_API_END
}

// true when the 3x3 / stride-2 convolution cin -> cout with output rows of `wo` pixels runs in its phase-split
// tensor-core forms (the extractor plan takes the same decision, mil_extractor.cu)
static bool s2_split_ok(int cin, int cout, int wo) {
  MilTcShape sh;
  if (cin < 9 || mil_tc_shape_s2(cin, cout, &sh) != 0) return false;
  return mil_conv_tc_fits(sh, wo + 1) && 2 * ((cin + 7) / 8) <= 10;
}

struct ConvWs {
  size_t off_tc, off_tc2, off_aux, off_partial, total;
};
// layer-level workspace: [fp32 packed weights][tc weights][tc weights 2][aux map][partial records]
static ConvWs conv_ws_layout(int n, int cin, int hi, int wi, int cout, int ho, int wo, int ks) {
  const MilPF8 gi = mil_pf8(n, cin, hi, wi), go = mil_pf8(n, cout, ho, wo);
  const size_t wp = ((size_t)ks * ks * gi.cb * 8 * go.cb * 8 * sizeof(float) + 255) / 256 * 256;
  size_t tc = 0;
  if (cin >= 9 && cout >= 9) {
    MilTcShape a, b;
    if (mil_tc_shape(cin, cout, ks, &a) == 0 && mil_tc_shape(cout, cin, ks, &b) == 0)
      tc = std::max(mil_tc_wpack_bytes(a), mil_tc_wpack_bytes(b));
    if (ks == 3 && mil_tc_shape_s2(cin, cout, &a) == 0) tc = std::max(tc, mil_tc_wpack_bytes(a));
    for (int r = 0; r < 2; ++r)
      if (ks == 3 && 2 * gi.cb <= 10 && mil_tc_shape_s2_dgrad(cout, cin, r, &a) == 0) tc = std::max(tc, mil_tc_wpack_bytes(a));
    tc = (tc + 255) / 256 * 256;
  }
  ConvWs w;
  w.off_tc = wp;
  w.off_tc2 = wp + tc;
  w.off_aux = wp + 2 * tc;
  // aux: the phase-split / even-position copy of the input of a stride-2 convolution (at most 4 * cb planes at the
  // output resolution), or the zero-stuffed half-resolution residual of its data gradient (cin channels, input size)
  size_t aux = 0;
  if (hi != ho || wi != wo)
    aux = std::max(mil_pf8_bytes(mil_split2_geom(n, cin, ho), MIL_F32), mil_pf8_bytes(gi, MIL_F32));
  aux = (aux + 255) / 256 * 256;
  w.off_partial = w.off_aux + aux;
  size_t pf = mil_wgrad_direct_partial_floats(gi, go, ks);
  pf = std::max(pf, mil_wgrad_tc_partial_floats(mil_pf8(n, cin, ho, wo), go, ks));
  w.total = w.off_partial + pf * sizeof(float) + 1024;
  return w;
}

int mil_upsample2_pf8(const void* in, int n, int c, int h, int w, void* out, int ho, int wo, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(in && out, "mil_upsample2_pf8: null pointer argument");
  return mil_launch_upsample2(in, mil_pf8(n, c, h, w), out, mil_pf8(n, c, ho, wo), (cudaStream_t)stream);
  MIL_API_END
}

int mil_minmax_normalize(const float* in, float* out, long long count, float* minmax, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(in && out && minmax && count > 0, "mil_minmax_normalize: null pointer argument");
  return mil_launch_minmax_normalize(in, out, count, minmax, (cudaStream_t)stream);
  MIL_API_END
}

int mil_adam_step(float* params_flat, const float* grads_flat, float* exp_avg, float* exp_avg_sq, long long count,
                  float step_size, float beta1, float beta2, float bc2_sqrt, float eps, float weight_decay,
                  void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(params_flat && grads_flat && exp_avg && exp_avg_sq && count > 0, "mil_adam_step: null pointer argument");
  return mil_launch_adam_step(params_flat, grads_flat, exp_avg, exp_avg_sq, count, step_size, beta1, beta2, bc2_sqrt,
                              eps, weight_decay, (cudaStream_t)stream);
  MIL_API_END
}

int mil_adam_step_dev(float* params_flat, const float* grads_flat, float* exp_avg, float* exp_avg_sq, long long count,
                      const float* hyper, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(params_flat && grads_flat && exp_avg && exp_avg_sq && hyper && count > 0,
              "mil_adam_step_dev: null pointer argument");
  return mil_launch_adam_step_dev(params_flat, grads_flat, exp_avg, exp_avg_sq, count, hyper, (cudaStream_t)stream);
  MIL_API_END
}

int mil_ingest_tiles_u8(const void* rois, int n_tiles, int roi, const int* crops, int pad, const unsigned char* flips,
                        int side, const int* bounds, const int* coef, int ksize, const int* bounds_host, void* out,
                        void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(rois && bounds && coef && bounds_host && out, "mil_ingest_tiles_u8: null pointer argument");
  MIL_REQUIRE(pad >= 0 && (crops != nullptr || flips == nullptr || true), "mil_ingest_tiles_u8: bad arguments");
  return mil_launch_ingest_u8((const uint8_t*)rois, n_tiles, roi, crops, pad, flips, side, bounds, coef, ksize,
                              bounds_host, (uint8_t*)out, (cudaStream_t)stream);
  MIL_API_END
}

size_t mil_conv_workspace_bytes(int n, int cin, int hi, int wi, int cout, int ho, int wo, int ks) {
  return conv_ws_layout(n, cin, hi, wi, cout, ho, wo, ks).total;
}

int mil_conv_pf8(int dtype, int impl, int transposed, const void* x, int n, int cx, int hx, int wx, const float* w,
                 int cout, int cin, int ks, int stride, const float* bias, const void* res, const void* act,
                 void* out, int ho, int wo, int epi, void* ws, size_t ws_bytes, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(x && w && out && ws, "mil_conv_pf8: null pointer argument");
  MIL_REQUIRE(cx == (transposed ? cout : cin), "mil_conv_pf8: x has %d channels, expected %d", cx,
              transposed ? cout : cin);
  const MilPF8 gx = mil_pf8(n, cx, hx, wx);
  const MilPF8 go = mil_pf8(n, transposed ? cin : cout, ho, wo);
  const ConvWs L = transposed ? conv_ws_layout(n, cin, ho, wo, cout, hx, wx, ks) : conv_ws_layout(n, cin, hx, wx, cout, ho, wo, ks);
  MIL_REQUIRE(ws_bytes >= L.total, "mil_conv_pf8: workspace too small (%zu < %zu)", ws_bytes, L.total);
  cudaStream_t s = (cudaStream_t)stream;
  char* wsb = reinterpret_cast<char*>(ws);
  MIL_TRY(mil_launch_pack_conv_w(w, (float*)ws, cout, cin, ks, transposed, s));
  const bool tc_ok = mil_tc_supported(dtype, ks, stride, gx.c, go.c) && !(stride == 2 && transposed);
  // the stride-2 3x3 convolution in the forms the extractor runs: forward on the phase-split input, data gradient
  // per input row parity (res: the HALF-resolution gradient of the projection branch, added at the even positions)
  const bool split = dtype == MIL_BF16 && ks == 3 && stride == 2 &&
                     (transposed ? s2_split_ok(cin, cout, wx) : s2_split_ok(cin, cout, wo));
  MIL_REQUIRE(impl != 2 || tc_ok || split, "mil_conv_pf8: impl=2 (tcgen05) does not support dtype=%d ks=%d stride=%d",
              dtype, ks, stride);
  const bool use_tc = impl != 1 && (tc_ok || split) && (impl == 2 || mil_tc_enabled());
  if (stride == 2 && transposed && res != nullptr && !(use_tc && split)) {
    // CUDA-core / zero-stuffed path: the half-resolution residual becomes a full-resolution map first
    const MilPF8 gh = mil_pf8(n, cin, hx, wx);
    void* up = wsb + L.off_aux;
    MIL_CHECK_CUDA(cudaMemsetAsync(up, 0, mil_pf8_bytes(go, dtype), s));
    if (dtype == MIL_BF16) MIL_TRY(mil_launch_upsample2(res, gh, up, go, s));
    else MIL_REQUIRE(false, "mil_conv_pf8: stride-2 data gradient with a residual needs dtype bf16");
    res = up;
  }
  if (!use_tc)
    return mil_launch_conv_direct(dtype, transposed, x, gx, (const float*)ws, bias, res, act, out, go, ks, stride,
                                  epi, s);
  if (split && !transposed) {
    const MilPF8 gs = mil_split2_geom(n, cin, ho);
    void* xs2 = wsb + L.off_aux;
    MIL_TRY(mil_zero_guards(dtype, xs2, gs, s));
    MIL_TRY(mil_launch_split2(x, gx, xs2, gs, s));
    MilTcPackJob job = {w, wsb + L.off_tc, cout, cin, 3, 0, 1};
    MIL_TRY(mil_launch_pack_tc_table(&job, 1, s));
    MilTcShape sh;
    MIL_TRY(mil_tc_shape_s2(cin, cout, &sh));
    return mil_launch_conv_tc(0, xs2, gs, wsb + L.off_tc, sh, bias, res, act, out, go, epi, 0, s);
  }
  if (split && transposed) {
    MIL_REQUIRE(epi == MIL_EPI_DGRAD && bias == nullptr, "mil_conv_pf8: the stride-2 data gradient has the DGRAD epilogue");
    const MilPF8 gh = mil_pf8(n, cin, hx, wx);  // geometry of the half-resolution residual
    MIL_CHECK_CUDA(cudaMemsetAsync(out, 0, mil_pf8_bytes(go, dtype), s));  // pads + guards of the full-resolution map
    for (int a = 0; a < 2; ++a) {
      char* wtc = wsb + (a ? L.off_tc2 : L.off_tc);
      MilTcPackJob job = {w, wtc, cout, cin, 3, 1, 2 + a};
      MIL_TRY(mil_launch_pack_tc_table(&job, 1, s));
      MilTcShape sh;
      MIL_TRY(mil_tc_shape_s2_dgrad(cout, cin, a, &sh));
      MIL_TRY(mil_launch_conv_tc(1, x, gx, wtc, sh, nullptr, (a == 0) ? res : nullptr, act, out, go, MIL_EPI_DGRAD, 0, s,
                                 (a == 0 && res != nullptr) ? &gh : nullptr, a, nullptr));
    }
    return 0;
  }
  MilTcShape sh;
  MIL_TRY(mil_tc_shape(gx.c, go.c, ks, &sh));
  void* wtc = wsb + L.off_tc;
  MIL_TRY(mil_launch_pack_tc((const float*)ws, wtc, sh, s));
  return mil_launch_conv_tc(transposed, x, gx, wtc, sh, bias, res, act, out, go, epi, stride == 2, s);
  MIL_API_END
}

int mil_conv_wgrad_pf8(int dtype, int impl, const void* x, int n, int cin, int hi, int wi, const void* dz, int cout,
                       int ho, int wo, int ks, int stride, float* dw, float* db, void* ws, size_t ws_bytes,
                       void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(x && dz && dw && ws, "mil_conv_wgrad_pf8: null pointer argument");
  const MilPF8 gi = mil_pf8(n, cin, hi, wi), go = mil_pf8(n, cout, ho, wo);
  const ConvWs L = conv_ws_layout(n, cin, hi, wi, cout, ho, wo, ks);
  MIL_REQUIRE(ws_bytes >= L.total, "mil_conv_wgrad_pf8: workspace too small (%zu < %zu)", ws_bytes, L.total);
  char* wsb = reinterpret_cast<char*>(ws);
  float* partial = reinterpret_cast<float*>(wsb + L.off_partial);
  cudaStream_t s = (cudaStream_t)stream;
  const bool tc_ok = mil_wgrad_tc_supported(dtype, ks, stride, cin, cout);
  // stride 2 in the forms the extractor runs: 3x3 on the phase-split input (nine single-tap MMAs), 1x1 on the
  // even-position copy -- everything at the OUTPUT resolution
  const bool s2_3 = dtype == MIL_BF16 && ks == 3 && stride == 2 && s2_split_ok(cin, cout, wo);
  const bool s2_1 = dtype == MIL_BF16 && ks == 1 && stride == 2 && mil_wgrad_tc_supported(dtype, 1, 1, cin, cout);
  MIL_REQUIRE(impl != 2 || tc_ok || s2_3 || s2_1,
              "mil_conv_wgrad_pf8: impl=2 (tcgen05) does not support dtype=%d ks=%d stride=%d", dtype, ks, stride);
  if (impl == 1 || !(tc_ok || s2_3 || s2_1) || (impl == 0 && !mil_tc_enabled()))
    return mil_launch_wgrad_direct(dtype, x, gi, dz, go, partial, dw, db, ks, stride, s);
  if (s2_3) {
    const MilPF8 gs = mil_split2_geom(n, cin, ho);
    void* xs2 = wsb + L.off_aux;
    MIL_TRY(mil_zero_guards(dtype, xs2, gs, s));
    MIL_TRY(mil_launch_split2(x, gi, xs2, gs, s));
    return mil_launch_wgrad_tc_s2(xs2, gs, dz, go, partial, dw, db, cin, s);
  }
  if (s2_1) {
    const MilPF8 gsub = mil_pf8(n, cin, ho, wo);
    void* xsub = wsb + L.off_aux;
    MIL_TRY(mil_zero_guards(dtype, xsub, gsub, s));
    MIL_TRY(mil_launch_subsample2(x, gi, xsub, gsub, s));
    return mil_launch_wgrad_tc(xsub, gsub, dz, go, partial, dw, db, 1, s);
  }
  return mil_launch_wgrad_tc(x, gi, dz, go, partial, dw, db, ks, s);
  MIL_API_END
}

// ---- stem at layer level ----------------------------------------------------------------------------------
struct StemWs {
  size_t off_argmax, off_xs, off_cv, off_wp, off_wtc, off_partial, total;
};
static StemWs stem_ws_layout(int n, int side, int dtype) {
  const MilGeom geo = mil_geom(side);
  const MilPF8 gp = mil_pf8(n, 20, geo.h[0], geo.h[0]);
  StemWs L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
  L.off_argmax = take(std::max((size_t)n * geo.h[0] * geo.h[0] * 20, mil_stem_tc_argmax_bytes(gp)));
  L.off_xs = take(mil_pf8_bytes(mil_stem_tc_geom_in(n, side), MIL_BF16));
  L.off_cv = take(mil_pf8_bytes(mil_stem_tc_geom_conv(n, side), MIL_BF16));
  L.off_wp = take(mil_stem_tc_wpack_floats() * sizeof(float));
  L.off_wtc = take(mil_stem_tc_wtc_bytes());
  L.off_partial = take(std::max(mil_stem_bwd_partial_floats(), mil_stem_tc_partial_floats(n, side)) * sizeof(float));
  L.total = off;
  (void)dtype;
  return L;
}
static bool stem_use_tc(int dtype, int impl) { return dtype == MIL_BF16 && impl != 1 && (impl == 2 || mil_tc_enabled()); }

size_t mil_stem_workspace_bytes(int n, int side, int dtype) {
  if (n < 1 || side < 8) return 0;
  return stem_ws_layout(n, side, dtype).total;
}

int mil_stem_forward(int dtype, int impl, const float* bag, int n, int side, const float* w, const float* b, void* pooled,
                     void* ws, size_t ws_bytes, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(bag && w && b && pooled && ws, "mil_stem_forward: null pointer argument");
  MIL_REQUIRE(n >= 1 && side >= 8, "mil_stem_forward: bad shape n=%d side=%d", n, side);
  MIL_REQUIRE(impl != 2 || dtype == MIL_BF16, "mil_stem_forward: impl=2 (tcgen05) needs dtype bf16");
  const StemWs L = stem_ws_layout(n, side, dtype);
  MIL_REQUIRE(ws_bytes >= L.total, "mil_stem_forward: workspace too small (%zu < %zu)", ws_bytes, L.total);
  const MilGeom geo = mil_geom(side);
  const MilPF8 gp = mil_pf8(n, 20, geo.h[0], geo.h[0]);
  char* wsb = reinterpret_cast<char*>(ws);
  cudaStream_t s = (cudaStream_t)stream;
  if (!stem_use_tc(dtype, impl))
    return mil_launch_stem_fwd(dtype, bag, nullptr, n, side, w, b, pooled, gp, (uint8_t*)(wsb + L.off_argmax), s);
  MIL_TRY(mil_zero_guards(MIL_BF16, wsb + L.off_xs, mil_stem_tc_geom_in(n, side), s));
  MIL_TRY(mil_zero_guards(MIL_BF16, wsb + L.off_cv, mil_stem_tc_geom_conv(n, side), s));
  return mil_launch_stem_tc_fwd(bag, 0, nullptr, n, side, w, b, wsb + L.off_xs, wsb + L.off_cv, (float*)(wsb + L.off_wp),
                                wsb + L.off_wtc, pooled, gp, (uint8_t*)(wsb + L.off_argmax), s, nullptr);
  MIL_API_END
}

int mil_stem_backward(int dtype, int impl, const float* bag, int n, int side, const void* g, float* dw, float* db, void* ws,
                      size_t ws_bytes, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(bag && g && dw && db && ws, "mil_stem_backward: null pointer argument");
  MIL_REQUIRE(n >= 1 && side >= 8, "mil_stem_backward: bad shape n=%d side=%d", n, side);
  const StemWs L = stem_ws_layout(n, side, dtype);
  MIL_REQUIRE(ws_bytes >= L.total, "mil_stem_backward: workspace too small (%zu < %zu)", ws_bytes, L.total);
  const MilGeom geo = mil_geom(side);
  const MilPF8 gp = mil_pf8(n, 20, geo.h[0], geo.h[0]);
  char* wsb = reinterpret_cast<char*>(ws);
  cudaStream_t s = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(wsb + L.off_partial);
  if (!stem_use_tc(dtype, impl))
    return mil_launch_stem_bwd(dtype, bag, nullptr, n, side, g, gp, (const uint8_t*)(wsb + L.off_argmax), partial, dw, db, s);
  return mil_launch_stem_tc_bwd(wsb + L.off_xs, n, side, g, gp, (const uint8_t*)(wsb + L.off_argmax), wsb + L.off_cv, partial,
                                dw, db, s);
  MIL_API_END
}

// ---- wide-channel kernels at layer level (alt_resnet.py parameterisation; tests/test_gpu_wide.py) ----------------
size_t mil_wide_conv_workspace_bytes(int mode, int transposed, int wcout, int wcin, int ks) {
  MilWideShape sh;
  if (mil_wide_shape(mode, transposed, wcout, wcin, ks, &sh) != 0) return 0;
  return mil_wide_wpack_bytes(sh) + 256;
}

int mil_wide_conv_pf8(int mode, int transposed, const void* x, int n, int cx, int h, int w, const float* wt, int wcout,
                      int wcin, int ks, const float* bias, const void* res, const void* act, void* out, int epi,
                      float slope, int tm, void* ws, size_t ws_bytes, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(x && wt && out && ws, "mil_wide_conv_pf8: null pointer argument");
  MilWideShape sh;
  MIL_TRY(mil_wide_shape(mode, transposed, wcout, wcin, ks, &sh));
  MIL_REQUIRE(ws_bytes >= mil_wide_wpack_bytes(sh), "mil_wide_conv_pf8: workspace too small (%zu < %zu)", ws_bytes,
              mil_wide_wpack_bytes(sh));
  MIL_REQUIRE(cx == sh.kin * sh.ngroups, "mil_wide_conv_pf8: x has %d channels, expected %d", cx, sh.kin * sh.ngroups);
  const MilPF8 gx = mil_pf8(n, cx, h, w), go = mil_pf8(n, sh.nout, h, w);
  cudaStream_t s = (cudaStream_t)stream;
  MIL_TRY(mil_launch_wide_pack(wt, ws, sh, s));
  return mil_launch_wide_conv(x, gx, ws, sh, bias, res, act, out, go, epi, slope, tm, s);
  MIL_API_END
}

size_t mil_wide_wgrad_workspace_bytes(int n, int cin, int cout, int h, int w, int ks) {
  return mil_wide_wgrad_partial_floats(mil_pf8(n, cin, h, w), mil_pf8(n, cout, h, w), ks) * sizeof(float) + 256;
}

int mil_wide_wgrad_pf8(const void* x, int n, int cin, int h, int w, const void* dz, int cout, int ks, float* dw, void* ws,
                       size_t ws_bytes, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(x && dz && dw && ws, "mil_wide_wgrad_pf8: null pointer argument");
  const MilPF8 gx = mil_pf8(n, cin, h, w), gz = mil_pf8(n, cout, h, w);
  const size_t need = mil_wide_wgrad_partial_floats(gx, gz, ks) * sizeof(float);
  MIL_REQUIRE(need > 0, "mil_wide_wgrad_pf8: unsupported shape cin=%d cout=%d ks=%d", cin, cout, ks);
  MIL_REQUIRE(ws_bytes >= need, "mil_wide_wgrad_pf8: workspace too small (%zu < %zu)", ws_bytes, need);
  return mil_launch_wide_wgrad(x, gx, dz, gz, (float*)ws, dw, nullptr, ks, (cudaStream_t)stream);
  MIL_API_END
}

size_t mil_wide_wgrad_s2_workspace_bytes(int n, int cin, int cout, int ho, int wo) {
  return mil_wide_wgrad_partial_floats(mil_split2_geom(n, cin, ho), mil_pf8(n, cout, ho, wo), 3, 1) * sizeof(float) + 256;
}

int mil_wide_wgrad_s2_pf8(const void* xs2, int n, int cin, int ho, int wo, const void* dz, int cout, float* dw, void* ws,
                          size_t ws_bytes, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(xs2 && dz && dw && ws, "mil_wide_wgrad_s2_pf8: null pointer argument");
  MIL_REQUIRE(ho == wo, "mil_wide_wgrad_s2_pf8: square maps only");
  const MilPF8 gs = mil_split2_geom(n, cin, ho), gz = mil_pf8(n, cout, ho, wo);
  const size_t need = mil_wide_wgrad_partial_floats(gs, gz, 3, 1) * sizeof(float);
  MIL_REQUIRE(need > 0, "mil_wide_wgrad_s2_pf8: unsupported shape cin=%d cout=%d", cin, cout);
  MIL_REQUIRE(ws_bytes >= need, "mil_wide_wgrad_s2_pf8: workspace too small (%zu < %zu)", ws_bytes, need);
  return mil_launch_wide_wgrad(xs2, gs, dz, gz, (float*)ws, dw, nullptr, 3, (cudaStream_t)stream, 1);
  MIL_API_END
}

int mil_merge2_pf8(const void* in, int n, int c, int h, int w, void* out, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(in && out, "mil_merge2_pf8: null pointer argument");
  return mil_launch_merge2(in, mil_split2_geom(n, c, (h - 1) / 2 + 1), out, mil_pf8(n, c, h, w), (cudaStream_t)stream);
  MIL_API_END
}

int mil_split2_pf8(const void* in, int n, int c, int h, int w, void* out, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(in && out, "mil_split2_pf8: null pointer argument");
  const MilPF8 gin = mil_pf8(n, c, h, w);
  return mil_launch_split2(in, gin, out, mil_split2_geom(n, c, (h - 1) / 2 + 1), (cudaStream_t)stream);
  MIL_API_END
}

// ---- the wide extractor as a whole ------------------------------------------------------------------------------
int mil_wide_param_count(const MilWideDesc* desc) {
  if (desc == nullptr || mil_wide_check_desc(*desc) != 0) return 0;
  return (int)mil_wide_param_table(*desc).size();
}
int mil_wide_param_info(const MilWideDesc* desc, int i, char* name_out, int name_cap, int* ndim, long long shape4[4],
                        long long* offset) {
  MIL_API_BEGIN
  MIL_REQUIRE(desc && name_out && ndim && shape4 && offset && name_cap > 0, "mil_wide_param_info: null pointer argument");
  MIL_TRY(mil_wide_check_desc(*desc));
  const auto t = mil_wide_param_table(*desc);
  MIL_REQUIRE(i >= 0 && i < (int)t.size(), "mil_wide_param_info: index %d out of range", i);
  snprintf(name_out, (size_t)name_cap, "%s", t[i].name.c_str());
  *ndim = t[i].ndim;
  for (int k = 0; k < 4; ++k) shape4[k] = t[i].shape[k];
  *offset = t[i].offset;
  return 0;
  MIL_API_END
}
long long mil_wide_param_total(const MilWideDesc* desc) {
  if (desc == nullptr || mil_wide_check_desc(*desc) != 0) return 0;
  const auto t = mil_wide_param_table(*desc);
  return t.back().offset + t.back().numel;
}
size_t mil_wide_workspace_bytes(const MilWideDesc* desc, int n_tiles, int side) {
  try {
    MilWidePlan pl;
    if (desc == nullptr || mil_wide_make_plan(*desc, n_tiles, side, &pl) != 0) return 0;
    return pl.total_bytes;
  } catch (...) {
    return 0;
  }
}
int mil_wide_forward(const MilWideDesc* desc, const void* const* params, const void* bag, int bag_is_u8, const int32_t* idx,
                     int n_tiles, int side, void* ws, size_t ws_bytes, float* H, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(desc && params && bag && ws && H, "mil_wide_forward: null pointer argument");
  MilWidePlan pl;
  MIL_TRY(mil_wide_make_plan(*desc, n_tiles, side, &pl));
  MIL_REQUIRE(ws_bytes >= pl.total_bytes, "mil_wide_forward: workspace too small (%zu < %zu)", ws_bytes, pl.total_bytes);
  return mil_wide_forward_impl(params, bag, bag_is_u8, idx, pl, ws, H, (cudaStream_t)stream);
  MIL_API_END
}
int mil_wide_backward(const MilWideDesc* desc, const void* const* params, int n_tiles, int side, void* ws, size_t ws_bytes,
                      const float* dH, float* grads_flat, void* stream) {
  MIL_API_BEGIN
  MIL_TRY(require_device());
  MIL_REQUIRE(desc && params && ws && dH && grads_flat, "mil_wide_backward: null pointer argument");
  MilWidePlan pl;
  MIL_TRY(mil_wide_make_plan(*desc, n_tiles, side, &pl));
  MIL_REQUIRE(ws_bytes >= pl.total_bytes, "mil_wide_backward: workspace too small (%zu < %zu)", ws_bytes, pl.total_bytes);
  return mil_wide_backward_impl(params, pl, ws, dH, grads_flat, (cudaStream_t)stream);
  MIL_API_END
}

}  // extern "C"
