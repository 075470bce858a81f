// Host-side model description of the ResNet-26 extractor + head parameters (reference gbm/model.py:14-61,
// 118-159): the 65 tensors of the reference state dict in its own order, the workspace plan, and the
// whole-network forward / backward drivers (mil_extractor.cu).
#pragma once
#include <string>
#include <vector>

#include "mil_common.cuh"
#include "mil_conv_tc.cuh"

struct MilParamInfo {
  std::string name;
  int ndim;
  long long shape[4];
  long long numel;
  long long offset;  // float offset inside the flat gradient buffer (state-dict order, densely packed)
};
const std::vector<MilParamInfo>& mil_param_table();
int mil_param_index(const char* name);  // -1 if unknown

// one convolution of the extractor
struct MilConvDesc {
  int layer, block;   // 0..3, 0..2
  int which;          // 0 = conv1, 1 = conv2, 2 = downsample
  int cin, cout, ks, stride;
  int p_w, p_b;       // parameter indices (p_b = -1: no bias)
  size_t wp_off, wpt_off;  // float offsets of the packed normal / transposed weights inside the pack area
  bool tc;                 // forward and data gradient run on the tcgen05 kernel (bf16 mode, 3x3 stride 1)
  size_t wtc_off, wtct_off;  // byte offsets of the bf16 UMMA-layout weights (normal / transposed) in the tc area
  size_t wtct_s2_off[2];     // stride-2 3x3 only: the data-gradient weights of the two input row parities
  bool fold;                 // conv2 of a stride-2 block: the block's 1x1 projection runs as extra K-groups of this convolution
  size_t wtc_fold_off;       //   (mil_tc_shape_fold); its operand blocks
};

// the phase-split copy of a c-channel map whose stride-2 output is ho x ho: 4 * cb chunk planes at that resolution
static inline MilPF8 mil_split2_geom(int n, int c, int ho) { return mil_pf8(n, 4 * ((c + 7) / 8) * 8, ho, ho); }
// view of its phase (0,0) = the first cb planes (same plane stride): the map at its even positions
static inline MilPF8 mil_split2_phase0(MilPF8 g, int c) {
  g.c = c;
  g.cb = (c + 7) / 8;
  return g;
}

struct MilPlan;
MilPF8 mil_xs2_geom(const MilPlan& pl, int l);  // geometry of the saved stride-2 block input of layer l (split or even-only)

struct MilPlan {
  int n, side, dtype;
  bool infer;                   // forward-only plan: three rotating map buffers, no sign masks, no arg-max records
  MilGeom geo;
  MilPF8 g[4];                  // activation geometry of layer1..4
  std::vector<MilConvDesc> convs;  // 27 entries, forward order
  size_t off_pooled, off_argmax, off_h[12], off_y[12], off_avg, off_grad[3], off_up[2], off_xs2[4], off_mh[12], off_my[12], off_mpool, off_wpack, off_wtc, off_partial;
  bool stem_tc;                 // stem on the tensor cores (bf16 mode)
  size_t off_xs, off_cv, off_stem_wp, off_stem_wtc;
  bool pool_mask;    // the fused stem kernel also writes the sign mask of the pooled map (off_mpool)
  bool masks;        // sign masks of the saved activations exist (bf16 tensor-core path): off_mh / off_my
  bool s2_split[4];  // stride-2 block of layer l runs on the phase-split input (else: full-resolution evaluation)
  size_t wpack_floats, wtc_bytes, partial_floats, partial_arena_floats, grad_bytes, up_bytes;
  size_t total_bytes;
};
int mil_make_plan(int n, int side, int dtype, MilPlan* plan, bool infer = false);

int mil_extractor_forward_impl(const void* const* params, const void* bag, int bag_u8, const int* idx,
                               const MilPlan& pl, void* ws, float* H, cudaStream_t s);
int mil_extractor_backward_impl(const void* const* params, const float* bag, const int* idx, const MilPlan& pl,
                                void* ws, const float* dH, float* grads, cudaStream_t s,
                                const cudaEvent_t* layer_events = nullptr);

void mil_debug_request_dump(int layer, int block, int which, float* dst);

// conv dispatch: tcgen05 implicit GEMM where supported (bf16), CUDA-core direct kernel otherwise
// wtc: bf16 UMMA-layout weights (mil_launch_pack_tc) or NULL -> CUDA-core kernel
int mil_conv_dispatch(int dtype, int transposed, const void* x, const MilPF8& gi, const float* wp, const void* wtc,
                      const float* bias,
                      const void* res, const void* act, void* out, const MilPF8& go, int ks, int stride, int epi,
                      cudaStream_t s, const void* mask_in = nullptr, void* mask_out = nullptr);
int mil_wgrad_dispatch(int dtype, const void* x, const MilPF8& gi, const void* dz, const MilPF8& go, float* partial,
                       float* dw, float* db, int ks, int stride, cudaStream_t s);
bool mil_tc_enabled();  // false when MIL_B200_DISABLE_TC=1 (debugging aid: CUDA-core kernels only)
int mil_zero_guards(int dtype, void* buf, const MilPF8& g, cudaStream_t s);
