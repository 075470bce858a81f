// Host-side model description of the WIDE extractor parameterisation (alt_resnet.py:70-145) -- parameter table,
// workspace plan and the whole-network forward / backward drivers (mil_wide_net.cu).
#pragma once
#include <string>
#include <vector>

#include "../../include/mil_b200.h"
#include "mil_common.cuh"
#include "mil_extractor.cuh"

struct MilWideConv {
  int layer, block, which;  // which: 0 = conv1, 1 = conv2, 2 = downsample.0
  int cin, cout, ks, stride;
  int p_w;                  // parameter index
  size_t wf_off, wt_off;    // byte offsets of the packed forward / data-gradient operand blocks in the pack area
};

struct MilWidePlan {
  MilWideDesc d;
  int n, side;
  int hc, h[4];             // conv1 output side; sides of layer1..4
  MilPF8 g[4], gxs, gcv;    // layer maps; space-to-depth input (48 ch); four-phase conv map (4 * stem ch)
  std::vector<MilParamInfo> params;
  std::vector<MilWideConv> convs;
  std::vector<std::vector<int>> first_conv;  // [layer][block] -> index of its conv1 in `convs`
  std::vector<std::vector<size_t>> off_h, off_y;
  size_t off_xs, off_pooled, off_argmax, off_xs2[4], off_avg, off_cv, off_up, off_tsub, off_tfull, off_grad[3], off_wpack,
      off_partial, stem_w_off;
  size_t s2_wt_off[4][4];   // [layer][input parity phase]: operand blocks of the stride-2 convolution's phase-wise data gradient
  size_t wpack_bytes, partial_floats, total_bytes;
};

int mil_wide_check_desc(const MilWideDesc& d);
std::vector<MilParamInfo> mil_wide_param_table(const MilWideDesc& d);
int mil_wide_make_plan(const MilWideDesc& d, int n, int side, MilWidePlan* plan);
int mil_wide_forward_impl(const void* const* params, const void* bag, int bag_u8, const int* idx, const MilWidePlan& pl,
                          void* ws, float* H, cudaStream_t s);
int mil_wide_backward_impl(const void* const* params, const MilWidePlan& pl, void* ws, const float* dH, float* grads,
                           cudaStream_t s);
