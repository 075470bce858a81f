// Wide-channel kernel family (64 .. 512 channels): the alt_resnet.py parameterisation of the extractor
// (reference alt_resnet.py:24-32,35-67,70-145: torchvision's ResNet with the BatchNorm layers stripped -- conv3x3 /
// conv1x1 without bias, ReLU, widths 64/128/256/512, fc with bias) on the PF8 layout.
//
// conv_tc_kernel / wgrad_tc_kernel (the 20..80-channel network of gbm/model.py) keep ALL weights of a convolution in
// shared memory and all its output channels in one MMA; beyond 80 channels neither fits.  Here the K loop runs over
// pairs of 8-channel chunks with the matching WEIGHT SLAB streamed next to the input planes, the output channels are
// tiled by 128 (64 for the 64-channel layer), and several 128-pixel M-tiles share one slab so that the L2 -> shared
// fill stays below what a CTA can pull (B200_PROFILING: ~25-40 B/clk/SM).
#pragma once
#include "mil_common.cuh"

#define WIDE_MAX_GROUPS 4
#define WIDE_MAX_TAPS 9

// K-loop description of one convolution as the wide kernels run it.
//   group = a set of input chunk planes read with the same tap list (one group, or the four parity phases of the
//   phase-split input of a stride-2 3x3 convolution); K-step = (group, pair of chunks): one K = 16 MMA per tap
struct MilWideShape {
  int mode;        // 0: ks = 1 / 3, stride 1 (forward or data gradient);  1: 3x3 / stride 2 forward on the phase-split
                   // input (mil_launch_split2);  2: stem 7x7 / stride 2 in space-to-depth-by-4 form (3x3 taps, 48 -> 4*C);
                   // 3 + 2a + b: data gradient of the 3x3 / stride-2 convolution for the input pixels of parity phase (a, b)
                   // (x = the output gradient, out = that phase of the input gradient, both at the OUTPUT resolution);
                   // 7: the same for ALL four phases at once, as 4 * wcin output channels (phase, ci) = the phase-split map
  int transposed;  // mode 0: data gradient (kernel input = the conv's output channels, taps mirrored)
  int wcout, wcin, ks;  // the PyTorch weight [wcout][wcin][ks][ks]
  int kin, nout;        // kernel-side input channels PER GROUP / output channels
  int ngroups, npairs;  // K-steps = ngroups * npairs
  int gntaps[WIDE_MAX_GROUPS], ga[WIDE_MAX_GROUPS], gb[WIDE_MAX_GROUPS];  // taps and parity phase (a, b) of each group
  signed char tdy[WIDE_MAX_GROUPS][WIDE_MAX_TAPS], tdx[WIDE_MAX_GROUPS][WIDE_MAX_TAPS];
  int nt, n_ntiles;     // output-channel tile (64 / 128) and their count
  int blocks_per_ntile; // weight blocks (tap x chunk pair) per N-tile = sum_g npairs * gntaps[g]
};
int mil_wide_shape(int mode, int transposed, int wcout, int wcin, int ks, MilWideShape* out);
size_t mil_wide_wpack_bytes(const MilWideShape& sh);
// w: PyTorch layout fp32 [wcout][wcin][ks][ks] -> bf16 operand blocks [N-tile][group][pair][tap][K half][nt rows][8]
int mil_launch_wide_pack(const float* w, void* wpk, const MilWideShape& sh, cudaStream_t s);
// out = epilogue(conv(x)):  epi / res / act as in mil_launch_conv_tc, slope = the activation's negative slope (0 = ReLU);
// x and out share their pixel geometry (stride-2 forward: x = the phase-split input at the OUTPUT resolution);
// tm = 128-pixel M-tiles per weight slab (0 = choose)
int mil_launch_wide_conv(const void* x, const MilPF8& gx, const void* wpk, const MilWideShape& sh, const float* bias,
                         const void* res, const void* act, void* out, const MilPF8& go, int epi, float slope, int tm,
                         cudaStream_t s, int res_chunks = 0);  // res_chunks > 0: the residual covers only the first chunks

// weight gradient dW[co][ci][tap] += sum_q x[q + shift_tap][ci] * dz[q][co] (stride 1; a stride-2 convolution hands in
// the zero-stuffed dz), db[co] += sum_q dz[q][co] (db may be NULL).  ks = 1 / 3, or 7 with x = the stem's
// space-to-depth input (48 channels) and dz = the four-phase gradient map (4 * C channels).  s2 = 1: the 3x3 / stride-2
// convolution on its phase-split input (x = mil_launch_split2's 4 * cb planes, dz at the output resolution).
size_t mil_wide_wgrad_partial_floats(const MilPF8& gx, const MilPF8& gz, int ks, int s2 = 0);
int mil_launch_wide_wgrad(const void* x, const MilPF8& gx, const void* dz, const MilPF8& gz, float* partial, float* dw,
                          float* db, int ks, cudaStream_t s, int s2 = 0);

// out (full resolution) = the four parity phases in `in` (4 * cb planes at half resolution, mil_launch_split2's layout)
// interleaved back; pad pixels of `out` are written as zeros (bf16)
int mil_launch_merge2(const void* in, const MilPF8& gin, void* out, const MilPF8& gout, cudaStream_t s);
