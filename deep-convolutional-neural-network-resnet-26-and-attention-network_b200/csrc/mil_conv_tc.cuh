// Host-side declarations of the tcgen05 implicit-GEMM convolution (mil_conv_tc.cu).
#pragma once
#include "mil_common.cuh"

#define MIL_TC_MAX_MMA 46  // ceil(9 taps * 10 chunks / 2) + 1

// K-loop description shared by the weight pre-pack and the kernel: MMA j multiplies K-groups order[2j], order[2j+1]
struct MilTcShape {
  int ks;           // 3 (nine taps) or 1 (pointwise)
  int cbin, cbout;  // 8-channel chunks of the kernel's input / output
  int npad;         // UMMA N (output channels padded to a multiple of 16)
  int nmma;         // number of K=16 MMAs per tile
  unsigned char g_tap[2 * MIL_TC_MAX_MMA];    // tap (0..8) of each K-group, 0xFF = zero padding group
  unsigned char g_chunk[2 * MIL_TC_MAX_MMA];  // input chunk of each K-group
};

bool mil_tc_supported(int dtype, int ks, int stride, int cin, int cout);
int mil_tc_shape(int cin, int cout, int ks, MilTcShape* out);  // cin/cout = the KERNEL's input/output channels
size_t mil_tc_wpack_bytes(const MilTcShape& sh);
// wp: fp32 packed weights [tap][kin_pad][nout_pad] (mil_launch_pack_conv_w, normal or transposed)
int mil_launch_pack_tc(const float* wp, void* wtc, const MilTcShape& sh, cudaStream_t s);
int mil_launch_conv_tc(int transposed, const void* x, const MilPF8& gx, const void* wtc, const MilTcShape& sh,
                       const float* bias, const void* res, const void* act, void* out, const MilPF8& go, int epi,
                       int sub, cudaStream_t s);
// out (geometry 2x) = zero-stuffed copy of in: out(n, 2y, 2x) = in(n, y, x), zero elsewhere (bf16)
int mil_launch_upsample2(const void* in, const MilPF8& gin, void* out, const MilPF8& gout, cudaStream_t s);

// tcgen05 weight gradient (mil_wgrad_tc.cu): 3x3 / 1x1, stride 1, bf16
bool mil_wgrad_tc_supported(int dtype, int ks, int stride, int cin, int cout);
size_t mil_wgrad_tc_partial_floats(const MilPF8& gx, const MilPF8& gz, int ks);
int mil_launch_wgrad_tc(const void* x, const MilPF8& gx, const void* dz, const MilPF8& gz, float* partial, float* dw,
                        float* db, int ks, cudaStream_t s);
