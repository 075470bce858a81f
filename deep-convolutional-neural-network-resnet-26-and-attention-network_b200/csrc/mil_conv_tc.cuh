// Host-side declarations of the tcgen05 implicit-GEMM convolution (mil_conv_tc.cu).
#pragma once
#include "mil_common.cuh"

#define MIL_TC_MAX_MMA 50  // ceil((9 taps * 10 chunks + 8 projection chunks) / 2) + 1

#define MIL_TC_MAX_TAPS 16
// Window + K-loop description shared by the weight pre-pack and the kernels.  A tap is an offset (dy, dx) of the
// flat pixel index (shift = dy * wp + dx; the data gradient negates it); MMA j multiplies the K-groups
// (tap, 8-channel chunk) number 2j and 2j+1 of the order below.
struct MilTcShape {
  int ks;           // 3: nine taps -1..1;  1: pointwise;  7: the stem in space-to-depth form, 4x4 taps -2..1
  int ntaps;
  signed char t_dy[MIL_TC_MAX_TAPS], t_dx[MIL_TC_MAX_TAPS];
  int cbin, cbout;  // 8-channel chunks of the kernel's input / output
  int npad;         // UMMA N (output channels padded to a multiple of 16)
  int nmma;         // number of K=16 MMAs per tile
  unsigned char g_tap[2 * MIL_TC_MAX_MMA];    // tap index of each K-group, 0xFF = zero padding group
  unsigned char g_chunk[2 * MIL_TC_MAX_MMA];  // input chunk of each K-group
};
// largest |shift| of the window on a map whose padded row length is wp
static inline int mil_tc_halo(const MilTcShape& sh, int wp) {
  int m = 0;
  for (int t = 0; t < sh.ntaps; ++t) {
    const int s = sh.t_dy[t] * wp + sh.t_dx[t];
    m = s > m ? s : (-s > m ? -s : m);
  }
  return m;
}

// per-MMA descriptor templates, precomputed on the host and passed as kernel parameters so that the issuing
// warp reads them through the uniform datapath (constant bank): A relative to the start of an A stage, B
// relative to the start of the weight block (M-tile = 128 pixels, planes of 128 + 2 * halo pixels)
struct TcIssue {
  uint64_t a_desc[MIL_TC_MAX_MMA];
  uint64_t b_desc[MIL_TC_MAX_MMA];
};
int mil_tc_build_issue(const MilTcShape& sh, int wp, int halo, int transposed, TcIssue* out);

bool mil_tc_supported(int dtype, int ks, int stride, int cin, int cout);
int mil_tc_shape(int cin, int cout, int ks, MilTcShape* out);  // cin/cout = the KERNEL's input/output channels
// 3x3 / stride-2 convolution on the PHASE-SPLIT input (mil_launch_split2: 4 * cb chunk planes at the output
// resolution, plane = phase * cb + chunk, phase = (row parity) * 2 + (column parity)): a 2x2 window of taps
// (-1..0) in which tap (dy, dx) of phase (a, b) carries the weight w[2 dy + a + 1][2 dx + b + 1] -- the same nine
// K-groups per chunk as a stride-1 3x3 convolution, evaluated at a quarter of the pixels.
int mil_tc_shape_s2(int cin, int cout, MilTcShape* out);
// conv2 of a stride-2 block with the block's 1x1 projection folded in as extra K-groups (nnBlocks.py:183-187):
// the kernel's input is the map h (cmain channels) FOLLOWED IN MEMORY by the projection's input (cproj channels, the even
// positions of the block input = the first planes of the phase-split copy) at the same geometry; K-groups = nine taps x
// the chunks of h + the centre tap x the projection chunks.  out = conv3x3(h) + conv1x1(x_even) in one accumulator.
int mil_tc_shape_fold(int cmain, int cproj, int cout, MilTcShape* out);
// data gradient of that convolution for the input rows of parity a, BOTH column parities at once: the kernel's output
// chunks are (column parity b, input chunk) = 2 * cb chunks, so that a thread stores the two neighbouring pixels
// (2Y + a, 2X) and (2Y + a, 2X + 1) together; taps (dY, dX) in {0, a} x {0, 1} of the output gradient, with
// w[ky][kx] where ky = a ? (dY ? 0 : 2) : 1 and kx = b ? (dX ? 0 : 2) : (dX ? none : 1)
int mil_tc_shape_s2_dgrad(int cout_conv, int cin_conv, int a, MilTcShape* out);
size_t mil_tc_wpack_bytes(const MilTcShape& sh);
// wp: fp32 packed weights [tap][kin_pad][nout_pad] (mil_launch_pack_conv_w, normal or transposed)
int mil_launch_pack_tc(const float* wp, void* wtc, const MilTcShape& sh, cudaStream_t s);
// true when the kernel's shared-memory ring fits for this window on a map of padded row length wp
bool mil_conv_tc_fits(const MilTcShape& sh, int wp);
int mil_launch_conv_tc(int transposed, const void* x, const MilPF8& gx, const void* wtc, const MilTcShape& sh,
                       const float* bias, const void* res, const void* act, void* out, const MilPF8& go, int epi,
                       int sub, cudaStream_t s, const MilPF8* gres_half = nullptr, int up_row = -1,
                       const void* mask_in = nullptr, void* mask_out = nullptr);
// Sign mask of an activation map (1 bit per element: value > 0), written by the forward epilogues (mask_out) and read
// by the data-gradient epilogues (mask_in) in place of the 16-bit activations: word w of a pixel holds chunks 4w..4w+3
// (byte = chunk, bit = channel), words are planes of PS pixels like the chunk planes of the map itself.
static inline size_t mil_sign_mask_bytes(const MilPF8& g) { return (size_t)((g.cb + 3) / 4) * g.PS * 4; }
// every convolution of a pass packed in one launch, straight from the PyTorch weight layout [cout][cin][ks][ks]
struct MilTcPackJob {
  const float* w;
  void* wtc;
  int cout, cin, ks, transposed;
  int s2;  // 0: plain;  1: forward of the stride-2 3x3 on the phase-split input;  2 + a: its data gradient, rows a
  const float* w2;  // != NULL: folded form (mil_tc_shape_fold): the 1x1 projection weight [cout][cin2][1][1]
  int cin2;
};
int mil_launch_pack_tc_table(const MilTcPackJob* jobs, int count, cudaStream_t s);
// out = the four (row parity, column parity) phases of `in` at half resolution, as 4 * cb chunk planes (bf16)
int mil_launch_split2(const void* in, const MilPF8& gin, void* out, const MilPF8& gout, cudaStream_t s);
// out (half resolution) = in at the even (y, x) positions (bf16): the input of a stride-2 1x1 projection
int mil_launch_subsample2(const void* in, const MilPF8& gin, void* out, const MilPF8& gout, cudaStream_t s);
// out (geometry 2x) = zero-stuffed copy of in: out(n, 2y, 2x) = in(n, y, x), zero elsewhere (bf16)
int mil_launch_upsample2(const void* in, const MilPF8& gin, void* out, const MilPF8& gout, cudaStream_t s);

// tcgen05 weight gradient (mil_wgrad_tc.cu): 3x3 / 1x1, stride 1, bf16
bool mil_wgrad_tc_supported(int dtype, int ks, int stride, int cin, int cout);
size_t mil_wgrad_tc_partial_floats(const MilPF8& gx, const MilPF8& gz, int ks);
int mil_launch_wgrad_tc_partials(const void* x, const MilPF8& gx, const void* dz, const MilPF8& gz, float* partial,
                                 int ks, int* ctas_out, long long* rec_out, cudaStream_t s);
int mil_launch_wgrad_tc(const void* x, const MilPF8& gx, const void* dz, const MilPF8& gz, float* partial, float* dw,
                        float* db, int ks, cudaStream_t s);

// stem backward with the un-pool fused into the A-operand construction (mil_stem_wgrad.cu): partial records of the
// space-to-depth weight gradient from xs, the POOLED gradient g and the arg-max records am (mil_stem_unpool.cuh)
bool mil_stem_wgrad_supported(const MilPF8& gp);
size_t mil_stem_wgrad_partial_floats();
int mil_launch_stem_wgrad(const void* xs, const MilPF8& gx, const void* g, const MilPF8& gp, const void* am, float* partial,
                          int* ctas_out, long long* rec_out, cudaStream_t s);
// 3x3 / stride-2 convolution: x given as its phase-split copy xs2 (4 * cb planes), dz at the output resolution
int mil_launch_wgrad_tc_s2(const void* xs2, const MilPF8& gs, const void* dz, const MilPF8& gz, float* partial, float* dw,
                           float* db, int cin, cudaStream_t s);

// stem convolution with the max-pool fused into its epilogue (mil_stem_pool.cu): xs -> pooled map + arg-max (+ sign
// mask); supported for even conv-map sizes and padded rows of at most 127 pixels
bool mil_stem_conv_pool_supported(const MilPF8& gp, int hc);
int mil_launch_stem_conv_pool(const void* xs, const MilPF8& gi, const void* wtc, const float* bias4, void* pooled,
                              const MilPF8& gp, void* argmax, void* mask_out, int hc, cudaStream_t s);

// stem on the tensor cores (mil_stem_tc.cu)
MilPF8 mil_stem_tc_geom_in(int n, int side);    // space-to-depth input: 12 channels, conv-output resolution, pad 2
MilPF8 mil_stem_tc_geom_conv(int n, int side);  // conv1 output: 20 channels, pad 2
size_t mil_stem_tc_wpack_floats();
size_t mil_stem_tc_wtc_bytes();
size_t mil_stem_tc_partial_floats(int n, int side);
size_t mil_stem_tc_argmax_bytes(const MilPF8& gp);  // arg-max records of the pooled map (mil_stem_unpool.cuh)
bool mil_stem_tc_fused_pool(const MilPF8& gp, int side);
int mil_launch_stem_tc_fwd(const void* x, int x_u8, const int* idx, int n, int side, const float* w, const float* b, void* xs,
                           void* convout, float* wp, void* wtc, void* pooled, const MilPF8& gp, uint8_t* argmax,
                           cudaStream_t s, void* pooled_mask = nullptr);
int mil_launch_stem_tc_bwd(const void* xs, int n, int side, const void* g, const MilPF8& gp, const uint8_t* argmax,
                           void* dy, float* partial, float* dw, float* db, cudaStream_t s);

// the stem's layout / pool kernels on their own, for any even channel count C (conv map = 4 * C channels (co, a, b) at
// the pooled resolution; arg-max records of mil_stem_tc_argmax_bytes(gp) bytes)
int mil_launch_stem_s2d4(const void* x, int x_u8, const int* idx, int side, void* xs, const MilPF8& gi, cudaStream_t s);
int mil_launch_stem_pool4(const void* cv, const MilPF8& gc, int hc, void* pooled, const MilPF8& gp, void* argmax,
                          cudaStream_t s);
int mil_launch_stem_unpool4(const void* g, const MilPF8& gp, const void* argmax, void* dy, const MilPF8& gc, cudaStream_t s);
