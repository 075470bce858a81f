/* mil_b200.h -- C ABI of the B200-native ResNet-26 + attention-MIL hot path (libmil_b200.so).
 *
 * The reference (frankenz/Deep-convolutional-neural-network-ResNet-26-and-Attention-network) is pure Python
 * and has no FFI; the drop-in boundary is its `model.Attention` nn.Module (gbm/model.py:114-264) as used by
 * gbm/classify_combined.py:432,446-447,518.  This header is the native interface a binding for that class
 * calls; each entry point names the reference lines it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`; the library allocates nothing and
 *     keeps no state between calls: the caller owns inputs, outputs, parameters, gradients and workspaces;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no call synchronises;
 *   - return value 0 = success; non-zero = error, message available from mil_last_error() (thread local);
 *     nothing throws across this boundary;
 *   - parameters are the reference's 65 state-dict tensors, fp32, PyTorch layouts, passed as an array of 65
 *     device pointers in state-dict order (mil_param_name(i) lists them);  gradients are ACCUMULATED (+=)
 *     into one flat fp32 buffer of mil_param_total() floats, tensor i at float offset mil_param_offset(i);
 *   - dtype: MIL_DTYPE_F32 = fp32 storage + FFMA kernels (the 1e-4 check mode),
 *            MIL_DTYPE_BF16 = bf16 activations, fp32 accumulation, tcgen05 tensor-core convolutions.
 *   - there is NO CPU fallback: every compute entry point needs an sm_100a device.
 */
#ifndef MIL_B200_H_
#define MIL_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIL_ABI_VERSION 1
#define MIL_DTYPE_F32 0
#define MIL_DTYPE_BF16 1
#define MIL_NUM_PARAMS 65
#define MIL_FEATURES 80     /* L, gbm/model.py:120 */
#define MIL_HEAD_SUMS 16    /* doubles exchanged by AR-2 (see mil_head_scores) */
#define MIL_HEAD_SCALARS 32 /* floats written by mil_head_finalize */
/* offsets inside the MIL_HEAD_SCALARS block */
#define MIL_SCAL_MTERM 0   /* [3] Mterm == slide logits (gbm/model.py:227-233) */
#define MIL_SCAL_YPRED 3   /* [3] softmax(logits)         (gbm/model.py:235) */
#define MIL_SCAL_DM 6      /* [3] d loss / d Mterm (kept for the backward pass) */
#define MIL_SCAL_SUMG 9    /* [3] sum_n g_nk (the L1 normaliser of F.normalize, gbm/model.py:213) */
#define MIL_SCAL_LOSS 12   /* label-smoothed weighted CE (gbm/model.py:241, nnBlocks.py:87-138) */
#define MIL_SCAL_ATERM_MU 13
#define MIL_SCAL_ATERM_VAR 14
#define MIL_SCAL_KLD 15
#define MIL_SCAL_YHAT 16   /* argmax class as float */
#define MIL_SCAL_ERROR 17  /* 1 - [yhat == Y] */

/* ---- library / model description ------------------------------------------------------------------ */
int mil_abi_version(void);
const char* mil_last_error(void);
long long mil_kernel_launch_count(void);          /* kernels this library has launched in this process so far */
int mil_param_count(void);                        /* 65 */
const char* mil_param_name(int i);                /* reference state-dict key, e.g. "cnn.module.layer1.0.conv1.weight" */
int mil_param_shape(int i, int* ndim, long long shape4[4]);
long long mil_param_offset(int i);                /* float offset inside the flat gradient buffer */
long long mil_param_total(void);                  /* 640967 */

/* Runtime switches (cross-checks; the environment variables MIL_B200_<NAME> give the initial values):
 *   "disable_tc"   1 = CUDA-core (FFMA) kernels only, with the tensor-core path's bf16 rounding points
 *   "stem_unfused" 1 = stem as separate conv / pool / unpool / weight-gradient kernels
 *   "no_pdl"       1 = ordinary stream-ordered launches instead of programmatic dependent launches of the tcgen05 kernels
 * A forward pass and its backward pass must run under the same settings (they size the workspace). */
int mil_set_option(const char* name, int value);
int mil_get_option(const char* name, int* value);

/* ---- feature extractor: ResNet.forward (gbm/model.py:50-61) + BasicResBlock (nnBlocks.py:175-189) ----
 * bag   : fp32 NCHW [n_bag,3,side,side], the caller's tensor as handed to Attention.forward (gbm/model.py:189)
 * idx   : optional int32[n_tiles] gather list = the train-mode 20 % subsample (gbm/model.py:193-194);
 *         NULL => tiles 0..n_tiles-1
 * H     : fp32 [n_tiles,80] features ('Fterm')
 * ws    : workspace of mil_extractor_workspace_bytes() bytes; forward leaves the activations backward needs in
 *         it, so pass the SAME workspace (untouched) to mil_extractor_backward.                              */
size_t mil_extractor_workspace_bytes(int n_tiles, int side, int dtype);
int mil_extractor_forward(const void* const* params_host, const float* bag, const int32_t* idx, int n_tiles, int side,
                          int dtype, void* ws, size_t ws_bytes, float* H, void* stream);
/* Same with raw 8-bit tiles (uint8 NCHW [n_bag,3,side,side]): the reference's CPU transform ToTensor() +
 * Normalize(0.5, 0.5) (RoiBuilder.py:199-202) is fused into the stem's load -- a quarter of the host->device bytes.
 * bf16 mode only.  The backward entry points never read the bag in bf16 mode (pass the same pointer).            */
int mil_extractor_forward_u8(const void* const* params_host, const uint8_t* bag, const int32_t* idx, int n_tiles,
                             int side, int dtype, void* ws, size_t ws_bytes, float* H, void* stream);
/* Forward only -- attention-map extraction / validation under torch.no_grad() (gbm/classify_combined.py:221-298,
 * :302-357): same kernels, but nothing is kept for a backward pass: three rotating map buffers instead of 25 saved
 * maps, no sign masks, no arg-max records (workspace ~0.5 MB per 224x224 tile instead of ~3.5 MB).  bag_is_u8 selects
 * raw 8-bit tiles (bf16 mode).  The workspace cannot be handed to mil_extractor_backward.                        */
size_t mil_extractor_infer_workspace_bytes(int n_tiles, int side, int dtype);
int mil_extractor_infer(const void* const* params_host, const void* bag, int bag_is_u8, const int32_t* idx, int n_tiles,
                        int side, int dtype, void* ws, size_t ws_bytes, float* H, void* stream);
/* autograd of the above (gbm/classify_combined.py:447): dH fp32 [n_tiles,80] -> grads_flat += d/d(params).
 * No gradient flows to `bag` (it is detached, gbm/model.py:194,196).                                         */
int mil_extractor_backward(const void* const* params_host, const float* bag, const int32_t* idx, int n_tiles, int side,
                           int dtype, void* ws, size_t ws_bytes, const float* dH, float* grads_flat, void* stream);

/* Same, for the bucketed gradient all-reduce overlapped with backward (north_star; replaces DataParallel's
 * reduce-add onto GPU 0): layer_events_host = 4 cudaEvent_t handles (or NULL entries).  Event [l] is recorded on
 * `stream` as soon as every gradient of layer l+1 is final -- [3] also covers fc and the head's parameters, which
 * precede the extractor in backward order -- so the caller can all-reduce that slice of grads_flat on another
 * stream while the remaining layers are still being processed.  conv1 / layer1 are final when the call's work is. */
int mil_extractor_backward_staged(const void* const* params_host, const float* bag, const int32_t* idx, int n_tiles,
                                  int side, int dtype, void* ws, size_t ws_bytes, const float* dH, float* grads_flat,
                                  void* const* layer_events_host, void* stream);

/* test / debugging aid: copy one saved activation out of a forward workspace as fp32 NCHW.
 * which = -1: stem output (after max-pool); 2*(3*layer+block): the block's inner activation h;
 * 2*(3*layer+block)+1: the block's output y  (layer 0..3, block 0..2).                                        */
int mil_extractor_read_activation(int n_tiles, int side, int dtype, const void* ws, int which, float* nchw,
                                  void* stream);

/* test / debugging aid: the NEXT mil_extractor_backward calls on this thread also write one intermediate
 * gradient as fp32 NCHW to `nchw` (NULL switches it off).  which = 0: gradient w.r.t. the pre-activation that
 * produced the input of block (layer, block); which = 1: w.r.t. the pre-activation of its first convolution. */
int mil_debug_dump_gradient(int layer, int block, int which, float* nchw);

/* ---- MIL head: Attention.forward after the CNN (gbm/model.py:200-246) ------------------------------
 * The head works on the LOCAL shard (n tiles) of a bag of n_global tiles; between the phases the caller
 * all-reduces (sum) the small double arrays across the ranks that share the bag (single GPU: nothing to do).
 *   stats  double[160] : sum_n x, sum_n x^2 per feature        (BatchNorm1d batch statistics, gbm/model.py:105-109)
 *   sums   double[16]  : sum g[3], sum g*b[3], sum raw[3], Gram(raw) upper triangle[6]  (gbm/model.py:211-229)
 *   bnsums double[160] : sum dHz, sum dHz*xhat per feature     (BatchNorm1d backward)                         */
size_t mil_head_workspace_bytes(int n);
int mil_head_stats(const float* H, int n, void* ws, size_t ws_bytes, double* stats, void* stream);
/* drop: optional fp32 {0,1} keep mask [n,80] of Dropout(0.25) (gbm/model.py:107,110); NULL in eval mode.
 * raw,g: fp32 [n,3] (attention scores before / after softplus+mask mix); b: fp32 [n] ('Bterm').              */
int mil_head_scores(const void* const* params_host, const float* H, const float* drop, int n, long long n_global,
                    const double* stats, float* raw, float* g, float* b, void* ws, size_t ws_bytes, double* sums,
                    void* stream);
/* Y: int64[1] label; class_w: optional fp32[3] (CrossEntropyWithProbs weight, gbm/model.py:127);
 * A ('Aterm'), wroi ('wROIs'): fp32 [3,n]; scal: fp32[MIL_HEAD_SCALARS].                                     */
int mil_head_finalize(const double* sums, const double* stats, long long n_global, const long long* Y,
                      const float* class_w, int n, const float* g, const float* b, float* A, float* wroi,
                      float* scal, void* stream);
/* gloss: optional fp32[1] upstream gradient of the loss (NULL => 1).  Accumulates the head's parameter
 * gradients (local contributions) into grads_flat; writes dHz, dHi fp32 [n,80] and bnsums.                   */
int mil_head_backward_a(const void* const* params_host, const float* H, const float* drop, int n, long long n_global,
                        const double* stats, const float* raw, const float* g, const float* b, const float* scal,
                        const float* gloss, float* dHz, float* dHi, float* grads_flat, void* ws, size_t ws_bytes,
                        double* bnsums, void* stream);
int mil_head_backward_b(const void* const* params_host, const float* H, int n, long long n_global, const double* stats,
                        const double* bnsums, const float* dHz, const float* dHi, float* dH, void* stream);

/* ---- layer-level operators on PF8 activation buffers (used by the layer-wise parity tests) ------------
 * PF8 = padded-flat 8-channel-chunk layout, see csrc/mil_common.cuh.  Buffers must be zero-filled once by the
 * caller (guards) and are mil_pf8_bytes() large.  impl: 0 = auto, 1 = CUDA-core direct, 2 = tcgen05.        */
size_t mil_pf8_bytes(int n, int c, int h, int w, int dtype);
int mil_to_pf8(int dtype, const float* nchw, void* pf8, int n, int c, int h, int w, void* stream);
int mil_from_pf8(int dtype, const void* pf8, float* nchw, int n, int c, int h, int w, void* stream);
/* bf16 only: out (ho x wo = the stride-2 convolution's INPUT size) = zero-stuffed copy of in (h x w):
 * out(n,2y,2x) = in(n,y,x), zero elsewhere.  The stride-2 data / weight gradients are the stride-1 ones of it. */
int mil_upsample2_pf8(const void* in, int n, int c, int h, int w, void* out, int ho, int wo, void* stream);
size_t mil_conv_workspace_bytes(int n, int cin, int hi, int wi, int cout, int ho, int wo, int ks);
/* out = epilogue(conv(x, w)):  epi 0: lrelu(acc+bias+res), 1: (acc+res)*lrelu'(act), 2: acc+bias+res.
 * transposed=1 computes the data gradient (x has the conv's OUTPUT geometry, out its INPUT geometry).
 * w is the PyTorch [cout,cin,ks,ks] fp32 weight.  res has out's geometry, EXCEPT for the stride-2 data gradient
 * (transposed=1, stride=2), where it is the half-resolution gradient of the block's 1x1 projection branch
 * ([n,cin] at x's spatial size) and is added at the even positions -- how nnBlocks.py:183-187's two branches meet in
 * backward.  In bf16 mode the stride-2 3x3 convolution runs in the extractor's own tensor-core forms (forward on the
 * phase-split input, data gradient per input row parity) whenever they fit.                                  */
int mil_conv_pf8(int dtype, int impl, int transposed, const void* x, int n, int cx, int hx, int wx, const float* w,
                 int cout, int cin, int ks, int stride, const float* bias, const void* res, const void* act,
                 void* out, int ho, int wo, int epi, void* ws, size_t ws_bytes, void* stream);
/* dw += wgrad(x, dz), db += sum dz (db may be NULL) */
int mil_conv_wgrad_pf8(int dtype, int impl, const void* x, int n, int cin, int hi, int wi, const void* dz, int cout,
                       int ho, int wo, int ks, int stride, float* dw, float* db, void* ws, size_t ws_bytes,
                       void* stream);

/* Stem at layer level (gbm/model.py:24-26,51-53: conv 7x7 / stride 2 / pad 3 + bias, LeakyReLU(0.1), max-pool 3x3 /
 * stride 2 / pad 1) and its backward pass (weight + bias gradient; the bag is detached, gbm/model.py:194,196).
 * bag: fp32 NCHW [n,3,side,side]; pooled / g: PF8 [n,20,h0,h0] (the pooled map / the gradient w.r.t. its
 * PRE-activation, i.e. already multiplied by LeakyReLU'); dw [20,3,7,7], db [20] are ACCUMULATED.  impl: 0 = what the
 * extractor runs, 1 = CUDA-core kernels, 2 = tcgen05 (bf16).  mil_stem_backward needs the workspace of the
 * mil_stem_forward call on the same bag (space-to-depth input, arg-max records).                                 */
size_t mil_stem_workspace_bytes(int n, int side, int dtype);
int mil_stem_forward(int dtype, int impl, const float* bag, int n, int side, const float* w, const float* b, void* pooled,
                     void* ws, size_t ws_bytes, void* stream);
int mil_stem_backward(int dtype, int impl, const float* bag, int n, int side, const void* g, float* dw, float* db, void* ws,
                      size_t ws_bytes, void* stream);

/* ---- wide-channel kernel family at layer level -------------------------------------
 * The alt_resnet.py parameterisation of the extractor (alt_resnet.py:24-32 conv3x3 / conv1x1 without bias, :35-67
 * BasicBlock with ReLU, :70-145 widths 64/128/256/512) on its own tcgen05 kernels: the K loop streams weight slabs
 * next to the input planes, output channels are tiled by 128.  bf16 only; PF8 buffers as above.
 *   mode 0: ks = 1 / 3, stride 1 -- forward (transposed = 0) or data gradient (transposed = 1; x = the gradient
 *           w.r.t. the conv's output);  mode 1: 3x3 / stride 2 forward, x = the phase-split input (mil_split2_pf8:
 *           4 * cin channels at the OUTPUT resolution h x w);  mode 2: the 7x7 / stride-2 stem in space-to-depth-by-4
 *           form, x = [n, 48, h, w] with x[(c,ry,rx)][Y][X] = tile[c][4Y+ry][4X+rx], out = [n, 4*wcout, h, w] with channel
 *           (co, a, b) = conv1(tile)[co][2Y+a][2X+b].
 * wt: PyTorch weight [wcout][wcin][ks][ks] fp32.  epi: 0 = act(acc + bias + res), 1 = (acc + res) * act'(act map),
 * 2 = acc + bias + res;  slope: negative slope of the activation (0 = ReLU, alt_resnet.py:49).  tm: 128-pixel tiles per
 * weight slab (0 = default).  The stride-2 data / weight gradients are the stride-1 ones of the zero-stuffed output
 * gradient (mil_upsample2_pf8).                                                                                 */
size_t mil_wide_conv_workspace_bytes(int mode, int transposed, int wcout, int wcin, int ks);
int mil_wide_conv_pf8(int mode, int transposed, const void* x, int n, int cx, int h, int w, const float* wt, int wcout,
                      int wcin, int ks, const float* bias, const void* res, const void* act, void* out, int epi,
                      float slope, int tm, void* ws, size_t ws_bytes, void* stream);
/* dw[cout][cin][ks][ks] += wgrad(x, dz) (no bias gradient: alt_resnet's convolutions have none); cout a multiple of
 * 128; ks = 7: the stem form (x = [n,48,h,w] space-to-depth input, dz = [n, 4*C, h, w], dw = [C][3][7][7]).       */
size_t mil_wide_wgrad_workspace_bytes(int n, int cin, int cout, int h, int w, int ks);
int mil_wide_wgrad_pf8(const void* x, int n, int cin, int h, int w, const void* dz, int cout, int ks, float* dw, void* ws,
                       size_t ws_bytes, void* stream);
/* out = the four (row, column) parity phases of in at half resolution, as 4 * c channels (plane = phase * c/8 + chunk) */
int mil_split2_pf8(const void* in, int n, int c, int h, int w, void* out, void* stream);
/* the inverse: out [n, c, h, w] = the phases of in interleaved back (pad pixels zero) */
int mil_merge2_pf8(const void* in, int n, int c, int h, int w, void* out, void* stream);
/* The stride-2 3x3 convolution's gradients at the OUTPUT resolution ho x wo (no zero-stuffing):
 *   weight gradient from the phase-split input xs2 (mil_split2_pf8) and the output gradient dz;
 *   data gradient: mil_wide_conv_pf8 with mode = 3 + 2a + b computes input parity phase (a, b) from x = dz
 *   (cx = wcout channels), out = [n, wcin, ho, wo]; act / res are maps of that phase.                               */
size_t mil_wide_wgrad_s2_workspace_bytes(int n, int cin, int cout, int ho, int wo);
int mil_wide_wgrad_s2_pf8(const void* xs2, int n, int cin, int ho, int wo, const void* dz, int cout, float* dw, void* ws,
                          size_t ws_bytes, void* stream);

/* ---- the wide extractor as a whole --------------------------------------------------
 * alt_resnet.py's ResNet as the tile feature extractor of the same MIL head: conv1 7x7/2 (3 -> stem, no bias), ReLU,
 * max-pool 3x3/2, four layers of BasicBlocks (alt_resnet.py:35-67; 1x1/2 projections where the shape changes, :107-123),
 * average pool, fc widths[3] -> features WITH bias (:90).  `slope` is the activation's negative slope (0 = ReLU, :49).
 * resnet18 of alt_resnet.py:157-165 = layers {2,2,2,2}, widths {64,128,256,512}, stem 64; features must be 80 (L).
 * Parameters: the state dict of Attention with cnn = DataParallel(alt_resnet.ResNet(BasicBlock, layers, num_classes=80)),
 * in its own order (mil_wide_param_*), fp32; gradients accumulate into one flat buffer like the ResNet-26 entry points.
 * bf16 only (bag: fp32 or uint8 NCHW).  The head's entry points (mil_head_*) are shared: hand them an array of
 * MIL_NUM_PARAMS pointers with the head's tensors at their ResNet-26 indices.                                      */
typedef struct MilWideDesc {
  int layers[4];
  int widths[4];
  int stem;
  int features;
  float slope;
} MilWideDesc;
int mil_wide_param_count(const MilWideDesc* desc);
int mil_wide_param_info(const MilWideDesc* desc, int i, char* name_out, int name_cap, int* ndim, long long shape4[4],
                        long long* offset);
long long mil_wide_param_total(const MilWideDesc* desc);
size_t mil_wide_workspace_bytes(const MilWideDesc* desc, int n_tiles, int side);
int mil_wide_forward(const MilWideDesc* desc, const void* const* params_host, const void* bag, int bag_is_u8,
                     const int32_t* idx, int n_tiles, int side, void* ws, size_t ws_bytes, float* H, void* stream);
int mil_wide_backward(const MilWideDesc* desc, const void* const* params_host, int n_tiles, int side, void* ws,
                      size_t ws_bytes, const float* dH, float* grads_flat, void* stream);

/* ---- training-loop glue ------------------------------------------------------
 * One Adam step over the FLAT parameter / gradient buffers (state-dict order, mil_param_offset) in one launch:
 * what `optim.Adam(classifier.parameters(), lr=2e-4)` + `optimizer.step()` do tensor by tensor in the reference
 * (gbm/classify_combined.py:519, :450-452).  Same arithmetic as torch.optim.Adam (L2 weight decay, no amsgrad):
 *   g += weight_decay * p;  m += (1 - beta1) * (g - m);  v = beta2 * v + (1 - beta2) * g * g;
 *   p -= step_size * m / (sqrt(v) / bc2_sqrt + eps),   step_size = lr / (1 - beta1^t),  bc2_sqrt = sqrt(1 - beta2^t)
 * (the caller computes step_size and bc2_sqrt in double precision, as torch does).                           */
int mil_adam_step(float* params_flat, const float* grads_flat, float* exp_avg, float* exp_avg_sq, long long count,
                  float step_size, float beta1, float beta2, float bc2_sqrt, float eps, float weight_decay,
                  void* stream);

/* The same step with the six scalars {step_size, beta1, beta2, bc2_sqrt, eps, weight_decay} read from DEVICE memory
 * (float hyper[6]): the form a captured CUDA graph replays -- the host refreshes `hyper` between replays.        */
int mil_adam_step_dev(float* params_flat, const float* grads_flat, float* exp_avg, float* exp_avg_sq, long long count,
                      const float* hyper, void* stream);

/* ---- tile ingest -----------------------------------------------------------------
 * The per-tile finalisation of the reference's loader (RoiBuilder.py:193-210: ToPILImage, Pad(100), RandomCrop(roi),
 * Resize(side), RandomHorizontalFlip, RandomVerticalFlip; :203-208 without the augmentations for validation) on the
 * cached 8-bit tiles, bit-identical to Pillow's antialiased bilinear resampling:
 *   rois   : device uint8 [n_tiles, roi, roi, 3] (HWC, the `.npy` tile cache, RoiBuilder.py:215-238)
 *   crops  : device int32 [n_tiles, 2] = (top, left) of the crop inside the padded (roi + 2 pad) image, or NULL for
 *            no pad / crop;  flips: device uint8 [n_tiles], bit 0 = horizontal, bit 1 = vertical flip, or NULL
 *   bounds : device int32 [side, 2] = (first input pixel, count) and coef: device int32 [side, ksize] = Pillow's
 *            fixed-point triangle weights (PRECISION_BITS = 22) per output position; bounds_host = host copy of bounds
 *   out    : device uint8 [n_tiles, 3, side, side] -- the 8-bit bag mil_extractor_forward_u8 consumes (ToTensor +
 *            Normalize(.5, .5) are fused into the stem's load)                                                     */
int mil_ingest_tiles_u8(const void* rois, int n_tiles, int roi, const int* crops, int pad, const unsigned char* flips,
                        int side, const int* bounds, const int* coef, int ksize, const int* bounds_host, void* out,
                        void* stream);

/* ---- attention-map export -------------------------------------------------------
 * out = (in - min(in)) / (max(in) - min(in)) over all `count` elements: the `plt.Normalize()(attn)` /
 * `(A - A.min()) / (A.max() - A.min())` scaling the reference applies to an attention map before writing the per-tile
 * `x y weight` heat-map files (gbm/classify.py:207-212, gbm/classify_combined.py:163).  minmax: device float[2]
 * scratch that receives (min, max).  A constant map (max == min) comes back as zeros.                          */
int mil_minmax_normalize(const float* in, float* out, long long count, float* minmax, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIL_B200_H_ */
