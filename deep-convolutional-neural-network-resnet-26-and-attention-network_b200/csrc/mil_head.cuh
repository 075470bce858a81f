// Host-side declarations of the MIL-head launchers (mil_head.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#define MIL_HEAD_NSUMS 16  // sum g[3], sum g*b[3], sum raw[3], raw Gram upper triangle[6], spare
// layout of the small fp32 scalar block written by head_finalize
#define MIL_SC_M 0      // Mterm / logits [3]
#define MIL_SC_P 3      // y_pred [3]
#define MIL_SC_DM 6     // d loss / d Mterm [3]
#define MIL_SC_S 9      // sum_n g [3]
#define MIL_SC_LOSS 12
#define MIL_SC_AMU 13
#define MIL_SC_AVAR 14
#define MIL_SC_KLD 15
#define MIL_SC_YHAT 16
#define MIL_SC_ERR 17
#define MIL_SC_COUNT 32

// device pointers to the head's parameters (fp32, PyTorch layouts; reference gbm/model.py:105,140-153)
struct MilHeadParams {
  const float* weight_mask;  // [3]
  const float* bn_w;         // context.bn.weight [80]
  const float* bn_b;         // context.bn.bias   [80]
  const float* att_w1;       // attention.lin1.weight [40][80]
  const float* att_b1;       // [40]
  const float* att_w2;       // attention.lin2.weight [3][40]
  const float* att_b2;       // [3]
  const float* buf_w1;       // buffer.lin1.weight [40][80]
  const float* buf_b1;       // [40]
  const float* buf_w2;       // buffer.classifier.weight [1][40]
  const float* buf_b2;       // [1]
};
struct MilHeadGrads {
  float *weight_mask, *bn_w, *bn_b, *att_w1, *att_b1, *att_w2, *att_b2, *buf_w1, *buf_b1, *buf_w2, *buf_b2;
};

size_t mil_head_part_doubles(int n);
size_t mil_head_bwd_partial_floats(int n);
int mil_launch_head_stats(const float* H, int n, double* part_ws, double* stats, cudaStream_t s);
int mil_launch_head_scores(const MilHeadParams& P, const float* H, const float* drop, int n, long long n_global,
                           const double* stats, float* raw, float* g, float* b, double* part_ws, double* sums,
                           cudaStream_t s);
int mil_launch_head_finalize(const double* sums, const double* stats, long long n_global, const long long* Y,
                             const float* class_w, int n, const float* g, const float* b, float* A, float* wroi,
                             float* scal, cudaStream_t s);
int mil_launch_head_bwd_a(const MilHeadParams& P, const MilHeadGrads& G, const float* H, const float* drop, int n,
                          long long n_global, const double* stats, const float* raw, const float* g, const float* b,
                          const float* scal, const float* gloss, float* dHz, float* dHi, float* part_ws,
                          double* bnsums, cudaStream_t s);
int mil_launch_head_bwd_b(const float* bn_w, const float* H, int n, long long n_global, const double* stats,
                          const double* bnsums, const float* dHz, const float* dHi, float* dH, cudaStream_t s);
