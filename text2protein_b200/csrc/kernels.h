// Internal C++ launch API of the t2p CUDA kernels (one function per kernel family).
// Activations are NHWC ("pixels x channels") so that every convolution / projection is a
// row-major [M = B*H*W, K] x [N, K]^T contraction; weights are packed [N][taps][Cin].
#pragma once
#include "common.h"

namespace t2p {

// ----------------------------------------------------------------------------- implicit GEMM
// out[m, n] = alpha * ( sum_k A[m, k] * Wt[n, k] + bias[n] + rowbias[m / rows_per_sample, n]
//                       + residual[m, n] )
// A is the im2col view (ksize 1 or 3, stride 1, zero pad ksize/2) of up to two NHWC sources
// concatenated along channels (the UNet skip concat is never materialised).
struct ConvGemmArgs {
  const void* a0 = nullptr;
  int c0 = 0;
  const void* a1 = nullptr;
  int c1 = 0;
  int B = 1, H = 1, W = 1;  // geometry of the A sources; M = B*H*W
  int ksize = 1;
  const void* w = nullptr;  // [N][ksize*ksize*(c0+c1)], same dtype as A
  int N = 0;
  const float* bias = nullptr;     // [N]
  const float* rowbias = nullptr;  // [M / rows_per_sample][N]
  int rows_per_sample = 0;
  const void* residual = nullptr;  // [M][N], out dtype
  int res_up = 0;                  // 1: residual is [B][H/2][W/2][N] and is nearest-upsampled x2
  float alpha = 1.f;
  void* out = nullptr;
  int out_dtype = kBF16;
  // optional per-(sample, channel) sum / sum-of-squares of the stored output (GroupNorm stats)
  float* stat_sum = nullptr;  // [M / rows_per_sample][N]
  float* stat_sq = nullptr;
};

// bf16 tcgen05 / TMEM / TMA path (sm_100a).  A, W are bf16.
void conv_gemm_tc(const ConvGemmArgs& a, cudaStream_t st);
// fp32 or bf16 SIMT path (verification mode and tiny-channel edge layers). dtype of A and W.
void conv_gemm_simt(const ConvGemmArgs& a, int in_dtype, cudaStream_t st);

}  // namespace t2p
