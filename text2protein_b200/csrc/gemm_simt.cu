// CUDA-core implicit GEMM with the same contract as conv_gemm_tc (kernels.h).  Used for
//  * the fp32 verification mode of the score network (1e-5 parity against the reference), and
//  * the edge layers whose channel count is not a multiple of 64 (first conv, Cin = 5 or 8).
// 64x64 output tile per 256-thread block, 4x4 micro-tile per thread, K staged 16 at a time.
#include "kernels.h"

namespace t2p {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

struct SimtParams {
  const void* a0; const void* a1;
  int c0, c1;
  int B, H, W, ksize;
  const void* w;
  int M, N;
  const float* bias; const float* rowbias; int rows_per_sample;
  const void* residual; int res_up; float alpha;
  void* out; int out_fp32;
  int rowbias_ld; int out_nchw;
};

template <typename T>
__global__ void __launch_bounds__(256) conv_gemm_simt_kernel(SimtParams p) {
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int ctot = p.c0 + p.c1;
  const int taps = p.ksize * p.ksize;
  const int Ktot = taps * ctot;
  const int pad = p.ksize / 2;
  const int hw = p.H * p.W;

  // each thread stages 4 A elements and 4 B elements per K step: row = tid / 4 (+0), k = (tid % 4) * 4 ..
  const int lrow = tid >> 2;        // 0..63
  const int lk = (tid & 3) * 4;     // 0,4,8,12
  const int am = m0 + lrow;
  int ab = 0, ah = 0, aw = 0;
  if (am < p.M) {
    ab = am / hw;
    const int rem = am - ab * hw;
    ah = rem / p.W;
    aw = rem - ah * p.W;
  }
  const int bn = n0 + lrow;

  float acc[4][4] = {};
  for (int k0 = 0; k0 < Ktot; k0 += TK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lk + j;
      float av = 0.f, bv = 0.f;
      if (k < Ktot) {
        const int tap = k / ctot;
        const int c = k - tap * ctot;
        const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
        const int ih = ah + kh - pad, iw = aw + kw - pad;
        if (am < p.M && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
          const long long pix = (static_cast<long long>(ab) * p.H + ih) * p.W + iw;
          if (c < p.c0) av = ldf(static_cast<const T*>(p.a0) + pix * p.c0 + c);
          else av = ldf(static_cast<const T*>(p.a1) + pix * p.c1 + (c - p.c0));
        }
        if (bn < p.N) bv = ldf(static_cast<const T*>(p.w) + static_cast<long long>(bn) * Ktot + k);
      }
      As[lk + j][lrow] = av;
      Bs[lk + j][lrow] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const int sample = p.rows_per_sample > 0 ? m / p.rows_per_sample : 0;
    long long rrow = m;
    if (p.res_up) {
      const int b = m / hw;
      const int rem = m - b * hw;
      const int h = rem / p.W, w = rem - h * p.W;
      rrow = (static_cast<long long>(b) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.rowbias) v += p.rowbias[static_cast<long long>(sample) * p.rowbias_ld + n];
      if (p.residual) {
        if (p.out_fp32) v += static_cast<const float*>(p.residual)[rrow * p.N + n];
        else v += __bfloat162float(static_cast<const __nv_bfloat16*>(p.residual)[rrow * p.N + n]);
      }
      v *= p.alpha;
      if (p.out_nchw) {
        const int b = m / hw;
        static_cast<float*>(p.out)[(static_cast<long long>(b) * p.N + n) * hw + (m - b * hw)] = v;
      } else if (p.out_fp32) static_cast<float*>(p.out)[static_cast<long long>(m) * p.N + n] = v;
      else static_cast<__nv_bfloat16*>(p.out)[static_cast<long long>(m) * p.N + n] = __float2bfloat16(v);
    }
  }
}

}  // namespace

void conv_gemm_simt(const ConvGemmArgs& a, int in_dtype, cudaStream_t st) {
  T2P_CHECK(a.ksize == 1 || a.ksize == 3, "ksize must be 1 or 3");
  T2P_CHECK(a.stat_part == nullptr, "SIMT path does not fuse GroupNorm statistics");
  SimtParams p{};
  p.a0 = a.a0; p.a1 = a.a1; p.c0 = a.c0; p.c1 = a.c1;
  p.B = a.B; p.H = a.H; p.W = a.W; p.ksize = a.ksize;
  p.w = a.w;
  const long long M = static_cast<long long>(a.B) * a.H * a.W;
  T2P_CHECK(M > 0 && M < (1ll << 31), "M out of range");
  p.M = static_cast<int>(M); p.N = a.N;
  p.bias = a.bias; p.rowbias = a.rowbias; p.rows_per_sample = a.rows_per_sample;
  p.residual = a.residual; p.res_up = a.res_up; p.alpha = a.alpha;
  p.out = a.out; p.out_fp32 = (a.out_dtype == kF32);
  p.rowbias_ld = a.rowbias_ld > 0 ? a.rowbias_ld : a.N;
  p.out_nchw = a.out_nchw;
  if (a.out_nchw) T2P_CHECK(a.out_dtype == kF32 && a.residual == nullptr, "out_nchw is fp32-only, without residual");
  dim3 grid(cdiv(p.M, TM), cdiv(p.N, TN));
  if (in_dtype == kF32) conv_gemm_simt_kernel<float><<<grid, 256, 0, st>>>(p);
  else if (in_dtype == kBF16) conv_gemm_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else T2P_CHECK(false, "unsupported input dtype");
  T2P_LAUNCH_CHECK();
}

}  // namespace t2p
