// Fused (flash-style, never materialising [T, Tk]) softmax attention over map tokens / text tokens.
//
//   out[b, t, h*dv : (h+1)*dv] = softmax_j( scale * <q[b,t,h,:], k[b,j,h,:]> ) @ v[b,j,h,:]
//
// q / k / v / out are strided token-major views (row pitch ld*, head h at column h*d), so the fused
// QKV projection output is consumed in place.  Covers AttnBlockpp (1 head, d = C, scale C^-0.5;
// layers.py:160-176) and CrossAttention self / cross (8 heads, d = C/8, scale d^-0.5; attention.py:170-193).
//
// Two implementations:
//  * attention_simt : fp32 accumulate, one warp per query row; the exact path used by the fp32
//    verification mode and as the in-library reference for the tensor-core kernel.
//  * attention_mma  : bf16 tensor-core kernel (attention_mma.cu).
#include "kernels.h"

namespace t2p {
namespace {

template <typename T>
__device__ __forceinline__ float ldv(const T* p);
template <>
__device__ __forceinline__ float ldv<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldv<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stv(float* p, float v) { *p = v; }
__device__ __forceinline__ void stv(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

constexpr int MAXE = 32;  // d <= 1024

template <typename T>
__global__ void attention_simt_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                      T* __restrict__ out, int B, int heads, int Tq, int Tk, int d, long long ldq,
                                      long long ldk, long long ldv_, long long ldo, float scale) {
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(B) * heads * Tq;
  if (warp >= total) return;
  const int t = static_cast<int>(warp % Tq);
  const int h = static_cast<int>((warp / Tq) % heads);
  const int b = static_cast<int>(warp / (static_cast<long long>(Tq) * heads));
  const T* qr = q + (static_cast<long long>(b) * Tq + t) * ldq + h * d;
  float qv[MAXE], acc[MAXE];
#pragma unroll
  for (int e = 0; e < MAXE; ++e) {
    const int i = e * 32 + lane;
    qv[e] = (i < d) ? ldv(qr + i) * scale : 0.f;
    acc[e] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j < Tk; ++j) {
    const T* kr = k + (static_cast<long long>(b) * Tk + j) * ldk + h * d;
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
      const int i = e * 32 + lane;
      if (i < d) s = fmaf(qv[e], ldv(kr + i), s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, s);
    const float corr = __expf(m - mn);
    const float p = __expf(s - mn);
    l = l * corr + p;
    const T* vr = v + (static_cast<long long>(b) * Tk + j) * ldv_ + h * d;
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
      const int i = e * 32 + lane;
      if (i < d) acc[e] = fmaf(p, ldv(vr + i), acc[e] * corr);
    }
    m = mn;
  }
  const float inv = 1.f / l;
  T* orow = out + (static_cast<long long>(b) * Tq + t) * ldo + h * d;
#pragma unroll
  for (int e = 0; e < MAXE; ++e) {
    const int i = e * 32 + lane;
    if (i < d) stv(orow + i, acc[e] * inv);
  }
}

}  // namespace

void attention_simt(const AttnArgs& a, int dtype, cudaStream_t st) {
  T2P_CHECK(a.d <= 32 * MAXE, "head dim too large");
  const long long warps = static_cast<long long>(a.B) * a.heads * a.Tq;
  const unsigned blocks = static_cast<unsigned>(cdiv64(warps, 4));
  if (dtype == kF32)
    attention_simt_kernel<float><<<blocks, 128, 0, st>>>(
        static_cast<const float*>(a.q), static_cast<const float*>(a.k), static_cast<const float*>(a.v),
        static_cast<float*>(a.out), a.B, a.heads, a.Tq, a.Tk, a.d, a.ldq, a.ldk, a.ldv, a.ldo, a.scale);
  else
    attention_simt_kernel<__nv_bfloat16><<<blocks, 128, 0, st>>>(
        static_cast<const __nv_bfloat16*>(a.q), static_cast<const __nv_bfloat16*>(a.k),
        static_cast<const __nv_bfloat16*>(a.v), static_cast<__nv_bfloat16*>(a.out), a.B, a.heads, a.Tq, a.Tk, a.d,
        a.ldq, a.ldk, a.ldv, a.ldo, a.scale);
  T2P_LAUNCH_CHECK();
}

}  // namespace t2p
