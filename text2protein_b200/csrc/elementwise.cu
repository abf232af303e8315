// Small helper kernels: timestep-embedding MLP, weight repacking, layout / dtype conversion.
#include <cmath>

#include "kernels.h"

namespace t2p {
namespace {

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

// One block per sample: sinusoidal embedding (layers.py:97-111) -> Linear(nf,4nf) -> Linear(4nf,4nf)
// (no activation in between, ncsnpp.py:227-228) -> SiLU (every consumer applies act(temb) first,
// layers.py:316).  fp32 throughout; output [B][4nf].
__global__ void temb_mlp_kernel(const long long* __restrict__ labels, const float* __restrict__ timesteps, int nf,
                                float neg_coef,
                                const float* __restrict__ w0,
                                const float* __restrict__ b0, const float* __restrict__ w1,
                                const float* __restrict__ b1, float* __restrict__ out) {
  extern __shared__ float sm[];  // emb[nf] | h0[4nf]
  float* emb = sm;
  float* h0 = sm + nf;
  const int b = blockIdx.x;
  const int half = nf / 2;
  // the embedded value is the time conditioning itself (a float for VP models, layers.py:106); integer labels
  // are converted like `timesteps.float()`
  const float t = timesteps ? timesteps[b] : static_cast<float>(labels[b]);
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = expf(static_cast<float>(i) * neg_coef);
    const float a = t * f;
    emb[i] = sinf(a);
    emb[half + i] = cosf(a);
  }
  if ((nf & 1) && threadIdx.x == 0) emb[nf - 1] = 0.f;
  __syncthreads();
  const int d = 4 * nf;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int n = warp; n < d; n += nw) {
    float acc = 0.f;
    for (int k = lane; k < nf; k += 32) acc = fmaf(emb[k], w0[static_cast<long long>(n) * nf + k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) h0[n] = acc + b0[n];
  }
  __syncthreads();
  // the second (d x d) layer is sliced over blockIdx.y; the cheap first layer is recomputed per slice
  const int per = (d + gridDim.y - 1) / gridDim.y;
  const int n_begin = blockIdx.y * per, n_end = min(d, n_begin + per);
  for (int n = n_begin + warp; n < n_end; n += nw) {
    float acc = 0.f;
    for (int k = lane; k < d; k += 32) acc = fmaf(h0[k], w1[static_cast<long long>(n) * d + k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[static_cast<long long>(b) * d + n] = silu_f(acc + b1[n]);
  }
}

template <typename TO>
__device__ __forceinline__ void stv(TO* p, float v);
template <>
__device__ __forceinline__ void stv<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stv<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// conv weight [Cout][Cin][k][k] fp32 -> [Cout][k*k][cin_pad] (channels innermost, zero padded)
// (row pitch `ld` elements: the packed row may be longer than k*k*cin_pad when extra K columns follow)
template <typename TO>
__global__ void pack_conv_kernel(const float* __restrict__ w, int cout, int cin, int kk, int cin_pad, long long ld,
                                 TO* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(cout) * kk * cin_pad;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % cin_pad);
  const int t = static_cast<int>((idx / cin_pad) % kk);
  const int o = static_cast<int>(idx / (static_cast<long long>(cin_pad) * kk));
  const float v = (c < cin) ? w[(static_cast<long long>(o) * cin + c) * kk + t] : 0.f;
  stv(out + static_cast<long long>(o) * ld + static_cast<long long>(t) * cin_pad + c, v);
}

// out[r][col0 + c] = (r == c) for an N x N identity block inside rows of pitch ld
template <typename TO>
__global__ void pack_identity_kernel(int n, long long ld, TO* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(n) * n) return;
  const int c = static_cast<int>(idx % n);
  const int r = static_cast<int>(idx / n);
  stv(out + static_cast<long long>(r) * ld + c, r == c ? 1.f : 0.f);
}

// out[r][c] = in[c][r] (NIN.W is [in, out]; GEMM wants [out][in]); or plain copy when !transpose.
template <typename TO>
__global__ void pack_matrix_kernel(const float* __restrict__ w, int rows, int cols, int transpose, long long ld,
                                   TO* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * cols) return;
  const int c = static_cast<int>(idx % cols);
  const int r = static_cast<int>(idx / cols);
  const float v = transpose ? w[static_cast<long long>(c) * rows + r] : w[idx];
  stv(out + static_cast<long long>(r) * ld + c, v);
}

__global__ void add_vectors_kernel(const float* __restrict__ a, const float* __restrict__ b, int n,
                                   float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (a ? a[i] : 0.f) + (b ? b[i] : 0.f);
}

template <typename TO>
__global__ void convert_kernel(const float* __restrict__ in, long long n, TO* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx < n) stv(out + idx, in[idx]);
}

template <typename T>
__device__ __forceinline__ float ldv(const T* p);
template <>
__device__ __forceinline__ float ldv<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldv<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// NHWC (any dtype) -> NCHW fp32, for debug taps and per-op tests
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, int B, int HW, int C, float* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(B) * HW * C) return;
  const int p = static_cast<int>(idx % HW);
  const int c = static_cast<int>((idx / HW) % C);
  const int b = static_cast<int>(idx / (static_cast<long long>(HW) * C));
  out[idx] = ldv(in + (static_cast<long long>(b) * HW + p) * C + c);
}

// NCHW fp32 -> NHWC dtype T with channel padding (zero)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, int B, int HW, int C, int cpad,
                                    T* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(B) * HW * cpad) return;
  const int c = static_cast<int>(idx % cpad);
  const int p = static_cast<int>((idx / cpad) % HW);
  const int b = static_cast<int>(idx / (static_cast<long long>(HW) * cpad));
  stv(out + idx, c < C ? in[(static_cast<long long>(b) * C + c) * HW + p] : 0.f);
}

// First conv as a plain GEMM: A[m][k] with k = tap * C + c (zero padded to kpad), from the fp32 NCHW state.
// 3x3, stride 1, zero padding 1.  One thread writes 8 consecutive k (16 bytes).
// One thread per pixel, all 9 * C taps of its row of the im2col matrix (C is a compile-time 5 or 8: the tap loops
// unroll, there is no per-element index arithmetic); consecutive lanes are consecutive pixels, so the fp32 NCHW reads
// are coalesced per (channel, tap) and each thread writes its KPAD-wide bf16 row as 16-byte vectors.
template <int C, int KPAD>
__global__ void __launch_bounds__(256) im2col3x3_pixel_kernel(const float* __restrict__ x, int B, int H, int W,
                                                              __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long m = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (m >= static_cast<long long>(B) * H * W) return;
  const int w = static_cast<int>(m % W);
  const int h = static_cast<int>((m / W) % H);
  const int b = static_cast<int>(m / (static_cast<long long>(W) * H));
  const float* xb = x + static_cast<long long>(b) * C * H * W;
  __align__(16) __nv_bfloat16 vals[KPAD];
#pragma unroll
  for (int k = 9 * C; k < KPAD; ++k) vals[k] = __float2bfloat16(0.f);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int ih = h + tap / 3 - 1, iw = w + tap % 3 - 1;
    const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
#pragma unroll
    for (int c = 0; c < C; ++c)
      vals[tap * C + c] = __float2bfloat16(ok ? xb[(static_cast<long long>(c) * H + ih) * W + iw] : 0.f);
  }
  uint4* dst = reinterpret_cast<uint4*>(out + m * KPAD);
#pragma unroll
  for (int i = 0; i < KPAD / 8; ++i) dst[i] = reinterpret_cast<const uint4*>(vals)[i];
}

__global__ void im2col3x3_nchw_kernel(const float* __restrict__ x, int B, int C, int H, int W, int kpad,
                                      __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int kv = kpad >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(B) * H * W * kv;
  if (idx >= total) return;
  const int v = static_cast<int>(idx % kv);
  long long m = idx / kv;
  const int w = static_cast<int>(m % W);
  m /= W;
  const int h = static_cast<int>(m % H);
  const int b = static_cast<int>(m / H);
  __nv_bfloat16 vals[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = v * 8 + j;
    float f = 0.f;
    if (k < 9 * C) {
      const int tap = k / C, c = k - tap * C;
      const int ih = h + tap / 3 - 1, iw = w + tap % 3 - 1;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) f = x[((static_cast<long long>(b) * C + c) * H + ih) * W + iw];
    }
    vals[j] = __float2bfloat16(f);
  }
  *reinterpret_cast<uint4*>(out + idx * 8) = *reinterpret_cast<const uint4*>(vals);
}

// conv weight [Cout][C][3][3] fp32 -> [Cout][kpad] bf16 with k = tap * C + c
__global__ void pack_first_conv_kernel(const float* __restrict__ w, int cout, int C, int kpad,
                                       __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cout * kpad) return;
  const int k = idx % kpad, o = idx / kpad;
  float f = 0.f;
  if (k < 9 * C) {
    const int tap = k / C, c = k - tap * C;
    f = w[(static_cast<long long>(o) * C + c) * 9 + tap];
  }
  out[idx] = __float2bfloat16(f);
}

// out[b,c,h,w] = h_nhwc[b,h,w,c] / sigma[label[b]]  in float64 (ncsnpp.py:259-261)
template <typename TO>
__global__ void scale_by_sigma_kernel(const float* __restrict__ h, const long long* __restrict__ labels,
                                      const double* __restrict__ sigmas, int B, int HW, int C, int do_scale,
                                      TO* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(B) * HW * C) return;
  const int b = static_cast<int>(idx / (static_cast<long long>(HW) * C));
  double v = static_cast<double>(h[idx]);  // h is NCHW, like out
  if (do_scale) v = v / sigmas[labels[b]];
  out[idx] = static_cast<TO>(v);
}

// rows 1 .. B-1 of a [B][n] fp32 matrix take row 0
__global__ void broadcast_row_kernel(float* __restrict__ buf, int n, long long total) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx < total) buf[n + idx] = buf[idx % n];
}

}  // namespace

void broadcast_row_f32(float* buf, int n, int B, cudaStream_t st) {
  if (B <= 1) return;
  const long long total = static_cast<long long>(B - 1) * n;
  broadcast_row_kernel<<<static_cast<unsigned>(cdiv64(total, 256)), 256, 0, st>>>(buf, n, total);
  T2P_LAUNCH_CHECK();
}

void temb_mlp(const long long* labels, const float* timesteps, int B, int nf, const float* w0, const float* b0, const float* w1,
              const float* b1, float* out, cudaStream_t st) {
  const size_t smem = sizeof(float) * (nf + 4 * nf);
  // -(ln 10000 / (half - 1)) evaluated in double then rounded once, as Python does (layers.py:101-103)
  const float neg_coef = static_cast<float>(-(log(10000.0) / static_cast<double>(nf / 2 - 1)));
  // a single sample (uniform labels) is latency-bound: more, thinner slices of the second layer
  temb_mlp_kernel<<<dim3(B, B == 1 ? 32 : 8), 256, smem, st>>>(labels, timesteps, nf, neg_coef, w0, b0, w1, b1, out);
  T2P_LAUNCH_CHECK();
}

void pack_conv_weight(const float* w, int cout, int cin, int k, int cin_pad, int out_dtype, void* out,
                      cudaStream_t st, long long ld) {
  const long long total = static_cast<long long>(cout) * k * k * cin_pad;
  if (ld <= 0) ld = static_cast<long long>(k) * k * cin_pad;
  const unsigned blocks = static_cast<unsigned>(cdiv64(total, 256));
  if (out_dtype == kF32) pack_conv_kernel<float><<<blocks, 256, 0, st>>>(w, cout, cin, k * k, cin_pad, ld, static_cast<float*>(out));
  else pack_conv_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(w, cout, cin, k * k, cin_pad, ld, static_cast<__nv_bfloat16*>(out));
  T2P_LAUNCH_CHECK();
}

void pack_matrix(const float* w, int rows, int cols, int transpose, int out_dtype, void* out, cudaStream_t st,
                 long long ld) {
  if (ld <= 0) ld = cols;
  const unsigned blocks = static_cast<unsigned>(cdiv64(static_cast<long long>(rows) * cols, 256));
  if (out_dtype == kF32) pack_matrix_kernel<float><<<blocks, 256, 0, st>>>(w, rows, cols, transpose, ld, static_cast<float*>(out));
  else pack_matrix_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(w, rows, cols, transpose, ld, static_cast<__nv_bfloat16*>(out));
  T2P_LAUNCH_CHECK();
}

void pack_identity(int n, long long ld, int out_dtype, void* out, cudaStream_t st) {
  const unsigned blocks = static_cast<unsigned>(cdiv64(static_cast<long long>(n) * n, 256));
  if (out_dtype == kF32) pack_identity_kernel<float><<<blocks, 256, 0, st>>>(n, ld, static_cast<float*>(out));
  else pack_identity_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(n, ld, static_cast<__nv_bfloat16*>(out));
  T2P_LAUNCH_CHECK();
}

void add_vectors_f32(const float* a, const float* b, int n, float* out, cudaStream_t st) {
  add_vectors_kernel<<<cdiv(n, 256), 256, 0, st>>>(a, b, n, out);
  T2P_LAUNCH_CHECK();
}

void convert_f32(const float* in, long long n, int out_dtype, void* out, cudaStream_t st) {
  const unsigned blocks = static_cast<unsigned>(cdiv64(n, 256));
  if (out_dtype == kF32) convert_kernel<float><<<blocks, 256, 0, st>>>(in, n, static_cast<float*>(out));
  else convert_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(in, n, static_cast<__nv_bfloat16*>(out));
  T2P_LAUNCH_CHECK();
}

void nhwc_to_nchw_f32(const void* in, int dtype, int B, int HW, int C, float* out, cudaStream_t st) {
  const unsigned blocks = static_cast<unsigned>(cdiv64(static_cast<long long>(B) * HW * C, 256));
  if (dtype == kF32) nhwc_to_nchw_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(in), B, HW, C, out);
  else nhwc_to_nchw_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), B, HW, C, out);
  T2P_LAUNCH_CHECK();
}

void nchw_f32_to_nhwc(const float* in, int B, int HW, int C, int cpad, int out_dtype, void* out, cudaStream_t st) {
  const unsigned blocks = static_cast<unsigned>(cdiv64(static_cast<long long>(B) * HW * cpad, 256));
  if (out_dtype == kF32) nchw_to_nhwc_kernel<float><<<blocks, 256, 0, st>>>(in, B, HW, C, cpad, static_cast<float*>(out));
  else nchw_to_nhwc_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(in, B, HW, C, cpad, static_cast<__nv_bfloat16*>(out));
  T2P_LAUNCH_CHECK();
}

void im2col3x3_nchw(const float* x, int B, int C, int H, int W, int kpad, void* out, cudaStream_t st) {
  T2P_CHECK(kpad % 8 == 0 && kpad >= 9 * C, "bad im2col padding");
  const long long pixels = static_cast<long long>(B) * H * W;
  if (C == 5 && kpad == 64) {
    launch_pdl(im2col3x3_pixel_kernel<5, 64>, dim3(static_cast<unsigned>(cdiv64(pixels, 256))), dim3(256), 0, st, x, B, H, W,
               static_cast<__nv_bfloat16*>(out));
    return;
  }
  if (C == 8 && kpad == 128) {
    launch_pdl(im2col3x3_pixel_kernel<8, 128>, dim3(static_cast<unsigned>(cdiv64(pixels, 256))), dim3(256), 0, st, x, B, H, W,
               static_cast<__nv_bfloat16*>(out));
    return;
  }
  const long long total = static_cast<long long>(B) * H * W * (kpad / 8);
  launch_pdl(im2col3x3_nchw_kernel, dim3(static_cast<unsigned>(cdiv64(total, 256))), dim3(256), 0, st, x, B, C, H, W, kpad,
             static_cast<__nv_bfloat16*>(out));
}

void pack_first_conv(const float* w, int cout, int C, int kpad, void* out, cudaStream_t st) {
  pack_first_conv_kernel<<<cdiv(cout * kpad, 256), 256, 0, st>>>(w, cout, C, kpad, static_cast<__nv_bfloat16*>(out));
  T2P_LAUNCH_CHECK();
}

void scale_by_sigma(const float* h_nchw, const long long* labels, const double* sigmas, int B, int HW, int C,
                    int do_scale, int out_dtype, void* out, cudaStream_t st) {
  const unsigned blocks = static_cast<unsigned>(cdiv64(static_cast<long long>(B) * HW * C, 256));
  if (out_dtype == kF64)
    scale_by_sigma_kernel<double><<<blocks, 256, 0, st>>>(h_nchw, labels, sigmas, B, HW, C, do_scale, static_cast<double*>(out));
  else if (out_dtype == kF32)
    scale_by_sigma_kernel<float><<<blocks, 256, 0, st>>>(h_nchw, labels, sigmas, B, HW, C, do_scale, static_cast<float*>(out));
  else T2P_CHECK(false, "score output must be fp64 or fp32");
  T2P_LAUNCH_CHECK();
}

}  // namespace t2p
