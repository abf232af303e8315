// Callers either side of the sampling loop, as device kernels working from small integer inputs:
//  * condition builders -- length masks, inpainting masks and the sampler's conditional mask generated on the
//    device from per-sample (length, residue ranges) integers instead of being built on the host and copied as
//    [B, N, N] / [B, C, N, N] bool tensors (reference utils.py:62-81,83-106,139-148; sampling.py:258-281);
//  * embed_tokens gather of the text-encoder embedding table (reference sampling_6d.py:134-137);
//  * the post-processing that turns a sampled 6D map into restraints (reference sampling_rosetta.py:69-96).
// All of it is integer / byte / single-rounding float work: results are bit-exact against the oracle.
#include "kernels.h"

namespace t2p {
namespace {

// ranges: [B or 1][R][2] inclusive residue index ranges (start, end); a position is selected if it lies in any
__device__ __forceinline__ bool in_ranges(const int* __restrict__ r, int R, int i) {
  bool hit = false;
  for (int k = 0; k < R; ++k) hit |= (i >= r[2 * k] && i <= r[2 * k + 1]);
  return hit;
}

// out[b][i][j] = i < len[b] && j < len[b]                                  (utils.py:89-93,139-148)
__global__ void length_mask_kernel(const int* __restrict__ lengths, int B, int N, unsigned char* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(B) * N * N) return;
  const int j = static_cast<int>(idx % N);
  const int i = static_cast<int>((idx / N) % N);
  const int b = static_cast<int>(idx / (static_cast<long long>(N) * N));
  const int l = lengths[b];
  out[idx] = (i < l && j < l) ? 1 : 0;
}

// out[b][i][j] = sel(i) || sel(j), sel = union of inclusive ranges          (utils.py:62-81)
__global__ void inpaint_mask_kernel(const int* __restrict__ ranges, int R, int per_sample, int B, int N,
                                    unsigned char* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(B) * N * N) return;
  const int j = static_cast<int>(idx % N);
  const int i = static_cast<int>((idx / N) % N);
  const int b = static_cast<int>(idx / (static_cast<long long>(N) * N));
  const int* r = ranges + (per_sample ? static_cast<long long>(b) * R * 2 : 0);
  out[idx] = (in_ranges(r, R, i) || in_ranges(r, R, j)) ? 1 : 0;
}

// The sampler's conditional mask (1 = free to evolve), sampling.py:258-281 for condition keys length / ss /
// inpainting built from integers: channel C-1 (padding) and 4..6 (ss) are fixed, length crops to the top-left
// len x len block, inpainting frees rows / columns of the selected residues only.
__global__ void condition_mask_kernel(const int* __restrict__ lengths, const int* __restrict__ ranges, int R,
                                      int per_sample, int has_ss, int B, int C, int N,
                                      unsigned char* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(B) * C * N * N) return;
  const int j = static_cast<int>(idx % N);
  const int i = static_cast<int>((idx / N) % N);
  const int c = static_cast<int>((idx / (static_cast<long long>(N) * N)) % C);
  const int b = static_cast<int>(idx / (static_cast<long long>(N) * N * C));
  bool free_ = true;
  if (lengths) {
    const int l = lengths[b];
    free_ = free_ && (i < l && j < l) && (c != C - 1);
  }
  if (has_ss) free_ = free_ && !(c >= 4 && c < 7);
  if (ranges) {
    const int* r = ranges + (per_sample ? static_cast<long long>(b) * R * 2 : 0);
    free_ = free_ && (in_ranges(r, R, i) || in_ranges(r, R, j));
  }
  out[idx] = free_ ? 1 : 0;
}

template <typename T>
__global__ void embed_gather_kernel(const T* __restrict__ table, long long V, int D, const long long* __restrict__ tok,
                                    long long n, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int vec = D >> 2;
  if (idx >= n * vec) return;
  const long long t = idx / vec;
  const int d4 = static_cast<int>(idx - t * vec) << 2;
  long long id = tok[t];
  id = id < 0 ? 0 : (id >= V ? V - 1 : id);
  float v[4];
  if constexpr (sizeof(T) == 4) {
    const float4 q = *reinterpret_cast<const float4*>(table + id * D + d4);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
    const uint2 q = *reinterpret_cast<const uint2*>(table + id * D + d4);
    v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
    v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
  }
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + t * D + d4) = make_float4(v[0], v[1], v[2], v[3]);
  if (out_bf16) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(out_bf16 + t * D + d4) = o;
  }
}

// One block per sample (sampling_rosetta.py:69-96): msk = round(x[C-1]) == 1 (round half to even, as numpy);
// L = sqrt(count) must be an integer; the masked entries of channels 0..3, in raster order, clipped to [-1, 1],
// are compacted into out[b][ch][0 .. L*L) and their inverse scalings into out[b][4 + ch][..]:
//   dist_abs = (dist + 1) * 10, omega_abs = omega * pi, theta_abs = theta * pi, phi_abs = (phi + 1) * pi / 2
// with fp32 arithmetic and float32(pi), exactly as numpy evaluates them on float32 arrays.
__global__ void __launch_bounds__(1024) postprocess_kernel(const float* __restrict__ x, int C, int N,
                                                          float* __restrict__ out, int* __restrict__ L_out) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int b = blockIdx.x;
  const int NN = N * N;
  const float* xb = x + static_cast<long long>(b) * C * NN;
  float* ob = out + static_cast<long long>(b) * 8 * NN;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const float pi = 3.14159274101257324f;  // float32(math.pi)
  for (int base = 0; base < NN; base += blockDim.x) {
    const int e = base + threadIdx.x;
    bool m = false;
    if (e < NN) m = (rintf(xb[static_cast<long long>(C - 1) * NN + e]) == 1.f);
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    const int within = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int before = carry;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    if (m) {
      const int pos = before + within;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float v = xb[static_cast<long long>(ch) * NN + e];
        v = fminf(fmaxf(v, -1.f), 1.f);
        ob[static_cast<long long>(ch) * NN + pos] = v;
        float a;
        if (ch == 0) a = (v + 1.f) * 10.f;
        else if (ch == 3) a = ((v + 1.f) * pi) / 2.f;
        else a = v * pi;
        ob[static_cast<long long>(4 + ch) * NN + pos] = a;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < nw; ++w) t += warp_tot[w];
      carry += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int cnt = carry;
    int l = static_cast<int>(sqrtf(static_cast<float>(cnt)) + 0.5f);
    while (l * l > cnt) --l;
    while ((l + 1) * (l + 1) <= cnt) ++l;
    L_out[b] = (l * l == cnt) ? l : -1;  // -1: "improper masking channel" (sampling_rosetta.py:71-73)
  }
}

}  // namespace

void length_mask(const int* lengths, int B, int N, unsigned char* out, cudaStream_t st) {
  const long long n = static_cast<long long>(B) * N * N;
  length_mask_kernel<<<static_cast<unsigned>(cdiv64(n, 256)), 256, 0, st>>>(lengths, B, N, out);
  T2P_LAUNCH_CHECK();
}

void inpaint_mask(const int* ranges, int R, int per_sample, int B, int N, unsigned char* out, cudaStream_t st) {
  const long long n = static_cast<long long>(B) * N * N;
  inpaint_mask_kernel<<<static_cast<unsigned>(cdiv64(n, 256)), 256, 0, st>>>(ranges, R, per_sample, B, N, out);
  T2P_LAUNCH_CHECK();
}

void condition_mask(const int* lengths, const int* ranges, int R, int per_sample, int has_ss, int B, int C, int N,
                    unsigned char* out, cudaStream_t st) {
  const long long n = static_cast<long long>(B) * C * N * N;
  condition_mask_kernel<<<static_cast<unsigned>(cdiv64(n, 256)), 256, 0, st>>>(lengths, ranges, R, per_sample, has_ss, B,
                                                                              C, N, out);
  T2P_LAUNCH_CHECK();
}

void embed_gather(const void* table, int table_dtype, long long V, int D, const long long* tokens, long long n,
                  float* out_f32, void* out_bf16, cudaStream_t st) {
  T2P_CHECK(D % 4 == 0, "embedding width must be a multiple of 4");
  T2P_CHECK(table_dtype == kF32 || table_dtype == kBF16, "embedding table must be fp32 or bf16");
  const unsigned blocks = static_cast<unsigned>(cdiv64(n * (D / 4), 256));
  if (table_dtype == kF32)
    embed_gather_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(table), V, D, tokens, n, out_f32,
                                                       static_cast<__nv_bfloat16*>(out_bf16));
  else
    embed_gather_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(table), V, D, tokens, n,
                                                               out_f32, static_cast<__nv_bfloat16*>(out_bf16));
  T2P_LAUNCH_CHECK();
}

void postprocess_6d(const float* x, int B, int C, int N, float* out, int* L_out, cudaStream_t st) {
  T2P_CHECK(C >= 5, "a 6D map has at least 4 coordinate channels and the padding channel");
  T2P_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * 8 * static_cast<size_t>(B) * N * N, st));
  postprocess_kernel<<<B, 1024, 0, st>>>(x, C, N, out, L_out);
  T2P_LAUNCH_CHECK();
}

}  // namespace t2p
