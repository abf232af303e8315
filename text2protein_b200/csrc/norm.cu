// GroupNorm (+SiLU, +2x resampling, +skip concat), LayerNorm and GEGLU for NHWC / token-major tensors.
//
// All three are HBM-bound: one coalesced 16-byte-per-thread read and one write per element, with the
// statistics kept per (sample, channel) so that the producer GEMM's epilogue can accumulate them and the
// skip concat of the UNet up path (channels of two tensors, group boundaries straddling both) needs no copy.
// Replaces nn.GroupNorm + nn.SiLU + naive_up/downsample_2d + torch.cat (score_sde_pytorch/models/layers.py:
// 179-188,282-311; ncsnpp.py:250), nn.LayerNorm and GEGLU (model/attention.py:37-44,203-205).
#include <algorithm>

#include "kernels.h"

namespace t2p {
namespace {

template <typename T>
struct Vec8;
template <>
struct Vec8<float> {
  float v[8];
  struct Raw { float4 a, b; };
  __device__ static Raw load_raw(const float* p) {
    Raw r;
    r.a = reinterpret_cast<const float4*>(p)[0];
    r.b = reinterpret_cast<const float4*>(p)[1];
    return r;
  }
  __device__ static Vec8 unpack(const Raw& t) {
    Vec8 r;
    r.v[0] = t.a.x; r.v[1] = t.a.y; r.v[2] = t.a.z; r.v[3] = t.a.w;
    r.v[4] = t.b.x; r.v[5] = t.b.y; r.v[6] = t.b.z; r.v[7] = t.b.w;
    return r;
  }
  __device__ static Vec8 load(const float* p) {
    Vec8 r;
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  __device__ void store(float* p) const {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <>
struct Vec8<__nv_bfloat16> {
  float v[8];
  using Raw = uint4;
  __device__ static Raw load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ static Vec8 unpack(const Raw& t) {
    Vec8 r;
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r.v[2 * i] = __uint_as_float(w[i] << 16);
      r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    return r;
  }
  __device__ static Vec8 load(const __nv_bfloat16* p) {
    Vec8 r;
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r.v[2 * i] = __uint_as_float(w[i] << 16);
      r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    return r;
  }
  __device__ void store(__nv_bfloat16* p) const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }

// SiLU for the bf16 pipeline: x * sigmoid(x) = h + h * tanh(h), h = x / 2 -- one MUFU op per element; the
// approximation error (~5e-4 relative) is below the bf16 rounding of the stored result.
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
template <typename T>
__device__ __forceinline__ float silu_t(float x);
template <>
__device__ __forceinline__ float silu_t<float>(float x) { return silu_f(x); }
template <>
__device__ __forceinline__ float silu_t<__nv_bfloat16>(float x) { return silu_tanh(x); }


// ------------------------------------------------------------------ statistics: per (sample, channel) sums
// Deterministic two-level reduction (no atomics): each block reduces a contiguous run of pixels to
// part[b][blk][c][{sum, sumsq}] in a fixed order; gn_finalize adds the blocks in order, in double.
template <typename T>
__global__ void gn_stats_kernel(const T* __restrict__ a0, int c0, const T* __restrict__ a1, int c1, int HW,
                                int pix_per_block, float* __restrict__ part) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];  // [rows][2 * ctot]
  const int ctot = c0 + c1;
  const int slots = ctot >> 3;
  const int rows = blockDim.x / slots;
  const int slot = threadIdx.x % slots;
  const int row = threadIdx.x / slots;
  const int b = blockIdx.y;
  float s[8] = {}, q[8] = {};
  const int ch = slot << 3;
  {
    const T* src;
    int cs, co;
    if (ch < c0) { src = a0; cs = c0; co = ch; }
    else { src = a1; cs = c1; co = ch - c0; }
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(HW, p0 + pix_per_block);
    for (int p = p0 + row; p < p1; p += rows) {
      const Vec8<T> v = Vec8<T>::load(src + (static_cast<long long>(b) * HW + p) * cs + co);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += v.v[i]; q[i] += v.v[i] * v.v[i]; }
    }
  }
  float* mine = sm + static_cast<size_t>(row) * 2 * ctot;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mine[2 * (ch + i)] = s[i];
    mine[2 * (ch + i) + 1] = q[i];
  }
  __syncthreads();
  float* dst = part + (static_cast<long long>(b) * gridDim.x + blockIdx.x) * 2 * ctot;
  for (int i = threadIdx.x; i < 2 * ctot; i += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < rows; ++r) acc += sm[static_cast<size_t>(r) * 2 * ctot + i];
    dst[i] = acc;
  }
}

// per (sample, group): mean / rstd -> per (sample, channel) affine  y = x * scale + shift.
// One warp per (sample, group); lanes stride over (channel, block) pairs and combine in double in a fixed
// order, so the result does not depend on scheduling.
__global__ void gn_finalize_kernel(const float* __restrict__ part0, int nblk0, int c0,
                                   const float* __restrict__ part1, int nblk1, int c1,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, int G, int HW,
                                   float eps, float* __restrict__ scale, float* __restrict__ shift, int total) {
  pdl_trigger();
  pdl_wait();
  const int C = c0 + c1;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= total) return;
  const int g = wid % G, b = wid / G;
  const int cpg = C / G;
  double s = 0, q = 0;
  // Channels are taken four at a time so that every lane has four independent L2 loads in flight (a plain
  // channel-by-channel loop serialises one round trip per partial).  The order of the additions is fixed.
  auto locate = [&](int c, const float*& base, long long& pitch, int& nblk) {
    const bool first = c < c0;
    const int cs = first ? c0 : c1;
    nblk = first ? nblk0 : nblk1;
    base = (first ? part0 : part1) + (static_cast<long long>(b) * nblk * cs + (first ? c : c - c0)) * 2;
    pitch = static_cast<long long>(cs) * 2;
  };
  int ci = 0;
  {
    for (; ci + 4 <= cpg; ci += 4) {
      const int c = g * cpg + ci;
      if (c < c0 && c + 3 >= c0) break;  // the four channels straddle the two sources: finish one by one
      const float* base;
      long long pitch;
      int nblk;
      locate(c, base, pitch, nblk);
      for (int k = lane; k < nblk; k += 32) {
        const float* pk = base + k * pitch;
        const float2 v0 = __ldg(reinterpret_cast<const float2*>(pk));
        const float2 v1 = __ldg(reinterpret_cast<const float2*>(pk + 2));
        const float2 v2 = __ldg(reinterpret_cast<const float2*>(pk + 4));
        const float2 v3 = __ldg(reinterpret_cast<const float2*>(pk + 6));
        s += static_cast<double>(v0.x); q += static_cast<double>(v0.y);
        s += static_cast<double>(v1.x); q += static_cast<double>(v1.y);
        s += static_cast<double>(v2.x); q += static_cast<double>(v2.y);
        s += static_cast<double>(v3.x); q += static_cast<double>(v3.y);
      }
    }
  }
  for (; ci < cpg; ++ci) {
    const float* base;
    long long pitch;
    int nblk;
    locate(g * cpg + ci, base, pitch, nblk);
    for (int k = lane; k < nblk; k += 32) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(base + k * pitch));
      s += static_cast<double>(v.x);
      q += static_cast<double>(v.y);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const double n = static_cast<double>(cpg) * HW;
  const double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0) var = 0;
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  for (int ci = lane; ci < cpg; ci += 32) {
    const int c = g * cpg + ci;
    const float a = gamma[c] * rstd;
    scale[static_cast<long long>(b) * C + c] = a;
    shift[static_cast<long long>(b) * C + c] = beta[c] - static_cast<float>(mean) * a;
  }
}

// ------------------------------------------------------------------ apply
// MODE 0: same resolution; 1: 2x2 mean AFTER the activation (also emits the 2x2 mean of the raw input);
// 2: nearest x2 upsample of the activated tensor.
template <typename T, int MODE>
__global__ void gn_apply_kernel(const T* __restrict__ a0, int c0, const T* __restrict__ a1, int c1, int B, int H,
                                int W, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                T* __restrict__ out, T* __restrict__ raw_out) {
  pdl_trigger();
  pdl_wait();
  const int ctot = c0 + c1;
  const int slots = ctot >> 3;
  const int OH = (MODE == 1) ? H >> 1 : H, OW = (MODE == 1) ? W >> 1 : W;  // iteration space
  const long long total = static_cast<long long>(B) * OH * OW * slots;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int slot = static_cast<int>(idx % slots);
  long long pix = idx / slots;
  const int ow = static_cast<int>(pix % OW);
  pix /= OW;
  const int oh = static_cast<int>(pix % OH);
  const int b = static_cast<int>(pix / OH);
  const int ch = slot << 3;
  const T* src;
  int cs, co;
  if (ch < c0) { src = a0; cs = c0; co = ch; }
  else { src = a1; cs = c1; co = ch - c0; }
  float sc[8], sh[8];
  {
    const float4* ps = reinterpret_cast<const float4*>(scale + static_cast<long long>(b) * ctot + ch);
    const float4* ph = reinterpret_cast<const float4*>(shift + static_cast<long long>(b) * ctot + ch);
    const float4 s0 = ps[0], s1 = ps[1], h0 = ph[0], h1 = ph[1];
    sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
    sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
  }
  if (MODE == 1) {
    Vec8<T> acc, raw;
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc.v[i] = 0.f; raw.v[i] = 0.f; }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const long long ip = (static_cast<long long>(b) * H + (2 * oh + dy)) * W + (2 * ow + dx);
        const Vec8<T> v = Vec8<T>::load(src + ip * cs + co);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float y = fmaf(v.v[i], sc[i], sh[i]);
          if (act) y = silu_t<T>(y);
          acc.v[i] += y;
          raw.v[i] += v.v[i];
        }
      }
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc.v[i] *= 0.25f; raw.v[i] *= 0.25f; }
    const long long op = (static_cast<long long>(b) * OH + oh) * OW + ow;
    acc.store(out + op * ctot + ch);
    if (raw_out) raw.store(raw_out + op * ctot + ch);
  } else {
    const long long ip = (static_cast<long long>(b) * H + oh) * W + ow;
    Vec8<T> v = Vec8<T>::load(src + ip * cs + co);
    const Vec8<T> raw = v;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = fmaf(v.v[i], sc[i], sh[i]);
      v.v[i] = act ? silu_t<T>(y) : y;
    }
    if (MODE == 0) {
      v.store(out + ip * ctot + ch);
    } else {
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const long long op = (static_cast<long long>(b) * (2 * H) + (2 * oh + dy)) * (2 * W) + (2 * ow + dx);
          v.store(out + op * ctot + ch);
          if (raw_out) raw.store(raw_out + op * ctot + ch);  // nearest x2 of the raw input (folded skip path)
        }
    }
  }
}

// MODE 0 streaming path.  A block owns `ppb` consecutive pixels of one sample; thread = (pixel row, 8-channel
// slot), so its 16 scale/shift values live in registers for the whole run and the block's accesses are one
// contiguous stream per source.  UNROLL independent 16-byte loads are in flight per thread.
template <typename T, int UNROLL>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 4 : 2) gn_apply_rows_kernel(const T* __restrict__ a0, int c0, const T* __restrict__ a1,
                                                            int c1, int HW, int ppb, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, int act,
                                                            T* __restrict__ out, int reverse) {
  pdl_trigger();
  pdl_wait();
  const int ctot = c0 + c1;
  const int slots = ctot >> 3;
  const int rows = blockDim.x / slots;
  const int slot = threadIdx.x % slots;
  const int row = threadIdx.x / slots;
  // blocks are dispatched in (x fastest, then y) order: reversed, the kernel sweeps memory from the end
  const int b = reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int bx = reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int ch = slot << 3;
  const T* src;
  int cs, co;
  if (ch < c0) { src = a0; cs = c0; co = ch; }
  else { src = a1; cs = c1; co = ch - c0; }
  float sc[8], sh[8];
  {
    const float4* ps = reinterpret_cast<const float4*>(scale + static_cast<long long>(b) * ctot + ch);
    const float4* ph = reinterpret_cast<const float4*>(shift + static_cast<long long>(b) * ctot + ch);
    const float4 s0 = __ldg(ps), s1 = __ldg(ps + 1), h0 = __ldg(ph), h1 = __ldg(ph + 1);
    sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
    sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
  }
  const int p0 = bx * ppb;
  const int p1 = min(HW, p0 + ppb);
  // pointer-bumped streams: sp / op advance by one round (rows * UNROLL pixels) per iteration
  const T* sp = src + (static_cast<long long>(b) * HW + p0 + row) * cs + co;
  T* op = out + (static_cast<long long>(b) * HW + p0 + row) * ctot + ch;
  const long long sstep = static_cast<long long>(rows) * cs, ostep = static_cast<long long>(rows) * ctot;
  for (int p = p0 + row; p < p1; p += rows * UNROLL, sp += UNROLL * sstep, op += UNROLL * ostep) {
    // raw (still packed) vectors first: UNROLL independent loads in flight at 4 registers each for bf16
    typename Vec8<T>::Raw raw[UNROLL];
    const bool full = p + (UNROLL - 1) * rows < p1;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (full || p + u * rows < p1) raw[u] = Vec8<T>::load_raw(sp + u * sstep);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (full || p + u * rows < p1) {
        Vec8<T> v = Vec8<T>::unpack(raw[u]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float y = fmaf(v.v[i], sc[i], sh[i]);
          v.v[i] = act ? silu_t<T>(y) : y;
        }
        v.store(op + u * ostep);
      }
    }
  }
}

// ------------------------------------------------------------------ one-launch GroupNorm for small tensors
// At 32 x 32 and below the statistics / finalize / apply chain is three launches of a few microseconds each for a
// tensor that fits in the registers of one thread block per (sample, 32-channel slice): 79 of the 403 launches of a
// forward at cond_length.yml.  Here a block loads its [HW pixels x 32 channels] slice ONCE (16 bytes per thread and
// pass, all passes in flight), reduces sum / sum of squares per channel (warp shuffles over the pixels of a warp,
// shared memory over the warps, fixed order -> deterministic), folds the channels of a group, and writes
// act((x - mean) * rstd * gamma + beta) from the registers.  Groups of 4, 8, 16 or 32 channels (every GroupNorm of
// the network); two-source channel concat as in the other kernels.
constexpr int kGnsCS = 32;       // channels per block
constexpr int kGnsMaxPass = 16;  // 256 threads x 16 passes x 8 channels = 1024 pixels x 32 channels

template <typename T>
__global__ void __launch_bounds__(256) gn_small_kernel(const T* __restrict__ a0, int c0, const T* __restrict__ a1, int c1,
                                                       int HW, int cpg, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps, int act,
                                                       T* __restrict__ out) {
  __shared__ float red[8][4][16];  // [warp][slot][8 sums | 8 sums of squares]
  __shared__ float stat[kGnsCS][2];  // per channel of the slice: scale, shift
  pdl_trigger();  // (a GEMM launched behind this kernel with programmatic serialisation may start its prologue)
  const int ctot = c0 + c1;
  const int b = blockIdx.y;
  const int ch0 = blockIdx.x * kGnsCS;       // first channel of the slice (in the concat)
  const int slot = threadIdx.x & 3;          // 8-channel slot within the slice
  const int prow = threadIdx.x >> 2;         // pixel within a pass of 64
  const T* src;
  int cs, co;
  if (ch0 < c0) { src = a0; cs = c0; co = ch0; }
  else { src = a1; cs = c1; co = ch0 - c0; }
  const T* sp = src + (static_cast<long long>(b) * HW + prow) * cs + co + slot * 8;
  typename Vec8<T>::Raw raw[kGnsMaxPass];
#pragma unroll
  for (int i = 0; i < kGnsMaxPass; ++i)
    if (prow + 64 * i < HW) raw[i] = Vec8<T>::load_raw(sp + static_cast<long long>(64 * i) * cs);
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s[k] = 0.f; q[k] = 0.f; }
#pragma unroll
  for (int i = 0; i < kGnsMaxPass; ++i) {
    if (prow + 64 * i < HW) {
      const Vec8<T> v = Vec8<T>::unpack(raw[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] += v.v[k]; q[k] = fmaf(v.v[k], v.v[k], q[k]); }
    }
  }
  // lanes with the same slot: xor 4, 8, 16
#pragma unroll
  for (int o = 4; o <= 16; o <<= 1)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
      q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
    }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 4) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { red[warp][lane][k] = s[k]; red[warp][lane][8 + k] = q[k]; }
  }
  __syncthreads();
  if (threadIdx.x < kGnsCS) {
    // lane c: channel c of the slice.  Warp partials in fp32 (each is a pairwise sum of <= 128 values), the group
    // combination -- E[x^2] - E[x]^2 cancels -- in double, by xor-shuffles over the cpg lanes of the group.
    const int c = threadIdx.x;
    float cs_ = 0.f, cq_ = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      cs_ += red[w][c >> 3][c & 7];
      cq_ += red[w][c >> 3][8 + (c & 7)];
    }
    double gs = static_cast<double>(cs_), gq = static_cast<double>(cq_);
    for (int o = 1; o < cpg; o <<= 1) {
      gs += __shfl_xor_sync(0xffffffffu, gs, o);
      gq += __shfl_xor_sync(0xffffffffu, gq, o);
    }
    const double n = static_cast<double>(HW) * cpg;
    const double mean = gs / n;
    const float var = static_cast<float>(fmax(gq / n - mean * mean, 0.0));
    const float rstd = rsqrtf(var + eps);
    const float ga = gamma[ch0 + c], be = beta[ch0 + c];
    stat[c][0] = rstd * ga;
    stat[c][1] = be - static_cast<float>(mean) * rstd * ga;
  }
  __syncthreads();
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = stat[slot * 8 + k][0]; sh[k] = stat[slot * 8 + k][1]; }
  T* op = out + (static_cast<long long>(b) * HW + prow) * ctot + ch0 + slot * 8;
#pragma unroll
  for (int i = 0; i < kGnsMaxPass; ++i) {
    if (prow + 64 * i < HW) {
      Vec8<T> v = Vec8<T>::unpack(raw[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float y = fmaf(v.v[k], sc[k], sh[k]);
        v.v[k] = act ? silu_t<T>(y) : y;
      }
      v.store(op + static_cast<long long>(64 * i) * ctot);
    }
  }
}

// ------------------------------------------------------------------ LayerNorm over the last dim
// One warp normalises ROWS rows at a time (their loads are issued together, the shuffle reductions interleave), the
// affine parameters are read once per warp as 16-byte vectors.  C = 256 * nv, nv <= MAXV.
template <typename T, int MAXV, int ROWS>
__global__ void layernorm_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int M, int C, float eps, T* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int row0 = warp * ROWS;
  if (row0 >= M) return;
  const int nv = C >> 8;  // 8-element vectors per lane
  float v[ROWS][MAXV][8];
  float s[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    s[r] = 0.f;
    const bool ok = row0 + r < M;
#pragma unroll
    for (int j = 0; j < MAXV; ++j)
      if (j < nv) {
        Vec8<T> t;
        if (ok) t = Vec8<T>::load(x + static_cast<long long>(row0 + r) * C + (j * 32 + lane) * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[r][j][i] = ok ? t.v[i] : 0.f; s[r] += v[r][j][i]; }
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < ROWS; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
  float q[ROWS], mean[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    mean[r] = s[r] / C;
    q[r] = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j)
      if (j < nv) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[r][j][i] - mean[r]; q[r] += d * d; }
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < ROWS; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
#pragma unroll
  for (int j = 0; j < MAXV; ++j)
    if (j < nv) {
      const int c = (j * 32 + lane) * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (row0 + r >= M) break;
        const float rstd = rsqrtf(q[r] / C + eps);
        Vec8<T> t;
#pragma unroll
        for (int i = 0; i < 8; ++i) t.v[i] = (v[r][j][i] - mean[r]) * rstd * gm[i] + bt[i];
        t.store(y + static_cast<long long>(row0 + r) * C + c);
      }
    }
}

template <typename T>
__device__ __forceinline__ float ld1(const T* p);
template <>
__device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// generic (any C): one warp per row, two passes over global memory
template <typename T>
__global__ void layernorm_generic_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, int M, int C, float eps,
                                         T* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const T* r = x + static_cast<long long>(warp) * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += ld1(r + c);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = ld1(r + c) - mean; q += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  for (int c = lane; c < C; c += 32)
    st1(y + static_cast<long long>(warp) * C + c, (ld1(r + c) - mean) * rstd * gamma[c] + beta[c]);
}

// ------------------------------------------------------------------ GEGLU: out = a * gelu_erf(gate)
// erf for the bf16 engine: Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7 absolute (below the fp32 rounding of
// 1 + erf, far below the bf16 rounding of the result) in ~14 instructions with two MUFU ops; erff() costs about
// twice that and made this kernel issue-bound (16.8 M outputs per launch).  The fp32 engine keeps erff().
__device__ __forceinline__ float erf_as(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
  float pl = fmaf(1.061405429f, t, -1.453152027f);
  pl = fmaf(pl, t, 1.421413741f);
  pl = fmaf(pl, t, -0.284496736f);
  pl = fmaf(pl, t, 0.254829592f);
  const float r = fmaf(-pl * t, __expf(-ax * ax), 1.f);
  return copysignf(r, x);
}
template <typename T>
__device__ __forceinline__ float gelu_erf(float g) {
  if (sizeof(T) == 2) return 0.5f * g * (1.f + erf_as(g * 0.70710678118654752f));
  return 0.5f * g * (1.f + erff(g * 0.70710678118654752f));
}

template <typename T>
__global__ void geglu_kernel(const T* __restrict__ z, long long M, int D, T* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int slots = D >> 3;
  if (idx >= M * slots) return;
  const int slot = static_cast<int>(idx % slots);
  const long long m = idx / slots;
  const Vec8<T> a = Vec8<T>::load(z + m * 2 * D + slot * 8);
  const Vec8<T> g = Vec8<T>::load(z + m * 2 * D + D + slot * 8);
  Vec8<T> o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = a.v[i] * gelu_erf<T>(g.v[i]);
  o.store(out + m * D + slot * 8);
}

}  // namespace

// ===================================================================================== host launchers
int gn_stats_blocks(int B, int HW) {
  // enough blocks to fill the machine, each with a contiguous run of pixels
  int blocks = std::max(1, std::min(cdiv(HW, 64), cdiv(148 * 8, B)));
  const int ppb = cdiv(HW, blocks);
  return cdiv(HW, ppb);
}

void gn_stats(const void* a0, int c0, const void* a1, int c1, int B, int HW, int dtype, float* part,
              cudaStream_t st) {
  const int ctot = c0 + c1;
  T2P_CHECK(c0 % 8 == 0 && c1 % 8 == 0 && ctot > 0, "GroupNorm channels must be multiples of 8");
  const int slots = ctot / 8;
  T2P_CHECK(slots <= 256, "too many channels for gn_stats");
  const int threads = slots * (256 / slots);
  const int blocks = gn_stats_blocks(B, HW);
  const int ppb = cdiv(HW, blocks);
  dim3 grid(blocks, B);
  const size_t smem = sizeof(float) * 2 * ctot * (threads / slots);
  if (dtype == kF32) {
    static bool cfgd = false;
    if (!cfgd) { T2P_CUDA(cudaFuncSetAttribute(gn_stats_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); cfgd = true; }
    launch_pdl(gn_stats_kernel<float>, grid, dim3(threads), smem, st, static_cast<const float*>(a0), c0,
               static_cast<const float*>(a1), c1, HW, ppb, part);
  } else {
    static bool cfgd = false;
    if (!cfgd) { T2P_CUDA(cudaFuncSetAttribute(gn_stats_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); cfgd = true; }
    launch_pdl(gn_stats_kernel<__nv_bfloat16>, grid, dim3(threads), smem, st, static_cast<const __nv_bfloat16*>(a0), c0,
               static_cast<const __nv_bfloat16*>(a1), c1, HW, ppb, part);
  }
  T2P_LAUNCH_CHECK();
}

void gn_finalize(const float* part0, int nblk0, int c0, const float* part1, int nblk1, int c1, const float* gamma,
                 const float* beta, int B, int G, int HW, float eps, float* scale, float* shift, cudaStream_t st) {
  T2P_CHECK((c0 + c1) % G == 0, "channels not divisible by groups");
  const int total = B * G;
  launch_pdl<2>(gn_finalize_kernel, dim3(cdiv(total, 4)), dim3(128), 0, st, part0, nblk0, c0, part1, nblk1, c1, gamma, beta,
             G, HW, eps, scale, shift, total);
}

template <typename T>
static void gn_apply_t(const void* a0, int c0, const void* a1, int c1, int B, int H, int W, const float* scale,
                       const float* shift, int act, int mode, void* out, void* raw_out, cudaStream_t st, int reverse) {
  const int slots = (c0 + c1) / 8;
  const int OH = mode == 1 ? H / 2 : H, OW = mode == 1 ? W / 2 : W;
  const long long total = static_cast<long long>(B) * OH * OW * slots;
  const unsigned blocks = static_cast<unsigned>(cdiv64(total, 256));
  const T* p0 = static_cast<const T*>(a0);
  const T* p1 = static_cast<const T*>(a1);
  T* o = static_cast<T*>(out);
  T* r = static_cast<T*>(raw_out);
  if (mode == 0 && slots <= 256) {
    constexpr int UNROLL = 4;
    const int HW = H * W;
    const int rows = 256 / slots;
    const int threads = rows * slots;
    // enough blocks to fill the machine several times over, up to 4 rounds of UNROLL vectors per thread
    const int want_blocks = std::max(1, (148 * 8) / B);
    int iters = HW / (rows * UNROLL * want_blocks);
    iters = std::max(1, std::min(iters, 4));
    const int ppb = rows * UNROLL * iters;
    dim3 grid(cdiv(HW, ppb), B);
    launch_pdl<4>(gn_apply_rows_kernel<T, UNROLL>, grid, dim3(threads), 0, st, p0, c0, p1, c1, HW, ppb, scale, shift, act, o, reverse);
  } else if (mode == 0) launch_pdl<4>(gn_apply_kernel<T, 0>, dim3(blocks), dim3(256), 0, st, p0, c0, p1, c1, B, H, W, scale, shift, act, o, r);
  else if (mode == 1) launch_pdl<4>(gn_apply_kernel<T, 1>, dim3(blocks), dim3(256), 0, st, p0, c0, p1, c1, B, H, W, scale, shift, act, o, r);
  else launch_pdl<4>(gn_apply_kernel<T, 2>, dim3(blocks), dim3(256), 0, st, p0, c0, p1, c1, B, H, W, scale, shift, act, o, r);
}

void gn_apply(const void* a0, int c0, const void* a1, int c1, int B, int H, int W, int dtype, const float* scale,
              const float* shift, int act, int mode, void* out, void* raw_out, cudaStream_t st, int reverse) {
  T2P_CHECK(mode >= 0 && mode <= 2, "bad resample mode");
  if (mode == 1) T2P_CHECK(H % 2 == 0 && W % 2 == 0, "downsample needs even H, W");
  if (dtype == kF32) gn_apply_t<float>(a0, c0, a1, c1, B, H, W, scale, shift, act, mode, out, raw_out, st, reverse);
  else gn_apply_t<__nv_bfloat16>(a0, c0, a1, c1, B, H, W, scale, shift, act, mode, out, raw_out, st, reverse);
}

bool gn_small_supported(int c0, int c1, int HW, int G) {
  const int C = c0 + c1;
  if (G <= 0 || C % G) return false;
  const int cpg = C / G;
  return HW > 0 && HW <= 64 * kGnsMaxPass && C % kGnsCS == 0 && c0 % kGnsCS == 0 &&
         (cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32);
}

void gn_small(const void* a0, int c0, const void* a1, int c1, int B, int HW, int dtype, int G, float eps,
              const float* gamma, const float* beta, int act, void* out, cudaStream_t st) {
  T2P_CHECK(gn_small_supported(c0, c1, HW, G), "shape not supported by the one-launch GroupNorm");
  const int C = c0 + c1;
  dim3 grid(C / kGnsCS, B);
  if (dtype == kF32)
    gn_small_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(a0), c0, static_cast<const float*>(a1), c1, HW,
                                                 C / G, gamma, beta, eps, act, static_cast<float*>(out));
  else
    gn_small_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(a0), c0,
                                                         static_cast<const __nv_bfloat16*>(a1), c1, HW, C / G, gamma, beta,
                                                         eps, act, static_cast<__nv_bfloat16*>(out));
  T2P_LAUNCH_CHECK();
}

void layernorm(const void* x, const float* gamma, const float* beta, long long M, int C, float eps, int dtype,
               void* y, cudaStream_t st) {
  const bool fast = (C % 256 == 0) && C <= 1024;
  const bool narrow = fast && C <= 512;  // 4 rows per warp while the row cache stays in registers
  const int rows_per_block = narrow ? 8 * 4 : 8;
  const unsigned blocks = static_cast<unsigned>(cdiv64(M, rows_per_block));
  if (dtype == kF32) {
    if (narrow) launch_pdl(layernorm_kernel<float, 2, 4>, dim3(blocks), dim3(256), 0, st, static_cast<const float*>(x), gamma, beta, (int)M, C, eps, static_cast<float*>(y));
    else if (fast) launch_pdl(layernorm_kernel<float, 4, 1>, dim3(blocks), dim3(256), 0, st, static_cast<const float*>(x), gamma, beta, (int)M, C, eps, static_cast<float*>(y));
    else launch_pdl(layernorm_generic_kernel<float>, dim3(blocks), dim3(256), 0, st, static_cast<const float*>(x), gamma, beta, (int)M, C, eps, static_cast<float*>(y));
  } else {
    if (narrow) launch_pdl(layernorm_kernel<__nv_bfloat16, 2, 4>, dim3(blocks), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(x), gamma, beta, (int)M, C, eps, static_cast<__nv_bfloat16*>(y));
    else if (fast) launch_pdl(layernorm_kernel<__nv_bfloat16, 4, 1>, dim3(blocks), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(x), gamma, beta, (int)M, C, eps, static_cast<__nv_bfloat16*>(y));
    else launch_pdl(layernorm_generic_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(x), gamma, beta, (int)M, C, eps, static_cast<__nv_bfloat16*>(y));
  }
  T2P_LAUNCH_CHECK();
}

void geglu(const void* z, long long M, int D, int dtype, void* out, cudaStream_t st) {
  T2P_CHECK(D % 8 == 0, "GEGLU width must be a multiple of 8");
  const unsigned blocks = static_cast<unsigned>(cdiv64(M * (D / 8), 256));
  if (dtype == kF32) launch_pdl(geglu_kernel<float>, dim3(blocks), dim3(256), 0, st, static_cast<const float*>(z), M, D, static_cast<float*>(out));
  else launch_pdl(geglu_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(z), M, D, static_cast<__nv_bfloat16*>(out));
  T2P_LAUNCH_CHECK();
}

}  // namespace t2p
