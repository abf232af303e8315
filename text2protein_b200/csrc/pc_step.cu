// Fused predictor / corrector half-steps of the reverse-SDE PC sampler, with in-kernel Philox noise.
//
// One launch per half-step replaces the ~25 ATen launches of the reference loop body
// (score_sde_pytorch/sampling.py:162-167 predictor, :179-199 corrector, :283-287 mask + .float();
// sde_lib.py:96-101,237-245 reverse-SDE discretisation):
//   predictor : x_mean = x - f + G^2 * score * (0.5 if probability_flow) ; x = x_mean + G * z
//   corrector : step = (snr * mean_b||z_b|| / mean_b||score_b||)^2 * 2 * alpha ; x_mean = x + step * score ;
//               x = x_mean + sqrt(2 step) * z          (batch-mean norms)
//   both      : x = where(mask, x, x_initial).float()
// The arithmetic follows the reference's dtype promotions (score is float64): products with the
// score run in double, noise terms in float, the state is rounded to float once per half-step.
// The corrector is a cooperative kernel: phase 1 reduces the per-sample squared norms of score and noise
// into per-chunk partials (no atomics, no zeroing), grid.sync(), phase 2 regenerates the same Philox
// normals and applies the update.  Memory traffic is the algorithmic 12 B/element (+1 B mask).
#include <cooperative_groups.h>

#include "kernels.h"
#include "philox.cuh"

namespace cg = cooperative_groups;

namespace t2p {
namespace {

struct StepParams {
  float* x;
  const void* score;
  int score_f64;
  int score_nhwc;
  const double* sigmas;     // optional: score = raw / sigmas[labels[b]]
  const long long* labels;
  const float* G;           // [B] predictor diffusion coefficient
  const float* sqrt_alpha;  // [B] VP drift (f = sqrt_alpha * x - x) or null (VE, f = 0)
  const float* alpha;       // [B] corrector alpha or null (= 1)
  float drift_scale;        // 1, or 0.5 for the probability-flow ODE
  int add_noise;            // 0 for the probability-flow ODE predictor
  float snr;
  const unsigned char* mask;  // [B*E] 1 = free to evolve, or null
  const float* x_init;
  float* x_mean_out;          // optional: masked x_mean (float), written by the predictor
  unsigned long long seed;
  long long stream_base, stream_mul;  // stream = stream_base + iter * stream_mul
  const long long* iter_ptr;          // device iteration counter or null (iter = 0)
  long long sample_offset;            // global index of local sample 0 (multi-GPU sharding)
  int B, C, HW;
  long long E;              // C*HW
  int chunks;               // chunks per sample (corrector)
  double* partial;          // [B*chunks][2]
};

__device__ __forceinline__ unsigned long long stream_of(const StepParams& p) {
  const long long it = p.iter_ptr ? *p.iter_ptr : 0;
  return static_cast<unsigned long long>(p.stream_base + it * p.stream_mul);
}

// score of 4 consecutive elements [e0, e0 + 4) of sample b, as the reference's float64 `h / used_sigmas`
// (1 / sigma is formed once per sample: 1 ulp of float64, far below the single float rounding of the state)
template <bool FAST>
__device__ __forceinline__ void load_score4(const StepParams& p, int b, long long e0, double inv_sigma, double (&s)[4]) {
  if (FAST || (!p.score_nhwc && !p.score_f64)) {
    const float4 v = *reinterpret_cast<const float4*>(static_cast<const float*>(p.score) + static_cast<long long>(b) * p.E + e0);
    s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
  } else if constexpr (!FAST) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      long long idx;
      if (p.score_nhwc) {
        const int c = static_cast<int>((e0 + i) / p.HW);
        const int pix = static_cast<int>((e0 + i) - static_cast<long long>(c) * p.HW);
        idx = (static_cast<long long>(b) * p.HW + pix) * p.C + c;
      } else {
        idx = static_cast<long long>(b) * p.E + e0 + i;
      }
      s[i] = p.score_f64 ? static_cast<const double*>(p.score)[idx]
                         : static_cast<double>(static_cast<const float*>(p.score)[idx]);
    }
  }
  if (p.sigmas) {
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] *= inv_sigma;
  }
}

// rounds the 4 updated values once to float, applies the condition mask (bit-exact: masked-out positions take
// x_initial) and stores x (and x_mean) as one 16-byte vector each
__device__ __forceinline__ void finish4(const StepParams& p, long long gi0, const double (&xn)[4], const double (&xm)[4]) {
  float xf[4], mf[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xf[i] = static_cast<float>(xn[i]);
    mf[i] = static_cast<float>(xm[i]);
  }
  if (p.mask) {
    const uchar4 m = *reinterpret_cast<const uchar4*>(p.mask + gi0);
    if (!(m.x && m.y && m.z && m.w)) {
      const float4 xi = *reinterpret_cast<const float4*>(p.x_init + gi0);
      if (!m.x) { xf[0] = xi.x; mf[0] = xi.x; }
      if (!m.y) { xf[1] = xi.y; mf[1] = xi.y; }
      if (!m.z) { xf[2] = xi.z; mf[2] = xi.z; }
      if (!m.w) { xf[3] = xi.w; mf[3] = xi.w; }
    }
  }
  *reinterpret_cast<float4*>(p.x + gi0) = make_float4(xf[0], xf[1], xf[2], xf[3]);
  if (p.x_mean_out) *reinterpret_cast<float4*>(p.x_mean_out + gi0) = make_float4(mf[0], mf[1], mf[2], mf[3]);
}

// grid = (slices of a sample, B): no per-element index arithmetic, 16-byte accesses throughout
// FAST: fp32 NCHW score, VE SDE (no drift), noise on -- the configuration of every shipped sampling config; the
// generic instantiation keeps the fp64 / NHWC score, VP drift and probability-flow variants of the API.
template <bool FAST>
__global__ void __launch_bounds__(256) predictor_kernel(const StepParams p) {
  const int b = blockIdx.y;
  const long long qps = p.E / 4;  // quads per sample
  const unsigned long long stream = stream_of(p);
  const float G = p.G[b];
  const float g2 = G * G;  // fp32, as G[:, None, None, None] ** 2
  const double gs = static_cast<double>(g2) * static_cast<double>(p.drift_scale);
  const double inv_sigma = p.sigmas ? 1.0 / p.sigmas[p.labels[b]] : 1.0;
  const float sa = p.sqrt_alpha ? p.sqrt_alpha[b] : 0.f;
  for (long long ql = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; ql < qps;
       ql += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long gi0 = static_cast<long long>(b) * p.E + ql * 4;
    const float4 xv = *reinterpret_cast<const float4*>(p.x + gi0);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    double s[4], xn[4], xm[4];
    load_score4<FAST>(p, b, ql * 4, inv_sigma, s);
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (FAST || p.add_noise)
      philox_normal4(p.seed, stream, static_cast<unsigned long long>((p.sample_offset + b) * qps + ql), z);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float f = 0.f;
      if (!FAST && p.sqrt_alpha) f = __fsub_rn(__fmul_rn(sa, xs[i]), xs[i]);
      const double rev_f = static_cast<double>(f) - gs * s[i];  // f == 0 for VE: exactly x + G^2 * score
      xm[i] = static_cast<double>(xs[i]) - rev_f;
      xn[i] = (FAST || p.add_noise) ? xm[i] + static_cast<double>(__fmul_rn(G, z[i])) : xm[i];
    }
    finish4(p, gi0, xn, xm);
  }
}

template <bool FAST>
__global__ void __launch_bounds__(256) corrector_kernel(const StepParams p) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[2][8];
  __shared__ float step_sh;
  const unsigned long long stream = stream_of(p);
  const long long items = static_cast<long long>(p.B) * p.chunks;
  const long long quads_per_chunk = p.E / 4 / p.chunks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- phase 1: partial squared norms of score and noise per (sample, chunk)
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = static_cast<int>(item / p.chunks);
    const int ck = static_cast<int>(item - static_cast<long long>(b) * p.chunks);
    double sg = 0.0, sn = 0.0;
    const double inv_sigma = p.sigmas ? 1.0 / p.sigmas[p.labels[b]] : 1.0;
    for (long long ql = threadIdx.x; ql < quads_per_chunk; ql += blockDim.x) {
      const long long qs = ck * quads_per_chunk + ql;  // quad within the sample
      float z[4];
      philox_normal4(p.seed, stream, static_cast<unsigned long long>((p.sample_offset + b) * (p.E / 4) + qs), z);
      double s[4];
      load_score4<FAST>(p, b, qs * 4, inv_sigma, s);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sg += s[i] * s[i];
        sn += static_cast<double>(z[i]) * static_cast<double>(z[i]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sn += __shfl_xor_sync(0xffffffffu, sn, o);
    }
    if (lane == 0) { red[0][warp] = sg; red[1][warp] = sn; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, c = 0.0;
      for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) { a += red[0][w]; c += red[1][w]; }
      p.partial[2 * item] = a;
      p.partial[2 * item + 1] = c;
    }
    __syncthreads();
  }
  grid.sync();

  // ---- step size from the batch-mean norms (every block recomputes it; B*chunks doubles)
  {
    double gsum = 0.0;
    float nsum = 0.f;
    for (int b = threadIdx.x; b < p.B; b += blockDim.x) {
      double a = 0.0, c = 0.0;
      for (int ck = 0; ck < p.chunks; ++ck) {
        a += p.partial[2 * (static_cast<long long>(b) * p.chunks + ck)];
        c += p.partial[2 * (static_cast<long long>(b) * p.chunks + ck) + 1];
      }
      gsum += sqrt(a);                        // ||grad_b||  (float64)
      nsum += static_cast<float>(sqrt(c));    // ||noise_b|| (float32 tensor in the reference)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
      nsum += __shfl_xor_sync(0xffffffffu, nsum, o);
    }
    if (lane == 0) { red[0][warp] = gsum; red[1][warp] = static_cast<double>(nsum); }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0;
      float c = 0.f;
      for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) { a += red[0][w]; c += static_cast<float>(red[1][w]); }
      const double grad_norm = a / p.B;
      const float noise_norm = c / static_cast<float>(p.B);
      const float sn = __fmul_rn(p.snr, noise_norm);  // python float * fp32 0-dim tensor -> fp32
      const double r = static_cast<double>(sn) / grad_norm;
      step_sh = static_cast<float>(r * r * 2.0);      // * alpha (fp32 [B]) demotes the 0-dim double
    }
    __syncthreads();
  }
  const float step0 = step_sh;

  // ---- phase 2: apply (one (sample, chunk) item per block iteration, as in phase 1)
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = static_cast<int>(item / p.chunks);
    const int ck = static_cast<int>(item - static_cast<long long>(b) * p.chunks);
    const double inv_sigma = p.sigmas ? 1.0 / p.sigmas[p.labels[b]] : 1.0;
    const float step = p.alpha ? __fmul_rn(step0, p.alpha[b]) : step0;
    const float nscale = sqrtf(__fmul_rn(step, 2.f));
    const double dstep = static_cast<double>(step);
    for (long long ql = threadIdx.x; ql < quads_per_chunk; ql += blockDim.x) {
      const long long qs = ck * quads_per_chunk + ql;
      const long long gi0 = static_cast<long long>(b) * p.E + qs * 4;
      float z[4];
      philox_normal4(p.seed, stream, static_cast<unsigned long long>((p.sample_offset + b) * (p.E / 4) + qs), z);
      const float4 xv = *reinterpret_cast<const float4*>(p.x + gi0);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      double s[4], xn[4], xm[4];
      load_score4<FAST>(p, b, qs * 4, inv_sigma, s);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xm[i] = static_cast<double>(xs[i]) + dstep * s[i];
        xn[i] = xm[i] + static_cast<double>(__fmul_rn(nscale, z[i]));
      }
      finish4(p, gi0, xn, xm);
    }
  }
}

__global__ void philox_fill_kernel(unsigned long long seed, unsigned long long stream, long long first_quad,
                                   long long quads, float scale, float* out) {
  for (long long qi = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; qi < quads;
       qi += static_cast<long long>(gridDim.x) * blockDim.x) {
    float z[4];
    philox_normal4(seed, stream, static_cast<unsigned long long>(first_quad + qi), z);
    *reinterpret_cast<float4*>(out + qi * 4) = make_float4(z[0] * scale, z[1] * scale, z[2] * scale, z[3] * scale);
  }
}

__global__ void philox_bits_kernel(unsigned long long seed, unsigned long long stream, long long first_quad,
                                   long long quads, unsigned int* out) {
  const long long qi = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (qi >= quads) return;
  const uint4 r = philox4x32_10(make_uint4(static_cast<unsigned>(first_quad + qi), static_cast<unsigned>((first_quad + qi) >> 32),
                                           static_cast<unsigned>(stream), static_cast<unsigned>(stream >> 32)),
                                make_uint2(static_cast<unsigned>(seed), static_cast<unsigned>(seed >> 32)));
  reinterpret_cast<uint4*>(out)[qi] = r;
}

// per-iteration scalars of a graph-replayed run: labels[b], G[b] from host-built tables, iteration counter
__global__ void run_prep_kernel(long long* state /*[0]=iter, [1]=next*/, const long long* label_table,
                                const float* g_table, int B, long long* labels, float* G) {
  const long long it = state[1];
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    labels[b] = label_table[it];
    G[b] = g_table[it];
  }
  __syncthreads();
  if (threadIdx.x == 0) { state[0] = it; state[1] = it + 1; }
}

// sampling.py:260-275 condition application on the prior sample, bit-exact mask semantics:
// x = where(mask, x, x_fixed); used once per run.
__global__ void apply_mask_kernel(float* x, const unsigned char* mask, const float* fixed, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n && !mask[i]) x[i] = fixed[i];
}

StepParams to_params(const PcStepArgs& a) {
  StepParams p{};
  p.x = a.x; p.score = a.score; p.score_f64 = (a.score_dtype == kF64); p.score_nhwc = a.score_nhwc;
  T2P_CHECK(a.score_dtype == kF64 || a.score_dtype == kF32, "score must be fp32 or fp64");
  p.sigmas = a.sigmas; p.labels = a.labels; p.G = a.G; p.sqrt_alpha = a.sqrt_alpha; p.alpha = a.alpha;
  p.drift_scale = a.probability_flow ? 0.5f : 1.f;
  p.add_noise = a.probability_flow ? 0 : 1;
  p.snr = a.snr; p.mask = a.mask; p.x_init = a.x_init; p.x_mean_out = a.x_mean_out;
  p.seed = a.seed; p.stream_base = a.stream_base; p.stream_mul = a.stream_mul; p.iter_ptr = a.iter_ptr;
  p.sample_offset = a.sample_offset;
  p.B = a.B; p.C = a.C; p.HW = a.HW; p.E = static_cast<long long>(a.C) * a.HW;
  T2P_CHECK(p.E % 4 == 0, "C*N*N must be a multiple of 4");
  T2P_CHECK((reinterpret_cast<uintptr_t>(a.x) & 15) == 0, "x must be 16-byte aligned");
  T2P_CHECK((reinterpret_cast<uintptr_t>(a.score) & 15) == 0, "score must be 16-byte aligned");
  T2P_CHECK((reinterpret_cast<uintptr_t>(a.mask) & 3) == 0, "mask must be 4-byte aligned");
  T2P_CHECK((reinterpret_cast<uintptr_t>(a.x_init) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.x_mean_out) & 15) == 0,
            "x_init / x_mean_out must be 16-byte aligned");
  if (a.sigmas) T2P_CHECK(a.labels != nullptr, "labels required with sigmas");
  if (a.mask) T2P_CHECK(a.x_init != nullptr, "x_init required with mask");
  p.partial = a.partial;
  return p;
}

int num_sms() {
  static int n = [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

}  // namespace

void pc_predictor_step(const PcStepArgs& a, cudaStream_t st) {
  StepParams p = to_params(a);
  T2P_CHECK(a.G != nullptr, "predictor needs G");
  const long long qps = p.E / 4;
  const int bx = static_cast<int>(std::max<long long>(1, std::min<long long>(cdiv64(qps, 256), cdiv64(num_sms() * 8LL, p.B))));
  const bool fast = !p.score_nhwc && !p.score_f64 && p.sqrt_alpha == nullptr && p.add_noise;
  if (fast) predictor_kernel<true><<<dim3(bx, p.B), 256, 0, st>>>(p);
  else predictor_kernel<false><<<dim3(bx, p.B), 256, 0, st>>>(p);
  T2P_LAUNCH_CHECK();
}

int pc_corrector_chunks(int B, long long E) {
  const long long q = E / 4;
  int chunks = 1;
  // several (sample, chunk) items per resident block, each chunk still at least one quad per thread
  while (static_cast<long long>(B) * chunks < 8LL * num_sms() && (q % (chunks * 2) == 0) && q / (chunks * 2) >= 256)
    chunks *= 2;
  return chunks;
}

void pc_corrector_step(const PcStepArgs& a, cudaStream_t st) {
  StepParams p = to_params(a);
  T2P_CHECK(a.partial != nullptr && a.chunks > 0, "corrector needs the partial-sum workspace");
  T2P_CHECK((p.E / 4) % a.chunks == 0, "chunks must divide the quads of a sample");
  p.chunks = a.chunks;
  const bool fast = !p.score_nhwc && !p.score_f64;
  static int max_blocks[2] = {0, 0};
  if (!max_blocks[fast]) {
    int per_sm = 0;
    if (fast) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, corrector_kernel<true>, 256, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, corrector_kernel<false>, 256, 0);
    max_blocks[fast] = std::max(1, per_sm) * num_sms();
  }
  // cooperative launch: every block must be resident; blocks loop over the (sample, chunk) items
  const int blocks = static_cast<int>(std::min<long long>(static_cast<long long>(p.B) * p.chunks, max_blocks[fast]));
  void* args[] = {&p};
  void* fn = fast ? reinterpret_cast<void*>(corrector_kernel<true>) : reinterpret_cast<void*>(corrector_kernel<false>);
  T2P_CUDA(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(256), args, 0, st));
}

void philox_normal_fill(unsigned long long seed, unsigned long long stream, long long first_element, long long count,
                        float scale, float* out, cudaStream_t st) {
  T2P_CHECK(first_element % 4 == 0 && count % 4 == 0, "philox fill works on whole quads");
  const long long quads = count / 4;
  const int blocks = static_cast<int>(std::min<long long>(cdiv64(quads, 256), 148LL * 16));
  philox_fill_kernel<<<std::max(blocks, 1), 256, 0, st>>>(seed, stream, first_element / 4, quads, scale, out);
  T2P_LAUNCH_CHECK();
}

void philox_bits_fill(unsigned long long seed, unsigned long long stream, long long first_quad, long long quads,
                      unsigned int* out, cudaStream_t st) {
  philox_bits_kernel<<<static_cast<unsigned>(cdiv64(quads, 256)), 256, 0, st>>>(seed, stream, first_quad, quads, out);
  T2P_LAUNCH_CHECK();
}

void run_prep(long long* state, const long long* label_table, const float* g_table, int B, long long* labels, float* G,
              cudaStream_t st) {
  run_prep_kernel<<<1, 256, 0, st>>>(state, label_table, g_table, B, labels, G);
  T2P_LAUNCH_CHECK();
}

void apply_mask(float* x, const unsigned char* mask, const float* fixed, long long n, cudaStream_t st) {
  apply_mask_kernel<<<static_cast<unsigned>(cdiv64(n, 256)), 256, 0, st>>>(x, mask, fixed, n);
  T2P_LAUNCH_CHECK();
}

}  // namespace t2p
