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
//
// Work decomposition (both kernels): a sample is cut into rows of kThreads quads (4 consecutive elements, one
// 16-byte access per tensor); the B * rows_per_sample rows are dealt out in contiguous, equal (+-1) ranges to a
// single wave of resident blocks (2 x 512 threads per SM), so there is no partial second wave and no
// per-element index arithmetic (the sample index is uniform per row).
//
// The corrector is a cooperative kernel.  Phase 1 generates the normals of the block's rows ONCE, keeps them in
// shared memory (8 KB per row, up to 13 rows per block; rows beyond the cache are regenerated in phase 2),
// accumulates sum(score^2) and sum(z^2) per (block, sample) into a partial slot (no atomics, no zeroing, fixed
// order) and prefetches the block's x rows into L2; grid.sync(); every block folds the partials into the
// batch-mean step size; phase 2 applies the update.  DRAM traffic is the algorithmic 12 B/element (+1 B mask).
#include <cooperative_groups.h>

#include "kernels.h"
#include "philox.cuh"

namespace cg = cooperative_groups;

namespace t2p {
namespace {

constexpr int kThreads = 512;       // threads per block = quads per row
constexpr int kMaxCacheRows = 13;   // 13 x 8 KB of cached normals per block, two blocks per SM
constexpr int kMaxBlocks = 4096;    // bound used to size the corrector's partial workspace

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
  float* x_mean_out;          // optional: masked x_mean (float)
  unsigned long long seed;
  long long stream_base, stream_mul;  // stream = stream_base + iter * stream_mul
  const long long* iter_ptr;          // device iteration counter or null (iter = 0)
  long long sample_offset;            // global index of local sample 0 (multi-GPU sharding)
  int B, C, HW;
  long long E;              // C*HW
  int qps;                  // quads per sample, E / 4
  int rps;                  // rows per sample, ceil(qps / kThreads)
  int rows;                 // B * rps
  int cache_rows;           // corrector: rows of normals a block keeps in shared memory
  int prefetch;             // corrector: phase 1 prefetches the x rows of phase 2 into L2
  int skip_conditioned;     // fully conditioned quads copy x_initial without loading x / score or drawing noise
  double* partial;          // corrector: [gridDim.x + B][2], slot = block + sample
};

__device__ __forceinline__ unsigned long long stream_of(const StepParams& p) {
  const long long it = p.iter_ptr ? *p.iter_ptr : 0;
  return static_cast<unsigned long long>(p.stream_base + it * p.stream_mul);
}

// contiguous row range of this block: [rows * j / G, rows * (j + 1) / G)
__device__ __forceinline__ void block_rows(const StepParams& p, int& r0, int& r1) {
  r0 = static_cast<int>(static_cast<long long>(p.rows) * blockIdx.x / gridDim.x);
  r1 = static_cast<int>(static_cast<long long>(p.rows) * (blockIdx.x + 1) / gridDim.x);
}

// the block whose range holds row r (inverse of block_rows)
__device__ __forceinline__ int block_of_row(const StepParams& p, int r) {
  return static_cast<int>((static_cast<long long>(r + 1) * gridDim.x + p.rows - 1) / p.rows) - 1;
}

// RAW network output of 4 consecutive elements [e0, e0 + 4) of sample b as doubles; the 1 / sigma of the
// reference's float64 `h / used_sigmas` is folded into the per-sample coefficient that multiplies it
// (1 / sigma is formed once per sample: 1 ulp of float64, far below the single float rounding of the state)
template <bool FAST>
__device__ __forceinline__ void load_score4(const StepParams& p, int b, long long e0, double (&s)[4]) {
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
}

__device__ __forceinline__ double inv_sigma_of(const StepParams& p, int b) {
  return p.sigmas ? 1.0 / p.sigmas[p.labels[b]] : 1.0;
}

// rounds the 4 updated values once to float, applies the condition mask (bit-exact: masked-out positions take
// x_initial) and stores x (and x_mean) as one 16-byte vector each
__device__ __forceinline__ void finish4(const StepParams& p, long long gi0, uchar4 m, const double (&xn)[4], const double (&xm)[4]) {
  float xf[4], mf[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xf[i] = static_cast<float>(xn[i]);
    mf[i] = static_cast<float>(xm[i]);
  }
  if (!(m.x && m.y && m.z && m.w)) {
    const float4 xi = *reinterpret_cast<const float4*>(p.x_init + gi0);
    if (!m.x) { xf[0] = xi.x; mf[0] = xi.x; }
    if (!m.y) { xf[1] = xi.y; mf[1] = xi.y; }
    if (!m.z) { xf[2] = xi.z; mf[2] = xi.z; }
    if (!m.w) { xf[3] = xi.w; mf[3] = xi.w; }
  }
  *reinterpret_cast<float4*>(p.x + gi0) = make_float4(xf[0], xf[1], xf[2], xf[3]);
  if (p.x_mean_out) *reinterpret_cast<float4*>(p.x_mean_out + gi0) = make_float4(mf[0], mf[1], mf[2], mf[3]);
}

__device__ __forceinline__ uchar4 load_mask4(const StepParams& p, long long gi0) {
  return p.mask ? *reinterpret_cast<const uchar4*>(p.mask + gi0) : make_uchar4(1, 1, 1, 1);
}

// walks the rows [r0, r1) of a block keeping (sample, row within the sample) without a division per row
struct RowIter {
  int r, r1, b, rq;
  __device__ __forceinline__ RowIter(const StepParams& p, int r0_, int r1_) : r(r0_), r1(r1_) {
    b = r0_ / p.rps;
    rq = r0_ - b * p.rps;
  }
  __device__ __forceinline__ bool done() const { return r >= r1; }
  __device__ __forceinline__ void next(const StepParams& p) {
    ++r;
    if (++rq == p.rps) { rq = 0; ++b; }
  }
  __device__ __forceinline__ int quad() const { return rq * kThreads + static_cast<int>(threadIdx.x); }  // within the sample
  __device__ __forceinline__ bool valid(const StepParams& p) const { return r < r1 && quad() < p.qps; }
  __device__ __forceinline__ long long element(const StepParams& p) const {
    return static_cast<long long>(b) * p.E + static_cast<long long>(quad()) * 4;
  }
};

// The mask of a row is loaded one row ahead: a quad whose 4 positions are all conditioned (mask == 0) takes
// x_initial whatever the update would be, so its x / score loads and its Philox + Box-Muller work are skipped
// (with a length condition more than half of the quads, whole warps at a time: one image row of 128 residues is
// one warp).  The result is bit-identical to computing the update and discarding it.
__device__ __forceinline__ bool all_conditioned(uchar4 m) { return !(m.x | m.y | m.z | m.w); }

__device__ __forceinline__ void copy_initial4(const StepParams& p, long long gi0) {
  const float4 xi = *reinterpret_cast<const float4*>(p.x_init + gi0);
  *reinterpret_cast<float4*>(p.x + gi0) = xi;
  if (p.x_mean_out) *reinterpret_cast<float4*>(p.x_mean_out + gi0) = xi;
}

// FAST: fp32 NCHW score, VE SDE (no drift), noise on -- the configuration of every shipped sampling config; the
// generic instantiation keeps the fp64 / NHWC score, VP drift and probability-flow variants of the API.
template <bool FAST>
__global__ void __launch_bounds__(kThreads, 2) predictor_kernel(const StepParams p) {
  int r0, r1;
  block_rows(p, r0, r1);
  const unsigned long long stream = stream_of(p);
  int cur_b = -1;
  float G = 0.f, sa = 0.f;
  double coef = 0.0;  // G^2 (fp32, as G[:, None, None, None] ** 2) * drift_scale / sigma
  RowIter it(p, r0, r1);
  uchar4 m = it.valid(p) ? load_mask4(p, it.element(p)) : make_uchar4(1, 1, 1, 1);
  while (!it.done()) {
    RowIter nx = it;
    nx.next(p);
    const uchar4 m_next = nx.valid(p) ? load_mask4(p, nx.element(p)) : make_uchar4(1, 1, 1, 1);
    const int b = it.b, ql = it.quad();
    if (b != cur_b) {  // uniform: a block crosses a sample boundary at most every rps rows
      cur_b = b;
      G = p.G[b];
      coef = static_cast<double>(__fmul_rn(G, G)) * static_cast<double>(p.drift_scale) * inv_sigma_of(p, b);
      sa = p.sqrt_alpha ? p.sqrt_alpha[b] : 0.f;
    }
    if (ql < p.qps) {
      const long long gi0 = it.element(p);
      if (p.skip_conditioned && all_conditioned(m)) {
        copy_initial4(p, gi0);
      } else {
        const float4 xv = *reinterpret_cast<const float4*>(p.x + gi0);
        double s[4], xn[4], xm[4];
        load_score4<FAST>(p, b, static_cast<long long>(ql) * 4, s);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (FAST || p.add_noise)
          philox_normal4(p.seed, stream, static_cast<unsigned long long>((p.sample_offset + b) * p.qps + ql), z);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          double base = static_cast<double>(xs[i]);  // x - f; f == 0 for VE: exactly x + G^2 * score
          if (!FAST && p.sqrt_alpha) base -= static_cast<double>(__fsub_rn(__fmul_rn(sa, xs[i]), xs[i]));
          xm[i] = fma(coef, s[i], base);
          xn[i] = (FAST || p.add_noise) ? xm[i] + static_cast<double>(__fmul_rn(G, z[i])) : xm[i];
        }
        finish4(p, gi0, m, xn, xm);
      }
    }
    m = m_next;
    it = nx;
  }
}

// block-wide sum of two doubles (result valid in thread 0); `red` is reused, hence the trailing barrier
__device__ __forceinline__ void block_sum2(double& a, double& c, double (*red)[kThreads / 32]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if (lane == 0) { red[0][warp] = a; red[1][warp] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0.0; c = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { a += red[0][w]; c += red[1][w]; }
  }
  __syncthreads();
}

template <bool FAST>
__global__ void __launch_bounds__(kThreads, 2) corrector_kernel(const StepParams p) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float4 zcache[];  // [cache_rows][kThreads] normals of this block's first rows
  __shared__ double red[2][kThreads / 32];
  __shared__ float step_sh;
  const unsigned long long stream = stream_of(p);
  int r0, r1;
  block_rows(p, r0, r1);

  // ---- phase 1: normals (kept), partial squared norms of the RAW score and of the noise per (block, sample)
  {
    double sg = 0.0, sn = 0.0;
    int cur_b = r0 < r1 ? r0 / p.rps : 0;
    auto flush = [&](int b) {
      block_sum2(sg, sn, red);
      if (threadIdx.x == 0) {
        const double inv = inv_sigma_of(p, b);
        p.partial[2 * (blockIdx.x + b)] = sg * inv * inv;  // sum (h / sigma)^2
        p.partial[2 * (blockIdx.x + b) + 1] = sn;
      }
      sg = 0.0; sn = 0.0;
    };
    int b = cur_b, rq = r0 - b * p.rps;
    for (int r = r0; r < r1; ++r, ++rq) {
      if (rq == p.rps) { rq = 0; ++b; }
      if (b != cur_b) { flush(cur_b); cur_b = b; }
      const int ql = rq * kThreads + static_cast<int>(threadIdx.x);
      if (ql >= p.qps) continue;
      const long long gi0 = static_cast<long long>(b) * p.E + static_cast<long long>(ql) * 4;
      double s[4];
      load_score4<FAST>(p, b, static_cast<long long>(ql) * 4, s);
      if (p.prefetch) {  // phase 2's first-touch operands: x (one 128-byte line per 8 threads) and the mask (per warp)
        if ((threadIdx.x & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + gi0));
        if ((threadIdx.x & 31) == 0 && p.mask) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.mask + gi0));
      }
      float z[4];
      philox_normal4(p.seed, stream, static_cast<unsigned long long>((p.sample_offset + b) * p.qps + ql), z);
      if (r - r0 < p.cache_rows) zcache[(r - r0) * kThreads + threadIdx.x] = make_float4(z[0], z[1], z[2], z[3]);
      if constexpr (FAST) {
        // squares of 4 elements summed in float (the inputs are floats), the running sums in double
        const float h0 = static_cast<float>(s[0]), h1 = static_cast<float>(s[1]), h2 = static_cast<float>(s[2]),
                    h3 = static_cast<float>(s[3]);
        sg += static_cast<double>(__fmaf_rn(h3, h3, __fmaf_rn(h2, h2, __fmaf_rn(h1, h1, __fmul_rn(h0, h0)))));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) sg += s[i] * s[i];
      }
      sn += static_cast<double>(__fmaf_rn(z[3], z[3], __fmaf_rn(z[2], z[2], __fmaf_rn(z[1], z[1], __fmul_rn(z[0], z[0])))));
    }
    if (r0 < r1) flush(cur_b);
  }
  grid.sync();

  // ---- step size from the batch-mean norms (every block recomputes it from the (block, sample) partials)
  {
    double gsum = 0.0, nsum_d = 0.0;
    float nsum = 0.f;
    for (int b = threadIdx.x; b < p.B; b += kThreads) {
      const int j0 = block_of_row(p, b * p.rps), j1 = block_of_row(p, (b + 1) * p.rps - 1);
      double a = 0.0, c = 0.0;
      for (int j = j0; j <= j1; ++j) {
        a += p.partial[2 * (j + b)];
        c += p.partial[2 * (j + b) + 1];
      }
      gsum += sqrt(a);                        // ||grad_b||  (float64)
      nsum += static_cast<float>(sqrt(c));    // ||noise_b|| (float32 tensor in the reference)
    }
    nsum_d = static_cast<double>(nsum);
    block_sum2(gsum, nsum_d, red);
    if (threadIdx.x == 0) {
      const double grad_norm = gsum / p.B;
      const float noise_norm = static_cast<float>(nsum_d) / static_cast<float>(p.B);
      const float sn = __fmul_rn(p.snr, noise_norm);  // python float * fp32 0-dim tensor -> fp32
      const double r = static_cast<double>(sn) / grad_norm;
      step_sh = static_cast<float>(r * r * 2.0);      // * alpha (fp32 [B]) demotes the 0-dim double
    }
    __syncthreads();
  }
  const float step0 = step_sh;

  // ---- phase 2: apply
  int cur_b = -1;
  float nscale = 0.f;
  double coef = 0.0;  // step / sigma
  RowIter it(p, r0, r1);
  uchar4 m = it.valid(p) ? load_mask4(p, it.element(p)) : make_uchar4(1, 1, 1, 1);
  while (!it.done()) {
    RowIter nx = it;
    nx.next(p);
    const uchar4 m_next = nx.valid(p) ? load_mask4(p, nx.element(p)) : make_uchar4(1, 1, 1, 1);
    const int b = it.b, ql = it.quad(), slot = it.r - r0;
    if (b != cur_b) {
      cur_b = b;
      const float step = p.alpha ? __fmul_rn(step0, p.alpha[b]) : step0;
      nscale = sqrtf(__fmul_rn(step, 2.f));
      coef = static_cast<double>(step) * inv_sigma_of(p, b);
    }
    if (ql < p.qps) {
      const long long gi0 = it.element(p);
      if (p.skip_conditioned && all_conditioned(m)) {
        copy_initial4(p, gi0);
      } else {
        const float4 xv = *reinterpret_cast<const float4*>(p.x + gi0);
        double s[4], xn[4], xm[4];
        load_score4<FAST>(p, b, static_cast<long long>(ql) * 4, s);
        float z[4];
        if (slot < p.cache_rows) {
          const float4 zv = zcache[slot * kThreads + threadIdx.x];
          z[0] = zv.x; z[1] = zv.y; z[2] = zv.z; z[3] = zv.w;
        } else {
          philox_normal4(p.seed, stream, static_cast<unsigned long long>((p.sample_offset + b) * p.qps + ql), z);
        }
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          xm[i] = fma(coef, s[i], static_cast<double>(xs[i]));
          xn[i] = xm[i] + static_cast<double>(__fmul_rn(nscale, z[i]));
        }
        finish4(p, gi0, m, xn, xm);
      }
    }
    m = m_next;
    it = nx;
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
  T2P_CHECK(a.B > 0 && p.E > 0 && p.E % 4 == 0, "C*N*N must be a positive multiple of 4");
  T2P_CHECK(p.E / 4 < (1LL << 30), "sample too large");
  p.qps = static_cast<int>(p.E / 4);
  p.rps = (p.qps + kThreads - 1) / kThreads;
  T2P_CHECK(static_cast<long long>(a.B) * p.rps < (1LL << 31), "batch too large");
  p.rows = a.B * p.rps;
  T2P_CHECK((reinterpret_cast<uintptr_t>(a.x) & 15) == 0, "x must be 16-byte aligned");
  T2P_CHECK((reinterpret_cast<uintptr_t>(a.score) & 15) == 0, "score must be 16-byte aligned");
  T2P_CHECK((reinterpret_cast<uintptr_t>(a.mask) & 3) == 0, "mask must be 4-byte aligned");
  T2P_CHECK((reinterpret_cast<uintptr_t>(a.x_init) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.x_mean_out) & 15) == 0,
            "x_init / x_mean_out must be 16-byte aligned");
  if (a.sigmas) T2P_CHECK(a.labels != nullptr, "labels required with sigmas");
  if (a.mask) T2P_CHECK(a.x_init != nullptr, "x_init required with mask");
  p.partial = a.partial;
  static const bool no_skip = getenv("T2P_STEP_NOSKIP") != nullptr;  // A/B knob
  p.skip_conditioned = (a.mask != nullptr && !no_skip) ? 1 : 0;
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

// resident blocks of `fn` over the whole device (one wave), at `smem` bytes of dynamic shared memory
int resident_blocks(const void* fn, size_t smem) {
  int per_sm = 0;
  T2P_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, smem));
  T2P_CHECK(per_sm > 0, "step kernel does not fit on an SM");
  return std::min(per_sm * num_sms(), kMaxBlocks);
}

}  // namespace

void pc_predictor_step(const PcStepArgs& a, cudaStream_t st) {
  StepParams p = to_params(a);
  T2P_CHECK(a.G != nullptr, "predictor needs G");
  const bool fast = !p.score_nhwc && !p.score_f64 && p.sqrt_alpha == nullptr && p.add_noise;
  static int wave[2] = {0, 0};
  if (!wave[fast])
    wave[fast] = resident_blocks(fast ? reinterpret_cast<const void*>(predictor_kernel<true>)
                                      : reinterpret_cast<const void*>(predictor_kernel<false>), 0);
  const int blocks = std::min(wave[fast], p.rows);
  if (fast) predictor_kernel<true><<<blocks, kThreads, 0, st>>>(p);
  else predictor_kernel<false><<<blocks, kThreads, 0, st>>>(p);
  T2P_LAUNCH_CHECK();
}

// doubles of partial-sum workspace the corrector needs for a batch of B samples: one (score, noise) slot per
// (block, sample) pair that occurs, indexed block + sample
long long pc_corrector_workspace_doubles(int B) { return 2LL * (kMaxBlocks + static_cast<long long>(B)); }

void pc_corrector_step(const PcStepArgs& a, cudaStream_t st) {
  StepParams p = to_params(a);
  T2P_CHECK(a.partial != nullptr, "corrector needs the partial-sum workspace");
  const bool fast = !p.score_nhwc && !p.score_f64;
  const void* fn = fast ? reinterpret_cast<const void*>(corrector_kernel<true>)
                        : reinterpret_cast<const void*>(corrector_kernel<false>);
  constexpr size_t kRowBytes = sizeof(float4) * kThreads;
  static bool attr_set[2] = {false, false};
  if (!attr_set[fast]) {
    T2P_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaxCacheRows * kRowBytes)));
    T2P_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set[fast] = true;
  }
  // shared-memory cache sized for the rows a block gets when two blocks per SM are resident
  static const bool no_cache = getenv("T2P_STEP_NOCACHE") != nullptr;
  static const bool no_prefetch = getenv("T2P_STEP_NOPREFETCH") != nullptr;
  const int want = static_cast<int>(cdiv64(p.rows, 2LL * num_sms()));
  p.cache_rows = no_cache ? 0 : std::min(want, kMaxCacheRows);
  p.prefetch = no_prefetch ? 0 : 1;
  const size_t smem = p.cache_rows * kRowBytes;
  static int wave[2][kMaxCacheRows + 1] = {};
  if (!wave[fast][p.cache_rows]) wave[fast][p.cache_rows] = resident_blocks(fn, smem);
  // cooperative launch: every block must be resident
  const int blocks = std::min(wave[fast][p.cache_rows], p.rows);
  void* args[] = {&p};
  T2P_CUDA(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(kThreads), args, smem, st));
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
