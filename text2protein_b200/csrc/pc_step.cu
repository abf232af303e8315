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
// Work decomposition (all kernels): a sample is cut into rows of kThreads quads (4 consecutive elements, one
// 16-byte access per tensor).  A single wave of resident blocks (2 x 512 threads per SM) takes the rows
// INTERLEAVED (block j: rows j, j + G, j + 2G, ...): a quad whose 4 positions are all conditioned takes x_initial
// without its x / score loads and without Philox + Box-Muller, so with a length condition the cost of a row
// depends on where it lies in its sample; contiguous row ranges left a third of the SMs idle (ncu: SM-active
// cycles 64 % of elapsed), interleaving spreads every sample over all SMs.
//
// The corrector is a cooperative kernel.  Phase 1 generates the normals of the block's rows ONCE and keeps them
// in shared memory (8 KB per row; rows beyond the cache are regenerated in phase 2), writes sum(score^2) and
// sum(z^2) of every row to a partial slot (warp shuffles in float, 16 warps summed in double: no atomics, fixed
// order, independent of the grid size) and prefetches what phase 2 reads into L2; grid.sync(); every block folds
// the row partials into the batch-mean step size; phase 2 applies the update.
#include <cooperative_groups.h>

#include "kernels.h"
#include "philox.cuh"

namespace cg = cooperative_groups;

namespace t2p {
namespace {

#ifndef T2P_STEP_THREADS
#define T2P_STEP_THREADS 512
#endif
#ifndef T2P_STEP_BLOCKS
#define T2P_STEP_BLOCKS 2
#endif
constexpr int kThreads = T2P_STEP_THREADS;     // threads per block = quads per row
constexpr int kBlocksPerSM = T2P_STEP_BLOCKS;  // resident blocks per SM the kernels are compiled for (register cap)
constexpr int kWarps = kThreads / 32;
// corrector: rows of cached normals per block (16 B per thread and row), kBlocksPerSM blocks sharing 227 KB with
// their static shared memory and the 1 KB the system reserves per block
constexpr int kMaxCacheRows = (227 * 1024 / kBlocksPerSM - 4096) / (16 * kThreads);
constexpr int kMaxBlocks = 4096;
constexpr int kSumChunk = 16;       // corrector phase 1: rows between two flushes of the per-row sums

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
  int prefetch;             // corrector: phase 1 prefetches the operands of phase 2 into L2
  int skip_conditioned;     // fully conditioned quads copy x_initial without loading x / score or drawing noise
  int in_place;             // conditioned positions of x_out (and x_mean_out) already hold x_initial: such quads are not touched
  double* partial;          // corrector: [rows][2] = sum (h / sigma)^2, sum z^2 of each row
  float* x_out;             // new state (== x unless the caller asked for an out-of-place step)
  int symmetrize;           // channels 0, 1: symmetric part wherever (i, j) and (j, i) are both free (needs x_out != x)
  int W;                    // image width (symmetrize: HW == W * W)
  const long long* last_iter_ptr;  // x_mean_out is written only in iteration *last_iter_ptr (null: always)
  PeerGroup peers;          // world > 1: the step size is the mean over the global batch (mailbox exchange)
  const long long* tag_base_ptr;
};

__device__ __forceinline__ unsigned long long stream_of(const StepParams& p) {
  const long long it = p.iter_ptr ? *p.iter_ptr : 0;
  return static_cast<unsigned long long>(p.stream_base + it * p.stream_mul);
}

// walks the rows blockIdx.x, blockIdx.x + gridDim.x, ... of a block keeping (sample, row within the sample)
// without a division per row; k counts the block's rows
struct RowIter {
  int r, b, rq, k, db, drq;
  __device__ __forceinline__ explicit RowIter(const StepParams& p) {
    r = blockIdx.x;
    b = r / p.rps;
    rq = r - b * p.rps;
    k = 0;
    db = gridDim.x / p.rps;
    drq = gridDim.x - db * p.rps;
  }
  __device__ __forceinline__ bool done(const StepParams& p) const { return r >= p.rows; }
  __device__ __forceinline__ void next(const StepParams& p) {
    r += gridDim.x;
    ++k;
    b += db;
    rq += drq;
    if (rq >= p.rps) { rq -= p.rps; ++b; }
  }
  __device__ __forceinline__ int quad() const { return rq * kThreads + static_cast<int>(threadIdx.x); }  // within the sample
  __device__ __forceinline__ bool valid(const StepParams& p) const { return r < p.rows && quad() < p.qps; }
  __device__ __forceinline__ long long element(const StepParams& p) const {
    return static_cast<long long>(b) * p.E + static_cast<long long>(quad()) * 4;
  }
};

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
__device__ __forceinline__ void finish4(const StepParams& p, float* xmean, long long gi0, uchar4 m, const double (&xn)[4],
                                        const double (&xm)[4]) {
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
  *reinterpret_cast<float4*>(p.x_out + gi0) = make_float4(xf[0], xf[1], xf[2], xf[3]);
  if (xmean) *reinterpret_cast<float4*>(xmean + gi0) = make_float4(mf[0], mf[1], mf[2], mf[3]);
}

// x_mean destination of this launch: the sampler returns the x_mean of its last predictor step only
// (sampling.py:289), so inside a run the store is skipped in every iteration but the last
__device__ __forceinline__ float* xmean_of(const StepParams& p) {
  if (!p.x_mean_out) return nullptr;
  if (p.last_iter_ptr && p.iter_ptr && *p.iter_ptr != *p.last_iter_ptr) return nullptr;
  return p.x_mean_out;
}

__device__ __forceinline__ uchar4 load_mask4(const StepParams& p, long long gi0) {
  return p.mask ? *reinterpret_cast<const uchar4*>(p.mask + gi0) : make_uchar4(1, 1, 1, 1);
}

// A quad whose 4 positions are all conditioned (mask == 0) takes x_initial whatever the update would be: its
// x / score loads and its Philox + Box-Muller work are skipped (with a length condition more than half of the
// quads, whole warps at a time: one image row of 128 residues is one warp).  Bit-identical to computing the
// update and discarding it.  Inside t2p_pc_run the conditioned positions of x and x_mean hold x_initial from
// the start of the run and nothing else writes them (in_place): there such a quad costs its 4 mask bytes.
__device__ __forceinline__ bool all_conditioned(uchar4 m) { return !(m.x | m.y | m.z | m.w); }

__device__ __forceinline__ void copy_initial4(const StepParams& p, float* xmean, long long gi0) {
  if (p.in_place) return;
  const float4 xi = *reinterpret_cast<const float4*>(p.x_init + gi0);
  *reinterpret_cast<float4*>(p.x_out + gi0) = xi;
  if (xmean) *reinterpret_cast<float4*>(xmean + gi0) = xi;
}

// ---- symmetrisation (opt-in; the reference has none: downstream takes np.triu, rosetta_min/utils.py:140,157).
// The half-step is linear in (x, score, z) given the per-sample scalars, so the symmetric part of the updated map
// is 0.5 (u[i][j] + u[j][i]) with u the float64 update; both positions evaluate the same commutative sum and
// round it once, hence the stored channels 0 / 1 are EXACTLY symmetric.  A position is symmetrised only when it
// and its transpose are both free (every condition builder of the reference produces symmetric masks).
struct Upd { double xm, xn; };

__device__ __forceinline__ double load_score1(const StepParams& p, int b, long long e) {
  long long idx;
  if (p.score_nhwc) {
    const int c = static_cast<int>(e / p.HW);
    const int pix = static_cast<int>(e - static_cast<long long>(c) * p.HW);
    idx = (static_cast<long long>(b) * p.HW + pix) * p.C + c;
  } else {
    idx = static_cast<long long>(b) * p.E + e;
  }
  return p.score_f64 ? static_cast<const double*>(p.score)[idx]
                     : static_cast<double>(static_cast<const float*>(p.score)[idx]);
}

// one element of the half-step: x_mean = fma(coef, score, x - f), x = x_mean + nscale * z (z == 0: no noise)
__device__ __forceinline__ Upd update1(float x, double s, float z, double coef, float nscale, float sa, bool vp_drift,
                                       bool noise) {
  double base = static_cast<double>(x);
  if (vp_drift) base -= static_cast<double>(__fsub_rn(__fmul_rn(sa, x), x));
  Upd u;
  u.xm = fma(coef, s, base);
  u.xn = noise ? u.xm + static_cast<double>(__fmul_rn(nscale, z)) : u.xm;
  return u;
}

// the update of the transposed partner (c, j, i) of element e = (c, i, j) of sample b, or false when the pair is
// not to be symmetrised (channel >= 2, or the partner is conditioned)
__device__ __forceinline__ bool partner_update(const StepParams& p, int b, long long e, unsigned long long stream,
                                               double coef, float nscale, float sa, bool vp_drift, bool noise, Upd& u) {
  const int c = static_cast<int>(e / p.HW);
  if (c >= 2) return false;
  const int pix = static_cast<int>(e - static_cast<long long>(c) * p.HW);
  const int i = pix / p.W, j = pix - i * p.W;
  const long long et = static_cast<long long>(c) * p.HW + static_cast<long long>(j) * p.W + i;
  const long long gt = static_cast<long long>(b) * p.E + et;
  if (p.mask && !p.mask[gt]) return false;
  float z[4] = {0.f, 0.f, 0.f, 0.f};
  if (noise) philox_normal4(p.seed, stream, static_cast<unsigned long long>((p.sample_offset + b) * p.qps + (et >> 2)), z);
  u = update1(p.x[gt], load_score1(p, b, et), z[et & 3], coef, nscale, sa, vp_drift, noise);
  return true;
}

__device__ __forceinline__ void symmetrize4(const StepParams& p, int b, long long e0, uchar4 m, unsigned long long stream,
                                            double coef, float nscale, float sa, bool vp_drift, bool noise,
                                            double (&xn)[4], double (&xm)[4]) {
  const unsigned char mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    Upd t;
    if (mm[i] && partner_update(p, b, e0 + i, stream, coef, nscale, sa, vp_drift, noise, t)) {
      xm[i] = 0.5 * (xm[i] + t.xm);
      xn[i] = 0.5 * (xn[i] + t.xn);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// Direct-load kernels: every operand layout of the API (fp64 / NHWC score, VP drift, probability flow, any
// C*N*N % 4 == 0).  FAST: fp32 NCHW score, VE SDE, noise on.  The mask of a row is loaded one row ahead.
template <bool FAST>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM) predictor_kernel(const StepParams p) {
  const unsigned long long stream = stream_of(p);
  float* const xmean = xmean_of(p);
  int cur_b = -1;
  float G = 0.f, sa = 0.f;
  double coef = 0.0;  // G^2 (fp32, as G[:, None, None, None] ** 2) * drift_scale / sigma
  RowIter it(p);
  uchar4 m = it.valid(p) ? load_mask4(p, it.element(p)) : make_uchar4(1, 1, 1, 1);
  while (!it.done(p)) {
    RowIter nx = it;
    nx.next(p);
    const uchar4 m_next = nx.valid(p) ? load_mask4(p, nx.element(p)) : make_uchar4(1, 1, 1, 1);
    const int b = it.b, ql = it.quad();
    if (b != cur_b) {  // uniform
      cur_b = b;
      G = p.G[b];
      coef = static_cast<double>(__fmul_rn(G, G)) * static_cast<double>(p.drift_scale) * inv_sigma_of(p, b);
      sa = p.sqrt_alpha ? p.sqrt_alpha[b] : 0.f;
    }
    if (ql < p.qps) {
      const long long gi0 = it.element(p);
      if (p.skip_conditioned && all_conditioned(m)) {
        copy_initial4(p, xmean, gi0);
      } else {
        const float4 xv = *reinterpret_cast<const float4*>(p.x + gi0);
        double s[4], xn[4], xm[4];
        load_score4<FAST>(p, b, static_cast<long long>(ql) * 4, s);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        const bool noise = FAST || p.add_noise;
        if (noise)
          philox_normal4(p.seed, stream, static_cast<unsigned long long>((p.sample_offset + b) * p.qps + ql), z);
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // x - f; f == 0 for VE: exactly x + G^2 * score
          const Upd u = update1(xs[i], s[i], z[i], coef, G, sa, !FAST && p.sqrt_alpha != nullptr, noise);
          xm[i] = u.xm;
          xn[i] = u.xn;
        }
        if (!FAST && p.symmetrize)
          symmetrize4(p, b, static_cast<long long>(ql) * 4, m, stream, coef, G, sa, p.sqrt_alpha != nullptr, noise, xn, xm);
        finish4(p, xmean, gi0, m, xn, xm);
      }
    }
    m = m_next;
    it = nx;
  }
}

// block-wide sum of two doubles (result valid in thread 0); `red` is reused, hence the trailing barrier
__device__ __forceinline__ void block_sum2(double& a, double& c, double (*red)[kWarps]) {
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
    for (int w = 0; w < kWarps; ++w) { a += red[0][w]; c += red[1][w]; }
  }
  __syncthreads();
}

// ---- corrector phase 1 bookkeeping: per-row sums.  Every warp reduces the (score^2, z^2) of its 32 quads with
// float shuffles (the addends are floats: 4 squares summed per quad) into acc[row % kSumChunk][warp]; every
// kSumChunk rows, and after the last one, the 16 warp values of each row are summed in double, scaled by
// 1 / sigma^2 and written to the row's partial slot.
struct RowSums {
  float2 (*acc)[kWarps];
  __device__ __forceinline__ void add(int k, float hh, float zz) const {  // all lanes of every warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      hh += __shfl_xor_sync(0xffffffffu, hh, o);
      zz += __shfl_xor_sync(0xffffffffu, zz, o);
    }
    if ((threadIdx.x & 31) == 0) acc[k % kSumChunk][threadIdx.x >> 5] = make_float2(hh, zz);
  }
  // rows [k0, k1) of this block (k1 - k0 <= kSumChunk); all threads
  __device__ __forceinline__ void flush(const StepParams& p, int k0, int k1) const {
    __syncthreads();
    for (int k = k0 + static_cast<int>(threadIdx.x); k < k1; k += kThreads) {
      double sg = 0.0, sn = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        sg += static_cast<double>(acc[k % kSumChunk][w].x);
        sn += static_cast<double>(acc[k % kSumChunk][w].y);
      }
      const int r = static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x);
      const double inv = inv_sigma_of(p, r / p.rps);
      p.partial[2 * static_cast<long long>(r)] = sg * inv * inv;  // sum (h / sigma)^2
      p.partial[2 * static_cast<long long>(r) + 1] = sn;
    }
    __syncthreads();
  }
};

__device__ __forceinline__ float sum_sq4(float a, float b, float c, float d) {
  return __fmaf_rn(d, d, __fmaf_rn(c, c, __fmaf_rn(b, b, __fmul_rn(a, a))));
}

// ---- global-batch step size of a sharded run (opt-in).  Every rank owns a mailbox of
// [2 parities][world] slots {sum ||grad||, sum ||noise||, tag, -} mapped into all peers of the node
// (cudaIpcOpenMemHandle, NVLink).  Block 0 of rank r writes its sums into slot r of EVERY mailbox, payload first,
// then the tag with release semantics at system scope; every block of every rank polls ITS OWN mailbox (local
// memory) until all `world` tags equal this step's tag, and adds the payloads in rank order -- the same double
// sums on every rank and in every block.  Tags are unique per corrector step of a run and per run (tag base from
// the host), consecutive steps alternate slot parity, and a rank publishes step s + 2 only after it has seen every
// peer's step s + 1 tag, which a peer writes after ALL its blocks have left step s (kernel boundary): a slot is
// never overwritten before its readers are done with it.  The wait is bounded:
// a peer that never shows up traps the kernel after ~20 s instead of hanging the device.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ void exchange_sums(const StepParams& p, double& gsum, double& nsum) {  // one thread per block
  const PeerGroup& g = p.peers;
  const long long it = p.iter_ptr ? *p.iter_ptr : 0;
  // index of this corrector step within the run: iteration * n_steps + inner step (stream_mul = n_steps + 1)
  const unsigned long long step =
      static_cast<unsigned long long>(p.stream_mul > 0 ? it * (p.stream_mul - 1) + (p.stream_base - 1) : p.stream_base);
  const unsigned long long tag = static_cast<unsigned long long>(p.tag_base_ptr ? *p.tag_base_ptr : 0) + step + 1;
  const int parity = static_cast<int>(step & 1);
  if (blockIdx.x == 0) {
    for (int r = 0; r < g.world; ++r) {
      unsigned long long* slot = g.box[r] + (static_cast<size_t>(parity) * g.world + g.rank) * 4;
      asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(slot), "d"(gsum) : "memory");
      asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(slot + 1), "d"(nsum) : "memory");
      st_release_sys(slot + 2, tag);
    }
  }
  double gs = 0.0, ns = 0.0;
  const unsigned long long t0 = global_timer_ns();
  for (int r = 0; r < g.world; ++r) {
    const unsigned long long* slot = g.box[g.rank] + (static_cast<size_t>(parity) * g.world + r) * 4;
    while (ld_acquire_sys(slot + 2) != tag) {
      if (global_timer_ns() - t0 > 20000000000ull) __trap();
      __nanosleep(200);
    }
    double a, c;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(a) : "l"(slot) : "memory");
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(c) : "l"(slot + 1) : "memory");
    gs += a;
    ns += c;
  }
  gsum = gs;
  nsum = ns;
}

// step size (before alpha) from the batch-mean norms; every block recomputes it from the row partials: 8 lanes
// per sample, fixed order
__device__ __forceinline__ float batch_step_size(const StepParams& p, double (*red)[kWarps], float* step_sh) {
  double gsum = 0.0;
  float nsum = 0.f;
  const int total = (p.B * 8 + kThreads - 1) / kThreads * kThreads;
  for (int idx = threadIdx.x; idx < total; idx += kThreads) {
    const int b = idx >> 3, part = idx & 7;
    double a = 0.0, c = 0.0;
    if (b < p.B) {
      const double* row = p.partial + 2 * static_cast<long long>(b) * p.rps;
      for (int j = part; j < p.rps; j += 8) {
        a += row[2 * j];
        c += row[2 * j + 1];
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (part == 0 && b < p.B) {
      gsum += sqrt(a);                        // ||grad_b||  (float64)
      nsum += static_cast<float>(sqrt(c));    // ||noise_b|| (float32 tensor in the reference)
    }
  }
  double nsum_d = static_cast<double>(nsum);
  block_sum2(gsum, nsum_d, red);
  if (threadIdx.x == 0) {
    long long nb = p.B;
    if (p.peers.world > 1) {  // mean over the GLOBAL batch: the reference's torch.norm(...).mean() of an unsharded run
      exchange_sums(p, gsum, nsum_d);
      nb = p.peers.global_batch;
    }
    const double grad_norm = gsum / static_cast<double>(nb);
    const float noise_norm = static_cast<float>(nsum_d) / static_cast<float>(nb);
    const float sn = __fmul_rn(p.snr, noise_norm);  // python float * fp32 0-dim tensor -> fp32
    const double r = static_cast<double>(sn) / grad_norm;
    *step_sh = static_cast<float>(r * r * 2.0);     // * alpha (fp32 [B]) demotes the 0-dim double
  }
  __syncthreads();
  return *step_sh;
}

template <bool FAST>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM) corrector_kernel(const StepParams p) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float4 zcache[];  // [cache_rows][kThreads] normals of this block's first rows
  __shared__ double red[2][kWarps];
  __shared__ float2 acc[kSumChunk][kWarps];
  __shared__ float step_sh;
  const unsigned long long stream = stream_of(p);
  const RowSums sums{acc};

  // ---- phase 1: normals (kept), squared norms of the RAW score and of the noise per row
  {
    RowIter it(p);
    int k0 = 0;
    for (; !it.done(p); it.next(p)) {
      const int b = it.b, ql = it.quad();
      float hh = 0.f, zz = 0.f;
      if (ql < p.qps) {
        const long long gi0 = it.element(p);
        double s[4];
        load_score4<FAST>(p, b, static_cast<long long>(ql) * 4, s);
        if (p.prefetch) {  // phase 2's first-touch operands: x (one 128-byte line per 8 threads) and the mask (per warp)
          if ((threadIdx.x & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + gi0));
          if ((threadIdx.x & 31) == 0 && p.mask) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.mask + gi0));
        }
        float z[4];
        philox_normal4(p.seed, stream, static_cast<unsigned long long>((p.sample_offset + b) * p.qps + ql), z);
        if (it.k < p.cache_rows) zcache[it.k * kThreads + threadIdx.x] = make_float4(z[0], z[1], z[2], z[3]);
        if constexpr (FAST) {
          hh = sum_sq4(static_cast<float>(s[0]), static_cast<float>(s[1]), static_cast<float>(s[2]), static_cast<float>(s[3]));
        } else {
          hh = static_cast<float>(s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + s[3] * s[3]);
        }
        zz = sum_sq4(z[0], z[1], z[2], z[3]);
      }
      sums.add(it.k, hh, zz);
      if (it.k - k0 == kSumChunk - 1) { sums.flush(p, k0, it.k + 1); k0 = it.k + 1; }
    }
    sums.flush(p, k0, it.k);
  }
  grid.sync();
  const float step0 = batch_step_size(p, red, &step_sh);

  // ---- phase 2: apply
  float* const xmean = xmean_of(p);
  int cur_b = -1;
  float nscale = 0.f;
  double coef = 0.0;  // step / sigma
  RowIter it(p);
  uchar4 m = it.valid(p) ? load_mask4(p, it.element(p)) : make_uchar4(1, 1, 1, 1);
  while (!it.done(p)) {
    RowIter nx = it;
    nx.next(p);
    const uchar4 m_next = nx.valid(p) ? load_mask4(p, nx.element(p)) : make_uchar4(1, 1, 1, 1);
    const int b = it.b, ql = it.quad();
    if (b != cur_b) {
      cur_b = b;
      const float step = p.alpha ? __fmul_rn(step0, p.alpha[b]) : step0;
      nscale = sqrtf(__fmul_rn(step, 2.f));
      coef = static_cast<double>(step) * inv_sigma_of(p, b);
    }
    if (ql < p.qps) {
      const long long gi0 = it.element(p);
      if (p.skip_conditioned && all_conditioned(m)) {
        copy_initial4(p, xmean, gi0);
      } else {
        const float4 xv = *reinterpret_cast<const float4*>(p.x + gi0);
        double s[4], xn[4], xm[4];
        load_score4<FAST>(p, b, static_cast<long long>(ql) * 4, s);
        float z[4];
        if (it.k < p.cache_rows) {
          const float4 zv = zcache[it.k * kThreads + threadIdx.x];
          z[0] = zv.x; z[1] = zv.y; z[2] = zv.z; z[3] = zv.w;
        } else {
          philox_normal4(p.seed, stream, static_cast<unsigned long long>((p.sample_offset + b) * p.qps + ql), z);
        }
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const Upd u = update1(xs[i], s[i], z[i], coef, nscale, 0.f, false, true);
          xm[i] = u.xm;
          xn[i] = u.xn;
        }
        if (!FAST && p.symmetrize)
          symmetrize4(p, b, static_cast<long long>(ql) * 4, m, stream, coef, nscale, 0.f, false, true, xn, xm);
        finish4(p, xmean, gi0, m, xn, xm);
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
  p.in_place = (a.mask != nullptr && a.conditioned_in_place) ? 1 : 0;
  static const bool no_skip = env_knob_set("T2P_STEP_NOSKIP");  // A/B knob (knob builds only)
  p.skip_conditioned = (a.mask != nullptr && !no_skip) ? 1 : 0;
  p.x_out = a.x_out ? a.x_out : a.x;
  T2P_CHECK((reinterpret_cast<uintptr_t>(p.x_out) & 15) == 0, "x_out must be 16-byte aligned");
  p.symmetrize = a.symmetrize ? 1 : 0;
  p.W = a.W;
  if (p.symmetrize) {
    T2P_CHECK(p.x_out != a.x, "symmetrize needs an out-of-place step (x_out != x): the update reads the transposed state");
    T2P_CHECK(a.W > 0 && static_cast<long long>(a.W) * a.W == a.HW, "symmetrize needs a square map (W * W == HW)");
  }
  p.last_iter_ptr = a.last_iter_ptr;
  if (a.peers && a.peers->world > 1) {
    p.peers = *a.peers;
    p.tag_base_ptr = a.tag_base_ptr;
    T2P_CHECK(p.peers.world <= PeerGroup::kMaxWorld && p.peers.global_batch > 0, "bad peer group");
  } else {
    p.peers.world = 1;
  }
  return p;
}

int num_sms() { return device_sm_count(); }

// resident blocks of `fn` over the whole device (one wave), at `smem` bytes of dynamic shared memory
int resident_blocks(const void* fn, size_t smem) {
  int per_sm = 0;
  T2P_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, smem));
  T2P_CHECK(per_sm > 0, "step kernel does not fit on an SM");
  return std::min(per_sm * num_sms(), kMaxBlocks);
}

constexpr size_t kRowBytes = sizeof(float4) * kThreads;

}  // namespace

void pc_predictor_step(const PcStepArgs& a, cudaStream_t st) {
  StepParams p = to_params(a);
  T2P_CHECK(a.G != nullptr, "predictor needs G");
  const bool fast = !p.score_nhwc && !p.score_f64 && p.sqrt_alpha == nullptr && p.add_noise && !p.symmetrize;
  static int wave_dev[kMaxDevices][2] = {};
  int (&wave)[2] = wave_dev[current_device()];
  if (!wave[fast])
    wave[fast] = resident_blocks(fast ? reinterpret_cast<const void*>(predictor_kernel<true>)
                                      : reinterpret_cast<const void*>(predictor_kernel<false>), 0);
  const int blocks = std::min(wave[fast], p.rows);
  if (fast) predictor_kernel<true><<<blocks, kThreads, 0, st>>>(p);
  else predictor_kernel<false><<<blocks, kThreads, 0, st>>>(p);
  T2P_LAUNCH_CHECK();
}

// doubles of partial-sum workspace the corrector needs: (score^2, noise^2) per row of kThreads quads
long long pc_corrector_workspace_doubles(int B, long long E) { return 2LL * B * cdiv64(E / 4, kThreads); }

void pc_corrector_step(const PcStepArgs& a, cudaStream_t st) {
  StepParams p = to_params(a);
  T2P_CHECK(a.partial != nullptr, "corrector needs the partial-sum workspace");
  static const bool no_cache = env_knob_set("T2P_STEP_NOCACHE");
  static const bool no_prefetch = env_knob_set("T2P_STEP_NOPREFETCH");
  p.prefetch = no_prefetch ? 0 : 1;
  const bool fast = !p.score_nhwc && !p.score_f64 && !p.symmetrize;
  const void* fn = fast ? reinterpret_cast<const void*>(corrector_kernel<true>)
                        : reinterpret_cast<const void*>(corrector_kernel<false>);
  static bool attr_set_dev[kMaxDevices][2] = {};
  bool (&attr_set)[2] = attr_set_dev[current_device()];
  if (!attr_set[fast]) {
    T2P_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaxCacheRows * kRowBytes)));
    T2P_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set[fast] = true;
  }
  // shared-memory cache sized for the rows a block gets when two blocks per SM are resident
  const int want = static_cast<int>(cdiv64(p.rows, static_cast<long long>(kBlocksPerSM) * num_sms()));
  p.cache_rows = no_cache ? 0 : std::min(want, kMaxCacheRows);
  const size_t smem = p.cache_rows * kRowBytes;
  static int wave_dev[kMaxDevices][2][kMaxCacheRows + 1] = {};
  int (&wave)[2][kMaxCacheRows + 1] = wave_dev[current_device()];
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
