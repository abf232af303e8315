// Implicit-GEMM convolution / projection on the sm_100a tensor cores.
//
//   out[M, N] = epilogue( im2col(A)[M, K] * Wt[N, K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// One CTA owns a 128 x BN output tile.  Warp 0 drives TMA: for every filter tap (kh, kw) and every
// 64-channel slice it loads the *shifted* 128-pixel window of the NHWC activation tensor as a 4-D
// box (TMA zero-fills the halo, so padding costs nothing) plus the matching [BN x 64] weight slice.
// Warp 1 issues tcgen05.mma from the 128B-swizzled shared-memory tiles into a TMEM accumulator.
// Warps 2..5 drain TMEM with tcgen05.ld and apply the fused epilogue (bias, per-sample time
// embedding bias, residual add, 1/sqrt(2) rescale, optional GroupNorm statistics) before storing
// NHWC bf16 / fp32.  Replaces the cuDNN / cuBLAS calls behind nn.Conv2d, NIN and nn.Linear on the
// reference hot path (score_sde_pytorch/models/layers.py:82-95,128-137; model/attention.py:161-166).
#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"

namespace t2p {

namespace {

// (A/B, r02_fused_gn_ab.txt: one polling lane + __syncwarp instead of 32 spinning lanes changed nothing measurable
// for the plain kernels and made the fused one slower -- the spinning warps react faster.)
#ifndef T2P_EPI_WARP_WAIT
#define T2P_EPI_WARP_WAIT 0
#endif
#if T2P_EPI_WARP_WAIT
#define T2P_EPI_WAIT(bar, parity) ptx::mbar_wait_warp(bar, parity)
#else
#define T2P_EPI_WAIT(bar, parity) ptx::mbar_wait(bar, parity)
#endif

#ifdef T2P_TIMING_KNOBS
#define T2P_ESTAMP(i) do { if (p.trace && blockIdx.x == 0 && q == 0 && cf == 0 && tl == 0 && lane == 0) p.trace[i] = clock64(); } while (0)
#else
#define T2P_ESTAMP(i) do { } while (0)
#endif
constexpr int BM = 128;      // rows (pixels) per tile == UMMA M
constexpr int BK = 64;       // bf16 elements per 128-byte swizzled row
constexpr int UMMA_K = 16;   // K per tcgen05.mma for 16-bit inputs
constexpr int NUM_THREADS = 192;
constexpr int SMEM_TILE_BUDGET = 192 * 1024;

struct TcParams {
  CUtensorMap tm_a0;
  CUtensorMap tm_a1;
  CUtensorMap tm_w;
  CUtensorMap tm_out;    // channel-major kernel: [M][N] bf16 output, box 32 channels x 32 pixels (TMA store)
  CUtensorMap tm_res;    // channel-major kernel: residual tensor, same pixel box as A
  CUtensorMap tm_ident;  // channel-major kernel: 128 x 128 bf16 identity (residual added by the tensor core)
  CUtensorMap tm_x0;     // channel-major kernel: centre-tap-only sources (folded skip path)
  CUtensorMap tm_x1;
  int M, N;
  int c0, c1;
  int xc0, xc1;
  int taps;          // 1 or 9
  int H, W;          // image geometry for tile -> (b, h, w)
  int mode2d;        // A is a plain [M, K] matrix
  int rows_per_sample;
  const float* bias;
  const float* rowbias;
  const void* residual;
  int res_up;
  float alpha;
  void* out;
  int out_fp32;
  float* stat_part;  // [M tiles][N][2] per-tile column {sum, sum of squares} or null
  int rowbias_ld;
  int out_nchw;
  int n_tiles;       // tiles along N
  int num_tiles;     // m_tiles * n_tiles
  int reverse;       // walk the tiles from the last to the first (see ConvGemmArgs::reverse)
  // fused-GroupNorm halo kernel: the 3x3 sources are RAW activations; act(x * scale + shift) is applied on the way
  // into shared memory.  a0 / a1 global pointers and the per-(sample, channel) affine over the concat [B][c0 + c1]
  const __nv_bfloat16* raw0;
  const __nv_bfloat16* raw1;
  const float* gn_scale;
  const float* gn_shift;
  int hf_debug;      // timing experiments (knob builds): 1 = copy without arithmetic, 2 = no global loads either
  // GroupNorm + SiLU of the OUTPUT applied in the epilogue (see epilogue_role_gn): the consumer's GroupNorm
  const float* gno_gamma;   // [N] or null (mode off)
  const float* gno_beta;
  int gno_cpg;              // channels per group: 4 or 8
  float gno_eps;
  unsigned long long* gno_part;  // [sample][tile part][N / cpg] per-part {sum, sum of squares} of the groups as one word;
                                 // all ones (sentinel) on entry, left so on exit
  int* gno_flags;           // [sample * n_tiles * 4]: warps that left the slot -- zero on entry, left zero on exit
  int gno_parts;            // parts per sample and channel quadrant = pixel tiles per sample x epilogue warps per quadrant
  int gno_slots;            // samples * n_tiles * 4
  long long* trace;         // knob builds, T2P_TRACE_T: clock64 stamps of CTA 0 of a channel-major launch
  const char* pf_ptr;       // optional L2 prefetch (the next GEMM's weights), multiple of 16 bytes
  long long pf_bytes;
  // split-K (channel-major kernel, launches of few tiles): the k-blocks of a tile are shared out over `splits` CTAs
  int splits;               // >= 1
  float* sk_part;           // [tile][split][128 channels][PX pixels] fp32 partial accumulators
  int* sk_ticket;           // [tile][8 epilogue warps] arrivals -- zero on entry, left zero on exit
  int gno_debug;            // timing experiments (knob builds): 1 = fix-up warps do not touch the tile, 2 = nor wait for its stores
};

// The grid asks the L2 for [pf_ptr, pf_ptr + pf_bytes) in 4 KB pieces, one warp per CTA, before the kernel waits for its
// predecessor (the range does not depend on it).
__device__ __forceinline__ void prefetch_l2_range(const TcParams& p, int lane) {
  if (!p.pf_ptr) return;
  constexpr long long kChunk = 4096;
  const long long chunks = (p.pf_bytes + kChunk - 1) / kChunk;
  for (long long c = blockIdx.x + static_cast<long long>(gridDim.x) * lane; c < chunks; c += static_cast<long long>(gridDim.x) * 32) {
    const long long off = c * kChunk;
    const long long n = p.pf_bytes - off < kChunk ? p.pf_bytes - off : kChunk;
    ptx::bulk_prefetch_l2(p.pf_ptr + off, static_cast<uint32_t>(n));
  }
}

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (SMEM_TILE_BUDGET / STAGE_BYTES) < 8 ? (SMEM_TILE_BUDGET / STAGE_BYTES) : 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
  static constexpr int ACC_COLS = BN < 32 ? 32 : BN;   // TMEM columns of one accumulator
  static constexpr int TMEM_COLS = 2 * ACC_COLS;        // two accumulators: epilogue(i) overlaps mainloop(i + 1)
  static constexpr int CH = (BN >= 32) ? 32 : 16;       // accumulator columns per tcgen05.ld
};

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Per-row state of one epilogue thread for the current tile.
struct EpiRow {
  int m;               // global output row
  bool row_ok;
  int sample;
  long long res_row;   // row of the residual tensor (through the x2 nearest-upsample map when res_up)
  int n0;
  bool rb_uniform;
};

// Epilogue of CH accumulator columns [c * CH, (c + 1) * CH) of one row: + bias (+ time-embedding bias)
// + residual, * alpha, store, and (CH == 32) the per-warp column statistics of the stored values.
template <int CH>
__device__ __forceinline__ void epilogue_chunk(const TcParams& p, const EpiRow& er, const uint32_t (&r)[32], int c,
                                               const float* __restrict__ bsm, float2* __restrict__ stat_row, int lane) {
  const int nb = er.n0 + c * CH;
  float v[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(r[i]);
  const bool full = (nb + CH <= p.N) && ((p.N & 7) == 0);
  if (er.row_ok) {
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] += bsm[c * CH + i];
    if (p.rowbias && !er.rb_uniform) {
      const float* rb = p.rowbias + static_cast<long long>(er.sample) * p.rowbias_ld + nb;
#pragma unroll
      for (int i = 0; i < CH; ++i)
        if (nb + i < p.N) v[i] += __ldg(rb + i);
    }
    if (p.residual) {
      if (p.out_fp32) {
        const float* rs = static_cast<const float*>(p.residual) + er.res_row * p.N + nb;
#pragma unroll
        for (int i = 0; i < CH; ++i)
          if (nb + i < p.N) v[i] += rs[i];
      } else if (full) {
        const uint4* rs = reinterpret_cast<const uint4*>(
            static_cast<const __nv_bfloat16*>(p.residual) + er.res_row * p.N + nb);
        uint4 t[CH / 8];
#pragma unroll
        for (int i = 0; i < CH / 8; ++i) t[i] = rs[i];
#pragma unroll
        for (int i = 0; i < CH / 8; ++i) {
          v[8 * i + 0] += bf16_lo(t[i].x); v[8 * i + 1] += bf16_hi(t[i].x);
          v[8 * i + 2] += bf16_lo(t[i].y); v[8 * i + 3] += bf16_hi(t[i].y);
          v[8 * i + 4] += bf16_lo(t[i].z); v[8 * i + 5] += bf16_hi(t[i].z);
          v[8 * i + 6] += bf16_lo(t[i].w); v[8 * i + 7] += bf16_hi(t[i].w);
        }
      } else {
        const __nv_bfloat16* rs = static_cast<const __nv_bfloat16*>(p.residual) + er.res_row * p.N + nb;
#pragma unroll
        for (int i = 0; i < CH; ++i)
          if (nb + i < p.N) v[i] += __bfloat162float(rs[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] *= p.alpha;
    if (p.out_fp32) {
      float* o = static_cast<float*>(p.out) + static_cast<long long>(er.m) * p.N + nb;
      if (p.out_nchw) {
        const int hw = p.H * p.W;
        const int b = er.m / hw;
        const int pix = er.m - b * hw;
#pragma unroll
        for (int i = 0; i < CH; ++i)
          if (nb + i < p.N) static_cast<float*>(p.out)[(static_cast<long long>(b) * p.N + nb + i) * hw + pix] = v[i];
      } else if (full) {
#pragma unroll
        for (int i = 0; i < CH / 4; ++i)
          reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < CH; ++i)
          if (nb + i < p.N) o[i] = v[i];
      }
    } else {
      __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(er.m) * p.N + nb;
      if (full) {
#pragma unroll
        for (int i = 0; i < CH / 8; ++i) {
          uint4 t;
          t.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]);
          t.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
          t.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
          t.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
          reinterpret_cast<uint4*>(o)[i] = t;
        }
      } else {
#pragma unroll
        for (int i = 0; i < CH; ++i)
          if (nb + i < p.N) o[i] = __float2bfloat16(v[i]);
      }
      // statistics are taken over the values as stored (bf16-rounded), matching what the
      // GroupNorm consumer will read back
#pragma unroll
      for (int i = 0; i < CH; ++i) v[i] = __bfloat162float(__float2bfloat16(v[i]));
    }
  }
  if constexpr (CH == 32) {
    if (p.stat_part) {
      // column sums over this warp's 32 rows by a transposing butterfly: after the 5 exchange steps lane l
      // holds the totals of column nb + l (31 shuffles per quantity instead of 32 x 5)
      float sq[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (!er.row_ok) v[i] = 0.f;
        sq[i] = v[i] * v[i];
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
          const float send_v = upper ? v[i] : v[i + off];
          const float keep_v = upper ? v[i + off] : v[i];
          const float send_q = upper ? sq[i] : sq[i + off];
          const float keep_q = upper ? sq[i + off] : sq[i];
          v[i] = keep_v + __shfl_xor_sync(0xffffffffu, send_v, off);
          sq[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, off);
        }
      }
      stat_row[c * 32 + lane] = make_float2(v[0], sq[0]);
    }
  }
}

// Persistent kernel: grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...
// (N-tile fastest, so that CTAs running side by side share the activation tile in L2).  The TMA ring runs
// ahead across tile boundaries and the accumulator is double-buffered in TMEM, so the epilogue of tile i
// overlaps the main loop of tile i + 1.
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_gemm_tc_kernel(const __grid_constant__ TcParams p) {
  using C = Cfg<BN>;
  constexpr int CH = C::CH;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[C::STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];   // accumulator ready for the epilogue
  __shared__ __align__(8) uint64_t tempty_bar[2];  // accumulator drained by the epilogue
  __shared__ uint32_t tmem_base_slot;
  __shared__ float2 stat_sm[4][BN >= 32 ? BN : 32];
  __shared__ float bias_sm[2][BN >= 32 ? BN : 32];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tiles = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;

  const int ctot = p.c0 + p.c1;
  const int chunks_per_tap = ctot / BK;
  const int num_kb = p.taps * chunks_per_tap;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&tfull_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&tempty_bar[s]), 4);  // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (warp == 2) prefetch_l2_range(p, lane);
  pdl_wait();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::prefetch_tmap(&p.tm_a0);
      ptx::prefetch_tmap(&p.tm_w);
      if (p.c1 > 0) ptx::prefetch_tmap(&p.tm_a1);
      const int pad = (p.taps == 9) ? 1 : 0;
      const int hw = p.H * p.W;
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int tt = p.reverse ? p.num_tiles - 1 - t : t;
        const int mt = tt / p.n_tiles;
        const int m0 = mt * BM;
        const int n0 = (tt - mt * p.n_tiles) * BN;
        int b0 = 0, h0 = 0, w0 = m0;
        if (!p.mode2d) {
          b0 = m0 / hw;
          const int rem = m0 - b0 * hw;
          h0 = rem / p.W;
          w0 = rem - h0 * p.W;
        }
        int kb = 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int kh = (p.taps == 9) ? tap / 3 : 0;
          const int kw = (p.taps == 9) ? tap - kh * 3 : 0;
          for (int cc = 0; cc < chunks_per_tap; ++cc, ++kb) {
            ptx::mbar_wait(ptx::smem_u32(&empty_bar[s]), ph ^ 1);
            const uint32_t fb = ptx::smem_u32(&full_bar[s]);
            ptx::mbar_arrive_expect_tx(fb, C::STAGE_BYTES);
            const uint32_t sa = tiles + s * C::STAGE_BYTES;
            const uint32_t sb = sa + C::A_BYTES;
            const int ch = cc * BK;
            if (ch < p.c0)
              ptx::tma_load_4d(sa, &p.tm_a0, fb, ch, w0 + kw - pad, h0 + kh - pad, b0);
            else
              ptx::tma_load_4d(sa, &p.tm_a1, fb, ch - p.c0, w0 + kw - pad, h0 + kh - pad, b0);
            ptx::tma_load_4d(sb, &p.tm_w, fb, kb * BK, n0, 0, 0);  // all maps are encoded rank-4
            if (++s == C::STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BN);
      int s = 0;
      uint32_t ph = 0;
      uint32_t tl = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tl) {
        const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
        ptx::mbar_wait(ptx::smem_u32(&tempty_bar[as]), aph ^ 1);  // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * C::ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[s]), ph);
          ptx::tc_fence_after();
          const uint32_t sa = tiles + s * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t da = ptx::umma_desc_k_sw128(sa);
          const uint64_t db = ptx::umma_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advancing 16 bf16 = 32 B inside the swizzle atom: +2 in the (addr >> 4) field
            ptx::umma_bf16(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          ptx::umma_commit(ptx::smem_u32(&empty_bar[s]));  // frees the smem slot when the MMAs retire
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        ptx::umma_commit(ptx::smem_u32(&tfull_bar[as]));  // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int e = (warp - 2) * 32 + lane;
    const int hw = p.H * p.W;
    uint32_t tl = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tl) {
      const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
      const int tt = p.reverse ? p.num_tiles - 1 - t : t;
        const int mt = tt / p.n_tiles;
      const int m0 = mt * BM;
      EpiRow er;
      er.n0 = (tt - mt * p.n_tiles) * BN;
      er.m = m0 + q * 32 + lane;
      er.row_ok = er.m < p.M;
      er.sample = (p.rows_per_sample > 0) ? (er.m / p.rows_per_sample) : 0;
      // While the main loop runs, stage the per-column bias (+ the per-sample time-embedding bias when the
      // whole tile belongs to one sample) in shared memory: the epilogue then never waits on global loads.
      er.rb_uniform = p.rowbias != nullptr && (p.rows_per_sample % BM) == 0;
      float* bsm = bias_sm[as];
      {
        const float* rb0 =
            er.rb_uniform ? p.rowbias + static_cast<long long>(m0 / p.rows_per_sample) * p.rowbias_ld : nullptr;
        for (int ch = e; ch < BN; ch += 128) {
          const int n = er.n0 + ch;
          float bv = 0.f;
          if (n < p.N) {
            if (p.bias) bv = __ldg(p.bias + n);
            if (rb0) bv += __ldg(rb0 + n);
          }
          bsm[ch] = bv;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      er.res_row = er.m;
      if (p.res_up && er.row_ok) {
        const int b = er.m / hw;
        const int rem = er.m - b * hw;
        const int h = rem / p.W;
        const int w = rem - h * p.W;
        er.res_row = (static_cast<long long>(b) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
      }
      T2P_EPI_WAIT(ptx::smem_u32(&tfull_bar[as]), aph);
      ptx::tc_fence_after();
      const uint32_t tbase = tmem_base + as * C::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
      auto release_acc = [&]() {
        // every tcgen05.ld of this warp has completed: hand the accumulator back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tempty_bar[as]));
      };
      if constexpr (CH == 16) {
        uint32_t r[32];
        {
          uint32_t r16[16];
          ptx::tmem_ld_32x16(tbase, r16);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = r16[i];
        }
        release_acc();
        epilogue_chunk<16>(p, er, r, 0, bsm, stat_sm[q], lane);
      } else {
        constexpr int NCH = BN / 32;
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_32x32(tbase, r0);
#pragma unroll 1
        for (int c = 0; c < NCH; c += 2) {
          ptx::tmem_ld_wait();
          if (c + 1 < NCH) ptx::tmem_ld_32x32(tbase + (c + 1) * 32, r1);
          else release_acc();
          epilogue_chunk<32>(p, er, r0, c, bsm, stat_sm[q], lane);
          if (c + 1 < NCH) {
            ptx::tmem_ld_wait();
            if (c + 2 < NCH) ptx::tmem_ld_32x32(tbase + (c + 2) * 32, r0);
            else release_acc();
            epilogue_chunk<32>(p, er, r1, c + 1, bsm, stat_sm[q], lane);
          }
        }
        if (p.stat_part) {
          asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
          for (int ch = e; ch < BN; ch += 128) {
            if (er.n0 + ch < p.N) {
              // fixed order -> run-to-run deterministic statistics
              const float2 a0 = stat_sm[0][ch], a1 = stat_sm[1][ch], a2 = stat_sm[2][ch], a3 = stat_sm[3][ch];
              float* dst = p.stat_part + (static_cast<long long>(mt) * p.N + er.n0 + ch) * 2;
              dst[0] = (a0.x + a1.x) + (a2.x + a3.x);
              dst[1] = (a0.y + a1.y) + (a2.y + a3.y);
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// =====================================================================================================
// Channel-major variant:  out^T[N, M] = Wt[N, K] * im2col(A)[M, K]^T.
//
// The 128 UMMA rows (TMEM lanes) are OUTPUT CHANNELS and the UMMA N dimension is a tile of PX pixels, i.e. the
// weight slice is the "A" operand and the shifted pixel window the "B" operand.  Against the pixel-major
// kernel above this (i) halves the weight traffic per FLOP for the 128-channel layers (one 16 KB weight slice
// feeds 256 pixels instead of 128: the L2 -> SMEM operand stream is what bounds these layers), (ii) makes the
// epilogue channel-per-thread: bias and time-embedding bias are per-thread scalars, GroupNorm statistics are
// per-thread running sums (no shuffles, no shared memory, no barriers), and every store / residual load of a
// warp is one contiguous 64-byte run of the NHWC tensor.
template <int PX>
struct CfgT {
  static constexpr int W_BYTES = 128 * BK * 2;
  static constexpr int P_BYTES = PX * BK * 2;
  static constexpr int STAGE_BYTES = W_BYTES + P_BYTES;
  static constexpr int STAGES = (SMEM_TILE_BUDGET / STAGE_BYTES) < 8 ? (SMEM_TILE_BUDGET / STAGE_BYTES) : 8;
  static constexpr int EPI_WARPS = 8;  // two per TMEM lane quadrant: each takes half of the tile's pixel columns
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int OUT_BYTES = EPI_WARPS * 2 * 32 * 32 * 2;  // per epilogue warp: 2 buffers of 32 px x 32 ch bf16
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * PX;
};

// ----------------------------------------------------------------------------------------------------------
// GroupNorm + SiLU of the convolution's OUTPUT by the convolution kernel itself (ResnetBlockBigGANpp: h = act(
// GroupNorm_1(Conv_0(.) + temb)), layers.py:314-318): the separate statistics / gn_finalize / gn_apply passes over the
// tensor disappear.  The statistics of a sample need ALL its pixel tiles, which other CTAs hold:
//   publish  every epilogue warp reduces its pixels x 32 channels to per-group {sum, sum of squares} and lane 0 stores
//            each pair as ONE 64-bit word into gno_part[sample][part][group]
//   fold     every warp that needs the statistics of a (sample, n tile, quadrant) slot reads all parts of its groups and
//            adds them itself, in double and in the same fixed order in every warp and every run (deterministic;
//            bit-identical across the warps of a sample).  The buffer is its own flag: a word holds a sentinel (all
//            ones: a NaN no arithmetic produces) until written, and a lane re-reads the words that are still the
//            sentinel (bounded wait, back-off).  No counter, no fence, no release / acquire pair.
//   leave    one counter word per slot, off the critical path: the last reader writes the sentinel back and zeroes the
//            counter -- gno_part and gno_flags enter and leave every launch in the same state.
// Two ways to use them:
//  * epilogue_role_gn (channel-major kernel, 64 x 64 images and below): two passes over the accumulator (tcgen05.ld does
//    not consume it) -- statistics, publish, fold, then y = silu(v * scale + shift) -> bf16 -> TMA store.  The raw tensor
//    never exists, but the accumulator is held while the warp waits for the other CTAs of its sample.
//  * deferred (halo kernel, 128-pixel-wide images): the plain epilogue stores the RAW bf16 tile and publishes, the
//    accumulator is released at once, and four extra warps (fixup_role) normalise each tile IN PLACE a tile or more
//    later, out of L2.  Measured reason (profiles/r02_gn_out_epilogue_timeline.txt): with the two-pass form every wave
//    of tiles waits for the slowest of the 64 CTAs of a sample -- tile period 1.40 x the mean MMA time whatever the
//    protocol costs -- because one spare accumulator is all the slack TMEM has.
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long pack_stat(float sum, float sq) {
  return (static_cast<unsigned long long>(__float_as_uint(sq)) << 32) | __float_as_uint(sum);
}
constexpr unsigned long long kStatSentinel = ~0ull;
__device__ __forceinline__ float silu_tanh_f(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// geometry of one (sample, n tile, quadrant) slot of the exchange
struct GnoSlot {
  int cpg, gpw, ngroups;          // channels per group (4 / 8 / 16 / 32), groups per warp (32 channels), groups per sample
  int slot;                       // (sample * n_tiles + nt) * 4 + q
  unsigned long long* part;       // gno_part + (sample * parts * ngroups + first group of the warp): + part * ngroups + g
};
__device__ __forceinline__ GnoSlot gno_slot(const TcParams& p, int sample, int nt, int q) {
  GnoSlot s;
  s.cpg = p.gno_cpg;
  s.gpw = 32 / s.cpg;
  s.ngroups = p.N / s.cpg;
  s.slot = (sample * p.n_tiles + nt) * 4 + q;
  s.part = p.gno_part + static_cast<long long>(sample) * p.gno_parts * s.ngroups + (nt * 128 + q * 32) / s.cpg;
  return s;
}

// publish: ssum / ssq = this lane's sums over its pixels of the channels (lane >> 2) + 8 k, k = 0..3 of the warp's 32.
// Channel totals over the four lanes that share a channel, then the channels of a group: groups of 4 channels = lanes
// with equal (lane >> 4) of one slot k, groups of 8 = all lanes of a slot, groups of 16 / 32 = two / four slots.
__device__ __forceinline__ void gno_publish(const GnoSlot& gs, int part, float (&ssum)[4], float (&ssq)[4], int lane) {
  float gsum[8], gsq[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 1; o <= 8; o <<= 1) {
      ssum[k] += __shfl_xor_sync(0xffffffffu, ssum[k], o);
      ssq[k] += __shfl_xor_sync(0xffffffffu, ssq[k], o);
    }
    // lanes 0..15 hold the channels (lane >> 2) < 4 of slot k, lanes 16..31 the channels 4..7
    const float os = __shfl_xor_sync(0xffffffffu, ssum[k], 16), oq = __shfl_xor_sync(0xffffffffu, ssq[k], 16);
    gsum[2 * k] = ssum[k]; gsq[2 * k] = ssq[k];        // (as seen from lane 0: group 2k of 4 channels ...
    gsum[2 * k + 1] = os;  gsq[2 * k + 1] = oq;        //  ... and group 2k + 1)
  }
  if (lane != 0) return;
  unsigned long long* const dst = gs.part + static_cast<long long>(part) * gs.ngroups;
  if (gs.cpg == 4) {
#pragma unroll
    for (int g = 0; g < 8; ++g) st_relaxed_gpu_u64(dst + g, pack_stat(gsum[g], gsq[g]));
  } else if (gs.cpg == 8) {
#pragma unroll
    for (int g = 0; g < 4; ++g) st_relaxed_gpu_u64(dst + g, pack_stat(gsum[2 * g] + gsum[2 * g + 1], gsq[2 * g] + gsq[2 * g + 1]));
  } else if (gs.cpg == 16) {
#pragma unroll
    for (int g = 0; g < 2; ++g)
      st_relaxed_gpu_u64(dst + g, pack_stat((gsum[4 * g] + gsum[4 * g + 1]) + (gsum[4 * g + 2] + gsum[4 * g + 3]),
                                            (gsq[4 * g] + gsq[4 * g + 1]) + (gsq[4 * g + 2] + gsq[4 * g + 3])));
  } else {
    st_relaxed_gpu_u64(dst, pack_stat(((gsum[0] + gsum[1]) + (gsum[2] + gsum[3])) + ((gsum[4] + gsum[5]) + (gsum[6] + gsum[7])),
                                      ((gsq[0] + gsq[1]) + (gsq[2] + gsq[3])) + ((gsq[4] + gsq[5]) + (gsq[6] + gsq[7]))));
  }
}

// fold (blocking): lane l adds the parts r = l / gpw, l / gpw + 32 / gpw, ... of group l % gpw (coalesced: the groups of
// a part are adjacent), sixteen loads in flight, re-reading what has not arrived yet; a butterfly over the lanes of equal
// group completes the sum.  On return lane g (g < gpw) -- and every lane congruent to it -- holds mean / rstd of group g.
__device__ __forceinline__ void gno_fold(const TcParams& p, const GnoSlot& gs, int lane, float& mean_g, float& rstd_g) {
  const int g = lane & (gs.gpw - 1);
  const int rstep = gs.cpg;  // 32 / gpw
  const unsigned long long* const src = gs.part + g;
  double sum = 0.0, sq = 0.0;
  for (int r = lane / gs.gpw; r < p.gno_parts; r += 16 * rstep) {
    unsigned long long v[16];
    unsigned pending = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (r + i * rstep < p.gno_parts) pending |= 1u << i;
    const unsigned mine = pending;
    long long t0 = 0;
    unsigned spins = 0;
    while (true) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (pending & (1u << i)) v[i] = ld_relaxed_gpu_u64(src + static_cast<long long>(r + i * rstep) * gs.ngroups);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if ((pending & (1u << i)) && static_cast<unsigned>(v[i] >> 32) != 0xffffffffu) pending &= ~(1u << i);
      if (!pending) break;
      if (spins == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000LL) __trap();
      if (++spins > 2) __nanosleep(spins > 16 ? 400 : 100);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (mine & (1u << i)) {
        sum += static_cast<double>(__uint_as_float(static_cast<unsigned>(v[i])));
        sq += static_cast<double>(__uint_as_float(static_cast<unsigned>(v[i] >> 32)));
      }
    }
  }
  __syncwarp();
  for (int o = gs.gpw; o < 32; o <<= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  const double inv_cnt = 1.0 / (static_cast<double>(p.rows_per_sample) * gs.cpg);
  const double mean = sum * inv_cnt;
  const float var = static_cast<float>(fmax(sq * inv_cnt - mean * mean, 0.0));
  mean_g = static_cast<float>(mean);
  rstd_g = rsqrtf(var + p.gno_eps);
}

// leave: `ticket` = what lane 0's atomicAdd(&gno_flags[slot], 1) returned (asked for right after the fold, looked at
// here); `readers` warps read a slot.  The last one hands it back as it was found.
__device__ __forceinline__ void gno_leave(const TcParams& p, const GnoSlot& gs, int ticket, int readers, int lane) {
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  if (ticket != readers - 1) return;
  for (int i = lane; i < p.gno_parts * gs.gpw; i += 32)
    st_relaxed_gpu_u64(gs.part + static_cast<long long>(i / gs.gpw) * gs.ngroups + (i & (gs.gpw - 1)), kStatSentinel);
  if (lane == 0) p.gno_flags[gs.slot] = 0;
}

// Epilogue of the channel-major kernels.  The accumulator is read in the fragment layout of tcgen05.ld.16x256b:
// thread t of a warp owns channels nw0 + t / 4 + 8 k (k = 0..3) and, in every group of 8 pixel columns, the two
// adjacent pixels 2 (t % 4) + {0, 1} -- one packed bf16x2 conversion per pair, and a transposing stmatrix lays
// the pairs down as [pixel][channel] rows for the TMA store (4 stmatrix per 32 x 32 chunk).  The code is
// specialised on the two launch-uniform options that would otherwise cost issue slots in every chunk:
//   RBVAR  the per-sample bias changes inside a tile (tiles wider than a sample: the 8 x 8 and 4 x 4 levels)
//   STATS  GroupNorm statistics wanted
struct EpiQ {
  int nw0;        // first channel of this warp
  int m0;         // first pixel of the tile
  bool ok[4];     // channel < N
  float bch[4];   // (bias[n] + rowbias[sample][n] when the tile lies inside one sample) * alpha
  float ssum[4], ssq[4];
};

template <bool RBVAR>
__device__ __forceinline__ void epi_begin_tile(const TcParams& p, EpiQ& eq, int lane) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int n = eq.nw0 + (lane >> 2) + 8 * k;
    eq.ok[k] = n < p.N;
    float b = 0.f;
    eq.ssum[k] = 0.f;
    eq.ssq[k] = 0.f;
    if (eq.ok[k]) {
      if (p.bias) b = __ldg(p.bias + n);
      if (!RBVAR && p.rowbias)
        b += __ldg(p.rowbias + static_cast<long long>(eq.m0 / p.rows_per_sample) * p.rowbias_ld + n);
    }
    eq.bch[k] = b * p.alpha;
  }
}

// 32 channels x 32 pixels [mb, mb + 32): out = acc * alpha + bias * alpha -> bf16, statistics over the values as
// stored, then the chunk goes to `buf` ([32 px][32 ch] bf16, 64-byte rows, 64-byte swizzle = the output tensor
// map's) -- conflict-free: the 8 rows of each 8 x 8 matrix land in 8 different 16-byte slots of a 128-byte line.
template <bool RBVAR, bool STATS>
__device__ __forceinline__ void epilogue_chunk_q(const TcParams& p, EpiQ& eq, const uint32_t (&lo)[16],
                                                 const uint32_t (&hi)[16], int mb, uint32_t buf, int lane) {
  uint32_t pk[4][4];  // [channel slot k][column group n]
  const int px_t = 2 * (lane & 3);
  const bool full = mb + 32 <= p.M;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const float v[4][2] = {{__uint_as_float(lo[4 * n]), __uint_as_float(lo[4 * n + 1])},
                           {__uint_as_float(lo[4 * n + 2]), __uint_as_float(lo[4 * n + 3])},
                           {__uint_as_float(hi[4 * n]), __uint_as_float(hi[4 * n + 1])},
                           {__uint_as_float(hi[4 * n + 2]), __uint_as_float(hi[4 * n + 3])}};
    const int px = mb + 8 * n + px_t;  // this thread's pixel pair (px, px + 1)
    const bool in_m = full || px < p.M;  // M is even whenever it matters: a pair is inside or outside as a whole
    const float *rb_row0 = nullptr, *rb_row1 = nullptr;
    if (RBVAR) {  // the two pixels may belong to different samples (odd pixel counts per sample)
      rb_row0 = p.rowbias + static_cast<long long>(min(px, p.M - 1) / p.rows_per_sample) * p.rowbias_ld + eq.nw0 + (lane >> 2);
      rb_row1 = p.rowbias + static_cast<long long>(min(px + 1, p.M - 1) / p.rows_per_sample) * p.rowbias_ld + eq.nw0 + (lane >> 2);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float b = eq.bch[k], b1 = eq.bch[k];
      if (RBVAR && eq.ok[k]) {
        b = fmaf(__ldg(rb_row0 + 8 * k), p.alpha, b);
        b1 = fmaf(__ldg(rb_row1 + 8 * k), p.alpha, b1);
      }
      pk[k][n] = ptx::pack_bf16x2(fmaf(v[k][0], p.alpha, b), fmaf(v[k][1], p.alpha, RBVAR ? b1 : b));
      if (STATS && in_m) {  // over the values as stored
        const float f0 = __uint_as_float(pk[k][n] << 16), f1 = __uint_as_float(pk[k][n] & 0xffff0000u);
        eq.ssum[k] += f0 + f1;
        eq.ssq[k] = fmaf(f0, f0, fmaf(f1, f1, eq.ssq[k]));
      }
    }
  }
  const int mt = lane >> 3, row = lane & 7;
  const uint32_t rowaddr = buf + (8 * (mt >> 1) + row) * 64;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf)
#pragma unroll
    for (int ph = 0; ph < 2; ++ph)
      ptx::stmatrix_x4_trans(rowaddr + ph * (16 * 64) + ((((2 * hf + (mt & 1)) ^ (row >> 1)) & 3) << 4), pk[2 * hf][2 * ph],
                             pk[2 * hf + 1][2 * ph], pk[2 * hf][2 * ph + 1], pk[2 * hf + 1][2 * ph + 1]);
}

// per-channel statistics of the tile: the four threads sharing a channel hold disjoint pixel subsets
__device__ __forceinline__ void epi_write_stats(const TcParams& p, EpiQ& eq, long long stat_tile, int lane) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float s = eq.ssum[k], q = eq.ssq[k];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    if ((lane & 3) == 0 && eq.ok[k])
      *reinterpret_cast<float2*>(p.stat_part + (stat_tile * p.N + eq.nw0 + (lane >> 2) + 8 * k) * 2) = make_float2(s, q);
  }
}

// One epilogue warp's loop over the CTA's tiles.  The warp drains TMEM lane quadrant q, pixel chunks
// [cf, cf + NCH) of each PX-pixel accumulator (two accumulators, ping-pong), through its two staging buffers at
// `obuf`; statistics tile index = stat_mul * pixel_tile + stat_add.
// GND (deferred GroupNorm of the output, see gno_publish): STATS without RBVAR; the statistics go to the exchange buffer
// instead of stat_part, and the warp counts in the shared-memory word `ready_smem` the tiles whose stores have been
// PERFORMED -- the fix-up warp of this quadrant (fixup_role) rewrites them in place.
template <int PX, int NCH, bool RBVAR, bool STATS, bool GND = false>
__device__ __forceinline__ void epilogue_role(const TcParams& p, uint32_t tmem_base, uint64_t* tfull_bar,
                                              uint64_t* tempty_bar, uint32_t obuf, int q, int cf, int stat_mul,
                                              int stat_add, int lane, uint32_t ready_smem = 0, int acc_owner = -1) {
  // acc_owner >= 0 (CTA-pair kernel): the accumulator is handed back on the mbarrier of that CTA of the cluster
  EpiQ eq;
  uint32_t tl = 0;
  uint32_t nstore = 0;
  // `count` tiles have reached global memory: called after the stores of tile `count` have been ISSUED, it waits for
  // everything but those (the previous tile's stores had a whole tile time to complete: waiting for the tile just stored
  // instead cost the epilogue its overlap with the next accumulator)
  auto publish_stored = [&](uint32_t count, bool last) {
    if (lane == 0) {
      if (last) ptx::tma_store_wait<0>();
      else ptx::tma_store_wait<NCH>();
      ptx::fence_proxy_async_global();
      ptx::st_release_cta_shared(ready_smem, static_cast<int>(count));
    }
  };
  for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tl) {
    const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
    const int tt = p.reverse ? p.num_tiles - 1 - t : t;
    const int pt = tt / p.n_tiles;
    eq.nw0 = (tt - pt * p.n_tiles) * 128 + q * 32;  // first channel of this warp
    eq.m0 = pt * PX;
    epi_begin_tile<RBVAR>(p, eq, lane);
    T2P_EPI_WAIT(ptx::smem_u32(&tfull_bar[as]), aph);
    ptx::tc_fence_after();
    T2P_ESTAMP(6);
    const uint32_t tbase = tmem_base + as * PX + (static_cast<uint32_t>(q * 32) << 16);
    auto release_acc = [&]() {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (acc_owner < 0) ptx::mbar_arrive(ptx::smem_u32(&tempty_bar[as]));
        else ptx::mbar_arrive_cluster(ptx::smem_u32(&tempty_bar[as]), static_cast<uint32_t>(acc_owner));
      }
    };
    auto load_chunk = [&](int c, uint32_t (&lo)[16], uint32_t (&hi)[16]) {
      ptx::tmem_ld_16x256_x4(tbase + c * 32, lo);
      ptx::tmem_ld_16x256_x4(tbase + (16u << 16) + c * 32, hi);
    };
    // one chunk -> this warp's staging buffer -> one TMA store.  TMA clips rows >= M and channels >= N, so
    // ragged edges need no masks here.
    auto emit_chunk = [&](const uint32_t (&lo)[16], const uint32_t (&hi)[16], int c) {
      const uint32_t buf = obuf + (nstore & 1) * (32 * 32 * 2);
      // the store issued two chunks ago read this buffer: at most one (the previous chunk's) may be pending
      if (lane == 0) ptx::tma_store_wait_read<1>();
      __syncwarp();
      epilogue_chunk_q<RBVAR, STATS>(p, eq, lo, hi, eq.m0 + c * 32, buf, lane);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && eq.nw0 < p.N && eq.m0 + c * 32 < p.M) {
        ptx::tma_store_4d(&p.tm_out, buf, eq.nw0, eq.m0 + c * 32, 0, 0);
        ptx::tma_store_commit();
      }
      ++nstore;
    };
    uint32_t lo0[16], hi0[16], lo1[16], hi1[16];
    load_chunk(cf, lo0, hi0);
#pragma unroll 1
    for (int i = 0; i < NCH; i += 2) {
      ptx::tmem_ld_wait();
      if (i + 1 < NCH) load_chunk(cf + i + 1, lo1, hi1);
      else release_acc();
      emit_chunk(lo0, hi0, cf + i);
      if (i + 1 < NCH) {
        ptx::tmem_ld_wait();
        if (i + 2 < NCH) load_chunk(cf + i + 2, lo0, hi0);
        else release_acc();
        emit_chunk(lo1, hi1, cf + i + 1);
      }
    }
    if (GND) {
      const int sample = eq.m0 / p.rows_per_sample;
      const GnoSlot gs = gno_slot(p, sample, tt - pt * p.n_tiles, q);
      gno_publish(gs, (pt - sample * (p.rows_per_sample / PX)) * stat_mul + stat_add, eq.ssum, eq.ssq, lane);
      publish_stored(tl, false);
    } else if (STATS) {
      epi_write_stats(p, eq, static_cast<long long>(stat_mul) * pt + stat_add, lane);
    }
    T2P_ESTAMP(7);
  }
  if (GND) publish_stored(tl, true);
  if (lane == 0) ptx::tma_store_wait_read<0>();  // smem must stay valid until the last store has read it
}

// ----------------------------------------------------------------------------------------------------------
// Split-K epilogue (channel-major kernel).  The 3x3 convolutions at 16 x 16 and below have long K (2304-4608) and few
// tiles: one CTA per tile walks 36-72 k-blocks at the latency of its own TMA ring (0.5 us per k-block, 19-30 us per
// launch whatever the batch) while most SMs idle.  With `splits` > 1 a work item is (tile, K range) and consecutive CTAs
// take the ranges of one tile.  Every epilogue warp dumps its 32 channels x pixels of the partial accumulator as fp32
// ([128][PX] per item, 32-byte runs per four lanes), fences and takes a ticket for its (tile, warp) slot; the warp that
// draws the last ticket adds the partials in split order 0, 1, ... -- its own included, read back like the others, so
// the sum does not depend on who came last (deterministic) -- and runs the ordinary epilogue (bias, per-sample bias,
// statistics, bf16, TMA store) on the sums.  Tickets are handed back zeroed.
template <int PX, int NCH, bool RBVAR, bool STATS>
__device__ __forceinline__ void epilogue_role_splitk(const TcParams& p, uint32_t tmem_base, uint64_t* tfull_bar,
                                                     uint64_t* tempty_bar, uint32_t obuf, int q, int cf, int stat_mul,
                                                     int stat_add, int lane) {
  EpiQ eq;
  uint32_t tl = 0;
  uint32_t nstore = 0;
  const int S = p.splits;
  const int num_items = p.num_tiles * S;
  const int wslot = q + 4 * stat_add;  // (stat_add = which half of the pixel columns this warp drains)
  // this thread's element (channel slot k, column group n, pixel pair) of a 32-pixel chunk inside a [128][PX] partial
  const int frag_off = (q * 32 + (lane >> 2)) * PX + 2 * (lane & 3);
  for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++tl) {
    const int t = it / S;
    const int split = it - t * S;
    const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
    const int tt = p.reverse ? p.num_tiles - 1 - t : t;
    const int pt = tt / p.n_tiles;
    eq.nw0 = (tt - pt * p.n_tiles) * 128 + q * 32;
    eq.m0 = pt * PX;
    float* const tile_part = p.sk_part + static_cast<long long>(tt) * S * (128 * PX);
    float* const mine = tile_part + static_cast<long long>(split) * (128 * PX) + frag_off;
    T2P_EPI_WAIT(ptx::smem_u32(&tfull_bar[as]), aph);
    ptx::tc_fence_after();
    const uint32_t tbase = tmem_base + as * PX + (static_cast<uint32_t>(q * 32) << 16);
    // ---- dump the partial accumulator
#pragma unroll 1
    for (int i = 0; i < NCH; ++i) {
      uint32_t lo[16], hi[16];
      ptx::tmem_ld_16x256_x4(tbase + (cf + i) * 32, lo);
      ptx::tmem_ld_16x256_x4(tbase + (16u << 16) + (cf + i) * 32, hi);
      ptx::tmem_ld_wait();
      float* const dst = mine + (cf + i) * 32;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        __stcg(reinterpret_cast<float2*>(dst + 8 * n), make_float2(__uint_as_float(lo[4 * n]), __uint_as_float(lo[4 * n + 1])));
        __stcg(reinterpret_cast<float2*>(dst + 8 * PX + 8 * n), make_float2(__uint_as_float(lo[4 * n + 2]), __uint_as_float(lo[4 * n + 3])));
        __stcg(reinterpret_cast<float2*>(dst + 16 * PX + 8 * n), make_float2(__uint_as_float(hi[4 * n]), __uint_as_float(hi[4 * n + 1])));
        __stcg(reinterpret_cast<float2*>(dst + 24 * PX + 8 * n), make_float2(__uint_as_float(hi[4 * n + 2]), __uint_as_float(hi[4 * n + 3])));
      }
    }
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tempty_bar[as]));
    __threadfence();
    __syncwarp();
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(&p.sk_ticket[tt * 8 + wslot], 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket != S - 1) continue;
    // ---- last arrival for this (tile, warp): reduce and finish
    __threadfence();
    if (lane == 0) p.sk_ticket[tt * 8 + wslot] = 0;
    epi_begin_tile<RBVAR>(p, eq, lane);
#pragma unroll 1
    for (int i = 0; i < NCH; ++i) {
      const float* const src = tile_part + frag_off + (cf + i) * 32;
      float acc[4][4][2];  // [k][n][pixel of the pair]
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[k][n][0] = acc[k][n][1] = 0.f;
#pragma unroll 1
      for (int sp = 0; sp < S; sp += 2) {  // two partials (32 loads of 8 bytes per lane) in flight
        float2 v[2][4][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (sp + u < S) {
            const float* const sb = src + static_cast<long long>(sp + u) * (128 * PX);
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int n = 0; n < 4; ++n) v[u][k][n] = __ldcg(reinterpret_cast<const float2*>(sb + 8 * k * PX + 8 * n));
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (sp + u < S) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int n = 0; n < 4; ++n) {
                acc[k][n][0] += v[u][k][n].x;
                acc[k][n][1] += v[u][k][n].y;
              }
          }
        }
      }
      uint32_t lo[16], hi[16];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        lo[4 * n] = __float_as_uint(acc[0][n][0]); lo[4 * n + 1] = __float_as_uint(acc[0][n][1]);
        lo[4 * n + 2] = __float_as_uint(acc[1][n][0]); lo[4 * n + 3] = __float_as_uint(acc[1][n][1]);
        hi[4 * n] = __float_as_uint(acc[2][n][0]); hi[4 * n + 1] = __float_as_uint(acc[2][n][1]);
        hi[4 * n + 2] = __float_as_uint(acc[3][n][0]); hi[4 * n + 3] = __float_as_uint(acc[3][n][1]);
      }
      const uint32_t buf = obuf + (nstore & 1) * (32 * 32 * 2);
      if (lane == 0) ptx::tma_store_wait_read<1>();
      __syncwarp();
      epilogue_chunk_q<RBVAR, STATS>(p, eq, lo, hi, eq.m0 + (cf + i) * 32, buf, lane);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && eq.nw0 < p.N && eq.m0 + (cf + i) * 32 < p.M) {
        ptx::tma_store_4d(&p.tm_out, buf, eq.nw0, eq.m0 + (cf + i) * 32, 0, 0);
        ptx::tma_store_commit();
      }
      ++nstore;
    }
    if (STATS) epi_write_stats(p, eq, static_cast<long long>(stat_mul) * pt + stat_add, lane);
  }
  if (lane == 0) ptx::tma_store_wait_read<0>();
}

template <int PX, int NCH>
__device__ __forceinline__ void epilogue_splitk_dispatch(const TcParams& p, uint32_t tmem_base, uint64_t* tfull_bar,
                                                         uint64_t* tempty_bar, uint32_t obuf, int q, int cf, int stat_mul,
                                                         int stat_add, int lane) {
  const bool rbvar = p.rowbias != nullptr && (p.rows_per_sample % PX) != 0;
  const bool stats = p.stat_part != nullptr;
  if (rbvar) {
    if (stats) epilogue_role_splitk<PX, NCH, true, true>(p, tmem_base, tfull_bar, tempty_bar, obuf, q, cf, stat_mul, stat_add, lane);
    else epilogue_role_splitk<PX, NCH, true, false>(p, tmem_base, tfull_bar, tempty_bar, obuf, q, cf, stat_mul, stat_add, lane);
  } else {
    if (stats) epilogue_role_splitk<PX, NCH, false, true>(p, tmem_base, tfull_bar, tempty_bar, obuf, q, cf, stat_mul, stat_add, lane);
    else epilogue_role_splitk<PX, NCH, false, false>(p, tmem_base, tfull_bar, tempty_bar, obuf, q, cf, stat_mul, stat_add, lane);
  }
}

// In-accumulator form (see the banner above gno_publish): one epilogue warp's loop over the CTA's tiles.
// While a warp waits, the MMA warp works on the next tile in the other accumulator.  No deadlock: CTAs are all
// resident, take their tiles in index order, and a sample's tiles span fewer consecutive indices than there are CTAs
// (checked on the host), so the tile a CTA must finish BEFORE one of sample s always belongs to an earlier sample.
template <int PX, int NCH>
__device__ __forceinline__ void epilogue_role_gn(const TcParams& p, uint32_t tmem_base, uint64_t* tfull_bar,
                                                 uint64_t* tempty_bar, uint32_t obuf, int q, int cf, int part_mul,
                                                 int part_add, int lane, int acc_owner = -1) {
  uint32_t tl = 0;
  uint32_t nstore = 0;
  const int tps = p.rows_per_sample / PX;       // pixel tiles per sample
  for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tl) {
    const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
    const int tt = p.reverse ? p.num_tiles - 1 - t : t;
    const int pt = tt / p.n_tiles;
    const int nt = tt - pt * p.n_tiles;
    const int nw0 = nt * 128 + q * 32;
    const int m0 = pt * PX;
    const int sample = m0 / p.rows_per_sample;
    const int part = (pt - sample * tps) * part_mul + part_add;
    const GnoSlot gs = gno_slot(p, sample, nt, q);
    float bch[4], ga[4], be[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int n = nw0 + (lane >> 2) + 8 * k;
      float b = p.bias ? __ldg(p.bias + n) : 0.f;
      if (p.rowbias) b += __ldg(p.rowbias + static_cast<long long>(sample) * p.rowbias_ld + n);
      bch[k] = b * p.alpha;
      ga[k] = __ldg(p.gno_gamma + n);
      be[k] = __ldg(p.gno_beta + n);
    }
    T2P_EPI_WAIT(ptx::smem_u32(&tfull_bar[as]), aph);
    ptx::tc_fence_after();
    const uint32_t tbase = tmem_base + as * PX + (static_cast<uint32_t>(q * 32) << 16);
    auto load_chunk = [&](int c, uint32_t (&lo)[16], uint32_t (&hi)[16]) {
      ptx::tmem_ld_16x256_x4(tbase + c * 32, lo);
      ptx::tmem_ld_16x256_x4(tbase + (16u << 16) + c * 32, hi);
    };
    uint32_t lo0[16], hi0[16], lo1[16], hi1[16];
    // ---- pass 1: statistics of v over this warp's pixels (thread slot k holds channel (lane >> 2) + 8 k)
    float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
    auto add_chunk = [&](const uint32_t (&lo)[16], const uint32_t (&hi)[16]) {
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const float v[4][2] = {{__uint_as_float(lo[4 * n]), __uint_as_float(lo[4 * n + 1])},
                               {__uint_as_float(lo[4 * n + 2]), __uint_as_float(lo[4 * n + 3])},
                               {__uint_as_float(hi[4 * n]), __uint_as_float(hi[4 * n + 1])},
                               {__uint_as_float(hi[4 * n + 2]), __uint_as_float(hi[4 * n + 3])}};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float f0 = fmaf(v[k][0], p.alpha, bch[k]), f1 = fmaf(v[k][1], p.alpha, bch[k]);
          ssum[k] += f0 + f1;
          ssq[k] = fmaf(f0, f0, fmaf(f1, f1, ssq[k]));
        }
      }
    };
    load_chunk(cf, lo0, hi0);
#pragma unroll 1
    for (int i = 0; i < NCH; i += 2) {
      ptx::tmem_ld_wait();
      if (i + 1 < NCH) load_chunk(cf + i + 1, lo1, hi1);
      add_chunk(lo0, hi0);
      if (i + 1 < NCH) {
        ptx::tmem_ld_wait();
        if (i + 2 < NCH) load_chunk(cf + i + 2, lo0, hi0);
        add_chunk(lo1, hi1);
      }
    }
    gno_publish(gs, part, ssum, ssq, lane);
    float mean_g, rstd_g;
    gno_fold(p, gs, lane, mean_g, rstd_g);
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(&p.gno_flags[gs.slot], 1);  // (its result is looked at after pass 2)
    float sc[4], sh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int src_lane = ((lane >> 2) + 8 * k) / gs.cpg;  // a lane that holds this channel's group
      const float m = __shfl_sync(0xffffffffu, mean_g, src_lane), rs = __shfl_sync(0xffffffffu, rstd_g, src_lane);
      sc[k] = rs * ga[k];
      sh[k] = be[k] - m * sc[k];
    }
    // ---- pass 2: normalise, activate, store
    auto release_acc = [&]() {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (acc_owner < 0) ptx::mbar_arrive(ptx::smem_u32(&tempty_bar[as]));
        else ptx::mbar_arrive_cluster(ptx::smem_u32(&tempty_bar[as]), static_cast<uint32_t>(acc_owner));
      }
    };
    auto emit_chunk = [&](const uint32_t (&lo)[16], const uint32_t (&hi)[16], int c) {
      const uint32_t buf = obuf + (nstore & 1) * (32 * 32 * 2);
      if (lane == 0) ptx::tma_store_wait_read<1>();
      __syncwarp();
      uint32_t pk[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const float v[4][2] = {{__uint_as_float(lo[4 * n]), __uint_as_float(lo[4 * n + 1])},
                               {__uint_as_float(lo[4 * n + 2]), __uint_as_float(lo[4 * n + 3])},
                               {__uint_as_float(hi[4 * n]), __uint_as_float(hi[4 * n + 1])},
                               {__uint_as_float(hi[4 * n + 2]), __uint_as_float(hi[4 * n + 3])}};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float y0 = fmaf(fmaf(v[k][0], p.alpha, bch[k]), sc[k], sh[k]);
          const float y1 = fmaf(fmaf(v[k][1], p.alpha, bch[k]), sc[k], sh[k]);
          pk[k][n] = ptx::pack_bf16x2(silu_tanh_f(y0), silu_tanh_f(y1));
        }
      }
      const int mt = lane >> 3, row = lane & 7;
      const uint32_t rowaddr = buf + (8 * (mt >> 1) + row) * 64;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int ph = 0; ph < 2; ++ph)
          ptx::stmatrix_x4_trans(rowaddr + ph * (16 * 64) + ((((2 * hf + (mt & 1)) ^ (row >> 1)) & 3) << 4), pk[2 * hf][2 * ph],
                                 pk[2 * hf + 1][2 * ph], pk[2 * hf][2 * ph + 1], pk[2 * hf + 1][2 * ph + 1]);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && nw0 < p.N && m0 + c * 32 < p.M) {
        ptx::tma_store_4d(&p.tm_out, buf, nw0, m0 + c * 32, 0, 0);
        ptx::tma_store_commit();
      }
      ++nstore;
    };
    load_chunk(cf, lo0, hi0);
#pragma unroll 1
    for (int i = 0; i < NCH; i += 2) {
      ptx::tmem_ld_wait();
      if (i + 1 < NCH) load_chunk(cf + i + 1, lo1, hi1);
      else release_acc();
      emit_chunk(lo0, hi0, cf + i);
      if (i + 1 < NCH) {
        ptx::tmem_ld_wait();
        if (i + 2 < NCH) load_chunk(cf + i + 2, lo0, hi0);
        else release_acc();
        emit_chunk(lo1, hi1, cf + i + 1);
      }
    }
    gno_leave(p, gs, ticket, p.gno_parts, lane);
  }
  if (lane == 0) ptx::tma_store_wait_read<0>();
}

// Deferred form: one fix-up warp per TMEM lane quadrant q normalises, in place and out of L2, the 32 channels x PX pixels
// that epilogue warp q stored raw (bf16) for each of the CTA's tiles.  `ready_smem`: the shared-memory word in which
// that epilogue warp counts the tiles whose stores have been performed (it waits for its bulk groups before it takes
// the next accumulator -- time it would spend waiting for the MMA anyway).  Lane = (pixel % 8, 8-channel slot): a warp
// instruction moves 8 pixels x 64 bytes, eight 16-byte loads in flight per lane.
template <int PX>
__device__ __forceinline__ void fixup_role(const TcParams& p, uint32_t ready_smem, int q, int lane) {
  const int tps = p.rows_per_sample / PX;
  const int c8 = lane & 3;
  uint32_t tl = 0;
  for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tl) {
    const int tt = p.reverse ? p.num_tiles - 1 - t : t;
    const int pt = tt / p.n_tiles;
    const int nt = tt - pt * p.n_tiles;
    const int ch0 = nt * 128 + q * 32 + c8 * 8;
    const int m0 = pt * PX;
    const int sample = m0 / p.rows_per_sample;
    const GnoSlot gs = gno_slot(p, sample, nt, q);
    const float4 ga0 = __ldg(reinterpret_cast<const float4*>(p.gno_gamma + ch0)), ga1 = __ldg(reinterpret_cast<const float4*>(p.gno_gamma + ch0 + 4));
    const float4 be0 = __ldg(reinterpret_cast<const float4*>(p.gno_beta + ch0)), be1 = __ldg(reinterpret_cast<const float4*>(p.gno_beta + ch0 + 4));
    float mean_g, rstd_g;
    gno_fold(p, gs, lane, mean_g, rstd_g);
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(&p.gno_flags[gs.slot], 1);
    // h = y / 2 = x * (scale / 2) + shift / 2;  silu(y) = h + h * tanh(h)
    const float ga[8] = {ga0.x, ga0.y, ga0.z, ga0.w, ga1.x, ga1.y, ga1.z, ga1.w};
    const float be[8] = {be0.x, be0.y, be0.z, be0.w, be1.x, be1.y, be1.z, be1.w};
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int src_lane = (c8 * 8 + e) / gs.cpg;
      const float m = __shfl_sync(0xffffffffu, mean_g, src_lane), rs = __shfl_sync(0xffffffffu, rstd_g, src_lane);
      sc[e] = 0.5f * rs * ga[e];
      sh[e] = 0.5f * be[e] - m * sc[e];
    }
#ifdef T2P_TIMING_KNOBS
    if (p.gno_debug & 1) {
      if (!(p.gno_debug & 2) && lane == 0) while (ptx::ld_acquire_cta_shared(ready_smem) <= static_cast<int>(tl)) __nanosleep(200);
      __syncwarp();
      gno_leave(p, gs, ticket, tps, lane);
      continue;
    }
#endif
    if (lane == 0) {  // the raw tile must have reached global memory
      const long long t0 = clock64();
      while (ptx::ld_acquire_cta_shared(ready_smem) <= static_cast<int>(tl)) {
        __nanosleep(200);
        if (clock64() - t0 > 4000000000LL) __trap();
      }
    }
    __syncwarp();
    char* const base = static_cast<char*>(p.out) + (static_cast<long long>(m0 + (lane >> 2)) * p.N + ch0) * 2;
    const long long pitch = 8LL * p.N * 2;  // eight pixels on
    // sixteen 16-byte loads per lane in flight (half of the warp's 16 KB slice): an L2 round trip is ~2 k clocks under
    // this kernel's operand traffic; with eight in flight the four round trips per tile made the fix-up warps as slow as
    // the MMA (all thirty-two do not fit the register file next to the other roles)
    constexpr int NLD = 16;
#pragma unroll 1
    for (int i = 0; i < PX / 8; i += NLD) {
      uint4 d[NLD];
#pragma unroll
      for (int b = 0; b < NLD; ++b) d[b] = __ldcg(reinterpret_cast<const uint4*>(base + (i + b) * pitch));
#pragma unroll
      for (int b = 0; b < NLD; ++b) {
        uint32_t w[4] = {d[b].x, d[b].y, d[b].z, d[b].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float h0 = fmaf(__uint_as_float(w[j] << 16), sc[2 * j], sh[2 * j]);
          const float h1 = fmaf(__uint_as_float(w[j] & 0xffff0000u), sc[2 * j + 1], sh[2 * j + 1]);
          float t0, t1;
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
          w[j] = ptx::pack_bf16x2(fmaf(h0, t0, h0), fmaf(h1, t1, h1));
        }
        __stcg(reinterpret_cast<uint4*>(base + (i + b) * pitch), make_uint4(w[0], w[1], w[2], w[3]));
      }
    }
    gno_leave(p, gs, ticket, tps, lane);
  }
}

// GNO: 0 = the kernel never normalises its output, 1 = in-accumulator form, 2 = deferred form (kernels with fix-up warps)
template <int PX, int NCH, int GNO = 1>
__device__ __forceinline__ void epilogue_dispatch(const TcParams& p, uint32_t tmem_base, uint64_t* tfull_bar,
                                                  uint64_t* tempty_bar, uint32_t obuf, int q, int cf, int stat_mul,
                                                  int stat_add, int lane, uint32_t ready_smem = 0, int acc_owner = -1) {
  if (GNO == 2) {
    epilogue_role<PX, NCH, false, true, true>(p, tmem_base, tfull_bar, tempty_bar, obuf, q, cf, stat_mul, stat_add, lane, ready_smem);
    return;
  }
  if (GNO == 1 && p.gno_gamma) {  // (part index = stat_mul * tile + stat_add: whole tiles, or the halves of the eight-warp layout)
    epilogue_role_gn<PX, NCH>(p, tmem_base, tfull_bar, tempty_bar, obuf, q, cf, stat_mul, stat_add, lane, acc_owner);
    return;
  }
  const bool rbvar = p.rowbias != nullptr && (p.rows_per_sample % PX) != 0;
  const bool stats = p.stat_part != nullptr;
  if (rbvar) {
    if (stats) epilogue_role<PX, NCH, true, true>(p, tmem_base, tfull_bar, tempty_bar, obuf, q, cf, stat_mul, stat_add, lane, 0, acc_owner);
    else epilogue_role<PX, NCH, true, false>(p, tmem_base, tfull_bar, tempty_bar, obuf, q, cf, stat_mul, stat_add, lane, 0, acc_owner);
  } else {
    if (stats) epilogue_role<PX, NCH, false, true>(p, tmem_base, tfull_bar, tempty_bar, obuf, q, cf, stat_mul, stat_add, lane, 0, acc_owner);
    else epilogue_role<PX, NCH, false, false>(p, tmem_base, tfull_bar, tempty_bar, obuf, q, cf, stat_mul, stat_add, lane, 0, acc_owner);
  }
}

template <int PX>
__global__ void __launch_bounds__(CfgT<PX>::THREADS, 1) conv_gemm_tcT_kernel(const __grid_constant__ TcParams p) {
  using C = CfgT<PX>;
#ifdef T2P_TIMING_KNOBS
#define T2P_TSTAMP(i) do { if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) p.trace[i] = clock64(); } while (0)
#else
#define T2P_TSTAMP(i) do { } while (0)
#endif
  if (threadIdx.x == 0) T2P_TSTAMP(0);
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[C::STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tiles = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t out_stage = tiles + C::STAGES * C::STAGE_BYTES;  // 1024-aligned (stage sizes are multiples of 1 KB)

  const int ctot = p.c0 + p.c1;
  const int chunks_per_tap = ctot / BK;
  const int x_chunks = (p.xc0 + p.xc1) / BK;
  const int num_kb = p.taps * chunks_per_tap + x_chunks + (p.residual ? 2 : 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&tfull_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&tempty_bar[s]), C::EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (warp == 2) prefetch_l2_range(p, lane);
  if (threadIdx.x == 0) T2P_TSTAMP(1);
  pdl_wait();
  if (threadIdx.x == 0) T2P_TSTAMP(2);  // everything above overlapped the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::prefetch_tmap(&p.tm_a0);
      ptx::prefetch_tmap(&p.tm_w);
      if (p.c1 > 0) ptx::prefetch_tmap(&p.tm_a1);
      if (p.residual) {
        ptx::prefetch_tmap(&p.tm_res);
        ptx::prefetch_tmap(&p.tm_ident);
      }
      if (p.xc0 > 0) ptx::prefetch_tmap(&p.tm_x0);
      if (p.xc1 > 0) ptx::prefetch_tmap(&p.tm_x1);
      const int pad = (p.taps == 9) ? 1 : 0;
      const int hw = p.H * p.W;
      int s = 0;
      uint32_t ph = 0;
      // work item = (tile, K split): with p.splits > 1 the k-blocks of a tile are shared out over consecutive CTAs
      const int S = p.splits;
      const int num_items = p.num_tiles * S;
      const int kb_taps = p.taps * chunks_per_tap;
      for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
        const int t = it / S;
        const int split = it - t * S;
        const int kb0 = split * num_kb / S, kb1 = (split + 1) * num_kb / S;
        const int tt = p.reverse ? p.num_tiles - 1 - t : t;
        const int pt = tt / p.n_tiles;
        const int m0 = pt * PX;
        const int n0 = (tt - pt * p.n_tiles) * 128;
        int b0 = 0, h0 = 0, w0 = m0;
        if (!p.mode2d) {
          b0 = m0 / hw;
          const int rem = m0 - b0 * hw;
          h0 = rem / p.W;
          w0 = rem - h0 * p.W;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t fb = ptx::smem_u32(&full_bar[s]);
          ptx::mbar_arrive_expect_tx(fb, C::STAGE_BYTES);
          const uint32_t sw = tiles + s * C::STAGE_BYTES;
          const uint32_t sp = sw + C::W_BYTES;
          if (kb < kb_taps) {
            const int tap = kb / chunks_per_tap;
            const int ch = (kb - tap * chunks_per_tap) * BK;
            const int kh = (p.taps == 9) ? tap / 3 : 0;
            const int kw = (p.taps == 9) ? tap - kh * 3 : 0;
            ptx::tma_load_4d(sw, &p.tm_w, fb, kb * BK, n0, 0, 0);
            if (ch < p.c0)
              ptx::tma_load_4d(sp, &p.tm_a0, fb, ch, w0 + kw - pad, h0 + kh - pad, b0);
            else
              ptx::tma_load_4d(sp, &p.tm_a1, fb, ch - p.c0, w0 + kw - pad, h0 + kh - pad, b0);
          } else if (kb < kb_taps + x_chunks) {
            // centre-tap-only sources (the skip path's 1x1 convolution folded into this GEMM)
            const int ch = (kb - kb_taps) * BK;
            ptx::tma_load_4d(sw, &p.tm_w, fb, kb * BK, n0, 0, 0);
            if (ch < p.xc0)
              ptx::tma_load_4d(sp, &p.tm_x0, fb, ch, w0, h0, b0);
            else
              ptx::tma_load_4d(sp, &p.tm_x1, fb, ch - p.xc0, w0, h0, b0);
          } else {
            // residual[m][n0 .. n0 + 128) enters the accumulator through two identity k-blocks: exact in fp32,
            // fetched by TMA like any operand, no epilogue loads
            const int j = kb - kb_taps - x_chunks;
            ptx::tma_load_4d(sw, &p.tm_ident, fb, j * BK, 0, 0, 0);
            ptx::tma_load_4d(sp, &p.tm_res, fb, n0 + j * BK, w0, h0, b0);
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(PX);  // M = 128 channels, N = PX pixels
      int s = 0;
      uint32_t ph = 0;
      uint32_t tl = 0;
      const int S = p.splits;
      const int num_items = p.num_tiles * S;
      for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++tl) {
        const int split = it % S;
        const int kb0 = split * num_kb / S, kb1 = (split + 1) * num_kb / S;
        const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
        ptx::mbar_wait(ptx::smem_u32(&tempty_bar[as]), aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * PX;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[s]), ph);
          ptx::tc_fence_after();
          if (tl == 0 && kb == kb0) T2P_TSTAMP(3);
          if (tl == 0 && kb == kb1 - 1) T2P_TSTAMP(4);
          const uint32_t sw = tiles + s * C::STAGE_BYTES;
          const uint64_t dw = ptx::umma_desc_k_sw128(sw);
          const uint64_t dp = ptx::umma_desc_k_sw128(sw + C::W_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) ptx::umma_bf16(tmem_acc, dw + 2 * k, dp + 2 * k, idesc, kb > kb0 || k != 0);
          ptx::umma_commit(ptx::smem_u32(&empty_bar[s]));
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        ptx::umma_commit(ptx::smem_u32(&tfull_bar[as]));
        if (tl == 0) T2P_TSTAMP(5);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9): thread = channel.
    // Two warps share each TMEM lane quadrant (warp % 4) and split the tile's pixel columns in halves: a CTA with a
    // single tile (the many latency-bound launches at 16 x 16 and below) drains its accumulator twice as fast.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int HC = PX / 64;  // 32-pixel chunks per warp (half of the tile)
    // this warp's two [32 px][32 ch] bf16 staging buffers; statistics tile = its half of the pixel tile: 2 * pt + half
    if (p.splits > 1)
      epilogue_splitk_dispatch<PX, HC>(p, tmem_base, tfull_bar, tempty_bar, out_stage + (warp - 2) * (2 * 32 * 32 * 2), q,
                                       half * HC, 2, half, lane);
    else
      epilogue_dispatch<PX, HC>(p, tmem_base, tfull_bar, tempty_bar, out_stage + (warp - 2) * (2 * 32 * 32 * 2), q,
                                half * HC, 2, half, lane);
  }

  if (warp == 2) T2P_TSTAMP(8);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  if (threadIdx.x == 0) T2P_TSTAMP(9);
}

// =====================================================================================================
// CTA-pair variant of the channel-major kernel (tcgen05.mma.cta_group::2) for layers of 256 output channels and more:
// the two CTAs of a cluster work on ONE 256-pixel tile and TWO adjacent 128-channel tiles.  Each holds its own 128 rows of
// the weight matrix (A, M = 256 over the pair) and HALF of the pixel tile (B, N = 256: 128 pixels each); the leader's MMA
// reads both halves.  Per CTA and k-block 16 + 16 KB come in from L2 instead of 16 + 32: the channel-major kernel is
// bound by what one SM takes from L2 (profiles/r02_tcT_trace_B8.txt), not by the tensor core.
//   * barriers: `full` lives in the leader (both CTAs' TMA loads count their bytes there: cta_group::2 loads with the
//     peer bit of the barrier address cleared; the leader arms it with the bytes of BOTH); `empty` and `tfull` exist in
//     both CTAs and the leader's tcgen05.commit arrives on both (multicast); `tempty` lives in the leader and takes the
//     epilogue warps of both CTAs (the follower's arrive remotely).
//   * tile order: CTA b takes tiles b, b + grid, ...; with an even number of channel tiles and an even grid the CTAs
//     2c and 2c + 1 always hold the same pixel tile and adjacent channel tiles.
//   * residual k-blocks: the follower's identity block sits 128 columns further left (TMA fills what is outside with
//     zeros), so each CTA adds the residual channels of its own rows.
// Epilogue: each CTA drains its own 128 lanes of TMEM exactly as the one-CTA kernel does.
struct CfgP {
  static constexpr int PX = 256;
  static constexpr int W_BYTES = 128 * BK * 2;
  static constexpr int P_BYTES = (PX / 2) * BK * 2;
  static constexpr int STAGE_BYTES = W_BYTES + P_BYTES;  // per CTA
  static constexpr int STAGES = 6;
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int OUT_BYTES = EPI_WARPS * 2 * 32 * 32 * 2;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * PX;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CfgP::THREADS, 1)
    conv_gemm_tcP_kernel(const __grid_constant__ TcParams p) {
  using C = CfgP;
  constexpr int PX = C::PX;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[C::STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();  // 0 = leader of the pair
  const uint32_t tiles = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t out_stage = tiles + C::STAGES * C::STAGE_BYTES;

  const int ctot = p.c0 + p.c1;
  const int chunks_per_tap = ctot / BK;
  const int x_chunks = (p.xc0 + p.xc1) / BK;
  const int num_kb = p.taps * chunks_per_tap + x_chunks + (p.residual ? 4 : 0);  // (residual: the pair's 256 channels)

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&tfull_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&tempty_bar[s]), 2 * C::EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(ptx::smem_u32(&tmem_base_slot), C::TMEM_COLS);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // the peer's barriers are initialised before anything arrives on them
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (warp == 2) prefetch_l2_range(p, lane);
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      ptx::prefetch_tmap(&p.tm_a0);
      ptx::prefetch_tmap(&p.tm_w);
      if (p.c1 > 0) ptx::prefetch_tmap(&p.tm_a1);
      if (p.residual) {
        ptx::prefetch_tmap(&p.tm_res);
        ptx::prefetch_tmap(&p.tm_ident);
      }
      if (p.xc0 > 0) ptx::prefetch_tmap(&p.tm_x0);
      if (p.xc1 > 0) ptx::prefetch_tmap(&p.tm_x1);
      const int pad = (p.taps == 9) ? 1 : 0;
      const int hw = p.H * p.W;
      const int kb_taps = p.taps * chunks_per_tap;
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int tt = p.reverse ? p.num_tiles - 1 - t : t;
        const int pt = tt / p.n_tiles;
        const int n0 = (tt - pt * p.n_tiles) * 128;
        const int m0 = pt * PX + static_cast<int>(rank) * (PX / 2);  // this CTA's half of the pixel tile
        int b0 = 0, h0 = 0, w0 = m0;
        if (!p.mode2d) {
          b0 = m0 / hw;
          const int rem = m0 - b0 * hw;
          h0 = rem / p.W;
          w0 = rem - h0 * p.W;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t fb = ptx::smem_u32(&full_bar[s]);
          if (rank == 0) ptx::mbar_arrive_expect_tx(fb, 2 * C::STAGE_BYTES);  // the bytes of both CTAs land on the leader's barrier
          const uint32_t sw = tiles + s * C::STAGE_BYTES;
          const uint32_t sp = sw + C::W_BYTES;
          if (kb < kb_taps) {
            const int tap = kb / chunks_per_tap;
            const int ch = (kb - tap * chunks_per_tap) * BK;
            const int kh = (p.taps == 9) ? tap / 3 : 0;
            const int kw = (p.taps == 9) ? tap - kh * 3 : 0;
            ptx::tma_load_4d_pair(sw, &p.tm_w, fb, kb * BK, n0, 0, 0);
            if (ch < p.c0)
              ptx::tma_load_4d_pair(sp, &p.tm_a0, fb, ch, w0 + kw - pad, h0 + kh - pad, b0);
            else
              ptx::tma_load_4d_pair(sp, &p.tm_a1, fb, ch - p.c0, w0 + kw - pad, h0 + kh - pad, b0);
          } else if (kb < kb_taps + x_chunks) {
            const int ch = (kb - kb_taps) * BK;
            ptx::tma_load_4d_pair(sw, &p.tm_w, fb, kb * BK, n0, 0, 0);
            if (ch < p.xc0)
              ptx::tma_load_4d_pair(sp, &p.tm_x0, fb, ch, w0, h0, b0);
            else
              ptx::tma_load_4d_pair(sp, &p.tm_x1, fb, ch - p.xc0, w0, h0, b0);
          } else {
            // residual of the pair's 256 channels through the identity: k-block j covers residual channels
            // [c_lo + 64 j, + 64) with c_lo = the pair's first channel; this CTA's rows are channels n0 .. n0 + 127
            const int j = kb - kb_taps - x_chunks;  // 0..3
            const int c_lo = n0 - (n0 & 128);
            ptx::tma_load_4d_pair(sw, &p.tm_ident, fb, c_lo + j * BK - n0, 0, 0, 0);
            ptx::tma_load_4d_pair(sp, &p.tm_res, fb, c_lo + j * BK, w0, h0, b0);
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
      }
      // Tail: the leader's commits arrive on THIS CTA's `empty` barriers asynchronously, and nothing else waits for the
      // last ring-full of them.  A CTA must not exit (its shared memory handed to the next CTA) while an arrive from its
      // peer is still under way: wait for every stage's last release before leaving.
      for (int i = 0; i < C::STAGES; ++i) {
        ptx::mbar_wait(ptx::smem_u32(&empty_bar[s]), ph ^ 1);
        if (++s == C::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader only)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_m256(PX);
      int s = 0;
      uint32_t ph = 0;
      uint32_t tl = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tl) {
        const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
        ptx::mbar_wait(ptx::smem_u32(&tempty_bar[as]), aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * PX;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[s]), ph);
          ptx::tc_fence_after();
          const uint32_t sw = tiles + s * C::STAGE_BYTES;
          const uint64_t dw = ptx::umma_desc_k_sw128(sw);
          const uint64_t dp = ptx::umma_desc_k_sw128(sw + C::W_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) ptx::umma_bf16_pair(tmem_acc, dw + 2 * k, dp + 2 * k, idesc, (kb | k) != 0);
          ptx::umma_commit_pair(ptx::smem_u32(&empty_bar[s]));
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        ptx::umma_commit_pair(ptx::smem_u32(&tfull_bar[as]));
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9 of both CTAs)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int HC = PX / 64;
    epilogue_dispatch<PX, HC, 1>(p, tmem_base, tfull_bar, tempty_bar, out_stage + (warp - 2) * (2 * 32 * 32 * 2), q,
                                 half * HC, 2, half, lane, 0, /*acc_owner=*/0);
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // neither CTA leaves (or frees tensor memory) while the other may still reach into it
  if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
}

// =====================================================================================================
// Halo variant of the channel-major kernel for 3x3 convolutions on 128-pixel-wide images (the layers that hold
// most of the FLOPs).  A pixel tile is two whole image rows; for each (kh, channel chunk) ONE box of
// 2 rows x 130 pixels (one-pixel halo left and right, zero-filled by TMA at the image border) is staged and the
// three horizontal taps read it through descriptors shifted by kw rows of 128 bytes -- the 128-byte swizzle is a
// function of the absolute shared-memory address, so a row-shifted window of a swizzled tile is itself a valid
// operand (probe: csrc/probe_shift.cu).  The L2 -> SMEM stream per (kh, chunk) drops from 3 x (16 + 32) KB to
// 3 x 16 + 32.5 KB.  Pixels and weights run through separate TMA rings.
template <int EW, bool GND = false>
struct CfgHT {
  static constexpr int W_BYTES = 128 * BK * 2;           // 16 KB weight tile (128 channels x 64 k)
  static constexpr int P_ROWPITCH = 130;                 // pixels per staged image row (128 + halo)
  static constexpr int P_LOAD_BYTES = 2 * P_ROWPITCH * BK * 2;  // 33 280
  static constexpr int P_BYTES = 33 * 1024;              // buffer pitch (1 KB multiple)
  static constexpr int NP = 3;                           // pixel buffers
  static constexpr int EPI_WARPS = EW;                   // 4, or 8: two per TMEM lane quadrant, half a tile each
  static constexpr int NW = EPI_WARPS == 8 ? 5 : 6;      // weight buffers (the staging of eight warps takes one)
  static constexpr int FIX_WARPS = GND ? 4 : 0;          // deferred GroupNorm of the output: one fix-up warp per quadrant
  static constexpr int THREADS = 64 + 32 * EPI_WARPS + 32 * FIX_WARPS;
  static constexpr int OUT_BYTES = EPI_WARPS * 2 * 32 * 32 * 2;
  static constexpr int SMEM_BYTES = NP * P_BYTES + NW * W_BYTES + OUT_BYTES + 1024;
  static constexpr int PX = 256;
  static constexpr int TMEM_COLS = 2 * PX;
};
#ifndef T2P_H_EPI_WARPS
#define T2P_H_EPI_WARPS 4
#endif
// (the plain epilogue is fastest with four warps; eight measured slower: one weight buffer less)
using CfgH = CfgHT<T2P_H_EPI_WARPS>;

// GND: the launch normalises its own output in the deferred form (see gno_publish): four epilogue warps store the raw
// tile and publish its statistics, four more warps rewrite it in place.
template <int EW, bool GND>
__global__ void __launch_bounds__(CfgHT<EW, GND>::THREADS, 1) conv_gemm_tcH_kernel(const __grid_constant__ TcParams p) {
  using C = CfgHT<EW, GND>;
  constexpr int PX = C::PX;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t pfull_bar[C::NP], pempty_bar[C::NP];
  __shared__ __align__(8) uint64_t wfull_bar[C::NW], wempty_bar[C::NW];
  __shared__ __align__(8) uint64_t tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ int stored_tiles[4];  // GND: per quadrant, tiles whose raw stores have been performed

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t pbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t wbase = pbase + C::NP * C::P_BYTES;
  const uint32_t out_stage = wbase + C::NW * C::W_BYTES;

  const int chunks = (p.c0 + p.c1) / BK;      // channel chunks of the 3x3 sources
  const int x_chunks = (p.xc0 + p.xc1) / BK;  // centre-tap-only sources
  const int r_chunks = p.residual ? 2 : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NP; ++s) {
      ptx::mbar_init(ptx::smem_u32(&pfull_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&pempty_bar[s]), 1);
    }
    for (int s = 0; s < C::NW; ++s) {
      ptx::mbar_init(ptx::smem_u32(&wfull_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&wempty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&tfull_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&tempty_bar[s]), C::EPI_WARPS);
    }
    for (int s = 0; s < 4; ++s) stored_tiles[s] = 0;
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (warp == 2) prefetch_l2_range(p, lane);
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::prefetch_tmap(&p.tm_a0);
      ptx::prefetch_tmap(&p.tm_w);
      if (p.c1 > 0) ptx::prefetch_tmap(&p.tm_a1);
      if (p.xc0 > 0) ptx::prefetch_tmap(&p.tm_x0);
      if (p.xc1 > 0) ptx::prefetch_tmap(&p.tm_x1);
      if (p.residual) {
        ptx::prefetch_tmap(&p.tm_res);
        ptx::prefetch_tmap(&p.tm_ident);
      }
      const int hw = p.H * p.W;
      int ps = 0, ws = 0;
      uint32_t pph = 0, wph = 0;
      auto load_w = [&](const CUtensorMap* tm, int c0, int c1) {
        ptx::mbar_wait(ptx::smem_u32(&wempty_bar[ws]), wph ^ 1);
        const uint32_t fb = ptx::smem_u32(&wfull_bar[ws]);
        ptx::mbar_arrive_expect_tx(fb, C::W_BYTES);
        ptx::tma_load_4d(wbase + ws * C::W_BYTES, tm, fb, c0, c1, 0, 0);
        if (++ws == C::NW) { ws = 0; wph ^= 1; }
      };
      auto load_p = [&](const CUtensorMap* tm, int bytes, int ch, int w, int h, int b) {
        ptx::mbar_wait(ptx::smem_u32(&pempty_bar[ps]), pph ^ 1);
        const uint32_t fb = ptx::smem_u32(&pfull_bar[ps]);
        ptx::mbar_arrive_expect_tx(fb, bytes);
        ptx::tma_load_4d(pbase + ps * C::P_BYTES, tm, fb, ch, w, h, b);
        if (++ps == C::NP) { ps = 0; pph ^= 1; }
      };
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int tt = p.reverse ? p.num_tiles - 1 - t : t;
        const int pt = tt / p.n_tiles;
        const int m0 = pt * PX;
        const int n0 = (tt - pt * p.n_tiles) * 128;
        const int b0 = m0 / hw;
        const int h0 = (m0 - b0 * hw) / p.W;  // even; the tile is image rows h0, h0 + 1
        for (int kh = 0; kh < 3; ++kh) {
          for (int cc = 0; cc < chunks; ++cc) {
            const int ch = cc * BK;
            if (ch < p.c0) load_p(&p.tm_a0, C::P_LOAD_BYTES, ch, -1, h0 + kh - 1, b0);
            else load_p(&p.tm_a1, C::P_LOAD_BYTES, ch - p.c0, -1, h0 + kh - 1, b0);
            for (int kw = 0; kw < 3; ++kw) load_w(&p.tm_w, ((kh * 3 + kw) * chunks + cc) * BK, n0);
          }
        }
        int kb = 9 * chunks;
        for (int cc = 0; cc < x_chunks; ++cc, ++kb) {  // folded skip path: centre tap, plain 2 x 128 pixel box
          const int ch = cc * BK;
          if (ch < p.xc0) load_p(&p.tm_x0, PX * BK * 2, ch, 0, h0, b0);
          else load_p(&p.tm_x1, PX * BK * 2, ch - p.xc0, 0, h0, b0);
          load_w(&p.tm_w, kb * BK, n0);
        }
        for (int j = 0; j < r_chunks; ++j) {  // residual through identity k-blocks
          load_p(&p.tm_res, PX * BK * 2, n0 + j * BK, 0, h0, b0);
          load_w(&p.tm_ident, j * BK, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128);  // M = 128 channels, N = 128 pixels (one image row)
      int ps = 0, ws = 0;
      uint32_t pph = 0, wph = 0;
      uint32_t tl = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tl) {
        const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
        ptx::mbar_wait(ptx::smem_u32(&tempty_bar[as]), aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * PX;
        bool first = true;
        // one weight tile against the two image rows of a staged pixel buffer (row pitch / tap shift in pixels)
        auto mma_block = [&](uint32_t pbuf, int rowpitch, int shift) {
          ptx::mbar_wait(ptx::smem_u32(&wfull_bar[ws]), wph);
          ptx::tc_fence_after();
          const uint64_t dw = ptx::umma_desc_k_sw128(wbase + ws * C::W_BYTES);
          const uint64_t d0 = ptx::umma_desc_k_sw128(pbuf + shift * 128);
          const uint64_t d1 = ptx::umma_desc_k_sw128(pbuf + (rowpitch + shift) * 128);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t acc = (first && k == 0) ? 0u : 1u;
            ptx::umma_bf16(tmem_acc, dw + 2 * k, d0 + 2 * k, idesc, acc);
            ptx::umma_bf16(tmem_acc + 128, dw + 2 * k, d1 + 2 * k, idesc, acc);
          }
          first = false;
          ptx::umma_commit(ptx::smem_u32(&wempty_bar[ws]));
          if (++ws == C::NW) { ws = 0; wph ^= 1; }
        };
        for (int kc = 0; kc < 3 * chunks; ++kc) {
          ptx::mbar_wait(ptx::smem_u32(&pfull_bar[ps]), pph);
          ptx::tc_fence_after();
          const uint32_t pbuf = pbase + ps * C::P_BYTES;
          for (int kw = 0; kw < 3; ++kw) mma_block(pbuf, C::P_ROWPITCH, kw);
          ptx::umma_commit(ptx::smem_u32(&pempty_bar[ps]));
          if (++ps == C::NP) { ps = 0; pph ^= 1; }
        }
        for (int kc = 0; kc < x_chunks + r_chunks; ++kc) {
          ptx::mbar_wait(ptx::smem_u32(&pfull_bar[ps]), pph);
          ptx::tc_fence_after();
          mma_block(pbase + ps * C::P_BYTES, 128, 0);
          ptx::umma_commit(ptx::smem_u32(&pempty_bar[ps]));
          if (++ps == C::NP) { ps = 0; pph ^= 1; }
        }
        ptx::umma_commit(ptx::smem_u32(&tfull_bar[as]));
      }
    }
  } else if (warp < 2 + C::EPI_WARPS) {
    // ------------------------------------------------------------ epilogue (warps 2..5): thread = channel
    const int q = warp & 3;
    if constexpr (GND) {
      static_assert(!GND || C::EPI_WARPS == 4, "deferred GroupNorm: one epilogue warp per quadrant");
      epilogue_dispatch<PX, PX / 32, 2>(p, tmem_base, tfull_bar, tempty_bar, out_stage + q * (2 * 32 * 32 * 2), q, 0, 1, 0, lane,
                                        ptx::smem_u32(&stored_tiles[q]));
    } else if constexpr (C::EPI_WARPS == 8) {
      const int half = (warp - 2) >> 2;
      epilogue_dispatch<PX, PX / 64, 1>(p, tmem_base, tfull_bar, tempty_bar, out_stage + (warp - 2) * (2 * 32 * 32 * 2), q,
                                        half * (PX / 64), 2, half, lane);
    } else {
      epilogue_dispatch<PX, PX / 32, 1>(p, tmem_base, tfull_bar, tempty_bar, out_stage + q * (2 * 32 * 32 * 2), q, 0, 1, 0, lane);
    }
  } else {
    // ------------------------------------------------------------ fix-up (GND; warps 6..9), quadrant = the epilogue warp's
    const int q = warp & 3;
    fixup_role<PX>(p, ptx::smem_u32(&stored_tiles[q]), q, lane);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}


// =====================================================================================================
// Halo kernel with GroupNorm-apply + SiLU fused into the operand path (ResnetBlockBigGANpp: h = act(GroupNorm(x))
// feeding Conv_0 / Conv_1, layers.py:305,318).  The normalised copy of the activation -- one HBM write and one HBM
// read of the largest tensors of the network per convolution -- never exists: eight extra warps fetch the RAW
// pixels of each (kh, channel chunk) box with 16-byte global loads, apply y = silu(x * scale + shift) and store the
// result straight into the 128-byte-swizzled tile the MMA warp reads (shared-memory traffic is what the TMA write of
// the plain kernel was; the first attempt, r01_fused_gn_halo_ab.txt, rewrote TMA-staged tiles in place and paid a
// read-modify-write of every box in a shared-memory pipe the UMMA operand reads already fill).  The arithmetic is
// packed bf16: with the affine pre-halved, h = fma(x, s/2, b/2), silu(2h) = h * tanh(h) + h -- one HFMA2, one
// MUFU.TANH (bf16x2), one HFMA2 per PAIR of elements.  The conv's zero padding applies to the activation, so halo
// pixels outside the image are stored as zeros, not as silu(shift).  Weights, the folded skip path (centre tap,
// raw) and residual k-blocks still arrive by TMA; both producers fill the same ring of pixel buffers in a fixed
// global order.  The two groups of four transform warps take alternate buffers, so each has two MMA periods to
// hide its load latency.
struct CfgHF {
  static constexpr int W_BYTES = CfgH::W_BYTES;
  static constexpr int P_ROWPITCH = CfgH::P_ROWPITCH;
  static constexpr int P_ROWS = 2 * P_ROWPITCH;          // pixel rows (128 B each) of a staged box
  static constexpr int P_BYTES = CfgH::P_BYTES;
  static constexpr int NP = 3;
  static constexpr int NW = 6;
  static constexpr int OUT_BYTES = CfgH::OUT_BYTES;
  static constexpr int SMEM_BYTES = NP * P_BYTES + NW * W_BYTES + OUT_BYTES + 1024;
  static constexpr int PX = 256;
  static constexpr int TMEM_COLS = 2 * PX;
#ifndef T2P_HF_GROUPS
#define T2P_HF_GROUPS 3
#endif
  static constexpr int T_GROUPS = T2P_HF_GROUPS;         // transform groups of four warps, taking the buffers in turn
  static constexpr int T_WARPS = 4 * T_GROUPS;
  static constexpr int THREADS = NUM_THREADS + 32 * T_WARPS;
};

__device__ __forceinline__ uint32_t bf16x2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t bf16x2_tanh(uint32_t a) {
  uint32_t d;
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

__global__ void __launch_bounds__(CfgHF::THREADS, 1) conv_gemm_tcHF_kernel(const __grid_constant__ TcParams p) {
  using C = CfgHF;
  constexpr int PX = C::PX;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t pfull_bar[C::NP], pempty_bar[C::NP];
  __shared__ __align__(8) uint64_t wfull_bar[C::NW], wempty_bar[C::NW];
  __shared__ __align__(8) uint64_t tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t pbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t wbase = pbase + C::NP * C::P_BYTES;
  const uint32_t out_stage = wbase + C::NW * C::W_BYTES;

  const int chunks = (p.c0 + p.c1) / BK;
  const int x_chunks = (p.xc0 + p.xc1) / BK;
  const int r_chunks = p.residual ? 2 : 0;
  const int bufs_per_tile = 3 * chunks + x_chunks + r_chunks;  // pixel buffers per tile, in ring order

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NP; ++s) {
      ptx::mbar_init(ptx::smem_u32(&pfull_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&pempty_bar[s]), 1);
    }
    for (int s = 0; s < C::NW; ++s) {
      ptx::mbar_init(ptx::smem_u32(&wfull_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&wempty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&tfull_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&tempty_bar[s]), 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int hw = p.H * p.W;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: weights + centre-tap / residual pixels
    if (lane == 0) {
      ptx::prefetch_tmap(&p.tm_w);
      if (p.xc0 > 0) ptx::prefetch_tmap(&p.tm_x0);
      if (p.xc1 > 0) ptx::prefetch_tmap(&p.tm_x1);
      if (p.residual) {
        ptx::prefetch_tmap(&p.tm_res);
        ptx::prefetch_tmap(&p.tm_ident);
      }
      int ws = 0;
      uint32_t wph = 0;
      long long pseq = 0;  // global index of the next pixel buffer of this CTA
      auto load_w = [&](const CUtensorMap* tm, int c0, int c1) {
        ptx::mbar_wait(ptx::smem_u32(&wempty_bar[ws]), wph ^ 1);
        const uint32_t fb = ptx::smem_u32(&wfull_bar[ws]);
        ptx::mbar_arrive_expect_tx(fb, C::W_BYTES);
        ptx::tma_load_4d(wbase + ws * C::W_BYTES, tm, fb, c0, c1, 0, 0);
        if (++ws == C::NW) { ws = 0; wph ^= 1; }
      };
      auto load_p = [&](const CUtensorMap* tm, int ch, int h, int b) {
        const int ps = static_cast<int>(pseq % C::NP);
        const uint32_t pph = static_cast<uint32_t>((pseq / C::NP) & 1);
        ptx::mbar_wait(ptx::smem_u32(&pempty_bar[ps]), pph ^ 1);
        const uint32_t fb = ptx::smem_u32(&pfull_bar[ps]);
        ptx::mbar_arrive_expect_tx(fb, PX * BK * 2);
        ptx::tma_load_4d(pbase + ps * C::P_BYTES, tm, fb, ch, 0, h, b);
        ++pseq;
      };
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int tt = p.reverse ? p.num_tiles - 1 - t : t;
        const int pt = tt / p.n_tiles;
        const int m0 = pt * PX;
        const int n0 = (tt - pt * p.n_tiles) * 128;
        const int b0 = m0 / hw;
        const int h0 = (m0 - b0 * hw) / p.W;
        for (int kh = 0; kh < 3; ++kh)
          for (int cc = 0; cc < chunks; ++cc)
            for (int kw = 0; kw < 3; ++kw) load_w(&p.tm_w, ((kh * 3 + kw) * chunks + cc) * BK, n0);
        pseq += 3 * chunks;  // filled by the transform warps
        int kb = 9 * chunks;
        for (int cc = 0; cc < x_chunks; ++cc, ++kb) {
          const int ch = cc * BK;
          if (ch < p.xc0) load_p(&p.tm_x0, ch, h0, b0);
          else load_p(&p.tm_x1, ch - p.xc0, h0, b0);
          load_w(&p.tm_w, kb * BK, n0);
        }
        for (int j = 0; j < r_chunks; ++j) {
          load_p(&p.tm_res, n0 + j * BK, h0, b0);
          load_w(&p.tm_ident, j * BK, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (as in conv_gemm_tcH_kernel)
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128);
      int ps = 0, ws = 0;
      uint32_t pph = 0, wph = 0;
      uint32_t tl = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tl) {
        const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
        ptx::mbar_wait(ptx::smem_u32(&tempty_bar[as]), aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * PX;
        bool first = true;
        auto mma_block = [&](uint32_t pbuf, int rowpitch, int shift) {
          ptx::mbar_wait(ptx::smem_u32(&wfull_bar[ws]), wph);
          ptx::tc_fence_after();
          const uint64_t dw = ptx::umma_desc_k_sw128(wbase + ws * C::W_BYTES);
          const uint64_t d0 = ptx::umma_desc_k_sw128(pbuf + shift * 128);
          const uint64_t d1 = ptx::umma_desc_k_sw128(pbuf + (rowpitch + shift) * 128);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t acc = (first && k == 0) ? 0u : 1u;
            ptx::umma_bf16(tmem_acc, dw + 2 * k, d0 + 2 * k, idesc, acc);
            ptx::umma_bf16(tmem_acc + 128, dw + 2 * k, d1 + 2 * k, idesc, acc);
          }
          first = false;
          ptx::umma_commit(ptx::smem_u32(&wempty_bar[ws]));
          if (++ws == C::NW) { ws = 0; wph ^= 1; }
        };
        for (int kc = 0; kc < 3 * chunks; ++kc) {
          ptx::mbar_wait(ptx::smem_u32(&pfull_bar[ps]), pph);
          ptx::tc_fence_after();
          const uint32_t pbuf = pbase + ps * C::P_BYTES;
          for (int kw = 0; kw < 3; ++kw) mma_block(pbuf, C::P_ROWPITCH, kw);
          ptx::umma_commit(ptx::smem_u32(&pempty_bar[ps]));
          if (++ps == C::NP) { ps = 0; pph ^= 1; }
        }
        for (int kc = 0; kc < x_chunks + r_chunks; ++kc) {
          ptx::mbar_wait(ptx::smem_u32(&pfull_bar[ps]), pph);
          ptx::tc_fence_after();
          mma_block(pbase + ps * C::P_BYTES, 128, 0);
          ptx::umma_commit(ptx::smem_u32(&pempty_bar[ps]));
          if (++ps == C::NP) { ps = 0; pph ^= 1; }
        }
        ptx::umma_commit(ptx::smem_u32(&tfull_bar[as]));
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------ epilogue (warps 2..5): thread = channel
    const int q = warp & 3;
    epilogue_dispatch<PX, PX / 32, 0>(p, tmem_base, tfull_bar, tempty_bar, out_stage + q * (2 * 32 * 32 * 2), q, 0, 1, 0, lane);
  } else {
    // ------------------------------------------------------------ transform warps (6..13)
    const int grp = (warp - 6) >> 2;                       // groups of four warps
    const int tg = static_cast<int>(threadIdx.x) - NUM_THREADS - grp * 128;  // 0..127 within the group
    const int c = tg & 7;                                  // logical 16-byte chunk = channels 8 c .. 8 c + 7 of the slice
    const int r0 = tg >> 3;                                // first pixel row of this thread; then every 16th
    const int ctot = p.c0 + p.c1;
    constexpr int PASSES = (C::P_ROWS + 15) / 16;          // 17
    long long seq_base = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, seq_base += bufs_per_tile) {
      const int tt = p.reverse ? p.num_tiles - 1 - t : t;
      const int pt = tt / p.n_tiles;
      const int m0 = pt * PX;
      const int b0 = m0 / hw;
      const int h0 = (m0 - b0 * hw) / p.W;
      for (int idx = 0; idx < 3 * chunks; ++idx) {
        const long long seq = seq_base + idx;
        if (static_cast<int>(seq % C::T_GROUPS) != grp) continue;
        const int kh = idx / chunks, cc = idx - kh * chunks;
        const int ch = cc * BK;
        // this thread's 8 channels: pre-halved affine as packed bf16 pairs
        const float* sc = p.gn_scale + static_cast<long long>(b0) * ctot + ch + c * 8;
        const float* sh = p.gn_shift + static_cast<long long>(b0) * ctot + ch + c * 8;
        const __nv_bfloat16* src = (ch < p.c0) ? p.raw0 : p.raw1;
        const int cs = (ch < p.c0) ? p.c0 : p.c1;
        const int cof = ((ch < p.c0) ? ch : ch - p.c0) + c * 8;
        // All loads of the box go out BEFORE the wait for a free buffer: they land in registers, so their latency
        // overlaps the time the MMA warp still needs for the buffer being recycled.
        uint4 v[PASSES];
        unsigned okmask = 0;
#pragma unroll
        for (int i = 0; i < PASSES; ++i) {
          const int pr = r0 + 16 * i;                          // pixel row of the box: rr * 130 + (px + 1)
          const int rr = pr >= C::P_ROWPITCH ? 1 : 0;
          const int px = pr - rr * C::P_ROWPITCH - 1;
          const int hh = h0 + kh - 1 + rr;
          const bool ok = pr < C::P_ROWS && px >= 0 && px < p.W && hh >= 0 && hh < p.H;
          v[i] = make_uint4(0u, 0u, 0u, 0u);
          if (ok) {
            okmask |= 1u << i;
            if (!(p.hf_debug & 2))
              v[i] = __ldg(reinterpret_cast<const uint4*>(src + ((static_cast<long long>(b0) * p.H + hh) * p.W + px) * cs + cof));
          }
        }
        const float4 sc4[2] = {__ldg(reinterpret_cast<const float4*>(sc)), __ldg(reinterpret_cast<const float4*>(sc) + 1)};
        const float4 sh4[2] = {__ldg(reinterpret_cast<const float4*>(sh)), __ldg(reinterpret_cast<const float4*>(sh) + 1)};
        const int ps = static_cast<int>(seq % C::NP);
        const uint32_t pph = static_cast<uint32_t>((seq / C::NP) & 1);
        ptx::mbar_wait_relaxed(ptx::smem_u32(&pempty_bar[ps]), pph ^ 1);
        // (the affine is fetched AFTER the pixel loads went out and consumed only here: its latency overlaps theirs)
        uint32_t hs[4], hb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          hs[i] = ptx::pack_bf16x2(0.5f * (i & 1 ? sc4[i >> 1].z : sc4[i >> 1].x),
                                   0.5f * (i & 1 ? sc4[i >> 1].w : sc4[i >> 1].y));
          hb[i] = ptx::pack_bf16x2(0.5f * (i & 1 ? sh4[i >> 1].z : sh4[i >> 1].x), 0.5f * (i & 1 ? sh4[i >> 1].w : sh4[i >> 1].y));
        }
        const uint32_t pbuf = pbase + ps * C::P_BYTES;
        // arithmetic for ALL rows first (68 independent chains per thread: the MUFU / HFMA2 latencies overlap), then
        // the stores -- interleaving them row by row put a volatile store between the chains and left each warp with
        // the instruction-level parallelism of a single row (the transform then took ~3.5 us per box and starved the
        // MMA warp: r02_fused_gn_ab.txt)
#pragma unroll
        for (int i = 0; i < PASSES; ++i) {
          uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
          if (!(p.hf_debug & 1)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t h = bf16x2_fma(w[k], hs[k], hb[k]);
              w[k] = bf16x2_fma(h, bf16x2_tanh(h), h);
            }
          }
          const bool on = (okmask >> i) & 1u;  // halo pixels outside the image stay zero: padding follows the activation
          v[i] = make_uint4(on ? w[0] : 0u, on ? w[1] : 0u, on ? w[2] : 0u, on ? w[3] : 0u);
        }
#pragma unroll
        for (int i = 0; i < PASSES; ++i) {
          const int pr = r0 + 16 * i;
          if (pr < C::P_ROWS) {
            const uint32_t dst = pbuf + pr * 128 + ((c ^ (pr & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[i].x), "r"(v[i].y), "r"(v[i].z),
                         "r"(v[i].w)
                         : "memory");
          }
        }
        ptx::fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");  // the four warps of this group
        if (tg == 0) ptx::mbar_arrive(ptx::smem_u32(&pfull_bar[ps]));
      }
      // The centre-tap / residual buffers that follow are filled by the TMA warp.  This group still WAITS for the
      // release of the ones on its alternation: mbarrier waits are by phase parity, so a producer may never get two
      // revolutions of the ring ahead of the consumer -- skipping over a long TMA-owned stretch straight to the next
      // tile would let its next wait pass on a stale phase and overwrite a buffer the MMA warp has not read yet.
      for (int idx = 3 * chunks; idx < bufs_per_tile; ++idx) {
        const long long seq = seq_base + idx;
        if (static_cast<int>(seq % C::T_GROUPS) != grp) continue;
        ptx::mbar_wait_relaxed(ptx::smem_u32(&pempty_bar[seq % C::NP]), static_cast<uint32_t>((seq / C::NP) & 1) ^ 1);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------ host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    T2P_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    T2P_CHECK(f != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

struct TmapKey {
  const void* ptr;
  uint64_t d[4];
  uint32_t b[4];
  int swizzle;
  bool operator==(const TmapKey& o) const {
    if (ptr != o.ptr || swizzle != o.swizzle) return false;
    for (int i = 0; i < 4; ++i)
      if (d[i] != o.d[i] || b[i] != o.b[i]) return false;
    return true;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.ptr) * 0x9E3779B97F4A7C15ull + k.swizzle;
    for (int i = 0; i < 4; ++i) {
      h ^= (k.d[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
      h ^= (k.b[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    }
    return static_cast<size_t>(h);
  }
};

// bf16 tensor with dims d[0] (innermost, contiguous) .. d[3]; dense strides; box b[0..3]; 128-byte swizzle
// (operand tiles) or none (the epilogue's store tiles).
// swizzle: 0 none, 2 = 64-byte, 3 = 128-byte (operand tiles)
CUtensorMap make_tmap_bf16(const void* ptr, const uint64_t d[4], const uint32_t b[4], int swizzle = 3) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, {d[0], d[1], d[2], d[3]}, {b[0], b[1], b[2], b[3]}, swizzle};
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  CUtensorMap tm;
  cuuint64_t dims[4] = {d[0], d[1], d[2], d[3]};
  cuuint64_t strides[3] = {d[0] * 2, d[0] * d[1] * 2, d[0] * d[1] * d[2] * 2};
  cuuint32_t box[4] = {b[0], b[1], b[2], b[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  T2P_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base must be 16-byte aligned");
  T2P_CHECK((strides[0] & 15) == 0, "TMA row pitch must be a multiple of 16 bytes");
  CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B
                                        : (swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  T2P_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(int(r)));
  if (cache.size() > 65536) cache.clear();
  cache.emplace(key, tm);
  return tm;
}

// 128 x 128 bf16 identity: the channel-major kernel adds the residual through two identity k-blocks
const void* identity128() {  // one copy per device
  static void* dev_ptr[kMaxDevices] = {};
  void*& d = dev_ptr[current_device()];
  if (!d) {
    std::vector<__nv_bfloat16> h(128 * 128, __float2bfloat16(0.f));
    for (int i = 0; i < 128; ++i) h[i * 128 + i] = __float2bfloat16(1.f);
    T2P_CUDA(cudaMalloc(&d, h.size() * 2));
    T2P_CUDA(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  }
  return d;
}

int sm_count() { return device_sm_count(); }

// Programmatic dependent launch, decided per launch.  On every tcgen05 GEMM it was slower (round 1, T2P_PDL bit 1);
// on GEMM launches that leave SMs idle (fewer CTAs than SMs: the latency-bound layers at 16 x 16 and below) the next
// kernel's set-up -- barrier init, TMEM allocation, descriptor prefetch -- runs under the tail of its predecessor:
// neutral at 64 maps per GPU, -4.6 % per PC iteration at 8 (profiles/r02_gn_small_ab.txt).  T2P_PDL_SMALL (knob
// builds) overrides the CTA threshold.
bool pdl_for(int grid) {
  static const bool all = (env_knob("T2P_PDL", 0) & 1) != 0;
  static const int small = env_knob("T2P_PDL_SMALL", 147);
  return all || grid <= small;
}

template <int BN>
void launch(TcParams& p, cudaStream_t st) {
  using C = Cfg<BN>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    T2P_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::SMEM_BYTES));
  }
  p.n_tiles = cdiv(p.N, BN);
  p.num_tiles = cdiv(p.M, BM) * p.n_tiles;
  const int grid = std::min(p.num_tiles, sm_count());
  launch_pdl_dyn(pdl_for(grid), conv_gemm_tc_kernel<BN>, dim3(grid), dim3(NUM_THREADS), C::SMEM_BYTES, st, p);
}

template <int PX>
void launch_t(TcParams& p, cudaStream_t st) {
  using C = CfgT<PX>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    T2P_CUDA(cudaFuncSetAttribute(conv_gemm_tcT_kernel<PX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::SMEM_BYTES));
  }
  p.n_tiles = cdiv(p.N, 128);
  p.num_tiles = cdiv(p.M, PX) * p.n_tiles;
  const int grid = std::min(p.num_tiles * p.splits, sm_count());
#ifdef T2P_TIMING_KNOBS
  // T2P_TRACE_T=1 (knob builds): time line of CTA 0 of every launch of at most one wave, printed per launch
  static const bool trace = env_knob("T2P_TRACE_T", 0) != 0;
  static long long* tbuf = nullptr;
  if (trace && p.num_tiles * p.splits <= sm_count()) {
    if (!tbuf) T2P_CUDA(cudaMalloc(reinterpret_cast<void**>(&tbuf), 16 * sizeof(long long)));
    T2P_CUDA(cudaMemsetAsync(tbuf, 0, 16 * sizeof(long long), st));
    p.trace = tbuf;
  }
#endif
  launch_pdl_dyn(pdl_for(grid), conv_gemm_tcT_kernel<PX>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, p);
#ifdef T2P_TIMING_KNOBS
  if (p.trace) {
    T2P_CUDA(cudaStreamSynchronize(st));
    long long h[16];
    T2P_CUDA(cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost));
    const int kb = (p.taps * (p.c0 + p.c1) + p.xc0 + p.xc1) / BK + (p.residual ? 2 : 0);
    std::fprintf(stderr, "[tcT trace] PX=%d M=%d N=%d kb=%d splits=%d grid=%d stats=%d gno=%d | prologue %lld | pdl wait %lld | first data %lld | last data %lld | "
                 "commit %lld | epilogue sees acc %lld | epilogue done %lld | stores read %lld | exit %lld (clocks from entry)\n",
                 PX, p.M, p.N, kb, p.splits, grid, p.stat_part ? 1 : 0, p.gno_gamma ? 1 : 0, h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0],
                 h[5] - h[0], h[6] - h[0], h[7] - h[0], h[8] - h[0], h[9] - h[0]);
  }
#endif
}

// CTA-pair kernel: grid = 2 x (clusters that fit the device at once), even number of tiles
void launch_p(TcParams& p, cudaStream_t st) {
  using C = CfgP;
  static bool configured[kMaxDevices] = {};
  static int max_clusters[kMaxDevices] = {};
  const int dev = current_device();
  if (first_use_on_device(configured)) {
    T2P_CUDA(cudaFuncSetAttribute(conv_gemm_tcP_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * 74);
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    T2P_CUDA(cudaOccupancyMaxActiveClusters(&n, conv_gemm_tcP_kernel, &cfg));
    max_clusters[dev] = std::max(1, n);
  }
  p.n_tiles = cdiv(p.N, 128);
  p.num_tiles = cdiv(p.M, C::PX) * p.n_tiles;
  const int grid = 2 * std::min(p.num_tiles / 2, max_clusters[dev]);
  launch_pdl_dyn(pdl_for(grid), conv_gemm_tcP_kernel, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, p);
}

template <int EW, bool GND>
void launch_h_ew(TcParams& p, cudaStream_t st) {
  using C = CfgHT<EW, GND>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    T2P_CUDA(cudaFuncSetAttribute(conv_gemm_tcH_kernel<EW, GND>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  p.n_tiles = cdiv(p.N, 128);
  p.num_tiles = cdiv(p.M, C::PX) * p.n_tiles;
  const int grid = std::min(p.num_tiles, sm_count());
  launch_pdl<1>(conv_gemm_tcH_kernel<EW, GND>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, p);
}

// Form of the output GroupNorm in the halo kernel.  Deferred (fix-up warps) when the CTAs work through several waves of
// tiles: the in-accumulator form then pays the slowest CTA of a sample in every wave.  In-accumulator for launches of a
// wave or two (small batches), where the deferred form's extra pass over the last tile is exposed.  Measured, cfg2,
// ms per PC iteration (profiles/r02_gn_out_ab.txt): B = 64 off 22.95, deferred 22.36, in-accumulator 22.57; B = 8 off 6.29,
// 6.23, 6.15.
bool gno_h_deferred(long long num_tiles) {
  static const int knob = env_knob("T2P_GNO_H_DEFER", -1);  // (knob builds: A/B)
  if (knob >= 0) return knob != 0;
  return num_tiles > 2LL * sm_count();
}

void launch_h(TcParams& p, cudaStream_t st) {
  if (p.gno_gamma && gno_h_deferred(cdiv(p.M, CfgH::PX) * cdiv(p.N, 128))) launch_h_ew<4, true>(p, st);
  else launch_h_ew<CfgH::EPI_WARPS, false>(p, st);
}

void launch_hf(TcParams& p, cudaStream_t st) {
  using C = CfgHF;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    T2P_CUDA(cudaFuncSetAttribute(conv_gemm_tcHF_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  p.n_tiles = cdiv(p.N, 128);
  p.num_tiles = cdiv(p.M, C::PX) * p.n_tiles;
  const int grid = std::min(p.num_tiles, sm_count());
  launch_pdl<1>(conv_gemm_tcHF_kernel, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, p);
}

// `rows` consecutive pixels (b, h, w raster order) as one TMA box (tw, th, tb) of the NHWC tensor
bool pixel_box(const ConvGemmArgs& a, int rows, uint32_t& tw, uint32_t& th, uint32_t& tb) {
  if (a.ksize == 1) { tw = rows; th = 1; tb = 1; return true; }
  if (a.W >= rows) {
    if (a.W % rows) return false;
    tw = rows; th = 1; tb = 1;
    return true;
  }
  if (rows % a.W) return false;
  tw = a.W;
  th = std::min<uint32_t>(a.H, rows / a.W);
  if (a.H % th) return false;
  tb = rows / (tw * th);
  return tb == 1 || th == static_cast<uint32_t>(a.H);  // a tile spanning samples must cover whole images
}

struct Plan {
  bool halo = false;  // channel-major halo kernel (3x3, W == 128, two image rows per tile)
  bool channel_major = false;
  int rows = BM;  // pixels per tile
  int bn = 128;   // pixel-major kernel: channels per tile
  uint32_t tw = 0, th = 0, tb = 0;
  bool stats_ok = false;
};

Plan make_plan(const ConvGemmArgs& a) {
  Plan pl;
  const long long M = static_cast<long long>(a.B) * a.H * a.W;
  const bool want_stats = a.stat_part != nullptr;
  if (a.N >= 128 && a.N % 8 == 0 && a.out_dtype == kBF16 && !a.out_nchw && !a.res_up) {
    // channel-major kernel: widest pixel tile that still spreads the problem over (most of) the SMs
    int pick = 0;
    for (int px : {256, 128, 64}) {
      uint32_t tw, th, tb;
      if (!pixel_box(a, px, tw, th, tb)) continue;
      if (want_stats && (a.rows_per_sample % px) != 0) continue;
      pick = px;
      if (cdiv64(M, px) * cdiv(a.N, 128) >= 96) break;
    }
    if (pick) {
      pl.channel_major = true;
      pl.rows = pick;
      pixel_box(a, pick, pl.tw, pl.th, pl.tb);
      static const bool halo_on = env_knob("T2P_HALO", 1) != 0;
      pl.halo = halo_on && pick == 256 && a.ksize == 3 && a.W == 128 && a.H % 2 == 0;
      pl.stats_ok = want_stats && a.out_dtype == kBF16;
      return pl;
    }
  }
  T2P_CHECK(pixel_box(a, BM, pl.tw, pl.th, pl.tb), "unsupported image geometry for the 128-pixel tile");
  pl.rows = BM;
  const int m_tiles = static_cast<int>(cdiv64(M, BM));
  if (a.N <= 16) pl.bn = 16;
  else if (a.N <= 32) pl.bn = 32;
  else if (a.N <= 64) pl.bn = 64;
  else {
    pl.bn = (a.N % 256 == 0) ? 256 : 128;
    while (pl.bn > 32 && m_tiles * cdiv(a.N, pl.bn) < 96) pl.bn >>= 1;
  }
  pl.stats_ok = want_stats && a.rows_per_sample > 0 && a.rows_per_sample % BM == 0 && a.out_dtype == kBF16 && a.N >= 32;
  return pl;
}

}  // namespace

// K splits of a channel-major launch with few tiles and long K (>= 32 k-blocks: the 3x3 convolutions at 16 x 16 and
// below): as many as keep the (tile, split) items within one wave of CTAs, at least four k-blocks each, at most eight
// (the last arrival adds the partials two at a time).  Measured (profiles/r02_splitk_ab.txt): such a kernel is mostly
// fixed cost -- a 4-k-block launch takes 12 us under ncu, a 36-k-block one 20 us, 72 k-blocks 33 us -- so splitting
// buys 3 us (K = 2304) to 13 us (K = 4608) per launch: -2.8 % per PC iteration at 8 maps per GPU, nothing at 64.
static int plan_splits(const ConvGemmArgs& a, const Plan& pl) {
  static const bool on = env_knob("T2P_SPLITK", 1) != 0;  // (knob builds: A/B)
  if (!on || !pl.channel_major || pl.halo || a.gn_scale) return 1;
  const long long tiles = cdiv64(static_cast<long long>(a.B) * a.H * a.W, pl.rows) * cdiv(a.N, 128);
  const int num_kb = (a.ksize * a.ksize * (a.c0 + a.c1) + a.xc0 + a.xc1) / BK + (a.residual ? 2 : 0);
  const long long cap = std::min<long long>(sm_count(), kSplitKMaxItems);
  static const int min_kb = env_knob("T2P_SPLITK_MINKB", 32);  // (knob builds: A/B of the threshold)
  if (tiles * 2 > cap || num_kb < min_kb) return 1;
  const int s = static_cast<int>(std::min<long long>({8, cap / tiles, num_kb / 4}));
  return s >= 2 ? s : 1;
}

int conv_gemm_tc_splits(const ConvGemmArgs& a) {
  ConvGemmArgs q = a;
  q.gno_gamma = nullptr;
  return plan_splits(q, make_plan(q));
}

bool conv_gemm_tc_channel_major(const ConvGemmArgs& a) { return make_plan(a).channel_major; }

bool conv_gemm_tc_fuses_gn(const ConvGemmArgs& a) { return make_plan(a).halo; }

// parts (statistic slices) per sample and channel quadrant: pixel tiles per sample x epilogue warps per quadrant
static int gn_out_parts(const ConvGemmArgs& a, const Plan& pl) {
  const long long tiles = cdiv64(static_cast<long long>(a.B) * a.H * a.W, pl.rows) * cdiv(a.N, 128);
  const int warps_per_quadrant = pl.halo ? (CfgH::EPI_WARPS == 8 && !gno_h_deferred(tiles) ? 2 : 1) : 2;
  return a.rows_per_sample / pl.rows * warps_per_quadrant;
}

bool conv_gemm_tc_gn_out_ok(const ConvGemmArgs& a, int groups) {
  if (groups <= 0 || a.N % 128 != 0 || a.N % groups != 0) return false;
  const int cpg = a.N / groups;
  if (cpg != 4 && cpg != 8 && cpg != 16 && cpg != 32) return false;
  if (a.out_dtype != kBF16 || a.out_nchw || a.residual || a.res_up || a.rows_per_sample <= 0) return false;
  ConvGemmArgs q = a;
  q.stat_part = nullptr;
  const Plan pl = make_plan(q);
  if (!pl.channel_major || a.rows_per_sample % pl.rows != 0) return false;
  if (plan_splits(q, pl) > 1) return false;  // launches of few tiles split K instead (and keep the one-launch GroupNorm)
  {
    // which launches normalise their output (knob builds: T2P_GNO_POLICY bit mask for the A/B): 1 = 128-pixel-wide images
    // with K < 2304, 2 = with K >= 2304, 4 = 64 x 64 images, 8 = 32 x 32 and below
    static const int policy = env_knob("T2P_GNO_POLICY", 15);
    const int ktot = a.ksize * a.ksize * (a.c0 + a.c1) + a.xc0 + a.xc1;
    const int cls = pl.halo ? (ktot < 2304 ? 1 : 2) : (a.rows_per_sample >= 4096 ? 4 : 8);
    if (!(policy & cls)) return false;
  }
  // deadlock freedom of the per-sample wait (see epilogue_role_gn): the tiles of one sample must span fewer
  // consecutive tile indices than there are CTAs
  const long long span = static_cast<long long>(a.rows_per_sample / pl.rows) * (a.N / 128);
  if (span <= sm_count()) return true;
  // (the deferred form never waits in an epilogue: any span will do)
  return pl.halo && gno_h_deferred(cdiv64(static_cast<long long>(a.B) * a.H * a.W, pl.rows) * cdiv(a.N, 128));
}

long long conv_gemm_tc_gn_out_part_floats(const ConvGemmArgs& a, int groups) {
  ConvGemmArgs q = a;
  q.stat_part = nullptr;
  const Plan pl = make_plan(q);
  return 2LL * a.B * gn_out_parts(a, pl) * groups;
}

long long conv_gemm_tc_gn_out_flag_ints(const ConvGemmArgs& a) { return 1LL * a.B * cdiv(a.N, 128) * 4; }

int conv_gemm_tc_stat_tile(const ConvGemmArgs& a) {
  ConvGemmArgs q = a;
  if (!q.stat_part) q.stat_part = reinterpret_cast<float*>(16);  // plan as if statistics were requested
  const Plan pl = make_plan(q);
  if (!pl.stats_ok) return 0;
  // the channel-major kernel's epilogue warps each own half a pixel tile; the halo kernel's own a whole one
  if (pl.halo) return (CfgH::EPI_WARPS == 8 && !a.gn_scale) ? pl.rows / 2 : pl.rows;  // (the fused variant: four warps)
  return pl.channel_major ? pl.rows / 2 : pl.rows;
}

void conv_gemm_tc(const ConvGemmArgs& a, cudaStream_t st) {
  T2P_CHECK(a.ksize == 1 || a.ksize == 3, "ksize must be 1 or 3");
  T2P_CHECK(a.c0 > 0 && a.c0 % BK == 0 && a.c1 % BK == 0 && a.xc0 % BK == 0 && a.xc1 % BK == 0,
            "channel counts must be multiples of 64");
  T2P_CHECK((a.xc0 == 0 || a.x0) && (a.xc1 == 0 || (a.x1 && a.xc0 > 0)), "bad extra sources");
  T2P_CHECK(a.out_dtype == kBF16 || a.out_dtype == kF32, "out dtype must be bf16 or fp32");
  const int ctot = a.c0 + a.c1;
  const int taps = a.ksize * a.ksize;
  const long long M = static_cast<long long>(a.B) * a.H * a.W;
  T2P_CHECK(M > 0 && M < (1ll << 31), "M out of range");

  TcParams p{};
  p.M = static_cast<int>(M);
  p.N = a.N;
  p.c0 = a.c0;
  p.c1 = a.c1;
  p.xc0 = a.xc0;
  p.xc1 = a.xc1;
  p.taps = taps;
  p.H = a.H;
  p.W = a.W;
  p.rows_per_sample = a.rows_per_sample;
  p.bias = a.bias;
  p.rowbias = a.rowbias;
  p.residual = a.residual;
  p.res_up = a.res_up;
  p.alpha = a.alpha;
  p.out = a.out;
  p.out_fp32 = (a.out_dtype == kF32);
  p.stat_part = a.stat_part;
  p.rowbias_ld = a.rowbias_ld > 0 ? a.rowbias_ld : a.N;
  p.out_nchw = a.out_nchw;
  p.reverse = a.reverse;
  p.mode2d = (a.ksize == 1) ? 1 : 0;
  if (a.out_nchw) T2P_CHECK(a.out_dtype == kF32 && a.residual == nullptr, "out_nchw is fp32-only, without residual");
  if (a.rowbias) T2P_CHECK(a.rows_per_sample > 0, "rows_per_sample required");
  if (a.res_up) T2P_CHECK(a.ksize == 3 && (a.H % 2 == 0) && (a.W % 2 == 0), "res_up needs an even image");

  Plan pl = make_plan(a);
  if (a.xc0 > 0) T2P_CHECK(pl.channel_major, "centre-tap sources need the channel-major kernel (N >= 128, bf16 out)");
  // CTA-pair kernel (conv_gemm_tcP_kernel): 256-pixel tiles of layers with a multiple of 256 output channels; each CTA
  // stages HALF of the pixel tile, so the pixel boxes are those of a 128-pixel tile
  bool pair = false;
  {
    static const bool pair_on = env_knob("T2P_PAIR", 1) != 0;  // (knob builds: A/B)
    uint32_t tw, th, tb;
    if (pair_on && pl.channel_major && !pl.halo && pl.rows == 256 && a.N % 256 == 0 && !a.gn_scale &&
        M % 256 == 0 && M / 256 * (a.N / 128) >= 2 && pixel_box(a, 128, tw, th, tb) &&
        !(a.sk_part && a.sk_ticket && plan_splits(a, pl) > 1)) {
      pair = true;
      pl.tw = tw; pl.th = th; pl.tb = tb;
    }
  }
  const int box_rows = pair ? 128 : pl.rows;
  if (a.stat_part)
    T2P_CHECK(pl.stats_ok, "fused GroupNorm statistics need whole pixel tiles per sample and bf16 output "
                           "(ask conv_gemm_tc_stat_tile first)");

  auto amap = [&](const void* ptr, int c) {
    if (p.mode2d) {
      uint64_t d[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(M), 1, 1};
      uint32_t b[4] = {BK, static_cast<uint32_t>(box_rows), 1, 1};
      return make_tmap_bf16(ptr, d, b);
    }
    uint64_t d[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(a.W), static_cast<uint64_t>(a.H),
                     static_cast<uint64_t>(a.B)};
    uint32_t b[4] = {BK, pl.tw, pl.th, pl.tb};
    return make_tmap_bf16(ptr, d, b);
  };
  auto amap_halo = [&](const void* ptr, int c) {  // 2 image rows x (W + 2) pixels
    uint64_t d[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(a.W), static_cast<uint64_t>(a.H),
                     static_cast<uint64_t>(a.B)};
    uint32_t b[4] = {BK, static_cast<uint32_t>(CfgH::P_ROWPITCH), 2, 1};
    return make_tmap_bf16(ptr, d, b);
  };
  p.tm_a0 = pl.halo ? amap_halo(a.a0, a.c0) : amap(a.a0, a.c0);
  p.tm_a1 = (a.c1 > 0) ? (pl.halo ? amap_halo(a.a1, a.c1) : amap(a.a1, a.c1)) : p.tm_a0;
  {
    uint64_t d[4] = {static_cast<uint64_t>(taps) * ctot + a.xc0 + a.xc1, static_cast<uint64_t>(a.N), 1, 1};
    uint32_t b[4] = {BK, static_cast<uint32_t>(pl.channel_major ? 128 : pl.bn), 1, 1};
    p.tm_w = make_tmap_bf16(a.w, d, b);
  }
  if (a.gno_gamma) {
    T2P_CHECK(conv_gemm_tc_gn_out_ok(a, a.gno_groups) && a.gno_beta && a.gno_part && a.gno_flags &&
                  !a.stat_part && !a.gn_scale,
              "this launch cannot normalise its own output (ask conv_gemm_tc_gn_out_ok first)");
    p.gno_gamma = a.gno_gamma;
    p.gno_beta = a.gno_beta;
    p.gno_cpg = a.N / a.gno_groups;
    p.gno_eps = a.gno_eps;
    p.gno_part = static_cast<unsigned long long*>(a.gno_part);
    p.gno_flags = a.gno_flags;
    p.gno_parts = gn_out_parts(a, pl);
    p.gno_slots = a.B * (a.N / 128) * 4;
    p.gno_debug = env_knob("T2P_GNO_DEBUG", 0);
  }
  p.pf_ptr = static_cast<const char*>(a.l2_prefetch);
  p.pf_bytes = a.l2_prefetch ? (a.l2_prefetch_bytes & ~15LL) : 0;
  p.splits = 1;
  if (a.sk_part && a.sk_ticket && !a.gno_gamma) {
    p.splits = plan_splits(a, pl);
    p.sk_part = a.sk_part;
    p.sk_ticket = a.sk_ticket;
  }
  if (pl.channel_major) {
    {
      uint64_t d[4] = {static_cast<uint64_t>(a.N), static_cast<uint64_t>(M), 1, 1};
      uint32_t b[4] = {32, 32, 1, 1};
      p.tm_out = make_tmap_bf16(a.out, d, b, 2);  // 64-byte swizzle: the epilogue's staging layout
    }
    if (a.xc0 > 0) p.tm_x0 = amap(a.x0, a.xc0);
    if (a.xc1 > 0) p.tm_x1 = amap(a.x1, a.xc1);
    if (a.residual) {
      p.tm_res = amap(a.residual, a.N);
      uint64_t d[4] = {128, 128, 1, 1};
      uint32_t b[4] = {BK, 128, 1, 1};
      p.tm_ident = make_tmap_bf16(identity128(), d, b);
    }
    if (a.gn_scale) {
      T2P_CHECK(pl.halo && a.gn_shift, "fused GroupNorm needs the halo kernel (ask conv_gemm_tc_fuses_gn first)");
      p.raw0 = static_cast<const __nv_bfloat16*>(a.a0);
      p.raw1 = static_cast<const __nv_bfloat16*>(a.a1);
      p.gn_scale = a.gn_scale;
      p.gn_shift = a.gn_shift;
      p.hf_debug = env_knob("T2P_HF_DEBUG", 0);
      launch_hf(p, st);
      return;
    }
    if (pl.halo) {
      launch_h(p, st);
      return;
    }
    if (pair) {
      launch_p(p, st);
      return;
    }
    switch (pl.rows) {
      case 256: launch_t<256>(p, st); break;
      case 128: launch_t<128>(p, st); break;
      default: launch_t<64>(p, st); break;
    }
    return;
  }
  switch (pl.bn) {
    case 16: launch<16>(p, st); break;
    case 32: launch<32>(p, st); break;
    case 64: launch<64>(p, st); break;
    case 128: launch<128>(p, st); break;
    default: launch<256>(p, st); break;
  }
}

}  // namespace t2p
