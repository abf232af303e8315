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
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"

namespace t2p {

namespace {

constexpr int BM = 128;      // rows (pixels) per tile == UMMA M
constexpr int BK = 64;       // bf16 elements per 128-byte swizzled row
constexpr int UMMA_K = 16;   // K per tcgen05.mma for 16-bit inputs
constexpr int NUM_THREADS = 192;

struct TcParams {
  CUtensorMap tm_a0;
  CUtensorMap tm_a1;
  CUtensorMap tm_w;
  int M, N;
  int c0, c1;
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
  int debug_mode;  // 0 = normal; 1..3 = bring-up bisection (see T2P_TC_DEBUG)
};

template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN <= 128) ? 3 : 4;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS) conv_gemm_tc_kernel(const __grid_constant__ TcParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[C::STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float2 stat_sm[4][BN >= 32 ? BN : 32];
  __shared__ float bias_sm[BN >= 32 ? BN : 32];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tiles = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;

  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int ctot = p.c0 + p.c1;
  const int chunks_per_tap = ctot / BK;
  const int num_kb = p.taps * chunks_per_tap;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(&accum_bar), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;

  if (p.debug_mode == 1) {
    // alloc / dealloc only
  } else if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::prefetch_tmap(&p.tm_a0);
      ptx::prefetch_tmap(&p.tm_w);
      if (p.c1 > 0) ptx::prefetch_tmap(&p.tm_a1);
      int b0 = 0, h0 = 0, w0 = m0;
      if (!p.mode2d) {
        const int hw = p.H * p.W;
        b0 = m0 / hw;
        const int rem = m0 - b0 * hw;
        h0 = rem / p.W;
        w0 = rem - h0 * p.W;
      }
      const int pad = (p.taps == 9) ? 1 : 0;
      int kb = 0;
      for (int tap = 0; tap < p.taps; ++tap) {
        const int kh = (p.taps == 9) ? tap / 3 : 0;
        const int kw = (p.taps == 9) ? tap - kh * 3 : 0;
        for (int cc = 0; cc < chunks_per_tap; ++cc, ++kb) {
          const int s = kb % C::STAGES;
          const uint32_t ph = (kb / C::STAGES) & 1;
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
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        ptx::mbar_wait(ptx::smem_u32(&full_bar[s]), ph);
        ptx::tc_fence_after();
        const uint32_t sa = tiles + s * C::STAGE_BYTES;
        const uint32_t sb = sa + C::A_BYTES;
        const uint64_t da = ptx::umma_desc_k_sw128(sa);
        const uint64_t db = ptx::umma_desc_k_sw128(sb);
        if (p.debug_mode == 2) {
          ptx::mbar_arrive(ptx::smem_u32(&empty_bar[s]));
          continue;
        }
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advancing 16 bf16 = 32 B inside the swizzle atom: +2 in the (addr >> 4) field
          ptx::umma_bf16(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        ptx::umma_commit(ptx::smem_u32(&empty_bar[s]));  // frees the smem slot when the MMAs retire
      }
      if (p.debug_mode == 2) ptx::mbar_arrive(ptx::smem_u32(&accum_bar));
      else ptx::umma_commit(ptx::smem_u32(&accum_bar));  // accumulator complete
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < p.M;
    const int sample = (p.rows_per_sample > 0) ? (m / p.rows_per_sample) : 0;
    // While the main loop runs, stage the per-column bias (+ the per-sample time-embedding bias when the
    // whole tile belongs to one sample) in shared memory: the epilogue then never waits on global loads.
    const bool rb_uniform = p.rowbias != nullptr && (p.rows_per_sample % BM) == 0;
    {
      const int e = (warp - 2) * 32 + lane;
      const float* rb0 = rb_uniform ? p.rowbias + static_cast<long long>(m0 / p.rows_per_sample) * p.rowbias_ld : nullptr;
      for (int ch = e; ch < BN; ch += 128) {
        const int n = n0 + ch;
        float bv = 0.f;
        if (n < p.N) {
          if (p.bias) bv = __ldg(p.bias + n);
          if (rb0) bv += __ldg(rb0 + n);
        }
        bias_sm[ch] = bv;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    ptx::mbar_wait(ptx::smem_u32(&accum_bar), 0);
    ptx::tc_fence_after();
    long long res_row = m;
    if (p.res_up && row_ok) {
      const int hw = p.H * p.W;
      const int b = m / hw;
      const int rem = m - b * hw;
      const int h = rem / p.W;
      const int w = rem - h * p.W;
      res_row = (static_cast<long long>(b) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
    }
    constexpr int CH = (BN >= 32) ? 32 : 16;
    const int nchunks = (p.debug_mode >= 2) ? 0 : BN / CH;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      uint32_t r[32];
      const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c * CH;
      if constexpr (CH == 32) {
        ptx::tmem_ld_32x32(taddr, r);
      } else {
        uint32_t r16[16];
        ptx::tmem_ld_32x16(taddr, r16);
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = r16[i];
      }
      ptx::tmem_ld_wait();
      const int nb = n0 + c * CH;
      float v[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(r[i]);
      const bool full = (nb + CH <= p.N) && ((p.N & 7) == 0);
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] += bias_sm[c * CH + i];
        if (p.rowbias && !rb_uniform) {
          const float* rb = p.rowbias + static_cast<long long>(sample) * p.rowbias_ld + nb;
#pragma unroll
          for (int i = 0; i < CH; ++i)
            if (nb + i < p.N) v[i] += __ldg(rb + i);
        }
        if (p.residual) {
          if (p.out_fp32) {
            const float* rs = static_cast<const float*>(p.residual) + res_row * p.N + nb;
#pragma unroll
            for (int i = 0; i < CH; ++i)
              if (nb + i < p.N) v[i] += rs[i];
          } else if (full) {
            const uint4* rs = reinterpret_cast<const uint4*>(
                static_cast<const __nv_bfloat16*>(p.residual) + res_row * p.N + nb);
#pragma unroll
            for (int i = 0; i < CH / 8; ++i) {
              const uint4 t = rs[i];
              v[8 * i + 0] += bf16_lo(t.x); v[8 * i + 1] += bf16_hi(t.x);
              v[8 * i + 2] += bf16_lo(t.y); v[8 * i + 3] += bf16_hi(t.y);
              v[8 * i + 4] += bf16_lo(t.z); v[8 * i + 5] += bf16_hi(t.z);
              v[8 * i + 6] += bf16_lo(t.w); v[8 * i + 7] += bf16_hi(t.w);
            }
          } else {
            const __nv_bfloat16* rs = static_cast<const __nv_bfloat16*>(p.residual) + res_row * p.N + nb;
#pragma unroll
            for (int i = 0; i < CH; ++i)
              if (nb + i < p.N) v[i] += __bfloat162float(rs[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] *= p.alpha;
        if (p.out_fp32) {
          float* o = static_cast<float*>(p.out) + static_cast<long long>(m) * p.N + nb;
          if (p.out_nchw) {
            const int hw = p.H * p.W;
            const int b = m / hw;
            const int pix = m - b * hw;
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
          __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(m) * p.N + nb;
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
            if (!row_ok) v[i] = 0.f;
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
          stat_sm[q][c * 32 + lane] = make_float2(v[0], sq[0]);
        }
      }
    }
    if constexpr (CH == 32) {
      if (p.stat_part) {
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
        const int e = (warp - 2) * 32 + lane;
        for (int ch = e; ch < BN; ch += 128) {
          if (n0 + ch < p.N) {
            // fixed order -> run-to-run deterministic statistics
            const float2 a0 = stat_sm[0][ch], a1 = stat_sm[1][ch], a2 = stat_sm[2][ch], a3 = stat_sm[3][ch];
            float* dst = p.stat_part + (static_cast<long long>(blockIdx.x) * p.N + n0 + ch) * 2;
            dst[0] = (a0.x + a1.x) + (a2.x + a3.x);
            dst[1] = (a0.y + a1.y) + (a2.y + a3.y);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_acc, C::TMEM_COLS);
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
  bool operator==(const TmapKey& o) const {
    if (ptr != o.ptr) return false;
    for (int i = 0; i < 4; ++i)
      if (d[i] != o.d[i] || b[i] != o.b[i]) return false;
    return true;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 4; ++i) {
      h ^= (k.d[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
      h ^= (k.b[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    }
    return static_cast<size_t>(h);
  }
};

// bf16 tensor with dims d[0] (innermost, contiguous) .. d[3]; dense strides; box b[0..3].
CUtensorMap make_tmap_bf16(const void* ptr, const uint64_t d[4], const uint32_t b[4]) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, {d[0], d[1], d[2], d[3]}, {b[0], b[1], b[2], b[3]}};
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
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  T2P_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(int(r)));
  if (cache.size() > 65536) cache.clear();
  cache.emplace(key, tm);
  return tm;
}

template <int BN>
void launch(const TcParams& p, cudaStream_t st) {
  using C = Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    T2P_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::SMEM_BYTES));
    configured = true;
  }
  dim3 grid(cdiv(p.M, BM), cdiv(p.N, BN));
  conv_gemm_tc_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(p);
  T2P_LAUNCH_CHECK();
}

}  // namespace

void conv_gemm_tc(const ConvGemmArgs& a, cudaStream_t st) {
  T2P_CHECK(a.ksize == 1 || a.ksize == 3, "ksize must be 1 or 3");
  T2P_CHECK(a.c0 > 0 && a.c0 % BK == 0 && a.c1 % BK == 0, "channel counts must be multiples of 64");
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
  if (a.out_nchw) T2P_CHECK(a.out_dtype == kF32 && a.residual == nullptr, "out_nchw is fp32-only, without residual");
  {
    static const int dbg = [] { const char* e = getenv("T2P_TC_DEBUG"); return e ? atoi(e) : 0; }();
    p.debug_mode = dbg;
  }
  if (a.rowbias) T2P_CHECK(a.rows_per_sample > 0, "rows_per_sample required");
  if (a.stat_part)
    T2P_CHECK(a.rows_per_sample > 0 && a.rows_per_sample % BM == 0 && a.out_dtype == kBF16 && a.N >= 32,
              "fused GroupNorm statistics need whole 128-row tiles per sample, bf16 output and N >= 32");
  if (a.res_up) T2P_CHECK(a.ksize == 3 && (a.H % 2 == 0) && (a.W % 2 == 0), "res_up needs an even image");

  // M-tile = box of 128 pixels in (b, h, w) raster order
  uint32_t tw, th, tb;
  if (a.ksize == 1) {
    p.mode2d = 1;
    tw = BM; th = 1; tb = 1;
  } else {
    p.mode2d = 0;
    if (a.W >= BM) {
      T2P_CHECK(a.W % BM == 0, "W must be a multiple of 128 when >= 128");
      tw = BM; th = 1; tb = 1;
    } else {
      T2P_CHECK(BM % a.W == 0, "W must divide 128");
      tw = a.W;
      th = std::min<uint32_t>(a.H, BM / a.W);
      T2P_CHECK(a.H % th == 0, "H must be a multiple of the tile height");
      tb = BM / (tw * th);
      T2P_CHECK(tb == 1 || th == static_cast<uint32_t>(a.H), "tile must cover whole images when spanning samples");
    }
  }
  auto amap = [&](const void* ptr, int c) {
    if (p.mode2d) {
      uint64_t d[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(M), 1, 1};
      uint32_t b[4] = {BK, BM, 1, 1};
      return make_tmap_bf16(ptr, d, b);
    }
    uint64_t d[4] = {static_cast<uint64_t>(c), static_cast<uint64_t>(a.W), static_cast<uint64_t>(a.H),
                     static_cast<uint64_t>(a.B)};
    uint32_t b[4] = {BK, tw, th, tb};
    return make_tmap_bf16(ptr, d, b);
  };
  p.tm_a0 = amap(a.a0, a.c0);
  p.tm_a1 = (a.c1 > 0) ? amap(a.a1, a.c1) : p.tm_a0;

  int bn;
  if (a.N <= 16) bn = 16;
  else if (a.N <= 32) bn = 32;
  else if (a.N <= 64) bn = 64;
  else if (a.N % 256 == 0 && cdiv(p.M, BM) * (a.N / 256) >= 148) bn = 256;
  else bn = 128;
  {
    uint64_t d[4] = {static_cast<uint64_t>(taps) * ctot, static_cast<uint64_t>(a.N), 1, 1};
    uint32_t b[4] = {BK, static_cast<uint32_t>(bn), 1, 1};
    p.tm_w = make_tmap_bf16(a.w, d, b);
  }
  switch (bn) {
    case 16: launch<16>(p, st); break;
    case 32: launch<32>(p, st); break;
    case 64: launch<64>(p, st); break;
    case 128: launch<128>(p, st); break;
    default: launch<256>(p, st); break;
  }
}

}  // namespace t2p
