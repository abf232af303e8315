// Flash-style attention on the sm_100a tensor cores: S = Q K^T and O = P V as tcgen05.mma with the score tile and
// the output accumulator in TMEM, operands staged by TMA.  Replaces softmax(q k^T) v of AttnBlockpp
// (score_sde_pytorch/models/layers.py:160-176: one head, d = C) and of CrossAttention, self and text-conditioned
// (model/attention.py:170-193: n_heads heads of d = C / n_heads; no mask is ever passed, padded text tokens are
// attended to).  The [T, Tk] score matrix never exists outside TMEM.
//
// One CTA = 128 query rows of one (sample, head[, 256-wide slice of the value dimension]).  Roles:
//   warp 0     TMA producer: per 128-key block the Q / K tiles in 64-channel slices (K loop of the score GEMM),
//              then the V tile in two 64-key slices (K loop of the output GEMM), through one ring of stages
//   warp 1     tcgen05.mma issuer (one elected thread) + TMEM allocation
//   warps 2-5  softmax: thread = query row = TMEM lane.  Reads its 128 fp32 scores with four tcgen05.ld in flight
//              (one wait), takes the running maximum, rescales the O accumulator in TMEM when the maximum moved
//              (tcgen05.ld / tcgen05.st), writes P as bf16 into a 128-byte-swizzled K-major shared-memory tile
//              (the A operand of the output GEMM) and finally normalises and stores O.
// Operand layouts: Q, K and P are K-major (the reduction index is contiguous); V is consumed exactly as it lies
// in memory, [key][channel] with the channel contiguous -- an MN-major B operand (instruction-descriptor bit 16),
// so no transposed copy of V is made.  Heads of 32 channels are 64-byte rows (SWIZZLE_64B tiles), wider heads use
// 64-channel slices of 128-byte rows (SWIZZLE_128B).
//
// Per 128-key block j:   MMA: S = Q K_j^T -> [s_full]     softmax: m, alpha; O *= alpha; P_j -> [p_full]
//                        MMA: O += P_j V_j ; S = Q K_{j+1}^T ...
// tcgen05.commit orders completion, so s_full of block j + 1 also says that O += P_j V_j has retired: P and S are
// single-buffered; overlap comes from two CTAs per SM (TMEM: 128 score columns + DV <= 128 output columns each).
#include <cuda.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>

#include "kernels.h"
#include "ptx.cuh"

namespace t2p {
namespace {

constexpr int BQ = 128;   // query rows per CTA == UMMA M
constexpr int BKEY = 128; // keys per block == UMMA N of the score GEMM
constexpr int THREADS = 192;

struct AttnTcParams {
  CUtensorMap tm_q, tm_k, tm_v;
  __nv_bfloat16* out;
  long long ldo;
  int Tq, Tk, heads, d, dv_splits;
  float scale_log2;  // scale * log2(e)
  int q_tiles, hy, items;  // work items = B x (heads * dv_splits) x q_tiles, query tile fastest
};

template <int D, int DV>
struct ACfg {
  static constexpr bool NARROW = (D == 32);            // 64-byte operand rows (SWIZZLE_64B)
  static constexpr int CH = NARROW ? 32 : 64;           // channels per K-loop slice of the score GEMM
  static constexpr int ROWB = CH * 2;                   // bytes per operand row
  static constexpr int QK_STAGE = 2 * 128 * ROWB;       // Q slice + K slice
  static constexpr int V_STAGE = 64 * DV * 2;           // 64 keys x DV channels
  static constexpr int STAGE = QK_STAGE > V_STAGE ? QK_STAGE : V_STAGE;
  static constexpr int TMEM_COLS = (BKEY + DV) <= 256 ? 256 : 512;
  static constexpr int NS = TMEM_COLS == 512 ? 4 : (STAGE <= 16384 ? 4 : 2);
  static constexpr int P_BYTES = BQ * BKEY * 2;         // 32 KB: two [128 rows][64 keys] SW128 tiles
  static constexpr int SMEM = NS * STAGE + P_BYTES + 1024;
  static constexpr int KSTEPS = CH / 16;                // tcgen05.mma K = 16 per instruction
};

// K-major operand tile, rows of ROWB bytes, swizzle = row size (64 or 128 bytes); 8-row groups are 8 * ROWB apart
template <int ROWB>
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((8 * ROWB) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(ROWB == 128 ? 2 : 4) << 61;  // SWIZZLE_128B : SWIZZLE_64B
  return d;
}
// MN-major operand (V as it lies in memory: [key][channel], channel contiguous): canonical layout
// ((8 | 4 x 16 B, n), (8 keys, k)) : ((1, LBO), (row, SBO)) -- a 64- (32-) channel group of 8 keys is one swizzle
// atom; SBO = next 8 keys, LBO = next channel group (one 64-key x 64-channel box = 8 KB further)
template <int ROWB>
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((64 * ROWB) >> 4) << 16;      // LBO: next channel group (box of 64 keys)
  d |= static_cast<uint64_t>((8 * ROWB) >> 4) << 32;       // SBO: next 8 keys
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(ROWB == 128 ? 2 : 4) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(128 >> 4) << 24);
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void tma_load_3d_to(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  ptx::tma_load_3d(dst, tm, bar, c0, c1, c2);
}

template <int D, int DV>
__global__ void __launch_bounds__(THREADS, (BKEY + DV) <= 256 ? 2 : 1) attention_tc_kernel(const __grid_constant__ AttnTcParams p) {
  pdl_trigger();  // (a GEMM launched behind this kernel with programmatic serialisation may start its prologue)
  using C = ACfg<D, DV>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C::NS], empty_bar[C::NS];
  __shared__ __align__(8) uint64_t sfull_bar, pfull_bar, ofull_bar, oempty_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t pbuf = ring + C::NS * C::STAGE;  // 1024-aligned: every stage size is a multiple of 1 KB

  const int nd = p.d / C::CH;                 // K-loop slices of the score GEMM
  const int nkb = (p.Tk + BKEY - 1) / BKEY;   // key blocks
  // Persistent: the CTA walks work items blockIdx.x, + gridDim.x, ...; barriers, TMEM and the TMA ring live across
  // items, so the producer prefetches the next item's first tiles while the softmax warps finish the current one,
  // and the set-up (barrier init, TMEM allocation, descriptor prefetch) is paid once per CTA instead of per tile.
  struct Item { int q0, b, qk_col0, v_col0; };
  auto item_of = [&](int it) {
    const int qt = it % p.q_tiles;
    const int r = it / p.q_tiles;
    const int hy = r % p.hy, b = r / p.hy;
    const int head = hy / p.dv_splits, dvs = hy - head * p.dv_splits;
    return Item{qt * BQ, b, head * p.d, head * p.d + dvs * DV};
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NS; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(&sfull_bar), 1);
    ptx::mbar_init(ptx::smem_u32(&pfull_bar), 4);  // one arrival per softmax warp
    ptx::mbar_init(ptx::smem_u32(&ofull_bar), 1);
    ptx::mbar_init(ptx::smem_u32(&oempty_bar), 4);  // the softmax warps have read O of the finished item
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_s = tmem;          // scores: columns [0, 128)
  const uint32_t tmem_o = tmem + BKEY;   // output accumulator: columns [128, 128 + DV)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::prefetch_tmap(&p.tm_q);
      ptx::prefetch_tmap(&p.tm_k);
      ptx::prefetch_tmap(&p.tm_v);
      int s = 0;
      uint32_t ph = 0;
      for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
      const Item w = item_of(it);
      const int q0 = w.q0, b = w.b, qk_col0 = w.qk_col0, v_col0 = w.v_col0;
      for (int j = 0; j < nkb; ++j) {
        for (int c = 0; c < nd; ++c) {  // Q slice (re-read per key block: it stays in L2) + K slice
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t fb = ptx::smem_u32(&full_bar[s]);
          ptx::mbar_arrive_expect_tx(fb, C::QK_STAGE);
          const uint32_t dst = ring + s * C::STAGE;
          tma_load_3d_to(dst, &p.tm_q, fb, qk_col0 + c * C::CH, q0, b);
          tma_load_3d_to(dst + 128 * C::ROWB, &p.tm_k, fb, qk_col0 + c * C::CH, j * BKEY, b);
          if (++s == C::NS) { s = 0; ph ^= 1; }
        }
        for (int kc = 0; kc < 2; ++kc) {  // V: 64 keys x DV channels, one box per 64- (32-) channel group
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t fb = ptx::smem_u32(&full_bar[s]);
          ptx::mbar_arrive_expect_tx(fb, C::V_STAGE);
          const uint32_t dst = ring + s * C::STAGE;
#pragma unroll
          for (int g = 0; g < DV / C::CH; ++g)
            tma_load_3d_to(dst + g * 64 * C::ROWB, &p.tm_v, fb, v_col0 + g * C::CH, j * BKEY + kc * 64, b);
          if (++s == C::NS) { s = 0; ph ^= 1; }
        }
      }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_bf16(BKEY, false);
      constexpr uint32_t idesc_o = idesc_bf16(DV, true);
      int s = 0;
      uint32_t ph = 0;
      uint32_t blk = 0, nit = 0;  // key blocks / items processed by this CTA (barrier phases)
      for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++nit) {
      for (int j = 0; j < nkb; ++j, ++blk) {
        // S = Q K_j^T.  (The softmax warps have finished reading the previous S: p_full of the block before.)
        for (int c = 0; c < nd; ++c) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[s]), ph);
          ptx::tc_fence_after();
          const uint32_t sq = ring + s * C::STAGE;
          const uint64_t dq = desc_kmajor<C::ROWB>(sq);
          const uint64_t dk = desc_kmajor<C::ROWB>(sq + 128 * C::ROWB);
#pragma unroll
          for (int k = 0; k < C::KSTEPS; ++k) ptx::umma_bf16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, (c | k) != 0);
          ptx::umma_commit(ptx::smem_u32(&empty_bar[s]));
          if (++s == C::NS) { s = 0; ph ^= 1; }
        }
        ptx::umma_commit(ptx::smem_u32(&sfull_bar));
        // O += P_j V_j once the softmax warps have written P_j (and rescaled O)
        ptx::mbar_wait(ptx::smem_u32(&pfull_bar), blk & 1);
        if (j == 0 && nit > 0) ptx::mbar_wait(ptx::smem_u32(&oempty_bar), (nit - 1) & 1);  // O of the last item was read
        ptx::tc_fence_after();
        for (int kc = 0; kc < 2; ++kc) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[s]), ph);
          ptx::tc_fence_after();
          const uint32_t sv = ring + s * C::STAGE;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 16 keys per instruction
            const uint64_t dp = ptx::umma_desc_k_sw128(pbuf + kc * (BQ * 128)) + 2 * k;
            const uint64_t dv = desc_mnmajor<C::ROWB>(sv + k * 16 * C::ROWB);
            ptx::umma_bf16(tmem_o, dp, dv, idesc_o, (j | kc | k) != 0);
          }
          ptx::umma_commit(ptx::smem_u32(&empty_bar[s]));
          if (++s == C::NS) { s = 0; ph ^= 1; }
        }
      }
      ptx::umma_commit(ptx::smem_u32(&ofull_bar));
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    const int qd = warp & 3;               // TMEM lane quadrant of this warp
    const int row = qd * 32 + lane;        // query row within the tile
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const float sl2 = p.scale_log2;
    uint32_t blk = 0, nit = 0;
    for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++nit) {
    const Item w = item_of(it);
    const int q0 = w.q0, b = w.b, v_col0 = w.v_col0;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nkb; ++j, ++blk) {
      const int valid = min(BKEY, p.Tk - j * BKEY);  // keys of this block that exist (TMA zero-fills the rest)
      ptx::mbar_wait_warp(ptx::smem_u32(&sfull_bar), blk & 1);
      ptx::tc_fence_after();
      // the whole score row (128 fp32) comes into registers with ONE wait: four tcgen05.ld in flight together
      uint32_t r[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::tmem_ld_32x32(tmem_s + lane_off + c * 32, r[c]);
      ptx::tmem_ld_wait();
      // row maximum of the raw scores (keys beyond Tk read as zero scores: masked out)
      float mx = m;
      if (valid == BKEY) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[c][i]));
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(r[c][i]));
      }
      // correction of what has been accumulated so far (skipped by a warp none of whose rows moved)
      if (j > 0) {
        const float alpha = ex2((m - mx) * sl2);
        l *= alpha;
        if (__any_sync(0xffffffffu, mx > m)) {
#pragma unroll 1
          for (int c = 0; c < DV / 32; ++c) {
            uint32_t o[32];
            ptx::tmem_ld_32x32(tmem_o + lane_off + c * 32, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(tmem_o + lane_off + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      m = mx;
      const float moff = -m * sl2;
      // P = exp2(s * sl2 - m * sl2) as bf16 into the swizzled K-major tile; row sum in fp32
      const uint32_t prow = pbuf + row * 128;
#pragma unroll
      for (int c = 0; c < BKEY / 32; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2(fmaf(__uint_as_float(r[c][i]), sl2, moff));
          float p1 = ex2(fmaf(__uint_as_float(r[c][i + 1]), sl2, moff));
          if (valid != BKEY) {
            if (c * 32 + i >= valid) p0 = 0.f;
            if (c * 32 + i + 1 >= valid) p1 = 0.f;
          }
          l += p0 + p1;
          pk[i >> 1] = ptx::pack_bf16x2(p0, p1);
        }
        // columns [32 c, 32 c + 32) = 16-byte chunks 4 (c % 2) .. + 3 of the 64-key tile c / 2
        const uint32_t tile = prow + (c >> 1) * (BQ * 128);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t chunk = static_cast<uint32_t>(((c & 1) * 4 + q4) ^ (row & 7));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + chunk * 16), "r"(pk[4 * q4]),
                       "r"(pk[4 * q4 + 1]), "r"(pk[4 * q4 + 2]), "r"(pk[4 * q4 + 3])
                       : "memory");
        }
      }
      ptx::fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async proxy
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&pfull_bar));
    }
    // epilogue: O / l -> bf16 -> global (rows beyond Tq exist only in the tile)
    ptx::mbar_wait_warp(ptx::smem_u32(&ofull_bar), nit & 1);
    ptx::tc_fence_after();
    const float inv = 1.f / l;
    const int t = q0 + row;
    __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.Tq + t) * p.ldo + v_col0;
#pragma unroll 1
    for (int c = 0; c < DV / 32; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_o + lane_off + c * 32, r);
      ptx::tmem_ld_wait();
      if (t < p.Tq) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 o;
          o.x = ptx::pack_bf16x2(__uint_as_float(r[8 * q4 + 0]) * inv, __uint_as_float(r[8 * q4 + 1]) * inv);
          o.y = ptx::pack_bf16x2(__uint_as_float(r[8 * q4 + 2]) * inv, __uint_as_float(r[8 * q4 + 3]) * inv);
          o.z = ptx::pack_bf16x2(__uint_as_float(r[8 * q4 + 4]) * inv, __uint_as_float(r[8 * q4 + 5]) * inv);
          o.w = ptx::pack_bf16x2(__uint_as_float(r[8 * q4 + 6]) * inv, __uint_as_float(r[8 * q4 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + q4 * 8) = o;
        }
      }
    }
    ptx::tc_fence_before();  // every tcgen05.ld of O has completed: the next item may overwrite it
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&oempty_bar));
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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

struct MapKey {
  const void* ptr;
  uint64_t cols, rows, batch, ld;
  uint32_t box0, box1;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && cols == o.cols && rows == o.rows && batch == o.batch && ld == o.ld && box0 == o.box0 &&
           box1 == o.box1;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    for (uint64_t v : {k.cols, k.rows, k.batch, k.ld, static_cast<uint64_t>(k.box0), static_cast<uint64_t>(k.box1)})
      h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    return static_cast<size_t>(h);
  }
};

// token-major strided view [batch][rows][cols] (row pitch ld elements) as a rank-3 map, box = box0 channels x box1
// rows, swizzle = box row bytes (64 or 128); rows beyond `rows` read as zeros (they never alias the next sample)
CUtensorMap view_map(const void* ptr, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t ld, uint32_t box0,
                     uint32_t box1) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, cols, rows, batch, ld, box0, box1};
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  CUtensorMap tm;
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {ld * 2, rows * ld * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           box0 * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  T2P_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (attention) failed with code " + std::to_string(int(r)));
  if (cache.size() > 65536) cache.clear();
  cache.emplace(key, tm);
  return tm;
}

template <int D, int DV>
void launch(const AttnArgs& a, cudaStream_t st) {
  using C = ACfg<D, DV>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured))
    T2P_CUDA(cudaFuncSetAttribute(attention_tc_kernel<D, DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  AttnTcParams p{};
  const uint64_t width = static_cast<uint64_t>(a.heads) * a.d;
  p.tm_q = view_map(a.q, width, a.Tq, a.B, a.ldq, C::CH, 128);
  p.tm_k = view_map(a.k, width, a.Tk, a.B, a.ldk, C::CH, 128);
  p.tm_v = view_map(a.v, width, a.Tk, a.B, a.ldv, C::CH, 64);
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.ldo = a.ldo;
  p.Tq = a.Tq; p.Tk = a.Tk; p.heads = a.heads; p.d = a.d;
  p.dv_splits = a.d / DV;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.q_tiles = cdiv(a.Tq, BQ);
  p.hy = a.heads * p.dv_splits;
  const long long items = static_cast<long long>(a.B) * p.hy * p.q_tiles;
  T2P_CHECK(items < (1ll << 31), "attention problem too large");
  p.items = static_cast<int>(items);
  const int resident = device_sm_count() * (C::TMEM_COLS == 256 ? 2 : 1);  // CTAs that fit at once
  attention_tc_kernel<D, DV><<<std::min(p.items, resident), THREADS, C::SMEM, st>>>(p);
  T2P_LAUNCH_CHECK();
}

}  // namespace

bool attention_tc_supported(const AttnArgs& a) {
  const bool dims = a.d == 32 || a.d == 64 || a.d == 128 || (a.d % 256 == 0 && a.d <= 4096);
  const bool aligned = (a.ldq % 8 == 0) && (a.ldk % 8 == 0) && (a.ldv % 8 == 0) && (a.ldo % 8 == 0) &&
                       ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) |
                         reinterpret_cast<uintptr_t>(a.v) | reinterpret_cast<uintptr_t>(a.out)) % 16 == 0);
  return dims && aligned && a.Tq > 0 && a.Tk > 0;
}

void attention_tc(const AttnArgs& a, cudaStream_t st) {
  T2P_CHECK(attention_tc_supported(a), "unsupported shape / alignment for the tcgen05 attention kernel");
  if (a.d == 32) launch<32, 32>(a, st);
  else if (a.d == 64) launch<64, 64>(a, st);
  else if (a.d == 128) launch<128, 128>(a, st);
  else launch<256, 256>(a, st);  // d = 256 k: 256-wide value slices, scores over the full d in every slice
}

}  // namespace t2p
