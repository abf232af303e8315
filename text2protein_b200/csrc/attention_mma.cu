// Flash-style attention on the bf16 tensor cores (mma.sync m16n8k16, fp32 softmax / accumulators).
// Same contract as attention_simt (attention.cu).  One CTA = 64 query rows x one head x one DV-wide
// slice of the value dimension; K/V are streamed in 64-key tiles, the [64 x 64] score tile lives in
// registers, P is re-used as the A operand of P.V without leaving the register file.
//   * head dim d (QK) is consumed in DK-wide chunks so that the single-head AttnBlockpp (d = C up to
//     1024) fits in shared memory; when d > DV the value dimension is split over blockIdx.y and the
//     score tile is recomputed per slice (attention is < 1.5 % of the network FLOPs).
// TODO(round 2): tcgen05 / TMEM version (S and O accumulators in TMEM, K/V by TMA).
#include "kernels.h"

namespace t2p {
namespace {

constexpr int BM = 64;   // query rows per CTA (16 per warp)
constexpr int BN = 64;   // keys per tile
constexpr int NT = 128;  // threads

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

struct Params {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* out;
  int heads, Tq, Tk, d;
  long long ldq, ldk, ldv, ldo;
  float scale_log2;  // scale * log2(e)
};

// copies a [64 x W] bf16 tile (rows r0.., row pitch ld, column offset c0) into smem with pitch W + 8;
// rows >= rmax are zero-filled
template <int W>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld, int r0, int rmax,
                                          int c0) {
  constexpr int VPR = W / 8;  // 16-byte vectors per row
  for (int i = threadIdx.x; i < 64 * VPR; i += NT) {
    const int r = i / VPR, cv = i - r * VPR;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r0 + r < rmax) val = *reinterpret_cast<const uint4*>(src + static_cast<long long>(r0 + r) * ld + c0 + cv * 8);
    *reinterpret_cast<uint4*>(dst + r * (W + 8) + cv * 8) = val;
  }
}

template <int DK, int DV>
__global__ void __launch_bounds__(NT) attention_mma_kernel(const Params p) {
  __shared__ __align__(16) __nv_bfloat16 Qs[BM * (DK + 8)];
  __shared__ __align__(16) __nv_bfloat16 Ks[BN * (DK + 8)];
  __shared__ __align__(16) __nv_bfloat16 Vs[BN * (DV + 8)];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int slices = p.d / DV;
  const int h = blockIdx.y / slices, sl = blockIdx.y - h * slices;
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * BM;
  const __nv_bfloat16* qb = p.q + static_cast<long long>(b) * p.Tq * p.ldq + h * p.d;
  const __nv_bfloat16* kb = p.k + static_cast<long long>(b) * p.Tk * p.ldk + h * p.d;
  const __nv_bfloat16* vb = p.v + static_cast<long long>(b) * p.Tk * p.ldv + h * p.d + sl * DV;

  float o[DV / 8][4];
#pragma unroll
  for (int i = 0; i < DV / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};

  for (int n0 = 0; n0 < p.Tk; n0 += BN) {
    float s[BN / 8][4];
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
    // ---- S = Q K^T over the head dim in DK-wide chunks
    for (int dc = 0; dc < p.d; dc += DK) {
      __syncthreads();  // previous users of Qs / Ks / Vs are done
      load_tile<DK>(Qs, qb, p.ldq, m0, p.Tq, dc);
      load_tile<DK>(Ks, kb, p.ldk, n0, p.Tk, dc);
      if (dc == 0) load_tile<DV>(Vs, vb, p.ldv, n0, p.Tk, 0);
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < DK / 16; ++kk) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4(smem_u32(Qs + (warp * 16 + (lane & 15)) * (DK + 8) + kk * 16 + (lane >> 4) * 8), a0, a1, a2, a3);
#pragma unroll
        for (int np = 0; np < BN / 16; ++np) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(smem_u32(Ks + (np * 16 + (lane & 7) + ((lane >> 4) << 3)) * (DK + 8) + kk * 16 + ((lane >> 3) & 1) * 8),
                  b0, b1, b2, b3);
          mma16816(s[2 * np], a0, a1, a2, a3, b0, b1);
          mma16816(s[2 * np + 1], a0, a1, a2, a3, b2, b3);
        }
      }
    }
    // ---- online softmax (rows g and g + 8 of this warp's 16)
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) {
      const int key = n0 + i * 8 + tq * 2;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool valid = (key + (e & 1)) < p.Tk;
        s[i][e] = valid ? s[i][e] * p.scale_log2 : -INFINITY;
        tmax[e >> 1] = fmaxf(tmax[e >> 1], s[i][e]);
      }
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
      const float mn = fmaxf(mrow[r], tmax[r]);
      corr[r] = exp2f(mrow[r] - mn);
      mrow[r] = mn;
    }
    float psum[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < BN / 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[i][e] = exp2f(s[i][e] - mrow[e >> 1]);
        psum[e >> 1] += s[i][e];
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) lrow[r] = lrow[r] * corr[r] + psum[r];
#pragma unroll
    for (int i = 0; i < DV / 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // ---- O += P V  (P from registers as the A operand; V^T fragments via ldmatrix.trans)
#pragma unroll
    for (int j = 0; j < BN / 16; ++j) {
      const uint32_t a0 = pack2(s[2 * j][0], s[2 * j][1]);
      const uint32_t a1 = pack2(s[2 * j][2], s[2 * j][3]);
      const uint32_t a2 = pack2(s[2 * j + 1][0], s[2 * j + 1][1]);
      const uint32_t a3 = pack2(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
      for (int np = 0; np < DV / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(Vs + (j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * (DV + 8) + np * 16 + (lane >> 4) * 8),
                  b0, b1, b2, b3);
        mma16816(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma16816(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
  }
  // ---- normalise and store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
  const float inv[2] = {1.f / lrow[0], 1.f / lrow[1]};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = m0 + warp * 16 + g + r * 8;
    if (row >= p.Tq) continue;
    __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.Tq + row) * p.ldo + h * p.d + sl * DV;
#pragma unroll
    for (int i = 0; i < DV / 8; ++i)
      *reinterpret_cast<uint32_t*>(orow + i * 8 + tq * 2) = pack2(o[i][2 * r] * inv[r], o[i][2 * r + 1] * inv[r]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Pipelined variant for head dims that fit shared memory whole (D <= 256): Q is staged once, K / V tiles are
// double-buffered with cp.async so the loads of key tile i + 1 overlap the math of tile i.  Same math and
// fragment layout as the kernel above.
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int W, int ROWS, int NTHREADS>
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld, int r0,
                                                int rmax, int c0) {
  constexpr int VPR = W / 8;
  for (int i = threadIdx.x; i < ROWS * VPR; i += NTHREADS) {
    const int r = i / VPR, cv = i - r * VPR;
    const bool ok = r0 + r < rmax;
    const __nv_bfloat16* g = src + static_cast<long long>(ok ? r0 + r : 0) * ld + c0 + cv * 8;
    cp_async16(smem_u32(dst + r * (W + 8) + cv * 8), g, ok);
  }
}

// WARPS x 16 query rows per CTA (each warp owns 16 rows; K / V tiles are shared by all warps)
template <int D, int DV, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) attention_pipe_kernel(const Params p) {
  constexpr int QROWS = WARPS * 16;
  constexpr int NTH = WARPS * 32;
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_dyn);
  __nv_bfloat16* Ks = Qs + QROWS * (D + 8);            // [2][BN][D + 8]
  __nv_bfloat16* Vs = Ks + 2 * BN * (D + 8);        // [2][BN][DV + 8]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  constexpr int slices = D / DV;
  const int h = blockIdx.y / slices, sl = blockIdx.y - h * slices;
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * QROWS;
  const __nv_bfloat16* qb = p.q + static_cast<long long>(b) * p.Tq * p.ldq + h * D;
  const __nv_bfloat16* kb = p.k + static_cast<long long>(b) * p.Tk * p.ldk + h * D;
  const __nv_bfloat16* vb = p.v + static_cast<long long>(b) * p.Tk * p.ldv + h * D + sl * DV;

  pdl_trigger();
  pdl_wait();
  load_tile_async<D, QROWS, NTH>(Qs, qb, p.ldq, m0, p.Tq, 0);
  load_tile_async<D, BN, NTH>(Ks, kb, p.ldk, 0, p.Tk, 0);
  load_tile_async<DV, BN, NTH>(Vs, vb, p.ldv, 0, p.Tk, 0);
  cp_async_commit();

  float o[DV / 8][4];
#pragma unroll
  for (int i = 0; i < DV / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};

  const int ntiles = (p.Tk + BN - 1) / BN;
  for (int it = 0; it < ntiles; ++it) {
    const int n0 = it * BN;
    const int cur = it & 1;
    if (it + 1 < ntiles) {  // prefetch the next key tile into the other buffer (its readers finished last iteration)
      load_tile_async<D, BN, NTH>(Ks + (cur ^ 1) * BN * (D + 8), kb, p.ldk, n0 + BN, p.Tk, 0);
      load_tile_async<DV, BN, NTH>(Vs + (cur ^ 1) * BN * (DV + 8), vb, p.ldv, n0 + BN, p.Tk, 0);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* Kc = Ks + cur * BN * (D + 8);
    const __nv_bfloat16* Vc = Vs + cur * BN * (DV + 8);
    float s[BN / 8][4];
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4(smem_u32(Qs + (warp * 16 + (lane & 15)) * (D + 8) + kk * 16 + (lane >> 4) * 8), a0, a1, a2, a3);
#pragma unroll
      for (int np = 0; np < BN / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(Kc + (np * 16 + (lane & 7) + ((lane >> 4) << 3)) * (D + 8) + kk * 16 + ((lane >> 3) & 1) * 8),
                b0, b1, b2, b3);
        mma16816(s[2 * np], a0, a1, a2, a3, b0, b1);
        mma16816(s[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    // (this kernel is issue-bound on the softmax arithmetic: the key-range mask is applied on the ragged last
    // tile only, the scaling is folded into the exponent's FMA and exp2 is the single-instruction MUFU form)
    float tmax[2] = {-INFINITY, -INFINITY};
    if (n0 + BN > p.Tk) {
#pragma unroll
      for (int i = 0; i < BN / 8; ++i) {
        const int key = n0 + i * 8 + tq * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (key + (e & 1) >= p.Tk) s[i][e] = -INFINITY;
      }
    }
#pragma unroll
    for (int i = 0; i < BN / 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) tmax[e >> 1] = fmaxf(tmax[e >> 1], s[i][e]);
    float corr[2], moff[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
      const float mn = fmaxf(mrow[r], tmax[r] * p.scale_log2);  // running max of the SCALED scores (scale > 0)
      corr[r] = fast_exp2(mrow[r] - mn);
      mrow[r] = mn;
      moff[r] = -mn;
    }
    float psum[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < BN / 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[i][e] = fast_exp2(fmaf(s[i][e], p.scale_log2, moff[e >> 1]));
        psum[e >> 1] += s[i][e];
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) lrow[r] = lrow[r] * corr[r] + psum[r];
#pragma unroll
    for (int i = 0; i < DV / 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
#pragma unroll
    for (int j = 0; j < BN / 16; ++j) {
      const uint32_t a0 = pack2(s[2 * j][0], s[2 * j][1]);
      const uint32_t a1 = pack2(s[2 * j][2], s[2 * j][3]);
      const uint32_t a2 = pack2(s[2 * j + 1][0], s[2 * j + 1][1]);
      const uint32_t a3 = pack2(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
      for (int np = 0; np < DV / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(Vc + (j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * (DV + 8) + np * 16 + (lane >> 4) * 8),
                  b0, b1, b2, b3);
        mma16816(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma16816(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    __syncthreads();  // everyone is done with buffer `cur` before the next iteration's prefetch overwrites it
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
  const float inv[2] = {1.f / lrow[0], 1.f / lrow[1]};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = m0 + warp * 16 + g + r * 8;
    if (row >= p.Tq) continue;
    __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.Tq + row) * p.ldo + h * D + sl * DV;
#pragma unroll
    for (int i = 0; i < DV / 8; ++i)
      *reinterpret_cast<uint32_t*>(orow + i * 8 + tq * 2) = pack2(o[i][2 * r] * inv[r], o[i][2 * r + 1] * inv[r]);
  }
}

Params make_params(const AttnArgs& a) {
  Params p;
  p.q = static_cast<const __nv_bfloat16*>(a.q);
  p.k = static_cast<const __nv_bfloat16*>(a.k);
  p.v = static_cast<const __nv_bfloat16*>(a.v);
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.heads = a.heads; p.Tq = a.Tq; p.Tk = a.Tk; p.d = a.d;
  p.ldq = a.ldq; p.ldk = a.ldk; p.ldv = a.ldv; p.ldo = a.ldo;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  return p;
}

template <int D, int DV, int WARPS>
void launch_pipe(const AttnArgs& a, cudaStream_t st) {
  constexpr int QROWS = WARPS * 16;
  constexpr size_t smem = sizeof(__nv_bfloat16) * (QROWS * (D + 8) + 2 * BN * (D + 8) + 2 * BN * (DV + 8));
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    T2P_CUDA(cudaFuncSetAttribute(attention_pipe_kernel<D, DV, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  }
  const Params p = make_params(a);
  dim3 grid(cdiv(a.Tq, QROWS), a.heads * (D / DV), a.B);
  launch_pdl(attention_pipe_kernel<D, DV, WARPS>, grid, dim3(WARPS * 32), smem, st, p);
}

template <int DK, int DV>
void launch(const AttnArgs& a, cudaStream_t st) {
  Params p;
  p.q = static_cast<const __nv_bfloat16*>(a.q);
  p.k = static_cast<const __nv_bfloat16*>(a.k);
  p.v = static_cast<const __nv_bfloat16*>(a.v);
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.heads = a.heads; p.Tq = a.Tq; p.Tk = a.Tk; p.d = a.d;
  p.ldq = a.ldq; p.ldk = a.ldk; p.ldv = a.ldv; p.ldo = a.ldo;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  dim3 grid(cdiv(a.Tq, BM), a.heads * (a.d / DV), a.B);
  attention_mma_kernel<DK, DV><<<grid, NT, 0, st>>>(p);
  T2P_LAUNCH_CHECK();
}

}  // namespace

bool attention_mma_supported(const AttnArgs& a) {
  const bool aligned = (a.ldq % 8 == 0) && (a.ldk % 8 == 0) && (a.ldv % 8 == 0) && (a.ldo % 2 == 0) &&
                       ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) |
                         reinterpret_cast<uintptr_t>(a.v)) % 16 == 0);
  return aligned && (a.d == 16 || a.d == 32 || a.d % 64 == 0) && a.d <= 4096;
}

void attention_mma(const AttnArgs& a, cudaStream_t st) {
  T2P_CHECK(attention_mma_supported(a), "unsupported shape / alignment for the tensor-core attention kernel");
  // 128 query rows per CTA when the sequence has them (K / V tiles are then read half as often), else 64; the
  // 256-wide single head of AttnBlockpp keeps its whole value dimension in one CTA (scores computed once)
  const bool wide = a.Tq >= 128;
  if (a.d == 16) { if (wide) launch_pipe<16, 16, 8>(a, st); else launch_pipe<16, 16, 4>(a, st); }
  else if (a.d == 32) { if (wide) launch_pipe<32, 32, 8>(a, st); else launch_pipe<32, 32, 4>(a, st); }
  else if (a.d == 64) { if (wide) launch_pipe<64, 64, 8>(a, st); else launch_pipe<64, 64, 4>(a, st); }
  else if (a.d == 128) { if (wide) launch_pipe<128, 128, 8>(a, st); else launch_pipe<128, 128, 4>(a, st); }
  else if (a.d == 256) { if (wide) launch_pipe<256, 256, 8>(a, st); else launch_pipe<256, 128, 4>(a, st); }
  else if (a.d == 16) launch<16, 16>(a, st);
  else if (a.d == 32) launch<32, 32>(a, st);
  else if (a.d == 64) launch<64, 64>(a, st);
  else if (a.d % 128 == 0) launch<64, 128>(a, st);
  else launch<64, 64>(a, st);
}

}  // namespace t2p
