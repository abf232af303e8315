// Hardware probe (not part of libt2p.so): can tcgen05.mma read a 128-byte-swizzled K-major operand whose start
// address is shifted by s rows (s * 128 bytes, i.e. not aligned to the 1024-byte swizzle atom)?  The conv kernel
// needs this to reuse ONE halo tile in shared memory for the three horizontal filter taps.
//   D[128 x 128] = A[128 x 64] * B[rows s .. s+127 of a 136 x 64 tile]^T
// Variants: base_offset field = s (PTX: (addr >> 7) & 7) or 0.  Build/run: see tools/run_probe.sh
#include <cuda.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.h"
#include "ptx.cuh"

using namespace t2p;

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                    const __grid_constant__ CUtensorMap tm_b, int shift, int use_base,
                                                    float* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + 128 * 128;  // A: 16 KB, B: 136 rows * 128 B = 17 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&full_bar), 1);
    ptx::mbar_init(ptx::smem_u32(&done_bar), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t fb = ptx::smem_u32(&full_bar);
    ptx::mbar_arrive_expect_tx(fb, 128 * 128 + 136 * 128);
    ptx::tma_load_4d(sa, &tm_a, fb, 0, 0, 0, 0);
    ptx::tma_load_4d(sb, &tm_b, fb, 0, 0, 0, 0);
    ptx::mbar_wait(fb, 0);
    ptx::tc_fence_after();
    const uint64_t da = ptx::umma_desc_k_sw128(sa);
    uint64_t db = ptx::umma_desc_k_sw128(sb + shift * 128);
    if (use_base) db |= static_cast<uint64_t>(((sb + shift * 128) >> 7) & 7) << 49;
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128);
    for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
    ptx::umma_commit(ptx::smem_u32(&done_bar));
  }
  ptx::mbar_wait(ptx::smem_u32(&done_bar), 0);
  ptx::tc_fence_after();
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    ptx::tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 128 + c * 32 + i] = __uint_as_float(r[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 128);
}

int main() {
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<EncodeFn>(f);
  const int RB = 136;
  std::vector<__nv_bfloat16> ha(128 * 64), hb(RB * 64);
  srand(7);
  for (auto& v : ha) v = __float2bfloat16((rand() % 17 - 8) / 8.f);
  for (auto& v : hb) v = __float2bfloat16((rand() % 17 - 8) / 8.f);
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dout, 128 * 128 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  auto mk = [&](void* p, int rows) {
    CUtensorMap tm;
    cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(rows), 1, 1};
    cuuint64_t strides[3] = {128, static_cast<cuuint64_t>(128) * rows, static_cast<cuuint64_t>(128) * rows};
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(rows), 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", int(r)); exit(2); }
    return tm;
  };
  const CUtensorMap tma = mk(da, 128), tmb = mk(db, RB);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  std::vector<float> hout(128 * 128);
  for (int use_base = 0; use_base < 2; ++use_base)
    for (int s = 0; s < 8; ++s) {
      cudaMemset(dout, 0, 128 * 128 * 4);
      probe_kernel<<<1, 128, 40 * 1024>>>(tma, tmb, s, use_base, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d base %d: CUDA error %s\n", s, use_base, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hout.data(), dout, 128 * 128 * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += double(__bfloat162float(ha[m * 64 + k])) * double(__bfloat162float(hb[(n + s) * 64 + k]));
          maxerr = fmax(maxerr, fabs(ref - hout[m * 128 + n]));
        }
      printf("shift %d rows, base_offset field %s: max err %.3e %s\n", s, use_base ? "(addr>>7)&7" : "0", maxerr,
             maxerr < 1e-3 ? "MATCH" : "mismatch");
    }
  return 0;
}
