// Philox4x32-10 counter-based generator + Box-Muller, device side.  Stream layout is documented in
// oracle/philox_ref.py (the CPU restatement the parity tests compare against): counter =
// (quad_lo, quad_hi, stream_lo, stream_hi), key = (seed_lo, seed_hi), one call -> 4 normals.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace t2p {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ float philox_uniform(unsigned int x) {
  // float32(x) * 2^-32 + 2^-33, in (0, 1]; identical to the numpy restatement
  return __fmaf_rn(static_cast<float>(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

__device__ __forceinline__ void philox_normal4(unsigned long long seed, unsigned long long stream,
                                               unsigned long long quad, float (&z)[4]) {
  const uint4 r = philox4x32_10(
      make_uint4(static_cast<unsigned int>(quad), static_cast<unsigned int>(quad >> 32),
                 static_cast<unsigned int>(stream), static_cast<unsigned int>(stream >> 32)),
      make_uint2(static_cast<unsigned int>(seed), static_cast<unsigned int>(seed >> 32)));
  const float u0 = philox_uniform(r.x), u1 = philox_uniform(r.y);
  const float u2 = philox_uniform(r.z), u3 = philox_uniform(r.w);
  const float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
  // angle 2 pi u in (0, 2 pi]: evaluated as -(cos, sin)(2 pi u - pi) so that the MUFU sine / cosine run inside
  // [-pi, pi], where their absolute error is 2^-21.4 (the normals stay within 3e-6 of the numpy restatement)
  float s0, c0, s1, c1;
  __sincosf(fmaf(6.283185307179586f, u1, -3.14159265358979f), &s0, &c0);
  __sincosf(fmaf(6.283185307179586f, u3, -3.14159265358979f), &s1, &c1);
  z[0] = -r0 * c0;
  z[1] = -r0 * s0;
  z[2] = -r1 * c1;
  z[3] = -r1 * s1;
}

}  // namespace t2p
