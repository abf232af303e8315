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

// -2 ln u for a philox_uniform() value, u in [2^-33, 1]: the argument is never zero, denormal, infinite or NaN,
// so this is the classic exponent split (mantissa in [2/3, 4/3)) + ln(1 + f) = f - f^2/2 + f^3 q(f) without the
// special-case handling of logf (22 -> 15 instructions).  q: degree-6 least-squares fit on [-1/3, 1/3] weighted
// for the relative error of ln(1 + f), 4.3e-8 in exact arithmetic.  The result is never negative.
__device__ __forceinline__ float neg2_log_uniform(float u) {
  const int ix = __float_as_int(u);
  const int e = (ix - 0x3f2aaaab) & 0xff800000;
  const float f = __fsub_rn(__int_as_float(ix - e), 1.0f);
  const float k = __fmul_rn(static_cast<float>(e), 1.1920928955078125e-07f);  // exponent, e / 2^23
  float q = 0.1401381939649582f;
  q = __fmaf_rn(q, f, -0.15139269828796387f);
  q = __fmaf_rn(q, f, 0.13999086618423462f);
  q = __fmaf_rn(q, f, -0.1646934598684311f);
  q = __fmaf_rn(q, f, 0.20010896027088165f);
  q = __fmaf_rn(q, f, -0.2500436305999756f);
  q = __fmaf_rn(q, f, 0.3333320617675781f);
  const float t = __fmul_rn(f, f);
  const float p = __fmaf_rn(__fmaf_rn(q, f, -0.5f), t, f);
  return __fmul_rn(__fmaf_rn(k, 0.6931471805599453f, p), -2.0f);
}

__device__ __forceinline__ float sqrt_approx(float v) {  // MUFU.SQRT, relative error 2^-23; sqrt(+-0) = +-0
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// Every product below is an explicit intrinsic: the normals must not depend on how the compiler contracts
// multiply-adds in the kernel this is inlined into (the step kernels and t2p_philox_normal must agree bit for bit).
__device__ __forceinline__ void philox_normal4(unsigned long long seed, unsigned long long stream,
                                               unsigned long long quad, float (&z)[4]) {
  const uint4 r = philox4x32_10(
      make_uint4(static_cast<unsigned int>(quad), static_cast<unsigned int>(quad >> 32),
                 static_cast<unsigned int>(stream), static_cast<unsigned int>(stream >> 32)),
      make_uint2(static_cast<unsigned int>(seed), static_cast<unsigned int>(seed >> 32)));
  const float u0 = philox_uniform(r.x), u1 = philox_uniform(r.y);
  const float u2 = philox_uniform(r.z), u3 = philox_uniform(r.w);
  const float r0 = sqrt_approx(neg2_log_uniform(u0)), r1 = sqrt_approx(neg2_log_uniform(u2));
  // angle 2 pi u in (0, 2 pi]: evaluated as -(cos, sin)(2 pi u - pi) so that the MUFU sine / cosine run inside
  // [-pi, pi], where their absolute error is 2^-21.4 (the normals stay within 3e-6 of the numpy restatement)
  float s0, c0, s1, c1;
  __sincosf(__fmaf_rn(6.283185307179586f, u1, -3.14159265358979f), &s0, &c0);
  __sincosf(__fmaf_rn(6.283185307179586f, u3, -3.14159265358979f), &s1, &c1);
  z[0] = __fmul_rn(-r0, c0);
  z[1] = __fmul_rn(-r0, s0);
  z[2] = __fmul_rn(-r1, c1);
  z[3] = __fmul_rn(-r1, s1);
}

}  // namespace t2p
