// Throughput of the special-function ops a fused SiLU could use (lanes per clock per SM), sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
  uint32_t ua = __float_as_uint(a), ub = ua + 1, uc = ua + 2, ud = ua + 3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(b)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(c)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(d)); }
    if (OP == 1) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d)); }
    if (OP == 2) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(b)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(d)); }
    if (OP == 3) { asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ua)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ub)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(uc)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ud)); }
    if (OP == 4) { asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ua)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ub)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(uc)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ud)); }
    if (OP == 5) { asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ua)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ub)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(uc)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ud)); }
    if (OP == 6) { asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(ua)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(ub)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(uc)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(ud)); }
    if (OP == 7) { asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(ua)); asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(ub)); asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(uc)); asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(ud)); }
    if (OP == 8) { asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(ua)); asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(ub)); asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(uc)); asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(ud)); }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __uint_as_float(ua ^ ub ^ uc ^ ud);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <int OP>
void run(const char* name, int elems_per_instr) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  const int iters = 4096;
  k<OP><<<148, 1024>>>(d, iters);  // warm
  k<OP><<<148, 1024>>>(d, iters);
  cudaDeviceSynchronize();
  float clk; cudaMemcpy(&clk, d, 4, cudaMemcpyDeviceToHost);
  double lanes = 1024.0 * 4 * iters;  // thread-instructions per SM
  printf("%-24s %7.2f thread-instr/clk/SM  = %7.2f elements/clk/SM   (%s)\n", name, lanes / clk, lanes * elems_per_instr / clk, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  run<0>("tanh.approx.f32", 1); run<1>("ex2.approx.f32", 1); run<2>("rcp.approx.f32", 1);
  run<3>("tanh.approx.bf16x2", 2); run<4>("tanh.approx.f16x2", 2); run<5>("ex2.approx.bf16x2", 2); run<6>("ex2.approx.f16x2", 2);
  run<7>("fma.rn.bf16x2", 2); run<8>("fma.rn.f16x2", 2);
  return 0;
}
