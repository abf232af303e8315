// Stand-alone device self-test for the implicit-GEMM kernels (no torch, no oracle): compares the
// tcgen05 path and the SIMT path against a one-thread-per-output naive kernel on random inputs.
// Build: see __graft_entry__.build();  run on a B200:  build/selftest_gemm
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

#include "kernels.h"

using namespace t2p;

static void fail_if(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    fprintf(stderr, "CUDA error at %s: %s\n", what, cudaGetErrorString(e));
    exit(2);
  }
}

__global__ void naive_conv(const __nv_bfloat16* a0, int c0, const __nv_bfloat16* a1, int c1, int B, int H,
                           int W, int ks, const __nv_bfloat16* w, int N, const float* bias,
                           const float* rowbias, int rps, const __nv_bfloat16* res, int res_up, float alpha,
                           float* out) {
  const long long M = 1ll * B * H * W;
  const long long idx = blockIdx.x * 1ll * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const int n = idx % N;
  const long long m = idx / N;
  const int hw = H * W;
  const int b = m / hw, rem = m % hw, h = rem / W, x = rem % W;
  const int ctot = c0 + c1, pad = ks / 2;
  double acc = 0.0;
  for (int kh = 0; kh < ks; ++kh)
    for (int kw = 0; kw < ks; ++kw) {
      const int ih = h + kh - pad, iw = x + kw - pad;
      if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
      const long long pix = (1ll * b * H + ih) * W + iw;
      const __nv_bfloat16* wr = w + (1ll * n * ks * ks + kh * ks + kw) * ctot;
      for (int c = 0; c < c0; ++c) acc += double(__bfloat162float(a0[pix * c0 + c])) * double(__bfloat162float(wr[c]));
      for (int c = 0; c < c1; ++c)
        acc += double(__bfloat162float(a1[pix * c1 + c])) * double(__bfloat162float(wr[c0 + c]));
    }
  float v = float(acc);
  if (bias) v += bias[n];
  if (rowbias) v += rowbias[(m / rps) * N + n];
  if (res) {
    long long rr = m;
    if (res_up) rr = (1ll * b * (H / 2) + h / 2) * (W / 2) + x / 2;
    v += __bfloat162float(res[rr * N + n]);
  }
  out[idx] = v * alpha;
}

static float frand() { return (rand() / float(RAND_MAX)) * 2.f - 1.f; }

template <typename T>
static T* dev_upload(const std::vector<T>& h) {
  T* d;
  fail_if(cudaMalloc(&d, h.size() * sizeof(T) + 256), "malloc");
  fail_if(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice), "h2d");
  return d;
}

static std::vector<__nv_bfloat16> rand_bf16(size_t n, float scale) {
  std::vector<__nv_bfloat16> v(n);
  for (auto& x : v) x = __float2bfloat16(frand() * scale);
  return v;
}

struct Case {
  const char* name;
  int B, H, W, c0, c1, ks, N;
  bool bias, rowbias, res, res_up, out_fp32, stats;
};

static int run_case(const Case& c) {
  const long long M = 1ll * c.B * c.H * c.W;
  const int ctot = c.c0 + c.c1, K = c.ks * c.ks * ctot;
  auto ha0 = rand_bf16(M * c.c0, 1.f);
  auto ha1 = rand_bf16(c.c1 ? M * c.c1 : 1, 1.f);
  auto hw = rand_bf16(1ll * c.N * K, 1.f / sqrtf(float(K)));
  std::vector<float> hb(c.N), hrb(1ll * c.B * c.N);
  for (auto& x : hb) x = frand();
  for (auto& x : hrb) x = frand();
  const long long res_rows = c.res_up ? M / 4 : M;
  auto hres = rand_bf16(res_rows * c.N, 1.f);
  std::vector<float> hres32(hres.size());
  for (size_t i = 0; i < hres.size(); ++i) hres32[i] = __bfloat162float(hres[i]);

  auto* a0 = dev_upload(ha0);
  auto* a1 = dev_upload(ha1);
  auto* w = dev_upload(hw);
  auto* bias = dev_upload(hb);
  auto* rb = dev_upload(hrb);
  auto* res = dev_upload(hres);
  auto* res32 = dev_upload(hres32);
  float *ref, *out32, *ssum, *ssq;
  __nv_bfloat16* out16;
  fail_if(cudaMalloc(&ref, M * c.N * 4), "malloc");
  fail_if(cudaMalloc(&out32, M * c.N * 4), "malloc");
  fail_if(cudaMalloc(&out16, M * c.N * 2), "malloc");
  const long long tiles = (M + 31) / 32;  // upper bound on the number of statistics tiles (>= 32 pixels each)
  fail_if(cudaMalloc(&ssum, tiles * c.N * 8), "malloc");
  ssq = nullptr;
  cudaMemset(out32, 0xff, M * c.N * 4);
  cudaMemset(out16, 0xff, M * c.N * 2);

  const float alpha = c.res ? 0.70710678f : 1.f;
  const int rps = c.H * c.W;
  naive_conv<<<(unsigned)((M * c.N + 255) / 256), 256>>>(a0, c.c0, c.c1 ? a1 : nullptr, c.c1, c.B, c.H, c.W, c.ks, w,
                                                        c.N, c.bias ? bias : nullptr, c.rowbias ? rb : nullptr, rps,
                                                        c.res ? res : nullptr, c.res_up, alpha, ref);
  fail_if(cudaGetLastError(), "naive launch");

  ConvGemmArgs g;
  g.a0 = a0; g.c0 = c.c0; g.a1 = c.c1 ? a1 : nullptr; g.c1 = c.c1;
  g.B = c.B; g.H = c.H; g.W = c.W; g.ksize = c.ks; g.w = w; g.N = c.N;
  g.bias = c.bias ? bias : nullptr;
  g.rowbias = c.rowbias ? rb : nullptr;
  g.rows_per_sample = rps;
  g.res_up = c.res_up;
  g.alpha = alpha;
  if (c.out_fp32) { g.out = out32; g.out_dtype = kF32; g.residual = c.res ? (const void*)res32 : nullptr; }
  else { g.out = out16; g.out_dtype = kBF16; g.residual = c.res ? (const void*)res : nullptr; }
  const int stile = (c.stats && !c.out_fp32) ? conv_gemm_tc_stat_tile(g) : 0;
  const bool do_stats = stile > 0;
  if (do_stats) g.stat_part = ssum;

  int rc = 0;
  std::vector<float> href(M * c.N), hout(M * c.N);
  for (int pass = 0; pass < 2; ++pass) {  // 0: tcgen05, 1: SIMT
    if (pass == 1) g.stat_part = nullptr;
    try {
      if (pass == 0) conv_gemm_tc(g, 0);
      else conv_gemm_simt(g, kBF16, 0);
    } catch (const std::exception& e) {
      printf("  [%s] %s: EXCEPTION %s\n", pass ? "simt" : "tc", c.name, e.what());
      return 1;
    }
    fail_if(cudaDeviceSynchronize(), pass ? "simt kernel" : "tc kernel");
    fail_if(cudaMemcpy(href.data(), ref, M * c.N * 4, cudaMemcpyDeviceToHost), "d2h");
    if (c.out_fp32) fail_if(cudaMemcpy(hout.data(), out32, M * c.N * 4, cudaMemcpyDeviceToHost), "d2h");
    else {
      std::vector<__nv_bfloat16> t(M * c.N);
      fail_if(cudaMemcpy(t.data(), out16, M * c.N * 2, cudaMemcpyDeviceToHost), "d2h");
      for (long long i = 0; i < M * c.N; ++i) hout[i] = __bfloat162float(t[i]);
    }
    double maxerr = 0, maxref = 0;
    for (long long i = 0; i < M * c.N; ++i) {
      double e = fabs(double(hout[i]) - double(href[i]));
      if (!(e == e)) e = 1e30;
      if (e > maxerr) maxerr = e;
      if (fabs(href[i]) > maxref) maxref = fabs(href[i]);
    }
    const double tol = (c.out_fp32 ? 2e-4 : 1.2e-2) * (maxref > 1 ? maxref : 1);
    const bool ok = maxerr <= tol;
    printf("  [%s] %-28s M=%lld N=%d K=%d  maxerr=%.3e (max|ref|=%.2f) %s\n", pass ? "simt" : "tc  ", c.name, M, c.N,
           K, maxerr, maxref, ok ? "PASS" : "FAIL");
    if (!ok) rc = 1;
    if (pass == 0 && do_stats) {
      const long long stiles = M / stile;
      std::vector<float> hp(stiles * c.N * 2), hs(c.B * c.N, 0.f), hq(c.B * c.N, 0.f);
      cudaMemcpy(hp.data(), ssum, stiles * c.N * 8, cudaMemcpyDeviceToHost);
      for (long long t = 0; t < stiles; ++t)
        for (int n = 0; n < c.N; ++n) {
          const int b = int(t * stile / rps);
          hs[b * c.N + n] += hp[(t * c.N + n) * 2];
          hq[b * c.N + n] += hp[(t * c.N + n) * 2 + 1];
        }
      double es = 0, eq = 0;
      for (int b = 0; b < c.B; ++b)
        for (int n = 0; n < c.N; ++n) {
          double s = 0, q = 0;
          for (int r = 0; r < rps; ++r) {
            const double v = hout[(1ll * b * rps + r) * c.N + n];
            s += v; q += v * v;
          }
          es = fmax(es, fabs(s - hs[b * c.N + n]) / (1 + fabs(s)));
          eq = fmax(eq, fabs(q - hq[b * c.N + n]) / (1 + fabs(q)));
        }
      const bool sok = es < 1e-3 && eq < 1e-3;
      printf("  [tc  ] %-28s stats rel err sum=%.2e sq=%.2e %s\n", c.name, es, eq, sok ? "PASS" : "FAIL");
      if (!sok) rc = 1;
    }
    cudaMemset(out32, 0xff, M * c.N * 4);
    cudaMemset(out16, 0xff, M * c.N * 2);
  }
  cudaFree(a0); cudaFree(a1); cudaFree(w); cudaFree(bias); cudaFree(rb); cudaFree(res); cudaFree(res32);
  cudaFree(ref); cudaFree(out32); cudaFree(out16); cudaFree(ssum); cudaFree(ssq);
  return rc;
}

__global__ void fill_bf16(__nv_bfloat16* p, long long n, float scale, unsigned seed) {
  const long long i = blockIdx.x * 1ll * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned h = static_cast<unsigned>(i) * 2654435761u + seed;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  p[i] = __float2bfloat16(((h & 0xffff) / 32768.f - 1.f) * scale);
}

// random (not zero) operands so that the power draw, and hence the clock, is representative;
// epi = 0: plain store, 1: + bias + per-sample bias + GroupNorm statistics, 2: + residual
static void bench_case(int B, int H, int W, int cin, int N, int ks, int epi, int iters = 10) {
  const long long M = 1ll * B * H * W;
  const int K = ks * ks * cin;
  __nv_bfloat16 *a, *w, *o, *res;
  float *bias, *rb, *st;
  fail_if(cudaMalloc(&a, M * cin * 2), "malloc");
  fail_if(cudaMalloc(&w, 1ll * N * K * 2), "malloc");
  fail_if(cudaMalloc(&o, M * N * 2), "malloc");
  fail_if(cudaMalloc(&res, M * N * 2), "malloc");
  fail_if(cudaMalloc(&bias, N * 4), "malloc");
  fail_if(cudaMalloc(&rb, 1ll * B * N * 4), "malloc");
  fail_if(cudaMalloc(&st, (M / 32 + 1) * N * 8), "malloc");  // statistics tiles are >= 32 pixels
  fill_bf16<<<(unsigned)((M * cin + 255) / 256), 256>>>(a, M * cin, 1.f, 1u);
  fill_bf16<<<(unsigned)((1ll * N * K + 255) / 256), 256>>>(w, 1ll * N * K, 0.05f, 2u);
  fill_bf16<<<(unsigned)((M * N + 255) / 256), 256>>>(res, M * N, 1.f, 3u);
  cudaMemset(bias, 0, N * 4);
  cudaMemset(rb, 0, 1ll * B * N * 4);
  ConvGemmArgs g;
  g.a0 = a; g.c0 = cin; g.B = B; g.H = H; g.W = W; g.ksize = ks; g.w = w; g.N = N; g.out = o;
  g.rows_per_sample = H * W;
  if (epi >= 1) {
    g.bias = bias; g.rowbias = rb;
    if (conv_gemm_tc_stat_tile(g) > 0) g.stat_part = st;
  }
  if (epi >= 2) { g.residual = res; g.alpha = 0.70710678f; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) conv_gemm_tc(g, 0);
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) conv_gemm_tc(g, 0);
  cudaEventRecord(e1);
  fail_if(cudaDeviceSynchronize(), "bench");
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= iters;
  printf("  bench B=%d %dx%d cin=%d N=%d k=%d epi=%d: %.3f ms  %.1f TFLOP/s\n", B, H, W, cin, N, ks, epi, ms,
         2.0 * M * N * K / (ms * 1e-3) / 1e12);
  cudaFree(a); cudaFree(w); cudaFree(o); cudaFree(res); cudaFree(bias); cudaFree(rb); cudaFree(st);
}

int main(int argc, char** argv) {
  srand(1234);
  const Case cases[] = {
      {"gemm 256x128x128", 1, 1, 256, 128, 0, 1, 128, false, false, false, false, false, false},
      {"gemm M=200 ragged", 1, 1, 200, 128, 0, 1, 128, true, false, false, false, false, false},
      {"gemm 1024x2048x256 fp32", 1, 1, 1024, 256, 0, 1, 2048, true, false, false, false, true, false},
      {"conv3 128x128 c128->128", 2, 128, 128, 128, 0, 3, 128, true, true, false, false, false, true},
      {"conv3 64x64 c128+128->128", 2, 64, 64, 128, 128, 3, 128, true, true, true, false, false, true},
      {"conv3 32x32 c256->256", 3, 32, 32, 256, 0, 3, 256, true, false, true, false, false, true},
      {"conv3 16x16 c256+256->256", 3, 16, 16, 256, 256, 3, 256, true, true, true, false, false, true},
      {"conv3 8x8 c256->256", 3, 8, 8, 256, 0, 3, 256, true, true, false, false, false, true},
      {"conv3 4x4 c256->256", 5, 4, 4, 256, 0, 3, 256, true, true, true, false, false, true},
      {"conv3 2x2 c512->512", 5, 2, 2, 512, 0, 3, 512, true, true, true, false, false, false},
      {"conv3 128x128 c128->5 fp32", 2, 128, 128, 128, 0, 3, 5, true, false, false, false, true, false},
      {"conv3 32x32 res_up", 2, 32, 32, 256, 0, 3, 256, true, false, true, true, false, false},
      {"conv1 64x64 c384->128", 2, 64, 64, 256, 128, 1, 128, true, false, false, false, false, false},
      {"conv3 256x256 c256->256", 1, 256, 256, 256, 0, 3, 256, true, true, true, false, false, true},
      {"conv3 128x128 c64+64->128 res", 3, 128, 128, 64, 64, 3, 128, true, true, true, false, false, true},
      {"conv3 4x4 c128->128 B=5 ragged", 5, 4, 4, 128, 0, 3, 128, true, true, true, false, false, false},
      {"conv3 64x64 c64->192", 2, 64, 64, 64, 0, 3, 192, true, true, true, false, false, true},
      {"conv3 16x16 res_up c128->128", 3, 16, 16, 128, 0, 3, 128, true, true, true, true, false, true},
      {"gemm M=300 K=4096 N=512", 1, 1, 300, 4096, 0, 1, 512, false, false, false, false, false, false},
      {"conv1 8x8 c256->256 B=70", 70, 8, 8, 256, 0, 1, 256, true, true, true, false, true, false},
  };
  if (argc > 1 && std::string(argv[1]) == "one") {
    // single shape for ncu captures: one <B> <H> <W> <cin> <N> <ks> <epi>
    if (argc < 9) { fprintf(stderr, "usage: selftest_gemm one B H W cin N ks epi [iters]\n"); return 2; }
    bench_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]),
               argc > 9 ? atoi(argv[9]) : 10);
    return 0;
  }
  int rc = 0;
  const char* only = getenv("T2P_SELFTEST_CASE");
  int ci = 0;
  for (const auto& c : cases) {
    if (!only || atoi(only) == ci) rc |= run_case(c);
    ++ci;
  }
  if (argc > 1) {
    for (int epi = 0; epi < 3; ++epi) {
      bench_case(64, 128, 128, 128, 128, 3, epi);
      bench_case(64, 128, 128, 256, 128, 3, epi);
      bench_case(64, 64, 64, 128, 128, 3, epi);
      bench_case(64, 32, 32, 256, 256, 3, epi);
      bench_case(64, 16, 16, 256, 256, 3, epi);
      bench_case(64, 8, 8, 256, 256, 3, epi);
      bench_case(64, 4, 4, 256, 256, 3, epi);
      bench_case(64, 128, 128, 256, 128, 1, epi);
    }
  }
  printf(rc ? "SELFTEST FAILED\n" : "SELFTEST OK\n");
  return rc;
}
