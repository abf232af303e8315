// Host-side common definitions for the t2p CUDA library: error plumbing and tensor views.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace t2p {

enum DType : int { kF32 = 0, kBF16 = 1, kF64 = 2, kI64 = 3, kU8 = 4 };

inline size_t dtype_size(int dt) {
  switch (dt) {
    case kF32: return 4;
    case kBF16: return 2;
    case kF64: return 8;
    case kI64: return 8;
    case kU8: return 1;
  }
  return 0;
}

struct Error : std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

void set_last_error(const std::string& m);

#define T2P_CHECK(cond, msg)                                                                   \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      throw ::t2p::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
    }                                                                                          \
  } while (0)

#define T2P_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      throw ::t2p::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + #expr + \
                         " -> " + cudaGetErrorString(_e));                                      \
    }                                                                                           \
  } while (0)

#define T2P_LAUNCH_CHECK() T2P_CUDA(cudaGetLastError())

// Timing / A-B knobs (T2P_DEBUG_SKIP, T2P_DEBUG_DUP, T2P_HALO, T2P_PDL, T2P_STEP_*, ...) are environment variables that
// only a library compiled with -DT2P_TIMING_KNOBS reads (libt2p_knobs.so, built for tools/ and `bench.py --lib knobs`).
// The shipped libt2p.so never looks at the environment: a stray variable cannot drop work from a timed region.
inline int env_knob(const char* name, int dflt) {
#ifdef T2P_TIMING_KNOBS
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
#else
  (void)name;
  return dflt;
#endif
}
inline bool env_knob_set(const char* name) {
#ifdef T2P_TIMING_KNOBS
  return getenv(name) != nullptr;
#else
  (void)name;
  return false;
#endif
}

// ---- programmatic dependent launch (PDL).  A kernel launched through launch_pdl() may become resident while
// its predecessor in the stream is still draining: it calls pdl_trigger() first (its own successor may be
// scheduled once every CTA of this grid has started), runs its prologue (barrier init, TMEM allocation,
// tensor-map prefetch) and calls pdl_wait() BEFORE its first global-memory access that depends on a predecessor;
// pdl_wait() returns when the preceding grids have completed and flushed.  Captured into CUDA graphs as
// programmatic edges.  Measured on B200 (profiles/r01_pdl_ab.txt): 3 % SLOWER for this network (27.58 vs 26.74 ms per
// PC iteration), so it is opt-in (T2P_PDL=<class mask>, 15 = all); by default the same kernels launch with ordinary stream ordering.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// CLS: kernel class bit tested against the T2P_PDL mask (1 tcgen05 GEMMs, 2 gn_finalize, 4 gn_apply, 8 everything else)
template <int CLS = 8, typename... P, typename... A>
inline void launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  static const bool on = (env_knob("T2P_PDL", 0) & CLS) != 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = on ? 1 : 0;
  T2P_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...));
}

// same with the programmatic-serialisation attribute decided per launch (size-dependent experiments)
template <typename... P, typename... A>
inline void launch_pdl_dyn(bool on, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = on ? 1 : 0;
  T2P_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...));
}
#endif

// Kernel attributes (opt-in shared memory, carve-out), occupancy results and SM counts belong to a DEVICE: caches of
// them are indexed by the current device, so a process that drives several GPUs configures each one.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  T2P_CUDA(cudaGetDevice(&dev));
  T2P_CHECK(dev >= 0 && dev < kMaxDevices, "device index out of range");
  return dev;
}
inline int device_sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (!n[dev]) T2P_CUDA(cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev));
  return n[dev];
}
// true exactly once per (call site's flag array, device)
inline bool first_use_on_device(bool (&flags)[kMaxDevices]) {
  const int dev = current_device();
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace t2p
