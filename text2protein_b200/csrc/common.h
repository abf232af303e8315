// Host-side common definitions for the t2p CUDA library: error plumbing and tensor views.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
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

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace t2p
