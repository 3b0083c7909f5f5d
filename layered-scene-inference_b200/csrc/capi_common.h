// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/lsi_b200.h"

namespace lsi {

void set_error(const char* fmt, ...);
void count_launch(unsigned n = 1);
// prepared-weights memo (capi.cu): the version the caller set for this conv call (0 = none), and the device buffer that holds / will hold
// the re-laid-out filter for (w, sig); *hit = its content is current
unsigned long long take_weight_version();
void* prep_cache_get(const void* w, unsigned long long ver, const int* sig, size_t bytes, bool* hit);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Optional per-kernel timing for bench.py's roofline: CUDA events on the launching stream around a launch.
enum KernelKind { kSplatFwd = 0, kNormalize = 1, kBwdTarget = 2, kBwdSource = 3, kConvTc = 4, kConvFp32 = 5, kWgrad = 6,
                  kNnOther = 7, kNumKinds = 8 };
struct ScopedTiming {
  cudaStream_t st; cudaEvent_t a, b; int kind; bool on;
  ScopedTiming(int kind, cudaStream_t s);
  ~ScopedTiming();
};

}  // namespace lsi

#define LSI_REQUIRE(cond, ...)              \
  do {                                      \
    if (!(cond)) {                          \
      lsi::set_error(__VA_ARGS__);          \
      return LSI_B200_EINVAL;               \
    }                                       \
  } while (0)

#define LSI_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      lsi::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return LSI_B200_ECUDA;                                                                  \
    }                                                                                         \
  } while (0)

#define LSI_LAUNCH_CHECK()                                                                    \
  do {                                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess) {                                                                 \
      lsi::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return LSI_B200_ECUDA;                                                                  \
    }                                                                                         \
    lsi::count_launch();                                                                      \
  } while (0)
