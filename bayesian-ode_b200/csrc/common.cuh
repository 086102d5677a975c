// Shared helpers for the bode_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/bode_b200.h"

namespace bode {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define BODE_CUDA(call)                                            \
  do {                                                             \
    int _st = ::bode::check_cuda((call), #call);                   \
    if (_st != BODE_OK) return _st;                                \
  } while (0)

#define BODE_REQUIRE(cond, ...)                                    \
  do {                                                             \
    if (!(cond)) {                                                 \
      ::bode::set_error(__VA_ARGS__);                              \
      return BODE_ERR_ARG;                                         \
    }                                                              \
  } while (0)

// 2^x on the SFU (MUFU.EX2); rel. error <= 2^-22, flushes denormals.
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace bode
