// Shared helpers for libsvgp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/svgp_b200.h"

namespace svgp {

void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> SVGP_ERR_CUDA

#define SVGP_REQUIRE(cond, msg)                       \
  do {                                                \
    if (!(cond)) {                                    \
      svgp::set_error("%s: %s", __func__, msg);       \
      return SVGP_ERR_ARG;                            \
    }                                                 \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// round-to-nearest TF32 (10 explicit mantissa bits), result kept in an fp32 container
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

}  // namespace svgp
