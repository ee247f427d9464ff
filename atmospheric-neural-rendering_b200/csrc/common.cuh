// common.cuh -- error handling and small device helpers for libatmonr_b200.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "device_math.cuh"

namespace atm {

extern thread_local char g_last_error[512];

inline int fail(const char* what, const char* detail) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, detail ? detail : "");
  return -1;
}

#define ATM_CHECK_LAUNCH(name)                                      \
  do {                                                              \
    cudaError_t e__ = cudaGetLastError();                           \
    if (e__ != cudaSuccess) return atm::fail(name, cudaGetErrorString(e__)); \
  } while (0)

#define ATM_REQUIRE(cond, name, msg) \
  do {                               \
    if (!(cond)) return atm::fail(name, msg); \
  } while (0)

constexpr int kTile = 128;  // samples per CTA tile == threads per CTA in the MLP kernels

__device__ __forceinline__ float round_f16(float v) { return __half2float(__float2half_rn(v)); }

// Vector reduction into global memory: one RED for both features of a table entry (sm_90+).
__device__ __forceinline__ void red_add_f32x2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_prod(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inclusive scans across the warp
__device__ __forceinline__ float warp_scan_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}
__device__ __forceinline__ float warp_scan_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

inline int grid_for(int64_t n, int block, int max_blocks = 1 << 30) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace atm
