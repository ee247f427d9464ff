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

// One aligned group of FOUR consecutive bins of a ray: stratified sample -> ECEF -> geodetic ->
// [0,1]^3 (instant_ngp.py:139-160). q = ray * (N/4) + group. Shared by k_ngp_sample_points4 and by
// the sampler warp of the field backward kernel (which computes the NEXT batch's points).
struct SamplerJob {
  atmonr_frame_t f;
  GeoFrame gf;
  const float* o;
  const float* d;
  const float* len;
  const float* u;      // mode 1
  const float* bins;   // optional bin edges (N floats, 16-byte aligned)
  int64_t groups;      // B * N / 4; 0 = no job
  int N4, mode;
  uint64_t seed, base;
  float alt_compress;
  float* x01;
  float* z;
};

__device__ __forceinline__ void sample_group4(const SamplerJob& j, int64_t q) {
  int64_t ray;
  if (j.groups <= 0xffffffffll) ray = (int64_t)((uint32_t)q / (uint32_t)j.N4);
  else ray = q / j.N4;
  const int g4 = (int)(q - ray * j.N4);
  const int N = j.N4 * 4, i0 = g4 * 4;
  float t[4] = {0.5f, 0.5f, 0.5f, 0.5f};
  if (j.mode == 1) {
    const float4 uu = reinterpret_cast<const float4*>(j.u)[q];
    t[0] = uu.x, t[1] = uu.y, t[2] = uu.z, t[3] = uu.w;
  } else if (j.mode == 2) {
    philox_uniform4(j.seed, j.base + (uint64_t)ray, (uint32_t)g4, t);
  }
  float lo[4];
  if (j.bins) {
    const float4 bb = reinterpret_cast<const float4*>(j.bins)[g4];
    lo[0] = bb.x, lo[1] = bb.y, lo[2] = bb.z, lo[3] = bb.w;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) lo[k] = (float)(i0 + k) / (float)N;
  }
  const float ln = j.len[ray];
  const float ox = j.o[ray * 3], oy = j.o[ray * 3 + 1], oz = j.o[ray * 3 + 2];
  const float dx = j.d[ray * 3], dy = j.d[ray * 3 + 1], dz = j.d[ray * 3 + 2];
  float zz[4], out[12];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    zz[k] = stratified_z(lo[k], t[k], N, ln);
    const float px = ox + dx * zz[k], py = oy + dy * zz[k], pz = oz + dz * zz[k];
    float c0 = px, c1 = py, c2 = pz;
    if (j.f.enabled) preprocess_f32(j.f, j.gf, px, py, pz, c0, c1, c2);
    to_unit_cube(c0, c1, c2, j.alt_compress, out[3 * k], out[3 * k + 1], out[3 * k + 2]);
  }
  reinterpret_cast<float4*>(j.z)[q] = make_float4(zz[0], zz[1], zz[2], zz[3]);
  float4* xo = reinterpret_cast<float4*>(j.x01) + 3 * q;
  xo[0] = make_float4(out[0], out[1], out[2], out[3]);
  xo[1] = make_float4(out[4], out[5], out[6], out[7]);
  xo[2] = make_float4(out[8], out[9], out[10], out[11]);
}

// Host side: fill a job (returns false when the fast 4-bin form does not apply).
inline bool make_sampler_job(SamplerJob& j, const atmonr_frame_t* f, const float* origin, const float* dir,
                             const float* len, const float* u, const float* bins, int64_t B, int N, int mode,
                             uint64_t seed, uint64_t ray_index_base, float alt_compress, float* x01, float* z) {
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (N % 4 != 0 || !aligned16(x01) || !aligned16(z) || !aligned16(u) || !aligned16(bins) ||
      B * (int64_t)(N / 4) >= (1ll << 37))
    return false;
  j.f = *f;
  j.gf = make_geo_frame(*f);
  j.o = origin, j.d = dir, j.len = len, j.u = u, j.bins = bins;
  j.groups = B * (int64_t)(N / 4);
  j.N4 = N / 4, j.mode = mode, j.seed = seed, j.base = ray_index_base, j.alt_compress = alt_compress;
  j.x01 = x01, j.z = z;
  return true;
}

inline int grid_for(int64_t n, int block, int max_blocks = 1 << 30) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace atm
