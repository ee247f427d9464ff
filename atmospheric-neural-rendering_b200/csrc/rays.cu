// rays.cu -- the ray table of a granule and the per-step batch gather (sm_100a).
//   atmonr_get_rays     : wgs_84.py:223-290 get_rays -- entry into the atmosphere shell, direction
//                         and length to the surface of every pixel/view
//   atmonr_filter_rays / atmonr_ray_extent / atmonr_normalize_origins :
//                         wgs_84.py:293-339 filter_rays, normalize_rays
//   atmonr_gather_batch : harp2.py:392-420 __getitem__/__getbatch__ -- the seven per-ray gathers of
//                         a batch in one launch
// Interface contract: include/atmonr_b200.h.
#include "common.cuh"
#include "ray_setup.cuh"

namespace atm {

// work layout: [0, n) current length, [n, 2n) altitude of the ray's top end at that length (float64)
// phase 0: first guess; phase 1: one refinement `len *= H / height`; both raise *flag when a ray's
// top end is further than `tol` from the shell. The reference refines EVERY ray of a chunk for as
// long as ANY ray of the chunk is out of tolerance (`(err > tol).any()`), so the loop lives on the
// host and a converged ray keeps being refined with its chunk.
__global__ void k_rays_refine(const float* __restrict__ lat, const float* __restrict__ lon,
                              const float* __restrict__ alt, const float* __restrict__ thetav,
                              const float* __restrict__ phiv, int64_t n, float origin_height, double tol,
                              int phase, double* __restrict__ work, int* __restrict__ flag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  bool out_of_tol = false;
  if (i < n) {
    RaySetup r;
    ray_setup(lat[i], lon[i], alt[i], thetav[i], phiv[i], origin_height, r);
    const double H = (double)origin_height;
    const double len = phase == 0 ? r.len0 : work[i] * H / work[n + i];
    const double h = ray_height(r, len);
    work[i] = len;
    work[n + i] = h;
    out_of_tol = fabs(H - h) > tol;  // false for NaN, like torch's comparison
  }
  if (__any_sync(0xffffffffu, out_of_tol) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

__global__ void k_rays_finish(const float* __restrict__ lat, const float* __restrict__ lon,
                              const float* __restrict__ alt, const float* __restrict__ thetav,
                              const float* __restrict__ phiv, int64_t n, float origin_height,
                              const double* __restrict__ work, float* __restrict__ origin,
                              float* __restrict__ dir, float* __restrict__ len) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  RaySetup r;
  ray_setup(lat[i], lon[i], alt[i], thetav[i], phiv[i], origin_height, r);
  float o[3], d[3], l;
  ray_outputs(r, work[i], o, d, l);
#pragma unroll
  for (int k = 0; k < 3; ++k) origin[3 * i + k] = o[k], dir[3 * i + k] = d[k];
  len[i] = l;
}

// wgs_84.py:293-313: one thread per ray, the mask is a torch.bool tensor (one byte per ray)
__global__ void k_filter_rays(const float* __restrict__ origin, const float* __restrict__ dir,
                              const float* __restrict__ rad, int64_t n, uint8_t* __restrict__ valid) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float o[3] = {origin[3 * i], origin[3 * i + 1], origin[3 * i + 2]};
  const float d[3] = {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
  valid[i] = ray_is_valid(o, d, rad[i]) ? 1 : 0;
}

// wgs_84.py:333-335: bounding box of the ray origins and lower ends. Grid-stride partial boxes,
// shuffle + shared-memory reduction per block; partial layout per block: hi[3], lo[3], NaN axes, pad.
constexpr int kExtentThreads = 256;
constexpr int kExtentMaxBlocks = 148 * 8;
static_assert(kExtentMaxBlocks * 8 * sizeof(float) == ATMONR_RAY_EXTENT_WORK_BYTES, "header constant out of date");

__device__ __forceinline__ void extent_warp_reduce(RayExtent& e) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    RayExtent t;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      t.hi[k] = __shfl_xor_sync(0xffffffffu, e.hi[k], o);
      t.lo[k] = __shfl_xor_sync(0xffffffffu, e.lo[k], o);
    }
    t.nan_axes = __shfl_xor_sync(0xffffffffu, e.nan_axes, o);
    extent_merge(e, t);
  }
}

__device__ __forceinline__ void extent_block_reduce(RayExtent& e, RayExtent* warp_boxes) {
  extent_warp_reduce(e);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_boxes[warp] = e;
  __syncthreads();
  if (warp == 0) {
    extent_init(e);
    if (lane < (int)(blockDim.x >> 5)) e = warp_boxes[lane];
    extent_warp_reduce(e);
  }
}

__global__ void __launch_bounds__(kExtentThreads)
k_ray_extent_partial(const float* __restrict__ origin, const float* __restrict__ dir, const float* __restrict__ len,
                     int64_t n, float* __restrict__ partial) {
  __shared__ RayExtent warp_boxes[kExtentThreads / 32];
  RayExtent e;
  extent_init(e);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float o[3] = {origin[3 * i], origin[3 * i + 1], origin[3 * i + 2]};
    const float d[3] = {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
    extent_add_ray(e, o, d, len[i]);
  }
  extent_block_reduce(e, warp_boxes);
  if (threadIdx.x == 0) {
    float* p = partial + 8 * (size_t)blockIdx.x;
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = e.hi[k], p[3 + k] = e.lo[k];
    p[6] = __uint_as_float(e.nan_axes);
    p[7] = 0.0f;
  }
}

// one block: partial boxes -> hi_lo[0..3) = max, hi_lo[3..6) = min (NaN on an axis that saw a NaN)
__global__ void __launch_bounds__(kExtentThreads)
k_ray_extent_final(const float* __restrict__ partial, int n_partial, float* __restrict__ hi_lo) {
  __shared__ RayExtent warp_boxes[kExtentThreads / 32];
  RayExtent e;
  extent_init(e);
  for (int b = threadIdx.x; b < n_partial; b += blockDim.x) {
    RayExtent t;
    const float* p = partial + 8 * (size_t)b;
#pragma unroll
    for (int k = 0; k < 3; ++k) t.hi[k] = p[k], t.lo[k] = p[3 + k];
    t.nan_axes = __float_as_uint(p[6]);
    extent_merge(e, t);
  }
  extent_block_reduce(e, warp_boxes);
  if (threadIdx.x == 0) {
    const float nan = __int_as_float(0x7fc00000);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const bool bad = (e.nan_axes >> k) & 1u;
      hi_lo[k] = bad ? nan : e.hi[k];
      hi_lo[3 + k] = bad ? nan : e.lo[k];
    }
  }
}

// wgs_84.py:338, one thread per coordinate; offset is the reference's float64[3] device tensor
__global__ void k_normalize_origins(const float* __restrict__ origin, int64_t n3, const double* __restrict__ offset,
                                    double scale, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n3) return;
  out[i] = normalize_coord(origin[i], offset[i % 3], scale);
}

// one thread per gathered ray; negative indices count from the end like torch indexing
__global__ void k_gather_batch(const float* __restrict__ origin, const float* __restrict__ dir,
                               const float* __restrict__ alt, const float* __restrict__ rad,
                               const float* __restrict__ len, const int32_t* __restrict__ ray_idx,
                               const int64_t* __restrict__ band, const int64_t* __restrict__ index, int64_t B,
                               int64_t R, float* __restrict__ o_origin, float* __restrict__ o_dir,
                               float* __restrict__ o_alt, float* __restrict__ o_rad, float* __restrict__ o_len,
                               int32_t* __restrict__ o_idx, int64_t* __restrict__ o_band, int* __restrict__ bad) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B) return;
  int64_t j = index[i];
  if (j < 0) j += R;
  if (j < 0 || j >= R) {  // reported by the caller; nothing is read out of bounds
    atomicOr(bad, 1);
    j = 0;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) o_origin[3 * i + k] = origin[3 * j + k], o_dir[3 * i + k] = dir[3 * j + k];
  if (o_alt) o_alt[i] = alt[j];
  o_rad[i] = rad[j];
  o_len[i] = len[j];
  if (o_idx) o_idx[i] = ray_idx[j];
  o_band[i] = band[j];
}

}  // namespace atm

using namespace atm;
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int atmonr_get_rays(const float* lat, const float* lon, const float* alt, const float* thetav,
                    const float* phiv, int64_t n, float ray_origin_height, double tol, int max_iters,
                    float* origin, float* dir, float* len, void* work, int* n_iters_host, void* stream) {
  ATM_REQUIRE(n >= 0 && max_iters >= 0, "atmonr_get_rays", "negative size");
  ATM_REQUIRE(n == 0 || (lat && lon && alt && thetav && phiv && origin && dir && len && work), "atmonr_get_rays",
              "null pointer");
  if (n_iters_host) *n_iters_host = 0;
  if (n == 0) return 0;
  double* w = reinterpret_cast<double*>(work);
  int* flag = reinterpret_cast<int*>(w + 2 * n);
  const int grid = grid_for(n, 128);
  int iters = 0;
  for (int phase = 0;; phase = 1) {
    if (cudaMemsetAsync(flag, 0, sizeof(int), S(stream)) != cudaSuccess) return fail("atmonr_get_rays", "memset failed");
    k_rays_refine<<<grid, 128, 0, S(stream)>>>(lat, lon, alt, thetav, phiv, n, ray_origin_height, tol, phase, w, flag);
    ATM_CHECK_LAUNCH("atmonr_get_rays");
    if (phase == 1) ++iters;
    int any = 0;
    cudaError_t e = cudaMemcpyAsync(&any, flag, sizeof(int), cudaMemcpyDeviceToHost, S(stream));
    if (e == cudaSuccess) e = cudaStreamSynchronize(S(stream));
    if (e != cudaSuccess) return fail("atmonr_get_rays", cudaGetErrorString(e));
    if (!any || iters >= max_iters) break;
  }
  k_rays_finish<<<grid, 128, 0, S(stream)>>>(lat, lon, alt, thetav, phiv, n, ray_origin_height, w, origin, dir, len);
  ATM_CHECK_LAUNCH("atmonr_get_rays");
  if (n_iters_host) *n_iters_host = iters;
  return 0;
}

int atmonr_filter_rays(const float* origin, const float* dir, const float* rad, int64_t n, uint8_t* valid,
                       void* stream) {
  ATM_REQUIRE(n >= 0, "atmonr_filter_rays", "negative size");
  if (n == 0) return 0;
  ATM_REQUIRE(origin && dir && rad && valid, "atmonr_filter_rays", "null pointer");
  k_filter_rays<<<grid_for(n, 256), 256, 0, S(stream)>>>(origin, dir, rad, n, valid);
  ATM_CHECK_LAUNCH("atmonr_filter_rays");
  return 0;
}

int atmonr_ray_extent(const float* origin, const float* dir, const float* len, int64_t n, float* hi_lo, void* work,
                      void* stream) {
  ATM_REQUIRE(n > 0, "atmonr_ray_extent", "the bounding box of no rays is undefined (torch.max of an empty tensor raises)");
  ATM_REQUIRE(origin && dir && len && hi_lo && work, "atmonr_ray_extent", "null pointer");
  const int blocks = grid_for(n, kExtentThreads, kExtentMaxBlocks);
  float* partial = reinterpret_cast<float*>(work);
  k_ray_extent_partial<<<blocks, kExtentThreads, 0, S(stream)>>>(origin, dir, len, n, partial);
  ATM_CHECK_LAUNCH("atmonr_ray_extent");
  k_ray_extent_final<<<1, kExtentThreads, 0, S(stream)>>>(partial, blocks, hi_lo);
  ATM_CHECK_LAUNCH("atmonr_ray_extent");
  return 0;
}

int atmonr_normalize_origins(const float* origin, int64_t n, const double* offset, double scale, float* out,
                             void* stream) {
  ATM_REQUIRE(n >= 0, "atmonr_normalize_origins", "negative size");
  if (n == 0) return 0;
  ATM_REQUIRE(origin && offset && out, "atmonr_normalize_origins", "null pointer");
  k_normalize_origins<<<grid_for(3 * n, 256), 256, 0, S(stream)>>>(origin, 3 * n, offset, scale, out);
  ATM_CHECK_LAUNCH("atmonr_normalize_origins");
  return 0;
}

int atmonr_gather_batch(const float* origin, const float* dir, const float* alt, const float* rad,
                        const float* len, const int32_t* ray_idx, const int64_t* band, const int64_t* index,
                        int64_t B, int64_t R, float* out_origin, float* out_dir, float* out_alt, float* out_rad,
                        float* out_len, int32_t* out_ray_idx, int64_t* out_band, int* bad_index, void* stream) {
  ATM_REQUIRE(B >= 0 && R >= 0, "atmonr_gather_batch", "negative size");
  if (B == 0) return 0;
  ATM_REQUIRE(R > 0, "atmonr_gather_batch", "index into an empty ray table");
  ATM_REQUIRE(origin && dir && rad && len && band && index && out_origin && out_dir && out_rad && out_len &&
                  out_band && bad_index,
              "atmonr_gather_batch", "null pointer");
  ATM_REQUIRE((out_alt == nullptr) == (alt == nullptr) || out_alt == nullptr, "atmonr_gather_batch", "out_alt without alt");
  ATM_REQUIRE(out_ray_idx == nullptr || ray_idx != nullptr, "atmonr_gather_batch", "out_ray_idx without ray_idx");
  k_gather_batch<<<grid_for(B, 256), 256, 0, S(stream)>>>(origin, dir, alt, rad, len, ray_idx, band, index, B, R,
                                                          out_origin, out_dir, out_alt, out_rad, out_len, out_ray_idx,
                                                          out_band, bad_index);
  ATM_CHECK_LAUNCH("atmonr_gather_batch");
  return 0;
}

}  // extern "C"
