// rays.cu -- the ray table of a granule and the per-step batch gather (sm_100a).
//   atmonr_get_rays     : wgs_84.py:223-290 get_rays -- entry into the atmosphere shell, direction
//                         and length to the surface of every pixel/view
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
