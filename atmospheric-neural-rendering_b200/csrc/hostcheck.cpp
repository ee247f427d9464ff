// hostcheck.cpp -- host build of the per-sample arithmetic in device_math.cuh.
//
// NOT a CPU fallback: nothing in the product imports this library. The CPU test-suite uses
// it to check, without a GPU, that the exact code the kernels execute (indexing, geodesy,
// Philox, stratified sampling) agrees with the oracle bit for bit / within tolerance.
#include <vector>

#include "device_math.cuh"
#include "nerf_points.cuh"
#include "ray_setup.cuh"

template <int D>
static void indices_impl(const atmonr_grid_t* g, const float* x, int xs, int64_t M, uint32_t* idx, float* w) {
  for (int64_t i = 0; i < M; ++i) {
    float p[D];
    for (int k = 0; k < D; ++k) p[k] = x[i * xs + k];
    for (int l = 0; l < g->n_levels; ++l) {
      uint32_t cell[D];
      float frac[D];
      atm::grid_cell<D>(p, g->scale[l], cell, frac);
      for (int c = 0; c < (1 << D); ++c) {
        uint32_t e;
        float wc;
        atm::grid_corner<D>(cell, frac, c, g->res[l], g->size[l], e, wc);
        const int64_t at = (i * g->n_levels + l) * (1 << D) + c;
        idx[at] = g->offset[l] + e;
        if (w) w[at] = wc;
      }
    }
  }
}

extern "C" {

void hc_preprocess_f32(const atmonr_frame_t* f, const float* p, float* out, int64_t n) {
  const atm::GeoFrame gf = atm::make_geo_frame(*f);
  for (int64_t i = 0; i < n; ++i)
    atm::preprocess_f32(*f, gf, p[3 * i], p[3 * i + 1], p[3 * i + 2], out[3 * i], out[3 * i + 1], out[3 * i + 2]);
}

void hc_preprocess_f64(const atmonr_frame_t* f, const double* p, double* out, int64_t n) {
  const atm::GeoFrame gf = atm::make_geo_frame(*f);
  for (int64_t i = 0; i < n; ++i)
    atm::preprocess_f64(*f, gf, p[3 * i], p[3 * i + 1], p[3 * i + 2], out[3 * i], out[3 * i + 1], out[3 * i + 2]);
}

void hc_ngp_sample_points(const atmonr_frame_t* f, const float* o, const float* d, const float* len,
                          const float* u, const float* bins, int64_t B, int N, int mode, uint64_t seed,
                          uint64_t base, float alt_compress, float* x01, float* z) {
  const atm::GeoFrame gf = atm::make_geo_frame(*f);
  for (int64_t ray = 0; ray < B; ++ray)
    for (int i = 0; i < N; ++i) {
      const int64_t idx = ray * N + i;
      const float t = mode == 0 ? 0.5f : (mode == 1 ? u[idx] : atm::philox_uniform(seed, base + ray, (uint32_t)i));
      const float lo = bins ? bins[i] : (float)i / (float)N;
      const float zz = atm::stratified_z(lo, t, N, len[ray]);
      z[idx] = zz;
      float p[3];
      for (int k = 0; k < 3; ++k) p[k] = o[ray * 3 + k] + d[ray * 3 + k] * zz;
      float c0 = p[0], c1 = p[1], c2 = p[2];
      if (f->enabled) atm::preprocess_f32(*f, gf, p[0], p[1], p[2], c0, c1, c2);
      atm::to_unit_cube(c0, c1, c2, alt_compress, x01[idx * 3], x01[idx * 3 + 1], x01[idx * 3 + 2]);
    }
}

void hc_hashgrid_indices(const atmonr_grid_t* g, const float* x, int xs, int64_t M, uint32_t* idx, float* w) {
  if (g->n_dims == 2) indices_impl<2>(g, x, xs, M, idx, w);
  else if (g->n_dims == 3) indices_impl<3>(g, x, xs, M, idx, w);
  else indices_impl<4>(g, x, xs, M, idx, w);
}

void hc_philox(uint64_t seed, uint64_t ray0, int64_t B, int N, float* out) {
  for (int64_t r = 0; r < B; ++r)
    for (int i = 0; i < N; ++i) out[r * N + i] = atm::philox_uniform(seed, ray0 + r, (uint32_t)i);
}

// wgs_84.py:223-290 with the loop structure of atmonr_get_rays (every ray of the call refined
// while any ray is out of tolerance); returns the number of refinements
int hc_get_rays(const float* lat, const float* lon, const float* alt, const float* thetav, const float* phiv,
                int64_t n, float origin_height, double tol, int max_iters, float* origin, float* dir, float* len) {
  std::vector<atm::RaySetup> rs(n);
  std::vector<double> cur(n), height(n);
  const double H = (double)origin_height;
  bool any = false;
  for (int64_t i = 0; i < n; ++i) {
    atm::ray_setup(lat[i], lon[i], alt[i], thetav[i], phiv[i], origin_height, rs[i]);
    cur[i] = rs[i].len0;
    height[i] = atm::ray_height(rs[i], cur[i]);
    any = any || fabs(H - height[i]) > tol;
  }
  int iters = 0;
  while (iters < max_iters && any) {
    any = false;
    for (int64_t i = 0; i < n; ++i) {
      cur[i] = cur[i] * H / height[i];
      height[i] = atm::ray_height(rs[i], cur[i]);
      any = any || fabs(H - height[i]) > tol;
    }
    ++iters;
  }
  for (int64_t i = 0; i < n; ++i) atm::ray_outputs(rs[i], cur[i], origin + 3 * i, dir + 3 * i, len[i]);
  return iters;
}

// wgs_84.py:293-339 with the per-ray functions of csrc/rays.cu (k_filter_rays, k_ray_extent_*, k_normalize_origins)
void hc_filter_rays(const float* origin, const float* dir, const float* rad, int64_t n, uint8_t* valid) {
  for (int64_t i = 0; i < n; ++i) valid[i] = atm::ray_is_valid(origin + 3 * i, dir + 3 * i, rad[i]) ? 1 : 0;
}

void hc_ray_extent(const float* origin, const float* dir, const float* len, int64_t n, float* hi_lo) {
  // two partial boxes merged at the end, like the kernel's partial / final passes
  atm::RayExtent a, b;
  atm::extent_init(a), atm::extent_init(b);
  for (int64_t i = 0; i < n; ++i) atm::extent_add_ray((i & 1) ? b : a, origin + 3 * i, dir + 3 * i, len[i]);
  atm::extent_merge(a, b);
  for (int k = 0; k < 3; ++k) {
    const bool bad = (a.nan_axes >> k) & 1u;
    hi_lo[k] = bad ? NAN : a.hi[k];
    hi_lo[3 + k] = bad ? NAN : a.lo[k];
  }
}

void hc_normalize_origins(const float* origin, int64_t n, const double* offset, double scale, float* out) {
  for (int64_t i = 0; i < 3 * n; ++i) out[i] = atm::normalize_coord(origin[i], offset[i % 3], scale);
}

// csrc/nerf_points.cu, per sample on the host: rows [pos | dir] and preprocessed points; then dL/dz
void hc_nerf_encode(const atmonr_frame_t* f, const float* o, const float* d, const float* z, int64_t B, int N,
                    const int32_t* pos_freqs, int dir_freqs, float* x, int ldx, float* pts_n) {
  const atm::GeoFrame gf = atm::make_geo_frame(*f);
  atm::NerfEncCfg c;
  atm::nerf_enc_cfg(pos_freqs, dir_freqs, c);
  for (int64_t i = 0; i < B * N; ++i)
    atm::nerf_encode_sample(*f, gf, o + 3 * (i / N), d + 3 * (i / N), z[i], c, x + i * ldx, pts_n + 3 * i);
}

void hc_nerf_encode_bwd(const atmonr_frame_t* f, const float* o, const float* d, const float* z, const float* pts_n,
                        const float* g, int ldg, int64_t B, int N, const int32_t* pos_freqs, float* gz) {
  const atm::GeoFrame gf = atm::make_geo_frame(*f);
  atm::NerfEncCfg c;
  atm::nerf_enc_cfg(pos_freqs, 0, c);
  for (int64_t i = 0; i < B * N; ++i)
    gz[i] = atm::nerf_encode_sample_bwd(*f, gf, o + 3 * (i / N), d + 3 * (i / N), z[i], pts_n + 3 * i, g + i * ldg, c);
}

}  // extern "C"
