// ray_setup.cuh -- per-ray arithmetic of the dataset-construction step: every pixel/view's
// entry into the atmosphere shell (origin at `ray_origin_height`), its direction and its length
// down to the surface (wgs_84.py:223-290 get_rays, with :24-53, :56-97, :100-160, :189-220).
//
// Host-compilable like device_math.cuh (csrc/hostcheck.cpp builds the same functions with g++ so
// the CPU suite can pin them to the reference's known answers). The dtype flow of the reference is
// kept on purpose, it decides the float32 results:
//   * the surface point is computed in float64 and ROUNDED to float32 before it is used;
//   * the local view direction is float64, the local-frame rotation (from 90 - lat, 90 - lon) is
//     evaluated in float32 and only then widened;
//   * the first length guess is the float32 difference H - alt over the float32 cos(theta_v),
//     divided in float64;
//   * the fixed-point refinement `len *= H / height(len)` runs in float64.
#pragma once

#include "device_math.cuh"

namespace atm {

// wgs_84.py:24-53 horizontal_to_cartesian (float64)
ATM_HD void geodetic_to_ecef(double lat_deg, double lon_deg, double h, double& x, double& y, double& z) {
  const double A = ATM_WGS_A, B = ATM_WGS_B;
  const double E_SQ = (A * A - B * B) / (A * A);
  const double PI = 3.141592653589793;
  const double phi = lat_deg * PI / 180.0, lam = lon_deg * PI / 180.0;
  const double sp = sin(phi), cp = cos(phi);
  const double n = A / sqrt(1.0 - (E_SQ * (sp * sp)));
  x = (n + h) * cp * cos(lam);
  y = (n + h) * cp * sin(lam);
  z = (n * (1.0 - E_SQ) + h) * sp;
}

// wgs_84.py:56-97 cartesian_to_horizontal, height only, literal form (this runs once per run: the
// algebraic shortcuts of ecef_to_geodetic are not needed here)
ATM_HD double ecef_height(double x, double y, double z) {
  const double A = ATM_WGS_A, B = ATM_WGS_B;
  const double E_SQ = (A * A - B * B) / (A * A);
  const double EP_SQ = (A * A - B * B) / (B * B);
  const double lam = atan2(y, x);
  const double horiz = sqrt(x * x + y * y);
  const double u = atan2(z / horiz, A / B);
  const double su = sin(u), cu = cos(u);
  const double phi = atan2(z + (EP_SQ * B) * ((su * su) * su), horiz - (E_SQ * A) * ((cu * cu) * cu));
  const double s = sin(phi);
  const double n = A / sqrt(1.0 - (E_SQ * (s * s)));
  return x / (cos(phi) * cos(lam)) - n;
}

struct RaySetup {
  float sx, sy, sz;   // surface point (float32, as the reference stores it)
  double dx, dy, dz;  // unit direction, pointing from the top of the shell to the surface
  double len0;        // first length guess
};

ATM_HD void ray_setup(float lat, float lon, float alt, float thetav, float phiv, float origin_height,
                      RaySetup& r) {
  const double PI = 3.141592653589793;
  const float PI_F = (float)PI;
  double x, y, z;
  geodetic_to_ecef((double)lat, (double)lon, (double)alt, x, y, z);
  r.sx = (float)x, r.sy = (float)y, r.sz = (float)z;
  // local +z-up frame (wgs_84.py:135-160): the rotation of (-theta, -phi) applied to (0,0,1)
  const double t = -(double)thetav * PI / 180.0, p = -(double)phiv * PI / 180.0;
  const double st = sin(t), ct = cos(t), sp = sin(p), cp = cos(p);
  // times the (-1,-1,1) turn between the scene and the WGS convention (wgs_84.py:189-220)
  const double v0 = -(sp * st), v1 = cp * st, v2 = ct;
  // local frame -> ECEF: rotation of (-(90 - lat), -(90 - lon)), evaluated in float32
  const float t2 = -(90.0f - lat) * PI_F / 180.0f, p2 = -(90.0f - lon) * PI_F / 180.0f;
  const float st2 = sinf(t2), ct2 = cosf(t2), sp2 = sinf(p2), cp2 = cosf(p2);
  const double r00 = (double)cp2, r01 = (double)(-sp2 * ct2), r02 = (double)(sp2 * st2);
  const double r10 = (double)sp2, r11 = (double)(cp2 * ct2), r12 = (double)(-cp2 * st2);
  const double r21 = (double)st2, r22 = (double)ct2;
  // the view direction points up from the surface; rays run the other way
  r.dx = -((r00 * v0 + r01 * v1) + r02 * v2);
  r.dy = -((r10 * v0 + r11 * v1) + r12 * v2);
  r.dz = -((0.0 * v0 + r21 * v1) + r22 * v2);
  // `(H - alt) / cos(..).double()`: float32 difference, float32 cosine, float64 quotient
  r.len0 = (double)(origin_height - alt) / (double)cosf(thetav * PI_F / 180.0f);
}

// altitude of the point `len` up the ray from the surface
ATM_HD double ray_height(const RaySetup& r, double len) {
  return ecef_height((double)r.sx - len * r.dx, (double)r.sy - len * r.dy, (double)r.sz - len * r.dz);
}

// final float32 outputs (wgs_84.py:286-290)
ATM_HD void ray_outputs(const RaySetup& r, double len, float* origin, float* dir, float& len_out) {
  const float lf = (float)len;
  len_out = lf;
  origin[0] = (float)((double)r.sx - r.dx * (double)lf);
  origin[1] = (float)((double)r.sy - r.dy * (double)lf);
  origin[2] = (float)((double)r.sz - r.dz * (double)lf);
  dir[0] = (float)r.dx, dir[1] = (float)r.dy, dir[2] = (float)r.dz;
}

// ---- wgs_84.py:293-313 filter_rays / :316-339 normalize_rays -----------------
// A ray is kept when its origin, its direction and its radiance hold no NaN.
ATM_HD bool ray_is_valid(const float* origin, const float* dir, float rad) {
  bool bad = rad != rad;
  for (int k = 0; k < 3; ++k) bad = bad || origin[k] != origin[k] || dir[k] != dir[k];
  return !bad;
}

// lower end of the ray, `origin + dir * len` in float32: product and sum round separately (the
// build disables contraction), as the two eager torch kernels of wgs_84.py:333 do
ATM_HD float ray_end_coord(float origin, float dir, float len) { return origin + dir * len; }

// running bounding box of the points of wgs_84.py:333-335; a NaN coordinate makes torch's max/min
// NaN, so NaNs are counted per axis instead of being dropped by the comparisons
struct RayExtent {
  float hi[3], lo[3];
  uint32_t nan_axes;
};
ATM_HD void extent_init(RayExtent& e) {
  for (int k = 0; k < 3; ++k) e.hi[k] = -INFINITY, e.lo[k] = INFINITY;
  e.nan_axes = 0u;
}
ATM_HD void extent_add(RayExtent& e, int k, float v) {
  if (v != v) {
    e.nan_axes |= 1u << k;
  } else {
    e.hi[k] = v > e.hi[k] ? v : e.hi[k];
    e.lo[k] = v < e.lo[k] ? v : e.lo[k];
  }
}
ATM_HD void extent_add_ray(RayExtent& e, const float* origin, const float* dir, float len) {
  for (int k = 0; k < 3; ++k) {
    extent_add(e, k, origin[k]);
    extent_add(e, k, ray_end_coord(origin[k], dir[k], len));
  }
}
ATM_HD void extent_merge(RayExtent& e, const RayExtent& o) {
  for (int k = 0; k < 3; ++k) {
    e.hi[k] = o.hi[k] > e.hi[k] ? o.hi[k] : e.hi[k];
    e.lo[k] = o.lo[k] < e.lo[k] ? o.lo[k] : e.lo[k];
  }
  e.nan_axes |= o.nan_axes;
}

// wgs_84.py:338 `clamp((origin - offset) / scale, -1, 1).float()`: float32 origin minus float64 offset,
// float64 quotient; a NaN fails both comparisons and stays NaN, like torch.clamp
ATM_HD float normalize_coord(float origin, double offset, double scale) {
  double v = ((double)origin - offset) / scale;
  v = v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v);
  return (float)v;
}

}  // namespace atm
