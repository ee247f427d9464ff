// device_math.cuh -- per-sample arithmetic shared by every kernel of libatmonr_b200.
//
// Everything here is __host__ __device__ and free of CUDA-only intrinsics, so the same code
// is compiled by g++ into a host self-check library (csrc/hostcheck.cpp; used by the CPU test
// suite to verify index/geodesy logic without a GPU). The library is built with
// -fmad=false (nvcc) / -ffp-contract=off (g++): every a*b+c below is two rounded operations
// exactly like the eager torch ops of the reference; fused multiply-adds are explicit.
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/atmonr_b200.h"

#if defined(__CUDACC__)
#define ATM_HD __host__ __device__ __forceinline__
#else
#define ATM_HD inline
#endif

namespace atm {

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator: one (seed, ray, bin) -> one uniform in [0,1).
// Partition-invariant: the draw of a sample depends only on its global ray index and bin.
// ---------------------------------------------------------------------------------------
ATM_HD void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  c[1] = (uint32_t)p1;
  c[3] = (uint32_t)p0;
  c[0] = n0;
  c[2] = n2;
}

// One Philox block gives the uniforms of FOUR consecutive bins: counter word 2 is bin / 4 and the
// bin's draw is output word bin % 4, so a thread that handles an aligned group of four bins runs
// the generator once (philox_uniform4); philox_uniform is the same stream, one bin at a time.
ATM_HD void philox_uniform4(uint64_t seed, uint64_t ray, uint32_t bin_group, float (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ray, (uint32_t)(ray >> 32), bin_group, 0x5eedu};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) out[k] = (float)(c[k] >> 8) * (1.0f / 16777216.0f);  // 24 bits -> [0,1)
}

ATM_HD float philox_uniform(uint64_t seed, uint64_t ray, uint32_t bin) {
  float v[4];
  philox_uniform4(seed, ray, bin >> 2, v);
  const uint32_t k = bin & 3u;
  return k == 0 ? v[0] : (k == 1 ? v[1] : (k == 2 ? v[2] : v[3]));
}

// ---------------------------------------------------------------------------------------
// samplers.py:34-45: stratified distance and the point on the ray.
// ---------------------------------------------------------------------------------------
ATM_HD float stratified_z(float bin_lo, float t, int n_bins, float len) {
  return (bin_lo + t / (float)n_bins) * len;
}

// ---------------------------------------------------------------------------------------
// wgs_84.py:56-97 cartesian_to_horizontal: one Bowring iteration, float64.
// ---------------------------------------------------------------------------------------
#define ATM_WGS_A 6378137.0
#define ATM_WGS_B 6356752.314245

ATM_HD double atm_rsqrt(double v) {
#if defined(__CUDA_ARCH__)
  return rsqrt(v);
#else
  return 1.0 / sqrt(v);
#endif
}

ATM_HD void ecef_to_geodetic(double x, double y, double z, double& lat_deg, double& lon_deg,
                             double& alt) {
  // Same quantities as the reference, with the sines / cosines of the two auxiliary angles taken
  // algebraically instead of through atan2 -> sin/cos:
  //   u   = atan2(z/d, a/b)  is only used as sin(u), cos(u)       = p/h, q/h with h = hypot(p, q)
  //   phi = atan2(num, den)  is returned, and used as sin/cos(phi) = num/g, den/g
  //   cos(lam) = x / d
  // This removes one atan2 and five sin/cos evaluations per sample (the kernel is bound by the
  // FP64 pipe); the results agree with the literal form to a few float64 ulps (~1e-15 relative,
  // five orders of magnitude below the float32 rounding of the outputs).
  const double A = ATM_WGS_A, B = ATM_WGS_B;
  const double E_SQ = (A * A - B * B) / (A * A);
  const double EP_SQ = (A * A - B * B) / (B * B);
  const double PI = 3.141592653589793;
  // Divisions and square roots are the bulk of the FP64 work (each a ~25-45 instruction sequence):
  // every 1/sqrt(.) below is one reciprocal-square-root sequence and every quotient by the same
  // denominator shares it, 4 rsqrt + 1 division instead of 4 sqrt + 8 divisions.
  const double lam = atan2(y, x);
  const double dd = x * x + y * y;
  const double rd = atm_rsqrt(dd);      // 1 / d
  const double d = dd * rd;             // sqrt(x^2 + y^2)
  const double p = z * rd, q = A / B;
  const double rh = atm_rsqrt(p * p + q * q);
  const double su = p * rh, cu = q * rh;
  const double num = z + (EP_SQ * B) * ((su * su) * su);
  const double den = d - (E_SQ * A) * ((cu * cu) * cu);
  const double phi = atan2(num, den);
  const double rg = atm_rsqrt(num * num + den * den);
  const double sp = num * rg, cp = den * rg;
  const double n = A * atm_rsqrt(1.0 - (E_SQ * (sp * sp)));
  // x == 0 is the reference's singular meridian (cos(lam) ~ 6e-17): keep its literal behaviour
  const double cl = x != 0.0 ? x * rd : cos(lam);
  alt = x / (cp * cl) - n;
  lat_deg = phi * 180.0 / PI;
  lon_deg = lam * 180.0 / PI;
}

// ---------------------------------------------------------------------------------------
// The same conversion, specialised to points near a reference direction (the granule centre).
// ---------------------------------------------------------------------------------------
// The sampler evaluates the conversion for every sample of every step, and the kernel is bound by
// FP64 instruction issue: the two atan2 calls and the six divisions of the literal form are ~60 %
// of its instructions. Every sample of a granule lies within a few degrees of the granule centre
// (lat0, lon0), so both angles are taken RELATIVE to it:
//   lam - lam0 = asin((y cos lam0 - x sin lam0) / d)        phi - phi0 = asin(sp cos phi0 - cp sin phi0)
// with sp, cp = sin/cos(phi) (already needed) and asin of a small argument a 9-term odd series
// (|s| <= 0.17, i.e. 9.8 degrees: truncation < 7e-19; real HARP2 swaths stay below 8 degrees);
// divisions by constants become multiplications by host-computed reciprocals and the one true
// division (1 / cos phi) a reciprocal with two Newton steps. Points outside the window (or x == 0)
// take the literal form above. Agreement with it: a few float64 ulps (~1e-15 relative), i.e. the
// float32 outputs differ in the last bit for ~1e-5 of the samples (tests/test_abi_and_host.py).
struct GeoFrame {
  double sin_lam0, cos_lam0, lam0;  // reference meridian, radians
  double sin_phi0, cos_phi0, phi0;  // reference parallel, radians
  double lat_k, lon_k, alt_k;       // 2 / lat_range, 2 / lon_range, 2 / origin_height
};

inline GeoFrame make_geo_frame(const atmonr_frame_t& f) {
  GeoFrame g{};
  if (!f.enabled) return g;
  const double PI = 3.141592653589793;
  const double lat0 = f.lat_min + 0.5 * f.lat_range;
  double lon0 = f.lon_min + 0.5 * f.lon_range;
  if (f.shift_lon) lon0 += 180.0;  // the frame's longitudes are shifted by 180 degrees (harp2.py:366-370)
  lon0 = fmod(lon0, 360.0);
  if (lon0 > 180.0) lon0 -= 360.0;
  if (lon0 <= -180.0) lon0 += 360.0;
  g.lam0 = lon0 * PI / 180.0, g.phi0 = lat0 * PI / 180.0;
  g.sin_lam0 = sin(g.lam0), g.cos_lam0 = cos(g.lam0);
  g.sin_phi0 = sin(g.phi0), g.cos_phi0 = cos(g.phi0);
  g.lat_k = 2.0 / f.lat_range, g.lon_k = 2.0 / f.lon_range, g.alt_k = 2.0 / f.origin_height;
  return g;
}

// 1/sqrt(v) and 1/v for normal positive v: hardware seed (about 20 bits) + two Newton steps.
ATM_HD double fast_rsqrt(double v) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
  const double h = 0.5 * v;
  y = fma(y, fma(-(h * y), y, 0.5), y);  // y + y (1/2 - h y^2)
  y = fma(y, fma(-(h * y), y, 0.5), y);
  return y;
#else
  return 1.0 / sqrt(v);
#endif
}
ATM_HD double fast_rcp(double v) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
  r = fma(r, fma(-v, r, 1.0), r);
  r = fma(r, fma(-v, r, 1.0), r);
  return r;
#else
  return 1.0 / v;
#endif
}

// asin(s) for |s| <= 0.17: s + s^3/6 + 3 s^5/40 + ... (9 terms)
ATM_HD double asin_small(double s) {
  const double t = s * s;
  double p = 12155.0 / 1245184.0;
  p = fma(p, t, 6435.0 / 557056.0);
  p = fma(p, t, 143.0 / 10240.0);
  p = fma(p, t, 231.0 / 13312.0);
  p = fma(p, t, 63.0 / 2816.0);
  p = fma(p, t, 35.0 / 1152.0);
  p = fma(p, t, 5.0 / 112.0);
  p = fma(p, t, 3.0 / 40.0);
  p = fma(p, t, 1.0 / 6.0);
  return fma(s * t, p, s);
}

#define ATM_GEO_WINDOW 0.17

ATM_HD void ecef_to_geodetic_local(const GeoFrame& g, double x, double y, double z, double& lat_deg,
                                   double& lon_deg, double& alt) {
  const double A = ATM_WGS_A, B = ATM_WGS_B;
  const double E_SQ = (A * A - B * B) / (A * A);
  const double EP_SQ = (A * A - B * B) / (B * B);
  const double PI = 3.141592653589793, R2D = 180.0 / 3.141592653589793;
  const double dd = x * x + y * y;
  const double rd = fast_rsqrt(dd);
  const double d = dd * rd;
  const double p = z * rd, q = A / B;
  const double rh = fast_rsqrt(p * p + q * q);
  const double su = p * rh, cu = q * rh;
  const double num = z + (EP_SQ * B) * ((su * su) * su);
  const double den = d - (E_SQ * A) * ((cu * cu) * cu);
  const double rg = fast_rsqrt(num * num + den * den);
  const double sp = num * rg, cp = den * rg;
  const double s_lam = (y * g.cos_lam0 - x * g.sin_lam0) * rd, c_lam = (x * g.cos_lam0 + y * g.sin_lam0) * rd;
  const double s_phi = sp * g.cos_phi0 - cp * g.sin_phi0, c_phi = cp * g.cos_phi0 + sp * g.sin_phi0;
  const bool near = fabs(s_lam) <= ATM_GEO_WINDOW && fabs(s_phi) <= ATM_GEO_WINDOW && c_lam > 0.0 &&
                    c_phi > 0.0 && x != 0.0 && den > 0.0;
  if (!near) {  // also catches NaN / zero / infinite inputs
    ecef_to_geodetic(x, y, z, lat_deg, lon_deg, alt);
    return;
  }
  double lam = g.lam0 + asin_small(s_lam);
  if (lam > PI) lam -= 2.0 * PI;
  if (lam < -PI) lam += 2.0 * PI;
  const double phi = g.phi0 + asin_small(s_phi);
  const double n = A * fast_rsqrt(1.0 - (E_SQ * (sp * sp)));
  alt = d * fast_rcp(cp) - n;  // x / (cos(phi) cos(lam)) with cos(lam) = x / d
  lat_deg = phi * R2D;
  lon_deg = lam * R2D;
}

ATM_HD double py_mod(double a, double m) {  // torch remainder: sign follows the divisor
  double r = fmod(a, m);
  if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
  return r;
}

// harp2.py:372-386 preprocess_coords after `coords * scale + offset`; returns UNCLIPPED
// float64 normalised (lat, lon, alt).
ATM_HD void horizontal_normalise(const atmonr_frame_t& f, const GeoFrame& g, double x, double y, double z,
                                 double& o0, double& o1, double& o2) {
  double lat, lon, alt;
  ecef_to_geodetic_local(g, x, y, z, lat, lon, alt);
  // torch remainder(lon, 360) - 180 for lon in [-180, 180]: the sign follows the divisor
  if (f.shift_lon) lon = (lon < 0.0 ? lon + 360.0 : lon) - 180.0;
  o0 = (lat - f.lat_min) * g.lat_k - 1.0;
  o1 = (lon - f.lon_min) * g.lon_k - 1.0;
  o2 = alt * g.alt_k - 1.0;
}

ATM_HD float clip1(float v) { return v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v); }
ATM_HD double clip1(double v) { return v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v); }

// float32 flavour (training): the multiply by `scale` is a float32 op (tensor * python
// float keeps float32), the offset add promotes to float64, the result is cast back to
// float32 and clipped.
ATM_HD void preprocess_f32(const atmonr_frame_t& f, const GeoFrame& g, float px, float py, float pz,
                           float& o0, float& o1, float& o2) {
  const float s = (float)f.scale;
  const double x = (double)(px * s) + f.offset[0];
  const double y = (double)(py * s) + f.offset[1];
  const double z = (double)(pz * s) + f.offset[2];
  double a, b, c;
  horizontal_normalise(f, g, x, y, z, a, b, c);
  o0 = clip1((float)a);
  o1 = clip1((float)b);
  o2 = clip1((float)c);
}

// float64 flavour (extract path, scripts/extract.py:205 feeds float64 points).
ATM_HD void preprocess_f64(const atmonr_frame_t& f, const GeoFrame& g, double px, double py, double pz,
                           double& o0, double& o1, double& o2) {
  double a, b, c;
  horizontal_normalise(f, g, px * f.scale + f.offset[0], py * f.scale + f.offset[1],
                       pz * f.scale + f.offset[2], a, b, c);
  o0 = clip1(a);
  o1 = clip1(b);
  o2 = clip1(c);
}

// instant_ngp.py:149,160: [-1,1] -> [0,1], altitude divided by alt_compress_factor.
ATM_HD void to_unit_cube(float c0, float c1, float c2, float alt_compress, float& x0, float& x1,
                         float& x2) {
  x0 = (c0 + 1.0f) / 2.0f;
  x1 = (c1 + 1.0f) / 2.0f;
  x2 = ((c2 + 1.0f) / 2.0f) / alt_compress;
}

// ---------------------------------------------------------------------------------------
// Multiresolution hash grid indexing (tiny-cuda-nn grid.h, upstream-recalled).
// ---------------------------------------------------------------------------------------
// tcnn evaluates exp2f(level * log2f(per_level_scale)) * base - 1 in float32 (and its host and
// device libm disagree in the last bits). Here the expression is evaluated in float64 and rounded
// ONCE to float32, so that independent implementations (this library, the oracle) agree bit for
// bit; the result is within a few float32 ulps of tcnn's.
inline float grid_level_scale(int level, float per_level_scale, int base_resolution) {
  return (float)(exp2((double)level * log2((double)per_level_scale)) * (double)base_resolution - 1.0);
}

template <int D>
ATM_HD uint32_t grid_entry(const uint32_t (&g)[D], uint32_t res, uint32_t size) {
  uint32_t stride = 1, index = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < D; ++k) {
    if (stride <= size) {
      index += g[k] * stride;
      stride *= res;
    }
  }
  if (size < stride) {
    const uint32_t primes[4] = {1u, 2654435761u, 805459861u, 3674653429u};
    index = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < D; ++k) index ^= g[k] * primes[k];
  }
  return index % size;
}

// Cell and fractional position of x in [0,1]^D at one level: pos = fmaf(scale, x, 0.5).
template <int D>
ATM_HD void grid_cell(const float* x, float scale, uint32_t (&cell)[D], float (&frac)[D]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < D; ++k) {
    const float pos = fmaf(scale, x[k], 0.5f);
    const float fl = floorf(pos);
    cell[k] = (uint32_t)(int)fl;
    frac[k] = pos - fl;
  }
}

// Entry index (within the level) and trilinear weight of corner `c` (bit k set -> +1 in dim k).
template <int D>
ATM_HD void grid_corner(const uint32_t (&cell)[D], const float (&frac)[D], int c, uint32_t res,
                        uint32_t size, uint32_t& entry, float& w) {
  uint32_t g[D];
  w = 1.0f;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < D; ++k) {
    if (c & (1 << k)) {
      g[k] = cell[k] + 1u;
      w *= frac[k];
    } else {
      g[k] = cell[k];
      w *= 1.0f - frac[k];
    }
  }
  entry = grid_entry<D>(g, res, size);
}

// tcnn SphericalHarmonics degree 2 on v = 2*d - 1 (4 outputs), upstream-recalled.
ATM_HD void sh_degree2(float dx, float dy, float dz, float (&out)[4]) {
  const float vx = dx * 2.0f - 1.0f, vy = dy * 2.0f - 1.0f, vz = dz * 2.0f - 1.0f;
  out[0] = 0.28209479177387814f;
  out[1] = -0.48860251190291987f * vy;
  out[2] = 0.48860251190291987f * vz;
  out[3] = -0.48860251190291987f * vx;
}

}  // namespace atm
