// nerf_points.cuh -- per-sample arithmetic of the NeRF point encoder (forward and backward), shared by
// csrc/nerf_points.cu (the kernels) and csrc/hostcheck.cpp (host build for the CPU test-suite).
#pragma once

#include "device_math.cuh"

namespace atm {

struct NerfEncCfg {
  int32_t freqs[3];   // frequencies per axis of the point encoding (list-L layout: [sin x L | cos x L] per axis)
  int32_t col0[3];    // first column of each axis
  int32_t pos_width;  // 2 * sum(freqs)
  int32_t dir_freqs;  // int-L layout of the direction: per axis, per frequency [sin, cos]
};

inline int nerf_enc_cfg(const int32_t* pos_freqs, int dir_freqs, NerfEncCfg& c) {
  if (!pos_freqs || dir_freqs < 0) return -1;
  int col = 0;
  for (int a = 0; a < 3; ++a) {
    if (pos_freqs[a] < 0) return -1;
    c.freqs[a] = pos_freqs[a];
    c.col0[a] = col;
    col += 2 * pos_freqs[a];
  }
  c.pos_width = col;
  c.dir_freqs = dir_freqs;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// forward, one sample: row = [pos encoding | dir encoding], pn = preprocessed point
// ------------------------------------------------------------------------------------------------
// sin and cos of one phase with ONE argument reduction on the device (the phases reach 2^13 pi)
ATM_HD void sin_cos(float arg, float& sn, float& cs) {
#if defined(__CUDA_ARCH__)
  sincosf(arg, &sn, &cs);
#else
  sn = sinf(arg), cs = cosf(arg);
#endif
}

ATM_HD void nerf_encode_sample(const atmonr_frame_t& f, const GeoFrame& gf, const float* o, const float* d, float zz,
                               const NerfEncCfg& cfg, float* row, float* pn) {
  // samplers.py:101 / :45: origin + dir * z, two float32 operations
  float p[3] = {o[0] + d[0] * zz, o[1] + d[1] * zz, o[2] + d[2] * zz};
  if (f.enabled) preprocess_f32(f, gf, p[0], p[1], p[2], p[0], p[1], p[2]);
  pn[0] = p[0], pn[1] = p[1], pn[2] = p[2];
  const float PI_F = 3.14159265358979323846f;
  for (int a = 0; a < 3; ++a) {
    const int L = cfg.freqs[a];
    float fr = 1.0f;
    for (int l = 0; l < L; ++l) {
      float sn, cs;
      sin_cos((fr * PI_F) * p[a], sn, cs);
      row[cfg.col0[a] + l] = sn;
      row[cfg.col0[a] + L + l] = cs;
      fr *= 2.0f;
    }
  }
  float* drow = row + cfg.pos_width;
  for (int a = 0; a < 3; ++a) {
    float fr = 1.0f;
    for (int l = 0; l < cfg.dir_freqs; ++l) {
      float sn, cs;
      sin_cos((fr * PI_F) * d[a], sn, cs);
      drow[a * 2 * cfg.dir_freqs + 2 * l] = sn;
      drow[a * 2 * cfg.dir_freqs + 2 * l + 1] = cs;
      fr *= 2.0f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, one sample: dL/d(row) -> dL/dz
// ------------------------------------------------------------------------------------------------
// Geodetic Jacobian. The reference differentiates wgs_84.py:56-97 with autograd; its three outputs are
// (one Bowring step away from) latitude, longitude and ellipsoidal height, whose derivatives w.r.t. the
// ECEF position are the rows of the local north / east / up frame:
//   d lat = north . dx / (M + h)      d lon = east . dx / ((N + h) cos lat)      d h = up . dx
// with M = a (1 - e^2) / W^3, N = a / W, W = sqrt(1 - e^2 sin^2 lat). Agreement with autograd through the
// reference's expressions: 2e-7 relative (the Bowring step's own error), tests/test_abi_and_host.py.
ATM_HD float nerf_encode_sample_bwd(const atmonr_frame_t& f, const GeoFrame& gf, const float* o, const float* d,
                                    float zz, const float* pn, const float* grow, const NerfEncCfg& cfg) {
  const float PI_F = 3.14159265358979323846f;
  float gp[3];
  for (int a = 0; a < 3; ++a) {
    const int L = cfg.freqs[a];
    const float p = pn[a];
    float fr = 1.0f, acc = 0.0f;
    for (int l = 0; l < L; ++l) {
      const float c = fr * PI_F;
      float sn, cs;
      sin_cos(c * p, sn, cs);
      // d sin(c p) = c cos(c p), d cos(c p) = -c sin(c p)
      acc += (grow[cfg.col0[a] + l] * cs - grow[cfg.col0[a] + L + l] * sn) * c;
      fr *= 2.0f;
    }
    gp[a] = acc;
  }
  if (f.enabled) {
    const float s = (float)f.scale;
    const float px = o[0] + d[0] * zz, py = o[1] + d[1] * zz, pz = o[2] + d[2] * zz;
    const double X = (double)(px * s) + f.offset[0], Y = (double)(py * s) + f.offset[1],
                 Z = (double)(pz * s) + f.offset[2];
    double lat_deg, lon_deg, h;
    ecef_to_geodetic_local(gf, X, Y, Z, lat_deg, lon_deg, h);
    // the clip of harp2.py:386 passes the gradient where the float32 value lies inside [-1, 1]
    double lon_s = lon_deg;
    if (f.shift_lon) lon_s = (lon_s < 0.0 ? lon_s + 360.0 : lon_s) - 180.0;
    const float v0 = (float)((lat_deg - f.lat_min) * gf.lat_k - 1.0), v1 = (float)((lon_s - f.lon_min) * gf.lon_k - 1.0),
                v2 = (float)(h * gf.alt_k - 1.0);
    const double R2D = 180.0 / 3.141592653589793;
    const double A = ATM_WGS_A, B = ATM_WGS_B;
    const double E_SQ = (A * A - B * B) / (A * A);
    const double phi = lat_deg / R2D, lam = lon_deg / R2D;
    const double sp = sin(phi), cp = cos(phi), sl = sin(lam), cl = cos(lam);
    const double w2 = 1.0 - E_SQ * sp * sp, w = sqrt(w2);
    const double Nn = A / w, Mm = A * (1.0 - E_SQ) / (w2 * w);
    const double g_lat = (v0 >= -1.0f && v0 <= 1.0f) ? (double)gp[0] * gf.lat_k * R2D / (Mm + h) : 0.0;
    const double g_lon = (v1 >= -1.0f && v1 <= 1.0f) ? (double)gp[1] * gf.lon_k * R2D / ((Nn + h) * cp) : 0.0;
    const double g_alt = (v2 >= -1.0f && v2 <= 1.0f) ? (double)gp[2] * gf.alt_k : 0.0;
    const double gx = g_lat * (-sp * cl) + g_lon * (-sl) + g_alt * (cp * cl);
    const double gy = g_lat * (-sp * sl) + g_lon * cl + g_alt * (cp * sl);
    const double gzz = g_lat * cp + g_alt * sp;
    gp[0] = (float)gx * s, gp[1] = (float)gy * s, gp[2] = (float)gzz * s;
  }
  return (gp[0] * d[0] + gp[1] * d[1]) + gp[2] * d[2];
}

}  // namespace atm
