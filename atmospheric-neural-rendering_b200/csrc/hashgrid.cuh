// hashgrid.cuh -- device side of the multiresolution hash encoding (F = 2 features/entry).
//
// Table layout in HBM/L2: one contiguous array of entries, level after level (offset[] from
// atmonr_grid_layout), each entry = 2 features. The forward reads the fp16 shadow (__half2 =
// one 4-byte gather per corner, 84.6 MB for the 3-D grid -> L2 resident on B200); the
// backward adds into an fp32 gradient array with one vector RED per corner.
#pragma once

#include "common.cuh"

namespace atm {

// Encode one point: out[2*l], out[2*l+1] = fp16-rounded interpolated features of level l.
template <int D>
__device__ __forceinline__ void hash_encode(const atmonr_grid_t& g, const __half2* __restrict__ table,
                                            const float (&x)[D], __half2 (&out)[ATMONR_MAX_LEVELS]) {
#pragma unroll
  for (int l = 0; l < ATMONR_MAX_LEVELS; ++l) {
    if (l < g.n_levels) {
      uint32_t cell[D];
      float frac[D];
      grid_cell<D>(x, g.scale[l], cell, frac);
      const __half2* lvl = table + g.offset[l];
      __half2 v[1 << D];
      float w[1 << D];
#pragma unroll
      for (int c = 0; c < (1 << D); ++c) {
        uint32_t e;
        grid_corner<D>(cell, frac, c, g.res[l], g.size[l], e, w[c]);
        v[c] = __ldg(lvl + e);
      }
      float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
      for (int c = 0; c < (1 << D); ++c) {
        const float2 f = __half22float2(v[c]);
        a0 = fmaf(w[c], f.x, a0);
        a1 = fmaf(w[c], f.y, a1);
      }
      out[l] = __floats2half2_rn(a0, a1);
    } else {
      out[l] = __floats2half2_rn(0.0f, 0.0f);
    }
  }
}

// Scatter dL/d(features) of one point into the fp32 gradient table.
template <int D>
__device__ __forceinline__ void hash_scatter(const atmonr_grid_t& g, float* __restrict__ dtable,
                                             const float (&x)[D], const float* denc, float scale) {
#pragma unroll
  for (int l = 0; l < ATMONR_MAX_LEVELS; ++l) {
    if (l < g.n_levels) {
      const float d0 = denc[2 * l] * scale, d1 = denc[2 * l + 1] * scale;
      if (d0 != 0.0f || d1 != 0.0f) {
        uint32_t cell[D];
        float frac[D];
        grid_cell<D>(x, g.scale[l], cell, frac);
        float* lvl = dtable + 2 * (size_t)g.offset[l];
#pragma unroll
        for (int c = 0; c < (1 << D); ++c) {
          uint32_t e;
          float w;
          grid_corner<D>(cell, frac, c, g.res[l], g.size[l], e, w);
          red_add_f32x2(lvl + 2 * (size_t)e, w * d0, w * d1);
        }
      }
    }
  }
}

}  // namespace atm
