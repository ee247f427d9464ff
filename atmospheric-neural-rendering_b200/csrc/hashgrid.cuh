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

// ---- fast path -------------------------------------------------------------------------------
// All 2^D corner entries and weights of one level with the per-corner work reduced to adds/xors:
// the per-dimension terms (g*stride for dense levels, g*prime for hashed levels) are formed once
// for g and g+1. Bit-identical to grid_corner()/grid_entry():
//   dense level  : index = sum_k g_k * res^k  (< 2*size unless x is outside [0,1]), then % size
//   hashed level : index = xor_k g_k * prime_k, size is a power of two -> mask
// Weights keep the reference association ((1*w0)*w1)*w2.
template <int D>
__device__ __forceinline__ void level_corners(const atmonr_grid_t& g, int l, const float (&x)[D],
                                              uint32_t (&e)[1 << D], float (&w)[1 << D]) {
  const float scale = g.scale[l];
  const uint32_t res = g.res[l], size = g.size[l];
  uint32_t cell[D];
  float frac[D];
  grid_cell<D>(x, scale, cell, frac);
  // does the level hash? (grid_entry: stride after the last included dimension exceeds size)
  uint64_t dense = 1;
#pragma unroll
  for (int k = 0; k < D; ++k) dense *= res;
  const bool hashed = dense > (uint64_t)size;
  uint32_t t0[D], t1[D];  // per-dimension index terms for g and g+1
  if (hashed) {
    const uint32_t primes[4] = {1u, 2654435761u, 805459861u, 3674653429u};
#pragma unroll
    for (int k = 0; k < D; ++k) {
      t0[k] = cell[k] * primes[k];
      t1[k] = t0[k] + primes[k];
    }
  } else {
    uint32_t stride = 1;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      t0[k] = cell[k] * stride;
      t1[k] = t0[k] + stride;
      stride *= res;
    }
  }
  const bool pow2 = (size & (size - 1u)) == 0u;
#pragma unroll
  for (int c = 0; c < (1 << D); ++c) {
    uint32_t idx = 0;
    float wc = 1.0f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const bool up = (c >> k) & 1;
      const uint32_t t = up ? t1[k] : t0[k];
      idx = hashed ? (idx ^ t) : (idx + t);
      wc *= up ? frac[k] : 1.0f - frac[k];
    }
    if (hashed && pow2) {
      idx &= size - 1u;
    } else if (idx >= size) {
      idx = (idx - size < size) ? idx - size : idx % size;
    }
    e[c] = idx;
    w[c] = wc;
  }
}

// Encode one point with level_corners(); same result as hash_encode().
template <int D>
__device__ __forceinline__ void hash_encode_fast(const atmonr_grid_t& g, const __half2* __restrict__ table,
                                                 const float (&x)[D], __half2 (&out)[ATMONR_MAX_LEVELS]) {
#pragma unroll
  for (int l = 0; l < ATMONR_MAX_LEVELS; ++l) {
    if (l < g.n_levels) {
      uint32_t e[1 << D];
      float w[1 << D];
      level_corners<D>(g, l, x, e, w);
      const __half2* lvl = table + g.offset[l];
      __half2 v[1 << D];
#pragma unroll
      for (int c = 0; c < (1 << D); ++c) v[c] = __ldg(lvl + e[c]);
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

template <int D>
__device__ __forceinline__ void hash_scatter_fast(const atmonr_grid_t& g, float* __restrict__ dtable,
                                                  const float (&x)[D], const float* denc, float scale) {
#pragma unroll
  for (int l = 0; l < ATMONR_MAX_LEVELS; ++l) {
    if (l < g.n_levels) {
      const float d0 = denc[2 * l] * scale, d1 = denc[2 * l + 1] * scale;
      if (d0 != 0.0f || d1 != 0.0f) {
        uint32_t e[1 << D];
        float w[1 << D];
        level_corners<D>(g, l, x, e, w);
        float* lvl = dtable + 2 * (size_t)g.offset[l];
#pragma unroll
        for (int c = 0; c < (1 << D); ++c) red_add_f32x2(lvl + 2 * (size_t)e[c], w[c] * d0, w[c] * d1);
      }
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
