// hashgrid.cuh -- device side of the multiresolution hash encoding (F = 2 features/entry).
//
// Table layout in HBM/L2: one contiguous array of entries, level after level (offset[] from
// atmonr_grid_layout), each entry = 2 features. The forward reads the fp16 shadow (__half2 =
// one 4-byte gather per corner, 84.6 MB for the 3-D grid -> L2 resident on B200); the
// backward adds into an fp32 gradient array with one vector RED per corner.
#pragma once

#include "common.cuh"

namespace atm {

// N-linear interpolation of one level, tiny-cuda-nn's arithmetic (grid.h kernel_grid,
// upstream-recalled; SURVEY 8c "tcnn accumulates in fp16"): the fp32 corner weight is rounded to
// fp16 and the two features of the entry are accumulated with ONE packed half-precision FMA per
// corner, corners in index order (bit k of the corner id = +1 in dimension k), starting from 0.
// (Also the cheapest form: 2 instructions per corner instead of 2 converts + 2 FFMA.)
template <int NC>
__device__ __forceinline__ __half2 interp_corners(const __half2 (&v)[NC], const float (&w)[NC]) {
  __half2 acc = __floats2half2_rn(0.0f, 0.0f);
#pragma unroll
  for (int c = 0; c < NC; ++c) acc = __hfma2(__float2half2_rn(w[c]), v[c], acc);
  return acc;
}

// Encode one point: out[l] = the two interpolated fp16 features of level l.
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
      out[l] = interp_corners<(1 << D)>(v, w);
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
      out[l] = interp_corners<(1 << D)>(v, w);
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

// ---- rolled-loop variants (small code footprint) ------------------------------------------------
// The fully unrolled 16-level loops above cost 100+ KB of SASS per kernel and make the fused
// kernels instruction-fetch bound. These variants keep the level loop rolled: per-level constants
// come from a shared-memory copy of the level table (dynamic index), features are written straight
// into the activation tile / read straight from TMEM instead of living in a register array.
// rare path of the index reduction (coordinates outside [0,1]); kept out of line on purpose
static __device__ __noinline__ uint32_t index_mod_slow(uint32_t idx, uint32_t size) { return idx % size; }

struct LevelRow {
  float scale;
  uint32_t res, size, offset, hashed;  // hashed: bit0 = level hashes, bit1 = size is a power of two
  uint32_t stride1, stride2, pad;      // res, res*res (dense index strides)
};

__device__ __forceinline__ void load_level_table(const atmonr_grid_t& g, LevelRow* rows) {
  if (threadIdx.x < ATMONR_MAX_LEVELS) {
    const int l = threadIdx.x;
    LevelRow r;
    r.scale = g.scale[l];
    r.res = g.res[l];
    r.size = g.size[l];
    r.offset = g.offset[l];
    uint64_t dense = 1;
    for (int k = 0; k < g.n_dims; ++k) dense *= r.res;
    r.hashed = (dense > (uint64_t)r.size ? 1u : 0u) | (((r.size & (r.size - 1u)) == 0u) ? 2u : 0u);
    r.stride1 = r.res;
    r.stride2 = r.res * r.res;
    r.pad = 0;
    rows[l] = r;
  }
}

// the two halves of level_corners3, for callers that need the entries only once per cell run
__device__ __forceinline__ void corner_weights3(const float (&frac)[3], float (&w)[8]) {
  const float wx[2] = {1.0f - frac[0], frac[0]};
  const float wy[2] = {1.0f - frac[1], frac[1]};
  const float wz[2] = {1.0f - frac[2], frac[2]};
  const float wxy[4] = {wx[0] * wy[0], wx[1] * wy[0], wx[0] * wy[1], wx[1] * wy[1]};
#pragma unroll
  for (int c = 0; c < 8; ++c) w[c] = wxy[c & 3] * wz[c >> 2];
}

// Entries of the 8 corners of `cell`, bit-identical to grid_entry() of every corner. Two tight
// paths cover a regular table: hashed levels with a power-of-two size (three-input XORs and one
// mask per corner) and dense levels whose 8 corners lie inside the level (one three-input add per
// corner: the corner with all +1 is the largest, so one compare proves that no modulo is needed).
// Anything else (coordinates outside [0,1], wrap at the upper boundary, odd sizes) takes the
// general path.
__device__ __forceinline__ uint32_t xor_and(uint32_t a, uint32_t b, uint32_t mask) {  // (a ^ b) & mask, one LOP3
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(r) : "r"(a), "r"(b), "r"(mask));
  return r;
}
__device__ __forceinline__ uint32_t xor2(uint32_t a, uint32_t b) {  // kept opaque so the pair terms are shared
  uint32_t r;
  asm("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void corner_entries3(const LevelRow& lv, const uint32_t (&cell)[3], uint32_t (&e)[8]) {
  const uint32_t x[2] = {cell[0], cell[0] + 1u};
  if (lv.hashed == 3u) {  // hashed, power-of-two size
    const uint32_t y0 = cell[1] * 2654435761u, y1 = y0 + 2654435761u;
    const uint32_t z0 = cell[2] * 805459861u, z1 = z0 + 805459861u;
    const uint32_t yz[4] = {xor2(y0, z0), xor2(y1, z0), xor2(y0, z1), xor2(y1, z1)};
    const uint32_t mask = lv.size - 1u;
#pragma unroll
    for (int c = 0; c < 8; ++c) e[c] = xor_and(x[c & 1], yz[c >> 1], mask);
  } else if (lv.hashed & 1u) {  // hashed, any size (not produced by atmonr_grid_layout)
    const uint32_t y0 = cell[1] * 2654435761u, z0 = cell[2] * 805459861u;
    const uint32_t y[2] = {y0, y0 + 2654435761u}, z[2] = {z0, z0 + 805459861u};
#pragma unroll
    for (int c = 0; c < 8; ++c) e[c] = index_mod_slow(x[c & 1] ^ y[(c >> 1) & 1] ^ z[c >> 2], lv.size);
  } else {  // dense
    const uint32_t y0 = cell[1] * lv.stride1, y1 = y0 + lv.stride1;
    const uint32_t z0 = cell[2] * lv.stride2, z1 = z0 + lv.stride2;
    const uint32_t yz[4] = {y0 + z0, y1 + z0, y0 + z1, y1 + z1};
#pragma unroll
    for (int c = 0; c < 8; ++c) e[c] = x[c & 1] + yz[c >> 1];
    // cells below 2^16 cannot wrap uint32 (res^3 of a dense level fits 32 bits), so e[7] is the largest
    if ((cell[0] | cell[1] | cell[2]) >= 65536u || e[7] >= lv.size) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (e[c] >= lv.size) e[c] = (e[c] - lv.size < lv.size) ? e[c] - lv.size : index_mod_slow(e[c], lv.size);
    }
  }
}

// L2 eviction policies: the table (and its gradient) should stay L2-resident while tens of GB of
// per-sample data stream through the same cache every step.
__device__ __forceinline__ uint64_t l2_policy_keep() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_stream() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint32_t ldg_u32_hint(const uint32_t* ptr, uint64_t policy) {
  uint32_t v;
  asm("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ float ldg_f32_hint(const float* ptr, uint64_t policy) {
  float v;
  asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ uint4 ldg_u128_hint(const uint4* ptr, uint64_t policy) {
  uint4 v;
  asm("ld.global.nc.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
      : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ void stg_u128_hint(uint4* ptr, uint4 v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void stg_f32_hint(float* ptr, float v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(ptr), "f"(v), "l"(policy) : "memory");
}

// &table[entry] with one 32x32+64-bit multiply-add (the compiler otherwise spends four
// instructions per corner on 64-bit address arithmetic)
template <typename T>
__device__ __forceinline__ T* entry_ptr(T* base, uint32_t entry) {
  T* p;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(entry), "n"(sizeof(T)), "l"(base));
  return p;
}

// corner entries + weights of one level from a LevelRow (3-D); bit-identical to level_corners<3>
__device__ __forceinline__ void level_corners3(const LevelRow& lv, const float (&x)[3], uint32_t (&e)[8],
                                               float (&w)[8], uint32_t (&cell)[3]) {
  float frac[3];
  grid_cell<3>(x, lv.scale, cell, frac);
  corner_entries3(lv, cell, e);
  corner_weights3(frac, w);
}
__device__ __forceinline__ void level_corners3(const LevelRow& lv, const float (&x)[3], uint32_t (&e)[8],
                                               float (&w)[8]) {
  uint32_t cell[3];
  level_corners3(lv, x, e, w, cell);
}

// Warp-aggregated scatter of one level. The 32 lanes of a warp hold consecutive samples of a ray,
// so lanes that fall into the same grid cell form contiguous runs; the 8 corner contributions
// (2 features each) of a run are summed with a segmented shuffle reduction and written by the
// run's first lane with one vector RED per corner. When the warp has little duplication the
// plain per-lane REDs are cheaper and are used instead. Must be called by all 32 lanes.
__device__ __forceinline__ void scatter_level_aggregated(const LevelRow& lv, float* __restrict__ dtable,
                                                         const float (&x)[3], float d0, float d1, bool valid) {
  constexpr uint32_t full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  uint32_t e[8], cell[3];
  float w[8];
  level_corners3(lv, x, e, w, cell);
  if (!valid) d0 = 0.0f, d1 = 0.0f;
  const uint32_t p0 = __shfl_up_sync(full, cell[0], 1), p1 = __shfl_up_sync(full, cell[1], 1),
                 p2 = __shfl_up_sync(full, cell[2], 1);
  const bool head = lane == 0 || p0 != cell[0] || p1 != cell[1] || p2 != cell[2];
  const uint32_t heads = __ballot_sync(full, head);
  float* base = dtable + 2 * (size_t)lv.offset;
  if (__popc(heads) > 16) {  // mostly distinct cells: no gain from aggregation
    if (d0 != 0.0f || d1 != 0.0f) {
#pragma unroll
      for (int c = 0; c < 8; ++c) red_add_f32x2(base + 2 * (size_t)e[c], w[c] * d0, w[c] * d1);
    }
    return;
  }
  const uint32_t above = lane == 31 ? 0u : (heads & ~((2u << lane) - 1u));
  const int run_end = above ? (__ffs(above) - 2) : 31;
  float v[16];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[2 * c] = w[c] * d0, v[2 * c + 1] = w[c] * d1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const bool take = lane + o <= run_end;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float t = __shfl_down_sync(full, v[i], o);
      if (take) v[i] += t;
    }
  }
  if (head) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (v[2 * c] != 0.0f || v[2 * c + 1] != 0.0f) red_add_f32x2(base + 2 * (size_t)e[c], v[2 * c], v[2 * c + 1]);
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
