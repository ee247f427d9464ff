// atmonr_b200.cu -- kernels and C ABI of libatmonr_b200.so (sm_100a).
// Interface contract: include/atmonr_b200.h. Reference citations are relative to the
// nasa/atmospheric-neural-rendering tree.
#include "common.cuh"
#include "hashgrid.cuh"
#include "mlp_simt.cuh"

// The translation unit is compiled in four parts (in parallel, see build.py): each part
// instantiates a subset of the kernels. ATM_PART unset = everything.
#ifndef ATM_PART
#define ATM_PART_BASIC 1
#define ATM_PART_MLP 1
#define ATM_PART_FIELD 1
#define ATM_PART_SURF 1
#else
#define ATM_PART_BASIC (ATM_PART == 0)
#define ATM_PART_MLP (ATM_PART == 1)
#define ATM_PART_FIELD (ATM_PART == 2)
#define ATM_PART_SURF (ATM_PART == 3)
#endif

namespace atm {
#if ATM_PART_BASIC
thread_local char g_last_error[512] = "";
#endif

#if ATM_PART_BASIC
// =========================================================================================
// Samplers and the point preprocessor
// =========================================================================================
__device__ __forceinline__ float draw_t(int mode, const float* u, int64_t idx, uint64_t seed,
                                        uint64_t ray, int bin) {
  if (mode == 0) return 0.5f;
  if (mode == 1) return u[idx];
  return philox_uniform(seed, ray, (uint32_t)bin);
}

// samplers.py:8-47
__global__ void k_sample_uniform(const float* __restrict__ o, const float* __restrict__ d,
                                 const float* __restrict__ len, const float* __restrict__ u,
                                 const float* __restrict__ bins, int64_t total, int N, int mode,
                                 uint64_t seed, uint64_t base, float* __restrict__ pts,
                                 float* __restrict__ z) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t ray = idx / N;
  const int i = (int)(idx - ray * N);
  const float t = draw_t(mode, u, idx, seed, base + ray, i);
  const float lo = bins ? bins[i] : (float)i / (float)N;
  const float zz = stratified_z(lo, t, N, len[ray]);
  z[idx] = zz;
#pragma unroll
  for (int k = 0; k < 3; ++k) pts[idx * 3 + k] = o[ray * 3 + k] + d[ray * 3 + k] * zz;
}

// harp2.py:372-386
__global__ void k_preprocess_f32(atmonr_frame_t f, GeoFrame gf, const float* __restrict__ p,
                                 float* __restrict__ out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a, b, c;
  preprocess_f32(f, gf, p[3 * i], p[3 * i + 1], p[3 * i + 2], a, b, c);
  out[3 * i] = a;
  out[3 * i + 1] = b;
  out[3 * i + 2] = c;
}
__global__ void k_preprocess_f64(atmonr_frame_t f, GeoFrame gf, const double* __restrict__ p,
                                 double* __restrict__ out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a, b, c;
  preprocess_f64(f, gf, p[3 * i], p[3 * i + 1], p[3 * i + 2], a, b, c);
  out[3 * i] = a;
  out[3 * i + 1] = b;
  out[3 * i + 2] = c;
}

// instant_ngp.py:139-160 in one pass: stratified sample -> ECEF -> geodetic -> [0,1]^3 with
// compressed altitude. One thread per sample; consecutive threads walk along a ray.
// (General form: any N, any pointer alignment. The training path uses k_ngp_sample_points4.)
// HEIGHT (`include_height`, instant_ngp.py:155-156 -> samplers.py:168-195): a fourth coordinate, the
// ellipsoidal height of (x0, x1, x2) * scale + offset over ray_origin_height, taken BEFORE the altitude
// compression of the third one (instant_ngp.py:160 comes after append_heights); x01 then has 4 columns.
struct HeightArgs {
  double scale, ox, oy, oz, origin_height;
};
template <bool HEIGHT>
__global__ void k_ngp_sample_points(atmonr_frame_t f, GeoFrame gf, const float* __restrict__ o,
                                    const float* __restrict__ d, const float* __restrict__ len,
                                    const float* __restrict__ u, const float* __restrict__ bins,
                                    int64_t total, int N, int mode, uint64_t seed, uint64_t base,
                                    float alt_compress, HeightArgs ha, float* __restrict__ x01, float* __restrict__ z) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t ray = idx / N;
  const int i = (int)(idx - ray * N);
  const float t = draw_t(mode, u, idx, seed, base + ray, i);
  const float lo = bins ? bins[i] : (float)i / (float)N;
  const float zz = stratified_z(lo, t, N, len[ray]);
  z[idx] = zz;
  float p[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) p[k] = o[ray * 3 + k] + d[ray * 3 + k] * zz;
  float c0 = p[0], c1 = p[1], c2 = p[2];
  if (f.enabled) preprocess_f32(f, gf, p[0], p[1], p[2], c0, c1, c2);
  float x0, x1, x2;
  to_unit_cube(c0, c1, c2, alt_compress, x0, x1, x2);
  if (HEIGHT) {
    const float x2u = (c2 + 1.0f) / 2.0f;   // uncompressed
    double lat, lon, alt;
    ecef_to_geodetic((double)x0 * ha.scale + ha.ox, (double)x1 * ha.scale + ha.oy, (double)x2u * ha.scale + ha.oz, lat,
                     lon, alt);
    reinterpret_cast<float4*>(x01)[idx] = make_float4(x0, x1, x2, (float)(alt / ha.origin_height));
  } else {
    x01[idx * 3] = x0;
    x01[idx * 3 + 1] = x1;
    x01[idx * 3 + 2] = x2;
  }
}

// Same arithmetic, one thread per aligned group of FOUR consecutive bins of a ray (N % 4 == 0,
// 16-byte aligned buffers): one Philox block, one index division and one load of the ray per
// four samples, 16-byte loads/stores, and four independent FP64 chains per thread to hide the
// latency of the FP64 pipe. The kernel is bound by FP64 issue, not by its 16 B/sample of HBM writes.
__global__ void __launch_bounds__(128) k_ngp_sample_points4(SamplerJob job) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q < job.groups) sample_group4(job, q);
}

// =========================================================================================
// Hash grid (modular operators)
// =========================================================================================
template <int D>
__global__ void k_hashgrid_fwd(atmonr_grid_t g, const float* __restrict__ x, int xs,
                               const __half2* __restrict__ table, int64_t M, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  float p[D];
#pragma unroll
  for (int k = 0; k < D; ++k) p[k] = x[i * xs + k];
  __half2 enc[ATMONR_MAX_LEVELS];
  hash_encode<D>(g, table, p, enc);
  float* row = out + i * 2 * g.n_levels;
#pragma unroll
  for (int l = 0; l < ATMONR_MAX_LEVELS; ++l)
    if (l < g.n_levels) {
      const float2 v = __half22float2(enc[l]);
      row[2 * l] = v.x;
      row[2 * l + 1] = v.y;
    }
}

template <int D>
__global__ void k_hashgrid_bwd(atmonr_grid_t g, const float* __restrict__ x, int xs,
                               const float* __restrict__ dout, int64_t M, float* __restrict__ dtable) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  float p[D];
#pragma unroll
  for (int k = 0; k < D; ++k) p[k] = x[i * xs + k];
  float d[2 * ATMONR_MAX_LEVELS];
#pragma unroll
  for (int l = 0; l < 2 * ATMONR_MAX_LEVELS; ++l) d[l] = l < 2 * g.n_levels ? dout[i * 2 * g.n_levels + l] : 0.0f;
  hash_scatter<D>(g, dtable, p, d, 1.0f);
}

template <int D>
__global__ void k_hashgrid_indices(atmonr_grid_t g, const float* __restrict__ x, int xs, int64_t M,
                                   uint32_t* __restrict__ idx) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  float p[D];
#pragma unroll
  for (int k = 0; k < D; ++k) p[k] = x[i * xs + k];
  for (int l = 0; l < g.n_levels; ++l) {
    uint32_t cell[D];
    float frac[D];
    grid_cell<D>(p, g.scale[l], cell, frac);
#pragma unroll
    for (int c = 0; c < (1 << D); ++c) {
      uint32_t e;
      float w;
      grid_corner<D>(cell, frac, c, g.res[l], g.size[l], e, w);
      idx[(i * g.n_levels + l) * (1 << D) + c] = g.offset[l] + e;
    }
  }
}

#endif  // ATM_PART_BASIC
#if ATM_PART_MLP
// =========================================================================================
// MLP (modular operators)
// =========================================================================================
template <int IN>
__device__ __forceinline__ void load_padded_input(const float* __restrict__ x, int n_in, bool valid,
                                                  __half2 (&xh)[IN / 2]) {
#pragma unroll
  for (int i = 0; i < IN / 2; ++i) {
    const float a = !valid ? 0.0f : (2 * i < n_in ? x[2 * i] : 1.0f);
    const float b = !valid ? 0.0f : (2 * i + 1 < n_in ? x[2 * i + 1] : 1.0f);
    xh[i] = __floats2half2_rn(a, b);
  }
}

template <int IN, int NH>
__global__ void __launch_bounds__(kTile) k_mlp_fwd(const __half* __restrict__ w, const float* __restrict__ x,
                                                   int n_in, int n_out, int64_t M, float* __restrict__ out) {
  using S = MlpShape<IN, NH>;
  extern __shared__ __align__(16) float smem[];
  load_weights(w, smem, S::kNumWeights);
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile * kTile < M; tile += gridDim.x) {
    const int64_t i = tile * kTile + threadIdx.x;
    if (i >= M) continue;
    __half2 xh[IN / 2], h[NH][kWidth / 2];
    load_padded_input<IN>(x + i * n_in, n_in, true, xh);
    float y[kOutPad];
    mlp_forward<IN, NH, kOutPad>(smem, xh, h, y);
#pragma unroll
    for (int o = 0; o < kOutPad; ++o)
      if (o < n_out) out[i * n_out + o] = y[o];
  }
}

template <int IN, int NH>
__global__ void __launch_bounds__(kTile) k_mlp_bwd(const __half* __restrict__ w, const float* __restrict__ x,
                                                   const float* __restrict__ dout, int n_in, int n_out,
                                                   int64_t M, float* __restrict__ dx, float* __restrict__ dw) {
  using S = MlpShape<IN, NH>;
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;
  float* sdW = sW + S::kNumWeights;
  float* scratch = sdW + S::kNumWeights;
  load_weights(w, sW, S::kNumWeights);
  for (int i = threadIdx.x; i < S::kNumWeights; i += blockDim.x) sdW[i] = 0.0f;
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile * kTile < M; tile += gridDim.x) {
    const int64_t i = tile * kTile + threadIdx.x;
    const bool valid = i < M;
    __half2 xh[IN / 2], h[NH][kWidth / 2];
    load_padded_input<IN>(x + (valid ? i : 0) * n_in, n_in, valid, xh);
    float y[kOutPad], dy[kOutPad], dxi[IN];
    mlp_forward<IN, NH, kOutPad>(sW, xh, h, y);
#pragma unroll
    for (int o = 0; o < kOutPad; ++o) dy[o] = (valid && o < n_out) ? dout[i * n_out + o] : 0.0f;
    mlp_backward<IN, NH>(sW, sdW, scratch, xh, h, dy, dxi);
    if (valid && dx) {
#pragma unroll
      for (int k = 0; k < IN; ++k)
        if (k < n_in) dx[i * n_in + k] = dxi[k];
    }
  }
  flush_dw(sdW, dw, S::kNumWeights);
}

#endif  // ATM_PART_MLP
// =========================================================================================
// Fused radiance field: hash grid -> pos_mlp -> [SH2(dir) | features | 1-pad] -> dir_mlp
// =========================================================================================
using PosMlp = MlpShape<32, 1>;
using DirMlp = MlpShape<32, 2>;
constexpr int kNumBands = 4;

#if ATM_PART_FIELD
// Assemble the padded dir_mlp input from the ray direction and the pos_mlp output
// (instant_ngp.py:165-169; tcnn Composite{SH2, Identity} then 1.0 padding to a multiple of 16).
// V = number of densities (`multi_band_extinction`: V = 4): the first V outputs of pos_mlp are densities,
// the other 16 - V are features, so the input is 4 + 16 - V wide: 19 -> 32 (V = 1), 16 -> 16 (V = 4).
template <int V>
struct DirIn {
  static constexpr int kWidthIn = V == 1 ? 32 : 16;
};
template <int V>
__device__ __forceinline__ void build_dir_input(const float* __restrict__ dir, const float (&po)[kOutPad],
                                                __half2 (&din)[DirIn<V>::kWidthIn / 2]) {
  constexpr int W = DirIn<V>::kWidthIn;
  float v[W];
  float sh[4];
  sh_degree2(dir[0], dir[1], dir[2], sh);
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = sh[k];
#pragma unroll
  for (int k = V; k < kOutPad; ++k) v[4 + k - V] = po[k];
#pragma unroll
  for (int k = 20 - V; k < W; ++k) v[k] = 1.0f;
#pragma unroll
  for (int k = 0; k < W / 2; ++k) din[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
}

// D = grid dimensionality (4 with `include_height`), V = densities per sample (4 with
// `multi_band_extinction`); the shipped configuration is <3, 1>.
template <int D, int V>
__global__ void __launch_bounds__(kTile)
k_field_fwd(atmonr_grid_t g, const __half2* __restrict__ table, const __half* __restrict__ pos_w,
            const __half* __restrict__ dir_w, const float* __restrict__ x01, const float* __restrict__ dirs,
            int64_t M, int N, float* __restrict__ sigma_raw, float* __restrict__ color_raw) {
  constexpr int DIN = DirIn<V>::kWidthIn;
  using DirM = MlpShape<DIN, 2>;
  __shared__ __align__(16) float sW[PosMlp::kNumWeights + DirM::kNumWeights];
  float* sWp = sW;
  float* sWd = sW + PosMlp::kNumWeights;
  load_weights(pos_w, sWp, PosMlp::kNumWeights);
  load_weights(dir_w, sWd, DirM::kNumWeights);
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile * kTile < M; tile += gridDim.x) {
    const int64_t i = tile * kTile + threadIdx.x;
    if (i >= M) continue;
    float p[D];
#pragma unroll
    for (int k = 0; k < D; ++k) p[k] = x01[D * i + k];
    __half2 enc[16], hp[1][16], din[DIN / 2], hd[2][16];
    hash_encode<D>(g, table, p, enc);
    float po[kOutPad];
    mlp_forward<32, 1, kOutPad>(sWp, enc, hp, po);
#pragma unroll
    for (int v = 0; v < V; ++v) sigma_raw[V * i + v] = po[v];
    build_dir_input<V>(dirs + (i / N) * 3, po, din);
    float c[kNumBands];
    mlp_forward<DIN, 2, kNumBands>(sWd, din, hd, c);
    *reinterpret_cast<float4*>(color_raw + 4 * i) = make_float4(c[0], c[1], c[2], c[3]);
  }
}

template <int D, int V>
__global__ void __launch_bounds__(kTile)
k_field_bwd(atmonr_grid_t g, const __half2* __restrict__ table, const __half* __restrict__ pos_w,
            const __half* __restrict__ dir_w, const float* __restrict__ x01, const float* __restrict__ dirs,
            const float* __restrict__ dsigma_raw, const float* __restrict__ dcolor_raw, int64_t M, int N,
            float* __restrict__ dtable, float* __restrict__ dpos_w, float* __restrict__ ddir_w) {
  constexpr int DIN = DirIn<V>::kWidthIn;
  using DirM = MlpShape<DIN, 2>;
  extern __shared__ __align__(16) float smem[];
  float* sWp = smem;
  float* sWd = sWp + PosMlp::kNumWeights;
  float* sdWp = sWd + DirM::kNumWeights;
  float* sdWd = sdWp + PosMlp::kNumWeights;
  float* scratch = sdWd + DirM::kNumWeights;
  load_weights(pos_w, sWp, PosMlp::kNumWeights);
  load_weights(dir_w, sWd, DirM::kNumWeights);
  for (int i = threadIdx.x; i < PosMlp::kNumWeights + DirM::kNumWeights; i += blockDim.x) sdWp[i] = 0.0f;
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile * kTile < M; tile += gridDim.x) {
    const int64_t i = tile * kTile + threadIdx.x;
    const bool valid = i < M;
    const int64_t j = valid ? i : 0;
    float p[D];
#pragma unroll
    for (int k = 0; k < D; ++k) p[k] = x01[D * j + k];
    __half2 enc[16], hp[1][16], din[DIN / 2], hd[2][16];
    hash_encode<D>(g, table, p, enc);
    float po[kOutPad];
    mlp_forward<32, 1, kOutPad>(sWp, enc, hp, po);
    build_dir_input<V>(dirs + (j / N) * 3, po, din);
    float c[kNumBands];
    mlp_forward<DIN, 2, kNumBands>(sWd, din, hd, c);

    float dout[kOutPad];
#pragma unroll
    for (int k = 0; k < kOutPad; ++k) dout[k] = 0.0f;
    if (valid) {
      const float4 dc = *reinterpret_cast<const float4*>(dcolor_raw + 4 * i);
      dout[0] = dc.x; dout[1] = dc.y; dout[2] = dc.z; dout[3] = dc.w;
    }
    float ddin[DIN];
    mlp_backward<DIN, 2>(sWd, sdWd, scratch, din, hd, dout, ddin);
    float dpo[kOutPad];
#pragma unroll
    for (int v = 0; v < V; ++v) dpo[v] = valid ? dsigma_raw[V * i + v] : 0.0f;
#pragma unroll
    for (int k = V; k < kOutPad; ++k) dpo[k] = ddin[4 + k - V];
    float denc[32];
    mlp_backward<32, 1>(sWp, sdWp, scratch, enc, hp, dpo, denc);
    if (valid) hash_scatter<D>(g, dtable, p, denc, 1.0f);
  }
  flush_dw(sdWp, dpos_w, PosMlp::kNumWeights);
  flush_dw(sdWd, ddir_w, DirM::kNumWeights);
}

#endif  // ATM_PART_FIELD
// =========================================================================================
// Surface branch (one row per ray): [hash2d(end point xy) | SH2(dir)] -> surf_mlp (48 wide in)
// =========================================================================================
using SurfMlp = MlpShape<48, 2>;
#if ATM_PART_SURF

__device__ __forceinline__ void surface_input(const atmonr_grid_t& g, const __half2* __restrict__ table,
                                              const float* __restrict__ o, const float* __restrict__ d,
                                              float len, float (&xy)[2], __half2 (&x)[24]) {
  // instant_ngp.py:140,150: end point of the ray in normalised Cartesian, mapped to [0,1]
#pragma unroll
  for (int k = 0; k < 2; ++k) xy[k] = ((o[k] + d[k] * len) + 1.0f) / 2.0f;
  __half2 enc[ATMONR_MAX_LEVELS];
  hash_encode<2>(g, table, xy, enc);
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = enc[k];
  float sh[4];
  sh_degree2(d[0], d[1], d[2], sh);
  x[16] = __floats2half2_rn(sh[0], sh[1]);
  x[17] = __floats2half2_rn(sh[2], sh[3]);
#pragma unroll
  for (int k = 18; k < 24; ++k) x[k] = __floats2half2_rn(1.0f, 1.0f);
}

__global__ void __launch_bounds__(kTile)
k_surface_fwd(atmonr_grid_t g, const __half2* __restrict__ table, const __half* __restrict__ w,
              const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ len,
              int64_t B, float* __restrict__ color_surf_raw) {
  extern __shared__ __align__(16) float smem[];
  load_weights(w, smem, SurfMlp::kNumWeights);
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile * kTile < B; tile += gridDim.x) {
    const int64_t i = tile * kTile + threadIdx.x;
    if (i >= B) continue;
    float xy[2];
    __half2 x[24], h[2][16];
    surface_input(g, table, o + 3 * i, d + 3 * i, len[i], xy, x);
    float c[kNumBands];
    mlp_forward<48, 2, kNumBands>(smem, x, h, c);
    *reinterpret_cast<float4*>(color_surf_raw + 4 * i) = make_float4(c[0], c[1], c[2], c[3]);
  }
}

__global__ void __launch_bounds__(kTile)
k_surface_bwd(atmonr_grid_t g, const __half2* __restrict__ table, const __half* __restrict__ w,
              const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ len,
              const float* __restrict__ dcs, int64_t B, float* __restrict__ dtable, float* __restrict__ dw) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;
  float* sdW = sW + SurfMlp::kNumWeights;
  float* scratch = sdW + SurfMlp::kNumWeights;
  load_weights(w, sW, SurfMlp::kNumWeights);
  for (int i = threadIdx.x; i < SurfMlp::kNumWeights; i += blockDim.x) sdW[i] = 0.0f;
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile * kTile < B; tile += gridDim.x) {
    const int64_t i = tile * kTile + threadIdx.x;
    const bool valid = i < B;
    const int64_t j = valid ? i : 0;
    float xy[2];
    __half2 x[24], h[2][16];
    surface_input(g, table, o + 3 * j, d + 3 * j, len[j], xy, x);
    float c[kNumBands];
    mlp_forward<48, 2, kNumBands>(sW, x, h, c);
    float dout[kOutPad];
#pragma unroll
    for (int k = 0; k < kOutPad; ++k) dout[k] = (valid && k < kNumBands) ? dcs[4 * i + k] : 0.0f;
    float dx[48];
    mlp_backward<48, 2>(sW, sdW, scratch, x, h, dout, dx);
    if (valid) hash_scatter<2>(g, dtable, xy, dx, 1.0f);
  }
  flush_dw(sdW, dw, SurfMlp::kNumWeights);
}

#endif  // ATM_PART_SURF
#if ATM_PART_BASIC
// =========================================================================================
// Emission-absorption compositing: one warp per ray (graphics_utils.py:6-77)
// =========================================================================================
struct RaySample {
  float delta;          // Voronoi cell width (km)
  float sig[4];         // post-ReLU density per V
  float col[4];         // post-ReLU colour per K
};

__device__ __forceinline__ float voronoi_delta(const float* __restrict__ zr, int i, int N, float zs) {
  const float zi = zr[i] * zs;
  const float lo = i == 0 ? zi * 0.0f : (zr[i - 1] * zs + zi) / 2.0f;
  const float hi = i == N - 1 ? zi : (zi + zr[i + 1] * zs) / 2.0f;
  return hi - lo;
}

// Raw inputs of one sample. The compositing kernels walk a ray 32 samples at a time with a serial
// dependence (the scan carry), so the loads of the following chunks are issued two chunks ahead
// (register prefetch): the kernels are bound by HBM latency x bytes in flight otherwise.
template <int K, int V>
struct RawSample {
  float zm, zi, zp;  // z[i-1], z[i], z[i+1] (normalised)
  float c[K];
  float s[V];
};

template <int K, int V>
__device__ __forceinline__ RawSample<K, V> load_raw(const float* __restrict__ zr, const float* __restrict__ color,
                                                    const float* __restrict__ sigma, int64_t row0, int i, int N) {
  RawSample<K, V> r;
  const bool in = i < N;
  r.zi = in ? zr[i] : 0.0f;
  r.zm = (in && i > 0) ? zr[i - 1] : 0.0f;
  r.zp = (in && i < N - 1) ? zr[i + 1] : 0.0f;
  if (K == 4) {
    const float4 cc = in ? *reinterpret_cast<const float4*>(color + (row0 + i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    r.c[0] = cc.x, r.c[1 % K] = cc.y, r.c[2 % K] = cc.z, r.c[3 % K] = cc.w;
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) r.c[k] = in ? color[(row0 + i) * K + k] : 0.0f;
  }
#pragma unroll
  for (int v = 0; v < V; ++v) r.s[v] = in ? sigma[(row0 + i) * V + v] : 0.0f;
  return r;
}

// same arithmetic as voronoi_delta, from prefetched neighbours
template <int K, int V>
__device__ __forceinline__ float raw_delta(const RawSample<K, V>& r, int i, int N, float zs) {
  const float zi = r.zi * zs;
  const float lo = i == 0 ? zi * 0.0f : (r.zm * zs + zi) / 2.0f;
  const float hi = i == N - 1 ? zi : (zi + r.zp * zs) / 2.0f;
  return hi - lo;
}

template <int K, int V>
__global__ void __launch_bounds__(128)
k_composite_fwd(const float* __restrict__ z, const float* __restrict__ color, const float* __restrict__ sigma,
                const float* __restrict__ color_surf, float zs, int64_t B, int N, int relu,
                float* __restrict__ cmap, float* __restrict__ catmo, float* __restrict__ csurf,
                float* __restrict__ tsurf, float* __restrict__ weights, float* __restrict__ alpha_out) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (ray >= B) return;
  const float* zr = z + ray * N;
  float carry[V], surf[V], acc[K];
#pragma unroll
  for (int v = 0; v < V; ++v) carry[v] = 1.0f, surf[v] = 1.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0f;
  RawSample<K, V> r0 = load_raw<K, V>(zr, color, sigma, ray * N, lane, N);
  RawSample<K, V> r1 = load_raw<K, V>(zr, color, sigma, ray * N, 32 + lane, N);
  for (int base = 0; base < N; base += 32) {
    const int i = base + lane;
    const bool in = i < N;
    const RawSample<K, V> r2 = load_raw<K, V>(zr, color, sigma, ray * N, i + 64, N);
    const float delta = in ? raw_delta(r0, i, N, zs) : 0.0f;
    float c[K];
#pragma unroll
    for (int k = 0; k < K; ++k) c[k] = relu ? fmaxf(r0.c[k], 0.0f) : r0.c[k];
    float sraw[V];
#pragma unroll
    for (int v = 0; v < V; ++v) sraw[v] = r0.s[v];
    r0 = r1;
    r1 = r2;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float raw = sraw[v];
      const float s = relu ? fmaxf(raw, 0.0f) : raw;
      const float a = 1.0f - expf(-s * delta);
      const float t = in ? (1.0f - a) + 1e-10f : 1.0f;
      const float incl = warp_scan_prod(t, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = carry[v] * excl;
      const float w = a * T;
      carry[v] *= __shfl_sync(0xffffffffu, incl, 31);
      surf[v] *= in ? (1.0f - a) : 1.0f;
      if (in) {
        if (weights) weights[(ray * N + i) * V + v] = w;
        if (alpha_out) alpha_out[(ray * N + i) * V + v] = a;
      }
      if (V == 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += w * c[k];
      } else {
        acc[v] += w * c[v];
      }
    }
  }
  float S[V];
#pragma unroll
  for (int v = 0; v < V; ++v) S[v] = warp_prod(surf[v]);
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0) {
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (tsurf) tsurf[ray * V + v] = S[v];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float cs = 0.0f;
      if (color_surf) {
        const float raw = color_surf[ray * K + k];
        cs = S[V == 1 ? 0 : k] * (relu ? fmaxf(raw, 0.0f) : raw);
      }
      if (catmo) catmo[ray * K + k] = acc[k];
      if (csurf) csurf[ray * K + k] = cs;
      cmap[ray * K + k] = acc[k] + cs;
    }
  }
}

// Backward: front-to-back with the saved totals. For sample i and density channel v
// (q = sum_k g_atmo[k] c[k] over the colour channels that v attenuates):
//   dL/dc[k]  = g_atmo[k] * w_i
//   dL/dsig_v = delta_i * ( e_i * (T_i q_i - (Q_v - Qpre_i) / t_i) - S_v G_v )
// with e = exp(-sig delta), t = (1-a)+1e-10, Q_v = sum_i w_i q_i (= g_atmo . C_atmo),
// Qpre the inclusive prefix of w q, S_v = prod(1-a), G_v = sum_k g_surf[k] c_surf[k].
// COMPACT: dsigma / dcolor are written as a dense LIST of the samples that can carry a gradient
// (with the ReLUs of instant_ngp.py:178-184 a sample whose raw density is <= 0 has alpha = 0, hence
// weight 0, hence dL/dcolour = 0, and the ReLU zeroes dL/dsigma), their sample indices go to
// active_idx and the list length to *n_active (zeroed by the caller). A ray's samples stay
// consecutive and in order (one atomicAdd per ray reserves its block). Empty space costs the field
// backward nothing that way.
// GW: the per-sample weights w_i,v = alpha T are an output that somebody differentiates as well (NeRF
// coarse pass: they define the fine sampler's CDF, samplers.py:72-74): g_weights (B,N,V) = dL/dw joins
// q_i, and Q_v gains sum_i w_i g_weights_i (one pass over the saved weights in front).
template <int K, int V, bool COMPACT, bool GW = false>
__global__ void __launch_bounds__(128)
k_composite_bwd(const float* __restrict__ z, const float* __restrict__ color, const float* __restrict__ sigma,
                const float* __restrict__ color_surf, const float* __restrict__ catmo,
                const float* __restrict__ tsurf, const float* __restrict__ g_atmo,
                const float* __restrict__ g_surf, float zs, int64_t B, int N, int relu,
                float* __restrict__ dcolor, float* __restrict__ dsigma, float* __restrict__ dcolor_surf,
                float* __restrict__ ddelta, float* __restrict__ grad_absmax, uint32_t* __restrict__ active_idx,
                uint32_t* __restrict__ n_active, const float* __restrict__ g_weights = nullptr,
                const float* __restrict__ weights = nullptr) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (ray >= B) return;
  const float* zr = z + ray * N;
  uint32_t list_at = 0;  // COMPACT: next free position of this ray's block of the list
  if (COMPACT) {
    uint32_t cnt = 0;
    if (!relu) {
      cnt = (uint32_t)N;
    } else if (V == 1 && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(sigma) & 15u) == 0) {
      // 16-byte loads, four in flight per lane: the count costs one memory round trip per ray
      const float4* s4 = reinterpret_cast<const float4*>(sigma + ray * N);
      uint32_t mine = 0;
#pragma unroll 4
      for (int q = lane; q < (N >> 2); q += 32) {
        const float4 sv = s4[q];
        mine += (sv.x > 0.0f) + (sv.y > 0.0f) + (sv.z > 0.0f) + (sv.w > 0.0f);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
      cnt = mine;
    } else {
      for (int base = 0; base < N; base += 32) {
        const int i = base + lane;
        bool act = false;
        if (i < N) {
#pragma unroll
          for (int v = 0; v < V; ++v) act = act || sigma[(ray * N + i) * V + v] > 0.0f;
        }
        cnt += __popc(__ballot_sync(0xffffffffu, act));
      }
    }
    if (lane == 0) list_at = atomicAdd(n_active, cnt);
    list_at = __shfl_sync(0xffffffffu, list_at, 0);
  }
  float ga[K], gs[K], Qtot[V], SG[V], carryT[V], carryQ[V];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    ga[k] = g_atmo[ray * K + k];
    gs[k] = (g_surf && color_surf) ? g_surf[ray * K + k] : 0.0f;
  }
#pragma unroll
  for (int v = 0; v < V; ++v) Qtot[v] = 0.0f, SG[v] = 0.0f, carryT[v] = 1.0f, carryQ[v] = 0.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int v = V == 1 ? 0 : k;
    Qtot[v] += ga[k] * catmo[ray * K + k];
    if (color_surf) {
      const float raw = color_surf[ray * K + k];
      const float cs = relu ? fmaxf(raw, 0.0f) : raw;
      const float S = tsurf[ray * V + v];
      SG[v] += S * gs[k] * cs;
      if (lane == 0 && dcolor_surf) dcolor_surf[ray * K + k] = (relu && !(raw > 0.0f)) ? 0.0f : gs[k] * S;
    }
  }
  if (GW) {
    float extra[V];
#pragma unroll
    for (int v = 0; v < V; ++v) extra[v] = 0.0f;
    for (int i = lane; i < N; i += 32) {
#pragma unroll
      for (int v = 0; v < V; ++v) extra[v] += weights[(ray * N + i) * V + v] * g_weights[(ray * N + i) * V + v];
    }
#pragma unroll
    for (int v = 0; v < V; ++v) Qtot[v] += warp_sum(extra[v]);
  }
  float amax = 0.0f;  // largest |gradient| written by this lane (for the fp16 scale of field_bwd_tc)
  RawSample<K, V> r0 = load_raw<K, V>(zr, color, sigma, ray * N, lane, N);
  RawSample<K, V> r1 = load_raw<K, V>(zr, color, sigma, ray * N, 32 + lane, N);
  for (int base = 0; base < N; base += 32) {
    const int i = base + lane;
    const bool in = i < N;
    const RawSample<K, V> r2 = load_raw<K, V>(zr, color, sigma, ray * N, i + 64, N);
    const float delta = in ? raw_delta(r0, i, N, zs) : 0.0f;
    float c[K], craw[K], sraw[V];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      craw[k] = r0.c[k];
      c[k] = relu ? fmaxf(craw[k], 0.0f) : craw[k];
    }
#pragma unroll
    for (int v = 0; v < V; ++v) sraw[v] = r0.s[v];
    r0 = r1;
    r1 = r2;
    uint32_t at = 0;  // COMPACT: this sample's list position
    bool listed = false;
    if (COMPACT) {
      listed = in;
      if (in && relu) {
        listed = false;
#pragma unroll
        for (int v = 0; v < V; ++v) listed = listed || sraw[v] > 0.0f;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, listed);
      at = list_at + __popc(m & ((1u << lane) - 1u));
      list_at += __popc(m);
      if (listed) active_idx[at] = (uint32_t)(ray * N + i);
    }
    float dc[K];
    float dd = 0.0f;  // dL/d(delta_i), summed over density channels
#pragma unroll
    for (int k = 0; k < K; ++k) dc[k] = 0.0f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float raw = sraw[v];
      const float s = relu ? fmaxf(raw, 0.0f) : raw;
      const float e = expf(-s * delta);
      const float a = 1.0f - e;
      const float t = in ? (1.0f - a) + 1e-10f : 1.0f;
      const float incl = warp_scan_prod(t, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = carryT[v] * excl;
      const float w = a * T;
      float q = 0.0f;
      if (V == 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) q += ga[k] * c[k], dc[k] = ga[k] * w;
      } else {
        q = ga[v] * c[v];
        dc[v] = ga[v] * w;
      }
      if (GW && in) q += g_weights[(ray * N + i) * V + v];
      const float wq = in ? w * q : 0.0f;
      const float qincl = carryQ[v] + warp_scan_sum(wq, lane);
      carryT[v] *= __shfl_sync(0xffffffffu, incl, 31);
      carryQ[v] = __shfl_sync(0xffffffffu, qincl, 31);
      const float core = e * (T * q - (Qtot[v] - qincl) / t) - SG[v];  // dL/d(sigma*delta)
      float ds = delta * core;
      dd += s * core;
      if (relu && !(raw > 0.0f)) ds = 0.0f;
      if (COMPACT) {
        if (listed) dsigma[(size_t)at * V + v] = ds, amax = fmaxf(amax, fabsf(ds));
      } else if (in) {
        dsigma[(ray * N + i) * V + v] = ds, amax = fmaxf(amax, fabsf(ds));
      }
    }
    if (in && ddelta) ddelta[ray * N + i] = dd;
    if (COMPACT ? listed : in) {
      const size_t orow = COMPACT ? (size_t)at : (size_t)(ray * N + i);
      float dk[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        dk[k] = (relu && !(craw[k] > 0.0f)) ? 0.0f : dc[k];
        amax = fmaxf(amax, fabsf(dk[k]));
      }
      if (K == 4) {
        *reinterpret_cast<float4*>(dcolor + orow * 4) = make_float4(dk[0], dk[1 % K], dk[2 % K], dk[3 % K]);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) dcolor[orow * K + k] = dk[k];
      }
    }
  }
  if (grad_absmax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    // non-negative floats order like their bit patterns
    if (lane == 0 && amax > 0.0f && amax < INFINITY) atomicMax(reinterpret_cast<int*>(grad_absmax), __float_as_int(amax));
  }
}

// =========================================================================================
// Per-band loss + gradient (instant_ngp.py:249-263, losses.py:5-33)
// =========================================================================================
__device__ __forceinline__ void loss_term(int kind, float p, float g, float max_i, float& val, float& dp) {
  const float eps = 1e-3f * max_i;
  float v_mse = 0, d_mse = 0, v_hdr = 0, d_hdr = 0, v_l1 = 0, d_l1 = 0;
  if (kind == 4 || kind == 5) {
    const float r = p / max_i - g / max_i;
    v_mse = r * r;
    d_mse = 2.0f * r / max_i;
  }
  if (kind == 2 || kind == 3) {
    const float r = p / max_i - g / max_i;
    v_l1 = fabsf(r);
    d_l1 = (r > 0.0f ? 1.0f : (r < 0.0f ? -1.0f : 0.0f)) / max_i;
  }
  if (kind == 1 || kind == 3 || kind == 5) {
    const float r = logf(g + eps) - logf(p + eps);
    v_hdr = r * r;
    d_hdr = -2.0f * r / (p + eps);
  }
  if (kind == 0) {
    const float r = (p - g) / (p + eps);
    val = r * r;
    dp = 2.0f * r / (p + eps);
    return;
  }
  if (kind == 1) { val = v_hdr; dp = d_hdr; }
  else if (kind == 2) { val = v_l1; dp = d_l1; }
  else if (kind == 3) { val = v_l1 + 0.2f * v_hdr; dp = d_l1 + 0.2f * d_hdr; }
  else if (kind == 4) { val = v_mse; dp = d_mse; }
  else { val = v_mse + 0.2f * v_hdr; dp = d_mse + 0.2f * d_hdr; }
}

__global__ void __launch_bounds__(256)
k_band_loss(const float* __restrict__ cmap, const int64_t* __restrict__ band, const float* __restrict__ rad,
            float max_i, int kind, int64_t B, int K, float gscale, float* __restrict__ dcmap,
            float* __restrict__ partial) {
  __shared__ float red[8];
  float sum = 0.0f;
  const float invB = 1.0f / (float)B;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)band[i];
    float val, dp;
    loss_term(kind, cmap[i * K + b], rad[i], max_i, val, dp);
    sum += val;
    if (dcmap) {
      for (int k = 0; k < K; ++k) dcmap[i * K + k] = k == b ? dp * invB * gscale : 0.0f;
    }
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int w = 0; w < 8; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}
__global__ void k_loss_finish(const float* __restrict__ partial, int n, int64_t B, float* __restrict__ loss) {
  // single warp, fixed order -> deterministic
  float s = 0.0f;
  for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) loss[0] = s / (float)B;
}

// =========================================================================================
// Fused AdamW (dense; mirrors torch.optim.AdamW single-tensor arithmetic)
// =========================================================================================
struct AdamConsts {
  float decay_mul;      // 1 - lr * weight_decay
  float w1;             // 1 - beta1 (lerp weight)
  float beta2;
  float w2;             // 1 - beta2
  float bc2_sqrt;       // sqrt(1 - beta2^step)
  float neg_step_size;  // -lr / (1 - beta1^step)
  float eps;
  float gscale;
};

__device__ __forceinline__ void adam_one(const AdamConsts& c, float& p, float g, float& m, float& v) {
  const float gr = g * c.gscale;
  p = p * c.decay_mul;                                                             // param.mul_(1 - lr*wd)
  m = c.w1 < 0.5f ? m + c.w1 * (gr - m) : gr - (gr - m) * (1.0f - c.w1);           // exp_avg.lerp_(grad, 1-beta1)
  v = v * c.beta2;                                                                 // exp_avg_sq.mul_(beta2)
  v = v + (c.w2 * gr) * gr;                                                        //   .addcmul_(g, g, 1-beta2)
  const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;
  p = p + (c.neg_step_size * m) / denom;                                           // param.addcdiv_(m, denom, -step)
}

__global__ void __launch_bounds__(256)
k_adamw(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
        __half* __restrict__ p16, int64_t n, AdamConsts c, int zero_grad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      float4 P = *reinterpret_cast<float4*>(p + i), G = *reinterpret_cast<float4*>(g + i);
      float4 Mv = *reinterpret_cast<float4*>(m + i), Vv = *reinterpret_cast<float4*>(v + i);
      adam_one(c, P.x, G.x, Mv.x, Vv.x);
      adam_one(c, P.y, G.y, Mv.y, Vv.y);
      adam_one(c, P.z, G.z, Mv.z, Vv.z);
      adam_one(c, P.w, G.w, Mv.w, Vv.w);
      *reinterpret_cast<float4*>(p + i) = P;
      *reinterpret_cast<float4*>(m + i) = Mv;
      *reinterpret_cast<float4*>(v + i) = Vv;
      if (zero_grad) *reinterpret_cast<float4*>(g + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p16) {
        *reinterpret_cast<__half2*>(p16 + i) = __floats2half2_rn(P.x, P.y);
        *reinterpret_cast<__half2*>(p16 + i + 2) = __floats2half2_rn(P.z, P.w);
      }
    } else {
      for (int64_t j = i; j < n; ++j) {
        float pj = p[j], mj = m[j], vj = v[j];
        adam_one(c, pj, g[j], mj, vj);
        p[j] = pj; m[j] = mj; v[j] = vj;
        if (zero_grad) g[j] = 0.0f;
        if (p16) p16[j] = __float2half_rn(pj);
      }
    }
  }
}

// =========================================================================================
// Extraction: float64 point -> preprocess -> hash -> pos_mlp -> max(sigma, 0)
// =========================================================================================
__global__ void __launch_bounds__(kTile)
k_extract_sigma(atmonr_frame_t f, GeoFrame gf, atmonr_grid_t g, const __half2* __restrict__ table,
                const __half* __restrict__ pos_w, const double* __restrict__ pts, int64_t n,
                float alt_compress, float* __restrict__ sigma) {
  __shared__ __align__(16) float sW[PosMlp::kNumWeights];
  load_weights(pos_w, sW, PosMlp::kNumWeights);
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile * kTile < n; tile += gridDim.x) {
    const int64_t i = tile * kTile + threadIdx.x;
    if (i >= n) continue;
    double c0 = pts[3 * i], c1 = pts[3 * i + 1], c2 = pts[3 * i + 2];
    if (f.enabled) preprocess_f64(f, gf, c0, c1, c2, c0, c1, c2);
    // instant_ngp.py:224-233 stay in float64; tcnn casts its input to float32
    const float p[3] = {(float)((c0 + 1.0) / 2.0), (float)((c1 + 1.0) / 2.0),
                        (float)(((c2 + 1.0) / 2.0) / (double)alt_compress)};
    __half2 enc[16], hp[1][16];
    hash_encode<3>(g, table, p, enc);
    float po[1];
    mlp_forward<32, 1, 1>(sW, enc, hp, po);
    sigma[i] = fmaxf(po[0], 0.0f);
  }
}

// =========================================================================================
// NeRF helpers: positional encoding and inverse-CDF sampling
// =========================================================================================
struct PeCfg {
  int32_t freqs[4];
  int32_t col0[4];
  int32_t C, width, interleaved;
};

// T = float: the training path (float32 points, float32 phases). T = double: the extract path, where
// the reference keeps the float64 points of scripts/extract.py through the encoder (nerf.py:209-213):
// the float32 factor 2^l * pi is widened, the phase and sin / cos are float64, the result is rounded
// once to float32.
template <typename T>
__global__ void k_positional_encoding(const T* __restrict__ pts, int64_t M, PeCfg cfg, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float PI_F = 3.14159265358979323846f;
  float* row = out + i * cfg.width;
  for (int a = 0; a < cfg.C; ++a) {
    const T p = pts[i * cfg.C + a];
    const int L = cfg.freqs[a];
    float f = 1.0f;
    for (int l = 0; l < L; ++l) {
      float sn, cs;
      if constexpr (sizeof(T) == 8) {
        const double arg = (double)(f * PI_F) * p;
        sn = (float)sin(arg), cs = (float)cos(arg);
      } else {
        const float arg = (f * PI_F) * p;
        sn = sinf(arg), cs = cosf(arg);
      }
      if (cfg.interleaved) {
        row[cfg.col0[a] + 2 * l] = sn;
        row[cfg.col0[a] + 2 * l + 1] = cs;
      } else {
        row[cfg.col0[a] + l] = sn;
        row[cfg.col0[a] + L + l] = cs;
      }
      f *= 2.0f;
    }
  }
}

// samplers.py:72-101: one warp per ray. Shared memory: cdf[Nc-1], mids[Nc-1], merged[P] with
// P the next power of two >= Nc+Nf (bitonic sort).
__global__ void k_sample_pdf(const float* __restrict__ weights, const float* __restrict__ zc,
                             const float* __restrict__ u, int64_t B, int Nc, int Nf, int P,
                             float* __restrict__ zout, int64_t* __restrict__ inds) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x;
  const int64_t ray = blockIdx.x;
  if (ray >= B) return;
  const int nb = Nc - 2;       // pdf bins
  const int ncdf = Nc - 1;     // cdf entries == number of mid points
  float* cdf = sm;
  float* mids = sm + ncdf;
  float* merged = mids + ncdf;
  const float* w = weights + ray * Nc;
  const float* zr = zc + ray * Nc;
  if (lane == 0) {
    // torch CPU accumulates sum/cumsum of float32 in a wider type and rounds on store
    double tot_d = 0.0;
    for (int j = 0; j < nb; ++j) tot_d += (double)(w[1 + j] + 1e-8f);
    const float tot = (float)tot_d;
    double run = 0.0;
    cdf[0] = 0.0f;
    for (int j = 0; j < nb; ++j) {
      run += (double)((w[1 + j] + 1e-8f) / tot);
      cdf[1 + j] = (float)run;
    }
  }
  for (int j = lane; j < ncdf; j += 32) mids[j] = 0.5f * (zr[j + 1] + zr[j]);
  for (int j = lane; j < P; j += 32) merged[j] = j < Nc ? zr[j] : INFINITY;
  __syncwarp();
  for (int s = lane; s < Nf; s += 32) {
    const float uu = u[ray * Nf + s];
    int lo = 0, hi = ncdf;  // first index with cdf[idx] > uu  (right=True)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= uu) lo = mid + 1; else hi = mid;
    }
    const int ind = lo;
    const int below = max(ind - 1, 0), above = min(ind, ncdf - 1);
    float den = cdf[above] - cdf[below];
    if (den < 1e-8f) den = 1.0f;
    const float t = (uu - cdf[below]) / den;
    merged[Nc + s] = mids[below] + t * (mids[above] - mids[below]);
    if (inds) inds[ray * Nf + s] = ind;
  }
  __syncwarp();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < P; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = merged[i], b = merged[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { merged[i] = b; merged[ixj] = a; }
        }
      }
      __syncwarp();
    }
  for (int j = lane; j < Nc + Nf; j += 32) zout[ray * (Nc + Nf) + j] = merged[j];
}

#endif  // ATM_PART_BASIC
}  // namespace atm

// =========================================================================================
// C ABI
// =========================================================================================
using namespace atm;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename Kern>
static int set_smem(Kern k, size_t bytes, const char* name) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(name, cudaGetErrorString(e));
  }
  return 0;
}

static bool mlp_supported(const atmonr_mlp_t* m) {
  return m && m->width == kWidth && m->out_pad == kOutPad && (m->n_hidden == 1 || m->n_hidden == 2) &&
         (m->in_pad == 16 || m->in_pad == 32 || m->in_pad == 48) && m->n_in <= m->in_pad && m->n_out <= m->out_pad;
}
static bool is_shape(const atmonr_mlp_t* m, int in_pad, int nh) {
  return mlp_supported(m) && m->in_pad == in_pad && m->n_hidden == nh;
}

#if ATM_PART_MLP
template <int IN, int NH>
static int launch_mlp_fwd(const atmonr_mlp_t* m, const void* w, const float* x, int64_t M, float* out, void* stream) {
  using Sh = MlpShape<IN, NH>;
  const size_t smem = Sh::kNumWeights * sizeof(float);
  if (set_smem(k_mlp_fwd<IN, NH>, smem, "atmonr_mlp_fwd")) return -1;
  const int grid = grid_for((M + kTile - 1) / kTile, 1, num_sms() * 8);
  k_mlp_fwd<IN, NH><<<grid, kTile, smem, S(stream)>>>((const __half*)w, x, m->n_in, m->n_out, M, out);
  ATM_CHECK_LAUNCH("atmonr_mlp_fwd");
  return 0;
}
template <int IN, int NH>
static int launch_mlp_bwd(const atmonr_mlp_t* m, const void* w, const float* x, const float* dout, int64_t M,
                          float* dx, float* dw, void* stream) {
  using Sh = MlpShape<IN, NH>;
  const size_t smem = (2 * Sh::kNumWeights + BwdScratch<IN>::kFloats) * sizeof(float);
  if (set_smem(k_mlp_bwd<IN, NH>, smem, "atmonr_mlp_bwd")) return -1;
  const int grid = grid_for((M + kTile - 1) / kTile, 1, num_sms() * 2);
  k_mlp_bwd<IN, NH><<<grid, kTile, smem, S(stream)>>>((const __half*)w, x, dout, m->n_in, m->n_out, M, dx, dw);
  ATM_CHECK_LAUNCH("atmonr_mlp_bwd");
  return 0;
}

#endif  // ATM_PART_MLP
extern "C" {

#if ATM_PART_BASIC
int atmonr_abi_version(void) { return ATMONR_ABI_VERSION; }
const char* atmonr_last_error(void) { return g_last_error; }

int atmonr_l2_persist(const void* ptr, size_t bytes, float hit_ratio, void* stream) {
  // cudaAccessPolicyWindow on `stream`: accesses to [ptr, ptr + bytes) by kernels launched on the stream
  // afterwards are kept in the L2's persisting set-aside; bytes == 0 clears the window.
  int dev = 0, max_window = 0, max_persist = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail("atmonr_l2_persist", "no device");
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  if (bytes == 0 || !ptr) {
    attr.accessPolicyWindow.num_bytes = 0;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    cudaStreamSetAttribute(S(stream), cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaCtxResetPersistingL2Cache();
    return 0;
  }
  ATM_REQUIRE(max_window > 0 && max_persist > 0, "atmonr_l2_persist", "the device has no persisting L2");
  size_t want = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
  if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
    cudaGetLastError();
    return fail("atmonr_l2_persist", "cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize) failed");
  }
  attr.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
  attr.accessPolicyWindow.num_bytes = bytes < (size_t)max_window ? bytes : (size_t)max_window;
  attr.accessPolicyWindow.hitRatio = hit_ratio * (float)((double)want / (double)attr.accessPolicyWindow.num_bytes < 1.0
                                                        ? (double)want / (double)attr.accessPolicyWindow.num_bytes : 1.0);
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(S(stream), cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) {
    cudaGetLastError();
    return fail("atmonr_l2_persist", "cudaStreamSetAttribute(access policy window) failed");
  }
  return 0;
}

int atmonr_grid_layout(int n_dims, int n_levels, int log2_hashmap_size, int base_resolution,
                       float per_level_scale, atmonr_grid_t* out) {
  ATM_REQUIRE(out, "atmonr_grid_layout", "null output");
  ATM_REQUIRE(n_dims >= 2 && n_dims <= 4, "atmonr_grid_layout", "n_dims must be 2, 3 or 4");
  ATM_REQUIRE(n_levels >= 1 && n_levels <= ATMONR_MAX_LEVELS, "atmonr_grid_layout", "n_levels out of range");
  ATM_REQUIRE(log2_hashmap_size >= 3 && log2_hashmap_size <= 30, "atmonr_grid_layout", "log2_hashmap_size out of range");
  memset(out, 0, sizeof(*out));
  out->n_dims = n_dims;
  out->n_levels = n_levels;
  out->n_feat = 2;
  const uint32_t cap = 1u << log2_hashmap_size;
  const uint32_t max_params = 0xFFFFFFFFu / 2;
  uint32_t offset = 0;
  for (int l = 0; l < n_levels; ++l) {
    const float scale = grid_level_scale(l, per_level_scale, base_resolution);
    const uint32_t res = (uint32_t)ceilf(scale) + 1u;
    uint32_t n;
    if (powf((float)res, (float)n_dims) > (float)max_params) {
      n = max_params;
    } else {
      n = 1;
      for (int k = 0; k < n_dims; ++k) n *= res;
    }
    n = (n + 7u) / 8u * 8u;
    if (n > cap) n = cap;
    out->scale[l] = scale;
    out->res[l] = res;
    out->size[l] = n;
    out->offset[l] = offset;
    offset += n;
  }
  out->offset[n_levels] = offset;
  return 0;
}

int atmonr_sample_uniform(const float* origin, const float* dir, const float* len, const float* u,
                          const float* bins, int64_t B, int N, int mode, uint64_t seed,
                          uint64_t ray_index_base, float* pts, float* z, void* stream) {
  ATM_REQUIRE(mode >= 0 && mode <= 2 && (mode != 1 || u), "atmonr_sample_uniform", "bad mode / missing u");
  if (B * N == 0) return 0;
  k_sample_uniform<<<grid_for(B * N, 256), 256, 0, S(stream)>>>(origin, dir, len, u, bins, B * N, N, mode, seed,
                                                               ray_index_base, pts, z);
  ATM_CHECK_LAUNCH("atmonr_sample_uniform");
  return 0;
}

int atmonr_preprocess_horizontal(const atmonr_frame_t* f, const void* pts, void* out, int64_t n, int is_f64,
                                 void* stream) {
  ATM_REQUIRE(f && f->enabled, "atmonr_preprocess_horizontal", "frame missing or disabled");
  if (n == 0) return 0;
  const GeoFrame gf = make_geo_frame(*f);
  if (is_f64)
    k_preprocess_f64<<<grid_for(n, 256), 256, 0, S(stream)>>>(*f, gf, (const double*)pts, (double*)out, n);
  else
    k_preprocess_f32<<<grid_for(n, 256), 256, 0, S(stream)>>>(*f, gf, (const float*)pts, (float*)out, n);
  ATM_CHECK_LAUNCH("atmonr_preprocess_horizontal");
  return 0;
}

int atmonr_ngp_sample_points(const atmonr_frame_t* f, const float* origin, const float* dir, const float* len,
                             const float* u, const float* bins, int64_t B, int N, int mode, uint64_t seed,
                             uint64_t ray_index_base, float alt_compress, float* x01, float* z, void* stream) {
  ATM_REQUIRE(f, "atmonr_ngp_sample_points", "null frame");
  ATM_REQUIRE(mode >= 0 && mode <= 2 && (mode != 1 || u), "atmonr_ngp_sample_points", "bad mode / missing u");
  if (B * N == 0) return 0;
  SamplerJob job;
  if (make_sampler_job(job, f, origin, dir, len, u, bins, B, N, mode, seed, ray_index_base, alt_compress, x01, z))
    k_ngp_sample_points4<<<grid_for(job.groups, 128), 128, 0, S(stream)>>>(job);
  else
    k_ngp_sample_points<false><<<grid_for(B * N, 256), 256, 0, S(stream)>>>(*f, make_geo_frame(*f), origin, dir, len, u,
                                                                           bins, B * N, N, mode, seed, ray_index_base,
                                                                           alt_compress, HeightArgs{}, x01, z);
  ATM_CHECK_LAUNCH("atmonr_ngp_sample_points");
  return 0;
}

int atmonr_ngp_sample_points_height(const atmonr_frame_t* f, const float* origin, const float* dir, const float* len,
                                    const float* u, const float* bins, int64_t B, int N, int mode, uint64_t seed,
                                    uint64_t ray_index_base, float alt_compress, double scale, const double* offset_host,
                                    double ray_origin_height, float* x01, float* z, void* stream) {
  ATM_REQUIRE(f && offset_host, "atmonr_ngp_sample_points_height", "null frame / offset");
  ATM_REQUIRE(mode >= 0 && mode <= 2 && (mode != 1 || u), "atmonr_ngp_sample_points_height", "bad mode / missing u");
  ATM_REQUIRE((reinterpret_cast<uintptr_t>(x01) & 15u) == 0, "atmonr_ngp_sample_points_height", "x01 must be 16-byte aligned");
  if (B * N == 0) return 0;
  const HeightArgs ha{scale, offset_host[0], offset_host[1], offset_host[2], ray_origin_height};
  k_ngp_sample_points<true><<<grid_for(B * N, 256), 256, 0, S(stream)>>>(*f, make_geo_frame(*f), origin, dir, len, u, bins,
                                                                          B * N, N, mode, seed, ray_index_base,
                                                                          alt_compress, ha, x01, z);
  ATM_CHECK_LAUNCH("atmonr_ngp_sample_points_height");
  return 0;
}

// one instantiation per grid dimensionality: 2 (surface), 3 (positions), 4 (positions + height,
// `include_height`, samplers.py:168-195)
#define ATM_GRID_DISPATCH(g, KERNEL, ...)                           \
  if ((g)->n_dims == 2) { KERNEL<2> __VA_ARGS__; }                  \
  else if ((g)->n_dims == 3) { KERNEL<3> __VA_ARGS__; }             \
  else { KERNEL<4> __VA_ARGS__; }

int atmonr_hashgrid_fwd(const atmonr_grid_t* g, const float* x, int xs, const void* table, int64_t M, float* out,
                        void* stream) {
  ATM_REQUIRE(g && g->n_feat == 2, "atmonr_hashgrid_fwd", "bad grid");
  if (M == 0) return 0;
  const int grid = grid_for(M, 128);
  ATM_GRID_DISPATCH(g, k_hashgrid_fwd, <<<grid, 128, 0, S(stream)>>>(*g, x, xs, (const __half2*)table, M, out));
  ATM_CHECK_LAUNCH("atmonr_hashgrid_fwd");
  return 0;
}

int atmonr_hashgrid_bwd(const atmonr_grid_t* g, const float* x, int xs, const float* dout, int64_t M,
                        float* dtable, void* stream) {
  ATM_REQUIRE(g && g->n_feat == 2, "atmonr_hashgrid_bwd", "bad grid");
  if (M == 0) return 0;
  const int grid = grid_for(M, 128);
  ATM_GRID_DISPATCH(g, k_hashgrid_bwd, <<<grid, 128, 0, S(stream)>>>(*g, x, xs, dout, M, dtable));
  ATM_CHECK_LAUNCH("atmonr_hashgrid_bwd");
  return 0;
}

int atmonr_hashgrid_indices(const atmonr_grid_t* g, const float* x, int xs, int64_t M, uint32_t* idx,
                            void* stream) {
  ATM_REQUIRE(g, "atmonr_hashgrid_indices", "bad grid");
  if (M == 0) return 0;
  const int grid = grid_for(M, 128);
  ATM_GRID_DISPATCH(g, k_hashgrid_indices, <<<grid, 128, 0, S(stream)>>>(*g, x, xs, M, idx));
  ATM_CHECK_LAUNCH("atmonr_hashgrid_indices");
  return 0;
}

#endif  // ATM_PART_BASIC
#if ATM_PART_MLP
int atmonr_mlp_fwd(const atmonr_mlp_t* m, const void* w, const float* x, int64_t M, float* out, void* stream) {
  ATM_REQUIRE(mlp_supported(m), "atmonr_mlp_fwd", "unsupported MLP shape (width 32, in_pad 16|32|48, 1|2 hidden layers)");
  if (M == 0) return 0;
  if (m->in_pad == 16 && m->n_hidden == 1) return launch_mlp_fwd<16, 1>(m, w, x, M, out, stream);
  if (m->in_pad == 16 && m->n_hidden == 2) return launch_mlp_fwd<16, 2>(m, w, x, M, out, stream);
  if (m->in_pad == 32 && m->n_hidden == 1) return launch_mlp_fwd<32, 1>(m, w, x, M, out, stream);
  if (m->in_pad == 32 && m->n_hidden == 2) return launch_mlp_fwd<32, 2>(m, w, x, M, out, stream);
  if (m->in_pad == 48 && m->n_hidden == 1) return launch_mlp_fwd<48, 1>(m, w, x, M, out, stream);
  return launch_mlp_fwd<48, 2>(m, w, x, M, out, stream);
}

int atmonr_mlp_bwd(const atmonr_mlp_t* m, const void* w, const float* x, const float* dout, int64_t M, float* dx,
                   float* dw, void* stream) {
  ATM_REQUIRE(mlp_supported(m), "atmonr_mlp_bwd", "unsupported MLP shape (width 32, in_pad 16|32|48, 1|2 hidden layers)");
  ATM_REQUIRE(dw, "atmonr_mlp_bwd", "null dw");
  if (M == 0) return 0;
  if (m->in_pad == 16 && m->n_hidden == 1) return launch_mlp_bwd<16, 1>(m, w, x, dout, M, dx, dw, stream);
  if (m->in_pad == 16 && m->n_hidden == 2) return launch_mlp_bwd<16, 2>(m, w, x, dout, M, dx, dw, stream);
  if (m->in_pad == 32 && m->n_hidden == 1) return launch_mlp_bwd<32, 1>(m, w, x, dout, M, dx, dw, stream);
  if (m->in_pad == 32 && m->n_hidden == 2) return launch_mlp_bwd<32, 2>(m, w, x, dout, M, dx, dw, stream);
  if (m->in_pad == 48 && m->n_hidden == 1) return launch_mlp_bwd<48, 1>(m, w, x, dout, M, dx, dw, stream);
  return launch_mlp_bwd<48, 2>(m, w, x, dout, M, dx, dw, stream);
}

#endif  // ATM_PART_MLP
#if ATM_PART_FIELD
// -> number of densities V (1 or 4), or -1 with the error recorded
static int check_field(const atmonr_grid_t* g, const atmonr_mlp_t* pm, const atmonr_mlp_t* dm, const char* name) {
  ATM_REQUIRE(g && (g->n_dims == 3 || g->n_dims == 4) && g->n_feat == 2 && g->n_levels == 16, name,
              "field needs a 3-D or 4-D grid with 16 levels x 2 features");
  ATM_REQUIRE(is_shape(pm, 32, 1) && pm->n_in == 32 && pm->n_out == 16, name, "pos_mlp must be 32 -> [32] -> 16");
  const bool v1 = is_shape(dm, 32, 2) && dm->n_in == 19 && dm->n_out == 4;
  const bool v4 = is_shape(dm, 16, 2) && dm->n_in == 16 && dm->n_out == 4;
  ATM_REQUIRE(v1 || v4, name, "dir_mlp must be 19 -> [32,32] -> 4 (one density) or 16 -> [32,32] -> 4 (four densities)");
  return v1 ? 1 : 4;
}

#define ATM_DV_DISPATCH(D, V, CALL)                                        \
  if ((D) == 3 && (V) == 1) { CALL(3, 1); } else if ((D) == 4 && (V) == 1) { CALL(4, 1); } \
  else if ((D) == 3) { CALL(3, 4); } else { CALL(4, 4); }

int atmonr_ngp_field_fwd(const atmonr_grid_t* g, const void* table, const atmonr_mlp_t* pm, const void* pos_w,
                         const atmonr_mlp_t* dm, const void* dir_w, const float* x01, const float* dirs, int64_t B,
                         int N, float* sigma_raw, float* color_raw, void* stream) {
  const int V = check_field(g, pm, dm, "atmonr_ngp_field_fwd");
  if (V < 0) return -1;
  const int64_t M = B * N;
  if (M == 0) return 0;
  const int grid = grid_for((M + kTile - 1) / kTile, 1, num_sms() * 16);
#define CALL(DD, VV)                                                                                           \
  k_field_fwd<DD, VV><<<grid, kTile, 0, S(stream)>>>(*g, (const __half2*)table, (const __half*)pos_w,          \
                                                    (const __half*)dir_w, x01, dirs, M, N, sigma_raw, color_raw)
  ATM_DV_DISPATCH(g->n_dims, V, CALL)
#undef CALL
  ATM_CHECK_LAUNCH("atmonr_ngp_field_fwd");
  return 0;
}

int atmonr_ngp_field_bwd(const atmonr_grid_t* g, const void* table, const atmonr_mlp_t* pm, const void* pos_w,
                         const atmonr_mlp_t* dm, const void* dir_w, const float* x01, const float* dirs,
                         const float* dsigma_raw, const float* dcolor_raw, int64_t B, int N, float* dtable,
                         float* dpos_w, float* ddir_w, void* stream) {
  const int V = check_field(g, pm, dm, "atmonr_ngp_field_bwd");
  if (V < 0) return -1;
  const int64_t M = B * N;
  if (M == 0) return 0;
  const int dir_weights = V == 1 ? MlpShape<32, 2>::kNumWeights : MlpShape<16, 2>::kNumWeights;
  const size_t smem = (2 * (PosMlp::kNumWeights + dir_weights) + BwdScratch<32>::kFloats) * sizeof(float);
  const int grid = grid_for((M + kTile - 1) / kTile, 1, num_sms() * 3);
#define CALL(DD, VV)                                                                                               \
  {                                                                                                                \
    if (set_smem(k_field_bwd<DD, VV>, smem, "atmonr_ngp_field_bwd")) return -1;                                    \
    k_field_bwd<DD, VV><<<grid, kTile, smem, S(stream)>>>(*g, (const __half2*)table, (const __half*)pos_w,         \
                                                         (const __half*)dir_w, x01, dirs, dsigma_raw, dcolor_raw, \
                                                         M, N, dtable, dpos_w, ddir_w);                            \
  }
  ATM_DV_DISPATCH(g->n_dims, V, CALL)
#undef CALL
  ATM_CHECK_LAUNCH("atmonr_ngp_field_bwd");
  return 0;
}

#endif  // ATM_PART_FIELD
#if ATM_PART_SURF
static int check_surface(const atmonr_grid_t* g, const atmonr_mlp_t* m, const char* name) {
  ATM_REQUIRE(g && g->n_dims == 2 && g->n_feat == 2 && g->n_levels == 16, name, "surface needs a 2-D grid with 16 levels x 2 features");
  ATM_REQUIRE(is_shape(m, 48, 2) && m->n_in == 36 && m->n_out == 4, name, "surf_mlp must be 36 -> [32,32] -> 4");
  return 0;
}

int atmonr_ngp_surface_fwd(const atmonr_grid_t* g, const void* table, const atmonr_mlp_t* m, const void* w,
                           const float* origin, const float* dir, const float* len, int64_t B,
                           float* color_surf_raw, void* stream) {
  if (check_surface(g, m, "atmonr_ngp_surface_fwd")) return -1;
  if (B == 0) return 0;
  const size_t smem = SurfMlp::kNumWeights * sizeof(float);
  const int grid = grid_for((B + kTile - 1) / kTile, 1, num_sms() * 8);
  k_surface_fwd<<<grid, kTile, smem, S(stream)>>>(*g, (const __half2*)table, (const __half*)w, origin, dir, len, B,
                                                 color_surf_raw);
  ATM_CHECK_LAUNCH("atmonr_ngp_surface_fwd");
  return 0;
}

int atmonr_ngp_surface_bwd(const atmonr_grid_t* g, const void* table, const atmonr_mlp_t* m, const void* w,
                           const float* origin, const float* dir, const float* len, const float* dcs, int64_t B,
                           float* dtable, float* dw, void* stream) {
  if (check_surface(g, m, "atmonr_ngp_surface_bwd")) return -1;
  if (B == 0) return 0;
  const size_t smem = (2 * SurfMlp::kNumWeights + BwdScratch<48>::kFloats) * sizeof(float);
  if (set_smem(k_surface_bwd, smem, "atmonr_ngp_surface_bwd")) return -1;
  const int grid = grid_for((B + kTile - 1) / kTile, 1, num_sms() * 2);
  k_surface_bwd<<<grid, kTile, smem, S(stream)>>>(*g, (const __half2*)table, (const __half*)w, origin, dir, len, dcs,
                                                 B, dtable, dw);
  ATM_CHECK_LAUNCH("atmonr_ngp_surface_bwd");
  return 0;
}

#endif  // ATM_PART_SURF
#if ATM_PART_BASIC
#define ATM_KV_DISPATCH(K, V, CALL)                                                       \
  if (K == 4 && V == 1) { CALL(4, 1); } else if (K == 4 && V == 4) { CALL(4, 4); }          \
  else if (K == 3 && V == 1) { CALL(3, 1); } else if (K == 3 && V == 3) { CALL(3, 3); }     \
  else if (K == 2 && V == 1) { CALL(2, 1); } else if (K == 2 && V == 2) { CALL(2, 2); }     \
  else if (K == 1 && V == 1) { CALL(1, 1); } else { return fail("atmonr_composite", "unsupported (K, V); need K<=4 and V in {1, K}"); }

int atmonr_composite_fwd(const float* z, const float* color, const float* sigma, const float* color_surf,
                         float z_scale, int64_t B, int N, int K, int V, int relu, float* color_map,
                         float* color_map_atmo, float* color_map_surf, float* trans_surf, float* weights,
                         float* alpha, void* stream) {
  if (B == 0) return 0;
  ATM_REQUIRE(color_map, "atmonr_composite_fwd", "null color_map");
  ATM_REQUIRE(K != 4 || (reinterpret_cast<uintptr_t>(color) & 15u) == 0, "atmonr_composite_fwd", "color must be 16-byte aligned");
  const int grid = grid_for(B * 32, 128);
#define CALL(KK, VV)                                                                                      \
  k_composite_fwd<KK, VV><<<grid, 128, 0, S(stream)>>>(z, color, sigma, color_surf, z_scale, B, N, relu,   \
                                                      color_map, color_map_atmo, color_map_surf, trans_surf, \
                                                      weights, alpha)
  ATM_KV_DISPATCH(K, V, CALL)
#undef CALL
  ATM_CHECK_LAUNCH("atmonr_composite_fwd");
  return 0;
}

int atmonr_composite_bwd(const float* z, const float* color, const float* sigma, const float* color_surf,
                         const float* color_map_atmo, const float* trans_surf, const float* d_atmo,
                         const float* d_surf, float z_scale, int64_t B, int N, int K, int V, int relu,
                         float* dcolor, float* dsigma, float* dcolor_surf, float* ddelta, float* grad_absmax,
                         void* stream) {
  if (B == 0) return 0;
  ATM_REQUIRE(color_map_atmo && d_atmo && dcolor && dsigma, "atmonr_composite_bwd", "null argument");
  ATM_REQUIRE(!color_surf || trans_surf, "atmonr_composite_bwd", "trans_surf required with a surface");
  ATM_REQUIRE(K != 4 || ((reinterpret_cast<uintptr_t>(color) | reinterpret_cast<uintptr_t>(dcolor)) & 15u) == 0,
              "atmonr_composite_bwd", "color and dcolor must be 16-byte aligned");
  const int grid = grid_for(B * 32, 128);
#define CALL(KK, VV)                                                                                                \
  k_composite_bwd<KK, VV, false><<<grid, 128, 0, S(stream)>>>(z, color, sigma, color_surf, color_map_atmo, trans_surf, \
                                                             d_atmo, d_surf, z_scale, B, N, relu, dcolor, dsigma,    \
                                                             dcolor_surf, ddelta, grad_absmax, nullptr, nullptr)
  ATM_KV_DISPATCH(K, V, CALL)
#undef CALL
  ATM_CHECK_LAUNCH("atmonr_composite_bwd");
  return 0;
}

int atmonr_composite_bwd_weights(const float* z, const float* color, const float* sigma, const float* color_surf,
                                 const float* color_map_atmo, const float* trans_surf, const float* weights,
                                 const float* d_atmo, const float* d_surf, const float* d_weights, float z_scale,
                                 int64_t B, int N, int K, int V, int relu, float* dcolor, float* dsigma,
                                 float* dcolor_surf, float* ddelta, void* stream) {
  if (B == 0) return 0;
  ATM_REQUIRE(color_map_atmo && d_atmo && dcolor && dsigma && weights && d_weights, "atmonr_composite_bwd_weights",
              "null argument");
  ATM_REQUIRE(!color_surf || trans_surf, "atmonr_composite_bwd_weights", "trans_surf required with a surface");
  ATM_REQUIRE(K != 4 || ((reinterpret_cast<uintptr_t>(color) | reinterpret_cast<uintptr_t>(dcolor)) & 15u) == 0,
              "atmonr_composite_bwd_weights", "color and dcolor must be 16-byte aligned");
  const int grid = grid_for(B * 32, 128);
#define CALL(KK, VV)                                                                                                  \
  k_composite_bwd<KK, VV, false, true><<<grid, 128, 0, S(stream)>>>(                                                  \
      z, color, sigma, color_surf, color_map_atmo, trans_surf, d_atmo, d_surf, z_scale, B, N, relu, dcolor, dsigma,    \
      dcolor_surf, ddelta, nullptr, nullptr, nullptr, d_weights, weights)
  ATM_KV_DISPATCH(K, V, CALL)
#undef CALL
  ATM_CHECK_LAUNCH("atmonr_composite_bwd_weights");
  return 0;
}

int atmonr_composite_bwd_compact(const float* z, const float* color, const float* sigma, const float* color_surf,
                                 const float* color_map_atmo, const float* trans_surf, const float* d_atmo,
                                 const float* d_surf, float z_scale, int64_t B, int N, int K, int V, int relu,
                                 uint32_t* active_idx, uint32_t* n_active, float* dcolor_c, float* dsigma_c,
                                 float* dcolor_surf, float* grad_absmax, void* stream) {
  if (B == 0) return 0;
  ATM_REQUIRE(color_map_atmo && d_atmo && dcolor_c && dsigma_c && active_idx && n_active, "atmonr_composite_bwd_compact",
              "null argument");
  ATM_REQUIRE(!color_surf || trans_surf, "atmonr_composite_bwd_compact", "trans_surf required with a surface");
  ATM_REQUIRE(B * (int64_t)N < ((int64_t)1 << 32), "atmonr_composite_bwd_compact", "B*N must be below 2^32");
  ATM_REQUIRE(K != 4 || ((reinterpret_cast<uintptr_t>(color) | reinterpret_cast<uintptr_t>(dcolor_c)) & 15u) == 0,
              "atmonr_composite_bwd_compact", "color and dcolor_c must be 16-byte aligned");
  const int grid = grid_for(B * 32, 128);
#define CALL(KK, VV)                                                                                               \
  k_composite_bwd<KK, VV, true><<<grid, 128, 0, S(stream)>>>(z, color, sigma, color_surf, color_map_atmo, trans_surf, \
                                                            d_atmo, d_surf, z_scale, B, N, relu, dcolor_c, dsigma_c, \
                                                            dcolor_surf, nullptr, grad_absmax, active_idx, n_active)
  ATM_KV_DISPATCH(K, V, CALL)
#undef CALL
  ATM_CHECK_LAUNCH("atmonr_composite_bwd_compact");
  return 0;
}

int atmonr_band_loss(const float* color_map, const int64_t* band, const float* rad, float max_i, int kind,
                     int64_t B, int K, float grad_scale, float* loss, float* dcolor_map, float* partial,
                     void* stream) {
  ATM_REQUIRE(kind >= 0 && kind <= 5, "atmonr_band_loss", "unknown loss kind");
  ATM_REQUIRE(B > 0 && loss && partial, "atmonr_band_loss", "empty batch or null argument");
  const int grid = grid_for(B, 256, 1024);
  k_band_loss<<<grid, 256, 0, S(stream)>>>(color_map, band, rad, max_i, kind, B, K, grad_scale, dcolor_map, partial);
  ATM_CHECK_LAUNCH("atmonr_band_loss");
  k_loss_finish<<<1, 32, 0, S(stream)>>>(partial, grid, B, loss);
  ATM_CHECK_LAUNCH("atmonr_band_loss");
  return 0;
}

int atmonr_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* param_f16, int64_t n,
                      double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                      double grad_scale, int zero_grad, void* stream) {
  ATM_REQUIRE(step >= 1, "atmonr_adamw_step", "step must be >= 1");
  if (n == 0) return 0;
  // torch/optim/adam.py _single_tensor_adam: the hyper-parameters are python floats (doubles)
  // combined in double precision and only then cast to the parameter dtype.
  AdamConsts c;
  c.decay_mul = (float)(1.0 - lr * weight_decay);
  c.w1 = (float)(1.0 - beta1);
  c.beta2 = (float)beta2;
  c.w2 = (float)(1.0 - beta2);
  c.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  c.neg_step_size = (float)(-(lr / (1.0 - pow(beta1, (double)step))));
  c.eps = (float)eps;
  c.gscale = (float)grad_scale;
  const int grid = grid_for((n + 3) / 4, 256, num_sms() * 16);
  k_adamw<<<grid, 256, 0, S(stream)>>>(param, grad, exp_avg, exp_avg_sq, (__half*)param_f16, n, c, zero_grad);
  ATM_CHECK_LAUNCH("atmonr_adamw_step");
  return 0;
}

int atmonr_extract_sigma(const atmonr_frame_t* f, const atmonr_grid_t* g, const void* table,
                         const atmonr_mlp_t* pm, const void* pos_w, const double* pts, int64_t n,
                         float alt_compress, float* sigma, void* stream) {
  ATM_REQUIRE(f && g && g->n_dims == 3 && g->n_levels == 16, "atmonr_extract_sigma", "bad frame/grid");
  ATM_REQUIRE(is_shape(pm, 32, 1) && pm->n_in == 32, "atmonr_extract_sigma", "pos_mlp must be 32 -> [32] -> 16");
  if (n == 0) return 0;
  const int grid = grid_for((n + kTile - 1) / kTile, 1, num_sms() * 16);
  k_extract_sigma<<<grid, kTile, 0, S(stream)>>>(*f, make_geo_frame(*f), *g, (const __half2*)table, (const __half*)pos_w, pts, n,
                                                alt_compress, sigma);
  ATM_CHECK_LAUNCH("atmonr_extract_sigma");
  return 0;
}

static int pe_config(int C, const int32_t* freqs, int interleaved, PeCfg& cfg) {
  if (!(C >= 1 && C <= 4 && freqs)) return -1;
  int col = 0;
  for (int a = 0; a < 4; ++a) {
    cfg.freqs[a] = a < C ? freqs[a] : 0;
    cfg.col0[a] = col;
    col += 2 * cfg.freqs[a];
  }
  cfg.C = C;
  cfg.width = col;
  cfg.interleaved = interleaved;
  return 0;
}

int atmonr_positional_encoding(const float* pts, int64_t M, int C, const int32_t* freqs, int interleaved,
                               float* out, void* stream) {
  PeCfg cfg;
  ATM_REQUIRE(pe_config(C, freqs, interleaved, cfg) == 0, "atmonr_positional_encoding", "C must be 1..4");
  if (M == 0) return 0;
  k_positional_encoding<float><<<grid_for(M, 256), 256, 0, S(stream)>>>(pts, M, cfg, out);
  ATM_CHECK_LAUNCH("atmonr_positional_encoding");
  return 0;
}

int atmonr_positional_encoding_f64(const double* pts, int64_t M, int C, const int32_t* freqs, int interleaved,
                                   float* out, void* stream) {
  PeCfg cfg;
  ATM_REQUIRE(pe_config(C, freqs, interleaved, cfg) == 0, "atmonr_positional_encoding_f64", "C must be 1..4");
  if (M == 0) return 0;
  k_positional_encoding<double><<<grid_for(M, 256), 256, 0, S(stream)>>>(pts, M, cfg, out);
  ATM_CHECK_LAUNCH("atmonr_positional_encoding_f64");
  return 0;
}

int atmonr_sample_pdf(const float* weights, const float* z_coarse, const float* u, int64_t B, int Nc, int Nf,
                      float* z_sorted, int64_t* inds, void* stream) {
  ATM_REQUIRE(Nc >= 3 && Nf >= 1, "atmonr_sample_pdf", "need Nc >= 3 and Nf >= 1");
  if (B == 0) return 0;
  int P = 1;
  while (P < Nc + Nf) P <<= 1;
  const size_t smem = (2 * (size_t)(Nc - 1) + P) * sizeof(float);
  ATM_REQUIRE(smem <= 48 * 1024, "atmonr_sample_pdf", "Nc + Nf too large");
  k_sample_pdf<<<(unsigned)B, 32, smem, S(stream)>>>(weights, z_coarse, u, B, Nc, Nf, P, z_sorted, inds);
  ATM_CHECK_LAUNCH("atmonr_sample_pdf");
  return 0;
}

#endif  // ATM_PART_BASIC
}  // extern "C"
