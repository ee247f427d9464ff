// nerf_points.cu -- the sample-point side of the NeRF training path (sm_100a), forward AND backward.
//
//   atmonr_nerf_encode      : pipelines/nerf.py:104-135 -- point on the ray, geodetic preprocessing
//                             (harp2.py:372-386), positional encoding of the point (encoders.py:21-27) and of
//                             the ray direction (encoders.py:14-20), written as ONE row [pos | dir] per sample
//   atmonr_nerf_encode_bwd  : d loss / d z of the same chain (the reference keeps the fine sample distances
//                             differentiable: samplers.py:96 detaches only the bin width), i.e. the positional
//                             encoding's derivative, the float64 Jacobian of the geodetic conversion and the
//                             projection on the ray direction, per sample in one pass
//   atmonr_sample_pdf_train : samplers.py:72-101 with the by-products the backward needs (CDF, sort permutation)
//   atmonr_sample_pdf_bwd   : gradient of the sorted distances w.r.t. the coarse weights (through the CDF) and
//                             the coarse distances
//   atmonr_composite_dz     : dL/d(Voronoi widths) -> dL/dz (graphics_utils.py:30-36)
//   atmonr_append_heights   : samplers.py:168-195
// Interface contract: include/atmonr_b200.h.
#include "common.cuh"
#include "nerf_points.cuh"

namespace atm {

// A thread encodes one sample; its row goes through shared memory (pitch = width + 1: no bank conflicts) and
// the warp then writes its 32 rows with row-contiguous 128-byte accesses (a thread writing its own 400-byte
// row touches 32 sectors per store instruction: 3.6 % of the NeRF step went there).
__global__ void __launch_bounds__(128)
k_nerf_encode(atmonr_frame_t f, GeoFrame gf, const float* __restrict__ origin, const float* __restrict__ dir,
              const float* __restrict__ z, int64_t M, int N, NerfEncCfg cfg, int width, float* __restrict__ x, int ldx,
              float* __restrict__ pts_n) {
  extern __shared__ float srows[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pitch = width + 1;
  float* mine = srows + (size_t)(warp * 32) * pitch;
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + warp * 32;     // first sample of this warp
  const int64_t i = i0 + lane;
  if (i < M) {
    const int64_t ray = i / N;
    float pn[3];
    nerf_encode_sample(f, gf, origin + 3 * ray, dir + 3 * ray, z[i], cfg, mine + lane * pitch, pn);
    pts_n[3 * i] = pn[0], pts_n[3 * i + 1] = pn[1], pts_n[3 * i + 2] = pn[2];
  }
  __syncwarp();
  const int rows = (int)min((int64_t)32, M - i0);
  for (int r = 0; r < rows; ++r) {
    float* dst = x + (i0 + r) * (int64_t)ldx;
    for (int c = lane; c < width; c += 32) dst[c] = mine[r * pitch + c];
  }
}

__global__ void __launch_bounds__(128)
k_nerf_encode_bwd(atmonr_frame_t f, GeoFrame gf, const float* __restrict__ origin, const float* __restrict__ dir,
                  const float* __restrict__ z, const float* __restrict__ pts_n, const float* __restrict__ g,
                  int ldg, int64_t M, int N, NerfEncCfg cfg, float* __restrict__ gz) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int64_t ray = i / N;
  gz[i] = nerf_encode_sample_bwd(f, gf, origin + 3 * ray, dir + 3 * ray, z[i], pts_n + 3 * i, g + i * (int64_t)ldg, cfg);
}

// ------------------------------------------------------------------------------------------------
// inverse-CDF sampling with its by-products, and its backward
// ------------------------------------------------------------------------------------------------
// samplers.py:72-101, one warp per ray (same arithmetic as k_sample_pdf of atmonr_b200.cu: torch's CPU
// cumsum accumulates float32 in a wider type and rounds on store). Extra outputs: the CDF the bin search
// ran on (so that a caller can check `inds == searchsorted(cdf, u, right=True)` exactly) and, for every
// position of the sorted output, which input it came from (0..Nc-1: coarse sample, Nc+s: fine sample s).
// Shared memory: cdf[Nc-1], mids[Nc-1], merged[P], src[P]; P = next power of two >= Nc+Nf.
__global__ void k_sample_pdf_train(const float* __restrict__ weights, const float* __restrict__ zc,
                                   const float* __restrict__ u, int64_t B, int Nc, int Nf, int P,
                                   float* __restrict__ zout, int64_t* __restrict__ inds, float* __restrict__ cdf_out,
                                   int32_t* __restrict__ src_out) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x;
  const int64_t ray = blockIdx.x;
  if (ray >= B) return;
  const int nb = Nc - 2, ncdf = Nc - 1;
  float* cdf = sm;
  float* mids = sm + ncdf;
  float* merged = mids + ncdf;
  int* src = reinterpret_cast<int*>(merged + P);
  const float* w = weights + ray * Nc;
  const float* zr = zc + ray * Nc;
  if (lane == 0) {
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
  for (int j = lane; j < P; j += 32) {
    merged[j] = j < Nc ? zr[j] : INFINITY;
    src[j] = j;
  }
  __syncwarp();
  if (cdf_out)
    for (int j = lane; j < ncdf; j += 32) cdf_out[ray * ncdf + j] = cdf[j];
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
          if ((a > b) == up) {
            merged[i] = b, merged[ixj] = a;
            const int t = src[i];
            src[i] = src[ixj], src[ixj] = t;
          }
        }
      }
      __syncwarp();
    }
  for (int j = lane; j < Nc + Nf; j += 32) {
    zout[ray * (Nc + Nf) + j] = merged[j];
    if (src_out) src_out[ray * (Nc + Nf) + j] = src[j];
  }
}

// Backward of the above for g = dL/d(sorted z). With b = below, a = above, den = cdf[a] - cdf[b] (1 where
// it was replaced, then constant), t = (u - cdf[b]) / den and the bin width detached (samplers.py:96):
//   fine = mids[b] + t * width      d fine / d mids[b] = 1
//   d fine / d cdf[b] = width * (-1/den + t/den)      d fine / d cdf[a] = -width * t / den   (den not replaced)
//   d fine / d cdf[b] = -width / den                                                          (den replaced)
// cdf[1+j] = sum_{i<=j} pdf_i, pdf_j = (w_j + 1e-8) / S: dL/dw_j = (G_j - sum_i G_i pdf_i) / S with
// G_j = sum_{i>=j} dL/dcdf[1+i]. mids[j] = (zc[j] + zc[j+1]) / 2; the coarse z also reach the output directly.
// Shared memory: gsrc[Nc+Nf], gcdf[Nc-1], gmid[Nc-1], pdf[Nc-2].
__global__ void k_sample_pdf_bwd(const float* __restrict__ gz_sorted, const int32_t* __restrict__ src,
                                 const float* __restrict__ weights, const float* __restrict__ zc,
                                 const float* __restrict__ u, const float* __restrict__ cdf_in,
                                 const int64_t* __restrict__ inds, int64_t B, int Nc, int Nf,
                                 float* __restrict__ d_weights, float* __restrict__ d_zc) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x;
  const int64_t ray = blockIdx.x;
  if (ray >= B) return;
  const int nb = Nc - 2, ncdf = Nc - 1, T = Nc + Nf;
  float* gsrc = sm;
  float* gcdf = gsrc + T;
  float* gmid = gcdf + ncdf;
  float* pdf = gmid + ncdf;
  const float* cdf = cdf_in + ray * ncdf;
  const float* zr = zc + ray * Nc;
  for (int j = lane; j < T; j += 32) gsrc[src[ray * T + j]] = gz_sorted[ray * T + j];
  for (int j = lane; j < ncdf; j += 32) gcdf[j] = 0.0f, gmid[j] = 0.0f;
  __syncwarp();
  for (int s = lane; s < Nf; s += 32) {
    const float gg = gsrc[Nc + s];
    const float uu = u[ray * Nf + s];
    const int ind = (int)inds[ray * Nf + s];
    const int below = max(ind - 1, 0), above = min(ind, ncdf - 1);
    const float cb = cdf[below], ca = cdf[above];
    float den = ca - cb;
    const bool replaced = den < 1e-8f;
    if (replaced) den = 1.0f;
    const float t = (uu - cb) / den;
    const float width = 0.5f * (zr[above + 1] + zr[above]) - 0.5f * (zr[below + 1] + zr[below]);
    const float gt = gg * width;
    atomicAdd(&gmid[below], gg);
    float g_cb = -gt / den;
    if (!replaced) {
      const float g_den = -gt * t / den;
      atomicAdd(&gcdf[above], g_den);
      g_cb -= g_den;
    }
    atomicAdd(&gcdf[below], g_cb);
  }
  __syncwarp();
  if (d_zc)
    for (int j = lane; j < Nc; j += 32) {
      float v = gsrc[j];
      if (j < ncdf) v += 0.5f * gmid[j];
      if (j >= 1) v += 0.5f * gmid[j - 1];
      d_zc[ray * Nc + j] = v;
    }
  __syncwarp();
  if (d_weights) {
    const float* w = weights + ray * Nc;
    if (lane == 0) {
      double tot_d = 0.0;
      for (int j = 0; j < nb; ++j) tot_d += (double)(w[1 + j] + 1e-8f);
      const float tot = (float)tot_d;
      float run = 0.0f, dot = 0.0f;
      for (int j = nb - 1; j >= 0; --j) {  // G_j = reverse cumulative sum of gcdf[1..]
        run += gcdf[1 + j];
        const float pj = (w[1 + j] + 1e-8f) / tot;
        pdf[j] = run;
        dot += run * pj;
      }
      gcdf[0] = dot;   // (cdf[0] is the constant 0: its slot carries the dot product to the other lanes)
      gmid[0] = tot;
    }
    __syncwarp();
    const float dot = gcdf[0], tot = gmid[0];
    for (int j = lane; j < Nc; j += 32)
      d_weights[ray * Nc + j] = (j >= 1 && j <= nb) ? (pdf[j - 1] - dot) / tot : 0.0f;
  }
}

// delta_i = hi_i - lo_i with hi_i = (z_i + z_{i+1}) / 2 (last: z_{N-1}) and lo_i = (z_{i-1} + z_i) / 2
// (first: 0), in km: dz_j = zs * (dd_j ([j<N-1] ? 1/2 : 1) - dd_j [j>0] / 2 + dd_{j-1} / 2 - dd_{j+1} / 2)
__global__ void k_composite_dz(const float* __restrict__ ddelta, int64_t B, int N, float zs, float* __restrict__ dz) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int j = (int)(i % N);
  const float dd = ddelta[i];
  float v = dd * (j < N - 1 ? 0.5f : 1.0f);
  if (j > 0) v += 0.5f * ddelta[i - 1] - 0.5f * dd;
  if (j < N - 1) v -= 0.5f * ddelta[i + 1];
  dz[i] = v * zs;
}

// samplers.py:168-195: ellipsoidal height of pts * scale + offset over ray_origin_height, appended as a
// fourth coordinate (float64 Bowring step, rounded once to float32)
__global__ void k_append_heights(const float* __restrict__ pts, int64_t M, double scale, double ox, double oy,
                                 double oz, double origin_height, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float p0 = pts[3 * i], p1 = pts[3 * i + 1], p2 = pts[3 * i + 2];
  double lat, lon, alt;
  ecef_to_geodetic((double)p0 * scale + ox, (double)p1 * scale + oy, (double)p2 * scale + oz, lat, lon, alt);
  reinterpret_cast<float4*>(out)[i] = make_float4(p0, p1, p2, (float)(alt / origin_height));
}

}  // namespace atm

using namespace atm;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int atmonr_nerf_encode(const atmonr_frame_t* f, const float* origin, const float* dir, const float* z, int64_t B,
                       int N, const int32_t* pos_freqs, int dir_freqs, float* x, int ldx, float* pts_n,
                       void* stream) {
  NerfEncCfg c;
  ATM_REQUIRE(f && origin && dir && z && x && pts_n, "atmonr_nerf_encode", "null argument");
  ATM_REQUIRE(nerf_enc_cfg(pos_freqs, dir_freqs, c) == 0, "atmonr_nerf_encode", "bad frequency counts");
  ATM_REQUIRE(ldx >= c.pos_width + 6 * dir_freqs, "atmonr_nerf_encode", "row stride smaller than the encoding");
  if (B * N == 0) return 0;
  const int width = c.pos_width + 6 * dir_freqs;
  const size_t smem = (size_t)128 * (width + 1) * sizeof(float);
  ATM_REQUIRE(smem <= 200 * 1024, "atmonr_nerf_encode", "encoding too wide");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_nerf_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("atmonr_nerf_encode", cudaGetErrorString(e));
  }
  k_nerf_encode<<<grid_for(B * N, 128), 128, smem, S(stream)>>>(*f, make_geo_frame(*f), origin, dir, z, B * N, N, c, width,
                                                                x, ldx, pts_n);
  ATM_CHECK_LAUNCH("atmonr_nerf_encode");
  return 0;
}

int atmonr_nerf_encode_bwd(const atmonr_frame_t* f, const float* origin, const float* dir, const float* z,
                           const float* pts_n, const float* g_x, int ldg, int64_t B, int N,
                           const int32_t* pos_freqs, float* g_z, void* stream) {
  NerfEncCfg c;
  ATM_REQUIRE(f && origin && dir && z && pts_n && g_x && g_z, "atmonr_nerf_encode_bwd", "null argument");
  ATM_REQUIRE(nerf_enc_cfg(pos_freqs, 0, c) == 0, "atmonr_nerf_encode_bwd", "bad frequency counts");
  ATM_REQUIRE(ldg >= c.pos_width, "atmonr_nerf_encode_bwd", "row stride smaller than the encoding");
  if (B * N == 0) return 0;
  k_nerf_encode_bwd<<<grid_for(B * N, 128), 128, 0, S(stream)>>>(*f, make_geo_frame(*f), origin, dir, z, pts_n, g_x,
                                                                 ldg, B * N, N, c, g_z);
  ATM_CHECK_LAUNCH("atmonr_nerf_encode_bwd");
  return 0;
}

int atmonr_sample_pdf_train(const float* weights, const float* z_coarse, const float* u, int64_t B, int Nc, int Nf,
                            float* z_sorted, int64_t* inds, float* cdf, int32_t* src, void* stream) {
  ATM_REQUIRE(Nc >= 3 && Nf >= 1, "atmonr_sample_pdf_train", "need Nc >= 3 and Nf >= 1");
  ATM_REQUIRE(weights && z_coarse && u && z_sorted, "atmonr_sample_pdf_train", "null argument");
  if (B == 0) return 0;
  int P = 1;
  while (P < Nc + Nf) P <<= 1;
  const size_t smem = (2 * (size_t)(Nc - 1) + 2 * (size_t)P) * sizeof(float);
  ATM_REQUIRE(smem <= 48 * 1024, "atmonr_sample_pdf_train", "Nc + Nf too large");
  k_sample_pdf_train<<<(unsigned)B, 32, smem, S(stream)>>>(weights, z_coarse, u, B, Nc, Nf, P, z_sorted, inds, cdf, src);
  ATM_CHECK_LAUNCH("atmonr_sample_pdf_train");
  return 0;
}

int atmonr_sample_pdf_bwd(const float* g_z_sorted, const int32_t* src, const float* weights, const float* z_coarse,
                          const float* u, const float* cdf, const int64_t* inds, int64_t B, int Nc, int Nf,
                          float* d_weights, float* d_z_coarse, void* stream) {
  ATM_REQUIRE(Nc >= 3 && Nf >= 1, "atmonr_sample_pdf_bwd", "need Nc >= 3 and Nf >= 1");
  ATM_REQUIRE(g_z_sorted && src && weights && z_coarse && u && cdf && inds, "atmonr_sample_pdf_bwd", "null argument");
  if (B == 0) return 0;
  const size_t smem = ((size_t)(Nc + Nf) + 2 * (size_t)(Nc - 1) + (size_t)(Nc - 2)) * sizeof(float);
  ATM_REQUIRE(smem <= 48 * 1024, "atmonr_sample_pdf_bwd", "Nc + Nf too large");
  k_sample_pdf_bwd<<<(unsigned)B, 32, smem, S(stream)>>>(g_z_sorted, src, weights, z_coarse, u, cdf, inds, B, Nc, Nf,
                                                        d_weights, d_z_coarse);
  ATM_CHECK_LAUNCH("atmonr_sample_pdf_bwd");
  return 0;
}

int atmonr_composite_dz(const float* ddelta, int64_t B, int N, float z_scale, float* dz, void* stream) {
  ATM_REQUIRE(ddelta && dz && N >= 1, "atmonr_composite_dz", "null argument");
  if (B == 0) return 0;
  k_composite_dz<<<grid_for(B * N, 256), 256, 0, S(stream)>>>(ddelta, B, N, z_scale, dz);
  ATM_CHECK_LAUNCH("atmonr_composite_dz");
  return 0;
}

int atmonr_append_heights(const float* pts, int64_t M, double scale, const double* offset_host,
                          double ray_origin_height, float* out, void* stream) {
  ATM_REQUIRE(pts && out && offset_host, "atmonr_append_heights", "null argument");
  ATM_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "atmonr_append_heights", "out must be 16-byte aligned");
  if (M == 0) return 0;
  k_append_heights<<<grid_for(M, 256), 256, 0, S(stream)>>>(pts, M, scale, offset_host[0], offset_host[1],
                                                           offset_host[2], ray_origin_height, out);
  ATM_CHECK_LAUNCH("atmonr_append_heights");
  return 0;
}

}  // extern "C"
