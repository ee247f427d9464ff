/*
 * atmonr_b200.h -- C ABI of libatmonr_b200.so, the sm_100a implementation of the AtmoNR
 * training / extraction hot path (nasa/atmospheric-neural-rendering).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = ok, negative = error; atmonr_last_error() gives the message
 *     (thread-local);
 *   - no allocation happens inside the library, no torch types cross this boundary;
 *   - the reference has no native interface: each entry point names the Python function or
 *     third-party (tiny-cuda-nn) module of the reference it replaces (file:line relative to
 *     the reference tree).
 */
#ifndef ATMONR_B200_H
#define ATMONR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATMONR_MAX_LEVELS 16
#define ATMONR_ABI_VERSION 1

/* Level table of a multiresolution hash grid (tiny-cuda-nn GridEncoding, used at
 * src/atmonr/pipelines/instant_ngp.py:60-63 and :78-80). Filled by atmonr_grid_layout. */
typedef struct {
  int32_t n_dims;    /* 2 or 3 */
  int32_t n_levels;  /* <= ATMONR_MAX_LEVELS */
  int32_t n_feat;    /* features per entry; only 2 is supported */
  int32_t reserved;
  float scale[ATMONR_MAX_LEVELS];     /* exp2f(l*log2f(per_level_scale))*base - 1 */
  uint32_t res[ATMONR_MAX_LEVELS];    /* ceilf(scale)+1 */
  uint32_t size[ATMONR_MAX_LEVELS];   /* entries in the level */
  uint32_t offset[ATMONR_MAX_LEVELS + 1]; /* first entry of the level; [n_levels] = total */
} atmonr_grid_t;

/* Constants of the 'horizontal' point preprocessor closure,
 * src/atmonr/datasets/harp2.py:351-388. enabled == 0 means no preprocessor. */
typedef struct {
  double scale;      /* metres per normalised unit (dataset.scale) */
  double offset[3];  /* dataset.offset, float64 */
  double lat_min, lat_range, lon_min, lon_range; /* float32 granule stats, up-cast */
  double origin_height;                          /* config ray_origin_height */
  int32_t shift_lon;                             /* granule crosses the dateline */
  int32_t enabled;
} atmonr_frame_t;

/* Bias-free fully fused MLP (tiny-cuda-nn FullyFusedMLP behind tcnn.Network,
 * instant_ngp.py:64-68,73-77,81-85). Weights: fp16 copies of the row-major [out][in]
 * matrices, concatenated first to last; the input is padded to in_pad with 1.0. */
typedef struct {
  int32_t n_in;      /* logical input width  */
  int32_t in_pad;    /* multiple of 16       */
  int32_t width;     /* hidden width (32)    */
  int32_t n_hidden;  /* hidden layers (1|2)  */
  int32_t n_out;     /* logical output width */
  int32_t out_pad;   /* multiple of 16 (16)  */
} atmonr_mlp_t;

int atmonr_abi_version(void);
const char* atmonr_last_error(void);

/* Optional tuning call (no reference counterpart): keep [ptr, ptr + bytes) -- the fp16 hash table the
 * field kernels gather from -- in the L2's persisting set-aside for kernels launched on `stream`
 * afterwards (cudaAccessPolicyWindow; hit_ratio in (0, 1], scaled down when the set-aside is smaller
 * than the range). bytes == 0 clears the window. */
int atmonr_l2_persist(const void* ptr, size_t bytes, float hit_ratio, void* stream);

/* Host only (no GPU needed): level table of tcnn's GridEncoding. */
int atmonr_grid_layout(int n_dims, int n_levels, int log2_hashmap_size, int base_resolution,
                       float per_level_scale, atmonr_grid_t* out_host);

/* ---- geospatial/wgs_84.py:223-290 get_rays (once per run, per chunk of pixels) ----------------
 * The ray table of a granule: for each of the n = P*A pixel/view pairs (lat, lon, alt, thetav,
 * phiv: float32, degrees / metres) the surface point, the direction from the top of the shell to
 * the surface, and the length for which the ray's upper end lies at ray_origin_height above the
 * ellipsoid, found by the reference's fixed point `len *= H / height(len)`: every ray of the call
 * is refined for as long as ANY ray of the call is further than tol metres from the shell, at
 * most max_iters times (so a call = one chunk of datasets/harp2.py:219-239). Outputs origin (n,3),
 * dir (n,3), len (n) float32. work: (2n + 1) * 8 bytes of device scratch. The iteration count
 * goes to *n_iters_host (may be NULL). NaN inputs give NaN rays (filtered by the caller,
 * wgs_84.py:293-313). UNLIKE the other calls this one synchronises `stream` (it reads the
 * convergence flag after every refinement). */
int atmonr_get_rays(const float* lat, const float* lon, const float* alt, const float* thetav,
                    const float* phiv, int64_t n, float ray_origin_height, double tol,
                    int max_iters, float* origin, float* dir, float* len, void* work,
                    int* n_iters_host, void* stream);

/* ---- geospatial/wgs_84.py:293-313 filter_rays (once per run) ----------------------------------
 * valid[i] = 1 when none of origin[i] (3), dir[i] (3), rad[i] is NaN, else 0; `valid` is one byte per
 * ray (the storage of a torch.bool tensor). n = 0 is a no-op. */
int atmonr_filter_rays(const float* origin, const float* dir, const float* rad, int64_t n,
                       uint8_t* valid, void* stream);

/* ---- geospatial/wgs_84.py:316-339 normalize_rays (once per run) --------------------------------
 * First half (:333-335): the bounding box of the ray origins and of the rays' lower ends
 * `origin + dir * len` (float32, product and sum rounded separately, as the eager reference does).
 * hi_lo (device, 6 floats) receives max x,y,z then min x,y,z; an axis that holds a NaN gives NaN like
 * torch.max / torch.min. n must be > 0. work: ATMONR_RAY_EXTENT_WORK_BYTES of device scratch.
 * The caller forms `scale = max(hi - lo) / 2` and `offset = (hi + lo) / 2` in float64 (:336-337: a
 * Python float and a float64[3] device tensor), then calls the second half (:338):
 * out[i][k] = float(clamp((double(origin[i][k]) - offset[k]) / scale, -1, 1)); offset is a DEVICE
 * pointer to three doubles. */
#define ATMONR_RAY_EXTENT_WORK_BYTES 37888
int atmonr_ray_extent(const float* origin, const float* dir, const float* len, int64_t n,
                      float* hi_lo, void* work, void* stream);
int atmonr_normalize_origins(const float* origin, int64_t n, const double* offset, double scale,
                             float* out, void* stream);

/* ---- datasets/harp2.py:392-420 __getitem__ / __getbatch__ (every step) ----------------------
 * The seven advanced-index gathers of a batch in one launch. Tables of R rays: origin (R,3),
 * dir (R,3), alt (R), rad (R), len (R) float32, ray_idx (R) int32, band (R) int64 (irgb_idx);
 * index (B) int64, negative values count from the end. alt / ray_idx and their outputs may be
 * NULL. *bad_index (device int, zeroed by the caller) is set when an index is out of range (the
 * entry is then read from ray 0). */
int atmonr_gather_batch(const float* origin, const float* dir, const float* alt, const float* rad,
                        const float* len, const int32_t* ray_idx, const int64_t* band,
                        const int64_t* index, int64_t B, int64_t R, float* out_origin,
                        float* out_dir, float* out_alt, float* out_rad, float* out_len,
                        int32_t* out_ray_idx, int64_t* out_band, int* bad_index, void* stream);

/* ---- samplers.py:8-47 sample_uniform_bins ------------------------------------------------
 * mode 0: bin mid-points (random=False); 1: uniforms read from u (B,N); 2: in-kernel Philox
 * keyed by (seed, ray_index_base + ray, bin). bins: N floats = linspace(0,1,N+1)[:-1] or NULL
 * (then i/N is used). pts (B,N,3), z (B,N). */
int atmonr_sample_uniform(const float* origin, const float* dir, const float* len,
                          const float* u, const float* bins, int64_t B, int N, int mode,
                          uint64_t seed, uint64_t ray_index_base, float* pts, float* z,
                          void* stream);

/* ---- harp2.py:372-386 preprocess_coords (+ wgs_84.py:56-97) ------------------------------
 * is_f64 selects float64 in/out (extract path) or float32 in/out (training path). n points. */
int atmonr_preprocess_horizontal(const atmonr_frame_t* frame_host, const void* pts, void* out,
                                 int64_t n, int is_f64, void* stream);

/* ---- instant_ngp.py:139-160 fused: sampler -> preprocessor -> (p+1)/2 -> z /= alt_compress
 * x01 (B*N,3) float32 is the hash-grid input; z (B,N) are the normalised sample distances. */
int atmonr_ngp_sample_points(const atmonr_frame_t* frame_host, const float* origin,
                             const float* dir, const float* len, const float* u,
                             const float* bins, int64_t B, int N, int mode, uint64_t seed,
                             uint64_t ray_index_base, float alt_compress, float* x01, float* z,
                             void* stream);
/* The same with `include_height` (instant_ngp.py:155-156 -> samplers.py:168-195): x01 (B*N,4), the fourth
 * column = ellipsoidal height of the [0,1]^3 point (before the altitude compression) * scale + offset
 * over ray_origin_height. offset_host: 3 doubles on the host. */
int atmonr_ngp_sample_points_height(const atmonr_frame_t* frame_host, const float* origin,
                                    const float* dir, const float* len, const float* u,
                                    const float* bins, int64_t B, int N, int mode, uint64_t seed,
                                    uint64_t ray_index_base, float alt_compress, double scale,
                                    const double* offset_host, double ray_origin_height, float* x01,
                                    float* z, void* stream);

/* ---- tcnn.Encoding HashGrid forward/backward (instant_ngp.py:163,236) --------------------
 * x (M, x_stride) float32, the first n_dims columns are used. table: fp16 shadow of the
 * parameters, (entries, 2). out (M, 2*n_levels) float32 holding fp16-rounded values.
 * bwd accumulates dL/dtable into a float32 (entries, 2) buffer with vector atomics. */
int atmonr_hashgrid_fwd(const atmonr_grid_t* grid_host, const float* x, int x_stride,
                        const void* table_f16, int64_t M, float* out, void* stream);
int atmonr_hashgrid_bwd(const atmonr_grid_t* grid_host, const float* x, int x_stride,
                        const float* dout, int64_t M, float* dtable, void* stream);
/* Debug/parity: uint32 entry index of every (sample, level, corner): (M, L, 2^D). */
int atmonr_hashgrid_indices(const atmonr_grid_t* grid_host, const float* x, int x_stride,
                            int64_t M, uint32_t* idx, void* stream);

/* ---- tcnn.Network forward/backward -------------------------------------------------------
 * x (M, n_in) float32; out (M, n_out) float32. bwd recomputes the activations; dx may be
 * NULL; dw (float32, same layout as the weights) is accumulated into. */
int atmonr_mlp_fwd(const atmonr_mlp_t* mlp_host, const void* w_f16, const float* x, int64_t M,
                   float* out, void* stream);
int atmonr_mlp_bwd(const atmonr_mlp_t* mlp_host, const void* w_f16, const float* x,
                   const float* dout, int64_t M, float* dx, float* dw, void* stream);

/* ---- fused radiance field: hash grid -> pos_mlp -> [SH2(dir) | feat] -> dir_mlp ------------
 * instant_ngp.py:163-171,178 per sample. x01 (M,3); dirs (B,3) with M = B*N; outputs are the
 * RAW (pre-ReLU) density sigma (M) and colour (M,4).  n_sigma == 1, num_bands == 4. */
int atmonr_ngp_field_fwd(const atmonr_grid_t* grid_host, const void* table_f16,
                         const atmonr_mlp_t* pos_mlp_host, const void* pos_w_f16,
                         const atmonr_mlp_t* dir_mlp_host, const void* dir_w_f16,
                         const float* x01, const float* dirs, int64_t B, int N, float* sigma_raw,
                         float* color_raw, void* stream);
int atmonr_ngp_field_bwd(const atmonr_grid_t* grid_host, const void* table_f16,
                         const atmonr_mlp_t* pos_mlp_host, const void* pos_w_f16,
                         const atmonr_mlp_t* dir_mlp_host, const void* dir_w_f16,
                         const float* x01, const float* dirs, const float* dsigma_raw,
                         const float* dcolor_raw, int64_t B, int N, float* dtable, float* dpos_w,
                         float* ddir_w, void* stream);

/* tcgen05 implementations of the two calls above (same contract; dense layers on the 5th-gen
 * tensor cores, accumulators in TMEM). enc (optional, (M,32) fp16): the forward stores the encoded
 * features there, the backward reads them instead of re-gathering the table. grad_absmax: device
 * scalar holding max|dsigma_raw|,|dcolor_raw| (as produced by atmonr_composite_bwd); it sets the
 * power-of-two scale under which the fp16 gradient operands are formed. */
int atmonr_ngp_field_fwd_tc(const atmonr_grid_t* grid_host, const void* table_f16,
                            const atmonr_mlp_t* pos_mlp_host, const void* pos_w_f16,
                            const atmonr_mlp_t* dir_mlp_host, const void* dir_w_f16,
                            const float* x01, const float* dirs, int64_t B, int N,
                            float* sigma_raw, float* color_raw, void* enc_f16, void* stream);
int atmonr_ngp_field_bwd_tc(const atmonr_grid_t* grid_host, const void* table_f16,
                            const atmonr_mlp_t* pos_mlp_host, const void* pos_w_f16,
                            const atmonr_mlp_t* dir_mlp_host, const void* dir_w_f16,
                            const float* x01, const float* dirs, const void* enc_f16,
                            const float* dsigma_raw, const float* dcolor_raw,
                            const float* grad_absmax, int64_t B, int N, float* dtable,
                            float* dpos_w, float* ddir_w, void* stream);

/* The tcgen05 backward over a LIST of samples (those whose incoming gradient can be non-zero; see
 * atmonr_composite_bwd_compact): active_idx (n,) sample indices, *n_active = n on the device,
 * dsigma_c (n,) / dcolor_c (n,4) indexed by list position. A sample with zero incoming gradient
 * contributes exactly zero to dtable / dpos_w / ddir_w, so the outputs equal those of
 * atmonr_ngp_field_bwd_tc on the dense arrays (up to the order of the fp32 atomics). enc_f16 is
 * required. */
int atmonr_ngp_field_bwd_tc_compact(const atmonr_grid_t* grid_host, const atmonr_mlp_t* pos_mlp_host,
                                    const void* pos_w_f16, const atmonr_mlp_t* dir_mlp_host,
                                    const void* dir_w_f16, const float* x01, const float* dirs,
                                    const void* enc_f16, const uint32_t* active_idx,
                                    const uint32_t* n_active, const float* dsigma_c,
                                    const float* dcolor_c, const float* grad_absmax, int64_t B, int N,
                                    float* dtable, float* dpos_w, float* ddir_w, void* stream);

/* ---- surface branch, per ray: [hash2d(pts_surf.xy) | SH2(dir)] -> surf_mlp ----------------
 * instant_ngp.py:140,150,173-174. color_surf_raw (B,4) pre-ReLU. */
int atmonr_ngp_surface_fwd(const atmonr_grid_t* grid2d_host, const void* table_f16,
                           const atmonr_mlp_t* mlp_host, const void* w_f16, const float* origin,
                           const float* dir, const float* len, int64_t B, float* color_surf_raw,
                           void* stream);
int atmonr_ngp_surface_bwd(const atmonr_grid_t* grid2d_host, const void* table_f16,
                           const atmonr_mlp_t* mlp_host, const void* w_f16, const float* origin,
                           const float* dir, const float* len, const float* dcolor_surf_raw,
                           int64_t B, float* dtable, float* dw, void* stream);

/* ---- graphics_utils.py:6-77 render / render_with_surface (+ the ReLUs of
 * instant_ngp.py:178-184 when relu != 0) -------------------------------------------------
 * z (B,N) normalised distances, multiplied by z_scale (= scale/1000, km); color (B,N,K);
 * sigma (B,N,V) with V == 1 or V == K; color_surf (B,K) or NULL. Outputs: color_map,
 * color_map_atmo, color_map_surf (B,K); trans_surf (B,V) = prod(1-alpha); optional weights
 * and alpha (B,N,V). K <= 4. */
int atmonr_composite_fwd(const float* z, const float* color, const float* sigma,
                         const float* color_surf, float z_scale, int64_t B, int N, int K, int V,
                         int relu, float* color_map, float* color_map_atmo,
                         float* color_map_surf, float* trans_surf, float* weights, float* alpha,
                         void* stream);
/* Backward given dL/dcolor_map_atmo (B,K) and dL/dcolor_map_surf (B,K). Produces dcolor
 * (B,N,K), dsigma (B,N,V) and dcolor_surf (B,K) w.r.t. the RAW inputs when relu != 0, and
 * optionally ddelta (B,N) = dL/d(Voronoi cell width in km) for callers that differentiate
 * through the sample distances (NeRF fine pass, samplers.py:96 keeps that path alive), and
 * optionally grad_absmax (1 float, zero-initialised by the caller) = max |dcolor|, |dsigma|. */
int atmonr_composite_bwd(const float* z, const float* color, const float* sigma,
                         const float* color_surf, const float* color_map_atmo,
                         const float* trans_surf, const float* d_atmo, const float* d_surf,
                         float z_scale, int64_t B, int N, int K, int V, int relu, float* dcolor,
                         float* dsigma, float* dcolor_surf, float* ddelta, float* grad_absmax,
                         void* stream);

/* atmonr_composite_bwd for callers that ALSO differentiate the per-sample weights it returned
 * (NeRF coarse pass: the weights define the fine sampler's CDF, samplers.py:72-74): weights (B,N,V) as
 * written by atmonr_composite_fwd, d_weights (B,N,V) = dL/dweights; their contribution is added to
 * dsigma / ddelta (the weights do not depend on colour). */
int atmonr_composite_bwd_weights(const float* z, const float* color, const float* sigma,
                                 const float* color_surf, const float* color_map_atmo,
                                 const float* trans_surf, const float* weights, const float* d_atmo,
                                 const float* d_surf, const float* d_weights, float z_scale,
                                 int64_t B, int N, int K, int V, int relu, float* dcolor,
                                 float* dsigma, float* dcolor_surf, float* ddelta, void* stream);

/* The same backward, writing dcolor / dsigma only for the samples that can carry a gradient, as a
 * list: active_idx (capacity B*N) receives their sample indices (a ray's samples stay consecutive
 * and ordered), dcolor_c (capacity B*N, K) and dsigma_c (capacity B*N, V) their gradients by list
 * position, *n_active (device, zeroed by the caller) the list length. With relu != 0 a sample is
 * listed iff one of its raw densities is > 0 (otherwise alpha = 0 => weight 0 => dcolor = 0, and the
 * ReLU zeroes dsigma); with relu == 0 every sample is listed. */
int atmonr_composite_bwd_compact(const float* z, const float* color, const float* sigma,
                                 const float* color_surf, const float* color_map_atmo,
                                 const float* trans_surf, const float* d_atmo, const float* d_surf,
                                 float z_scale, int64_t B, int N, int K, int V, int relu,
                                 uint32_t* active_idx, uint32_t* n_active, float* dcolor_c,
                                 float* dsigma_c, float* dcolor_surf, float* grad_absmax,
                                 void* stream);

/* ---- per-band loss and its gradient (instant_ngp.py:249-263, losses.py:5-33) ---------------
 * kind: 0 dark, 1 hdr, 2 l1, 3 l1_plus_hdr, 4 mse, 5 mse_plus_hdr. color_map (B,K); band
 * (B) int64; rad (B). Writes loss (1 float, mean over B) and dcolor_map (B,K) scaled by
 * grad_scale. partial: scratch of at least 1024 floats. */
int atmonr_band_loss(const float* color_map, const int64_t* band, const float* rad, float max_i,
                     int kind, int64_t B, int K, float grad_scale, float* loss,
                     float* dcolor_map, float* partial, void* stream);

/* ---- fused AdamW (torch.optim.AdamW as configured at instant_ngp.py:107-127) --------------
 * Dense update of n parameters; grad is multiplied by grad_scale first; optionally writes the
 * fp16 shadow and zeroes the gradient. step is the 1-based step count. */
int atmonr_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                      void* param_f16, int64_t n, double lr, double beta1, double beta2, double eps,
                      double weight_decay, int64_t step, double grad_scale, int zero_grad,
                      void* stream);

/* ---- extract (instant_ngp.py:208-247; loop at scripts/extract.py:203-209) -----------------
 * pts (n,3) float64 normalised scene coordinates -> sigma (n) float32 = max(pos_mlp[...,0],0). */
int atmonr_extract_sigma(const atmonr_frame_t* frame_host, const atmonr_grid_t* grid_host,
                         const void* table_f16, const atmonr_mlp_t* pos_mlp_host,
                         const void* pos_w_f16, const double* pts, int64_t n, float alt_compress,
                         float* sigma, void* stream);
/* The same query with the two dense layers on tcgen05 (the training forward kernel without its
 * colour branch); same arguments, same results up to the accumulation order of the dense layers. */
int atmonr_extract_sigma_tc(const atmonr_frame_t* frame_host, const atmonr_grid_t* grid_host,
                         const void* table_f16, const atmonr_mlp_t* pos_mlp_host,
                         const void* pos_w_f16, const double* pts, int64_t n, float alt_compress,
                         float* sigma, void* stream);

/* ---- NeRF path helpers ----------------------------------------------------------------------
 * encoders.py:4-28 positional_encoding; list variant (per-axis frequency counts, layout
 * [sin x L | cos x L] per axis) when interleaved == 0, int variant ([sin,cos] per frequency)
 * otherwise. pts (M,C) float32, C <= 4. */
int atmonr_positional_encoding(const float* pts, int64_t M, int C, const int32_t* freqs_host,
                               int interleaved, float* out, void* stream);
/* The same for float64 points (the extract path, pipelines/nerf.py:209-213: scripts/extract.py feeds
 * float64 points and the reference encodes them before casting): phases and sin / cos in float64,
 * output rounded once to float32. */
int atmonr_positional_encoding_f64(const double* pts, int64_t M, int C, const int32_t* freqs_host,
                                   int interleaved, float* out, void* stream);
/* samplers.py:50-103 sample_pdf up to and including the sort: weights (B,Nc) (the V == 1
 * column), z_coarse (B,Nc), u (B,Nf) -> z_sorted (B,Nc+Nf), inds (B,Nf) int64. */
int atmonr_sample_pdf(const float* weights, const float* z_coarse, const float* u, int64_t B,
                      int Nc, int Nf, float* z_sorted, int64_t* inds, void* stream);
/* The same with the by-products the training path needs: cdf (B,Nc-1) = the CDF the bin search ran on
 * (inds == searchsorted(cdf, u, right=True) exactly), src (B,Nc+Nf) int32 = for every position of
 * z_sorted the input it came from (0..Nc-1: coarse sample, Nc+s: fine sample s). Either may be NULL. */
int atmonr_sample_pdf_train(const float* weights, const float* z_coarse, const float* u, int64_t B,
                            int Nc, int Nf, float* z_sorted, int64_t* inds, float* cdf,
                            int32_t* src, void* stream);
/* Backward of sample_pdf as the reference's autograd graph defines it (samplers.py:72-101; only the
 * bin width is detached, :96): g_z_sorted (B,Nc+Nf) -> d_weights (B,Nc) (zero in the first and last
 * column, which samplers.py:72 drops) and d_z_coarse (B,Nc); either output may be NULL. */
int atmonr_sample_pdf_bwd(const float* g_z_sorted, const int32_t* src, const float* weights,
                          const float* z_coarse, const float* u, const float* cdf,
                          const int64_t* inds, int64_t B, int Nc, int Nf, float* d_weights,
                          float* d_z_coarse, void* stream);
/* pipelines/nerf.py:104-135 for one network pass, per sample i of ray r = i / N:
 *   p = origin[r] + dir[r] * z[i]  ->  harp2.py:372-386 `horizontal` preprocessing when frame->enabled
 *   x[i] = [ positional_encoding(p, pos_freqs) (list layout, encoders.py:21-27)
 *          | positional_encoding(dir[r], dir_freqs) (int layout, encoders.py:14-20) ],  row stride ldx
 *   pts_n[i] = the preprocessed point (B*N,3).
 * The direction encoding is evaluated per sample from the ray's direction (nothing of shape (B,N,3)
 * is repeated or concatenated). */
int atmonr_nerf_encode(const atmonr_frame_t* frame, const float* origin, const float* dir,
                       const float* z, int64_t B, int N, const int32_t* pos_freqs_host, int dir_freqs,
                       float* x, int ldx, float* pts_n, void* stream);
/* dL/dz (B*N) of the chain above given g_x = dL/dx (row stride ldg; the first 2*sum(pos_freqs)
 * columns are read): derivative of the positional encoding, Jacobian of the geodetic conversion in
 * float64 (wgs_84.py:56-97 differentiated, clip mask of harp2.py:386 included) and projection on the
 * ray direction. The reference obtains the same numbers from autograd through its float64 torch
 * expressions. */
int atmonr_nerf_encode_bwd(const atmonr_frame_t* frame, const float* origin, const float* dir,
                           const float* z, const float* pts_n, const float* g_x, int ldg, int64_t B,
                           int N, const int32_t* pos_freqs_host, float* g_z, void* stream);
/* dL/d(Voronoi cell widths, km) as written by atmonr_composite_bwd -> dL/dz (graphics_utils.py:30-36
 * differentiated; z_scale = km per unit of z). */
int atmonr_composite_dz(const float* ddelta, int64_t B, int N, float z_scale, float* dz, void* stream);
/* samplers.py:168-195 append_heights: out (M,4) = [pts | height(pts * scale + offset) / ray_origin_height]
 * with the float64 Bowring step of wgs_84.py:56-97. offset_host: 3 doubles on the host. */
int atmonr_append_heights(const float* pts, int64_t M, double scale, const double* offset_host,
                          double ray_origin_height, float* out, void* stream);

/* ---- AtmoNeRF dense layers (models/nerf.py:6-93: eleven biased nn.Linear) on tcgen05 -----------
 * Float32 in and out, float32-accurate: every operand is split into three bfloat16 terms and each
 * product is assembled from the six significant partial products on the tensor cores (float32
 * accumulation in TMEM). ld*: row strides in elements (column slices of wider tensors are fine).
 *
 * atmonr_linear_fwd_tc:  Y (M, n_out) = act(X' (M, k_in) * B (n_out, k_in)^T + bias)
 *   X = x, or, with x2 != NULL, the concatenation [x (M, k_split) | x2 (M, k_in - k_split)] read
 *   in place (models/nerf.py:62,86 `torch.cat([x, x_pos])`; k_split a multiple of 8).
 *   X' = X, or, with mask != NULL, X where mask (M, k_in) > 0 and 0 elsewhere (the ReLU
 *   derivative of a layer's output applied to the incoming gradient while it is staged).
 *   B is given as `planes`, produced once per step by atmonr_linear_prep from w (n_out, k_in)
 *   row-major, or, with transpose != 0, from the transpose of w (k_in, n_out) row-major (the
 *   input-gradient product dX = dY * W is this call on the planes of W^T). planes:
 *       ceil(n_out / 256) * ceil(k_in / 32) * terms * 16384 bytes.
 *   bias (n_out) may be NULL; act: 0 none, 1 ReLU.
 *   out_mask (M, n_out; may be NULL): Y is zeroed where out_mask <= 0, applied while the rows are written
 *   out. In the backward chain of an MLP the input gradient dX = dY * W of layer l+1 IS the gradient of
 *   layer l's post-ReLU output, whose values are out_mask: the result is then layer l's pre-activation
 *   gradient and neither product of layer l has to read a mask in its main loop.
 *   bits_out (M, n_out / 32 uint32; may be NULL; needs act == 1 and n_out % 128 == 0): receives the sign bits
 *   (Y > 0) of this product's ReLU output; bits_in: such an array used in place of out_mask (32 bytes per
 *   256-wide row instead of 1 KB). The bit layout is private to this entry point.
 *   terms: 3 = the float32-exact flavour above (six products); 2 = two bf16 terms per operand and the
 *   three products hi*hi + hi*lo + lo*hi (product error ~2^-16 relative: half the tensor-core work,
 *   two CTAs per SM); the planes must have been prepared with the same `terms`.
 * atmonr_linear_dw_tc:   dW (n_out, k_in) += dY' (M, n_out)^T * X (M, k_in)   (dW contiguous,
 *   zeroed or pre-loaded by the caller; mask (M, n_out) as above, applied to dY; x / x2 / k_split
 *   as above). db (n_out, may be NULL) += column sums of dY' (the bias gradient, accumulated by
 *   the threads that stage dY). */
int atmonr_linear_prep(const float* w, int n_out, int k_in, int transpose, int terms, void* planes,
                       void* stream);
int atmonr_linear_fwd_tc(const float* x, int64_t ldx, const float* x2, int64_t ldx2, int k_split,
                         const float* mask, int64_t ldm, const void* planes, const float* bias,
                         int64_t M, int n_out, int k_in, int act, int terms, const float* out_mask,
                         int64_t ldom, const void* bits_in, void* bits_out, float* y, int64_t ldy,
                         void* stream);
int atmonr_linear_dw_tc(const float* dy, int64_t ldy, const float* mask, int64_t ldm,
                        const float* x, int64_t ldx, const float* x2, int64_t ldx2, int k_split,
                        int64_t M, int n_out, int k_in, int terms, float* dw, float* db,
                        void* stream);

/* ---- tensor-core self test -------------------------------------------------------------------
 * One 128-row tile through the three tcgen05 operand configurations of the fused kernels.
 * a, b: (128, 32) fp16 row-major; d: (128, 32) float32.
 *   mode 0: d = a @ b[:32].T   (forward layer: A and B K-major)
 *   mode 1: d = a @ b[:32]     (input gradient: B read MN-major)
 *   mode 2: d[:32] = a.T @ b   (weight gradient: A and B MN-major, K = 128 rows)
 *   mode 3: d is (128, 64); d[:32, :32] + d[32:64, 32:64] = a.T @ b            (weight gradient with two
 *   mode 4: d is (128, 64); d[:32, :16] + d[32:64, 16:32] = a.T @ b[:, :16]     sample groups per MMA) */
int atmonr_tc_probe(const void* a_f16, const void* b_f16, int mode, float* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ATMONR_B200_H */
