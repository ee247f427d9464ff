// linear_tc.cu -- float32-accurate dense layer on the 5th-gen tensor cores (sm_100a), for the
// 256-wide AtmoNeRF MLP (reference: src/atmonr/models/nerf.py:6-93).
//
//   Y[M, n_out] = act(X[M, k_in] * B[n_out, k_in]^T + bias)            (all float32 in HBM)
//
// The reference computes these layers in float32 and north_star asks for 1e-3 relative agreement of
// the rendered radiances, which single-pass TF32/bf16 tensor-core products miss (measured with
// cuBLAS TF32: 1.7e-3..3.8e-3). So every float32 operand is split into three
// bfloat16 terms  v = hi + mid + lo  (each the bf16 rounding of what the previous ones left: 24
// significand bits in total, and bf16 has float32's exponent range, so there is no scaling to
// manage) and the product is assembled from the six significant partial products
//      hi*hi + hi*mid + mid*hi + mid*mid + hi*lo + lo*hi                (dropped terms <= 2^-25)
// as six tcgen05.mma.kind::f16 (bf16 inputs, float32 accumulate) into ONE accumulator in TMEM.
// TERMS = 2 is the two-term flavour  v = hi + lo  (16 significand bits) with the three products
//      hi*hi + hi*lo + lo*hi                                            (dropped terms <= 2^-17)
// i.e. a product error of ~2^-16 relative, sixty times below what one TF32 pass delivers and
// measured at < 1e-4 of the rendered radiances (tests/test_zz_gpu_linear_tc.py): half the tensor-core
// work and two thirds of the staging, and the stages shrink enough for TWO CTAs per SM, whose
// independent tiles hide each other's global-load latency. The NeRF pipeline trains with TERMS = 2;
// TERMS = 3 stays available (ATMONR_LINEAR_TERMS=3) as the float32-exact cross-check.
//
// Tile: 128 rows of X x up to 256 columns of Y per CTA (grid.y walks wider layers: fc9 has 256+V
// outputs), K in chunks of 32. Operand tiles live in shared memory in the no-swizzle core-matrix
// layout of tc_common.cuh; two stages, so the loads and the split of chunk c+1 run underneath the
// twelve MMAs of chunk c (a stage is recycled when the commit of the MMAs that read it has arrived).
// B is split ONCE per step by atmonr_linear_prep (optionally transposed: the input-gradient product
// dX = dY * W is the same kernel on the planes of W^T) and stored in HBM already in tile order, so
// its staging is ONE bulk asynchronous copy per plane (cp.async.bulk, completion on an mbarrier: no
// thread touches the weights); X is split on the fly by the loading threads. The output tile goes
// through shared memory (the operand stages are free by then) so that the global stores are whole
// contiguous rows (a thread owns one ROW of the accumulator: direct stores would touch 32 rows per
// warp instruction; measured round 2: 100 k cycles per tile with per-thread weight copies and
// direct stores, 12.4 k of which are tensor-core time).
//
// Where a tile's time went with TWO register sets (-DATM_LIN_TIMING + scripts/diag_lin_timing.py,
// 786 432 x 256 x 256, two CTAs per SM): 47 k cycles per 128-row tile = prologue 1.5 k, first chunk 8 k
// (the first activations come from DRAM), seven more chunks 3.3 k each, of which 2 k were spent waiting
// for activations requested two chunks (6 k cycles) earlier, epilogue 10-14 k; the six MMAs of a chunk
// take 0.8 k. Hence three sets (three chunks, 48 KB per CTA, in flight): layer 0.55 -> 0.51 ms, NeRF step
// 25.4 -> 23.8 ms; a fourth set fits the 128 registers but was slower (24.4 ms). The tensor pipe is 29 %
// busy; raising it further needs more activation bytes in flight per SM than registers allow (bulk copies
// of float32 rows into a shared-memory ring need the space the second CTA occupies), i.e. bf16
// activations between the layers.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace atm {
using namespace tc;

namespace lin {
constexpr int kRows = 128;                        // rows of X per CTA == TMEM lanes
constexpr int kCols = 256;                        // columns of Y per CTA (max N of one MMA)
constexpr int kChunk = 32;                        // K per stage
constexpr int kThreads = 256;
constexpr int kATile = kRows * kChunk * 2;        // 8 KB: one bf16 plane of the X chunk
constexpr int kBTile = kCols * kChunk * 2;        // 16 KB: one bf16 plane of the B chunk
constexpr int kHalf = 128;                        // columns of one epilogue pass
constexpr int kOutTile = kRows * (kHalf + 4) * 4; // 66 KB: staged output, one column half
constexpr uint32_t kTmemCols = 256;
template <int TERMS>
struct Map {
  static constexpr int kStage = TERMS * kATile + TERMS * kBTile;   // 72 KB (3 terms) / 48 KB (2 terms)
  static constexpr int kBar = 2 * kStage;   // bar[s]: stage s consumed by its MMAs; full[s]: B planes of stage s landed
  static constexpr int kTmemPtr = kBar + 32;
  static constexpr int kBias = kTmemPtr + 16;   // this CTA's 256 bias values (zero where there is none)
  static constexpr int kBytes = kBias + kCols * 4;
  static_assert(kOutTile <= 2 * kStage, "the staged output half must fit in the operand stages");
};
}  // namespace lin

__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// v[8] -> the three bf16 planes, 8 values (16 bytes) each
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& mid, uint4& lo) {
  uint32_t* h = reinterpret_cast<uint32_t*>(&hi);
  uint32_t* m = reinterpret_cast<uint32_t*>(&mid);
  uint32_t* l = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a = v[2 * j], b = v[2 * j + 1];
    const __nv_bfloat162 ph = __floats2bfloat162_rn(a, b);
    const float ra = a - __bfloat162float(ph.x), rb = b - __bfloat162float(ph.y);   // exact
    const __nv_bfloat162 pm = __floats2bfloat162_rn(ra, rb);
    const float sa = ra - __bfloat162float(pm.x), sb = rb - __bfloat162float(pm.y); // exact
    const __nv_bfloat162 pl = __floats2bfloat162_rn(sa, sb);
    h[j] = *reinterpret_cast<const uint32_t*>(&ph);
    m[j] = *reinterpret_cast<const uint32_t*>(&pm);
    l[j] = *reinterpret_cast<const uint32_t*>(&pl);
  }
}

// v[8] -> two bf16 planes (hi, lo)
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t* h = reinterpret_cast<uint32_t*>(&hi);
  uint32_t* l = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a = v[2 * j], b = v[2 * j + 1];
    const __nv_bfloat162 ph = __floats2bfloat162_rn(a, b);
    const __nv_bfloat162 pl = __floats2bfloat162_rn(a - __bfloat162float(ph.x), b - __bfloat162float(ph.y));
    h[j] = *reinterpret_cast<const uint32_t*>(&ph);
    l[j] = *reinterpret_cast<const uint32_t*>(&pl);
  }
}
// split 8 values into TERMS planes `pitch` bytes apart, at 16-byte chunk (r, cc) of a [rows][width] tile
template <int TERMS>
__device__ __forceinline__ void split_store(const float (&v)[8], uint8_t* tile, int pitch, int r, int cc, int width) {
  if (TERMS == 3) {
    uint4 hi, mid, lo;
    split8(v, hi, mid, lo);
    st_chunk(tile, r, cc, width, hi);
    st_chunk(tile + pitch, r, cc, width, mid);
    st_chunk(tile + 2 * pitch, r, cc, width, lo);
  } else {
    uint4 hi, lo;
    split8(v, hi, lo);
    st_chunk(tile, r, cc, width, hi);
    st_chunk(tile + pitch, r, cc, width, lo);
  }
}
// (A plane, B plane) of the significant partial products
template <int TERMS> struct Products;
template <> struct Products<3> {
  static constexpr int kN = 6;
  __device__ static constexpr int a(int t) { return t == 0 ? 0 : t == 1 ? 0 : t == 2 ? 1 : t == 3 ? 1 : t == 4 ? 0 : 2; }
  __device__ static constexpr int b(int t) { return t == 0 ? 0 : t == 1 ? 1 : t == 2 ? 0 : t == 3 ? 1 : t == 4 ? 2 : 0; }
};
template <> struct Products<2> {
  static constexpr int kN = 3;
  __device__ static constexpr int a(int t) { return t == 2 ? 1 : 0; }
  __device__ static constexpr int b(int t) { return t == 1 ? 1 : 0; }
};

// ---- staging in groups of FOUR values ------------------------------------------------------------
// The activations are staged by warp instructions that read 8 rows x 64 contiguous bytes (lane / 4 =
// row of an 8-row group, lane % 4 = 16-byte piece): whole 32-byte sectors, each requested once, and the
// 8-byte shared-memory stores of a warp fill two core matrices exactly (128 contiguous bytes each: no
// bank conflicts). The first version read 8 rows x four 16-byte pieces 32 bytes apart and the other
// halves with a second instruction: every sector twice, and the L1 data pipe at 70-85 % of its
// wavefront rate was what bound both dense-layer kernels (ncu round 2).
// four consecutive float32 of a row (`left` = columns left in the row; fewer than 4 -> zero fill),
// optionally zeroed where the matching entry of `m` is not positive
__device__ __forceinline__ void load4(const float* __restrict__ src, const float* __restrict__ m, int left, bool vec_ok,
                                      float (&v)[4]) {
  if (vec_ok && left >= 4) {
    // (ld.global.cg, past the L1, was measured: 8 % slower)
    const float4 p = *reinterpret_cast<const float4*>(src);
    v[0] = p.x, v[1] = p.y, v[2] = p.z, v[3] = p.w;
    if (m) {
      const float4 a = *reinterpret_cast<const float4*>(m);
      if (!(a.x > 0.0f)) v[0] = 0.0f;
      if (!(a.y > 0.0f)) v[1] = 0.0f;
      if (!(a.z > 0.0f)) v[2] = 0.0f;
      if (!(a.w > 0.0f)) v[3] = 0.0f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < left) v[j] = (m && !(m[j] > 0.0f)) ? 0.0f : src[j];
  }
}
// split 4 values into TERMS planes `pitch` bytes apart; `at` = the 8-byte slot of the values in plane 0
template <int TERMS>
__device__ __forceinline__ void split_store4(const float (&v)[4], uint8_t* at, int pitch) {
  uint2 pl[TERMS];
  float r[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
  for (int t = 0; t < TERMS; ++t) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(r[0], r[1]), b = __floats2bfloat162_rn(r[2], r[3]);
    pl[t] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    if (t + 1 < TERMS) {   // what this term leaves (exact in float32)
      r[0] -= __bfloat162float(a.x), r[1] -= __bfloat162float(a.y);
      r[2] -= __bfloat162float(b.x), r[3] -= __bfloat162float(b.y);
    }
  }
#pragma unroll
  for (int t = 0; t < TERMS; ++t) *reinterpret_cast<uint2*>(at + t * pitch) = pl[t];
}

// B (n_out, k_in) float32 row-major, or its transpose when `transpose` (then the source is
// (k_in, n_out) row-major) -> planes[(tile * k_chunks + chunk) * TERMS + plane][256 x 32 bf16, tile layout]
template <int TERMS>
__global__ void k_linear_prep(const float* __restrict__ w, int n_out, int k_in, int transpose, int k_chunks,
                              int n_tiles, uint8_t* __restrict__ planes) {
  const int groups = k_chunks * 4;  // 8-column groups per row
  const int64_t total = (int64_t)n_tiles * lin::kCols * groups;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int row = (int)(i / groups), grp = (int)(i % groups);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = grp * 8 + j;
    float x = 0.0f;
    if (row < n_out && k < k_in) x = transpose ? w[(size_t)k * n_out + row] : w[(size_t)row * k_in + k];
    v[j] = x;
  }
  const int tile = row / lin::kCols, r = row % lin::kCols, chunk = grp >> 2, cc = grp & 3;
  uint8_t* base = planes + ((size_t)tile * k_chunks + chunk) * TERMS * lin::kBTile;
  split_store<TERMS>(v, base, lin::kBTile, r, cc, lin::kChunk);
}

#ifdef ATM_LIN_TIMING
// phase timestamps (SM clock) of 64 CTAs from the middle of the grid: tuning aid, not part of the product build
__device__ long long g_lin_timing[64 * 8];
__device__ long long g_lin_timing2[64 * 8];
#define ATM_LIN_STAMP2(slot)                                                                      \
  if (c == 5 && tid == 0 && blockIdx.y == 0 && blockIdx.x >= gridDim.x / 2 && blockIdx.x < gridDim.x / 2 + 64) \
    g_lin_timing2[(blockIdx.x - gridDim.x / 2) * 8 + (slot)] = clock64();
#define ATM_LIN_STAMP(slot)                                                                       \
  if (tid == 0 && blockIdx.y == 0 && blockIdx.x >= gridDim.x / 2 && blockIdx.x < gridDim.x / 2 + 64) \
    g_lin_timing[(blockIdx.x - gridDim.x / 2) * 8 + (slot)] = clock64();
#else
#define ATM_LIN_STAMP(slot)
#define ATM_LIN_STAMP2(slot)
#endif

template <int TERMS>
__global__ void __launch_bounds__(lin::kThreads, TERMS == 2 ? 2 : 1)
k_linear_tc(const float* __restrict__ x, int64_t ldx, const float* __restrict__ x2, int64_t ldx2, int k_split,
            const float* __restrict__ mask, int64_t ldm, const uint8_t* __restrict__ planes, const float* __restrict__ bias, int64_t M, int n_out, int k_in,
            int k_chunks, int act, const float* __restrict__ out_mask, int64_t ldom, const uint32_t* __restrict__ bits_in,
            uint32_t* __restrict__ bits_out, float* __restrict__ y, int64_t ldy) {
  extern __shared__ __align__(128) uint8_t smem[];
  using MapT = lin::Map<TERMS>;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + MapT::kBar);          // bar[s]: MMAs that read stage s are done
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + MapT::kTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * lin::kRows;
  const int n0 = blockIdx.y * lin::kCols;                                  // first output column of this CTA
  ATM_LIN_STAMP(0)
  if (warp == 0) tmem_alloc<lin::kTmemCols>(tmem_ptr);
  uint64_t* full = bar + 2;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    mbar_init(full, 1);
    mbar_init(full + 1, 1);
    fence_mbar_init();
  }
  float* sbias = reinterpret_cast<float*>(smem + MapT::kBias);
  sbias[tid] = (bias && n0 + tid < n_out) ? __ldg(bias + n0 + tid) : 0.0f;   // kThreads == kCols
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t acc = *tmem_ptr;
  const uint32_t sbase = smem_u32(smem);

  const int n_cols = min(lin::kCols, ((n_out + 15) / 16) * 16 - n0);       // MMA N (multiple of 16)
  const uint32_t idesc = make_idesc_bf16(lin::kRows, n_cols);
  const uint8_t* b_src = planes + (size_t)blockIdx.y * k_chunks * TERMS * lin::kBTile;
  const bool vec_ok = (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                      (!mask || ((ldm & 3) == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0)) &&
                      (!x2 || ((ldx2 & 3) == 0 && (reinterpret_cast<uintptr_t>(x2) & 15) == 0));

  // this thread's 2 x 8 values of a chunk: a quarter warp (8 lanes) owns the 8 rows of ONE core matrix =
  // 128 contiguous bytes of shared memory (no bank conflicts); in global memory the warp reads 8 rows x
  // 128 contiguous bytes
  // The values are requested THREE chunks ahead (three register sets, set = chunk % 3): one
  // chunk's worth of bytes in flight per CTA (16 KB) did not cover the global-load latency (ncu round 2:
  // 35 % of the warp samples on the long scoreboard, DRAM at a third of its rate).
  float xa[4][4], xb[4][4], xc[4][4];   // chunks c, c + 1, c + 2 in flight (set = chunk % 3)
  // This thread's part of a chunk (128 rows x 32 columns): 4 consecutive columns (16 bytes) of four rows
  // 32 apart. A warp instruction reads 8 rows x 64 bytes: warp w takes the 64-byte half (w & 1) of the row
  // groups (w >> 1) + 4 it. Row and column are the same for every chunk: the pointers are formed once per
  // tile (the per-chunk 64-bit products were a quarter of all issued instructions).
  // (lanes 0-15 fill one core matrix, lanes 16-31 the next one: each half-warp's 8-byte stores cover 128
  // contiguous bytes of shared memory, no bank conflict)
  const int lane_f = tid & 31;
  const int piece = (lane_f & 1) | ((lane_f >> 4) << 1);           // 16-byte piece of the 64-byte half row
  const int kq = ((warp & 1) << 4) | (piece << 2);                 // first of this thread's 4 columns in a chunk
  const int fr0 = ((warp >> 1) << 3) | ((lane_f >> 1) & 7);        // its first row in the tile
  const int64_t frow = row0 + fr0;
  const float* xr = x + frow * ldx + kq;
  const float* x2r = x2 ? x2 + frow * ldx2 + (kq - k_split) : nullptr;
  const float* mr = mask ? mask + frow * ldm + kq : nullptr;
  const int64_t step_x = 32 * ldx, step_x2 = 32 * ldx2, step_m = 32 * ldm;
  // byte offset of (row fr0, column kq) in a plane of the A tile; + 2048 per 32 rows
  const uint32_t a_off = (uint32_t)((fr0 >> 3) * (lin::kChunk >> 3) * kCore + (kq >> 3) * kCore + (fr0 & 7) * 16 + (kq & 7) * 2);
  auto fetch_x = [&](int c, float (&xv)[4][4]) {
    const int k0 = c * lin::kChunk + kq;
    const bool second = k0 >= k_split;
    const int left = (second ? k_in : k_split) - k0;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[it][j] = 0.0f;
      if (frow + it * 32 < M && c < k_chunks) {
        // columns [0, k_split) come from x, [k_split, k_in) from x2 (k_split is a multiple of 8: a group
        // of 4 never straddles the two)
        const float* src = (second ? x2r + it * step_x2 : xr + it * step_x) + c * lin::kChunk;
        load4(src, mr ? mr + it * step_m + c * lin::kChunk : nullptr, left, vec_ok, xv[it]);
      }
    }
  };
  ATM_LIN_STAMP(1)
  fetch_x(0, xa);
  fetch_x(1, xb);
  fetch_x(2, xc);

  // ---- B chunks: TERMS planes, already split and in tile order: the first n_cols rows of a plane are its
  // first n_cols * 64 bytes -> one bulk copy per plane, announced on full[stage]. The copy of chunk c + 1 is
  // issued during iteration c (as soon as the MMAs of chunk c - 1, which read that stage, have committed),
  // the first two right here: issued at the top of its own iteration, a copy's L2 round trip was exposed in
  // every chunk (8 x ~1.7 us of a 27 us tile).
  auto issue_b = [&](int c) {
    const int sb = c & 1;
    const uint8_t* src = b_src + (size_t)c * TERMS * lin::kBTile;
    uint8_t* dst = smem + sb * MapT::kStage + TERMS * lin::kATile;
    const uint32_t bytes = (uint32_t)n_cols * 64u;
    mbar_expect_tx(full + sb, (uint32_t)TERMS * bytes);
#pragma unroll
    for (int p = 0; p < TERMS; ++p) bulk_g2s(dst + p * lin::kBTile, src + p * lin::kBTile, bytes, full + sb);
  };
  if (tid == 0) {
    issue_b(0);
    if (k_chunks > 1) issue_b(1);
  }

  int set3 = 0;
  for (int c = 0; c < k_chunks; ++c) {
    const int s = c & 1;
    uint8_t* stage = smem + s * MapT::kStage;
    if (c == 1) { ATM_LIN_STAMP(2) }
    if (c == 4) { ATM_LIN_STAMP(3) }
    // the MMAs of chunk c-2 read this stage: wait for their commit (completion number (c>>1)-1 of bar[s])
    ATM_LIN_STAMP2(0)
    if (c >= 2) mbar_wait(bar + s, (uint32_t)(((c >> 1) - 1) & 1));
    ATM_LIN_STAMP2(1)
    // ---- X chunk: 128 rows x 32 columns float32 -> TERMS bf16 planes (4 groups of 4 values per thread)
    // the set of chunk c is split and stored, then refilled with chunk c + 3: three chunks' worth of loads
    // (48 KB per CTA) stay in flight -- with two, 2 k of every 3 k-cycle chunk were spent waiting for them
    if (set3 == 0) {
#pragma unroll
      for (int it = 0; it < 4; ++it) split_store4<TERMS>(xa[it], stage + a_off + it * 2048, lin::kATile);
      fetch_x(c + 3, xa);
    } else if (set3 == 1) {
#pragma unroll
      for (int it = 0; it < 4; ++it) split_store4<TERMS>(xb[it], stage + a_off + it * 2048, lin::kATile);
      fetch_x(c + 3, xb);
    } else {
#pragma unroll
      for (int it = 0; it < 4; ++it) split_store4<TERMS>(xc[it], stage + a_off + it * 2048, lin::kATile);
      fetch_x(c + 3, xc);
    }
    set3 = set3 == 2 ? 0 : set3 + 1;
    ATM_LIN_STAMP2(2)
    ATM_LIN_STAMP2(3)
    if (tid == 0 && c >= 1 && c + 1 < k_chunks) {
      // chunk c - 1 (the other stage) is the ((c-1)>>1)-th completion of its barrier
      mbar_wait(bar + (s ^ 1), (uint32_t)(((c - 1) >> 1) & 1));
      issue_b(c + 1);
    }
    ATM_LIN_STAMP2(4)
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    ATM_LIN_STAMP2(5)
    if (warp == 0) {
      mbar_wait(full + s, (uint32_t)((c >> 1) & 1));   // the weight planes of this chunk have landed
      ATM_LIN_STAMP2(6)
      tc_fence_after();
      const uint32_t a0 = sbase + s * MapT::kStage, b0 = a0 + TERMS * lin::kATile;
      // (A plane, B plane): hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi  /  hi*hi, hi*lo, lo*hi
      using Pr = Products<TERMS>;
#pragma unroll
      for (int k = 0; k < lin::kChunk / 16; ++k) {
#pragma unroll
        for (int t = 0; t < Pr::kN; ++t) {
          const uint64_t ad = desc_k_major(a0 + Pr::a(t) * lin::kATile + k * 2 * kCore, lin::kChunk);
          const uint64_t bd = desc_k_major(b0 + Pr::b(t) * lin::kBTile + k * 2 * kCore, lin::kChunk);
          umma_f16(acc, ad, bd, idesc, (c | k | t) != 0 ? 1u : 0u);
        }
      }
      umma_commit(bar + s);
    }
  }
  ATM_LIN_STAMP(4)
  // every MMA has been issued by warp 0 in order; the last commit covers them all
  {
    const int last = k_chunks - 1;
    mbar_wait(bar + (last & 1), (uint32_t)((last >> 1) & 1));
    tc_fence_after();
  }
  ATM_LIN_STAMP(5)
  // ---- epilogue, one 128-column half of the accumulator at a time: thread (warp w, lane) owns row
  // (w % 4) * 32 + lane and columns (w / 4) * 64 .. + 64 of the half. bias + activation, then the half is
  // staged in shared memory (the operand stages are free: every MMA is complete; row pitch 132 floats:
  // the 16-byte stores of 8 lanes = 8 rows fall into distinct banks) and written out row by row, a
  // warp covering 512 contiguous bytes per instruction.
  {
    float* tile = reinterpret_cast<float*>(smem);
    constexpr int pitch = lin::kHalf + 4;
    const int r = (warp & 3) * 32 + (tid & 31);
    const int n_valid = min(n_cols, n_out - n0);         // real output columns of this CTA
    const int rows = (int)min((int64_t)lin::kRows, M - row0);
    const bool vec = (ldy & 3) == 0 && (n0 & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
    const bool om_vec = !out_mask || ((ldom & 3) == 0 && (reinterpret_cast<uintptr_t>(out_mask) & 15) == 0);
    const int wpr = (n_out >> 7) << 2;                   // sign-bit words per row: 4 per 128 columns
    for (int h0 = 0; h0 < n_cols; h0 += lin::kHalf) {
      if (h0 > 0) __syncthreads();                       // the previous half has been written out
      const int w_valid = min(lin::kHalf, n_valid - h0);   // real columns of this half
      // out_mask (M, n_out) floats / bits_in (M, n_out / 32) words: the output is zeroed where the mask is
      // <= 0 / the bit is clear -- the ReLU derivative of the layer whose OUTPUT this product's result is
      // the gradient of, applied where the rows are written out with coalesced accesses (so that no
      // downstream product has to read a mask in its main loop). bits_out: the sign bits of this
      // product's own ReLU output, for the backward pass to use that way (32 bytes per 256-wide row
      // instead of re-reading 1 KB of activations). Bit layout, shared by writer and reader: row r,
      // 128-column half h, word k (0..3), bit q  <->  column 128 h + 4 q + k, i.e. exactly the (lane,
      // component) that handles the column in the loops below.
      const float* omh = out_mask ? out_mask + n0 + h0 : nullptr;
      // Whole, aligned half (every 256-/128-wide layer): thread t handles the 16-byte groups t, t + 256, ...
      // = row (t / 32) + 8 u, group lane. Its 16 mask groups (or sign words) are requested HERE, before
      // the accumulator is staged, so that their latency sits underneath the TMEM reads and the
      // shared-memory pass (one dependent global load per written group measured 62 % of the warp
      // samples on the long scoreboard).
      const bool fast = vec && om_vec && w_valid == lin::kHalf;
      const int lane = tid & 31, wrow = tid >> 5;
      const int bword = ((n0 + h0) >> 7) << 2;
      float4 mreg[16];
      if (fast && omh) {
        const float* mp = omh + (row0 + wrow) * ldom + 4 * lane;
        const int64_t mstep = 8 * ldom;
#pragma unroll
        for (int u = 0; u < 16; ++u)
          if (wrow + 8 * u < rows) mreg[u] = *reinterpret_cast<const float4*>(mp + u * mstep);
      } else if (fast && bits_in) {
        const uint32_t* bp = bits_in + (row0 + wrow) * wpr + bword;
#pragma unroll
        for (int u = 0; u < 16; ++u)
          if (wrow + 8 * u < rows) {
            const uint4 b = *reinterpret_cast<const uint4*>(bp + u * 8 * wpr);
            mreg[u] = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w));
          }
      }
#pragma unroll 1
      for (int cb = 0; cb < 64; cb += 16) {
        const int lc = (warp >> 2) * 64 + cb, col = h0 + lc;
        if (col >= n_cols) break;                        // warp-uniform
        float v[16];
        tmem_ld16(tmem_addr(acc, warp, col), v);
        const float4* bq = reinterpret_cast<const float4*>(sbias + col);   // same address in every lane: broadcast
        float4* dst = reinterpret_cast<float4*>(tile + r * pitch + lc);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 bb = bq[q];
          float4 o = make_float4(v[4 * q] + bb.x, v[4 * q + 1] + bb.y, v[4 * q + 2] + bb.z, v[4 * q + 3] + bb.w);
          if (act == 1) o = make_float4(fmaxf(o.x, 0.0f), fmaxf(o.y, 0.0f), fmaxf(o.z, 0.0f), fmaxf(o.w, 0.0f));
          dst[q] = o;
        }
      }
      __syncthreads();
      if (w_valid <= 0) continue;                          // uniform
      float* yh = y + n0 + h0;
      if (fast) {
        float* yp = yh + (row0 + wrow) * ldy + 4 * lane;
        const int64_t ystep = 8 * ldy;
        uint32_t* bo = bits_out ? bits_out + (row0 + wrow) * wpr + bword : nullptr;
        const float* tp = tile + wrow * pitch + 4 * lane;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int rr = wrow + 8 * u;
          if (rr >= rows) break;                           // warp-uniform
          float4 v = *reinterpret_cast<const float4*>(tp + u * 8 * pitch);
          if (omh) {
            const float4 m = mreg[u];
            v.x = m.x > 0.0f ? v.x : 0.0f, v.y = m.y > 0.0f ? v.y : 0.0f;
            v.z = m.z > 0.0f ? v.z : 0.0f, v.w = m.w > 0.0f ? v.w : 0.0f;
          } else if (bits_in) {
            const float4 m = mreg[u];
            v.x = ((__float_as_uint(m.x) >> lane) & 1u) ? v.x : 0.0f, v.y = ((__float_as_uint(m.y) >> lane) & 1u) ? v.y : 0.0f;
            v.z = ((__float_as_uint(m.z) >> lane) & 1u) ? v.z : 0.0f, v.w = ((__float_as_uint(m.w) >> lane) & 1u) ? v.w : 0.0f;
          }
          *reinterpret_cast<float4*>(yp + u * ystep) = v;
          if (bo) {
            const uint32_t b0 = __ballot_sync(0xffffffffu, v.x > 0.0f), b1 = __ballot_sync(0xffffffffu, v.y > 0.0f);
            const uint32_t b2 = __ballot_sync(0xffffffffu, v.z > 0.0f), b3 = __ballot_sync(0xffffffffu, v.w > 0.0f);
            if (lane == 0) *reinterpret_cast<uint4*>(bo + u * 8 * wpr) = make_uint4(b0, b1, b2, b3);
          }
        }
      } else if (vec && om_vec) {
        const int q_per_row = w_valid >> 2;                // whole float4 groups
        for (int i = tid; i < rows * q_per_row; i += lin::kThreads) {
          const int rr = i / q_per_row, q = i - rr * q_per_row;
          float4 v = *reinterpret_cast<const float4*>(tile + rr * pitch + 4 * q);
          if (omh) {
            const float4 m = *reinterpret_cast<const float4*>(omh + (row0 + rr) * ldom + 4 * q);
            v.x = m.x > 0.0f ? v.x : 0.0f, v.y = m.y > 0.0f ? v.y : 0.0f;
            v.z = m.z > 0.0f ? v.z : 0.0f, v.w = m.w > 0.0f ? v.w : 0.0f;
          }
          *reinterpret_cast<float4*>(yh + (row0 + rr) * ldy + 4 * q) = v;
        }
        const int tail = w_valid & 3;
        for (int i = tid; i < rows * tail; i += lin::kThreads) {
          const int rr = i / tail, cc = (w_valid & ~3) + (i - rr * tail);
          const float v = tile[rr * pitch + cc];
          yh[(row0 + rr) * ldy + cc] = (omh && !(omh[(row0 + rr) * ldom + cc] > 0.0f)) ? 0.0f : v;
        }
      } else {
        for (int i = tid; i < rows * w_valid; i += lin::kThreads) {
          const int rr = i / w_valid, cc = i - rr * w_valid;
          const float v = tile[rr * pitch + cc];
          yh[(row0 + rr) * ldy + cc] = (omh && !(omh[(row0 + rr) * ldom + cc] > 0.0f)) ? 0.0f : v;
        }
      }
    }
  }
  ATM_LIN_STAMP(6)
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<lin::kTmemCols>(acc);
  ATM_LIN_STAMP(7)
}

// =========================================================================================
// weight gradient: dW[n_out, k_in] += sum over rows m of dY[m, n_out] * X[m, k_in]
// =========================================================================================
// The reduction runs over the ROWS, so both operands are read MN-major straight from their
// row-major chunks (32 rows per stage; tc_common.cuh: the same bytes serve both views). A CTA owns a
// contiguous slab of rows, accumulates a 256 x 256 block of dW in TMEM (two M = 128 accumulators,
// all 512 columns) and adds it to dW with float32 REDs at the end; grid.y / grid.z walk wider
// layers in blocks of 256 (fc9: n_out = 256 + V, fc6: k_in = 256 + 76). Same six-product bf16
// split as the forward; the ReLU derivative is applied to dY while it is staged (`mask` = the
// layer's output).
namespace ldw {
constexpr int kChunk = 32;                         // rows per stage (the MMA K dimension)
constexpr int kWide = 256;                         // columns of a staged operand block
constexpr int kThreads = 512;                         // 16 warps: the staging code is a chain of dependent conversions,
                                                   // 8 warps left the SM at an IPC of 0.9 (ncu round 2)
constexpr int kTile = kChunk * kWide * 2;          // 16 KB: one bf16 plane of one operand
constexpr uint32_t kTmemCols = 512;
template <int TERMS>
struct Map {
  static constexpr int kStage = 2 * TERMS * kTile;   // 96 KB (3 terms) / 64 KB (2 terms)
  static constexpr int kBar = 2 * kStage;
  static constexpr int kTmemPtr = kBar + 16;
  static constexpr int kBytes = kTmemPtr + 16;
};
}  // namespace ldw

template <int TERMS>
__global__ void __launch_bounds__(ldw::kThreads, 1)
k_linear_dw_tc(const float* __restrict__ dy, int64_t ldy, const float* __restrict__ mask, int64_t ldm,
               const float* __restrict__ x, int64_t ldx, const float* __restrict__ x2, int64_t ldx2, int k_split,
               int64_t M, int n_out, int k_in, float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ __align__(128) uint8_t smem[];
  using MapT = ldw::Map<TERMS>;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + MapT::kBar);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + MapT::kTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5;
  // this CTA's slab of 32-row chunks
  const int64_t chunks = (M + ldw::kChunk - 1) / ldw::kChunk;
  const int64_t per = (chunks + gridDim.x - 1) / gridDim.x;
  const int64_t c_lo = (int64_t)blockIdx.x * per, c_hi = min(chunks, c_lo + per);
  if (c_lo >= c_hi) return;                                      // uniform: nothing allocated yet
  if (warp == 0) tmem_alloc<ldw::kTmemCols>(tmem_ptr);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    fence_mbar_init();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t acc = *tmem_ptr;
  const uint32_t sbase = smem_u32(smem);
  const int n0 = blockIdx.y * ldw::kWide, k0 = blockIdx.z * ldw::kWide;     // block of dW owned by this CTA
  const int m_halves = (min(n_out - n0, ldw::kWide) + 127) / 128;           // M = 128 accumulators in use
  const int n_cols = min(ldw::kWide, ((k_in - k0 + 15) / 16) * 16);         // MMA N
  const uint32_t idesc = make_idesc_bf16(128, n_cols) | (1u << 15) | (1u << 16);   // A and B MN-major
  const bool vy = (ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
                  (!mask || ((ldm & 3) == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0));
  const bool vx = (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                  (!x2 || ((ldx2 & 3) == 0 && (reinterpret_cast<uintptr_t>(x2) & 15) == 0));

  // bias gradient = column sums of the (masked) dY: the staging threads see every value anyway; the
  // CTAs of the first k_in block add theirs
  float colsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  const bool want_db = db != nullptr && blockIdx.z == 0;

  // This thread's values of the chunks being staged, requested TWO chunks ahead: set A serves the even
  // iterations, set B the odd ones; a set is refilled (chunk it + 2) right after its values have been split
  // and stored, so two chunks (128 KB per SM) are in flight while the tensor core works on a third. (One
  // chunk ahead left the kernel at a third of the DRAM rate, the staging warps waiting on the loads.)
  float ady[4][4], axx[4][4], bdy[4][4], bxx[4][4];
  // A thread's part of a chunk (32 rows x 256 columns per operand): 4 consecutive columns (16 bytes) of the
  // rows 8 it + lane / 4; a warp instruction reads 8 rows x 64 bytes, warp w the columns 16 w .. 16 w + 15
  // (see load4). Columns are the same for every row group and every chunk, the operand pointers advance by
  // 32 rows per fetch (the fetches are issued in chunk order): no 64-bit product inside the loop.
  static_assert(ldw::kThreads == 512, "staging map: 16 warps x 16 columns = 256 columns");
  // (lanes 0-15 fill one core matrix, lanes 16-31 the next one: no shared-memory bank conflict)
  const int lane_ = tid & 31, ccol = (warp << 4) | ((((lane_ & 1) | ((lane_ >> 4) << 1))) << 2), r8 = (lane_ >> 1) & 7;
  const int kdy = n0 + ccol, kx = k0 + ccol;
  const bool dy_on = kdy < n_out, x_on = kx < k_in, x_second = kx >= k_split;
  const int dy_left = n_out - kdy, x_left = (x_second ? k_in : k_split) - kx;
  const int64_t xld = x_second ? ldx2 : ldx;
  const float* dyq = dy + (c_lo * ldw::kChunk + r8) * ldy + kdy;
  const float* mkq = mask ? mask + (c_lo * ldw::kChunk + r8) * ldm + kdy : nullptr;
  const float* xq = (x_second ? x2 + (kx - k_split) : x + kx) + (c_lo * ldw::kChunk + r8) * xld;
  const int64_t ldy8 = 8 * ldy, ldm8 = 8 * ldm, xld8 = 8 * xld;
  // byte offset of (row r8, column ccol) in a plane of a [32][256] tile; + 4096 per 8 rows
  const uint32_t t_off = (uint32_t)((ccol >> 3) * kCore + r8 * 16 + (ccol & 7) * 2);
  auto fetch = [&](int64_t c, float (&vdy)[4][4], float (&vxx)[4][4]) {
    const int64_t row_lo = c * ldw::kChunk + r8;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int j = 0; j < 4; ++j) vdy[a][j] = 0.0f, vxx[a][j] = 0.0f;
      if (c < c_hi && row_lo + 8 * a < M) {
        if (dy_on) load4(dyq + a * ldy8, mkq ? mkq + a * ldm8 : nullptr, dy_left, vy, vdy[a]);
        if (x_on) load4(xq + a * xld8, nullptr, x_left, vx, vxx[a]);
      }
    }
    dyq += ldw::kChunk * ldy, xq += ldw::kChunk * xld;
    if (mkq) mkq += ldw::kChunk * ldm;
  };
  auto stage_rows = [&](const float (&vdy)[4][4], const float (&vxx)[4][4], uint8_t* stage) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (want_db) {
#pragma unroll
        for (int j = 0; j < 4; ++j) colsum[j] += vdy[a][j];
      }
      split_store4<TERMS>(vdy[a], stage + t_off + a * 4096, ldw::kTile);
      split_store4<TERMS>(vxx[a], stage + TERMS * ldw::kTile + t_off + a * 4096, ldw::kTile);
    }
  };
  fetch(c_lo, ady, axx);
  fetch(c_lo + 1, bdy, bxx);
  int it = 0;
  for (int64_t c = c_lo; c < c_hi; ++c, ++it) {
    const int s = it & 1;
    uint8_t* stage = smem + s * MapT::kStage;
    if (it >= 2) mbar_wait(bar + s, (uint32_t)(((it >> 1) - 1) & 1));
    if (s == 0) {
      stage_rows(ady, axx, stage);
      fetch(c + 2, ady, axx);
    } else {
      stage_rows(bdy, bxx, stage);
      fetch(c + 2, bdy, bxx);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      const uint32_t a0 = sbase + s * MapT::kStage, b0 = a0 + TERMS * ldw::kTile;
      using Pr = Products<TERMS>;
      for (int h = 0; h < m_halves; ++h) {
#pragma unroll
        for (int k = 0; k < ldw::kChunk / 16; ++k) {
#pragma unroll
          for (int t = 0; t < Pr::kN; ++t) {
            // two 8-row groups per K step; columns 128 h .. of dY are 16 h core matrices further on
            const uint32_t koff = k * 2 * (ldw::kWide / 8) * kCore;
            const uint64_t ad = desc_mn_major(a0 + Pr::a(t) * ldw::kTile + koff + h * 16 * kCore, ldw::kWide);
            const uint64_t bd = desc_mn_major(b0 + Pr::b(t) * ldw::kTile + koff, ldw::kWide);
            umma_f16(acc + h * 256, ad, bd, idesc, (it | k | t) != 0 ? 1u : 0u);
          }
        }
      }
      umma_commit(bar + s);
    }
  }
  {
    const int last = it - 1;
    mbar_wait(bar + (last & 1), (uint32_t)((last >> 1) & 1));
    tc_fence_after();
  }
  if (want_db) {  // the 8 lanes that differ in bits 1-3 hold the same 4 columns for different rows
    const int lane = tid & 31;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = colsum[j];
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      if ((lane & 14) == 0 && kdy + j < n_out) atomicAdd(db + kdy + j, v);
    }
  }
  // ---- epilogue: accumulator h, TMEM lane = row of dW (n_out index), columns = k_in index. A thread owns
  // one ROW of the accumulator, so REDs issued straight from the registers touch 32 rows = 32 sectors
  // per warp instruction (all CTAs add into the same 256 x 256 block: 148 x 65 536 scalar REDs = 9.7 M
  // sectors, ~0.1 ms of a 0.14 ms coarse-pass launch). The block goes through shared memory instead (the
  // operand stages are free: every MMA is complete), one 128-column half at a time, and is added with
  // red.global.add.v4.f32 along the rows: a warp instruction covers 512 contiguous bytes (16 sectors for
  // 128 values instead of 32 sectors for 32).
  {
    float* tile = reinterpret_cast<float*>(smem);
    constexpr int kHalfCols = 128, pitch = kHalfCols + 4;          // 128 x 132 floats = 66 KB <= one stage pair
    static_assert(128 * pitch * 4 <= 2 * MapT::kStage, "the staged half must fit in the operand stages");
    const int lane = tid & 31, wrow = tid >> 5;
    const bool vec = (k_in & 3) == 0 && (k0 & 3) == 0 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0;
    for (int h = 0; h < m_halves; ++h) {
      for (int h0 = 0; h0 < n_cols; h0 += kHalfCols) {
        __syncthreads();                                           // the previous half has been added
        const int r = (warp & 3) * 32 + lane;                      // TMEM lane of this thread
#pragma unroll 1
        for (int cb = 0; cb < 32; cb += 16) {                      // 4 warps per lane quarter x 32 columns
          const int lc = (warp >> 2) * 32 + cb, col = h0 + lc;
          if (col >= n_cols) break;                                // warp-uniform
          float v[16];
          tmem_ld16(tmem_addr(acc, warp, h * 256 + col), v);
          float4* dst = reinterpret_cast<float4*>(tile + r * pitch + lc);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        __syncthreads();
        const int w_valid = min(kHalfCols, k_in - k0 - h0);        // real columns of this half
        const int rows = min(128, n_out - n0 - h * 128);           // real rows of this accumulator
        if (w_valid <= 0 || rows <= 0) continue;                   // uniform
        float* base = dw + (size_t)(n0 + h * 128) * k_in + k0 + h0;
        if (vec && w_valid == kHalfCols) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {                            // 128 rows x 32 groups of 4 over 512 threads
            const int rr = wrow + 16 * u;
            if (rr >= rows) break;                                 // warp-uniform
            const float4 v = *reinterpret_cast<const float4*>(tile + rr * pitch + 4 * lane);
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + (size_t)rr * k_in + 4 * lane),
                         "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                         : "memory");
          }
        } else {
          for (int i = tid; i < rows * w_valid; i += ldw::kThreads) {
            const int rr = i / w_valid, cc = i - rr * w_valid;
            atomicAdd(base + (size_t)rr * k_in + cc, tile[rr * pitch + cc]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<ldw::kTmemCols>(acc);
}

}  // namespace atm

using namespace atm;
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

#define ATM_TERMS_DISPATCH(terms, CALL) \
  if ((terms) == 3) { CALL(3); } else { CALL(2); }

int atmonr_linear_prep(const float* w, int n_out, int k_in, int transpose, int terms, void* planes, void* stream) {
  ATM_REQUIRE(w && planes, "atmonr_linear_prep", "null pointer");
  ATM_REQUIRE(n_out > 0 && k_in > 0, "atmonr_linear_prep", "empty matrix");
  ATM_REQUIRE(terms == 2 || terms == 3, "atmonr_linear_prep", "terms must be 2 or 3");
  const int n_tiles = (n_out + lin::kCols - 1) / lin::kCols, k_chunks = (k_in + lin::kChunk - 1) / lin::kChunk;
  const int64_t total = (int64_t)n_tiles * lin::kCols * k_chunks * 4;
#define CALL(T) \
  k_linear_prep<T><<<grid_for(total, 256), 256, 0, S(stream)>>>(w, n_out, k_in, transpose, k_chunks, n_tiles, \
                                                                reinterpret_cast<uint8_t*>(planes))
  ATM_TERMS_DISPATCH(terms, CALL)
#undef CALL
  ATM_CHECK_LAUNCH("atmonr_linear_prep");
  return 0;
}

int atmonr_linear_fwd_tc(const float* x, int64_t ldx, const float* x2, int64_t ldx2, int k_split, const float* mask,
                         int64_t ldm, const void* planes, const float* bias, int64_t M, int n_out, int k_in, int act,
                         int terms, const float* out_mask, int64_t ldom, const void* bits_in, void* bits_out, float* y,
                         int64_t ldy, void* stream) {
  ATM_REQUIRE(M >= 0 && n_out > 0 && k_in > 0, "atmonr_linear_fwd_tc", "bad shape");
  ATM_REQUIRE(act == 0 || act == 1, "atmonr_linear_fwd_tc", "act must be 0 (none) or 1 (ReLU)");
  ATM_REQUIRE(terms == 2 || terms == 3, "atmonr_linear_fwd_tc", "terms must be 2 or 3");
  if (M == 0) return 0;
  ATM_REQUIRE(x && planes && y, "atmonr_linear_fwd_tc", "null pointer");
  if (!x2) k_split = k_in;
  ATM_REQUIRE(k_split > 0 && k_split <= k_in && (k_split == k_in || k_split % 8 == 0), "atmonr_linear_fwd_tc",
              "k_split must be a multiple of 8 inside (0, k_in]");
  ATM_REQUIRE(ldx >= k_split && (!x2 || ldx2 >= k_in - k_split) && ldy >= n_out && (!mask || ldm >= k_in) &&
                  (!out_mask || ldom >= n_out),
              "atmonr_linear_fwd_tc", "row stride smaller than the row");
  ATM_REQUIRE((M + lin::kRows - 1) / lin::kRows < (1ll << 31), "atmonr_linear_fwd_tc", "too many rows");
  if (bits_in || bits_out) {
    // the sign-bit forms live on the whole-half, 16-byte path of the epilogue
    ATM_REQUIRE(n_out % 128 == 0 && (ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "atmonr_linear_fwd_tc",
                "sign bits need n_out % 128 == 0 and 16-byte aligned output rows");
    ATM_REQUIRE(!(bits_in && out_mask), "atmonr_linear_fwd_tc", "give out_mask or bits_in, not both");
    ATM_REQUIRE(!bits_out || act == 1, "atmonr_linear_fwd_tc", "bits_out records a ReLU output (act must be 1)");
    ATM_REQUIRE(((reinterpret_cast<uintptr_t>(bits_in) | reinterpret_cast<uintptr_t>(bits_out)) & 15) == 0,
                "atmonr_linear_fwd_tc", "sign-bit arrays must be 16-byte aligned");
  }
  const int n_tiles = (n_out + lin::kCols - 1) / lin::kCols, k_chunks = (k_in + lin::kChunk - 1) / lin::kChunk;
  dim3 grid((unsigned)((M + lin::kRows - 1) / lin::kRows), (unsigned)n_tiles);
#define CALL(T)                                                                                                        \
  {                                                                                                                    \
    cudaError_t e = cudaFuncSetAttribute(k_linear_tc<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, lin::Map<T>::kBytes); \
    if (e != cudaSuccess) return fail("atmonr_linear_fwd_tc", cudaGetErrorString(e));                                  \
    k_linear_tc<T><<<grid, lin::kThreads, lin::Map<T>::kBytes, S(stream)>>>(                                           \
        x, ldx, x2, ldx2, k_split, mask, ldm, reinterpret_cast<const uint8_t*>(planes), bias, M, n_out, k_in, k_chunks, \
        act, out_mask, ldom, reinterpret_cast<const uint32_t*>(bits_in), reinterpret_cast<uint32_t*>(bits_out), y, ldy); \
  }
  ATM_TERMS_DISPATCH(terms, CALL)
#undef CALL
  ATM_CHECK_LAUNCH("atmonr_linear_fwd_tc");
  return 0;
}

int atmonr_linear_dw_tc(const float* dy, int64_t ldy, const float* mask, int64_t ldm, const float* x, int64_t ldx,
                        const float* x2, int64_t ldx2, int k_split, int64_t M, int n_out, int k_in, int terms,
                        float* dw, float* db, void* stream) {
  ATM_REQUIRE(M >= 0 && n_out > 0 && k_in > 0, "atmonr_linear_dw_tc", "bad shape");
  ATM_REQUIRE(terms == 2 || terms == 3, "atmonr_linear_dw_tc", "terms must be 2 or 3");
  if (M == 0) return 0;
  ATM_REQUIRE(dy && x && dw, "atmonr_linear_dw_tc", "null pointer");
  if (!x2) k_split = k_in;
  ATM_REQUIRE(k_split > 0 && k_split <= k_in && (k_split == k_in || k_split % 8 == 0), "atmonr_linear_dw_tc",
              "k_split must be a multiple of 8 inside (0, k_in]");
  ATM_REQUIRE(ldy >= n_out && ldx >= k_split && (!x2 || ldx2 >= k_in - k_split) && (!mask || ldm >= n_out),
              "atmonr_linear_dw_tc", "row stride smaller than the row");
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return fail("atmonr_linear_dw_tc", "cannot query the device");
  const int ny = (n_out + ldw::kWide - 1) / ldw::kWide, nz = (k_in + ldw::kWide - 1) / ldw::kWide;
  const int64_t chunks = (M + ldw::kChunk - 1) / ldw::kChunk;
  int64_t slabs = sms / (ny * nz);           // one CTA per SM over all blocks of dW
  if (slabs < 1) slabs = 1;
  if (slabs > chunks) slabs = chunks;
  dim3 grid((unsigned)slabs, (unsigned)ny, (unsigned)nz);
#define CALL(T)                                                                                                           \
  {                                                                                                                       \
    cudaError_t e = cudaFuncSetAttribute(k_linear_dw_tc<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, ldw::Map<T>::kBytes); \
    if (e != cudaSuccess) return fail("atmonr_linear_dw_tc", cudaGetErrorString(e));                                      \
    k_linear_dw_tc<T><<<grid, ldw::kThreads, ldw::Map<T>::kBytes, S(stream)>>>(dy, ldy, mask, ldm, x, ldx, x2, ldx2,      \
                                                                              k_split, M, n_out, k_in, dw, db);           \
  }
  ATM_TERMS_DISPATCH(terms, CALL)
#undef CALL
  ATM_CHECK_LAUNCH("atmonr_linear_dw_tc");
  return 0;
}

#ifdef ATM_LIN_TIMING
int atmonr_debug_lin_timing(long long* out_host) {
  return cudaMemcpyFromSymbol(out_host, atm::g_lin_timing, sizeof(long long) * 64 * 8) == cudaSuccess ? 0 : -1;
}
int atmonr_debug_lin_timing2(long long* out_host) {
  return cudaMemcpyFromSymbol(out_host, atm::g_lin_timing2, sizeof(long long) * 64 * 8) == cudaSuccess ? 0 : -1;
}
#endif

}  // extern "C"
