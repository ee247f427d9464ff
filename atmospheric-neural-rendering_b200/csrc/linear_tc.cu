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
//
// Tile: 128 rows of X x up to 256 columns of Y per CTA (grid.y walks wider layers: fc9 has 256+V
// outputs), K in chunks of 32. Operand tiles live in shared memory in the no-swizzle core-matrix
// layout of tc_common.cuh; two stages, so the loads and the split of chunk c+1 run underneath the
// twelve MMAs of chunk c (a stage is recycled when the commit of the MMAs that read it has arrived).
// B is split ONCE per step by atmonr_linear_prep (optionally transposed: the input-gradient product
// dX = dY * W is the same kernel on the planes of W^T) and stored in HBM already in tile order, so
// its staging is a plain 16-byte copy; X is split on the fly by the loading threads.
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
constexpr int kStage = 3 * kATile + 3 * kBTile;   // 72 KB
constexpr int kBar = 2 * kStage;                  // two mbarriers
constexpr int kTmemPtr = kBar + 16;
constexpr int kBytes = kTmemPtr + 16;
constexpr uint32_t kTmemCols = 256;
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

// B (n_out, k_in) float32 row-major, or its transpose when `transpose` (then the source is
// (k_in, n_out) row-major) -> planes[(tile * k_chunks + chunk) * 3 + plane][256 x 32 bf16, tile layout]
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
  uint4 hi, mid, lo;
  split8(v, hi, mid, lo);
  const int tile = row / lin::kCols, r = row % lin::kCols, chunk = grp >> 2, cc = grp & 3;
  uint8_t* base = planes + ((size_t)tile * k_chunks + chunk) * 3 * lin::kBTile;
  st_chunk(base, r, cc, lin::kChunk, hi);
  st_chunk(base + lin::kBTile, r, cc, lin::kChunk, mid);
  st_chunk(base + 2 * lin::kBTile, r, cc, lin::kChunk, lo);
}

__global__ void __launch_bounds__(lin::kThreads, 1)
k_linear_tc(const float* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ planes,
            const float* __restrict__ bias, int64_t M, int n_out, int k_in, int k_chunks, int act,
            float* __restrict__ y, int64_t ldy) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + lin::kBar);          // bar[s]: MMAs that read stage s are done
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + lin::kTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<lin::kTmemCols>(tmem_ptr);
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

  const int64_t row0 = (int64_t)blockIdx.x * lin::kRows;
  const int n0 = blockIdx.y * lin::kCols;                                  // first output column of this CTA
  const int n_cols = min(lin::kCols, ((n_out + 15) / 16) * 16 - n0);       // MMA N (multiple of 16)
  const uint32_t idesc = make_idesc_bf16(lin::kRows, n_cols);
  const uint8_t* b_src = planes + (size_t)blockIdx.y * k_chunks * 3 * lin::kBTile;
  const bool vec_ok = (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;

  for (int c = 0; c < k_chunks; ++c) {
    const int s = c & 1;
    uint8_t* stage = smem + s * lin::kStage;
    // the MMAs of chunk c-2 read this stage: wait for their commit (completion number (c>>1)-1 of bar[s])
    if (c >= 2) mbar_wait(bar + s, (uint32_t)(((c >> 1) - 1) & 1));
    // ---- X chunk: 128 rows x 32 columns float32 -> three bf16 planes (2 groups of 8 values per thread)
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int item = tid + it * lin::kThreads;        // 512 items = 128 rows x 4 groups
      const int r = item >> 2, cc = item & 3;
      const int64_t row = row0 + r;
      const int k0 = c * lin::kChunk + cc * 8;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.0f;
      if (row < M) {
        const float* src = x + row * ldx + k0;
        if (vec_ok && k0 + 8 <= k_in) {
          const float4 p = *reinterpret_cast<const float4*>(src), q = *reinterpret_cast<const float4*>(src + 4);
          v[0] = p.x, v[1] = p.y, v[2] = p.z, v[3] = p.w, v[4] = q.x, v[5] = q.y, v[6] = q.z, v[7] = q.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (k0 + j < k_in) v[j] = src[j];
        }
      }
      uint4 hi, mid, lo;
      split8(v, hi, mid, lo);
      st_chunk(stage, r, cc, lin::kChunk, hi);
      st_chunk(stage + lin::kATile, r, cc, lin::kChunk, mid);
      st_chunk(stage + 2 * lin::kATile, r, cc, lin::kChunk, lo);
    }
    // ---- B chunk: three planes, already split and in tile order: 16-byte copies of the rows in use
    {
      const uint4* src = reinterpret_cast<const uint4*>(b_src + (size_t)c * 3 * lin::kBTile);
      uint4* dst = reinterpret_cast<uint4*>(stage + 3 * lin::kATile);
      const int per_plane = n_cols * 4;                 // 16-byte chunks of the rows in use (rows are 64 B)
      for (int p = 0; p < 3; ++p)
        for (int i = tid; i < per_plane; i += lin::kThreads)
          dst[p * (lin::kBTile / 16) + i] = __ldg(src + p * (lin::kBTile / 16) + i);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      const uint32_t a0 = sbase + s * lin::kStage, b0 = a0 + 3 * lin::kATile;
      // (A plane, B plane): hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi
      const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};
#pragma unroll
      for (int k = 0; k < lin::kChunk / 16; ++k) {
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          const uint64_t ad = desc_k_major(a0 + pa[t] * lin::kATile + k * 2 * kCore, lin::kChunk);
          const uint64_t bd = desc_k_major(b0 + pb[t] * lin::kBTile + k * 2 * kCore, lin::kChunk);
          umma_f16(acc, ad, bd, idesc, (c | k | t) != 0 ? 1u : 0u);
        }
      }
      umma_commit(bar + s);
    }
  }
  // every MMA has been issued by warp 0 in order; the last commit covers them all
  {
    const int last = k_chunks - 1;
    mbar_wait(bar + (last & 1), (uint32_t)((last >> 1) & 1));
    tc_fence_after();
  }
  // ---- epilogue: thread (warp w, lane) owns row (w % 4) * 32 + lane, columns (w / 4) * 128 .. + 128
  {
    const int r = (warp & 3) * 32 + (tid & 31);
    const int64_t row = row0 + r;
    const int c_lo = (warp >> 2) * 128;
#pragma unroll 1
    for (int cb = 0; cb < 128; cb += 16) {
      const int col = c_lo + cb;
      if (col >= n_cols) break;                          // warp-uniform
      float v[16];
      tmem_ld16(tmem_addr(acc, warp, col), v);
      if (row < M) {
        float* dst = y + row * ldy + n0 + col;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n0 + col + j;
          if (n < n_out) {
            float o = v[j] + (bias ? bias[n] : 0.0f);
            if (act == 1) o = fmaxf(o, 0.0f);
            dst[j] = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<lin::kTmemCols>(acc);
}

}  // namespace atm

using namespace atm;
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int atmonr_linear_prep(const float* w, int n_out, int k_in, int transpose, void* planes, void* stream) {
  ATM_REQUIRE(w && planes, "atmonr_linear_prep", "null pointer");
  ATM_REQUIRE(n_out > 0 && k_in > 0, "atmonr_linear_prep", "empty matrix");
  const int n_tiles = (n_out + lin::kCols - 1) / lin::kCols, k_chunks = (k_in + lin::kChunk - 1) / lin::kChunk;
  const int64_t total = (int64_t)n_tiles * lin::kCols * k_chunks * 4;
  k_linear_prep<<<grid_for(total, 256), 256, 0, S(stream)>>>(w, n_out, k_in, transpose, k_chunks, n_tiles,
                                                             reinterpret_cast<uint8_t*>(planes));
  ATM_CHECK_LAUNCH("atmonr_linear_prep");
  return 0;
}

int atmonr_linear_fwd_tc(const float* x, int64_t ldx, const void* planes, const float* bias, int64_t M, int n_out,
                         int k_in, int act, float* y, int64_t ldy, void* stream) {
  ATM_REQUIRE(M >= 0 && n_out > 0 && k_in > 0, "atmonr_linear_fwd_tc", "bad shape");
  ATM_REQUIRE(act == 0 || act == 1, "atmonr_linear_fwd_tc", "act must be 0 (none) or 1 (ReLU)");
  if (M == 0) return 0;
  ATM_REQUIRE(x && planes && y, "atmonr_linear_fwd_tc", "null pointer");
  ATM_REQUIRE(ldx >= k_in && ldy >= n_out, "atmonr_linear_fwd_tc", "row stride smaller than the row");
  ATM_REQUIRE((M + lin::kRows - 1) / lin::kRows < (1ll << 31), "atmonr_linear_fwd_tc", "too many rows");
  cudaError_t e = cudaFuncSetAttribute(k_linear_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, lin::kBytes);
  if (e != cudaSuccess) return fail("atmonr_linear_fwd_tc", cudaGetErrorString(e));
  const int n_tiles = (n_out + lin::kCols - 1) / lin::kCols, k_chunks = (k_in + lin::kChunk - 1) / lin::kChunk;
  dim3 grid((unsigned)((M + lin::kRows - 1) / lin::kRows), (unsigned)n_tiles);
  k_linear_tc<<<grid, lin::kThreads, lin::kBytes, S(stream)>>>(x, ldx, reinterpret_cast<const uint8_t*>(planes), bias, M,
                                                               n_out, k_in, k_chunks, act, y, ldy);
  ATM_CHECK_LAUNCH("atmonr_linear_fwd_tc");
  return 0;
}

}  // extern "C"
