// ngp_fused.cu -- tcgen05 (5th-generation tensor core) kernels of libatmonr_b200.
//
//   atmonr_tc_probe        : one 128-row tile through each of the three operand configurations
//                            the fused kernels use (forward, input-gradient, weight-gradient);
//                            parity-tested against a matmul so descriptor mistakes are caught in
//                            isolation.
//   atmonr_ngp_field_fwd_tc: fused radiance field forward (hash grid -> pos_mlp -> SH -> dir_mlp)
//                            with the five dense layers on tcgen05.mma, accumulators in TMEM.
//   atmonr_ngp_field_bwd_tc: its backward: recompute, input gradients and weight gradients on
//                            tcgen05.mma, table gradients scattered with vector REDs.
//
// Thread/row mapping: CTA = 128 threads = one 128-sample tile; thread t owns sample row t, which
// is TMEM lane t of every accumulator, so the epilogue between two layers (ReLU, fp16 pack,
// store as the next layer's A operand) is thread-local.
#include "common.cuh"
#include "hashgrid.cuh"
#include "tc_common.cuh"

namespace atm {

using namespace tc;

// =========================================================================================
// probe
// =========================================================================================
__global__ void __launch_bounds__(128) k_tc_probe(const __half* __restrict__ A, const __half* __restrict__ B,
                                                  int mode, float* __restrict__ D) {
  __shared__ __align__(1024) uint8_t sA[128 * 32 * 2 + 2048];
  __shared__ __align__(1024) uint8_t sB[128 * 32 * 2 + 2048];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (int)sizeof(sA) / 16; i += 128) {
    reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(sB)[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  load_matrix_tile(A, sA, 128, 32);
  load_matrix_tile(B, sB, 128, 32);
  if (warp == 0) tmem_alloc<32>(&tmem_base_s);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    if (mode == 0) {  // D = A[128x32] * W[32x32]^T : both K-major, K = 32
      const uint32_t idesc = make_idesc(128, 32, 0, 0);
      for (int k = 0; k < 2; ++k)
        umma_f16(tmem, desc_k_major(a0 + k * 2 * kCore, 32), desc_k_major(b0 + k * 2 * kCore, 32), idesc, k > 0);
    } else if (mode == 1) {  // D = A[128x32] * W[32x32] : A K-major, B MN-major (rows of W are K)
      const uint32_t idesc = make_idesc(128, 32, 0, 1);
      for (int k = 0; k < 2; ++k)
        umma_f16(tmem, desc_k_major(a0 + k * 2 * kCore, 32), desc_mn_major(b0 + k * 2 * 4 * kCore, 32), idesc, k > 0);
    } else {  // D[m][n] = sum_s A[s][m] * B[s][n] : both MN-major, K = 128 samples
      const uint32_t idesc = make_idesc(128, 32, 1, 1);
      for (int k = 0; k < 8; ++k)
        umma_f16(tmem, desc_mn_major(a0 + k * 2 * 4 * kCore, 32), desc_mn_major(b0 + k * 2 * 4 * kCore, 32), idesc,
                 k > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem_addr(tmem, warp, 0), v);
#pragma unroll
  for (int j = 0; j < 32; ++j) D[tid * 32 + j] = v[j];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tmem);
}


// =========================================================================================
// fused radiance field, forward
// =========================================================================================
namespace fwd {
// shared-memory map (bytes)
constexpr int kW1P = 0;             // pos_mlp layer 0  [32][32]
constexpr int kW2P = kW1P + 2048;   // pos_mlp output   [16][32]
constexpr int kWD1 = kW2P + 1024;   // dir_mlp layer 0  [32][32]
constexpr int kWD2 = kWD1 + 2048;   // dir_mlp layer 1  [32][32]
constexpr int kWD3 = kWD2 + 2048;   // dir_mlp output   [16][32]
constexpr int kA0 = kWD3 + 1024;    // activation tile  [128][32]
constexpr int kA1 = kA0 + 8192;     // activation tile  [128][32]
constexpr int kBar = kA1 + 8192;
constexpr int kTmemPtr = kBar + 8;
constexpr int kBytes = kTmemPtr + 8;
constexpr uint32_t kTmemCols = 64;  // [0,32): hidden accumulator, [32,48): 16-wide outputs
}  // namespace fwd

// all five weight matrices -> tile layout
__device__ __forceinline__ void load_field_weights(uint8_t* smem, const __half* __restrict__ pos_w,
                                                   const __half* __restrict__ dir_w) {
  load_matrix_tile(pos_w, smem + fwd::kW1P, 32, 32);
  load_matrix_tile(pos_w + 1024, smem + fwd::kW2P, 16, 32);
  load_matrix_tile(dir_w, smem + fwd::kWD1, 32, 32);
  load_matrix_tile(dir_w + 1024, smem + fwd::kWD2, 32, 32);
  load_matrix_tile(dir_w + 2048, smem + fwd::kWD3, 16, 32);
}

// store one 32-wide fp16 row (optionally after ReLU) into an activation tile
template <bool RELU>
__device__ __forceinline__ void store_row32(uint8_t* tile, int r, const float (&v)[32]) {
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    uint4 q;
    uint32_t* qp = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = v[cc * 8 + 2 * j], b = v[cc * 8 + 2 * j + 1];
      if (RELU) a = fmaxf(a, 0.0f), b = fmaxf(b, 0.0f);
      qp[j] = pack_h2(a, b);
    }
    st_chunk(tile, r, cc, 32, q);
  }
}
__device__ __forceinline__ void store_row32_h2(uint8_t* tile, int r, const __half2 (&h)[16]) {
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    uint4 q;
    uint32_t* qp = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) qp[j] = *reinterpret_cast<const uint32_t*>(&h[cc * 4 + j]);
    st_chunk(tile, r, cc, 32, q);
  }
}

// dir_mlp input row: [SH2(dir) | pos_out[1..15] | 1.0 x 13] (instant_ngp.py:165-169 + tcnn padding)
__device__ __forceinline__ void dir_input_row(const float* __restrict__ dir, const float (&po)[16], float (&v)[32]) {
  float sh[4];
  sh_degree2(dir[0], dir[1], dir[2], sh);
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = sh[k];
#pragma unroll
  for (int k = 1; k < 16; ++k) v[3 + k] = po[k];
#pragma unroll
  for (int k = 19; k < 32; ++k) v[k] = 1.0f;
}

// One dense layer on the tensor core: acc[tmem_col .. +N) = A_tile[128][32] * W_tile[N][32]^T.
// Called by ONE thread after the CTA-wide barrier that published the A tile.
template <int N>
__device__ __forceinline__ void issue_layer(uint32_t tmem_col_addr, uint32_t a_tile, uint32_t w_tile, uint64_t* bar) {
  tc_fence_after();
  constexpr uint32_t idesc = make_idesc(128, N, 0, 0);
#pragma unroll
  for (int k = 0; k < 2; ++k)
    umma_f16(tmem_col_addr, desc_k_major(a_tile + k * 2 * kCore, 32), desc_k_major(w_tile + k * 2 * kCore, 32), idesc, k);
  umma_commit(bar);
}

// publish this thread's shared-memory writes to the tensor core and join the CTA
__device__ __forceinline__ void publish_and_sync() {
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
}

__global__ void __launch_bounds__(128, 4)
k_field_fwd_tc(atmonr_grid_t g, const __half2* __restrict__ table, const __half* __restrict__ pos_w,
               const __half* __restrict__ dir_w, const float* __restrict__ x01, const float* __restrict__ dirs,
               int64_t M, int N, float* __restrict__ sigma_raw, float* __restrict__ color_raw,
               __half* __restrict__ enc_out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + fwd::kBar);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + fwd::kTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5;
  load_field_weights(smem, pos_w, dir_w);
  if (warp == 0) tmem_alloc<fwd::kTmemCols>(tmem_ptr);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  publish_and_sync();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t acc32 = tmem, acc16 = tmem + 32;
  const uint32_t my32 = tmem_addr(tmem, warp, 0), my16 = tmem_addr(tmem, warp, 32);
  const uint32_t sbase = smem_u32(smem);
  uint8_t* A0 = smem + fwd::kA0;
  uint8_t* A1 = smem + fwd::kA1;
  uint32_t phase = 0;

  for (int64_t tile = blockIdx.x; tile * kTile < M; tile += gridDim.x) {
    const int64_t i = tile * kTile + tid;
    const bool valid = i < M;
    const int64_t j = valid ? i : M - 1;
    // ---- hash-grid encoding -> A0
    {
      const float p[3] = {x01[3 * j], x01[3 * j + 1], x01[3 * j + 2]};
      __half2 enc[16];
      hash_encode_fast<3>(g, table, p, enc);
      store_row32_h2(A0, tid, enc);
      if (enc_out && valid) {
        uint4* dst = reinterpret_cast<uint4*>(enc_out + i * 32);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) dst[cc] = ld_chunk(A0, tid, cc, 32);
      }
    }
    publish_and_sync();
    if (tid == 0) issue_layer<32>(acc32, sbase + fwd::kA0, sbase + fwd::kW1P, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    float v[32];
    tmem_ld32(my32, v);
    store_row32<true>(A1, tid, v);
    publish_and_sync();
    if (tid == 0) issue_layer<16>(acc16, sbase + fwd::kA1, sbase + fwd::kW2P, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    float po[16];
    tmem_ld16(my16, po);
    if (valid) sigma_raw[i] = po[0];
    dir_input_row(dirs + (j / N) * 3, po, v);
    store_row32<false>(A0, tid, v);
    publish_and_sync();
    if (tid == 0) issue_layer<32>(acc32, sbase + fwd::kA0, sbase + fwd::kWD1, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    tmem_ld32(my32, v);
    store_row32<true>(A1, tid, v);
    publish_and_sync();
    if (tid == 0) issue_layer<32>(acc32, sbase + fwd::kA1, sbase + fwd::kWD2, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    tmem_ld32(my32, v);
    store_row32<true>(A0, tid, v);
    publish_and_sync();
    if (tid == 0) issue_layer<16>(acc16, sbase + fwd::kA0, sbase + fwd::kWD3, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    float c[4];
    tmem_ld4(my16, c);
    if (valid) *reinterpret_cast<float4*>(color_raw + 4 * i) = make_float4(c[0], c[1], c[2], c[3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<fwd::kTmemCols>(tmem);
}

}  // namespace atm

using namespace atm;

extern "C" {

// Debug / parity entry point (declared in include/atmonr_b200.h).
int atmonr_tc_probe(const void* a_f16, const void* b_f16, int mode, float* d, void* stream) {
  ATM_REQUIRE(mode >= 0 && mode <= 2, "atmonr_tc_probe", "mode must be 0, 1 or 2");
  k_tc_probe<<<1, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const __half*)a_f16, (const __half*)b_f16, mode, d);
  ATM_CHECK_LAUNCH("atmonr_tc_probe");
  return 0;
}


static int tc_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int check_field_shapes(const atmonr_grid_t* g, const atmonr_mlp_t* pm, const atmonr_mlp_t* dm, const char* name) {
  ATM_REQUIRE(g && g->n_dims == 3 && g->n_feat == 2 && g->n_levels == 16, name, "field needs a 3-D grid with 16 levels x 2 features");
  ATM_REQUIRE(pm && pm->in_pad == 32 && pm->width == 32 && pm->n_hidden == 1 && pm->n_out == 16, name, "pos_mlp must be 32 -> [32] -> 16");
  ATM_REQUIRE(dm && dm->in_pad == 32 && dm->width == 32 && dm->n_hidden == 2 && dm->n_in == 19 && dm->n_out == 4, name, "dir_mlp must be 19 -> [32,32] -> 4");
  return 0;
}

// Same contract as atmonr_ngp_field_fwd; the dense layers run on tcgen05. enc_out (optional,
// (M,32) fp16) receives the encoded features so the backward pass can skip the table gathers.
int atmonr_ngp_field_fwd_tc(const atmonr_grid_t* g, const void* table, const atmonr_mlp_t* pm, const void* pos_w,
                            const atmonr_mlp_t* dm, const void* dir_w, const float* x01, const float* dirs, int64_t B,
                            int N, float* sigma_raw, float* color_raw, void* enc_out, void* stream) {
  if (check_field_shapes(g, pm, dm, "atmonr_ngp_field_fwd_tc")) return -1;
  const int64_t M = B * N;
  if (M == 0) return 0;
  const int64_t tiles = (M + kTile - 1) / kTile;
  const int grid = (int)(tiles < (int64_t)tc_num_sms() * 4 ? tiles : (int64_t)tc_num_sms() * 4);
  k_field_fwd_tc<<<grid, kTile, fwd::kBytes, reinterpret_cast<cudaStream_t>(stream)>>>(
      *g, (const __half2*)table, (const __half*)pos_w, (const __half*)dir_w, x01, dirs, M, N, sigma_raw, color_raw,
      (__half*)enc_out);
  ATM_CHECK_LAUNCH("atmonr_ngp_field_fwd_tc");
  return 0;
}

}  // extern "C"
