// tc_common.cuh -- thin inline-PTX layer over the sm_100a tensor-core path:
// tcgen05.mma (UMMA) with operands in shared memory and accumulators in tensor memory (TMEM),
// tcgen05.ld for the epilogue, mbarrier completion tracking.
//
// Shared-memory operand layout used throughout this library (no swizzle, "interleaved" canonical
// layout): a tile of R rows x C fp16 columns is stored as 8x8 "core matrices" of 128 contiguous
// bytes (8 rows x 16 B); core matrices that are neighbours along the columns are kColStep = 128 B
// apart, neighbours along the rows are (C/8)*128 B apart:
//
//     addr(r, c) = (r/8) * (C/8)*128 + (c/8) * 128 + (r%8) * 16 + (c%8) * 2
//
// The same bytes serve as a K-major operand (rows = M or N, columns = K) and as an MN-major
// operand (rows = K, columns = M or N): only the two strides in the descriptor swap roles.
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

namespace atm {
namespace tc {

constexpr uint32_t kCore = 128;  // bytes of one 8x8 fp16 core matrix

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// byte offset of element (r, c) in a tile with C columns
__device__ __forceinline__ uint32_t tile_off(int r, int c, int C) {
  return (uint32_t)((r >> 3) * (C >> 3) * kCore + (c >> 3) * kCore + (r & 7) * 16 + (c & 7) * 2);
}

// ---- descriptors ------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading and
// stride byte offsets in 16-byte units, descriptor version 1 (Blackwell), no swizzle.
// Every tile address here is a multiple of 16 below 256 KB, so the address field is simply
// saddr >> 4: descriptors of neighbouring tiles / K steps differ by a compile-time constant and
// the issuing warp forms each one with a single add.
__host__ __device__ constexpr uint64_t desc_fields(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
// K-major operand: rows = M/N (groups of 8 rows are (C/8)*128 B apart), K chunks 128 B apart.
__device__ __forceinline__ uint64_t desc_k_major(uint32_t saddr, int C) {
  return desc_fields(kCore, (uint32_t)(C >> 3) * kCore) | (uint64_t)(saddr >> 4);
}
// MN-major operand over the same bytes: rows = K, columns = M/N.
__device__ __forceinline__ uint64_t desc_mn_major(uint32_t saddr, int C) {
  return desc_fields((uint32_t)(C >> 3) * kCore, kCore) | (uint64_t)(saddr >> 4);
}

// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: fp16 A and B, fp32 D.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                       // c_format = F32
         | (0u << 7) | (0u << 10)        // a_format = b_format = F16
         | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM allocation -----------------------------------------------------------------------
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// ---- fences ------------------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() {  // generic-proxy smem writes -> async proxy (UMMA)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- mbarrier ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// The hint lets the hardware park the thread until the phase completes (or the hint expires)
// instead of returning after the short default window (the re-polls were 15 % of the backward
// kernel's issued instructions; they only used otherwise idle issue slots: run time unchanged).
#ifndef ATM_SUSPEND_HINT_NS
#define ATM_SUSPEND_HINT_NS 0x989680
#endif
constexpr uint32_t kSuspendHintNs = ATM_SUSPEND_HINT_NS;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(kSuspendHintNs)
      : "memory");
}

// ---- bulk asynchronous copy global -> shared (TMA engine, no tensor map) ------------------------
// One thread announces `bytes` on the mbarrier and issues the copy; the barrier's phase completes
// when the bytes have landed (and every other expected arrival has happened). Source, destination
// and size are multiples of 16 bytes. The data is written through the async proxy, which is the
// proxy tcgen05.mma reads operands through: no extra fence between the barrier wait and the MMA.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// MN-major operand whose K direction skips row groups: the two 8-row K groups of one MMA are
// `k_group_stride` row groups apart (see issue_dweight_t in ngp_fused.cu).
__device__ __forceinline__ uint64_t desc_mn_major_strided(uint32_t saddr, int C, int k_group_stride) {
  return desc_fields((uint32_t)(C >> 3) * kCore * (uint32_t)k_group_stride, kCore) | (uint64_t)(saddr >> 4);
}

// ---- MMA ---------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T. Called by ALL 32 lanes of one converged warp with identical
// (warp-uniform) operands; elect.sync picks the lane that issues. Predicating the instruction on
// the elect predicate (instead of branching on a lane id around it) lets ptxas keep descriptors in
// uniform registers and emit one UTCHMMA; inside a divergent `if (tid == 0)` it wraps every
// UTCHMMA in an elect-and-retry loop that costs ~150 cycles per MMA (measured, scratch/ubench).
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Make the mbarrier track completion of all MMAs issued so far by the elected lane (same calling
// convention as umma_f16: the whole converged warp).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(smem_u32(bar))
      : "memory");
}

// ---- TMEM -> registers (each thread reads its own lane: 32 lanes per warp, N consecutive columns)
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float (&v)[2]) {
  uint32_t r[2];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
  tmem_ld_wait();
  v[0] = __uint_as_float(r[0]);
  v[1] = __uint_as_float(r[1]);
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> TMEM: 128 consecutive columns of this warp's 32 lanes set to zero
__device__ __forceinline__ void tmem_st_zero128(uint32_t taddr) {
#pragma unroll
  for (int c = 0; c < 128; c += 8)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + c), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TMEM address of (lane quarter of this warp, column)
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int warp, int col) {
  return base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col;
}

// Ask the L2 for `bytes` (a multiple of 16) at the 16-byte aligned global address p; nothing waits on it.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- staging helpers ---------------------------------------------------------------------------
// Row `r` of a tile with C columns: write 8 fp16 values (one 16-byte chunk) at column chunk `cc`.
__device__ __forceinline__ void st_chunk(uint8_t* tile, int r, int cc, int C, uint4 v) {
  *reinterpret_cast<uint4*>(tile + (r >> 3) * (C >> 3) * kCore + cc * kCore + (r & 7) * 16) = v;
}
__device__ __forceinline__ uint4 ld_chunk(const uint8_t* tile, int r, int cc, int C) {
  return *reinterpret_cast<const uint4*>(tile + (r >> 3) * (C >> 3) * kCore + cc * kCore + (r & 7) * 16);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Copy a row-major [R][C] fp16 matrix from global memory into the tile layout (whole CTA).
__device__ __forceinline__ void load_matrix_tile(const __half* __restrict__ g, uint8_t* tile, int R, int C) {
  const int chunks = R * (C >> 3);
  for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
    const int r = i / (C >> 3), cc = i % (C >> 3);
    st_chunk(tile, r, cc, C, *reinterpret_cast<const uint4*>(g + r * C + cc * 8));
  }
}

}  // namespace tc
}  // namespace atm
