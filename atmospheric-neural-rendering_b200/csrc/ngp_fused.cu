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
#include <stdlib.h>

#include "common.cuh"
#include "hashgrid.cuh"
#include "tc_common.cuh"

// debug switches for timing experiments (never defined in the shipped build)
#ifdef ATM_DEBUG_NO_DW
#define ATM_DW if (false)
#else
#define ATM_DW
#endif
#ifdef ATM_DEBUG_NO_SCATTER
#define ATM_SCATTER_ON (invS == 123456.0f)
#else
#define ATM_SCATTER_ON true
#endif

namespace atm {

using namespace tc;

// =========================================================================================
// weight-gradient MMAs with sample-group concatenation
// =========================================================================================
// dW[o][i] = sum_s delta[s][o] * act[s][i] has a tiny output (<= 32 x 32) and a huge K (samples),
// while a tcgen05.mma of M = 128 costs ~40 cycles whatever N <= 64 is (micro-benchmark, round 1): the cost
// is the number of MMAs. Both operands are read MN-major from [rows][C] tiles whose 8-row groups
// are contiguous blocks of (C/8)*128 bytes, so column group C/8 + j of row group g IS column
// group j of row group g + 1: reading "too many" columns concatenates the following row groups
// for free. With the MMA's two K row groups WAYS groups apart,
//     A = act tile  (M = 128: row groups g .. g+3 side by side, 32 columns each),
//     B = delta tile (N = WAYS * KD: row groups g .. g+WAYS-1 side by side),
// D[32 q + i][KD q' + o] = sum over the rows of groups {g+q, g+WAYS+q} x {g+q', g+WAYS+q'}; the
// diagonal blocks q = q' < WAYS hold the gradient of 16 * WAYS samples per MMA (the off-diagonal
// blocks are ignored). The flush sums the diagonal blocks: dW[o][i] = sum_q D[32 q + i][KD q + o].
template <int KD, int WAYS, int ROWS>
__device__ __forceinline__ void issue_dweight_t(uint32_t acc, uint32_t act_tile, uint32_t d_tile, uint32_t accumulate) {
  constexpr uint32_t idesc = make_idesc(128, WAYS * KD, 1, 1);
  constexpr int kGroupsPerMma = 2 * WAYS;
#pragma unroll
  for (int j = 0; j < ROWS / (8 * kGroupsPerMma); ++j)
    umma_f16(acc, desc_mn_major_strided(act_tile + j * kGroupsPerMma * 4 * kCore, 32, WAYS),
             desc_mn_major_strided(d_tile + j * kGroupsPerMma * (KD / 8) * kCore, KD, WAYS), idesc,
             accumulate | (uint32_t)(j > 0));
}
// Add the diagonal blocks of one accumulator into dW (row-major [KD_real rows][32]); called by the
// first WAYS warps of the CTA (warp q owns TMEM lanes 32 q ..).
template <int KD, int WAYS>
__device__ __forceinline__ void flush_dweight_t(uint32_t tmem, int col, int warp, int lane, int rows, float scale,
                                                float* __restrict__ dW) {
  if (warp >= WAYS) return;
  float v[KD];
  if constexpr (KD == 32) {
    tmem_ld32(tmem_addr(tmem, warp, col + KD * warp), v);
  } else {
    tmem_ld16(tmem_addr(tmem, warp, col + KD * warp), v);
  }
#pragma unroll
  for (int o = 0; o < KD; ++o)
    if (o < rows) atomicAdd(dW + o * 32 + lane, v[o] * scale);
}

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
  if (mode == 4) {  // [128][16] tile from the first 16 columns of B
    for (int i = tid; i < 128 * 2; i += 128)
      st_chunk(sB, i >> 1, i & 1, 16, *reinterpret_cast<const uint4*>(B + (i >> 1) * 32 + (i & 1) * 8));
  } else {
    load_matrix_tile(B, sB, 128, 32);
  }
  if (warp == 0) tmem_alloc<64>(&tmem_base_s);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    if (mode == 0) {  // D = A[128x32] * W[32x32]^T : both K-major, K = 32
      const uint32_t idesc = make_idesc(128, 32, 0, 0);
      for (int k = 0; k < 2; ++k)
        umma_f16(tmem, desc_k_major(a0 + k * 2 * kCore, 32), desc_k_major(b0 + k * 2 * kCore, 32), idesc, k > 0);
    } else if (mode == 1) {  // D = A[128x32] * W[32x32] : A K-major, B MN-major (rows of W are K)
      const uint32_t idesc = make_idesc(128, 32, 0, 1);
      for (int k = 0; k < 2; ++k)
        umma_f16(tmem, desc_k_major(a0 + k * 2 * kCore, 32), desc_mn_major(b0 + k * 2 * 4 * kCore, 32), idesc, k > 0);
    } else if (mode == 2) {  // D[m][n] = sum_s A[s][m] * B[s][n] : both MN-major, K = 128 samples
      const uint32_t idesc = make_idesc(128, 32, 1, 1);
      for (int k = 0; k < 8; ++k)
        umma_f16(tmem, desc_mn_major(a0 + k * 2 * 4 * kCore, 32), desc_mn_major(b0 + k * 2 * 4 * kCore, 32), idesc,
                 k > 0);
    } else if (mode == 3) {  // sample-group concatenation, 32-column delta tile (B), two ways
      issue_dweight_t<32, 2, 128>(tmem, a0, b0, 0);
    } else {                 // the same with a 16-column delta tile: B holds [128][16] (b[:, :16] repacked)
      issue_dweight_t<16, 2, 128>(tmem, a0, b0, 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem_addr(tmem, warp, 0), v);
#pragma unroll
  for (int j = 0; j < 32; ++j) D[tid * (mode >= 3 ? 64 : 32) + j] = v[j];
  if (mode >= 3) {
    tmem_ld32(tmem_addr(tmem, warp, 32), v);
#pragma unroll
    for (int j = 0; j < 32; ++j) D[tid * 64 + 32 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
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
constexpr int kA = kWD3 + 1024;     // activation tile  [128][32], rewritten in place layer after layer
constexpr int kLv = kA + 8192;      // level table, 16 x 32 B
constexpr int kBar = kLv + 512;
constexpr int kTmemPtr = kBar + 8;
constexpr int kBytes = kTmemPtr + 8;
// one 32-column accumulator; the 16-wide outputs reuse its first columns (a layer's accumulator
// has been read by every thread before the next layer is issued)
constexpr uint32_t kTmemCols = 32;
#ifndef ATM_FWD_CTAS
#define ATM_FWD_CTAS 10
#endif
constexpr int kCtasPerSm = ATM_FWD_CTAS;  // 10 x 4 warps per SM at <= 48 registers; smem 17 KB and 32 TMEM columns per CTA
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

// two fp32 -> packed fp16 pair (low half = a), optionally through ReLU, in one instruction
template <bool RELU>
__device__ __forceinline__ uint32_t pack_h2_act(float a, float b) {
  uint32_t r;
  if (RELU)
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// store one 32-wide fp16 row (optionally after ReLU) into an activation tile
template <bool RELU>
__device__ __forceinline__ void store_row32(uint8_t* tile, int r, const float (&v)[32]) {
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    uint4 q;
    uint32_t* qp = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) qp[j] = pack_h2_act<RELU>(v[cc * 8 + 2 * j], v[cc * 8 + 2 * j + 1]);
    st_chunk(tile, r, cc, 32, q);
  }
}

// columns [16*half, 16*half+16) of row r
template <bool RELU>
__device__ __forceinline__ void store_half16(uint8_t* tile, int r, int half, const float (&v)[16]) {
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    uint4 q;
    uint32_t* qp = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) qp[j] = pack_h2_act<RELU>(v[cc * 8 + 2 * j], v[cc * 8 + 2 * j + 1]);
    st_chunk(tile, r, 2 * half + cc, 32, q);
  }
}
// hidden-layer epilogue: this thread's 32 accumulator columns -> ReLU -> fp16 row of the tile, in
// two 16-column halves (a 32-column tcgen05.ld would pin 32 registers at once)
__device__ __forceinline__ void relu_acc_to_tile(uint32_t taddr, uint8_t* tile, int r) {
  float h[16];
  tmem_ld16(taddr, h);
  store_half16<true>(tile, r, 0, h);
  tmem_ld16(taddr + 16, h);
  store_half16<true>(tile, r, 1, h);
}

// ---- hash-grid encoding of one point, software-pipelined over the levels ------------------------
// level_issue() computes the 8 corner entries of a level and issues the 8 table gathers;
// level_finish() interpolates. The loop issues level l+1 before it consumes level l, so a thread
// always has 8-16 gathers in flight. Same arithmetic as hash_encode().
__device__ __forceinline__ void level_issue(const LevelRow& L, const __half2* __restrict__ table, const float (&p)[3],
                                            uint32_t (&v)[8], float (&frac)[3], uint64_t keep) {
  uint32_t cell[3], e[8];
  grid_cell<3>(p, L.scale, cell, frac);
  const uint32_t* base = reinterpret_cast<const uint32_t*>(table + L.offset);
#ifdef ATM_L2_HINTS
#define ATM_GATHER(ptr) ldg_u32_hint(ptr, keep)
#else
#define ATM_GATHER(ptr) __ldg(ptr)
#endif
#ifndef ATM_NO_DENSE_PAIRS
  if (L.hashed == 0u) {
    // Dense level: the +1 corner in x is the NEXT entry in memory, so the 8 gathers need 4 addresses
    // (the second load of a pair is an immediate offset). Same entries as corner_entries3: corner c =
    // e0 + (c & 1) + (c & 2 ? stride1 : 0) + (c & 4 ? stride2 : 0); the all-+1 corner is the largest,
    // so one compare proves that no corner wraps (cells below 2^16 cannot overflow uint32).
    const uint32_t e0 = cell[0] + cell[1] * L.stride1 + cell[2] * L.stride2;
    if ((cell[0] | cell[1] | cell[2]) < 65536u && e0 + 1u + L.stride1 + L.stride2 < L.size) {
      const uint32_t* q0 = entry_ptr(base, e0);
      const uint32_t* q1 = entry_ptr(base, e0 + L.stride1);
      const uint32_t* q2 = entry_ptr(base, e0 + L.stride2);
      const uint32_t* q3 = entry_ptr(base, e0 + L.stride1 + L.stride2);
      v[0] = ATM_GATHER(q0), v[1] = ATM_GATHER(q0 + 1);
      v[2] = ATM_GATHER(q1), v[3] = ATM_GATHER(q1 + 1);
      v[4] = ATM_GATHER(q2), v[5] = ATM_GATHER(q2 + 1);
      v[6] = ATM_GATHER(q3), v[7] = ATM_GATHER(q3 + 1);
      return;
    }
  }
#endif
  corner_entries3(L, cell, e);
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = ATM_GATHER(entry_ptr(base, e[c]));
#undef ATM_GATHER
}
__device__ __forceinline__ uint32_t level_finish(const uint32_t (&v)[8], const float (&frac)[3]) {
  float w[8];
  corner_weights3(frac, w);
  __half2 h[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) h[c] = *reinterpret_cast<const __half2*>(&v[c]);
  const __half2 acc = interp_corners<8>(h, w);
  return *reinterpret_cast<const uint32_t*>(&acc);
}
__device__ __forceinline__ void encode_to_tile(const LevelRow* __restrict__ lv, const __half2* __restrict__ table,
                                               const float (&p)[3], uint8_t* tile, int r, uint64_t keep) {
  uint32_t vn[8];
  float fn[3];
  level_issue(lv[0], table, p, vn, fn, keep);
  uint8_t* row = tile + tile_off(r, 0, 32);
  // (an explicit two-register-set version of this loop measured 4 % slower than the copies below)
#pragma unroll 2
  for (int l = 0; l < ATMONR_MAX_LEVELS; ++l) {
    uint32_t vc[8];
    float fc[3];
#pragma unroll
    for (int c = 0; c < 8; ++c) vc[c] = vn[c];
    fc[0] = fn[0], fc[1] = fn[1], fc[2] = fn[2];
    if (l + 1 < ATMONR_MAX_LEVELS) level_issue(lv[l + 1], table, p, vn, fn, keep);
    // features 2l, 2l+1 of row r: chunk l/4 (128 B apart), 4 bytes per level inside the chunk
    *reinterpret_cast<uint32_t*>(row + (l >> 2) * kCore + (l & 3) * 4) = level_finish(vc, fc);
  }
}

// ---- the same loop for the forward / extraction kernels, trimmed for instruction issue -----------
// Those kernels are bound by instruction issue, and ~20 of the ~120 instructions a level costs were
// address bookkeeping: the shared-memory window base rebuilt per level (the level table and the tile
// were reached through generic pointers), the row's tile offset rebuilt from the thread index per
// level, the level's table base formed from the kernel parameter and an offset per level. Here the
// level table holds the 64-bit base pointer of every level, shared memory is addressed with 32-bit
// shared-window addresses formed once per tile, and the row address is kept in a register.
// Arithmetic, corner order and entries are those of level_issue() / level_finish(): bit-identical.
struct FwdLevel {  // 32 bytes per level
  float scale;
  uint32_t hashed, size, stride1;  // `hashed` as in LevelRow
  uint32_t stride2, pad;
  const uint32_t* base;  // first entry of the level in the fp16 table (one entry = 4 bytes)
};
static_assert(sizeof(FwdLevel) == sizeof(LevelRow), "the forward kernels keep FwdLevel rows where LevelRow rows were");

__device__ __forceinline__ void load_fwd_levels(const atmonr_grid_t& g, const __half2* __restrict__ table, FwdLevel* rows) {
  if (threadIdx.x < ATMONR_MAX_LEVELS) {
    const int l = threadIdx.x;
    FwdLevel r;
    r.scale = g.scale[l];
    r.size = g.size[l];
    uint64_t dense = 1;
    for (int k = 0; k < g.n_dims; ++k) dense *= g.res[l];
    r.hashed = (dense > (uint64_t)r.size ? 1u : 0u) | (((r.size & (r.size - 1u)) == 0u) ? 2u : 0u);
    r.stride1 = g.res[l];
    r.stride2 = g.res[l] * g.res[l];
    r.pad = 0;
    r.base = reinterpret_cast<const uint32_t*>(table + g.offset[l]);
    rows[l] = r;
  }
}

__device__ __forceinline__ FwdLevel lds_fwd_level(uint32_t saddr) {
  FwdLevel L;
  uint32_t sc, lo, hi;
  // volatile: must not be hoisted above the barrier that publishes the table
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(sc), "=r"(L.hashed), "=r"(L.size), "=r"(L.stride1) : "r"(saddr));
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+16];" : "=r"(L.stride2), "=r"(L.pad), "=r"(lo), "=r"(hi) : "r"(saddr));
  L.scale = __uint_as_float(sc);
  L.base = reinterpret_cast<const uint32_t*>(((uint64_t)hi << 32) | lo);
  return L;
}

__device__ __forceinline__ void level_issue_s(uint32_t level_saddr, const float (&p)[3], uint32_t (&v)[8],
                                              float (&frac)[3], uint64_t keep) {
  const FwdLevel L = lds_fwd_level(level_saddr);
  uint32_t cell[3], e[8];
  grid_cell<3>(p, L.scale, cell, frac);
#ifdef ATM_L2_HINTS
#define ATM_GATHER(ptr) ldg_u32_hint(ptr, keep)
#else
#define ATM_GATHER(ptr) __ldg(ptr)
#endif
  if (L.hashed == 0u) {  // dense level: 4 addresses for the 8 gathers (see level_issue)
    const uint32_t e0 = cell[0] + cell[1] * L.stride1 + cell[2] * L.stride2;
    if ((cell[0] | cell[1] | cell[2]) < 65536u && e0 + 1u + L.stride1 + L.stride2 < L.size) {
      const uint32_t* q0 = entry_ptr(L.base, e0);
      const uint32_t* q1 = entry_ptr(L.base, e0 + L.stride1);
      const uint32_t* q2 = entry_ptr(L.base, e0 + L.stride2);
      const uint32_t* q3 = entry_ptr(L.base, e0 + L.stride1 + L.stride2);
      v[0] = ATM_GATHER(q0), v[1] = ATM_GATHER(q0 + 1);
      v[2] = ATM_GATHER(q1), v[3] = ATM_GATHER(q1 + 1);
      v[4] = ATM_GATHER(q2), v[5] = ATM_GATHER(q2 + 1);
      v[6] = ATM_GATHER(q3), v[7] = ATM_GATHER(q3 + 1);
      return;
    }
  }
  LevelRow R;
  R.hashed = L.hashed, R.size = L.size, R.stride1 = L.stride1, R.stride2 = L.stride2;
  corner_entries3(R, cell, e);
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = ATM_GATHER(entry_ptr(L.base, e[c]));
#undef ATM_GATHER
}

// lv: FwdLevel rows in shared memory; tile: the activation tile; r: this thread's row
__device__ __forceinline__ void encode_to_tile_s(const FwdLevel* lv, const float (&p)[3], uint8_t* tile, int r,
                                                 uint64_t keep) {
  const uint32_t lv_s = smem_u32(lv);
  uint32_t row_s = smem_u32(tile) + tile_off(r, 0, 32);
  asm volatile("" : "+r"(row_s));  // opaque: held in a register instead of being rebuilt from the thread index per level
  // two register sets, two levels per iteration: level l + 1 is issued before level l is interpolated
  uint32_t va[8], vb[8];
  float fa[3], fb[3];
  constexpr uint32_t kRow = (uint32_t)sizeof(FwdLevel);
  level_issue_s(lv_s, p, va, fa, keep);
#pragma unroll 1
  for (int l = 0; l < ATMONR_MAX_LEVELS; l += 2) {
    level_issue_s(lv_s + (uint32_t)(l + 1) * kRow, p, vb, fb, keep);
    // features 2l, 2l+1 of row r: chunk l/4 (128 B apart), 4 bytes per level inside the chunk
    const uint32_t at = row_s + (uint32_t)((l >> 2) * kCore + (l & 3) * 4);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(at), "r"(level_finish(va, fa)) : "memory");
    if (l + 2 < ATMONR_MAX_LEVELS) level_issue_s(lv_s + (uint32_t)(l + 2) * kRow, p, va, fa, keep);
    asm volatile("st.shared.b32 [%0+4], %1;" ::"r"(at), "r"(level_finish(vb, fb)) : "memory");
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

__global__ void __launch_bounds__(128, fwd::kCtasPerSm)
k_field_fwd_tc(atmonr_grid_t g, const __half2* __restrict__ table, const __half* __restrict__ pos_w,
               const __half* __restrict__ dir_w, const float* __restrict__ x01, const float* __restrict__ dirs,
               int64_t M, int N, float* __restrict__ sigma_raw, float* __restrict__ color_raw,
               __half* __restrict__ enc_out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + fwd::kBar);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + fwd::kTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5;
  FwdLevel* lv = reinterpret_cast<FwdLevel*>(smem + fwd::kLv);
  load_fwd_levels(g, table, lv);
  load_field_weights(smem, pos_w, dir_w);
  if (warp == 0) tmem_alloc<fwd::kTmemCols>(tmem_ptr);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  publish_and_sync();
  tc_fence_after();
  const uint32_t acc = *tmem_ptr;
  const uint32_t mine = tmem_addr(acc, warp, 0);
  const uint32_t sbase = smem_u32(smem), sa = sbase + fwd::kA;
  uint8_t* A = smem + fwd::kA;
  uint32_t phase = 0;
  const uint64_t keep_pol = l2_policy_keep(), stream_pol = l2_policy_stream();

  for (int64_t tile = blockIdx.x; tile * kTile < M; tile += gridDim.x) {
    const int64_t i = tile * kTile + tid;
    const bool valid = i < M;
    const int64_t j = valid ? i : M - 1;
    // ---- hash-grid encoding -> A (the previous tile's last MMA has been waited for)
    {
#ifdef ATM_L2_HINTS
      const float p[3] = {ldg_f32_hint(x01 + 3 * j, stream_pol), ldg_f32_hint(x01 + 3 * j + 1, stream_pol),
                          ldg_f32_hint(x01 + 3 * j + 2, stream_pol)};
#else
      const float p[3] = {x01[3 * j], x01[3 * j + 1], x01[3 * j + 2]};
#endif
      encode_to_tile_s(lv, p, A, tid, keep_pol);
      if (enc_out && valid) {
        uint4* dst = reinterpret_cast<uint4*>(enc_out + i * 32);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
#ifdef ATM_L2_HINTS
          stg_u128_hint(dst + cc, ld_chunk(A, tid, cc, 32), stream_pol);
#else
          dst[cc] = ld_chunk(A, tid, cc, 32);
#endif
        }
      }
    }
    publish_and_sync();
    if (warp == 0) issue_layer<32>(acc, sa, sbase + fwd::kW1P, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    relu_acc_to_tile(mine, A, tid);
    publish_and_sync();
    if (warp == 0) issue_layer<16>(acc, sa, sbase + fwd::kW2P, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    {
      // dir_mlp input row: [SH2(dir) | pos_out[1..15] | 1.0 x 13] (instant_ngp.py:165-169 + tcnn padding)
      float po[16], h[16];
      tmem_ld16(mine, po);
      if (valid) sigma_raw[i] = po[0];
      const float* dir = dirs + (size_t)((uint32_t)j / (uint32_t)N) * 3;
      float sh[4];
      sh_degree2(dir[0], dir[1], dir[2], sh);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = sh[k];
#pragma unroll
      for (int k = 4; k < 16; ++k) h[k] = po[k - 3];
      store_half16<false>(A, tid, 0, h);
#pragma unroll
      for (int k = 0; k < 3; ++k) h[k] = po[13 + k];
#pragma unroll
      for (int k = 3; k < 16; ++k) h[k] = 1.0f;
      store_half16<false>(A, tid, 1, h);
    }
    publish_and_sync();
    if (warp == 0) issue_layer<32>(acc, sa, sbase + fwd::kWD1, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    relu_acc_to_tile(mine, A, tid);
    publish_and_sync();
    if (warp == 0) issue_layer<32>(acc, sa, sbase + fwd::kWD2, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    relu_acc_to_tile(mine, A, tid);
    publish_and_sync();
    if (warp == 0) issue_layer<16>(acc, sa, sbase + fwd::kWD3, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    float c[4];
    tmem_ld4(mine, c);
    if (valid) *reinterpret_cast<float4*>(color_raw + 4 * i) = make_float4(c[0], c[1], c[2], c[3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<fwd::kTmemCols>(acc);
}


// =========================================================================================
// extraction: float64 scene point -> geodetic preprocessing -> hash grid -> pos_mlp -> max(sigma, 0)
// =========================================================================================
// instant_ngp.py:208-247 with the two dense layers of pos_mlp on tcgen05: the forward kernel above
// without the colour branch, fed from float64 points (scripts/extract.py:205) through the float64
// flavour of the preprocessor. Same shared-memory map; 6 CTAs per SM (the FP64 chain needs registers).
__global__ void __launch_bounds__(128, 6)
k_extract_sigma_tc(atmonr_frame_t f, GeoFrame gf, atmonr_grid_t g, const __half2* __restrict__ table,
                   const __half* __restrict__ pos_w, const double* __restrict__ pts, int64_t n,
                   float alt_compress, float* __restrict__ sigma) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + fwd::kBar);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + fwd::kTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5;
  FwdLevel* lv = reinterpret_cast<FwdLevel*>(smem + fwd::kLv);
  load_fwd_levels(g, table, lv);
  load_matrix_tile(pos_w, smem + fwd::kW1P, 32, 32);
  load_matrix_tile(pos_w + 1024, smem + fwd::kW2P, 16, 32);
  if (warp == 0) tmem_alloc<fwd::kTmemCols>(tmem_ptr);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  publish_and_sync();
  tc_fence_after();
  const uint32_t acc = *tmem_ptr;
  const uint32_t mine = tmem_addr(acc, warp, 0);
  const uint32_t sbase = smem_u32(smem), sa = sbase + fwd::kA;
  uint8_t* A = smem + fwd::kA;
  uint32_t phase = 0;
  const uint64_t keep_pol = l2_policy_keep();
  for (int64_t tile = blockIdx.x; tile * kTile < n; tile += gridDim.x) {
    const int64_t i = tile * kTile + tid;
    const bool valid = i < n;
    const int64_t j = valid ? i : n - 1;
    {
      double c0 = pts[3 * j], c1 = pts[3 * j + 1], c2 = pts[3 * j + 2];
      if (f.enabled) preprocess_f64(f, gf, c0, c1, c2, c0, c1, c2);
      // instant_ngp.py:224-233 stay in float64; tcnn casts its input to float32
      const float p[3] = {(float)((c0 + 1.0) / 2.0), (float)((c1 + 1.0) / 2.0),
                          (float)(((c2 + 1.0) / 2.0) / (double)alt_compress)};
      encode_to_tile_s(lv, p, A, tid, keep_pol);
    }
    publish_and_sync();
    if (warp == 0) issue_layer<32>(acc, sa, sbase + fwd::kW1P, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    relu_acc_to_tile(mine, A, tid);
    publish_and_sync();
    if (warp == 0) issue_layer<16>(acc, sa, sbase + fwd::kW2P, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    float po[4];
    tmem_ld4(mine, po);
    if (valid) sigma[i] = fmaxf(po[0], 0.0f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<fwd::kTmemCols>(acc);
}

// =========================================================================================
// fused radiance field, backward
// =========================================================================================
namespace bwd {
// shared-memory map (bytes); the 16-column tiles come first and a pad closes the map so that the
// 128-row MN-major reads of the weight-gradient MMAs (which run past a 16/32-column tile) stay
// inside the allocation.
constexpr int kW = 0;                 // the five weight tiles, same order as fwd:: (8192 B)
constexpr int kDO = 8192;             // dL/d(dir_mlp out)  [128][16]
constexpr int kDP = kDO + 4096;       // dL/d(pos_mlp out)  [128][16]
constexpr int kENC = kDP + 4096;      // activations [128][32] ...
constexpr int kH = kENC + 8192;
constexpr int kDIN = kH + 8192;
constexpr int kH1 = kDIN + 8192;
constexpr int kH2 = kH1 + 8192;
constexpr int kDA = kH2 + 8192;       // delta tiles [128][32]
constexpr int kDB = kDA + 8192;
constexpr int kPad = kDB + 8192;
constexpr int kLv = kPad + 2048;      // level table
constexpr int kBar = kLv + 512;
constexpr int kBar2 = kBar + 8;
constexpr int kTmemPtr = kBar2 + 8;
constexpr int kBytes = kTmemPtr + 8;
constexpr uint32_t kTmemCols = 256;
// TMEM columns
constexpr int cAcc32 = 0, cAcc16 = 32, cDWd3 = 64, cDWd2 = 96, cDWd1 = 128, cDW2p = 160, cDW1p = 192;
}  // namespace bwd

// delta[128][KD] (K-major A) times W (read MN-major: rows = K = out, cols = N = in = 32) -> acc32
template <int KD>
__device__ __forceinline__ void issue_dinput(uint32_t acc, uint32_t d_tile, uint32_t w_tile) {
  constexpr uint32_t idesc = make_idesc(128, 32, 0, 1);
#pragma unroll
  for (int k = 0; k < KD / 16; ++k)
    umma_f16(acc, desc_k_major(d_tile + k * 2 * kCore, KD), desc_mn_major(w_tile + k * 2 * 4 * kCore, 32), idesc, k);
}
// dW[o][i] += sum_s delta[s][o] * act[s][i]: both tiles read MN-major, K = 128 samples
template <int KD>
__device__ __forceinline__ void issue_dweight(uint32_t acc, uint32_t d_tile, uint32_t a_tile, uint32_t accumulate) {
  constexpr uint32_t idesc = make_idesc(128, 32, 1, 1);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    umma_f16(acc, desc_mn_major(d_tile + k * 2 * (KD / 8) * kCore, KD), desc_mn_major(a_tile + k * 2 * 4 * kCore, 32),
             idesc, accumulate | (uint32_t)(k > 0));
}

// dst row r = fp16(v) where this thread's row of the activation tile `act` is positive, else 0:
// the ReLU derivative applied in the packed-half domain (one convert, one compare-to-mask and one
// AND per pair of values; the values equal "mask in fp32, then round")
__device__ __forceinline__ void store_row32_masked(uint8_t* dst, const uint8_t* act, int r, const float (&v)[32]) {
  const __half2 zero = __floats2half2_rn(0.0f, 0.0f);
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const uint4 a = ld_chunk(act, r, cc, 32);
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    uint4 q;
    uint32_t* qp = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      qp[j] = pack_h2_act<false>(v[cc * 8 + 2 * j], v[cc * 8 + 2 * j + 1]) & __hgt2_mask(ah[j], zero);
    st_chunk(dst, r, cc, 32, q);
  }
}

__device__ __forceinline__ void store_row16(uint8_t* tile, int r, const float (&v)[16]) {
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    uint4 q;
    uint32_t* qp = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) qp[j] = pack_h2(v[cc * 8 + 2 * j], v[cc * 8 + 2 * j + 1]);
    st_chunk(tile, r, cc, 16, q);
  }
}

__global__ void __launch_bounds__(128, 2)
k_field_bwd_tc(atmonr_grid_t g, const __half2* __restrict__ table, const __half* __restrict__ pos_w,
               const __half* __restrict__ dir_w, const float* __restrict__ x01, const float* __restrict__ dirs,
               const __half* __restrict__ enc_in, const float* __restrict__ dsigma_raw,
               const float* __restrict__ dcolor_raw, const float* __restrict__ grad_absmax, int64_t M, int N,
               float* __restrict__ dtable, float* __restrict__ dpos_w, float* __restrict__ ddir_w) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + bwd::kBar);
  uint64_t* bar2 = reinterpret_cast<uint64_t*>(smem + bwd::kBar2);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + bwd::kTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5;
  LevelRow* lv = reinterpret_cast<LevelRow*>(smem + bwd::kLv);
  load_level_table(g, lv);
  // the pad is read (as don't-care rows) by the weight-gradient MMAs: keep it finite
  for (int i = tid; i < 2048 / 16; i += 128) reinterpret_cast<uint4*>(smem + bwd::kPad)[i] = make_uint4(0, 0, 0, 0);
  load_field_weights(smem + bwd::kW, pos_w, dir_w);
  if (warp == 0) tmem_alloc<bwd::kTmemCols>(tmem_ptr);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_mbar_init();
  }
  publish_and_sync();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t my32 = tmem_addr(tmem, warp, bwd::cAcc32), my16 = tmem_addr(tmem, warp, bwd::cAcc16);
  uint32_t phase2 = 0;
  const uint32_t sb = smem_u32(smem);
  // power-of-two scale that lifts the incoming gradients into fp16 range
  float S = 1.0f;
  if (grad_absmax) {
    const float amax = *grad_absmax;
    if (amax > 0.0f && amax < INFINITY) S = exp2f(fminf(fmaxf(floorf(log2f(2048.0f / amax)), -60.0f), 60.0f));
  }
  const float invS = 1.0f / S;
  uint32_t phase = 0, seen_tile = 0;

  for (int64_t tile = blockIdx.x; tile * kTile < M; tile += gridDim.x, seen_tile = 1) {
    const int64_t i = tile * kTile + tid;
    const bool valid = i < M;
    const int64_t j = valid ? i : M - 1;
    const float p[3] = {x01[3 * j], x01[3 * j + 1], x01[3 * j + 2]};
    // the previous tile's weight-gradient MMAs still read its buffers: wait before overwriting
    if (seen_tile) mbar_wait(bar2, phase2), phase2 ^= 1;
    // ---------------- recompute the activations ----------------
    if (enc_in) {
      const uint4* src = reinterpret_cast<const uint4*>(enc_in + j * 32);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) st_chunk(smem + bwd::kENC, tid, cc, 32, src[cc]);
    } else {
      encode_to_tile(lv, table, p, smem + bwd::kENC, tid, l2_policy_keep());
    }
    publish_and_sync();
    if (warp == 0) issue_layer<32>(tmem + bwd::cAcc32, sb + bwd::kENC, sb + bwd::kW + fwd::kW1P, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    float v[32];
    tmem_ld32(my32, v);
    store_row32<true>(smem + bwd::kH, tid, v);
    publish_and_sync();
    if (warp == 0) issue_layer<16>(tmem + bwd::cAcc16, sb + bwd::kH, sb + bwd::kW + fwd::kW2P, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    {
      float po[16];
      tmem_ld16(my16, po);
      dir_input_row(dirs + (size_t)((uint32_t)j / (uint32_t)N) * 3, po, v);
    }
    store_row32<false>(smem + bwd::kDIN, tid, v);
    publish_and_sync();
    if (warp == 0) issue_layer<32>(tmem + bwd::cAcc32, sb + bwd::kDIN, sb + bwd::kW + fwd::kWD1, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    tmem_ld32(my32, v);
    store_row32<true>(smem + bwd::kH1, tid, v);
    publish_and_sync();
    if (warp == 0) issue_layer<32>(tmem + bwd::cAcc32, sb + bwd::kH1, sb + bwd::kW + fwd::kWD2, bar);
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    tmem_ld32(my32, v);
    store_row32<true>(smem + bwd::kH2, tid, v);
    // ---------------- backward ----------------
    {
      float dout[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) dout[k] = 0.0f;
      if (valid) {
        const float4 dc = *reinterpret_cast<const float4*>(dcolor_raw + 4 * i);
        dout[0] = dc.x * S, dout[1] = dc.y * S, dout[2] = dc.z * S, dout[3] = dc.w * S;
      }
      store_row16(smem + bwd::kDO, tid, dout);
    }
    publish_and_sync();
    if (warp == 0) {
      tc_fence_after();
      issue_dinput<16>(tmem + bwd::cAcc32, sb + bwd::kDO, sb + bwd::kW + fwd::kWD3);
      umma_commit(bar);  // the next epilogue only waits for the input gradient ...
      issue_dweight<16>(tmem + bwd::cDWd3, sb + bwd::kDO, sb + bwd::kH2, seen_tile);  // ... this overlaps it
    }
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    tmem_ld32(my32, v);
    store_row32_masked(smem + bwd::kDA, smem + bwd::kH2, tid, v);
    publish_and_sync();
    if (warp == 0) {
      tc_fence_after();
      issue_dinput<32>(tmem + bwd::cAcc32, sb + bwd::kDA, sb + bwd::kW + fwd::kWD2);
      umma_commit(bar);  // the next epilogue only waits for the input gradient ...
      issue_dweight<32>(tmem + bwd::cDWd2, sb + bwd::kDA, sb + bwd::kH1, seen_tile);  // ... this overlaps it
    }
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    tmem_ld32(my32, v);
    store_row32_masked(smem + bwd::kDB, smem + bwd::kH1, tid, v);
    publish_and_sync();
    if (warp == 0) {
      tc_fence_after();
      issue_dinput<32>(tmem + bwd::cAcc32, sb + bwd::kDB, sb + bwd::kW + fwd::kWD1);
      umma_commit(bar);  // the next epilogue only waits for the input gradient ...
      issue_dweight<32>(tmem + bwd::cDWd1, sb + bwd::kDB, sb + bwd::kDIN, seen_tile);  // ... this overlaps it
    }
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    tmem_ld32(my32, v);   // dL/d(dir_mlp input): columns 4..18 are the pos_mlp features
    {
      float dpo[16];
      dpo[0] = valid ? dsigma_raw[i] * S : 0.0f;
#pragma unroll
      for (int k = 1; k < 16; ++k) dpo[k] = v[3 + k];
      store_row16(smem + bwd::kDP, tid, dpo);
    }
    publish_and_sync();
    if (warp == 0) {
      tc_fence_after();
      issue_dinput<16>(tmem + bwd::cAcc32, sb + bwd::kDP, sb + bwd::kW + fwd::kW2P);
      umma_commit(bar);  // the next epilogue only waits for the input gradient ...
      issue_dweight<16>(tmem + bwd::cDW2p, sb + bwd::kDP, sb + bwd::kH, seen_tile);  // ... this overlaps it
    }
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    tmem_ld32(my32, v);
    store_row32_masked(smem + bwd::kDA, smem + bwd::kH, tid, v);
    publish_and_sync();
    if (warp == 0) {
      tc_fence_after();
      issue_dinput<32>(tmem + bwd::cAcc32, sb + bwd::kDA, sb + bwd::kW + fwd::kW1P);
      umma_commit(bar);  // the next epilogue only waits for the input gradient ...
      issue_dweight<32>(tmem + bwd::cDW1p, sb + bwd::kDA, sb + bwd::kENC, seen_tile);  // ... this overlaps it
      umma_commit(bar2);  // tile boundary: every MMA that reads this tile's buffers
    }
    mbar_wait(bar, phase), phase ^= 1;
    tc_fence_after();
    // dL/d(encoded features) (scaled by S) sits in TMEM columns [0,32): scatter level by level
#pragma unroll 2
    for (int l = 0; l < ATMONR_MAX_LEVELS; ++l) {
      float d[2];
      tmem_ld2(my32 + 2 * l, d);
      const float d0 = d[0] * invS, d1 = d[1] * invS;
#ifdef ATM_SCATTER_AGGREGATED
      scatter_level_aggregated(lv[l], dtable, p, d0, d1, valid);
#else
      if (valid && (d0 != 0.0f || d1 != 0.0f)) {
        uint32_t e[8];
        float w[8];
        level_corners3(lv[l], p, e, w);
        float* base = dtable + 2 * (size_t)lv[l].offset;
#pragma unroll
        for (int c = 0; c < 8; ++c) red_add_f32x2(base + 2 * (size_t)e[c], w[c] * d0, w[c] * d1);
      }
#endif
    }
  }

  // ---------------- flush the weight gradients (TMEM lanes = output neuron) ----------------
  if (seen_tile) mbar_wait(bar2, phase2);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && seen_tile) {
    const int o = tid;  // lane = row of the weight matrix
    float w[32];
    tmem_ld32(tmem_addr(tmem, 0, bwd::cDW1p), w);
#pragma unroll
    for (int c = 0; c < 32; ++c) atomicAdd(dpos_w + o * 32 + c, w[c] * invS);
    tmem_ld32(tmem_addr(tmem, 0, bwd::cDW2p), w);
    if (o < 16) {
#pragma unroll
      for (int c = 0; c < 32; ++c) atomicAdd(dpos_w + 1024 + o * 32 + c, w[c] * invS);
    }
    tmem_ld32(tmem_addr(tmem, 0, bwd::cDWd1), w);
#pragma unroll
    for (int c = 0; c < 32; ++c) atomicAdd(ddir_w + o * 32 + c, w[c] * invS);
    tmem_ld32(tmem_addr(tmem, 0, bwd::cDWd2), w);
#pragma unroll
    for (int c = 0; c < 32; ++c) atomicAdd(ddir_w + 1024 + o * 32 + c, w[c] * invS);
    tmem_ld32(tmem_addr(tmem, 0, bwd::cDWd3), w);
    if (o < 16) {
#pragma unroll
      for (int c = 0; c < 32; ++c) atomicAdd(ddir_w + 2048 + o * 32 + c, w[c] * invS);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<bwd::kTmemCols>(tmem);
}


// =========================================================================================
// fused radiance field, backward, FOUR independent 128-row pipelines per CTA (one CTA per SM)
// =========================================================================================
// Same arithmetic as k_field_bwd_tc (the 128-row kernel above, which re-gathers the table when the
// forward's encoding cache is absent); this one reads the cached encoding and is built for throughput.
//
// A tile (128 sample rows) goes through 9 tensor-core round trips: 4 recompute layers (activations
// are recomputed from the cached encoding, not stored) and 5 backward stages, each = "epilogue warps
// write an operand tile -> named barrier -> issuer warp issues the input-gradient MMAs, commits, then
// issues the stage's weight-gradient MMAs (which run while the epilogue warps already process the
// committed result) -> mbarrier -> tcgen05.ld". Shared memory is reused along the chain:
//   X   : encoded features (until layer 0 is done)  -> dL/dh2          -> fp32 staging of the scatter
//   H   : pos_mlp hidden                                                -> fp32 staging of the scatter
//   DIN : dir_mlp input                              -> encoded features again (for dW of layer 0)
//   H1  : dir_mlp hidden 0                           -> dL/dh
//   H2  : dir_mlp hidden 1                           -> dL/dh1
//   DO  : dL/d(dir_mlp out)                          -> dL/d(pos_mlp out)
// A buffer is only overwritten after the mbarrier wait that proves its last MMA reader is done
// (tcgen05.commit covers every MMA issued before it by the same thread).
//
// One round trip is ~1400 cycles with ~100 instructions of work per warp in it, so what matters is how
// many INDEPENDENT tiles an SM has in flight and how freely they drift apart. Measured (round 2, B200,
// 2^18 rays x 1024): two CTAs per SM with 256-row tiles (two halves in lock-step behind one issuer
// warp): 32.4 ms, of which 23.0 ms is the recompute + input-gradient chain alone, +4 ms weight-gradient
// MMAs, +5.5 ms scatter. This kernel: ONE CTA per SM runs four 128-row pipelines (pipeline p = epilogue
// warps 4p .. 4p+3, TMEM lanes = its rows, accumulator columns 32 p ..) and FOUR issuer warps, one per
// pipeline, each blocking only on its own pipeline's named barrier: 27.0 ms. (One issuer warp serving
// the four pipelines in a fixed order: 39 ms, with a static software-pipelined slot schedule: 47-50 ms;
// a pipeline in its scatter phase holds up the other three.) The weights are shared, and so are the
// weight-gradient accumulators: they are zeroed once and every pipeline's MMAs accumulate into them
// (one tensor-pipe queue per SM; consecutive accumulations into one TMEM region are the ordinary
// K-loop dependence). With 512 TMEM columns per CTA every layer's weight gradient concatenates two
// sample groups per MMA (issue_dweight_t, WAYS = 2): 20 weight-gradient MMAs per tile.
//
// COMPACT: the rows are not samples 0..M-1 but the `*n_active` samples listed in active_idx (the
// ones whose incoming gradient can be non-zero, written by k_composite_bwd in compact mode together
// with their gradients dsigma_raw / dcolor_raw, which are then indexed by ROW, not by sample).
// A sample with a zero incoming gradient contributes exactly zero to every output of this kernel,
// so leaving it out changes nothing but the time.
namespace bwd4 {
constexpr int kPipes = 4;
constexpr int kRows = 128;                       // rows per pipeline
constexpr int kThreads = kPipes * kRows + 32 * kPipes;    // 16 epilogue warps + one MMA-issuer warp per pipeline
// per-pipeline shared-memory region
constexpr int kX = 0;
constexpr int kH = kX + 8192;
constexpr int kDIN = kH + 8192;
constexpr int kH1 = kDIN + 8192;
constexpr int kH2 = kH1 + 8192;
constexpr int kDO = kH2 + 8192;                  // [128][16]
constexpr int kPos = kDO + 4096;                 // 3 x 136 floats (rows skewed by row / 16)
constexpr int kPipeBytes = kPos + 3 * 136 * 4 + 32;   // 46720: a multiple of 128
// CTA-wide
constexpr int kW = kPipes * kPipeBytes;          // the five weight tiles (also absorb the last over-read)
constexpr int kLv = kW + 8192;
constexpr int kBar = kLv + 512;                  // done[4], free[4]
constexpr int kTmemPtr = kBar + 64;
constexpr int kBytes = kTmemPtr + 16;
constexpr uint32_t kTmemCols = 512;
constexpr int cAcc = 0;                          // + 32 p
constexpr int cDWd3 = 128, cDW2p = 160, cDWd2 = 192, cDWd1 = 256, cDW1p = 320;
static_assert(kPipeBytes % 128 == 0, "pipeline regions must keep the 128-byte tile alignment");
}  // namespace bwd4

template <int N>
__device__ __forceinline__ void issue_layer1(uint32_t acc, uint32_t a_tile, uint32_t w_tile) {
  constexpr uint32_t idesc = make_idesc(128, N, 0, 0);
#pragma unroll
  for (int k = 0; k < 2; ++k)
    umma_f16(acc, desc_k_major(a_tile + k * 2 * kCore, 32), desc_k_major(w_tile + k * 2 * kCore, 32), idesc, k);
}
// named barriers 1 + p: the 128 epilogue threads of pipeline p arrive, the issuer warp syncs (count 160);
// named barriers 5 + p: the 128 epilogue threads of pipeline p among themselves
__device__ __forceinline__ void pipe_arrive(int p) {
  fence_async_smem();
  tc_fence_before();
  asm volatile("bar.arrive %0, 160;" ::"r"(1 + p) : "memory");
}
__device__ __forceinline__ void pipe_wait(int p) {
  asm volatile("bar.sync %0, 160;" ::"r"(1 + p) : "memory");
  tc_fence_after();
}
__device__ __forceinline__ void pipe_sync(int p) { asm volatile("bar.sync %0, 128;" ::"r"(5 + p) : "memory"); }

template <bool COMPACT>
__global__ void __launch_bounds__(bwd4::kThreads, 1)
k_field_bwd_tc4(atmonr_grid_t g, const __half* __restrict__ pos_w, const __half* __restrict__ dir_w,
                const float* __restrict__ x01, const float* __restrict__ dirs, const __half* __restrict__ enc_in,
                const float* __restrict__ dsigma_raw, const float* __restrict__ dcolor_raw,
                const float* __restrict__ grad_absmax, int64_t M_samples, int N, float* __restrict__ dtable,
                float* __restrict__ dpos_w, float* __restrict__ ddir_w, const uint32_t* __restrict__ active_idx,
                const uint32_t* __restrict__ n_active) {
  const int64_t M = COMPACT ? (int64_t)*n_active : M_samples;  // rows
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + bwd4::kBar);        // done[p]: committed MMAs of pipeline p's stage
  uint64_t* free_ = done + bwd4::kPipes;                                   // free[p]: every MMA reading p's tile is complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + bwd4::kTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5;
  LevelRow* lv = reinterpret_cast<LevelRow*>(smem + bwd4::kLv);
  load_level_table(g, lv);
  load_field_weights(smem + bwd4::kW, pos_w, dir_w);
  if (warp == 0) tmem_alloc<bwd4::kTmemCols>(tmem_ptr);
  if (tid == 0) {
    for (int p = 0; p < 2 * bwd4::kPipes; ++p) mbar_init(done + p, 1);
    fence_mbar_init();
  }
  publish_and_sync();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t sb = smem_u32(smem), sw = sb + bwd4::kW;
  float S = 1.0f;
  if (grad_absmax) {
    const float amax = *grad_absmax;
    if (amax > 0.0f && amax < INFINITY) S = exp2f(fminf(fmaxf(floorf(log2f(2048.0f / amax)), -60.0f), 60.0f));
  }
  const float invS = 1.0f / S;
  // zero the (shared) weight-gradient accumulators: columns 128 .. 383 of all 128 lanes
  if (warp < 4) {
    tmem_st_zero128(tmem_addr(tmem, warp, bwd4::cDWd3));
    tmem_st_zero128(tmem_addr(tmem, warp, bwd4::cDWd3 + 128));
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int64_t n_tiles = (M + bwd4::kRows - 1) / bwd4::kRows;
  // tile of pipeline p in iteration `it` of this CTA: ((it * gridDim + blockIdx) * 4 + p)
  const int64_t group_stride = (int64_t)gridDim.x * bwd4::kPipes;
  const int64_t first_tile = (int64_t)blockIdx.x * bwd4::kPipes;

  if (warp >= bwd4::kPipes * 4) {
    // ------------------------------- MMA issuer warps: one per pipeline ----------------------
    // (all 32 lanes run this code converged; umma_f16 / umma_commit elect the issuing lane)
    // Each issuer only ever blocks on ITS pipeline's barrier, so the pipelines drift apart freely (a
    // single issuer serving the four pipelines in a fixed order was measured at 39 - 50 ms: one
    // pipeline in its scatter phase held up the other three). The weight-gradient accumulators are
    // shared: they are zeroed below and every MMA accumulates (the tensor pipe executes the MMAs of
    // all four issuers in one queue; consecutive accumulations into one TMEM region are the ordinary
    // K-loop dependence).
    const int p = warp - bwd4::kPipes * 4;
    const uint32_t pb = sb + p * bwd4::kPipeBytes;
    const uint32_t acc = tmem + bwd4::cAcc + 32 * p;
    for (int64_t tile = first_tile + p; tile < n_tiles; tile += group_stride) {
      pipe_wait(p);
      issue_layer1<32>(acc, pb + bwd4::kX, sw + fwd::kW1P); umma_commit(done + p);
      pipe_wait(p);
      issue_layer1<16>(acc, pb + bwd4::kH, sw + fwd::kW2P); umma_commit(done + p);
      pipe_wait(p);
      issue_layer1<32>(acc, pb + bwd4::kDIN, sw + fwd::kWD1); umma_commit(done + p);
      pipe_wait(p);
      issue_layer1<32>(acc, pb + bwd4::kH1, sw + fwd::kWD2); umma_commit(done + p);
      pipe_wait(p);  // S0: DO = dL/d(dir_mlp out), H2
      issue_dinput<16>(acc, pb + bwd4::kDO, sw + fwd::kWD3);
      umma_commit(done + p);
      ATM_DW issue_dweight_t<16, 2, 128>(tmem + bwd4::cDWd3, pb + bwd4::kH2, pb + bwd4::kDO, 1u);
      pipe_wait(p);  // S1: X = dL/dh2
      issue_dinput<32>(acc, pb + bwd4::kX, sw + fwd::kWD2);
      umma_commit(done + p);
      ATM_DW issue_dweight_t<32, 2, 128>(tmem + bwd4::cDWd2, pb + bwd4::kH1, pb + bwd4::kX, 1u);
      pipe_wait(p);  // S2: H2 = dL/dh1
      issue_dinput<32>(acc, pb + bwd4::kH2, sw + fwd::kWD1);
      umma_commit(done + p);
      ATM_DW issue_dweight_t<32, 2, 128>(tmem + bwd4::cDWd1, pb + bwd4::kDIN, pb + bwd4::kH2, 1u);
      pipe_wait(p);  // S3: DO = dL/d(pos_mlp out)
      issue_dinput<16>(acc, pb + bwd4::kDO, sw + fwd::kW2P);
      umma_commit(done + p);
      ATM_DW issue_dweight_t<16, 2, 128>(tmem + bwd4::cDW2p, pb + bwd4::kH, pb + bwd4::kDO, 1u);
      pipe_wait(p);  // S4: H1 = dL/dh, DIN = encoded features again
      issue_dinput<32>(acc, pb + bwd4::kH1, sw + fwd::kW1P);
      umma_commit(done + p);
      ATM_DW issue_dweight_t<32, 2, 128>(tmem + bwd4::cDW1p, pb + bwd4::kDIN, pb + bwd4::kH1, 1u);
      umma_commit(free_ + p);  // tile boundary: every MMA that reads this tile's buffers
      __syncwarp();
    }
  } else {
    // ------------------------------- epilogue warps ---------------------------------------------
    const int p = warp >> 2;                    // pipeline
    const int row = tid & (bwd4::kRows - 1);    // row of the pipeline's tile == TMEM lane
    uint8_t* base = smem + p * bwd4::kPipeBytes;
    uint8_t* X = base + bwd4::kX;
    uint8_t* H = base + bwd4::kH;
    uint8_t* DIN = base + bwd4::kDIN;
    uint8_t* H1 = base + bwd4::kH1;
    uint8_t* H2 = base + bwd4::kH2;
    uint8_t* DO = base + bwd4::kDO;
    float* pos = reinterpret_cast<float*>(base + bwd4::kPos);
    uint64_t* my_done = done + p;
    uint64_t* my_free = free_ + p;
    const uint32_t my32 = tmem_addr(tmem, warp, bwd4::cAcc + 32 * p);
    uint32_t phase = 0, phase2 = 0, seen_tile = 0;
    const int64_t my_first = first_tile + p;
    // This thread's inputs of the NEXT tile (encoded features, position) are loaded into registers
    // before the scatter phase of the current tile, so a tile never starts with an exposed global load.
    uint4 nx[4];
    float npos[3];
    int64_t nsample = 0;     // sample index of the next tile's row
    uint32_t idx_ahead = 0;  // COMPACT: list entry of this thread's row one tile further
    auto load_idx = [&](int64_t t) -> uint32_t {
      if (!COMPACT || t >= n_tiles) return 0u;
      const int64_t ii = t * bwd4::kRows + row;
      return __ldg(active_idx + (ii < M ? ii : M - 1));
    };
    auto fetch_inputs = [&](int64_t t) {
      const int64_t ii = t * bwd4::kRows + row;
      int64_t jj = ii < M ? ii : M - 1;
      if (COMPACT) {
        jj = (int64_t)idx_ahead;
        idx_ahead = load_idx(t + group_stride);
      }
      nsample = jj;
      const uint4* src = reinterpret_cast<const uint4*>(enc_in + jj * 32);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) nx[cc] = __ldg(src + cc);
      npos[0] = __ldg(x01 + 3 * jj), npos[1] = __ldg(x01 + 3 * jj + 1), npos[2] = __ldg(x01 + 3 * jj + 2);
    };
    if (my_first < n_tiles) {
      idx_ahead = load_idx(my_first);
      fetch_inputs(my_first);
    }

    for (int64_t tile = my_first; tile < n_tiles; tile += group_stride, seen_tile = 1) {
      const int64_t i = tile * bwd4::kRows + row;  // row (indexes the incoming gradients)
      const bool valid = i < M;
      const int64_t j = nsample;                   // sample (indexes enc_in, x01, dirs)
      if (seen_tile) mbar_wait(my_free, phase2), phase2 ^= 1;
      const uint4* enc_row = reinterpret_cast<const uint4*>(enc_in + j * 32);
      // ---------------- recompute the activations ----------------
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) st_chunk(X, row, cc, 32, nx[cc]);
      pipe_arrive(p);
      {
        const int at = row + (row >> 4);
        pos[at] = npos[0], pos[136 + at] = npos[1], pos[272 + at] = npos[2];
      }
      float4 dc_in = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      float ds_in = 0.0f;
      if (valid) dc_in = __ldg(reinterpret_cast<const float4*>(dcolor_raw + 4 * i)), ds_in = __ldg(dsigma_raw + i);
      const float* dptr = dirs + (size_t)((uint32_t)j / (uint32_t)N) * 3;  // the ray's direction (dir_mlp input)
      const float dr[3] = {__ldg(dptr), __ldg(dptr + 1), __ldg(dptr + 2)};
      mbar_wait(my_done, phase), phase ^= 1;
      tc_fence_after();
      float v[32];
      tmem_ld32(my32, v);
      store_row32<true>(H, row, v);
      pipe_arrive(p);
      mbar_wait(my_done, phase), phase ^= 1;
      tc_fence_after();
      {
        float po[16];
        tmem_ld16(my32, po);
        dir_input_row(dr, po, v);
      }
      store_row32<false>(DIN, row, v);
      pipe_arrive(p);
      mbar_wait(my_done, phase), phase ^= 1;
      tc_fence_after();
      tmem_ld32(my32, v);
      store_row32<true>(H1, row, v);
      pipe_arrive(p);
      mbar_wait(my_done, phase), phase ^= 1;
      tc_fence_after();
      tmem_ld32(my32, v);
      store_row32<true>(H2, row, v);
      // ---------------- backward ----------------
      {
        float dout[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) dout[k] = 0.0f;
        dout[0] = dc_in.x * S, dout[1] = dc_in.y * S, dout[2] = dc_in.z * S, dout[3] = dc_in.w * S;
        store_row16(DO, row, dout);
      }
      pipe_arrive(p);  // S0
      mbar_wait(my_done, phase), phase ^= 1;
      tc_fence_after();
      tmem_ld32(my32, v);
      store_row32_masked(X, H2, row, v);  // dL/dh2 -> X (layer 0 finished with the encoded features)
      pipe_arrive(p);  // S1
      mbar_wait(my_done, phase), phase ^= 1;  // covers dW(d3): DO and H2 are free
      tc_fence_after();
      tmem_ld32(my32, v);
      store_row32_masked(H2, H1, row, v);  // dL/dh1 -> H2
      pipe_arrive(p);  // S2
      mbar_wait(my_done, phase), phase ^= 1;  // covers dW(d2): X and H1 are free
      tc_fence_after();
      tmem_ld32(my32, v);
      {
        float dpo[16];
        dpo[0] = ds_in * S;
#pragma unroll
        for (int k = 1; k < 16; ++k) dpo[k] = v[3 + k];
        store_row16(DO, row, dpo);  // dL/d(pos_mlp out) -> DO
      }
      pipe_arrive(p);  // S3
      mbar_wait(my_done, phase), phase ^= 1;  // covers dW(d1): H2 and DIN are free
      tc_fence_after();
      tmem_ld32(my32, v);
      store_row32_masked(H1, H, row, v);  // dL/dh -> H1
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) st_chunk(DIN, row, cc, 32, enc_row[cc]);  // encoded features again -> DIN
      pipe_arrive(p);  // S4
      mbar_wait(my_done, phase), phase ^= 1;  // covers dW(2p): DO and H are free
      tc_fence_after();
      // ---- table-gradient scatter with run-length merging --------------------------------------
      // dL/d(encoded features) of the tile is staged in shared memory (fp32, X and H are free now),
      // then the work is re-mapped: thread (g, q) walks the 16 CONSECUTIVE samples 16g..16g+15 of
      // level q. Consecutive samples of a ray mostly stay in the same grid cell, so their 8 corner
      // contributions are merged in registers and written with one vector RED per corner per cell
      // run instead of one per sample. Most samples carry NO gradient (density <= 0 kills both
      // dL/dsigma and the compositing weight: 64 % of the rows on the benchmark, exactly zero in
      // float32 as well, profiles/r3_backward_rows.json), so the rows with a non-zero gradient are
      // found first (16 independent shared-memory reads) and only those are walked.
      // (Pairing the x / x+1 corners of a cell into one red.v4.f32 when their entries share an aligned
      // 16-byte pair was measured: +2 ms. The REDs are not what bounds the kernel: without any RED it
      // is only 5.5 ms faster.)
      tmem_ld32(my32, v);
      tc_fence_before();
      {
        float* srow = reinterpret_cast<float*>(X) + row * 32;      // X and H: 16 KB of fp32 staging
        const int swz = (row ^ (row >> 4)) & 7;
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 16-byte chunks, XOR-swizzled by the row to spread banks
          *reinterpret_cast<float4*>(srow + ((k ^ swz) << 2)) =
              make_float4(v[4 * k] * invS, v[4 * k + 1] * invS, v[4 * k + 2] * invS, v[4 * k + 3] * invS);
      }
      pipe_sync(p);
      if (tile + group_stride < n_tiles) fetch_inputs(tile + group_stride);
      {
        const int grp = row & 7, lvl = row >> 3;   // 8 groups of 16 rows x 16 levels
        const LevelRow L = lv[lvl];
        float2* tbase = reinterpret_cast<float2*>(dtable) + L.offset;
        const float* stage = reinterpret_cast<const float*>(X);
        const int64_t row0 = tile * bwd4::kRows + grp * 16;
        const float* srow0 = stage + grp * 16 * 32 + ((lvl & 1) << 1);
        uint32_t nz = 0;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float2 d = *reinterpret_cast<const float2*>(srow0 + r * 32 + ((((lvl >> 1) ^ ((r ^ grp) & 7))) << 2));
          nz |= (d.x != 0.0f || d.y != 0.0f) ? (1u << r) : 0u;
        }
        if (row0 + 16 > M) nz &= row0 < M ? (1u << (int)(M - row0)) - 1u : 0u;
        if (!ATM_SCATTER_ON) nz = 0;
        float acc[16];
        uint32_t c_run[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};
        bool open = false;
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = 0.0f;
        auto flush = [&]() {  // entries are only needed here, once per run of samples in one cell
          uint32_t e[8];
          corner_entries3(L, c_run, e);
#pragma unroll
          for (int c = 0; c < 8; ++c) red_add_f32x2(reinterpret_cast<float*>(entry_ptr(tbase, e[c])), acc[2 * c], acc[2 * c + 1]);
        };
#pragma unroll 1
        while (nz) {
          const int r = __ffs(nz) - 1;
          nz &= nz - 1u;
          const int rr = grp * 16 + r;
          const float2 d = *reinterpret_cast<const float2*>(srow0 + r * 32 + ((((lvl >> 1) ^ ((r ^ grp) & 7))) << 2));
          const float* pp = pos + rr + (rr >> 4);
          const float q[3] = {pp[0], pp[136], pp[272]};
          uint32_t cell[3];
          float frac[3], w[8];
          grid_cell<3>(q, L.scale, cell, frac);
          corner_weights3(frac, w);
          if (open && (cell[0] != c_run[0] || cell[1] != c_run[1] || cell[2] != c_run[2])) {
            flush();
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] = 0.0f;
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            acc[2 * c] = fmaf(w[c], d.x, acc[2 * c]);
            acc[2 * c + 1] = fmaf(w[c], d.y, acc[2 * c + 1]);
          }
          c_run[0] = cell[0], c_run[1] = cell[1], c_run[2] = cell[2];
          open = true;
        }
        if (open) flush();
      }
      pipe_sync(p);  // the staging area is the next tile's X/H
    }
    // this pipeline's last tile: its free[] completion has not been consumed by anybody
    if (seen_tile) mbar_wait(my_free, phase2);
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 2 && first_tile < n_tiles) {
    const int lane = tid & 31;
    flush_dweight_t<32, 2>(tmem, bwd4::cDW1p, warp, lane, 32, invS, dpos_w);
    flush_dweight_t<16, 2>(tmem, bwd4::cDW2p, warp, lane, 16, invS, dpos_w + 1024);
    flush_dweight_t<32, 2>(tmem, bwd4::cDWd1, warp, lane, 32, invS, ddir_w);
    flush_dweight_t<32, 2>(tmem, bwd4::cDWd2, warp, lane, 32, invS, ddir_w + 1024);
    flush_dweight_t<16, 2>(tmem, bwd4::cDWd3, warp, lane, 16, invS, ddir_w + 2048);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<bwd4::kTmemCols>(tmem);
}


}  // namespace atm

using namespace atm;

extern "C" {

// Debug / parity entry point (declared in include/atmonr_b200.h).
int atmonr_tc_probe(const void* a_f16, const void* b_f16, int mode, float* d, void* stream) {
  ATM_REQUIRE(mode >= 0 && mode <= 4, "atmonr_tc_probe", "mode must be 0..4");
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
  ATM_REQUIRE(M < ((int64_t)1 << 31), "atmonr_ngp_field_fwd_tc", "B*N must be below 2^31 per call (chunk the batch)");
  const int64_t tiles = (M + kTile - 1) / kTile;
  const int64_t max_ctas = (int64_t)tc_num_sms() * fwd::kCtasPerSm;
  const int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
  k_field_fwd_tc<<<grid, kTile, fwd::kBytes, reinterpret_cast<cudaStream_t>(stream)>>>(
      *g, (const __half2*)table, (const __half*)pos_w, (const __half*)dir_w, x01, dirs, M, N, sigma_raw, color_raw,
      (__half*)enc_out);
  ATM_CHECK_LAUNCH("atmonr_ngp_field_fwd_tc");
  return 0;
}


// atmonr_extract_sigma on the tensor-core path (declared in include/atmonr_b200.h).
int atmonr_extract_sigma_tc(const atmonr_frame_t* f, const atmonr_grid_t* g, const void* table, const atmonr_mlp_t* pm,
                            const void* pos_w, const double* pts, int64_t n, float alt_compress, float* sigma,
                            void* stream) {
  ATM_REQUIRE(f && g && g->n_dims == 3 && g->n_feat == 2 && g->n_levels == 16, "atmonr_extract_sigma_tc", "bad frame/grid");
  ATM_REQUIRE(pm && pm->in_pad == 32 && pm->width == 32 && pm->n_hidden == 1 && pm->n_out == 16, "atmonr_extract_sigma_tc",
              "pos_mlp must be 32 -> [32] -> 16");
  if (n == 0) return 0;
  const int64_t tiles = (n + kTile - 1) / kTile;
  const int64_t max_ctas = (int64_t)tc_num_sms() * 6;
  const int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
  k_extract_sigma_tc<<<grid, kTile, fwd::kBytes, reinterpret_cast<cudaStream_t>(stream)>>>(
      *f, make_geo_frame(*f), *g, (const __half2*)table, (const __half*)pos_w, pts, n, alt_compress, sigma);
  ATM_CHECK_LAUNCH("atmonr_extract_sigma_tc");
  return 0;
}

// Same contract as atmonr_ngp_field_bwd. enc (optional): features saved by the forward pass;
// grad_absmax (optional device scalar): max |incoming gradient|, sets the fp16 operand scale.
int atmonr_ngp_field_bwd_tc(const atmonr_grid_t* g, const void* table, const atmonr_mlp_t* pm, const void* pos_w,
                            const atmonr_mlp_t* dm, const void* dir_w, const float* x01, const float* dirs,
                            const void* enc, const float* dsigma_raw, const float* dcolor_raw,
                            const float* grad_absmax, int64_t B, int N, float* dtable, float* dpos_w, float* ddir_w,
                            void* stream) {
  if (check_field_shapes(g, pm, dm, "atmonr_ngp_field_bwd_tc")) return -1;
  const int64_t M = B * N;
  if (M == 0) return 0;
  ATM_REQUIRE(M < ((int64_t)1 << 31), "atmonr_ngp_field_bwd_tc", "B*N must be below 2^31 per call (chunk the batch)");
  const bool use_wide = getenv("ATMONR_BWD_NARROW") == nullptr;  // read per call: tests toggle it
  if (enc && use_wide) {
    cudaError_t e4 = cudaFuncSetAttribute(k_field_bwd_tc4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd4::kBytes);
    if (e4 != cudaSuccess) return fail("atmonr_ngp_field_bwd_tc", cudaGetErrorString(e4));
    const int64_t groups = (M + bwd4::kPipes * bwd4::kRows - 1) / (bwd4::kPipes * bwd4::kRows);
    const int grid4 = (int)(groups < (int64_t)tc_num_sms() ? groups : (int64_t)tc_num_sms());
    k_field_bwd_tc4<false><<<grid4, bwd4::kThreads, bwd4::kBytes, reinterpret_cast<cudaStream_t>(stream)>>>(
        *g, (const __half*)pos_w, (const __half*)dir_w, x01, dirs, (const __half*)enc, dsigma_raw, dcolor_raw,
        grad_absmax, M, N, dtable, dpos_w, ddir_w, nullptr, nullptr);
    ATM_CHECK_LAUNCH("atmonr_ngp_field_bwd_tc");
    return 0;
  }
  cudaError_t e = cudaFuncSetAttribute(k_field_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kBytes);
  if (e != cudaSuccess) return fail("atmonr_ngp_field_bwd_tc", cudaGetErrorString(e));
  const int64_t tiles = (M + kTile - 1) / kTile;
  const int grid = (int)(tiles < (int64_t)tc_num_sms() * 2 ? tiles : (int64_t)tc_num_sms() * 2);
  k_field_bwd_tc<<<grid, kTile, bwd::kBytes, reinterpret_cast<cudaStream_t>(stream)>>>(
      *g, (const __half2*)table, (const __half*)pos_w, (const __half*)dir_w, x01, dirs, (const __half*)enc, dsigma_raw,
      dcolor_raw, grad_absmax, M, N, dtable, dpos_w, ddir_w);
  ATM_CHECK_LAUNCH("atmonr_ngp_field_bwd_tc");
  return 0;
}

// The same backward over the samples listed in active_idx only (atmonr_composite_bwd_compact):
// dsigma_c (n_active,) and dcolor_c (n_active, 4) are indexed by list position, *n_active is read
// on the device (no host synchronisation). enc (the forward's cached features) is required.
int atmonr_ngp_field_bwd_tc_compact(const atmonr_grid_t* g, const atmonr_mlp_t* pm, const void* pos_w,
                                    const atmonr_mlp_t* dm, const void* dir_w, const float* x01, const float* dirs,
                                    const void* enc, const uint32_t* active_idx, const uint32_t* n_active,
                                    const float* dsigma_c, const float* dcolor_c, const float* grad_absmax, int64_t B,
                                    int N, float* dtable, float* dpos_w, float* ddir_w, void* stream) {
  if (check_field_shapes(g, pm, dm, "atmonr_ngp_field_bwd_tc_compact")) return -1;
  const int64_t M = B * N;
  if (M == 0) return 0;
  ATM_REQUIRE(M < ((int64_t)1 << 31), "atmonr_ngp_field_bwd_tc_compact", "B*N must be below 2^31 per call (chunk the batch)");
  ATM_REQUIRE(enc && active_idx && n_active && dsigma_c && dcolor_c, "atmonr_ngp_field_bwd_tc_compact", "null argument");
  cudaError_t e4 = cudaFuncSetAttribute(k_field_bwd_tc4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd4::kBytes);
  if (e4 != cudaSuccess) return fail("atmonr_ngp_field_bwd_tc_compact", cudaGetErrorString(e4));
  const int64_t groups = (M + bwd4::kPipes * bwd4::kRows - 1) / (bwd4::kPipes * bwd4::kRows);  // upper bound
  const int grid4 = (int)(groups < (int64_t)tc_num_sms() ? groups : (int64_t)tc_num_sms());
  k_field_bwd_tc4<true><<<grid4, bwd4::kThreads, bwd4::kBytes, reinterpret_cast<cudaStream_t>(stream)>>>(
      *g, (const __half*)pos_w, (const __half*)dir_w, x01, dirs, (const __half*)enc, dsigma_c, dcolor_c, grad_absmax, M,
      N, dtable, dpos_w, ddir_w, active_idx, n_active);
  ATM_CHECK_LAUNCH("atmonr_ngp_field_bwd_tc_compact");
  return 0;
}

}  // extern "C"
