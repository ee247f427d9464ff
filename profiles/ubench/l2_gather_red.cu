// l2_gather_red.cu -- micro-benchmark behind the "achievable L2 bandwidth" denominators of bench.py
// ("measure achievable random-4-B-gather L2 BW with a micro-benchmark").
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_gather_red l2_gather_red.cu && ./l2_gather_red
//
// Four access patterns over tables of the sizes the Instant-NGP field kernels touch:
//   gather4   : every lane reads 4 bytes (one __half2 entry) at an independent pseudo-random index of an
//               84.6 MB table (the fp16 hash table: L2-resident)         -> 32 sectors per warp request
//   gather4x8 : the same, but the 32 lanes of a warp read from 8 distinct random cells (4 lanes per
//               32-byte sector), which is what consecutive samples of a ray do on the hashed levels
//   red8      : every lane adds 8 bytes (red.global.add.v2.f32, one gradient entry) at an independent
//               random index of a 169 MB float2 table (the gradient table: larger than L2's half)
//   red8x8    : the same with 4 lanes per sector
// Reported: lane-operations/s, the algorithmic bytes/s (4 or 8 B per lane-op, the way SURVEY 8d counts
// them) and the L2 sector traffic/s (32 B per distinct sector per request).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int LANES_PER_SECTOR, int ITERS>
__global__ void __launch_bounds__(256) k_gather(const uint32_t* __restrict__ table, uint32_t n_entries, uint32_t* out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t acc = 0;
  uint32_t key = (LANES_PER_SECTOR == 1 ? tid : (tid / LANES_PER_SECTOR)) * 2654435761u;
#pragma unroll 8
  for (int i = 0; i < ITERS; ++i) {
    key = mix(key + i);
    uint32_t idx = key % n_entries;
    if (LANES_PER_SECTOR > 1) idx = (idx & ~7u) | (lane % LANES_PER_SECTOR);   // same 32-byte sector, own word
    acc += __ldg(table + idx);
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <int LANES_PER_SECTOR, int ITERS>
__global__ void __launch_bounds__(256) k_red(float2* __restrict__ table, uint32_t n_entries) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t key = (LANES_PER_SECTOR == 1 ? tid : (tid / LANES_PER_SECTOR)) * 2654435761u;
#pragma unroll 8
  for (int i = 0; i < ITERS; ++i) {
    key = mix(key + i);
    uint32_t idx = key % n_entries;
    if (LANES_PER_SECTOR > 1) idx = (idx & ~3u) | (lane % LANES_PER_SECTOR);   // 4 float2 per 32-byte sector
    float* p = reinterpret_cast<float*>(table + idx);
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(1.0f), "f"(0.5f) : "memory");
  }
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch(); launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t n_half2 = 21141696u;           // entries of the 3-D hash table (84.6 MB as __half2)
  uint32_t* table; float2* gtable; uint32_t* out;
  cudaMalloc(&table, (size_t)n_half2 * 4); cudaMemset(table, 1, (size_t)n_half2 * 4);
  cudaMalloc(&gtable, (size_t)n_half2 * 8); cudaMemset(gtable, 0, (size_t)n_half2 * 8);
  cudaMalloc(&out, 4);
  constexpr int ITERS = 256;
  const int grid = sms * 8 * 4, block = 256;     // 8 CTAs x 256 threads per SM, 4 waves
  const double lane_ops = (double)grid * block * ITERS;
  struct Row { const char* name; float ms; double bytes_per_op; double sectors_per_warp_req; };
  Row rows[4];
  rows[0] = {"gather4", time_ms([&] { k_gather<1, ITERS><<<grid, block>>>(table, n_half2, out); }, 5), 4.0, 32.0};
  rows[1] = {"gather4x8", time_ms([&] { k_gather<4, ITERS><<<grid, block>>>(table, n_half2, out); }, 5), 4.0, 8.0};
  rows[2] = {"red8", time_ms([&] { k_red<1, ITERS><<<grid, block>>>(gtable, n_half2); }, 5), 8.0, 32.0};
  rows[3] = {"red8x8", time_ms([&] { k_red<4, ITERS><<<grid, block>>>(gtable, n_half2); }, 5), 8.0, 8.0};
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
  printf("{\"sms\": %d, \"table_mb\": %.1f, \"grad_table_mb\": %.1f, \"patterns\": {", sms, n_half2 * 4 / 1e6, n_half2 * 8 / 1e6);
  for (int i = 0; i < 4; ++i) {
    const double s = rows[i].ms * 1e-3;
    printf("%s\"%s\": {\"ms\": %.3f, \"lane_ops_per_s\": %.4g, \"algorithmic_GBps\": %.1f, \"sector_GBps\": %.1f}",
           i ? ", " : "", rows[i].name, rows[i].ms, lane_ops / s, lane_ops * rows[i].bytes_per_op / s / 1e9,
           lane_ops / 32.0 * rows[i].sectors_per_warp_req * 32.0 / s / 1e9);
  }
  printf("}}\n");
  return 0;
}
