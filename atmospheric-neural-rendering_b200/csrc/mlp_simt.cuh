// mlp_simt.cuh -- thread-per-sample SIMT implementation of the bias-free 32-wide MLPs.
//
// Numerics (identical to oracle/tcnn_spec.py Network.forward(fp16=True)): weights, the padded
// input and every hidden activation are fp16 values; products are accumulated in fp32; the
// output stays fp32. One thread owns one sample (row); weights live in shared memory as fp32
// (converted once per CTA), read as broadcast float4. Weight gradients are reduced per tile
// through shared-memory staging and accumulated in a CTA-private shared copy that is flushed
// to global memory once per CTA.
//
// This is the portable fallback of the tcgen05 path in mlp_tc.cuh: same layout, same rounding
// points.
#pragma once

#include "common.cuh"

namespace atm {

constexpr int kWidth = 32;    // hidden width
constexpr int kOutPad = 16;   // padded output width

template <int IN, int NH>
struct MlpShape {
  static constexpr int kIn = IN;
  static constexpr int kHidden = NH;
  static constexpr int kOffHidden = kWidth * IN;                       // [32][32] (NH == 2)
  static constexpr int kOffOut = kWidth * IN + (NH - 1) * kWidth * kWidth;  // [16][32]
  static constexpr int kNumWeights = kOffOut + kOutPad * kWidth;
};

// fp16 global weights -> fp32 shared memory (all threads of the CTA).
__device__ __forceinline__ void load_weights(const __half* __restrict__ w, float* s, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = __half2float(w[i]);
}

// y[o] = sum_i W[o][i] * x[i], x packed as fp16 pairs.
template <int IN, int OUT>
__device__ __forceinline__ void dense(const float* __restrict__ W, const __half2 (&x)[IN / 2],
                                      float (&y)[OUT]) {
  float xf[IN];
#pragma unroll
  for (int i = 0; i < IN / 2; ++i) {
    const float2 v = __half22float2(x[i]);
    xf[2 * i] = v.x;
    xf[2 * i + 1] = v.y;
  }
#pragma unroll
  for (int o = 0; o < OUT; ++o) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < IN; i += 4) {
      const float4 w = *reinterpret_cast<const float4*>(W + o * IN + i);
      acc = fmaf(w.x, xf[i], acc);
      acc = fmaf(w.y, xf[i + 1], acc);
      acc = fmaf(w.z, xf[i + 2], acc);
      acc = fmaf(w.w, xf[i + 3], acc);
    }
    y[o] = acc;
  }
}

// dx[i] = sum_o W[o][i] * dy[o]
template <int IN, int OUT>
__device__ __forceinline__ void dense_t(const float* __restrict__ W, const float (&dy)[OUT],
                                        float (&dx)[IN]) {
#pragma unroll
  for (int i = 0; i < IN; ++i) dx[i] = 0.0f;
#pragma unroll
  for (int o = 0; o < OUT; ++o) {
    const float g = dy[o];
#pragma unroll
    for (int i = 0; i < IN; i += 4) {
      const float4 w = *reinterpret_cast<const float4*>(W + o * IN + i);
      dx[i] = fmaf(w.x, g, dx[i]);
      dx[i + 1] = fmaf(w.y, g, dx[i + 1]);
      dx[i + 2] = fmaf(w.z, g, dx[i + 2]);
      dx[i + 3] = fmaf(w.w, g, dx[i + 3]);
    }
  }
}

template <int N>
__device__ __forceinline__ void relu_pack(const float (&y)[N], __half2 (&h)[N / 2]) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i)
    h[i] = __floats2half2_rn(fmaxf(y[2 * i], 0.0f), fmaxf(y[2 * i + 1], 0.0f));
}

template <int N>
__device__ __forceinline__ void relu_mask(const __half2 (&h)[N / 2], float (&d)[N]) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float2 v = __half22float2(h[i]);
    if (!(v.x > 0.0f)) d[2 * i] = 0.0f;
    if (!(v.y > 0.0f)) d[2 * i + 1] = 0.0f;
  }
}

// Forward through one MLP keeping the hidden activations (fp16) for the backward pass.
template <int IN, int NH, int OUT_ROWS>
__device__ __forceinline__ void mlp_forward(const float* __restrict__ W,
                                            const __half2 (&x)[IN / 2],
                                            __half2 (&h)[NH][kWidth / 2],
                                            float (&out)[OUT_ROWS]) {
  using S = MlpShape<IN, NH>;
  float y[kWidth];
  dense<IN, kWidth>(W, x, y);
  relu_pack<kWidth>(y, h[0]);
  if (NH == 2) {
    dense<kWidth, kWidth>(W + S::kOffHidden, h[0], y);
    relu_pack<kWidth>(y, h[NH - 1]);
  }
  dense<kWidth, OUT_ROWS>(W + S::kOffOut, h[NH - 1], out);
}

// ---- weight-gradient reduction over one 128-row tile -------------------------------------
// sA: [128][IN+4] activations, sD: [128][OUT+4] deltas (rows of invalid samples are zero).
// Thread (warp w, lane l) owns dW[o][i] for o in [w*OUT/4, (w+1)*OUT/4), i = l (+32).
template <int IN, int OUT>
__device__ __forceinline__ void dw_accumulate(const float* sA, const float* sD, float* sdW) {
  constexpr int LDA = IN + 4, LDD = OUT + 4, OPT = OUT / 4;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int ic = 0; ic < IN; ic += 32) {
    const int i = ic + lane;
    if (i < IN) {
      float acc[OPT];
#pragma unroll
      for (int j = 0; j < OPT; ++j) acc[j] = 0.0f;
#pragma unroll 4
      for (int s = 0; s < kTile; ++s) {
        const float a = sA[s * LDA + i];
        const float4* dp = reinterpret_cast<const float4*>(sD + s * LDD + w * OPT);
#pragma unroll
        for (int q = 0; q < OPT / 4; ++q) {
          const float4 dv = dp[q];
          acc[4 * q] = fmaf(dv.x, a, acc[4 * q]);
          acc[4 * q + 1] = fmaf(dv.y, a, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(dv.z, a, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(dv.w, a, acc[4 * q + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < OPT; ++j) sdW[(w * OPT + j) * IN + i] += acc[j];
    }
  }
}

template <int N>
__device__ __forceinline__ void stage_half(float* s, int ld, const __half2 (&h)[N / 2]) {
  float* row = s + threadIdx.x * ld;
#pragma unroll
  for (int i = 0; i < N / 4; ++i) {
    const float2 a = __half22float2(h[2 * i]), b = __half22float2(h[2 * i + 1]);
    *reinterpret_cast<float4*>(row + 4 * i) = make_float4(a.x, a.y, b.x, b.y);
  }
}
template <int N>
__device__ __forceinline__ void stage_float(float* s, int ld, const float (&v)[N]) {
  float* row = s + threadIdx.x * ld;
#pragma unroll
  for (int i = 0; i < N / 4; ++i)
    *reinterpret_cast<float4*>(row + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// Shared-memory scratch of the backward tile: activations and deltas.
template <int IN>
struct BwdScratch {
  static constexpr int kActCols = (IN > kWidth ? IN : kWidth) + 4;  // the activation area also stages hidden rows
  static constexpr int kFloats = kTile * kActCols + kTile * (kWidth + 4);
};

// Backward through one MLP for the CTA's tile. Every thread of the CTA must call it (it
// synchronises); threads without a valid sample pass dout == 0.
//   W   : fp32 weights in smem          sdW: CTA-private fp32 weight-gradient accumulator
//   x,h : this thread's input and hidden activations (from mlp_forward)
//   dout: dL/d(output rows), length 16 (unused rows zero)
//   dx  : dL/d(padded input)
template <int IN, int NH>
__device__ __forceinline__ void mlp_backward(const float* __restrict__ W, float* sdW, float* scratch,
                                             const __half2 (&x)[IN / 2],
                                             const __half2 (&h)[NH][kWidth / 2],
                                             const float (&dout)[kOutPad], float (&dx)[IN]) {
  using S = MlpShape<IN, NH>;
  float* sA = scratch;
  float* sD = scratch + kTile * BwdScratch<IN>::kActCols;
  float dh[kWidth];

  // output layer: dW_out += dout^T h_last ; dh = W_out^T dout, masked by the ReLU
  __syncthreads();
  stage_half<kWidth>(sA, kWidth + 4, h[NH - 1]);
  stage_float<kOutPad>(sD, kOutPad + 4, dout);
  __syncthreads();
  dw_accumulate<kWidth, kOutPad>(sA, sD, sdW + S::kOffOut);
  dense_t<kWidth, kOutPad>(W + S::kOffOut, dout, dh);
  relu_mask<kWidth>(h[NH - 1], dh);

  if (NH == 2) {
    float dh0[kWidth];
    __syncthreads();
    stage_half<kWidth>(sA, kWidth + 4, h[0]);
    stage_float<kWidth>(sD, kWidth + 4, dh);
    __syncthreads();
    dw_accumulate<kWidth, kWidth>(sA, sD, sdW + S::kOffHidden);
    dense_t<kWidth, kWidth>(W + S::kOffHidden, dh, dh0);
    relu_mask<kWidth>(h[0], dh0);
#pragma unroll
    for (int i = 0; i < kWidth; ++i) dh[i] = dh0[i];
  }

  // first layer
  __syncthreads();
  stage_half<IN>(sA, IN + 4, x);
  stage_float<kWidth>(sD, kWidth + 4, dh);
  __syncthreads();
  dw_accumulate<IN, kWidth>(sA, sD, sdW);
  dense_t<IN, kWidth>(W, dh, dx);
}

// Flush the CTA-private gradient accumulator to global memory (fp32 atomics).
__device__ __forceinline__ void flush_dw(const float* sdW, float* dw, int n) {
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = sdW[i];
    if (v != 0.0f) atomicAdd(dw + i, v);
  }
}

}  // namespace atm
