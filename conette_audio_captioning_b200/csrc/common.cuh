// Shared helpers for the conette_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace cnb {

// ---- error plumbing: every C-ABI entry point returns 0 or a negative code and records a message -----------------
void set_error(const std::string& msg);
#define CNB_CUDA_OK(expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      cnb::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                     ":" + std::to_string(__LINE__) + ")");                                        \
      return -2;                                                                                   \
    }                                                                                              \
  } while (0)
#define CNB_REQUIRE(cond, msg)                                                         \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      cnb::set_error(std::string("invalid argument: ") + (msg) + " [" #cond "]");     \
      return -1;                                                                       \
    }                                                                                  \
  } while (0)
void count_launch();
#define CNB_LAUNCH_OK()              \
  do {                               \
    cnb::count_launch();             \
    CNB_CUDA_OK(cudaGetLastError()); \
  } while (0)

constexpr int kNumSMs = 148;

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// exact (erf) GELU, as nn.GELU() / F.gelu default (reference convnext.py:48, get.py:24-28)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// 16-bit GEMM operand type of precision "fast": IEEE fp16 (11-bit significand = 8x less operand rounding than bf16 at the
// same tcgen05 kind::f16 rate).  Conversions saturate at +-65504 instead of producing inf: ConvNeXt operands are LayerNorm
// outputs and GELU hidden units, far inside the range.
using act16 = __half;
using act16x2 = __half2;
__host__ __device__ __forceinline__ float act2float(act16 v) { return __half2float(v); }
__host__ __device__ __forceinline__ act16 float2act(float v) {
  return __float2half_rn(v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v));
}
__device__ __forceinline__ act16x2 floats2act2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return *reinterpret_cast<act16x2*>(&r);
}

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ act16 from_float<act16>(float v) { return float2act(v); }
__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(act16 v) { return act2float(v); }

}  // namespace cnb
