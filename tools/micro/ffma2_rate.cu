// Microbenchmark: issue rate of packed fma.rn.f32x2 (FFMA2) vs scalar FFMA on sm_100a, in the operand pattern of the
// depthwise-conv inner loop (7 accumulators x taps, one weight shared by 7 consecutive instructions).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(384, 1) k(const float2* __restrict__ src, float2* __restrict__ dst, int iters) {
  float2 in[13], w[7], acc[14];
  for (int i = 0; i < 13; ++i) in[i] = src[threadIdx.x + 384 * i];
  for (int i = 0; i < 7; ++i) w[i] = src[threadIdx.x + 384 * (13 + i)];
  for (int i = 0; i < 14; ++i) acc[i] = make_float2(0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 7; ++kk) {
#pragma unroll
      for (int p = 0; p < 7; ++p) {
        if (MODE == 0) {  // packed, weight shared by 14 consecutive instructions
          acc[p] = __ffma2_rn(in[p + kk], w[kk], acc[p]);
          acc[p + 7] = __ffma2_rn(in[p + kk], w[(kk + 1) % 7], acc[p + 7]);
        } else if (MODE == 1) {  // scalar
          acc[p].x = fmaf(in[p + kk].x, w[kk].x, acc[p].x);
          acc[p].y = fmaf(in[p + kk].y, w[kk].y, acc[p].y);
          acc[p + 7].x = fmaf(in[p + kk].x, w[(kk + 1) % 7].x, acc[p + 7].x);
          acc[p + 7].y = fmaf(in[p + kk].y, w[(kk + 1) % 7].y, acc[p + 7].y);
        } else if (MODE == 2) {  // packed, accumulate in place with both multiplicands equal (2 distinct register pairs)
          acc[p] = __ffma2_rn(w[kk], w[kk], acc[p]);
          acc[p + 7] = __ffma2_rn(w[kk], w[kk], acc[p + 7]);
        } else {  // packed, scalar weight broadcast built once: in * (wx, wx)
          acc[p] = __ffma2_rn(in[p + kk], in[(p + kk + 1) % 13], acc[p]);
          acc[p + 7] = __ffma2_rn(in[p + kk], in[(p + kk + 2) % 13], acc[p + 7]);
        }
      }
    }
  }
  float2 s = make_float2(0.f, 0.f);
  for (int i = 0; i < 14; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
  dst[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int threads, const float2* src, float2* dst) {
  const int iters = 2000, blocks = 148;
  k<MODE><<<blocks, threads>>>(src, dst, 10);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<blocks, threads>>>(src, dst, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double fma = (double)blocks * threads * iters * 98 * 2;  // scalar FMAs
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-34s threads/SM %3d: %.3f ms, %.1f TFMA/s, %.1f FMA/clk/SM (at %d MHz nominal)\n", name, threads, ms,
         fma / ms * 1e-9, fma / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
}

int main() {
  float2 *src, *dst;
  cudaMalloc(&src, 384 * 20 * sizeof(float2));
  cudaMemset(src, 0, 384 * 20 * sizeof(float2));
  cudaMalloc(&dst, 148 * 384 * sizeof(float2));
  for (int threads : {128, 256, 384, 512}) {
    run<0>("FFMA2 (in, w shared, acc)", threads, src, dst);
    run<1>("FFMA scalar", threads, src, dst);
    run<2>("FFMA2 (w, w, acc)", threads, src, dst);
    run<3>("FFMA2 (3 distinct pairs)", threads, src, dst);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
