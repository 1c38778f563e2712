// Microbenchmark: how fast can one SM pull an L2-resident weight set into shared memory with TMA tile loads, in the shape of
// the cluster decoder's weight ring (32 KB ring stages filled by 2-D boxes of a [rows x K] fp16 matrix, 128-byte swizzle)?
// Every CTA streams the SAME 38.7 MB buffer (the decoder's weights: L2 hits after the first pass) through an S-stage ring with
// no consumer work beyond the mbarrier hand-shake.  Prints GB/s per SM and in aggregate for several grid sizes / box shapes.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_stream tma_stream.cu -lcuda && ./tma_stream
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { auto e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(x), "r"(y), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}

constexpr int kStageBytes = 32768;

// 1-D bulk copies: every stage is `pieces` contiguous blobs of kStageBytes / pieces bytes (the buffer as pre-tiled smem images)
__global__ void __launch_bounds__(128, 1) bulk_kernel(const uint8_t* __restrict__ buf, long long total, int pieces, int stages, int passes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ unsigned long long full[8], empty[8];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_addr(&full[s]), 1);
      mbar_init(smem_addr(&empty[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int per_pass = (int)(total / kStageBytes);
  const int n_chunks = per_pass * passes;
  const int pb = kStageBytes / pieces;
  if (tid == 0) {
    int c = 0;
    for (int g = 0; g < n_chunks; ++g) {
      const int s = g % stages;
      if (g >= stages) mbar_wait(smem_addr(&empty[s]), ((g / stages) - 1) & 1);
      const uint32_t bar = smem_addr(&full[s]);
      mbar_expect_tx(bar, kStageBytes);
      for (int b = 0; b < pieces; ++b)
        bulk_load_1d(smem_addr(smem + (size_t)s * kStageBytes + (size_t)b * pb), buf + (size_t)c * kStageBytes + (size_t)b * pb, pb, bar);
      if (++c == per_pass) c = 0;
    }
  } else if (tid == 32) {
    for (int g = 0; g < n_chunks; ++g) {
      const int s = g % stages;
      mbar_wait(smem_addr(&full[s]), (g / stages) & 1);
      mbar_arrive(smem_addr(&empty[s]));
    }
  }
  __syncthreads();
}

// tensor boxes issued by NP producer threads (one per warp), producer p owning chunks g % NP == p
template <int BR, int NP>
__global__ void __launch_bounds__(32 * (NP + 1), 1) stream_mp_kernel(const __grid_constant__ CUtensorMap map, int rows, int k, int stages,
                                                                     int passes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ unsigned long long full[8], empty[8];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_addr(&full[s]), 1);
      mbar_init(smem_addr(&empty[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  constexpr int kBoxes = kStageBytes / (BR * 128);
  const int kchunks = k / 64, rtiles = rows / BR;
  const int n_boxes = kchunks * rtiles;
  const int n_chunks = n_boxes / kBoxes * passes;
  const int warp = tid >> 5;
  if (warp < NP && (tid & 31) == 0) {
    // stages % NP == 0: producer p always fills stages p, p + NP, ...
    int s = warp, ph = 0;
    int box = warp * kBoxes;
    for (int g = warp; g < n_chunks; g += NP) {
      if (g >= stages) mbar_wait(smem_addr(&empty[s]), ph ^ 1);
      const uint32_t bar = smem_addr(&full[s]);
      mbar_expect_tx(bar, kStageBytes);
#pragma unroll
      for (int b = 0; b < kBoxes; ++b) {
        const int bb = box + b;
        const int rt = (bb >> 2) % rtiles, kc = bb & 3;  // k = 256: 4 k-chunks
        tma_load_2d(smem_addr(smem + (size_t)s * kStageBytes + (size_t)b * BR * 128), &map, kc * 64, rt * BR, bar);
      }
      box += NP * kBoxes;
      s += NP;
      if (s >= stages) { s -= stages; ph ^= 1; }
    }
  } else if (tid == 32 * NP) {
    for (int g = 0; g < n_chunks; ++g) {
      const int s = g % stages;
      mbar_wait(smem_addr(&full[s]), (g / stages) & 1);
      mbar_arrive(smem_addr(&empty[s]));
    }
  }
  __syncthreads();
}

// boxes of [BR rows x 64 k] halves (BR * 128 bytes each); a stage = kStageBytes / (BR * 128) boxes
template <int BR>
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap map, int rows, int k, int stages, int passes,
                                                        unsigned long long* clk_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ unsigned long long full[8], empty[8];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_addr(&full[s]), 1);
      mbar_init(smem_addr(&empty[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  constexpr int kBoxes = kStageBytes / (BR * 128);
  const int kchunks = k / 64, rtiles = rows / BR;
  const int n_boxes = kchunks * rtiles;
  const int n_chunks = n_boxes / kBoxes * passes;
  const unsigned long long t0 = clock64();
  if (tid == 0) {  // producer
    int s = 0, ph = 0, rt = 0, kc = 0;  // running stage / phase / box coordinates: no divisions in the issue loop
    for (int g = 0; g < n_chunks; ++g) {
      if (g >= stages) mbar_wait(smem_addr(&empty[s]), ph ^ 1);
      const uint32_t bar = smem_addr(&full[s]);
      mbar_expect_tx(bar, kStageBytes);
#pragma unroll
      for (int b = 0; b < kBoxes; ++b) {
        tma_load_2d(smem_addr(smem + (size_t)s * kStageBytes + (size_t)b * BR * 128), &map, kc * 64, rt * BR, bar);
        if (++kc == kchunks) { kc = 0; if (++rt == rtiles) rt = 0; }
      }
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else if (tid == 32) {  // consumer: releases a stage as soon as it has landed
    for (int g = 0; g < n_chunks; ++g) {
      const int s = g % stages;
      mbar_wait(smem_addr(&full[s]), (g / stages) & 1);
      mbar_arrive(smem_addr(&empty[s]));
    }
  }
  __syncthreads();
  if (tid == 0) clk_out[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BR>
int run(EncodeFn enc, void* buf, int rows, int k, int grid, int stages, unsigned long long* clk) {
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)BR};
  const cuuint32_t es[2] = {1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("tensor map failed\n");
    return 1;
  }
  auto kern = stream_kernel<BR>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, stages * kStageBytes));
  const int passes = 4;
  kern<<<grid, 128, stages * kStageBytes>>>(map, rows, k, stages, 1, clk);  // warm the L2
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  kern<<<grid, 128, stages * kStageBytes>>>(map, rows, k, stages, passes, clk);
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double bytes = (double)rows * k * 2 * passes;
  printf("box %3d x 64  stages %d  grid %3d: %.3f ms  %.1f GB/s per SM  %.2f TB/s aggregate\n", BR, stages, grid, ms, bytes / ms * 1e-6,
         bytes * grid / ms * 1e-9);
  return 0;
}

int run_bulk(const uint8_t* buf, long long total, int pieces, int grid, int stages) {
  CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, stages * kStageBytes));
  const int passes = 4;
  bulk_kernel<<<grid, 128, stages * kStageBytes>>>(buf, total, pieces, stages, 1);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  bulk_kernel<<<grid, 128, stages * kStageBytes>>>(buf, total, pieces, stages, passes);
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double bytes = (double)total * passes;
  printf("1-D bulk %5d B  stages %d  grid %3d: %.3f ms  %.1f GB/s per SM  %.2f TB/s aggregate\n", kStageBytes / pieces, stages, grid, ms,
         bytes / ms * 1e-6, bytes * grid / ms * 1e-9);
  return 0;
}

template <int BR, int NP>
int run_mp(EncodeFn enc, void* buf, int rows, int k, int grid, int stages) {
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)BR};
  const cuuint32_t es[2] = {1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 1;
  auto kern = stream_mp_kernel<BR, NP>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, stages * kStageBytes));
  const int passes = 4;
  kern<<<grid, 32 * (NP + 1), stages * kStageBytes>>>(map, rows, k, stages, 1);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  kern<<<grid, 32 * (NP + 1), stages * kStageBytes>>>(map, rows, k, stages, passes);
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double bytes = (double)rows * k * 2 * passes;
  printf("box %3d x 64  %d producers  stages %d  grid %3d: %.3f ms  %.1f GB/s per SM  %.2f TB/s aggregate\n", BR, NP, stages, grid, ms,
         bytes / ms * 1e-6, bytes * grid / ms * 1e-9);
  return 0;
}

int main() {
  const int k = 256, rows = 75776;  // 75776 x 256 halves = 38.8 MB
  void* buf;
  CK(cudaMalloc(&buf, (size_t)rows * k * 2));
  CK(cudaMemset(buf, 0, (size_t)rows * k * 2));
  unsigned long long* clk;
  CK(cudaMalloc(&clk, 148 * 8));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  EncodeFn enc = (EncodeFn)fn;
  for (int grid : {1, 8, 104, 148})
    for (int stages : {3, 5, 6}) {
      if (run<128>(enc, buf, rows, k, grid, stages, clk)) return 1;
    }
  for (int grid : {1, 104})
    for (int stages : {5}) {
      if (run<32>(enc, buf, rows, k, grid, stages, clk)) return 1;
      if (run<256>(enc, buf, rows, k, grid, stages, clk)) return 1;
    }
  for (int grid : {1, 104})
    for (int pieces : {1, 2, 8}) {
      if (run_bulk((const uint8_t*)buf, (long long)rows * k * 2, pieces, grid, 5)) return 1;
    }
  if (run_mp<128, 2>(enc, buf, rows, k, 104, 6)) return 1;
  if (run_mp<128, 3>(enc, buf, rows, k, 104, 6)) return 1;
  if (run_mp<32, 2>(enc, buf, rows, k, 104, 6)) return 1;
  if (run_mp<32, 3>(enc, buf, rows, k, 104, 6)) return 1;
  return 0;
}
