// Epilogue helpers shared by the tcgen05 kernels (gemm_tc.cu, mlp_fused.cu): tanh-fit GELU on packed fp32 pairs, TMA bulk
// tensor stores and their group bookkeeping, vector shared-memory accesses, the epilogue-wide named barrier.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace cnb {

constexpr int kEpiWarps = 16;                      // four warps per TMEM lane quarter, interleaved 16-column chunks

// erf-GELU through one MUFU.TANH per element: coefficients fitted so that max |approx - 0.5x(1+erf(x/sqrt2))| = 2.5e-5
// over [-8, 8]; evaluated two elements at a time with the packed fp32 pipe (fma.rn.f32x2 / mul.rn.f32x2).
__device__ __forceinline__ float2 gelu_tanh_fit2(float2 x) {
  const float2 x2 = __fmul2_rn(x, x);
  float2 p = __ffma2_rn(x2, make_float2(-3.51516789e-04f, -3.51516789e-04f), make_float2(3.70056460e-02f, 3.70056460e-02f));
  p = __ffma2_rn(x2, p, make_float2(7.97507884e-01f, 7.97507884e-01f));
  const float2 u = __fmul2_rn(x, p);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(hx, t, hx);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x), "r"(y),
               "r"(src)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(32 * kEpiWarps) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

}  // namespace cnb
