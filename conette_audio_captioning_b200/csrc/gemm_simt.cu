// fp32 CUDA-core GEMM with fused epilogues:  out[m,n] = epi( sum_k A[m,k] * W[n,k] ).
// Used by the fp32 "parity" encoder mode and by the decoder (whose arithmetic stays fp32 so that greedy token ids match
// the reference bit-for-bit).  Both operands are K-major exactly as PyTorch stores activations (M,K)
// and nn.Linear weights (N,K) (reference convnext.py:66-69, torch nn.TransformerDecoderLayer).
#include "common.cuh"
#include "kernels.h"

namespace cnb {

constexpr int kBK = 16;

template <int EPI, typename OutT>
__device__ __forceinline__ void epilogue_store(float acc, int m, int n, const EpiParams& ep, OutT* out, int64_t ldo) {
  float v = acc + (ep.bias ? ep.bias[n] : 0.f);
  if (EPI == EPI_BIAS_GELU) v = gelu_erf(v);
  if (EPI == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
  if (EPI == EPI_SCALE_RESID) v = ep.resid[(int64_t)m * ldo + n] + ep.scale[n] * v;
  out[(int64_t)m * ldo + n] = from_float<OutT>(v);
}

template <int BM, int BN, int TM, int TN, int EPI, typename OutT>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ Wt, int M, int N, int K, EpiParams ep,
                OutT* __restrict__ out, int64_t ldo) {
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  __shared__ float As[kBK][BM + 4];
  __shared__ float Bs[kBK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += kBK) {
    for (int i = tid; i < BM * kBK / 4; i += 256) {
      const int r = i / (kBK / 4), kq = (i % (kBK / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M && k0 + kq < K) v = *reinterpret_cast<const float4*>(A + (int64_t)(m0 + r) * lda + k0 + kq);
      As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
    }
    for (int i = tid; i < BN * kBK / 4; i += 256) {
      const int r = i / (kBK / 4), kq = (i % (kBK / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + r < N && k0 + kq < K) v = *reinterpret_cast<const float4*>(Wt + (int64_t)(n0 + r) * K + k0 + kq);
      Bs[kq + 0][r] = v.x; Bs[kq + 1][r] = v.y; Bs[kq + 2][r] = v.z; Bs[kq + 3][r] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < N) epilogue_store<EPI, OutT>(acc[i][j], m, n, ep, out, ldo);
    }
  }
}

template <int EPI, typename OutT>
static int launch_epi(const float* a, int64_t lda, const float* w, int m, int n, int k, const EpiParams& ep, OutT* out,
                      int64_t ldo, cudaStream_t stream) {
  if (m > 512) {
    dim3 grid((unsigned)ceil_div(n, 128), (unsigned)ceil_div(m, 128));
    gemm_f32_kernel<128, 128, 8, 8, EPI, OutT><<<grid, 256, 0, stream>>>(a, lda, w, m, n, k, ep, out, ldo);
  } else {
    dim3 grid((unsigned)ceil_div(n, 64), (unsigned)ceil_div(m, 32));
    gemm_f32_kernel<32, 64, 2, 4, EPI, OutT><<<grid, 256, 0, stream>>>(a, lda, w, m, n, k, ep, out, ldo);
  }
  CNB_LAUNCH_OK();
  return 0;
}

template <typename OutT>
int launch_gemm_f32(const float* a, int64_t lda, const float* w, int m, int n, int k, Epilogue epi, const EpiParams& ep,
                    OutT* out, int64_t ldo, cudaStream_t stream) {
  CNB_REQUIRE(k % 4 == 0 && lda % 4 == 0, "gemm_f32 needs K and lda to be multiples of 4");
  if (m == 0 || n == 0) return 0;
  switch (epi) {
    case EPI_BIAS: return launch_epi<EPI_BIAS, OutT>(a, lda, w, m, n, k, ep, out, ldo, stream);
    case EPI_BIAS_GELU: return launch_epi<EPI_BIAS_GELU, OutT>(a, lda, w, m, n, k, ep, out, ldo, stream);
    case EPI_BIAS_RELU: return launch_epi<EPI_BIAS_RELU, OutT>(a, lda, w, m, n, k, ep, out, ldo, stream);
    case EPI_SCALE_RESID: return launch_epi<EPI_SCALE_RESID, OutT>(a, lda, w, m, n, k, ep, out, ldo, stream);
  }
  set_error("gemm_f32: unknown epilogue");
  return -1;
}
template int launch_gemm_f32<float>(const float*, int64_t, const float*, int, int, int, Epilogue, const EpiParams&, float*,
                                    int64_t, cudaStream_t);

}  // namespace cnb
