// fp32 CUDA-core GEMM with fused epilogues:  out[m,n] = epi( sum_k A[m,k] * W[n,k] ).
// Used by the fp32 "parity" encoder mode and by the decoder (whose arithmetic stays fp32 so that greedy token ids match
// the reference bit-for-bit).  Both operands are K-major exactly as PyTorch stores activations (M,K)
// and nn.Linear weights (N,K) (reference convnext.py:66-69, torch nn.TransformerDecoderLayer).
//
// Tiles are staged K-contiguous in shared memory by 16-byte cp.async copies (two stages, so the loads of k-block i+1 fly
// under the FMAs of k-block i); every thread owns an interleaved TM x TN micro-tile (rows ty + i*BM/TM, cols tx + j*BN/TN)
// and reads 4 k-values per LDS.128.  The row stride of 36 floats (9 x 16 B, odd) keeps the quarter-warp LDS.128 accesses
// conflict-free.  gridDim.z > 1 = split-K: slice z writes raw partial sums to out + z*M*ldo (epilogue deferred to the
// consumer, e.g. the fused reduce + bias + residual + LayerNorm of the decoder), which stays deterministic.
#include "common.cuh"
#include "kernels.h"

namespace cnb {

constexpr int kBK = 32;
constexpr int kLds = kBK + 4;  // padded row stride (floats)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int EPI, typename OutT>
__device__ __forceinline__ void epilogue_store(float acc, int m, int n, const EpiParams& ep, OutT* out, int64_t ldo) {
  float v = acc + (ep.bias ? ep.bias[n] : 0.f);
  if (EPI == EPI_BIAS_GELU) v = gelu_erf(v);
  if (EPI == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
  if (EPI == EPI_SCALE_RESID) v = ep.resid[(int64_t)m * ldo + n] + ep.scale[n] * v;
  out[(int64_t)m * ldo + n] = from_float<OutT>(v);
}

template <int BM, int BN, int TM, int TN, int EPI, typename OutT>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_f32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ Wt, int M, int N, int K, int k_per_split,
                EpiParams ep, OutT* __restrict__ out, int64_t ldo) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int TX = BN / TN, TY = BM / TM;
  __shared__ __align__(16) float As[2][BM][kLds];
  __shared__ __align__(16) float Bs[2][BN][kLds];
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const int n_kb = (k_end - k_begin + kBK - 1) / kBK;

  auto load_stage = [&](int stage, int kb) {
    const int k0 = k_begin + kb * kBK;
    for (int i = tid; i < BM * (kBK / 4); i += NT) {
      const int r = i / (kBK / 4), kq = (i % (kBK / 4)) * 4;
      const bool ok = (m0 + r < M) && (k0 + kq < k_end);
      cp_async16(&As[stage][r][kq], ok ? A + (int64_t)(m0 + r) * lda + k0 + kq : A, ok);
    }
    for (int i = tid; i < BN * (kBK / 4); i += NT) {
      const int r = i / (kBK / 4), kq = (i % (kBK / 4)) * 4;
      const bool ok = (n0 + r < N) && (k0 + kq < k_end);
      cp_async16(&Bs[stage][r][kq], ok ? Wt + (int64_t)(n0 + r) * K + k0 + kq : Wt, ok);
    }
    cp_async_commit();
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_stage(0, 0);
  for (int kb = 0; kb < n_kb; ++kb) {
    const int st = kb & 1;
    if (kb + 1 < n_kb) {
      load_stage(st ^ 1, kb + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; kk += 4) {
      float4 a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(&As[st][ty + i * TY][kk]);
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const float4*>(&Bs[st][tx + j * TX][kk]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
    __syncthreads();
  }
  if (gridDim.z > 1) {  // split-K: raw partial sums, consumer applies the epilogue
    float* part = reinterpret_cast<float*>(out) + (int64_t)blockIdx.z * M * ldo;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty + i * TY;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx + j * TX;
        if (n < N) part[(int64_t)m * ldo + n] = acc[i][j];
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty + i * TY;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx + j * TX;
      if (n < N) epilogue_store<EPI, OutT>(acc[i][j], m, n, ep, out, ldo);
    }
  }
}

// ---- latency-optimised variant for the decoder: K-slice <= 256 per CTA ------------------------------------------------
// The decoder's GEMMs are tiny (M = live rows ~ 192, K = 256 or a 256-wide split of 2048) and strictly sequential, so a
// k-block loop pays one L2 round trip per block.  Here the whole (BM + BN) x 256 operand panel is requested up front in
// two cp.async groups (k < 128, k >= 128): one L2 latency, then compute on the first half while the second lands.
constexpr int kKP = 256;
constexpr int kKPLds = kKP + 4;

template <int BM, int BN, int TM, int TN, int EPI>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_f32_panel_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ Wt, int M, int N, int K,
                      EpiParams ep, float* __restrict__ out, int64_t ldo) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int TX = BN / TN, TY = BM / TM;
  extern __shared__ __align__(16) float s_panel[];
  float (*As)[kKPLds] = reinterpret_cast<float (*)[kKPLds]>(s_panel);
  float (*Bs)[kKPLds] = reinterpret_cast<float (*)[kKPLds]>(s_panel + BM * kKPLds);
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * kKP;
  const int ks = min(kKP, K - k_begin);  // multiple of 4

  auto load_half = [&](int kh0, int kh1) {
    const int nq = (kh1 - kh0) / 4;
    if (nq <= 0) return;
    for (int i = tid; i < BM * nq; i += NT) {
      const int r = i / nq, kq = kh0 + (i - r * nq) * 4;
      const bool ok = m0 + r < M;
      cp_async16(&As[r][kq], ok ? A + (int64_t)(m0 + r) * lda + k_begin + kq : A, ok);
    }
    for (int i = tid; i < BN * nq; i += NT) {
      const int r = i / nq, kq = kh0 + (i - r * nq) * 4;
      const bool ok = n0 + r < N;
      cp_async16(&Bs[r][kq], ok ? Wt + (int64_t)(n0 + r) * K + k_begin + kq : Wt, ok);
    }
  };
  const int khalf = min(ks, kKP / 2);
  load_half(0, khalf);
  cp_async_commit();
  load_half(khalf, ks);
  cp_async_commit();

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto compute = [&](int k0, int k1) {
#pragma unroll 4
    for (int kk = k0; kk < k1; kk += 4) {
      float4 a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(&As[ty + i * TY][kk]);
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const float4*>(&Bs[tx + j * TX][kk]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
  };
  cp_async_wait<1>();
  __syncthreads();
  compute(0, khalf);
  cp_async_wait<0>();
  __syncthreads();
  compute(khalf, ks);

  float* dst = out + (gridDim.z > 1 ? (int64_t)blockIdx.z * M * ldo : 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty + i * TY;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx + j * TX;
      if (n >= N) continue;
      if (gridDim.z > 1) dst[(int64_t)m * ldo + n] = acc[i][j];  // split-K: raw partial sums
      else epilogue_store<EPI, float>(acc[i][j], m, n, ep, out, ldo);
    }
  }
}

template <int BM, int BN, int TM, int TN, int EPI>
static int launch_panel_cfg(const float* a, int64_t lda, const float* w, int m, int n, int k, int splits, const EpiParams& ep,
                            float* out, int64_t ldo, cudaStream_t stream) {
  constexpr int smem = (BM + BN) * kKPLds * (int)sizeof(float);
  auto kern = gemm_f32_panel_kernel<BM, BN, TM, TN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    CNB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div(n, BN), (unsigned)ceil_div(m, BM), splits);
  kern<<<grid, (BM / TM) * (BN / TN), smem, stream>>>(a, lda, w, m, n, k, ep, out, ldo);
  CNB_LAUNCH_OK();
  return 0;
}

template <int EPI>
static int launch_panel(const float* a, int64_t lda, const float* w, int m, int n, int k, const EpiParams& ep, float* out,
                        int64_t ldo, cudaStream_t stream) {
  const int splits = (int)ceil_div(k, kKP);
  // enough CTAs to cover the chip: big tiles only when the problem is wide
  if (ceil_div(m, 64) * ceil_div(n, 64) * splits >= 96)
    return launch_panel_cfg<64, 64, 4, 4, EPI>(a, lda, w, m, n, k, splits, ep, out, ldo, stream);
  return launch_panel_cfg<32, 32, 2, 2, EPI>(a, lda, w, m, n, k, splits, ep, out, ldo, stream);
}

// Decoder entry: K <= 256 runs with the fused epilogue; K > 256 is split into 256-wide slices whose raw partial sums go
// to out + z*M*ldo (ceil(K/256) slabs) for the consumer to reduce.
int launch_gemm_f32_panel(const float* a, int64_t lda, const float* w, int m, int n, int k, Epilogue epi, const EpiParams& ep,
                          float* out, int64_t ldo, cudaStream_t stream) {
  CNB_REQUIRE(k % 4 == 0 && lda % 4 == 0, "gemm_f32_panel needs K and lda to be multiples of 4");
  if (m == 0 || n == 0) return 0;
  switch (epi) {
    case EPI_BIAS: return launch_panel<EPI_BIAS>(a, lda, w, m, n, k, ep, out, ldo, stream);
    case EPI_BIAS_GELU: return launch_panel<EPI_BIAS_GELU>(a, lda, w, m, n, k, ep, out, ldo, stream);
    case EPI_BIAS_RELU: return launch_panel<EPI_BIAS_RELU>(a, lda, w, m, n, k, ep, out, ldo, stream);
    default: break;
  }
  set_error("gemm_f32_panel: unsupported epilogue");
  return -1;
}

template <int EPI, typename OutT>
static int launch_epi(const float* a, int64_t lda, const float* w, int m, int n, int k, int splits, const EpiParams& ep,
                      OutT* out, int64_t ldo, cudaStream_t stream) {
  const int kps = (int)ceil_div(ceil_div(k, splits), kBK) * kBK;
  const int64_t ctas_64 = ceil_div(m, 64) * ceil_div(n, 64) * splits;
  if (m > 512 || ctas_64 >= 96) {
    dim3 grid((unsigned)ceil_div(n, 64), (unsigned)ceil_div(m, 64), splits);
    gemm_f32_kernel<64, 64, 4, 4, EPI, OutT><<<grid, 256, 0, stream>>>(a, lda, w, m, n, k, kps, ep, out, ldo);
  } else {  // few rows and few columns: small tiles so that the launch still covers many SMs
    dim3 grid((unsigned)ceil_div(n, 32), (unsigned)ceil_div(m, 32), splits);
    gemm_f32_kernel<32, 32, 2, 2, EPI, OutT><<<grid, 256, 0, stream>>>(a, lda, w, m, n, k, kps, ep, out, ldo);
  }
  CNB_LAUNCH_OK();
  return 0;
}

template <typename OutT>
int launch_gemm_f32(const float* a, int64_t lda, const float* w, int m, int n, int k, Epilogue epi, const EpiParams& ep,
                    OutT* out, int64_t ldo, cudaStream_t stream, int splits) {
  CNB_REQUIRE(k % 4 == 0 && lda % 4 == 0, "gemm_f32 needs K and lda to be multiples of 4");
  CNB_REQUIRE(splits >= 1 && (splits == 1 || sizeof(OutT) == 4), "split-K writes fp32 partial sums");
  if (m == 0 || n == 0) return 0;
  switch (epi) {
    case EPI_BIAS: return launch_epi<EPI_BIAS, OutT>(a, lda, w, m, n, k, splits, ep, out, ldo, stream);
    case EPI_BIAS_GELU: return launch_epi<EPI_BIAS_GELU, OutT>(a, lda, w, m, n, k, splits, ep, out, ldo, stream);
    case EPI_BIAS_RELU: return launch_epi<EPI_BIAS_RELU, OutT>(a, lda, w, m, n, k, splits, ep, out, ldo, stream);
    case EPI_SCALE_RESID: return launch_epi<EPI_SCALE_RESID, OutT>(a, lda, w, m, n, k, splits, ep, out, ldo, stream);
  }
  set_error("gemm_f32: unknown epilogue");
  return -1;
}
template int launch_gemm_f32<float>(const float*, int64_t, const float*, int, int, int, Epilogue, const EpiParams&, float*,
                                    int64_t, cudaStream_t, int);

}  // namespace cnb
