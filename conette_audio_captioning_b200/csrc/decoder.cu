// Transformer-decoder step kernels (everything that is not a GEMM): token embedding + positional encoding, KV-cached
// causal self-attention with beam back-pointers, cross-attention over the encoder frames of the row's clip, residual
// add + LayerNorm.  fp32 throughout.
// Reference: nn/decoders/aac_tfmer.py:100-116 (embedding * sqrt(d) + PE, then torch nn.TransformerDecoder with
// post-norm layers, eps 1e-5, 8 heads x 32, no final norm); the reference recomputes all i+1 positions and re-projects
// the memory for every beam at every step -- here K/V are cached.
#include "attention.cuh"
#include "common.cuh"
#include "kernels.h"

namespace cnb {

constexpr int kD = 256;
constexpr int kHeadDim = 32;
constexpr int kHeads = 8;

// ---- x[r,:] = emb[token[r,pos],:] * sqrt(256) + PE[pos,:] ---------------------------------------------------------------
__global__ void embed_kernel(const int* __restrict__ tokens, int tok_stride, int pos, const float* __restrict__ emb,
                             const float* __restrict__ pe, const int* __restrict__ done, float* __restrict__ x, int rows) {
  const int r = blockIdx.x, c = threadIdx.x;
  const int tok = tokens[(int64_t)r * tok_stride + pos];
  x[(int64_t)r * kD + c] = emb[(int64_t)tok * kD + c] * 16.0f + pe[(int64_t)pos * kD + c];
}

// ---- self-attention / cross-attention: one warp per (row, head); arithmetic in attention.cuh -----------------------------
__global__ void __launch_bounds__(256)
self_attn_kernel(const float* __restrict__ qkv, float* kcache, float* vcache, const int* __restrict__ src_row, int pos,
                 int max_len, const int* __restrict__ done, float* __restrict__ attn, int rows) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= rows * kHeads) return;
  self_attention_task<false>(qkv, kcache, vcache, src_row, pos, max_len, attn, gw / kHeads, gw % kHeads, lane);
}

__global__ void __launch_bounds__(256)
cross_attn_kernel(const float* __restrict__ q, const float* __restrict__ ck, const float* __restrict__ cv, int64_t kv_stride,
                  const int* __restrict__ lens, int beam, int tp, const int* __restrict__ done, float* __restrict__ attn,
                  int rows) {
  extern __shared__ float s_sc[];  // (8 warps, tp)
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= rows * kHeads) return;
  const int r = gw / kHeads, h = gw % kHeads;
  const int clip = r / beam;
  cross_attention_task<false>(s_sc + wib * tp, q, ck, cv, kv_stride, lens[clip], clip, tp, attn, r, h, lane);
}

// ---- x = LayerNorm(x + bias + sum_s delta[s]) (eps 1e-5, biased variance): one warp per row of 256 ---------------------
// delta holds nsplit slabs of (rows, 256): the raw split-K partial sums of the preceding GEMM, added in a fixed order.
__global__ void __launch_bounds__(256)
add_ln_kernel(float* __restrict__ x, const float* __restrict__ delta, int nsplit, const float* __restrict__ bias,
              const float* __restrict__ g, const float* __restrict__ b, const int* __restrict__ done, int rows) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const int64_t slab = (int64_t)rows * kD;
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lane + 32 * j;
    float d = bias ? bias[c] : 0.f;
    for (int sp = 0; sp < nsplit; ++sp) d += delta[sp * slab + (int64_t)r * kD + c];
    v[j] = x[(int64_t)r * kD + c] + d;
    s += v[j];
  }
  const float mean = warp_sum(s) * (1.f / kD);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) q += (v[j] - mean) * (v[j] - mean);
  const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / kD) + 1e-5f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lane + 32 * j;
    x[(int64_t)r * kD + c] = (v[j] - mean) * rstd * g[c] + b[c];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
int launch_embed(const int* tokens, int pos, const float* emb, const float* pe, float* x, const DecoderDims& dd,
                 const int* done, cudaStream_t stream) {
  embed_kernel<<<dd.rows, kD, 0, stream>>>(tokens, dd.max_len + 1, pos, emb, pe, done, x, dd.rows);
  CNB_LAUNCH_OK();
  return 0;
}

int launch_self_attn(const float* qkv, float* kcache, float* vcache, const int* src_row, int pos, float* attn,
                     const DecoderDims& dd, const int* done, cudaStream_t stream) {
  const int warps = dd.rows * kHeads;
  self_attn_kernel<<<(warps + 7) / 8, 256, 0, stream>>>(qkv, kcache, vcache, src_row, pos, dd.max_len, done, attn, dd.rows);
  CNB_LAUNCH_OK();
  return 0;
}

int launch_cross_attn(const float* q, const float* ck, const float* cv, int64_t kv_stride, const int* lens, float* attn,
                      const DecoderDims& dd, const int* done, cudaStream_t stream) {
  const int warps = dd.rows * kHeads;
  const size_t smem = (size_t)8 * dd.tp * sizeof(float);
  cross_attn_kernel<<<(warps + 7) / 8, 256, smem, stream>>>(q, ck, cv, kv_stride, lens, dd.beam, dd.tp, done, attn, dd.rows);
  CNB_LAUNCH_OK();
  return 0;
}

int launch_add_ln(float* x, const float* delta, int nsplit, const float* bias, const float* g, const float* b, int rows,
                  const int* done, cudaStream_t stream) {
  add_ln_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(x, delta, nsplit, bias, g, b, done, rows);
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace cnb
