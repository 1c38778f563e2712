// Decoder attention for one (row, head) per warp -- shared by every decoder execution mode so that they stay bit-identical.
// Latency-oriented: a decode step is a chain of L2 round trips, so every load that does not depend on another load's value
// is issued in the same batch (lane = cached position for the self-attention scores, 8-deep batches for the value sums).
// Reference: torch nn.MultiheadAttention inside nn.TransformerDecoderLayer (8 heads x 32, scale 1/sqrt(32)), called from
// src/conette/nn/decoders/aac_tfmer.py:100-116 with a causal mask (self) and a key-padding mask t >= len (cross).
#pragma once
#include "common.cuh"

namespace cnb {

constexpr int kAttD = 256, kAttHeads = 8, kAttHeadDim = 32;
constexpr float kAttScale = 0.17677669529663687f;  // 1/sqrt(32)

// Self-attention of row r, head h at position `pos` over the KV cache.  kcache/vcache: (R, max_len, 256) of this layer;
// src_row[r][p] = physical row holding position p of row r's history (beam back-pointers).  max_len <= 64.
// COHERENT: read activations through L2 (ld.global.cg) -- needed inside the persistent kernel where peers wrote them.
template <bool COHERENT>
__device__ __forceinline__ float att_ld(const float* p) {
  return COHERENT ? __ldcg(p) : *p;
}
template <bool COHERENT>
__device__ __forceinline__ float4 att_ld4(const float* p) {
  return COHERENT ? __ldcg(reinterpret_cast<const float4*>(p)) : *reinterpret_cast<const float4*>(p);
}

template <bool COHERENT>
__device__ __forceinline__ void self_attention_task(const float* qkv, float* kcache, float* vcache, const int* src_row, int pos,
                                                    int max_len, float* attn, int r, int h, int lane) {
  const int col = h * kAttHeadDim + lane;
  const float* qrow = qkv + (int64_t)r * 768 + h * kAttHeadDim;
  // lane d: this position's q/k/v element d (k, v go to the cache); every lane: the whole 32-dim q (broadcast loads)
  const float q_d = att_ld<COHERENT>(qkv + (int64_t)r * 768 + col);
  const float k_d = att_ld<COHERENT>(qkv + (int64_t)r * 768 + 256 + col);
  const float v_d = att_ld<COHERENT>(qkv + (int64_t)r * 768 + 512 + col);
  float qv[kAttHeadDim];
#pragma unroll
  for (int d = 0; d < kAttHeadDim; d += 4) {
    const float4 t = att_ld4<COHERENT>(qrow + d);
    qv[d] = t.x; qv[d + 1] = t.y; qv[d + 2] = t.z; qv[d + 3] = t.w;
  }
  kcache[((int64_t)r * max_len + pos) * kAttD + col] = k_d;
  vcache[((int64_t)r * max_len + pos) * kAttD + col] = v_d;
  // scores: lane owns cached positions lane and lane+32 (< pos); the new position's score comes from registers
  float sc[2];
  int pr[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int p = lane + 32 * i;
    sc[i] = -INFINITY;
    pr[i] = r;
    if (p < pos) {
      pr[i] = src_row[(int64_t)r * max_len + p];
      const float* kr = kcache + ((int64_t)pr[i] * max_len + p) * kAttD + h * kAttHeadDim;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < kAttHeadDim; d += 4) {
        const float4 kk = att_ld4<COHERENT>(kr + d);
        a = fmaf(qv[d], kk.x, a); a = fmaf(qv[d + 1], kk.y, a); a = fmaf(qv[d + 2], kk.z, a); a = fmaf(qv[d + 3], kk.w, a);
      }
      sc[i] = a * kAttScale;
    }
  }
  const float s_new = warp_sum(q_d * k_d) * kAttScale;
  if ((pos & 31) == lane) sc[pos >> 5] = s_new;
  const float mx = warp_max(fmaxf(sc[0], sc[1]));
  const float e0 = (sc[0] == -INFINITY) ? 0.f : expf(sc[0] - mx);
  const float e1 = (sc[1] == -INFINITY) ? 0.f : expf(sc[1] - mx);
  const float inv = 1.f / warp_sum(e0 + e1);
  // values: acc[d] = sum_p w_p * V[p][d]; loads batched 8 deep (addresses come from shuffles, not from memory)
  float acc = 0.f;
  for (int p0 = 0; p0 < pos; p0 += 8) {
    float vv[8], ww[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = p0 + u;
      const int src = __shfl_sync(0xffffffffu, (p >> 5) ? pr[1] : pr[0], p & 31);
      ww[u] = __shfl_sync(0xffffffffu, (p >> 5) ? e1 : e0, p & 31);
      vv[u] = (p < pos) ? att_ld<COHERENT>(vcache + ((int64_t)src * max_len + p) * kAttD + col) : 0.f;
      if (p >= pos) ww[u] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc = fmaf(ww[u], vv[u], acc);
  }
  const float w_new = __shfl_sync(0xffffffffu, (pos >> 5) ? e1 : e0, pos & 31);
  acc = fmaf(w_new, v_d, acc);
  attn[(int64_t)r * kAttD + col] = acc * inv;
}

// Cross-attention of row r, head h over the T' encoder frames of its clip.  ck/cv rows have stride kv_stride floats.
// sc: warp-private shared scratch of tp floats.
template <bool COHERENT>
__device__ __forceinline__ void cross_attention_task(float* sc, const float* q, const float* ck, const float* cv,
                                                     int64_t kv_stride, int len, int clip, int tp, float* attn, int r, int h,
                                                     int lane) {
  const float* qh = q + (int64_t)r * kAttD + h * kAttHeadDim;
  float qv[kAttHeadDim];
#pragma unroll
  for (int d = 0; d < kAttHeadDim; d += 4) {
    const float4 t = att_ld4<COHERENT>(qh + d);
    qv[d] = t.x; qv[d + 1] = t.y; qv[d + 2] = t.z; qv[d + 3] = t.w;
  }
  float mx = -INFINITY;
  for (int t = lane; t < tp; t += 32) {
    float s = -INFINITY;
    if (t < len) {  // key_padding_mask: frames t >= len are masked (reference pl_modules/conette.py:460-462)
      const float* kr = ck + ((int64_t)clip * tp + t) * kv_stride + h * kAttHeadDim;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < kAttHeadDim; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(kr + d);
        a = fmaf(qv[d], kk.x, a); a = fmaf(qv[d + 1], kk.y, a); a = fmaf(qv[d + 2], kk.z, a); a = fmaf(qv[d + 3], kk.w, a);
      }
      s = a * kAttScale;
    }
    sc[t] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int t = lane; t < tp; t += 32) {
    const float e = (sc[t] == -INFINITY) ? 0.f : expf(sc[t] - mx);
    sc[t] = e;
    sum += e;
  }
  const float inv = 1.f / warp_sum(sum);
  __syncwarp();
  float acc = 0.f;
  const int tmax = len < tp ? len : tp;
  const float* vbase = cv + (int64_t)clip * tp * kv_stride + h * kAttHeadDim + lane;
  for (int t0 = 0; t0 < tmax; t0 += 8) {
    float vv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) vv[u] = (t0 + u < tmax) ? vbase[(int64_t)(t0 + u) * kv_stride] : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) acc = fmaf((t0 + u < tmax) ? sc[t0 + u] : 0.f, vv[u], acc);
  }
  attn[(int64_t)r * kAttD + h * kAttHeadDim + lane] = acc * inv;
  __syncwarp();
}

}  // namespace cnb
