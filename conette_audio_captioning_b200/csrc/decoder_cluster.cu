// K-DEC cluster: the whole beam-search decode of a group of clips inside ONE thread-block cluster, one launch per call.
//
// Why one launch: a decode step is ~50 strictly dependent tiny operations on R = clips x beam rows; as separate kernels the
// 20-step decode is a chain of ~1400 launches (launch-latency bound, ~10 ms for 64 clips and for 8 clips alike).  Beam search
// never mixes clips, so the batch is cut into groups of NR / beam clips (NR = 16 or 32 decoder rows) and every group is decoded
// start to finish by one cluster of 8 CTAs that never talks to the rest of the grid:
//   * CTA h owns attention head h, 1/8 of every projection, 1/8 of the FF hidden units and 1/8 of the vocabulary;
//   * GEMMs run on the tensor cores with the WEIGHTS as the M operand and the decoder rows as N, in fp32-ACCURATE arithmetic:
//     every weight matrix is stored as two fp16 matrices W1 = fp16(W), W2 = fp16((W - W1) * 2048), activations are split the same
//     way on the fly (x1, x2), and  W.x = W1.x1 + 2^-11 (W1.x2 + W2.x1)  (+ O(2^-22)) is evaluated as TWO tcgen05.mma kind::f16
//     per k16 step: W1 . [x1 | x2] (N = 2 NR, the hi and lo operand blocks are adjacent in shared memory) into accumulator
//     columns [hi | lo], then W2 . x1 accumulated onto the lo columns; the epilogue returns hi + 2^-11 lo.  Same weight bytes as
//     fp32 (4 B / weight), fp32 accumulation in TMEM, logits within ~2e-5 of the fp32 CPU reference -- the earlier tf32 form
//     truncated operands to 11 bits (biased, 1e-3 on scores) and is gone;
//   * weights stream L2 -> shared memory through a 3 x 32 KB TMA ring that runs ahead across phase boundaries (the schedule of a
//     step is static); one warp feeds the ring, one warp issues the MMAs, twelve warps read TMEM for the epilogues;
//   * the two attention output projections are K-SPLIT: CTA h multiplies its own head's attention output (K = 32, no
//     all-gather of the heads) with its 32 columns of W_o for ALL 256 outputs (8 MMAs instead of 32 three-quarters-empty ones),
//     the partial sums are reduce-scattered to the owner of each 32-column slice (the FF2 GEMM works the same way over this CTA's
//     256 hidden units), the owner adds bias + residual and all-gathers the pre-LayerNorm slice; every CTA normalises all rows
//     redundantly IN PLACE (fp32 staging -> hi/lo fp16 operand, same bytes).  6 exchanges per layer, all one-sided DSMEM pushes:
//     st.async completing bytes on the RECEIVER's mbarrier -- no cluster barrier, no fence, no L1 flush;
//   * the classifier is tiled over the vocabulary (rounds of 256 words per CTA): per round the logits tile is scanned for the
//     running per-row max / sum-exp / top-`beam` words, so any vocabulary size works; one exchange later every CTA merges the 8
//     partial results redundantly and deterministically (beam state replicated in shared memory);
// Reference semantics: nn/decoders/aac_tfmer.py:100-116 (embedding*16 + PE, post-norm nn.TransformerDecoder, eps 1e-5),
// nn/decoding/beam.py:113-203 and :230-269 (see beam.cu for the fixed-slot formulation this mirrors).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace cg = cooperative_groups;

namespace cnb {

namespace {

constexpr int kCl = 8;          // CTAs per cluster = attention heads
constexpr int kCThreads = 512;
constexpr int kCWarps = kCThreads / 32;
constexpr int kCD = 256, kCLayers = 6, kCHead = 32;
constexpr int kCMaxBeam = 8;
constexpr int kCMaxLen = 64;
constexpr int kCMaxTp = 128;    // encoder frames per clip handled by the in-register cross-attention scores
constexpr int kCPad = 0, kCEos = 2;
constexpr unsigned kFull = 0xffffffffu;
constexpr float kCAttScale = 0.17677669529663687f;  // 1/sqrt(32)
constexpr int kTrSlots = 20;
constexpr int kStageBytes = 32768;
// weight ring depth: bytes in flight bound the L2 -> shared-memory stream (~32 KB per microsecond and stage); 16-row clusters
// have room for five stages, 32-row clusters for three
template <int NR> struct RingDepth { static constexpr int value = NR <= 16 ? 5 : 3; };
constexpr int kHalfStage = 16384;                  // W1 tiles in the first half of a stage, W2 tiles in the second
constexpr int kChunksPerLayer = 4 + 1 + 1 + 1 + 8 + 8;  // QKV | sa_out | ca_q | ca_out | FF1 | FF2
constexpr int kTileSlots = 4;   // TMEM accumulator tiles (2 per GEMM phase, double-buffered across classifier rounds)
constexpr int kTmemCols = 256;  // 4 slots x 2 NR columns (NR = 32)
constexpr int kIssuerWarp = 12, kProducerWarp = 15, kEpiGroups = 3;  // warps 0..11 read TMEM (three groups of four)
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;
constexpr int kClsRound = 256;  // words per CTA and classifier round (two 128-word tiles)

__device__ __forceinline__ unsigned long long cl_global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct CCand {
  float v;
  int idx;
};
// order of candidates: higher value first, ties -> lower index (torch.topk is unspecified on ties; the oracle's margins cover it)

// The same order as one unsigned 64-bit comparison (no branches): key(a) > key(b) <=> cbetter(a, b).  High word = the float mapped
// monotonically onto unsigned integers, low word = 0x7fffffff - idx (ties -> lower index).  NaN scores and the "nothing" sentinel
// (idx = 0x7fffffff) map to key 0, below every real candidate (-inf maps to 0x007fffff in the high word).
__device__ __forceinline__ unsigned long long ckey(float v, int idx) {
  uint32_t u = __float_as_uint(v);
  u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;
  const bool none = (v != v) || idx == 0x7fffffff;
  return none ? 0ull : ((unsigned long long)u << 32) | (uint32_t)(0x7fffffff - idx);
}
__device__ __forceinline__ CCand ckey_decode(unsigned long long k) {
  if (k == 0ull) return CCand{-INFINITY, 0x7fffffff};
  uint32_t u = (uint32_t)(k >> 32);
  u ^= (u >> 31) ? 0x80000000u : 0xffffffffu;
  return CCand{__uint_as_float(u), 0x7fffffff - (int)(uint32_t)k};
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
    k = other > k ? other : k;
  }
  return k;
}

// ---- DSMEM push with receiver-side completion ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_peer(uint32_t addr, int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async16(uint32_t peer_addr, float4 v, uint32_t peer_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(peer_addr),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)),
               "r"(peer_bar)
               : "memory");
}
__device__ __forceinline__ void st_async8(uint32_t peer_addr, uint32_t a, uint32_t b, uint32_t peer_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(peer_addr), "r"(a),
               "r"(b), "r"(peer_bar)
               : "memory");
}
__device__ __forceinline__ void st_async4(uint32_t peer_addr, float v, uint32_t peer_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(peer_addr),
               "r"(__float_as_uint(v)), "r"(peer_bar)
               : "memory");
}

// ---- fp16 hi/lo split --------------------------------------------------------------------------------------------------
// x = x1 + 2^-11 x2 with x1 = fp16(x), x2 = fp16((x - x1) * 2048): 22 significand bits; |x2| <= |x|, so the lo part cannot
// overflow where the hi part does not (saturating conversions; decoder activations are LayerNorm outputs, attention averages
// and GELU hidden units, |x| << 65504).
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const act16x2 p = floats2act2(v[2 * i], v[2 * i + 1]);
    const float2 f = __half22float2(p);
    const act16x2 q = floats2act2((v[2 * i] - f.x) * kLoScale, (v[2 * i + 1] - f.y) * kLoScale);
    h[i] = *reinterpret_cast<const uint32_t*>(&p);
    l[i] = *reinterpret_cast<const uint32_t*>(&q);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---- operand layouts (UMMA K-major) ---------------------------------------------------------------------------------------
// K = 256 operand of NR rows ("opx" / "oph"): 4 k-chunks of 64, each [hi: NR rows x 128 B | lo: NR rows x 128 B], 128B swizzle
// (16-byte group index XOR (row & 7)).  hi and lo blocks are adjacent, so one MMA with N = 2 NR reads both.
template <int NR> __device__ __forceinline__ int op_off(int r, int k) {  // byte offset of the 16-byte group holding hi(r, k..k+7)
  return (k >> 6) * (2 * NR * 128) + r * 128 + ((((k >> 3) & 7) ^ (r & 7)) << 4);
}
// K = 32 operand ("opa", one head): [hi: NR rows x 64 B | lo: NR rows x 64 B], 64B swizzle (group index XOR ((row >> 1) & 3))
__device__ __forceinline__ int opa_off(int r, int k) { return r * 64 + ((((k >> 3) & 3) ^ ((r >> 1) & 3)) << 4); }

template <int NR> struct CSmem {
  static constexpr int kStages = RingDepth<NR>::value;
  alignas(1024) uint8_t ring[kStages][kStageBytes];  // weight tiles (A operand)
  // layer input x as hi/lo operand; between an all-gather and its LayerNorm the same bytes hold the pre-LN rows as fp32 [NR][256]
  alignas(1024) uint8_t opx[NR * 1024];
  // FF hidden slice (256 units of this CTA) as hi/lo operand | q [NR][32] + k|v [NR][64] of this head (fp32) during the
  // attention phases | the fp32 logits tile [NR][256] of a classifier round
  alignas(1024) uint8_t oph[NR * 1024];
  alignas(1024) uint8_t opa[NR * 128];               // attention output of this head as hi/lo operand (K = 32)
  // reduce-scatter landing zone: partial sums for this CTA's 32 columns, one slab per peer.  During the beam exchange the same
  // bytes hold stat [kCl][NR][2] (max, sum-exp per peer slice) followed by cnd [kCl][NR][kCMaxBeam] (best words per peer slice)
  float recv[kCl][NR][kCHead];
  float xr[NR][kCHead];                              // this CTA's 32 columns of the residual stream, exact fp32
  float st_stat[NR][2];                              // running max / sum-exp of this CTA's vocabulary slice
  CCand st_cnd[NR][kCMaxBeam];                       // running best words of this CTA's vocabulary slice (by logit)
  CCand win[kCWarps][kCMaxBeam];
  int2 winj[kCWarps][kCMaxBeam];                     // (previous beam position, word) of every winner
  uint16_t tokens[2][NR][kCMaxLen + 2];
  uint8_t src[2][NR][kCMaxLen];                      // local row holding position p of this row's history (beam back-pointers)
  float sum_lp[NR];
  int live[NR];
  int any_live;
  uint32_t tmem_slot;
  unsigned long long bar_rs, bar_ag, bar_beam;       // exchange mbarriers: reduce-scatter, all-gather, beam step
  unsigned long long full[kStages], empty[kStages], tile_full[kTileSlots];
  unsigned long long tr_acc[kTrSlots];
  unsigned long long tr_last;
  unsigned long long tr_wait_full;  // debug trace: SM cycles the MMA issuer of the first CTA spent waiting for weight chunks

  __device__ __forceinline__ float (*q())[kCHead] { return reinterpret_cast<float(*)[kCHead]>(oph); }
  __device__ __forceinline__ float (*kv())[2 * kCHead] { return reinterpret_cast<float(*)[2 * kCHead]>(oph + NR * kCHead * 4); }
  __device__ __forceinline__ float (*lt())[kClsRound] { return reinterpret_cast<float(*)[kClsRound]>(oph); }
  __device__ __forceinline__ float (*stage())[kCD] { return reinterpret_cast<float(*)[kCD]>(opx); }
  __device__ __forceinline__ float (*stat())[NR][2] { return reinterpret_cast<float(*)[NR][2]>(&recv[0][0][0]); }
  __device__ __forceinline__ CCand (*cnd())[NR][kCMaxBeam] {
    return reinterpret_cast<CCand(*)[NR][kCMaxBeam]>(reinterpret_cast<uint8_t*>(&recv[0][0][0]) + kCl * NR * 2 * sizeof(float));
  }
};

// state of the weight pipeline (tracked by every thread; `load` is meaningful in the producer warp)
struct Pipe {
  uint32_t load = 0;  // chunks requested so far (running over all steps)
  uint32_t use = 0;   // chunks consumed so far
};

// Global loads / stores of the attention data (self-attention caches, cross-attention K / V) with an explicit L2 eviction priority
// (ClusterArgs::kv_policy, always a valid policy: kL2EvictNormal = the default behaviour)
__device__ __forceinline__ float4 ldg_f4_pol(const float* p, unsigned long long pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ldg_f32_pol(const float* p, unsigned long long pol) {
  float v;
  asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg_f32_pol(float* p, float v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

// Attention of NRW rows at once by one warp over up to NCH*32 cached keys each (+ optionally one extra key/value held in
// shared memory: the position being decoded).  kptr(rr, j) / vptr(rr, j) give the 32-float head slice of key / value j of
// row rr.  Scores: lane = key (8 x LDG.128 each, all rows / chunks requested before the first use).  Values: lane =
// (key group of 4, 4-dim quad): 16-byte loads, all requested up front, then a 2-step shuffle reduction over the groups.
// The result (32 floats per row) is written as hi/lo fp16 into the K = 32 operand buffer `opa` (row index out_row[rr]).
// KT = true (cross-attention): kptr(rr, 0) is the base of the clip's TRANSPOSED key slab [32 dims][tpad frames] (written once per
// call by cross_k_transpose_kernel), so lane = frame reads 32 coalesced words (one L1 wavefront each) instead of eight 16-byte
// pieces of its own 128-byte row (32 wavefronts per instruction: the L1 request queue was the bottleneck of the phase).
template <int NR, int NRW, int NCH, bool KT, typename KPtr, typename VPtr>
__device__ __forceinline__ void attend(const float* const (&q)[NRW], const int (&n)[NRW], const bool (&valid)[NRW], KPtr kptr,
                                       VPtr vptr, const float* const (&kv_new)[NRW], bool has_new, uint8_t* opa,
                                       const int (&out_row)[NRW], int lane, unsigned long long pol, int tpad = 0) {
  float sc[NRW][NCH];
  // one row x one chunk per warp (10 s clips in 16-row clusters): the value rows are requested together with the keys, so the
  // phase costs one L2 round trip instead of two (the addresses do not depend on the scores)
  constexpr bool kHoist = NRW * NCH == 1;
  float4 vh[kHoist ? 8 : 1];
  if (kHoist) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = (lane >> 3) + 4 * u;
      vh[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid[0] && j < n[0]) vh[u] = ldg_f4_pol(vptr(0, j) + 4 * (lane & 7), pol);
    }
  }
#pragma unroll
  for (int rr = 0; rr < NRW; ++rr) {
    float qv[kCHead];
#pragma unroll
    for (int d = 0; d < kCHead; d += 4) {
      const float4 t = valid[rr] ? *reinterpret_cast<const float4*>(q[rr] + d) : make_float4(0.f, 0.f, 0.f, 0.f);
      qv[d] = t.x; qv[d + 1] = t.y; qv[d + 2] = t.z; qv[d + 3] = t.w;
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int j = lane + 32 * ch;
      sc[rr][ch] = -INFINITY;
      if (KT) {
        if (valid[rr] && j < n[rr]) {
          const float* kb = kptr(rr, 0) + j;
          float kk[kCHead];
#pragma unroll
          for (int d = 0; d < kCHead; ++d) kk[d] = ldg_f32_pol(kb + d * tpad, pol);
          float a = 0.f;
#pragma unroll
          for (int d = 0; d < kCHead; ++d) a = fmaf(qv[d], kk[d], a);
          sc[rr][ch] = a * kCAttScale;
        }
      } else if (valid[rr] && j < n[rr]) {
        const float* kr = kptr(rr, j);
        float4 kk[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) kk[d] = ldg_f4_pol(kr + 4 * d, pol);
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          a = fmaf(qv[4 * d], kk[d].x, a); a = fmaf(qv[4 * d + 1], kk[d].y, a);
          a = fmaf(qv[4 * d + 2], kk[d].z, a); a = fmaf(qv[4 * d + 3], kk[d].w, a);
        }
        sc[rr][ch] = a * kCAttScale;
      }
    }
  }
  float e[NRW][NCH], e_new[NRW], inv[NRW];
#pragma unroll
  for (int rr = 0; rr < NRW; ++rr) {
    float s_new = -INFINITY;
    if (has_new) s_new = warp_sum(valid[rr] ? q[rr][lane] * kv_new[rr][lane] : 0.f) * kCAttScale;
    float mx = s_new;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) mx = fmaxf(mx, sc[rr][ch]);
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      e[rr][ch] = (sc[rr][ch] == -INFINITY) ? 0.f : expf(sc[rr][ch] - mx);
      sum += e[rr][ch];
    }
    sum = warp_sum(sum);
    e_new[rr] = has_new ? expf(s_new - mx) : 0.f;
    inv[rr] = 1.f / (sum + e_new[rr]);
  }
  // values
  const int dq = lane & 7, pg = lane >> 3;
#pragma unroll
  for (int rr = 0; rr < NRW; ++rr) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float4 vv[8];
      float ww[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int jl = pg + 4 * u;  // key index inside the chunk
        const int j = jl + 32 * ch;
        ww[u] = __shfl_sync(kFull, e[rr][ch], jl);
        if (kHoist) {
          vv[u] = vh[u];
        } else {
          vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid[rr] && j < n[rr]) vv[u] = ldg_f4_pol(vptr(rr, j) + 4 * dq, pol);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc.x = fmaf(ww[u], vv[u].x, acc.x); acc.y = fmaf(ww[u], vv[u].y, acc.y);
        acc.z = fmaf(ww[u], vv[u].z, acc.z); acc.w = fmaf(ww[u], vv[u].w, acc.w);
      }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      acc.x += __shfl_xor_sync(kFull, acc.x, o); acc.y += __shfl_xor_sync(kFull, acc.y, o);
      acc.z += __shfl_xor_sync(kFull, acc.z, o); acc.w += __shfl_xor_sync(kFull, acc.w, o);
    }
    if (has_new && valid[rr] && lane < 8) {
      const float4 vn = *reinterpret_cast<const float4*>(kv_new[rr] + kCHead + 4 * dq);
      acc.x = fmaf(e_new[rr], vn.x, acc.x); acc.y = fmaf(e_new[rr], vn.y, acc.y);
      acc.z = fmaf(e_new[rr], vn.z, acc.z); acc.w = fmaf(e_new[rr], vn.w, acc.w);
    }
    acc.x *= inv[rr]; acc.y *= inv[rr]; acc.z *= inv[rr]; acc.w *= inv[rr];
    // lanes 0..7 hold dims [4 dq, +4): even lanes fetch their neighbour's quad and write one 16-byte hi group + one lo group
    const float nx = __shfl_down_sync(kFull, acc.x, 1), ny = __shfl_down_sync(kFull, acc.y, 1);
    const float nz = __shfl_down_sync(kFull, acc.z, 1), nw = __shfl_down_sync(kFull, acc.w, 1);
    if (valid[rr] && lane < 8 && (lane & 1) == 0) {
      const float v8[8] = {acc.x, acc.y, acc.z, acc.w, nx, ny, nz, nw};
      uint4 hi, lo;
      split8(v8, hi, lo);
      const int off = opa_off(out_row[rr], 4 * dq);
      *reinterpret_cast<uint4*>(opa + off) = hi;
      *reinterpret_cast<uint4*>(opa + NR * 64 + off) = lo;
    }
  }
}

#define CL_TR(slot)                                \
  if (tr_on) {                                     \
    const unsigned long long n_ = cl_global_ns();  \
    S.tr_acc[slot] += n_ - S.tr_last;              \
    S.tr_last = n_;                                \
  }

// ---- weight pipeline --------------------------------------------------------------------------------------------------
// The chunk stream of one decode step is static (chunk = one 32 KB ring stage: W1 tiles in the first 16 KB, W2 tiles in the
// second 16 KB, all fp16, k-chunks of 64 = 128-byte rows):
//   per layer  QKV      4 chunks: k-chunk c, three boxes [32 rows x 64 k] per half (q / k / v rows of this head: one 96-row tile)
//              sa_out   1 chunk : head-packed W_o[:, 32 h .. +32]: two boxes [128 rows x 32 k] per half (64-byte rows, 64B swizzle)
//              ca_q     1 chunk : four boxes [32 rows x 64 k] per half (the whole K = 256 of this head's 32 query rows)
//              ca_out   1 chunk : like sa_out
//              FF1      8 chunks: tile t (128 hidden units) x k-chunk c: one box [128 rows x 64 k] per half
//              FF2      8 chunks: tile t (128 outputs) x k-chunk c of this CTA's K slice [256 rank, +256)
//   classifier 8 chunks per round of 256 words: tile t x k-chunk c; rows beyond the vocabulary are zero-filled by TMA
// Chunk g (running index) lives in ring stage g % kStages.
// weight-ring load: plain, or with an L2 eviction-priority hint when the launch asks for one (ClusterArgs::w_policy != 0)
__device__ __forceinline__ void tma_w(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar, unsigned long long policy) {
  if (policy == 0) tma_load_2d(dst, map, x, y, bar);
  else tma_load_2d_hint(dst, map, x, y, bar, policy);
}

// Chunk g into its ring stage.  (Pulling the boxes into L2 a few chunks ahead of the ring with cp.async.bulk.prefetch.tensor was
// tried and removed: the prefetches queue in front of the ring's own loads in the TMA unit -- issuer wait 160 us -> 1000 us per
// decode, kernel 3.85 -> 5.2 ms at every distance from 4 to 23 chunks.)
template <int kStages>
__device__ __noinline__ void issue_chunk(uint8_t* ring, unsigned long long* full, const ClusterArgs& a, uint32_t g, int rank,
                                         int v0, int chunks_per_step) {
  const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(a.tmaps);
  const int s = (int)(g % kStages);
  const int idx = (int)(g % (uint32_t)chunks_per_step);
  const uint32_t dst = smem_addr(ring + (size_t)s * kStageBytes);
  const uint32_t bar = smem_addr(&full[s]);
  auto box = [&](uint32_t to, const CUtensorMap* m, int x, int y) { tma_w(to, m, x, y, bar, a.w_policy); };
  if (idx < kCLayers * kChunksPerLayer) {
    const int l = idx / kChunksPerLayer, j = idx - l * kChunksPerLayer;
    const CUtensorMap* lm = maps + kDecMapsPerLayer * l;
    if (j < 4) {  // QKV
      mbar_expect_tx(bar, 6 * 4096);
#pragma unroll
      for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int part = 0; part < 3; ++part)
          box(dst + half * kHalfStage + part * 4096, lm + 0 + half, j * 64, part * kCD + rank * kCHead);
    } else if (j == 4 || j == 6) {  // sa_out / ca_out (head-packed, K = 32)
      mbar_expect_tx(bar, 4 * 8192);
      const CUtensorMap* m = lm + (j == 4 ? 2 : 6);
#pragma unroll
      for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int t = 0; t < 2; ++t) box(dst + half * kHalfStage + t * 8192, m + half, 0, rank * kCD + t * 128);
    } else if (j == 5) {  // ca_q
      mbar_expect_tx(bar, 8 * 4096);
#pragma unroll
      for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int c = 0; c < 4; ++c) box(dst + half * kHalfStage + c * 4096, lm + 4 + half, c * 64, rank * kCHead);
    } else if (j < 15) {  // FF1: hidden units [256 rank + 128 t, +128)
      const int t = (j - 7) >> 2, c = (j - 7) & 3;
      mbar_expect_tx(bar, 2 * kHalfStage);
#pragma unroll
      for (int half = 0; half < 2; ++half) box(dst + half * kHalfStage, lm + 8 + half, c * 64, rank * kCD + t * 128);
    } else {  // FF2: outputs [128 t, +128), this CTA's K slice
      const int t = (j - 15) >> 2, c = (j - 15) & 3;
      mbar_expect_tx(bar, 2 * kHalfStage);
#pragma unroll
      for (int half = 0; half < 2; ++half) box(dst + half * kHalfStage, lm + 10 + half, rank * kCD + c * 64, t * 128);
    }
  } else {  // classifier
    const int jj = idx - kCLayers * kChunksPerLayer;
    const int rd = jj >> 3, t = (jj >> 2) & 1, c = jj & 3;
    mbar_expect_tx(bar, 2 * kHalfStage);
#pragma unroll
    for (int half = 0; half < 2; ++half)
      box(dst + half * kHalfStage, maps + kCLayers * kDecMapsPerLayer + half, c * 64, v0 + rd * kClsRound + t * 128);
  }
}

enum GemmKind : int { G_QKV = 0, G_P32 = 1, G_OUT = 2, G_FF = 3 };

// One GEMM phase on the tensor cores.  `b_addr` = shared-memory address of the B operand (hi/lo activations); accumulator tile t
// of the phase lives in TMEM slot slot0 + t.  epi(t, quarter, lane, r0, v[16]) is called by the warps that own accumulator rows
// [32 quarter, +32) of tile t, once per block of 16 decoder rows: v[j] = sum_k W[row][k] * x[r0 + j][k] (hi + 2^-11 lo).
// Ends with all threads having finished their TMEM reads (callers __syncthreads() before touching what the epilogue wrote).
template <int NR, typename Epi>
__device__ __forceinline__ void tc_gemm(CSmem<NR>& S, const ClusterArgs& a, Pipe& pp, uint32_t tmem_base, uint32_t& tile_par,
                                        int kind, int n_chunks, uint32_t b_addr, int slot0, int rank, int v0,
                                        int chunks_per_step, Epi epi) {
  constexpr int kStages = CSmem<NR>::kStages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  fence_proxy_async_smem();  // activations were written through the generic / st.async path: make them visible to the MMA
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const int n_tiles = kind == G_FF ? n_chunks / 4 : (kind == G_OUT ? 2 : 1);
  if (warp == kProducerWarp) {
    // producer: this phase's chunks, then run ahead into the next phases' weights as far as the ring allows.  Every wait is
    // on MMAs that the issuer issues without depending on this warp beyond the current phase: no deadlock.
    const uint32_t target = pp.use + (uint32_t)n_chunks + (kStages - 1);
    while (pp.load < target) {
      const int s = (int)(pp.load % kStages);
      mbar_wait(smem_addr(&S.empty[s]), ((pp.load / kStages) & 1u) ^ 1u);  // the MMAs that read this stage have completed
      if (elect_one()) issue_chunk<kStages>(&S.ring[0][0], S.full, a, pp.load, rank, v0, chunks_per_step);
      __syncwarp();
      ++pp.load;
    }
  } else if (warp == kIssuerWarp) {
    // MMA issuer: the whole warp runs the loop with warp-uniform values, one elected lane issues (tc_ptx.cuh elect_one)
    constexpr uint32_t idesc_hl = make_idesc(128, 2 * NR), idesc_l = make_idesc(128, NR);
    const uint32_t tb = __shfl_sync(kFull, tmem_base, 0);
    for (int i = 0; i < n_chunks; ++i) {
      const uint32_t u = pp.use + (uint32_t)i;
      const int s = (int)(u % kStages);
      if (a.trace != nullptr && blockIdx.x == 0) {
        const long long c0 = clock64();
        mbar_wait(smem_addr(&S.full[s]), (u / kStages) & 1u);
        if (lane == 0) S.tr_wait_full += (unsigned long long)(clock64() - c0);
      } else {
        mbar_wait(smem_addr(&S.full[s]), (u / kStages) & 1u);
      }
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_addr(&S.ring[s][0]);
        if (kind == G_OUT) {
          const uint64_t bdesc = make_smem_desc_sw64(b_addr);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint64_t a1 = make_smem_desc_sw64(st + t * 8192), a2 = make_smem_desc_sw64(st + kHalfStage + t * 8192);
            const uint32_t d = tb + (uint32_t)((slot0 + t) * 2 * NR);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              tcgen05_mma_f16(d, a1 + 2 * k, bdesc + 2 * k, idesc_hl, k != 0);
              tcgen05_mma_f16(d + NR, a2 + 2 * k, bdesc + 2 * k, idesc_l, 1u);
            }
            tcgen05_commit(smem_addr(&S.tile_full[slot0 + t]));
          }
        } else if (kind == G_P32) {
          const uint32_t d = tb + (uint32_t)(slot0 * 2 * NR);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint64_t a1 = make_smem_desc(st + c * 4096), a2 = make_smem_desc(st + kHalfStage + c * 4096);
            const uint64_t bdesc = make_smem_desc(b_addr + c * (2 * NR * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              tcgen05_mma_f16(d, a1 + 2 * k, bdesc + 2 * k, idesc_hl, (c | k) != 0);
              tcgen05_mma_f16(d + NR, a2 + 2 * k, bdesc + 2 * k, idesc_l, 1u);
            }
          }
          tcgen05_commit(smem_addr(&S.tile_full[slot0]));
        } else {  // G_QKV: chunk i = k-chunk i of the single tile; G_FF: chunk i = (tile i / 4, k-chunk i % 4)
          const int t = kind == G_FF ? (i >> 2) : 0, c = kind == G_FF ? (i & 3) : i;
          const uint32_t d = tb + (uint32_t)((slot0 + t) * 2 * NR);
          const uint64_t a1 = make_smem_desc(st), a2 = make_smem_desc(st + kHalfStage);
          const uint64_t bdesc = make_smem_desc(b_addr + c * (2 * NR * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tcgen05_mma_f16(d, a1 + 2 * k, bdesc + 2 * k, idesc_hl, (c | k) != 0);
            tcgen05_mma_f16(d + NR, a2 + 2 * k, bdesc + 2 * k, idesc_l, 1u);
          }
          if (c == 3) tcgen05_commit(smem_addr(&S.tile_full[slot0 + t]));
        }
        tcgen05_commit(smem_addr(&S.empty[s]));  // frees the stage once these MMAs have read it
      }
      __syncwarp();
    }
  }
  pp.use += (uint32_t)n_chunks;
  __syncwarp();
  // accumulator rows [32 q, +32) are only reachable from warps with id % 4 == q: warps 4 g .. 4 g + 3 read tiles g, g + 3, ...
  if (warp < 4 * kEpiGroups) {
    const int quarter = warp & 3;
    for (int t = warp >> 2; t < n_tiles; t += kEpiGroups) {
      const int slot = slot0 + t;
      mbar_wait(smem_addr(&S.tile_full[slot]), (tile_par >> slot) & 1u);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * 2 * NR);
#pragma unroll
      for (int r0 = 0; r0 < NR; r0 += 16) {
        float v[16], w[16];
        tmem_ld_32x16(taddr + r0, v);
        tmem_ld_32x16(taddr + NR + r0, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(w[j], kLoInv, v[j]);
        epi(t, quarter, lane, r0, v);
      }
    }
  }
  tile_par ^= ((1u << n_tiles) - 1u) << slot0;  // every thread tracks the phase parity of every accumulator slot
  tcgen05_fence_before();
}

// counters of completed exchanges (parity of the mbarrier phases); identical in every thread of every CTA of the cluster
struct ExCount {
  uint32_t rs = 0, ag = 0, beam = 0;
};

// Epilogue of a K-split GEMM (sa_out / ca_out / FF2): partial sum of output column n = 128 t + 32 quarter + lane goes to the CTA
// that owns the column (n / 32) -- slab `rank` of its reduce-scatter zone.
template <int NR>
__device__ __forceinline__ void rs_push(CSmem<NR>& S, int rank, int t, int quarter, int lane, int r0, const float (&v)[16]) {
  const int owner = 4 * t + quarter;
  if (owner == rank) {
#pragma unroll
    for (int j = 0; j < 16; ++j) S.recv[rank][r0 + j][lane] = v[j];
  } else {
    const uint32_t dst = map_peer(smem_addr(&S.recv[rank][r0][lane]), owner);
    const uint32_t bar = map_peer(smem_addr(&S.bar_rs), owner);
#pragma unroll
    for (int j = 0; j < 16; ++j) st_async4(dst + j * (kCHead * 4), v[j], bar);
  }
}

// After a K-split GEMM: wait for the 7 peers' partial sums, x_pre[r][own 32 columns] = residual + bias + sum of the 8 partials
// (fixed order), all-gather the pre-LayerNorm slices into every CTA's staging area (= opx), LayerNorm all rows in place:
// opx becomes the hi/lo operand of the next GEMM, xr the exact fp32 residual slice.
template <int NR>
__device__ __forceinline__ void reduce_gather_ln(CSmem<NR>& S, ExCount& ex, int rank, const float* __restrict__ bias,
                                                 const float* __restrict__ g, const float* __restrict__ b, bool tr_on) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kBytes = (kCl - 1) * NR * kCHead * 4;
  const uint32_t bar_rs = smem_addr(&S.bar_rs), bar_ag = smem_addr(&S.bar_ag);
  if (tid == 0) {
    mbar_expect_tx(bar_rs, kBytes);
    mbar_expect_tx(bar_ag, kBytes);
  }
  mbar_wait(bar_rs, ex.rs & 1u);
  ++ex.rs;
  __syncthreads();  // this CTA's own slab was written with ordinary stores by the epilogue warps
  CL_TR(15);
  float (*stg)[kCD] = S.stage();
  for (int idx = tid; idx < NR * 8; idx += kCThreads) {
    const int r = idx >> 3, q4 = idx & 7;
    float4 y = __ldg(reinterpret_cast<const float4*>(bias + rank * kCHead) + q4);
    const float4 x = *reinterpret_cast<const float4*>(&S.xr[r][4 * q4]);
    y.x += x.x; y.y += x.y; y.z += x.z; y.w += x.w;
#pragma unroll
    for (int i = 0; i < kCl; ++i) {  // fixed order
      const float4 p = *reinterpret_cast<const float4*>(&S.recv[i][r][4 * q4]);
      y.x += p.x; y.y += p.y; y.z += p.z; y.w += p.w;
    }
    float* dst = &stg[r][rank * kCHead + 4 * q4];
    *reinterpret_cast<float4*>(dst) = y;
    const uint32_t da = smem_addr(dst);
#pragma unroll
    for (int p = 1; p < kCl; ++p) {
      const int peer = (rank + p) & (kCl - 1);
      st_async16(map_peer(da, peer), y, map_peer(bar_ag, peer));
    }
  }
  CL_TR(16);
  mbar_wait(bar_ag, ex.ag & 1u);
  ++ex.ag;
  __syncthreads();  // own slice (ordinary stores) visible to every warp
  CL_TR(17);
  // LayerNorm (eps 1e-5, biased variance): one warp per row, lane holds columns [8 lane, +8)
  float v[NR / kCWarps][8];
#pragma unroll
  for (int rr = 0; rr < NR / kCWarps; ++rr) {
    const int r = warp + kCWarps * rr;
    const float4 p0 = *reinterpret_cast<const float4*>(&stg[r][8 * lane]);
    const float4 p1 = *reinterpret_cast<const float4*>(&stg[r][8 * lane + 4]);
    const float x[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j];
    const float mean = warp_sum(s) * (1.f / kCD);
    float qq = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) qq += (x[j] - mean) * (x[j] - mean);
    const float rstd = 1.f / sqrtf(warp_sum(qq) * (1.f / kCD) + 1e-5f);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g) + 2 * lane), g1 = __ldg(reinterpret_cast<const float4*>(g) + 2 * lane + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + 2 * lane), b1 = __ldg(reinterpret_cast<const float4*>(b) + 2 * lane + 1);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[rr][j] = (x[j] - mean) * rstd * gg[j] + bb[j];
  }
  __syncthreads();  // every row has been read: the staging bytes may now be overwritten by the operand layout
#pragma unroll
  for (int rr = 0; rr < NR / kCWarps; ++rr) {
    const int r = warp + kCWarps * rr;
    uint4 hi, lo;
    split8(v[rr], hi, lo);
    const int off = op_off<NR>(r, 8 * lane);
    *reinterpret_cast<uint4*>(S.opx + off) = hi;
    *reinterpret_cast<uint4*>(S.opx + NR * 128 + off) = lo;
    if ((lane >> 2) == rank) {
      *reinterpret_cast<float4*>(&S.xr[r][8 * (lane & 3)]) = make_float4(v[rr][0], v[rr][1], v[rr][2], v[rr][3]);
      *reinterpret_cast<float4*>(&S.xr[r][8 * (lane & 3) + 4]) = make_float4(v[rr][4], v[rr][5], v[rr][6], v[rr][7]);
    }
  }
}

// x[r] = emb[token of row r at `pos`] * 16 + PE[pos] -> opx (hi/lo operand) + xr (own residual slice); rows >= R are zero
template <int NR>
__device__ __forceinline__ void embed_rows(CSmem<NR>& S, const ClusterArgs& a, int rank, int R, int cur, int pos) {
  for (int i = threadIdx.x; i < NR * 32; i += kCThreads) {
    const int r = i >> 5, ch = i & 31;  // columns [8 ch, +8)
    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r < R) {
      const int tok = S.tokens[cur][r][pos];
      const float4* e = reinterpret_cast<const float4*>(a.emb + (int64_t)tok * kCD) + 2 * ch;
      const float4* p = reinterpret_cast<const float4*>(a.pe + (int64_t)pos * kCD) + 2 * ch;
      const float4 e0 = __ldg(e), e1 = __ldg(e + 1), p0 = __ldg(p), p1 = __ldg(p + 1);
      x[0] = fmaf(e0.x, 16.f, p0.x); x[1] = fmaf(e0.y, 16.f, p0.y); x[2] = fmaf(e0.z, 16.f, p0.z); x[3] = fmaf(e0.w, 16.f, p0.w);
      x[4] = fmaf(e1.x, 16.f, p1.x); x[5] = fmaf(e1.y, 16.f, p1.y); x[6] = fmaf(e1.z, 16.f, p1.z); x[7] = fmaf(e1.w, 16.f, p1.w);
    }
    uint4 hi, lo;
    split8(x, hi, lo);
    const int off = op_off<NR>(r, 8 * ch);
    *reinterpret_cast<uint4*>(S.opx + off) = hi;
    *reinterpret_cast<uint4*>(S.opx + NR * 128 + off) = lo;
    if ((ch >> 2) == rank) {
      *reinterpret_cast<float4*>(&S.xr[r][8 * (ch & 3)]) = make_float4(x[0], x[1], x[2], x[3]);
      *reinterpret_cast<float4*>(&S.xr[r][8 * (ch & 3) + 4]) = make_float4(x[4], x[5], x[6], x[7]);
    }
  }
}

// One decoder layer for the rows of this cluster (all 8 CTAs execute it in lock step through the six exchanges).
template <int NR>
__device__ __noinline__ void decode_layer(CSmem<NR>& S, const ClusterArgs& a, Pipe& pp, ExCount& ex, uint32_t tmem_base,
                                          uint32_t& tile_par, int l, int rank, int R, int grow0, int clip0, int pos, int cur,
                                          int v0, int chunks_per_step, bool tr_on) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int beam = a.beam, max_len = a.max_len, tp = a.tp;
  const int64_t cache_l = (int64_t)a.rows * max_len * kCD;
  const int64_t kv_stride = (int64_t)kCLayers * 2 * kCD;
  const ClusterLayer& L = a.layers[l];
  constexpr int NRW = NR / kCWarps;  // rows per warp in the attention phases
  float (*Q)[kCHead] = S.q();
  float (*KV)[2 * kCHead] = S.kv();
  const uint32_t opx = smem_addr(S.opx), oph = smem_addr(S.oph), opa = smem_addr(S.opa);

  // ---- P1: q | k | v of head `rank` (96 weight rows), then self-attention for the rows of this head
  tc_gemm<NR>(S, a, pp, tmem_base, tile_par, G_QKV, 4, opx, 0, rank, v0, chunks_per_step,
              [&](int, int quarter, int d, int r0, const float (&v)[16]) {
                if (quarter >= 3) return;
                const float bias = __ldg(L.sa_in_b + quarter * kCD + rank * kCHead + d);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (quarter == 0) Q[r0 + j][d] = v[j] + bias;
                  else KV[r0 + j][(quarter - 1) * kCHead + d] = v[j] + bias;
                }
              });
  __syncthreads();
  CL_TR(0);
  {
    float* kc = a.kc + l * cache_l;
    float* vc = a.vc + l * cache_l;
    const float* qq[NRW];
    const float* kvn[NRW];
    int orow[NRW], nn[NRW];
    bool valid[NRW];
    const uint8_t* s0[NRW];
#pragma unroll
    for (int rr = 0; rr < NRW; ++rr) {
      const int r = warp + kCWarps * rr;
      valid[rr] = r < R;
      qq[rr] = &Q[r][0];
      kvn[rr] = &KV[r][0];
      orow[rr] = r;
      nn[rr] = pos;
      s0[rr] = &S.src[cur][r][0];
      if (valid[rr]) {  // the new position goes to this head's slice of the global cache (read back in later steps only)
        const int64_t o = ((int64_t)(grow0 + r) * max_len + pos) * kCD + rank * kCHead + lane;
        stg_f32_pol(kc + o, kvn[rr][lane], a.kv_policy);
        stg_f32_pol(vc + o, kvn[rr][kCHead + lane], a.kv_policy);
      }
    }
    auto kp = [&](int rr, int j) { return kc + ((int64_t)(grow0 + s0[rr][j]) * max_len + j) * kCD + rank * kCHead; };
    auto vp = [&](int rr, int j) { return vc + ((int64_t)(grow0 + s0[rr][j]) * max_len + j) * kCD + rank * kCHead; };
    if (valid[0]) {  // warp-uniform (rows of a warp: r, r + 16)
      if (pos <= 32) attend<NR, NRW, 1, false>(qq, nn, valid, kp, vp, kvn, true, S.opa, orow, lane, a.kv_policy);
      else attend<NR, NRW, 2, false>(qq, nn, valid, kp, vp, kvn, true, S.opa, orow, lane, a.kv_policy);
    }
  }
  CL_TR(1);
  // ---- P2: self-attention output projection, K-split over the heads; reduce-scatter, + bias + residual, all-gather, LayerNorm 1
  tc_gemm<NR>(S, a, pp, tmem_base, tile_par, G_OUT, 1, opa, 0, rank, v0, chunks_per_step,
              [&](int t, int quarter, int ln, int r0, const float (&v)[16]) { rs_push<NR>(S, rank, t, quarter, ln, r0, v); });
  CL_TR(2);
  reduce_gather_ln<NR>(S, ex, rank, L.sa_out_b, L.n1_g, L.n1_b, tr_on);
  CL_TR(3);
  // ---- P3: cross-attention query of head `rank`, cross-attention over the clip's encoder frames
  tc_gemm<NR>(S, a, pp, tmem_base, tile_par, G_P32, 1, opx, 0, rank, v0, chunks_per_step,
              [&](int, int quarter, int j, int r0, const float (&v)[16]) {
                if (quarter != 0) return;
                const float bias = __ldg(L.ca_q_b + rank * kCHead + j);
#pragma unroll
                for (int i = 0; i < 16; ++i) Q[r0 + i][j] = v[i] + bias;
              });
  __syncthreads();
  CL_TR(4);
  {
    const float* ck = a.ckv + (int64_t)l * 2 * kCD + rank * kCHead;  // key-padding mask: frames >= len are skipped
    const float* qq[NRW];
    const float* kvn[NRW];
    int orow[NRW], nn[NRW], c0[NRW];
    bool valid[NRW];
#pragma unroll
    for (int rr = 0; rr < NRW; ++rr) {
      const int r = warp + kCWarps * rr;
      valid[rr] = r < R;
      qq[rr] = &Q[r][0];
      kvn[rr] = nullptr;
      orow[rr] = r;
      c0[rr] = clip0 + (valid[rr] ? r / beam : 0);
      nn[rr] = valid[rr] ? min(a.lens[c0[rr]], tp) : 0;
    }
    const int tpad = a.tpad;
    const float* kt = a.ckt + ((int64_t)l * kCl + rank) * kCHead * tpad;  // [clip][layer][head][32][tpad]
    auto kp = [&](int rr, int) { return kt + (int64_t)c0[rr] * kCLayers * kCD * tpad; };
    auto vp = [&](int rr, int j) { return ck + ((int64_t)c0[rr] * tp + j) * kv_stride + kCD; };
    if (valid[0]) {
      if (tp <= 32) attend<NR, NRW, 1, true>(qq, nn, valid, kp, vp, kvn, false, S.opa, orow, lane, a.kv_policy, tpad);
      else if (tp <= 64) attend<NR, NRW, 2, true>(qq, nn, valid, kp, vp, kvn, false, S.opa, orow, lane, a.kv_policy, tpad);
      else attend<NR, NRW, 4, true>(qq, nn, valid, kp, vp, kvn, false, S.opa, orow, lane, a.kv_policy, tpad);
    }
  }
  CL_TR(5);
  // ---- P4: cross-attention output projection (K-split), reduce-scatter, all-gather, LayerNorm 2
  tc_gemm<NR>(S, a, pp, tmem_base, tile_par, G_OUT, 1, opa, 0, rank, v0, chunks_per_step,
              [&](int t, int quarter, int ln, int r0, const float (&v)[16]) { rs_push<NR>(S, rank, t, quarter, ln, r0, v); });
  CL_TR(6);
  reduce_gather_ln<NR>(S, ex, rank, L.ca_out_b, L.n2_g, L.n2_b, tr_on);
  CL_TR(7);
  // ---- P5: FF1 slice (256 hidden units of this CTA) + GELU -> oph (hi/lo operand)
  tc_gemm<NR>(S, a, pp, tmem_base, tile_par, G_FF, 8, opx, 0, rank, v0, chunks_per_step,
              [&](int t, int quarter, int d, int r0, const float (&v)[16]) {
                const int j = t * 128 + quarter * 32 + d;  // hidden unit = k index of the FF2 operand
                const float bias = __ldg(L.l1_b + rank * kCD + j);
                uint8_t* base = S.oph + (j >> 6) * (2 * NR * 128) + (j & 7) * 2;
                const int grp = (j >> 3) & 7;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int r = r0 + i;
                  const float h = gelu_erf(v[i] + bias);
                  const act16 h1 = float2act(h);
                  const act16 h2 = float2act((h - act2float(h1)) * kLoScale);
                  uint8_t* p = base + r * 128 + ((grp ^ (r & 7)) << 4);
                  *reinterpret_cast<act16*>(p) = h1;
                  *reinterpret_cast<act16*>(p + NR * 128) = h2;
                }
              });
  CL_TR(8);
  // ---- P6: FF2 partial sums over this CTA's K slice for all 256 outputs, reduce-scatter, + bias + residual, all-gather, LN 3
  tc_gemm<NR>(S, a, pp, tmem_base, tile_par, G_FF, 8, oph, 0, rank, v0, chunks_per_step,
              [&](int t, int quarter, int ln, int r0, const float (&v)[16]) { rs_push<NR>(S, rank, t, quarter, ln, r0, v); });
  CL_TR(9);
  reduce_gather_ln<NR>(S, ex, rank, L.l2_b, L.n3_g, L.n3_b, tr_on);
  CL_TR(10);
}

// Classifier slice (rounds of 256 words) + distributed beam step (one exchange).
template <int NR>
__device__ __noinline__ void decode_select(CSmem<NR>& S, const ClusterArgs& a, Pipe& pp, ExCount& ex, uint32_t tmem_base,
                                           uint32_t& tile_par, int rank, int R, int grow0, int nclips, int step, int cur,
                                           int vs, int chunks_per_step, bool tr_on) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int beam = a.beam, max_len = a.max_len, V = a.vocab;
  const int v0 = rank * vs;
  const int n_rounds = vs / kClsRound;
  float (*LT)[kClsRound] = S.lt();
  // ---- classifier rounds: logits tile -> running per-row max / sum-exp / top-`beam` words of this vocabulary slice
  for (int rd = 0; rd < n_rounds; ++rd) {
    const int base = v0 + rd * kClsRound;
    tc_gemm<NR>(S, a, pp, tmem_base, tile_par, G_FF, 8, smem_addr(S.opx), (rd & 1) * 2, rank, v0, chunks_per_step,
                [&](int t, int quarter, int d, int r0, const float (&v)[16]) {
                  const int j = t * 128 + quarter * 32 + d;
                  const int w = base + j;
                  const float bias = w < V ? __ldg(a.cls_b + w) : 0.f;
#pragma unroll
                  for (int i = 0; i < 16; ++i) LT[r0 + i][j] = w < V ? v[i] + bias : -INFINITY;
                });
    __syncthreads();
    CL_TR(11);
    for (int r = warp; r < R; r += kCWarps) {
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = LT[r][lane + 32 * u];
      if (a.tap != nullptr) {  // raw logits (before the masks), reference seam AACDecoder.__call__
        float* tp_ = a.tap + ((int64_t)step * a.rows + grow0 + r) * V;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (base + lane + 32 * u < V) tp_[base + lane + 32 * u] = x[u];
      }
      if (step < a.min_len) {  // beam.py:129-130
        const int e = kCEos - base;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (e == lane + 32 * u) x[u] = -INFINITY;
      }
      if (a.forbid != nullptr) {  // beam.py:146-156: lane p fetches history token p and its flag, then all lanes walk the list
        for (int p0 = 0; p0 <= step; p0 += 32) {
          int my_e = -1;
          if (p0 + lane <= step) {
            const int tok = S.tokens[cur][r][p0 + lane];
            const int e = tok - base;
            if (e >= 0 && e < kClsRound && a.forbid[tok]) my_e = e;
          }
          unsigned hit = __ballot_sync(kFull, my_e >= 0);
          while (hit) {
            const int src = __ffs(hit) - 1;
            hit &= hit - 1;
            const int e = __shfl_sync(kFull, my_e, src);
            if ((e & 31) == lane) {
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (u == (e >> 5)) x[u] = -INFINITY;
            }
          }
        }
      }
      float mx = x[0];
#pragma unroll
      for (int u = 1; u < 8; ++u) mx = fmaxf(mx, x[u]);
      mx = warp_max(mx);
      float sm = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) sm += (x[u] == -INFINITY) ? 0.f : expf(x[u] - mx);
      sm = warp_sum(sm);
      if (mx == -INFINITY) sm = 0.f;
      // the previous rounds' list joins as one extra candidate per lane (lane k holds entry k)
      unsigned long long old_key = 0ull;
      if (rd > 0 && lane < beam) {
        const CCand oc = S.st_cnd[r][lane];
        old_key = ckey(oc.v, oc.idx);
      }
      float om = -INFINITY, os = 0.f;
      if (rd > 0) {
        om = S.st_stat[r][0];
        os = S.st_stat[r][1];
      }
      __syncwarp();
      if (lane == 0) {
        const float m2 = fmaxf(om, mx);
        float s2 = 0.f;
        if (m2 != -INFINITY) s2 = (om == -INFINITY ? 0.f : os * expf(om - m2)) + (mx == -INFINITY ? 0.f : sm * expf(mx - m2));
        S.st_stat[r][0] = m2;
        S.st_stat[r][1] = s2;
      }
      // `beam` rounds of: every lane's best candidate strictly after the previous winner, then a warp arg-max (64-bit keys)
      unsigned long long keys[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) keys[u] = (base + lane + 32 * u < V) ? ckey(x[u], base + lane + 32 * u) : 0ull;
      unsigned long long prev = ~0ull;
      for (int k = 0; k < beam; ++k) {
        unsigned long long best = old_key < prev ? old_key : 0ull;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const unsigned long long c = keys[u] < prev ? keys[u] : 0ull;
          best = c > best ? c : best;
        }
        best = warp_max_u64(best);
        if (lane == 0) S.st_cnd[r][k] = ckey_decode(best);
        prev = best;  // 0 = exhausted (or NaN logits): every later round emits the sentinel too
      }
      __syncwarp();
    }
    CL_TR(12);
  }
  __syncthreads();
  {
    const uint32_t bar = smem_addr(&S.bar_beam);
    constexpr int kUnits = 1 + kCMaxBeam;  // 8-byte units per row: (max, sum-exp) + kCMaxBeam candidates
    if (tid == 0) mbar_expect_tx(bar, (kCl - 1) * NR * kUnits * 8);
    float (*STAT)[NR][2] = S.stat();
    CCand (*CND)[NR][kCMaxBeam] = S.cnd();
    for (int idx = tid; idx < kCl * NR * kUnits; idx += kCThreads) {
      const int p = idx / (NR * kUnits), rem = idx - p * (NR * kUnits);
      const int r = rem / kUnits, u = rem - r * kUnits;
      const int peer = (rank + p) & (kCl - 1);
      const uint32_t* srcw = u == 0 ? reinterpret_cast<const uint32_t*>(&S.st_stat[r][0])
                                    : reinterpret_cast<const uint32_t*>(&S.st_cnd[r][u - 1]);
      void* dst = u == 0 ? static_cast<void*>(&STAT[rank][r][0]) : static_cast<void*>(&CND[rank][r][u - 1]);
      if (peer == rank) {
        reinterpret_cast<uint32_t*>(dst)[0] = srcw[0];
        reinterpret_cast<uint32_t*>(dst)[1] = srcw[1];
      } else {
        st_async8(map_peer(smem_addr(dst), peer), srcw[0], srcw[1], map_peer(bar, peer));
      }
    }
    mbar_wait(bar, ex.beam & 1u);
    ++ex.beam;
    __syncthreads();
  }
  CL_TR(13);
  // ---- beam step, part B (replicated): merge, flat top-k per clip, history / back-pointer update, finish bookkeeping
  float (*STAT)[NR][2] = S.stat();
  CCand (*CND)[NR][kCMaxBeam] = S.cnd();
  const int nxt = cur ^ 1;
  for (int lc = warp; lc < nclips; lc += kCWarps) {
    const int r0 = lc * beam;
    int live_label[kCMaxBeam];
    float prev_sum[kCMaxBeam];
    int nlive = 0;
#pragma unroll
    for (int q = 0; q < kCMaxBeam; ++q) {
      live_label[q] = 0;
      prev_sum[q] = 0.f;
    }
#pragma unroll
    for (int lb = 0; lb < kCMaxBeam; ++lb)
      if (lb < beam && S.live[r0 + lb]) {
#pragma unroll
        for (int q = 0; q < kCMaxBeam; ++q)
          if (q == nlive) {
            live_label[q] = lb;
            prev_sum[q] = S.sum_lp[r0 + lb];
          }
        ++nlive;
      }
    if (nlive == 0) continue;  // warp-uniform
    const int nrows_used = (step == 0) ? 1 : nlive;  // step 0: only the first row (beam.py:243-246)
    const int k_sel = nlive;
    auto label_at = [&](int q) {
      int r = 0;
#pragma unroll
      for (int i = 0; i < kCMaxBeam; ++i)
        if (i == q) r = live_label[i];
      return r;
    };
    // log-sum-exp of every used row from the 8 slice statistics (fixed order)
    float row_mx[kCMaxBeam], row_lg[kCMaxBeam];
#pragma unroll
    for (int j = 0; j < kCMaxBeam; ++j) {
      row_mx[j] = 0.f;
      row_lg[j] = 0.f;
      if (j < nrows_used) {
        const int r = r0 + label_at(j);
        float m = STAT[0][r][0];
#pragma unroll
        for (int i = 1; i < kCl; ++i) m = fmaxf(m, STAT[i][r][0]);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kCl; ++i) s += STAT[i][r][1] * expf(STAT[i][r][0] - m);
        row_mx[j] = m;
        row_lg[j] = logf(s);
      }
    }
    // candidates: lane = (peer i = lane & 7, k in {lane >> 3, (lane >> 3) + 4}) for every used row j; value = cumulative score,
    // index = j * V + word (beam.py:256-263).  k_sel rounds of "best candidate strictly after the previous winner".
    const int pi = lane & 7, kg = lane >> 3;
    unsigned long long prev_key = ~0ull;
    for (int rsel = 0; rsel < k_sel; ++rsel) {
      unsigned long long best = 0ull;
#pragma unroll
      for (int j = 0; j < kCMaxBeam; ++j) {
        if (j < nrows_used) {
          const int r = r0 + label_at(j);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int k = kg + 4 * kk;
            if (k < beam) {
              const CCand raw = CND[pi][r][k];
              const float lsm = (raw.v - row_mx[j]) - row_lg[j];
              unsigned long long c = raw.idx == 0x7fffffff ? 0ull : ckey(step == 0 ? lsm : prev_sum[j] + lsm, j * V + raw.idx);
              c = c < prev_key ? c : 0ull;
              best = c > best ? c : best;
            }
          }
        }
      }
      best = warp_max_u64(best);
      prev_key = best;
      CCand w = ckey_decode(best);
      if (w.idx == 0x7fffffff) w.idx = 0;  // NaN logits (fully masked clip, len 0): stay memory-safe like a no-op pick
      if (lane == 0) {
        const int pj = w.idx / V;  // row of the previous beam, word
        S.win[warp][rsel] = w;
        S.winj[warp][rsel] = make_int2(pj, w.idx - pj * V);
      }
    }
    __syncwarp();
    // candidate r -> r-th live label (beam.py:165-176); histories via back-pointers: lane = position
    for (int r = 0; r < k_sel; ++r) {
      const int row = r0 + label_at(r);
      const int2 pw = S.winj[warp][r];
      const int srow = r0 + label_at(pw.x);
      for (int p = lane; p <= step + 1; p += 32) {
        if (p <= step) {
          S.tokens[nxt][row][p] = S.tokens[cur][srow][p];
          if (p < max_len) S.src[nxt][row][p] = S.src[cur][srow][p];
        } else {
          S.tokens[nxt][row][p] = (uint16_t)pw.y;
          if (p < max_len) S.src[nxt][row][p] = (uint8_t)row;
        }
      }
    }
    __syncwarp();
    if (lane < k_sel) {
      const int r = lane;
      const int row = r0 + label_at(r);
      const CCand w = S.win[warp][r];
      const int word = S.winj[warp][r].y;
      S.sum_lp[row] = w.v;
      if (word == kCEos || step == max_len - 1) {  // beam.py:173-190
        if (rank == 0) {
          for (int p = 0; p <= step; ++p)
            a.bs.out_preds[(int64_t)(grow0 + row) * max_len + p] = S.tokens[nxt][row][p + 1];
          a.bs.out_lp[grow0 + row] = w.v / (float)(step + 1);
        }
        S.live[row] = 0;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  CL_TR(14);
}

template <int NR>
__global__ void __launch_bounds__(kCThreads, 1)
decoder_cluster_kernel(const __grid_constant__ ClusterArgs a, int clips_per_group, int n_groups, int vs /*vocabulary slice*/) {
  constexpr int kStages = CSmem<NR>::kStages;
  extern __shared__ uint8_t smem_raw[];
  // operand tiles need 1024-byte alignment (128B swizzle atoms); the offset is the same in every CTA of the launch
  uint8_t* smem_al = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  CSmem<NR>& S = *reinterpret_cast<CSmem<NR>*>(smem_al);

  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();  // = attention head owned by this CTA
  const int cluster_id = blockIdx.x / kCl, n_clusters = gridDim.x / kCl;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int beam = a.beam, max_len = a.max_len;
  const int v0 = rank * vs;
  const int chunks_per_step = kCLayers * kChunksPerLayer + (vs / kClsRound) * 8;
  int steps_max = 0;
  for (int i = tid; i < (int)(sizeof(CSmem<NR>) / 4); i += kCThreads) reinterpret_cast<uint32_t*>(smem_al)[i] = 0u;
  __syncthreads();
  if (tid == 0) {
    mbar_init(smem_addr(&S.bar_rs), 1);
    mbar_init(smem_addr(&S.bar_ag), 1);
    mbar_init(smem_addr(&S.bar_beam), 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_addr(&S.full[s]), 1);
      mbar_init(smem_addr(&S.empty[s]), 1);
    }
    for (int t = 0; t < kTileSlots; ++t) mbar_init(smem_addr(&S.tile_full[t]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async_smem();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S.tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&S.tmem_slot);
  // debug trace (CNB_DEC_TRACE): thread 0 of the first CTA accumulates the time between phase marks
  const bool tr_on = a.trace != nullptr && blockIdx.x == 0 && tid == 0;
  const long long tr_clk0 = clock64();
  if (tid == 0) S.tr_wait_full = 0ull;
  if (tr_on) S.tr_last = cl_global_ns();
  const unsigned long long tr_ns0 = tr_on ? S.tr_last : 0ull;
  cl.sync();  // every CTA's mbarriers are initialised before any peer pushes data at them
  Pipe pp;
  ExCount ex;
  uint32_t tile_par = 0;

  for (int g = cluster_id; g < n_groups; g += n_clusters) {
    const int clip0 = g * clips_per_group;
    const int nclips = min(clips_per_group, a.batch - clip0);
    const int R = nclips * beam;       // live local rows (<= NR)
    const int grow0 = clip0 * beam;    // first global row of the group

    // ---- init: beam state (replicated), outputs (rank 0), first embedding
    for (int i = tid; i < 2 * NR * (kCMaxLen + 2); i += kCThreads) (&S.tokens[0][0][0])[i] = kCPad;
    for (int i = tid; i < 2 * NR * kCMaxLen; i += kCThreads) (&S.src[0][0][0])[i] = (uint8_t)((i / kCMaxLen) % NR);
    if (tid < NR) {
      S.sum_lp[tid] = 0.f;
      S.live[tid] = tid < R ? 1 : 0;
    }
    __syncthreads();
    if (tid < R) S.tokens[0][tid][0] = (uint16_t)a.bos_ids[clip0 + tid / beam];
    if (rank == 0) {
      for (int i = tid; i < R * max_len; i += kCThreads) a.bs.out_preds[(int64_t)grow0 * max_len + i] = kCPad;
      if (tid < R) a.bs.out_lp[grow0 + tid] = 0.f;
    }
    __syncthreads();
    embed_rows<NR>(S, a, rank, R, 0, 0);
    __syncthreads();

    int cur = 0, steps_done = max_len;
    for (int step = 0; step < max_len; ++step) {
      for (int l = 0; l < kCLayers; ++l)
        decode_layer<NR>(S, a, pp, ex, tmem_base, tile_par, l, rank, R, grow0, clip0, step, cur, v0, chunks_per_step, tr_on);
      decode_select<NR>(S, a, pp, ex, tmem_base, tile_par, rank, R, grow0, nclips, step, cur, vs, chunks_per_step, tr_on);
      cur ^= 1;
      // The beam exchange landed in the reduce-scatter zone: no peer may start the next step's first reduce-scatter before
      // every CTA has finished merging.  Split cluster barrier: arrive here, wait right before the first K-split GEMM could
      // push (i.e. before the next layer 0) -- the embedding and the early-exit test run in its shadow.
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      // ---- continue?  (state is replicated, so every CTA of the cluster takes the same branch)
      if (tid == 0) {
        int any = 0;
        for (int r = 0; r < R; ++r) any |= S.live[r];
        S.any_live = any;
      }
      __syncthreads();
      const bool go_on = S.any_live != 0;
      if (go_on && step + 1 < max_len) embed_rows<NR>(S, a, rank, R, cur, step + 1);
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
      if (!go_on) {
        steps_done = step + 1;
        break;
      }
      __syncthreads();
    }
    steps_max = max(steps_max, steps_done);
    __syncthreads();
  }
  // drain the weight chunks that were requested ahead but never consumed
  if (warp == kProducerWarp)
    for (uint32_t u = pp.use; u < pp.load; ++u) mbar_wait(smem_addr(&S.full[u % kStages]), (u / kStages) & 1u);
  if (tr_on) {
    for (int i = 0; i < 18; ++i) a.trace[i] = S.tr_acc[i];
    a.trace[18] = (unsigned long long)(clock64() - tr_clk0);  // SM cycles / elapsed ns = the clock this SM really ran at
    a.trace[19] = cl_global_ns() - tr_ns0;
    a.trace[20] = S.tr_wait_full;
  }
  if (rank == 0 && tid == 0 && steps_max > 0) atomicMax(&a.bs.done[1], steps_max);
  tcgen05_fence_before();
  cl.sync();  // no CTA may exit while a peer can still push into its shared memory
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

template <int NR> size_t cluster_smem() { return sizeof(CSmem<NR>) + 1024; }

struct ClusterPlan {
  int nr = 0, clips_per_group = 0, n_groups = 0, vs = 0, n_clusters = 0;
  size_t smem = 0;
};

template <int NR> int max_clusters_for(int* out) {
  static int cached = -1;
  if (cached < 0) {
    const size_t smem = cluster_smem<NR>();
    CNB_CUDA_OK(cudaFuncSetAttribute(decoder_cluster_kernel<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t lc = {};
    lc.blockDim = dim3(kCThreads);
    lc.gridDim = dim3(kCl * 64);
    lc.dynamicSmemBytes = smem;
    cudaLaunchAttribute la[1];
    la[0].id = cudaLaunchAttributeClusterDimension;
    la[0].val.clusterDim.x = kCl;
    la[0].val.clusterDim.y = 1;
    la[0].val.clusterDim.z = 1;
    lc.attrs = la;
    lc.numAttrs = 1;
    int n = 0;
    CNB_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, decoder_cluster_kernel<NR>, &lc));
    cached = n;
    if (getenv("CNB_DEC_TRACE")) fprintf(stderr, "[dec cluster] NR %d: max active clusters %d, smem %zu B\n", NR, n, smem);
  }
  *out = cached;
  return 0;
}

// Group size policy.  Fewest rows per cluster = least attention / epilogue / LayerNorm work per step (those phases scale with
// the rows; measured: 15 rows take 4.0 ms in a 16-row cluster and 5.3 ms in a 32-row one), so the clips are spread over 16-row
// clusters as long as one wave of them fits the device; beyond that 32-row clusters (half as many groups, deeper per step) beat
// a second wave.  CNB_DEC_NR=16|32 overrides.  0 ok, 1 unsupported, < 0 error.
int cluster_plan(const ClusterArgs& a, ClusterPlan* p) {
  if (a.beam < 1 || a.beam > kCMaxBeam || a.max_len > kCMaxLen || a.tp > kCMaxTp || a.vocab <= 4 || a.vocab > 65535 ||
      a.tmaps == nullptr)
    return 1;
  p->vs = (int)ceil_div(ceil_div(a.vocab, kCl), kClsRound) * kClsRound;
  int mc16 = 0, mc32 = 0;
  if (int rc = max_clusters_for<16>(&mc16)) return rc;
  if (int rc = max_clusters_for<32>(&mc32)) return rc;
  if (mc16 <= 0 && mc32 <= 0) return 1;
  const char* env_nr = getenv("CNB_DEC_NR");
  const int forced = env_nr ? atoi(env_nr) : 0;
  const int need16 = (int)ceil_div(a.batch, 16 / a.beam);
  int nr = need16 > mc16 ? 32 : 16;
  if (forced == 16 || forced == 32) nr = forced;
  if (nr == 32 && (mc32 <= 0 || a.batch * a.beam <= 16)) nr = 16;
  if (nr == 16 && mc16 <= 0) nr = 32;
  const int mc = nr == 16 ? mc16 : mc32;
  const int gmax = nr / a.beam;
  int g = (int)ceil_div(a.batch, mc);
  if (g > gmax || getenv("CNB_DEC_FILL")) g = gmax;  // CNB_DEC_FILL: experiments with full clusters
  if (const char* eg = getenv("CNB_DEC_GROUP")) g = std::max(g, std::min(atoi(eg), gmax));  // ... or a given clips per cluster
  p->nr = nr;
  p->clips_per_group = g;
  p->n_groups = (int)ceil_div(a.batch, g);
  p->n_clusters = p->n_groups < mc ? p->n_groups : mc;
  p->smem = nr == 16 ? cluster_smem<16>() : cluster_smem<32>();
  return 0;
}

}  // namespace

bool decoder_cluster_supported(const ClusterArgs& a) {
  ClusterPlan p;
  return cluster_plan(a, &p) == 0;
}

int decoder_cluster_sms(const ClusterArgs& a) {
  ClusterPlan p;
  return cluster_plan(a, &p) == 0 ? p.n_clusters * kCl : 0;
}

int launch_decoder_cluster(const ClusterArgs& a, cudaStream_t stream) {
  ClusterPlan p;
  const int prc = cluster_plan(a, &p);
  if (prc < 0) return prc;
  CNB_REQUIRE(prc == 0, "decoder cluster mode: unsupported beam (<= 8) / max_len (<= 64) / T' (<= 128) / vocabulary (<= 65535)");
  if (getenv("CNB_DEC_TRACE"))
    fprintf(stderr, "[dec cluster] batch %d beam %d -> NR %d, %d clips/group, %d groups on %d clusters, vocabulary slice %d\n",
            a.batch, a.beam, p.nr, p.clips_per_group, p.n_groups, p.n_clusters, p.vs);
  cudaLaunchConfig_t lc = {};
  lc.blockDim = dim3(kCThreads);
  lc.gridDim = dim3(p.n_clusters * kCl);
  lc.dynamicSmemBytes = p.smem;
  lc.stream = stream;
  cudaLaunchAttribute la[1];
  la[0].id = cudaLaunchAttributeClusterDimension;
  la[0].val.clusterDim.x = kCl;
  la[0].val.clusterDim.y = 1;
  la[0].val.clusterDim.z = 1;
  lc.attrs = la;
  lc.numAttrs = 1;
  CNB_CUDA_OK(cudaMemsetAsync(a.bs.done, 0, 4 * sizeof(int), stream));
  if (p.nr == 16) CNB_CUDA_OK(cudaLaunchKernelEx(&lc, decoder_cluster_kernel<16>, a, p.clips_per_group, p.n_groups, p.vs));
  else CNB_CUDA_OK(cudaLaunchKernelEx(&lc, decoder_cluster_kernel<32>, a, p.clips_per_group, p.n_groups, p.vs));
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace cnb
