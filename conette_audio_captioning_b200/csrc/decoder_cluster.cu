// K-DEC cluster: the whole beam-search decode of a group of clips inside ONE thread-block cluster, one launch per call.
//
// Why: a decode step is ~50 strictly dependent tiny operations on R = clips x beam rows.  As separate kernels (CUDA-graph
// replayed) the 20-step decode is a chain of ~1000 launches at ~9 us each = 9.8 ms for 64 clips -- and 8.3 ms for 8 clips:
// pure latency.  Beam search never mixes clips, so the batch is cut into groups of G = 12 / beam clips (R <= 12 rows) and
// every group is decoded start to finish by one cluster of 8 CTAs that never talks to the rest of the grid:
//   * CTA h of the cluster owns attention head h, 1/8 of every projection's output columns, 1/8 of the FF hidden units
//     and 1/8 of the vocabulary; activations (R x 256 fp32) are replicated in every CTA's shared memory;
//   * a phase boundary is a DSMEM slice broadcast + one hardware cluster barrier (~0.2 us) instead of a kernel launch;
//     7 barriers per layer-step (6 per layer + 1 for the distributed beam step);
//   * weights are never staged: they stream L2 -> registers (each CTA reads its own 1/8 slice, 4.7 MB per step), x is
//     register-stationary (lane = k, 96 registers hold the 12 x 256 panel), 12 dot products per weight row are reduced
//     with a transposing butterfly (18 shuffles);
//   * the beam step is distributed: every CTA masks + scans its vocabulary slice (per-row max / sum-exp / top-k by
//     logit), one barrier later every CTA merges the 8 partial results redundantly and deterministically, so the beam
//     state (token histories, KV back-pointers, scores) is replicated in shared memory and needs no further exchange.
// All arithmetic is fp32 (FFMA2 on packed pairs); results agree with the graph / persistent modes up to fp32 summation
// order (tests guard near-ties by margin).
// Reference semantics: nn/decoders/aac_tfmer.py:100-116 (embedding*16 + PE, post-norm nn.TransformerDecoder, eps 1e-5),
// nn/decoding/beam.py:113-203 and :230-269 (see beam.cu for the fixed-slot formulation this mirrors).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "attention.cuh"
#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace cnb {

namespace {

constexpr int kCl = 8;          // CTAs per cluster = attention heads
constexpr int kRm = 12;         // beam rows per cluster
constexpr int kCThreads = 256;
constexpr int kCWarps = kCThreads / 32;
constexpr int kCD = 256, kCFF = 2048, kCLayers = 6, kCHead = 32;
constexpr int kCMaxBeam = 8;
constexpr int kCMaxLen = 64;
constexpr int kCPad = 0, kCEos = 2;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ unsigned long long cl_global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr int kTrSlots = 20;

struct CCand {
  float v;
  int idx;
};
__device__ __forceinline__ bool cbetter(const CCand& a, const CCand& b) {
  return a.v > b.v || (a.v == b.v && a.idx < b.idx);
}

// ---- skinny GEMM phase: out[r][c] = sum_k x[r][k] * W[n(c)][k], 12 rows, K = 256 per CTA ------------------------------------
// Lane = output column pair (no cross-lane reduction), the K range is split KS ways across threads so that all 256 threads
// work whatever the column count; x comes from shared memory as broadcast LDS.128, weights stream L2 -> registers from the
// k4-packed copy  Wp[(k/4) * N + n][4]  (a thread's two columns x four k's are two adjacent 16-byte loads, a warp reads 1 KB
// contiguous).  FFMA2 pairs run along k: (x[k], x[k+1]) * (w[k], w[k+1]) with no repacking.  24 independent accumulator
// chains per thread; the next block of four k-quads is in flight while the current one is consumed.
__device__ __forceinline__ float4 ldw4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 ldw4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int KS, typename ColMap, typename Epi>
__device__ __forceinline__ void gemm_phase(const float* __restrict__ wp_base, int N, int k4off, const float* xs, int ncols,
                                           ColMap colmap, float* red, int tid, Epi epi) {
  constexpr int K4 = 64 / KS;  // k-quads per thread
  constexpr int U = 4;
  static_assert(K4 % U == 0, "k split");
  const int P = ncols >> 1;
  const int ks = tid / P, pair = tid - ks * P;
  const bool active = tid < P * KS;
  float2 acc[kRm][2];
#pragma unroll
  for (int r = 0; r < kRm; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
  if (active) {
    const float4* wp = reinterpret_cast<const float4*>(wp_base) + (int64_t)(k4off + ks * K4) * N + colmap(2 * pair);
    const float* xk = xs + 4 * ks * K4;
    float4 wa[U], wb[U], na[U], nb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      wa[u] = ldw4(wp + (int64_t)u * N);
      wb[u] = ldw4(wp + (int64_t)u * N + 1);
    }
#pragma unroll 1
    for (int i0 = 0; i0 < K4; i0 += U) {
      if (i0 + U < K4) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          na[u] = ldw4(wp + (int64_t)(i0 + U + u) * N);
          nb[u] = ldw4(wp + (int64_t)(i0 + U + u) * N + 1);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float2 a01 = make_float2(wa[u].x, wa[u].y), a23 = make_float2(wa[u].z, wa[u].w);
        const float2 b01 = make_float2(wb[u].x, wb[u].y), b23 = make_float2(wb[u].z, wb[u].w);
#pragma unroll
        for (int r = 0; r < kRm; ++r) {
          const float4 xv = *reinterpret_cast<const float4*>(xk + r * kCD + 4 * (i0 + u));
          const float2 x01 = make_float2(xv.x, xv.y), x23 = make_float2(xv.z, xv.w);
          acc[r][0] = __ffma2_rn(x01, a01, acc[r][0]);
          acc[r][1] = __ffma2_rn(x01, b01, acc[r][1]);
          acc[r][0] = __ffma2_rn(x23, a23, acc[r][0]);
          acc[r][1] = __ffma2_rn(x23, b23, acc[r][1]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        wa[u] = na[u];
        wb[u] = nb[u];
      }
    }
  }
  if (KS == 1) {
    if (active) {
#pragma unroll
      for (int r = 0; r < kRm; ++r) {
        epi(r, 2 * pair, acc[r][0].x + acc[r][0].y);
        epi(r, 2 * pair + 1, acc[r][1].x + acc[r][1].y);
      }
    }
  } else {
    if (active) {
#pragma unroll
      for (int r = 0; r < kRm; ++r)
        *reinterpret_cast<float2*>(red + (ks * kRm + r) * ncols + 2 * pair) =
            make_float2(acc[r][0].x + acc[r][0].y, acc[r][1].x + acc[r][1].y);
    }
    __syncthreads();
    for (int idx = tid; idx < kRm * ncols; idx += kCThreads) {
      const int r = idx / ncols, c = idx - r * ncols;
      float v = red[r * ncols + c];
#pragma unroll
      for (int q = 1; q < KS; ++q) v += red[(q * kRm + r) * ncols + c];  // fixed order
      epi(r, c, v);
    }
  }
}

// ---- shared-memory carve-up -------------------------------------------------------------------------------------------
struct CSmem {
  float xs[kRm][kCD];                 // layer input / residual stream (replicated in every CTA)
  float ga[kRm][kCD];                 // attention outputs of all heads (gathered) | FF2 partial sums (local)
  float gb[kRm][kCD];                 // pre-LayerNorm rows (gathered)
  float qb[kRm][kCD];                 // q of this CTA's head (only columns [32h, 32h+32) are used)
  float hs[kRm][kCD];                 // FF1 hidden slice (local)
  float recv[kCl][kRm][kCHead];       // FF2 partial sums for this CTA's 32 columns, one slab per peer
  float kv[kRm][2 * kCHead];          // k | v of the current position, this head
  float red[16 * kRm * kCHead];       // split-K partial sums of a GEMM phase (max: 16 x 12 x 32 = 2 x 12 x 256)
  float stat[kCl][kRm][2];            // per peer: max / sum-exp of its vocabulary slice
  CCand cnd[kCl][kRm][kCMaxBeam];     // per peer: its best words per row (by logit)
  float st_stat[kRm][2];              // local staging of the two above
  CCand st_cnd[kRm][kCMaxBeam];
  CCand win[kCWarps][kCMaxBeam];
  int tokens[2][kRm][kCMaxLen + 1];
  int src[2][kRm][kCMaxLen];          // local row holding position p of this row's history (beam back-pointers)
  float sum_lp[kRm];
  int live[kRm];
  int any_live;
  unsigned long long tr_acc[kTrSlots];
  unsigned long long tr_last;
};

__device__ __forceinline__ void bcast_slice(cg::cluster_group& cl, float* buf, int col0, int rank, int tid) {
  // buf is a [kRm][256] array at the same offset in every CTA: copy columns [col0, col0+32) of all rows to the 7 peers
  for (int idx = tid; idx < (kCl - 1) * kRm * 8; idx += kCThreads) {
    const int p = idx / (kRm * 8), rem = idx - p * (kRm * 8);
    const int r = rem >> 3, q = rem & 7;
    const int peer = (rank + 1 + p) & (kCl - 1);
    float* src = buf + r * kCD + col0 + 4 * q;
    const float4 v = *reinterpret_cast<const float4*>(src);
    *reinterpret_cast<float4*>(cl.map_shared_rank(src, peer)) = v;
  }
}

// x = LayerNorm(gb) (eps 1e-5, biased variance), one warp per row
__device__ __forceinline__ void ln_rows(const float (*gb)[kCD], float (*xs)[kCD], const float* __restrict__ g,
                                        const float* __restrict__ b, int warp, int lane) {
  for (int r = warp; r < kRm; r += kCWarps) {
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = gb[r][lane + 32 * j];
      s += v[j];
    }
    const float mean = warp_sum(s) * (1.f / kCD);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) q += (v[j] - mean) * (v[j] - mean);
    const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / kCD) + 1e-5f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      xs[r][c] = (v[j] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
    }
  }
}

// self-attention of local row r, head h at position pos; q/k/v of the new position come from shared memory, older K/V from
// this head's slice of the global cache (written by this CTA in earlier steps).  Same arithmetic as attention.cuh.
__device__ __forceinline__ void self_attn_local(const float* q, const float* kvn, float* kcache, float* vcache, const int* src,
                                                int grow0, int r, int pos, int max_len, int h, float* out, int lane) {
  const int col = h * kCHead + lane;
  const float q_d = q[lane], k_d = kvn[lane], v_d = kvn[kCHead + lane];
  float qv[kCHead];
#pragma unroll
  for (int d = 0; d < kCHead; d += 4) {
    const float4 t = *reinterpret_cast<const float4*>(q + d);
    qv[d] = t.x; qv[d + 1] = t.y; qv[d + 2] = t.z; qv[d + 3] = t.w;
  }
  kcache[((int64_t)(grow0 + r) * max_len + pos) * kCD + col] = k_d;
  vcache[((int64_t)(grow0 + r) * max_len + pos) * kCD + col] = v_d;
  float sc[2];
  int pr[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int p = lane + 32 * i;
    sc[i] = -INFINITY;
    pr[i] = grow0 + r;
    if (p < pos) {
      pr[i] = grow0 + src[p];
      const float* kr = kcache + ((int64_t)pr[i] * max_len + p) * kCD + h * kCHead;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < kCHead; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(kr + d);
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
  float acc = 0.f;
  for (int p0 = 0; p0 < pos; p0 += 8) {
    float vv[8], ww[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = p0 + u;
      const int srow = __shfl_sync(kFull, (p >> 5) ? pr[1] : pr[0], p & 31);
      ww[u] = __shfl_sync(kFull, (p >> 5) ? e1 : e0, p & 31);
      vv[u] = (p < pos) ? vcache[((int64_t)srow * max_len + p) * kCD + col] : 0.f;
      if (p >= pos) ww[u] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc = fmaf(ww[u], vv[u], acc);
  }
  const float w_new = __shfl_sync(kFull, (pos >> 5) ? e1 : e0, pos & 31);
  acc = fmaf(w_new, v_d, acc);
  out[lane] = acc * inv;
}

__global__ void __launch_bounds__(kCThreads, 1)
decoder_cluster_kernel(const PersistentArgs a, int clips_per_group, int n_groups, int vs /*vocabulary slice width*/) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  CSmem& S = *reinterpret_cast<CSmem*>(smem_raw);
  float* s_logits = reinterpret_cast<float*>(smem_raw + sizeof(CSmem));  // [kRm][vs]
  float* s_sc = s_logits + kRm * vs;                                     // [kCWarps][tp] cross-attention scratch

  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();  // = attention head owned by this CTA
  const int cluster_id = blockIdx.x / kCl, n_clusters = gridDim.x / kCl;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int beam = a.beam, max_len = a.max_len, V = a.vocab, tp = a.tp;
  const int v0 = rank * vs;
  const int ncls = min(vs, V - v0) > 0 ? min(vs, V - v0) : 0;
  const int64_t cache_l = (int64_t)a.rows * max_len * kCD;
  const int64_t kv_stride = (int64_t)kCLayers * 2 * kCD;
  int steps_max = 0;
  for (int i = tid; i < (int)(sizeof(CSmem) / 4); i += kCThreads) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
  __syncthreads();
  // debug trace (CNB_DEC_TRACE): thread 0 of the first CTA accumulates the time between phase marks
  const bool tr_on = a.trace != nullptr && blockIdx.x == 0 && tid == 0;
  if (tr_on) S.tr_last = cl_global_ns();
#define CL_TR(slot)                                \
  if (tr_on) {                                     \
    const unsigned long long n_ = cl_global_ns();  \
    S.tr_acc[slot] += n_ - S.tr_last;              \
    S.tr_last = n_;                                \
  }

  for (int g = cluster_id; g < n_groups; g += n_clusters) {
    const int clip0 = g * clips_per_group;
    const int nclips = min(clips_per_group, a.batch - clip0);
    const int R = nclips * beam;       // live local rows (<= kRm)
    const int grow0 = clip0 * beam;    // first global row of the group

    // ---- init: beam state (replicated), outputs (rank 0), first embedding
    for (int i = tid; i < 2 * kRm * (kCMaxLen + 1); i += kCThreads) (&S.tokens[0][0][0])[i] = kCPad;
    for (int i = tid; i < 2 * kRm * kCMaxLen; i += kCThreads) (&S.src[0][0][0])[i] = (i / kCMaxLen) % kRm;
    if (tid < kRm) {
      S.sum_lp[tid] = 0.f;
      S.live[tid] = tid < R ? 1 : 0;
    }
    __syncthreads();
    if (tid < R) S.tokens[0][tid][0] = (int)a.bos_ids[clip0 + tid / beam];
    if (rank == 0) {
      for (int i = tid; i < R * max_len; i += kCThreads) a.bs.out_preds[(int64_t)grow0 * max_len + i] = kCPad;
      if (tid < R) a.bs.out_lp[grow0 + tid] = 0.f;
    }
    __syncthreads();
    for (int i = tid; i < kRm * (kCD / 4); i += kCThreads) {
      const int r = i / (kCD / 4), q = i % (kCD / 4);
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < R) {
        const int tok = S.tokens[0][r][0];
        const float4 e = ldw4(a.emb + (int64_t)tok * kCD + 4 * q);
        const float4 p = ldw4(a.pe + 4 * q);
        o = make_float4(fmaf(e.x, 16.f, p.x), fmaf(e.y, 16.f, p.y), fmaf(e.z, 16.f, p.z), fmaf(e.w, 16.f, p.w));
      }
      *reinterpret_cast<float4*>(&S.xs[r][4 * q]) = o;
    }
    // peers may still be reading this CTA's gather buffers of the previous group: one barrier separates the groups
    cl.sync();

    int cur = 0, steps_done = max_len;
    for (int step = 0; step < max_len; ++step) {
      const int pos = step;
      for (int l = 0; l < kCLayers; ++l) {
        const PLayer& L = a.layers[l];
        // ---- P1: q | k | v of head `rank` (96 columns), then self-attention for the 12 rows of this head
        gemm_phase<4>(L.sa_in_p, 3 * kCD, 0, &S.xs[0][0], 96,
                      [&](int c) { return (c >> 5) * kCD + rank * kCHead + (c & 31); }, S.red, tid,
                      [&](int r, int c, float acc) {
                        const int part = c >> 5, d = c & 31;
                        const float v = acc + __ldg(L.sa_in_b + part * kCD + rank * kCHead + d);
                        if (part == 0) S.qb[r][rank * kCHead + d] = v;
                        else S.kv[r][(part - 1) * kCHead + d] = v;
                      });
        __syncthreads();
        CL_TR(0);
        for (int r = warp; r < R; r += kCWarps)
          self_attn_local(&S.qb[r][rank * kCHead], &S.kv[r][0], a.kc + l * cache_l, a.vc + l * cache_l, &S.src[cur][r][0], grow0,
                          r, pos, max_len, rank, &S.ga[r][rank * kCHead], lane);
        __syncthreads();
        CL_TR(1);
        bcast_slice(cl, &S.ga[0][0], rank * kCHead, rank, tid);
        cl.sync();  // #1
        CL_TR(2);
        // ---- P2: self-attention output projection (32 columns) + residual, gather, LayerNorm 1
        gemm_phase<16>(L.sa_out_p, kCD, 0, &S.ga[0][0], kCHead, [&](int c) { return rank * kCHead + c; }, S.red, tid,
                       [&](int r, int j, float acc) {
                         const int c = rank * kCHead + j;
                         S.gb[r][c] = S.xs[r][c] + (acc + __ldg(L.sa_out_b + c));
                       });
        __syncthreads();
        CL_TR(3);
        bcast_slice(cl, &S.gb[0][0], rank * kCHead, rank, tid);
        cl.sync();  // #2
        CL_TR(4);
        ln_rows(S.gb, S.xs, L.n1_g, L.n1_b, warp, lane);
        __syncthreads();
        CL_TR(5);
        // ---- P3: cross-attention query of head `rank`, cross-attention over the clip's encoder frames
        gemm_phase<16>(L.ca_q_p, kCD, 0, &S.xs[0][0], kCHead, [&](int c) { return rank * kCHead + c; }, S.red, tid,
                       [&](int r, int j, float acc) {
                         S.qb[r][rank * kCHead + j] = acc + __ldg(L.ca_q_b + rank * kCHead + j);
                       });
        __syncthreads();
        CL_TR(6);
        for (int r = warp; r < R; r += kCWarps) {
          const int clip = clip0 + r / beam;
          cross_attention_task<false>(s_sc + warp * tp, &S.qb[0][0], a.ckv + (int64_t)l * 2 * kCD,
                                      a.ckv + (int64_t)l * 2 * kCD + kCD, kv_stride, a.lens[clip], clip, tp, &S.ga[0][0], r, rank,
                                      lane);
        }
        __syncthreads();
        CL_TR(7);
        bcast_slice(cl, &S.ga[0][0], rank * kCHead, rank, tid);
        cl.sync();  // #3
        CL_TR(8);
        // ---- P4: cross-attention output projection + residual, gather, LayerNorm 2
        gemm_phase<16>(L.ca_out_p, kCD, 0, &S.ga[0][0], kCHead, [&](int c) { return rank * kCHead + c; }, S.red, tid,
                       [&](int r, int j, float acc) {
                         const int c = rank * kCHead + j;
                         S.gb[r][c] = S.xs[r][c] + (acc + __ldg(L.ca_out_b + c));
                       });
        __syncthreads();
        CL_TR(9);
        bcast_slice(cl, &S.gb[0][0], rank * kCHead, rank, tid);
        cl.sync();  // #4
        CL_TR(10);
        ln_rows(S.gb, S.xs, L.n2_g, L.n2_b, warp, lane);
        __syncthreads();
        // ---- P5: FF1 slice (256 hidden units of this CTA) + GELU
        gemm_phase<2>(L.l1_p, kCFF, 0, &S.xs[0][0], kCD, [&](int c) { return rank * kCD + c; }, S.red, tid,
                      [&](int r, int j, float acc) { S.hs[r][j] = gelu_erf(acc + __ldg(L.l1_b + rank * kCD + j)); });
        __syncthreads();
        CL_TR(11);
        // ---- P6: FF2 partial sums over this CTA's K slice for all 256 outputs, reduce-scatter, + bias + residual, gather, LN 3
        gemm_phase<2>(L.l2_p, kCD, rank * (kCD / 4), &S.hs[0][0], kCD, [&](int c) { return c; }, S.red, tid,
                      [&](int r, int j, float acc) { S.ga[r][j] = acc; });
        __syncthreads();
        CL_TR(12);
        for (int idx = tid; idx < kCl * kRm * 8; idx += kCThreads) {
          const int p = idx / (kRm * 8), rem = idx - p * (kRm * 8);
          const int r = rem >> 3, q = rem & 7;
          const int peer = (rank + p) & (kCl - 1);
          const float4 v = *reinterpret_cast<const float4*>(&S.ga[r][peer * kCHead + 4 * q]);
          *reinterpret_cast<float4*>(cl.map_shared_rank(&S.recv[rank][r][4 * q], peer)) = v;
        }
        cl.sync();  // #5
        CL_TR(13);
        for (int idx = tid; idx < kRm * kCHead; idx += kCThreads) {
          const int r = idx >> 5, c = idx & 31;
          float y = __ldg(L.l2_b + rank * kCHead + c);
#pragma unroll
          for (int i = 0; i < kCl; ++i) y += S.recv[i][r][c];  // fixed order
          S.gb[r][rank * kCHead + c] = S.xs[r][rank * kCHead + c] + y;
        }
        __syncthreads();
        bcast_slice(cl, &S.gb[0][0], rank * kCHead, rank, tid);
        cl.sync();  // #6
        ln_rows(S.gb, S.xs, L.n3_g, L.n3_b, warp, lane);
        __syncthreads();
        CL_TR(14);
      }

      // ---- classifier slice: logits[r][c] for words v0 + c
      {
        // the packed classifier has vpad >= 8 * vs columns (zero beyond V): every slice is a whole number of column pairs
        gemm_phase<1>(a.cls_p, a.vpad, 0, &S.xs[0][0], vs, [&](int c) { return v0 + c; }, S.red, tid,
                      [&](int r, int j, float acc) {
                        if (j < ncls) s_logits[r * vs + j] = acc + __ldg(a.cls_b + v0 + j);
                      });
      }
      __syncthreads();
      CL_TR(15);
      // ---- beam step, part A (local): masks, per-row max / sum-exp / top-k of this vocabulary slice
      for (int r = warp; r < R; r += kCWarps) {
        float* lg = s_logits + r * vs;
        if (lane == 0 && step < a.min_len && kCEos >= v0 && kCEos < v0 + ncls) lg[kCEos - v0] = -INFINITY;  // beam.py:129-130
        if (a.forbid != nullptr) {                                                                           // beam.py:146-156
          for (int p = lane; p <= step; p += 32) {
            const int tok = S.tokens[cur][r][p];
            if (a.forbid[tok] && tok >= v0 && tok < v0 + ncls) lg[tok - v0] = -INFINITY;
          }
        }
        __syncwarp();
        CCand loc[kCMaxBeam];
#pragma unroll
        for (int i = 0; i < kCMaxBeam; ++i) loc[i] = CCand{-INFINITY, 0x7fffffff};
        float mx = -INFINITY;
        for (int c = lane; c < ncls; c += 32) {
          const float t = lg[c];
          mx = fmaxf(mx, t);
          const CCand cc{t, v0 + c};
          if (cbetter(cc, loc[kCMaxBeam - 1])) {
            loc[kCMaxBeam - 1] = cc;
#pragma unroll
            for (int i = kCMaxBeam - 1; i > 0; --i)
              if (cbetter(loc[i], loc[i - 1])) {
                const CCand tt = loc[i];
                loc[i] = loc[i - 1];
                loc[i - 1] = tt;
              }
          }
        }
        mx = warp_max(mx);
        float sm = 0.f;
        for (int c = lane; c < ncls; c += 32) sm += expf(lg[c] - mx);
        sm = warp_sum(sm);
        if (lane == 0) {
          S.st_stat[r][0] = mx;
          S.st_stat[r][1] = (mx == -INFINITY) ? 0.f : sm;
        }
        for (int k = 0; k < beam; ++k) {
          CCand best = loc[0];
          int owner = lane;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const CCand other{__shfl_xor_sync(kFull, best.v, o), __shfl_xor_sync(kFull, best.idx, o)};
            const int oo = __shfl_xor_sync(kFull, owner, o);
            if (cbetter(other, best)) {
              best = other;
              owner = oo;
            }
          }
          if (lane == 0) S.st_cnd[r][k] = best;
          if (lane == owner) {
#pragma unroll
            for (int i = 0; i < kCMaxBeam - 1; ++i) loc[i] = loc[i + 1];
            loc[kCMaxBeam - 1] = CCand{-INFINITY, 0x7fffffff};
          }
        }
      }
      __syncthreads();
      CL_TR(16);
      for (int idx = tid; idx < kCl * kRm * (2 + 2 * kCMaxBeam); idx += kCThreads) {
        const int peer = idx / (kRm * (2 + 2 * kCMaxBeam)), rem = idx % (kRm * (2 + 2 * kCMaxBeam));
        const int r = rem / (2 + 2 * kCMaxBeam), w = rem % (2 + 2 * kCMaxBeam);
        if (w < 2) {
          *cl.map_shared_rank(&S.stat[rank][r][w], peer) = S.st_stat[r][w];
        } else {
          const int* srcw = reinterpret_cast<const int*>(&S.st_cnd[r][0]) + (w - 2);
          int* dstw = reinterpret_cast<int*>(&S.cnd[rank][r][0]) + (w - 2);
          *cl.map_shared_rank(dstw, peer) = *srcw;
        }
      }
      cl.sync();  // #7
      CL_TR(17);
      // ---- beam step, part B (replicated): merge, flat top-k per clip, history / back-pointer update, finish bookkeeping
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
            float m = S.stat[0][r][0];
#pragma unroll
            for (int i = 1; i < kCl; ++i) m = fmaxf(m, S.stat[i][r][0]);
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < kCl; ++i) s += S.stat[i][r][1] * expf(S.stat[i][r][0] - m);
            row_mx[j] = m;
            row_lg[j] = logf(s);
          }
        }
        // candidates: (used row j, peer i, k) -> value; k_sel rounds of "best candidate strictly after the previous winner"
        const int n_c = nrows_used * kCl * beam;
        CCand prev_win{INFINITY, -1};
        for (int rsel = 0; rsel < k_sel; ++rsel) {
          CCand best{-INFINITY, 0x7fffffff};
          for (int ci = lane; ci < n_c; ci += 32) {
            const int j = ci / (kCl * beam), rem = ci - j * (kCl * beam);
            const int i = rem / beam, k = rem - i * beam;
            const int r = r0 + label_at(j);
            const CCand raw = S.cnd[i][r][k];
            if (raw.idx == 0x7fffffff) continue;
            float mxj = 0.f, lgj = 0.f, pv = 0.f;
#pragma unroll
            for (int q = 0; q < kCMaxBeam; ++q)
              if (q == j) {
                mxj = row_mx[q];
                lgj = row_lg[q];
                pv = prev_sum[q];
              }
            const float lsm = (raw.v - mxj) - lgj;
            const CCand c{step == 0 ? lsm : pv + lsm, j * V + raw.idx};
            if (cbetter(c, best) && cbetter(prev_win, c)) best = c;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const CCand other{__shfl_xor_sync(kFull, best.v, o), __shfl_xor_sync(kFull, best.idx, o)};
            if (cbetter(other, best)) best = other;
          }
          prev_win = best;
          if (best.idx == 0x7fffffff) {  // NaN logits (fully masked clip, len 0): stay memory-safe like a no-op pick
            best.idx = 0;
            prev_win = CCand{-INFINITY, 0x7ffffffe};
          }
          if (lane == 0) S.win[warp][rsel] = best;
        }
        __syncwarp();
        // candidate r -> r-th live label (beam.py:165-176); histories via back-pointers
        for (int item = lane; item < k_sel * (step + 2); item += 32) {
          const int r = item / (step + 2), p = item - r * (step + 2);
          const int row = r0 + label_at(r);
          const int prev_pos = S.win[warp][r].idx / V;
          const int word = S.win[warp][r].idx - prev_pos * V;
          const int srow = r0 + label_at(prev_pos);
          if (p <= step) {
            S.tokens[nxt][row][p] = S.tokens[cur][srow][p];
            if (p < max_len) S.src[nxt][row][p] = S.src[cur][srow][p];
          } else {
            S.tokens[nxt][row][p] = word;
            if (p < max_len) S.src[nxt][row][p] = row;
          }
        }
        __syncwarp();
        if (lane < k_sel) {
          const int r = lane;
          const int row = r0 + label_at(r);
          const CCand w = S.win[warp][r];
          const int prev_pos = w.idx / V;
          const int word = w.idx - prev_pos * V;
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
      CL_TR(18);
      cur = nxt;
      // ---- continue?  (state is replicated, so every CTA of the cluster takes the same branch)
      if (tid == 0) {
        int any = 0;
        for (int r = 0; r < R; ++r) any |= S.live[r];
        S.any_live = any;
      }
      __syncthreads();
      if (!S.any_live) {
        steps_done = step + 1;
        break;
      }
      // ---- next embedding: x[r] = emb[token at position step+1] * 16 + PE[step+1]
      if (step + 1 < max_len) {
        for (int i = tid; i < R * (kCD / 4); i += kCThreads) {
          const int r = i / (kCD / 4), q = i % (kCD / 4);
          const int tok = S.tokens[cur][r][step + 1];
          const float4 e = ldw4(a.emb + (int64_t)tok * kCD + 4 * q);
          const float4 p = ldw4(a.pe + (int64_t)(step + 1) * kCD + 4 * q);
          *reinterpret_cast<float4*>(&S.xs[r][4 * q]) =
              make_float4(fmaf(e.x, 16.f, p.x), fmaf(e.y, 16.f, p.y), fmaf(e.z, 16.f, p.z), fmaf(e.w, 16.f, p.w));
        }
      }
      __syncthreads();
    }
    steps_max = max(steps_max, steps_done);
  }
  if (tr_on)
    for (int i = 0; i < kTrSlots; ++i) a.trace[i] = S.tr_acc[i];
  if (rank == 0 && tid == 0 && steps_max > 0) atomicMax(&a.bs.done[1], steps_max);
  cl.sync();  // no CTA may exit while a peer can still write into its shared memory
}

}  // namespace

// rows per cluster / vocabulary slice / shared-memory need; returns false when this mode does not apply to the shape
static bool cluster_plan(const PersistentArgs& a, int* clips_per_group, int* n_groups, int* vs, size_t* smem) {
  if (a.beam < 1 || a.beam > kCMaxBeam || a.max_len > kCMaxLen) return false;
  *clips_per_group = kRm / a.beam;
  *n_groups = (a.batch + *clips_per_group - 1) / *clips_per_group;
  *vs = a.vpad / kCl;  // = round_up(ceil(V / 8), 4), fixed when the classifier was packed
  *smem = sizeof(CSmem) + (size_t)kRm * *vs * sizeof(float) + (size_t)kCWarps * a.tp * sizeof(float);
  return *smem <= 220 * 1024;
}

bool decoder_cluster_supported(const PersistentArgs& a) {
  int cpg, ng, vs;
  size_t smem;
  return cluster_plan(a, &cpg, &ng, &vs, &smem);
}

int launch_decoder_cluster(const PersistentArgs& a, cudaStream_t stream) {
  int cpg, ng, vs;
  size_t smem;
  CNB_REQUIRE(cluster_plan(a, &cpg, &ng, &vs, &smem), "decoder cluster mode: unsupported beam / max_len / vocabulary / T'");
  static size_t smem_set = 0;
  if (smem > smem_set) {
    CNB_CUDA_OK(cudaFuncSetAttribute(decoder_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  cudaLaunchConfig_t lc = {};
  lc.blockDim = dim3(kCThreads);
  lc.dynamicSmemBytes = smem;
  lc.stream = stream;
  cudaLaunchAttribute la[1];
  la[0].id = cudaLaunchAttributeClusterDimension;
  la[0].val.clusterDim.x = kCl;
  la[0].val.clusterDim.y = 1;
  la[0].val.clusterDim.z = 1;
  lc.attrs = la;
  lc.numAttrs = 1;
  static int max_clusters = 0;
  static size_t max_clusters_smem = 0;
  if (max_clusters == 0 || max_clusters_smem != smem) {
    lc.gridDim = dim3(kCl * 64);
    int n = 0;
    CNB_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, decoder_cluster_kernel, &lc));
    CNB_REQUIRE(n > 0, "decoder cluster mode: no cluster of 8 CTAs fits on this device");
    max_clusters = n;
    if (getenv("CNB_DEC_TRACE")) fprintf(stderr, "[dec cluster] max active clusters %d, smem %zu B, groups %d\n", n, smem, ng);
    max_clusters_smem = smem;
  }
  const int n_clusters = ng < max_clusters ? ng : max_clusters;
  lc.gridDim = dim3(n_clusters * kCl);
  CNB_CUDA_OK(cudaMemsetAsync(a.bs.done, 0, 4 * sizeof(int), stream));
  CNB_CUDA_OK(cudaLaunchKernelEx(&lc, decoder_cluster_kernel, a, cpg, ng, vs));
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace cnb
