// K-DEC cluster: the whole beam-search decode of a group of clips inside ONE thread-block cluster, one launch per call.
//
// Why: a decode step is ~50 strictly dependent tiny operations on R = clips x beam rows.  As separate kernels (CUDA-graph
// replayed) the 20-step decode is a chain of ~1000 launches at ~9 us each = 9.8 ms for 64 clips -- and 8.3 ms for 8 clips:
// pure latency.  Beam search never mixes clips, so the batch is cut into groups of G clips (G x beam <= 16 rows) and every
// group is decoded start to finish by one cluster of 8 CTAs that never talks to the rest of the grid:
//   * CTA h of the cluster owns attention head h, 1/8 of every projection's output columns, 1/8 of the FF hidden units
//     and 1/8 of the vocabulary; activations (16 x 256 fp32) are replicated in every CTA's shared memory;
//   * GEMMs run on the tensor cores with the WEIGHTS as the M operand: D^T (128 weight rows x 16 decoder rows) +=
//     W_tile (128 x 8, TMA-loaded fp32 straight from the nn.Linear layout, 128B swizzle) x X^T (8 x 16, the activations
//     kept in shared memory in UMMA K-major layout), tcgen05.mma kind::tf32, fp32 accumulators in TMEM.  One thread feeds a
//     4-stage TMA ring that runs ahead across phase boundaries (the weight schedule of a step is static) and issues the
//     MMAs; the epilogue (bias / residual / GELU) reads TMEM with one lane per weight row.  An fp32 CUDA-core version of
//     these GEMMs was measured first: FFMA2 issues at half rate with three register-pair operands, which left the cluster
//     decoder no faster than the launch-bound graph (profiles/r1_decoder_cluster_notes.md);
//   * a phase boundary is a one-sided DSMEM push: every CTA writes its 32-column slice into the 7 peers' buffers with
//     st.async, each 16-byte store completing bytes on the RECEIVER's mbarrier -- no cluster barrier, no fence, no L1
//     flush; 7 exchanges per layer-step (6 per layer + 1 for the distributed beam step).  Buffer reuse is safe because
//     every exchange is all-to-all: a peer can only be one exchange ahead, and consecutive exchanges alternate buffers;
//   * attention: one warp per row; every K row and every V row is requested before the first use (two L2 round trips);
//   * the beam step is distributed: every CTA masks + scans its vocabulary slice (per-row max / sum-exp / top-k by
//     logit), one exchange later every CTA merges the 8 partial results redundantly and deterministically, so the beam
//     state (token histories, KV back-pointers, scores) is replicated in shared memory and needs no further exchange.
// Precision: GEMM operands are truncated to tf32 by the tensor core (fp32 accumulate); everything else is fp32.  This
// mode belongs to precision "fast" (whose encoder GEMMs are bf16); precision "parity" uses the fp32 graph decoder.
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
constexpr int kRm = 16;         // decoder rows per cluster = N of the MMAs
constexpr int kCThreads = 512;
constexpr int kCWarps = kCThreads / 32;
constexpr int kCD = 256, kCFF = 2048, kCLayers = 6, kCHead = 32;
constexpr int kCMaxBeam = 8;
constexpr int kCMaxLen = 64;
constexpr int kCMaxTp = 128;    // encoder frames per clip handled by the in-register cross-attention scores
constexpr int kCPad = 0, kCEos = 2;
constexpr unsigned kFull = 0xffffffffu;
constexpr float kCAttScale = 0.17677669529663687f;  // 1/sqrt(32)
constexpr int kNumEx = 7;       // exchange slots (one mbarrier each)
constexpr int kTrSlots = 20;
constexpr int kStages = 3;      // weight ring: 3 x 32 KB
constexpr int kStageFloats = 8192;
constexpr int kChunksPerLayer = 4 + 3 + 8 + 8;  // QKV | sa_out, ca_q, ca_out | FF1 | FF2
constexpr int kMaxClsTiles = 4;
constexpr int kChains = 4;      // independent TMEM accumulators per tile (k8 step j of every k-chunk goes to chain j)
constexpr int kTmemCols = kMaxClsTiles * kChains * 16;
constexpr int kProducerTid = 32 * 15;  // warp 15 feeds the weight ring and has no epilogue duty (warp 11 covers its quarter)

__device__ __forceinline__ unsigned long long cl_global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct CCand {
  float v;
  int idx;
};
__device__ __forceinline__ bool cbetter(const CCand& a, const CCand& b) {
  return a.v > b.v || (a.v == b.v && a.idx < b.idx);
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
__device__ __forceinline__ void cbar_expect(uint32_t bar, uint32_t bytes) { mbar_expect_tx(bar, bytes); }
__device__ __forceinline__ void cbar_wait(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }

// ---- activation buffers that feed the tensor cores: UMMA K-major, 128B swizzle ---------------------------------------------
// element (row r, k) of a 16 x 256 operand lives at float index xo(r, k): 8 chunks of 32 k, each 16 rows x 128 bytes, the
// 16-byte group index XOR-ed with (r & 7).  Chunk h is exactly the 32 columns owned by head / CTA h (2 KB contiguous).
__device__ __forceinline__ int xo(int r, int k) {
  return ((k >> 5) << 9) + (r << 5) + (((((k >> 2) & 7) ^ (r & 7)) << 2) | (k & 3));
}

struct CSmem {
  float ring[kStages][kStageFloats];  // weight tiles (A operand), 1024-byte aligned
  float xs[kRm * kCD];                // layer input / residual stream, operand layout (replicated in every CTA)
  float ga[kRm * kCD];                // attention outputs of all heads (gathered), operand layout
  union {                             // never live at the same time (hs: FF1 epilogue -> last FF2 MMA)
    float hs[kRm * kCD];              // FF1 hidden slice (local), operand layout
    float gb[kRm][kCD];               // pre-LayerNorm rows (gathered) | FF2 partial sums (local), plain layout
  };
  float recv[kCl][kRm][kCHead];       // FF2 partial sums for this CTA's 32 columns, one slab per peer
  float q[kRm][kCHead];               // q of this CTA's head
  float kv[kRm][2 * kCHead];          // k | v of the current position, this head
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
  uint32_t tmem_slot;
  unsigned long long bars[kNumEx];    // one mbarrier per exchange slot
  unsigned long long full[kStages], empty[kStages], tile_full[kMaxClsTiles];
  unsigned long long tr_acc[kTrSlots];
  unsigned long long tr_last;
  unsigned long long tr2[4];         // debug (thread 0): cycles waiting for weights / issuing MMAs / waiting for MMAs / epilogue
};

// state of the weight pipeline, owned by thread 0 of the CTA (plain registers / local memory)
struct Pipe {
  uint32_t load = 0;  // chunks requested so far (running over all steps; meaningful in the producer thread)
  uint32_t use = 0;   // chunks consumed so far (tracked by every thread)
};

// x = LayerNorm(gb) (eps 1e-5, biased variance), one warp per row; result in operand layout
__device__ __forceinline__ void ln_rows(const float (*gb)[kCD], float* xs, const float* __restrict__ g,
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
      xs[xo(r, c)] = (v[j] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
    }
  }
}

// Attention of NR rows at once by one warp over up to NCH*32 cached keys each (+ optionally one extra key/value held in
// shared memory: the position being decoded).  kptr(rr, j) / vptr(rr, j) give the 32-float head slice of key / value j of
// row rr.  Scores: lane = key (8 x LDG.128 each, all rows / chunks requested before the first use).  Values: lane =
// (key group of 4, 4-dim quad): 16-byte loads, all requested up front, then a 2-step shuffle reduction over the groups.
template <int NR, int NCH, typename KPtr, typename VPtr>
__device__ __forceinline__ void attend(const float* const (&q)[NR], const int (&n)[NR], const bool (&valid)[NR], KPtr kptr,
                                       VPtr vptr, const float* const (&kv_new)[NR], bool has_new, float* const (&out)[NR],
                                       const int (&out_sw)[NR], int lane) {
  float sc[NR][NCH];
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    float qv[kCHead];
#pragma unroll
    for (int d = 0; d < kCHead; d += 4) {
      const float4 t = *reinterpret_cast<const float4*>(q[rr] + d);
      qv[d] = t.x; qv[d + 1] = t.y; qv[d + 2] = t.z; qv[d + 3] = t.w;
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int j = lane + 32 * ch;
      sc[rr][ch] = -INFINITY;
      if (valid[rr] && j < n[rr]) {
        const float4* kr = reinterpret_cast<const float4*>(kptr(rr, j));
        float4 kk[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) kk[d] = kr[d];
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
  float e[NR][NCH], e_new[NR], inv[NR];
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    float s_new = -INFINITY;
    if (has_new) s_new = warp_sum(q[rr][lane] * kv_new[rr][lane]) * kCAttScale;
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
  for (int rr = 0; rr < NR; ++rr) {
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
        vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid[rr] && j < n[rr]) vv[u] = *reinterpret_cast<const float4*>(vptr(rr, j) + 4 * dq);
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
    if (valid[rr] && lane < 8) {
      if (has_new) {
        const float4 vn = *reinterpret_cast<const float4*>(kv_new[rr] + kCHead + 4 * dq);
        acc.x = fmaf(e_new[rr], vn.x, acc.x); acc.y = fmaf(e_new[rr], vn.y, acc.y);
        acc.z = fmaf(e_new[rr], vn.z, acc.z); acc.w = fmaf(e_new[rr], vn.w, acc.w);
      }
      *reinterpret_cast<float4*>(out[rr] + 4 * (dq ^ out_sw[rr])) =
          make_float4(acc.x * inv[rr], acc.y * inv[rr], acc.z * inv[rr], acc.w * inv[rr]);
    }
  }
}

// per-thread-0 cycle counters inside the GEMM phases (tr2): compile with -DCNB_DEC_TR2=1 to collect them; off by default
// because the clock reads and the divergent shared-memory updates sit on the MMA issue path of warp 0
#ifndef CNB_DEC_TR2
#define CNB_DEC_TR2 0
#endif
#define CL_TR(slot)                                \
  if (tr_on) {                                     \
    const unsigned long long n_ = cl_global_ns();  \
    S.tr_acc[slot] += n_ - S.tr_last;              \
    S.tr_last = n_;                                \
  }

// ---- weight pipeline --------------------------------------------------------------------------------------------------
// The chunk stream of one decode step is static (chunk = one 32 KB ring stage):
//   per layer  QKV      4 chunks: two k-chunks of 32, each 3 boxes [32 rows x 32 k] (q / k / v rows of this head)
//              sa_out, ca_q, ca_out   1 chunk each: eight boxes [32 rows x 32 k] at a 4 KB pitch (the whole K = 256)
//              FF1, FF2 8 chunks each: one box [256 rows x 32 k] = both 128-row tiles of one k-chunk
//   classifier 8 chunks per pair of 128-word tiles
// Chunk g (running index) lives in ring stage g % kStages.  Producer = thread 32, MMA issuer = thread 0.
__device__ __forceinline__ void issue_chunk(CSmem& S, const PersistentArgs& a, uint32_t g, int rank, int v0, int chunks_per_step) {
  const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(a.tmaps);
  const int s = (int)(g % kStages);
  const int idx = (int)(g % (uint32_t)chunks_per_step);
  const uint32_t dst = smem_addr(&S.ring[s][0]);
  const uint32_t bar = smem_addr(&S.full[s]);
  if (idx < kCLayers * kChunksPerLayer) {
    const int l = idx / kChunksPerLayer, j = idx - l * kChunksPerLayer;
    if (j < 4) {  // QKV
      mbar_expect_tx(bar, 6 * 4096);
#pragma unroll
      for (int sub = 0; sub < 2; ++sub)
#pragma unroll
        for (int part = 0; part < 3; ++part)
          tma_load_2d(dst + sub * 16384 + part * 4096, maps + 6 * l, (2 * j + sub) * 32, part * kCD + rank * kCHead, bar);
    } else if (j < 7) {  // sa_out / ca_q / ca_out
      mbar_expect_tx(bar, 8 * 4096);
#pragma unroll
      for (int sub = 0; sub < 8; ++sub) tma_load_2d(dst + sub * 4096, maps + 6 * l + 1 + (j - 4), sub * 32, rank * kCHead, bar);
    } else if (j < 15) {  // FF1: hidden units [256 rank, +256)
      mbar_expect_tx(bar, 32768);
      tma_load_2d(dst, maps + 6 * l + 4, (j - 7) * 32, rank * kCD, bar);
    } else {  // FF2: all 256 outputs, this CTA's K slice [256 rank, +256)
      mbar_expect_tx(bar, 32768);
      tma_load_2d(dst, maps + 6 * l + 5, rank * kCD + (j - 15) * 32, 0, bar);
    }
  } else {  // classifier: words [v0 + 256 pair, +256); rows beyond the vocabulary are zero-filled by TMA
    const int jj = idx - kCLayers * kChunksPerLayer;
    mbar_expect_tx(bar, 32768);
    tma_load_2d(dst, maps + 36, (jj & 7) * 32, v0 + (jj >> 3) * 256, bar);
  }
}

// k8 step j of one k-chunk for TPC tiles (accumulator chain j).  Issuing a tcgen05.mma costs the issuing thread ~100 cycles
// (measured), far more than the ~24 tensor-pipe cycles of a 128 x 16 x 8 MMA: four threads issue, one per chain.
template <int TPC>
__device__ __forceinline__ void mma_block(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, int j) {
  constexpr uint32_t idesc = make_idesc_tf32(128, kRm);
#pragma unroll
  for (int tt = 0; tt < TPC; ++tt)
    tcgen05_mma_tf32(tmem_d + (uint32_t)((tt * kChains + j) * kRm), adesc + (uint64_t)(2 * j + tt * 1024), bdesc + 2 * j, idesc,
                     accumulate);
}

// Geometry of a GEMM phase: n_chunks ring stages; a chunk carries `tpc` tiles (16 KB apart) x `ns` k-chunks of 32 (sub_pitch
// bytes apart); 8 / ns chunks complete a group of tpc accumulator tiles.
struct GemmShape {
  int n_chunks, tpc, ns, sub_pitch;
};

// One GEMM phase on the tensor cores; xbuf = B operand (16 x 256, operand layout).  epi(tile, quarter, lane, v[16]) is
// called by the four warps that own the accumulator rows [32 quarter, +32) of that tile: v[r] = sum_k W[row][k] * x[r][k].
// Ends with all threads having finished their TMEM reads (callers __syncthreads() before touching what the epilogue wrote).
template <typename Epi>
__device__ __forceinline__ void tc_gemm(CSmem& S, const PersistentArgs& a, Pipe& pp, uint32_t tmem_base, uint32_t& tile_par,
                                        const GemmShape g, const float* xbuf, int rank, int v0, int chunks_per_step, Epi epi) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  fence_proxy_async_smem();  // activations were written through the generic / st.async path: make them visible to the MMA
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const int cpg = 8 / g.ns;  // chunks per tile group
  const int n_tiles = (g.n_chunks / cpg) * g.tpc;
  // Producer and MMA issuers run their loops on WHOLE warps with warp-uniform values and one elected lane issuing (elect_one,
  // tc_ptx.cuh): under a single-thread branch every TMA / MMA operand went through a vector -> uniform register waterfall
  // loop, ~75-100 cycles per instruction (the decode spent 19 % of its time issuing 24 k MMAs per CTA).
  if (warp == kCWarps - 1) {
    // producer: this phase's chunks, then run ahead into the next phases' weights as far as the ring allows.  Every wait is
    // on MMAs that the issuers issue without depending on this warp beyond the current phase: no deadlock.
    const uint32_t target = pp.use + (uint32_t)g.n_chunks + (kStages - 1);
    while (pp.load < target) {
      const int s = (int)(pp.load % kStages);
      mbar_wait(smem_addr(&S.empty[s]), ((pp.load / kStages) & 1u) ^ 1u);  // the MMAs that read this stage have completed
      if (elect_one()) issue_chunk(S, a, pp.load, rank, v0, chunks_per_step);
      __syncwarp();
      ++pp.load;
    }
  } else if ((warp & 3) == 0) {
    // MMA issuers (warps 0, 4, 8, 12; issuer j owns accumulator chain j): descriptors advance by plain adds.
    const int chain = __shfl_sync(kFull, warp >> 2, 0);
    const uint32_t tb = __shfl_sync(kFull, tmem_base, 0);
    const uint64_t bdesc0 = make_smem_desc(smem_addr(xbuf));
    const uint32_t a_step = (uint32_t)g.sub_pitch >> 4;
    for (int i = 0; i < g.n_chunks; ++i) {
      const uint32_t u = pp.use + (uint32_t)i;
      const int s = (int)(u % kStages);
#if CNB_DEC_TR2
      const unsigned long long tf0 = clock64();
      const bool tr0 = tid == 0;
#endif
      mbar_wait(smem_addr(&S.full[s]), (u / kStages) & 1u);
      tcgen05_fence_after();
#if CNB_DEC_TR2
      const unsigned long long tf1 = clock64();
      if (tr0) S.tr2[0] += tf1 - tf0;
#endif
      const int grp = i / cpg, ic = i - grp * cpg;
      if (elect_one()) {
        uint64_t adesc = make_smem_desc(smem_addr(&S.ring[s][0]));
        uint64_t bdesc = bdesc0 + (uint64_t)(ic * g.ns * 128);  // 2048 bytes per k-chunk of the B operand
        const uint32_t tmem_d = tb + (uint32_t)(grp * g.tpc * kChains * kRm);
        if (g.tpc == 2) {
          for (int sub = 0; sub < g.ns; ++sub, adesc += a_step, bdesc += 128) mma_block<2>(tmem_d, adesc, bdesc, (ic | sub) != 0, chain);
        } else {
          for (int sub = 0; sub < g.ns; ++sub, adesc += a_step, bdesc += 128) mma_block<1>(tmem_d, adesc, bdesc, (ic | sub) != 0, chain);
        }
        tcgen05_commit(smem_addr(&S.empty[s]));  // frees the stage once these MMAs have read it
        if (ic == cpg - 1)
          for (int tt = 0; tt < g.tpc; ++tt) tcgen05_commit(smem_addr(&S.tile_full[grp * g.tpc + tt]));
      }
      __syncwarp();
#if CNB_DEC_TR2
      if (tr0) S.tr2[1] += clock64() - tf1;
#endif
    }
  }
  pp.use += (uint32_t)g.n_chunks;
  __syncwarp();
  // accumulator rows [32 q, +32) are only reachable from warps with id % 4 == q: warps 4t .. 4t+3 read tile t, except that
  // warp 15 (the producer) hands tile 3 / quarter 3 to warp 11
  const int quarter = warp & 3;
  for (int t = warp >> 2; t < n_tiles && warp != 15; t += (warp == 11 ? 1 : kMaxClsTiles)) {
#if CNB_DEC_TR2
    const unsigned long long tw0 = clock64();
#endif
    mbar_wait(smem_addr(&S.tile_full[t]), (tile_par >> t) & 1u);
    tcgen05_fence_after();
#if CNB_DEC_TR2
    if (tid == 0) S.tr2[2] += clock64() - tw0;
#endif
    float v[kRm], w[kRm];
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * kChains * kRm);
    tmem_ld_32x16(taddr, v);
#pragma unroll
    for (int c = 1; c < kChains; ++c) {
      tmem_ld_32x16(taddr + c * kRm, w);
      tmem_ld_wait();
#pragma unroll
      for (int r = 0; r < kRm; ++r) v[r] += w[r];
    }
    epi(t, quarter, lane, v);
#if CNB_DEC_TR2
    if (tid == 0) S.tr2[3] += clock64() - tw0;
#endif
  }
  tile_par ^= (1u << n_tiles) - 1u;  // every thread tracks the phase parity of every accumulator tile
  tcgen05_fence_before();
}

// One decoder layer for the rows of this cluster (all 8 CTAs execute it in lock step through the six exchanges).
// `par` = parity of this call's exchange mbarrier completions (every slot 0..5 completes exactly once per layer).
__device__ __noinline__ void decode_layer(CSmem& S, const PersistentArgs& a, Pipe& pp, uint32_t tmem_base, uint32_t& tile_par,
                                          int l, int rank, int R, int grow0, int clip0, int pos, int cur, uint32_t par,
                                          int v0, int chunks_per_step, bool tr_on) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int beam = a.beam, max_len = a.max_len, tp = a.tp;
  const int64_t cache_l = (int64_t)a.rows * max_len * kCD;
  const int64_t kv_stride = (int64_t)kCLayers * 2 * kCD;
  const PLayer& L = a.layers[l];
  // push this CTA's 2 KB chunk `rank` of an operand-layout buffer (same offset in every CTA) to the 7 peers on exchange
  // slot e, then wait until the 7 peers' chunks have landed here.  Callers __syncthreads() before (the chunk is complete).
  auto exchange_chunk = [&](float* buf, int e) {
    const uint32_t bar = smem_addr(&S.bars[e]);
    if (tid == 0) cbar_expect(bar, (kCl - 1) * kRm * kCHead * 4);
    for (int idx = tid; idx < (kCl - 1) * kRm * 8; idx += kCThreads) {
      const int p = idx / (kRm * 8), q = idx - p * (kRm * 8);
      const int peer = (rank + 1 + p) & (kCl - 1);
      float* src = buf + rank * (kRm * kCHead) + 4 * q;
      st_async16(map_peer(smem_addr(src), peer), *reinterpret_cast<const float4*>(src), map_peer(bar, peer));
    }
    cbar_wait(bar, par);
  };
  // same for columns [32 rank, +32) of the plain-layout gb
  auto exchange_slice = [&](float* buf, int e) {
    const uint32_t bar = smem_addr(&S.bars[e]);
    if (tid == 0) cbar_expect(bar, (kCl - 1) * kRm * kCHead * 4);
    for (int idx = tid; idx < (kCl - 1) * kRm * 8; idx += kCThreads) {
      const int p = idx / (kRm * 8), rem = idx - p * (kRm * 8);
      const int r = rem >> 3, q = rem & 7;
      const int peer = (rank + 1 + p) & (kCl - 1);
      float* src = buf + r * kCD + rank * kCHead + 4 * q;
      st_async16(map_peer(smem_addr(src), peer), *reinterpret_cast<const float4*>(src), map_peer(bar, peer));
    }
    cbar_wait(bar, par);
  };

  // ---- P1: q | k | v of head `rank` (96 weight rows), then self-attention for the rows of this head
  tc_gemm(S, a, pp, tmem_base, tile_par, GemmShape{4, 1, 2, 16384}, S.xs, rank, v0, chunks_per_step,
          [&](int, int quarter, int d, const float (&v)[kRm]) {
            if (quarter >= 3) return;
            const float bias = __ldg(L.sa_in_b + quarter * kCD + rank * kCHead + d);
#pragma unroll
            for (int r = 0; r < kRm; ++r) {
              if (quarter == 0) S.q[r][d] = v[r] + bias;
              else S.kv[r][(quarter - 1) * kCHead + d] = v[r] + bias;
            }
          });
  __syncthreads();
  CL_TR(0);
  {
    float* kc = a.kc + l * cache_l;
    float* vc = a.vc + l * cache_l;
    for (int r = warp; r < R; r += kCWarps) {
      const float* const qq[1] = {&S.q[r][0]};
      const float* const kvn[1] = {&S.kv[r][0]};
      float* const oo[1] = {S.ga + rank * (kRm * kCHead) + r * kCHead};
      const int osw[1] = {r & 7};
      const int nn[1] = {pos};
      const bool valid[1] = {true};
      const int* s0 = &S.src[cur][r][0];
      // the new position goes to this head's slice of the global cache (read back in later steps only)
      const int64_t o = ((int64_t)(grow0 + r) * max_len + pos) * kCD + rank * kCHead + lane;
      kc[o] = kvn[0][lane];
      vc[o] = kvn[0][kCHead + lane];
      auto kp = [&](int, int j) { return kc + ((int64_t)(grow0 + s0[j]) * max_len + j) * kCD + rank * kCHead; };
      auto vp = [&](int, int j) { return vc + ((int64_t)(grow0 + s0[j]) * max_len + j) * kCD + rank * kCHead; };
      if (pos <= 32) attend<1, 1>(qq, nn, valid, kp, vp, kvn, true, oo, osw, lane);
      else attend<1, 2>(qq, nn, valid, kp, vp, kvn, true, oo, osw, lane);
    }
  }
  __syncthreads();
  CL_TR(1);
  exchange_chunk(S.ga, 0);
  CL_TR(2);
  // ---- P2: self-attention output projection (32 columns) + residual, gather, LayerNorm 1
  tc_gemm(S, a, pp, tmem_base, tile_par, GemmShape{1, 1, 8, 4096}, S.ga, rank, v0, chunks_per_step,
          [&](int, int quarter, int j, const float (&v)[kRm]) {
            if (quarter != 0) return;
            const int c = rank * kCHead + j;
            const float bias = __ldg(L.sa_out_b + c);
#pragma unroll
            for (int r = 0; r < kRm; ++r) S.gb[r][c] = S.xs[xo(r, c)] + (v[r] + bias);
          });
  __syncthreads();
  CL_TR(3);
  exchange_slice(&S.gb[0][0], 1);
  CL_TR(4);
  ln_rows(S.gb, S.xs, L.n1_g, L.n1_b, warp, lane);
  CL_TR(5);
  // ---- P3: cross-attention query of head `rank`, cross-attention over the clip's encoder frames
  tc_gemm(S, a, pp, tmem_base, tile_par, GemmShape{1, 1, 8, 4096}, S.xs, rank, v0, chunks_per_step,
          [&](int, int quarter, int j, const float (&v)[kRm]) {
            if (quarter != 0) return;
            const float bias = __ldg(L.ca_q_b + rank * kCHead + j);
#pragma unroll
            for (int r = 0; r < kRm; ++r) S.q[r][j] = v[r] + bias;
          });
  __syncthreads();
  CL_TR(6);
  {
    const float* ck = a.ckv + (int64_t)l * 2 * kCD + rank * kCHead;  // key-padding mask: frames >= len are skipped
    for (int r = warp; r < R; r += kCWarps) {
      const float* const qq[1] = {&S.q[r][0]};
      const float* const kvn[1] = {nullptr};
      float* const oo[1] = {S.ga + rank * (kRm * kCHead) + r * kCHead};
      const int osw[1] = {r & 7};
      const int c0 = clip0 + r / beam;
      const int nn[1] = {min(a.lens[c0], tp)};
      const bool valid[1] = {true};
      auto kp = [&](int, int j) { return ck + ((int64_t)c0 * tp + j) * kv_stride; };
      auto vp = [&](int, int j) { return ck + ((int64_t)c0 * tp + j) * kv_stride + kCD; };
      if (tp <= 32) attend<1, 1>(qq, nn, valid, kp, vp, kvn, false, oo, osw, lane);
      else if (tp <= 64) attend<1, 2>(qq, nn, valid, kp, vp, kvn, false, oo, osw, lane);
      else attend<1, 4>(qq, nn, valid, kp, vp, kvn, false, oo, osw, lane);
    }
  }
  __syncthreads();
  CL_TR(7);
  exchange_chunk(S.ga, 2);
  CL_TR(8);
  // ---- P4: cross-attention output projection + residual, gather, LayerNorm 2
  tc_gemm(S, a, pp, tmem_base, tile_par, GemmShape{1, 1, 8, 4096}, S.ga, rank, v0, chunks_per_step,
          [&](int, int quarter, int j, const float (&v)[kRm]) {
            if (quarter != 0) return;
            const int c = rank * kCHead + j;
            const float bias = __ldg(L.ca_out_b + c);
#pragma unroll
            for (int r = 0; r < kRm; ++r) S.gb[r][c] = S.xs[xo(r, c)] + (v[r] + bias);
          });
  __syncthreads();
  CL_TR(9);
  exchange_slice(&S.gb[0][0], 3);
  CL_TR(10);
  ln_rows(S.gb, S.xs, L.n2_g, L.n2_b, warp, lane);
  // ---- P5: FF1 slice (256 hidden units of this CTA) + GELU -> hs (operand layout)
  tc_gemm(S, a, pp, tmem_base, tile_par, GemmShape{8, 2, 1, 0}, S.xs, rank, v0, chunks_per_step,
          [&](int t, int quarter, int d, const float (&v)[kRm]) {
            const int j = t * 128 + quarter * 32 + d;
            const float bias = __ldg(L.l1_b + rank * kCD + j);
#pragma unroll
            for (int r = 0; r < kRm; ++r) S.hs[xo(r, j)] = gelu_erf(v[r] + bias);
          });
  CL_TR(11);
  // ---- P6: FF2 partial sums over this CTA's K slice for all 256 outputs, reduce-scatter, + bias + residual, gather, LN 3
  tc_gemm(S, a, pp, tmem_base, tile_par, GemmShape{8, 2, 1, 0}, S.hs, rank, v0, chunks_per_step,
          [&](int t, int quarter, int d, const float (&v)[kRm]) {
            const int n = t * 128 + quarter * 32 + d;
#pragma unroll
            for (int r = 0; r < kRm; ++r) S.gb[r][n] = v[r];
          });
  __syncthreads();
  CL_TR(12);
  {
    const uint32_t bar = smem_addr(&S.bars[4]);
    if (tid == 0) cbar_expect(bar, (kCl - 1) * kRm * kCHead * 4);
    for (int idx = tid; idx < kCl * kRm * 8; idx += kCThreads) {
      const int p = idx / (kRm * 8), rem = idx - p * (kRm * 8);
      const int r = rem >> 3, q = rem & 7;
      const int peer = (rank + p) & (kCl - 1);
      const float4 v = *reinterpret_cast<const float4*>(&S.gb[r][peer * kCHead + 4 * q]);
      float* dst = &S.recv[rank][r][4 * q];
      if (peer == rank) *reinterpret_cast<float4*>(dst) = v;
      else st_async16(map_peer(smem_addr(dst), peer), v, map_peer(bar, peer));
    }
    cbar_wait(bar, par);
    __syncthreads();  // own slab was written with ordinary stores; every thread is done reading the partial sums in gb
  }
  CL_TR(13);
  for (int idx = tid; idx < kRm * kCHead; idx += kCThreads) {
    const int r = idx >> 5, c = idx & 31;
    float y = __ldg(L.l2_b + rank * kCHead + c);
#pragma unroll
    for (int i = 0; i < kCl; ++i) y += S.recv[i][r][c];  // fixed order
    S.gb[r][rank * kCHead + c] = S.xs[xo(r, rank * kCHead + c)] + y;
  }
  __syncthreads();
  exchange_slice(&S.gb[0][0], 5);
  ln_rows(S.gb, S.xs, L.n3_g, L.n3_b, warp, lane);
  CL_TR(14);
}

// Classifier slice + distributed beam step (one exchange on slot 6, parity `par`).
__device__ __noinline__ void decode_select(CSmem& S, const PersistentArgs& a, Pipe& pp, uint32_t tmem_base, uint32_t& tile_par,
                                           float* s_logits, int rank, int R, int grow0, int nclips, int step, int cur, int vs,
                                           uint32_t par, int chunks_per_step, bool tr_on) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int beam = a.beam, max_len = a.max_len, V = a.vocab;
  const int v0 = rank * vs;
  const int ncls = min(vs, V - v0) > 0 ? min(vs, V - v0) : 0;
  // ---- classifier slice: logits[r][c] for words v0 + c (tiles of 128 words)
  {
    const int cls_pairs = (vs + 255) / 256;
    tc_gemm(S, a, pp, tmem_base, tile_par, GemmShape{cls_pairs * 8, 2, 1, 0}, S.xs, rank, v0, chunks_per_step,
            [&](int t, int quarter, int d, const float (&v)[kRm]) {
              const int j = t * 128 + quarter * 32 + d;
              if (j < ncls) {
                const float bias = __ldg(a.cls_b + v0 + j);
#pragma unroll
                for (int r = 0; r < kRm; ++r) s_logits[r * vs + j] = v[r] + bias;
              }
            });
  }
  __syncthreads();
  CL_TR(15);
  // ---- beam step, part A (local): masks, per-row max / sum-exp / top-`beam` words of this vocabulary slice
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
    float mx = -INFINITY;
    for (int c = lane; c < ncls; c += 32) mx = fmaxf(mx, lg[c]);
    mx = warp_max(mx);
    float sm = 0.f;
    for (int c = lane; c < ncls; c += 32) sm += expf(lg[c] - mx);
    sm = warp_sum(sm);
    if (lane == 0) {
      S.st_stat[r][0] = mx;
      S.st_stat[r][1] = (mx == -INFINITY) ? 0.f : sm;
    }
    // `beam` rounds of: every lane's best word strictly after the previous winner, then a warp arg-max
    CCand prev{INFINITY, -1};
    for (int k = 0; k < beam; ++k) {
      CCand best{-INFINITY, 0x7fffffff};
      for (int c = lane; c < ncls; c += 32) {
        const CCand cc{lg[c], v0 + c};
        if (cbetter(cc, best) && cbetter(prev, cc)) best = cc;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const CCand other{__shfl_xor_sync(kFull, best.v, o), __shfl_xor_sync(kFull, best.idx, o)};
        if (cbetter(other, best)) best = other;
      }
      if (lane == 0) S.st_cnd[r][k] = best;
      prev = best;
      if (best.idx == 0x7fffffff) prev = CCand{-INFINITY, 0x7ffffffe};  // exhausted (or NaN logits): keep emitting sentinels
    }
  }
  __syncthreads();
  CL_TR(16);
  {
    const uint32_t bar = smem_addr(&S.bars[6]);
    constexpr int kUnits = 1 + kCMaxBeam;  // 8-byte units per row: (max, sum-exp) + kCMaxBeam candidates
    if (tid == 0) cbar_expect(bar, (kCl - 1) * kRm * kUnits * 8);
    for (int idx = tid; idx < kCl * kRm * kUnits; idx += kCThreads) {
      const int p = idx / (kRm * kUnits), rem = idx - p * (kRm * kUnits);
      const int r = rem / kUnits, u = rem - r * kUnits;
      const int peer = (rank + p) & (kCl - 1);
      const uint32_t* srcw = u == 0 ? reinterpret_cast<const uint32_t*>(&S.st_stat[r][0])
                                    : reinterpret_cast<const uint32_t*>(&S.st_cnd[r][u - 1]);
      void* dst = u == 0 ? static_cast<void*>(&S.stat[rank][r][0]) : static_cast<void*>(&S.cnd[rank][r][u - 1]);
      if (peer == rank) {
        reinterpret_cast<uint32_t*>(dst)[0] = srcw[0];
        reinterpret_cast<uint32_t*>(dst)[1] = srcw[1];
      } else {
        st_async8(map_peer(smem_addr(dst), peer), srcw[0], srcw[1], map_peer(bar, peer));
      }
    }
    cbar_wait(bar, par);
    __syncthreads();
  }
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
}

__global__ void __launch_bounds__(kCThreads, 1)
decoder_cluster_kernel(const __grid_constant__ PersistentArgs a, int clips_per_group, int n_groups, int vs /*vocabulary slice*/) {
  extern __shared__ uint8_t smem_raw[];
  // operand tiles need 1024-byte alignment (128B swizzle atoms); the offset is the same in every CTA of the launch
  uint8_t* smem_al = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  CSmem& S = *reinterpret_cast<CSmem*>(smem_al);
  float* s_logits = reinterpret_cast<float*>(smem_al + sizeof(CSmem));  // [kRm][vs]

  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();  // = attention head owned by this CTA
  const int cluster_id = blockIdx.x / kCl, n_clusters = gridDim.x / kCl;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int beam = a.beam, max_len = a.max_len;
  const int v0 = rank * vs;
  const int chunks_per_step = kCLayers * kChunksPerLayer + ((vs + 255) / 256) * 8;
  int steps_max = 0;
  for (int i = tid; i < (int)(sizeof(CSmem) / 4); i += kCThreads) reinterpret_cast<uint32_t*>(smem_al)[i] = 0u;
  __syncthreads();
  if (tid == 0) {
    for (int e = 0; e < kNumEx; ++e) mbar_init(smem_addr(&S.bars[e]), 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_addr(&S.full[s]), 1);
      mbar_init(smem_addr(&S.empty[s]), kChains);  // one tcgen05.commit per MMA issuer
    }
    for (int t = 0; t < kMaxClsTiles; ++t) mbar_init(smem_addr(&S.tile_full[t]), kChains);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async_smem();
  }
  if (warp == 0) {  // TMEM: 4 tiles x 4 chains x 16 columns
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
  if (tr_on) S.tr_last = cl_global_ns();
  cl.sync();  // every CTA's mbarriers are initialised before any peer pushes data at them
  Pipe pp;
  uint32_t tile_par = 0;
  int n_layers = 0, n_steps = 0;  // completed exchange rounds: parity of the exchange mbarrier phases

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
        const float4 e = __ldg(reinterpret_cast<const float4*>(a.emb + (int64_t)tok * kCD) + q);
        const float4 p = __ldg(reinterpret_cast<const float4*>(a.pe) + q);
        o = make_float4(fmaf(e.x, 16.f, p.x), fmaf(e.y, 16.f, p.y), fmaf(e.z, 16.f, p.z), fmaf(e.w, 16.f, p.w));
      }
      *reinterpret_cast<float4*>(&S.xs[xo(r, 4 * q)]) = o;
    }
    __syncthreads();

    int cur = 0, steps_done = max_len;
    for (int step = 0; step < max_len; ++step) {
      const int pos = step;
      for (int l = 0; l < kCLayers; ++l) {
        decode_layer(S, a, pp, tmem_base, tile_par, l, rank, R, grow0, clip0, pos, cur, (uint32_t)(n_layers & 1), v0,
                     chunks_per_step, tr_on);
        ++n_layers;
      }
      decode_select(S, a, pp, tmem_base, tile_par, s_logits, rank, R, grow0, nclips, step, cur, vs, (uint32_t)(n_steps & 1),
                    chunks_per_step, tr_on);
      ++n_steps;
      cur ^= 1;
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
          const float4 e = __ldg(reinterpret_cast<const float4*>(a.emb + (int64_t)tok * kCD) + q);
          const float4 p = __ldg(reinterpret_cast<const float4*>(a.pe + (int64_t)(step + 1) * kCD) + q);
          *reinterpret_cast<float4*>(&S.xs[xo(r, 4 * q)]) =
              make_float4(fmaf(e.x, 16.f, p.x), fmaf(e.y, 16.f, p.y), fmaf(e.z, 16.f, p.z), fmaf(e.w, 16.f, p.w));
        }
      }
      __syncthreads();
    }
    steps_max = max(steps_max, steps_done);
  }
  // drain the weight chunks that were requested ahead but never consumed
  if (warp == kCWarps - 1)
    for (uint32_t u = pp.use; u < pp.load; ++u) mbar_wait(smem_addr(&S.full[u % kStages]), (u / kStages) & 1u);
  if (tr_on)
    for (int i = 0; i < kTrSlots; ++i) a.trace[i] = S.tr_acc[i];
  if (tr_on)
    for (int i = 0; i < 4; ++i) a.trace[kTrSlots + i] = S.tr2[i];
  if (rank == 0 && tid == 0 && steps_max > 0) atomicMax(&a.bs.done[1], steps_max);
  tcgen05_fence_before();
  cl.sync();  // no CTA may exit while a peer can still push into its shared memory
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

size_t cluster_smem(int vs) { return sizeof(CSmem) + (size_t)kRm * vs * sizeof(float) + 1024; }

struct ClusterPlan {
  int clips_per_group = 0, n_groups = 0, vs = 0, n_clusters = 0;
  size_t smem = 0;
};

int max_clusters_for(size_t smem, int* out) {
  static int cached = 0;
  static size_t cached_smem = 0;
  if (cached == 0 || cached_smem != smem) {
    CNB_CUDA_OK(cudaFuncSetAttribute(decoder_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    CNB_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, decoder_cluster_kernel, &lc));
    cached = n;
    cached_smem = smem;
    if (getenv("CNB_DEC_TRACE")) fprintf(stderr, "[dec cluster] max active clusters %d, smem %zu B\n", n, smem);
  }
  *out = cached;
  return 0;
}

// Group size policy: spread the clips over as many co-resident clusters as the device offers (fewest rows per cluster =
// least attention / beam work per step), in one wave when 16 rows per cluster allow it.  0 ok, 1 unsupported, < 0 error.
int cluster_plan(const PersistentArgs& a, ClusterPlan* p) {
  if (a.beam < 1 || a.beam > kCMaxBeam || a.max_len > kCMaxLen || a.tp > kCMaxTp || a.vpad <= 0 || a.tmaps == nullptr) return 1;
  p->vs = a.vpad / kCl;
  if (p->vs > 128 * kMaxClsTiles) return 1;
  p->smem = cluster_smem(p->vs);
  if (p->smem > 227 * 1024) return 1;
  int mc = 0;
  if (int rc = max_clusters_for(p->smem, &mc)) return rc;
  if (mc <= 0) return 1;
  const int gmax = kRm / a.beam;
  int g = (a.batch + mc - 1) / mc;
  if (g > gmax) g = gmax;
  p->clips_per_group = g;
  p->n_groups = (a.batch + g - 1) / g;
  p->n_clusters = p->n_groups < mc ? p->n_groups : mc;
  return 0;
}

}  // namespace

bool decoder_cluster_supported(const PersistentArgs& a) {
  ClusterPlan p;
  return cluster_plan(a, &p) == 0;
}

int launch_decoder_cluster(const PersistentArgs& a, cudaStream_t stream) {
  ClusterPlan p;
  const int prc = cluster_plan(a, &p);
  if (prc < 0) return prc;
  CNB_REQUIRE(prc == 0, "decoder cluster mode: unsupported beam / max_len / vocabulary / T'");
  if (getenv("CNB_DEC_TRACE"))
    fprintf(stderr, "[dec cluster] batch %d beam %d -> %d clips/group, %d groups on %d clusters\n", a.batch, a.beam,
            p.clips_per_group, p.n_groups, p.n_clusters);
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
  CNB_CUDA_OK(cudaLaunchKernelEx(&lc, decoder_cluster_kernel, a, p.clips_per_group, p.n_groups, p.vs));
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace cnb
