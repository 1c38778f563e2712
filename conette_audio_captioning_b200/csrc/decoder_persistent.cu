// K-DEC persistent: the whole beam-search decode (all steps x 6 layers + classifier + beam selection) in ONE cooperative
// kernel launch, 1 CTA per SM, phases separated by a software grid barrier.
//
// Why: a decode step is ~70 tiny, strictly dependent operations on R = clips x beam (~192) rows.  As separate kernels
// (even replayed from a CUDA graph) each pays ~3 us of launch/drain latency on B200 and the 20-step decode costs ~12 ms;
// here a phase boundary is one atomic + one L2 poll (~0.5 us) and every operand stays in L2.
//
// Phases per layer (B = grid barrier):
//   QKV GEMM [LayerNorm-on-load of the previous layer's FF2 partial sums]  B  self-attention (KV cache, beam back-pointers)
//   B  SA-out GEMM  B  LN1-on-load + cross-q GEMM  B  cross-attention  B  CA-out GEMM  B  LN2-on-load + FF1 (GELU)  B
//   FF2 split-K (8 x 256)  B   ... then LN3-on-load + classifier GEMM  B  beam step (1 CTA per clip) + next embedding  B
// GEMMs are fp32 CUDA-core (bit-faithful greedy ids): 256-wide K panels staged by cp.async.cg (L2-coherent),
// interleaved TM x TN micro-tiles.  LayerNorm-on-load: every tile normalises its own rows straight into the shared-memory
// A panel (x is double-buffered in global so that tiles of the same rows never read what a sibling writes).
// Reference semantics: nn/decoders/aac_tfmer.py:100-116, nn/decoding/beam.py:113-203 (see decoder.cu / beam.cu).
#include <cooperative_groups.h>

#include <utility>

#include "attention.cuh"
#include "common.cuh"
#include "kernels.h"

namespace cnb {

constexpr int kPD = 256, kPHeads = 8, kPHeadDim = 32, kPFF = 2048, kPLayers = 6;
constexpr int kPThreads = 256;
constexpr int kPanelLds = 256 + 4;
constexpr int kPMaxBeam = 8;
constexpr int kPSplits = 8;  // FF2 K-slices

// ---- software grid barrier (monotonic arrival counter; bar[] zeroed by the host before launch) ----------------------
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// trace (debug, optional): block 0 records (phase id, time before the barrier, time after it) for every barrier
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& gen, unsigned long long* trace = nullptr,
                                             int phase = 0) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (trace && blockIdx.x == 0) {
      trace[3 * gen] = (unsigned long long)phase;
      trace[3 * gen + 1] = global_ns();
    }
    gen += 1;
    __threadfence();  // publish this CTA's writes (gpu scope; also invalidates this SM's L1 so later loads see peers' data)
    const unsigned int arrived = atomicAdd(&bar[0], 1u);
    if (arrived == gridDim.x * gen - 1u) {
      atomicExch(&bar[1], gen);
    } else {
      unsigned int spins = 0;
      while (*reinterpret_cast<volatile unsigned int*>(&bar[1]) < gen) {
        if (++spins > (1u << 26)) {
          printf("conette_b200: grid barrier timed out (block %d gen %u)\n", blockIdx.x, gen);
          __trap();
        }
      }
    }
    __threadfence();
    if (trace && blockIdx.x == 0) trace[3 * (gen - 1) + 2] = global_ns();
  }
  __syncthreads();
}

__device__ __forceinline__ void cp_async16_cg(void* smem, const void* gmem, bool valid) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// GEMM tile: out[m0.., n0..] = epi(A[m, k_begin..k_begin+256) . W[n, k_begin..+256))
//   LN_MODE 0: A rows come from global `a` (row stride lda)
//   LN_MODE n>0: A row = LayerNorm(xin[m] + ln_bias + sum_{s<n} delta[s][m]) * g + b   (k_begin must be 0, K = 256);
//              tiles with n0 == 0 also store the normalised rows to xout
//   EPI 0: + bias   1: gelu_erf(+ bias)   2: raw store to slab z (split-K partial)
// ---------------------------------------------------------------------------------------------------------------------
struct LnArgs {
  const float* xin;
  float* xout;
  const float* delta;
  int nsplit;
  const float* ln_bias;
  const float* g;
  const float* b;
};

// PDL = launched with programmatic stream serialization: the weight panel (independent of the previous kernel) is requested
// first, then griddepcontrol.wait orders everything that reads or writes activations after the predecessor's completion.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <int BM, int BN, int TM, int TN, int LN_MODE, int EPI, bool PDL = false>
__device__ __forceinline__ void gemm_tile(float* s_panel, const float* a, int64_t lda, const LnArgs& ln, const float* W, int K,
                                          int k_begin, const float* bias, float* out, int64_t ldo, int M, int N, int m0,
                                          int n0) {
  constexpr int TX = BN / TN, TY = BM / TM;
  static_assert(TX * TY == kPThreads, "256 threads");
  float (*As)[kPanelLds] = reinterpret_cast<float (*)[kPanelLds]>(s_panel);
  float (*Bs)[kPanelLds] = reinterpret_cast<float (*)[kPanelLds]>(s_panel + BM * kPanelLds);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid % TX, ty = tid / TX;

  // W panel (and the A panel when it is plain data): 16-byte cp.async through L2
  for (int i = tid; i < BN * 64; i += kPThreads) {
    const int r = i >> 6, kq = (i & 63) * 4;
    const bool ok = n0 + r < N;
    cp_async16_cg(&Bs[r][kq], ok ? W + (int64_t)(n0 + r) * K + k_begin + kq : W, ok);
  }
  if (PDL) pdl_wait();
  if (LN_MODE == 0) {
    for (int i = tid; i < BM * 64; i += kPThreads) {
      const int r = i >> 6, kq = (i & 63) * 4;
      const bool ok = m0 + r < M;
      cp_async16_cg(&As[r][kq], ok ? a + (int64_t)(m0 + r) * lda + k_begin + kq : a, ok);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (LN_MODE != 0) {
    // LayerNorm-on-load: warp w owns rows w, w+8, ... of the tile.  LN_MODE = number of delta slabs (1 or 8), a compile-time
    // constant so that every load of every row is issued before the first reduction (one L2 round trip, not one per row).
    constexpr int RPW = BM / (kPThreads / 32);
    constexpr int NSPLIT = LN_MODE;
    float v[RPW][8];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
      const int m = m0 + warp + rr * (kPThreads / 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        // same association as add_ln_kernel: x + (bias + delta_0 + delta_1 + ...), so the execution modes stay bit-identical
        float d = ln.ln_bias ? ln.ln_bias[c] : 0.f;
        float xv = 0.f;
        if (m < M) {
          xv = __ldcg(&ln.xin[(int64_t)m * kPD + c]);
#pragma unroll
          for (int sp = 0; sp < NSPLIT; ++sp) d += __ldcg(&ln.delta[((int64_t)sp * M + m) * kPD + c]);
        }
        v[rr][j] = xv + d;
      }
    }
    float gg[8], bb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      gg[j] = ln.g[c];
      bb[j] = ln.b[c];
    }
    float mean[RPW], rstd[RPW];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[rr][j];
      mean[rr] = warp_sum(s) * (1.f / kPD);
    }
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) q += (v[rr][j] - mean[rr]) * (v[rr][j] - mean[rr]);
      rstd[rr] = 1.f / sqrtf(warp_sum(q) * (1.f / kPD) + 1e-5f);
    }
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
      const int r = warp + rr * (kPThreads / 32);
      const int m = m0 + r;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        const float y = (m < M) ? (v[rr][j] - mean[rr]) * rstd[rr] * gg[j] + bb[j] : 0.f;
        if (m < M && n0 == 0) ln.xout[(int64_t)m * kPD + c] = y;
        As[r][c] = y;
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int kk = 0; kk < 256; kk += 4) {
    float4 av[TM], bv[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) av[i] = *reinterpret_cast<const float4*>(&As[ty + i * TY][kk]);
#pragma unroll
    for (int j = 0; j < TN; ++j) bv[j] = *reinterpret_cast<const float4*>(&Bs[tx + j * TX][kk]);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]);
        acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
        acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]);
        acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
      }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty + i * TY;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx + j * TX;
      if (n >= N) continue;
      float v = acc[i][j];
      if (EPI != 2) v += bias[n];
      if (EPI == 1) v = gelu_erf(v);
      out[(int64_t)m * ldo + n] = v;
    }
  }
  __syncthreads();  // the panel is reused by the next tile / phase
}

template <int BM, int BN, int TM, int TN, int LN_MODE, int EPI, bool PDL = false>
__device__ __forceinline__ void gemm_phase(float* s_panel, const float* a, int64_t lda, const LnArgs& ln, const float* W, int K,
                                           const float* bias, float* out, int64_t ldo, int M, int N) {
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const int k_slices = (EPI == 2) ? K / 256 : 1;
  const int n_tiles = tiles_m * tiles_n * k_slices;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int z = t / (tiles_m * tiles_n);
    const int rem = t - z * tiles_m * tiles_n;
    const int tm = rem / tiles_n, tn = rem - tm * tiles_n;
    gemm_tile<BM, BN, TM, TN, LN_MODE, EPI, PDL>(s_panel, a, lda, ln, W, K, z * 256, bias, out + (int64_t)z * M * ldo, ldo, M,
                                                 N, tm * BM, tn * BN);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// attention phases (one warp per (row, head)); same arithmetic as decoder.cu
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void self_attn_phase(const float* qkv, float* kcache, float* vcache, const int* src_row, int pos,
                                                int max_len, float* attn, int rows) {
  const int lane = threadIdx.x & 31;
  const int n_tasks = rows * kPHeads;
  for (int gw = blockIdx.x * (kPThreads / 32) + (threadIdx.x >> 5); gw < n_tasks; gw += gridDim.x * (kPThreads / 32))
    self_attention_task<true>(qkv, kcache, vcache, src_row, pos, max_len, attn, gw / kPHeads, gw % kPHeads, lane);
}

__device__ __forceinline__ void cross_attn_phase(float* s_scores, const float* q, const float* ck, const float* cv,
                                                 int64_t kv_stride, const int* lens, int beam, int tp, float* attn, int rows) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int n_tasks = rows * kPHeads;
  for (int gw = blockIdx.x * (kPThreads / 32) + wib; gw < n_tasks; gw += gridDim.x * (kPThreads / 32)) {
    const int r = gw / kPHeads, h = gw % kPHeads;
    const int clip = r / beam;
    cross_attention_task<true>(s_scores + wib * tp, q, ck, cv, kv_stride, lens[clip], clip, tp, attn, r, h, lane);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// beam step for one clip (same selection logic as beam.cu::beam_step_kernel) + embedding of the next input token
// ---------------------------------------------------------------------------------------------------------------------
struct PCand {
  float v;
  int idx;
};
__device__ __forceinline__ bool pbetter(const PCand& a, const PCand& b) { return a.v > b.v || (a.v == b.v && a.idx < b.idx); }

struct BeamSmem {
  float lse_max[kPMaxBeam], lse_log[kPMaxBeam];
  float red[2][kPMaxBeam][kPThreads / 32];
  PCand cand[2][kPThreads / 32];
  int owner[2][kPThreads / 32];
  PCand win[kPMaxBeam];
};

__device__ void beam_clip(BeamSmem& sm, float* logits, const uint8_t* forbid, const BeamState& st, int step, int cur, int min_len,
                          int beam, int max_len, int vocab, int clip, const float* emb, const float* pe, float* x_next) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = clip * beam;
  const int tstride = max_len + 1;
  const int* tok_cur = st.tokens[cur];
  int* tok_new = st.tokens[cur ^ 1];
  const int* src_cur = st.src_row[cur];
  int* src_new = st.src_row[cur ^ 1];
  int live_label[kPMaxBeam];
  float prev_sum[kPMaxBeam];
  int nlive = 0;
#pragma unroll
  for (int l = 0; l < kPMaxBeam; ++l) {
    live_label[l] = 0;
    prev_sum[l] = 0.f;
  }
#pragma unroll
  for (int l = 0; l < kPMaxBeam; ++l)
    if (l < beam && st.live[row0 + l]) {
#pragma unroll
      for (int q = 0; q < kPMaxBeam; ++q)
        if (q == nlive) {
          live_label[q] = l;
          prev_sum[q] = st.sum_lp[row0 + l];
        }
      ++nlive;
    }
  auto label_at = [&](int q) {
    int r = 0;
#pragma unroll
    for (int i = 0; i < kPMaxBeam; ++i)
      if (i == q) r = live_label[i];
    return r;
  };
  __syncthreads();  // all threads have read live / sum_lp before anybody updates them below
  if (nlive > 0) {
    const int nrows_used = (step == 0) ? 1 : nlive;
    const int k_sel = nlive;
    for (int j = 0; j < nrows_used; ++j) {
      const int row = row0 + label_at(j);
      float* lg = logits + (int64_t)row * vocab;
      if (tid == 0 && step < min_len) lg[2] = -INFINITY;
      if (forbid != nullptr && tid <= step) {
        const int tok = tok_cur[(int64_t)row * tstride + tid];
        if (forbid[tok]) lg[tok] = -INFINITY;
      }
    }
    __syncthreads();
    // log-softmax statistics, all 8 warps per row: unrolled strided loads (8 in flight per thread), two-level reductions
    for (int j = 0; j < nrows_used; ++j) {
      const float* lg = logits + (int64_t)(row0 + label_at(j)) * vocab;
      float mx = -INFINITY;
      for (int v0 = tid; v0 < vocab; v0 += 8 * kPThreads) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = (v0 + u * kPThreads < vocab) ? lg[v0 + u * kPThreads] : -INFINITY;
#pragma unroll
        for (int u = 0; u < 8; ++u) mx = fmaxf(mx, t[u]);
      }
      mx = warp_max(mx);
      if (lane == 0) sm.red[0][j][warp] = mx;
    }
    __syncthreads();
    for (int j = 0; j < nrows_used; ++j) {
      const float* lg = logits + (int64_t)(row0 + label_at(j)) * vocab;
      float mx = sm.red[0][j][0];
#pragma unroll
      for (int i = 1; i < kPThreads / 32; ++i) mx = fmaxf(mx, sm.red[0][j][i]);
      float s = 0.f;
      for (int v0 = tid; v0 < vocab; v0 += 8 * kPThreads) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = (v0 + u * kPThreads < vocab) ? lg[v0 + u * kPThreads] : -INFINITY;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += expf(t[u] - mx);
      }
      s = warp_sum(s);
      if (lane == 0) sm.red[1][j][warp] = s;
      if (tid == 0) sm.lse_max[j] = mx;
    }
    __syncthreads();
    if (tid < nrows_used) {
      float s = 0.f;
      for (int i = 0; i < kPThreads / 32; ++i) s += sm.red[1][tid][i];  // fixed order
      sm.lse_log[tid] = logf(s);
    }
    __syncthreads();
    PCand loc[kPMaxBeam];
#pragma unroll
    for (int i = 0; i < kPMaxBeam; ++i) loc[i] = PCand{-INFINITY, 0x7fffffff};
    for (int j = 0; j < nrows_used; ++j) {
      const float* lg = logits + (int64_t)(row0 + label_at(j)) * vocab;
      const float mx = sm.lse_max[j], lg_sum = sm.lse_log[j];
      float prev = 0.f;
#pragma unroll
      for (int i = 0; i < kPMaxBeam; ++i)
        if (i == j) prev = prev_sum[i];
      for (int v0 = tid; v0 < vocab; v0 += 8 * kPThreads) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = (v0 + u * kPThreads < vocab) ? lg[v0 + u * kPThreads] : -INFINITY;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int v = v0 + u * kPThreads;
          if (v >= vocab) break;
          const float lsm = (t[u] - mx) - lg_sum;
          const PCand c{step == 0 ? lsm : prev + lsm, j * vocab + v};
          if (pbetter(c, loc[kPMaxBeam - 1])) {
            loc[kPMaxBeam - 1] = c;
#pragma unroll
            for (int i = kPMaxBeam - 1; i > 0; --i)
              if (pbetter(loc[i], loc[i - 1])) {
                const PCand tt = loc[i];
                loc[i] = loc[i - 1];
                loc[i - 1] = tt;
              }
          }
        }
      }
    }
    for (int r = 0; r < k_sel; ++r) {
      PCand best = loc[0];
      int owner = tid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const PCand other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.idx, o)};
        const int oo = __shfl_xor_sync(0xffffffffu, owner, o);
        if (pbetter(other, best)) {
          best = other;
          owner = oo;
        }
      }
      const int buf = r & 1;
      if (lane == 0) {
        sm.cand[buf][warp] = best;
        sm.owner[buf][warp] = owner;
      }
      __syncthreads();
      PCand bw = sm.cand[buf][0];
      int ow = sm.owner[buf][0];
#pragma unroll
      for (int i = 1; i < kPThreads / 32; ++i)
        if (pbetter(sm.cand[buf][i], bw)) {
          bw = sm.cand[buf][i];
          ow = sm.owner[buf][i];
        }
      if (bw.idx == 0x7fffffff) bw.idx = 0;
      if (tid == 0) sm.win[r] = bw;
      if (tid == ow) {
#pragma unroll
        for (int i = 0; i < kPMaxBeam - 1; ++i) loc[i] = loc[i + 1];
        loc[kPMaxBeam - 1] = PCand{-INFINITY, 0x7fffffff};
      }
    }
    __syncthreads();
    for (int item = tid; item < k_sel * (step + 2); item += kPThreads) {
      const int r = item / (step + 2), p = item - r * (step + 2);
      const int row = row0 + label_at(r);
      const int prev_pos = sm.win[r].idx / vocab;
      const int word = sm.win[r].idx - prev_pos * vocab;
      const int src = row0 + label_at(prev_pos);
      if (p <= step) {
        tok_new[(int64_t)row * tstride + p] = tok_cur[(int64_t)src * tstride + p];
        src_new[(int64_t)row * max_len + p] = src_cur[(int64_t)src * max_len + p];
      } else {
        tok_new[(int64_t)row * tstride + p] = word;
        if (p < max_len) src_new[(int64_t)row * max_len + p] = row;
      }
    }
    __syncthreads();
    if (tid < k_sel) {
      const int r = tid;
      const int row = row0 + label_at(r);
      const int prev_pos = sm.win[r].idx / vocab;
      const int word = sm.win[r].idx - prev_pos * vocab;
      st.sum_lp[row] = sm.win[r].v;
      if (word == 2 || step == max_len - 1) {
        for (int p = 0; p <= step; ++p) st.out_preds[(int64_t)row * max_len + p] = tok_new[(int64_t)row * tstride + p + 1];
        st.out_lp[row] = sm.win[r].v / (float)(step + 1);
        st.live[row] = 0;
        const int left = atomicSub(&st.done[2], 1) - 1;
        if (left == 0) {
          st.done[1] = step + 1;
          __threadfence();
          st.done[0] = 1;
        }
      }
    }
    __syncthreads();
  }
  // embedding of the next input token for every row of the clip (dead rows keep a valid stale token)
  if (step + 1 < max_len) {
    for (int i = tid; i < beam * kPD; i += kPThreads) {
      const int row = row0 + i / kPD, c = i % kPD;
      const int tok = tok_new[(int64_t)row * tstride + step + 1];
      x_next[(int64_t)row * kPD + c] = emb[(int64_t)tok * kPD + c] * 16.0f + pe[(int64_t)(step + 1) * kPD + c];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPThreads, 1) decoder_persistent_kernel(const PersistentArgs p) {
  extern __shared__ __align__(16) float s_dyn[];
  __shared__ BeamSmem s_beam;
  unsigned int gen = 0;
  const int R = p.rows;
  const int64_t cache_l = (int64_t)R * p.max_len * kPD;
  const int64_t kv_stride = kPLayers * 2 * kPD;
  const int tstride = p.max_len + 1;
  float* x_cur = p.xa;  // residual stream (double-buffered across LayerNorm phases)
  float* x_alt = p.xb;

  // ---- init (what beam_init_kernel + the first embedding do) ------------------------------------------------------
  for (int r = blockIdx.x * kPThreads + threadIdx.x; r < R; r += gridDim.x * kPThreads) {
    for (int q = 0; q <= p.max_len; ++q) {
      p.bs.tokens[0][(int64_t)r * tstride + q] = 0;
      p.bs.tokens[1][(int64_t)r * tstride + q] = 0;
    }
    p.bs.tokens[0][(int64_t)r * tstride] = (int)p.bos_ids[r / p.beam];
    for (int q = 0; q < p.max_len; ++q) {
      p.bs.src_row[0][(int64_t)r * p.max_len + q] = r;
      p.bs.src_row[1][(int64_t)r * p.max_len + q] = r;
      p.bs.out_preds[(int64_t)r * p.max_len + q] = 0;
    }
    p.bs.sum_lp[r] = 0.f;
    p.bs.live[r] = 1;
    p.bs.out_lp[r] = 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.bs.done[0] = 0;
    p.bs.done[1] = p.max_len;
    p.bs.done[2] = R;
  }
  for (int i = blockIdx.x * kPThreads + threadIdx.x; i < R * kPD; i += gridDim.x * kPThreads) {
    const int r = i / kPD, c = i - r * kPD;
    const int tok = (int)p.bos_ids[r / p.beam];
    x_cur[i] = p.emb[(int64_t)tok * kPD + c] * 16.0f + p.pe[c];
  }
  grid_barrier(p.bar, gen, p.trace, 0);

  int cur = 0;
  for (int step = 0; step < p.max_len; ++step) {
    const int* src_row = p.bs.src_row[cur];
    for (int l = 0; l < kPLayers; ++l) {
      const PLayer& L = p.layers[l];
      LnArgs ln{};
      // ---- QKV (layer 0: x is the embedding; later layers: LN3 of the previous layer on load)
      if (l == 0) {
        gemm_phase<32, 32, 2, 2, 0, 0>(s_dyn, x_cur, kPD, ln, L.sa_in_w, kPD, L.sa_in_b, p.qkv, 768, R, 768);
      } else {
        const PLayer& P = p.layers[l - 1];
        ln = LnArgs{x_cur, x_alt, p.part, kPSplits, P.l2_b, P.n3_g, P.n3_b};
        gemm_phase<32, 32, 2, 2, kPSplits, 0>(s_dyn, nullptr, 0, ln, L.sa_in_w, kPD, L.sa_in_b, p.qkv, 768, R, 768);
        float* t = x_cur; x_cur = x_alt; x_alt = t;
      }
      grid_barrier(p.bar, gen, p.trace, 1);
      self_attn_phase(p.qkv, p.kc + l * cache_l, p.vc + l * cache_l, src_row, step, p.max_len, p.attn, R);
      grid_barrier(p.bar, gen, p.trace, 2);
      gemm_phase<32, 32, 2, 2, 0, 0>(s_dyn, p.attn, kPD, ln, L.sa_out_w, kPD, L.sa_out_b, p.tmp, kPD, R, kPD);
      grid_barrier(p.bar, gen, p.trace, 3);
      // ---- LN1 on load + cross-attention query projection (q goes to the free qkv buffer, row stride 256)
      ln = LnArgs{x_cur, x_alt, p.tmp, 1, nullptr, L.n1_g, L.n1_b};
      gemm_phase<32, 32, 2, 2, 1, 0>(s_dyn, nullptr, 0, ln, L.ca_q_w, kPD, L.ca_q_b, p.qkv, kPD, R, kPD);
      { float* t = x_cur; x_cur = x_alt; x_alt = t; }
      grid_barrier(p.bar, gen, p.trace, 4);
      cross_attn_phase(s_dyn, p.qkv, p.ckv + (int64_t)l * 2 * kPD, p.ckv + (int64_t)l * 2 * kPD + kPD, kv_stride, p.lens, p.beam,
                       p.tp, p.attn, R);
      grid_barrier(p.bar, gen, p.trace, 5);
      gemm_phase<32, 32, 2, 2, 0, 0>(s_dyn, p.attn, kPD, ln, L.ca_out_w, kPD, L.ca_out_b, p.tmp, kPD, R, kPD);
      grid_barrier(p.bar, gen, p.trace, 6);
      // ---- LN2 on load + FF1 (GELU)
      ln = LnArgs{x_cur, x_alt, p.tmp, 1, nullptr, L.n2_g, L.n2_b};
      gemm_phase<48, 64, 3, 4, 1, 1>(s_dyn, nullptr, 0, ln, L.l1_w, kPD, L.l1_b, p.ff, kPFF, R, kPFF);
      { float* t = x_cur; x_cur = x_alt; x_alt = t; }
      grid_barrier(p.bar, gen, p.trace, 7);
      // ---- FF2: 8 K-slices of raw partial sums (bias + residual + LN3 happen on the next load)
      gemm_phase<64, 64, 4, 4, 0, 2>(s_dyn, p.ff, kPFF, ln, L.l2_w, kPFF, nullptr, p.part, kPD, R, kPD);
      grid_barrier(p.bar, gen, p.trace, 8);
    }
    {  // ---- LN3 of the last layer on load + classifier
      const PLayer& P = p.layers[kPLayers - 1];
      LnArgs ln{x_cur, x_alt, p.part, kPSplits, P.l2_b, P.n3_g, P.n3_b};
      gemm_phase<64, 96, 4, 6, kPSplits, 0>(s_dyn, nullptr, 0, ln, p.cls_w, kPD, p.cls_b, p.logits, p.vocab, R, p.vocab);
      grid_barrier(p.bar, gen, p.trace, 9);
    }
    // ---- beam step: one CTA per clip; it also writes the next step's embedding into x_cur (free: LN wrote x_alt)
    for (int clip = blockIdx.x; clip < p.batch; clip += gridDim.x)
      beam_clip(s_beam, p.logits, p.forbid, p.bs, step, cur, p.min_len, p.beam, p.max_len, p.vocab, clip, p.emb, p.pe, x_cur);
    cur ^= 1;
    grid_barrier(p.bar, gen, p.trace, 10);
    if (*reinterpret_cast<volatile int*>(&p.bs.done[0])) break;  // uniform: read after the barrier by every CTA
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same phases as stand-alone kernels ("fused graph" decoder mode): LayerNorm-on-load GEMMs, attention phases and the
// beam step (+ next-token embedding) are launched per phase and replayed from a CUDA graph -- 50 launches per step instead
// of 69 -- with bit-identical arithmetic to the persistent kernel.
// ---------------------------------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN, int LN_MODE, int EPI>
__global__ void __launch_bounds__(kPThreads)
gemm_phase_kernel(const float* a, int64_t lda, LnArgs ln, const float* W, int K, const float* bias, float* out, int64_t ldo,
                  int M, int N) {
  extern __shared__ __align__(16) float s_dyn[];
  pdl_launch_dependents();  // the next phase may start its prologue (weight panel loads) while this one drains
  gemm_phase<BM, BN, TM, TN, LN_MODE, EPI, true>(s_dyn, a, lda, ln, W, K, bias, out, ldo, M, N);
}

__global__ void __launch_bounds__(kPThreads)
self_attn_phase_kernel(const float* qkv, float* kcache, float* vcache, const int* src_row, int pos, int max_len, float* attn,
                       int rows) {
  pdl_launch_dependents();
  pdl_wait();
  self_attn_phase(qkv, kcache, vcache, src_row, pos, max_len, attn, rows);
}

__global__ void __launch_bounds__(kPThreads)
cross_attn_phase_kernel(const float* q, const float* ck, const float* cv, int64_t kv_stride, const int* lens, int beam, int tp,
                        float* attn, int rows) {
  extern __shared__ __align__(16) float s_dyn[];
  pdl_launch_dependents();
  pdl_wait();
  cross_attn_phase(s_dyn, q, ck, cv, kv_stride, lens, beam, tp, attn, rows);
}

__global__ void __launch_bounds__(kPThreads)
beam_embed_kernel(float* logits, const uint8_t* forbid, BeamState st, int step, int cur, int min_len, int beam, int max_len,
                  int vocab, const float* emb, const float* pe, float* x_next) {
  __shared__ BeamSmem sm;
  pdl_launch_dependents();
  pdl_wait();
  if (st.done[0]) return;
  beam_clip(sm, logits, forbid, st, step, cur, min_len, beam, max_len, vocab, blockIdx.x, emb, pe, x_next);
}

static bool g_use_pdl = true;
void decoder_set_pdl(bool on) { g_use_pdl = on; }

// launch with the programmatic-stream-serialization attribute (PDL); also valid inside stream capture
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int BM, int BN, int TM, int TN, int LN_MODE, int EPI>
static int launch_phase_cfg(const float* a, int64_t lda, const LnArgs& ln, const float* W, int K, const float* bias, float* out,
                            int64_t ldo, int M, int N, cudaStream_t stream) {
  constexpr int smem = (BM + BN) * kPanelLds * (int)sizeof(float);
  auto kern = gemm_phase_kernel<BM, BN, TM, TN, LN_MODE, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    CNB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int n_tiles = (int)(ceil_div(M, BM) * ceil_div(N, BN)) * (EPI == 2 ? K / 256 : 1);
  CNB_CUDA_OK(launch_pdl(kern, dim3(n_tiles), dim3(kPThreads), smem, stream, a, lda, ln, W, K, bias, out, ldo, M, N));
  count_launch();
  return 0;
}

// one decoder step (position `step`) as 50 phase launches; x_cur / x_alt are swapped in place like in the persistent kernel
int launch_decoder_step_fused(const PersistentArgs& p, int step, int cur, float** x_cur_io, float** x_alt_io,
                              cudaStream_t st) {
  float* x_cur = *x_cur_io;
  float* x_alt = *x_alt_io;
  const int R = p.rows;
  const int64_t cache_l = (int64_t)R * p.max_len * kPD;
  const int64_t kv_stride = kPLayers * 2 * kPD;
  const int* src_row = p.bs.src_row[cur];
  const int attn_blocks = (int)ceil_div((int64_t)R * kPHeads, kPThreads / 32);
  for (int l = 0; l < kPLayers; ++l) {
    const PLayer& L = p.layers[l];
    LnArgs ln{};
    if (l == 0) {
      if (int rc = launch_phase_cfg<32, 32, 2, 2, 0, 0>(x_cur, kPD, ln, L.sa_in_w, kPD, L.sa_in_b, p.qkv, 768, R, 768, st)) return rc;
    } else {
      const PLayer& P = p.layers[l - 1];
      ln = LnArgs{x_cur, x_alt, p.part, kPSplits, P.l2_b, P.n3_g, P.n3_b};
      if (int rc = launch_phase_cfg<32, 32, 2, 2, kPSplits, 0>(nullptr, 0, ln, L.sa_in_w, kPD, L.sa_in_b, p.qkv, 768, R, 768, st)) return rc;
      std::swap(x_cur, x_alt);
    }
    CNB_CUDA_OK(launch_pdl(self_attn_phase_kernel, dim3(attn_blocks), dim3(kPThreads), 0, st, (const float*)p.qkv,
                           p.kc + l * cache_l, p.vc + l * cache_l, src_row, step, p.max_len, p.attn, R));
    count_launch();
    if (int rc = launch_phase_cfg<32, 32, 2, 2, 0, 0>(p.attn, kPD, ln, L.sa_out_w, kPD, L.sa_out_b, p.tmp, kPD, R, kPD, st)) return rc;
    ln = LnArgs{x_cur, x_alt, p.tmp, 1, nullptr, L.n1_g, L.n1_b};
    if (int rc = launch_phase_cfg<32, 32, 2, 2, 1, 0>(nullptr, 0, ln, L.ca_q_w, kPD, L.ca_q_b, p.qkv, kPD, R, kPD, st)) return rc;
    std::swap(x_cur, x_alt);
    CNB_CUDA_OK(launch_pdl(cross_attn_phase_kernel, dim3(attn_blocks), dim3(kPThreads), (size_t)8 * p.tp * sizeof(float), st,
                           (const float*)p.qkv, p.ckv + (int64_t)l * 2 * kPD, p.ckv + (int64_t)l * 2 * kPD + kPD, kv_stride,
                           p.lens, p.beam, p.tp, p.attn, R));
    count_launch();
    if (int rc = launch_phase_cfg<32, 32, 2, 2, 0, 0>(p.attn, kPD, ln, L.ca_out_w, kPD, L.ca_out_b, p.tmp, kPD, R, kPD, st)) return rc;
    ln = LnArgs{x_cur, x_alt, p.tmp, 1, nullptr, L.n2_g, L.n2_b};
    if (int rc = launch_phase_cfg<48, 64, 3, 4, 1, 1>(nullptr, 0, ln, L.l1_w, kPD, L.l1_b, p.ff, kPFF, R, kPFF, st)) return rc;
    std::swap(x_cur, x_alt);
    if (int rc = launch_phase_cfg<64, 64, 4, 4, 0, 2>(p.ff, kPFF, ln, L.l2_w, kPFF, nullptr, p.part, kPD, R, kPD, st)) return rc;
  }
  {
    const PLayer& P = p.layers[kPLayers - 1];
    LnArgs ln{x_cur, x_alt, p.part, kPSplits, P.l2_b, P.n3_g, P.n3_b};
    if (int rc = launch_phase_cfg<64, 96, 4, 6, kPSplits, 0>(nullptr, 0, ln, p.cls_w, kPD, p.cls_b, p.logits, p.vocab, R, p.vocab, st))
      return rc;
  }
  CNB_CUDA_OK(launch_pdl(beam_embed_kernel, dim3(p.batch), dim3(kPThreads), 0, st, p.logits, p.forbid, p.bs, step, cur,
                         p.min_len, p.beam, p.max_len, p.vocab, p.emb, p.pe, x_cur));
  count_launch();
  *x_cur_io = x_cur;
  *x_alt_io = x_alt;
  return 0;
}

// beam-state initialisation + embedding of the task BOS tokens into x (what the persistent kernel does before step 0)
__global__ void decoder_init_kernel(const PersistentArgs p) {
  const int R = p.rows, tstride = p.max_len + 1;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    for (int q = 0; q <= p.max_len; ++q) {
      p.bs.tokens[0][(int64_t)r * tstride + q] = 0;
      p.bs.tokens[1][(int64_t)r * tstride + q] = 0;
    }
    p.bs.tokens[0][(int64_t)r * tstride] = (int)p.bos_ids[r / p.beam];
    for (int q = 0; q < p.max_len; ++q) {
      p.bs.src_row[0][(int64_t)r * p.max_len + q] = r;
      p.bs.src_row[1][(int64_t)r * p.max_len + q] = r;
      p.bs.out_preds[(int64_t)r * p.max_len + q] = 0;
    }
    p.bs.sum_lp[r] = 0.f;
    p.bs.live[r] = 1;
    p.bs.out_lp[r] = 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.bs.done[0] = 0;
    p.bs.done[1] = p.max_len;
    p.bs.done[2] = R;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * kPD; i += gridDim.x * blockDim.x) {
    const int r = i / kPD, c = i - r * kPD;
    const int tok = (int)p.bos_ids[r / p.beam];
    p.xa[i] = p.emb[(int64_t)tok * kPD + c] * 16.0f + p.pe[c];
  }
}

int launch_decoder_init(const PersistentArgs& p, cudaStream_t st) {
  decoder_init_kernel<<<(p.rows * kPD + 255) / 256, 256, 0, st>>>(p);
  CNB_LAUNCH_OK();
  return 0;
}

int launch_decoder_persistent(const PersistentArgs& args, cudaStream_t stream) {
  static int max_blocks_per_sm = -1;
  const size_t smem = (size_t)(64 + 96) * kPanelLds * sizeof(float);
  if (max_blocks_per_sm < 0) {
    CNB_CUDA_OK(cudaFuncSetAttribute(decoder_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CNB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, decoder_persistent_kernel, kPThreads, smem));
  }
  CNB_REQUIRE(max_blocks_per_sm >= 1, "persistent decoder kernel does not fit on an SM");
  CNB_REQUIRE((size_t)8 * args.tp * sizeof(float) <= smem, "too many encoder frames for the cross-attention score buffer");
  int dev = 0, n_sm = 0;
  CNB_CUDA_OK(cudaGetDevice(&dev));
  CNB_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  CNB_CUDA_OK(cudaMemsetAsync(args.bar, 0, 2 * sizeof(unsigned int), stream));
  void* kargs[] = {const_cast<PersistentArgs*>(&args)};
  CNB_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(decoder_persistent_kernel), dim3(n_sm), dim3(kPThreads), kargs,
                                          smem, stream));
  count_launch();
  return 0;
}

}  // namespace cnb
