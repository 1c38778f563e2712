// K-BEAM: one CTA per clip does the whole per-step selection of the reference's batched beam search without any host
// synchronisation: EOS / no-repeat masking, log-softmax over the vocabulary for every live beam, flat top-k over
// (live beams x V) by warp-shuffle reductions, history / KV-cache back-pointer update and finish bookkeeping.
// Reference: nn/decoding/beam.py:113-203 (step loop) and :230-269 (_select_k_next_toks); semantics in SURVEY.md Appendix B.
// Fixed-slot formulation: physical row = clip * beam + label; labels stick to row positions, finished labels leave the
// live set, the r-th best candidate goes to the r-th live label (shown equal to the reference's row-compacting code by
// tests/test_oracle_vs_reference.py on the CPU restatement this kernel mirrors).
#include "common.cuh"
#include "kernels.h"

namespace cnb {

constexpr int kBeamThreads = 256;
constexpr int kMaxBeam = 8;
constexpr int kPad = 0, kEos = 2;

__global__ void beam_init_kernel(const int64_t* __restrict__ bos_ids, BeamState st, int rows, int beam, int max_len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0) {
    st.done[0] = 0;
    st.done[1] = max_len;
    st.done[2] = rows;  // live rows remaining
  }
  if (r >= rows) return;
  for (int p = 0; p <= max_len; ++p) {
    st.tokens[0][(int64_t)r * (max_len + 1) + p] = kPad;
    st.tokens[1][(int64_t)r * (max_len + 1) + p] = kPad;
  }
  st.tokens[0][(int64_t)r * (max_len + 1)] = (int)bos_ids[r / beam];
  for (int p = 0; p < max_len; ++p) {
    st.src_row[0][(int64_t)r * max_len + p] = r;
    st.src_row[1][(int64_t)r * max_len + p] = r;
    st.out_preds[(int64_t)r * max_len + p] = kPad;
  }
  st.sum_lp[r] = 0.f;
  st.live[r] = 1;
  st.out_lp[r] = 0.f;
}

struct Cand {
  float v;
  int idx;  // flat index: live_position * V + word
};
__device__ __forceinline__ bool better(const Cand& a, const Cand& b) {
  return a.v > b.v || (a.v == b.v && a.idx < b.idx);
}

__global__ void __launch_bounds__(kBeamThreads)
beam_step_kernel(float* __restrict__ logits, const uint8_t* __restrict__ forbid, BeamState st, int step, int cur, int min_len,
                 int beam, int max_len, int vocab) {
  if (st.done[0]) return;
  __shared__ float s_lse_max[kMaxBeam];
  __shared__ float s_lse_log[kMaxBeam];
  __shared__ float s_redw[2][kMaxBeam][kBeamThreads / 32];
  __shared__ Cand s_cand[2][kBeamThreads / 32];
  __shared__ int s_owner[2][kBeamThreads / 32];
  __shared__ Cand s_win[kMaxBeam];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int clip = blockIdx.x;
  const int row0 = clip * beam;
  const int tstride = max_len + 1;
  const int* tok_cur = st.tokens[cur];
  int* tok_new = st.tokens[cur ^ 1];
  const int* src_cur = st.src_row[cur];
  int* src_new = st.src_row[cur ^ 1];

  // live labels of this clip in ascending order (every thread derives the same list: no broadcast barrier needed)
  int live_label[kMaxBeam];
  float prev_sum[kMaxBeam];
  int nlive = 0;
#pragma unroll
  for (int l = 0; l < kMaxBeam; ++l) {
    live_label[l] = 0;
    prev_sum[l] = 0.f;
  }
#pragma unroll
  for (int l = 0; l < kMaxBeam; ++l)
    if (l < beam && st.live[row0 + l]) {
#pragma unroll
      for (int q = 0; q < kMaxBeam; ++q)
        if (q == nlive) {
          live_label[q] = l;
          prev_sum[q] = st.sum_lp[row0 + l];
        }
      ++nlive;
    }
  if (nlive == 0) return;
  const int nrows_used = (step == 0) ? 1 : nlive;  // step 0: only the first row (beam.py:243-246)
  const int k_sel = nlive;                         // number of candidates to select (= beam at step 0)
  auto label_at = [&](int q) {
    int r = 0;
#pragma unroll
    for (int i = 0; i < kMaxBeam; ++i)
      if (i == q) r = live_label[i];
    return r;
  };

  // ---- masks: EOS before min_len (beam.py:129-130), no-repeat of forbidden tokens already in the history (:146-156)
  for (int j = 0; j < nrows_used; ++j) {
    const int row = row0 + label_at(j);
    float* lg = logits + (int64_t)row * vocab;
    if (tid == 0 && step < min_len) lg[kEos] = -INFINITY;
    if (forbid != nullptr && tid <= step) {
      const int tok = tok_cur[(int64_t)row * tstride + tid];
      if (forbid[tok]) lg[tok] = -INFINITY;
    }
  }
  __syncthreads();

  // ---- log-softmax statistics, all 8 warps per row: unrolled strided loads (8 in flight per thread), two-level reductions
  for (int j = 0; j < nrows_used; ++j) {
    const float* lg = logits + (int64_t)(row0 + label_at(j)) * vocab;
    float mx = -INFINITY;
    for (int v0 = tid; v0 < vocab; v0 += 8 * kBeamThreads) {
      float t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = (v0 + u * kBeamThreads < vocab) ? lg[v0 + u * kBeamThreads] : -INFINITY;
#pragma unroll
      for (int u = 0; u < 8; ++u) mx = fmaxf(mx, t[u]);
    }
    mx = warp_max(mx);
    if (lane == 0) s_redw[0][j][warp] = mx;
  }
  __syncthreads();
  for (int j = 0; j < nrows_used; ++j) {
    const float* lg = logits + (int64_t)(row0 + label_at(j)) * vocab;
    float mx = s_redw[0][j][0];
#pragma unroll
    for (int i = 1; i < kBeamThreads / 32; ++i) mx = fmaxf(mx, s_redw[0][j][i]);
    float sm = 0.f;
    for (int v0 = tid; v0 < vocab; v0 += 8 * kBeamThreads) {
      float t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = (v0 + u * kBeamThreads < vocab) ? lg[v0 + u * kBeamThreads] : -INFINITY;
#pragma unroll
      for (int u = 0; u < 8; ++u) sm += expf(t[u] - mx);
    }
    sm = warp_sum(sm);
    if (lane == 0) s_redw[1][j][warp] = sm;
    if (tid == 0) s_lse_max[j] = mx;
  }
  __syncthreads();
  if (tid < nrows_used) {
    float t = 0.f;
    for (int i = 0; i < kBeamThreads / 32; ++i) t += s_redw[1][tid][i];  // fixed order
    s_lse_log[tid] = logf(t);
  }
  __syncthreads();

  // ---- thread-local top-k over the flat (row, word) candidates, kept sorted (best first)
  Cand loc[kMaxBeam];
#pragma unroll
  for (int i = 0; i < kMaxBeam; ++i) loc[i] = Cand{-INFINITY, 0x7fffffff};
  for (int j = 0; j < nrows_used; ++j) {
    const float* lg = logits + (int64_t)(row0 + label_at(j)) * vocab;
    const float mx = s_lse_max[j], lg_sum = s_lse_log[j];
    float prev = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxBeam; ++i)
      if (i == j) prev = prev_sum[i];
    for (int v0 = tid; v0 < vocab; v0 += 8 * kBeamThreads) {
      float t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = (v0 + u * kBeamThreads < vocab) ? lg[v0 + u * kBeamThreads] : -INFINITY;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int v = v0 + u * kBeamThreads;
        if (v >= vocab) break;
        const float lsm = (t[u] - mx) - lg_sum;
        const Cand c{step == 0 ? lsm : prev + lsm, j * vocab + v};
        if (better(c, loc[kMaxBeam - 1])) {
          loc[kMaxBeam - 1] = c;
#pragma unroll
          for (int i = kMaxBeam - 1; i > 0; --i)
            if (better(loc[i], loc[i - 1])) {
              const Cand tt = loc[i];
              loc[i] = loc[i - 1];
              loc[i - 1] = tt;
            }
        }
      }
    }
  }
  // ---- k_sel rounds of block arg-max over the heads of the thread-local lists (one barrier per round)
  for (int r = 0; r < k_sel; ++r) {
    Cand best = loc[0];
    int owner = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const Cand other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.idx, o)};
      const int oo = __shfl_xor_sync(0xffffffffu, owner, o);
      if (better(other, best)) {
        best = other;
        owner = oo;
      }
    }
    const int buf = r & 1;
    if (lane == 0) {
      s_cand[buf][warp] = best;
      s_owner[buf][warp] = owner;
    }
    __syncthreads();
    Cand bw = s_cand[buf][0];
    int ow = s_owner[buf][0];
#pragma unroll
    for (int i = 1; i < kBeamThreads / 32; ++i)
      if (better(s_cand[buf][i], bw)) {
        bw = s_cand[buf][i];
        ow = s_owner[buf][i];
      }
    if (bw.idx == 0x7fffffff) bw.idx = 0;  // NaN logits (fully masked clip, len 0): stay memory-safe like a no-op pick
    if (tid == 0) s_win[r] = bw;
    if (tid == ow) {
#pragma unroll
      for (int i = 0; i < kMaxBeam - 1; ++i) loc[i] = loc[i + 1];
      loc[kMaxBeam - 1] = Cand{-INFINITY, 0x7fffffff};
    }
  }
  __syncthreads();

  // ---- bookkeeping: candidate r -> r-th live label (beam.py:165-176), histories via back-pointers
  for (int item = tid; item < k_sel * (step + 2); item += kBeamThreads) {
    const int r = item / (step + 2), p = item - r * (step + 2);
    const int row = row0 + label_at(r);
    const int prev_pos = s_win[r].idx / vocab;
    const int word = s_win[r].idx - prev_pos * vocab;
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
    const int prev_pos = s_win[r].idx / vocab;
    const int word = s_win[r].idx - prev_pos * vocab;
    st.sum_lp[row] = s_win[r].v;
    if (word == kEos || step == max_len - 1) {  // beam.py:173-190
      for (int p = 0; p <= step; ++p)
        st.out_preds[(int64_t)row * max_len + p] = tok_new[(int64_t)row * tstride + p + 1];
      st.out_lp[row] = s_win[r].v / (float)(step + 1);
      st.live[row] = 0;
      const int left = atomicSub(&st.done[2], 1) - 1;
      if (left == 0) {  // every slot of every clip is finished (beam.py:192-194)
        st.done[1] = step + 1;
        __threadfence();
        st.done[0] = 1;
      }
    }
  }
}

// best beam per clip (first arg-max of the average log-prob, beam.py:218-220) + index of its first EOS
__global__ void beam_finalize_kernel(BeamState st, int64_t* __restrict__ best_preds, float* __restrict__ best_lp,
                                     int* __restrict__ best_len, int batch, int beam, int max_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  int best = 0;
  float bv = st.out_lp[b * beam];
  for (int l = 1; l < beam; ++l)
    if (st.out_lp[b * beam + l] > bv) {
      bv = st.out_lp[b * beam + l];
      best = l;
    }
  best_lp[b] = bv;
  int first_eos = max_len;
  for (int p = 0; p < max_len; ++p) {
    const int64_t t = st.out_preds[(int64_t)(b * beam + best) * max_len + p];
    best_preds[(int64_t)b * max_len + p] = t;
    if (t == kEos && first_eos == max_len) first_eos = p;
  }
  best_len[b] = first_eos;
}

int launch_beam_init(const int64_t* bos_ids, BeamState st, const DecoderDims& dd, cudaStream_t stream) {
  beam_init_kernel<<<(dd.rows + 127) / 128, 128, 0, stream>>>(bos_ids, st, dd.rows, dd.beam, dd.max_len);
  CNB_LAUNCH_OK();
  return 0;
}

int launch_beam_step(float* logits, const uint8_t* forbid, BeamState st, int step, int cur, int min_len,
                     const DecoderDims& dd, cudaStream_t stream) {
  CNB_REQUIRE(dd.beam <= kMaxBeam, "beam_size > 8 is not supported");
  CNB_REQUIRE(dd.max_len + 1 <= kBeamThreads, "max_pred_size too large");
  beam_step_kernel<<<dd.rows / dd.beam, kBeamThreads, 0, stream>>>(logits, forbid, st, step, cur, min_len, dd.beam,
                                                                   dd.max_len, dd.vocab);
  CNB_LAUNCH_OK();
  return 0;
}

int launch_beam_finalize(BeamState st, int64_t* best_preds, float* best_lp, int* best_len, const DecoderDims& dd,
                         cudaStream_t stream) {
  const int batch = dd.rows / dd.beam;
  beam_finalize_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(st, best_preds, best_lp, best_len, batch, dd.beam, dd.max_len);
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace cnb
