// Internal launcher declarations shared by the translation units of libconette_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace cnb {

// ---- front-end -------------------------------------------------------------------------------------------------
constexpr int kMelSchedMax = 48;      // entries per lane of the balanced mel schedule (CoNeTTE's bank needs 29)
struct FrontendParams {
  const float2* twiddle = nullptr;  // (1024) W_1024^j
  const int* mel_lo = nullptr;      // (224)
  const int* mel_cnt = nullptr;     // (224)
  const int* mel_off = nullptr;     // (224)
  const float* mel_w = nullptr;     // (nnz)
  const float2* mel_sched = nullptr; // (mel_sched_len, 32): lane-balanced schedule of the same non-zeros, see frontend.cu
  int mel_sched_len = 0;
  const float* bn_scale = nullptr;  // (224)
  const float* bn_shift = nullptr;  // (224)
  const float* ones = nullptr;      // (224)
  const float* zeros = nullptr;     // (224)
};
int launch_frontend(const float* wav, int batch, int64_t n_samples, const FrontendParams& p, bool apply_bn, float* out,
                    cudaStream_t stream);

// ---- polyphase resampler (input side of the boundary) ----------------------------------------------
// taps (new, n_taps) compact filter bank, lo (new) first dense column of every phase; lens_in device (B) or null
int launch_resample(const float* x, int batch, int64_t n_in, const int64_t* lens_in, const float* taps, const int* lo, int orig,
                    int nw, int n_taps, int width, float* out, int64_t n_out, cudaStream_t stream);

// ---- stem: Conv2d(1,96,4x4,s4,pad(4,0)) + LayerNorm(channels_first) -> NHWC f32 --------------------------------------
int launch_stem(const float* logmel_bn, int batch, int n_frames, int h1, const float* w_t /*(16,96)*/, const float* bias,
                const float* ln_g, const float* ln_b, float* out, cudaStream_t stream);

// ---- depthwise 7x7 + LayerNorm(C) -> (M, C) in OutT ---------------------------------------------------------------------
template <typename OutT>
int launch_dwconv_ln(const float* x, int batch, int h, int w, int c, const float* w_t /*(49,C)*/, const float* bias,
                     const float* ln_g, const float* ln_b, OutT* out, cudaStream_t stream);

// ---- LayerNorm(channels_first) + 2x2/s2 im2col pack: (B,H,W,C) f32 -> (B*(H/2)*(W/2), 4C) in OutT, K order (kh,kw,c) --
template <typename OutT>
int launch_ln_pack2x2(const float* x, int batch, int h, int w, int c, const float* ln_g, const float* ln_b, OutT* out,
                      cudaStream_t stream);

// ---- mean over the 7 frequency columns: (B,T',7,768) f32 -> (B,T',768) f32 ------------------------------------------
int launch_freq_mean(const float* x, int batch, int tp, int w, int c, float* out, cudaStream_t stream);

// ---- clip head: max_t + mean_t -> LN(768) -> Linear 527 -> sigmoid -------------------------------------------------
int launch_clip_head(const float* frame_embs, int batch, int tp, const float* ln_g, const float* ln_b, const float* w,
                     const float* bias, int n_cls, float* out, cudaStream_t stream);

// ---- GEMMs: out[m,n] = epilogue(sum_k A[m,k] * W[n,k]) ---------------------------------------------------------------
enum Epilogue : int {
  EPI_BIAS = 0,         // acc + bias[n]
  EPI_BIAS_GELU = 1,    // gelu_erf(acc + bias[n])
  EPI_BIAS_RELU = 2,    // relu(acc + bias[n])
  EPI_SCALE_RESID = 3,  // resid[m,n] + scale[n] * (acc + bias[n])      (layer-scale + residual; in-place allowed)
};
struct EpiParams {
  const float* bias = nullptr;
  const float* scale = nullptr;
  const float* resid = nullptr;  // (M, N) f32, row stride = ldo
};
// fp32 SIMT GEMM (parity mode + decoder).  A (M,K) f32 row stride lda; W (N,K) f32; out (M,N) OutT row stride ldo.
// splits > 1 = split-K: slice z writes raw partial sums to out + z*M*ldo and the epilogue is left to the consumer.
template <typename OutT>
int launch_gemm_f32(const float* a, int64_t lda, const float* w, int m, int n, int k, Epilogue epi, const EpiParams& ep,
                    OutT* out, int64_t ldo, cudaStream_t stream, int splits = 1);

// Latency-optimised fp32 GEMM for the decoder (whole 256-wide K panel in shared memory, one L2 round trip).  K > 256 is
// split into ceil(K/256) slices writing raw partial sums to out + z*M*ldo (epilogue left to the consumer).
int launch_gemm_f32_panel(const float* a, int64_t lda, const float* w, int m, int n, int k, Epilogue epi, const EpiParams& ep,
                          float* out, int64_t ldo, cudaStream_t stream);

// fp16-operand tcgen05 GEMM (fast mode).  A (M,K) fp16 contiguous, W (N,K) fp16 contiguous (both K-major, TMA-fed).
struct TcGemmPlan;  // opaque: TMA descriptors + tile configuration
template <typename OutT>
int launch_gemm_tc(const act16* a, const act16* w, int m, int n, int k, Epilogue epi, const EpiParams& ep,
                   OutT* out, int64_t ldo, cudaStream_t stream);
// fused pointwise MLP of ConvNeXt stage 1 (mlp_fused.cu): x (M, 96) f32 += scale * (W2 . GELU(W1 . y + b1) + b2), in place
int launch_mlp_fused_c96(const act16* y, const act16* w1, const act16* w2, const float* b1, const float* b2,
                         const float* scale, float* x, int m, cudaStream_t stream);
// fused pointwise MLP of ConvNeXt stages 2 / 3 (mlp_fused_pair.cu, C = 192 / 384): x (M, C) f32 += scale * (W2 . GELU(W1 . y + b1) + b2)
int launch_mlp_fused_pair(int c, const act16* y, const act16* w1, const act16* w2, const float* b1, const float* b2,
                          const float* scale, float* x, int m, cudaStream_t stream);
int gemm_tc_init();  // resolves cuTensorMapEncodeTiled; returns 0 on success
// 2-D row-major (rows, cols) tensor map into map_out (a 128-byte CUtensorMap): box = (box_rows, 128 bytes), 128B swizzle
int tc_make_map(void* map_out, const void* ptr, int64_t rows, int64_t cols, int box_rows, int elt_bytes);

// 2-D row-major fp16 tensor map with an explicit box: box_cols 64 -> 128B swizzle, 32 -> 64B swizzle
int tc_make_map_f16_box(void* map_out, const void* ptr, int64_t rows, int64_t cols, int box_rows, int box_cols);
// 4-D NHWC fp32 tensor map, box (box_c, box_w, 1, 1), no swizzle, zero OOB fill (used by the depthwise-conv ring loader)
int tc_make_map_nhwc_f32(void* map_out, const float* ptr, int batch, int h, int w, int c, int box_c, int box_w);
// TMA-ring depthwise 7x7 + LayerNorm (dwconv_ring.cu); returns 1 when (C, W) has no instantiation (caller falls back)
template <typename OutT>
int launch_dwconv_ln_tma(const float* x, int batch, int h, int w, int c, const float* w_t, const float* bias,
                         const float* ln_g, const float* ln_b, OutT* out, cudaStream_t stream);

// the encoder about to be enqueued overlaps the previous batch's decoder (streaming host API): prefer half-size CTAs
void dwconv_set_overlap_hint(bool on);
// SMs the persistent encoder kernels (gemm_tc, mlp_fused, mlp_fused_pair, dwconv_ring) size their grids for: 148 unless the
// streaming host API knows that a decoder holds some of them (a 148-CTA persistent grid on fewer free SMs runs as two waves)
void set_sm_budget(int sms);
int sm_budget();

// ---- decoder / beam search ------------------------------------------------------------------------------------------------
struct DecoderDims {
  int rows;      // B * beam
  int beam;
  int tp;        // encoder frames T'
  int max_len;   // max_pred_size
  int vocab;
};
struct BeamState {
  int* tokens[2];      // (R, max_len+1) ping-pong token histories (position 0 = task BOS id)
  int* src_row[2];     // (R, max_len) ping-pong: physical row whose KV cache holds position p of this row's history
  float* sum_lp;       // (R)
  uint8_t* live;       // (R)
  int64_t* out_preds;  // (R, max_len)
  float* out_lp;       // (R)
  int* done;           // [0] all-finished flag, [1] pred_size, [2] live rows remaining
};
// every step kernel returns immediately once done[0] is set (device-side early exit, no host sync)
int launch_embed(const int* tokens, int pos, const float* emb, const float* pe, float* x, const DecoderDims& dd,
                 const int* done, cudaStream_t stream);
int launch_self_attn(const float* qkv, float* kcache, float* vcache, const int* src_row, int pos, float* attn,
                     const DecoderDims& dd, const int* done, cudaStream_t stream);
int launch_cross_attn(const float* q, const float* ck, const float* cv, int64_t kv_stride, const int* lens, float* attn,
                      const DecoderDims& dd, const int* done, cudaStream_t stream);
// x = LayerNorm(x + bias + sum_s delta[s]) (eps 1e-5), in place; delta = nsplit slabs of (rows, 256), bias may be null
int launch_add_ln(float* x, const float* delta, int nsplit, const float* bias, const float* g, const float* b, int rows,
                  const int* done, cudaStream_t stream);
int launch_beam_init(const int64_t* bos_ids, BeamState st, const DecoderDims& dd, cudaStream_t stream);
int launch_beam_step(float* logits, const uint8_t* forbid, BeamState st, int step, int cur, int min_len,
                     const DecoderDims& dd, cudaStream_t stream);
int launch_beam_finalize(BeamState st, int64_t* best_preds, float* best_lp, int* best_len, const DecoderDims& dd,
                         cudaStream_t stream);

// ---- cluster decoder: the whole beam-search decode in ONE launch (decoder_cluster.cu) -------------------------------------
// Decoder weights as the tensor cores take them there: every nn.Linear matrix W is split into two fp16 matrices
// W1 = fp16(W), W2 = fp16((W - W1) * 2048) (22 significand bits together), each behind its own TMA descriptor.
constexpr int kDecMapsPerLayer = 12;  // {sa_in, sa_out (head-packed), ca_q, ca_out (head-packed), l1, l2} x {W1, W2}
constexpr int kDecMaps = 6 * kDecMapsPerLayer + 2;  // + classifier {W1, W2}
// L2 eviction-priority policies for the .L2::cache_hint forms of the TMA loads (the encodings createpolicy.fractional.L2::evict_*
// produces at fraction 1.0)
constexpr unsigned long long kL2EvictNormal = 0x1000000000000000ull, kL2EvictFirst = 0x12F0000000000000ull,
                             kL2EvictLast = 0x14F0000000000000ull;

struct ClusterLayer {
  const float *sa_in_b, *sa_out_b, *ca_q_b, *ca_out_b, *l1_b, *l2_b, *n1_g, *n1_b, *n2_g, *n2_b, *n3_g, *n3_b;
};
struct ClusterArgs {
  ClusterLayer layers[6];
  const float *emb, *pe, *cls_b;
  const void* tmaps;         // kDecMaps CUtensorMaps (device): layer l -> [12 l + 2 j + half], classifier at 72 + half
  const float* ckv;          // (B*T', 6*512) cross-attention K|V of all layers
  const float* ckt;          // cross-attention keys transposed per clip: [B][6 layers][8 heads][32 dims][tpad]
  int tpad;                  // T' rounded up to a multiple of 32 (128-byte rows)
  const int* lens;
  const int64_t* bos_ids;
  const uint8_t* forbid;
  float *kc, *vc;            // self-attention caches (6, rows, max_len, 256)
  BeamState bs;              // out_preds / out_lp / done are used
  float* tap;                // optional (tests): raw logits of every step, (max_len, rows, vocab)
  unsigned long long* trace; // optional (debug): phase times of the first CTA
  unsigned long long w_policy; // L2 eviction policy of the weight-ring loads (kL2Evict*; 0 = no hint)
  unsigned long long kv_policy; // ... of the attention data (K/V caches, cross-attention K/V): always a valid policy
  int rows, beam, tp, max_len, vocab, min_len, batch;
  int compact;               // 1 = prefer few, fat clusters (32 rows each): the decode shares the GPU with the next encoder
};
bool decoder_cluster_supported(const ClusterArgs& args);
int launch_decoder_cluster(const ClusterArgs& args, cudaStream_t stream);
// SMs the launch for `args` will occupy (8 per cluster); 0 when unsupported
int decoder_cluster_sms(const ClusterArgs& args);

// global launch counter (reported through cnb_launch_count)
void count_launch();

}  // namespace cnb
