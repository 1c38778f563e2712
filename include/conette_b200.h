/* conette_b200 -- C ABI of the B200-native CoNeTTE inference hot path.
 *
 * The reference (Labbeti/conette-audio-captioning) is pure Python/PyTorch and has no FFI of its own; the operator seams
 * this library replaces are:
 *   S1  ConvNeXt.forward(input_, input_shapes) -> {frame_embs, frame_embs_lens, clipwise_output}
 *         reference src/conette/nn/encoders/convnext.py:264-336           -> cnb_frontend / cnb_encoder
 *   S2  generate(decoder, pad_id, bos_id, eos_id, vocab_size, frame_embs, frame_embs_pad_mask, beam_size,
 *                min_pred_size, max_pred_size, forbid_rep_mask) -> 4-tuple
 *         reference src/conette/nn/decoding/beam.py:22-227                  -> cnb_decode
 *       (with the projection of pl_modules/conette.py:452-467 + common.py:59-78 folded in front of it)
 *   S3  AACDecoder.__call__ (decoding/common.py:9-29, nn/decoders/aac_tfmer.py:71-118), replaced wholesale by a
 *       KV-cached step; exposed only for stage-isolated parity as cnb_decoder_logits (teacher-forced logits).
 *   S1+S2 behind CoNeTTEModel.forward (huggingface/model.py:185-261)        -> cnb_caption / cnb_caption_host
 *   state-dict tensors by the reference's names       -> cnb_load_weight
 *
 * Conventions: plain C types only; every function returns 0 on success, a negative code on failure and records a
 * message retrievable with cnb_last_error().  Unless a name ends in _host, pointers are DEVICE pointers owned by the
 * caller (e.g. torch tensors) and work is enqueued on the given CUDA stream (a cudaStream_t passed as void*, NULL = legacy
 * default stream) without synchronising.  The library owns only its packed weights and an internal workspace.
 * There is no CPU fallback: every entry point needs a CUDA device of compute capability 10.x.
 */
#ifndef CONETTE_B200_H_
#define CONETTE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNB_ABI_VERSION 1

typedef struct cnb_handle cnb_handle;

typedef struct cnb_config {
  int32_t abi_version;   /* must be CNB_ABI_VERSION */
  int32_t device;        /* CUDA device ordinal */
  int32_t vocab_size;    /* V: rows of decoder.emb_layer.weight / decoder.classifier.weight */
  int32_t precision;     /* 0 = fast: fp16-operand tcgen05 encoder GEMMs (fp32 accumulate, fp32 residual stream, tanh-fit GELU)
                                and the one-launch cluster decoder, whose tensor-core GEMMs run on fp16 hi/lo split
                                operands (22 significand bits, fp32 accumulate: logits within ~2e-5 of fp32)
                            1 = parity: fp32 CUDA-core GEMMs everywhere (encoder and the CUDA-graph decoder) */
  int32_t enc_chunk;     /* clips encoded per pass (0 = library default); bounds workspace, keeps tiles in L2 */
  int32_t reserved[3];   /* [0] = decoder implementation: 0 default (cluster kernel in precision "fast" when the shape
                            fits: beam <= 8, max_len <= 64, T' <= 128, V <= 65535; else the fp32 graph), 1 fp32 CUDA graph,
                            2 fp32 eager launches, 3 cluster kernel or error */
} cnb_config;

enum { CNB_PRECISION_FAST = 0, CNB_PRECISION_PARITY = 1 };
/* Hard limits of the decode entry points (the reference accepts any value; these cover its configurations: beam 2/3/5,
 * max_pred_size 20/30).  Violations return -1 with a message naming the limit. */
#define CNB_MAX_BEAM 8        /* beam_size, and captions per clip of cnb_score_captions */
#define CNB_MAX_PRED_SIZE 64  /* max_pred_size */
enum { CNB_DTYPE_F32 = 0, CNB_DTYPE_I64 = 1, CNB_DTYPE_BOOL = 2, CNB_DTYPE_U8 = 3 };
/* encoder tap points for stage-isolated parity (cnb_encoder_tap): activation returned as fp32, NHWC */
enum { CNB_TAP_LOGMEL_BN = 0, CNB_TAP_STEM = 1, CNB_TAP_BLOCK = 2, CNB_TAP_DOWN = 3, CNB_TAP_DWLN = 4 };

const char* cnb_last_error(void);
int cnb_abi_version(void);

int cnb_create(const cnb_config* cfg, cnb_handle** out);
int cnb_destroy(cnb_handle* h);

/* Stage one tensor of the reference state dict (HOST pointer, contiguous, row-major). Names are the reference's own
 * (e.g. "preprocessor.encoder.stages.0.0.pwconv1.weight", "model.decoder.layers.3.linear1.bias"); tensors the CUDA path
 * does not consume (tokenizer state, task ids, forbid mask, num_batches_tracked) are accepted and ignored. */
int cnb_load_weight(cnb_handle* h, const char* name, const void* host_ptr, int32_t dtype, int32_t ndim,
                    const int64_t* shape);
/* Validate that every required tensor was staged, pack (transpose / fp16 copies / fp16 hi-lo split of the decoder / BN fold / sparse mel) and upload. */
int cnb_finalize_weights(cnb_handle* h);

/* Output geometry for a padded batch of n_samples: STFT frames T, ConvNeXt stage heights, output frames T'. */
int cnb_geometry(int64_t n_samples, int32_t* n_stft_frames, int32_t stage_heights[4], int32_t* n_out_frames);

/* S0 input side: polyphase sinc resampler, wav_in (B, n_in) f32 at orig_freq -> wav_out (B, n_out) f32 at
 * new_freq; replaces torchaudio.functional.resample as called by the reference at
 * src/conette/huggingface/preprocessor.py:139-141 (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99).
 * orig_freq / new_freq are already divided by their gcd.  taps (new_freq, n_taps) f32 = the non-zero support of every phase
 * of the reference's filter bank, tap_lo (new_freq) i32 = first dense column of that support, width = the bank's
 * half-width (dense bank = 2 * width + orig_freq columns; tap_lo[p] + n_taps <= that).  lens_in (B) i64 DEVICE true input
 * lengths or NULL (= n_in): samples at or beyond ceil(new * len / orig) are written as 0, i.e. the result equals
 * per-clip resampling followed by the reference's right zero-padding (nn/functional/pad.py:11-17). */
int cnb_resample(cnb_handle* h, const float* wav_in, const int64_t* lens_in, int32_t batch, int64_t n_in, const float* taps,
                 const int32_t* tap_lo, int32_t orig_freq, int32_t new_freq, int32_t n_taps, int32_t width, float* wav_out,
                 int64_t n_out, void* stream);

/* S1a front-end: wav (B, N) f32 -> log-mel (B, T, 224) f32, optionally through eval BatchNorm (bn0). */
int cnb_frontend(cnb_handle* h, const float* wav, int32_t batch, int64_t n_samples, int32_t apply_bn, float* logmel_out,
                 void* stream);
/* S1 encoder: wav (B, N) f32 -> frame_embs (B, T', 768) f32 [time-major, i.e. the reference's frame_embs transposed as
 * preprocessor.py:64 does] and, if clip_probs_out != NULL, AudioSet tag probabilities (B, 527). */
int cnb_encoder(cnb_handle* h, const float* wav, int32_t batch, int64_t n_samples, float* frame_embs_out,
                float* clip_probs_out, void* stream);
/* Debug/parity: run the encoder on (B <= enc_chunk) clips and copy one intermediate activation out (fp32, NHWC). */
int cnb_encoder_tap(cnb_handle* h, const float* wav, int32_t batch, int64_t n_samples, int32_t tap_kind, int32_t stage,
                    int32_t block, float* out, int64_t out_capacity_elems, void* stream);

/* S2 projection + beam search: frame_embs (B, T', 768) f32, lens (B) i32 valid frames, bos_ids (B) i64 task BOS token,
 * forbid_mask (V) u8 or NULL.  Outputs (device): preds (B, max_len) i64 best beam padded with 0; lprobs (B) f32;
 * mult_preds (B, beam, max_len) i64; mult_lprobs (B, beam) f32; info (2 + B) i32 = {pred_size, reserved,
 * first-EOS index of the best beam per clip (max_len if none)}.  The caller trims mult_preds to pred_size and preds to
 * max(first-EOS)+1 exactly as beam.py:205-225 does. */
int cnb_decode(cnb_handle* h, const float* frame_embs, const int32_t* lens, const int64_t* bos_ids,
               const uint8_t* forbid_mask, int32_t batch, int32_t n_frames, int32_t beam, int32_t min_len, int32_t max_len,
               int64_t* preds_out, float* lprobs_out, int64_t* mult_preds_out, float* mult_lprobs_out, int32_t* info_out,
               void* stream);
/* cnb_decode through the cluster kernel (error if the shape does not fit it) with the kernel's per-step logits tapped:
 * logits_out (max_len, B*beam, V) f32 = what the reference's decoder call returns for the last position at every step
 * (AACDecoder.__call__, src/conette/nn/decoding/common.py:9-29; beam.py:113-127), before the EOS / no-repeat masks.  Rows are
 * the fixed beam slots (clip * beam + label); steps after the early exit are left untouched.  Test hook. */
int cnb_decode_tap(cnb_handle* h, const float* frame_embs, const int32_t* lens, const int64_t* bos_ids,
                   const uint8_t* forbid_mask, int32_t batch, int32_t n_frames, int32_t beam, int32_t min_len, int32_t max_len,
                   int64_t* preds_out, float* lprobs_out, int64_t* mult_preds_out, float* mult_lprobs_out, int32_t* info_out,
                   float* logits_out, void* stream);
/* S3 (fp32 step kernels): teacher-forced decoder logits. tokens (B, steps) i64 -> logits (B, steps, V) f32. */
int cnb_decoder_logits(cnb_handle* h, const float* frame_embs, const int32_t* lens, const int64_t* tokens, int32_t batch,
                       int32_t n_frames, int32_t steps, float* logits_out, void* stream);
/* Teacher-forced scoring of given captions: replaces the loss loop of CoNeTTEPLM.test_step /
 * validation_step, reference src/conette/pl_modules/conette.py:293-318 (decode_audio(..., "forcing", caps_in=caps[:, :-1]),
 * nn/decoding/forcing.py:12-76) + CrossEntropyLossMean(ignore_index=pad_id, dim=1) (nn/modules/ce_mean.py:10-40).
 * captions (B, n_caps, cap_len) i64, position 0 = the clip's task BOS id, right-padded with pad_id 0.  Outputs:
 * token_lprobs (B, n_caps, cap_len-1) f32 = log-softmax of position p evaluated at token p+1 (0 where that token is pad);
 * losses (B, n_caps) f32 = -mean of token_lprobs over the non-pad targets.  The (B, cap_len-1, V) logits are never
 * written to HBM beyond one (B*n_caps, V) step buffer; cross-attention K|V are computed once per clip. */
int cnb_score_captions(cnb_handle* h, const float* frame_embs, const int32_t* lens, const int64_t* captions, int32_t batch,
                       int32_t n_frames, int32_t n_caps, int32_t cap_len, float* token_lprobs_out, float* losses_out,
                       void* stream);

/* S1+S2: waveform -> token ids, all buffers on the device. x_lens_host (B) i64 = true sample counts (HOST; NULL = all
 * n_samples) from which frame lens = round_half_even(len / (N // T')) are derived (convnext.py:312-315).
 * frame_embs_out / clip_probs_out may be NULL. */
int cnb_caption(cnb_handle* h, const float* wav, const int64_t* x_lens_host, const int64_t* bos_ids,
                const uint8_t* forbid_mask, int32_t batch, int64_t n_samples, int32_t beam, int32_t min_len, int32_t max_len,
                int64_t* preds_out, float* lprobs_out, int64_t* mult_preds_out, float* mult_lprobs_out, int32_t* info_out,
                float* clip_probs_out, void* stream);
/* Same with HOST buffers for everything: host->device copy of the waveforms, compute, device->host copy of the results
 * and a stream synchronise all happen inside the call (this is the end-to-end entry the Python wrapper uses). */
int cnb_caption_host(cnb_handle* h, const float* wav_host, const int64_t* x_lens_host, const int64_t* bos_ids_host,
                     const uint8_t* forbid_mask_host, int32_t batch, int64_t n_samples, int32_t beam, int32_t min_len,
                     int32_t max_len, int64_t* preds_out_host, float* lprobs_out_host, int64_t* mult_preds_out_host,
                     float* mult_lprobs_out_host, int32_t* info_out_host, float* clip_probs_out_host);
/* Split-phase form of cnb_caption_host for callers that stream batches (dataset captioning as conette-predict does over a
 * file list, serving): _begin enqueues the H2D copies, the whole path and the D2H copies and returns a ticket, _end blocks
 * until that batch's outputs are in its host buffers.  Two batches may be in flight, so the H2D copy of batch i+1 overlaps
 * the compute of batch i, and (fast precision) batch i decodes on a high-priority stream while batch i+1 is being encoded;
 * results are identical to cnb_caption_host.  wav_host may also be DEVICE memory (an already resident batch): it is then used
 * in place.  Input and output buffers must stay valid and untouched (host ones pinned, for real overlap) until _end; calling
 * _begin a third time without _end first waits for the oldest batch. */
int cnb_caption_host_begin(cnb_handle* h, const float* wav_host, const int64_t* x_lens_host, const int64_t* bos_ids_host,
                           const uint8_t* forbid_mask_host, int32_t batch, int64_t n_samples, int32_t beam, int32_t min_len,
                           int32_t max_len, int64_t* preds_host, float* lprobs_host, int64_t* mult_preds_host,
                           float* mult_lprobs_host, int32_t* info_host, float* clip_probs_host, int32_t* ticket_out);
int cnb_caption_host_end(cnb_handle* h, int32_t ticket);

/* Test hook: one GEMM with a fused epilogue, out (M,N) f32 = epi(A (M,K) f32 x W (N,K) f32 ^T). epi: 0 bias, 1 bias+GELU,
 * 2 bias+ReLU, 3 resid + scale*(acc+bias).  use_tc=1 runs the fp16-operand tcgen05 kernel (operands rounded to fp16 first;
 * out_bf16=1 also rounds the result through the 16-bit (fp16) epilogue), use_tc=0 the fp32 CUDA-core kernel. */
int cnb_debug_gemm(cnb_handle* h, const float* a, const float* w, const float* bias, const float* scale, const float* resid,
                   int32_t m, int32_t n, int32_t k, int32_t epi, int32_t use_tc, int32_t out_bf16, float* out, void* stream);

/* Test hook: the fused pointwise MLP of a ConvNeXt stage-1 block (reference convnext.py:66-73, C = 96):
 * x (M,96) f32 += scale * (W2 (96,384) . GELU(W1 (384,96) . y (M,96) + b1) + b2), in place, through the kernel the fast
 * precision mode uses (y, W1, W2 and the hidden activations are rounded to fp16, fp32 accumulation and residual). */
int cnb_debug_mlp_fused(cnb_handle* h, const float* y, const float* w1, const float* b1, const float* w2, const float* b2,
                        const float* scale, float* x, int32_t m, void* stream);
/* Test hook: the same for the fused stage-2 / stage-3 MLP kernel (C = 192 / 384, hidden 4C; CTA pairs, streamed weights):
 * y (M, C), w1 (4C, C), w2 (C, 4C), x (M, C) updated in place. */
int cnb_debug_mlp_fused_pair(cnb_handle* h, int32_t c, const float* y, const float* w1, const float* b1, const float* w2,
                             const float* b2, const float* scale, float* x, int32_t m, void* stream);

/* Per-kernel-class device timing: between cnb_profile_begin and cnb_profile_end every launch group issued through this
 * handle is bracketed by a CUDA event pair on the launching stream; _end synchronises and returns, per class, the summed
 * event time in ms and the number of brackets. Arrays must hold CNB_K_COUNT entries.  The block kernels are reported
 * per ConvNeXt stage (CNB_K_*_S0 + stage); the aggregate classes CNB_K_DWLN / GEMM_PW1 / GEMM_PW2 stay zero. */
enum {
  CNB_K_FRONTEND = 0, CNB_K_STEM = 1, CNB_K_DWLN = 2, CNB_K_GEMM_PW1 = 3, CNB_K_GEMM_PW2 = 4, CNB_K_DS_PACK = 5,
  CNB_K_DS_GEMM = 6, CNB_K_HEAD = 7, CNB_K_PROJ_KV = 8, CNB_K_DEC_GEMM = 9, CNB_K_DEC_ATTN = 10, CNB_K_DEC_CLS = 11,
  CNB_K_BEAM = 12,
  /* stage-resolved classes of the three ConvNeXt block kernels: base + stage (0..3) */
  CNB_K_DWLN_S0 = 13, CNB_K_GEMM_PW1_S0 = 17, CNB_K_GEMM_PW2_S0 = 21, CNB_K_COUNT = 25
};
int cnb_profile_begin(cnb_handle* h);
int cnb_profile_end(cnb_handle* h, float* ms_per_class, int64_t* brackets_per_class, int32_t n_classes);
/* Timeline form of the same brackets: unlike cnb_profile_begin it leaves the streaming overlap of cnb_caption_host_begin on
 * (decode of batch i on its own stream next to the encoder of batch i+1), and _end returns every bracket in issue order:
 * class, begin and end in ms after the first bracket's begin.  At most `cap` entries are written, *n_out = brackets recorded. */
int cnb_profile_timeline_begin(cnb_handle* h);
int cnb_profile_timeline_end(cnb_handle* h, int32_t* cls, float* t_begin_ms, float* t_end_ms, int32_t cap, int32_t* n_out);

/* Number of kernels launched by this handle since creation (bench.py reports the per-step delta as gpu_launches). */
int64_t cnb_launch_count(const cnb_handle* h);
/* Bytes of device memory currently held (weights + workspace). */
int64_t cnb_device_bytes(const cnb_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* CONETTE_B200_H_ */
