// K-FE: fused log-mel front-end.
//   waveform (B, N) f32 -> reflect-pad 512 -> frames (hop 320, win 1024, periodic Hann) -> real FFT-1024 -> power
//   -> sparse Slaney mel (513 -> 224) -> 10*log10(clamp(., 1e-10)) -> eval BatchNorm per mel bin -> (B, T, 224) f32
// Replaces torchlibrosa Spectrogram + LogmelFilterBank as called at reference nn/encoders/convnext.py:160-180,276-278
// and bn0 at :290-292.  The reference evaluates the DFT as two dense Conv1d (2.1 GFLOP per
// 10 s clip) and writes the (B,T,513) power tensor to HBM; here two frames are packed into one complex radix-4 FFT
// in shared memory and only the waveform is read / the normalised log-mel written (algorithmic bytes: 4N + 4*T*224).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace cnb {

constexpr int kFftN = 1024;
constexpr int kHop = 320;
constexpr int kBins = 513;
constexpr int kMels = 224;
constexpr int kFramesPerCta = 8;
constexpr int kSpan = (kFramesPerCta - 1) * kHop + kFftN;  // 3264 samples feed 8 frames
constexpr int kFeThreads = 256;

__device__ __forceinline__ int digit_reverse4x5(int k) {
  // base-4 digit reversal of a 10-bit index (radix-4 DIF leaves X[k] at position rev(k))
  int r = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    r = (r << 2) | (k & 3);
    k >>= 2;
  }
  return r;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void __launch_bounds__(kFeThreads)
frontend_kernel(const float* __restrict__ wav, int64_t n_samples, int n_frames,
                const float2* __restrict__ twiddle,     // W_1024^j = (cos, -sin)(2*pi*j/1024), j < 1024
                const int* __restrict__ mel_lo,         // first FFT bin of each mel filter
                const int* __restrict__ mel_cnt,        // number of bins
                const int* __restrict__ mel_off,        // offset into mel_w
                const float* __restrict__ mel_w,        // packed non-zero filter weights
                const float* __restrict__ bn_scale,     // gamma / sqrt(var + eps)   (or 1)
                const float* __restrict__ bn_shift,     // beta - mean * scale        (or 0)
                float* __restrict__ out) {              // (B, T, 224)
  __shared__ float s_wav[kSpan];
  __shared__ float2 s_z[kFftN];
  __shared__ float2 s_tw[kFftN];
  __shared__ float s_pow[2][kBins + 3];

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFramesPerCta;
  const float* x = wav + (int64_t)b * n_samples;

  for (int i = tid; i < kFftN; i += kFeThreads) s_tw[i] = twiddle[i];
  // samples of the reflect-padded row: padded index p = t0*320 + i  <->  original index p - 512
  const int64_t start = (int64_t)t0 * kHop - kFftN / 2;
  for (int i = tid; i < kSpan; i += kFeThreads) {
    int64_t idx = start + i;
    if (idx < 0) idx = -idx;                                   // reflect without repeating the edge sample
    if (idx >= n_samples) idx = 2 * (n_samples - 1) - idx;
    float v = 0.f;
    if (idx >= 0 && idx < n_samples) v = __ldg(x + idx);       // frames beyond T never get stored
    s_wav[i] = v;
  }
  __syncthreads();

  for (int pair = 0; pair < kFramesPerCta / 2; ++pair) {
    const int fa = 2 * pair, fb = 2 * pair + 1;
    // z[n] = hann[n] * (frame_a[n] + i * frame_b[n])
    for (int n = tid; n < kFftN; n += kFeThreads) {
      const float w = 0.5f - 0.5f * s_tw[n].x;
      s_z[n] = make_float2(w * s_wav[fa * kHop + n], w * s_wav[fb * kHop + n]);
    }
    __syncthreads();
    // radix-4 decimation-in-frequency, in place, one butterfly per thread per stage
#pragma unroll
    for (int stage = 0; stage < 5; ++stage) {
      const int L = kFftN >> (2 * stage);
      const int q = L >> 2;
      const int g = tid / q, pos = tid - g * q;
      const int base = g * L + pos;
      const float2 a0 = s_z[base], a1 = s_z[base + q], a2 = s_z[base + 2 * q], a3 = s_z[base + 3 * q];
      const float2 t0c = make_float2(a0.x + a2.x, a0.y + a2.y);
      const float2 t1c = make_float2(a0.x - a2.x, a0.y - a2.y);
      const float2 t2c = make_float2(a1.x + a3.x, a1.y + a3.y);
      const float2 t3c = make_float2(a1.y - a3.y, -(a1.x - a3.x));  // -i * (a1 - a3)
      float2 y0 = make_float2(t0c.x + t2c.x, t0c.y + t2c.y);
      float2 y1 = make_float2(t1c.x + t3c.x, t1c.y + t3c.y);
      float2 y2 = make_float2(t0c.x - t2c.x, t0c.y - t2c.y);
      float2 y3 = make_float2(t1c.x - t3c.x, t1c.y - t3c.y);
      if (stage < 4) {
        const int s = kFftN / L;
        y1 = cmul(y1, s_tw[(pos * s) & (kFftN - 1)]);
        y2 = cmul(y2, s_tw[(2 * pos * s) & (kFftN - 1)]);
        y3 = cmul(y3, s_tw[(3 * pos * s) & (kFftN - 1)]);
      }
      s_z[base] = y0;
      s_z[base + q] = y1;
      s_z[base + 2 * q] = y2;
      s_z[base + 3 * q] = y3;
      __syncthreads();
    }
    // split the packed spectrum: Xa = (Z[k] + conj Z[N-k]) / 2, Xb = (Z[k] - conj Z[N-k]) / 2i ; power = |X|^2
    for (int k = tid; k < kBins; k += kFeThreads) {
      const float2 zk = s_z[digit_reverse4x5(k)];
      const float2 zn = s_z[digit_reverse4x5((kFftN - k) & (kFftN - 1))];
      const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
      const float br = 0.5f * (zk.y + zn.y), bi = 0.5f * (zk.x - zn.x);
      s_pow[0][k] = ar * ar + ai * ai;
      s_pow[1][k] = br * br + bi * bi;
    }
    __syncthreads();
    // sparse mel + dB + BN; 2 frames x 224 mels
    for (int item = tid; item < 2 * kMels; item += kFeThreads) {
      const int f = item / kMels, m = item - f * kMels;
      const int t = t0 + 2 * pair + f;
      if (t < n_frames) {
        const int lo = mel_lo[m], cnt = mel_cnt[m];
        const float* w = mel_w + mel_off[m];
        float acc = 0.f;
        for (int i = 0; i < cnt; ++i) acc = fmaf(s_pow[f][lo + i], w[i], acc);
        const float db = 10.0f * log10f(fmaxf(acc, 1e-10f));
        out[((int64_t)b * n_frames + t) * kMels + m] = db * bn_scale[m] + bn_shift[m];
      }
    }
    __syncthreads();
  }
}

// =====================================================================================================================
// K-FE v2: one warp per frame pair, 1024-point complex FFT as 32 x 32 with both 32-point passes done in registers.
//   CTA = 8 warps x 2 pairs = 32 consecutive frames of one clip, whose 10 944 reflect-padded samples are staged once in
//   shared memory (hop 320 / window 1024 => 3.2x reuse).  Per pair: lane n2 loads z[32 n1 + n2] (conflict-free), runs a
//   radix-2 DIF 32-point FFT over n1 with compile-time twiddles, multiplies by W_1024^{n2 k1} (running power of W^{n2}),
//   transposes through a padded 32 x 33 warp-private tile, runs the second 32-point FFT over n2 and leaves X[k1 + 32 k2]
//   in shared memory; power of the two packed real spectra, sparse mel, dB and BatchNorm follow.  No block barrier after
//   the initial staging, no bank conflicts in the FFT, ~1100 warp instructions per frame (v1: ~6x more + 9 barriers).
// =====================================================================================================================
constexpr int kFe2Warps = 6;                                          // x 2 CTAs per SM = 12 warps, prologues overlap
constexpr int kFe2PairsPerWarp = 2;
constexpr int kFe2Frames = kFe2Warps * kFe2PairsPerWarp * 2;           // 32 frames per CTA
constexpr int kFe2Span = (kFe2Frames - 1) * kHop + kFftN;              // 10 944 samples
constexpr int kFe2WarpScratch = 32 * 33 * 2;                           // transpose / spectrum tiles; power spectra and mel sums reuse them

__device__ __forceinline__ constexpr int bitrev5(int k) {
  return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

// in-register radix-2 DIF FFT of 32 complex values; output X[k] ends up in element bitrev5(k)
__device__ __forceinline__ void fft32_regs(float (&re)[32], float (&im)[32]) {
  constexpr float kC[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                            0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f, 0.0f, -0.19509032201612825f,
                            -0.38268343236508978f, -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f,
                            -0.92387953251128674f, -0.98078528040323043f};
  constexpr float kS[16] = {0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f, 0.70710678118654752f,
                            0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f, 1.0f, 0.98078528040323043f,
                            0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
                            0.38268343236508978f, 0.19509032201612825f};
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
#pragma unroll
    for (int g = 0; g < 32; g += 2 * half) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const int i0 = g + j, i1 = g + j + half;
        const int m = j * (16 / half);  // twiddle W_32^m = (cos, -sin)(2 pi m / 32)
        const float ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
        re[i0] = ar + br;
        im[i0] = ai + bi;
        const float dr = ar - br, di = ai - bi;
        if (m == 0) {
          re[i1] = dr;
          im[i1] = di;
        } else if (m == 8) {  // times -i
          re[i1] = di;
          im[i1] = -dr;
        } else {              // (dr + i di) * (c - i s)
          re[i1] = dr * kC[m] + di * kS[m];
          im[i1] = di * kC[m] - dr * kS[m];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kFe2Warps * 32, 2)
frontend_warpfft_kernel(const float* __restrict__ wav, int64_t n_samples, int n_frames, const float2* __restrict__ twiddle,
                        const int* __restrict__ mel_lo, const int* __restrict__ mel_cnt, const int* __restrict__ mel_off,
                        const float* __restrict__ mel_w, const float2* __restrict__ mel_sched, int sched_len,
                        const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_fe[];
  float* s_wav = s_fe;                       // [kFe2Span]
  float* s_win = s_wav + kFe2Span;           // [1024] periodic Hann
  float2* s_tw = reinterpret_cast<float2*>(s_win + kFftN);  // [32] W_1024^{n2}
  float2* s_sched = s_tw + 32;               // [sched_len][32] lane-balanced mel schedule (weight, bin | end | filter)
  float* s_scr = reinterpret_cast<float*>(s_sched + sched_len * 32);  // per-warp scratch
  __shared__ __align__(8) unsigned long long s_bar;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFe2Frames;
  const float* x = wav + (int64_t)b * n_samples;
  const int64_t start = (int64_t)t0 * kHop - kFftN / 2;
  // Staging.  The span of an interior CTA is contiguous and 16-byte aligned in global memory (hop 320, 32 frames per CTA):
  // one bulk async copy (plus one for the mel schedule) instead of 43 dependent load/store round trips per thread -- with one
  // CTA per SM nothing else hides that latency, and the ncu source view had 40 % of the kernel's stall samples here.
  const bool interior = start >= 0 && start + kFe2Span <= n_samples && (reinterpret_cast<uintptr_t>(x + start) & 15) == 0;
  const uint32_t bar = smem_u32(&s_bar);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async_smem();
    const uint32_t sched_bytes = (uint32_t)sched_len * 32u * 8u;
    mbar_expect_tx(bar, sched_bytes + (interior ? (uint32_t)kFe2Span * 4u : 0u));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s_sched)),
                 "l"(mel_sched), "r"(sched_bytes), "r"(bar)
                 : "memory");
    if (interior)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s_wav)),
                   "l"(x + start), "r"((uint32_t)kFe2Span * 4u), "r"(bar)
                   : "memory");
  }
  if (!interior) {
    for (int i = tid; i < kFe2Span; i += kFe2Warps * 32) {
      int64_t idx = start + i;
      if (idx < 0) idx = -idx;                                   // reflect without repeating the edge sample
      if (idx >= n_samples) idx = 2 * (n_samples - 1) - idx;
      s_wav[i] = (idx >= 0 && idx < n_samples) ? __ldg(x + idx) : 0.f;
    }
  }
  for (int i = tid; i < kFftN; i += kFe2Warps * 32) s_win[i] = 0.5f - 0.5f * twiddle[i].x;
  if (tid < 32) s_tw[tid] = twiddle[tid];
  __syncthreads();
  mbar_wait(bar, 0);

  float* tile_re = s_scr + warp * kFe2WarpScratch;   // [32][33]   (later: spectrum re[1024])
  float* tile_im = tile_re + 32 * 33;                // [32][33]   (later: spectrum im[1024])
  const float2 w1 = s_tw[lane];

  for (int pp = 0; pp < kFe2PairsPerWarp; ++pp) {
    const int fa = (warp * kFe2PairsPerWarp + pp) * 2;  // frame index inside the CTA
    if (t0 + fa >= n_frames) break;                     // warp-uniform
    float re[32], im[32];
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
      const int n = 32 * n1 + lane;
      const float w = s_win[n];
      re[n1] = w * s_wav[fa * kHop + n];
      im[n1] = w * s_wav[(fa + 1) * kHop + n];
    }
    fft32_regs(re, im);
    // twiddle W_1024^{lane * k1} by running powers, then transpose: tile[k1][lane]
    {
      float cr = 1.f, ci = 0.f;
#pragma unroll
      for (int k1 = 0; k1 < 32; ++k1) {
        const int r = bitrev5(k1);
        const float yr = re[r] * cr - im[r] * ci, yi = re[r] * ci + im[r] * cr;
        tile_re[k1 * 33 + lane] = yr;
        tile_im[k1 * 33 + lane] = yi;
        const float nr = cr * w1.x - ci * w1.y, ni = cr * w1.y + ci * w1.x;
        cr = nr;
        ci = ni;
      }
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) {
      re[n2] = tile_re[lane * 33 + n2];
      im[n2] = tile_im[lane * 33 + n2];
    }
    __syncwarp();
    fft32_regs(re, im);
    // X[k1 + 32 k2] (k1 = lane) -> spectrum arrays (reuse the tile storage: 1024 <= 32*33)
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
      tile_re[32 * k2 + lane] = re[bitrev5(k2)];
      tile_im[32 * k2 + lane] = im[bitrev5(k2)];
    }
    __syncwarp();
    // split the packed spectrum: Xa = (Z[k] + conj Z[N-k]) / 2, Xb = (Z[k] - conj Z[N-k]) / 2i ; power = |X|^2
    for (int k = lane; k < kBins; k += 32) {
      const int kn = (kFftN - k) & (kFftN - 1);
      const float zr = tile_re[k], zi = tile_im[k], nr = tile_re[kn], ni = tile_im[kn];
      const float ar = 0.5f * (zr + nr), ai = 0.5f * (zi - ni);
      const float br = 0.5f * (zi + ni), bi = 0.5f * (zr - nr);
      // in place: entry k of the two tiles is read by this iteration only (its partner N-k >= 512 is never written)
      tile_re[k] = ar * ar + ai * ai;
      tile_im[k] = br * br + bi * bi;
    }
    __syncwarp();
    // sparse mel: every lane walks its balanced share of the 884 non-zeros as one flat list (api.cu builds it: the filters
    // are dealt longest-first to the least loaded lane).  Filter-per-lane rounds made the warp wait for the widest filter
    // of every round: 95 serial iterations instead of 29.  The sum order inside a filter is unchanged.
    {
      const float *pwa = tile_re, *pwb = tile_im;   // power spectra of the two frames, bins 0..512
      float* s_mel = tile_re + 520;   // [2][224] behind the first power spectrum (the upper half of the tile is dead by now)
      float a0 = 0.f, a1 = 0.f;
      for (int i = 0; i < sched_len; ++i) {
        const float2 e = s_sched[i * 32 + lane];
        const int code = __float_as_int(e.y);
        const int bin = code & 1023;
        a0 = fmaf(pwa[bin], e.x, a0);
        a1 = fmaf(pwb[bin], e.x, a1);
        if (code & 1024) {
          const int m = code >> 11;
          s_mel[m] = a0;
          s_mel[kMels + m] = a1;
          a0 = a1 = 0.f;
        }
      }
      __syncwarp();
      // dB + BN, coalesced
      const int ta = t0 + fa;
      for (int m = lane; m < kMels; m += 32) {
        const float sc = bn_scale[m], sh = bn_shift[m];
        out[((int64_t)b * n_frames + ta) * kMels + m] = 10.0f * log10f(fmaxf(s_mel[m], 1e-10f)) * sc + sh;
        if (ta + 1 < n_frames)
          out[((int64_t)b * n_frames + ta + 1) * kMels + m] = 10.0f * log10f(fmaxf(s_mel[kMels + m], 1e-10f)) * sc + sh;
      }
    }
    __syncwarp();
  }
}

int launch_frontend(const float* wav, int batch, int64_t n_samples, const FrontendParams& p, bool apply_bn,
                    float* out, cudaStream_t stream) {
  const int n_frames = (int)(n_samples / kHop) + 1;
  static const bool use_v1 = getenv("CNB_FRONTEND_V1") != nullptr;  // debugging aid: the block-wide radix-4 kernel
  if (use_v1) {
    dim3 grid((n_frames + kFramesPerCta - 1) / kFramesPerCta, batch);
    frontend_kernel<<<grid, kFeThreads, 0, stream>>>(wav, n_samples, n_frames, p.twiddle, p.mel_lo, p.mel_cnt, p.mel_off,
                                                     p.mel_w, apply_bn ? p.bn_scale : p.ones,
                                                     apply_bn ? p.bn_shift : p.zeros, out);
    CNB_LAUNCH_OK();
    return 0;
  }
  const size_t smem = (size_t)(kFe2Span + kFftN + 64 + p.mel_sched_len * 64 + kFe2Warps * kFe2WarpScratch) * sizeof(float);
  constexpr size_t smem_max = (size_t)(kFe2Span + kFftN + 64 + kMelSchedMax * 64 + kFe2Warps * kFe2WarpScratch) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    CNB_CUDA_OK(cudaFuncSetAttribute(frontend_warpfft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    attr_set = true;
  }
  dim3 grid((n_frames + kFe2Frames - 1) / kFe2Frames, batch);
  frontend_warpfft_kernel<<<grid, kFe2Warps * 32, smem, stream>>>(wav, n_samples, n_frames, p.twiddle, p.mel_lo, p.mel_cnt,
                                                                  p.mel_off, p.mel_w, p.mel_sched, p.mel_sched_len,
                                                                  apply_bn ? p.bn_scale : p.ones,
                                                                  apply_bn ? p.bn_shift : p.zeros, out);
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace cnb
