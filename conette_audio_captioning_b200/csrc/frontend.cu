// K-FE: fused log-mel front-end.
//   waveform (B, N) f32 -> reflect-pad 512 -> frames (hop 320, win 1024, periodic Hann) -> real FFT-1024 -> power
//   -> sparse Slaney mel (513 -> 224) -> 10*log10(clamp(., 1e-10)) -> eval BatchNorm per mel bin -> (B, T, 224) f32
// Replaces torchlibrosa Spectrogram + LogmelFilterBank as called at reference nn/encoders/convnext.py:160-180,276-278
// and bn0 at :290-292.  The reference evaluates the DFT as two dense Conv1d (2.1 GFLOP per
// 10 s clip) and writes the (B,T,513) power tensor to HBM; here two frames are packed into one complex radix-4 FFT
// in shared memory and only the waveform is read / the normalised log-mel written (algorithmic bytes: 4N + 4*T*224).
#include "common.cuh"
#include "kernels.h"

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

int launch_frontend(const float* wav, int batch, int64_t n_samples, const FrontendParams& p, bool apply_bn,
                    float* out, cudaStream_t stream) {
  const int n_frames = (int)(n_samples / kHop) + 1;
  dim3 grid((n_frames + kFramesPerCta - 1) / kFramesPerCta, batch);
  frontend_kernel<<<grid, kFeThreads, 0, stream>>>(wav, n_samples, n_frames, p.twiddle, p.mel_lo, p.mel_cnt, p.mel_off,
                                                   p.mel_w, apply_bn ? p.bn_scale : p.ones, apply_bn ? p.bn_shift : p.zeros,
                                                   out);
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace cnb
