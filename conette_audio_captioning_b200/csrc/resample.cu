// K-RS: polyphase sinc resampler to 32 kHz on the GPU (the input side of the boundary).
// Replaces torchaudio.functional.resample as called by the reference at huggingface/preprocessor.py:139-141
// (resampling_method "sinc_interp_hann", lowpass_filter_width 6, rolloff 0.99 -- torchaudio 0.13.1, un-vendored):
//   y[f * new + p] = sum_k kernel[p][k] * xpad[f * orig + k],  xpad = x zero-padded by (width, width + orig),
//   truncated to ceil(new * len / orig) samples per clip.
// The filter bank comes from the host (resample.py builds it with the published formula in the same dtypes) in compact
// form: phase p keeps only its non-zero support [lo[p], lo[p] + n_taps) -- the dense bank is new x (2 width + orig) wide
// but the Hann window clamps everything beyond 6 zero crossings to exactly 0 (441 -> 320: 459 columns, 18 non-zero).
// HBM-bound: a CTA stages the input span of its output frames in shared memory once (coalesced), every thread then
// produces outputs with n_taps FMAs in ascending tap order; outputs beyond a clip's resampled length are written as 0
// so that the result equals the reference's per-clip resample followed by right zero-padding.
#include "common.cuh"
#include "kernels.h"

namespace cnb {

constexpr int kRsThreads = 256;
constexpr int kRsMaxSpan = 10240;  // floats of staged input per CTA (40 KB)

__global__ void __launch_bounds__(kRsThreads)
resample_kernel(const float* __restrict__ x, int64_t n_in, const int64_t* __restrict__ lens_in, const float* __restrict__ taps,
                const int* __restrict__ lo, int orig, int nw, int n_taps, int width, int frames_per_cta, int span,
                float* __restrict__ out, int64_t n_out) {
  extern __shared__ float s_x[];
  const int b = blockIdx.y;
  const int64_t f0 = (int64_t)blockIdx.x * frames_per_cta;
  const int64_t len = lens_in ? min(lens_in[b], n_in) : n_in;
  const int64_t tgt = (len * nw + orig - 1) / orig;  // ceil(new * len / orig)
  const float* xb = x + (int64_t)b * n_in;
  const int64_t i0 = f0 * orig - width;  // input index of s_x[0]
  for (int i = threadIdx.x; i < span; i += kRsThreads) {
    const int64_t j = i0 + i;
    s_x[i] = (j >= 0 && j < len) ? __ldg(xb + j) : 0.f;
  }
  __syncthreads();
  float* ob = out + (int64_t)b * n_out;
  const int n_local = frames_per_cta * nw;
  for (int i = threadIdx.x; i < n_local; i += kRsThreads) {
    const int f = i / nw, p = i - f * nw;
    const int64_t o = (f0 + f) * nw + p;
    if (o >= n_out) break;
    float acc = 0.f;
    if (o < tgt) {
      const float* t = taps + (int64_t)p * n_taps;
      const float* sx = s_x + f * orig + lo[p];
      for (int s = 0; s < n_taps; ++s) acc = fmaf(__ldg(t + s), sx[s], acc);
    }
    ob[o] = acc;
  }
}

int launch_resample(const float* x, int batch, int64_t n_in, const int64_t* lens_in, const float* taps, const int* lo, int orig,
                    int nw, int n_taps, int width, float* out, int64_t n_out, cudaStream_t stream) {
  CNB_REQUIRE(orig > 0 && nw > 0 && n_taps > 0 && width >= 0, "resample: bad filter geometry");
  const int k_dense = 2 * width + orig;  // lo[p] + n_taps <= k_dense is the caller's contract
  int frames = 2048 / nw;
  if (frames < 1) frames = 1;
  while (frames > 1 && frames * orig + k_dense > kRsMaxSpan) --frames;
  const int span = frames * orig + k_dense;
  CNB_REQUIRE(span <= kRsMaxSpan, "resample: sample-rate ratio too large for the staged span (reduce by the gcd first)");
  const int64_t n_frames = ceil_div(n_out, nw);
  dim3 grid((unsigned)ceil_div(n_frames, frames), (unsigned)batch);
  resample_kernel<<<grid, kRsThreads, span * sizeof(float), stream>>>(x, n_in, lens_in, taps, lo, orig, nw, n_taps, width, frames,
                                                                     span, out, n_out);
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace cnb
