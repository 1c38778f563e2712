// ConvNeXt-Tiny CUDA-core kernels (everything that is not a GEMM): stem, depthwise 7x7 + LayerNorm, downsample
// LayerNorm + 2x2 im2col pack, frequency mean and the clip (tag) head.  Activations are NHWC; the residual stream is
// fp32, GEMM operands are OutT (fp16 in fast mode, f32 in parity mode).
// Reference: nn/encoders/convnext.py:61-74 (block), :207-217 (stem / downsample), :306-334 (mean + head),
// nn/modules/norm.py:35-40 (channels_first LayerNorm, biased variance, eps 1e-6).
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace cnb {

constexpr float kLnEps = 1e-6f;

// Sum 16 per-lane values across the warp with 16 shuffles: afterwards lane l holds the total of value
// idx(l) = 8*bit4(l) + 4*bit3(l) + 2*bit2(l) + bit1(l)  (lanes l and l^1 hold the same value).
__device__ __forceinline__ float warp_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool up = lane & 16;
    const float keep = up ? v[k + 8] : v[k], send = up ? v[k] : v[k + 8];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool up = lane & 8;
    const float keep = up ? v[k + 4] : v[k], send = up ? v[k] : v[k + 4];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const bool up = lane & 4;
    const float keep = up ? v[k + 2] : v[k], send = up ? v[k] : v[k + 2];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const bool up = lane & 2;
    const float keep = up ? v[1] : v[0], send = up ? v[0] : v[1];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}


// =====================================================================================================================
// K-STEM: one CTA per output row (b, h): 56 pixels x 96 channels; 8 warps x 7 pixels, lane owns channels l, l+32, l+64
// =====================================================================================================================
constexpr int kStemThreads = 256;
constexpr int kStemRows = 4;    // output rows per group (16 input frames x 224 mel bins = 14 KB of shared memory)
constexpr int kStemMaxGroups = 4;  // groups per CTA (launch parameter, 1..4): the 48 weight registers are loaded once and
                                   // the next group's frames are fetched with cp.async while the current group is computed

__device__ __forceinline__ void stem_cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;  // zero fill = the conv's zero padding in time
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kStemThreads, 3)
stem_kernel(const float* __restrict__ lm, int n_frames, int h1, int n_groups, const float* __restrict__ w_t,
            const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
            float* __restrict__ out) {
  __shared__ __align__(16) float s_in[2][kStemRows * 4][224];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int hb = blockIdx.x * (kStemRows * n_groups);
  const float* lmb = lm + (int64_t)b * n_frames * 224;
  auto fetch = [&](int grp, int buf) {
    const int h0 = hb + grp * kStemRows;
    if (h0 < h1) {
      for (int i = tid; i < kStemRows * 4 * 56; i += kStemThreads) {  // 16 frames x 56 float4
        const int r = i / 56, c4 = i - r * 56;
        const int t = 4 * h0 - 4 + r;  // Conv2d padding (4, 0): 4 zero frames before/after in time
        const bool ok = t >= 0 && t < n_frames;
        stem_cp_async16(&s_in[buf][r][4 * c4], ok ? lmb + (int64_t)t * 224 + 4 * c4 : lmb, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch(0, 0);
  float w[16][3], bia[3], g[3], be[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = lane + 32 * j;
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k][j] = w_t[k * 96 + c];
    bia[j] = bias[c];
    g[j] = ln_g[c];
    be[j] = ln_b[c];
  }
  // 8 warps x 7 pixels = one output row of 56 pixels.  The kernel is issue / latency bound, not HBM-bound, so the LayerNorm
  // statistics of a warp's 7 pixels go through ONE 16-value butterfly (7 sums | 7 sums of squares: 16 shuffles instead of
  // 70) and the normalisation is folded into one FMA per value.
  for (int grp = 0; grp < n_groups; ++grp) {
    const int h0 = hb + grp * kStemRows;
    if (h0 >= h1) break;
    const int buf = grp & 1;
    if (grp + 1 < n_groups) fetch(grp + 1, buf ^ 1);  // buffer buf^1 was released by the barrier that ended group grp-1
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    for (int rr = 0; rr < kStemRows; ++rr) {
      const int h = h0 + rr;
      if (h >= h1) break;
      float acc[7][3];
#pragma unroll
      for (int p = 0; p < 7; ++p) {
        acc[p][0] = bia[0], acc[p][1] = bia[1], acc[p][2] = bia[2];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 v = *reinterpret_cast<const float4*>(&s_in[buf][rr * 4 + r][4 * (warp * 7 + p)]);
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[p][j] = fmaf(vv[c], w[r * 4 + c][j], acc[p][j]);
        }
      }
      float st[16];
#pragma unroll
      for (int p = 0; p < 7; ++p) {
        st[p] = (acc[p][0] + acc[p][1]) + acc[p][2];
        st[p + 8] = fmaf(acc[p][0], acc[p][0], fmaf(acc[p][1], acc[p][1], acc[p][2] * acc[p][2]));
      }
      st[7] = st[15] = 0.f;
      const float tot = warp_sum16(st, lane);  // lane l holds value (l >> 1) & 15: sums 0..6 | squares 8..14
      float* o = out + (((int64_t)b * h1 + h) * 56 + warp * 7) * 96;
#pragma unroll
      for (int p = 0; p < 7; ++p) {
        const float s1 = __shfl_sync(0xffffffffu, tot, 2 * p);
        const float s2 = __shfl_sync(0xffffffffu, tot, 2 * (p + 8));
        const float mean = s1 * (1.f / 96.f);
        const float var = fmaxf(fmaf(s2, 1.f / 96.f, -mean * mean), 0.f);
        const float rstd = rsqrtf(var + kLnEps);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float a = rstd * g[j];
          o[p * 96 + lane + 32 * j] = fmaf(acc[p][j], a, fmaf(-mean, a, be[j]));
        }
      }
    }
    __syncthreads();  // everybody has finished reading s_in[buf] before the fetch of group grp+2 overwrites it
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

int launch_stem(const float* lm, int batch, int n_frames, int h1, const float* w_t, const float* bias, const float* ln_g,
                const float* ln_b, float* out, cudaStream_t stream) {
  // enough CTAs for ~6 per SM first, then up to 4 row groups per CTA (small batches / the sliced host path keep 1)
  const int64_t row_groups = (int64_t)batch * ceil_div(h1, kStemRows);
  int n_groups = (int)(row_groups / ((int64_t)kNumSMs * 6));
  n_groups = n_groups < 1 ? 1 : (n_groups > kStemMaxGroups ? kStemMaxGroups : n_groups);
  dim3 grid((unsigned)ceil_div(h1, kStemRows * n_groups), batch);
  stem_kernel<<<grid, kStemThreads, 0, stream>>>(lm, n_frames, h1, n_groups, w_t, bias, ln_g, ln_b, out);
  CNB_LAUNCH_OK();
  return 0;
}

// =====================================================================================================================
// K-DWLN: depthwise 7x7 (pad 3) + bias + LayerNorm over C, one CTA per output row (b, h).
//   thread = (channel pair cp, strip of 7 output pixels); weights for the pair live in registers (49 x float2);
//   each of the 7 kernel rows streams 13 input float2 through 49 FMAs per channel.
//   LayerNorm: two-pass (mean, then squared deviations): warp shuffles, then per-warp partials summed in fixed order.
// =====================================================================================================================
template <int C, int W, typename OutT>
__global__ void __launch_bounds__(((C / 2 + 31) / 32) * 32 * (W / 7))
dwconv_ln_kernel(const float* __restrict__ x, int H, const float* __restrict__ w_t, const float* __restrict__ bias,
                 const float* __restrict__ ln_g, const float* __restrict__ ln_b, OutT* __restrict__ out) {
  constexpr int CP = C / 2;
  constexpr int CP_PAD = ((CP + 31) / 32) * 32;
  constexpr int PW = 7;
  constexpr int NW = CP_PAD / 32;  // warps that share one strip of pixels
  __shared__ float s_sum[W][NW];   // per-warp partials, summed in a fixed order (deterministic, no atomics)
  __shared__ float s_sq[W][NW];

  const int tid = threadIdx.x;
  const int strip = tid / CP_PAD;
  const int cp = tid - strip * CP_PAD;
  const int wis = cp >> 5;  // warp index inside the strip
  const bool active = cp < CP;
  const int b = blockIdx.y, h = blockIdx.x;
  const int w0 = strip * PW;

  float2 acc[PW];
  if (active) {
    const float2 bi = *reinterpret_cast<const float2*>(bias + 2 * cp);
#pragma unroll
    for (int p = 0; p < PW; ++p) acc[p] = bi;
    const float* xb = x + (int64_t)b * H * W * C + 2 * cp;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int r = h + i - 3;
      if (r < 0 || r >= H) continue;  // zero padding at the padded-batch border
      float2 wr[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) wr[j] = __ldg(reinterpret_cast<const float2*>(w_t + (i * 7 + j) * C + 2 * cp));
      float2 in[PW + 6];
#pragma unroll
      for (int j = 0; j < PW + 6; ++j) {
        const int col = w0 + j - 3;
        in[j] = (col >= 0 && col < W) ? __ldg(reinterpret_cast<const float2*>(xb + ((int64_t)r * W + col) * C))
                                      : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int p = 0; p < PW; ++p) acc[p] = __ffma2_rn(in[p + j], wr[j], acc[p]);
    }
  } else {
#pragma unroll
    for (int p = 0; p < PW; ++p) acc[p] = make_float2(0.f, 0.f);
  }
  // pass 1: mean over C for each of the strip's pixels (all lanes of a warp share the strip)
#pragma unroll
  for (int p = 0; p < PW; ++p) {
    const float s = warp_sum(acc[p].x + acc[p].y);
    if ((tid & 31) == 0) s_sum[w0 + p][wis] = s;
  }
  __syncthreads();
  float mean[PW];
#pragma unroll
  for (int p = 0; p < PW; ++p) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) t += s_sum[w0 + p][i];
    mean[p] = t * (1.f / C);
    const float dx = active ? acc[p].x - mean[p] : 0.f, dy = active ? acc[p].y - mean[p] : 0.f;
    const float s = warp_sum(dx * dx + dy * dy);
    if ((tid & 31) == 0) s_sq[w0 + p][wis] = s;
  }
  __syncthreads();
  if (active) {
    const float2 g = *reinterpret_cast<const float2*>(ln_g + 2 * cp);
    const float2 be = *reinterpret_cast<const float2*>(ln_b + 2 * cp);
    OutT* o = out + (((int64_t)b * H + h) * W + w0) * C + 2 * cp;
#pragma unroll
    for (int p = 0; p < PW; ++p) {
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NW; ++i) q += s_sq[w0 + p][i];
      const float rstd = 1.f / sqrtf(q * (1.f / C) + kLnEps);
      const float y0 = (acc[p].x - mean[p]) * rstd * g.x + be.x;
      const float y1 = (acc[p].y - mean[p]) * rstd * g.y + be.y;
      if constexpr (sizeof(OutT) == 2) {
        *reinterpret_cast<act16x2*>(o + (int64_t)p * C) = floats2act2(y0, y1);
      } else {
        *reinterpret_cast<float2*>(o + (int64_t)p * C) = make_float2(y0, y1);
      }
    }
  }
}

// =====================================================================================================================
// K-DWLN v2 (stages 1-3): marching ring buffer, two output rows per iteration.
//   CTA = (clip, column strip of TW pixels, segment of SEG output rows).  A 10-row ring of (TW+6) x C fp32 input pixels
//   lives in shared memory; rows h+5, h+6 are prefetched with 16-byte cp.async (zero-filled outside the image = conv
//   padding) while output rows h, h+1 are computed, so every input row is fetched once per CTA (+ the 6-row halo at
//   segment starts).  thread = (channel pair, 7-pixel strip): its 49 x float2 weights stay in registers for the whole
//   march; each of the 8 live input rows is read once (13 LDS.64) and feeds both output rows: 686 packed FFMA2
//   (fma.rn.f32x2) per 104 conflict-free LDS.64.  LayerNorm: one pass (sum, sum of squares) with a 16-value butterfly
//   that costs 16 shuffles per 16 values, per-warp partials combined in fixed order through shared memory (1 barrier).
// =====================================================================================================================
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes) : "memory");
}

template <int C, int W, int TW, typename OutT>
__global__ void __launch_bounds__(((C / 2 + 31) / 32) * 32 * (TW / 7), 1)
dwconv_ln_ring_kernel(const float* __restrict__ x, int H, int seg_rows, const float* __restrict__ w_t,
                      const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                      OutT* __restrict__ out) {
  constexpr int CP = C / 2;
  constexpr int CP_PAD = ((CP + 31) / 32) * 32;
  constexpr int PW = 7;
  constexpr int NW = CP_PAD / 32;
  constexpr int NT = CP_PAD * (TW / PW);
  constexpr int RW = TW + 6;            // ring row width in pixels (3-pixel halo each side)
  constexpr int ROW_F4 = RW * C / 4;    // float4 per ring row
  constexpr int NS = 10;                // ring slots: 8 live rows + 2 in flight
  extern __shared__ __align__(16) float s_dyn[];
  float* ring = s_dyn;                              // [NS][RW][C] (+ pad so that idle lanes of a padded warp stay in bounds)
  float* s_part = s_dyn + NS * RW * C + 2 * (CP_PAD - CP);  // [2 rows][2 (sum, sq)][TW][NW]

  const int tid = threadIdx.x, lane = tid & 31;
  const int strip_t = tid / CP_PAD;
  const int cp = tid - strip_t * CP_PAD;
  const int wis = cp >> 5;
  const bool active = cp < CP;
  constexpr int STRIPS = W / TW;
  const int strip_c = blockIdx.x % STRIPS, seg = blockIdx.x / STRIPS;
  const int b = blockIdx.y;
  const int w_base = strip_c * TW;      // first output column of this CTA
  const int h0 = seg * seg_rows;
  const int h1 = min(H, h0 + seg_rows);
  if (h0 >= H) return;
  const float* xb = x + (int64_t)b * H * W * C;

  auto load_row = [&](int r, int slot) {
    float* dst = ring + slot * (RW * C);
    const bool row_ok = (r >= 0) && (r < H);
    for (int i = tid; i < ROW_F4; i += NT) {
      const int px = i / (C / 4), c4 = i - px * (C / 4);
      const int col = w_base + px - 3;
      const bool ok = row_ok && col >= 0 && col < W;
      cp_async16_zfill(dst + px * C + c4 * 4, ok ? xb + ((int64_t)r * W + col) * C + c4 * 4 : xb, ok);
    }
  };
  for (int i = 0; i < 8; ++i) load_row(h0 - 3 + i, i);
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (CP_PAD != CP) {  // idle lanes of a padded warp read past their pixel (times zero weights): keep that data finite
    for (int i = tid; i < 2 * RW * C + 2 * (CP_PAD - CP); i += NT) ring[8 * RW * C + i] = 0.f;  // slots 8, 9 + tail pad
  }

  float2 wr[49];
  float2 bi = make_float2(0.f, 0.f), g = bi, be = bi;
  if (active) {
#pragma unroll
    for (int t = 0; t < 49; ++t) wr[t] = __ldg(reinterpret_cast<const float2*>(w_t + t * C + 2 * cp));
    bi = *reinterpret_cast<const float2*>(bias + 2 * cp);
    g = *reinterpret_cast<const float2*>(ln_g + 2 * cp);
    be = *reinterpret_cast<const float2*>(ln_b + 2 * cp);
  } else {  // idle lanes of a padded warp run the same instruction stream on zero weights (no divergence)
#pragma unroll
    for (int t = 0; t < 49; ++t) wr[t] = make_float2(0.f, 0.f);
  }
  const int wl0 = strip_t * PW;         // first output column of this thread inside the CTA strip
  int slot0 = 0;                        // ring slot of input row h-3

  for (int h = h0; h < h1; h += 2) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                    // rows h-3..h+4 resident; everybody has left rows h-2, h-1
    {
      const int s8 = slot0 + 8 >= NS ? slot0 + 8 - NS : slot0 + 8;
      const int s9 = s8 + 1 >= NS ? s8 + 1 - NS : s8 + 1;
      if (h + 5 <= h1 + 2) load_row(h + 5, s8);  // overwrites rows h-5 / h-4
      if (h + 6 <= h1 + 2) load_row(h + 6, s9);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    float2 acc0[PW], acc1[PW];
#pragma unroll
    for (int p = 0; p < PW; ++p) acc0[p] = acc1[p] = bi;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int sl = slot0 + i >= NS ? slot0 + i - NS : slot0 + i;
      const float* row = ring + sl * (RW * C) + wl0 * C + 2 * cp;
      float2 in[PW + 6];
#pragma unroll
      for (int j = 0; j < PW + 6; ++j) in[j] = *reinterpret_cast<const float2*>(row + j * C);
      // tap-major order: consecutive FFMA2 hit 7 (14) different accumulators, so the issue stream has no dependent chains
#pragma unroll
      for (int j = 0; j < 7; ++j) {
#pragma unroll
        for (int p = 0; p < PW; ++p) {
          if (i < 7) acc0[p] = __ffma2_rn(in[p + j], wr[i * 7 + j], acc0[p]);
          if (i > 0) acc1[p] = __ffma2_rn(in[p + j], wr[(i - 1) * 7 + j], acc1[p]);
        }
      }
    }
    // ---- LayerNorm statistics: (sum, sumsq) for 14 pixels
    float sm[16], sq[16];
#pragma unroll
    for (int p = 0; p < PW; ++p) {
      sm[p] = acc0[p].x + acc0[p].y;
      sq[p] = fmaf(acc0[p].x, acc0[p].x, acc0[p].y * acc0[p].y);
      sm[p + 8] = acc1[p].x + acc1[p].y;
      sq[p + 8] = fmaf(acc1[p].x, acc1[p].x, acc1[p].y * acc1[p].y);
    }
    sm[7] = sm[15] = sq[7] = sq[15] = 0.f;
    const float tsum = warp_sum16(sm, lane);
    const float tsq = warp_sum16(sq, lane);
    if ((lane & 1) == 0) {
      const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
      const int row = idx >> 3, p = idx & 7;
      if (p < PW) {
        s_part[((row * 2 + 0) * TW + wl0 + p) * NW + wis] = tsum;
        s_part[((row * 2 + 1) * TW + wl0 + p) * NW + wis] = tsq;
      }
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        if (h + row >= h1) break;
        OutT* o = out + (((int64_t)b * H + h + row) * W + w_base + wl0) * C + 2 * cp;
#pragma unroll
        for (int p = 0; p < PW; ++p) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < NW; ++i) {
            s1 += s_part[((row * 2 + 0) * TW + wl0 + p) * NW + i];
            s2 += s_part[((row * 2 + 1) * TW + wl0 + p) * NW + i];
          }
          const float mean = s1 * (1.f / C);
          const float var = fmaxf(s2 * (1.f / C) - mean * mean, 0.f);
          const float rstd = 1.f / sqrtf(var + kLnEps);
          const float2 a = row ? acc1[p] : acc0[p];
          const float y0 = (a.x - mean) * rstd * g.x + be.x;
          const float y1 = (a.y - mean) * rstd * g.y + be.y;
          if constexpr (sizeof(OutT) == 2) {
            *reinterpret_cast<act16x2*>(o + (int64_t)p * C) = floats2act2(y0, y1);
          } else {
            *reinterpret_cast<float2*>(o + (int64_t)p * C) = make_float2(y0, y1);
          }
        }
      }
    }
    slot0 = slot0 + 2 >= NS ? slot0 + 2 - NS : slot0 + 2;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int C, int W, int TW, typename OutT>
static int launch_dwconv_ring_t(const float* x, int batch, int h, const float* w_t, const float* bias, const float* ln_g,
                                const float* ln_b, OutT* out, cudaStream_t stream) {
  constexpr int CP_PAD = ((C / 2 + 31) / 32) * 32;
  constexpr int threads = CP_PAD * (TW / 7);
  constexpr int NW = CP_PAD / 32;
  constexpr size_t smem = (size_t)(10 * (TW + 6) * C + 2 * (CP_PAD - C / 2) + 4 * TW * NW) * sizeof(float);
  auto kern = dwconv_ln_ring_kernel<C, W, TW, OutT>;
  static bool attr_set = false;
  if (!attr_set) {
    CNB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  constexpr int strips = W / TW;
  // about one wave of CTAs (1 CTA / SM): segments sized so that strips * batch * n_seg ~ #SMs, but at least 8 rows each
  int n_seg = kNumSMs / (strips * batch);
  if (n_seg < 1) n_seg = 1;
  int seg_rows = (int)ceil_div(h, n_seg);
  seg_rows += seg_rows & 1;  // rows are produced in pairs
  if (seg_rows < 8) seg_rows = 8;
  n_seg = (int)ceil_div(h, seg_rows);
  dim3 grid(strips * n_seg, batch);
  kern<<<grid, threads, smem, stream>>>(x, h, seg_rows, w_t, bias, ln_g, ln_b, out);
  CNB_LAUNCH_OK();
  return 0;
}

// =====================================================================================================================
// K-DWLN stage 4 (C = 768, W = 7): the whole image row (7 pixels x 768 channels = 21 KB of fp32) is one contiguous run of the
// NHWC tensor, every output pixel's 7x7 window spans the full width, and a clip has only T'/... = 31 rows.
//   CTA = 384 threads = one channel pair each, all 768 channels (LayerNorm stays inside the CTA); it walks a contiguous range
//   of the flattened (clip, row) space with its 49 x float2 weights in registers;
//   input rows live in a 10-slot shared-memory ring (7 live + 3 in flight), each filled by ONE cp.async.bulk of 21 504 bytes;
//   the horizontal zero padding is resolved at compile time (tap (p, j) exists iff 0 <= p + j - 3 < 7: 37 of 49 per kernel
//   row, 259 packed FMAs per output row instead of 343), the vertical one by skipping rows (CTA-uniform branch);
//   LayerNorm: one pass (sum | sum of squares) through the 16-value butterfly, per-warp partials combined in fixed order.
// The generic one-CTA-per-row kernel above re-read 7 input rows and 49 weights per output row from L1/L2 (57 us per launch at
// 64 clips, 4x its FP32 floor).
// =====================================================================================================================
constexpr int kW7Slots = 10;
constexpr int kW7RowFloats = 7 * 768;
constexpr int kW7RowBytes = kW7RowFloats * 4;
constexpr int kW7Smem = kW7Slots * kW7RowBytes + 2 * 16 * 12 * 4 + kW7Slots * 8 + 128;

__device__ __forceinline__ void w7_mbar_init(uint32_t bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); }
__device__ __forceinline__ void w7_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void w7_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (!ok && ++spins > (1u << 22)) {  // a lost row would otherwise hang the GPU: fail loudly instead
      printf("conette_b200: dwconv_ln_w7 row wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void w7_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

template <typename OutT>
__global__ void __launch_bounds__(384, 1)
dwconv_ln_w7_kernel(const float* __restrict__ x, int H, int rows_total, int quota, const float* __restrict__ w_t,
                    const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                    OutT* __restrict__ out) {
  constexpr int C = 768, W = 7;
  extern __shared__ uint8_t w7_raw[];
  const uint32_t base = ((uint32_t)__cvta_generic_to_shared(w7_raw) + 127u) & ~127u;
  uint8_t* sm = w7_raw + (base - (uint32_t)__cvta_generic_to_shared(w7_raw));
  const float* ring = reinterpret_cast<const float*>(sm);
  float* s_part = reinterpret_cast<float*>(sm + kW7Slots * kW7RowBytes);  // [2][16 values][12 warps]
  const uint32_t bar0 = base + kW7Slots * kW7RowBytes + 2 * 16 * 12 * 4;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kW7Slots; ++s) w7_mbar_init(bar0 + 8 * s);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float2 wr[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) wr[t] = __ldg(reinterpret_cast<const float2*>(w_t + t * C + 2 * tid));
  const float2 bi = *reinterpret_cast<const float2*>(bias + 2 * tid);
  const float2 gm = *reinterpret_cast<const float2*>(ln_g + 2 * tid);
  const float2 be = *reinterpret_cast<const float2*>(ln_b + 2 * tid);
  __syncthreads();

  const int u_begin = (int)blockIdx.x * quota, u_end = min(rows_total, u_begin + quota);
  uint32_t k_issue = 0;  // input rows requested so far by this CTA: request k lives in slot k % 10, barrier phase (k / 10) & 1
  uint32_t it = 0;       // output rows produced so far (partial buffer it & 1)
  for (int u = u_begin; u < u_end;) {
    const int b = u / H, h_first = u - b * H;
    const int h_last = min(H, h_first + (u_end - u)) - 1;  // last output row of this segment (same clip)
    u += h_last - h_first + 1;
    const int r_lo = max(0, h_first - 3), r_hi = min(H - 1, h_last + 3);
    const uint32_t k_seg = k_issue;  // request index of input row r_lo
    const float* xb = x + (int64_t)b * H * kW7RowFloats;
    int r_req = r_lo;                // next input row to request
    int r_seen = r_lo - 1;           // rows up to here have landed (as seen by this thread)
    for (int h = h_first; h <= h_last; ++h, ++it) {
      // slots of rows < h - 3 were released by the barrier that ended the previous output row
      const int want = min(r_hi, h + kW7Slots - 4);  // 7 live rows + (slots - 7) in flight
      if (tid == 0) {
        for (int r = r_req; r <= want; ++r) {
          const uint32_t s = (k_issue + (uint32_t)(r - r_req)) % kW7Slots, bar = bar0 + 8 * s;
          w7_mbar_expect(bar, kW7RowBytes);
          w7_bulk_load(base + s * kW7RowBytes, xb + (int64_t)r * kW7RowFloats, kW7RowBytes, bar);
        }
      }
      if (want >= r_req) {  // every thread tracks the request counters
        k_issue += (uint32_t)(want - r_req + 1);
        r_req = want + 1;
      }
      const int need = min(r_hi, h + 3);
      for (; r_seen < need; ++r_seen) {
        const uint32_t k = k_seg + (uint32_t)(r_seen + 1 - r_lo);
        w7_mbar_wait(bar0 + 8 * (k % kW7Slots), (k / kW7Slots) & 1u);
      }
      float2 acc[W];
#pragma unroll
      for (int p = 0; p < W; ++p) acc[p] = bi;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const int r = h + i - 3;
        if (r < 0 || r >= H) continue;  // zero padding above / below the clip (Appendix F.3: also at the padded-batch border)
        const uint32_t k = k_seg + (uint32_t)(r - r_lo);
        const float2* row = reinterpret_cast<const float2*>(ring + (size_t)(k % kW7Slots) * kW7RowFloats) + tid;
        float2 in[W];
#pragma unroll
        for (int q = 0; q < W; ++q) in[q] = row[q * (C / 2)];
#pragma unroll
        for (int j = 0; j < 7; ++j)
#pragma unroll
          for (int p = 0; p < W; ++p)
            if (p + j - 3 >= 0 && p + j - 3 < W) acc[p] = __ffma2_rn(in[p + j - 3], wr[i * 7 + j], acc[p]);
      }
      // LayerNorm over the 768 channels of each of the 7 pixels
      float st[16];
#pragma unroll
      for (int p = 0; p < W; ++p) {
        st[p] = acc[p].x + acc[p].y;
        st[p + 8] = fmaf(acc[p].x, acc[p].x, acc[p].y * acc[p].y);
      }
      st[7] = st[15] = 0.f;
      const float tot = warp_sum16(st, lane);  // lane l holds value (l >> 1) & 15
      float* part = s_part + (it & 1u) * (16 * 12);
      if ((lane & 1) == 0) part[(lane >> 1) * 12 + warp] = tot;
      __syncthreads();  // partials visible; every thread has also finished reading the ring for this output row
      // second stage in every warp: lane v < 16 adds the 12 per-warp partials of value v in fixed order (12 LDS per lane
      // instead of 168 broadcast loads per thread), the pixel statistics travel by shuffle
      float tv = 0.f;
      if (lane < 16) {
#pragma unroll
        for (int w = 0; w < 12; ++w) tv += part[lane * 12 + w];
      }
      OutT* o = out + ((int64_t)(b * H + h) * W) * C + 2 * tid;
#pragma unroll
      for (int p = 0; p < W; ++p) {
        const float s1 = __shfl_sync(0xffffffffu, tv, p), s2 = __shfl_sync(0xffffffffu, tv, p + 8);
        const float mean = s1 * (1.f / C);
        const float var = fmaxf(fmaf(s2, 1.f / C, -mean * mean), 0.f);
        const float rstd = 1.f / sqrtf(var + kLnEps);
        const float y0 = (acc[p].x - mean) * rstd * gm.x + be.x;
        const float y1 = (acc[p].y - mean) * rstd * gm.y + be.y;
        if constexpr (sizeof(OutT) == 2) {
          *reinterpret_cast<act16x2*>(o + (int64_t)p * C) = floats2act2(y0, y1);
        } else {
          *reinterpret_cast<float2*>(o + (int64_t)p * C) = make_float2(y0, y1);
        }
      }
    }
    __syncthreads();  // the next segment's first requests overwrite slots this segment may still be reading
  }
}

template <typename OutT>
static int launch_dwconv_ln_w7(const float* x, int batch, int h, const float* w_t, const float* bias, const float* ln_g,
                               const float* ln_b, OutT* out, cudaStream_t stream) {
  auto kern = dwconv_ln_w7_kernel<OutT>;
  static bool attr_set = false;
  if (!attr_set) {
    CNB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kW7Smem));
    attr_set = true;
  }
  const int total = batch * h;
  if (total == 0) return 0;
  // contiguous row ranges, one per SM; at least 4 rows each so that the 6 halo rows of a range stay a minority
  int ctas = std::min(sm_budget(), (int)ceil_div(total, 4));
  const int quota = (int)ceil_div(total, ctas);
  ctas = (int)ceil_div(total, quota);
  kern<<<ctas, 384, kW7Smem, stream>>>(x, h, total, quota, w_t, bias, ln_g, ln_b, out);
  CNB_LAUNCH_OK();
  return 0;
}

template <int C, int W, typename OutT>
static int launch_dwconv_ln_t(const float* x, int batch, int h, const float* w_t, const float* bias, const float* ln_g,
                              const float* ln_b, OutT* out, cudaStream_t stream) {
  constexpr int threads = ((C / 2 + 31) / 32) * 32 * (W / 7);
  dim3 grid(h, batch);
  dwconv_ln_kernel<C, W, OutT><<<grid, threads, 0, stream>>>(x, h, w_t, bias, ln_g, ln_b, out);
  CNB_LAUNCH_OK();
  return 0;
}

template <typename OutT>
int launch_dwconv_ln(const float* x, int batch, int h, int w, int c, const float* w_t, const float* bias, const float* ln_g,
                     const float* ln_b, OutT* out, cudaStream_t stream) {
  static const bool use_tma = getenv("CNB_DWCONV_V2") == nullptr;  // A/B switch: the cp.async ring kernels below
  if (use_tma) {
    const int rc = launch_dwconv_ln_tma<OutT>(x, batch, h, w, c, w_t, bias, ln_g, ln_b, out, stream);
    if (rc <= 0) return rc;  // 1 = no TMA-ring instantiation for this (C, W)
  }
  if (c == 96 && w == 56) return launch_dwconv_ring_t<96, 56, 28, OutT>(x, batch, h, w_t, bias, ln_g, ln_b, out, stream);
  if (c == 192 && w == 28) return launch_dwconv_ring_t<192, 28, 14, OutT>(x, batch, h, w_t, bias, ln_g, ln_b, out, stream);
  if (c == 384 && w == 14) return launch_dwconv_ring_t<384, 14, 7, OutT>(x, batch, h, w_t, bias, ln_g, ln_b, out, stream);
  static const bool w7_ring = getenv("CNB_DW_S4_GENERIC") == nullptr;  // A/B switch: the one-CTA-per-row kernel
  if (c == 768 && w == 7 && w7_ring) return launch_dwconv_ln_w7<OutT>(x, batch, h, w_t, bias, ln_g, ln_b, out, stream);
  if (c == 768 && w == 7) return launch_dwconv_ln_t<768, 7, OutT>(x, batch, h, w_t, bias, ln_g, ln_b, out, stream);
  set_error("dwconv_ln: unsupported (C, W) = (" + std::to_string(c) + ", " + std::to_string(w) + ")");
  return -1;
}
template int launch_dwconv_ln<float>(const float*, int, int, int, int, const float*, const float*, const float*,
                                     const float*, float*, cudaStream_t);
template int launch_dwconv_ln<act16>(const float*, int, int, int, int, const float*, const float*, const float*,
                                             const float*, act16*, cudaStream_t);

// =====================================================================================================================
// K-DS (part 1): LayerNorm(channels_first) per pixel + pack 2x2/stride-2 patches as GEMM rows.
//   out[(b, h', w'), (kh, kw, c)] = LN(x[b, 2h'+kh, 2w'+kw, :])[c]   (odd trailing row/col dropped: floor)
//   C / 12 lanes per input pixel (8 / 16 / 32), 32 / (C / 12) pixels per warp pass.
// =====================================================================================================================
template <int C, typename OutT>
__global__ void __launch_bounds__(256)
ln_pack2x2_kernel(const float* __restrict__ x, int batch, int H, int W, const float* __restrict__ ln_g,
                  const float* __restrict__ ln_b, OutT* __restrict__ out) {
  // LPP = C / 12 lanes share one input pixel (8 / 16 / 32 for C = 96 / 192 / 384): a lane owns the float4 groups
  // lip, lip + LPP, lip + 2 LPP of the pixel's channels (coalesced), so a warp normalises 32 / LPP pixels at once and the
  // two reductions take log2(LPP) shuffles.  The first version (one warp per pixel, 24 of 32 lanes active at C = 96,
  // 64-bit index divisions) was issue-bound: ncu counted 217 warp instructions per pixel at 25 % of the DRAM rate.
  constexpr int LPP = C / 12;
  constexpr int PPW = 32 / LPP;
  static_assert(C % 12 == 0 && (LPP & (LPP - 1)) == 0 && LPP <= 32, "lanes per pixel");
  const int lane = threadIdx.x & 31;
  const int lip = lane % LPP, sub = lane / LPP;
  const int Ho = H / 2, Wo = W / 2;
  const uint32_t n_pix = (uint32_t)batch * (uint32_t)(Ho * 2) * (uint32_t)(Wo * 2);
  const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  float4 g[3], be[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    g[i] = *reinterpret_cast<const float4*>(ln_g + 4 * (lip + LPP * i));
    be[i] = *reinterpret_cast<const float4*>(ln_b + 4 * (lip + LPP * i));
  }
  constexpr int U = 2;  // passes in flight per warp (all loads issued before the first reduction)
  for (uint32_t base = warp_id * (PPW * U); base < n_pix; base += n_warps * (PPW * U)) {
    float4 v[U][3];
    OutT* o[U];
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // enumerate only the pixels that are used: (b, h < 2*Ho, w < 2*Wo); odd trailing row / column dropped (floor)
      uint32_t pix = base + u * PPW + sub;
      valid[u] = pix < n_pix;
      if (!valid[u]) pix = n_pix - 1;
      const uint32_t rowq = pix / (uint32_t)(2 * Wo);
      const int wq = (int)(pix - rowq * (uint32_t)(2 * Wo));
      const uint32_t bq = rowq / (uint32_t)(2 * Ho);
      const int hq = (int)(rowq - bq * (uint32_t)(2 * Ho));
      const float* px = x + (((int64_t)bq * H + hq) * W + wq) * C;
      const int64_t row = ((int64_t)bq * Ho + hq / 2) * Wo + wq / 2;
      o[u] = out + row * (4 * (int64_t)C) + ((hq & 1) * 2 + (wq & 1)) * C;
#pragma unroll
      for (int i = 0; i < 3; ++i) v[u][i] = __ldg(reinterpret_cast<const float4*>(px + 4 * (lip + LPP * i)));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) s += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
#pragma unroll
      for (int ofs = LPP / 2; ofs > 0; ofs >>= 1) s += __shfl_xor_sync(0xffffffffu, s, ofs);
      const float mean = s * (1.f / C);
      float qq = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float dx = v[u][i].x - mean, dy = v[u][i].y - mean, dz = v[u][i].z - mean, dw = v[u][i].w - mean;
        qq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
#pragma unroll
      for (int ofs = LPP / 2; ofs > 0; ofs >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, ofs);
      const float rstd = 1.f / sqrtf(qq * (1.f / C) + kLnEps);
      if (!valid[u]) continue;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float y0 = (v[u][i].x - mean) * rstd * g[i].x + be[i].x, y1 = (v[u][i].y - mean) * rstd * g[i].y + be[i].y;
        const float y2 = (v[u][i].z - mean) * rstd * g[i].z + be[i].z, y3 = (v[u][i].w - mean) * rstd * g[i].w + be[i].w;
        OutT* dst = o[u] + 4 * (lip + LPP * i);
        if constexpr (sizeof(OutT) == 2) {
          act16x2 p0 = floats2act2(y0, y1), p1 = floats2act2(y2, y3);
          uint2 w2;
          w2.x = *reinterpret_cast<uint32_t*>(&p0);
          w2.y = *reinterpret_cast<uint32_t*>(&p1);
          *reinterpret_cast<uint2*>(dst) = w2;
        } else {
          *reinterpret_cast<float4*>(dst) = make_float4(y0, y1, y2, y3);
        }
      }
    }
  }
}

template <typename OutT>
int launch_ln_pack2x2(const float* x, int batch, int h, int w, int c, const float* ln_g, const float* ln_b, OutT* out,
                      cudaStream_t stream) {
  const int64_t n_pix = (int64_t)batch * (h / 2) * 2 * (w / 2) * 2;
  CNB_REQUIRE(n_pix < ((int64_t)1 << 31), "ln_pack2x2: too many pixels for 32-bit indexing (split the batch)");
  const int blocks = (int)std::min<int64_t>(ceil_div(n_pix, 8 * 2), (int64_t)kNumSMs * 8);
  if (c == 96) ln_pack2x2_kernel<96, OutT><<<blocks, 256, 0, stream>>>(x, batch, h, w, ln_g, ln_b, out);
  else if (c == 192) ln_pack2x2_kernel<192, OutT><<<blocks, 256, 0, stream>>>(x, batch, h, w, ln_g, ln_b, out);
  else if (c == 384) ln_pack2x2_kernel<384, OutT><<<blocks, 256, 0, stream>>>(x, batch, h, w, ln_g, ln_b, out);
  else {
    set_error("ln_pack2x2: unsupported channel count " + std::to_string(c));
    return -1;
  }
  CNB_LAUNCH_OK();
  return 0;
}
template int launch_ln_pack2x2<float>(const float*, int, int, int, int, const float*, const float*, float*, cudaStream_t);
template int launch_ln_pack2x2<act16>(const float*, int, int, int, int, const float*, const float*, act16*,
                                              cudaStream_t);

// =====================================================================================================================
// mean over frequency (reference convnext.py:306): (B, T', W, C) -> (B, T', C)
// =====================================================================================================================
__global__ void freq_mean_kernel(const float* __restrict__ x, int64_t n_rows, int W, int C, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * C) return;
  const int64_t row = i / C;
  const int c = (int)(i - row * C);
  float s = 0.f;
  for (int w = 0; w < W; ++w) s += x[(row * W + w) * C + c];
  out[i] = s / W;
}

int launch_freq_mean(const float* x, int batch, int tp, int w, int c, float* out, cudaStream_t stream) {
  const int64_t n = (int64_t)batch * tp * c;
  freq_mean_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(x, (int64_t)batch * tp, w, c, out);
  CNB_LAUNCH_OK();
  return 0;
}

// =====================================================================================================================
// clip head (reference convnext.py:324-334): max_t + mean_t -> LayerNorm(768, eps 1e-6) -> Linear(527) -> sigmoid
//   grid (clip, class slice), 256 threads; 768 = 3 channels per thread
// =====================================================================================================================
__global__ void __launch_bounds__(256)
clip_head_kernel(const float* __restrict__ fe, int tp, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                 const float* __restrict__ w, const float* __restrict__ bias, int n_cls, float* __restrict__ out) {
  __shared__ __align__(16) float s_v[768];
  __shared__ float s_red[8];
  __shared__ float s_stat[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* f = fe + (int64_t)blockIdx.x * tp * 768;
  float v[3];
  float part = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = tid + 256 * j;
    float mx = -INFINITY, sm = 0.f;
    for (int t = 0; t < tp; ++t) {
      const float a = f[(int64_t)t * 768 + c];
      mx = fmaxf(mx, a);
      sm += a;
    }
    v[j] = mx + sm / tp;
    part += v[j];
  }
  part = warp_sum(part);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += s_red[i];
    s_stat[0] = s / 768.f;
  }
  __syncthreads();
  const float mean = s_stat[0];
  part = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) part += (v[j] - mean) * (v[j] - mean);
  part = warp_sum(part);
  __syncthreads();
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += s_red[i];
    s_stat[1] = 1.f / sqrtf(s / 768.f + kLnEps);
  }
  __syncthreads();
  const float rstd = s_stat[1];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = tid + 256 * j;
    s_v[c] = (v[j] - mean) * rstd * ln_g[c] + ln_b[c];
  }
  __syncthreads();
  // classes [blockIdx.y * per, +per) of this clip: one warp per class, 6 x LDG.128 per lane, all in flight before the first use
  const int per = (n_cls + (int)gridDim.y - 1) / (int)gridDim.y;
  const int n_end = min(n_cls, ((int)blockIdx.y + 1) * per);
  float4 xv[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) xv[j] = *reinterpret_cast<const float4*>(&s_v[4 * (lane + 32 * j)]);
  for (int n = (int)blockIdx.y * per + warp; n < n_end; n += 8) {
    const float4* wr = reinterpret_cast<const float4*>(w + (int64_t)n * 768);
    float4 wv[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) wv[j] = __ldg(wr + lane + 32 * j);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      acc = fmaf(xv[j].x, wv[j].x, acc);
      acc = fmaf(xv[j].y, wv[j].y, acc);
      acc = fmaf(xv[j].z, wv[j].z, acc);
      acc = fmaf(xv[j].w, wv[j].w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[(int64_t)blockIdx.x * n_cls + n] = 1.f / (1.f + expf(-(acc + bias[n])));
  }
}

int launch_clip_head(const float* frame_embs, int batch, int tp, const float* ln_g, const float* ln_b, const float* w,
                     const float* bias, int n_cls, float* out, cudaStream_t stream) {
  // the pooling + LayerNorm prologue (95 KB of L2 reads per clip) is repeated by every class slice: cheaper than a second launch
  const int slices = batch >= 148 ? 4 : 8;
  clip_head_kernel<<<dim3(batch, slices), 256, 0, stream>>>(frame_embs, tp, ln_g, ln_b, w, bias, n_cls, out);
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace cnb
