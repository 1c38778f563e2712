// K-DWLN v3: depthwise 7x7 (pad 3) + bias + LayerNorm(C) for ConvNeXt stages 1-3 (reference convnext.py:61-66), sm_100a.
//
// The op is bound by the FP32 pipe, not by HBM (49 FMA per 6 algorithmic bytes), so the kernel is organised around keeping
// every FMA lane of every SM busy with packed fma.rn.f32x2:
//   * thread = (row group g, 7-pixel strip, channel PAIR): its 49 x float2 weights stay in registers for the whole launch;
//     per iteration it produces 2 output rows x 7 pixels x 2 channels from 8 input rows x 13 pixels (104 LDS.64, 686 FFMA2);
//   * CTA = 384 threads = 12 warps (3 per scheduler): 2 row groups x (TW/7 strips) x (C/2 pairs); no padded lanes -- for
//     C = 96 a strip owns 1.5 warps and the LayerNorm reduction works on half-warps;
//   * input rows live in a 14-slot shared-memory ring shared by both row groups (10 live rows + 4 in flight), filled by
//     ONE thread with cp.async.bulk.tensor (4-D NHWC map, box = C x (TW+6) pixels): the conv zero padding -- left/right
//     halo, rows above/below the clip -- is the TMA out-of-bounds fill, so there is no per-thread address arithmetic;
//   * work = the flattened (clip, column strip, row quad) space cut into equal contiguous ranges, one per SM (148 CTAs):
//     an SM crosses a column boundary at most a few times and all SMs finish together;
//   * C = 384 (stage 3) does not fit a ring of full-channel rows: a 2-CTA cluster splits the channels (192 each, same
//     smem / thread shape as stage 2) and the CTAs push their LayerNorm partials into each other's shared memory
//     (st.async pushes that complete bytes on the receiver's mbarrier); the partial buffers alternate so that a fast peer
//     can never overwrite sums that are still being read;
//   * LayerNorm: one pass (sum, sum of squares); a 16-value butterfly costs 15-16 shuffles per 16 values; per-(half-)warp
//     partials are combined in fixed order through shared memory (deterministic); normalisation with packed FFMA2.
#include <cuda.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace cnb {

namespace {

constexpr float kLnEps = 1e-6f;
constexpr int kRG = 2;        // row groups per CTA (2 output rows each)
constexpr int kNS = 14;       // ring slots
constexpr int kThreads = 384;  // CTA size of the one-CTA-per-SM shapes; narrow strips (THREADS = 192) run two CTAs per SM

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__device__ __forceinline__ uint32_t map_peer_smem(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

// One butterfly stage: n values survive; lanes with (lane & mask) keep the upper half.
template <int N>
__device__ __forceinline__ void bfly_stage(float (&v)[16], int lane, int mask) {
  const bool up = lane & mask;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const float keep = up ? v[k + N] : v[k], send = up ? v[k] : v[k + N];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
  }
}
// 16 values summed over the 16 lanes of a half-warp (15 shuffles): lane l ends with value (l & 15) bit-reversed-free index
// idx = 8*b3 + 4*b2 + 2*b1 + b0 of its half-warp.
__device__ __forceinline__ float half_sum16(float (&v)[16], int lane) {
  bfly_stage<8>(v, lane, 8);
  bfly_stage<4>(v, lane, 4);
  bfly_stage<2>(v, lane, 2);
  bfly_stage<1>(v, lane, 1);
  return v[0];
}
// 16 values summed over the 32 lanes of a warp (16 shuffles): lane l ends with value idx = (l >> 1) & 15 (bit order
// 8*b4 + 4*b3 + 2*b2 + b1); lanes l and l^1 hold the same total.
__device__ __forceinline__ float warp_sum16(float (&v)[16], int lane) {
  bfly_stage<8>(v, lane, 16);
  bfly_stage<4>(v, lane, 8);
  bfly_stage<2>(v, lane, 4);
  bfly_stage<1>(v, lane, 2);
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// C = channels handled by one CTA, CS = channel split (CTAs per cluster; the layer has C * CS channels)
template <int C, int CS, int W, int TW, typename OutT>
struct DwCfg {
  static constexpr int CT = C * CS;
  static constexpr int CP = C / 2;
  static constexpr int NSTRIP = TW / 7;
  static constexpr int GROUP_T = CP * NSTRIP;           // threads of one row group
  static constexpr bool HALF = (CP % 32) != 0;          // strips own a non-integral number of warps: reduce per half-warp
  static constexpr int NPART = HALF ? CP / 16 : CP / 32;  // partial sums per pixel produced by one CTA
  static constexpr int NPT = NPART * CS;                // ... per pixel in total
  static constexpr int PSTR = NPT <= 4 ? 8 : 16;        // floats per pixel in s_part: [sum x NPT | pad][sq x NPT | pad]
  static constexpr int RW = TW + 6;
  static constexpr int ROW_FLOATS = RW * C;
  static constexpr int ROW_BYTES = ROW_FLOATS * 4;
  static constexpr int STRIPS = W / TW;
  static constexpr int THREADS = GROUP_T * kRG;         // 384, or 192 for the 14-pixel strips of stage 1
  static constexpr int CTAS_PER_SM = kThreads / THREADS;  // 168 registers/thread: 384 threads per SM either way
  static constexpr int PART_FLOATS = kRG * NSTRIP * 16 * PSTR;
  static constexpr int PART_BUFS = CS;                  // cluster mode alternates two partial buffers
  static constexpr int RING_OFF = 0;
  static constexpr int PART_OFF = kNS * ROW_BYTES;
  static constexpr int VEC_OFF = PART_OFF + PART_BUFS * PART_FLOATS * 4;   // bias | gamma | beta
  static constexpr int BAR_OFF = VEC_OFF + 3 * C * 4;
  static constexpr int SMEM = BAR_OFF + 32 + 128;              // 2 ring barriers + 2 partial-exchange barriers + alignment slack
  // bytes the peer CTA pushes into this CTA's partial buffer per iteration: every warp has 16 writer lanes x (sum, sumsq)
  static constexpr int PEER_BYTES = (THREADS / 32) * 16 * 2 * 4;
  static_assert(CS == 1 || CS == 2, "channel split");
  static_assert(THREADS == kThreads || THREADS * 2 == kThreads, "CTA shape");
  static_assert(SMEM * CTAS_PER_SM <= 232448 - 1024 * CTAS_PER_SM, "shared memory budget per SM");
  static_assert(CP % 16 == 0 && TW % 7 == 0 && W % TW == 0, "tiling");
  static_assert(ROW_BYTES % 128 == 0, "TMA destination alignment");
  static_assert(NPT <= 8, "partials per pixel");
  static_assert(SMEM <= 232448, "shared memory budget");
};

template <int C, int CS, int W, int TW, typename OutT>
__global__ void __launch_bounds__(DwCfg<C, CS, W, TW, OutT>::THREADS, DwCfg<C, CS, W, TW, OutT>::CTAS_PER_SM)
dwconv_ln_tma_kernel(const __grid_constant__ CUtensorMap map_x, int H, int PQ, int total_units, int quota,
                     const float* __restrict__ w_t, const float* __restrict__ bias, const float* __restrict__ ln_g,
                     const float* __restrict__ ln_b, OutT* __restrict__ out) {
  using Cfg = DwCfg<C, CS, W, TW, OutT>;
  constexpr int CP = Cfg::CP, PW = 7, NPT = Cfg::NPT, PSTR = Cfg::PSTR, ROW_FLOATS = Cfg::ROW_FLOATS, CT = Cfg::CT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* sm = smem_raw + (base_u32 - smem_u32(smem_raw));
  const float* ring = reinterpret_cast<const float*>(sm + Cfg::RING_OFF);
  float* s_part = reinterpret_cast<float*>(sm + Cfg::PART_OFF);
  float* s_vec = reinterpret_cast<float*>(sm + Cfg::VEC_OFF);
  const uint32_t bar0 = base_u32 + Cfg::BAR_OFF;

  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid / Cfg::GROUP_T;
  const int rem = tid - g * Cfg::GROUP_T;
  const int strip_t = rem / CP;
  const int cp = rem - strip_t * CP;
  const int part = Cfg::HALF ? (cp >> 4) : (cp >> 5);
  const int wl0 = strip_t * PW;
  uint32_t rank = 0;  // which C-channel slice of the layer this CTA owns
  if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int c_off = (int)rank * C;

  for (int i = tid; i < C; i += Cfg::THREADS) {
    s_vec[i] = bias[c_off + i];
    s_vec[C + i] = ln_g[c_off + i];
    s_vec[2 * C + i] = ln_b[c_off + i];
  }
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    mbar_init(bar0 + 16, 1);   // CS > 1: the peer's LayerNorm partials of even / odd iterations have landed
    mbar_init(bar0 + 24, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async_smem();
  }
  float2 wr[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) wr[t] = __ldg(reinterpret_cast<const float2*>(w_t + t * CT + c_off + 2 * cp));
  if (CS > 1) {  // the peer must be running (its shared memory mapped) before anybody pushes partials into it
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }

  const int u_begin = (int)(blockIdx.x / CS) * quota;
  const int u_end = min(total_units, u_begin + quota);
  uint32_t q = 0;  // running batch counter: batch q is consumed by the q-th iteration of this CTA (barrier q&1, parity (q>>1)&1)

  for (int u = u_begin; u < u_end;) {
    const int col = u / PQ, q0 = u - col * PQ;
    const int nq = min(u_end - u, PQ - q0);
    const int b = col / Cfg::STRIPS, strip_c = col - b * Cfg::STRIPS;
    const int w_base = strip_c * TW;
    const int h0 = 4 * q0;
    const int h_end = min(H, h0 + 4 * nq);
    u += nq;

    __syncthreads();  // every thread has left the previous item's ring (and the setup above is visible)
    if (tid == 0) {
      const uint32_t bar = bar0 + 8 * (q & 1);
      mbar_expect_tx(bar, 10 * Cfg::ROW_BYTES);
#pragma unroll 1
      for (int v = 0; v < 10; ++v) tma_load_4d(base_u32 + Cfg::RING_OFF + v * Cfg::ROW_BYTES, &map_x, c_off, w_base - 3, h0 - 3 + v, b, bar);
    }
    int sbase = 0;  // ring slot of input row h - 3 (virtual row 4j)
#pragma unroll 1
    for (int j = 0; j < nq; ++j, ++q) {
      const int h = h0 + 4 * j;
      __syncthreads();  // iteration j-1 is finished everywhere: its four oldest rows and s_part can be overwritten
      if (tid == 0 && j + 1 < nq) {
        const uint32_t bar = bar0 + 8 * ((q + 1) & 1);
        mbar_expect_tx(bar, 4 * Cfg::ROW_BYTES);
        int sl = sbase + 10;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (sl >= kNS) sl -= kNS;
          tma_load_4d(base_u32 + Cfg::RING_OFF + sl * Cfg::ROW_BYTES, &map_x, c_off, w_base - 3, h + 7 + t, b, bar);
          ++sl;
        }
      }
      if (CS > 1 && tid == 0) mbar_expect_tx(bar0 + 16 + 8 * (q & 1), Cfg::PEER_BYTES);
      mbar_wait(bar0 + 8 * (q & 1), (q >> 1) & 1);

      float2 acc0[PW], acc1[PW];
      {
        const float2 bi = *reinterpret_cast<const float2*>(s_vec + 2 * cp);
#pragma unroll
        for (int p = 0; p < PW; ++p) acc0[p] = acc1[p] = bi;
      }
      int sl = sbase + 2 * g;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (sl >= kNS) sl -= kNS;
        const float* row = ring + sl * ROW_FLOATS + wl0 * C + 2 * cp;
        ++sl;
        float2 in[PW + 6];
#pragma unroll
        for (int k = 0; k < PW + 6; ++k) in[k] = *reinterpret_cast<const float2*>(row + k * C);
        // tap-major order: consecutive FFMA2 hit 7 (14) different accumulators, so the issue stream has no dependent chains
#pragma unroll
        for (int k = 0; k < 7; ++k) {
#pragma unroll
          for (int p = 0; p < PW; ++p) {
            if (i < 7) acc0[p] = __ffma2_rn(in[p + k], wr[i * 7 + k], acc0[p]);
            if (i > 0) acc1[p] = __ffma2_rn(in[p + k], wr[(i - 1) * 7 + k], acc1[p]);
          }
        }
      }
      // ---- LayerNorm statistics: (sum, sumsq) of 14 pixels, value index = 8 * row + p
      float smv[16], sqv[16];
#pragma unroll
      for (int p = 0; p < PW; ++p) {
        smv[p] = acc0[p].x + acc0[p].y;
        sqv[p] = fmaf(acc0[p].x, acc0[p].x, acc0[p].y * acc0[p].y);
        smv[p + 8] = acc1[p].x + acc1[p].y;
        sqv[p + 8] = fmaf(acc1[p].x, acc1[p].x, acc1[p].y * acc1[p].y);
      }
      smv[7] = smv[15] = sqv[7] = sqv[15] = 0.f;
      float* my_part = s_part + (CS > 1 ? (q & 1) * Cfg::PART_FLOATS : 0) + (g * Cfg::NSTRIP + strip_t) * 16 * PSTR;
      {
        float tsum, tsq;
        int idx;
        bool writer = true;
        if (Cfg::HALF) {
          tsum = half_sum16(smv, lane);
          tsq = half_sum16(sqv, lane);
          idx = lane & 15;
        } else {
          tsum = warp_sum16(smv, lane);
          tsq = warp_sum16(sqv, lane);
          idx = (lane >> 1) & 15;
          writer = (lane & 1) == 0;
        }
        if (writer) {
          float* dst = my_part + idx * PSTR + (int)rank * Cfg::NPART + part;
          dst[0] = tsum;
          dst[PSTR / 2] = tsq;
          if (CS > 1) {
            // one-sided push: each store completes its bytes on the RECEIVER's mbarrier -- no cluster barrier and no
            // cluster-scope fence per iteration (measured against a barrier.cluster release/acquire pair: -2.4 %)
            const uint32_t peer = map_peer_smem(smem_u32(dst), rank ^ 1u);
            const uint32_t peer_bar = map_peer_smem(bar0 + 16 + 8 * (q & 1), rank ^ 1u);
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(peer),
                         "r"(__float_as_uint(tsum)), "r"(peer_bar)
                         : "memory");
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(
                             peer + 4u * (PSTR / 2)),
                         "r"(__float_as_uint(tsq)), "r"(peer_bar)
                         : "memory");
          }
        }
      }
      __syncthreads();   // this CTA's partials are in place
      // the peer's partials: buffer q&1 was last read in iteration q-2, which every thread of this CTA finished before any
      // of them pushed iteration q-1, and the peer cannot push iteration q before it has received all of q-1
      if (CS > 1) mbar_wait(bar0 + 16 + 8 * (q & 1), (q >> 1) & 1);
      const float2 gm = *reinterpret_cast<const float2*>(s_vec + C + 2 * cp);
      const float2 be = *reinterpret_cast<const float2*>(s_vec + 2 * C + 2 * cp);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int hr = h + 2 * g + r;
        if (hr >= h_end) break;
        OutT* o = out + (((int64_t)b * H + hr) * W + w_base + wl0) * CT + c_off + 2 * cp;
#pragma unroll
        for (int p = 0; p < PW; ++p) {
          const float* pp = my_part + (r * 8 + p) * PSTR;
          float s1, s2;
          {
            const float4 a = *reinterpret_cast<const float4*>(pp), c4 = *reinterpret_cast<const float4*>(pp + PSTR / 2);
            s1 = NPT == 3 ? (a.x + a.y) + a.z : (a.x + a.y) + (a.z + a.w);
            s2 = NPT == 3 ? (c4.x + c4.y) + c4.z : (c4.x + c4.y) + (c4.z + c4.w);
            if (NPT > 4) {
              const float4 a2 = *reinterpret_cast<const float4*>(pp + 4), c2 = *reinterpret_cast<const float4*>(pp + PSTR / 2 + 4);
              s1 += NPT == 6 ? a2.x + a2.y : (a2.x + a2.y) + (a2.z + a2.w);
              s2 += NPT == 6 ? c2.x + c2.y : (c2.x + c2.y) + (c2.z + c2.w);
            }
          }
          const float mean = s1 * (1.f / CT);
          const float var = fmaxf(fmaf(s2, 1.f / CT, -mean * mean), 0.f);
          const float rstd = rsqrtf(var + kLnEps);
          const float nmr = -mean * rstd;
          const float2 a = r ? acc1[p] : acc0[p];
          const float2 t = __ffma2_rn(a, make_float2(rstd, rstd), make_float2(nmr, nmr));
          const float2 y = __ffma2_rn(t, gm, be);
          if constexpr (sizeof(OutT) == 2) {
            *reinterpret_cast<act16x2*>(o + (int64_t)p * CT) = floats2act2(y.x, y.y);
          } else {
            *reinterpret_cast<float2*>(o + (int64_t)p * CT) = y;
          }
        }
      }
      sbase += 4;
      if (sbase >= kNS) sbase -= kNS;
    }
  }
  if (CS > 1) {  // no CTA of the pair leaves while pushes into it could still be in flight
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

template <int C, int CS, int W, int TW, typename OutT>
int launch_t(const float* x, int batch, int h, const float* w_t, const float* bias, const float* ln_g, const float* ln_b,
             OutT* out, cudaStream_t stream) {
  using Cfg = DwCfg<C, CS, W, TW, OutT>;
  static_assert(Cfg::NPT == 3 || Cfg::NPT == 4 || Cfg::NPT == 6 || Cfg::NPT == 8, "partial layout");
  CUtensorMap map_x;
  if (int rc = tc_make_map_nhwc_f32(&map_x, x, batch, h, W, Cfg::CT, C, Cfg::RW)) return rc;
  auto kern = dwconv_ln_tma_kernel<C, CS, W, TW, OutT>;
  static bool attr_set = false;
  if (!attr_set) {
    CNB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  const int pq = (int)ceil_div(h, 4);
  const int total = batch * Cfg::STRIPS * pq;
  const int max_groups = sm_budget() * Cfg::CTAS_PER_SM / CS;  // one CTA (or CTA pair) per SM (pair); two for the narrow strips
  int groups = total < max_groups ? total : max_groups;
  const int quota = (int)ceil_div(total, groups);
  groups = (int)ceil_div(total, quota);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(groups * CS));
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CNB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, map_x, h, pq, total, quota, w_t, bias, ln_g, ln_b, out));
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace

// set by the streaming host API while it enqueues an encoder that will share the GPU with the previous batch's decoder
static std::atomic<int> g_sm_budget{kNumSMs};
void set_sm_budget(int sms) { g_sm_budget.store(sms < 8 ? 8 : sms > kNumSMs ? kNumSMs : sms, std::memory_order_relaxed); }
int sm_budget() { return g_sm_budget.load(std::memory_order_relaxed); }
static std::atomic<bool> g_overlap_hint{false};
void dwconv_set_overlap_hint(bool on) { g_overlap_hint.store(on, std::memory_order_relaxed); }

template <typename OutT>
int launch_dwconv_ln_tma(const float* x, int batch, int h, int w, int c, const float* w_t, const float* bias,
                         const float* ln_g, const float* ln_b, OutT* out, cudaStream_t stream) {
  if (c == 96 && w == 56) {
    // Stage 1 has two shapes.  28-pixel strips, one 384-thread CTA per SM: fastest alone (0.249 ms per launch at 64 clips).
    // 14-pixel strips, two 192-thread CTAs per SM: 11 % slower alone (wider halo share) but 296 half-size CTAs leave less
    // wave quantisation when the previous batch's decoder holds 104 of the 148 SMs -- the streaming API sets the hint
    // (measured: streaming step 9.21 -> 9.17 ms, end to end 66.5 k -> 67.6 k audio-s/s).  CNB_DW_S1_NARROW=0/1 forces one.
    static const int forced = [] { const char* e = getenv("CNB_DW_S1_NARROW"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    const bool narrow = forced >= 0 ? forced == 1 : g_overlap_hint.load(std::memory_order_relaxed);
    if (narrow) return launch_t<96, 1, 56, 14, OutT>(x, batch, h, w_t, bias, ln_g, ln_b, out, stream);
    return launch_t<96, 1, 56, 28, OutT>(x, batch, h, w_t, bias, ln_g, ln_b, out, stream);
  }
  if (c == 192 && w == 28) return launch_t<192, 1, 28, 14, OutT>(x, batch, h, w_t, bias, ln_g, ln_b, out, stream);
  if (c == 384 && w == 14) return launch_t<192, 2, 14, 14, OutT>(x, batch, h, w_t, bias, ln_g, ln_b, out, stream);
  return 1;
}
template int launch_dwconv_ln_tma<float>(const float*, int, int, int, int, const float*, const float*, const float*,
                                         const float*, float*, cudaStream_t);
template int launch_dwconv_ln_tma<act16>(const float*, int, int, int, int, const float*, const float*, const float*,
                                                 const float*, act16*, cudaStream_t);

}  // namespace cnb
