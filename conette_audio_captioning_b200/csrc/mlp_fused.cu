// Fused ConvNeXt pointwise MLP for stage 1 (C = 96): x += scale * (W2 . GELU(W1 . y + b1) + b2) in ONE kernel
// (reference convnext.py:66-73: pwconv1 -> GELU -> pwconv2 -> layer scale -> residual).
//
// Why: as two GEMM launches the stage-1 MLP is HBM-bound on the hidden activations -- pw1 writes (M, 384) bf16 = 693 MB per
// block at 64 clips, pw2 reads it back (ncu: pw2 at 81 % of the DRAM peak).  Here the hidden tile never leaves the SM:
//   * both weight matrices live in shared memory for the whole launch (72 KB each), loaded once per CTA; K = 96 is held as
//     one 64-wide k-block (128B swizzle) plus one 32-wide k-block (64B swizzle), so nothing is padded;
//   * per 128-row tile: GEMM1 runs in six 64-column chunks into two alternating TMEM accumulators; the epilogue warps add
//     the bias, apply GELU and write the chunk as bf16 into a 16 KB shared-memory tile in exactly the UMMA K-major 128B-swizzled
//     layout (the staging format of gemm_tc.cu's TMA stores), which GEMM2 consumes as its A operand, accumulating the
//     128 x 96 output in a third TMEM region; the hidden tile is double-buffered so the GELU of chunk j+1 overlaps GEMM2 of j;
//   * the fp32 residual rows are prefetched into registers at the start of the tile (each thread reads 96 contiguous bytes of
//     its row), the updated rows go through a swizzled staging tile (aliasing the hidden buffers) and leave by TMA store;
//     the A operand has its own buffer, so the next tile's loads and first GEMMs run under this tile's output phase.
// HBM traffic per block: y 173 MB + x 347 MB read + 347 MB written = 867 MB instead of 2 253 MB.
// Warp roles as in gemm_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warps 2..17 epilogue.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_epi.cuh"
#include "tc_ptx.cuh"

namespace cnb {

namespace {

constexpr int kC = 96, kHid = 384, kBM = 128, kCh = 64, kNCh = kHid / kCh;  // 6 hidden chunks of 64
constexpr int kThreadsF = 64 + 32 * kEpiWarps;
constexpr int kW1aBytes = kHid * 128;           // W1 k-block 0: [384 rows x 64 bf16], 128B swizzle
constexpr int kW1bBytes = kHid * 64;            // W1 k-block 1: [384 rows x 32 bf16], 64B swizzle
constexpr int kW2Bytes = kNCh * kC * 128;       // six k-blocks of [96 rows x 128 B]
constexpr int kA0Bytes = kBM * 128, kA1Bytes = kBM * 64;
constexpr int kHBytes = kBM * 128;              // one k-block (64 bf16) of the hidden tile
constexpr int kOffW1a = 0;
constexpr int kOffW1b = kOffW1a + kW1aBytes;
constexpr int kOffW2 = kOffW1b + kW1bBytes;
constexpr int kOffA0 = kOffW2 + kW2Bytes;
constexpr int kOffA1 = kOffA0 + kA0Bytes;
constexpr int kOffH = kOffA1 + kA1Bytes;        // H0 | H1 | 16 KB extra = the 48 KB fp32 output staging (3 boxes of 32 columns)
constexpr int kOffVec = kOffH + 3 * kHBytes;    // bias1 (384) | bias2 (96) | scale (96)
constexpr int kOffBar = kOffVec + (kHid + 2 * kC) * 4;
constexpr int kSmemF = kOffBar + 256 + 1024;
static_assert(3 * kHBytes == kBM * kC * 4, "the staging area holds the fp32 output tile exactly");
static_assert(kOffW1b % 1024 == 0 && kOffW2 % 1024 == 0 && kOffA0 % 1024 == 0 && kOffA1 % 1024 == 0 && kOffH % 1024 == 0, "align");
static_assert(kSmemF <= 232448, "shared memory budget");
constexpr int kTmemColsF = 256;                 // D1[0] 0..63 | D1[1] 64..127 | O 128..223

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__global__ void __launch_bounds__(kThreadsF, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap map_y0, const __grid_constant__ CUtensorMap map_y1,
                 const __grid_constant__ CUtensorMap map_w1a, const __grid_constant__ CUtensorMap map_w1b,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_x, int M,
                 const float* __restrict__ b1, const float* __restrict__ b2, const float* __restrict__ scale,
                 const float* __restrict__ x) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  float* s_b1 = reinterpret_cast<float*>(sm + kOffVec);
  float* s_b2 = s_b1 + kHid;
  float* s_sc = s_b2 + kC;
  const uint32_t bars = base + kOffBar;
  const uint32_t w_full = bars, a_full = bars + 8, a_empty = bars + 16, o_full = bars + 24, o_empty = bars + 32;
  auto d1_full = [&](int b) { return bars + 64 + 8u * b; };
  auto d1_empty = [&](int b) { return bars + 80 + 8u * b; };
  auto h_full = [&](int b) { return bars + 96 + 8u * b; };
  auto h_empty = [&](int b) { return bars + 112 + 8u * b; };
  const uint32_t tmem_slot = bars + 128;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + kOffBar + 128);

  for (int i = threadIdx.x; i < kHid; i += kThreadsF) s_b1[i] = b1[i];
  for (int i = threadIdx.x; i < kC; i += kThreadsF) {
    s_b2[i] = b2[i];
    s_sc[i] = scale[i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (M + kBM - 1) / kBM;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, kEpiWarps);
    for (int b = 0; b < 2; ++b) {
      mbar_init(d1_full(b), 1);
      mbar_init(d1_empty(b), kEpiWarps / 2);   // one arrival per warp of the group that owns buffer b
      mbar_init(h_full(b), kEpiWarps / 2);
      mbar_init(h_empty(b), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemColsF) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, kW1aBytes + kW1bBytes + kW2Bytes);
      for (int half = 0; half < 2; ++half) {  // TMA boxes hold at most 256 rows: 384 = 2 x 192
        tma_load_2d(base + kOffW1a + half * (192 * 128), &map_w1a, 0, half * 192, w_full);
        tma_load_2d(base + kOffW1b + half * (192 * 64), &map_w1b, 64, half * 192, w_full);
      }
      for (int j = 0; j < kNCh; ++j) tma_load_2d(base + kOffW2 + j * (kC * 128), &map_w2, j * 64, 0, w_full);
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        mbar_wait(a_empty, (it & 1) ^ 1);  // the GEMM1s of the previous tile have finished reading the A buffer
        mbar_expect_tx(a_full, kA0Bytes + kA1Bytes);
        tma_load_2d(base + kOffA0, &map_y0, 0, t * kBM, a_full);
        tma_load_2d(base + kOffA1, &map_y1, 64, t * kBM, a_full);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(kBM, kCh), idesc2 = make_idesc(kBM, kC);
      mbar_wait(w_full, 0);
      tcgen05_fence_after();
      // every accumulator / hidden buffer is used exactly three times per tile, so the phase parity of the k-th use inside
      // tile `it` is (3 it + k) & 1: no per-buffer counters (runtime-indexed arrays end up in local memory -- an earlier
      // version lost 25 % of its time waiting on those LDLs in the synchronisation path)
      int it = 0;
      auto gemm1 = [&](int j) {   // D1[j & 1] = A (128 x 96) . W1[64 j .. 64 j + 63]^T
        const int b = j & 1;
        mbar_wait(d1_empty(b), ((uint32_t)(3 * it + (j >> 1)) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(b * kCh);
        {
          const uint64_t adesc = make_smem_desc(base + kOffA0);
          const uint64_t bdesc = make_smem_desc(base + kOffW1a + j * (kCh * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) tcgen05_mma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc1, k != 0);
        }
        {
          const uint64_t adesc = make_smem_desc_sw64(base + kOffA1);
          const uint64_t bdesc = make_smem_desc_sw64(base + kOffW1b + j * (kCh * 64));
#pragma unroll
          for (int k = 0; k < 2; ++k) tcgen05_mma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc1, 1u);
        }
        tcgen05_commit(d1_full(b));
      };
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        mbar_wait(a_full, it & 1);
        tcgen05_fence_after();
        gemm1(0);
        gemm1(1);
#pragma unroll
        for (int j = 0; j < kNCh; ++j) {
          const int hb = j & 1;
          mbar_wait(h_full(hb), (uint32_t)(3 * it + (j >> 1)) & 1u);  // the epilogue has written hidden chunk j (bf16, operand layout)
          if (j == 0) mbar_wait(o_empty, (it & 1) ^ 1);  // the previous tile's output accumulator has been read
          tcgen05_fence_after();
          const uint64_t adesc = make_smem_desc(base + kOffH + hb * kHBytes);
          const uint64_t bdesc = make_smem_desc(base + kOffW2 + j * (kC * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) tcgen05_mma_bf16(tmem_base + 2 * kCh, adesc + 2 * k, bdesc + 2 * k, idesc2, (j | k) != 0);
          tcgen05_commit(h_empty(hb));
          if (j == kNCh - 1) tcgen05_commit(o_full);
          if (j + 2 < kNCh) gemm1(j + 2);
          if (j + 2 == kNCh - 1) tcgen05_commit(a_empty);  // the last GEMM1 of the tile has been issued: A is free once it completes
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int lane_grp = warp & 3;            // TMEM lanes [32 lane_grp, +32)
    const int sub = (warp - 2) >> 2;          // 24-column slice of the output
    // the 16 warps form two groups of 8 (two per TMEM lane quarter): group g owns accumulator D1[g] and hidden buffer H[g],
    // i.e. the chunks j = g, g + 2, g + 4 of every tile, each thread taking 32 of the chunk's 64 columns.  The groups only
    // meet at the output phase, so the GELU of one chunk runs under the TMEM load / barrier latency of the other, and every
    // barrier sees one arrival per warp (8) instead of one per thread (512 same-address shared-memory atomics per chunk
    // were the bottleneck of the first version).
    const int grp = sub & 1, half = sub >> 1;
    const bool leader = (warp == 2 && lane == 0);
    const int row = lane_grp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
    const uint32_t my_d1_full = d1_full(grp), my_d1_empty = d1_empty(grp), my_h_full = h_full(grp), my_h_empty = h_empty(grp);
    const uint32_t h_row = base + kOffH + (uint32_t)grp * kHBytes + (uint32_t)row * 128u;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      // residual rows: 24 consecutive floats of this thread's row (96 bytes = three full 32-byte sectors).  They are pulled
      // into L2 now and into registers only after the hidden chunks (holding them across the GELU loop spilled registers);
      // the L2 hit then hides under the wait for the last GEMM2.
      const int64_t grow = (int64_t)t * kBM + row;
      const float4* xp = reinterpret_cast<const float4*>(x + grow * kC + 24 * sub);
      if (grow < M) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(xp));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(xp + 5));
      }
#pragma unroll
      for (int k = 0; k < kNCh / 2; ++k) {
        const int j = 2 * k + grp;
        const uint32_t use = (uint32_t)(3 * it + k);  // this is use number `use` of accumulator / hidden buffer `grp`
        mbar_wait(my_d1_full, use & 1u);
        tcgen05_fence_after();
        float v[32];
        tmem_ld_32x16(lane_addr + (uint32_t)(grp * kCh + 32 * half), v);
        tmem_ld_32x16(lane_addr + (uint32_t)(grp * kCh + 32 * half + 16), v + 16);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(my_d1_empty);
        float2* v2 = reinterpret_cast<float2*>(v);
        const float2* sb2 = reinterpret_cast<const float2*>(s_b1 + j * kCh + 32 * half);
#pragma unroll
        for (int i = 0; i < 16; ++i) v2[i] = gelu_tanh_fit2(__fadd2_rn(v2[i], sb2[i]));
        if (k == 0) {
          // the hidden buffers double as the output staging of the previous tile: its TMA store must have read them
          if (leader) bulk_wait_read<0>();
          epi_bar(1);
        }
        mbar_wait(my_h_empty, (use & 1u) ^ 1u);  // the GEMM2 that last read this hidden buffer has completed
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * q + 0], v[8 * q + 1]);
          __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]);
          __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
          st_shared_v4(h_row + (uint32_t)(((4 * half + q) ^ (row & 7)) << 4), *reinterpret_cast<uint32_t*>(&p0),
                       *reinterpret_cast<uint32_t*>(&p1), *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
        }
        fence_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(my_h_full);
      }
      // ---- output: O (128 x 96) + bias2, layer scale, residual ----
      float4 xr[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) xr[q] = grow < M ? __ldg(xp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      mbar_wait(o_full, it & 1);  // every MMA of the tile has completed: the hidden buffers are dead and become the staging
      tcgen05_fence_after();
      float o[24];
      tmem_ld_32x16(lane_addr + (uint32_t)(2 * kCh + 24 * sub), o);
      tmem_ld_32x8(lane_addr + (uint32_t)(2 * kCh + 24 * sub + 16), o + 16);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const int c = 24 * sub + 4 * q;  // first of four consecutive output columns
        const uint32_t addr = base + kOffH + (uint32_t)(c >> 5) * (kBM * 128) + (uint32_t)row * 128u +
                              (uint32_t)((((c & 31) >> 2) ^ (row & 7)) << 4);
        const float4 bb = *reinterpret_cast<const float4*>(s_b2 + c);
        const float4 ss = *reinterpret_cast<const float4*>(s_sc + c);
        const float r0 = fmaf(ss.x, o[4 * q + 0] + bb.x, xr[q].x), r1 = fmaf(ss.y, o[4 * q + 1] + bb.y, xr[q].y);
        const float r2 = fmaf(ss.z, o[4 * q + 2] + bb.z, xr[q].z), r3 = fmaf(ss.w, o[4 * q + 3] + bb.w, xr[q].w);
        st_shared_v4(addr, __float_as_uint(r0), __float_as_uint(r1), __float_as_uint(r2), __float_as_uint(r3));
      }
      fence_async_smem();
      epi_bar(2);
      if (leader) {
#pragma unroll
        for (int bx = 0; bx < 3; ++bx) tma_store_2d(&map_x, base + kOffH + bx * (kBM * 128), 32 * bx, t * kBM);
        bulk_commit();
      }
    }
    if (leader) bulk_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemColsF) : "memory");
  }
}

}  // namespace

// y (M, 96) bf16, w1 (384, 96) bf16, w2 (96, 384) bf16, x (M, 96) fp32 updated in place
int launch_mlp_fused_c96(const __nv_bfloat16* y, const __nv_bfloat16* w1, const __nv_bfloat16* w2, const float* b1, const float* b2,
                         const float* scale, float* x, int m, cudaStream_t stream) {
  if (m == 0) return 0;
  CUtensorMap map_y0, map_y1, map_w1a, map_w1b, map_w2, map_x;
  if (int rc = tc_make_map_bf16_box(&map_y0, y, m, kC, kBM, 64)) return rc;
  if (int rc = tc_make_map_bf16_box(&map_y1, y, m, kC, kBM, 32)) return rc;
  if (int rc = tc_make_map_bf16_box(&map_w1a, w1, kHid, kC, 192, 64)) return rc;
  if (int rc = tc_make_map_bf16_box(&map_w1b, w1, kHid, kC, 192, 32)) return rc;
  if (int rc = tc_make_map(&map_w2, w2, kC, kHid, kC, 2)) return rc;
  if (int rc = tc_make_map(&map_x, x, m, kC, kBM, 4)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    CNB_CUDA_OK(cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemF));
    attr_set = true;
  }
  const int n_tiles = (int)ceil_div(m, kBM);
  const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  mlp_fused_kernel<<<grid, kThreadsF, kSmemF, stream>>>(map_y0, map_y1, map_w1a, map_w1b, map_w2, map_x, m, b1, b2, scale, x);
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace cnb
