// Fused ConvNeXt pointwise MLP for stages 2 and 3 (C = 192 / 384): x += scale * (W2 . GELU(W1 . y + b1) + b2) in ONE kernel
// (reference convnext.py:66-73: pwconv1 -> GELU -> pwconv2 -> layer scale -> residual).
//
// Why: as two GEMM launches the MLP writes the (M, 4C) fp16 hidden tensor and reads it back (stage 2: 347 MB per block at 64
// clips; pw2 sat at 92 % of the HBM rate and 43 % of the tensor rate).  Unlike stage 1 (mlp_fused.cu) neither the weights (590 KB /
// 2.4 MB) nor the hidden tile (768 / 1536 fp32 columns) fit on an SM, so:
//   * a CTA PAIR (one TPC) owns a 256-row tile, 128 rows per CTA, and runs every MMA as tcgen05.mma cta_group::2 (issued by the
//     even CTA): each CTA loads only HALF of every weight chunk, which halves the L2 -> shared-memory weight stream per row;
//   * the hidden dimension is walked in chunks of 64 units.  Chunk j: GEMM1 D1[j % kD1] (128 x 64 fp32 in tensor memory) =
//     y . W1[j]^T; the 16 epilogue warps add the bias, apply GELU and write the fp16 result back over the columns they have just
//     read (tcgen05.st); GEMM2 takes that as its A operand straight from tensor memory and accumulates O (128 x C fp32, tensor
//     memory) += H_j . W2[:, j]^T (C = 384: two N = 192 instructions per k16 step).  GEMM2 trails GEMM1 by kD1 - 1 chunks, so the
//     tensor pipe has work while the epilogue (tensor-memory load, GELU at the MUFU rate, store, an arrival that crosses the
//     pair: ~1 100 cycles) runs; tcgen05.mma instructions execute in issue order, which protects the in-place reuse;
//   * W1 half-chunks and W2 half-chunks stream through two separate TMA rings (a W1 slot is free as soon as its GEMM1 has
//     completed, a W2 slot only after the trailing GEMM2), the output leaves through TMA reduce-add stores into the fp32 residual
//     stream (no residual load) through a small staging buffer.
//   C = 192: y double buffered, 4 + 4 ring slots of 12 KB, three D1 buffers, O preloaded into registers, 64 columns per store pass.
//   C = 384: O fills 384 of the 512 tensor-memory columns (two D1 buffers), ONE 96 KB y buffer (the next tile is requested when
//            the last GEMM1 has read it, behind the last two GEMM2s and the output phase), 2 + 2 ring slots of 24 KB, 32 columns
//            per store pass read from tensor memory pass by pass.
// Warp roles (both CTAs): warp 0 TMA producer, warp 1 TMEM allocator (+ MMA issuer in the even CTA), warps 2..17 epilogue,
// warp 18 output stores.  Barriers that both CTAs feed (operand bytes, "hidden chunk written", "output read") live in the even CTA.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tc_epi.cuh"
#include "tc_ptx.cuh"

namespace cnb {

namespace {

constexpr int kBM2 = 128, kCh2 = 64;                   // rows per CTA, hidden units per chunk
constexpr int kThreads2 = 64 + 32 * kEpiWarps + 32;     // producer, MMA issuer, 16 epilogue warps, store warp
constexpr int kStgBox = kBM2 * 128;                     // [128 rows x 32 fp32]
constexpr int kTmemCols2 = 512;

template <int C> struct F2Cfg {
  static constexpr int kHid = 4 * C, kNCh = kHid / kCh2, kKB = C / 64, kNSplit = C / 192;
  static constexpr int kYBufs = C == 192 ? 2 : 1, kR1 = C == 192 ? 4 : 2, kR2 = C == 192 ? 4 : 2;
  static constexpr int kStgBoxes = C == 192 ? 2 : 1, kD1 = C == 192 ? 3 : 2, kLag = kD1 - 1;
  static constexpr bool kPreloadO = C == 192;
  static constexpr int kYBytes = kKB * kBM2 * 128;           // kKB boxes of [128 rows x 64 fp16]
  static constexpr int kW1Half = kKB * 32 * 128;             // kKB boxes of [32 rows x 64 fp16] (this CTA's half of the chunk)
  static constexpr int kW2Half = kNSplit * 96 * 128;         // kNSplit boxes of [96 output rows x 64 fp16]
  static constexpr int kPassCols = 32 * kStgBoxes, kPasses = C / kPassCols, kColsPerThread = kPassCols / 4;
  static constexpr int kOffY = 0;
  static constexpr int kOffW1 = kOffY + kYBufs * kYBytes;
  static constexpr int kOffW2 = kOffW1 + kR1 * kW1Half;
  static constexpr int kOffStg = kOffW2 + kR2 * kW2Half;
  static constexpr int kOffVec = kOffStg + kStgBoxes * kStgBox;   // b2 (C) | scale (C); b1 is read through the L1 (no room)
  static constexpr int kOffBar = kOffVec + 2 * C * 4;
  static constexpr int kSmem = kOffBar + 256 + 1024;
  static constexpr int kTmemD1 = C;                          // O 0..C-1 | D1[b] at C + 64 b
  static_assert(C == 192 || C == 384, "stage 2 / stage 3");
  static_assert(kNCh % kD1 == 0, "hidden chunks per buffer");
  static_assert(kOffW1 % 1024 == 0 && kOffW2 % 1024 == 0 && kOffStg % 1024 == 0 && kW1Half % 1024 == 0 && kW2Half % 1024 == 0, "align");
  static_assert(kSmem <= 232448, "shared memory budget");
  static_assert(C + kD1 * kCh2 <= 512, "tensor memory columns");
};
// barrier table (byte offsets from kOffBar; every CTA has its own copy, "lead" = the even CTA's copy is the one that counts)
constexpr uint32_t kBAFull = 0, kBAEmpty = 16, kBW1Full = 32, kBW1Empty = 64, kBW2Full = 96, kBW2Empty = 128, kBD1Full = 160,
                   kBHFull = 184, kBOFull = 208, kBOEmpty = 216, kBStgReady = 224, kBStgFree = 232, kBTmemSlot = 240;

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait2() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D (tmem) (+)= A (tmem: row = lane, two fp16 per 32-bit column, K-major) . B (smem descriptor)^T over the CTA pair
__device__ __forceinline__ void tcgen05_mma_f16_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
mlp_fused_pair_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_w1,
                      const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_x, int M,
                      const float* __restrict__ b1, const float* __restrict__ b2, const float* __restrict__ scale) {
  using F = F2Cfg<C>;
  constexpr int kNCh = F::kNCh, kKB = F::kKB, kR1 = F::kR1, kR2 = F::kR2, kYBufs = F::kYBufs, kD1 = F::kD1, kLag = F::kLag;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  float* s_b2 = reinterpret_cast<float*>(sm + F::kOffVec);
  float* s_sc = s_b2 + C;
  const uint32_t bars = base + F::kOffBar;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + F::kOffBar + kBTmemSlot);

  for (int i = threadIdx.x; i < C; i += kThreads2) {
    s_b2[i] = b2[i];
    s_sc[i] = scale[i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);
  const int n_tiles = (M + 2 * kBM2 - 1) / (2 * kBM2);   // 256-row tiles

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(bars + kBAFull + 8u * b, 1);     // even CTA: y tile b of both CTAs has landed
      mbar_init(bars + kBAEmpty + 8u * b, 1);    // both: the last GEMM1 reading y tile b has completed
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(bars + kBW1Full + 8u * s, 1);    // even CTA: W1 slot s of both CTAs has landed
      mbar_init(bars + kBW1Empty + 8u * s, 1);   // both: the GEMM1 reading W1 slot s has completed
      mbar_init(bars + kBW2Full + 8u * s, 1);
      mbar_init(bars + kBW2Empty + 8u * s, 1);   // both: the GEMM2 reading W2 slot s has completed
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(bars + kBD1Full + 8u * b, 1);               // both: GEMM1 into D1[b] has completed
      mbar_init(bars + kBHFull + 8u * b, 2 * kEpiWarps);    // even CTA: one arrival per epilogue warp of both CTAs
    }
    mbar_init(bars + kBOFull, 1);
    mbar_init(bars + kBOEmpty, 2 * kEpiWarps);
    mbar_init(bars + kBStgReady, kEpiWarps);
    mbar_init(bars + kBStgFree, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {  // the same warp of both CTAs allocates the same columns in both tensor memories
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + kBTmemSlot), "r"(kTmemCols2)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();   // the peer's barriers are initialised before anything is signalled across the pair
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own 128 rows of y, own half of every weight chunk) =====================
    const uint32_t sb = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t ubars = sb + F::kOffBar;
    const uint32_t urank = __shfl_sync(0xffffffffu, rank, 0);
    uint32_t g = 0;   // running weight-chunk counter: chunk g lives in W1 slot g % kR1 and W2 slot g % kR2
    int it = 0;
    for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
      const int b = it % kYBufs;
      mbar_wait(ubars + kBAEmpty + 8u * b, ((uint32_t)(it / kYBufs) & 1u) ^ 1u);   // this CTA's copy
      if (elect_one()) {
        const uint32_t full = ubars + kBAFull + 8u * b;
        const uint32_t lead = mapa_cluster(full, 0);
        if (urank == 0) mbar_expect_tx(full, 2 * F::kYBytes);
#pragma unroll
        for (int kb = 0; kb < kKB; ++kb)
          tma_load_2d_pair(sb + F::kOffY + b * F::kYBytes + kb * (kBM2 * 128), &map_y, kb * 64, t * (2 * kBM2) + (int)urank * kBM2, lead);
      }
      __syncwarp();
      for (int j = 0; j < kNCh; ++j, ++g) {
        const uint32_t s1 = g % kR1, s2 = g % kR2;
        mbar_wait(ubars + kBW1Empty + 8u * s1, ((g / kR1) & 1u) ^ 1u);
        if (elect_one()) {
          const uint32_t full = ubars + kBW1Full + 8u * s1;
          const uint32_t lead = mapa_cluster(full, 0);
          if (urank == 0) mbar_expect_tx(full, 2 * F::kW1Half);
#pragma unroll
          for (int kb = 0; kb < kKB; ++kb)
            tma_load_2d_pair(sb + F::kOffW1 + s1 * F::kW1Half + kb * (32 * 128), &map_w1, kb * 64, j * kCh2 + (int)urank * 32, lead);
        }
        __syncwarp();
        mbar_wait(ubars + kBW2Empty + 8u * s2, ((g / kR2) & 1u) ^ 1u);
        if (elect_one()) {
          const uint32_t full = ubars + kBW2Full + 8u * s2;
          const uint32_t lead = mapa_cluster(full, 0);
          if (urank == 0) mbar_expect_tx(full, 2 * F::kW2Half);
#pragma unroll
          for (int ns = 0; ns < F::kNSplit; ++ns)   // output columns [192 ns + 96 rank, +96) of W2's k-columns [64 j, +64)
            tma_load_2d_pair(sb + F::kOffW2 + s2 * F::kW2Half + ns * (96 * 128), &map_w2, j * kCh2, ns * 192 + (int)urank * 96, lead);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (even CTA only; whole-warp control flow, one elected lane issues) =====================
    if (rank == 0) {
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sb = __shfl_sync(0xffffffffu, base, 0);
      const uint32_t ubars = sb + F::kOffBar;
      constexpr uint32_t idesc1 = make_idesc(2 * kBM2, kCh2), idesc2 = make_idesc(2 * kBM2, 192);
      uint32_t g = 0;
      int it = 0;
      // GEMM2 of chunk jj (running index gj) of tile `cit`: O (+)= H_jj . W2[:, jj]^T
      auto gemm2 = [&](int jj, uint32_t gj, int cit) {
        const int hb = jj % kD1;
        const uint32_t s2 = gj % kR2;
        mbar_wait(ubars + kBHFull + 8u * hb, (uint32_t)(cit * (kNCh / kD1) + jj / kD1) & 1u);   // both epilogues wrote H_jj
        mbar_wait(ubars + kBW2Full + 8u * s2, (gj / kR2) & 1u);
        if (jj == 0) mbar_wait(ubars + kBOEmpty, ((uint32_t)cit & 1u) ^ 1u);                    // previous tile's output read
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a_col = tb + (uint32_t)(F::kTmemD1 + hb * kCh2);
#pragma unroll
          for (int ns = 0; ns < F::kNSplit; ++ns) {
            const uint64_t bdesc = make_smem_desc(sb + F::kOffW2 + s2 * F::kW2Half + ns * (96 * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)   // k16 step k: packed fp16 at columns [16 k, 16 k + 8) of the chunk (written by sub-warp k)
              tcgen05_mma_f16_ts_pair(tb + (uint32_t)(ns * 192), a_col + 16u * k, bdesc + 2 * k, idesc2, (jj | k) != 0);
          }
          tcgen05_commit_pair(ubars + kBW2Empty + 8u * s2);
          if (jj == kNCh - 1) tcgen05_commit_pair(ubars + kBOFull);
        }
        __syncwarp();
      };
      for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
        const int b = it % kYBufs;
        mbar_wait(ubars + kBAFull + 8u * b, (uint32_t)(it / kYBufs) & 1u);
        for (int j = 0; j < kNCh; ++j, ++g) {
          const uint32_t s1 = g % kR1;
          mbar_wait(ubars + kBW1Full + 8u * s1, (g / kR1) & 1u);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t d1 = tb + (uint32_t)(F::kTmemD1 + (j % kD1) * kCh2);
#pragma unroll
            for (int kb = 0; kb < kKB; ++kb) {
              const uint64_t adesc = make_smem_desc(sb + F::kOffY + b * F::kYBytes + kb * (kBM2 * 128));
              const uint64_t bdesc = make_smem_desc(sb + F::kOffW1 + s1 * F::kW1Half + kb * (32 * 128));
#pragma unroll
              for (int k = 0; k < 4; ++k) tcgen05_mma_f16_pair(d1, adesc + 2 * k, bdesc + 2 * k, idesc1, (kb | k) != 0);
            }
            tcgen05_commit_pair(ubars + kBD1Full + 8u * (j % kD1));
            tcgen05_commit_pair(ubars + kBW1Empty + 8u * s1);
            if (j == kNCh - 1) tcgen05_commit_pair(ubars + kBAEmpty + 8u * b);   // y tile b is free
          }
          __syncwarp();
          if (j >= kLag) gemm2(j - kLag, g - kLag, it);   // GEMM2 trails GEMM1 by kLag chunks
        }
#pragma unroll
        for (int r = kLag; r >= 1; --r) gemm2(kNCh - r, g - r, it);
      }
    }
  } else if (warp == kEpiWarps + 2) {
    // ===================== output stores: one pass of kPassCols columns at a time, TMA reduce-add into the residual stream ====
    const uint32_t sb = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t ubars = sb + F::kOffBar;
    const uint32_t urank = __shfl_sync(0xffffffffu, rank, 0);
    uint32_t pc = 0;
    for (int t = pair; t < n_tiles; t += n_pairs) {
      for (int p = 0; p < F::kPasses; ++p, ++pc) {
        mbar_wait(ubars + kBStgReady, pc & 1u);   // the 16 epilogue warps have written this pass
        if (elect_one()) {
#pragma unroll
          for (int bx = 0; bx < F::kStgBoxes; ++bx)
            tma_reduce_add_2d(&map_x, sb + F::kOffStg + bx * kStgBox, F::kPassCols * p + 32 * bx, t * (2 * kBM2) + (int)urank * kBM2);
          bulk_commit();
          bulk_wait_read<0>();             // the stores have read the staging buffer
          mbar_arrive(ubars + kBStgFree);
        }
        __syncwarp();
      }
    }
    if (elect_one()) bulk_wait_all();
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int lane_grp = warp & 3;            // TMEM lanes [32 lane_grp, +32)
    const int sub = (warp - 2) >> 2;          // 16-column slice of a hidden chunk / quarter of an output pass
    const int row = lane_grp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
    const uint32_t lead_h0 = mapa_cluster(bars + kBHFull, 0);
    const uint32_t lead_oe = mapa_cluster(bars + kBOEmpty, 0);
    constexpr int CPT = F::kColsPerThread;    // output columns per thread and pass (16 or 8)
    uint32_t pc = 0;
    int it = 0;
    for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
#pragma unroll 1
      for (int j = 0; j < kNCh; ++j) {
        const int hb = j % kD1;
        mbar_wait(bars + kBD1Full + 8u * hb, (uint32_t)(it * (kNCh / kD1) + j / kD1) & 1u);
        tcgen05_fence_after();
        const uint32_t taddr = lane_addr + (uint32_t)(F::kTmemD1 + hb * kCh2 + 16 * sub);
        float v[16];
        tmem_ld_32x16(taddr, v);
        tmem_ld_wait();
        float2* v2 = reinterpret_cast<float2*>(v);
        const float4* gb = reinterpret_cast<const float4*>(b1 + j * kCh2 + 16 * sub);   // 64 bytes, L1-resident after the first tile
        const float4 q0 = __ldg(gb), q1 = __ldg(gb + 1), q2 = __ldg(gb + 2), q3 = __ldg(gb + 3);
        const float2 bia[8] = {make_float2(q0.x, q0.y), make_float2(q0.z, q0.w), make_float2(q1.x, q1.y), make_float2(q1.z, q1.w),
                               make_float2(q2.x, q2.y), make_float2(q2.z, q2.w), make_float2(q3.x, q3.y), make_float2(q3.z, q3.w)};
        uint32_t h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 gl = gelu_tanh_fit2(__fadd2_rn(v2[i], bia[i]));
          act16x2 pk = floats2act2(gl.x, gl.y);   // k = 2 i in the low half, 2 i + 1 in the high half
          h[i] = *reinterpret_cast<uint32_t*>(&pk);
        }
        tmem_st_32x8(taddr, h);   // over the first 8 of the 16 columns this thread has just read
        tmem_st_wait2();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_h0 + 8u * hb);
      }
      // ---- output: O (128 x C) + bias2, layer scale; the residual add happens in the TMA reduce store ----
      mbar_wait(bars + kBOFull, (uint32_t)it & 1u);
      tcgen05_fence_after();
      float o[F::kPreloadO ? F::kPasses : 1][16];
      if constexpr (F::kPreloadO) {   // few enough columns to hold: the accumulator is handed back before the stores start
#pragma unroll
        for (int p = 0; p < F::kPasses; ++p) tmem_ld_32x16(lane_addr + (uint32_t)(F::kPassCols * p + CPT * sub), o[p]);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_oe);   // the accumulator may be overwritten by the next tile's GEMM2
      }
#pragma unroll
      for (int p = 0; p < F::kPasses; ++p, ++pc) {
        float ov[CPT];
        if constexpr (F::kPreloadO) {
#pragma unroll
          for (int i = 0; i < CPT; ++i) ov[i] = o[p][i];
        } else {
          tmem_ld_32x8(lane_addr + (uint32_t)(F::kPassCols * p + CPT * sub), ov);
          tmem_ld_wait();
          if (p == F::kPasses - 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_oe);
          }
        }
        mbar_wait(bars + kBStgFree, (pc & 1u) ^ 1u);   // the previous pass has left the staging buffer
        const int c0 = F::kPassCols * p + CPT * sub;   // first of this thread's output columns
        const int in_pass = CPT * sub;                 // column offset inside the pass
        const uint32_t box = base + F::kOffStg + (uint32_t)(in_pass >> 5) * kStgBox + (uint32_t)row * 128u;
#pragma unroll
        for (int q = 0; q < CPT / 4; ++q) {
          const int c = c0 + 4 * q;
          const float4 bb = *reinterpret_cast<const float4*>(s_b2 + c);
          const float4 ss = *reinterpret_cast<const float4*>(s_sc + c);
          const int grp16 = ((in_pass & 31) >> 2) + q;   // 16-byte group inside the 128-byte box row
          st_shared_v4(box + (uint32_t)((grp16 ^ (row & 7)) << 4), __float_as_uint(ss.x * (ov[4 * q + 0] + bb.x)),
                       __float_as_uint(ss.y * (ov[4 * q + 1] + bb.y)), __float_as_uint(ss.z * (ov[4 * q + 2] + bb.z)),
                       __float_as_uint(ss.w * (ov[4 * q + 3] + bb.w)));
        }
        fence_async_smem();   // generic-proxy writes -> visible to the TMA store issued by the store warp
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + kBStgReady);
      }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();   // no CTA of the pair leaves (or frees tensor memory) while the other may still signal it
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols2) : "memory");
  }
}

template <int C>
int launch_pair(const act16* y, const act16* w1, const act16* w2, const float* b1, const float* b2, const float* scale, float* x,
                int m, cudaStream_t stream) {
  using F = F2Cfg<C>;
  if (m == 0) return 0;
  CUtensorMap map_y, map_w1, map_w2, map_x;
  if (int rc = tc_make_map_f16_box(&map_y, y, m, C, kBM2, 64)) return rc;
  if (int rc = tc_make_map_f16_box(&map_w1, w1, F::kHid, C, 32, 64)) return rc;
  if (int rc = tc_make_map_f16_box(&map_w2, w2, C, F::kHid, 96, 64)) return rc;
  if (int rc = tc_make_map(&map_x, x, m, C, kBM2, 4)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    CNB_CUDA_OK(cudaFuncSetAttribute(mlp_fused_pair_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, F::kSmem));
    attr_set = true;
  }
  const int n_tiles = (int)ceil_div(m, 2 * kBM2);
  const int pairs = n_tiles < sm_budget() / 2 ? n_tiles : sm_budget() / 2;
  mlp_fused_pair_kernel<C><<<2 * pairs, kThreads2, F::kSmem, stream>>>(map_y, map_w1, map_w2, map_x, m, b1, b2, scale);
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace

// y (M, C) fp16, w1 (4C, C) fp16, w2 (C, 4C) fp16, x (M, C) fp32 updated in place; C = 192 (stage 2) or 384 (stage 3)
int launch_mlp_fused_pair(int c, const act16* y, const act16* w1, const act16* w2, const float* b1, const float* b2,
                          const float* scale, float* x, int m, cudaStream_t stream) {
  if (c == 192) return launch_pair<192>(y, w1, w2, b1, b2, scale, x, m, stream);
  if (c == 384) return launch_pair<384>(y, w1, w2, b1, b2, scale, x, m, stream);
  CNB_REQUIRE(false, "fused pair MLP: C must be 192 or 384");
  return -1;
}

}  // namespace cnb
