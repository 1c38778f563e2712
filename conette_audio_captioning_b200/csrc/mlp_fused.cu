// Fused ConvNeXt pointwise MLP for stage 1 (C = 96): x += scale * (W2 . GELU(W1 . y + b1) + b2) in ONE kernel
// (reference convnext.py:66-73: pwconv1 -> GELU -> pwconv2 -> layer scale -> residual).
//
// Why: as two GEMM launches the stage-1 MLP is HBM-bound on the hidden activations -- pw1 writes (M, 384) fp16 = 693 MB per
// block at 64 clips, pw2 reads it back (ncu: pw2 at 81 % of the DRAM peak).  Here the hidden tile never leaves the SM and
// never even reaches shared memory:
//   * both weight matrices live in shared memory for the whole launch (72 KB each), loaded once per CTA; K = 96 is held as
//     one 64-wide k-block (128B swizzle) plus one 32-wide k-block (64B swizzle), so nothing is padded;
//   * tensor memory holds the whole 128 x 384 fp32 hidden tile (six 64-column chunks, columns 0..383) plus the 128 x 96 output
//     accumulator (columns 384..479).  GEMM1 of tile i+1 is issued chunk by chunk right behind GEMM2 of tile i, so the
//     epilogue never waits for a GEMM1;
//   * per chunk the epilogue warps load the fp32 accumulator, add the bias, apply GELU and write the fp16 result back over
//     the first half of the columns they have just read (tcgen05.st, two fp16 per 32-bit column).  GEMM2 takes that as its A
//     operand straight from tensor memory (tcgen05.mma with a TMEM A operand, the layout of CUTLASS' SM100_MMA_F16BF16_TS),
//     so there is no hidden tile in shared memory, no proxy fence and no buffer hand-back: tcgen05.mma instructions of one
//     thread execute in issue order, which is all the protection the in-place reuse needs;
//   * the fp32 residual tile has its own 48 KB staging buffer: the producer warp TMA-loads it a tile ahead, the epilogue
//     updates it in place and the producer warp TMA-stores it (it alone waits for the store to drain).
// The 16 epilogue warps form two groups (chunks j = g, g + 2, g + 4) that only meet at the output phase; every mbarrier sees
// one arrival per warp.
// HBM traffic per block: y 173 MB + x 347 MB read + 347 MB written = 867 MB instead of 2 253 MB.
// Warp roles: warp 0 TMA producer (weights, A tiles), warp 1 MMA issuer, warps 2..17 epilogue, warp 18 residual in / output out.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "tc_epi.cuh"
#include "tc_ptx.cuh"

namespace cnb {

namespace {

constexpr int kC = 96, kHid = 384, kBM = 128, kCh = 64, kNCh = kHid / kCh;  // 6 hidden chunks of 64
constexpr int kThreadsF = 64 + 32 * kEpiWarps + 32;   // producer, MMA issuer, 16 epilogue warps, residual/output warp
constexpr int kW1aBytes = kHid * 128;           // W1 k-block 0: [384 rows x 64 fp16], 128B swizzle
constexpr int kW1bBytes = kHid * 64;            // W1 k-block 1: [384 rows x 32 fp16], 64B swizzle
constexpr int kW2Bytes = kNCh * kC * 128;       // six k-blocks of [96 rows x 128 B]
constexpr int kA0Bytes = kBM * 128, kA1Bytes = kBM * 64;
constexpr int kStgBytes = kBM * kC * 4;         // fp32 residual / output tile: 3 boxes of [128 rows x 32 columns], 128B swizzle
constexpr int kOffW1a = 0;
constexpr int kOffW1b = kOffW1a + kW1aBytes;
constexpr int kOffW2 = kOffW1b + kW1bBytes;
constexpr int kOffA0 = kOffW2 + kW2Bytes;
constexpr int kOffA1 = kOffA0 + kA0Bytes;
constexpr int kOffStg = kOffA1 + kA1Bytes;
constexpr int kOffVec = kOffStg + kStgBytes;    // bias1 (384) | bias2 (96) | scale (96)
constexpr int kOffBar = kOffVec + (kHid + 2 * kC) * 4;
constexpr int kSmemF = kOffBar + 256 + 1024;
static_assert(kOffW1b % 1024 == 0 && kOffW2 % 1024 == 0 && kOffA0 % 1024 == 0 && kOffA1 % 1024 == 0 && kOffStg % 1024 == 0, "align");
static_assert(kSmemF <= 232448, "shared memory budget");
constexpr int kTmemColsF = 512;                 // hidden chunks 0..383 | O 384..479
constexpr int kTmemO = kHid;
constexpr int kN1 = 192;                         // GEMM1 instruction width: three hidden chunks
constexpr int kTraceTiles = 8, kTraceRoles = 5;   // roles: A producer, MMA issuer, epilogue group 0 / 1 (one warp each), output warp

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D (tmem) (+)= A (tmem: row = lane, two fp16 per 32-bit column, K-major) . B (smem descriptor)^T
__device__ __forceinline__ void tcgen05_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                    uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kReduce: the residual add is done by the TMA store itself (cp.reduce.async.bulk ... .add.f32 into x): no residual load
template <bool kReduce>
__global__ void __launch_bounds__(kThreadsF, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap map_y0, const __grid_constant__ CUtensorMap map_y1,
                 const __grid_constant__ CUtensorMap map_w1a, const __grid_constant__ CUtensorMap map_w1b,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_x, int M,
                 const float* __restrict__ b1, const float* __restrict__ b2, const float* __restrict__ scale,
                 unsigned long long* __restrict__ trace) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  float* s_b1 = reinterpret_cast<float*>(sm + kOffVec);
  float* s_b2 = s_b1 + kHid;
  float* s_sc = s_b2 + kC;
  const uint32_t bars = base + kOffBar;
  const uint32_t w_full = bars, a_full = bars + 8, a_empty = bars + 16, o_full = bars + 24, o_empty = bars + 32;
  const uint32_t resid_full = bars + 40, stg_ready = bars + 48;
  auto d1_full = [&](int hf) { return bars + 64 + 8u * hf; }; // GEMM1 of chunks 3 hf .. 3 hf + 2 has completed (once per tile)
  auto h_full = [&](int j) { return bars + 112 + 8u * j; };   // the fp16 hidden chunk j is in tensor memory (once per tile)
  const uint32_t tmem_slot = bars + 160;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + kOffBar + 160);

  for (int i = threadIdx.x; i < kHid; i += kThreadsF) s_b1[i] = b1[i];
  for (int i = threadIdx.x; i < kC; i += kThreadsF) {
    s_b2[i] = b2[i];
    s_sc[i] = scale[i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (M + kBM - 1) / kBM;
  // optional timeline of CTA 0 (CNB_MLP_TRACE=1): trace[(role * kTraceTiles + tile) * 16 + event] = clock64()
  auto mark = [&](int role, int tile_it, int ev) {
    if (trace != nullptr && blockIdx.x == 0 && tile_it < kTraceTiles) trace[(role * kTraceTiles + tile_it) * 16 + ev] = clock64();
  };

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, kEpiWarps);
    mbar_init(resid_full, 1);
    mbar_init(stg_ready, kEpiWarps);
    for (int j = 0; j < kNCh; ++j) {
      if (j < 2) mbar_init(d1_full(j), 1);
      mbar_init(h_full(j), kEpiWarps / 2);   // one arrival per warp of the group that owns chunk j
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
    // ===================== TMA producer: weights once, then the A tiles =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, kW1aBytes + kW1bBytes + kW2Bytes);
      for (int half = 0; half < 2; ++half) {  // TMA boxes hold at most 256 rows: 384 = 2 x 192
        tma_load_2d(base + kOffW1a + half * (192 * 128), &map_w1a, 0, half * 192, w_full);
        tma_load_2d(base + kOffW1b + half * (192 * 64), &map_w1b, 64, half * 192, w_full);
      }
      for (int j = 0; j < kNCh; ++j) tma_load_2d(base + kOffW2 + j * (kC * 128), &map_w2, j * 64, 0, w_full);
      // A tiles: each is requested as soon as the GEMM1s of the previous tile have released the buffer (a whole tile before
      // the MMA warp needs it) and the one after it is pulled into L2
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        mbar_wait(a_empty, (it & 1) ^ 1);
        mark(0, it, 0);
        mbar_expect_tx(a_full, kA0Bytes + kA1Bytes);
        tma_load_2d(base + kOffA0, &map_y0, 0, t * kBM, a_full);
        tma_load_2d(base + kOffA1, &map_y1, 64, t * kBM, a_full);
        if (t + (int)gridDim.x < n_tiles) {
          tma_prefetch_2d(&map_y0, 0, (t + (int)gridDim.x) * kBM);
          tma_prefetch_2d(&map_y1, 64, (t + (int)gridDim.x) * kBM);
        }
      }
    }
  } else if (warp == kEpiWarps + 2) {
    // ===================== residual / output warp: the fp32 tile in (a tile ahead), the updated tile out =====================
    // (its own warp: chained behind the A loads, the wait for the previous tile's output delayed the next A tile, the MMA
    // warp stalled on it and every GEMM2 of the tile queued up behind that stall)
    if (lane == 0) {
      int it = 0, t_prev = -1;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        if (it > 0) {
          mbar_wait(stg_ready, (it & 1) ^ 1);  // the epilogue has written the previous tile's output into the staging buffer
          mark(4, it, 0);
#pragma unroll
          for (int bx = 0; bx < 3; ++bx) {
            if (kReduce) tma_reduce_add_2d(&map_x, base + kOffStg + bx * (kBM * 128), 32 * bx, t_prev * kBM);
            else tma_store_2d(&map_x, base + kOffStg + bx * (kBM * 128), 32 * bx, t_prev * kBM);
          }
          bulk_commit();
          bulk_wait_read<0>();                 // the store has read the buffer: it may be refilled
          mark(4, it, 1);
        }
        if (kReduce) {
          mbar_arrive(resid_full);   // the staging buffer is free
        } else {
          mbar_expect_tx(resid_full, kStgBytes);
#pragma unroll
          for (int bx = 0; bx < 3; ++bx) tma_load_2d(base + kOffStg + bx * (kBM * 128), &map_x, 32 * bx, t * kBM, resid_full);
        }
        if (t + (int)gridDim.x < n_tiles) {
#pragma unroll
          for (int bx = 0; bx < 3; ++bx) tma_prefetch_2d(&map_x, 32 * bx, (t + (int)gridDim.x) * kBM);
        }
        t_prev = t;
      }
      if (it > 0) {
        mbar_wait(stg_ready, (it & 1) ^ 1);
#pragma unroll
        for (int bx = 0; bx < 3; ++bx) {
          if (kReduce) tma_reduce_add_2d(&map_x, base + kOffStg + bx * (kBM * 128), 32 * bx, t_prev * kBM);
          else tma_store_2d(&map_x, base + kOffStg + bx * (kBM * 128), 32 * bx, t_prev * kBM);
        }
        bulk_commit();
        bulk_wait_all();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the control flow on values the compiler can prove warp-uniform (shuffled from lane 0), one elected
    // lane issues: descriptors then live in uniform registers and the MMAs go out back to back.  (Under `if (lane == 0)` every
    // operand went through a vector -> uniform register waterfall loop, ~90 cycles per MMA: at N = 64 / 96 that was twice the
    // time the tensor core needs for the instruction, and the issue rate paced the kernel.)
    {
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sb = __shfl_sync(0xffffffffu, base, 0);
      const uint32_t ubars = sb + kOffBar;
      constexpr uint32_t idesc1 = make_idesc(kBM, kN1), idesc2 = make_idesc(kBM, kC);
      mbar_wait(w_full, 0);
      tcgen05_fence_after();
      // GEMM1 in two halves of 192 columns (three hidden chunks per instruction): 12 MMAs per tile instead of 36
      auto gemm1 = [&](int hf) {   // hidden columns [192 hf, 192 hf + 192) (fp32) = A (128 x 96) . W1[192 hf ..]^T
        if (elect_one()) {
          const uint32_t d = tb + (uint32_t)(hf * kN1);
          {
            const uint64_t adesc = make_smem_desc(sb + kOffA0);
            const uint64_t bdesc = make_smem_desc(sb + kOffW1a + hf * (kN1 * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k) tcgen05_mma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc1, k != 0);
          }
          {
            const uint64_t adesc = make_smem_desc_sw64(sb + kOffA1);
            const uint64_t bdesc = make_smem_desc_sw64(sb + kOffW1b + hf * (kN1 * 64));
#pragma unroll
            for (int k = 0; k < 2; ++k) tcgen05_mma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc1, 1u);
          }
          tcgen05_commit(ubars + 64 + 8u * hf);   // d1_full(hf)
          if (hf == 1) tcgen05_commit(ubars + 16);  // a_empty: the last GEMM1 of the tile has been issued
        }
        __syncwarp();
      };
      if ((int)blockIdx.x < n_tiles) {
        mbar_wait(a_full, 0);
        tcgen05_fence_after();
        gemm1(0);
        gemm1(1);
      }
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const bool has_next = t + (int)gridDim.x < n_tiles;
#pragma unroll 1
        for (int j = 0; j < kNCh; ++j) {
          mbar_wait(ubars + 112 + 8u * j, it & 1);  // h_full(j): the epilogue has written hidden chunk j (fp16) into tensor memory
          mark(1, it, j);
          if (j == 0) mbar_wait(o_empty, (it & 1) ^ 1);  // the previous tile's output accumulator has been read
          if (j == 0) mark(1, it, 6);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t bdesc = make_smem_desc(sb + kOffW2 + j * (kC * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)   // k-steps 0,1 come from the columns written by the half-0 warps, 2,3 from half 1
              tcgen05_mma_f16_ts(tb + kTmemO, tb + (uint32_t)(j * kCh + (k >> 1) * 32 + (k & 1) * 8), bdesc + 2 * k, idesc2,
                                  (j | k) != 0);
            if (j == kNCh - 1) tcgen05_commit(ubars + 24);   // o_full
          }
          __syncwarp();
          if (has_next && (j == 2 || j == kNCh - 1)) {
            // GEMM1 of the next tile overwrites the hidden columns of three chunks: it is issued behind the GEMM2s that read
            // them.  The first half goes out in mid-tile (the epilogue starts the next tile's first chunk before this tile's
            // output phase), the second half after o_full so that o_full is not queued behind it.
            if (j == 2) {
              mbar_wait(a_full, (it + 1) & 1);
              tcgen05_fence_after();
              mark(1, it, 7);
            }
            gemm1(j / 3);
          }
          if (j == kNCh - 1) mark(1, it, 8);
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int lane_grp = warp & 3;            // TMEM lanes [32 lane_grp, +32)
    const int sub = (warp - 2) >> 2;          // 24-column slice of the output
    // two groups of 8 warps (two per TMEM lane quarter): group g handles the chunks j = g, g + 2, g + 4 of every tile, each
    // thread 32 of the chunk's 64 columns
    const int grp = sub & 1, half = sub >> 1;
    const int row = lane_grp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
    const bool tr = (warp == 2 || warp == 6) && lane == 0;   // traced threads: half 0 of groups 0 and 1
    auto do_chunk = [&](int cit, int k) {   // chunk j = 2 k + grp of the tile with iteration index cit
      const int j = 2 * k + grp;
      const uint32_t taddr = lane_addr + (uint32_t)(j * kCh + 32 * half);
      mbar_wait(d1_full(j / 3), cit & 1);   // GEMM1 runs in two 192-column halves
      if (tr) mark(2 + grp, cit, 2 * k);
      tcgen05_fence_after();
      float v[32];
      tmem_ld_32x16(taddr, v);
      tmem_ld_32x16(taddr + 16, v + 16);
      tmem_ld_wait();
      float2* v2 = reinterpret_cast<float2*>(v);
      const float2* sb2 = reinterpret_cast<const float2*>(s_b1 + j * kCh + 32 * half);
      uint32_t h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 g = gelu_tanh_fit2(__fadd2_rn(v2[i], sb2[i]));
        act16x2 p = floats2act2(g.x, g.y);   // k = 2 i in the low half, 2 i + 1 in the high half
        h[i] = *reinterpret_cast<uint32_t*>(&p);
      }
      tmem_st_32x16(taddr, h);   // over the first 16 of the 32 columns this thread has just read
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(h_full(j));
      if (tr) mark(2 + grp, cit, 2 * k + 1);
    };
    // Software pipeline across the tile boundary: the first chunk of tile i+1 is processed BEFORE the output phase of tile i,
    // so the wait for the last GEMM2 (o_full: wake-up of the MMA warp, 8 MMAs, completion, ~1 200 cycles) is covered by work.
    if ((int)blockIdx.x < n_tiles) do_chunk(0, 0);
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
#pragma unroll 1
      for (int k = 1; k < kNCh / 2; ++k) do_chunk(it, k);
      if (t + (int)gridDim.x < n_tiles) do_chunk(it + 1, 0);
      // ---- output: O (128 x 96) + bias2, layer scale, residual ----
      mbar_wait(o_full, it & 1);
      if (tr) mark(2 + grp, it, 6);
      tcgen05_fence_after();
      float o[24];
      tmem_ld_32x16(lane_addr + (uint32_t)(kTmemO + 24 * sub), o);
      tmem_ld_32x8(lane_addr + (uint32_t)(kTmemO + 24 * sub + 16), o + 16);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      mbar_wait(resid_full, it & 1);   // the residual rows of this tile have landed in the staging buffer
      if (tr) mark(2 + grp, it, 7);
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const int c = 24 * sub + 4 * q;  // first of four consecutive output columns
        const uint32_t addr = base + kOffStg + (uint32_t)(c >> 5) * (kBM * 128) + (uint32_t)row * 128u +
                              (uint32_t)((((c & 31) >> 2) ^ (row & 7)) << 4);
        const float4 xr = kReduce ? make_float4(0.f, 0.f, 0.f, 0.f) : ld_shared_v4(addr);
        const float4 bb = *reinterpret_cast<const float4*>(s_b2 + c);
        const float4 ss = *reinterpret_cast<const float4*>(s_sc + c);
        const float r0 = fmaf(ss.x, o[4 * q + 0] + bb.x, xr.x), r1 = fmaf(ss.y, o[4 * q + 1] + bb.y, xr.y);
        const float r2 = fmaf(ss.z, o[4 * q + 2] + bb.z, xr.z), r3 = fmaf(ss.w, o[4 * q + 3] + bb.w, xr.w);
        st_shared_v4(addr, __float_as_uint(r0), __float_as_uint(r1), __float_as_uint(r2), __float_as_uint(r3));
      }
      fence_async_smem();   // generic-proxy writes -> visible to the TMA store issued by the producer warp
      __syncwarp();
      if (lane == 0) mbar_arrive(stg_ready);
      if (tr) mark(2 + grp, it, 8);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemColsF) : "memory");
  }
}

}  // namespace

// y (M, 96) fp16, w1 (384, 96) fp16, w2 (96, 384) fp16, x (M, 96) fp32 updated in place
int launch_mlp_fused_c96(const act16* y, const act16* w1, const act16* w2, const float* b1, const float* b2,
                         const float* scale, float* x, int m, cudaStream_t stream) {
  if (m == 0) return 0;
  CUtensorMap map_y0, map_y1, map_w1a, map_w1b, map_w2, map_x;
  if (int rc = tc_make_map_f16_box(&map_y0, y, m, kC, kBM, 64)) return rc;
  if (int rc = tc_make_map_f16_box(&map_y1, y, m, kC, kBM, 32)) return rc;
  if (int rc = tc_make_map_f16_box(&map_w1a, w1, kHid, kC, 192, 64)) return rc;
  if (int rc = tc_make_map_f16_box(&map_w1b, w1, kHid, kC, 192, 32)) return rc;
  if (int rc = tc_make_map(&map_w2, w2, kC, kHid, kC, 2)) return rc;
  if (int rc = tc_make_map(&map_x, x, m, kC, kBM, 4)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    CNB_CUDA_OK(cudaFuncSetAttribute(mlp_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemF));
    CNB_CUDA_OK(cudaFuncSetAttribute(mlp_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemF));
    attr_set = true;
  }
  const int n_tiles = (int)ceil_div(m, kBM);
  const int grid = n_tiles < sm_budget() ? n_tiles : sm_budget();
  static const bool use_reduce = getenv("CNB_MLP_NO_REDUCE") == nullptr;   // default: residual add by the TMA store (6 % faster)
  static const bool want_trace = getenv("CNB_MLP_TRACE") != nullptr;
  static int traced = 0;
  unsigned long long* trace = nullptr;
  constexpr size_t kTraceWords = (size_t)kTraceRoles * kTraceTiles * 16;
  if (want_trace && traced < 2 && n_tiles >= 148 * kTraceTiles) {
    CNB_CUDA_OK(cudaMalloc(&trace, kTraceWords * 8));
    CNB_CUDA_OK(cudaMemsetAsync(trace, 0, kTraceWords * 8, stream));
  }
  if (use_reduce) mlp_fused_kernel<true><<<grid, kThreadsF, kSmemF, stream>>>(map_y0, map_y1, map_w1a, map_w1b, map_w2, map_x, m, b1, b2, scale, trace);
  else mlp_fused_kernel<false><<<grid, kThreadsF, kSmemF, stream>>>(map_y0, map_y1, map_w1a, map_w1b, map_w2, map_x, m, b1, b2, scale, trace);
  if (trace != nullptr) {   // debugging aid: timeline of CTA 0 (cycles relative to the first event), printed to stderr
    std::vector<unsigned long long> h(kTraceWords);
    CNB_CUDA_OK(cudaStreamSynchronize(stream));
    CNB_CUDA_OK(cudaMemcpy(h.data(), trace, kTraceWords * 8, cudaMemcpyDeviceToHost));
    cudaFree(trace);
    ++traced;
    unsigned long long t0 = ~0ull;
    for (auto v : h) if (v != 0 && v < t0) t0 = v;
    static const char* names[kTraceRoles] = {"A-producer", "mma", "epi-g0", "epi-g1", "out-warp"};
    for (int r = 0; r < kTraceRoles; ++r)
      for (int ti = 0; ti < kTraceTiles; ++ti) {
        fprintf(stderr, "mlp-trace %-10s tile %d:", names[r], ti);
        for (int e = 0; e < 9; ++e) {
          const unsigned long long v = h[((size_t)r * kTraceTiles + ti) * 16 + e];
          if (v) fprintf(stderr, " e%d=%llu", e, v - t0); else fprintf(stderr, " e%d=-", e);
        }
        fprintf(stderr, "\n");
      }
  }
  CNB_LAUNCH_OK();
  return 0;
}

}  // namespace cnb
