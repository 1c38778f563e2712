// fp16-operand tensor-core GEMM for sm_100a: TMA-fed tcgen05.mma with the accumulator in TMEM and fused epilogues.
//   out[m,n] = epi( sum_k A[m,k] * W[n,k] ),  A (M,K) fp16 and W (N,K) fp16 both K-major, fp32 accumulation.
// This is the ConvNeXt pointwise-MLP / downsample / projection hot loop (reference convnext.py:66-73, :212-217,
// pl_modules/common.py:71-78), ~92 % of the encoder FLOPs.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads of A (128 x 64) and W (BLOCK_N x 64) tiles, 128B swizzle
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16, kind::f16, fp16 in / f32 acc)
//   warps 2..17 epilogue: tcgen05.ld (32 lanes x 32 columns) -> bias / GELU / layer-scale+residual -> global stores
//               (GELU in this 16-bit path = 0.5x(1+tanh(x(a1+a3x^2+a5x^4))), a minimax fit of the erf form, max abs error
//               2.5e-5 + the 2^-11 relative error of tanh.approx -- of the order of the fp16 rounding of the stored hidden)
// CTA pairs (PAIR = true, large M): two CTAs of a 2-CTA cluster (one TPC) work on one 256 x BLOCK_N tile with
// tcgen05.mma.cta_group::2 issued by the even CTA.  Each CTA TMA-loads its own 128 rows of A but only HALF of the W tile
// (BLOCK_N/2 rows), so the L2 -> shared-memory fill per flop drops by 30 % at BLOCK_N = 192 (the fill rate, not the tensor
// pipe, bounded the 1-CTA kernel in stages 3-4: profiles/r1_ncu_gemm_tc_stage3.txt) and the smaller stages allow a deeper
// ring.  Cross-CTA protocol: both producers complete bytes on the LEADER's full barrier; tcgen05.commit multicasts the
// "stage free" / "accumulator ready" arrivals to both CTAs; the epilogue warps of both CTAs arrive on the leader's
// "accumulator drained" barrier; cluster barriers bracket barrier init / TMEM allocation and teardown.
// Pipelines: STAGES-deep smem ring (full/empty mbarriers) and a 2-deep TMEM accumulator ring (tmem_full/tmem_empty) so
// the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>

#include <limits.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"
#include "tc_epi.cuh"

namespace cnb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 fp16 = 128 bytes = one swizzle-128B atom row
constexpr int kUmmaK = 16;
constexpr int kEpiChunk = 16;                      // accumulator columns per tcgen05.ld
constexpr int kMaxN = 3072;                        // bias / layer-scale vectors are staged in shared memory
constexpr int kTcThreads = 64 + 32 * kEpiWarps;   // warp 0 TMA, warp 1 MMA, warps 2.. epilogue

// ---------------------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------------------
// Output path: the epilogue never touches global memory with per-thread row accesses (one thread = one output row would
// make every warp-level store hit 32 different cache lines -- measured: the kernel time scaled with the output bytes at
// ~1.8 TB/s).  Instead 64-column passes are written into 128B-swizzled staging tiles in shared memory and leave through
// TMA bulk tensor stores; the layer-scale/residual epilogue TMA-loads the fp32 residual tile into the same staging tiles
// at tile start and updates it in place.
template <int BLOCK_N, int EPI, typename OutT, bool PAIR = false>
struct TcCfg {
  // layer-scale + residual: by default scale * (acc + bias) leaves through a TMA REDUCE-ADD store into x (the add happens
  // in L2, no residual tile is loaded: 64 KB instead of 96 KB of staging at BLOCK_N = 192, i.e. one more ring stage, and a
  // third less SM <-> L2 traffic in the epilogue).  -DCNB_PW2_RESID_LOAD=1 restores the load / add / store form.
#ifndef CNB_PW2_RESID_LOAD
#define CNB_PW2_RESID_LOAD 0
#endif
  static constexpr bool kResid = (EPI == EPI_SCALE_RESID) && CNB_PW2_RESID_LOAD;
  static constexpr bool kReduce = (EPI == EPI_SCALE_RESID) && !CNB_PW2_RESID_LOAD;
  static constexpr int kElt = (int)sizeof(OutT);
  static constexpr int kBoxCols = 128 / kElt;                 // 64 fp16 / 32 fp32 per 128-byte swizzle row
  static constexpr int kBoxBytes = kBlockM * 128;             // 16 KB
  static constexpr int kBoxesPerPass = 64 / kBoxCols;         // 1 / 2
  static constexpr int kPassBytes = kBoxesPerPass * kBoxBytes;
  static constexpr int kPasses = (BLOCK_N + 63) / 64;
  // residual tiles are staged whole; plain fp16 outputs rotate through three 16 KB tiles (one CTA-wide barrier per pass: the
  // elected lane waits for the store of pass p-1 right after issuing the store of pass p, so whoever has passed the barrier of
  // pass p+1 knows that the tile of pass p-1 = the tile of pass p+2 is free); plain fp32 outputs (32 KB tiles) ping-pong
  static constexpr int kStagingBufs = kResid ? kPasses : (kElt == 2 ? 3 : 2);
  static constexpr bool kOneBar = !kResid && kStagingBufs == 3;
  static constexpr int kVecFloats = (kResid || kReduce) ? 2 * 768 : kMaxN; // bias (| scale) staged in smem
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kTileM = PAIR ? 2 * kBlockM : kBlockM;     // rows of one MMA tile (both CTAs of a pair)
  static constexpr int kBRows = PAIR ? BLOCK_N / 2 : BLOCK_N;     // W rows this CTA loads per k-block
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kFullTx = PAIR ? 2 * kStageBytes : kStageBytes;  // bytes completed on the (leader's) full barrier
  static constexpr int kMaxStages = PAIR ? 6 : 4;
  static constexpr int kFixed = kStagingBufs * kPassBytes + kVecFloats * 4 + 512 /*barriers*/ + 1024 /*align*/;
  static constexpr int kStagesFit = (232448 - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesFit > kMaxStages ? kMaxStages : kStagesFit;
  static_assert(kStages >= 2, "shared memory budget");
  static constexpr int kStagingOff = kStages * kStageBytes;
  static constexpr int kBarOff = kStagingOff + kStagingBufs * kPassBytes;
  static constexpr int kVecOff = kBarOff + 512;
  static constexpr int kTotal = kVecOff + kVecFloats * 4 + 1024;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
};

template <int BLOCK_N, int EPI, typename OutT, bool PAIR>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_resid, int M, int N, int K,
               EpiParams ep) {
  using C = TcCfg<BLOCK_N, EPI, OutT, PAIR>;
  constexpr int kStages = C::kStages;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;            // 0 = leader (issues the MMAs)
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // persistent worker index (a CTA or a CTA pair)
  const int n_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle-128B tiles need 1024-byte alignment
  uint8_t* smem_aligned = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t stg = base + C::kStagingOff;
  const uint32_t bars = base + C::kBarOff;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
  const uint32_t resid_bar = bars + 8u * (2 * kStages + 4);
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 5);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_aligned + C::kBarOff + 8 * (2 * kStages + 5));

  float* s_bias = reinterpret_cast<float*>(smem_aligned + C::kVecOff);
  float* s_scale = s_bias + 768;  // only used by the residual epilogue (N <= 768)
  for (int i = threadIdx.x; i < N; i += kTcThreads) {
    s_bias[i] = ep.bias[i];
    if (C::kResid || C::kReduce) s_scale[i] = ep.scale[i];
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = N / BLOCK_N;
  const int tiles_m = (M + C::kTileM - 1) / C::kTileM;
  const int n_tiles = tiles_m * tiles_n;
  const int k_blocks = (K + kBlockK - 1) / kBlockK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), PAIR ? 2 * kEpiWarps : kEpiWarps);   // one arrival per epilogue warp (of both CTAs)
    }
    mbar_init(resid_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {  // the same warp of both CTAs allocates the same columns in both tensor memories
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(C::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(C::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything is signalled across the pair
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====================
    {
      const uint32_t sb = __shfl_sync(0xffffffffu, base, 0);
      const uint32_t ubars = sb + C::kBarOff;
      int s = 0;
      uint32_t ph = 0;
      const uint32_t urank = __shfl_sync(0xffffffffu, cta_rank, 0);
      for (int t = unit; t < n_tiles; t += n_units) {
        const int m_blk = t / tiles_n, n_blk = t - m_blk * tiles_n;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(ubars + 8u * (kStages + s), ph ^ 1);   // empty (this CTA's own copy)
          if (elect_one()) {
            const uint32_t a_dst = sb + s * C::kStageBytes;
            const uint32_t b_dst = a_dst + C::kABytes;
            const uint32_t full = ubars + 8u * s;
            if (PAIR) {
              const uint32_t lead_full = mapa_cluster(full, 0);
              if (urank == 0) mbar_expect_tx(full, C::kFullTx);   // the bytes of both CTAs land on the leader's barrier
              tma_load_2d_pair(a_dst, &map_a, kb * kBlockK, m_blk * C::kTileM + (int)urank * kBlockM, lead_full);
              tma_load_2d_pair(b_dst, &map_w, kb * kBlockK, n_blk * BLOCK_N + (int)urank * C::kBRows, lead_full);
            } else {
              mbar_expect_tx(full, C::kStageBytes);
              tma_load_2d(a_dst, &map_a, kb * kBlockK, m_blk * kBlockM, full);
              tma_load_2d(b_dst, &map_w, kb * kBlockK, n_blk * BLOCK_N, full);
            }
          }
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pairs: the leader CTA only) =====================
    // the whole warp runs the loop on warp-uniform values, one elected lane issues (see elect_one in tc_ptx.cuh)
    if (cta_rank == 0) {
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sb = __shfl_sync(0xffffffffu, base, 0);
      const uint32_t ubars = sb + C::kBarOff;
      constexpr uint32_t idesc = make_idesc(C::kTileM, BLOCK_N);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int t = unit; t < n_tiles; t += n_units) {
        mbar_wait(ubars + 8u * (2 * kStages + 2 + as), aph ^ 1);  // tempty: the epilogue has drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t tmem_d = tb + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(ubars + 8u * s, ph);   // full
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = sb + s * C::kStageBytes;
            const uint64_t adesc = make_smem_desc(a_addr);
            const uint64_t bdesc = make_smem_desc(a_addr + C::kABytes);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              // advance 16 fp16 = 32 bytes inside the 128-byte swizzle atom: +2 in (addr >> 4) units
              if (PAIR) tcgen05_mma_f16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
              else tcgen05_mma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            }
            if (PAIR) {  // one arrival on the barrier at this offset in both CTAs
              tcgen05_commit_pair(ubars + 8u * (kStages + s));
              if (kb == k_blocks - 1) tcgen05_commit_pair(ubars + 8u * (2 * kStages + as));
            } else {
              tcgen05_commit(ubars + 8u * (kStages + s));  // empty: frees the smem stage once these MMAs have read it
              if (kb == k_blocks - 1) tcgen05_commit(ubars + 8u * (2 * kStages + as));   // tfull: accumulator complete -> epilogue
            }
          }
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        as ^= 1;
        if (as == 0) aph ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int lane_grp = warp & 3;            // TMEM lanes [32*lane_grp, +32) are accessible to this warp
    const int sub = (warp - 2) >> 2;          // the four warps of a lane quarter split every 64-column pass 4 x 16
    // TMA stores / residual loads are issued by one elected lane of warp 2, with warp-uniform operands (see elect_one in
    // tc_ptx.cuh: a plain `lane == 0` branch costs ~90 cycles per TMA instruction, in front of a barrier all 16 warps wait on).
    // Bulk async-groups belong to the issuing thread: every commit / wait below uses the same elected lane.
    const bool lead_warp = (warp == 2);
    const uint32_t ustg = __shfl_sync(0xffffffffu, stg, 0);
    const uint32_t uresid_bar = __shfl_sync(0xffffffffu, resid_bar, 0);
    const int row = lane_grp * 32 + lane;     // row of the tile owned by this thread
    int as = 0;
    uint32_t aph = 0, rph = 0;
    uint32_t pass_ctr = 0;                    // running pass counter: plain outputs ping-pong the two staging tiles
    const uint32_t lead_tempty0 = PAIR ? mapa_cluster(tempty_bar(0), 0) : tempty_bar(0);
    for (int t = unit; t < n_tiles; t += n_units) {
      const int m_blk = t / tiles_n, n_blk = t - m_blk * tiles_n;
      const int m_row0 = m_blk * C::kTileM + (int)cta_rank * kBlockM;   // first output row of this CTA's half of the tile
      if (C::kResid && lead_warp) {
        if (elect_one()) {
          // previous tile's stores must have finished reading the staging tiles; then fetch this tile's residual rows
          bulk_wait_read<0>();
          int n_boxes = 0;
#pragma unroll
          for (int i = 0; i < C::kPasses; ++i)
#pragma unroll
            for (int bx = 0; bx < C::kBoxesPerPass; ++bx)
              if (64 * i + bx * C::kBoxCols < BLOCK_N) ++n_boxes;
          mbar_expect_tx(uresid_bar, (uint32_t)n_boxes * C::kBoxBytes);
#pragma unroll
          for (int i = 0; i < C::kPasses; ++i)
#pragma unroll
            for (int bx = 0; bx < C::kBoxesPerPass; ++bx) {
              const int c = 64 * i + bx * C::kBoxCols;
              if (c < BLOCK_N)
                tma_load_2d(ustg + i * C::kPassBytes + bx * C::kBoxBytes, &map_resid, n_blk * BLOCK_N + c, m_row0, uresid_bar);
            }
        }
        __syncwarp();
      }
      mbar_wait(tfull_bar(as), aph);
      tcgen05_fence_after();
      // every TMEM chunk of this warp is requested before the first use and the accumulator stage is handed back to the MMA
      // warp as soon as the data sits in registers: the epilogue math / staging / stores overlap the next tile's MMAs
      float v[C::kPasses][kEpiChunk];
#pragma unroll
      for (int i = 0; i < C::kPasses; ++i) {
        const int c0 = 64 * i + kEpiChunk * sub;
        if (c0 < BLOCK_N)
          tmem_ld_32x16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(as * BLOCK_N + c0), v[i]);
      }
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(lead_tempty0 + 8u * as);   // the leader's MMA warp waits for both CTAs' epilogues
        else mbar_arrive(tempty_bar(as));
      }
      if (C::kResid) {
        mbar_wait(resid_bar, rph);
        rph ^= 1;
      }
#pragma unroll
      for (int i = 0; i < C::kPasses; ++i) {
        const int c0 = 64 * i + kEpiChunk * sub;           // first column (inside the tile) of this thread's chunk
        const uint32_t buf = C::kResid ? (uint32_t)i : (C::kOneBar ? pass_ctr % 3u : (pass_ctr & 1u));
        const uint32_t tile_smem = stg + buf * C::kPassBytes;
        if (!C::kResid && !C::kOneBar) {
          if (lead_warp) {
            if (elect_one()) bulk_wait_read<1>();          // the store that last used this staging tile has drained it
            __syncwarp();
          }
          epi_bar(1);
        }
        if (c0 < BLOCK_N) {
          const int n0 = n_blk * BLOCK_N + c0;
          float2* v2 = reinterpret_cast<float2*>(v[i]);
          const float2* sb2 = reinterpret_cast<const float2*>(s_bias + n0);
#pragma unroll
          for (int j = 0; j < kEpiChunk / 2; ++j) v2[j] = __fadd2_rn(v2[j], sb2[j]);
          if (EPI == EPI_BIAS_GELU) {
#pragma unroll
            for (int j = 0; j < kEpiChunk / 2; ++j) v2[j] = gelu_tanh_fit2(v2[j]);
          }
          if (EPI == EPI_BIAS_RELU) {
#pragma unroll
            for (int j = 0; j < kEpiChunk; ++j) v[i][j] = fmaxf(v[i][j], 0.f);
          }
          // 16 columns of row `row` inside the pass: byte offset of the first one, then 16-byte chunks XOR (row & 7)
          const int col_byte = (kEpiChunk * sub) * C::kElt;   // 0, 32, 64, 96 (fp16)  |  0, 64, 128, 192 (fp32)
          const uint32_t box_base = tile_smem + (uint32_t)(col_byte / 128) * C::kBoxBytes + (uint32_t)row * 128u;
          const int chunk0 = (col_byte % 128) / 16;
          if constexpr (C::kElt == 2) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              act16x2 p0 = floats2act2(v[i][8 * q + 0], v[i][8 * q + 1]);
              act16x2 p1 = floats2act2(v[i][8 * q + 2], v[i][8 * q + 3]);
              act16x2 p2 = floats2act2(v[i][8 * q + 4], v[i][8 * q + 5]);
              act16x2 p3 = floats2act2(v[i][8 * q + 6], v[i][8 * q + 7]);
              st_shared_v4(box_base + (uint32_t)(((chunk0 + q) ^ (row & 7)) << 4), *reinterpret_cast<uint32_t*>(&p0),
                           *reinterpret_cast<uint32_t*>(&p1), *reinterpret_cast<uint32_t*>(&p2),
                           *reinterpret_cast<uint32_t*>(&p3));
            }
          } else {
            const float2* ss2 = reinterpret_cast<const float2*>(s_scale + n0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t addr = box_base + (uint32_t)(((chunk0 + q) ^ (row & 7)) << 4);
              if (C::kResid) {
                const float4 x = ld_shared_v4(addr);
                v2[2 * q] = __ffma2_rn(ss2[2 * q], v2[2 * q], make_float2(x.x, x.y));
                v2[2 * q + 1] = __ffma2_rn(ss2[2 * q + 1], v2[2 * q + 1], make_float2(x.z, x.w));
              }
              if (C::kReduce) {
                v2[2 * q] = __fmul2_rn(ss2[2 * q], v2[2 * q]);
                v2[2 * q + 1] = __fmul2_rn(ss2[2 * q + 1], v2[2 * q + 1]);
              }
              st_shared_v4(addr, __float_as_uint(v[i][4 * q]), __float_as_uint(v[i][4 * q + 1]),
                           __float_as_uint(v[i][4 * q + 2]), __float_as_uint(v[i][4 * q + 3]));
            }
          }
        }
        fence_async_smem();   // generic-proxy writes -> visible to the TMA store
        epi_bar(2);
        if (lead_warp) {
          if (elect_one()) {
            const uint32_t utile = ustg + buf * C::kPassBytes;
#pragma unroll
            for (int bx = 0; bx < C::kBoxesPerPass; ++bx) {
              const int c = 64 * i + bx * C::kBoxCols;
              if (c < BLOCK_N) {
                if (C::kReduce) tma_reduce_add_2d(&map_out, utile + bx * C::kBoxBytes, n_blk * BLOCK_N + c, m_row0);
                else tma_store_2d(&map_out, utile + bx * C::kBoxBytes, n_blk * BLOCK_N + c, m_row0);
              }
            }
            bulk_commit();
            if (C::kOneBar) bulk_wait_read<1>();
          }
          __syncwarp();
        }
        ++pass_ctr;
      }
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
    if (lead_warp) {
      if (elect_one()) bulk_wait_all();
      __syncwarp();
    }
  }

  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();   // neither CTA leaves (or frees tensor memory) while the other may still signal it
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int gemm_tc_init() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CNB_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available in this driver");
    return -3;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

// 2-D row-major (rows, cols) tensor of 2- or 4-byte elements, box = (box_rows, 128 bytes of columns), 128B swizzle, zero OOB fill
static int make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int elt_bytes);
int tc_make_map(void* map_out, const void* ptr, int64_t rows, int64_t cols, int box_rows, int elt_bytes) {
  CNB_REQUIRE(g_encode != nullptr, "gemm_tc_init() was not called");
  return make_map(reinterpret_cast<CUtensorMap*>(map_out), ptr, (uint64_t)rows, (uint64_t)cols, (uint32_t)box_rows, elt_bytes);
}
// Encoding a tensor map costs a few microseconds of host time; the encoder launches ~100 GEMMs per pass over a handful of
// (pointer, shape) combinations, so encoded maps are memoised (host launches must stay ahead of ~15 us kernels).
struct MapKey {
  const void* ptr;
  uint64_t rows, cols;
  uint32_t box_rows;
  int elt;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && box_rows == o.box_rows && elt == o.elt;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h = h * 1000003u ^ std::hash<uint64_t>()(k.rows);
    h = h * 1000003u ^ std::hash<uint64_t>()(k.cols);
    h = h * 1000003u ^ std::hash<uint64_t>()(((uint64_t)k.box_rows << 8) | (uint64_t)k.elt);
    return h;
  }
};
static int make_map_uncached(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int elt_bytes);
static int make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int elt_bytes) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{ptr, rows, cols, box_rows, elt_bytes};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *map = it->second;
    return 0;
  }
  if (int rc = make_map_uncached(map, ptr, rows, cols, box_rows, elt_bytes)) return rc;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *map);
  return 0;
}
static int make_map_uncached(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int elt_bytes) {
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * (uint64_t)elt_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / elt_bytes), box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return -3;
  }
  return 0;
}

// 2-D row-major fp16 tensor, box = (box_rows, box_cols) with box_cols * 2 bytes == the swizzle span (64 or 128 bytes)
int tc_make_map_f16_box(void* map_out, const void* ptr, int64_t rows, int64_t cols, int box_rows, int box_cols) {
  CNB_REQUIRE(g_encode != nullptr, "gemm_tc_init() was not called");
  CNB_REQUIRE(box_cols == 32 || box_cols == 64, "fp16 box must span 64 or 128 bytes");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp16 box) failed with CUresult " + std::to_string((int)r));
    return -3;
  }
  return 0;
}

// 4-D NHWC fp32 activation (C fastest), box = (box_c, box_w, 1, 1), no swizzle, zero fill outside the image: one TMA op brings
// one haloed row strip of the depthwise-conv ring (dwconv_ring.cu); negative / overshooting coordinates are the padding
int tc_make_map_nhwc_f32(void* map_out, const float* ptr, int batch, int h, int w, int c, int box_c, int box_w) {
  CNB_REQUIRE(g_encode != nullptr, "gemm_tc_init() was not called");
  CNB_REQUIRE(box_c <= 256 && box_c <= c && box_w <= 256 && box_c % 4 == 0, "NHWC tensor map: box dimensions out of range");
  const cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)c * 4, (cuuint64_t)w * c * 4, (cuuint64_t)h * w * c * 4};
  const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr),
                        dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (NHWC) failed with CUresult " + std::to_string((int)r));
    return -3;
  }
  return 0;
}

template <int BLOCK_N, int EPI, typename OutT, bool PAIR>
static int launch_cfg(const act16* a, const act16* w, int m, int n, int k, const EpiParams& ep, OutT* out,
                      cudaStream_t stream) {
  using C = TcCfg<BLOCK_N, EPI, OutT, PAIR>;
  CUtensorMap map_a, map_w, map_out, map_resid;
  if (int rc = make_map(&map_a, a, (uint64_t)m, (uint64_t)k, kBlockM, 2)) return rc;
  if (int rc = make_map(&map_w, w, (uint64_t)n, (uint64_t)k, C::kBRows, 2)) return rc;
  if (int rc = make_map(&map_out, out, (uint64_t)m, (uint64_t)n, kBlockM, (int)sizeof(OutT))) return rc;
  map_resid = map_out;
  if (C::kResid)
    if (int rc = make_map(&map_resid, ep.resid, (uint64_t)m, (uint64_t)n, kBlockM, 4)) return rc;
  auto kern = gemm_tc_kernel<BLOCK_N, EPI, OutT, PAIR>;
  static bool attr_set = false;
  if (!attr_set) {
    CNB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_set = true;
  }
  const int n_tiles = (int)ceil_div(m, C::kTileM) * (n / BLOCK_N);
  if (PAIR) {
    const int pairs = n_tiles < sm_budget() / 2 ? n_tiles : sm_budget() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = C::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CNB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, map_a, map_w, map_out, map_resid, m, n, k, ep));
    CNB_LAUNCH_OK();
    return 0;
  }
  const int grid = n_tiles < sm_budget() ? n_tiles : sm_budget();
  kern<<<grid, kTcThreads, C::kTotal, stream>>>(map_a, map_w, map_out, map_resid, m, n, k, ep);
  CNB_LAUNCH_OK();
  return 0;
}

// CTA pairs pay off when the tile count fills the 74 pairs several times over; CNB_GEMM_PAIR=0 turns them off, =<rows> sets
// the minimum M (default 512 so that the ragged-M unit tests exercise the pair protocol too)
static int pair_min_rows() {
  static const int v = [] {
    const char* e = getenv("CNB_GEMM_PAIR");
    if (!e) return 512;
    const int x = atoi(e);
    return x <= 0 ? INT_MAX : x;
  }();
  return v;
}

template <int EPI, typename OutT>
static int launch_n(const act16* a, const act16* w, int m, int n, int k, const EpiParams& ep, OutT* out,
                    cudaStream_t stream) {
  // measured (64 x 10 s clips, pair vs 1-CTA): pw2 stages 3-4 -18 % / -22 %, pw1 stage 3 -9 %, stage 4 -12 %, downsample -9 %,
  // stage 2 equal (HBM-bound).  (With a .release.cluster accumulator hand-back the short-K shapes were 26-39 % SLOWER: the
  // cluster-scope fence serialised epilogue and MMAs at every tile boundary -- see mbar_arrive_cluster.)
  static const int pair_min_k = [] { const char* e = getenv("CNB_GEMM_PAIR_MINK"); return e ? atoi(e) : 0; }();
  if (m >= pair_min_rows() && k >= pair_min_k) {
    // pw1 (GELU, fp16 out): 256-column tiles when N allows (768 / 1536 / 3072): the A tile is re-fetched 6x instead of 8x per
    // 256 rows and the epilogue's per-tile hand-offs amortise over a third more columns (pw1 stages 2-4: -2 % / -3 % / -7 %)
    static const bool wide = [] { const char* e = getenv("CNB_GEMM_N256"); return !e || atoi(e) != 0; }();
    if constexpr (EPI == EPI_BIAS_GELU && sizeof(OutT) == 2)
      if (wide && n % 256 == 0) return launch_cfg<256, EPI, OutT, true>(a, w, m, n, k, ep, out, stream);
    if (n % 192 == 0) return launch_cfg<192, EPI, OutT, true>(a, w, m, n, k, ep, out, stream);
    if (n % 128 == 0) return launch_cfg<128, EPI, OutT, true>(a, w, m, n, k, ep, out, stream);
  }
  if (n % 192 == 0) return launch_cfg<192, EPI, OutT, false>(a, w, m, n, k, ep, out, stream);
  if (n % 128 == 0) return launch_cfg<128, EPI, OutT, false>(a, w, m, n, k, ep, out, stream);
  if (n == 96) return launch_cfg<96, EPI, OutT, false>(a, w, m, n, k, ep, out, stream);
  set_error("gemm_tc: N=" + std::to_string(n) + " is neither 96 nor a multiple of 128/192");
  return -1;
}

template <typename OutT>
int launch_gemm_tc(const act16* a, const act16* w, int m, int n, int k, Epilogue epi, const EpiParams& ep,
                   OutT* out, int64_t ldo, cudaStream_t stream) {
  CNB_REQUIRE(g_encode != nullptr, "gemm_tc_init() was not called");
  CNB_REQUIRE(k % 8 == 0, "gemm_tc needs K to be a multiple of 8 (16-byte TMA row stride)");
  CNB_REQUIRE(ep.bias != nullptr, "gemm_tc epilogues need a bias vector");
  CNB_REQUIRE(ldo == n, "gemm_tc writes through a TMA tensor map: the output must be contiguous (ldo == N)");
  CNB_REQUIRE(n <= (epi == EPI_SCALE_RESID ? 768 : kMaxN), "gemm_tc stages bias/scale in shared memory: N too large");
  if (m == 0) return 0;
  switch (epi) {
    case EPI_BIAS: return launch_n<EPI_BIAS, OutT>(a, w, m, n, k, ep, out, stream);
    case EPI_BIAS_GELU: return launch_n<EPI_BIAS_GELU, OutT>(a, w, m, n, k, ep, out, stream);
    case EPI_BIAS_RELU: return launch_n<EPI_BIAS_RELU, OutT>(a, w, m, n, k, ep, out, stream);
    case EPI_SCALE_RESID:
      if constexpr (sizeof(OutT) == 4) {
        CNB_REQUIRE(ep.scale != nullptr && ep.resid != nullptr, "gemm_tc: the layer-scale/residual epilogue needs scale and resid");
        // the reduce-add epilogue accumulates into `out`: the encoder updates x in place (resid == out); otherwise seed it
        if (!CNB_PW2_RESID_LOAD && (const void*)ep.resid != (const void*)out)
          CNB_CUDA_OK(cudaMemcpyAsync(out, ep.resid, (size_t)m * n * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        return launch_n<EPI_SCALE_RESID, OutT>(a, w, m, n, k, ep, out, stream);
      }
      break;
  }
  set_error("gemm_tc: unsupported epilogue / output type combination");
  return -1;
}
template int launch_gemm_tc<float>(const act16*, const act16*, int, int, int, Epilogue, const EpiParams&,
                                   float*, int64_t, cudaStream_t);
template int launch_gemm_tc<act16>(const act16*, const act16*, int, int, int, Epilogue,
                                           const EpiParams&, act16*, int64_t, cudaStream_t);

}  // namespace cnb
