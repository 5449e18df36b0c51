// rpx_gemm2.cuh — the 2-CTA (cta_group::2) form of the contraction core in rpx_gemm.cuh.
//
//   D[M, N] = A[M, K] * B[N, K]^T   on 256 x 256 output tiles computed by a PAIR of CTAs (a
//   2-CTA cluster on one TPC).  CTA r of the pair holds rows [128r, 128r+128) of the A tile and
//   rows [n/2 * r, n/2 * (r+1)) of the B tile in ITS shared memory; one tcgen05.mma.cta_group::2
//   issued by the leader CTA (rank 0) multiplies both halves and writes each CTA's 128 accumulator
//   rows into that CTA's TMEM.
//
// Why: with one CTA per tile every k-block needs 48 KB of operands (16 KB A + 32 KB B) for 512
// cycles of MMA, so the 4-stage ring hides only ~1.5k cycles of load latency and the tensor pipe
// starves on the operand streamed from HBM (measured 84 % tensor-active on the FFN-up GEMM).  A
// pair needs 32 KB per CTA for the same 512 cycles: 6 stages, 50 % more latency cover, a third
// less L2->SM traffic.
//
// Protocol (per stage / per tile), following the PTX ISA's cta_group::2 rules:
//   - TMA loads use the .cta_group::2 form and signal the LEADER's `full` barrier (peer bit of
//     the barrier address cleared); the leader arms it with the pair's total byte count.
//   - the leader's MMA thread commits with .multicast::cluster to the `empty` / `tfull` barriers
//     of BOTH CTAs, so each CTA's producer and epilogue warps run exactly as in the 1-CTA kernel.
//   - epilogue warps of both CTAs arrive on the LEADER's `tempty` barrier (remote arrive via mapa).
//   - TMEM is allocated / freed with the cta_group::2 forms by one warp of each CTA; the cluster
//     synchronises after barrier init and before teardown.
// The epilogue functors (rpx_gemm.cuh) are reused unchanged.
#pragma once
#include "rpx_gemm.cuh"

namespace rpx {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address

// 2-SM TMA load: data lands in THIS CTA's shared memory, completion bytes go to the leader's barrier.
RPX_DEVICE void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask),
        "r"(c0), "r"(c1)
      : "memory");
}
RPX_DEVICE void tma_load_2d_2sm_hint(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1,
                                     uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask),
        "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
RPX_DEVICE void tmem_alloc_2sm(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
RPX_DEVICE void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
RPX_DEVICE void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all prior MMAs of this thread retired) on the barrier at this smem offset in every CTA
// of `cta_mask`.
RPX_DEVICE void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
RPX_DEVICE void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}

constexpr int kPairM = 2 * kBlockM;  // 256 output rows per CTA pair

template <int STAGES>
struct Gemm2Cfg {
  static constexpr int kBlockN = 256;
  static constexpr int kABytes = kBlockM * kBlockK * 2;        // this CTA's 128 A rows
  static constexpr int kBBytes = (kBlockN / 2) * kBlockK * 2;  // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;        // 32 KB
  static constexpr int kTmemCols = 2 * kBlockN;
  static constexpr int kBarBytes = 256;
  static constexpr size_t smem_bytes(size_t epi_extra) {
    return (size_t)STAGES * kStageBytes + 1024 + kBarBytes + epi_extra;
  }
};

// Column range of n-tile `n_blk`: N is cut into tiles_n near-equal tiles in units of 32 columns
// (the wider ones last) instead of tiles_n - 1 full tiles and a short tail.  With N = 1472 the
// tail tile would be 192 wide: the pairs that visit it (tile = pair + i * n_pairs walks a fixed
// residue class of n_blk) ran ~8 % ahead of the others, the six pairs sharing an A row block fell
// out of step and A came from DRAM ~3x instead of once (ncu: 7.1 GB read for the FFN
// down-projection against 3.4 GB algorithmic).  224/224/256/256/256/256 gives every pair the same
// work per three tiles.
__device__ __forceinline__ void n_tile_range(int n_blk, int N, int tiles_n, int& n0, int& n_this) {
  const int units = (N + 31) >> 5;
  const int base = units / tiles_n;
  const int first_wide = tiles_n - (units - base * tiles_n);
  n0 = 32 * (n_blk * base + (n_blk > first_wide ? n_blk - first_wide : 0));
  n_this = 32 * (base + (n_blk >= first_wide ? 1 : 0));
}

// Tile -> (row block of 256, n-tile origin / width).  Default: n fastest, near-equal n-tiles.
// M_FASTEST (the similarity kernel): the pair grid is a multiple of tiles_m, so a pair keeps ONE row
// block (its two query blocks) for all its tiles and walks the corpus in fixed 256-wide tiles; the
// last tile may be ragged (width rounded up to 32: TMA zero-fills, the epilogue masks by N).
template <bool M_FASTEST>
__device__ __forceinline__ void decode_tile(int tile, int N, int tiles_m, int tiles_n, int& m_blk, int& n_blk,
                                            int& n0, int& n_this) {
  if (M_FASTEST) {
    m_blk = tile % tiles_m;
    n_blk = tile / tiles_m;
    n0 = n_blk * 256;
    const int rest = (N - n0 + 31) & ~31;
    n_this = rest < 256 ? rest : 256;
  } else {
    n_blk = tile % tiles_n;
    m_blk = tile / tiles_n;
    n_tile_range(n_blk, N, tiles_n, n0, n_this);
  }
}

template <int STAGES, class Epi, bool M_FASTEST = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gemm_threads<Epi>(), 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                int tiles_m, int tiles_n, const __grid_constant__ typename Epi::Params ep) {
  using Cfg = Gemm2Cfg<STAGES>;
  constexpr int BLOCK_N = Cfg::kBlockN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024 - (raw_addr & 1023)) & 1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint8_t* smem_extra = smem + STAGES * Cfg::kStageBytes + Cfg::kBarBytes;

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int num_tiles = tiles_m * tiles_n;  // tiles_m counts 256-row blocks
  const int num_kb = K / kBlockK;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);   // leader's producer arms it; both CTAs' TMA complete_tx on the leader's copy
        mbar_init(&empty[s], 1);  // one multicast commit per round
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull[s], 1);
        mbar_init(&tempty[s], 2 * Epi::kWarps);  // one arrive per epilogue warp of either CTA (leader's copy is used)
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // (see gemm_tc_kernel)
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs) {
        int m_blk, n_blk, n0, n_this;
        decode_tile<M_FASTEST>(tile, N, tiles_m, tiles_n, m_blk, n_blk, n0, n_this);
        const int a_row = m_blk * kPairM + (int)rank * kBlockM;
        const int b_row = n0 + (int)rank * (n_this / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1, 1);
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
          if (M_FASTEST) {
            // the row block is re-read by every tile of this pair, the other operand is streamed once
            // (same L2 policy as the 1-CTA similarity kernel: -20 % there)
            tma_load_2d_2sm_hint(sA + stage * Cfg::kABytes, &tmA, &full[stage], kb * kBlockK, a_row, kEvictLast);
            tma_load_2d_2sm_hint(sB + stage * Cfg::kBBytes, &tmB, &full[stage], kb * kBlockK, b_row, kEvictFirst);
          } else {
            tma_load_2d_2sm(sA + stage * Cfg::kABytes, &tmA, &full[stage], kb * kBlockK, a_row);
            tma_load_2d_2sm(sB + stage * Cfg::kBBytes, &tmB, &full[stage], kb * kBlockK, b_row);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs) {
        int m_blk, n_blk, n0, n_this;
        decode_tile<M_FASTEST>(tile, N, tiles_m, tiles_n, m_blk, n_blk, n0, n_this);
        const uint32_t idesc = make_idesc_bf16(kPairM, (uint32_t)n_this);
        mbar_wait(&tempty[as], aphase ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase, 3);
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc_kmajor_sw128(smem_u32(sA + stage * Cfg::kABytes));
          const uint64_t b_desc = make_smem_desc_kmajor_sw128(smem_u32(sB + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_bf16_ss_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2sm(&empty[stage], 3);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tfull[as], 3);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs)
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + (threadIdx.x & 31);
    const int part = (warp - kEpiWarp0) >> 2;
    Epi epi(ep, smem_extra, row, part);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = pair; tile < num_tiles; tile += n_pairs) {
      TileCtx t;
      int n_this;
      decode_tile<M_FASTEST>(tile, N, tiles_m, tiles_n, t.m_blk, t.n_blk, t.n0, n_this);
      t.m0 = t.m_blk * kPairM + (int)rank * kBlockM;
      t.n_cols = n_this < N - t.n0 ? n_this : ((N - t.n0 + 31) & ~31);
      t.row = row;
      t.rows_per_warp = 32;
      t.part = part;
      t.split = Epi::kWarps / 4;
      t.M = M;
      t.N = N;
      t.tmem = tmem_base + as * BLOCK_N + ((uint32_t)(lane_grp * 32) << 16);
      {
        const int nt = tile + n_pairs;
        if (nt < num_tiles) {
          int nm, nn;
          decode_tile<M_FASTEST>(nt, N, tiles_m, tiles_n, nm, nn, t.next_n0, t.next_cols);
          t.next_m0 = nm * kPairM + (int)rank * kBlockM;
          if (t.next_cols > N - t.next_n0) t.next_cols = N - t.next_n0;
        } else {
          t.next_m0 = -1;
          t.next_n0 = 0;
          t.next_cols = 0;
        }
      }
      epi.before_wait(t);
      mbar_wait(&tfull[as], aphase, 4);
      tc_fence_after();
      epi.tile(t);
      tc_fence_before();
      __syncwarp();
      if ((threadIdx.x & 31) == 0) {
        if (leader) mbar_arrive(&tempty[as]);
        else mbar_arrive_remote(&tempty[as], 0);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    epi.finish();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA tears down while its peer may still touch its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace rpx
