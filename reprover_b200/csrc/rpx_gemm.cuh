// rpx_gemm.cuh — the tcgen05 / TMA / TMEM contraction core shared by the encoder
// GEMMs and the similarity kernel.
//
//   D[M, N] (fp32, in TMEM) = A[M, K] * B[N, K]^T        A, B bf16, K contiguous
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      : TMA producer   — cp.async.bulk.tensor A/B tiles into a STAGES-deep
//                 128B-swizzled shared-memory ring, completion on `full[]` mbarriers
//   warp 1      : MMA issuer     — one elected thread issues tcgen05.mma (M=128,
//                 N<=256, K=16) x4 per stage; tcgen05.commit releases the stage
//                 (`empty[]`) and, after the last k-block, publishes the accumulator
//                 (`tfull[]`).  Also owns TMEM alloc/dealloc.
//   warps 2..5  : epilogue       — tcgen05.ld the accumulator (one row per thread)
//                 and run the fused epilogue functor; `tempty[]` hands the TMEM
//                 stage back.  Two accumulator stages (2 x BLOCK_N columns) let the
//                 epilogue of tile i overlap the mainloop of tile i+1.
//
// `n_blk_stride` > 1 makes the kernel visit only every stride-th B tile (the similarity
// kernel's sampling pass); it is 1 everywhere else.
//
// Tiles are visited in n-fastest order so that concurrently resident CTAs share
// the same A row-block through L2 (the B operand — the weights — is small and
// L2-resident).  The similarity kernel uses M_FASTEST instead: A is the (small)
// query block, B the streamed corpus, and a grid that is a multiple of tiles_m
// pins every CTA to one query block for its whole life.
//
// The reference has no counterpart: it calls torch `@` / nn.Linear (cuBLAS) —
// SURVEY.md §2.1 K3/K8/K9/K11.
#pragma once
#include "rpx_ptx.cuh"

namespace rpx {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int kUmmaK = 16;
#ifndef RPX_EPI_WARPS
#define RPX_EPI_WARPS 4
#endif
constexpr int kEpiWarp0 = 2;  // first epilogue warp
// threads of a launch: TMA warp + MMA warp + Epi::kWarps epilogue warps (4 or 8)
template <class Epi>
constexpr int gemm_threads() { return 64 + 32 * Epi::kWarps; }

template <int BLOCK_N, int STAGES, int BM = kBlockM>
struct GemmCfg {
  static_assert(BM == 64 || BM == 128, "UMMA M of a 1-CTA tile: 64 or 128");
  static_assert(2 * STAGES + 4 <= 30, "barrier block holds at most 13 stages");
  static constexpr int kABytes = BM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // two accumulator stages
  static_assert(kTmemCols == 64 || kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512,
                "TMEM allocation must be a power of two in [32, 512]");
  // ring + 1 KB alignment slack + barriers/tmem pointer
  static constexpr int kBarBytes = 256;
  static constexpr size_t smem_bytes(size_t epi_extra) {
    return (size_t)STAGES * kStageBytes + 1024 + kBarBytes + epi_extra;
  }
};

// Kernel argument of gemm_tc_kernel: CTAs [0, work_ctas) run the GEMM (persistent over the tiles with that
// stride); CTAs beyond prefetch `bytes` at `ptr` into L2 and leave.  work_ctas == gridDim.x: no helpers.
struct L2Prefetch {
  const void* ptr;
  uint32_t bytes;
  int work_ctas;
  unsigned long long* timeline;  // debug (rpx_debug_set_timeline): 8 %globaltimer stamps of CTA 0, or null
};

RPX_DEVICE unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define RPX_STAMP(pf, i)                                                   \
  do {                                                                     \
    if ((pf).timeline != nullptr && blockIdx.x == 0) (pf).timeline[i] = global_ns(); \
  } while (0)

// What an epilogue functor sees for one output tile.
struct TileCtx {
  uint32_t tmem;   // TMEM address of this thread's row, column 0 of the accumulator stage
  int m0, n0;      // tile origin in the output
  int n_cols;      // valid columns in this tile (multiple of 32)
  int row;         // this thread's row inside the tile, 0..127
  int m_blk, n_blk;
  int M, N;
  int next_m0, next_n0;  // origin of the next tile this CTA will process (next_m0 < 0: none)
  int next_cols;         // its width
  int part, split;       // this warp handles the 32-column chunks with (chunk % split) == part
  int rows_per_warp;     // tile rows held by one TMEM lane group: 32 (M = 128 tiles) or 16 (M = 64: lanes 16-31 idle)
};

// Epi must provide:
//   struct Params;                       (trivially copyable kernel argument)
//   static constexpr size_t kSmemBytes;  (extra dynamic shared memory, may be 0)
//   static constexpr int kWarps;         (4 or 8 epilogue warps; with 8, two warps share a TMEM lane
//                                         group and split the tile's 32-column chunks between them)
//   __device__ Epi(const Params&, uint8_t* smem_extra, int row /*tile row this thread owns, 0..127*/,
//                  int part /*column share of this warp, 0..kWarps/4-1*/);
//   __device__ void before_wait(const TileCtx&);  (work that may run while the MMAs of this tile are
//                                                  still in flight, e.g. prefetching)
//   __device__ void tile(const TileCtx&);     (all 128 epilogue threads, warp-converged)
//   __device__ void finish();
// SPLIT_B (the gated FFN up-projection on narrow tiles): the B tile is two boxes of BLOCK_N/2 rows — the
// gate rows and the linear-branch rows of the same BLOCK_N/2 hidden units.  The packed weight interleaves
// wi_0 / wi_1 in 128-row blocks (rows [256j, 256j+128) gate, [256j+128, 256j+256) linear, see
// rpx_encoder.cu), so n-tile t (units u0 = t * BLOCK_N/2) takes rows 256 (u0/128) + u0 % 128 and the same + 128;
// tmB must then be encoded with a box of BLOCK_N/2 rows.
//
// BM = 64 (latency path): 64-row tiles, tcgen05.mma M = 64.  The accumulator then occupies lanes 0-15 of each
// 32-lane TMEM group (row r of the tile sits in lane 32 (r / 16) + r % 16), so an epilogue warp owns 16 rows and
// its upper 16 lanes idle.  A CTA takes in half the activation bytes per k-block and the ring holds more
// stages — what a narrow GEMM's time is made of (see rpx_encoder.cu).
template <int BLOCK_N, int STAGES, class Epi, bool M_FASTEST = false, bool SPLIT_B = false, int BM = kBlockM>
__global__ void __launch_bounds__(gemm_threads<Epi>(), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               int M, int N, int K, int tiles_m, int tiles_n, int n_blk_stride, typename Epi::Params ep,
               L2Prefetch pf) {
  using Cfg = GemmCfg<BLOCK_N, STAGES, BM>;
  if (threadIdx.x == 0) RPX_STAMP(pf, 0);
  if ((int)blockIdx.x >= pf.work_ctas) {
    // Helper CTA (latency path): the GEMM itself keeps only a fraction of the SMs busy, so the launch is
    // widened to the whole GPU and the surplus CTAs pull the NEXT layer's weights into L2 (36 MB of 126 MB)
    // while this layer computes — its GEMMs then stream their B operand at L2 instead of DRAM latency.
    pdl_launch_dependents();
    const int n_help = (int)gridDim.x - pf.work_ctas;
    const size_t lines = ((size_t)pf.bytes + 127) >> 7;
    const char* base = static_cast<const char*>(pf.ptr);
    for (size_t i = (size_t)((int)blockIdx.x - pf.work_ctas) * blockDim.x + threadIdx.x; i < lines;
         i += (size_t)n_help * blockDim.x)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (i << 7)));
    return;
  }
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tile bases.
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
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = K / kBlockK;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull[s], 1);
        mbar_init(&tempty[s], 32 * Epi::kWarps);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above touched only this CTA's shared memory / TMEM.  Under programmatic dependent launch
  // (rpx_ptx.cuh) the preceding kernel may still be running: the A operand and whatever the epilogue reads
  // from global memory are its outputs and must wait for it (pdl_wait), but the B operand — the weights —
  // is never written by a kernel of the chain, so the producer streams the first ring's worth of B tiles
  // BEFORE it waits: by the time the predecessor retires, a third of this CTA's weights are already on chip.
  pdl_launch_dependents();
  if (threadIdx.x == 0) RPX_STAMP(pf, 1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int pre = 0;  // stages of the first tile whose barrier is armed and whose B tile is already in flight
      auto load_b = [&](int st, int n_blk, int kb) {
        if (M_FASTEST) {
          tma_load_2d_hint(sB + st * Cfg::kBBytes, &tmB, &full[st], kb * kBlockK, n_blk * n_blk_stride * BLOCK_N, kEvictFirst);
        } else if (SPLIT_B) {
          constexpr int H = BLOCK_N / 2;
          const int u0 = n_blk * H;
          const int gate_row = (u0 / 128) * 256 + (u0 % 128);
          tma_load_2d(sB + st * Cfg::kBBytes, &tmB, &full[st], kb * kBlockK, gate_row);
          tma_load_2d(sB + st * Cfg::kBBytes + H * kBlockK * 2, &tmB, &full[st], kb * kBlockK, gate_row + 128);
        } else {
          tma_load_2d(sB + st * Cfg::kBBytes, &tmB, &full[st], kb * kBlockK, n_blk * n_blk_stride * BLOCK_N);
        }
      };
      if (!M_FASTEST && (int)blockIdx.x < num_tiles) {
        const int n_blk0 = (int)blockIdx.x % tiles_n;
        pre = num_kb < STAGES ? num_kb : STAGES;
        for (int kb = 0; kb < pre; ++kb) {  // fresh barriers: every stage is free
          mbar_arrive_expect_tx(&full[kb], Cfg::kStageBytes);
          load_b(kb, n_blk0, kb);
        }
      }
      pdl_wait();
      RPX_STAMP(pf, 2);
      for (int tile = blockIdx.x; tile < num_tiles; tile += pf.work_ctas) {
        const int n_blk = M_FASTEST ? tile / tiles_m : tile % tiles_n;
        const int m_blk = M_FASTEST ? tile % tiles_m : tile / tiles_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (pre > 0) {
            --pre;  // armed, B in flight: only the A tile is missing
          } else {
            mbar_wait(&empty[stage], phase ^ 1, 1);
            mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
            load_b(stage, n_blk, kb);
          }
          // L2 eviction hints: streamed operand evict-first, re-used operand evict-last.  Measured
          // on B200 (A/B, one box): -20 % time for the similarity kernel (corpus streamed once, query
          // block re-read by every tile), +3 % for the encoder GEMMs -> M_FASTEST only.  (An L2
          // prefetch of the streamed operand a few k-blocks ahead of its TMA load was also measured:
          // 6-12 % slower at every distance tried, removed.)
          if (M_FASTEST) {
            tma_load_2d_hint(sA + stage * Cfg::kABytes, &tmA, &full[stage], kb * kBlockK, m_blk * BM, kEvictLast);
          } else {
            tma_load_2d(sA + stage * Cfg::kABytes, &tmA, &full[stage], kb * kBlockK, m_blk * BM);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += pf.work_ctas) {
        const int n_blk = M_FASTEST ? tile / tiles_m : tile % tiles_n;
        int n_this = N - n_blk * n_blk_stride * BLOCK_N;
        if (n_this > BLOCK_N) n_this = BLOCK_N;
        n_this = (n_this + 15) & ~15;
        const uint32_t idesc = make_idesc_bf16(BM, (uint32_t)n_this);
        mbar_wait(&tempty[as], aphase ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase, 3);
          if (kb == 0 && tile == (int)blockIdx.x) RPX_STAMP(pf, 3);
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc_kmajor_sw128(smem_u32(sA + stage * Cfg::kABytes));
          const uint64_t b_desc = make_smem_desc_kmajor_sw128(smem_u32(sB + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // +32 bytes (>>4 = 2) per K=16 step inside the swizzle atom
            umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);
          if (kb == num_kb - 1 && tile == (int)blockIdx.x) RPX_STAMP(pf, 4);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int lane_grp = warp & 3;                            // TMEM lane group this warp may read
    const int lane = threadIdx.x & 31;
    // row of the tile this thread owns; with 64-row tiles lanes 16-31 of a group hold nothing: their row lies
    // beyond every matrix, so the functors' `m < M` tests switch them off
    const int row = BM == kBlockM ? lane_grp * 32 + lane : (lane < 16 ? lane_grp * 16 + lane : (1 << 28));
    const int part = (warp - kEpiWarp0) >> 2;                 // which share of the columns (0 when 4 warps)
    pdl_wait();  // the epilogue reads (row scales, residual stream) and overwrites the predecessor's outputs
    Epi epi(ep, smem_extra, lane_grp * 32 + lane, part);      // (the functor's staging is per TMEM lane)
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += pf.work_ctas) {
      TileCtx t;
      t.n_blk = M_FASTEST ? tile / tiles_m : tile % tiles_n;
      t.m_blk = M_FASTEST ? tile % tiles_m : tile / tiles_n;
      t.m0 = t.m_blk * BM;
      t.rows_per_warp = BM / 4;
      t.n0 = t.n_blk * n_blk_stride * BLOCK_N;  // (n_blk stays the logical tile index)
      int n_this = N - t.n0;
      if (n_this > BLOCK_N) n_this = BLOCK_N;
      t.n_cols = n_this;
      t.row = row;
      t.part = part;
      t.split = Epi::kWarps / 4;
      t.M = M;
      t.N = N;
      t.tmem = tmem_base + as * BLOCK_N + ((uint32_t)(lane_grp * 32) << 16);
      {
        const int nt = tile + pf.work_ctas;
        if (nt < num_tiles) {
          t.next_m0 = (M_FASTEST ? nt % tiles_m : nt / tiles_n) * BM;
          t.next_n0 = (M_FASTEST ? nt / tiles_m : nt % tiles_n) * BLOCK_N;
          t.next_cols = N - t.next_n0 < BLOCK_N ? N - t.next_n0 : BLOCK_N;
        } else {
          t.next_m0 = -1;
          t.next_n0 = 0;
          t.next_cols = 0;
        }
      }
      epi.before_wait(t);
      mbar_wait(&tfull[as], aphase, 4);
      if (threadIdx.x == 64 && tile == (int)blockIdx.x) RPX_STAMP(pf, 5);
      tc_fence_after();
      epi.tile(t);
      if (threadIdx.x == 64 && tile == (int)blockIdx.x) RPX_STAMP(pf, 6);
      tc_fence_before();
      mbar_arrive(&tempty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    epi.finish();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
  if (threadIdx.x == 0) RPX_STAMP(pf, 7);
}

// ============================================================================ epilogues

// Row scale shared by the epilogues that consume an RMSNorm'd operand: the
// RMSNorm weight is folded into the GEMM's B operand at load time, so all that
// is left is rs[m] = rsqrt(mean_k(x[m,k]^2) + eps), applied to the fp32 accumulator.
// sumsq is delivered as `n_parts` partial sums per row (written by the producing
// epilogue, one per output n-block), summed here in a fixed order -> deterministic.
struct RowScale {
  const float* ss_parts;  // [n_parts][M] or nullptr (scale 1)
  int n_parts;
  int part_stride;        // elements between parts (>= M)
  float inv_dim;          // 1 / d_model
  float eps;
  __device__ float get(int m) const {
    if (ss_parts == nullptr) return 1.0f;
    float s = 0.f;
    // loads in batches of 24 (independent, all in flight together), additions in part order: the latency
    // path has 46 parts per row and would pay one L2 round trip per part with a plain loop
    constexpr int kBatch = 24;
    for (int p = 0; p < n_parts; p += kBatch) {
      float v[kBatch];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) v[j] = p + j < n_parts ? ss_parts[(size_t)(p + j) * part_stride + m] : 0.f;
#pragma unroll
      for (int j = 0; j < kBatch; ++j)
        if (p + j < n_parts) s += v[j];
    }
    return rsqrtf(s * inv_dim + eps);
  }
};

// C[m, n] = acc (fp32).  Generic; used by tests.
struct EpiStoreF32 {
  struct Params {
    float* C;
    int ldc;
  };
  static constexpr size_t kSmemBytes = 0;
  static constexpr int kWarps = RPX_EPI_WARPS;
  Params p;
  __device__ EpiStoreF32(const Params& p_, uint8_t*, int, int) : p(p_) {}
  __device__ void before_wait(const TileCtx&) {}
  __device__ void tile(const TileCtx& t) {
    const int m = t.m0 + t.row;
    const bool ok = m < t.M;
    for (int c = 32 * t.part; c < t.n_cols; c += 32 * t.split) {
      uint32_t v[32];
      tmem_ld_32x32(t.tmem + c, v);
      tmem_ld_wait();
      if (ok) {
        float4* dst = reinterpret_cast<float4*>(p.C + (size_t)m * p.ldc + t.n0 + c);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                               __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
    }
  }
  __device__ void finish() {}
};

// C[m, n] = bf16(acc * rs[m]).  QKV projection (RMSNorm folded: prologue of K3).
struct EpiStoreBF16 {
  struct Params {
    __nv_bfloat16* C;
    int ldc;
    RowScale rs;
  };
  static constexpr size_t kSmemBytes = 0;
  static constexpr int kWarps = RPX_EPI_WARPS;
  Params p;
  float rs = 0.f;  // this thread's row scale for the tile in flight
  __device__ EpiStoreBF16(const Params& p_, uint8_t*, int, int) : p(p_) {}
  // the row scale does not depend on the accumulator: its partial sums (up to 23 L2 round trips on the latency
  // path) are fetched while the MMAs of the tile are still running
  __device__ void before_wait(const TileCtx& t) {
    const int m = t.m0 + t.row;
    rs = m < t.M ? p.rs.get(m) : 0.f;
  }
  __device__ void tile(const TileCtx& t) {
    const int m = t.m0 + t.row;
    const bool ok = m < t.M;
    for (int c = 32 * t.part; c < t.n_cols; c += 32 * t.split) {
      uint32_t v[32];
      tmem_ld_32x32(t.tmem + c, v);
      tmem_ld_wait();
      if (ok) {
        uint4* dst = reinterpret_cast<uint4*>(p.C + (size_t)m * p.ldc + t.n0 + c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]) * rs, __uint_as_float(v[8 * i + 1]) * rs);
          o.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * rs, __uint_as_float(v[8 * i + 3]) * rs);
          o.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * rs, __uint_as_float(v[8 * i + 5]) * rs);
          o.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * rs, __uint_as_float(v[8 * i + 7]) * rs);
          dst[i] = o;
        }
      }
    }
  }
  __device__ void finish() {}
};

// Residual update (attention output projection K8 and FFN down projection K9):
//   h32[m, n] += acc;  h16[m, n] = bf16(h32[m, n]);  ss_out[part-of-n_blk][m] = sum_n h32[m, n]^2
// The fp32 copy is the residual stream; the bf16 copy is the next GEMM's A operand;
// ss_out feeds the next RMSNorm (see RowScale).
//
// The accumulator arrives one ROW per thread (TMEM lane == thread), which is the worst
// possible layout for global memory: a warp-wide 16-byte access would touch 32 different
// 128-byte lines.  Each warp therefore transposes its 32x32 fp32 block through a swizzled
// shared-memory tile and does the read-modify-write with lanes running along the row:
// one instruction covers 4 rows x 128 contiguous bytes (4 L1 wavefronts instead of 32).
// Residual loads run one chunk ahead.  (RPX_EPI_WARPS=8 — two warps per TMEM lane group splitting
// the chunks — was measured neutral to slightly negative on B200 and is off by default.  Measured
// and dropped: an L2 prefetch of the next tile's residual rows one tile ahead — the lines
// were evicted again before use, 6.0 GB read per launch against 3.4 GB algorithmic — and L2
// eviction-priority hints on the residual loads / stores, 0 %.)
// CHUNK_SS: one partial sum per 32-column chunk instead of one per tile (ss_out is then indexed by the chunk's
// position in the row, [N / 32][ss_stride]) — the latency path uses 32- or 64-wide tiles depending on the token
// count and must hand the next RMSNorm the same partial sums either way.
struct EpiResidualParams {
  float* h32;
  __nv_bfloat16* h16;
  int ld;
  float* ss_out;  // [tiles_n * kWarps / 4][ss_stride]
  int ss_stride;
};
template <bool CHUNK_SS>
struct EpiResidualT {
  using Params = EpiResidualParams;
  static constexpr int kWarps = RPX_EPI_WARPS;
  static constexpr size_t kSmemBytes = kWarps * 32 * 32 * sizeof(float);  // one 32x32 tile per warp
  Params p;
  float4* stg;  // this warp's staging tile: row r = 8 float4, stored at slot (j ^ (r & 7))
  int lane, grp;
  float4 h[8];  // residual values of the chunk in flight (loaded one chunk ahead)
  __device__ EpiResidualT(const Params& p_, uint8_t* smem_extra, int row, int part) : p(p_) {
    lane = row & 31;
    grp = row >> 5;
    stg = reinterpret_cast<float4*>(smem_extra) + (part * 4 + grp) * 32 * 8;
  }
  __device__ __forceinline__ void load_chunk(const TileCtx& t, int c, int row_base, int sub, int col4,
                                             float4 (&hh)[8]) const {
    const size_t col = (size_t)t.n0 + c + col4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = sub + 4 * i, m = row_base + r;
      if (r < t.rows_per_warp && m < t.M) hh[i] = *reinterpret_cast<const float4*>(p.h32 + (size_t)m * p.ld + col);
      else hh[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // the first chunk's residual values do not depend on the accumulator: they are fetched while the MMAs of
  // the tile are still running (on the latency path that L2 round trip was a third of the epilogue)
  __device__ void before_wait(const TileCtx& t) {
    const int c = 32 * t.part;
    if (c < t.n_cols) load_chunk(t, c, t.m0 + grp * t.rows_per_warp, lane >> 3, (lane & 7) * 4, h);
  }
  __device__ void tile(const TileCtx& t) {
    const int sub = lane >> 3;        // row within a group of 4
    const int j4 = lane & 7;          // which float4 of the 32-column chunk
    const int col4 = j4 * 4;
    const int row_base = t.m0 + grp * t.rows_per_warp;
    const int step = 32 * t.split;
    float ss[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ss[i] = 0.f;
    float4 hn[8];
    int c = 32 * t.part;
    for (; c < t.n_cols; c += step) {
      uint32_t v[32];
      tmem_ld_32x32(t.tmem + c, v);
      // next chunk's residual loads go out before this chunk is consumed
      if (c + step < t.n_cols) load_chunk(t, c + step, row_base, sub, col4, hn);
      const size_t col = (size_t)t.n0 + c + col4;
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        stg[lane * 8 + (j ^ (lane & 7))] =
            make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                        __uint_as_float(v[4 * j + 3]));
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = sub + 4 * i;
        const int m = row_base + r;
        const float4 a = stg[r * 8 + (j4 ^ (r & 7))];
        h[i].x += a.x;
        h[i].y += a.y;
        h[i].z += a.z;
        h[i].w += a.w;
        ss[i] += h[i].x * h[i].x + h[i].y * h[i].y + h[i].z * h[i].z + h[i].w * h[i].w;
        if (r < t.rows_per_warp && m < t.M) {
          *reinterpret_cast<float4*>(p.h32 + (size_t)m * p.ld + col) = h[i];
          *reinterpret_cast<uint2*>(p.h16 + (size_t)m * p.ld + col) =
              make_uint2(pack_bf16x2(h[i].x, h[i].y), pack_bf16x2(h[i].z, h[i].w));
        }
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = hn[i];
      if (CHUNK_SS) {
        write_ss(t, ss, row_base, sub, (t.n0 + c) >> 5);
#pragma unroll
        for (int i = 0; i < 8; ++i) ss[i] = 0.f;
      }
    }
    if (!CHUNK_SS) write_ss(t, ss, row_base, sub, t.n_blk * t.split + t.part);
  }
  // a row's partial sums sit in the 8 lanes that share `sub`
  __device__ __forceinline__ void write_ss(const TileCtx& t, float (&ss)[8], int row_base, int sub, int part_idx) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = ss[i];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      const int r = sub + 4 * i, m = row_base + r;
      if ((lane & 7) == 0 && r < t.rows_per_warp && m < t.M) p.ss_out[(size_t)part_idx * p.ss_stride + m] = v;
    }
  }
  __device__ void finish() {}
};
using EpiResidual = EpiResidualT<false>;

// The same residual update with the data movement handed to the TMA engine (2-CTA kernel only).
//
// Why: with K = 384 (attention output projection) the MMA of a tile takes ~2 us while its
// epilogue has to move 320 KB per CTA; four warps issuing their own loads keep only ~16 KB in
// flight per SM, and ncu's stall samples sat on the first use of the looked-ahead residual
// registers (the kernel ran at ~3.2 TB/s inside the step).  Here each epilogue warp owns a ring
// of R 32x32 fp32 boxes in shared memory: one lane keeps R-1 box loads in flight (across tile
// boundaries — the next tile's coordinates are known), the warp updates a box in place, row per
// thread (TMEM lane == thread == box row, 128-byte swizzle => conflict-free 16-byte accesses),
// and the fp32 box plus its bf16 copy leave again as two bulk tensor stores.  No thread ever
// waits on a global load or store; rows past M are zero-filled on load and clipped on store.
template <int R>
struct EpiResidualTma {
  struct alignas(64) Params {
    CUtensorMap tm_h32;  // fp32 [M, N], box 32 cols x 32 rows, 128-byte swizzle (load + store)
    CUtensorMap tm_h16;  // bf16 [M, N], box 32 cols x 32 rows,  64-byte swizzle (store)
    float* ss_out;       // [tiles_n][ss_stride]
    int ss_stride;
  };
  static constexpr int kWarps = 4;
  static constexpr int kInBytes = 32 * 32 * 4;
  static constexpr int kOutBytes = 32 * 32 * 2;
  static constexpr int kWarpBytes = R * kInBytes + 2 * kOutBytes;
  static constexpr size_t kSmemBytes = 1024 + (size_t)kWarps * kWarpBytes + 256;
  const Params* pp;
  uint8_t* in;     // this warp's ring
  uint8_t* out16;  // this warp's two bf16 staging boxes
  uint64_t* full;  // this warp's R "box landed" barriers
  int lane, grp;
  uint32_t seq = 0;  // boxes consumed so far (ring slot = seq % R, parity = (seq / R) & 1)
  bool primed = false;

  __device__ EpiResidualTma(const Params& p_, uint8_t* smem_extra, int row, int /*part*/) : pp(&p_) {
    lane = row & 31;
    grp = row >> 5;
    uint8_t* base = smem_extra + ((1024 - (smem_u32(smem_extra) & 1023)) & 1023);
    in = base + grp * kWarpBytes;
    out16 = in + R * kInBytes;
    full = reinterpret_cast<uint64_t*>(base + kWarps * kWarpBytes) + grp * R;
    if (lane == 0) {
      for (int s = 0; s < R; ++s) mbar_init(&full[s], 1);
      fence_mbar_init();
    }
    __syncwarp();
  }
  // Box `idx` counted from the first box of tile t (running on into the next tile) -> ring.
  __device__ __forceinline__ void issue_load(const TileCtx& t, int idx, uint32_t s) const {
    const int nch = t.n_cols >> 5;
    int c0, c1;
    if (idx < nch) {
      c0 = t.n0 + 32 * idx;
      c1 = t.m0 + grp * 32;
    } else {
      idx -= nch;
      if (t.next_m0 < 0 || idx >= (t.next_cols >> 5)) return;
      c0 = t.next_n0 + 32 * idx;
      c1 = t.next_m0 + grp * 32;
    }
    uint64_t* bar = &full[s % R];
    mbar_arrive_expect_tx(bar, kInBytes);
    tma_load_2d(in + (s % R) * kInBytes, &pp->tm_h32, bar, c0, c1);
  }
  __device__ void before_wait(const TileCtx& t) {
    if (primed) return;
    primed = true;
    if (lane == 0)
      for (int k = 0; k < R - 1; ++k) issue_load(t, k, seq + k);
  }
  __device__ void tile(const TileCtx& t) {
    const int nch = t.n_cols >> 5;
    const int m_w = t.m0 + grp * 32;
    const int sw = lane & 7;
    float ss = 0.f;
    for (int ci = 0; ci < nch; ++ci, ++seq) {
      const uint32_t slot = seq % R;
      mbar_wait<0>(&full[slot], (seq / R) & 1, 7);
      uint32_t v[32];
      tmem_ld_32x32(t.tmem + 32 * ci, v);
      tmem_ld_wait();
      float4* hrow = reinterpret_cast<float4*>(in + slot * kInBytes) + lane * 8;
      uint4* orow = reinterpret_cast<uint4*>(out16 + (seq & 1) * kOutBytes) + lane * 4;
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 x = hrow[j ^ sw];
        x.x += __uint_as_float(v[4 * j]);
        x.y += __uint_as_float(v[4 * j + 1]);
        x.z += __uint_as_float(v[4 * j + 2]);
        x.w += __uint_as_float(v[4 * j + 3]);
        ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        hrow[j ^ sw] = x;
        pk[2 * j] = pack_bf16x2(x.x, x.y);
        pk[2 * j + 1] = pack_bf16x2(x.z, x.w);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        orow[q ^ ((lane >> 1) & 3)] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&pp->tm_h32, in + slot * kInBytes, t.n0 + 32 * ci, m_w);
        tma_store_2d(&pp->tm_h16, out16 + (seq & 1) * kOutBytes, t.n0 + 32 * ci, m_w);
        bulk_commit_group();
        // the previous box's stores have drained their shared-memory reads: its ring slot takes the
        // load R-1 boxes ahead, and its bf16 staging box is free for the next iteration
        bulk_wait_group_read<1>();
        issue_load(t, ci + R - 1, seq + R - 1);
      }
      __syncwarp();
    }
    const int m = m_w + lane;
    if (m < t.M) pp->ss_out[(size_t)t.n_blk * pp->ss_stride + m] = ss;
  }
  __device__ void finish() {
    if (lane == 0) bulk_wait_group_read<0>();
    __syncwarp();
  }
};

// gelu_new (tanh form) — HF activations.py NewGELUActivation, used by T5 "gated-gelu".
__device__ __forceinline__ float gelu_new(float x) {
  const float k0 = 0.7978845608028654f;  // sqrt(2/pi)
  const float k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.0f + t);
}

// Gated-GELU FFN up projection (K9): accumulator columns [0, HALF) are the gate and [HALF, 2*HALF) the
// linear branch of the same HALF hidden units (HALF = 128: the 256-wide tiles of the throughput path, whose B
// operand interleaves wi_0 / wi_1 in 128-row blocks; HALF = 64 / 32: the SPLIT_B narrow tiles).
//   out[m, n_blk*HALF + j] = bf16( gelu_new(acc[j]*rs) * (acc[HALF+j]*rs) )
template <int HALF>
struct EpiGeGLUT {
  struct Params {
    __nv_bfloat16* out;  // [M, N/2]
    int ldo;
    RowScale rs;
  };
  static constexpr size_t kSmemBytes = 0;
  static constexpr int kWarps = RPX_EPI_WARPS;
  Params p;
  float rs = 0.f;  // (fetched ahead of the accumulator: see EpiStoreBF16)
  __device__ EpiGeGLUT(const Params& p_, uint8_t*, int, int) : p(p_) {}
  __device__ void before_wait(const TileCtx& t) {
    const int m = t.m0 + t.row;
    rs = m < t.M ? p.rs.get(m) : 0.f;
  }
  __device__ void tile(const TileCtx& t) {
    const int m = t.m0 + t.row;
    const bool ok = m < t.M;
    for (int c = 32 * t.part; c < HALF; c += 32 * t.split) {
      uint32_t g[32], u[32];
      tmem_ld_32x32(t.tmem + c, g);
      tmem_ld_32x32(t.tmem + HALF + c, u);
      tmem_ld_wait();
      if (ok) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)m * p.ldo + t.n_blk * HALF + c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float r[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            r[j] = gelu_new(__uint_as_float(g[8 * i + j]) * rs) * (__uint_as_float(u[8 * i + j]) * rs);
          uint4 o;
          o.x = pack_bf16x2(r[0], r[1]);
          o.y = pack_bf16x2(r[2], r[3]);
          o.z = pack_bf16x2(r[4], r[5]);
          o.w = pack_bf16x2(r[6], r[7]);
          dst[i] = o;
        }
      }
    }
  }
  __device__ void finish() {}
};
using EpiGeGLU = EpiGeGLUT<128>;

}  // namespace rpx
