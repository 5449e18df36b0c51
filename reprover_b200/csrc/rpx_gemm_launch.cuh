// rpx_gemm_launch.cuh — host-side launcher for gemm_tc_kernel.
#pragma once
#include "rpx_common.cuh"
#include "rpx_gemm.cuh"
#include "rpx_gemm2.cuh"

namespace rpx {

constexpr int kGemmStages = 4;

// A: [M, K] bf16 (row pitch lda), B: [N, K] bf16 (row pitch ldb).  K % 64 == 0, N % 32 == 0.
// `grid_limit` caps the persistent grid (0 = one CTA per SM).
// M_FASTEST kernels take the grid size verbatim from `grid_limit` (the caller sizes it as a
// multiple of tiles_m) and accept any N (the epilogue masks the ragged tail).
template <int BLOCK_N, class Epi, bool M_FASTEST = false, int STAGES = kGemmStages, bool SPLIT_B = false, int BM = kBlockM>
int launch_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K,
                const typename Epi::Params& ep, cudaStream_t stream, int grid_limit = 0, const void* prefetch_ptr = nullptr,
                size_t prefetch_bytes = 0) {
  using Cfg = GemmCfg<BLOCK_N, STAGES, BM>;
  static_assert(BM == kBlockM || (!M_FASTEST && Epi::kWarps == 4), "64-row tiles: encoder epilogues with one warp per lane group");
  RPX_REQUIRE(M > 0 && N > 0 && K > 0, RPX_ERR_INVALID, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  RPX_REQUIRE(K % kBlockK == 0, RPX_ERR_UNSUPPORTED, "gemm: K=%d must be a multiple of %d", K, kBlockK);
  RPX_REQUIRE(M_FASTEST || N % 32 == 0, RPX_ERR_UNSUPPORTED, "gemm: N=%d must be a multiple of 32", N);
  DeviceInfo dev;
  RPX_TRY(get_device_info(&dev));
  CUtensorMap tmA, tmB;
  RPX_TRY(make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM));
  RPX_TRY(make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, SPLIT_B ? BLOCK_N / 2 : BLOCK_N));
  RPX_REQUIRE(!SPLIT_B || N % BLOCK_N == 0, RPX_ERR_UNSUPPORTED, "gemm: split-B tiles need N %% %d == 0", BLOCK_N);
  const int tiles_m = ceil_div(M, BM);
  const int tiles_n = ceil_div(N, BLOCK_N);
  const size_t smem = Cfg::smem_bytes(Epi::kSmemBytes);
  RPX_REQUIRE(smem <= dev.smem_optin, RPX_ERR_UNSUPPORTED, "gemm: needs %zu B smem, device allows %zu",
              smem, dev.smem_optin);
  auto kern = gemm_tc_kernel<BLOCK_N, STAGES, Epi, M_FASTEST, SPLIT_B, BM>;
  static thread_local int configured_dev = -1;  // per-instantiation, per-thread
  if (configured_dev != dev.device) {
    RPX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_dev = dev.device;
  }
  int grid = tiles_m * tiles_n;
  int cap = grid_limit > 0 ? grid_limit : dev.num_sms;
  if (grid > cap) grid = cap;
  L2Prefetch pf{prefetch_ptr, (uint32_t)prefetch_bytes, grid, next_timeline_slot()};
  if (prefetch_ptr != nullptr && prefetch_bytes > 0 && prefetch_bytes < ((size_t)1 << 32) && grid < dev.num_sms)
    grid = dev.num_sms;  // surplus SMs run prefetch helpers
  RPX_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(gemm_threads<Epi>()), smem, stream, pdl_enabled(), tmA, tmB, M, N, K, tiles_m,
                         tiles_n, 1, ep, pf));
  return RPX_OK;
}


constexpr int kGemm2Stages = 6;

// 2-CTA (cta_group::2) launcher: 256 x 256 tiles, one CTA pair per tile, persistent over num_sms/2 pairs.
template <class Epi, int STAGES = kGemm2Stages>
int launch_gemm2(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K,
                 const typename Epi::Params& ep, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<STAGES>;
  RPX_REQUIRE(M > 0 && N > 0 && K > 0, RPX_ERR_INVALID, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  RPX_REQUIRE(K % kBlockK == 0, RPX_ERR_UNSUPPORTED, "gemm: K=%d must be a multiple of %d", K, kBlockK);
  RPX_REQUIRE(N % 32 == 0, RPX_ERR_UNSUPPORTED, "gemm: N=%d must be a multiple of 32", N);
  DeviceInfo dev;
  RPX_TRY(get_device_info(&dev));
  CUtensorMap tmA, tmB;
  RPX_TRY(make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBlockM));
  RPX_TRY(make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, Cfg::kBlockN / 2));
  const int tiles_m = ceil_div(M, kPairM);
  const int tiles_n = ceil_div(N, Cfg::kBlockN);
  const size_t smem = Cfg::smem_bytes(Epi::kSmemBytes);
  RPX_REQUIRE(smem <= dev.smem_optin, RPX_ERR_UNSUPPORTED, "gemm2: needs %zu B smem, device allows %zu", smem,
              dev.smem_optin);
  auto kern = gemm_tc2_kernel<STAGES, Epi>;
  static thread_local int configured_dev = -1;
  if (configured_dev != dev.device) {
    RPX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_dev = dev.device;
  }
  int pairs = tiles_m * tiles_n;
  if (pairs > dev.num_sms / 2) pairs = dev.num_sms / 2;
  RPX_CUDA_OK(launch_pdl(kern, dim3(2 * pairs), dim3(gemm_threads<Epi>()), smem, stream, pdl_enabled(), tmA, tmB, M, N, K,
                         tiles_m, tiles_n, ep));
  return RPX_OK;
}

}  // namespace rpx
