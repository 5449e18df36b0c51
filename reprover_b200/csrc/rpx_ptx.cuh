// rpx_ptx.cuh — thin inline-PTX wrappers for the sm_100a features the engine uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), fences.
//
// Everything here is sm_100a-only by design (no multi-arch dispatch).  Waits are
// bounded: a barrier that does not flip within ~2 s of SM clock traps instead of
// hanging the GPU (a hung box costs a gpurun strike).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace rpx {

#define RPX_DEVICE __device__ __forceinline__

RPX_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

RPX_DEVICE uint32_t lane_id() { return threadIdx.x & 31; }

// Returns 1 in exactly one (converged) lane of the warp.
RPX_DEVICE uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xFFFFFFFF;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred));
  return pred;
}

// ----------------------------------------------------------------------------- mbarrier
RPX_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
RPX_DEVICE void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
RPX_DEVICE void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
RPX_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
RPX_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifndef RPX_WAIT_HINT_NS
#define RPX_WAIT_HINT_NS 2000
#endif
// One mbarrier.try_wait probe.  HINT_NS > 0 adds a suspend-time hint: a failed probe may park the
// thread for up to that long (ptxas emits NANOSLEEP.SYNCS), which frees issue slots for the other
// warps of the sub-partition — good for the long waits of the GEMM pipelines (sim kernel: -19 %),
// bad for short latency-critical handshakes (attention), which use HINT_NS = 0.
template <uint32_t HINT_NS>
RPX_DEVICE uint32_t mbar_try_wait_t(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  if constexpr (HINT_NS > 0) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(HINT_NS)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
  return ok;
}
// Bounded wait.  `tag` identifies the call site in the trap message.  The loop body is kept to
// the bare minimum: these spin loops share their SM sub-partition's issue slots with other warps.
template <uint32_t HINT_NS = RPX_WAIT_HINT_NS>
RPX_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait_t<HINT_NS>(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait_t<HINT_NS>(bar, parity)) {
    if ((++spins & 0x3FFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000LL) {  // ~2 s: a pipeline bug, not a slow kernel
        printf("[rpx] mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, (int)blockIdx.x,
               (int)threadIdx.x, parity);
        __trap();
      }
    }
  }
}

// ----------------------------------------------------------------------------- TMA
// 2-D tiled load: coordinates are (c0 = innermost/contiguous dim, c1 = row).
RPX_DEVICE void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0,
                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1)
      : "memory");
}
// Same, with an L2 eviction-priority hint (createpolicy-encoded constants below).
RPX_DEVICE void tma_load_2d_hint(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0,
                                 int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// 2-D tiled store shared -> global, tracked by the issuing thread's bulk async-group.
RPX_DEVICE void tma_store_2d(const void* tmap, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
RPX_DEVICE void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// Until at most N of this thread's bulk groups still have shared-memory reads outstanding.
template <int N>
RPX_DEVICE void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;

RPX_DEVICE void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// ----------------------------------------------------------------------------- tcgen05 / TMEM
// One full warp allocates `ncols` (power of two >= 32) TMEM columns; the base
// address lands in shared memory at `dst`.
RPX_DEVICE void tmem_alloc(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
RPX_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
RPX_DEVICE void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
RPX_DEVICE void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; bf16 inputs, fp32 accumulate.  Issued by ONE thread.
RPX_DEVICE void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on `bar` once every tcgen05.mma issued so far by this thread has retired.
// (Implies tcgen05.fence::before_thread_sync.)  Issued by ONE thread.
RPX_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                   "r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns.
// taddr = (lane_base << 16) | column; lane_base must be 32*(warp_id % 4).
RPX_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 32 consecutive 32-bit columns.
RPX_DEVICE void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
        "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]),
        "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
RPX_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
RPX_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor for a K-major bf16 tile stored as rows of 128 B
// (64 elements) with the 128-byte swizzle TMA writes (CU_TENSOR_MAP_SWIZZLE_128B):
//   start address >> 4            bits [0,14)
//   leading byte offset >> 4      bits [16,30)   (ignored for swizzled K-major; 1)
//   stride byte offset >> 4       bits [32,46)   (8 rows * 128 B = 1024 B between row groups)
//   descriptor version = 1        bits [46,48)   (Blackwell)
//   layout type = 2 (SWIZZLE_128B) bits [61,64)
// The tile base must be 1024-byte aligned.  Advancing along K inside the 128-B
// swizzle atom is done by adding (k_bytes >> 4) to the low word.
RPX_DEVICE uint64_t make_smem_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, dense.
//   c_format  [4,6)  = 1 (f32)      a_format [7,10) = 1 (bf16)   b_format [10,13) = 1 (bf16)
//   a_major 15 = 0,  b_major 16 = 0 (K-major)
//   n_dim [17,23) = N >> 3          m_dim [24,29) = M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---- thread-block clusters: rank, barrier
RPX_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// All threads of all CTAs of the cluster; release/acquire at cluster scope.
RPX_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}


// ----------------------------------------------------------------------------- programmatic dependent launch
// Kernels of one encoder call are launched with programmatic stream serialization (rpx_common.cuh
// launch_pdl): a kernel may start while its predecessor in the stream is still running, so that its
// prologue (barrier init, TMEM allocation, descriptor prefetch) overlaps the predecessor's tail.
// pdl_wait() returns once the predecessor grid has completed and its writes are visible: every kernel
// calls it before the first access to global memory another kernel wrote.  pdl_launch_dependents() lets
// the successor be scheduled once every CTA of this grid has issued it.  Both are no-ops for a kernel
// launched without the attribute.
RPX_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
RPX_DEVICE void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ----------------------------------------------------------------------------- misc
RPX_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace rpx
