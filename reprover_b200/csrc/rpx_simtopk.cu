// rpx_simtopk.cu — fused similarity + top-k, the device half of
// `Corpus.get_nearest_premises` (reference common.py:299-326):
//     similarities = ctx_emb @ premise_emb.T        (:307)
//     argsort(descending) ... first k accessible    (:308-322)
//
// Stage 1  sim_topk (gemm_tc_kernel<..., EpiSimTopk, M_FASTEST>):
//   tcgen05 MMA of a 128-query block against 256-premise tiles streamed once from HBM
//   by TMA; the [Q, N] score matrix is never written.  Each epilogue thread owns one
//   query row in TMEM, compares its 256 scores against that query's running threshold
//   and appends the few survivors (score, index) to a per-(CTA, query) candidate list.
//   When a list fills up, the warp compacts it: a warp-shuffle bisection finds a
//   threshold that keeps ~KEEP best entries, and the threshold rises.
// Stage 2  select_rescore_kernel (one CTA per query):
//   gathers the query's lists from the CTAs that served its block, selects the KEEP
//   best by (fp32 score, index), re-scores those in fp64 with the canonical summation
//   order (bf16 products are exact in fp64) and emits the k best ordered by
//   (fp64 score desc, index asc) — the ordering contract in include/rpx.h.
// Stage 3  topk_merge_kernel: the k-way merge after the multi-GPU all-gather.
#include <math.h>

#include "rpx_gemm_launch.cuh"
#include "rpx_kernels.cuh"

namespace rpx {

namespace {

constexpr int kSimBlockN = 256;
#ifndef RPX_SIM_SAMPLE_TILES
#define RPX_SIM_SAMPLE_TILES 32
#endif
#ifndef RPX_SIM_SEL_SLACK
#define RPX_SIM_SEL_SLACK 16
#endif
constexpr int kSampleTiles = RPX_SIM_SAMPLE_TILES;  // 32 x 256 = 8192 sampled premises
#ifndef RPX_SIM_2CTA
#define RPX_SIM_2CTA 1  // 0: stage 1 always on the 1-CTA kernel
#endif
constexpr int kSelSlack = RPX_SIM_SEL_SLACK;        // stage 2 re-scores between n_res and n_res + kSelSlack rows
constexpr unsigned kFull = 0xffffffffu;

// Monotone map float bits -> uint32 (a > b  <=>  fkey(a) > fkey(b), -0 < +0).
__device__ __forceinline__ uint32_t fkey(uint32_t u) { return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u); }
__device__ __forceinline__ uint32_t unkey(uint32_t k) { return (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k; }

// EPL = candidate entries per lane: a list holds CAP = 32*EPL entries and is compacted back to
// about KEEP (<= CAP/2) when it fills.  KEEP is the size of the candidate superset a CTA
// guarantees for its slice of the corpus.
// Stage-1 epilogue parameters (shared by every EpiSimTopk instantiation).
struct SimTopkParams {
  uint2* cand;           // [grid][128][CAP]  (score bits, local index)
  int32_t* cnt;          // [grid][128]
  uint32_t* gthr;        // [tiles_m*128] shared per-query threshold (monotone key, atomicMax)
  uint32_t* gmin;        // [tiles_m*128] min over CTAs of their first published rank_r-th best key
  uint32_t* gcnt;        // [tiles_m*128] number of CTAs that have published into gmin
  int n_seg;             // CTAs per query block
  int rank_r;            // ceil(KEEP / n_seg)
  int final_max;         // longest list stage 2 accepts (CAP: no final compaction; else KEEP+SLACK)
  const uint32_t* mask;  // optional access bitmask [nq][mask_stride]
  int64_t mask_stride;
  int nq;
  int n;
  int tiles_m;
};

// PAIRED: the epilogue runs inside the 2-CTA (cta_group::2) kernel.  CTA `rank` of pair p serves query
// block 2*(p % (tiles_m/2)) + rank and corpus segment p / (tiles_m/2); lists, counters and stage 2 keep
// the 1-CTA numbering  cta = query_block + segment * tiles_m.
template <int EPL, int KEEP_, bool PAIRED = false>
struct EpiSimTopk {
  static constexpr int CAP = 32 * EPL;
  static constexpr int KEEP = KEEP_;
  static constexpr int SLACK = 16;
  static_assert(KEEP + SLACK + 32 <= CAP, "list too small");
  using Params = SimTopkParams;
  static constexpr size_t kSmemBytes = 0;
  static constexpr int kWarps = 4;  // one warp per TMEM lane group: a query's list has one writer

  Params p;
  float thr;        // pass rule: score > thr
  uint2* wptr;      // next free entry of this thread's list
  uint2* buf;       // this thread's (query's) list
  uint2* warp_buf;  // list of lane 0's query; lane l's list is warp_buf + l*CAP
  int q, lane, slot;
  bool active;
  bool published = false;  // this list has contributed to gmin
  bool have_gmin = false;  // the all-CTA bound has been adopted

  __device__ EpiSimTopk(const Params& p_, uint8_t*, int row, int) : p(p_) {
    lane = row & 31;
    int cta = blockIdx.x;
    if (PAIRED) {
      const int pair = blockIdx.x >> 1, half = p.tiles_m >> 1;
      cta = 2 * (pair % half) + (int)cluster_ctarank() + (pair / half) * p.tiles_m;
    }
    q = (cta % p.tiles_m) * kBlockM + row;
    active = q < p.nq;
    thr = -INFINITY;
    warp_buf = p.cand + ((size_t)cta * kBlockM + (row & ~31)) * CAP;
    buf = warp_buf + (size_t)lane * CAP;
    wptr = buf;
    slot = cta * kBlockM + row;
  }
  __device__ __forceinline__ int count() const { return (int)(wptr - buf); }

  // Branch-free conditional append: every lane executes the same three instructions, the store
  // and the pointer bump are predicated.  (A branchy append serialises the warp once per lane
  // that appends, which made the epilogue ~4x slower than the MMA it has to keep up with.)
  __device__ __forceinline__ void append_if_gt(uint32_t bits, uint32_t idx) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.gt.f32 p, %1, %2;\n"
        "@p st.global.v2.b32 [%0], {%3, %4};\n"
        "@p add.s64 %0, %0, 8;\n"
        "}\n"
        : "+l"(wptr)
        : "f"(__uint_as_float(bits)), "f"(thr), "r"(bits), "r"(idx)
        : "memory");
  }
  __device__ __forceinline__ void append_if_gt_masked(uint32_t bits, uint32_t idx, uint32_t bit) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.gt.f32 p, %1, %2;\n"
        "setp.ne.and.u32 p, %5, 0, p;\n"
        "@p st.global.v2.b32 [%0], {%3, %4};\n"
        "@p add.s64 %0, %0, 8;\n"
        "}\n"
        : "+l"(wptr)
        : "f"(__uint_as_float(bits)), "f"(thr), "r"(bits), "r"(idx), "r"(bit)
        : "memory");
  }

  // Loads lane-strided entries [i*32 + lane] of list `b` (count c) into registers.
  __device__ __forceinline__ void load_list(const uint2* b, int c, uint2 (&e)[EPL]) const {
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int pos = i * 32 + lane;
      e[i] = pos < c ? b[pos] : make_uint2(0u, 0u);
    }
  }

  // Largest T with count(key >= T) >= target, by bisection on [lo, hi) with
  // count(>= lo) >= target > count(>= hi).  Returns T and the count at T.
  __device__ __forceinline__ uint32_t kth_key(const uint32_t (&key)[EPL], int c, int target, int slack,
                                              uint64_t lo, uint64_t hi, int count_lo, int* count_out) const {
    while (count_lo > target + slack && hi - lo > 1) {
      const uint32_t mid = (uint32_t)(lo + (hi - lo) / 2);
      int m = 0;
#pragma unroll
      for (int i = 0; i < EPL; ++i) m += (i * 32 + lane < c && key[i] >= mid) ? 1 : 0;
      m = __reduce_add_sync(kFull, m);
      if (m >= target) {
        lo = mid;
        count_lo = m;
      } else {
        hi = mid;
      }
    }
    *count_out = count_lo;
    return (uint32_t)lo;
  }

  // Warp-cooperative compaction of every list in this warp that holds more than `min_count`
  // entries: a bisection over the monotone score keys finds a threshold that keeps between KEEP
  // and KEEP+SLACK entries (exactly KEEP, lowest indices first, when many scores tie).  The next
  // list's entries are fetched while the current one is processed.
  __device__ void compact_warp(int min_count) {
    const int my_cnt = count();
    unsigned todo = __ballot_sync(kFull, my_cnt > min_count);
    if (todo == 0u) return;
    uint2 e[EPL], en[EPL];
    int src = __ffs(todo) - 1;
    todo &= todo - 1;
    int c = __shfl_sync(kFull, my_cnt, src);
    load_list(warp_buf + (size_t)src * CAP, c, e);
    while (src >= 0) {
      const int nsrc = todo ? __ffs(todo) - 1 : -1;
      todo &= todo - 1;
      int cn = 0;
      if (nsrc >= 0) {
        cn = __shfl_sync(kFull, my_cnt, nsrc);
        load_list(warp_buf + (size_t)nsrc * CAP, cn, en);
      }
      uint2* b = warp_buf + (size_t)src * CAP;
      uint32_t key[EPL];
      uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        if (i * 32 + lane < c) {
          key[i] = fkey(e[i].x);
          kmin = min(kmin, key[i]);
          kmax = max(kmax, key[i]);
        } else {
          key[i] = 0u;
        }
      }
      kmin = __reduce_min_sync(kFull, kmin);
      kmax = __reduce_max_sync(kFull, kmax);
      // invariant: count(key >= lo) = count_lo >= KEEP ; count(key >= hi) < KEEP
      int count_lo;
      const uint32_t lo32 = kth_key(key, c, KEEP, SLACK, kmin, (uint64_t)kmax + 1, c, &count_lo);
      // one-time global bound: the R-th best of this list (R = ceil(KEEP / #CTAs of the block)); once
      // every CTA of the query block has published one, >= KEEP entries exceed their minimum.
      const bool publish = __shfl_sync(kFull, (int)!published, src) != 0;
      uint32_t rth = 0u;
      if (publish) {
        int dummy;
        rth = kth_key(key, c, p.rank_r, 0, lo32, (uint64_t)kmax + 1, count_lo, &dummy);
      }
      const bool tie_mode = count_lo > KEEP + SLACK;  // > SLACK entries share the key `lo`
      int need_eq = 0;
      if (tie_mode) {
        int gt = 0;
#pragma unroll
        for (int i = 0; i < EPL; ++i) gt += (i * 32 + lane < c && key[i] > lo32) ? 1 : 0;
        gt = __reduce_add_sync(kFull, gt);
        need_eq = KEEP - gt;  // > 0 by the invariant (count(>= lo+1) < KEEP)
      }
      const unsigned lt_mask = (1u << lane) - 1u;
      int out = 0, eq_seen = 0;
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        const bool valid = i * 32 + lane < c;
        bool keep;
        if (tie_mode) {
          // lists are filled in increasing index order, so "first" == lowest index
          const bool eq = valid && key[i] == lo32;
          const unsigned eqb = __ballot_sync(kFull, eq);
          keep = valid && (key[i] > lo32 || (eq && eq_seen + __popc(eqb & lt_mask) < need_eq));
          eq_seen += __popc(eqb);
        } else {
          keep = valid && key[i] >= lo32;
        }
        const unsigned kb = __ballot_sync(kFull, keep);
        if (keep) b[out + __popc(kb & lt_mask)] = e[i];
        out += __popc(kb);
      }
      if (lane == src) {
        wptr = buf + out;
        // pass rule is `score > thr`: ties of `lo` are shut out in tie mode (later ones have
        // higher indices than the KEEP entries held), admitted otherwise.
        const float t_new = __uint_as_float(unkey(tie_mode ? lo32 : lo32 - 1u));
        thr = fmaxf(thr, t_new);
        // This CTA holds >= KEEP entries with key >= lo, so no entry with key < lo can be in the
        // query's global top-KEEP: publish the bound for the other CTAs serving this query block.
        atomicMax(p.gthr + q, lo32 - 1u);
        if (publish) {
          atomicMin(p.gmin + q, rth);
          __threadfence();
          atomicAdd(p.gcnt + q, 1u);
          published = true;
        }
      }
      src = nsrc;
      c = cn;
#pragma unroll
      for (int i = 0; i < EPL; ++i) e[i] = en[i];
    }
    __syncwarp();
  }

  __device__ void before_wait(const TileCtx&) {}
  __device__ void tile(const TileCtx& t) {
    // adopt the best bound any CTA of this query block has published so far
    if (active) {
      uint32_t g = __ldcg(p.gthr + q);
      if (!have_gmin && __ldcg(p.gcnt + q) >= (uint32_t)p.n_seg) {
        __threadfence();
        const uint32_t gm = __ldcg(p.gmin + q);  // every CTA holds >= rank_r entries with key >= gm
        g = max(g, gm - 1u);
        have_gmin = true;
      }
      if (g > fkey(__float_as_uint(thr))) thr = __uint_as_float(unkey(g));
    }
    for (int c = 0; c < t.n_cols; c += 32) {
      // only the lists that are actually about to overflow are compacted: after the warm-up (when
      // all 32 fill together) that is typically a single lane, so the other warps are not held up
      if (__any_sync(kFull, count() > CAP - 32)) compact_warp(CAP - 64);
      uint32_t v[32];
      tmem_ld_32x32(t.tmem + c, v);
      tmem_ld_wait();
      const int base = t.n0 + c;
      uint32_t word = 0xFFFFFFFFu;
      if (p.mask != nullptr && active) word = p.mask[(size_t)q * p.mask_stride + (base >> 5)];
      if (base + 32 > p.n) word &= (1u << (p.n - base)) - 1u;  // ragged corpus tail (n - base in 1..31)
      if (!active) word = 0u;
      if (__all_sync(kFull, word == 0xFFFFFFFFu)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) append_if_gt(v[j], (uint32_t)(base + j));
      } else if (__any_sync(kFull, word != 0u)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) append_if_gt_masked(v[j], (uint32_t)(base + j), (word >> j) & 1u);
      }
    }
  }

  __device__ void finish() {
    // stage 2 keeps all lists of a query in shared memory: shorten them only if they would not fit
    if (p.final_max < CAP && __any_sync(kFull, count() > p.final_max)) compact_warp(p.final_max);
    p.cnt[slot] = count();
  }
};

// ------------------------------------------------------------------------------------ stage 0
// Sampling pass: scores of every query against a strided sample of corpus tiles (n_blk_stride > 1)
// are written out ([rows][ld] fp32, -inf where masked / out of range); sample_threshold_kernel
// then takes, per query, the n_res-th best sample score as the starting threshold of stage 1.
// With ~8k sampled premises the main pass admits ~1.4 % of the scores, so its lists hardly ever
// need compacting.  (A threshold from a sample is always valid: at least n_res premises — the
// sampled ones — score at or above it.)
struct EpiSampleScores {
  struct Params {
    float* S;
    int ld;
    const uint32_t* mask;
    int64_t mask_stride;
    int nq;
    int n;
  };
  static constexpr size_t kSmemBytes = 0;
  static constexpr int kWarps = 4;
  Params p;
  __device__ EpiSampleScores(const Params& p_, uint8_t*, int, int) : p(p_) {}
  __device__ void before_wait(const TileCtx&) {}
  __device__ void tile(const TileCtx& t) {
    const int q = t.m0 + t.row;
    const bool active = q < p.nq;
    float* dst = p.S + (size_t)q * p.ld + (size_t)t.n_blk * kSimBlockN;
    for (int c = 0; c < kSimBlockN; c += 32) {
      uint32_t v[32];
      if (c < t.n_cols) {
        tmem_ld_32x32(t.tmem + c, v);
        tmem_ld_wait();
      }
      const int base = t.n0 + c;
      uint32_t word = 0u;
      if (c < t.n_cols) {
        word = 0xFFFFFFFFu;
        if (p.mask != nullptr && active) word = p.mask[(size_t)q * p.mask_stride + (base >> 5)];
        if (base + 32 > p.n) word &= (1u << (p.n - base)) - 1u;
      }
      if (active) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          o.x = ((word >> (j + 0)) & 1u) ? __uint_as_float(v[j + 0]) : -INFINITY;
          o.y = ((word >> (j + 1)) & 1u) ? __uint_as_float(v[j + 1]) : -INFINITY;
          o.z = ((word >> (j + 2)) & 1u) ? __uint_as_float(v[j + 2]) : -INFINITY;
          o.w = ((word >> (j + 3)) & 1u) ? __uint_as_float(v[j + 3]) : -INFINITY;
          *reinterpret_cast<float4*>(dst + c + j) = o;
        }
      }
    }
  }
  __device__ void finish() {}
};

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, T* red, Op op, T identity);

// Block-wide sum of per-thread counts with ONE barrier per call: warp REDUX, one shared-memory
// atomic per warp, three rotating counters (slot i % 3 is used by call i and cleared during call
// i + 1, well before call i + 3 adds to it again).  `slots` must be zero on the first call.
__device__ __forceinline__ int block_count(int m, int* slots, int iter) {
  m = __reduce_add_sync(kFull, m);
  int* cur = slots + iter % 3;
  if ((threadIdx.x & 31) == 0 && m != 0) atomicAdd(cur, m);
  __syncthreads();
  const int total = *cur;
  if (threadIdx.x == 0) slots[(iter + 2) % 3] = 0;
  return total;
}

// One CTA per query: gthr[q] = (n_res-th largest sample key) - 1, or 0 when fewer than n_res
// sampled premises are admissible.
__global__ void __launch_bounds__(256)
sample_threshold_kernel(const float* __restrict__ S, int ld, int n_cols, int n_res, uint32_t* __restrict__ gthr) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  uint32_t* keys = reinterpret_cast<uint32_t*>(sm_raw);
  __shared__ int redi[32];
  __shared__ uint32_t redu[32];
  const int q = blockIdx.x, tid = threadIdx.x;
  const uint32_t kNegInf = fkey(0xFF800000u);
  uint32_t kmax = 0u;
  int valid = 0;
  for (int i = tid; i < n_cols; i += 256) {
    const uint32_t k = fkey(__float_as_uint(S[(size_t)q * ld + i]));
    keys[i] = k;
    kmax = max(kmax, k);
    valid += k > kNegInf ? 1 : 0;
  }
  valid = block_reduce<int>(valid, redi, [](int a, int b) { return a + b; }, 0);
  kmax = block_reduce<uint32_t>(kmax, redu, [](uint32_t a, uint32_t b) { return a > b ? a : b; }, 0u);
  if (valid < n_res) {
    if (tid == 0) gthr[q] = 0u;
    return;
  }
  // a T with n_res <= count(key >= T) <= n_res + 16 (any T with count >= n_res is a valid bound;
  // the slack saves most of the bisection steps).  Start from the smallest valid key.
  uint32_t kmin = 0xFFFFFFFFu;
  for (int i = tid; i < n_cols; i += 256) kmin = (keys[i] > kNegInf && keys[i] < kmin) ? keys[i] : kmin;
  kmin = block_reduce<uint32_t>(kmin, redu, [](uint32_t a, uint32_t b) { return a < b ? a : b; }, 0xFFFFFFFFu);
  uint64_t lo = kmin, hi = (uint64_t)kmax + 1;  // count(>= lo) = valid >= n_res; count(>= hi) = 0
  int count_lo = valid;
  __shared__ int cslots[3];
  if (tid < 3) cslots[tid] = 0;
  __syncthreads();
  for (int iter = 0; count_lo > n_res + 16 && hi - lo > 1; ++iter) {
    const uint32_t mid = (uint32_t)(lo + (hi - lo) / 2);
    int m = 0;
    for (int i = tid; i < n_cols; i += 256) m += keys[i] >= mid ? 1 : 0;
    m = block_count(m, cslots, iter);
    if (m >= n_res) {
      lo = mid;
      count_lo = m;
    } else {
      hi = mid;
    }
  }
  if (tid == 0) gthr[q] = (uint32_t)lo - 1u;
}

// ------------------------------------------------------------------------------------ stage 2

// Canonical fp64 dot product (identical in oracle/rpx_oracle.c::rpx_oracle_dot64):
// lane l accumulates, in increasing j then e order, the elements d = (j*32 + l)*8 + e
// (e = 0..7) with acc = acc + a*b — the bf16 x bf16 product is exact (even in fp32), so this is
// one rounding per addition — and the 32 partials are combined by the xor butterfly
// 16, 8, 4, 2, 1 (p = p + p_partner).
__device__ __forceinline__ double dot64_canonical(const __nv_bfloat16* __restrict__ qrow,  // smem or global
                                                  const __nv_bfloat16* __restrict__ erow, int d, int lane) {
  double acc = 0.0;
  const int chunks = d >> 3;
  // all of this lane's 16-byte loads of the (cold, DRAM-resident) index row go out before the first
  // dependent fma; the summation order is unchanged
  constexpr int kMaxIter = 8;  // d <= 8 * 32 * 8 = 2048 takes the batched path
  if (chunks <= kMaxIter * 32) {
    uint4 ev[kMaxIter];
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
      const int ch = lane + it * 32;
      ev[it] = ch < chunks ? *reinterpret_cast<const uint4*>(erow + ch * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
      const int ch = lane + it * 32;
      if (ch < chunks) {
        const uint4 qv = *reinterpret_cast<const uint4*>(qrow + ch * 8);
        const uint32_t ew[4] = {ev[it].x, ev[it].y, ev[it].z, ev[it].w};
        const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          // bf16 x bf16 is exact in fp32 (8 + 8 significand bits), so the fp32 product converted to
          // fp64 equals the exact product: one F2F per element instead of two (the fp32->fp64
          // conversion pipe, not HBM, was the limiter of this kernel)
          const float p0 = __uint_as_float(qw[w] << 16) * __uint_as_float(ew[w] << 16);
          const float p1 = __uint_as_float(qw[w] & 0xFFFF0000u) * __uint_as_float(ew[w] & 0xFFFF0000u);
          acc += (double)p0;
          acc += (double)p1;
        }
      }
    }
  } else {
    for (int ch = lane; ch < chunks; ch += 32) {
      const uint4 ev = *reinterpret_cast<const uint4*>(erow + ch * 8);
      const uint4 qv = *reinterpret_cast<const uint4*>(qrow + ch * 8);
      const uint32_t ew[4] = {ev.x, ev.y, ev.z, ev.w};
      const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float p0 = __uint_as_float(qw[w] << 16) * __uint_as_float(ew[w] << 16);
        const float p1 = __uint_as_float(qw[w] & 0xFFFF0000u) * __uint_as_float(ew[w] & 0xFFFF0000u);
        acc += (double)p0;
        acc += (double)p1;
      }
    }
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(kFull, acc, off);
  return acc;
}

// Threads per query in stage 2: 512 when few queries are in flight (one query's latency is what
// matters: Q = 1 0.192 vs 0.202 ms, Q = 64 0.166 vs 0.178 ms), 256 when there are enough queries to
// fill the GPU (4 CTAs per SM instead of 2 overlap the staging / bisection / gather phases of different
// queries: Q = 1024 0.607 vs 0.627 ms).
constexpr int kSelThreadsLatency = 512, kSelThreadsThroughput = 256, kSelThroughputMinQueries = 512;
constexpr int kSelMax = 288;  // >= largest re-score set (k + margin + selection slack)

__device__ __forceinline__ uint64_t ckey(uint2 e) {
  // composite: score (monotone) high, ~index low => larger key == better under (score desc, index asc)
  return ((uint64_t)fkey(e.x) << 32) | (uint64_t)(0xFFFFFFFFu - e.y);
}

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, T* red, Op op, T identity) {
  for (int off = 16; off; off >>= 1) v = op(v, __shfl_xor_sync(kFull, v, off));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : identity;
  if (threadIdx.x < 32) {
    for (int off = 16; off; off >>= 1) r = op(r, __shfl_xor_sync(kFull, r, off));
    if (threadIdx.x == 0) red[0] = r;
  }
  __syncthreads();
  r = red[0];
  __syncthreads();
  return r;
}

// One CTA per query.  The query's candidate lists (one per stage-1 CTA that served its block,
// each <= list_max entries) are staged in shared memory as 64-bit composite keys; a bisection
// picks the `n_res` best (+ <= 16), which are re-scored in fp64 and ranked exactly.
template <int kSelThreads>
__global__ void __launch_bounds__(kSelThreads)
select_rescore_kernel(const uint2* __restrict__ cand, const int32_t* __restrict__ cnt, int cap, int list_max,
                      int n_res, int grid_sim, int tiles_m, const __nv_bfloat16* __restrict__ Q,
                      const __nv_bfloat16* __restrict__ E, int d, int k, int64_t idx_offset,
                      float* __restrict__ out_scores, double* __restrict__ out_scores64,
                      int64_t* __restrict__ out_idx, int32_t* __restrict__ out_count) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int n_seg = grid_sim / tiles_m;
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(sm_raw);                                  // [d]
  double* sel_score = reinterpret_cast<double*>(sm_raw + (((size_t)d * 2 + 15) & ~(size_t)15));   // [kSelMax]
  uint32_t* sel_idx = reinterpret_cast<uint32_t*>(sel_score + kSelMax);                            // [kSelMax]
  int* seg_off = reinterpret_cast<int*>(sel_idx + kSelMax);                                        // [n_seg + 1]
  uint64_t* keys = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(seg_off + n_seg + 1) + 15) & ~(uintptr_t)15);                   // [n_seg * list_max]
  __shared__ uint64_t red64[32];
  __shared__ int cslots[3];
  __shared__ int n_sel;

  const int q = blockIdx.x;
  const int q_blk = q / kBlockM, row = q % kBlockM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < d / 8; i += kSelThreads)
    reinterpret_cast<uint4*>(sq)[i] = reinterpret_cast<const uint4*>(Q + (size_t)q * d)[i];
  if (tid < 3) cslots[tid] = 0;
  if (tid == 0) {
    n_sel = 0;
    int acc = 0;
    for (int s = 0; s < n_seg; ++s) {
      seg_off[s] = acc;
      int c = cnt[(q_blk + s * tiles_m) * kBlockM + row];
      acc += c < list_max ? c : list_max;  // (stage 1 guarantees c <= list_max)
    }
    seg_off[n_seg] = acc;
  }
  __syncthreads();
  const int total = seg_off[n_seg];

  // ---- stage the composite keys
  uint64_t kmin = ~0ull, kmax = 0ull;
  for (int s = warp; s < n_seg; s += kSelThreads / 32) {
    const int slot = (q_blk + s * tiles_m) * kBlockM + row;
    const int o = seg_off[s], c = seg_off[s + 1] - o;
    const uint2* b = cand + (size_t)slot * cap;
    for (int i = lane; i < c; i += 32) {
      const uint64_t key = ckey(b[i]);
      keys[o + i] = key;
      kmin = key < kmin ? key : kmin;
      kmax = key > kmax ? key : kmax;
    }
  }
  kmin = block_reduce<uint64_t>(kmin, red64, [](uint64_t a, uint64_t b) { return a < b ? a : b; }, ~0ull);
  kmax = block_reduce<uint64_t>(kmax, red64, [](uint64_t a, uint64_t b) { return a > b ? a : b; }, 0ull);

  // ---- threshold: count(key >= lo) in [n_res, n_res + 16] (keys are distinct, so it exists)
  uint64_t lo = kmin;
  if (total > n_res + kSelSlack) {
    uint64_t hi = kmax;  // count(>= kmax) = 1 < n_res
    int count_lo = total;
    // invariant: count(>= lo) = count_lo >= n_res, count(>= hi) < n_res
    for (int iter = 0; count_lo > n_res + kSelSlack && hi - lo > 1; ++iter) {
      const uint64_t mid = lo + (hi - lo) / 2;
      int m = 0;
      for (int i = tid; i < total; i += kSelThreads) m += keys[i] >= mid ? 1 : 0;
      m = block_count(m, cslots, iter);
      if (m >= n_res) {
        lo = mid;
        count_lo = m;
      } else {
        hi = mid;
      }
    }
  }
  // ---- collect the selected candidates
  for (int i = tid; i < total; i += kSelThreads) {
    const uint64_t key = keys[i];
    if (key >= lo) {
      const int pos = atomicAdd(&n_sel, 1);
      if (pos < kSelMax) sel_idx[pos] = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
    }
  }
  __syncthreads();
  const int ns = n_sel < kSelMax ? n_sel : kSelMax;

  // ---- exact fp64 re-scoring, one warp per candidate
  for (int c = warp; c < ns; c += kSelThreads / 32) {
    const double s = dot64_canonical(sq, E + (size_t)sel_idx[c] * d, d, lane);
    if (lane == 0) sel_score[c] = s;
  }
  __syncthreads();

  // ---- rank by counting under (score desc, index asc); ranks are a permutation
  for (int c = tid; c < ns; c += kSelThreads) {
    const double sc = sel_score[c];
    const uint32_t ic = sel_idx[c];
    int rank = 0;
    for (int j = 0; j < ns; ++j) {
      const double sj = sel_score[j];
      rank += (sj > sc || (sj == sc && sel_idx[j] < ic)) ? 1 : 0;
    }
    if (rank < k) {
      const size_t o = (size_t)q * k + rank;
      out_scores[o] = (float)sc;
      if (out_scores64) out_scores64[o] = sc;
      out_idx[o] = (int64_t)ic + idx_offset;
    }
  }
  const int valid = ns < k ? ns : k;
  for (int r = valid + tid; r < k; r += kSelThreads) {
    const size_t o = (size_t)q * k + r;
    out_scores[o] = -INFINITY;
    if (out_scores64) out_scores64[o] = -INFINITY;
    out_idx[o] = -1;
  }
  if (out_count && tid == 0) out_count[q] = valid;
}

// ------------------------------------------------------------------------------------ stage 3
__global__ void __launch_bounds__(256)
topk_merge_kernel(const double* __restrict__ scores, const int64_t* __restrict__ idx, int n_parts, int nq, int k,
                  float* __restrict__ out_scores, double* __restrict__ out_scores64, int64_t* __restrict__ out_idx,
                  int32_t* __restrict__ out_count) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int n = n_parts * k;
  double* s = reinterpret_cast<double*>(sm_raw);
  int64_t* ix = reinterpret_cast<int64_t*>(s + n);
  __shared__ int n_valid;
  const int q = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) n_valid = 0;
  __syncthreads();
  int local_valid = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    const int part = i / k, r = i % k;
    const size_t src = ((size_t)part * nq + q) * k + r;
    s[i] = scores[src];
    ix[i] = idx[src];
    local_valid += ix[i] >= 0 ? 1 : 0;
  }
  atomicAdd(&n_valid, local_valid);
  __syncthreads();
  // Every part is already sorted under the contract (valid entries first), so the global rank of
  // an entry is its position in its own part plus, for every other part, the number of entries that
  // beat it — found by binary search (R * log2 k steps instead of R * k).
  for (int c = tid; c < n; c += blockDim.x) {
    const int64_t ic = ix[c];
    if (ic < 0) continue;
    const double sc = s[c];
    const int own = c / k;
    int rank = c - own * k;
    for (int p2 = 0; p2 < n_parts; ++p2) {
      if (p2 == own) continue;
      const double* ps = s + p2 * k;
      const int64_t* pi = ix + p2 * k;
      int lo = 0, hi = k;  // first position whose entry does NOT beat (sc, ic)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int64_t im = pi[mid];
        const bool beats = im >= 0 && (ps[mid] > sc || (ps[mid] == sc && im < ic));
        if (beats) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      const size_t o = (size_t)q * k + rank;
      out_scores[o] = (float)sc;
      if (out_scores64) out_scores64[o] = sc;
      out_idx[o] = ic;
    }
  }
  const int valid = n_valid < k ? n_valid : k;
  for (int r = valid + tid; r < k; r += blockDim.x) {
    const size_t o = (size_t)q * k + r;
    out_scores[o] = -INFINITY;
    if (out_scores64) out_scores64[o] = -INFINITY;
    out_idx[o] = -1;
  }
  if (out_count && tid == 0) out_count[q] = valid;
}

struct SimPlan {
  int epl, cap, keep, list_max, n_res;
  size_t sel_smem;
  int tiles_m, grid;
  bool paired;  // stage 1 on the 2-CTA kernel: tiles_m is even, grid = 2 * pairs
  size_t cand_bytes, cnt_bytes, gthr_bytes, sample_bytes, total;
  int sample_tiles;  // corpus tiles scored by the sampling pass (0 = no sampling pass)
};

constexpr size_t kSelSmemBudget = 200 * 1024;

size_t sel_smem_bytes(int d, int n_seg, int list_max) {
  return (((size_t)d * 2 + 15) & ~(size_t)15) + kSelMax * (sizeof(double) + sizeof(uint32_t)) +
         ((size_t)n_seg + 1) * sizeof(int) + 16 + (size_t)n_seg * list_max * sizeof(uint64_t);
}

int plan_sim(int nq, int k, int d, int num_sms, SimPlan* pl) {
  RPX_REQUIRE(k >= 1 && k <= 200, RPX_ERR_UNSUPPORTED, "sim_topk: k=%d outside [1, 200]", k);
  RPX_REQUIRE(nq >= 1, RPX_ERR_INVALID, "sim_topk: nq=%d", nq);
  // per-CTA candidate superset KEEP >= re-score set n_res = k + margin; the margin absorbs
  // fp32 (tensor-core) vs fp64 rank flips at the k-th place
  pl->n_res = k + (k / 8 > 12 ? k / 8 : 12);
  pl->keep = k <= 100 ? 128 : 256;
  pl->epl = 16;
  pl->cap = 32 * pl->epl;
  int chunk_q = nq < num_sms * kBlockM ? nq : num_sms * kBlockM;  // queries per launch
  pl->tiles_m = ceil_div(chunk_q, kBlockM);
  // two or more query blocks: stage 1 runs on CTA pairs (256 queries x 256 premises per tcgen05
  // instruction); an odd block count is padded with an inactive block
  pl->paired = RPX_SIM_2CTA && pl->tiles_m >= 2;
  if (pl->paired) pl->tiles_m += pl->tiles_m & 1;
  int n_seg = pl->paired ? (num_sms / 2) / (pl->tiles_m / 2) : num_sms / pl->tiles_m;
  if (n_seg < 1) n_seg = 1;
  // stage 2 keeps one query's lists in shared memory: full-length lists if they fit, else lists
  // compacted to KEEP+16 at the end of stage 1, else fewer stage-1 CTAs per query block
  // (short lists => small stage-2 footprint => several stage-2 CTAs per SM to hide the gather latency)
  pl->list_max = pl->cap;
  while (pl->list_max > pl->keep + 16 && sel_smem_bytes(d, n_seg, pl->list_max) > (size_t)46 << 10) pl->list_max -= 8;
  while (n_seg > 1 && sel_smem_bytes(d, n_seg, pl->list_max) > kSelSmemBudget) --n_seg;
  pl->grid = n_seg * pl->tiles_m;
  pl->sel_smem = sel_smem_bytes(d, n_seg, pl->list_max);
  pl->cand_bytes = align_up((size_t)pl->grid * kBlockM * pl->cap * sizeof(uint2), 256);
  pl->cnt_bytes = align_up((size_t)pl->grid * kBlockM * sizeof(int32_t), 256);
  pl->gthr_bytes = align_up((size_t)pl->tiles_m * kBlockM * sizeof(uint32_t), 256);
  // sampling pass: up to kSampleTiles tiles of 256 premises, score matrix capped at 64 MB
  pl->sample_tiles = kSampleTiles;
  while (pl->sample_tiles > 4 &&
         (size_t)pl->tiles_m * kBlockM * pl->sample_tiles * kSimBlockN * sizeof(float) > (size_t)64 << 20)
    pl->sample_tiles /= 2;
  pl->sample_bytes = align_up((size_t)pl->tiles_m * kBlockM * pl->sample_tiles * kSimBlockN * sizeof(float), 256);
  pl->total = pl->cand_bytes + pl->cnt_bytes + 3 * pl->gthr_bytes + pl->sample_bytes;  // + gthr, gmin, gcnt
  return RPX_OK;
}

// Launch of stage 1 with the plan's (fixed) tiles_m / grid.  The A tensor map covers the true
// nq rows, so query rows beyond nq are zero-filled by TMA and flagged inactive in the epilogue.
// `tiles_n_override` / `stride`: visit only tiles 0, stride, 2*stride, ... (the sampling pass).
template <class Epi>
int launch_sim_epi(const __nv_bfloat16* Q, int nq, const __nv_bfloat16* E, int64_t n, int d,
                   const typename Epi::Params& ep, const SimPlan& pl, cudaStream_t st, int tiles_n_override = 0,
                   int stride = 1) {
  using Cfg = GemmCfg<kSimBlockN, kGemmStages>;
  DeviceInfo dev;
  RPX_TRY(get_device_info(&dev));
  CUtensorMap tmA, tmB;
  RPX_TRY(make_tmap_bf16_2d(&tmA, Q, (uint64_t)nq, (uint64_t)d, (uint64_t)d, kBlockM));
  RPX_TRY(make_tmap_bf16_2d(&tmB, E, (uint64_t)n, (uint64_t)d, (uint64_t)d, kSimBlockN));
  const int tiles_n = tiles_n_override > 0 ? tiles_n_override : (int)ceil_div64(n, kSimBlockN);
  const size_t smem = Cfg::smem_bytes(Epi::kSmemBytes);
  auto kern = gemm_tc_kernel<kSimBlockN, kGemmStages, Epi, true>;
  static thread_local int configured_dev = -1;
  if (configured_dev != dev.device) {
    RPX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_dev = dev.device;
  }
  int64_t tiles = (int64_t)pl.tiles_m * tiles_n;
  const int grid = tiles < pl.grid ? (int)tiles : pl.grid;  // stays a multiple of tiles_m
  kern<<<grid, gemm_threads<Epi>(), smem, st>>>(tmA, tmB, pl.tiles_m * kBlockM, (int)n, d, pl.tiles_m, tiles_n, stride,
                                                ep);
  RPX_CUDA_OK(cudaGetLastError());
  return RPX_OK;
}

// Stage 1 on the 2-CTA kernel (plan.paired): pairs = (tiles_m / 2) x segments.
template <class Epi>
int launch_sim_epi_paired(const __nv_bfloat16* Q, int nq, const __nv_bfloat16* E, int64_t n, int d,
                          const typename Epi::Params& ep, const SimPlan& pl, cudaStream_t st) {
  using Cfg = Gemm2Cfg<kGemm2Stages>;
  DeviceInfo dev;
  RPX_TRY(get_device_info(&dev));
  CUtensorMap tmA, tmB;
  RPX_TRY(make_tmap_bf16_2d(&tmA, Q, (uint64_t)nq, (uint64_t)d, (uint64_t)d, kBlockM));
  RPX_TRY(make_tmap_bf16_2d(&tmB, E, (uint64_t)n, (uint64_t)d, (uint64_t)d, Cfg::kBlockN / 2));
  const int tiles_m2 = pl.tiles_m / 2;
  const int tiles_n = (int)ceil_div64(n, kSimBlockN);
  const size_t smem = Cfg::smem_bytes(Epi::kSmemBytes);
  RPX_REQUIRE(smem <= dev.smem_optin, RPX_ERR_UNSUPPORTED, "sim: needs %zu B smem, device allows %zu", smem,
              dev.smem_optin);
  auto kern = gemm_tc2_kernel<kGemm2Stages, Epi, true>;
  static thread_local int configured_dev = -1;
  if (configured_dev != dev.device) {
    RPX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_dev = dev.device;
  }
  const int64_t tiles = (int64_t)tiles_m2 * tiles_n;
  const int pairs = tiles < pl.grid / 2 ? (int)tiles : pl.grid / 2;  // stays a multiple of tiles_m2
  kern<<<2 * pairs, gemm_threads<Epi>(), smem, st>>>(tmA, tmB, pl.tiles_m * kBlockM, (int)n, d, tiles_m2, tiles_n, ep);
  RPX_CUDA_OK(cudaGetLastError());
  return RPX_OK;
}

}  // namespace
}  // namespace rpx

using namespace rpx;

extern "C" {

size_t rpx_sim_topk_workspace_bytes(int32_t nq, int32_t k) {
  SimPlan pl;
  // sized for the largest Blackwell SM count so the query works without a device (d only
  // shrinks the plan, so the smallest legal d gives the upper bound)
  if (plan_sim(nq, k, 64, 160, &pl) != RPX_OK) return 0;
  return pl.total + 256;
}

int rpx_sim_topk(const void* d_Q, int32_t nq, const void* d_E, int64_t n, int32_t d, int32_t k,
                 const uint32_t* d_access_mask, int64_t mask_stride_words, float* d_out_scores,
                 double* d_out_scores64, int64_t* d_out_idx, int32_t* d_out_count, int64_t idx_offset,
                 void* d_workspace, size_t workspace_bytes, void* stream) {
  RPX_REQUIRE(d_Q && d_out_scores && d_out_idx && d_workspace, RPX_ERR_INVALID, "rpx_sim_topk: null argument");
  RPX_REQUIRE(d_E != nullptr || n == 0, RPX_ERR_INVALID, "rpx_sim_topk: null index");
  RPX_REQUIRE(n >= 0 && n < (int64_t)INT32_MAX - 512, RPX_ERR_UNSUPPORTED, "rpx_sim_topk: n=%lld out of range", (long long)n);
  RPX_REQUIRE(d > 0 && d % 64 == 0 && d <= 8192, RPX_ERR_UNSUPPORTED, "rpx_sim_topk: d=%d must be a multiple of 64 (<= 8192)", d);
  RPX_REQUIRE(d_access_mask == nullptr || mask_stride_words * 32 >= n, RPX_ERR_INVALID, "rpx_sim_topk: mask stride too small");
  DeviceInfo dev;
  RPX_TRY(get_device_info(&dev));
  SimPlan pl;
  RPX_TRY(plan_sim(nq, k, d, dev.num_sms, &pl));
  RPX_REQUIRE(pl.total <= workspace_bytes, RPX_ERR_WORKSPACE, "rpx_sim_topk: workspace %zu < %zu", workspace_bytes, pl.total);
  RPX_REQUIRE((reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, RPX_ERR_INVALID, "workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint2* cand = reinterpret_cast<uint2*>(d_workspace);
  int32_t* cnt = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(d_workspace) + pl.cand_bytes);
  uint32_t* gthr = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(d_workspace) + pl.cand_bytes + pl.cnt_bytes);
  uint32_t* gcnt = gthr + pl.gthr_bytes / 4;
  uint32_t* gmin = gcnt + pl.gthr_bytes / 4;
  float* sample = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(gmin) + pl.gthr_bytes);
  const int n_seg = pl.grid / pl.tiles_m;
  const int rank_r = ceil_div(pl.keep, n_seg);
  const __nv_bfloat16* Q = static_cast<const __nv_bfloat16*>(d_Q);
  const __nv_bfloat16* E = static_cast<const __nv_bfloat16*>(d_E);
  static thread_local int sel_configured = -1;
  if (sel_configured != dev.device) {
    RPX_CUDA_OK(cudaFuncSetAttribute(select_rescore_kernel<kSelThreadsLatency>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSelSmemBudget + 8 * 1024)));
    RPX_CUDA_OK(cudaFuncSetAttribute(select_rescore_kernel<kSelThreadsThroughput>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSelSmemBudget + 8 * 1024)));
    sel_configured = dev.device;
  }
  const int chunk_q = pl.tiles_m * kBlockM;
  for (int q0 = 0; q0 < nq; q0 += chunk_q) {
    const int nq_c = nq - q0 < chunk_q ? nq - q0 : chunk_q;
    // (only the last chunk can be smaller; its tiles_m may shrink but the plan's grid stays valid
    //  because we keep tiles_m fixed and let the surplus query blocks be empty)
    RPX_CUDA_OK(cudaMemsetAsync(cnt, 0, pl.cnt_bytes + 2 * pl.gthr_bytes, st));  // cnt, gthr, gcnt are adjacent
    RPX_CUDA_OK(cudaMemsetAsync(gmin, 0xFF, pl.gthr_bytes, st));
    const uint32_t* mask_c = d_access_mask ? d_access_mask + (size_t)q0 * mask_stride_words : nullptr;
    const int64_t tiles_n_all = ceil_div64(n, kSimBlockN);
    if (tiles_n_all >= 4 * (int64_t)pl.sample_tiles) {
      // stage 0: starting thresholds from a strided sample of the corpus (overwrites gthr)
      const int ld = pl.sample_tiles * kSimBlockN;
      EpiSampleScores::Params sp{sample, ld, mask_c, mask_stride_words, nq_c, (int)n};
      RPX_TRY((launch_sim_epi<EpiSampleScores>(Q + (size_t)q0 * d, nq_c, E, n, d, sp, pl, st, pl.sample_tiles,
                                                (int)(tiles_n_all / pl.sample_tiles))));
      sample_threshold_kernel<<<nq_c, 256, (size_t)ld * sizeof(uint32_t), st>>>(sample, ld, ld, pl.n_res, gthr);
      RPX_CUDA_OK(cudaGetLastError());
    }
    if (n > 0) {
      SimTopkParams ep{cand, cnt, gthr, gmin, gcnt, n_seg, rank_r, pl.list_max, mask_c, mask_stride_words, nq_c, (int)n, pl.tiles_m};
      if (pl.paired) {
        if (pl.keep == 128) {
          RPX_TRY((launch_sim_epi_paired<EpiSimTopk<16, 128, true>>(Q + (size_t)q0 * d, nq_c, E, n, d, ep, pl, st)));
        } else {
          RPX_TRY((launch_sim_epi_paired<EpiSimTopk<16, 256, true>>(Q + (size_t)q0 * d, nq_c, E, n, d, ep, pl, st)));
        }
      } else if (pl.keep == 128) {
        RPX_TRY((launch_sim_epi<EpiSimTopk<16, 128>>(Q + (size_t)q0 * d, nq_c, E, n, d, ep, pl, st)));
      } else {
        RPX_TRY((launch_sim_epi<EpiSimTopk<16, 256>>(Q + (size_t)q0 * d, nq_c, E, n, d, ep, pl, st)));
      }
    }
    if (nq_c >= kSelThroughputMinQueries) {
      select_rescore_kernel<kSelThreadsThroughput><<<nq_c, kSelThreadsThroughput, pl.sel_smem, st>>>(
          cand, cnt, pl.cap, pl.list_max, pl.n_res, pl.grid, pl.tiles_m, Q + (size_t)q0 * d, E, d, k, idx_offset,
          d_out_scores + (size_t)q0 * k, d_out_scores64 ? d_out_scores64 + (size_t)q0 * k : nullptr,
          d_out_idx + (size_t)q0 * k, d_out_count ? d_out_count + q0 : nullptr);
    } else {
      select_rescore_kernel<kSelThreadsLatency><<<nq_c, kSelThreadsLatency, pl.sel_smem, st>>>(
          cand, cnt, pl.cap, pl.list_max, pl.n_res, pl.grid, pl.tiles_m, Q + (size_t)q0 * d, E, d, k, idx_offset,
          d_out_scores + (size_t)q0 * k, d_out_scores64 ? d_out_scores64 + (size_t)q0 * k : nullptr,
          d_out_idx + (size_t)q0 * k, d_out_count ? d_out_count + q0 : nullptr);
    }
    RPX_CUDA_OK(cudaGetLastError());
  }
  return RPX_OK;
}

int rpx_topk_merge(const double* d_scores64, const int64_t* d_idx, int32_t n_parts, int32_t nq, int32_t k,
                   float* d_out_scores, double* d_out_scores64, int64_t* d_out_idx, int32_t* d_out_count,
                   void* stream) {
  RPX_REQUIRE(d_scores64 && d_idx && d_out_scores && d_out_idx, RPX_ERR_INVALID, "rpx_topk_merge: null argument");
  RPX_REQUIRE(n_parts >= 1 && nq >= 1 && k >= 1, RPX_ERR_INVALID, "rpx_topk_merge: bad sizes");
  const size_t smem = (size_t)n_parts * k * 16;
  RPX_REQUIRE(smem <= 96 * 1024, RPX_ERR_UNSUPPORTED, "rpx_topk_merge: n_parts*k=%d too large", n_parts * k);
  DeviceInfo dev;
  RPX_TRY(get_device_info(&dev));
  static thread_local int configured = -1;
  if (configured != dev.device) {
    RPX_CUDA_OK(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = dev.device;
  }
  topk_merge_kernel<<<nq, 256, smem, static_cast<cudaStream_t>(stream)>>>(d_scores64, d_idx, n_parts, nq, k, d_out_scores,
                                                                       d_out_scores64, d_out_idx, d_out_count);
  RPX_CUDA_OK(cudaGetLastError());
  return RPX_OK;
}

}  // extern "C"
