/* ORACLE — test infrastructure only.  Plain-C CPU restatement of the integer / byte /
 * ranking parts of the reference's premise-retrieval path.  Never linked into or
 * called by the product (reprover_b200/); used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg as the checker.
 *
 * What each function follows:
 *   rpx_oracle_tokenize         HF ByT5Tokenizer as called at reference retrieval/model.py:199-205
 *                               (transformers tokenization_byt5.py:195-208: id = byte + 3, EOS = 1,
 *                               truncation to max_length INCLUDING the EOS)
 *   rpx_oracle_relative_bucket  HF T5Attention._relative_position_bucket (modeling_t5.py:189-234),
 *                               bidirectional, float32 arithmetic as in the torch code
 *   rpx_oracle_dot64            common.py:307 `ctx_emb @ premise_emb.t()` for ONE pair, evaluated in
 *                               fp64 on bf16 operands with the engine's documented summation order
 *   rpx_oracle_sim_topk         common.py:307-322: all similarities, descending order, first k that
 *                               pass the accessibility test; ties -> lower index (deterministic
 *                               refinement of the reference's unspecified argsort tie order)
 *   rpx_oracle_topk_merge       the k-way merge a sharded index needs
 *
 * Pinning: the reference has no tests or golden vectors for this path.  The tokenizer
 * and the bucket function are pinned against the installed HF code in
 * tests/test_oracle_cpu.py (and the fixtures in tests/golden/); dot64 / top-k are pinned
 * against an independent numpy fp64 implementation there.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RPX_ORACLE_API __attribute__((visibility("default")))

static inline double bf16_to_double(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return (double)f;
}

/* Packed tokenisation.  cu_out[n+1] gets the token prefix sums; ids_out (may be NULL to only
 * count) gets cu_out[n] ids. */
RPX_ORACLE_API void rpx_oracle_tokenize(const uint8_t* bytes, const int64_t* offsets, int32_t n, int32_t max_len,
                                        int32_t* ids_out, int32_t* cu_out) {
  int32_t total = 0;
  cu_out[0] = 0;
  for (int32_t s = 0; s < n; ++s) {
    int64_t nb = offsets[s + 1] - offsets[s];
    int64_t keep = nb;
    if (keep + 1 > max_len) keep = max_len - 1; /* truncation leaves room for the EOS */
    if (ids_out) {
      for (int64_t p = 0; p < keep; ++p) ids_out[total + p] = (int32_t)bytes[offsets[s] + p] + 3;
      ids_out[total + keep] = 1; /* </s> */
    }
    total += (int32_t)keep + 1;
    cu_out[s + 1] = total;
  }
}

RPX_ORACLE_API int32_t rpx_oracle_relative_bucket(int32_t relative_position, int32_t num_buckets,
                                                  int32_t max_distance) {
  /* bidirectional=True: half the buckets for positive offsets */
  int32_t nb = num_buckets / 2;
  int32_t bucket = relative_position > 0 ? nb : 0;
  int32_t dist = abs(relative_position);
  int32_t max_exact = nb / 2;
  if (dist < max_exact) return bucket + dist;
  float ratio = logf((float)dist / (float)max_exact) / (float)log((double)max_distance / (double)max_exact);
  int32_t large = max_exact + (int32_t)(ratio * (float)(nb - max_exact));
  if (large > nb - 1) large = nb - 1;
  return bucket + large;
}

/* The canonical summation order of the ordering contract (include/rpx.h, rpx_sim_topk):
 * 32 partial sums; partial l adds, in increasing (j, e), the products of elements
 * (j*32 + l)*8 + e, e = 0..7, for every 8-element chunk j*32 + l < d/8; the partials are then
 * combined by a butterfly with strides 16, 8, 4, 2, 1.  Products of bf16 values are exact in
 * fp64, so each addition rounds once (the GPU's fma does the same). */
RPX_ORACLE_API double rpx_oracle_dot64(const uint16_t* q, const uint16_t* e, int32_t d) {
  double part[32];
  int32_t chunks = d / 8;
  for (int l = 0; l < 32; ++l) {
    double acc = 0.0;
    for (int32_t ch = l; ch < chunks; ch += 32)
      for (int k = 0; k < 8; ++k) {
        double prod = bf16_to_double(q[ch * 8 + k]) * bf16_to_double(e[ch * 8 + k]);
        acc = acc + prod;
      }
    part[l] = acc;
  }
  for (int off = 16; off; off >>= 1) {
    double next[32];
    for (int l = 0; l < 32; ++l) next[l] = part[l] + part[l ^ off];
    memcpy(part, next, sizeof(part));
  }
  return part[0];
}

/* better(a, b): a ranks before b under (score desc, index asc). */
static inline int better(double sa, int64_t ia, double sb, int64_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}

static void insert_sorted(double* bs, int64_t* bi, int32_t* count, int32_t k, double s, int64_t i) {
  int32_t c = *count;
  if (c == k && !better(s, i, bs[k - 1], bi[k - 1])) return;
  int32_t pos = c < k ? c : k - 1;
  while (pos > 0 && better(s, i, bs[pos - 1], bi[pos - 1])) {
    bs[pos] = bs[pos - 1];
    bi[pos] = bi[pos - 1];
    --pos;
  }
  bs[pos] = s;
  bi[pos] = i;
  if (c < k) *count = c + 1;
}

/* Q [nq, d], E [n, d] as raw bf16 bits.  mask: optional bitmask rows of mask_stride uint32 words.
 * Outputs [nq, k]: fp64 scores, global indices (local + idx_offset), -inf / -1 beyond out_count. */
RPX_ORACLE_API void rpx_oracle_sim_topk(const uint16_t* Q, int32_t nq, const uint16_t* E, int64_t n, int32_t d,
                                        int32_t k, const uint32_t* mask, int64_t mask_stride, int64_t idx_offset,
                                        double* out_scores, int64_t* out_idx, int32_t* out_count) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int32_t q = 0; q < nq; ++q) {
    double* bs = out_scores + (size_t)q * k;
    int64_t* bi = out_idx + (size_t)q * k;
    int32_t count = 0;
    for (int64_t i = 0; i < n; ++i) {
      if (mask && !((mask[(size_t)q * mask_stride + (i >> 5)] >> (i & 31)) & 1u)) continue;
      double s = rpx_oracle_dot64(Q + (size_t)q * d, E + (size_t)i * d, d);
      insert_sorted(bs, bi, &count, k, s, i);
    }
    for (int32_t r = 0; r < count; ++r) bi[r] += idx_offset;
    for (int32_t r = count; r < k; ++r) {
      bs[r] = -INFINITY;
      bi[r] = -1;
    }
    if (out_count) out_count[q] = count;
  }
}

/* scores/idx [n_parts, nq, k]; entries with idx < 0 are empty. */
RPX_ORACLE_API void rpx_oracle_topk_merge(const double* scores, const int64_t* idx, int32_t n_parts, int32_t nq,
                                          int32_t k, double* out_scores, int64_t* out_idx, int32_t* out_count) {
  for (int32_t q = 0; q < nq; ++q) {
    double* bs = out_scores + (size_t)q * k;
    int64_t* bi = out_idx + (size_t)q * k;
    int32_t count = 0;
    for (int32_t p = 0; p < n_parts; ++p)
      for (int32_t r = 0; r < k; ++r) {
        size_t src = ((size_t)p * nq + q) * k + r;
        if (idx[src] < 0) continue;
        insert_sorted(bs, bi, &count, k, scores[src], idx[src]);
      }
    for (int32_t r = count; r < k; ++r) {
      bs[r] = -INFINITY;
      bi[r] = -1;
    }
    if (out_count) out_count[q] = count;
  }
}
