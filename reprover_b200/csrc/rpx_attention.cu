// rpx_attention.cu — T5 self-attention over packed variable-length sequences.
//
// Replaces HF T5Attention.forward (modeling_t5.py:253-344):
//   scores = q k^T            (NO 1/sqrt(d) scaling in T5)
//          + position_bias    (bucketed relative bias, shared by all layers; K5)
//          + padding mask     (packed layout: keys simply stop at the sequence end)
//   out    = softmax_fp32(scores) v, heads merged to [T, heads*64]
// without ever materialising the [B, heads, L, L] score tensor: flash-style online
// softmax, one CTA per (64-query tile, head, sequence), K/V streamed through a
// double-buffered cp.async ring, QK^T and PV on the tensor cores.
//
// Round-1 note: this kernel uses the legacy warp-level `mma.sync` path (HMMA).
// Attention is 0.5-8 % of the encoder FLOPs; the tcgen05 version
// is listed as follow-up work in DESIGN.md.
#include "rpx_common.cuh"
#include "rpx_kernels.cuh"
#include "rpx_ptx.cuh"

namespace rpx {

namespace {

constexpr int kHD = 64;      // head dim (d_kv)
constexpr int kQT = 64;      // query rows per CTA
constexpr int kKT = 64;      // keys per pipeline step
constexpr int kAttnThreads = 128;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t dst = smem_u32(smem_dst);
  const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A [rows][64] bf16 tile: row r = 128 B = 8 chunks of 16 B; chunk c lives at c ^ (r & 7).
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// Loads `rows` x 64 bf16 from global (row pitch ld elems) into a swizzled tile; rows >= n_valid are zero.
__device__ __forceinline__ void load_tile(uint8_t* tile, const __nv_bfloat16* g, int64_t ld, int n_valid,
                                          int tid) {
  // 64 rows * 8 chunks = 512 chunks; 128 threads -> 4 each
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * kAttnThreads;
    const int r = idx >> 3, c = idx & 7;
    const bool ok = r < n_valid;
    const __nv_bfloat16* src = g + (int64_t)(ok ? r : 0) * ld + c * 8;
    cp_async16(tile + tile_off(r, c), src, ok);
  }
}

__global__ void __launch_bounds__(kAttnThreads)
t5_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                    const int32_t* __restrict__ cu_seqlens, const float* __restrict__ bias_lut,
                    int n_heads, int R, int ld_qkv, int ld_out) {
  const int seq = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
  const int t0 = cu_seqlens[seq];
  const int len = cu_seqlens[seq + 1] - t0;
  const int q0 = qt * kQT;
  if (q0 >= len) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((128 - (raw & 127)) & 127);
  uint8_t* sQ = smem;                 // 8 KB
  uint8_t* sK = smem + 8192;          // 2 x 8 KB
  uint8_t* sV = smem + 8192 * 3;      // 2 x 8 KB
  float* sBias = reinterpret_cast<float*>(smem + 8192 * 5);  // 2R+1 floats

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int inner = n_heads * kHD;

  const __nv_bfloat16* qbase = qkv + (int64_t)t0 * ld_qkv + head * kHD;
  const __nv_bfloat16* kbase = qbase + inner;
  const __nv_bfloat16* vbase = qbase + 2 * inner;

  for (int i = tid; i < 2 * R + 1; i += kAttnThreads) sBias[i] = bias_lut[head * (2 * R + 1) + i];

  load_tile(sQ, qbase + (int64_t)q0 * ld_qkv, ld_qkv, len - q0 < kQT ? len - q0 : kQT, tid);
  cp_async_commit();
  const int n_kt = (len + kKT - 1) / kKT;
  {
    const int nv = len < kKT ? len : kKT;
    load_tile(sK, kbase, ld_qkv, nv, tid);
    load_tile(sV, vbase, ld_qkv, nv, tid);
    cp_async_commit();
  }

  // Q fragments for this warp's 16 rows (4 k-steps of 16 dims)
  cp_async_wait<1>();
  __syncthreads();
  uint32_t qf[4][4];
  {
    const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qf[ks], smem_u32(sQ) + tile_off(r, ks * 2 + (lane >> 4)));
  }

  const float kLog2e = 1.4426950408889634f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

  const int row_lo = q0 + warp * 16 + g;  // sequence-relative query position of c0/c1 (c2/c3: +8)

  for (int kt = 0; kt < n_kt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_kt) {
      const int kb1 = (kt + 1) * kKT;
      const int nv = len - kb1 < kKT ? len - kb1 : kKT;
      load_tile(sK + (buf ^ 1) * 8192, kbase + (int64_t)kb1 * ld_qkv, ld_qkv, nv, tid);
      load_tile(sV + (buf ^ 1) * 8192, vbase + (int64_t)kb1 * ld_qkv, ld_qkv, nv, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    const uint32_t kT = smem_u32(sK + buf * 8192);
    const uint32_t vT = smem_u32(sV + buf * 8192);
    const int kb = kt * kKT;

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int nbp = 0; nbp < 4; ++nbp) {
        uint32_t kf[4];
        const int r = nbp * 16 + (lane & 7) + (lane >> 4) * 8;
        ldmatrix_x4(kf, kT + tile_off(r, ks * 2 + ((lane >> 3) & 1)));
        mma_bf16_16816(s[2 * nbp], qf[ks], kf[0], kf[1]);
        mma_bf16_16816(s[2 * nbp + 1], qf[ks], kf[2], kf[3]);
      }
    }

    // ---- + relative-position bias, key mask, online softmax (fp32)
    float m_new[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kb + nb * 8 + 2 * t + (e & 1);
        const int row = row_lo + (e >> 1) * 8;
        int d = key - row;
        d = d < -R ? -R : (d > R ? R : d);
        float v = s[nb][e] + sBias[d + R];
        v = key < len ? v : -INFINITY;
        s[nb][e] = v;
        m_new[e >> 1] = fmaxf(m_new[e >> 1], v);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      m_new[h] = fmaxf(m_new[h], __shfl_xor_sync(0xffffffffu, m_new[h], 1));
      m_new[h] = fmaxf(m_new[h], __shfl_xor_sync(0xffffffffu, m_new[h], 2));
    }
    float scale[2], mb[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // every tile holds at least one valid key, so m_new is finite from the first tile on
      scale[h] = exp2f((m_run[h] - m_new[h]) * kLog2e);
      mb[h] = m_new[h] * kLog2e;
      m_run[h] = m_new[h];
      l_run[h] *= scale[h];
    }
    uint32_t pf[4][4];  // P as A fragments for the 4 k-steps (16 keys each) of PV
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      float p0 = exp2f(s[nb][0] * kLog2e - mb[0]);
      float p1 = exp2f(s[nb][1] * kLog2e - mb[0]);
      float p2 = exp2f(s[nb][2] * kLog2e - mb[1]);
      float p3 = exp2f(s[nb][3] * kLog2e - mb[1]);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      const int kk = nb >> 1, hi = nb & 1;
      pf[kk][hi * 2 + 0] = pack_bf16x2(p0, p1);
      pf[kk][hi * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i][0] *= scale[0];
      o[i][1] *= scale[0];
      o[i][2] *= scale[1];
      o[i][3] *= scale[1];
    }

    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dbp = 0; dbp < 4; ++dbp) {
        uint32_t vf[4];
        const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldmatrix_x4_trans(vf, vT + tile_off(r, dbp * 2 + (lane >> 4)));
        mma_bf16_16816(o[2 * dbp], pf[kk], vf[0], vf[1]);
        mma_bf16_16816(o[2 * dbp + 1], pf[kk], vf[2], vf[3]);
      }
    }
    __syncthreads();  // all warps done with this buffer before it is refilled
  }

#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int r0 = row_lo, r1 = row_lo + 8;
  __nv_bfloat16* obase = out + (int64_t)t0 * ld_out + head * kHD;
#pragma unroll
  for (int db = 0; db < 8; ++db) {
    const int col = db * 8 + 2 * t;
    if (r0 < len)
      *reinterpret_cast<uint32_t*>(obase + (int64_t)r0 * ld_out + col) = pack_bf16x2(o[db][0] * inv0, o[db][1] * inv0);
    if (r1 < len)
      *reinterpret_cast<uint32_t*>(obase + (int64_t)r1 * ld_out + col) = pack_bf16x2(o[db][2] * inv1, o[db][3] * inv1);
  }
}

}  // namespace

int launch_t5_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, const int32_t* cu_seqlens,
                        const float* bias_lut, int n_seqs, int max_len, int n_heads, int d_kv,
                        int max_distance, cudaStream_t stream) {
  RPX_REQUIRE(d_kv == kHD, RPX_ERR_UNSUPPORTED, "attention: d_kv=%d (only 64 is implemented)", d_kv);
  RPX_REQUIRE(n_seqs > 0 && max_len > 0, RPX_ERR_INVALID, "attention: empty batch");
  RPX_REQUIRE(n_seqs <= 65535 && n_heads <= 65535, RPX_ERR_UNSUPPORTED, "attention: grid limits exceeded");
  const int inner = n_heads * d_kv;
  const size_t smem = 8192 * 5 + (size_t)(2 * max_distance + 1) * sizeof(float) + 128;
  static thread_local int configured = -1;
  int dev = 0;
  RPX_CUDA_OK(cudaGetDevice(&dev));
  if (configured != dev) {
    RPX_CUDA_OK(cudaFuncSetAttribute(t5_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    configured = dev;
  }
  RPX_REQUIRE(smem <= 64 * 1024, RPX_ERR_UNSUPPORTED, "attention: bias table too large");
  dim3 grid((max_len + kQT - 1) / kQT, n_heads, n_seqs);
  t5_attention_kernel<<<grid, kAttnThreads, smem, stream>>>(qkv, out, cu_seqlens, bias_lut, n_heads,
                                                            max_distance, 3 * inner, inner);
  RPX_CUDA_OK(cudaGetLastError());
  return RPX_OK;
}

}  // namespace rpx
