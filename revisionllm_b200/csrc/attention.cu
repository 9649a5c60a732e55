// Attention kernels for head_dim = 128, no GQA.
//
//  * attn_prefill: causal varlen self-attention over the packed qkv stream (flash-style: K/V tiles of
//    64 keys staged in shared memory with cp.async double buffering, QK^T and PV on the legacy
//    mma.sync tensor path, online softmax in fp32 registers with warp-quad shuffles).  Prefill
//    attention is 0.4 % of the prefill FLOPs at L=184 (8.9 of 2392 GFLOP per segment, BASELINE.md
//    section 3), so it is kept on mma.sync in this round; the GEMMs carry the tcgen05 work.
//  * attn_decode: one query row per (sequence, head) against the paged KV cache; HBM-bound
//    (reads (L+t) * 512 B per (seq, head)), 16-byte coalesced loads of whole key/value rows.
//
// Replaces the eager `matmul -> softmax(fp32) -> matmul` attention of transformers' Llama
// (modeling_llama.py eager_attention_forward) that the reference reaches from
// revisionllm/model/vtimellm_llama.py:79-90, and its DynamicCache concat.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "rvl_internal.h"
#include <cstdlib>
#include "rvl_ptx.cuh"

namespace rvl {

constexpr int kD = 128;        // head_dim
constexpr int kQT = 64;        // query rows per CTA
constexpr int kKT = 64;        // keys per tile

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = smem_u32(smem);
  const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 2^x on the SFU (2 ulp, flushes denormals, -inf -> +0): the probabilities are rounded to bf16 for the PV product anyway
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Blackwell packed fp32 FMA: d.{x,y} = a.{x,y} * b.{x,y} + c.{x,y} in one instruction (FFMA2)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// Tile in smem: [rows][128] bf16, 16-byte chunk c of row r stored at chunk (c ^ (r & 7)).
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return static_cast<uint32_t>(r * 256 + ((c ^ (r & 7)) << 4)); }

// Load the 64 x 128 tile of positions [pos0, pos0 + 64) of one sequence; positions >= L are zero-filled.  A sequence's
// positions [0, ctx_len) live in the rows of its context (a prompt prefix shared with other sequences, computed once),
// positions >= ctx_len in its own rows: `ctx` points at column 0 of position 0, `own_adj` at where position 0 WOULD be if
// the own rows started there (own - ctx_len rows), so both cases are base + pos * row_stride.  Each thread copies chunk
// (tid & 15) of rows (tid >> 4) + 8 i: the swizzled shared-memory address advances by 2 KB per i, and the address math
// is ~10 instructions per cp.async (the first version spent 44, more than half of the kernel's instructions).
__device__ __forceinline__ void load_tile(uint8_t* smem_tile, const __nv_bfloat16* ctx, const __nv_bfloat16* own_adj, int ctx_len,
                                          long long row_stride, int pos0, int L, int tid) {
  const int r0 = tid >> 4, c = tid & 15;
  const uint32_t sbase = smem_u32(smem_tile) + tile_off(r0, c);
  const int p_first = pos0 + r0;
  ctx += c * 8;
  own_adj += c * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int pos = p_first + i * 8;
    const bool ok = pos < L;
    const int pc = ok ? pos : L - 1;           // keep the (unread) source address inside the sequence
    const __nv_bfloat16* src = (pc < ctx_len ? ctx : own_adj) + static_cast<long long>(pc) * row_stride;
    const int sz = ok ? 16 : 0;                // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + i * 2048), "l"(src), "r"(sz) : "memory");
  }
}

__global__ void __launch_bounds__(128, 3) attn_prefill_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            __nv_bfloat16* __restrict__ out,
                                                            const int32_t* __restrict__ cu_seqlens, int n_heads,
                                                            float scale_log2, const int32_t* __restrict__ seq_pos0,
                                                            const int32_t* __restrict__ seq_ctx_row, int only_last) {
  pdl_trigger();
  pdl_wait();
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int s0 = cu_seqlens[seq];
  // optional external context: positions [0, p0) of this sequence are the rows [c0, c0 + p0) of the packed stream (a prompt
  // prefix shared by the batch, projected once); the sequence's own rows hold positions p0 .. L - 1
  const int p0 = seq_pos0 ? seq_pos0[seq] : 0;
  const int c0 = seq_ctx_row ? seq_ctx_row[seq] : 0;
  const int L = p0 + cu_seqlens[seq + 1] - s0;
  const int q0 = qt * kQT;
  if (q0 >= L || q0 + kQT <= p0) return;      // past the end, or a query tile that lies entirely inside the context
  if (only_last && q0 + kQT < L) return;      // last decoder layer of a generation prefill: only the last position is consumed
  extern __shared__ __align__(128) uint8_t smem[];
  // 64 KB: K [2][64][128] | V [2][64][128]; the Q tile borrows K's second buffer until its fragments sit in registers
  // (three CTAs per SM instead of two - the kernel is bound by the latency of each CTA's first loads, not by bandwidth)
  uint8_t* sK = smem;
  uint8_t* sV = smem + 16384 * 2;
  uint8_t* sQ = smem + 16384;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = n_heads * kD;
  const long long stride = 3LL * H;
  const __nv_bfloat16* own = qkv + (static_cast<long long>(s0) - p0) * stride + head * kD;  // where position 0 would be (own rows start at p0), q columns
  const __nv_bfloat16* ctx = qkv + static_cast<long long>(c0) * stride + head * kD;        // position 0 of the context, q columns

  const int n_tiles = min((L + kKT - 1) / kKT, qt + 1);  // causal: keys <= q0 + 63
  load_tile(sQ, ctx, own, p0, stride, q0, L, tid);
  load_tile(sK, ctx + H, own + H, p0, stride, 0, L, tid);
  load_tile(sV, ctx + 2 * H, own + 2 * H, p0, stride, 0, L, tid);
  cp_async_commit();

  uint32_t qf[8][4];
  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int g = lane >> 2, t4 = lane & 3;

  for (int j = 0; j < n_tiles; ++j) {
    const int buf = j & 1;
    if (j == 0) {
      cp_async_wait<0>();
      __syncthreads();
      // Q fragments (A operand, 16 rows of this warp x 16 dims per k-step)
      const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) ldmatrix_x4(qf[ks], smem_u32(sQ) + tile_off(r, ks * 2 + (lane >> 4)));
      if (n_tiles > 1) {
        __syncthreads();                 // every warp holds its Q fragments: K's second buffer may be overwritten
        load_tile(sK + 16384, ctx + H, own + H, p0, stride, kKT, L, tid);
        load_tile(sV + 16384, ctx + 2 * H, own + 2 * H, p0, stride, kKT, L, tid);
        cp_async_commit();
      }
    } else {
      if (j + 1 < n_tiles) {
        const int k0n = (j + 1) * kKT;
        load_tile(sK + (buf ^ 1) * 16384, ctx + H, own + H, p0, stride, k0n, L, tid);
        load_tile(sV + (buf ^ 1) * 16384, ctx + 2 * H, own + 2 * H, p0, stride, k0n, L, tid);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
    }
    const uint32_t kaddr = smem_u32(sK + buf * 16384);
    const uint32_t vaddr = smem_u32(sV + buf * 16384);
    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
        ldmatrix_x4(b, kaddr + tile_off(key, ks * 2 + ((lane >> 3) & 1)));
        mma_bf16_16816(s[2 * np], qf[ks], b[0], b[1]);
        mma_bf16_16816(s[2 * np + 1], qf[ks], b[2], b[3]);
      }
    }
    // ---- mask + online softmax (rows g and g+8 of this warp's 16).  The scores stay unscaled: the scale is positive, so
    // the row maximum is taken on the raw values and 1/sqrt(d) * log2(e) rides in the FFMA in front of each ex2.
    const int k0 = j * kKT;
    const int qrow0 = q0 + warp * 16 + g;  // local query index of c0/c1; c2/c3 are +8
    if (k0 + kKT - 1 > q0 + warp * 16 || k0 + kKT > L) {   // warp-uniform: tiles below the diagonal and inside L need no mask
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = k0 + nt * 8 + t4 * 2 + (e & 1);
          const int qr = qrow0 + (e >> 1) * 8;
          if (!(key <= qr && key < L)) s[nt][e] = -INFINITY;
        }
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
    }
    float corr[2], mneg[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      const float mnew = fmaxf(m_run[h], mx[h] * scale_log2);
      const float msafe = mnew == -INFINITY ? 0.f : mnew;
      corr[h] = ex2_approx(m_run[h] - msafe);  // m_run = -inf -> 0
      m_run[h] = mnew;
      mneg[h] = -msafe;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = ex2_approx(fmaf(s[nt][e], scale_log2, mneg[e >> 1]));
        s[nt][e] = p;
        rs[e >> 1] += p;
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * corr[h] + rs[h];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // ---- O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * ks][0], s[2 * ks][1]);
      pa[1] = pack_bf16x2(s[2 * ks][2], s[2 * ks][3]);
      pa[2] = pack_bf16x2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      pa[3] = pack_bf16x2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 8; ++dp) {
        uint32_t b[4];
        const int key = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldmatrix_x4_trans(b, vaddr + tile_off(key, dp * 2 + (lane >> 4)));
        mma_bf16_16816(o[2 * dp], pa, b[0], b[1]);
        mma_bf16_16816(o[2 * dp + 1], pa, b[2], b[3]);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled (tile j+2)
  }
  // ---- normalise and store
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int qr = q0 + warp * 16 + g + h * 8;
    if (qr < L && qr >= p0) {                  // context positions are produced by the context's own sequence
      const float inv = 1.f / l_run[h];
      __nv_bfloat16* dst = out + (static_cast<long long>(s0) + (qr - p0)) * H + head * kD;
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        *reinterpret_cast<uint32_t*>(dst + nt * 8 + t4 * 2) = pack_bf16x2(o[nt][2 * h] * inv, o[nt][2 * h + 1] * inv);
      }
    }
  }
}

void launch_attn_prefill(const void* qkv, void* out, const int32_t* cu_seqlens, int n_seq, int max_seqlen, int n_heads,
                         cudaStream_t st, const int32_t* seq_pos0, const int32_t* seq_ctx_row, int only_last, int64_t total_tokens,
                         int num_sms) {
  if (n_seq <= 0 || max_seqlen <= 0) return;
  // tcgen05 kernel (attention_tcgen05.cu) unless RVL_ATTN_PREFILL=0 asks for the mma.sync kernel below (= 2: only for
  // sequences that name an external context, the round-2 state before the tensor-core kernel learnt to follow one)
  const int mode = tuning().attn_prefill;
  if (mode != 0 && !(mode == 2 && (seq_pos0 || seq_ctx_row)) && total_tokens > 0 && num_sms > 0 &&
      launch_attn_prefill_tc(qkv, out, cu_seqlens, n_seq, total_tokens, max_seqlen, n_heads, num_sms, st, only_last, seq_pos0, seq_ctx_row))
    return;
  static bool attr = false;
  constexpr int smem = 16384 * 4;
  if (!attr) {
    cudaFuncSetAttribute(attn_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  dim3 grid((max_seqlen + kQT - 1) / kQT, n_heads, n_seq);
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(kD));
  launch_k(attn_prefill_kernel, grid, dim3(128), smem, st, reinterpret_cast<const __nv_bfloat16*>(qkv),
           reinterpret_cast<__nv_bfloat16*>(out), cu_seqlens, n_heads, scale_log2, seq_pos0, seq_ctx_row, only_last);
}

// ------------------------------------------------------------------------------------------- small MHA, head_dim 96
// nn.MultiheadAttention of the stage-2 ClipEncoder (revisionllm/model/adapter/transformer.py:210-223,271-305): 8 heads x 96
// dims, non-causal, fixed Tq / Tk per launch, optional key-padding mask, K / V optionally taken from another sequence index
// (all windows of a query share the query's text tokens).  Same structure as attn_prefill_kernel - 64 query rows per CTA,
// 64-key tiles in shared memory with cp.async double buffering, QK^T and PV on mma.sync, online softmax in registers - with
// six 16-dim k-steps instead of eight; rows keep the 256-byte pitch of the 128-dim tiles so the swizzle is unchanged.
// Replaces the CUDA-core kernel of round 1 (3.7 ms per self-attention launch for 700 windows x 251 tokens).
constexpr int kD96 = 96;

__device__ __forceinline__ void load_tile96(uint8_t* smem_tile, const __nv_bfloat16* base, long long row_stride, int n_valid, int tid) {
  // 64 rows x 12 chunks of 16 B
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int idx = tid + i * 128;
    const int r = idx / 12, c = idx - r * 12;
    const bool ok = r < n_valid;
    cp_async16(smem_tile + tile_off(r, c), base + (ok ? r : 0) * row_stride + c * 8, ok);
  }
}

__global__ void __launch_bounds__(128) mha96_mma_kernel(const __nv_bfloat16* __restrict__ q, long long q_stride,
                                                         const __nv_bfloat16* __restrict__ k, long long k_stride,
                                                         const __nv_bfloat16* __restrict__ v, long long v_stride,
                                                         __nv_bfloat16* __restrict__ out, long long out_stride, int Tq, int Tk,
                                                         const int32_t* __restrict__ kv_seq_idx,
                                                         const float* __restrict__ key_mask, float scale_log2) {
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int q0 = qt * kQT;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sK = smem;                  // [2][64][128 (96 used)]
  uint8_t* sV = smem + 16384 * 2;
  uint8_t* sQ = smem + 16384;          // borrows K's second buffer until the Q fragments are in registers (3 CTAs per SM)
  __shared__ float s_mask[2][kKT];     // additive key mask of the two tiles in flight (0 / -inf)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kv_seq = kv_seq_idx ? kv_seq_idx[seq] : seq;
  const __nv_bfloat16* qbase = q + (static_cast<long long>(seq) * Tq + q0) * q_stride + head * kD96;
  const __nv_bfloat16* kbase = k + static_cast<long long>(kv_seq) * Tk * k_stride + head * kD96;
  const __nv_bfloat16* vbase = v + static_cast<long long>(kv_seq) * Tk * v_stride + head * kD96;
  const float* mbase = key_mask ? key_mask + static_cast<long long>(kv_seq) * Tk : nullptr;
  const int n_tiles = (Tk + kKT - 1) / kKT;
  load_tile96(sQ, qbase, q_stride, min(kQT, Tq - q0), tid);
  load_tile96(sK, kbase, k_stride, min(kKT, Tk), tid);
  load_tile96(sV, vbase, v_stride, min(kKT, Tk), tid);
  cp_async_commit();
  if (tid < kKT) s_mask[0][tid] = (tid < Tk && (!mbase || mbase[tid] != 0.f)) ? 0.f : -INFINITY;

  uint32_t qf[6][4];
  float o[12][4];
#pragma unroll
  for (int i = 0; i < 12; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int g = lane >> 2, t4 = lane & 3;

  for (int j = 0; j < n_tiles; ++j) {
    const int buf = j & 1;
    if (j == 0) {
      cp_async_wait<0>();
      __syncthreads();
      const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int ks = 0; ks < 6; ++ks) ldmatrix_x4(qf[ks], smem_u32(sQ) + tile_off(r, ks * 2 + (lane >> 4)));
      if (n_tiles > 1) {
        __syncthreads();
        load_tile96(sK + 16384, kbase + kKT * k_stride, k_stride, min(kKT, Tk - kKT), tid);
        load_tile96(sV + 16384, vbase + kKT * v_stride, v_stride, min(kKT, Tk - kKT), tid);
        cp_async_commit();
        if (tid < kKT) s_mask[1][tid] = (kKT + tid < Tk && (!mbase || mbase[kKT + tid] != 0.f)) ? 0.f : -INFINITY;
      }
    } else {
      if (j + 1 < n_tiles) {
        const int k0n = (j + 1) * kKT;
        load_tile96(sK + (buf ^ 1) * 16384, kbase + k0n * k_stride, k_stride, min(kKT, Tk - k0n), tid);
        load_tile96(sV + (buf ^ 1) * 16384, vbase + k0n * v_stride, v_stride, min(kKT, Tk - k0n), tid);
        cp_async_commit();
        if (tid < kKT) s_mask[buf ^ 1][tid] = (k0n + tid < Tk && (!mbase || mbase[k0n + tid] != 0.f)) ? 0.f : -INFINITY;
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
    }
    const uint32_t kaddr = smem_u32(sK + buf * 16384);
    const uint32_t vaddr = smem_u32(sV + buf * 16384);
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
        ldmatrix_x4(b, kaddr + tile_off(key, ks * 2 + ((lane >> 3) & 1)));
        mma_bf16_16816(s[2 * np], qf[ks], b[0], b[1]);
        mma_bf16_16816(s[2 * np + 1], qf[ks], b[2], b[3]);
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float add = s_mask[buf][nt * 8 + t4 * 2 + (e & 1)];
        s[nt][e] = s[nt][e] * scale_log2 + add;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      mnew[h] = fmaxf(m_run[h], mx[h]);
      const float msafe = mnew[h] == -INFINITY ? 0.f : mnew[h];
      corr[h] = exp2f(m_run[h] - msafe);
      m_run[h] = mnew[h];
      mnew[h] = msafe;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f(s[nt][e] - mnew[e >> 1]);
        s[nt][e] = p;
        rs[e >> 1] += p;
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * corr[h] + rs[h];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * ks][0], s[2 * ks][1]);
      pa[1] = pack_bf16x2(s[2 * ks][2], s[2 * ks][3]);
      pa[2] = pack_bf16x2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      pa[3] = pack_bf16x2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 6; ++dp) {
        uint32_t b[4];
        const int key = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldmatrix_x4_trans(b, vaddr + tile_off(key, dp * 2 + (lane >> 4)));
        mma_bf16_16816(o[2 * dp], pa, b[0], b[1]);
        mma_bf16_16816(o[2 * dp + 1], pa, b[2], b[3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int qr = q0 + warp * 16 + g + h * 8;
    if (qr < Tq) {
      const float inv = 1.f / l_run[h];
      __nv_bfloat16* dst = out + (static_cast<long long>(seq) * Tq + qr) * out_stride + head * kD96;
#pragma unroll
      for (int nt = 0; nt < 12; ++nt)
        *reinterpret_cast<uint32_t*>(dst + nt * 8 + t4 * 2) = pack_bf16x2(o[nt][2 * h] * inv, o[nt][2 * h + 1] * inv);
    }
  }
}

int launch_mha96(const void* q, long long q_stride, const void* k, long long k_stride, const void* v, long long v_stride, void* out,
                 long long out_stride, int n_seq, int n_heads, int Tq, int Tk, const int32_t* kv_seq_idx, const float* key_mask,
                 cudaStream_t st, int n_kv_seq, int num_sms) {
  // tcgen05 kernel (attention_tcgen05.cu, the head_dim-96 instantiation) when the caller says how many key / value sequences
  // there are (the tensor maps need the row count) and the operands are 16-byte aligned; RVL_ATTN_MHA96=0: the kernel above
  if (!kv_seq_idx) n_kv_seq = n_seq;
  if (tuning().attn_mha96 != 0 && n_kv_seq > 0 && num_sms > 0 &&
      launch_mha96_tc(q, q_stride, k, k_stride, v, v_stride, out, out_stride, n_seq, n_kv_seq, n_heads, Tq, Tk, kv_seq_idx, key_mask, num_sms, st))
    return cudaGetLastError() == cudaSuccess ? RVL_OK : RVL_ERR_CUDA;
  static bool attr = false;
  constexpr int smem = 16384 * 4;
  if (!attr) {
    if (cudaFuncSetAttribute(mha96_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return RVL_ERR_CUDA;
    attr = true;
  }
  dim3 grid((Tq + kQT - 1) / kQT, n_heads, n_seq);
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(kD96));
  mha96_mma_kernel<<<grid, 128, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(q), q_stride, reinterpret_cast<const __nv_bfloat16*>(k),
                                            k_stride, reinterpret_cast<const __nv_bfloat16*>(v), v_stride,
                                            reinterpret_cast<__nv_bfloat16*>(out), out_stride, Tq, Tk, kv_seq_idx, key_mask, scale_log2);
  return cudaGetLastError() == cudaSuccess ? RVL_OK : RVL_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------- decode
// grid (n_heads, n_seq), 128 threads: one query row per (sequence, head) against the paged KV cache.
//   phase 0 (kFused): RoPE of this step's q and k (rotate-half, fp32 sincosf like LlamaRotaryEmbedding, results rounded
//            to bf16 exactly as rope_kv_kernel does) and append of k', v to the cache page - the separate RoPE / KV
//            kernel and its launch disappear from the decode step;
//   phase 1: scores = q . k_j, 8 threads per key (16 dims each), 16 keys per CTA iteration, four iterations' loads
//            in flight before the first FMA;
//   phase 2: softmax over the row in shared memory;
//   phase 3: out = sum p_j v_j, 16 threads per key cover 128 dims, 8 keys per iteration (eight loads in flight),
//            cross-group reduction in shared memory.
// Algorithmic bytes per (seq, head): n_keys * 128 * 2 * 2.  HBM-bound: reads whole 256-byte K / V rows.
// kU: 16-key groups whose K loads are in flight together (phase 1) - and 2 kU 8-key groups of V (phase 3).  kU = 4 (64 registers,
// 8 CTAs per SM) is the measured optimum on B200: kU = 12 (one round trip per context, 128 registers, 4 CTAs per SM) ran
// the 7B decode step at 10.7 ms against 9.3 ms.
constexpr int kPtSmem = 128;     // page ids kept in shared memory by the register decode kernel (4096 tokens at 32 per page)
template <bool kFused, int kU = 4, int kPS = 0>
__global__ void __launch_bounds__(128, 8) attn_decode_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                           const int32_t* __restrict__ seq_lens,
                                                           const int32_t* __restrict__ page_table, int max_pages,
                                                           __nv_bfloat16* k_pages, __nv_bfloat16* v_pages, int n_heads,
                                                           int page_size, float scale, float theta) {
  extern __shared__ float s_scores[];          // [n_keys_max]
  __shared__ float s_red[8][kD];
  __shared__ float s_stat[8];
  __shared__ __align__(16) float s_q[kD];               // this step's (rotated) query, bf16-rounded values
  __shared__ int32_t s_pt[kPtSmem];                     // this sequence's page ids: one global read instead of one per key group
  pdl_trigger();
  pdl_wait();
  const int head = blockIdx.x, seq = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pos = seq_lens[seq];               // tokens already cached = position of this step's token
  const int n_keys = pos + 1;                  // includes the token appended this step
  const int H = n_heads * kD;
  const int32_t* pt = page_table + static_cast<long long>(seq) * max_pages;
  // kPS: compile-time page size (32 in every shipped configuration) - shifts and masks instead of integer divisions in
  // front of every K / V load; 0 = use the runtime value
  const int ps = kPS ? kPS : page_size;
  const int n_pages = (n_keys + ps - 1) / ps;
  const bool pt_smem = n_pages <= kPtSmem;
  if (pt_smem && tid < n_pages) s_pt[tid] = pt[tid];     // visible after the __syncthreads of phase 0
  auto page_of = [&](int key) { const int i = key / ps; return pt_smem ? s_pt[i] : pt[i]; };
  const __nv_bfloat16* row = qkv + static_cast<long long>(seq) * 3 * H + head * kD;

  // ---- phase 0
  if (kFused) {
    // the rotated key goes straight to its cache slot and v is copied there, so that after ONE barrier every thread finds
    // this step's token in the cache like any other key (plain loads below: the read-only path would not see these writes)
    const int page = pt[pos / ps];
    const long long slot = ((static_cast<long long>(page) * n_heads + head) * ps + pos % ps) * kD;
    if (tid < kD / 2) {
      const float inv_freq = 1.0f / powf(theta, static_cast<float>(2 * tid) / static_cast<float>(kD));
      float sn, cs;
      sincosf(static_cast<float>(pos) * inv_freq, &sn, &cs);
      const float q0 = __bfloat162float(row[tid]), q1 = __bfloat162float(row[tid + 64]);
      const float k0 = __bfloat162float(row[H + tid]), k1 = __bfloat162float(row[H + tid + 64]);
      s_q[tid] = __bfloat162float(__float2bfloat16(q0 * cs - q1 * sn));
      s_q[tid + 64] = __bfloat162float(__float2bfloat16(q1 * cs + q0 * sn));
      k_pages[slot + tid] = __float2bfloat16(k0 * cs - k1 * sn);
      k_pages[slot + tid + 64] = __float2bfloat16(k1 * cs + k0 * sn);
    } else if (tid < kD / 2 + 16) {
      reinterpret_cast<uint4*>(v_pages + slot)[tid - kD / 2] = reinterpret_cast<const uint4*>(row + 2 * H)[tid - kD / 2];
    }
  } else {
    s_q[tid] = __bfloat162float(row[tid]);
  }
  __syncthreads();

  // ---- phase 1
  {
    const int dg = lane & 7;                   // dims [16 dg, 16 dg + 16)
    float qf[16];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 t = reinterpret_cast<const float4*>(s_q)[dg * 4 + e];
      qf[4 * e] = t.x; qf[4 * e + 1] = t.y; qf[4 * e + 2] = t.z; qf[4 * e + 3] = t.w;
    }
    const int ksub = warp * 4 + (lane >> 3);   // key inside a 16-key iteration
    for (int kbase = 0; kbase < n_keys; kbase += 16 * kU) {
      uint4 kv[kU][2];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int key = kbase + u * 16 + ksub;
        kv[u][0] = kv[u][1] = make_uint4(0u, 0u, 0u, 0u);
        if (key < n_keys) {
          const int page = page_of(key);
          const uint4* kp = reinterpret_cast<const uint4*>(k_pages + ((static_cast<long long>(page) * n_heads + head) * ps + key % ps) * kD + dg * 16);
          kv[u][0] = kp[0]; kv[u][1] = kp[1];
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int key = kbase + u * 16 + ksub;
        const uint32_t kw[8] = {kv[u][0].x, kv[u][0].y, kv[u][0].z, kv[u][0].w, kv[u][1].x, kv[u][1].y, kv[u][1].z, kv[u][1].w};
        float2 acc2 = make_float2(0.f, 0.f);       // even / odd dims accumulate side by side in one FFMA2 per bf16 pair
#pragma unroll
        for (int e = 0; e < 8; ++e) acc2 = ffma2(make_float2(qf[2 * e], qf[2 * e + 1]), make_float2(bf16_lo(kw[e]), bf16_hi(kw[e])), acc2);
        float acc = acc2.x + acc2.y;
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (dg == 0 && key < n_keys) s_scores[key] = acc * scale;
      }
    }
  }
  __syncthreads();
  // ---- phase 2: softmax statistics
  float mx = -INFINITY;
  for (int i = tid; i < n_keys; i += 128) mx = fmaxf(mx, s_scores[i]);
  mx = warp_max(mx);
  if (lane == 0) s_stat[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_stat[0], s_stat[1]), fmaxf(s_stat[2], s_stat[3]));
  float sum = 0.f;
  for (int i = tid; i < n_keys; i += 128) {
    const float p = __expf(s_scores[i] - mx);
    s_scores[i] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_stat[4 + warp] = sum;
  __syncthreads();
  const float inv = 1.f / (s_stat[4] + s_stat[5] + s_stat[6] + s_stat[7]);
  // ---- phase 3: PV
  {
    const int grp = tid >> 4;                  // 8 key groups
    const int dv = (tid & 15) * 8;             // dims [dv, dv + 8)
    float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    for (int kbase = 0; kbase < n_keys; kbase += 16 * kU) {
      uint4 vv[2 * kU];
      float pp[2 * kU];
#pragma unroll
      for (int u = 0; u < 2 * kU; ++u) {
        const int key = kbase + u * 8 + grp;
        vv[u] = make_uint4(0u, 0u, 0u, 0u);
        pp[u] = 0.f;
        if (key < n_keys) {
          const int page = page_of(key);
          vv[u] = *reinterpret_cast<const uint4*>(v_pages + ((static_cast<long long>(page) * n_heads + head) * ps + key % ps) * kD + dv);
          pp[u] = s_scores[key];
        }
      }
#pragma unroll
      for (int u = 0; u < 2 * kU; ++u) {
        const float2 p = make_float2(pp[u], pp[u]);
        acc[0] = ffma2(p, make_float2(bf16_lo(vv[u].x), bf16_hi(vv[u].x)), acc[0]);
        acc[1] = ffma2(p, make_float2(bf16_lo(vv[u].y), bf16_hi(vv[u].y)), acc[1]);
        acc[2] = ffma2(p, make_float2(bf16_lo(vv[u].z), bf16_hi(vv[u].z)), acc[2]);
        acc[3] = ffma2(p, make_float2(bf16_lo(vv[u].w), bf16_hi(vv[u].w)), acc[3]);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { s_red[grp][dv + 2 * e] = acc[e].x; s_red[grp][dv + 2 * e + 1] = acc[e].y; }
  }
  __syncthreads();
  {
    float r = 0.f;
#pragma unroll
    for (int gi = 0; gi < 8; ++gi) r += s_red[gi][tid];
    out[static_cast<long long>(seq) * H + head * kD + tid] = __float2bfloat16(r * inv);
  }
}

// ------------------------------------------------------------------------------------------- decode, staged
// Same arithmetic, different data movement: the cached K and V rows of the (sequence, head) are fetched with bulk async
// copies (one per page and per K / V: a (page, head) chunk is 8 KB contiguous) into shared memory, everything in flight
// at once and without occupying registers; QK^T starts when K has landed while V is still arriving.  The register-
// staged kernel above makes six dependent HBM round trips per CTA (three 64-key rounds for K, three for V) and was
// measured at 108 us per layer for 180 x 32 rows of ~190 keys (0.69 of HBM peak, latency- not bandwidth-bound: mapping
// 1/6 of the reads to L2-resident shared pages changed nothing).  Contexts longer than the buffer are processed in
// chunks with an online softmax.
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <bool kFused>
__global__ void __launch_bounds__(128) attn_decode_staged_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                                  const int32_t* __restrict__ seq_lens,
                                                                  const int32_t* __restrict__ page_table, int max_pages,
                                                                  __nv_bfloat16* k_pages, __nv_bfloat16* v_pages, int n_heads,
                                                                  int page_size, float scale, float theta, int cap_keys) {
  extern __shared__ __align__(128) uint8_t dsm[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(dsm);            // [cap_keys][128]
  __nv_bfloat16* sV = sK + static_cast<size_t>(cap_keys) * kD;           // [cap_keys][128]
  float* s_scores = reinterpret_cast<float*>(sV + static_cast<size_t>(cap_keys) * kD);   // [cap_keys]
  __shared__ float s_red[8][kD];
  __shared__ float s_stat[8];
  __shared__ __align__(16) float s_q[kD];
  __shared__ __align__(16) __nv_bfloat16 s_knew[kD];
  __shared__ __align__(8) uint64_t bars[2];
  pdl_trigger();
  pdl_wait();
  const int head = blockIdx.x, seq = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pos = seq_lens[seq];
  const int n_keys = pos + 1;
  const int n_cached = kFused ? pos : n_keys;     // rows to fetch from the cache (the fused kernel produces row `pos` itself)
  const int H = n_heads * kD;
  const int32_t* pt = page_table + static_cast<long long>(seq) * max_pages;
  const __nv_bfloat16* row = qkv + static_cast<long long>(seq) * 3 * H + head * kD;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int grp = tid >> 4;                  // phase 3: 8 key groups
  const int dv = (tid & 15) * 8;             //          dims [dv, dv + 8)
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float m_run = -INFINITY, l_run = 0.f;
  uint32_t phase = 0;
  for (int k0 = 0; k0 < n_keys; k0 += cap_keys) {
    const int nk = min(cap_keys, n_keys - k0);
    const int nk_cached = max(0, min(nk, n_cached - k0));
    if (tid == 0) {
      fence_proxy_async();                   // earlier generic reads of the buffers precede the async writes
      mbar_arrive_expect_tx(&bars[0], static_cast<uint32_t>(nk_cached) * 256u);
      mbar_arrive_expect_tx(&bars[1], static_cast<uint32_t>(nk_cached) * 256u);
    }
    {
      // one thread per (K | V, page piece) issues its copy, so the page-table reads run in parallel (a single issuing
      // thread spent ~500 cycles per copy waiting for its page id).  A copy may complete before thread 0's expect_tx:
      // the phase cannot end early because thread 0's arrival is still pending.
      const int first_slot = k0 % page_size;
      const int n_pieces = nk_cached > 0 ? (first_slot + nk_cached + page_size - 1) / page_size : 0;
      for (int i = tid; i < 2 * n_pieces; i += 128) {
        const int pass = i >= n_pieces ? 1 : 0;
        const int piece = i - pass * n_pieces;
        const int kk = piece == 0 ? 0 : piece * page_size - first_slot;      // first key (chunk-relative) of the piece
        const int key = k0 + kk;
        const int slot = key % page_size;
        const int n = min(page_size - slot, nk_cached - kk);
        const int page = pt[key / page_size];
        const __nv_bfloat16* pages = pass ? v_pages : k_pages;
        bulk_copy_g2s((pass ? sV : sK) + static_cast<size_t>(kk) * kD,
                      pages + ((static_cast<long long>(page) * n_heads + head) * page_size + slot) * kD,
                      static_cast<uint32_t>(n) * 256u, &bars[pass]);
      }
    }
    if (k0 == 0) {
      // ---- phase 0: this step's query (and, fused, RoPE + KV append)
      if (kFused) {
        if (tid < kD / 2) {
          const float inv_freq = 1.0f / powf(theta, static_cast<float>(2 * tid) / static_cast<float>(kD));
          float sn, cs;
          sincosf(static_cast<float>(pos) * inv_freq, &sn, &cs);
          const float q0 = __bfloat162float(row[tid]), q1 = __bfloat162float(row[tid + 64]);
          const float kx = __bfloat162float(row[H + tid]), ky = __bfloat162float(row[H + tid + 64]);
          s_q[tid] = __bfloat162float(__float2bfloat16(q0 * cs - q1 * sn));
          s_q[tid + 64] = __bfloat162float(__float2bfloat16(q1 * cs + q0 * sn));
          s_knew[tid] = __float2bfloat16(kx * cs - ky * sn);
          s_knew[tid + 64] = __float2bfloat16(ky * cs + kx * sn);
        }
      } else {
        s_q[tid] = __bfloat162float(row[tid]);
      }
      __syncthreads();
      if (kFused) {
        const int page = pt[pos / page_size];
        const long long slot = ((static_cast<long long>(page) * n_heads + head) * page_size + pos % page_size) * kD;
        if (tid < 16) reinterpret_cast<uint4*>(k_pages + slot)[tid] = reinterpret_cast<const uint4*>(s_knew)[tid];
        else if (tid < 32) reinterpret_cast<uint4*>(v_pages + slot)[tid - 16] = reinterpret_cast<const uint4*>(row + 2 * H)[tid - 16];
      }
    }
    if (kFused && pos >= k0 && pos < k0 + nk) {
      // the new token's row of this chunk comes from registers / qkv, not from the cache
      if (tid < 16) reinterpret_cast<uint4*>(sK + static_cast<size_t>(pos - k0) * kD)[tid] = reinterpret_cast<const uint4*>(s_knew)[tid];
      else if (tid < 32) reinterpret_cast<uint4*>(sV + static_cast<size_t>(pos - k0) * kD)[tid - 16] = reinterpret_cast<const uint4*>(row + 2 * H)[tid - 16];
    }
    __syncthreads();
    // ---- phase 1: scores (8 threads per key, each dims [8 dg, 8 dg + 8) and [64 + 8 dg, ...): conflict-free 128 B runs)
    mbar_wait(&bars[0], phase);
    {
      const int dg = lane & 7;
      float qf[16];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float4 t0 = reinterpret_cast<const float4*>(s_q)[e * 16 + dg * 2];
        const float4 t1 = reinterpret_cast<const float4*>(s_q)[e * 16 + dg * 2 + 1];
        qf[8 * e] = t0.x; qf[8 * e + 1] = t0.y; qf[8 * e + 2] = t0.z; qf[8 * e + 3] = t0.w;
        qf[8 * e + 4] = t1.x; qf[8 * e + 5] = t1.y; qf[8 * e + 6] = t1.z; qf[8 * e + 7] = t1.w;
      }
      const int ksub = warp * 4 + (lane >> 3);
      for (int kb = 0; kb < nk; kb += 16) {
        const int key = kb + ksub;
        float a = 0.f;
        if (key < nk) {
          const uint4* kp = reinterpret_cast<const uint4*>(sK + static_cast<size_t>(key) * kD);
          const uint4 ka = kp[dg], kb2 = kp[8 + dg];
          const uint32_t kw[8] = {ka.x, ka.y, ka.z, ka.w, kb2.x, kb2.y, kb2.z, kb2.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) a += qf[2 * e] * bf16_lo(kw[e]) + qf[2 * e + 1] * bf16_hi(kw[e]);
        }
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (dg == 0 && key < nk) s_scores[key] = a * scale;
      }
    }
    __syncthreads();
    // ---- phase 2: (online) softmax statistics of the chunk
    float mx = -INFINITY;
    for (int i = tid; i < nk; i += 128) mx = fmaxf(mx, s_scores[i]);
    mx = warp_max(mx);
    if (lane == 0) s_stat[warp] = mx;
    __syncthreads();
    const float m_new = fmaxf(m_run, fmaxf(fmaxf(s_stat[0], s_stat[1]), fmaxf(s_stat[2], s_stat[3])));
    float sum = 0.f;
    for (int i = tid; i < nk; i += 128) {
      const float p = __expf(s_scores[i] - m_new);
      s_scores[i] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    if (lane == 0) s_stat[4 + warp] = sum;
    __syncthreads();
    const float corr = __expf(m_run - m_new);      // first chunk: exp(-inf) = 0
    l_run = l_run * corr + (s_stat[4] + s_stat[5] + s_stat[6] + s_stat[7]);
    m_run = m_new;
    // ---- phase 3: PV
    mbar_wait(&bars[1], phase);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll 4
    for (int key = grp; key < nk; key += 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(sV + static_cast<size_t>(key) * kD + dv);
      const float p = s_scores[key];
      acc[0] += p * bf16_lo(v.x); acc[1] += p * bf16_hi(v.x);
      acc[2] += p * bf16_lo(v.y); acc[3] += p * bf16_hi(v.y);
      acc[4] += p * bf16_lo(v.z); acc[5] += p * bf16_hi(v.z);
      acc[6] += p * bf16_lo(v.w); acc[7] += p * bf16_hi(v.w);
    }
    phase ^= 1;
    __syncthreads();                             // buffers and scores are free for the next chunk
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_red[grp][dv + e] = acc[e];
  __syncthreads();
  {
    float r = 0.f;
#pragma unroll
    for (int gi = 0; gi < 8; ++gi) r += s_red[gi][tid];
    out[static_cast<long long>(seq) * H + head * kD + tid] = __float2bfloat16(r / l_run);
  }
}

// ------------------------------------------------------------------------------------------- decode, tensor-core tiles
// Same arithmetic once more, few instructions per byte: inside the 1-hour sweep the power-capped SM clock (1.1 - 1.3 GHz)
// makes the register kernel above SM-bound (115 us per layer in the step against 88 us at 1.9 GHz under ncu: every 16 bytes
// of K / V cost ~25 issue slots of bf16 unpacking and FFMA2 there).  Here 64-key K / V tiles stream through a double-buffered
// shared-memory ring with cp.async (straight out of the paged cache: a (page, head) chunk is 8 KB contiguous, two pages per
// tile) and both products run on mma.sync.m16n8k16 with the single query row padded to a 16-row tile: each of the four warps
// owns 16 keys of every tile (8 ldmatrix + 16 mma for its scores, the same for its share of P V), keeps its own online-softmax
// state, and the four partial (max, sum, output) triples are merged once at the end.  15 of the 16 tile rows are wasted
// tensor work - the tensor pipe is idle in a decode step anyway; the point is ~4x fewer issued instructions per byte.
// Page size 32 only.  P is rounded to bf16 for the P V product exactly like the prefill kernel does.
template <bool kFused>
__global__ void __launch_bounds__(128, 3) attn_decode_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                               const int32_t* __restrict__ seq_lens,
                                                               const int32_t* __restrict__ page_table, int max_pages,
                                                               __nv_bfloat16* k_pages, __nv_bfloat16* v_pages, int n_heads,
                                                               float scale_log2, float theta) {
  constexpr int kPS = 32;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sK = smem;                        // [2][64][128] bf16, swizzled like the prefill tiles
  uint8_t* sV = smem + 32768;                // [2][64][128]
  uint8_t* sQ = smem + 65536;                // [16][128]: row 0 = this step's query, rows 1 - 15 zero
  float* s_o = reinterpret_cast<float*>(smem + 65536 + 4096);   // [4 warps][128] partial outputs
  __shared__ float s_ml[8];                  // per warp: running max (log2 domain), running sum
  __shared__ __align__(16) __nv_bfloat16 s_knew[kD];
  pdl_trigger();
  pdl_wait();
  const int head = blockIdx.x, seq = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pos = seq_lens[seq];
  const int n_keys = pos + 1;
  const int n_cached = kFused ? pos : n_keys;                 // rows that come from the cache (the fused kernel makes row `pos` itself)
  const int H = n_heads * kD;
  const int32_t* pt = page_table + static_cast<long long>(seq) * max_pages;
  const __nv_bfloat16* row = qkv + static_cast<long long>(seq) * 3 * H + head * kD;
  const int n_tiles = (n_keys + kKT - 1) / kKT;

  // K / V rows of tile `t` into buffer `b`: thread -> chunk (tid & 15) of rows (tid >> 4) + 8 i of both pages of the tile
  const int r0 = tid >> 4, c = tid & 15;
  const uint32_t sw = static_cast<uint32_t>((c ^ r0) << 4);
  auto load_tile = [&](int t, int b) {
    const uint32_t kdst = smem_u32(sK) + b * 16384 + r0 * 256 + sw;
    const uint32_t vdst = smem_u32(sV) + b * 16384 + r0 * 256 + sw;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int key_base = (t * 2 + half) * kPS;
      const int page = key_base < n_keys ? __ldg(pt + t * 2 + half) : 0;
      const long long chunk0 = ((static_cast<long long>(page) * n_heads + head) * kPS + r0) * kD + c * 8;
      const __nv_bfloat16* ksrc = k_pages + chunk0;
      const __nv_bfloat16* vsrc = v_pages + chunk0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int key = key_base + r0 + 8 * i;
        const uint32_t off = static_cast<uint32_t>((half * 32 + 8 * i) * 256);
        if (kFused && key == pos) {
          // this step's own row: k' from shared memory (phase 0), v from the qkv row - never through the cache
          const uint4 kk = reinterpret_cast<const uint4*>(s_knew)[c];
          const uint4 vv = reinterpret_cast<const uint4*>(row + 2 * H)[c];
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(kdst + off), "r"(kk.x), "r"(kk.y), "r"(kk.z), "r"(kk.w) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(vdst + off), "r"(vv.x), "r"(vv.y), "r"(vv.z), "r"(vv.w) : "memory");
        } else {
          const int sz = key < n_cached ? 16 : 0;             // src-size 0 -> zero fill (rows past the context must be finite)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(kdst + off), "l"(ksrc + i * 8 * kD), "r"(sz) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vdst + off), "l"(vsrc + i * 8 * kD), "r"(sz) : "memory");
        }
      }
    }
  };

  // a first tile that does not hold this step's own row can be on its way while phase 0 runs
  const bool early = !kFused || pos >= kKT;
  if (early) {
    load_tile(0, 0);
    cp_async_commit();
  }
  // ---- phase 0: query tile (row 0 real), RoPE of q / k, KV append
  {
    uint4* qz = reinterpret_cast<uint4*>(sQ);
    qz[tid] = make_uint4(0u, 0u, 0u, 0u);
    qz[tid + 128] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  __nv_bfloat16* q0 = reinterpret_cast<__nv_bfloat16*>(sQ);   // row 0 is stored unswizzled (row & 7 == 0)
  if (kFused) {
    const int page = pt[pos / kPS];
    const long long slot = ((static_cast<long long>(page) * n_heads + head) * kPS + pos % kPS) * kD;
    if (tid < kD / 2) {
      const float inv_freq = 1.0f / powf(theta, static_cast<float>(2 * tid) / static_cast<float>(kD));
      float sn, cs;
      sincosf(static_cast<float>(pos) * inv_freq, &sn, &cs);
      const float qa = __bfloat162float(row[tid]), qb = __bfloat162float(row[tid + 64]);
      const float ka = __bfloat162float(row[H + tid]), kb = __bfloat162float(row[H + tid + 64]);
      q0[tid] = __float2bfloat16(qa * cs - qb * sn);
      q0[tid + 64] = __float2bfloat16(qb * cs + qa * sn);
      const __nv_bfloat16 k0 = __float2bfloat16(ka * cs - kb * sn), k1 = __float2bfloat16(kb * cs + ka * sn);
      s_knew[tid] = k0;
      s_knew[tid + 64] = k1;
      k_pages[slot + tid] = k0;
      k_pages[slot + tid + 64] = k1;
    } else if (tid < kD / 2 + 16) {
      reinterpret_cast<uint4*>(v_pages + slot)[tid - kD / 2] = reinterpret_cast<const uint4*>(row + 2 * H)[tid - kD / 2];
    }
  } else {
    q0[tid] = row[tid];
  }
  __syncthreads();
  if (!early) {
    load_tile(0, 0);
    cp_async_commit();
  }
  uint32_t qf[8][4];
  {
    const int r = (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) ldmatrix_x4(qf[ks], smem_u32(sQ) + tile_off(r, ks * 2 + (lane >> 4)));
  }
  float o[16][2];                                  // row 0 of this warp's partial output lives in lanes 0 - 3
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  const int t4 = lane & 3;
  const bool row0 = lane < 4;

  for (int j = 0; j < n_tiles; ++j) {
    const int buf = j & 1;
    if (j + 1 < n_tiles) {
      load_tile(j + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t kaddr = smem_u32(sK) + buf * 16384;
    const uint32_t vaddr = smem_u32(sV) + buf * 16384;
    const int key0 = j * kKT + warp * 16;
    if (key0 < n_keys) {                           // warp-uniform: a warp whose 16 keys lie past the context skips the tile
      float s[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        uint32_t b[4];
        ldmatrix_x4(b, kaddr + tile_off(warp * 16 + (lane & 7) + (lane >> 4) * 8, ks * 2 + ((lane >> 3) & 1)));
        mma_bf16_16816(s[0], qf[ks], b[0], b[1]);
        mma_bf16_16816(s[1], qf[ks], b[2], b[3]);
      }
      // row 0 of the score tile: lanes 0 - 3 hold keys key0 + nt * 8 + t4 * 2 + {0, 1} in s[nt][0 .. 1]
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = key0 + nt * 8 + t4 * 2 + e;
          if (key >= n_keys) s[nt][e] = -INFINITY;
          mx = fmaxf(mx, s[nt][e]);
        }
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = __shfl_sync(0xffffffffu, mx, 0);        // the maximum of row 0 (key0 < n_keys: at least one key is valid)
      const float m_new = fmaxf(m_run, mx * scale_log2);
      const float corr = ex2_approx(m_run - m_new);             // first tile: 2^-inf = 0
      m_run = m_new;
      float p[2][2];
      float rs = 0.f;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          p[nt][e] = ex2_approx(fmaf(s[nt][e], scale_log2, -m_new));
          rs += p[nt][e];
        }
      }
      l_run = l_run * corr + rs;
#pragma unroll
      for (int i = 0; i < 16; ++i) { o[i][0] *= corr; o[i][1] *= corr; }
      uint32_t pa[4];
      pa[0] = row0 ? pack_bf16x2(p[0][0], p[0][1]) : 0u;        // row g = 0 only; the other 15 rows of P are zero
      pa[1] = 0u;
      pa[2] = row0 ? pack_bf16x2(p[1][0], p[1][1]) : 0u;
      pa[3] = 0u;
#pragma unroll
      for (int dp = 0; dp < 8; ++dp) {
        uint32_t b[4];
        ldmatrix_x4_trans(b, vaddr + tile_off(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4)));
        // accumulator halves of rows 8 - 15 stay zero (their P rows are zero): two throw-away registers per mma
        float acc0[4] = {o[2 * dp][0], o[2 * dp][1], 0.f, 0.f};
        float acc1[4] = {o[2 * dp + 1][0], o[2 * dp + 1][1], 0.f, 0.f};
        mma_bf16_16816(acc0, pa, b[0], b[1]);
        mma_bf16_16816(acc1, pa, b[2], b[3]);
        o[2 * dp][0] = acc0[0]; o[2 * dp][1] = acc0[1];
        o[2 * dp + 1][0] = acc1[0]; o[2 * dp + 1][1] = acc1[1];
      }
    }
    __syncthreads();                               // the buffer is refilled two tiles later
  }
  // ---- merge the four warps' (max, sum, output) of row 0
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
  if (lane == 0) { s_ml[warp] = m_run; s_ml[4 + warp] = l_run; }
  if (row0) {
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      s_o[warp * kD + nt * 8 + t4 * 2] = o[nt][0];
      s_o[warp * kD + nt * 8 + t4 * 2 + 1] = o[nt][1];
    }
  }
  __syncthreads();
  {
    const float m = fmaxf(fmaxf(s_ml[0], s_ml[1]), fmaxf(s_ml[2], s_ml[3]));
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float f = ex2_approx(s_ml[w] - m);     // a warp that saw no key: 2^-inf = 0
      num += f * s_o[w * kD + tid];
      den += f * s_ml[4 + w];
    }
    out[static_cast<long long>(seq) * H + head * kD + tid] = __float2bfloat16(num / den);
  }
}

// fused != 0: qkv holds the un-rotated q, k of this step; the kernel applies RoPE at position seq_lens[i] and appends
// k', v to the cache itself (the decode step of the engine).  fused == 0: qkv is post-RoPE and the cache already
// holds this step's token (after rvl_rope_kv).
void launch_attn_decode(const void* qkv, void* out, const int32_t* seq_lens, int n_seq, const int32_t* page_table,
                        int max_pages, void* k_pages, void* v_pages, int n_heads, int page_size, int fused,
                        float theta, int max_kv_len, cudaStream_t st) {
  if (n_seq <= 0) return;
  dim3 grid(n_heads, n_seq);
  const bool pdl_dec = pdl_few_rows(n_seq, 8);
  const float scale = 1.0f / sqrtf(static_cast<float>(kD));
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  __nv_bfloat16* kp = reinterpret_cast<__nv_bfloat16*>(k_pages);
  __nv_bfloat16* vp = reinterpret_cast<__nv_bfloat16*>(v_pages);
  // Few (sequence, head) rows (stage 2 runs one query at a time): latency matters, the staged kernel wins (B = 1: 3.25 vs
  // 3.50 ms per 7B decode step); many rows: 8 register-staged CTAs per SM overlap better than 2 staged ones (B = 180: 9.4
  // vs 9.9 ms).  RVL_ATTN_DECODE = "regs" / "staged" forces one of them.
  const int forced = tuning().attn_decode;
  // Page size 32 (every shipped configuration): the tensor-core tile kernel 'm' at every batch size - measured per 7B decode
  // step inside CUDA graphs (tools/decode_ab.py, mma / regs / staged): B = 1: 2.87 / 3.04 / 2.93 ms, 8: 2.93 / 3.19 / 3.17,
  // 23: 3.22 / 3.42 / 3.83, 56: 3.85 / 4.05 / 3.89, 180: 6.51 / 6.76 / 11.5.  Other page sizes: staged for few rows, registers for
  // many.  RVL_ATTN_DECODE = regs / staged / mma forces one.
  char mode = forced ? static_cast<char>(forced) : (page_size == 32 ? 'm' : (n_seq * n_heads >= 1024 ? 'r' : 's'));
  if (mode == 'm' && page_size != 32) mode = 'r';
  if (mode == 'm') {
    constexpr int smem_m = 65536 + 4096 + 4 * kD * 4;
    static bool attr_m = false;
    if (!attr_m) {
      cudaFuncSetAttribute(attn_decode_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_m);
      cudaFuncSetAttribute(attn_decode_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_m);
      attr_m = true;
    }
    const float scale_log2 = 1.4426950408889634f * scale;
    if (fused)
      launch_pdl(pdl_dec, attn_decode_mma_kernel<true>, grid, dim3(128), smem_m, st, q, o, seq_lens, page_table, max_pages, kp, vp, n_heads, scale_log2, theta);
    else
      launch_pdl(pdl_dec, attn_decode_mma_kernel<false>, grid, dim3(128), smem_m, st, q, o, seq_lens, page_table, max_pages, kp, vp, n_heads, scale_log2, theta);
    return;
  }
  if (mode == 's') {
    // staged kernel: K and V rows of the whole context (or of a 432-key chunk) in shared memory.  Up to 220 keys two
    // CTAs share an SM (one computes while the other's copies are in flight).
    int keys = max_pages * page_size;
    if (max_kv_len > 0 && max_kv_len < keys) keys = max_kv_len;
    int cap = (keys + 7) / 8 * 8;
    if (cap > 432) cap = 432;
    const int smem = cap * (2 * kD * 2 + 4);
    static int attr_smem = 0;
    if (smem > attr_smem) {
      cudaFuncSetAttribute(attn_decode_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 432 * (2 * kD * 2 + 4));
      cudaFuncSetAttribute(attn_decode_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 432 * (2 * kD * 2 + 4));
      attr_smem = 432 * (2 * kD * 2 + 4);
    }
    if (fused)
      launch_pdl(pdl_dec, attn_decode_staged_kernel<true>, grid, dim3(128), smem, st, q, o, seq_lens, page_table, max_pages, kp, vp, n_heads,
               page_size, scale, theta, cap);
    else
      launch_pdl(pdl_dec, attn_decode_staged_kernel<false>, grid, dim3(128), smem, st, q, o, seq_lens, page_table, max_pages, kp, vp, n_heads,
               page_size, scale, theta, cap);
    return;
  }
  const int smem = max_pages * page_size * static_cast<int>(sizeof(float));
  static int attr_smem_r = 0;
  if (smem > 40000 && smem > attr_smem_r) {
    cudaFuncSetAttribute(attn_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(attn_decode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(attn_decode_kernel<true, 4, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(attn_decode_kernel<false, 4, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_smem_r = smem;
  }
  if (page_size == 32 && tuning().attn_ps32 != 0) {          // RVL_ATTN_PS32=0 (diagnostic): runtime page size arithmetic
    if (fused)
      launch_pdl(pdl_dec, attn_decode_kernel<true, 4, 32>, grid, dim3(128), smem, st, q, o, seq_lens, page_table, max_pages, kp, vp, n_heads, page_size,
               scale, theta);
    else
      launch_pdl(pdl_dec, attn_decode_kernel<false, 4, 32>, grid, dim3(128), smem, st, q, o, seq_lens, page_table, max_pages, kp, vp, n_heads, page_size,
               scale, theta);
    return;
  }
  if (fused)
    launch_pdl(pdl_dec, attn_decode_kernel<true>, grid, dim3(128), smem, st, q, o, seq_lens, page_table, max_pages, kp, vp, n_heads, page_size,
             scale, theta);
  else
    launch_pdl(pdl_dec, attn_decode_kernel<false>, grid, dim3(128), smem, st, q, o, seq_lens, page_table, max_pages, kp, vp, n_heads, page_size,
             scale, theta);
}

}  // namespace rvl
