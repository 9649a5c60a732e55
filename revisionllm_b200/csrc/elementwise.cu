// Bandwidth-bound kernels of the decoder stack: RMSNorm, embedding gather / splice scatter,
// RoPE + KV page write, SwiGLU.  All are vectorised (16-byte accesses), coalesced along the
// feature dimension and HBM-bound; grids are one CTA per token row (>> 148 SMs x resident CTAs).
//
// Replaces (reference runs unfused torch eager kernels through transformers' Llama):
//   LlamaRMSNorm, rotate_half RoPE, DynamicCache append, SiLU*mul, nn.Embedding and the torch.cat
//   splice of revisionllm/model/vtimellm_arch.py:165-238.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rvl_internal.h"
#include "rvl_ptx.cuh"

namespace rvl {

// ------------------------------------------------------------------------------------ RMSNorm
// y(bf16) = x(fp32) * rsqrt(mean(x^2) + eps) * w(bf16).  One CTA per row, the row is held in
// registers between the reduction and the scaling pass (read once, write once).
// Algorithmic bytes per row: dim*4 (read) + dim*2 (write) (+ dim*2 weight, L2 resident).
// Optional fused split-k reduction (decode): the row is first completed with the partial sums the preceding
// weight-streaming GEMM left in `partials` ([n_partials][rows][dim] fp32) and written back as the new residual.
// kPartials compiles the split-k reduction in; the prefill variant stays at 54 registers / 28 CTAs per SM without it (with the
// partial-sum registers in the same kernel it measured 87 registers, 14 CTAs and 244 instead of 113 us per 33120 x 4096 launch)
template <int THREADS, int kMaxVec, bool kPartials>   // kMaxVec float4 per thread: dim <= THREADS * kMaxVec * 4
__global__ void __launch_bounds__(THREADS) rmsnorm_kernel(const float* x, const __nv_bfloat16* __restrict__ w,
                                                           __nv_bfloat16* __restrict__ y, int dim, float eps,
                                                           const int32_t* __restrict__ rows,
                                                           const float* __restrict__ partials, int n_partials,
                                                           long long partial_stride, float* x_out) {
  constexpr int kMaxPartials = 8;
  pdl_trigger();
  pdl_wait();
  const long long src_row = rows ? rows[blockIdx.x] : blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x + src_row * dim);
  const int nvec = dim >> 2;
  float4 c[kMaxVec];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int idx = threadIdx.x + i * THREADS;
    if (idx < nvec) {
      c[i] = xr[idx];
      if (kPartials && n_partials > 0) {
        // all partial loads are issued before the first add: one round trip instead of n_partials
        float4 q[kMaxPartials];
#pragma unroll
        for (int p = 0; p < kMaxPartials; ++p)
          if (p < n_partials) q[p] = __ldcg(reinterpret_cast<const float4*>(partials + p * partial_stride + src_row * dim) + idx);
#pragma unroll
        for (int p = 0; p < kMaxPartials; ++p)
          if (p < n_partials) { c[i].x += q[p].x; c[i].y += q[p].y; c[i].z += q[p].z; c[i].w += q[p].w; }
      }
      if (kPartials && x_out) reinterpret_cast<float4*>(x_out + src_row * dim)[idx] = c[i];
      ss += c[i].x * c[i].x + c[i].y * c[i].y + c[i].z * c[i].z + c[i].w * c[i].w;
    }
  }
  // few rows (decode): the weight vector is fetched before the reduction so that its L2 latency is not paid after the barrier
  constexpr bool kEarlyW = kMaxVec <= 2;
  const uint2* wr = reinterpret_cast<const uint2*>(w);
  uint2 wv_early[kEarlyW ? kMaxVec : 1];
  if (kEarlyW) {
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int idx = threadIdx.x + i * THREADS;
      if (idx < nvec) wv_early[i] = __ldg(wr + idx);
    }
  }
  __shared__ float red[THREADS / 32];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < THREADS / 32; ++i) tot += red[i];
  const float inv = rsqrtf(tot / static_cast<float>(dim) + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<long long>(blockIdx.x) * dim);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int idx = threadIdx.x + i * THREADS;
    if (idx < nvec) {
      const uint2 wv = kEarlyW ? wv_early[i] : __ldg(wr + idx);
      uint2 o;
      o.x = pack_bf16x2(bf16_lo(wv.x) * (c[i].x * inv), bf16_hi(wv.x) * (c[i].y * inv));
      o.y = pack_bf16x2(bf16_lo(wv.y) * (c[i].z * inv), bf16_hi(wv.y) * (c[i].w * inv));
      yr[idx] = o;
    }
  }
}

void launch_rmsnorm(const float* x, const void* w, void* y, int64_t n_rows, int dim, float eps, const int32_t* rows,
                    cudaStream_t st, const float* partials, int n_partials, int64_t partial_stride, float* x_out) {
  if (n_rows <= 0) return;
  const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(w);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  const long long ps = partial_stride;
  // few rows (decode): one float4 per thread so that a row's loads are all in flight at once (the 128-thread
  // variant took 9.7 us for 180 rows with 4 split-k partials - latency, not bandwidth)
  const bool reduce = n_partials > 0 || x_out != nullptr;
  const bool pdl_few = pdl_few_rows(n_rows, 4);
  if (n_rows <= 1024 && dim >= 2048 && reduce)
    launch_pdl(pdl_few, rmsnorm_kernel<1024, 2, true>, dim3(static_cast<unsigned>(n_rows)), dim3(1024), 0, st, x, wp, yp, dim, eps, rows, partials, n_partials, ps, x_out);
  else if (n_rows <= 1024 && dim >= 2048)
    launch_pdl(pdl_few, rmsnorm_kernel<1024, 2, false>, dim3(static_cast<unsigned>(n_rows)), dim3(1024), 0, st, x, wp, yp, dim, eps, rows, partials, n_partials, ps, x_out);
  else if (reduce)
    launch_k(rmsnorm_kernel<256, 8, true>, dim3(static_cast<unsigned>(n_rows)), dim3(256), 0, st, x, wp, yp, dim, eps, rows, partials, n_partials, ps, x_out);
  else if (dim <= 128 * 32)
    // many rows (prefill): 128 threads x 8 float4, 28 CTAs per SM - measured 113 us per 33120 x 4096 launch (82 % of DRAM
    // peak); a 512-thread variant with the same bytes per thread ran at 250 us
    launch_k(rmsnorm_kernel<128, 8, false>, dim3(static_cast<unsigned>(n_rows)), dim3(128), 0, st, x, wp, yp, dim, eps, rows, partials, n_partials, ps, x_out);
  else
    launch_k(rmsnorm_kernel<256, 8, false>, dim3(static_cast<unsigned>(n_rows)), dim3(256), 0, st, x, wp, yp, dim, eps, rows, partials, n_partials, ps, x_out);
}

// ------------------------------------------------------------------------------------ embedding / splice rows
// out[dst_rows[i]] (fp32) = table[ids[i]] (bf16).  Text rows of the splice
// (vtimellm_arch.py:194 `embed_tokens(cat(text chunks))`) and the decode-step token embedding.
__global__ void embed_rows_kernel(const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ ids,
                                  const int32_t* __restrict__ dst_rows, int dim, int vocab, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x;
  int id = ids[i];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const long long dst = dst_rows ? dst_rows[i] : i;
  const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<long long>(id) * dim);
  float4* o = reinterpret_cast<float4*>(out + dst * dim);
  for (int v = threadIdx.x; v < (dim >> 3); v += blockDim.x) {
    const uint4 s = __ldg(src + v);
    o[2 * v] = make_float4(bf16_lo(s.x), bf16_hi(s.x), bf16_lo(s.y), bf16_hi(s.y));
    o[2 * v + 1] = make_float4(bf16_lo(s.z), bf16_hi(s.z), bf16_lo(s.w), bf16_hi(s.w));
  }
}
void launch_embed_rows(const void* table, const int32_t* ids, const int32_t* dst_rows, int n, int dim, int vocab,
                       float* out, cudaStream_t st) {
  if (n <= 0) return;
  launch_k(embed_rows_kernel, dim3(n), dim3(128), 0, st, reinterpret_cast<const __nv_bfloat16*>(table), ids, dst_rows, dim, vocab, out);
}

// out[dst_rows[i]] (fp32) = src[i] (bf16): already-projected visual rows (stage-2 CLS tokens).
__global__ void scatter_rows_kernel(const __nv_bfloat16* __restrict__ src, const int32_t* __restrict__ dst_rows, int dim,
                                    float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x;
  const long long dst = dst_rows ? dst_rows[i] : i;
  const uint4* s4 = reinterpret_cast<const uint4*>(src + static_cast<long long>(i) * dim);
  float4* o = reinterpret_cast<float4*>(out + dst * dim);
  for (int v = threadIdx.x; v < (dim >> 3); v += blockDim.x) {
    const uint4 s = __ldg(s4 + v);
    o[2 * v] = make_float4(bf16_lo(s.x), bf16_hi(s.x), bf16_lo(s.y), bf16_hi(s.y));
    o[2 * v + 1] = make_float4(bf16_lo(s.z), bf16_hi(s.z), bf16_lo(s.w), bf16_hi(s.w));
  }
}
void launch_scatter_rows_bf16(const void* src, const int32_t* dst_rows, int n, int dim, float* out, cudaStream_t st) {
  if (n <= 0) return;
  launch_k(scatter_rows_kernel, dim3(n), dim3(128), 0, st, reinterpret_cast<const __nv_bfloat16*>(src), dst_rows, dim, out);
}

// ------------------------------------------------------------------------------------ window gather (feature loader)
// out[i] (bf16) = src[idx[i]] (fp32): the per-window frame sampling `features[np.linspace(start, end, num_frames)]` of
// revisionllm/eval/eval_nlq_negative.py:224-235 fused with the fp32 -> bf16 cast the reference does later with
// `.to(torch.bfloat16)`; the movie's features cross PCIe once as stored (fp32) and the overlapping windows are built here.
// HBM-bound: 4 B read + 2 B written per element; one CTA per output row.
__global__ void gather_rows_f32_bf16_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int n_src, int dim,
                                            __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x;
  int r = idx[i];
  r = r < 0 ? 0 : (r >= n_src ? n_src - 1 : r);
  const float4* s4 = reinterpret_cast<const float4*>(src + static_cast<long long>(r) * dim);
  uint2* o = reinterpret_cast<uint2*>(out + static_cast<long long>(i) * dim);
  for (int v = threadIdx.x; v < (dim >> 2); v += blockDim.x) {
    const float4 x = __ldg(s4 + v);
    o[v] = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
  }
}
void launch_gather_rows_f32_bf16(const float* src, const int32_t* idx, int n_rows, int n_src, int dim, void* out, cudaStream_t st) {
  if (n_rows <= 0) return;
  launch_k(gather_rows_f32_bf16_kernel, dim3(n_rows), dim3(192), 0, st, src, idx, n_src, dim, reinterpret_cast<__nv_bfloat16*>(out));
}

// ------------------------------------------------------------------------------------ token -> sequence map
__global__ void token_seq_kernel(const int32_t* __restrict__ cu, int n_seq, int32_t* __restrict__ tok_seq,
                                 int32_t* __restrict__ last_rows, long long total) {
  pdl_trigger();
  pdl_wait();
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t < n_seq && last_rows) last_rows[t] = cu[t + 1] - 1;
  if (t >= total) return;
  int lo = 0, hi = n_seq;  // find s with cu[s] <= t < cu[s+1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cu[mid] <= t) lo = mid; else hi = mid;
  }
  tok_seq[t] = lo;
}
void launch_token_seq(const int32_t* cu_seqlens, int n_seq, int32_t* tok_seq, int32_t* last_rows, int64_t total,
                      cudaStream_t st) {
  const long long n = total > n_seq ? total : n_seq;
  if (n <= 0) return;
  launch_k(token_seq_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, cu_seqlens, n_seq, tok_seq, last_rows,
           static_cast<long long>(total));
}

// Last decoder layer of a generation prefill: only the last position of each sequence goes on, so its attention row
// (bf16) and its residual row (fp32) are copied out of the packed stream and the rest of the layer runs on n_seq rows.
__global__ void __launch_bounds__(256) gather_last_rows_kernel(const uint4* __restrict__ attn, const float4* __restrict__ hidden,
                                                               const int32_t* __restrict__ rows, int dim, uint4* __restrict__ attn_out,
                                                               float4* __restrict__ hidden_out) {
  pdl_trigger();
  pdl_wait();
  const long long src = rows[blockIdx.x];
  const int na = dim / 8, nh = dim / 4;
  for (int i = threadIdx.x; i < na; i += blockDim.x) attn_out[static_cast<long long>(blockIdx.x) * na + i] = attn[src * na + i];
  for (int i = threadIdx.x; i < nh; i += blockDim.x) hidden_out[static_cast<long long>(blockIdx.x) * nh + i] = hidden[src * nh + i];
}
void launch_gather_last_rows(const void* attn, const float* hidden, const int32_t* rows, int n_rows, int dim, void* attn_out,
                             float* hidden_out, cudaStream_t st) {
  if (n_rows <= 0) return;
  launch_k(gather_last_rows_kernel, dim3(n_rows), dim3(256), 0, st, reinterpret_cast<const uint4*>(attn),
           reinterpret_cast<const float4*>(hidden), rows, dim, reinterpret_cast<uint4*>(attn_out), reinterpret_cast<float4*>(hidden_out));
}

// ------------------------------------------------------------------------------------ RoPE + KV write
// In-place rotate-half RoPE on q and k of qkv [T, 3*H] (bf16) and write of the rotated k and of v into
// the paged cache [page][head][slot][128].  One CTA per token; cos/sin computed once per token in fp32
// (sincosf with full range reduction: positions reach 4096 rad) and shared by all heads.
// Algorithmic bytes per token: read 3H*2, write 2H*2 (q,k in place) + 2H*2 (cache).
__global__ void __launch_bounds__(128) rope_kv_kernel(__nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ positions,
                                                       const int32_t* __restrict__ tok_seq,
                                                       const int32_t* __restrict__ cu_seqlens,
                                                       const int32_t* __restrict__ page_table, int max_pages,
                                                       __nv_bfloat16* __restrict__ k_pages, __nv_bfloat16* __restrict__ v_pages,
                                                       int n_heads, int page_size, float theta,
                                                       const int32_t* __restrict__ seq_pos0) {
  constexpr int D = 128;
  __shared__ float s_cos[D / 2], s_sin[D / 2];
  pdl_trigger();
  pdl_wait();
  const long long tok = blockIdx.x;
  const int seq = tok_seq ? tok_seq[tok] : static_cast<int>(tok);
  // seq_pos0: the sequence's own rows start at this position (the positions before it belong to a shared context)
  const int pos = positions ? positions[tok] : static_cast<int>(tok - cu_seqlens[seq]) + (seq_pos0 ? seq_pos0[seq] : 0);
  if (threadIdx.x < D / 2) {
    // inv_freq_i = theta^(-2i/D)  (LlamaRotaryEmbedding), angle in fp32 like the reference
    const float inv_freq = 1.0f / powf(theta, static_cast<float>(2 * threadIdx.x) / static_cast<float>(D));
    float s, c;
    sincosf(static_cast<float>(pos) * inv_freq, &s, &c);
    s_cos[threadIdx.x] = c;
    s_sin[threadIdx.x] = s;
  }
  __syncthreads();
  const int H = n_heads * D;
  __nv_bfloat16* row = qkv + tok * 3LL * H;
  const int page = page_table[static_cast<long long>(seq) * max_pages + pos / page_size];
  const int slot = pos % page_size;
  // q and k: (qk, head, j) with j in 0..7 covering dims [8j, 8j+8) and partner [64+8j, ...)
  for (int it = threadIdx.x; it < 2 * n_heads * 8; it += blockDim.x) {
    const int j = it & 7;
    const int head = (it >> 3) % n_heads;
    const int qk = it / (8 * n_heads);
    __nv_bfloat16* base = row + qk * H + head * D;
    uint4 lo = *reinterpret_cast<const uint4*>(base + 8 * j);
    uint4 hi = *reinterpret_cast<const uint4*>(base + 64 + 8 * j);
    uint32_t* l = reinterpret_cast<uint32_t*>(&lo);
    uint32_t* h = reinterpret_cast<uint32_t*>(&hi);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d0 = 8 * j + 2 * e;
      const float a0 = bf16_lo(l[e]), a1 = bf16_hi(l[e]);
      const float b0 = bf16_lo(h[e]), b1 = bf16_hi(h[e]);
      const float c0 = s_cos[d0], s0 = s_sin[d0], c1 = s_cos[d0 + 1], s1 = s_sin[d0 + 1];
      l[e] = pack_bf16x2(a0 * c0 - b0 * s0, a1 * c1 - b1 * s1);
      h[e] = pack_bf16x2(b0 * c0 + a0 * s0, b1 * c1 + a1 * s1);
    }
    *reinterpret_cast<uint4*>(base + 8 * j) = lo;
    *reinterpret_cast<uint4*>(base + 64 + 8 * j) = hi;
    if (qk == 1) {
      __nv_bfloat16* kc = k_pages + ((static_cast<long long>(page) * n_heads + head) * page_size + slot) * D;
      *reinterpret_cast<uint4*>(kc + 8 * j) = lo;
      *reinterpret_cast<uint4*>(kc + 64 + 8 * j) = hi;
    }
  }
  // v: straight copy into the cache
  for (int it = threadIdx.x; it < n_heads * 16; it += blockDim.x) {
    const int j = it & 15;
    const int head = it >> 4;
    const uint4 val = *reinterpret_cast<const uint4*>(row + 2 * H + head * D + 8 * j);
    __nv_bfloat16* vc = v_pages + ((static_cast<long long>(page) * n_heads + head) * page_size + slot) * D;
    *reinterpret_cast<uint4*>(vc + 8 * j) = val;
  }
}
void launch_rope_kv(void* qkv, int64_t n_tokens, const int32_t* positions, const int32_t* tok_seq,
                    const int32_t* cu_seqlens, const int32_t* page_table, int max_pages, void* k_pages, void* v_pages,
                    int n_heads, int page_size, float theta, cudaStream_t st, const int32_t* seq_pos0) {
  if (n_tokens <= 0) return;
  launch_k(rope_kv_kernel, dim3(static_cast<unsigned>(n_tokens)), dim3(128), 0, st, reinterpret_cast<__nv_bfloat16*>(qkv), positions,
           tok_seq, cu_seqlens, page_table, max_pages, reinterpret_cast<__nv_bfloat16*>(k_pages),
           reinterpret_cast<__nv_bfloat16*>(v_pages), n_heads, page_size, theta, seq_pos0);
}

// ------------------------------------------------------------------------------------ SwiGLU
// act[t, i] = silu(gu[t, i]) * gu[t, I + i], bf16 in/out, fp32 math, 8 elements (16 B) per thread.
// Algorithmic bytes per token: 2I*2 read + I*2 write.
__global__ void swiglu_kernel(const __nv_bfloat16* __restrict__ gu, __nv_bfloat16* __restrict__ act, long long n_vec,
                              int inter) {
  pdl_trigger();
  pdl_wait();
  const int vec_per_row = inter >> 3;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < n_vec;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = v / vec_per_row;
    const int i = static_cast<int>(v - t * vec_per_row) << 3;
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(gu + t * 2LL * inter + i));
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(gu + t * 2LL * inter + inter + i));
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
    const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float g0 = bf16_lo(gw[e]), g1 = bf16_hi(gw[e]);
      const float r0 = __fdividef(g0, 1.f + __expf(-g0)) * bf16_lo(uw[e]);
      const float r1 = __fdividef(g1, 1.f + __expf(-g1)) * bf16_hi(uw[e]);
      o[e] = pack_bf16x2(r0, r1);
    }
    *reinterpret_cast<uint4*>(act + t * static_cast<long long>(inter) + i) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
void launch_swiglu(const void* gu, void* act, int64_t n_tokens, int inter, cudaStream_t st) {
  const long long n_vec = n_tokens * (inter >> 3);
  if (n_vec <= 0) return;
  long long blocks = (n_vec + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;  // grid-stride: a multiple of the SM count
  launch_k(swiglu_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(gu),
           reinterpret_cast<__nv_bfloat16*>(act), n_vec, inter);
}

}  // namespace rvl
