// Kernels of the stage-2 adapter `ClipEncoder` (d_model 768, 8 heads x 96, ffn 2048) that are not GEMMs:
// LayerNorm (+ optional positional add / casts) and a small multi-head attention with key padding.
//
// Replaces, in revisionllm/model/adapter/transformer.py of the reference:
//   nn.LayerNorm (norm1/norm2, :199-200,:255-256), `with_pos_embed` (:207-208), and the
//   nn.MultiheadAttention calls of T2V_TransformerEncoderLayer.forward_post (:271-305, queries = frames,
//   keys/values = text tokens, key_padding_mask) and TransformerEncoderLayer.forward_post (:210-223,
//   251 tokens attending to each other).  The linear layers around them run on the tcgen05 GEMM.
//
// The attention here is tiny (19 GFLOP per encoder layer for 100 segments x 251 tokens, 0.4 GFLOP for the
// text cross-attention) and memory/latency-bound, so it is a SIMT kernel: K and V of one (sequence, head)
// staged in shared memory, two threads per query row (48 dims each), online softmax in fp32.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rvl_internal.h"
#include "rvl_ptx.cuh"

namespace rvl {

// ------------------------------------------------------------------------------------ LayerNorm (+pos, casts)
// One warp per row.  y = (x - mean) * rsqrt(var + eps) * w + b (or y = x when w == nullptr).
// Outputs (each optional): y_f32 (may alias x), y_bf16, y_pos_bf16 = bf16(y + pos[row % period]).
__global__ void __launch_bounds__(128) layernorm_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                         const __nv_bfloat16* __restrict__ b, float* y_f32,
                                                         __nv_bfloat16* __restrict__ y_bf16, const float* __restrict__ pos,
                                                         __nv_bfloat16* __restrict__ y_pos_bf16, long long rows, int dim,
                                                         int period, float eps) {
  const long long row = blockIdx.x * 4LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * dim;
  constexpr int kMax = 8;  // float4 per lane -> dim <= 1024
  float4 v[kMax];
  const int nvec = dim >> 2;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      v[i] = reinterpret_cast<const float4*>(xr)[idx];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  float mean = 0.f, inv = 1.f;
  if (w != nullptr) {
    mean = warp_sum(s) / static_cast<float>(dim);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kMax; ++i) {
      const int idx = lane + i * 32;
      if (idx < nvec) {
        const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
        ss += a * a + c * c + d * d + e * e;
      }
    }
    inv = rsqrtf(warp_sum(ss) / static_cast<float>(dim) + eps);  // biased variance, like nn.LayerNorm
  }
  const float* pr = pos ? pos + static_cast<long long>(row % period) * dim : nullptr;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      float4 y = v[i];
      if (w != nullptr) {
        const uint2 wv = __ldg(reinterpret_cast<const uint2*>(w) + idx);
        const uint2 bv = __ldg(reinterpret_cast<const uint2*>(b) + idx);
        y.x = (y.x - mean) * inv * bf16_lo(wv.x) + bf16_lo(bv.x);
        y.y = (y.y - mean) * inv * bf16_hi(wv.x) + bf16_hi(bv.x);
        y.z = (y.z - mean) * inv * bf16_lo(wv.y) + bf16_lo(bv.y);
        y.w = (y.w - mean) * inv * bf16_hi(wv.y) + bf16_hi(bv.y);
      }
      if (y_f32) reinterpret_cast<float4*>(y_f32 + row * dim)[idx] = y;
      if (y_bf16) reinterpret_cast<uint2*>(y_bf16 + row * dim)[idx] = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
      if (y_pos_bf16) {
        const float4 p = reinterpret_cast<const float4*>(pr)[idx];
        reinterpret_cast<uint2*>(y_pos_bf16 + row * dim)[idx] =
            make_uint2(pack_bf16x2(y.x + p.x, y.y + p.y), pack_bf16x2(y.z + p.z, y.w + p.w));
      }
    }
  }
}

// ------------------------------------------------------------------------------------ small MHA, head_dim 96
// grid (n_heads, n_seq, ceil(Tq / 128)), 256 threads: thread pair (2r, 2r+1) owns query row r of the chunk,
// each thread 48 of the 96 dims.  K/V rows of (kv sequence, head) live in shared memory as bf16.
constexpr int kHD = 96;
constexpr int kHalf = 48;

__global__ void __launch_bounds__(256) mha96_kernel(const __nv_bfloat16* __restrict__ q, long long q_stride,
                                                     const __nv_bfloat16* __restrict__ k, long long k_stride,
                                                     const __nv_bfloat16* __restrict__ v, long long v_stride,
                                                     __nv_bfloat16* __restrict__ out, long long out_stride, int Tq, int Tk,
                                                     const int32_t* __restrict__ kv_seq_idx,
                                                     const float* __restrict__ key_mask, float scale) {
  extern __shared__ __align__(16) uint8_t smem_mha[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_mha);       // [Tk][96]
  __nv_bfloat16* sV = sK + static_cast<size_t>(Tk) * kHD;               // [Tk][96]
  float* sMask = reinterpret_cast<float*>(sV + static_cast<size_t>(Tk) * kHD);  // [Tk] additive (0 / -inf)
  const int head = blockIdx.x, seq = blockIdx.y;
  const int kv_seq = kv_seq_idx ? kv_seq_idx[seq] : seq;
  const int tid = threadIdx.x;
  // stage K, V (12 x 16 B per row)
  for (int i = tid; i < Tk * 12; i += 256) {
    const int r = i / 12, c = i - r * 12;
    const long long krow = static_cast<long long>(kv_seq) * Tk + r;
    reinterpret_cast<uint4*>(sK)[i] = __ldg(reinterpret_cast<const uint4*>(k + krow * k_stride + head * kHD) + c);
    reinterpret_cast<uint4*>(sV)[i] = __ldg(reinterpret_cast<const uint4*>(v + krow * v_stride + head * kHD) + c);
  }
  for (int i = tid; i < Tk; i += 256)
    sMask[i] = (key_mask == nullptr || key_mask[static_cast<long long>(kv_seq) * Tk + i] != 0.f) ? 0.f : -INFINITY;
  __syncthreads();
  const int qr = blockIdx.z * 128 + (tid >> 1);
  const int half = tid & 1;
  const bool active = qr < Tq;
  const long long qrow = static_cast<long long>(seq) * Tq + (active ? qr : 0);
  float qf[kHalf], acc[kHalf];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(q + qrow * q_stride + head * kHD + half * kHalf);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const uint4 t = __ldg(qp + c);
      qf[c * 8 + 0] = bf16_lo(t.x) * scale; qf[c * 8 + 1] = bf16_hi(t.x) * scale;
      qf[c * 8 + 2] = bf16_lo(t.y) * scale; qf[c * 8 + 3] = bf16_hi(t.y) * scale;
      qf[c * 8 + 4] = bf16_lo(t.z) * scale; qf[c * 8 + 5] = bf16_hi(t.z) * scale;
      qf[c * 8 + 6] = bf16_lo(t.w) * scale; qf[c * 8 + 7] = bf16_hi(t.w) * scale;
    }
  }
#pragma unroll
  for (int d = 0; d < kHalf; ++d) acc[d] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  for (int j0 = 0; j0 < Tk; j0 += 4) {
    float s[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = j0 + jj;
      float d = 0.f;
      if (j < Tk) {
        const uint32_t* kr = reinterpret_cast<const uint32_t*>(sK + static_cast<size_t>(j) * kHD + half * kHalf);
#pragma unroll
        for (int e = 0; e < kHalf / 2; ++e) {
          const uint32_t kk = kr[e];
          d += qf[2 * e] * bf16_lo(kk) + qf[2 * e + 1] * bf16_hi(kk);
        }
      }
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      s[jj] = (j < Tk) ? d + sMask[j] : -INFINITY;
    }
    const float cm = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
    const float m_new = fmaxf(m_run, cm);
    if (m_new == -INFINITY) continue;   // everything so far masked
    const float corr = __expf(m_run - m_new);
    l_run *= corr;
#pragma unroll
    for (int d = 0; d < kHalf; ++d) acc[d] *= corr;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = j0 + jj;
      const float p = __expf(s[jj] - m_new);   // exp(-inf) = 0 for masked / out-of-range keys
      l_run += p;
      if (j < Tk) {
        const uint32_t* vr = reinterpret_cast<const uint32_t*>(sV + static_cast<size_t>(j) * kHD + half * kHalf);
#pragma unroll
        for (int e = 0; e < kHalf / 2; ++e) {
          const uint32_t vv = vr[e];
          acc[2 * e] += p * bf16_lo(vv);
          acc[2 * e + 1] += p * bf16_hi(vv);
        }
      }
    }
    m_run = m_new;
  }
  if (active) {
    const float inv = 1.f / l_run;
    uint32_t* op = reinterpret_cast<uint32_t*>(out + qrow * out_stride + head * kHD + half * kHalf);
#pragma unroll
    for (int e = 0; e < kHalf / 2; ++e) op[e] = pack_bf16x2(acc[2 * e] * inv, acc[2 * e + 1] * inv);
  }
}

}  // namespace rvl

using namespace rvl;

extern "C" {

int rvl_layernorm(rvl_handle* h, const float* x, const void* w, const void* b, float* y_f32, void* y_bf16,
                  const float* pos, void* y_pos_bf16, int64_t rows, int32_t dim, int32_t period, float eps,
                  rvl_stream stream) {
  (void)h;
  if (!x || rows <= 0) return RVL_ERR_INVALID;
  if (dim % 4 || dim > 1024 || (w != nullptr) != (b != nullptr) || (y_pos_bf16 && (!pos || period <= 0))) return RVL_ERR_INVALID;
  layernorm_kernel<<<static_cast<unsigned>((rows + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<const __nv_bfloat16*>(b), y_f32,
      reinterpret_cast<__nv_bfloat16*>(y_bf16), pos, reinterpret_cast<__nv_bfloat16*>(y_pos_bf16), rows, dim,
      period > 0 ? period : 1, eps);
  return cudaGetLastError() == cudaSuccess ? RVL_OK : RVL_ERR_CUDA;
}

int rvl_mha96(rvl_handle* h, const void* q, int64_t q_stride, const void* k, int64_t k_stride, const void* v,
              int64_t v_stride, void* out, int64_t out_stride, int32_t n_seq, int32_t n_heads, int32_t Tq, int32_t Tk,
              const int32_t* kv_seq_idx, const float* key_mask, rvl_stream stream) {
  (void)h;
  if (!q || !k || !v || !out || n_seq <= 0 || Tq <= 0 || Tk <= 0) return RVL_ERR_INVALID;
  if (q_stride % 8 || k_stride % 8 || v_stride % 8 || out_stride % 2) return RVL_ERR_INVALID;
  const size_t smem = static_cast<size_t>(Tk) * kHD * 2 * 2 + static_cast<size_t>(Tk) * 4;
  if (smem > 200 * 1024) return RVL_ERR_INVALID;   // Tk <= ~520
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    if (cudaFuncSetAttribute(mha96_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return RVL_ERR_CUDA;
    attr = smem;
  }
  dim3 grid(n_heads, n_seq, (Tq + 127) / 128);
  mha96_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), q_stride, reinterpret_cast<const __nv_bfloat16*>(k), k_stride,
      reinterpret_cast<const __nv_bfloat16*>(v), v_stride, reinterpret_cast<__nv_bfloat16*>(out), out_stride, Tq, Tk,
      kv_seq_idx, key_mask, 1.0f / sqrtf(static_cast<float>(kHD)));
  return cudaGetLastError() == cudaSuccess ? RVL_OK : RVL_ERR_CUDA;
}

}  // extern "C"
