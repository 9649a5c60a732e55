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

// (the head_dim-96 attention of the adapter lives in attention.cu: mha96_mma_kernel)

}  // namespace rvl

using namespace rvl;

extern "C" {

int rvl_layernorm(rvl_handle* h, const float* x, const void* w, const void* b, float* y_f32, void* y_bf16,
                  const float* pos, void* y_pos_bf16, int64_t rows, int32_t dim, int32_t period, float eps,
                  rvl_stream stream) {
  if (!x || rows <= 0) return report_error(h, RVL_ERR_INVALID, "rvl_layernorm: null argument");
  if (dim % 4 || dim > 1024 || (w != nullptr) != (b != nullptr) || (y_pos_bf16 && (!pos || period <= 0)))
    return report_error(h, RVL_ERR_INVALID, "rvl_layernorm: dim must be a multiple of 4 and <= 1024, weight and bias come together, y_pos needs pos and period");
  layernorm_kernel<<<static_cast<unsigned>((rows + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<const __nv_bfloat16*>(b), y_f32,
      reinterpret_cast<__nv_bfloat16*>(y_bf16), pos, reinterpret_cast<__nv_bfloat16*>(y_pos_bf16), rows, dim,
      period > 0 ? period : 1, eps);
  return cudaGetLastError() == cudaSuccess ? RVL_OK : RVL_ERR_CUDA;
}

int rvl_mha96(rvl_handle* h, const void* q, int64_t q_stride, const void* k, int64_t k_stride, const void* v,
              int64_t v_stride, void* out, int64_t out_stride, int32_t n_seq, int32_t n_heads, int32_t Tq, int32_t Tk,
              const int32_t* kv_seq_idx, const float* key_mask, rvl_stream stream) {
  if (!q || !k || !v || !out || n_seq <= 0 || Tq <= 0 || Tk <= 0) return report_error(h, RVL_ERR_INVALID, "rvl_mha96: null argument");
  if (q_stride % 8 || k_stride % 8 || v_stride % 8 || out_stride % 2)
    return report_error(h, RVL_ERR_INVALID, "rvl_mha96: q / k / v row strides must be multiples of 8 elements, the out stride of 2");
  return launch_mha96(q, q_stride, k, k_stride, v, v_stride, out, out_stride, n_seq, n_heads, Tq, Tk, kv_seq_idx, key_mask,
                      static_cast<cudaStream_t>(stream));
}

}  // extern "C"
