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

// frames bf16 -> fp32 residual stream
__global__ void clip_cast_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long n4) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n4) return;
  const uint2 v = reinterpret_cast<const uint2*>(x)[i];
  reinterpret_cast<float4*>(y)[i] = make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
}
// x1[v, 0] = global token, x1[v, 1 + t] = x[v, t]   (transformer.py:127-133: `torch.cat([global_rep_token, src])`)
__global__ void clip_prepend_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ token, float* __restrict__ x1, int T,
                                    int dim4, long long rows1) {
  const long long row = blockIdx.x;                       // row of x1
  if (row >= rows1) return;
  const int t1 = static_cast<int>(row % (T + 1));
  const long long v = row / (T + 1);
  float4* dst = reinterpret_cast<float4*>(x1) + row * dim4;
  if (t1 == 0) {
    for (int i = threadIdx.x; i < dim4; i += blockDim.x) {
      const uint2 w = reinterpret_cast<const uint2*>(token)[i];
      dst[i] = make_float4(bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y));
    }
  } else {
    const float4* src = reinterpret_cast<const float4*>(x) + (v * T + t1 - 1) * dim4;
    for (int i = threadIdx.x; i < dim4; i += blockDim.x) dst[i] = src[i];
  }
}
// CLS rows: out[v] = x1_bf[v * (T + 1)]
__global__ void clip_cls_rows_kernel(const __nv_bfloat16* __restrict__ x1, __nv_bfloat16* __restrict__ out, int T1, int dim8) {
  const uint4* src = reinterpret_cast<const uint4*>(x1) + static_cast<long long>(blockIdx.x) * T1 * dim8;
  uint4* dst = reinterpret_cast<uint4*>(out) + static_cast<long long>(blockIdx.x) * dim8;
  for (int i = threadIdx.x; i < dim8; i += blockDim.x) dst[i] = src[i];
}

namespace {
constexpr int kClipD = 768, kClipHeads = 8, kClipFfn = 2048;
inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }
struct ClipWs {
  size_t x, x1, stage, cls, total;              // `stage`: buffers of the cross-attention part, re-used by the self-attention part
  size_t a_xbf, a_xpbf, a_q, a_att, a_h, a_kv;  // offsets inside `stage` (cross-attention part)
  size_t b_xbf, b_xpbf, b_qk, b_v, b_att, b_h;  // offsets inside `stage` (self-attention part)
};
ClipWs clip_ws(long long V, long long T, long long Q, long long Lq) {
  ClipWs w{};
  const size_t rows = static_cast<size_t>(V * T), rows1 = static_cast<size_t>(V * (T + 1)), D = kClipD;
  size_t off = 0;
  w.x = off; off += up256(rows * D * 4);
  w.x1 = off; off += up256(rows1 * D * 4);
  w.stage = off;
  size_t a = 0;
  w.a_xbf = a; a += up256(rows * D * 2);
  w.a_xpbf = a; a += up256(rows * D * 2);
  w.a_q = a; a += up256(rows * D * 2);
  w.a_att = a; a += up256(rows * D * 2);
  w.a_h = a; a += up256(rows * kClipFfn * 2);
  w.a_kv = a; a += up256(static_cast<size_t>(Q * Lq) * 2 * D * 2);
  size_t b = 0;
  w.b_xbf = b; b += up256(rows1 * D * 2);
  w.b_xpbf = b; b += up256(rows1 * D * 2);
  w.b_qk = b; b += up256(rows1 * 2 * D * 2);
  w.b_v = b; b += up256(rows1 * D * 2);
  w.b_att = b; b += up256(rows1 * D * 2);
  w.b_h = b; b += up256(rows1 * kClipFfn * 2);
  off += a > b ? a : b;
  w.cls = off; off += up256(static_cast<size_t>(V) * D * 2);
  w.total = off;
  return w;
}
}  // namespace

}  // namespace rvl

using namespace rvl;

extern "C" {

size_t rvl_clip_encoder_workspace_bytes(int32_t n_seg, int32_t n_frames, int32_t n_text, int32_t text_len) {
  if (n_seg <= 0 || n_frames <= 0 || n_text <= 0 || text_len <= 0) return 0;
  return clip_ws(n_seg, n_frames, n_text, text_len).total;
}

int rvl_clip_encoder(rvl_handle* h, const rvl_clip_weights* w, const void* frames, const void* text, const float* text_mask,
                     const int32_t* seg_text_idx, int32_t V, int32_t T, int32_t Q, int32_t Lq, void* ws, size_t ws_bytes, void* out,
                     rvl_stream stream) {
  if (!h || !w || !frames || !text || !text_mask || !ws || !out) return report_error(h, RVL_ERR_INVALID, "rvl_clip_encoder: null argument");
  if (V <= 0 || T <= 0 || Q <= 0 || Lq <= 0 || (!seg_text_idx && Q != V))
    return report_error(h, RVL_ERR_INVALID, "rvl_clip_encoder: bad sizes (seg_text_idx is required when n_text != n_seg)");
  const ClipWs l = clip_ws(V, T, Q, Lq);
  if (ws_bytes < l.total || (reinterpret_cast<uintptr_t>(ws) & 255)) return report_error(h, RVL_ERR_INVALID, "rvl_clip_encoder: workspace too small or misaligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* b = static_cast<uint8_t*>(ws);
  const int D = kClipD, F = kClipFfn;
  const long long rows = static_cast<long long>(V) * T, rows1 = static_cast<long long>(V) * (T + 1);
  float* x = reinterpret_cast<float*>(b + l.x);
  float* x1 = reinterpret_cast<float*>(b + l.x1);
  uint8_t* sg = b + l.stage;
  int rc;
  auto bf = [](const void* p, long long elems) { return static_cast<const void*>(static_cast<const uint16_t*>(p) + elems); };
  auto gemm = [&](const void* A, const void* W, const void* bias, void* o, long long M, int N, int K, int mode, int flags) -> int {
    return rvl_gemm_bf16(h, A, W, bias, o, M, N, K, N, mode, flags, nullptr, 1, stream);
  };
  // ---- frames -> fp32 residual stream; q input of the first layer = x + pos
  clip_cast_kernel<<<static_cast<unsigned>((rows * D / 4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(frames), x, rows * D / 4);
  if ((rc = rvl_layernorm(h, x, nullptr, nullptr, nullptr, nullptr, w->pos, sg + l.a_xpbf, rows, D, T, 1e-5f, stream))) return rc;
  // ---- 2 x text -> video cross-attention (T2V_TransformerEncoderLayer.forward_post, transformer.py:271-305)
  for (int i = 0; i < 2; ++i) {
    const rvl_clip_layer& L = w->t2v[i];
    void* q = sg + l.a_q; void* kv = sg + l.a_kv; void* att = sg + l.a_att; void* xbf = sg + l.a_xbf; void* hb = sg + l.a_h;
    if ((rc = gemm(sg + l.a_xpbf, L.in_proj_w, L.in_proj_b, q, rows, D, D, RVL_GEMM_OUT_BF16, 0))) return rc;                     // Q = (x + pos) Wq^T + bq
    if ((rc = gemm(text, bf(L.in_proj_w, 1LL * D * D), bf(L.in_proj_b, D), kv, 1LL * Q * Lq, 2 * D, D, RVL_GEMM_OUT_BF16, 0))) return rc;   // K | V of the text, once per query
    if ((rc = rvl_mha96(h, q, D, kv, 2 * D, bf(kv, D), 2 * D, att, D, V, kClipHeads, T, Lq, seg_text_idx, Q, text_mask, stream))) return rc;
    if ((rc = gemm(att, L.out_proj_w, L.out_proj_b, x, rows, D, D, RVL_GEMM_ADD_F32, 0))) return rc;                             // src2 = x + attn
    if ((rc = rvl_layernorm(h, x, L.norm1_w, L.norm1_b, nullptr, xbf, nullptr, nullptr, rows, D, 0, 1e-5f, stream))) return rc;
    if ((rc = gemm(xbf, L.linear1_w, L.linear1_b, hb, rows, F, D, RVL_GEMM_OUT_BF16, RVL_GEMM_FLAG_RELU))) return rc;
    if ((rc = gemm(hb, L.linear2_w, L.linear2_b, x, rows, D, F, RVL_GEMM_ADD_F32, 0))) return rc;
    if ((rc = rvl_layernorm(h, x, L.norm2_w, L.norm2_b, x, nullptr, w->pos, sg + l.a_xpbf, rows, D, T, 1e-5f, stream))) return rc;
  }
  // ---- prepend the global token; 2 x post-norm self-attention over 1 + T tokens (TransformerEncoderLayer.forward_post, :210-223)
  const int T1 = T + 1;
  clip_prepend_kernel<<<static_cast<unsigned>(rows1), 192, 0, st>>>(x, reinterpret_cast<const __nv_bfloat16*>(w->global_token), x1, T, D / 4, rows1);
  void* x1bf = sg + l.b_xbf; void* x1pbf = sg + l.b_xpbf; void* qk = sg + l.b_qk; void* vb = sg + l.b_v; void* att1 = sg + l.b_att; void* h1 = sg + l.b_h;
  if ((rc = rvl_layernorm(h, x1, nullptr, nullptr, nullptr, x1bf, w->pos_global, x1pbf, rows1, D, T1, 1e-5f, stream))) return rc;
  for (int i = 0; i < 2; ++i) {
    const rvl_clip_layer& L = w->enc[i];
    if ((rc = gemm(x1pbf, L.in_proj_w, L.in_proj_b, qk, rows1, 2 * D, D, RVL_GEMM_OUT_BF16, 0))) return rc;                      // q = k = x + pos
    if ((rc = gemm(x1bf, bf(L.in_proj_w, 2LL * D * D), bf(L.in_proj_b, 2 * D), vb, rows1, D, D, RVL_GEMM_OUT_BF16, 0))) return rc;   // v = x
    if ((rc = rvl_mha96(h, qk, 2 * D, bf(qk, D), 2 * D, vb, D, att1, D, V, kClipHeads, T1, T1, nullptr, 0, nullptr, stream))) return rc;
    if ((rc = gemm(att1, L.out_proj_w, L.out_proj_b, x1, rows1, D, D, RVL_GEMM_ADD_F32, 0))) return rc;
    if ((rc = rvl_layernorm(h, x1, L.norm1_w, L.norm1_b, x1, x1bf, nullptr, nullptr, rows1, D, 0, 1e-5f, stream))) return rc;
    if ((rc = gemm(x1bf, L.linear1_w, L.linear1_b, h1, rows1, F, D, RVL_GEMM_OUT_BF16, RVL_GEMM_FLAG_RELU))) return rc;
    if ((rc = gemm(h1, L.linear2_w, L.linear2_b, x1, rows1, D, F, RVL_GEMM_ADD_F32, 0))) return rc;
    if ((rc = rvl_layernorm(h, x1, L.norm2_w, L.norm2_b, x1, x1bf, w->pos_global, x1pbf, rows1, D, T1, 1e-5f, stream))) return rc;
  }
  // ---- CLS row -> Linear(768 -> hidden)
  clip_cls_rows_kernel<<<V, 96, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x1bf), reinterpret_cast<__nv_bfloat16*>(b + l.cls), T1, D / 8);
  if ((rc = gemm(b + l.cls, w->proj_w, w->proj_b, out, V, w->hidden, D, RVL_GEMM_OUT_BF16, 0))) return rc;
  return cudaGetLastError() == cudaSuccess ? RVL_OK : report_error(h, RVL_ERR_CUDA, "rvl_clip_encoder: launch failed");
}

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
              const int32_t* kv_seq_idx, int32_t n_kv_seq, const float* key_mask, rvl_stream stream) {
  if (!q || !k || !v || !out || n_seq <= 0 || Tq <= 0 || Tk <= 0) return report_error(h, RVL_ERR_INVALID, "rvl_mha96: null argument");
  if (q_stride % 8 || k_stride % 8 || v_stride % 8 || out_stride % 2)
    return report_error(h, RVL_ERR_INVALID, "rvl_mha96: q / k / v row strides must be multiples of 8 elements, the out stride of 2");
  if (kv_seq_idx && n_kv_seq <= 0) return report_error(h, RVL_ERR_INVALID, "rvl_mha96: kv_seq_idx needs n_kv_seq, the number of key / value sequences");
  return launch_mha96(q, q_stride, k, k_stride, v, v_stride, out, out_stride, n_seq, n_heads, Tq, Tk, kv_seq_idx, key_mask,
                      static_cast<cudaStream_t>(stream), kv_seq_idx ? n_kv_seq : n_seq, handle_num_sms(h));
}

}  // extern "C"
