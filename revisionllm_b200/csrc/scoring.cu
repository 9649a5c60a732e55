// CLIP text-to-frame cosine top-k scoring and stage-2 segment selection.
//
// Replaces revisionllm/eval/similarity.py:71-94 (`_topk_pooling`: sims = video @ text^T, top-k frames,
// gather + SUM of the top-k frame vectors) together with the caller arithmetic
// revisionllm/eval/eval_nlq_negative.py:309-316 (frames / frames.norm(dim=0), einsum with cls) and
// revisionllm/eval/eval_nlq_retrieval_e2e2.py:380-386 (per-frame norm), i.e.
//   score(proposal) = dot(sum_{f in top-k} normalised_frame_f, cls) = sum of the k largest sims.
// HBM-bound: each frame row (dim * 2 B) is read once (norm_axis = 1) or twice (norm_axis = 0, the second
// pass hits L2/L1).  One CTA per proposal.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rvl_internal.h"
#include "rvl_ptx.cuh"

namespace rvl {

constexpr int kScoreThreads = 256;
constexpr int kMaxK = 16;

__global__ void __launch_bounds__(kScoreThreads) cosine_topk_kernel(const __nv_bfloat16* __restrict__ frames,
                                                                     const int32_t* __restrict__ seg_offsets,
                                                                     const int32_t* __restrict__ seg_ends, int dim,
                                                                     const __nv_bfloat16* __restrict__ cls, int k,
                                                                     int norm_axis, float* __restrict__ scores_out,
                                                                     int32_t* __restrict__ topk_idx_out) {
  extern __shared__ float s_w[];  // [dim] per-dimension weight: cls_d (axis 1) or cls_d / colnorm_d (axis 0); then sims
  const int seg = blockIdx.x;
  const int r0 = seg_offsets[seg], r1 = seg_ends ? seg_ends[seg] : seg_offsets[seg + 1];
  const int n = r1 > r0 ? r1 - r0 : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __nv_bfloat16* base = frames + static_cast<long long>(r0) * dim;
  for (int d = tid; d < dim; d += kScoreThreads) {
    float w = __bfloat162float(cls[d]);
    if (norm_axis == 0) {
      float ss = 0.f;
      for (int f = 0; f < n; ++f) {
        const float v = __bfloat162float(base[static_cast<long long>(f) * dim + d]);
        ss += v * v;
      }
      w = w / sqrtf(ss);
    }
    s_w[d] = w;
  }
  __syncthreads();
  float* sims = s_w + dim;  // [n] for this proposal
  for (int f = warp; f < n; f += kScoreThreads / 32) {
    const __nv_bfloat16* row = base + static_cast<long long>(f) * dim;
    float dot = 0.f, ss = 0.f;
    for (int d = lane * 8; d < dim; d += 32 * 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + d));
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = bf16_lo(w4[e]), b = bf16_hi(w4[e]);
        dot += a * s_w[d + 2 * e] + b * s_w[d + 2 * e + 1];
        ss += a * a + b * b;
      }
    }
    dot = warp_sum(dot);
    ss = warp_sum(ss);
    if (lane == 0) sims[f] = norm_axis == 1 ? dot / sqrtf(ss) : dot;  // axis 0 is folded into s_w; 2 = raw dot
  }
  __syncthreads();
  // top-k by repeated arg-max over the (short) sims row; ties -> lowest index. Warp 0 only.
  if (warp == 0) {
    const int kk = k < n ? k : n;
    float score = 0.f;
    int chosen[kMaxK];
    for (int it = 0; it < kk; ++it) {
      float best = -INFINITY;
      int besti = 0x7fffffff;
      for (int f = lane; f < n; f += 32) {
        bool taken = false;
        for (int c = 0; c < it; ++c) taken |= (chosen[c] == f);
        const float v = sims[f];
        if (!taken && (v > best || (v == best && f < besti))) { best = v; besti = f; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
      }
      chosen[it] = besti;
      score += best;
      if (lane == 0 && topk_idx_out) topk_idx_out[seg * k + it] = besti;
    }
    if (lane == 0) {
      scores_out[seg] = score;
      if (topk_idx_out)
        for (int it = kk; it < k; ++it) topk_idx_out[seg * k + it] = -1;
    }
  }
}

void launch_cosine_topk(const void* frames, const int32_t* seg_offsets, const int32_t* seg_ends, int n_seg, int dim, const void* cls, int k,
                        int norm_axis, int max_seg_rows, float* scores_out, int32_t* topk_idx_out, cudaStream_t st) {
  if (n_seg <= 0) return;
  cosine_topk_kernel<<<n_seg, kScoreThreads, (dim + max_seg_rows) * sizeof(float), st>>>(
      reinterpret_cast<const __nv_bfloat16*>(frames), seg_offsets, seg_ends, dim, reinterpret_cast<const __nv_bfloat16*>(cls), k,
      norm_axis, scores_out, topk_idx_out);
}

// idx_out[rank] = i for the k best scores; rank_i = #{j : s_j > s_i or (s_j == s_i and j < i)}.
// Exact (comparison only), order-independent, so the selection is bit-reproducible.
__global__ void select_topk_kernel(const float* __restrict__ scores, int n, int k, int32_t* __restrict__ idx_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float si = scores[i];
  int rank = 0;
  for (int j = 0; j < n; ++j) {
    const float sj = __ldg(scores + j);
    rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
  }
  if (rank < k) idx_out[rank] = i;
}
void launch_select_topk(const float* scores, int n, int k, int32_t* idx_out, cudaStream_t st) {
  if (n <= 0 || k <= 0) return;
  select_topk_kernel<<<(n + 127) / 128, 128, 0, st>>>(scores, n, k, idx_out);
}

// ------------------------------------------------------------------------------------------- merge + rank
// The per-query scoring tail of the reference, on the device and in double precision like the Python floats it
// restates (so the ranking is bit-identical to a float64 numpy restatement):
//   1. over the kept proposals (keep != 0: the answer parsed to a span): cos /= max(cos), ent /= max(ent) when
//      `normalize` (revisionllm/eval/eval_nlq_negative.py:321-327);
//   2. merged = cos - ent (mode 0, --score_merge add) | cos / ent (1, multiply) | -ent (2) | cos (3)   (:328-336);
//   3. min-max normalisation of the merged scores when `minmax` and min != max
//      (revisionllm/eval/metric_retrieval_forward.py:146-152);
//   4. stage-2 cover filter (:119-141): if any kept proposal lies in cover1, only kept proposals inside cover_all
//      survive; otherwise all kept proposals do;
//   5. order = surviving proposals by descending score, ties in index order (Python's stable sorted(..., reverse=True) of
//      grounding_metrics_stream, :38).
// One CTA; n <= 8192 proposals per query (MAD: 57 - 143 windows).  scores_out[i] is NaN for proposals that are not kept.
constexpr int kRankThreads = 1024;

__device__ __forceinline__ double block_reduce_max(double v, double* red, bool want_min) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double x = __shfl_xor_sync(0xffffffffu, v, o);
    v = want_min ? fmin(v, x) : fmax(v, x);
  }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double r = red[0];
  for (int w = 1; w < kRankThreads / 32; ++w) r = want_min ? fmin(r, red[w]) : fmax(r, red[w]);
  return r;
}

__global__ void __launch_bounds__(kRankThreads) merge_rank_kernel(const float* __restrict__ cos, const float* __restrict__ ent,
                                                                   const int32_t* __restrict__ keep,
                                                                   const int32_t* __restrict__ cover1,
                                                                   const int32_t* __restrict__ cover_all, int n, int mode,
                                                                   int normalize, int minmax, double* __restrict__ scores_out,
                                                                   int32_t* __restrict__ order_out, int32_t* __restrict__ n_out) {
  __shared__ double red[kRankThreads / 32];
  __shared__ int s_any;
  const int tid = threadIdx.x;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  if (tid == 0) s_any = 0;
  double mc = -inf, me = -inf;
  for (int i = tid; i < n; i += kRankThreads)
    if (keep[i]) {
      if (cos) mc = fmax(mc, static_cast<double>(cos[i]));
      if (ent) me = fmax(me, static_cast<double>(ent[i]));
    }
  mc = block_reduce_max(mc, red, false);
  me = block_reduce_max(me, red, false);
  double lo = inf, hi = -inf;
  for (int i = tid; i < n; i += kRankThreads) {
    double sc = __longlong_as_double(0x7ff8000000000000LL);
    if (keep[i]) {
      double c = cos ? static_cast<double>(cos[i]) : 0.0, e = ent ? static_cast<double>(ent[i]) : 0.0;
      if (normalize) { c = c / mc; e = e / me; }
      sc = mode == 0 ? c - e : (mode == 1 ? c / e : (mode == 2 ? -e : c));
      lo = fmin(lo, sc);
      hi = fmax(hi, sc);
      if (cover1 && cover1[i]) s_any = 1;
    }
    scores_out[i] = sc;
  }
  lo = block_reduce_max(lo, red, true);
  hi = block_reduce_max(hi, red, false);
  __syncthreads();
  // the reference min-max normalises only inside the branch that applies the stage-2 filter (metric_retrieval_forward.py:137)
  if (minmax && lo != hi && (cover1 == nullptr || s_any != 0)) {
    for (int i = tid; i < n; i += kRankThreads)
      if (keep[i]) scores_out[i] = (scores_out[i] - lo) / (hi - lo);
  }
  __syncthreads();
  const bool filter = cover1 != nullptr && s_any != 0;
  const int32_t* cov = cover_all ? cover_all : cover1;
  int alive_local = 0;
  for (int i = tid; i < n; i += kRankThreads) {
    const bool alive = keep[i] && (!filter || cov[i]);
    if (!alive) continue;
    ++alive_local;
    const double si = scores_out[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      if (!(keep[j] && (!filter || cov[j]))) continue;
      const double sj = scores_out[j];
      rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
    }
    order_out[rank] = i;
  }
  __shared__ int s_cnt;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  atomicAdd(&s_cnt, alive_local);
  __syncthreads();
  if (tid == 0) *n_out = s_cnt;
}

void launch_merge_rank(const float* cos, const float* ent, const int32_t* keep, const int32_t* cover1, const int32_t* cover_all,
                       int n, int mode, int normalize, int minmax, double* scores_out, int32_t* order_out, int32_t* n_out,
                       cudaStream_t st) {
  merge_rank_kernel<<<1, kRankThreads, 0, st>>>(cos, ent, keep, cover1, cover_all, n, mode, normalize, minmax, scores_out,
                                                order_out, n_out);
}

}  // namespace rvl
