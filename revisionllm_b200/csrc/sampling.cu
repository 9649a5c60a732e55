// Greedy sampling, per-step entropy and EOS bookkeeping on the device.
//
// Replaces the per-step tail of the reference's generation loop
// (revisionllm/model/vtimellm_llama.py:321-362): `scores` = raw last-row logits, next token
// (argmax here instead of softmax + multinomial, BASELINE.json north_star), finished rows emit
// pad, a row finishes when it emits EOS; and the entropy of
// revisionllm/uncertainty/funs_get_feature_X.py:130-134: p = softmax(logits),
// H = -sum p * log(p + 1e-10), so the host never materialises [B, T, V].
// One CTA per row; the fp32 row (128 KB at V = 32000) is read three times (max/argmax, sum of
// exponentials, entropy) - passes 2 and 3 hit L2.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "rvl_internal.h"
#include "rvl_ptx.cuh"

namespace rvl {

constexpr int kSampleThreads = 256;

__global__ void __launch_bounds__(kSampleThreads) sample_greedy_kernel(const float* __restrict__ logits, int vocab,
                                                                        int32_t* __restrict__ unfinished, int eos_id,
                                                                        int pad_id, int32_t* __restrict__ next_tokens,
                                                                        float* __restrict__ entropy_out) {
  __shared__ float s_val[kSampleThreads / 32];
  __shared__ int s_idx[kSampleThreads / 32];
  __shared__ float s_sum[kSampleThreads / 32];
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x;
  const float* x = logits + static_cast<long long>(row) * vocab;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // pass 1: max and argmax, ties -> lowest index
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int i = tid; i < vocab; i += kSampleThreads) {
    const float v = x[i];
    if (v > best || (v == best && i < besti)) { best = v; besti = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  if (lane == 0) { s_val[warp] = best; s_idx[warp] = besti; }
  __syncthreads();
  best = s_val[0]; besti = s_idx[0];
#pragma unroll
  for (int w = 1; w < kSampleThreads / 32; ++w) {
    if (s_val[w] > best || (s_val[w] == best && s_idx[w] < besti)) { best = s_val[w]; besti = s_idx[w]; }
  }
  // pass 2: sum of exp
  float sum = 0.f;
  for (int i = tid; i < vocab; i += kSampleThreads) sum += expf(x[i] - best);
  sum = warp_sum(sum);
  if (lane == 0) s_sum[warp] = sum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kSampleThreads / 32; ++w) tot += s_sum[w];
  __syncthreads();
  // pass 3: entropy with the reference's +1e-10 inside the log
  const float inv = 1.f / tot;
  float h = 0.f;
  for (int i = tid; i < vocab; i += kSampleThreads) {
    const float p = expf(x[i] - best) * inv;
    h -= p * logf(p + 1e-10f);
  }
  h = warp_sum(h);
  if (lane == 0) s_sum[warp] = h;
  __syncthreads();
  if (tid == 0) {
    float H = 0.f;
#pragma unroll
    for (int w = 0; w < kSampleThreads / 32; ++w) H += s_sum[w];
    if (entropy_out) entropy_out[row] = H;
    int tok = besti;
    if (unfinished) {
      const int u = unfinished[row];
      tok = u ? tok : pad_id;                       // vtimellm_llama.py:343-347
      unfinished[row] = (u && tok != eos_id) ? 1 : 0;  // :352-356
    }
    next_tokens[row] = tok;
  }
}

void launch_sample_greedy(const float* logits, int n_seq, int vocab, int32_t* unfinished, int eos_id, int pad_id,
                          int32_t* next_tokens, float* entropy_out, int32_t* /*seq_lens*/, int32_t* /*n_unfinished*/,
                          cudaStream_t st) {
  if (n_seq <= 0) return;
  launch_k(sample_greedy_kernel, dim3(n_seq), dim3(kSampleThreads), 0, st, logits, vocab, unfinished, eos_id, pad_id, next_tokens,
           entropy_out);
}

}  // namespace rvl
