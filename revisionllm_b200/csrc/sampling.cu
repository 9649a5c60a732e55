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

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"): counter-based, so the draw of
// (seed, step, row) is reproducible whatever the batch composition.  Checked against the Random123 known-answer vectors
// (tests/test_host_logic.py through the oracle's restatement, tests/test_gpu_kernels.py on the device).
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__global__ void __launch_bounds__(kSampleThreads) sample_greedy_kernel(const float* __restrict__ logits, int vocab,
                                                                        int32_t* __restrict__ unfinished, int eos_id,
                                                                        int pad_id, int32_t* __restrict__ next_tokens,
                                                                        float* __restrict__ entropy_out, float inv_temperature,
                                                                        unsigned long long seed, uint32_t step,
                                                                        uint32_t* __restrict__ philox_out) {
  __shared__ float s_val[kSampleThreads / 32];
  __shared__ float s_scan[kSampleThreads];
  __shared__ int s_pick;
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
  // ---- multinomial draw from softmax(logits / T) (the reference's sample(): logits_warper temperature 0.05, softmax,
  // torch.multinomial - revisionllm/model/vtimellm_llama.py:312-338) by inverse CDF: every thread sums a contiguous chunk
  // of exp((x - max) / T), a block scan finds the chunk that holds u * total, its owner walks it.
  int sampled = besti;
  if (inv_temperature > 0.f) {
    const int chunk = (vocab + kSampleThreads - 1) / kSampleThreads;
    const int i0 = tid * chunk, i1 = min(vocab, i0 + chunk);
    float part = 0.f;
    for (int i = i0; i < i1; ++i) part += expf((x[i] - best) * inv_temperature);
    s_scan[tid] = part;
    if (tid == 0) s_pick = -1;
    __syncthreads();
    if (tid == 0) {
      uint32_t c[4] = {static_cast<uint32_t>(row), step, 0u, 0u};
      philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
      if (philox_out) { philox_out[row * 4] = c[0]; philox_out[row * 4 + 1] = c[1]; philox_out[row * 4 + 2] = c[2]; philox_out[row * 4 + 3] = c[3]; }
      const float u = static_cast<float>(c[0] >> 8) * (1.0f / 16777216.0f);      // [0, 1)
      float total = 0.f;
      for (int t = 0; t < kSampleThreads; ++t) total += s_scan[t];
      const float target = u * total;
      float run = 0.f;
      int owner = -1;
      for (int t = 0; t < kSampleThreads; ++t) {
        if (s_scan[t] > 0.f) owner = t;                         // last non-empty chunk catches target == total
        if (target < run + s_scan[t]) { owner = t; break; }
        run += s_scan[t];
      }
      s_pick = owner;
      s_scan[owner] = target - run;                              // what is left to walk inside the chunk
    }
    __syncthreads();
    if (tid == s_pick) {
      const float rest = s_scan[tid];
      float run = 0.f;
      int pick = i1 - 1;
      for (int i = i0; i < i1; ++i) {
        run += expf((x[i] - best) * inv_temperature);
        if (rest < run) { pick = i; break; }
      }
      s_pick = -2 - pick;
    }
    __syncthreads();
    sampled = -2 - s_pick;
  }
  if (tid == 0) {
    float H = 0.f;
#pragma unroll
    for (int w = 0; w < kSampleThreads / 32; ++w) H += s_sum[w];
    if (entropy_out) entropy_out[row] = H;
    int tok = sampled;
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
           entropy_out, 0.f, 0ull, 0u, static_cast<uint32_t*>(nullptr));
}

void launch_sample_multinomial(const float* logits, int n_seq, int vocab, float temperature, unsigned long long seed,
                               uint32_t step, int32_t* unfinished, int eos_id, int pad_id, int32_t* next_tokens,
                               float* entropy_out, uint32_t* philox_out, cudaStream_t st) {
  if (n_seq <= 0) return;
  launch_k(sample_greedy_kernel, dim3(n_seq), dim3(kSampleThreads), 0, st, logits, vocab, unfinished, eos_id, pad_id, next_tokens,
           entropy_out, 1.0f / temperature, seed, step, philox_out);
}

}  // namespace rvl
