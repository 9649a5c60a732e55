// Internal declarations shared by the .cu translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>

#include <string>
#include <utility>

#include "../../include/revisionllm_b200.h"

struct CUtensorMap_st;          // CUDA driver API tensor map (cuda.h), used by pointer only here

namespace rvl {

struct GemmCall {
  const void* A = nullptr;     // [M, K] bf16 activations (tokens)
  const void* W = nullptr;     // [N, K] bf16 weights (features)
  const void* bias = nullptr;  // [N] bf16 or null
  void* out = nullptr;         // [M(rowmap), ldc] bf16 / fp32
  int64_t M = 0, N = 0, K = 0, ldc = 0;
  int out_mode = RVL_GEMM_OUT_BF16;
  int flags = 0;
  const int32_t* rowmap = nullptr;
  int split_k = 1;             // explicit k split (tile mode); > 1 needs RVL_GEMM_ADD_F32 (atomics) or split_stride
  int64_t split_stride = 0;    // > 0: split-k partials written to out + s * split_stride (fp32, no atomics)
  int* split_used = nullptr;   // out: the split actually launched
  // stream-K workspace (weight-streaming orientation only): fp32 partial tiles and per-CTA flags
  float* stream_ws = nullptr;
  size_t stream_ws_bytes = 0;
  unsigned int* stream_flags = nullptr;   // per-CTA "partial published" flags, zero between launches (the consumer clears them)
};
int gemm_bf16(const GemmCall& c, int num_sms, cudaStream_t st, std::string* err);

// Diagnostic switches ("Diagnostic switches"): the RVL_* environment variables are read ONCE - at the first
// rvl_create of the process - into this table; nothing on the hot path calls getenv().  tools/ that flip a switch inside one
// process call rvl_reload_env() afterwards.  -1 / 0 = "not set" where a switch has a default that depends on the problem.
struct Tuning {
  int pdl = -1;              // RVL_PDL bit mask (-1: GEMMs always, few-row norm / attention up to 128 rows)
  int a_tiles = 0;           // RVL_A_TILES
  int stream_k = -1;         // RVL_STREAM_K
  int plan_debug = 0;        // RVL_PLAN_DEBUG
  int pair = -1;             // RVL_PAIR
  int spair = -1;            // RVL_SPAIR
  int spair_streamk = -1;    // RVL_SPAIR_STREAMK
  int staged = -1;           // RVL_STAGED
  int group_m = 0;           // RVL_GROUP_M
  int full_last_layer = 0;   // RVL_FULL_LAST_LAYER
  int attn_decode = 0;       // RVL_ATTN_DECODE: 'r' / 's' / 0
  int attn_ps32 = -1;        // RVL_ATTN_PS32
  int attn_prefill = -1;     // RVL_ATTN_PREFILL: 0 = mma.sync kernel, 2 = mma.sync only with an external context, else tcgen05
  int attn_mha96 = -1;       // RVL_ATTN_MHA96: 0 = mma.sync kernel, else tcgen05
  int norm_threads = 0;      // RVL_NORM_THREADS: threads of the few-row RMSNorm CTA (256 / 512 / 1024)
};
const Tuning& tuning();      // engine.cu
void reload_tuning();

// Programmatic dependent launch: RVL_PDL=0 in the environment switches it off (plain stream order).
// Bit 0: the GEMM launches, bit 1: every other kernel.  Default 1: measured on B200 (decode step, 7B shape, B = 180 / 32)
// 10.08 / 5.71 ms without, 9.72 / 5.18 ms with GEMMs only, 9.90 / 5.67 ms with everything - small kernels that become
// resident early only squat on the SM while the GEMM before them is still streaming.
// Bit 2: the few-row (decode) RMSNorm only, bit 3: the decode attention kernels only.  Tools A/B in one process through rvl_reload_env().
inline int pdl_mask() {
  const int e = tuning().pdl;
  return e >= 0 ? e : 1;
}
// Decode-step RMSNorm (bit 2) and attention (bit 3): early launch pays for few rows - measured per 7B decode step with both
// on: 3.12 -> 3.01 ms at B = 1, 4.24 -> 4.13 at 23, 5.36 -> 5.08 at 56, 6.43 -> 6.28 at 96 - and costs at B = 180
// (8.50 -> 8.77 ms: the 1024-thread RMSNorm CTAs cannot share an SM with a streaming GEMM CTA).  Default: on up to 128 rows.
inline bool pdl_few_rows(long long rows, int bit) {
  const int e = tuning().pdl;
  if (e >= 0) return (e & (2 | bit)) != 0;
  return rows <= 128;
}
// Launch `kernel` so that it may start while the previous kernel of `st` is still running (it must call pdl_wait()
// before touching anything earlier kernels write; see rvl_ptx.cuh).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_pdl((pdl_mask() & 2) != 0, kernel, grid, block, smem, st, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_gemm_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_pdl((pdl_mask() & 1) != 0, kernel, grid, block, smem, st, std::forward<Args>(args)...);
}

// engine.cu: sets rvl_last_error (handle and thread) and returns `code`
int report_error(const rvl_handle* h, int code, const char* msg);
int handle_num_sms(const rvl_handle* h);   // SM count of the handle's device (0 for a null handle)

// elementwise.cu
// y = rmsnorm(x + sum_p partials[p]) ; when x_out != null the summed row is written back (residual update)
void launch_rmsnorm(const float* x, const void* w, void* y, int64_t n_rows, int dim, float eps, const int32_t* rows,
                    cudaStream_t st, const float* partials = nullptr, int n_partials = 0, int64_t partial_stride = 0,
                    float* x_out = nullptr);
void launch_embed_rows(const void* table, const int32_t* ids, const int32_t* dst_rows, int n, int dim, int vocab,
                       float* out, cudaStream_t st);
void launch_scatter_rows_bf16(const void* src, const int32_t* dst_rows, int n, int dim, float* out, cudaStream_t st);
void launch_rope_kv(void* qkv, int64_t n_tokens, const int32_t* positions, const int32_t* tok_seq,
                    const int32_t* cu_seqlens, const int32_t* page_table, int max_pages, void* k_pages, void* v_pages,
                    int n_heads, int page_size, float theta, cudaStream_t st, const int32_t* seq_pos0 = nullptr);
void launch_gather_rows_f32_bf16(const float* src, const int32_t* idx, int n_rows, int n_src, int dim, void* out, cudaStream_t st);
void launch_gather_last_rows(const void* attn, const float* hidden, const int32_t* rows, int n_rows, int dim, void* attn_out,
                             float* hidden_out, cudaStream_t st);
void launch_swiglu(const void* gu, void* act, int64_t n_tokens, int inter, cudaStream_t st);
void launch_token_seq(const int32_t* cu_seqlens, int n_seq, int32_t* tok_seq, int32_t* last_rows, int64_t total,
                      cudaStream_t st);
// gemm_tcgen05.cu: 2D bf16 row-major [rows, cols] tensor map, box = [box_rows, 64 cols], 128-byte swizzle, OOB -> zeros;
// ld = row pitch in elements when the matrix is a column slice of a wider one (0: dense rows)
int make_tmap_bf16_2d(::CUtensorMap_st* tm, const void* base, int64_t rows, int64_t cols, int box_rows, std::string* err, int64_t ld = 0);
// attention.cu / attention_tcgen05.cu
void launch_attn_prefill(const void* qkv, void* out, const int32_t* cu_seqlens, int n_seq, int max_seqlen, int n_heads,
                         cudaStream_t st, const int32_t* seq_pos0 = nullptr, const int32_t* seq_ctx_row = nullptr, int only_last = 0,
                         int64_t total_tokens = 0, int num_sms = 0);
bool launch_attn_prefill_tc(const void* qkv, void* out, const int32_t* cu_seqlens, int n_seq, int64_t total_tokens, int max_seqlen,
                            int n_heads, int num_sms, cudaStream_t st, int only_last, const int32_t* seq_pos0 = nullptr,
                            const int32_t* seq_ctx_row = nullptr);
bool launch_mha96_tc(const void* q, long long q_stride, const void* k, long long k_stride, const void* v, long long v_stride, void* out,
                     long long out_stride, int n_seq, int n_kv_seq, int n_heads, int Tq, int Tk, const int32_t* kv_seq_idx,
                     const float* key_mask, int num_sms, cudaStream_t st);
void launch_attn_decode(const void* qkv, void* out, const int32_t* seq_lens, int n_seq, const int32_t* page_table,
                        int max_pages, void* k_pages, void* v_pages, int n_heads, int page_size, int fused, float theta,
                        int max_kv_len, cudaStream_t st);
int launch_mha96(const void* q, long long q_stride, const void* k, long long k_stride, const void* v, long long v_stride, void* out,
                 long long out_stride, int n_seq, int n_heads, int Tq, int Tk, const int32_t* kv_seq_idx, const float* key_mask,
                 cudaStream_t st, int n_kv_seq = 0, int num_sms = 0);
// sampling.cu
void launch_sample_greedy(const float* logits, int n_seq, int vocab, int32_t* unfinished, int eos_id, int pad_id,
                          int32_t* next_tokens, float* entropy_out, int32_t* seq_lens, int32_t* n_unfinished,
                          cudaStream_t st);
void launch_sample_multinomial(const float* logits, int n_seq, int vocab, float temperature, unsigned long long seed,
                               uint32_t step, int32_t* unfinished, int eos_id, int pad_id, int32_t* next_tokens,
                               float* entropy_out, uint32_t* philox_out, cudaStream_t st);
// scoring.cu
void launch_cosine_topk(const void* frames, const int32_t* seg_offsets, const int32_t* seg_ends, int n_seg, int dim, const void* cls, int k,
                        int norm_axis, int max_seg_rows, float* scores_out, int32_t* topk_idx_out, cudaStream_t st);
void launch_select_topk(const float* scores, int n, int k, int32_t* idx_out, cudaStream_t st);
void launch_merge_rank(const float* cos, const float* ent, const int32_t* keep, const int32_t* cover1, const int32_t* cover_all,
                       int n, int mode, int normalize, int minmax, double* scores_out, int32_t* order_out, int32_t* n_out,
                       cudaStream_t st);

}  // namespace rvl
