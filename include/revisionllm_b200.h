/*
 * revisionllm_b200 - C ABI of the Blackwell (sm_100a) kernels behind the
 * ReVisionLLM recursive segment-scoring inference path.
 *
 * The reference (Tanveer81/ReVisionLLM) has no FFI boundary of its own: its hot
 * path is Python calling HuggingFace `transformers` / torch.
 * This header is the boundary a maintainer binds (ctypes stub in
 * INTEGRATION.md); each entry point names the reference code it replaces.
 * All paths below are relative to the reference tree.
 *
 * Conventions
 *  - plain C, no C++ types, no exceptions across the boundary;
 *  - every function returns RVL_OK (0) or a negative rvl_status; the message is
 *    available from rvl_last_error(h) (or rvl_last_error(NULL) for handle-less calls);
 *  - every tensor is allocated and owned by the caller (PyTorch): the library
 *    receives raw device pointers + sizes, never frees them, and keeps pointers
 *    past a call only for bound weights, the workspace and the KV pages;
 *  - every call enqueues work on the caller's cudaStream_t and returns without
 *    synchronising; no host allocation or cudaMalloc on the hot path;
 *  - one handle per (process, GPU); a handle is not thread-safe;
 *  - there is NO CPU fallback: without an sm_100 device the calls fail.
 */
#ifndef REVISIONLLM_B200_H
#define REVISIONLLM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVL_ABI_VERSION 4

#if defined(__GNUC__)
#define RVL_API __attribute__((visibility("default")))
#else
#define RVL_API
#endif

typedef enum rvl_status {
  RVL_OK = 0,
  RVL_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  RVL_ERR_CUDA = -2,        /* CUDA runtime / driver error */
  RVL_ERR_STATE = -3,       /* weights / workspace / kv not bound */
  RVL_ERR_UNSUPPORTED = -4  /* device is not sm_100 */
} rvl_status;

typedef struct rvl_handle rvl_handle;
typedef void* rvl_stream; /* cudaStream_t */

/* Llama-2-7B / Vicuna-7B-v1.5 shape by default. head_dim must be 128. */
typedef struct rvl_config {
  int32_t hidden;        /* 4096 */
  int32_t n_layers;      /* 32 */
  int32_t n_heads;       /* 32 (no GQA) */
  int32_t head_dim;      /* 128 */
  int32_t intermediate;  /* 11008 */
  int32_t vocab;         /* 32000 */
  int32_t adapter_dim;   /* 768 (CLIP ViT-L/14) */
  int32_t max_pos;       /* 4096 */
  int32_t kv_page_size;  /* tokens per KV page; 32 */
  int32_t device;        /* CUDA device ordinal */
  float rms_eps;         /* 1e-5 */
  float rope_theta;      /* 10000 */
} rvl_config;

/* Per-layer weights, all bf16 row-major [out_features, in_features] like the HF
 * state dict; q/k/v and gate/up are concatenated along out_features by the host
 * once at load time:
 *   wqkv  = cat(self_attn.{q,k,v}_proj.weight)        [3*hidden, hidden]
 *   wo    = self_attn.o_proj.weight                   [hidden, hidden]
 *   wgu   = cat(mlp.gate_proj.weight, up_proj.weight) [2*intermediate, hidden]  (rvl_weights.wgu_layout 0), or the same
 *           rows interleaved in blocks of 32 = [gate 16b..16b+15 | up 16b..16b+15] (wgu_layout 1, intermediate % 16 == 0):
 *           SwiGLU is then fused into the gate/up GEMM epilogue and the [tokens, 2*intermediate] tensor is never written
 *   wdown = mlp.down_proj.weight                      [hidden, intermediate]
 *   ln1   = input_layernorm.weight, ln2 = post_attention_layernorm.weight  [hidden] */
typedef struct rvl_layer_weights {
  const void* wqkv;
  const void* wo;
  const void* wgu;
  const void* wdown;
  const void* ln1;
  const void* ln2;
} rvl_layer_weights;

typedef struct rvl_weights {
  const void* embed_tokens;        /* model.embed_tokens.weight [vocab, hidden] bf16 */
  const void* final_norm;          /* model.norm.weight [hidden] bf16 */
  const void* lm_head;             /* lm_head.weight [vocab, hidden] bf16 */
  const void* proj_w;              /* model.mm_projector.weight [hidden, adapter_dim] bf16 (stage 1) */
  const void* proj_b;              /* model.mm_projector.bias [hidden] bf16 */
  const rvl_layer_weights* layers; /* n_layers entries (host array, copied) */
  int32_t wgu_layout;              /* 0: cat(gate, up); 1: 16-row interleave (see rvl_layer_weights) */
} rvl_weights;

/* ---- lifecycle ------------------------------------------------------------------------- */
RVL_API int rvl_abi_version(void);
RVL_API const char* rvl_last_error(const rvl_handle* h);
RVL_API int rvl_create(const rvl_config* cfg, rvl_handle** out);
RVL_API void rvl_destroy(rvl_handle* h);
/* Replaces builder.py:21-67 `load_pretrained_model` at the device boundary: borrows the pointers. */
RVL_API int rvl_bind_weights(rvl_handle* h, const rvl_weights* w);
/* Scratch for up to max_tokens packed prompt tokens and max_seqs sequences. */
RVL_API size_t rvl_workspace_bytes(const rvl_handle* h, int64_t max_tokens, int32_t max_seqs);
RVL_API int rvl_set_workspace(rvl_handle* h, void* ws, size_t bytes, int64_t max_tokens, int32_t max_seqs);
/* Paged KV cache: [layer][k|v][page][head][slot][head_dim] bf16. */
RVL_API size_t rvl_kv_bytes(const rvl_handle* h, int32_t n_pages);
RVL_API int rvl_set_kv(rvl_handle* h, void* kv, int32_t n_pages);

/* ---- hot path --------------------------------------------------------------------------- */

/* Projector GEMM fused with the embedding splice.
 * Replaces vtimellm_arch.py:42,125 (`mm_projector`) + :149-244 (split ids at the
 * placeholder, embed text chunks, concat) of `prepare_inputs_labels_for_multimodal`.
 *   feats        [n_feat_rows, adapter_dim] bf16, all segments' frames concatenated
 *   feat_dst     [n_feat_rows] int32: destination row of each frame in the packed stream
 *   text_ids     [n_text] int32 token ids (placeholders removed)
 *   text_dst     [n_text] int32 destination rows
 *   hidden_out   [total_tokens, hidden] fp32 packed residual stream (written at the dst rows) */
RVL_API int rvl_project_splice(rvl_handle* h, const void* feats, const int32_t* feat_dst, int32_t n_feat_rows,
                       const int32_t* text_ids, const int32_t* text_dst, int32_t n_text,
                       float* hidden_out, int64_t total_tokens, rvl_stream stream);

/* Window builder of the feature loader: out[i] (bf16 [n_rows, dim]) = features[frame_idx[i]] (fp32 [n_frames, dim]).
 * Replaces the host-side `features[np.linspace(start, end, num_frames, dtype=int32)]` per window of
 * eval_nlq_negative.py:224-235 / eval_nlq_retrieval_e2e2.py:262-277 and the later cast to bf16: the movie's features
 * are copied to the GPU once, as stored, and every (overlapping) window is gathered there.  Indices are clamped. */
RVL_API int rvl_gather_windows(rvl_handle* h, const float* features, int32_t n_frames, int32_t dim,
                       const int32_t* frame_idx, int32_t n_rows, void* out, rvl_stream stream);

/* Same splice for already-projected visual rows (stage-2: one ClipEncoder CLS row per segment,
 * vtimellm_arch.py:114-121): vis [n_vis, hidden] bf16 scattered to vis_dst rows. */
RVL_API int rvl_splice_rows(rvl_handle* h, const void* vis, const int32_t* vis_dst, int32_t n_vis,
                    const int32_t* text_ids, const int32_t* text_dst, int32_t n_text,
                    float* hidden_out, int64_t total_tokens, rvl_stream stream);

/* Varlen causal prefill of the decoder stack over a packed token stream.
 * Replaces transformers LlamaModel/LlamaForCausalLM.forward reached from
 * vtimellm_llama.py:79-90 at step 0 (RMSNorm, QKV, RoPE, KV write, causal attention,
 * O, SwiGLU MLP, final norm, lm_head).
 *   hidden       [total_tokens, hidden] fp32 in/out (residual stream, overwritten; with all_logits == 0 the o projection
 *                and the MLP of the LAST layer run on the last row of each sequence only - nothing else reads the other
 *                rows - so on return `hidden` holds the input of the last layer, not its output)
 *   cu_seqlens   [n_seq+1] int32 device
 *   page_table   [n_seq, max_pages] int32 device (KV page ids per sequence)
 *   logits_out   fp32 [n_seq, vocab] (last token of each sequence) or, when all_logits != 0,
 *                [total_tokens, vocab]
 *   seq_pos0     optional int32 [n_seq] device: sequence i's own rows start at position seq_pos0[i]; its positions
 *                [0, seq_pos0[i]) are the rows [seq_ctx_row[i], +seq_pos0[i]) of the same packed stream - a prompt prefix
 *                shared by the batch that is embedded, projected and written to (shared) KV pages ONCE, as its own
 *                sequence, and attended to by every sequence that names it (max_seqlen counts positions, context included).
 *                seq_pos0[i] must be a multiple of the KV page size: a shared prefix consists of whole pages.
 *   seq_ctx_row  optional int32 [n_seq] device, see seq_pos0 (both NULL: every sequence is self-contained) */
RVL_API int rvl_prefill(rvl_handle* h, float* hidden, const int32_t* cu_seqlens, int32_t n_seq,
                int64_t total_tokens, int32_t max_seqlen, const int32_t* page_table, int32_t max_pages,
                float* logits_out, int32_t all_logits, const int32_t* seq_pos0, const int32_t* seq_ctx_row,
                rvl_stream stream);

/* One KV-cached decode step for n_seq sequences (one new token each).
 * Replaces the step t>0 forward: vtimellm_arch.py:88-100 (position = tokens so far)
 * + transformers Llama forward with past_key_values.
 *   token_ids    [n_seq] int32 device: token to embed and append
 *   seq_lens     [n_seq] int32 device, in/out: tokens already in the cache (= position of the new
 *                token); incremented by one at the end of the step
 *   max_kv_len   host-side upper bound of seq_lens[i] + 1 over the batch (sizes the attention kernel's
 *                shared-memory K/V staging; 0 = unknown: max_pages * kv_page_size is assumed)
 *   logits_out   [n_seq, vocab] fp32 */
RVL_API int rvl_decode_step(rvl_handle* h, const int32_t* token_ids, int32_t* seq_lens, int32_t n_seq,
                    const int32_t* page_table, int32_t max_pages, int32_t max_kv_len, float* logits_out,
                    rvl_stream stream);

/* Greedy sampling + per-step entropy + EOS bookkeeping on the device.
 * Replaces vtimellm_llama.py:337-362 (argmax instead of multinomial; finished rows emit pad;
 * a row finishes on EOS) and funs_get_feature_X.py:130-134 (H = -sum p*log(p+1e-10)).
 *   logits        [n_seq, vocab] fp32
 *   unfinished    [n_seq] int32 in/out (1 = still generating); may be NULL (no EOS handling)
 *   next_tokens   [n_seq] int32 out
 *   entropy_out   [n_seq] fp32 out (entropy of this step's distribution; may be NULL) */
RVL_API int rvl_sample_greedy(rvl_handle* h, const float* logits, int32_t n_seq, int32_t vocab,
                      int32_t* unfinished, int32_t eos_id, int32_t pad_id, int32_t* next_tokens,
                      float* entropy_out, rvl_stream stream);

/* The reference's own sampling rule: next token ~ softmax(logits / temperature) (vtimellm_llama.py:312-338 with
 * inference.py:47-48: do_sample=True, temperature=0.05), same EOS / pad bookkeeping and entropy output as
 * rvl_sample_greedy.  The uniform variate of (seed, step, row) comes from Philox4x32-10 with key = seed and counter =
 * (row, step, 0, 0), u = (x0 >> 8) / 2^24, and the draw is the inverse CDF of exp((x - max) / T) in index order, so a
 * run is reproducible and independent of batch composition (torch.multinomial's stream is not reproduced - no two
 * torch versions agree on it either).  philox_out (optional, [n_seq, 4] uint32) returns the raw Philox words (tests). */
RVL_API int rvl_sample_multinomial(rvl_handle* h, const float* logits, int32_t n_seq, int32_t vocab, float temperature,
                           uint64_t seed, uint32_t step, int32_t* unfinished, int32_t eos_id, int32_t pad_id,
                           int32_t* next_tokens, float* entropy_out, uint32_t* philox_out, rvl_stream stream);

/* CLIP text-to-frame cosine top-k score of each proposal.
 * Replaces similarity.py:71-94 `_topk_pooling` + the caller arithmetic
 * eval_nlq_negative.py:309-316 / eval_nlq_retrieval_e2e2.py:380-386:
 * normalise frames (norm_axis 1 = per frame, 0 = across the frame axis, the stage-1 driver's
 * quirk, 2 = none: raw dot products as in _topk_pooling itself), sims = frames . cls, top-k frames (ties -> lowest index), score = dot(sum of the
 * top-k normalised frames, cls) = sum of the top-k sims.
 *   frames   [n_rows, dim] bf16;  seg_offsets [n_seg+1] int32 (rows of each proposal)
 *   seg_ends [n_seg] int32 or NULL: when given, proposal i is rows [seg_offsets[i], seg_ends[i]) - slices that need not
 *            touch each other (eval_nlq_negative.py:309-310 scores `features[k][v0:v1+1]`, the predicted span of a window);
 *            an empty slice scores 0
 *   cls      [dim] bf16;  scores_out [n_seg] fp32;  topk_idx_out [n_seg, k] int32 (-1 padded, may be NULL)
 *   max_seg_rows: upper bound on the rows of one proposal (<= 8192), k <= 16 */
RVL_API int rvl_cosine_topk(rvl_handle* h, const void* frames, const int32_t* seg_offsets, const int32_t* seg_ends,
                    int32_t n_seg, int32_t dim, const void* cls, int32_t k, int32_t norm_axis, int32_t max_seg_rows,
                    float* scores_out, int32_t* topk_idx_out, rvl_stream stream);

/* Stage-2 segment selection: indices of the k largest fp32 scores, descending, ties -> lowest
 * index (bit-exact given identical scores; BASELINE.json north_star). n <= 65536. */
RVL_API int rvl_select_topk(rvl_handle* h, const float* scores, int32_t n, int32_t k, int32_t* idx_out,
                    rvl_stream stream);

/* Per-query scoring tail: normalise, merge, min-max, stage-2 cover filter and ranking of the stage-1 proposals.
 * Replaces eval_nlq_negative.py:317-336 (max-normalisation + `cos - entropy` / `cos / entropy` merge),
 * metric_retrieval_forward.py:119-152 (cover filter, min-max normalisation) and the descending stable sort of
 * grounding_metrics_stream (:38), in double precision like the Python floats of the reference.
 *   cos, ent     [n] fp32 device (either may be NULL when `mode` does not read it)
 *   keep         [n] int32: proposal parsed to a span (not "Not Present", not the 249-249 sentinel)
 *   cover1       [n] int32 or NULL: proposal lies in the windows kept by the first stage-2 log
 *   cover_all    [n] int32 or NULL (= cover1): ... by either stage-2 log
 *   mode         0: cos - ent, 1: cos / ent, 2: -ent, 3: cos;  normalize / minmax: 0 or 1
 *   scores_out   [n] fp64 merged scores (NaN where keep == 0)
 *   order_out    [n] int32: the first *n_out entries are the surviving proposals, best first (ties: lower index)
 * n <= 8192. */
RVL_API int rvl_merge_rank(rvl_handle* h, const float* cos, const float* ent, const int32_t* keep,
                   const int32_t* cover1, const int32_t* cover_all, int32_t n, int32_t mode, int32_t normalize,
                   int32_t minmax, double* scores_out, int32_t* order_out, int32_t* n_out, rvl_stream stream);

/* `n_steps` tokens of the greedy generation loop in one call (vtimellm_llama.py:287-369 for n_steps iterations): for
 * s = 0 .. n_steps - 1: token_ring[s] / entropy_ring[s] = greedy sample (+ entropy, EOS / pad bookkeeping as in
 * rvl_sample_greedy) of `logits`, then one rvl_decode_step on that token writes the next logits back to `logits`.
 * Every pointer is fixed for the whole call, so the call can be captured as ONE CUDA graph and replayed (what
 * engine.decode_chunk does); nothing synchronises.
 *   logits        [n_seq, vocab] fp32, in/out        token_ring   [n_steps, n_seq] int32 out
 *   entropy_ring  [n_steps, n_seq] fp32 out or NULL   unfinished   [n_seq] int32 in/out or NULL */
RVL_API int rvl_decode_n(rvl_handle* h, int32_t n_steps, float* logits, int32_t* token_ring, float* entropy_ring,
                 int32_t* unfinished, int32_t eos_id, int32_t pad_id, int32_t* seq_lens, int32_t n_seq,
                 const int32_t* page_table, int32_t max_pages, int32_t max_kv_len, rvl_stream stream);

/* ---- measurement ------------------------------------------------------------------------------ */
#define RVL_PROF_GEMM 0          /* tcgen05 GEMM, token-major (prefill) */
#define RVL_PROF_GEMM_SMALL_M 1  /* tcgen05 GEMM, weight-streaming orientation (decode / last rows) */
#define RVL_PROF_ATTN_PREFILL 2
#define RVL_PROF_ATTN_DECODE 3
/* Per-launch CUDA-event timing of the kernels rvl_prefill / rvl_decode_step enqueue, on the launching
 * stream (bench.py's roofline numbers).  Off by default; enabling it (re)starts the recording. */
RVL_API int rvl_profile_enable(rvl_handle* h, int32_t on, int32_t capacity);
/* Sums over the recorded launches of one category: device ms, algorithmic FLOPs and bytes. Synchronises. */
RVL_API int rvl_profile_read(rvl_handle* h, int32_t category, double* total_ms, double* total_flops,
                     double* total_bytes, int64_t* launches);

/* tools/gemm_timeline.py only: per-CTA clock64 timestamps of the GEMM roles (8 x uint64 per CTA, up to 160 CTAs).
 * Reads the previous launch's stamps into `out` (n entries) when out != NULL, then switches stamping on/off. */
RVL_API void rvl_debug_gemm_timestamps(int enable, unsigned long long* out, int n);
/* tools/attn_timeline.py only: globaltimer stamps of CTA 0 of the tcgen05 prefill attention kernel, [4 roles][256 events][4]
 * uint64 (see csrc/attention_tcgen05.cu); reads the previous launch's stamps into `out`, then switches stamping on / off. */
RVL_API void rvl_debug_attn_timestamps(int enable, unsigned long long* out, int n);
/* tools/phase_times.py only: enqueue a one-thread kernel that spins ~20 us and writes {SM cycles, nanoseconds} to out_dev
 * (2 x uint64, device memory): the SM clock the stream's kernels see at that point (the power-capped clock inside a sweep). */
RVL_API void rvl_debug_sm_clock(unsigned long long* out_dev, rvl_stream stream);
/* The RVL_* diagnostic switches are read from the environment once, at the first rvl_create of the process;
 * tools that change one afterwards call this to make the library read them again. */
RVL_API void rvl_reload_env(void);

/* ---- individual kernels (unit parity tests and host-side composition) ------------------------ */

#define RVL_GEMM_OUT_BF16 0
#define RVL_GEMM_OUT_F32 1
#define RVL_GEMM_ADD_F32 2      /* out(fp32) += A.B^T, in place, non-atomic */
#define RVL_GEMM_FLAG_RELU 1
#define RVL_GEMM_FLAG_SWAP 2    /* stream the weight as the 128-row MMA operand (small-M / decode) */
#define RVL_GEMM_FLAG_STREAMK 4 /* with SWAP and a bound workspace: always deal the k-blocks evenly to the SMs (stream-K);
                                   by default the library does so only when plain tiles leave a ragged last wave */
#define RVL_GEMM_FLAG_SWIGLU 16 /* W rows are interleaved in blocks of 32 = [16 gate rows | 16 up rows] (rvl_weights.wgu_layout 1):
                                   the epilogue stores silu(gate) * up as bf16 [M, N / 2] (ldc counts act columns); needs
                                   N % 32 == 0, RVL_GEMM_OUT_BF16, no bias / ReLU / rowmap / split_k */
#define RVL_GEMM_FLAG_W_CONST 8 /* W is a bound weight that no work queued earlier on the stream writes: with SWAP the kernel
                                   (launched with programmatic stream serialization) may prefetch W into shared memory
                                   while the kernel that produces A is still running */

/* out[M,N] = act(A[M,K] . W[N,K]^T + bias[N]); A, W, bias bf16; fp32 accumulation in TMEM.
 * Replaces every nn.Linear on the path (cuBLAS via torch): q/k/v/o, gate/up/down, lm_head,
 * mm_projector and the ClipEncoder linears.  K % 8 == 0, N % 8 == 0.
 * rowmap (optional, int32 [M]) scatters output rows. split_k > 1 needs RVL_GEMM_ADD_F32. */
RVL_API int rvl_gemm_bf16(rvl_handle* h, const void* A, const void* W, const void* bias, void* out,
                  int64_t M, int64_t N, int64_t K, int64_t ldc, int32_t out_mode, int32_t flags,
                  const int32_t* rowmap, int32_t split_k, rvl_stream stream);

/* y(bf16) = x(fp32) * rsqrt(mean(x^2)+eps) * w(bf16); rows optional gather (int32 [n_rows]). */
RVL_API int rvl_rmsnorm(rvl_handle* h, const float* x, const void* w, void* y, int64_t n_rows, int32_t dim,
                float eps, const int32_t* rows, rvl_stream stream);

/* RoPE (rotate-half) on q,k of qkv [n_tokens, 3*hidden] bf16 in place + KV page write.
 * positions: explicit int32 [n_tokens] device array (decode: seq_lens), or NULL with tok_seq and
 * cu_seqlens given (prefill: position = index inside the sequence).  tok_seq int32 [n_tokens] =
 * sequence of each token (NULL: token i belongs to sequence i). */
RVL_API int rvl_rope_kv(rvl_handle* h, void* qkv, int64_t n_tokens, const int32_t* positions,
                const int32_t* tok_seq, const int32_t* cu_seqlens, const int32_t* page_table,
                int32_t max_pages, int32_t layer, rvl_stream stream);

/* act[n, I] = silu(gu[n, 0:I]) * gu[n, I:2I], bf16. */
RVL_API int rvl_swiglu(rvl_handle* h, const void* gu, void* act, int64_t n_tokens, int32_t intermediate,
               rvl_stream stream);

/* Causal varlen attention over qkv [total_tokens, 3*hidden] (post-RoPE) -> out [total_tokens, hidden] bf16.
 * tcgen05 kernel (TMA-staged Q / K / V tiles, S and O in TMEM, fp32 online softmax); replaces the eager attention of
 * transformers' Llama reached from vtimellm_llama.py:79-90. */
RVL_API int rvl_attn_prefill(rvl_handle* h, const void* qkv, void* out, const int32_t* cu_seqlens,
                     int32_t n_seq, int32_t max_seqlen, int64_t total_tokens, rvl_stream stream);

/* Paged-KV decode attention: q from qkv [n_seq, 3*hidden], keys 0..seq_lens[i] (inclusive of this
 * step's token) -> out [n_seq, hidden] bf16.
 *   fused_rope == 0: qkv is post-RoPE and the cache already holds this step's token (after rvl_rope_kv);
 *   fused_rope != 0: qkv holds the un-rotated q, k of this step; the kernel applies RoPE at position
 *                    seq_lens[i] and appends k', v to the cache page itself (what rvl_decode_step does - the
 *                    separate apply_rotary_pos_emb + DynamicCache.update kernels of the reference disappear). */
RVL_API int rvl_attn_decode(rvl_handle* h, const void* qkv, void* out, const int32_t* seq_lens, int32_t n_seq,
                    const int32_t* page_table, int32_t max_pages, int32_t layer, int32_t fused_rope,
                    int32_t max_kv_len /* as in rvl_decode_step */, rvl_stream stream);

/* ---- stage-2 adapter (ClipEncoder) ------------------------------------------------------------- */

/* One encoder layer of the adapter (nn.MultiheadAttention + FFN + two LayerNorms; transformer.py:188-305), all bf16. */
typedef struct rvl_clip_layer {
  const void* in_proj_w;   /* [3*768, 768]: q | k | v */
  const void* in_proj_b;   /* [3*768] */
  const void* out_proj_w;  /* [768, 768] */
  const void* out_proj_b;
  const void* linear1_w;   /* [2048, 768] */
  const void* linear1_b;
  const void* linear2_w;   /* [768, 2048] */
  const void* linear2_b;
  const void* norm1_w;
  const void* norm1_b;
  const void* norm2_w;
  const void* norm2_b;
} rvl_clip_layer;

typedef struct rvl_clip_weights {
  rvl_clip_layer t2v[2];        /* text -> video cross-attention layers (t2v_encoder.layers.{0,1}) */
  rvl_clip_layer enc[2];        /* self-attention layers over [global token ; frames] (encoder.layers.{0,1}) */
  const void* global_token;     /* bf16 [768]  (global_rep_token) */
  const float* pos;             /* fp32 [T, 768]: PositionEmbeddingSine(normalize=True) of T frames (transformer.py:35-57) */
  const float* pos_global;      /* fp32 [T + 1, 768]: global_rep_pos followed by `pos` */
  const void* proj_w;           /* bf16 [hidden, 768]  (mm_projector.weight of the adapter) */
  const void* proj_b;           /* bf16 [hidden] */
  int32_t hidden;
} rvl_clip_weights;

/* Bytes of scratch rvl_clip_encoder needs for V segments of T frames and Q texts of Lq tokens. */
RVL_API size_t rvl_clip_encoder_workspace_bytes(int32_t n_seg, int32_t n_frames, int32_t n_text, int32_t text_len);

/* The whole stage-2 adapter in one call.  Replaces ClipEncoder.forward (transformer.py:94-145: two text->video
 * cross-attention layers :271-305, prepend the global token, two post-norm self-attention layers :210-223 over 1 + T tokens,
 * CLS row -> Linear(768 -> hidden)) as reached from the hierarchy branch of vtimellm_arch.py:114-121.
 *   frames        [n_seg, n_frames, 768] bf16      text       [n_text, text_len, 768] bf16
 *   text_mask     [n_text, text_len] fp32, 1 = valid
 *   seg_text_idx  [n_seg] int32 or NULL (n_text == n_seg, one to one): the text each segment attends to - the reference
 *                 repeats the query once per segment (vtimellm_arch.py:116-119), here the keys / values of a text are
 *                 projected once
 *   ws            scratch of at least rvl_clip_encoder_workspace_bytes(...) bytes, 256-byte aligned
 *   out           [n_seg, hidden] bf16: one projected CLS row per segment */
RVL_API int rvl_clip_encoder(rvl_handle* h, const rvl_clip_weights* w, const void* frames, const void* text,
                     const float* text_mask, const int32_t* seg_text_idx, int32_t n_seg, int32_t n_frames,
                     int32_t n_text, int32_t text_len, void* ws, size_t ws_bytes, void* out, rvl_stream stream);

/* the kernels the adapter is made of (unit parity tests) */

/* LayerNorm over the last dim (<= 1024, biased variance, like nn.LayerNorm; transformer.py:199-200):
 * y = (x - mean) * rsqrt(var + eps) * w + b, or y = x when w == b == NULL.  Any of the outputs may be
 * NULL: y_f32 (may alias x), y_bf16, y_pos_bf16 = bf16(y + pos[row % period]) (`with_pos_embed`,
 * transformer.py:207-208).  x fp32 [rows, dim]; w, b bf16 [dim]; pos fp32 [period, dim]. */
RVL_API int rvl_layernorm(rvl_handle* h, const float* x, const void* w, const void* b, float* y_f32, void* y_bf16,
                  const float* pos, void* y_pos_bf16, int64_t rows, int32_t dim, int32_t period, float eps,
                  rvl_stream stream);

/* Multi-head attention with head_dim 96 and optional key padding (nn.MultiheadAttention core of
 * transformer.py:216-217 and :288-289): for sequence s and head a,
 *   out[s*Tq + t, a*96:(a+1)*96] = softmax_j(q[s*Tq + t] . k[kv(s)*Tk + j] / sqrt(96) + mask) v[kv(s)*Tk + j]
 * q/k/v/out are bf16 with explicit row strides (elements), so they can be column slices of fused
 * projections.  kv_seq_idx int32 [n_seq] (NULL: kv(s) = s) lets many query sequences share one key/value
 * sequence (the text of a query is shared by all its segments); n_kv_seq = number of key / value sequences
 * (rows of k and v = n_kv_seq * Tk; ignored when kv_seq_idx is NULL).  key_mask fp32 [n_kv_seq, Tk], 0 = padded.
 * Runs on tcgen05 (TMA-staged Q / K / V tiles, scores and output in TMEM) when q, k, v, out are 16-byte aligned and
 * out_stride is a multiple of 8; otherwise on an mma.sync kernel. */
RVL_API int rvl_mha96(rvl_handle* h, const void* q, int64_t q_stride, const void* k, int64_t k_stride, const void* v,
              int64_t v_stride, void* out, int64_t out_stride, int32_t n_seq, int32_t n_heads, int32_t Tq,
              int32_t Tk, const int32_t* kv_seq_idx, int32_t n_kv_seq, const float* key_mask, rvl_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* REVISIONLLM_B200_H */
