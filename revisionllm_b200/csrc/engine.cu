// C ABI + host-side orchestration of the decoder stack (layer loop, workspace carving, KV pages).
// Everything here only enqueues kernels on the caller's stream: no allocation, no synchronisation.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "rvl_internal.h"

using namespace rvl;

struct rvl_handle {
  rvl_config cfg{};
  int num_sms = 0;
  bool weights_bound = false;
  rvl_weights w{};
  std::vector<rvl_layer_weights> layers;
  // workspace
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  int64_t max_tokens = 0;
  int32_t max_seqs = 0;
  void *xnorm = nullptr, *qkv = nullptr, *attn = nullptr, *gu = nullptr, *act = nullptr, *xlast = nullptr;
  float* dec_hidden = nullptr;
  float* stream_ws = nullptr;        // stream-K partial tiles of the weight-streaming GEMMs: [num_sms][256][256] fp32
  size_t stream_ws_bytes = 0;
  unsigned int* stream_flags = nullptr;  // [num_sms] "partial published" flags (zero between launches)
  float* partials = nullptr;             // [kMaxSplit][max_seqs][hidden] fp32 split-k partials of the decode o/down GEMMs
  int32_t *tok_seq = nullptr, *last_rows = nullptr;
  // kv
  uint8_t* kv = nullptr;
  int32_t n_pages = 0;
  // optional per-launch timing (bench.py roofline): CUDA event pairs on the launching stream
  bool prof_on = false;
  struct ProfRec { int cat; double flops, bytes; };
  std::vector<cudaEvent_t> prof_ev;      // 2 per record
  std::vector<ProfRec> prof_rec;
  size_t prof_cap = 0;
  mutable std::string err;
};

namespace {
struct ProfScope {
  rvl_handle* h; cudaStream_t st; size_t idx; bool on;
  ProfScope(const rvl_handle* hc, cudaStream_t s, int cat, double flops, double bytes)
      : h(const_cast<rvl_handle*>(hc)), st(s), idx(0), on(false) {
    if (!h->prof_on || h->prof_rec.size() >= h->prof_cap) return;
    on = true;
    idx = h->prof_rec.size();
    h->prof_rec.push_back({cat, flops, bytes});
    cudaEventRecord(h->prof_ev[2 * idx], st);
  }
  ~ProfScope() { if (on) cudaEventRecord(h->prof_ev[2 * idx + 1], st); }
};
}  // namespace

static thread_local std::string g_err;

static int fail(const rvl_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_err = msg;
  return code;
}
namespace rvl {
static Tuning g_tuning;
static bool g_tuning_loaded = false;
void reload_tuning() {
  auto geti = [](const char* name, int unset) { const char* e = getenv(name); return e ? atoi(e) : unset; };
  Tuning t;
  t.pdl = geti("RVL_PDL", -1);
  t.a_tiles = geti("RVL_A_TILES", 0);
  t.stream_k = geti("RVL_STREAM_K", -1);
  t.plan_debug = getenv("RVL_PLAN_DEBUG") ? 1 : 0;
  t.pair = geti("RVL_PAIR", -1);
  t.spair = geti("RVL_SPAIR", -1);
  t.spair_streamk = geti("RVL_SPAIR_STREAMK", -1);
  t.staged = geti("RVL_STAGED", -1);
  t.group_m = geti("RVL_GROUP_M", 0);
  t.full_last_layer = geti("RVL_FULL_LAST_LAYER", 0);
  const char* ad = getenv("RVL_ATTN_DECODE");
  t.attn_decode = ad ? ad[0] : 0;
  t.attn_ps32 = geti("RVL_ATTN_PS32", -1);
  t.attn_prefill = geti("RVL_ATTN_PREFILL", -1);
  t.attn_mha96 = geti("RVL_ATTN_MHA96", -1);
  t.norm_threads = geti("RVL_NORM_THREADS", 0);
  g_tuning = t;
  g_tuning_loaded = true;
}
const Tuning& tuning() {
  if (!g_tuning_loaded) reload_tuning();
  return g_tuning;
}
// error reporting for the entry points that live in other translation units (clip_encoder.cu)
int report_error(const rvl_handle* h, int code, const char* msg) { return fail(h, code, msg); }
int handle_num_sms(const rvl_handle* h) { return h ? h->num_sms : 0; }
}  // namespace rvl
static int check_cuda(const rvl_handle* h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, RVL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return RVL_OK;
}
static inline size_t align_up(size_t x, size_t a = 1024) { return (x + a - 1) / a * a; }

constexpr size_t kStreamWsBytes = 160ull * 2 * 8 * 8 * 128 * 16;   // up to 160 SMs x 2 A tiles x 128 KB of fp32 partial
constexpr int kMaxSplit = 8;
struct WsLayout {
  size_t xnorm, qkv, attn, gu, act, xlast, dec_hidden, stream_ws, stream_flags, partials, tok_seq, last_rows, total;
};
static WsLayout ws_layout(const rvl_config& c, int64_t T, int32_t S) {
  WsLayout l{};
  const size_t H = c.hidden, I = c.intermediate;
  const size_t rows = static_cast<size_t>(T > S ? T : S);
  size_t off = 0;
  l.xnorm = off; off += align_up(rows * H * 2);
  l.qkv = off; off += align_up(rows * 3 * H * 2);
  l.attn = off; off += align_up(rows * H * 2);
  l.gu = off; off += align_up(rows * 2 * I * 2);
  l.act = off; off += align_up(rows * I * 2);
  l.xlast = off; off += align_up(static_cast<size_t>(S) * H * 2);
  l.dec_hidden = off; off += align_up(static_cast<size_t>(S) * H * 4);
  l.stream_ws = off; off += align_up(kStreamWsBytes);
  l.stream_flags = off; off += align_up(1024);
  l.partials = off; off += align_up(static_cast<size_t>(kMaxSplit) * S * H * 4);
  l.tok_seq = off; off += align_up(rows * 4);
  l.last_rows = off; off += align_up(static_cast<size_t>(S) * 4);
  l.total = off;
  return l;
}

extern "C" {

int rvl_abi_version(void) { return RVL_ABI_VERSION; }

void rvl_reload_env(void) { rvl::reload_tuning(); }

const char* rvl_last_error(const rvl_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int rvl_create(const rvl_config* cfg, rvl_handle** out) {
  if (!cfg || !out) return fail(nullptr, RVL_ERR_INVALID, "rvl_create: null argument");
  if (cfg->head_dim != 128) return fail(nullptr, RVL_ERR_INVALID, "rvl_create: head_dim must be 128");
  if (cfg->hidden != cfg->n_heads * cfg->head_dim) return fail(nullptr, RVL_ERR_INVALID, "rvl_create: hidden != n_heads*head_dim");
  if (cfg->hidden % 64 || cfg->intermediate % 64 || cfg->vocab % 8 || cfg->adapter_dim % 8 || cfg->hidden > 8192)
    return fail(nullptr, RVL_ERR_INVALID, "rvl_create: unsupported dimensions");
  if (cfg->kv_page_size <= 0 || cfg->kv_page_size > 256) return fail(nullptr, RVL_ERR_INVALID, "rvl_create: bad kv_page_size");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, RVL_ERR_CUDA, std::string("rvl_create: no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, RVL_ERR_INVALID, "rvl_create: bad device ordinal");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, RVL_ERR_CUDA, "rvl_create: cudaGetDeviceProperties failed");
  if (prop.major != 10) {
    char buf[128];
    snprintf(buf, sizeof buf, "rvl_create: device is sm_%d%d; these kernels are sm_100a only", prop.major, prop.minor);
    return fail(nullptr, RVL_ERR_UNSUPPORTED, buf);
  }
  if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(nullptr, RVL_ERR_CUDA, "rvl_create: cudaSetDevice failed");
  rvl::tuning();                       // the RVL_* switches are read here, once (rvl_reload_env() re-reads them)
  rvl_handle* h = new rvl_handle();
  h->cfg = *cfg;
  h->num_sms = prop.multiProcessorCount;
  *out = h;
  return RVL_OK;
}

void rvl_destroy(rvl_handle* h) { delete h; }

int rvl_bind_weights(rvl_handle* h, const rvl_weights* w) {
  if (!h || !w || !w->layers) return fail(h, RVL_ERR_INVALID, "rvl_bind_weights: null argument");
  if (!w->embed_tokens || !w->final_norm || !w->lm_head) return fail(h, RVL_ERR_INVALID, "rvl_bind_weights: missing tensor");
  if (w->wgu_layout != 0 && w->wgu_layout != 1) return fail(h, RVL_ERR_INVALID, "rvl_bind_weights: wgu_layout must be 0 or 1");
  if (w->wgu_layout == 1 && h->cfg.intermediate % 16) return fail(h, RVL_ERR_INVALID, "rvl_bind_weights: wgu_layout 1 needs intermediate % 16 == 0");
  h->layers.assign(w->layers, w->layers + h->cfg.n_layers);
  for (auto& l : h->layers)
    if (!l.wqkv || !l.wo || !l.wgu || !l.wdown || !l.ln1 || !l.ln2) return fail(h, RVL_ERR_INVALID, "rvl_bind_weights: missing layer tensor");
  h->w = *w;
  h->w.layers = h->layers.data();
  h->weights_bound = true;
  return RVL_OK;
}

size_t rvl_workspace_bytes(const rvl_handle* h, int64_t max_tokens, int32_t max_seqs) {
  if (!h || max_tokens <= 0 || max_seqs <= 0) return 0;
  return ws_layout(h->cfg, max_tokens, max_seqs).total;
}

int rvl_set_workspace(rvl_handle* h, void* ws, size_t bytes, int64_t max_tokens, int32_t max_seqs) {
  if (!h || !ws) return fail(h, RVL_ERR_INVALID, "rvl_set_workspace: null argument");
  const WsLayout l = ws_layout(h->cfg, max_tokens, max_seqs);
  if (bytes < l.total) return fail(h, RVL_ERR_INVALID, "rvl_set_workspace: buffer too small");
  if (reinterpret_cast<uintptr_t>(ws) & 255) return fail(h, RVL_ERR_INVALID, "rvl_set_workspace: buffer must be 256-byte aligned");
  uint8_t* b = static_cast<uint8_t*>(ws);
  h->ws = b; h->ws_bytes = bytes; h->max_tokens = max_tokens; h->max_seqs = max_seqs;
  h->xnorm = b + l.xnorm; h->qkv = b + l.qkv; h->attn = b + l.attn; h->gu = b + l.gu; h->act = b + l.act;
  h->xlast = b + l.xlast; h->dec_hidden = reinterpret_cast<float*>(b + l.dec_hidden);
  h->stream_ws = reinterpret_cast<float*>(b + l.stream_ws);
  h->stream_ws_bytes = kStreamWsBytes;
  h->stream_flags = reinterpret_cast<unsigned int*>(b + l.stream_flags);
  h->partials = reinterpret_cast<float*>(b + l.partials);
  if (cudaMemset(h->stream_flags, 0, 1024) != cudaSuccess) return fail(h, RVL_ERR_CUDA, "rvl_set_workspace: cudaMemset failed");
  h->tok_seq = reinterpret_cast<int32_t*>(b + l.tok_seq); h->last_rows = reinterpret_cast<int32_t*>(b + l.last_rows);
  return RVL_OK;
}

static size_t kv_layer_half_bytes(const rvl_config& c, int32_t n_pages) {
  return static_cast<size_t>(n_pages) * c.n_heads * c.kv_page_size * c.head_dim * 2;
}
size_t rvl_kv_bytes(const rvl_handle* h, int32_t n_pages) {
  if (!h || n_pages <= 0) return 0;
  return kv_layer_half_bytes(h->cfg, n_pages) * 2 * h->cfg.n_layers;
}
int rvl_set_kv(rvl_handle* h, void* kv, int32_t n_pages) {
  if (!h || !kv || n_pages <= 0) return fail(h, RVL_ERR_INVALID, "rvl_set_kv: bad argument");
  h->kv = static_cast<uint8_t*>(kv);
  h->n_pages = n_pages;
  return RVL_OK;
}
static void* k_pages(const rvl_handle* h, int layer) { return h->kv + kv_layer_half_bytes(h->cfg, h->n_pages) * (2 * layer); }
static void* v_pages(const rvl_handle* h, int layer) { return h->kv + kv_layer_half_bytes(h->cfg, h->n_pages) * (2 * layer + 1); }

// ------------------------------------------------------------------------------------------ GEMM helper
// partial_buf != null (decode o_proj / down_proj): the GEMM leaves `*split_used` fp32 partial sums
// [split][tokens][features] there and the RMSNorm that follows adds them to the residual stream.
static int linear(const rvl_handle* h, const void* x, const void* w, const void* bias, void* out, int64_t tokens,
                  int64_t features, int64_t K, int64_t ldc, int mode, int flags, const int32_t* rowmap, cudaStream_t st,
                  float* partial_buf = nullptr, int* split_used = nullptr) {
  GemmCall c;
  c.A = x; c.W = w; c.bias = bias; c.out = out; c.M = tokens; c.N = features; c.K = K; c.ldc = ldc;
  c.out_mode = mode; c.flags = flags; c.rowmap = rowmap; c.split_k = 1;
  if (tokens <= 256) {
    // few tokens (decode, lm_head on last rows): stream the weight through the 128-row MMA slot once
    c.flags |= RVL_GEMM_FLAG_SWAP | RVL_GEMM_FLAG_W_CONST;   // w is a bound weight: prefetchable before the producer of x ends
    if (partial_buf) {
      const int tiles = static_cast<int>((features + 127) / 128);
      const int kb = static_cast<int>((K + 63) / 64);
      int sk = h->num_sms / tiles;                       // fill the SMs once: 32 tiles x 4 splits
      if (sk > kMaxSplit) sk = kMaxSplit;
      while (sk > 1 && kb / sk < 8) --sk;
      if (sk < 1) sk = 1;
      c.split_k = sk;
      c.out = partial_buf; c.out_mode = RVL_GEMM_OUT_F32; c.ldc = features; c.split_stride = tokens * features;
      c.split_used = split_used;
    } else if (h->stream_ws) {
      c.stream_ws = h->stream_ws; c.stream_ws_bytes = h->stream_ws_bytes; c.stream_flags = h->stream_flags;
    }
  }
  std::string err;
  const double out_b = (mode == RVL_GEMM_OUT_BF16 ? 2.0 : (mode == RVL_GEMM_ADD_F32 ? 8.0 : 4.0)) * tokens * features;
  ProfScope ps(h, st, tokens <= 256 ? RVL_PROF_GEMM_SMALL_M : RVL_PROF_GEMM, 2.0 * tokens * features * K,
               2.0 * features * K + 2.0 * tokens * K + out_b);
  int rc = gemm_bf16(c, h->num_sms, st, &err);
  if (rc) return fail(h, rc, err);
  return RVL_OK;
}

// act = silu(x . Wgate^T) * (x . Wup^T): one GEMM with the SwiGLU in its epilogue when the weight is interleaved
// (wgu_layout 1), else GEMM -> [tokens, 2I] -> swiglu kernel.
static int gate_up(const rvl_handle* h, const rvl_layer_weights& w, int64_t tokens, cudaStream_t st) {
  const int H = h->cfg.hidden, I = h->cfg.intermediate;
  if (h->w.wgu_layout == 1)
    return linear(h, h->xnorm, w.wgu, nullptr, h->act, tokens, 2 * I, H, I, RVL_GEMM_OUT_BF16, RVL_GEMM_FLAG_SWIGLU, nullptr, st);
  int rc = linear(h, h->xnorm, w.wgu, nullptr, h->gu, tokens, 2 * I, H, 2 * I, RVL_GEMM_OUT_BF16, 0, nullptr, st);
  if (rc) return rc;
  launch_swiglu(h->gu, h->act, tokens, I, st);
  return RVL_OK;
}

int rvl_gemm_bf16(rvl_handle* h, const void* A, const void* W, const void* bias, void* out, int64_t M, int64_t N,
                  int64_t K, int64_t ldc, int32_t out_mode, int32_t flags, const int32_t* rowmap, int32_t split_k,
                  rvl_stream stream) {
  if (!h || !A || !W || !out) return fail(h, RVL_ERR_INVALID, "rvl_gemm_bf16: null argument");
  GemmCall c;
  c.A = A; c.W = W; c.bias = bias; c.out = out; c.M = M; c.N = N; c.K = K; c.ldc = ldc;
  c.out_mode = out_mode; c.flags = flags; c.rowmap = rowmap; c.split_k = split_k < 1 ? 1 : split_k;
  if ((flags & RVL_GEMM_FLAG_SWAP) && c.split_k == 1 && h->stream_ws) {
    c.stream_ws = h->stream_ws; c.stream_ws_bytes = h->stream_ws_bytes; c.stream_flags = h->stream_flags;
  }
  std::string err;
  int rc = gemm_bf16(c, h->num_sms, static_cast<cudaStream_t>(stream), &err);
  if (rc) return fail(h, rc, err);
  return check_cuda(h, "rvl_gemm_bf16");
}

// ------------------------------------------------------------------------------------------ splice
int rvl_project_splice(rvl_handle* h, const void* feats, const int32_t* feat_dst, int32_t n_feat_rows,
                       const int32_t* text_ids, const int32_t* text_dst, int32_t n_text, float* hidden_out,
                       int64_t total_tokens, rvl_stream stream) {
  if (!h || !hidden_out) return fail(h, RVL_ERR_INVALID, "rvl_project_splice: null argument");
  if (!h->weights_bound || !h->w.proj_w || !h->w.proj_b) return fail(h, RVL_ERR_STATE, "rvl_project_splice: projector weights not bound");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  (void)total_tokens;
  if (n_text > 0) launch_embed_rows(h->w.embed_tokens, text_ids, text_dst, n_text, h->cfg.hidden, h->cfg.vocab, hidden_out, st);
  if (n_feat_rows > 0) {
    // the epilogue adds the bias and writes each projected frame straight to its spliced row
    GemmCall c;
    c.A = feats; c.W = h->w.proj_w; c.bias = h->w.proj_b; c.out = hidden_out;
    c.M = n_feat_rows; c.N = h->cfg.hidden; c.K = h->cfg.adapter_dim; c.ldc = h->cfg.hidden;
    c.out_mode = RVL_GEMM_OUT_F32; c.rowmap = feat_dst;
    if (n_feat_rows <= 256) c.flags |= RVL_GEMM_FLAG_SWAP;
    std::string err;
    int rc = gemm_bf16(c, h->num_sms, st, &err);
    if (rc) return fail(h, rc, err);
  }
  return check_cuda(h, "rvl_project_splice");
}

int rvl_splice_rows(rvl_handle* h, const void* vis, const int32_t* vis_dst, int32_t n_vis, const int32_t* text_ids,
                    const int32_t* text_dst, int32_t n_text, float* hidden_out, int64_t total_tokens, rvl_stream stream) {
  if (!h || !hidden_out) return fail(h, RVL_ERR_INVALID, "rvl_splice_rows: null argument");
  if (!h->weights_bound) return fail(h, RVL_ERR_STATE, "rvl_splice_rows: weights not bound");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  (void)total_tokens;
  if (n_text > 0) launch_embed_rows(h->w.embed_tokens, text_ids, text_dst, n_text, h->cfg.hidden, h->cfg.vocab, hidden_out, st);
  if (n_vis > 0) launch_scatter_rows_bf16(vis, vis_dst, n_vis, h->cfg.hidden, hidden_out, st);
  return check_cuda(h, "rvl_splice_rows");
}

// ------------------------------------------------------------------------------------------ decoder stack
static int ready(const rvl_handle* h, const char* who) {
  if (!h) return fail(h, RVL_ERR_INVALID, std::string(who) + ": null handle");
  if (!h->weights_bound) return fail(h, RVL_ERR_STATE, std::string(who) + ": weights not bound");
  if (!h->ws) return fail(h, RVL_ERR_STATE, std::string(who) + ": workspace not set");
  if (!h->kv) return fail(h, RVL_ERR_STATE, std::string(who) + ": kv pages not set");
  return RVL_OK;
}

int rvl_prefill(rvl_handle* h, float* hidden, const int32_t* cu_seqlens, int32_t n_seq, int64_t total_tokens,
                int32_t max_seqlen, const int32_t* page_table, int32_t max_pages, float* logits_out, int32_t all_logits,
                const int32_t* seq_pos0, const int32_t* seq_ctx_row, rvl_stream stream) {
  int rc = ready(h, "rvl_prefill");
  if (rc) return rc;
  if (!hidden || !cu_seqlens || !page_table || !logits_out) return fail(h, RVL_ERR_INVALID, "rvl_prefill: null argument");
  if (total_tokens > h->max_tokens || n_seq > h->max_seqs || n_seq <= 0 || total_tokens <= 0)
    return fail(h, RVL_ERR_INVALID, "rvl_prefill: batch exceeds the workspace");
  if (max_seqlen > max_pages * h->cfg.kv_page_size) return fail(h, RVL_ERR_INVALID, "rvl_prefill: page table too small for max_seqlen");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const rvl_config& c = h->cfg;
  const int64_t T = total_tokens;
  const int H = c.hidden, I = c.intermediate;
  launch_token_seq(cu_seqlens, n_seq, h->tok_seq, h->last_rows, T, st);
  const bool env_full_last = tuning().full_last_layer == 1;   // diagnostic: run the last layer on every row
  for (int l = 0; l < c.n_layers; ++l) {
    const rvl_layer_weights& w = h->layers[l];
    launch_rmsnorm(hidden, w.ln1, h->xnorm, T, H, c.rms_eps, nullptr, st);
    if ((rc = linear(h, h->xnorm, w.wqkv, nullptr, h->qkv, T, 3 * H, H, 3 * H, RVL_GEMM_OUT_BF16, 0, nullptr, st))) return rc;
    launch_rope_kv(h->qkv, T, nullptr, h->tok_seq, cu_seqlens, page_table, max_pages, k_pages(h, l), v_pages(h, l),
                   c.n_heads, c.kv_page_size, c.rope_theta, st, seq_pos0);
    // Generation prefill, last layer: every position's K / V is cached above, but only the LAST position's hidden state is
    // ever read (final norm + lm_head on the last rows), so attention runs for the query tile that holds it and the o
    // projection and the MLP run on n_seq gathered rows instead of T (1/32 of the o / gate|up / down FLOPs of a 7B prefill).
    const bool last_only = !all_logits && l == c.n_layers - 1 && H % 8 == 0 && !env_full_last;
    {
      // causal FLOPs need the per-sequence lengths (device side); use the uniform-length bound T*max_seqlen
      ProfScope ps(h, st, RVL_PROF_ATTN_PREFILL, 2.0 * T * max_seqlen * H, 2.0 * T * 4 * H);
      // an external context must end on a 32-key boundary for the tcgen05 kernel (it loads 32-key boxes); contexts are whole
      // KV pages, so any page size that is a multiple of 32 guarantees it - otherwise total_tokens = 0 selects the mma.sync kernel
      const bool ctx_ok = !seq_pos0 || c.kv_page_size % 32 == 0;
      launch_attn_prefill(h->qkv, h->attn, cu_seqlens, n_seq, max_seqlen, c.n_heads, st, seq_pos0, seq_ctx_row, last_only ? 1 : 0,
                          ctx_ok ? T : 0, h->num_sms);
    }
    if (last_only) {
      const int64_t n = n_seq;
      float* hid = h->dec_hidden;
      float* part = n <= 256 ? h->partials : nullptr;       // split-k partial buffers exist for the weight-streaming orientation
      int pending = 0;
      launch_gather_last_rows(h->attn, hidden, h->last_rows, n_seq, H, h->xlast, hid, st);
      if ((rc = linear(h, h->xlast, w.wo, nullptr, hid, n, H, H, H, RVL_GEMM_ADD_F32, 0, nullptr, st, part, &pending))) return rc;
      launch_rmsnorm(hid, w.ln2, h->xnorm, n, H, c.rms_eps, nullptr, st, h->partials, pending, n * H, pending ? hid : nullptr);
      if ((rc = gate_up(h, w, n, st))) return rc;
      if ((rc = linear(h, h->act, w.wdown, nullptr, hid, n, H, I, H, RVL_GEMM_ADD_F32, 0, nullptr, st, part, &pending))) return rc;
      launch_rmsnorm(hid, h->w.final_norm, h->xlast, n, H, c.rms_eps, nullptr, st, h->partials, pending, n * H, pending ? hid : nullptr);
      if ((rc = linear(h, h->xlast, h->w.lm_head, nullptr, logits_out, n, c.vocab, H, c.vocab, RVL_GEMM_OUT_F32, 0, nullptr, st))) return rc;
      return check_cuda(h, "rvl_prefill");
    }
    if ((rc = linear(h, h->attn, w.wo, nullptr, hidden, T, H, H, H, RVL_GEMM_ADD_F32, 0, nullptr, st))) return rc;
    launch_rmsnorm(hidden, w.ln2, h->xnorm, T, H, c.rms_eps, nullptr, st);
    if ((rc = gate_up(h, w, T, st))) return rc;
    if ((rc = linear(h, h->act, w.wdown, nullptr, hidden, T, H, I, H, RVL_GEMM_ADD_F32, 0, nullptr, st))) return rc;
  }
  if (all_logits) {
    // LlamaForCausalLM.forward semantics: logits for every position (fp32)
    launch_rmsnorm(hidden, h->w.final_norm, h->xnorm, T, H, c.rms_eps, nullptr, st);
    if ((rc = linear(h, h->xnorm, h->w.lm_head, nullptr, logits_out, T, c.vocab, H, c.vocab, RVL_GEMM_OUT_F32, 0, nullptr, st))) return rc;
  } else {
    // generation only needs the last row of each sequence: gather -> norm -> lm_head
    launch_rmsnorm(hidden, h->w.final_norm, h->xlast, n_seq, H, c.rms_eps, h->last_rows, st);
    if ((rc = linear(h, h->xlast, h->w.lm_head, nullptr, logits_out, n_seq, c.vocab, H, c.vocab, RVL_GEMM_OUT_F32, 0, nullptr, st))) return rc;
  }
  return check_cuda(h, "rvl_prefill");
}

__global__ void inc_kernel(int32_t* v, int n) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] += 1;
}

int rvl_decode_step(rvl_handle* h, const int32_t* token_ids, int32_t* seq_lens, int32_t n_seq,
                    const int32_t* page_table, int32_t max_pages, int32_t max_kv_len, float* logits_out, rvl_stream stream) {
  int rc = ready(h, "rvl_decode_step");
  if (rc) return rc;
  if (!token_ids || !seq_lens || !page_table || !logits_out) return fail(h, RVL_ERR_INVALID, "rvl_decode_step: null argument");
  if (n_seq <= 0 || n_seq > h->max_seqs) return fail(h, RVL_ERR_INVALID, "rvl_decode_step: n_seq exceeds the workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const rvl_config& c = h->cfg;
  const int H = c.hidden, I = c.intermediate;
  const int64_t n = n_seq;
  float* hid = h->dec_hidden;
  launch_embed_rows(h->w.embed_tokens, token_ids, nullptr, n_seq, H, c.vocab, hid, st);
  // o_proj / down_proj leave split-k partial sums in h->partials; the next RMSNorm adds them to the residual
  const int64_t pstride = n * H;
  int pending = 0;
  for (int l = 0; l < c.n_layers; ++l) {
    const rvl_layer_weights& w = h->layers[l];
    launch_rmsnorm(hid, w.ln1, h->xnorm, n, H, c.rms_eps, nullptr, st, h->partials, pending, pstride, pending ? hid : nullptr);
    if ((rc = linear(h, h->xnorm, w.wqkv, nullptr, h->qkv, n, 3 * H, H, 3 * H, RVL_GEMM_OUT_BF16, 0, nullptr, st))) return rc;
    {
      // RoPE of this step's q / k and the KV append happen inside the attention kernel
      ProfScope ps(h, st, RVL_PROF_ATTN_DECODE, 0.0, 0.0);   // bytes depend on seq_lens: the caller supplies them
      launch_attn_decode(h->qkv, h->attn, seq_lens, n_seq, page_table, max_pages, k_pages(h, l), v_pages(h, l), c.n_heads,
                         c.kv_page_size, 1, c.rope_theta, max_kv_len, st);
    }
    if ((rc = linear(h, h->attn, w.wo, nullptr, hid, n, H, H, H, RVL_GEMM_ADD_F32, 0, nullptr, st, h->partials, &pending))) return rc;
    launch_rmsnorm(hid, w.ln2, h->xnorm, n, H, c.rms_eps, nullptr, st, h->partials, pending, pstride, hid);
    if ((rc = gate_up(h, w, n, st))) return rc;
    if ((rc = linear(h, h->act, w.wdown, nullptr, hid, n, H, I, H, RVL_GEMM_ADD_F32, 0, nullptr, st, h->partials, &pending))) return rc;
  }
  launch_rmsnorm(hid, h->w.final_norm, h->xlast, n, H, c.rms_eps, nullptr, st, h->partials, pending, pstride, hid);
  if ((rc = linear(h, h->xlast, h->w.lm_head, nullptr, logits_out, n, c.vocab, H, c.vocab, RVL_GEMM_OUT_F32, 0, nullptr, st))) return rc;
  launch_k(inc_kernel, dim3((n_seq + 127) / 128), dim3(128), 0, st, seq_lens, static_cast<int>(n_seq));
  return check_cuda(h, "rvl_decode_step");
}

int rvl_sample_greedy(rvl_handle* h, const float* logits, int32_t n_seq, int32_t vocab, int32_t* unfinished,
                      int32_t eos_id, int32_t pad_id, int32_t* next_tokens, float* entropy_out, rvl_stream stream) {
  if (!h || !logits || !next_tokens) return fail(h, RVL_ERR_INVALID, "rvl_sample_greedy: null argument");
  launch_sample_greedy(logits, n_seq, vocab, unfinished, eos_id, pad_id, next_tokens, entropy_out, nullptr, nullptr,
                       static_cast<cudaStream_t>(stream));
  return check_cuda(h, "rvl_sample_greedy");
}

int rvl_decode_n(rvl_handle* h, int32_t n_steps, float* logits, int32_t* token_ring, float* entropy_ring, int32_t* unfinished,
                 int32_t eos_id, int32_t pad_id, int32_t* seq_lens, int32_t n_seq, const int32_t* page_table, int32_t max_pages,
                 int32_t max_kv_len, rvl_stream stream) {
  int rc = ready(h, "rvl_decode_n");
  if (rc) return rc;
  if (!logits || !token_ring || !seq_lens || !page_table || n_steps <= 0) return fail(h, RVL_ERR_INVALID, "rvl_decode_n: bad argument");
  for (int s = 0; s < n_steps; ++s) {
    int32_t* tok = token_ring + static_cast<size_t>(s) * n_seq;
    launch_sample_greedy(logits, n_seq, h->cfg.vocab, unfinished, eos_id, pad_id, tok,
                         entropy_ring ? entropy_ring + static_cast<size_t>(s) * n_seq : nullptr, nullptr, nullptr,
                         static_cast<cudaStream_t>(stream));
    if ((rc = rvl_decode_step(h, tok, seq_lens, n_seq, page_table, max_pages, max_kv_len, logits, stream))) return rc;
  }
  return RVL_OK;
}

int rvl_gather_windows(rvl_handle* h, const float* features, int32_t n_frames, int32_t dim, const int32_t* frame_idx, int32_t n_rows,
                       void* out, rvl_stream stream) {
  if (!h || !features || !frame_idx || !out) return fail(h, RVL_ERR_INVALID, "rvl_gather_windows: null argument");
  if (n_frames <= 0 || dim % 4) return fail(h, RVL_ERR_INVALID, "rvl_gather_windows: need n_frames > 0 and dim % 4 == 0");
  launch_gather_rows_f32_bf16(features, frame_idx, n_rows, n_frames, dim, out, static_cast<cudaStream_t>(stream));
  return check_cuda(h, "rvl_gather_windows");
}

int rvl_sample_multinomial(rvl_handle* h, const float* logits, int32_t n_seq, int32_t vocab, float temperature, uint64_t seed,
                            uint32_t step, int32_t* unfinished, int32_t eos_id, int32_t pad_id, int32_t* next_tokens,
                            float* entropy_out, uint32_t* philox_out, rvl_stream stream) {
  if (!h || !logits || !next_tokens) return fail(h, RVL_ERR_INVALID, "rvl_sample_multinomial: null argument");
  if (!(temperature > 0.f)) return fail(h, RVL_ERR_INVALID, "rvl_sample_multinomial: temperature must be > 0 (use rvl_sample_greedy for argmax)");
  launch_sample_multinomial(logits, n_seq, vocab, temperature, seed, step, unfinished, eos_id, pad_id, next_tokens, entropy_out,
                            philox_out, static_cast<cudaStream_t>(stream));
  return check_cuda(h, "rvl_sample_multinomial");
}

int rvl_cosine_topk(rvl_handle* h, const void* frames, const int32_t* seg_offsets, const int32_t* seg_ends, int32_t n_seg,
                    int32_t dim, const void* cls, int32_t k, int32_t norm_axis, int32_t max_seg_rows, float* scores_out,
                    int32_t* topk_idx_out, rvl_stream stream) {
  if (!h || !frames || !seg_offsets || !cls || !scores_out) return fail(h, RVL_ERR_INVALID, "rvl_cosine_topk: null argument");
  if (dim % 8 || k < 1 || k > 16 || (norm_axis < 0 || norm_axis > 2) || max_seg_rows < 1 || max_seg_rows > 8192)
    return fail(h, RVL_ERR_INVALID, "rvl_cosine_topk: need dim % 8 == 0, 1 <= k <= 16, norm_axis in {0,1,2}, max_seg_rows <= 8192");
  launch_cosine_topk(frames, seg_offsets, seg_ends, n_seg, dim, cls, k, norm_axis, max_seg_rows, scores_out, topk_idx_out,
                     static_cast<cudaStream_t>(stream));
  return check_cuda(h, "rvl_cosine_topk");
}

int rvl_select_topk(rvl_handle* h, const float* scores, int32_t n, int32_t k, int32_t* idx_out, rvl_stream stream) {
  if (!h || !scores || !idx_out) return fail(h, RVL_ERR_INVALID, "rvl_select_topk: null argument");
  if (n > 65536 || k > n) return fail(h, RVL_ERR_INVALID, "rvl_select_topk: need k <= n <= 65536");
  launch_select_topk(scores, n, k, idx_out, static_cast<cudaStream_t>(stream));
  return check_cuda(h, "rvl_select_topk");
}

int rvl_merge_rank(rvl_handle* h, const float* cos, const float* ent, const int32_t* keep, const int32_t* cover1,
                   const int32_t* cover_all, int32_t n, int32_t mode, int32_t normalize, int32_t minmax, double* scores_out,
                   int32_t* order_out, int32_t* n_out, rvl_stream stream) {
  if (!h || !keep || !scores_out || !order_out || !n_out) return fail(h, RVL_ERR_INVALID, "rvl_merge_rank: null argument");
  if (n <= 0 || n > 8192 || mode < 0 || mode > 3) return fail(h, RVL_ERR_INVALID, "rvl_merge_rank: need 0 < n <= 8192 and mode in 0..3");
  if ((mode <= 1 && (!cos || !ent)) || (mode == 2 && !ent) || (mode == 3 && !cos))
    return fail(h, RVL_ERR_INVALID, "rvl_merge_rank: the score arrays the mode reads are missing");
  launch_merge_rank(cos, ent, keep, cover1, cover_all, n, mode, normalize, minmax, scores_out, order_out, n_out,
                    static_cast<cudaStream_t>(stream));
  return check_cuda(h, "rvl_merge_rank");
}

// ------------------------------------------------------------------------------------------ profiling
int rvl_profile_enable(rvl_handle* h, int32_t on, int32_t capacity) {
  if (!h) return fail(h, RVL_ERR_INVALID, "rvl_profile_enable: null handle");
  if (on) {
    const size_t cap = capacity > 0 ? capacity : 16384;
    while (h->prof_ev.size() < 2 * cap) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return fail(h, RVL_ERR_CUDA, "rvl_profile_enable: cudaEventCreate failed");
      h->prof_ev.push_back(e);
    }
    h->prof_cap = cap;
    h->prof_rec.clear();
    h->prof_rec.reserve(cap);
  }
  h->prof_on = on != 0;
  return RVL_OK;
}

int rvl_profile_read(rvl_handle* h, int32_t category, double* total_ms, double* total_flops, double* total_bytes,
                     int64_t* launches) {
  if (!h || !total_ms || !total_flops || !total_bytes || !launches) return fail(h, RVL_ERR_INVALID, "rvl_profile_read: null argument");
  double ms = 0, fl = 0, by = 0;
  int64_t n = 0;
  for (size_t i = 0; i < h->prof_rec.size(); ++i) {
    if (h->prof_rec[i].cat != category) continue;
    if (cudaEventSynchronize(h->prof_ev[2 * i + 1]) != cudaSuccess) return fail(h, RVL_ERR_CUDA, "rvl_profile_read: event sync failed");
    float t = 0.f;
    if (cudaEventElapsedTime(&t, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]) != cudaSuccess) return fail(h, RVL_ERR_CUDA, "rvl_profile_read: elapsed failed");
    ms += t; fl += h->prof_rec[i].flops; by += h->prof_rec[i].bytes; ++n;
  }
  *total_ms = ms; *total_flops = fl; *total_bytes = by; *launches = n;
  return RVL_OK;
}

// tools/ only: SM clock seen by a kernel at this point of the stream (cycles and nanoseconds of a ~20 us spin on one thread)
__global__ void sm_clock_probe_kernel(unsigned long long* out) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  const long long c0 = clock64();
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < 20000ull);
  out[0] = static_cast<unsigned long long>(clock64() - c0);
  out[1] = t1 - t0;
}
void rvl_debug_sm_clock(unsigned long long* out_dev, rvl_stream stream) {
  if (out_dev) sm_clock_probe_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(out_dev);
}

// ------------------------------------------------------------------------------------------ single kernels
int rvl_rmsnorm(rvl_handle* h, const float* x, const void* w, void* y, int64_t n_rows, int32_t dim, float eps,
                const int32_t* rows, rvl_stream stream) {
  if (!h || !x || !w || !y) return fail(h, RVL_ERR_INVALID, "rvl_rmsnorm: null argument");
  if (dim % 4 || dim > 8192) return fail(h, RVL_ERR_INVALID, "rvl_rmsnorm: dim must be a multiple of 4 and <= 8192");
  launch_rmsnorm(x, w, y, n_rows, dim, eps, rows, static_cast<cudaStream_t>(stream));
  return check_cuda(h, "rvl_rmsnorm");
}

int rvl_rope_kv(rvl_handle* h, void* qkv, int64_t n_tokens, const int32_t* positions, const int32_t* tok_seq,
                const int32_t* cu_seqlens, const int32_t* page_table, int32_t max_pages, int32_t layer, rvl_stream stream) {
  if (!h || !qkv || !page_table) return fail(h, RVL_ERR_INVALID, "rvl_rope_kv: null argument");
  if (!h->kv) return fail(h, RVL_ERR_STATE, "rvl_rope_kv: kv pages not set");
  if (!positions && !(tok_seq && cu_seqlens)) return fail(h, RVL_ERR_INVALID, "rvl_rope_kv: need positions or (tok_seq, cu_seqlens)");
  if (layer < 0 || layer >= h->cfg.n_layers) return fail(h, RVL_ERR_INVALID, "rvl_rope_kv: bad layer");
  launch_rope_kv(qkv, n_tokens, positions, tok_seq, cu_seqlens, page_table, max_pages, k_pages(h, layer), v_pages(h, layer),
                 h->cfg.n_heads, h->cfg.kv_page_size, h->cfg.rope_theta, static_cast<cudaStream_t>(stream));
  return check_cuda(h, "rvl_rope_kv");
}

int rvl_swiglu(rvl_handle* h, const void* gu, void* act, int64_t n_tokens, int32_t intermediate, rvl_stream stream) {
  if (!h || !gu || !act) return fail(h, RVL_ERR_INVALID, "rvl_swiglu: null argument");
  if (intermediate % 8) return fail(h, RVL_ERR_INVALID, "rvl_swiglu: intermediate must be a multiple of 8");
  launch_swiglu(gu, act, n_tokens, intermediate, static_cast<cudaStream_t>(stream));
  return check_cuda(h, "rvl_swiglu");
}

int rvl_attn_prefill(rvl_handle* h, const void* qkv, void* out, const int32_t* cu_seqlens, int32_t n_seq,
                     int32_t max_seqlen, int64_t total_tokens, rvl_stream stream) {
  if (!h || !qkv || !out || !cu_seqlens) return fail(h, RVL_ERR_INVALID, "rvl_attn_prefill: null argument");
  launch_attn_prefill(qkv, out, cu_seqlens, n_seq, max_seqlen, h->cfg.n_heads, static_cast<cudaStream_t>(stream), nullptr, nullptr, 0,
                      total_tokens, h->num_sms);
  return check_cuda(h, "rvl_attn_prefill");
}

int rvl_attn_decode(rvl_handle* h, const void* qkv, void* out, const int32_t* seq_lens, int32_t n_seq,
                    const int32_t* page_table, int32_t max_pages, int32_t layer, int32_t fused_rope, int32_t max_kv_len,
                    rvl_stream stream) {
  if (!h || !qkv || !out || !seq_lens || !page_table) return fail(h, RVL_ERR_INVALID, "rvl_attn_decode: null argument");
  if (!h->kv) return fail(h, RVL_ERR_STATE, "rvl_attn_decode: kv pages not set");
  if (layer < 0 || layer >= h->cfg.n_layers) return fail(h, RVL_ERR_INVALID, "rvl_attn_decode: bad layer");
  launch_attn_decode(qkv, out, seq_lens, n_seq, page_table, max_pages, k_pages(h, layer), v_pages(h, layer), h->cfg.n_heads,
                     h->cfg.kv_page_size, fused_rope, h->cfg.rope_theta, max_kv_len, static_cast<cudaStream_t>(stream));
  return check_cuda(h, "rvl_attn_decode");
}

}  // extern "C"
