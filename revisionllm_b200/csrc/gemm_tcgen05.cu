// bf16 GEMM on the 5th-gen tensor cores:  D[m][n] = sum_k A[m][k] * B[n][k]   (both operands K-major).
//
// Replaces every nn.Linear on the ReVisionLLM scoring path (cuBLAS through torch in the reference):
// q/k/v/o, gate/up/down, lm_head (transformers Llama reached from
// revisionllm/model/vtimellm_llama.py:79-90), mm_projector (revisionllm/model/vtimellm_arch.py:42,125)
// and the ClipEncoder linears (revisionllm/model/adapter/transformer.py).
//
// Design (one CTA per SM, persistent, warp-specialised, 192 threads):
//   warp 0 : TMA producer  - per pipeline stage one cp.async.bulk.tensor 2D load of an (a_tiles x 128) x 64 A box
//                            and one of a BN x 64 B box (128-byte swizzle), mbarrier complete_tx
//   warp 1 : MMA issuer    - tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, four K=16 steps per stage and A tile,
//                            fp32 accumulators in TMEM
//   warps 2..5 : epilogue  - tcgen05.ld 32x32b.x32 -> registers -> bias / ReLU / residual / row scatter -> global
//   smem ring full/empty mbarriers (TMA <-> MMA) and tmem full/empty mbarriers (MMA <-> epilogue).
// Whole warps run the role loops and one elected lane issues the TMA / MMA instructions, so the loop state stays
// in uniform registers (an `if (lane == 0)` role body makes ptxas serialise every uniform-datapath instruction).
//
// Measured on B200 (tools/probes/tma_probe.cu, profiles/): one SM's TMA path delivers ~55-72 B/clk from L2 and
// ~25 B/clk from HBM whatever the pipeline depth, so
//   * a 128x256 tile (48 KB per k-block) is TMA-bound (620 clk per k-block against 512 clk of MMA).  With
//     a_tiles = 2 the CTA owns a 256 x 256 tile: two A tiles share each B tile, 64 KB feed 1024 clk of MMA;
//   * weight streaming (few tokens: decode, last-row lm_head) only reaches HBM speed when every SM streams full
//     128-row boxes all the time.  The host swaps operand roles so the weight is the A operand, and the k-blocks of
//     all tiles are dealt out evenly to the CTAs (stream-K); a CTA that starts in the middle of a tile leaves an
//     fp32 partial in a workspace and the CTA that owns the head of that tile adds it in a fixed order
//     (deterministic, no atomics).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "rvl_internal.h"
#include "rvl_ptx.cuh"

namespace rvl {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kGemmThreads = 192;
constexpr int kMaxStages = 12;
constexpr int kATileBytes = kBM * kBK * 2;  // 16 KB per 128-row A tile and k-block
constexpr int kEpiScratchBytes = 4 * 32 * 33 * 4;   // per epilogue warp a 32 x 33 fp32 transpose tile
constexpr int kSmemBudget = 227 * 1024 - 1024 /*align slack*/ - 512 /*barriers*/ - kEpiScratchBytes;
// Weight-streaming (transposed) epilogue: output tiles are staged in shared memory as [token][feature] and written with
// TMA stores, one 32-token box per accumulator chunk; the staging buffer replaces the transpose scratch.
constexpr int kStageOutBytes = 48 * 1024;
constexpr int kSmemBudgetStaged = 227 * 1024 - 1024 - 512 - kStageOutBytes;

struct GemmArgs {
  int M, N, K;          // A rows, B rows, reduction
  int tiles_m, tiles_n; // tiles_m counts (a_tiles x 128)-row tiles
  int k_blocks;         // ceil(K / 64)
  int group_m;          // tile rasterisation: this many m-tiles share each n-tile sweep (their A slab stays in L2)
  int split_k;          // tile mode: k-range split (atomics into fp32 out, or partial buffers via split_stride)
  int a_tiles;          // 1 or 2: 128-row A tiles per CTA tile (they share the B tile)
  int bn;               // B rows per tile = MMA N (multiple of 16, <= 256)
  int stages;           // smem ring depth
  int acc_stages;       // TMEM accumulator stages (2 = epilogue overlaps the next tile's MMAs)
  int sub_stride;       // TMEM columns per A tile accumulator
  int tmem_cols;        // power of two >= acc_stages * a_tiles * sub_stride
  int stream_a;         // A is a weight streamed once from HBM: L2 evict-first for A, evict-last for B
  int stream_k;         // deal k-blocks of all tiles evenly to the CTAs (weight streaming)
  int prefetch_a;       // A is a bound weight no earlier kernel writes: fill the smem ring with A before pdl_wait()
  int units_per_cta;    // stream-K: k-blocks per CTA
  float* ws;            // stream-K: per-CTA fp32 partial [a_tiles*128][ws_ld]
  unsigned int* flags;  // stream-K: per-CTA "partial published" flag; the head CTA that consumed the partial clears it again,
                        // so the flags are zero between launches and a captured CUDA graph can be replayed unchanged
  int ws_ld;
  long long split_stride;  // > 0: split-k partial s goes to out + s * split_stride (plain stores, no atomics)
  long long ldc;
  void* out;
  const __nv_bfloat16* bias;  // indexed by the feature dimension
  const int* rowmap;          // indexed by the token dimension (optional)
  int mode;                   // RVL_GEMM_OUT_*
  int relu;
  int transposed;             // 0: tokens = A rows (m), features = B rows (n); 1: the other way round
  int atomic;                 // split-k partial sums: atomicAdd into fp32 out
  int staged;                 // transposed epilogue through shared-memory staging + TMA stores (tmap_out); chunks per group
  int swiglu;                 // feature rows come in blocks of [16 gate | 16 up]: store silu(gate) * up, bf16, [tokens][features / 2]
};

// One unit of work of a CTA: k-blocks [kb0, kb1) of tile (m_blk, n_blk).
struct WorkItem {
  int m_blk, n_blk, kb0, kb1, ks;
  int kind;  // 0 = whole k-range handled here (normal epilogue); 1 = stream-K tail piece: write the partial;
             // 2 = stream-K head piece: add `followers` partials, then normal epilogue
  int followers;
};

__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int group_m, int& m_blk, int& n_blk) {
  const int group = group_m * tiles_n;
  const int gid = t / group;
  const int first_m = gid * group_m;
  const int gsz = min(tiles_m - first_m, group_m);
  const int r = t - gid * group;
  m_blk = first_m + r % gsz;
  n_blk = r / gsz;
}

// idx-th work item of this CTA; false when there is none.  All three roles call it with the same arguments.
template <bool kStreamK>
__device__ __forceinline__ bool get_work(const GemmArgs& a, int idx, WorkItem& w, int vid = -1) {
  // vid: the CTA's index among the CTAs that share the work (blockIdx.x, or the cluster index for the CTA-pair kernels)
  if (vid < 0) vid = static_cast<int>(blockIdx.x);
  const int tiles_mn = a.tiles_m * a.tiles_n;
  if (!kStreamK) {
    const int tile = blockIdx.x + idx * gridDim.x;
    if (tile >= tiles_mn * a.split_k) return false;
    const int kb_per_split = (a.k_blocks + a.split_k - 1) / a.split_k;
    w.ks = tile / tiles_mn;
    tile_coords(tile - w.ks * tiles_mn, a.tiles_m, a.tiles_n, a.group_m, w.m_blk, w.n_blk);
    w.kb0 = w.ks * kb_per_split;
    w.kb1 = min(a.k_blocks, w.kb0 + kb_per_split);
    w.kind = 0;
    w.followers = 0;
    return true;
  }
  // stream-K: units are (tile, k-block) in tile-major order; this CTA owns [u0, u1)
  const long long total = static_cast<long long>(tiles_mn) * a.k_blocks;
  const long long u0 = static_cast<long long>(vid) * a.units_per_cta;
  const long long u1 = min(total, u0 + a.units_per_cta);
  if (u0 >= u1) return false;
  const int t_first = static_cast<int>(u0 / a.k_blocks);
  const int tile = t_first + idx;
  const long long t_begin = static_cast<long long>(tile) * a.k_blocks;
  if (t_begin >= u1) return false;
  w.ks = 0;
  tile_coords(tile, a.tiles_m, a.tiles_n, a.group_m, w.m_blk, w.n_blk);
  w.kb0 = static_cast<int>(max(u0, t_begin) - t_begin);
  w.kb1 = static_cast<int>(min(u1, t_begin + a.k_blocks) - t_begin);
  w.followers = 0;
  if (w.kb0 > 0) {
    w.kind = 1;  // the head of this tile belongs to an earlier CTA
  } else if (w.kb1 < a.k_blocks) {
    w.kind = 2;  // later CTAs hold the rest: CTAs blockIdx.x+1 .. whose ranges start inside this tile
    const long long t_end = t_begin + a.k_blocks;
    int f = 0;
    while ((static_cast<long long>(vid) + f + 1) * a.units_per_cta < t_end) ++f;
    w.followers = f;
  } else {
    w.kind = 0;
  }
  return true;
}

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Optional per-CTA timestamps (clock64) for tools/: [start, producer done, mma done, first accumulator ready,
// follower flags seen, epilogue done, partial published, -]
__device__ unsigned long long g_gemm_dbg[160 * 8];
__device__ int g_gemm_dbg_on = 0;
#define DBG_T(slot) do { if (g_gemm_dbg_on && lane == 0) g_gemm_dbg[blockIdx.x * 8 + (slot)] = clock64(); } while (0)

// Store one 32-column chunk of the accumulator row(s) this thread owns.  v: the 32 fp32 values; m: A-row index of the
// thread; n0: first B-row index of the chunk; f_base: first A-row of this warp (transposed mode); sc: the warp's 32x33
// fp32 shared-memory transpose scratch.
__device__ __forceinline__ void store_chunk(const GemmArgs& args, float (&v)[32], int m, bool m_ok, int n0, int nvalid, int ks,
                                            int f_base, float* sc, int lane) {
  if (args.swiglu) {
    // Fused SwiGLU (replaces `act_fn(gate_proj(x)) * up_proj(x)` of LlamaMLP): the weight's rows are interleaved at bind
    // time in blocks of 32 = [gate i0..i0+15 | up i0..i0+15], so gate and up of one act element meet in one chunk.
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(args.out);
    if (!args.transposed) {
      // thread = token m, v[0..15] = gate, v[16..31] = up of act columns n0/2 .. n0/2+15 (host guarantees N % 32 == 0)
      if (!m_ok) return;
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float g0 = v[2 * j], g1 = v[2 * j + 1];
        o[j] = pack_bf16x2(__fdividef(g0, 1.f + __expf(-g0)) * v[16 + 2 * j], __fdividef(g1, 1.f + __expf(-g1)) * v[17 + 2 * j]);
      }
      uint4* dst = reinterpret_cast<uint4*>(out + static_cast<long long>(m) * args.ldc + (n0 >> 1));
      dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
      dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
      // thread = weight row m: lanes 0-15 hold gate, lanes 16-31 up of act columns f_base/2 + (lane & 15); v[j] = token
      // n0 + j.  Lower lanes finish the even tokens, upper lanes the odd ones: one shuffle per token pair.
      const bool lower = lane < 16;
      __nv_bfloat16* col = out + (f_base >> 1) + (lane & 15);
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float got = __shfl_xor_sync(0xffffffffu, lower ? v[j + 1] : v[j], 16);
        const float g = lower ? v[j] : got;
        const float u = lower ? got : v[j + 1];
        const int jj = j + (lower ? 0 : 1);
        if (jj < nvalid && m_ok) col[static_cast<long long>(n0 + jj) * args.ldc] = __float2bfloat16(__fdividef(g, 1.f + __expf(-g)) * u);
      }
    }
    return;
  }
  // ---- fast paths (full 32-column chunk, no bias / ReLU / row map): few instructions per element, because
  // the four epilogue warps run one per scheduler and every instruction's latency is exposed
  const bool plain = nvalid == 32 && args.bias == nullptr && !args.relu && args.rowmap == nullptr && !args.atomic;
  if (plain && !args.transposed) {
    if (!m_ok) return;
    if (args.mode == RVL_GEMM_OUT_BF16) {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(args.out) + static_cast<long long>(m) * args.ldc + n0);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        dst[q] = make_uint4(pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]), pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]),
                            pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
    } else {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + ks * args.split_stride + static_cast<long long>(m) * args.ldc + n0);
      if (args.mode == RVL_GEMM_ADD_F32) {
        float4 p[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) p[q] = dst[q];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          dst[q] = make_float4(v[q * 4] + p[q].x, v[q * 4 + 1] + p[q].y, v[q * 4 + 2] + p[q].z, v[q * 4 + 3] + p[q].w);
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      }
    }
    return;
  }
  const bool plain_any = args.bias == nullptr && !args.relu && args.rowmap == nullptr && !args.atomic;
  if (plain_any && args.transposed && f_base + 32 <= args.M) {
    // out[token][feature]: transpose the warp's 32 features x 32 tokens through shared memory so that each
    // lane stores 16 contiguous bytes of one token row
#pragma unroll
    for (int j = 0; j < 32; ++j) sc[lane * 33 + j] = v[j];
    __syncwarp();
    if (args.mode == RVL_GEMM_OUT_BF16) {
      const int f0 = (lane & 3) * 8, t0 = lane >> 2;
      __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(args.out) + static_cast<long long>(n0 + t0) * args.ldc + f_base + f0;
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int t = t0 + pass * 8;
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = sc[(f0 + i) * 33 + t];
        if (t < nvalid)   // ragged last chunk of the token dimension
          *reinterpret_cast<uint4*>(base + static_cast<long long>(pass) * 8 * args.ldc) =
              make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
      }
    } else {
      const int f0 = (lane & 7) * 4, t0 = lane >> 3;
      float* base = reinterpret_cast<float*>(args.out) + ks * args.split_stride + static_cast<long long>(n0 + t0) * args.ldc + f_base + f0;
#pragma unroll
      for (int pass = 0; pass < 8; ++pass) {
        const int t = t0 + pass * 4;
        float4 x = make_float4(sc[(f0 + 0) * 33 + t], sc[(f0 + 1) * 33 + t], sc[(f0 + 2) * 33 + t], sc[(f0 + 3) * 33 + t]);
        float4* dst = reinterpret_cast<float4*>(base + static_cast<long long>(pass) * 4 * args.ldc);
        if (t < nvalid) {
          if (args.mode == RVL_GEMM_ADD_F32) {
            const float4 p = *dst;
            x.x += p.x; x.y += p.y; x.z += p.z; x.w += p.w;
          }
          *dst = x;
        }
      }
    }
    __syncwarp();
    return;
  }
  if (!args.transposed) {
    // thread = token row m, 32 consecutive features n0..n0+31
    if (!m_ok) return;
    if (args.bias != nullptr && ks == 0) {
      if (nvalid == 32) {
        const uint4* bp = reinterpret_cast<const uint4*>(args.bias + n0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 b = __ldg(bp + q);
          v[q * 8 + 0] += bf16_lo(b.x); v[q * 8 + 1] += bf16_hi(b.x);
          v[q * 8 + 2] += bf16_lo(b.y); v[q * 8 + 3] += bf16_hi(b.y);
          v[q * 8 + 4] += bf16_lo(b.z); v[q * 8 + 5] += bf16_hi(b.z);
          v[q * 8 + 6] += bf16_lo(b.w); v[q * 8 + 7] += bf16_hi(b.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) v[j] += __bfloat162float(args.bias[n0 + j]);
      }
    }
    if (args.relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    const long long row = args.rowmap ? args.rowmap[m] : m;
    if (args.mode == RVL_GEMM_OUT_BF16) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(args.out) + row * args.ldc + n0;
      if (nvalid == 32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
          o.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
          o.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
          o.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
          reinterpret_cast<uint4*>(dst)[q] = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) dst[j] = __float2bfloat16(v[j]);
      }
    } else {
      float* dst = reinterpret_cast<float*>(args.out) + ks * args.split_stride + row * args.ldc + n0;
      if (args.atomic) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) atomicAdd(dst + j, v[j]);
      } else if (nvalid == 32) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 o = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          if (args.mode == RVL_GEMM_ADD_F32) {
            const float4 p = reinterpret_cast<const float4*>(dst)[q];
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
          }
          reinterpret_cast<float4*>(dst)[q] = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) dst[j] = (args.mode == RVL_GEMM_ADD_F32 ? dst[j] : 0.f) + v[j];
      }
    }
  } else {
    // thread = feature m, its 32 values are tokens n0..n0+31: out[token][feature], lanes -> consecutive
    // features (coalesced along the feature dimension)
    if (!m_ok) return;
    const float b = (args.bias != nullptr && ks == 0) ? __bfloat162float(args.bias[m]) : 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < nvalid) {
        float x = v[j] + b;
        if (args.relu) x = fmaxf(x, 0.f);
        const long long row = args.rowmap ? args.rowmap[n0 + j] : (n0 + j);
        if (args.mode == RVL_GEMM_OUT_BF16) {
          reinterpret_cast<__nv_bfloat16*>(args.out)[row * args.ldc + m] = __float2bfloat16(x);
        } else {
          float* dst = reinterpret_cast<float*>(args.out) + ks * args.split_stride + row * args.ldc + m;
          if (args.atomic) atomicAdd(dst, x);
          else if (args.mode == RVL_GEMM_ADD_F32) *dst += x;
          else *dst = x;
        }
      }
    }
  }
}

// kATiles: 128-row A tiles per CTA tile (compile time so the MMA / epilogue loops specialise);
// kStreamK: k-blocks dealt evenly to the CTAs with in-kernel fix-up of the partial tiles.
template <int kATiles, bool kStreamK>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out, const GemmArgs args) {
  const int kStages = args.stages;
  const int BN = args.bn;
  constexpr int a_tiles = kATiles;
  const int a_slot_bytes = a_tiles * kATileBytes;
  const int b_tile_bytes = BN * kBK * 2;
  const uint32_t stage_tx_bytes = static_cast<uint32_t>(a_slot_bytes + b_tile_bytes);
  const int acc_cols = a_tiles * args.sub_stride;          // TMEM columns of one accumulator stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                  // [kStages][a_tiles][128 x 64] bf16
  uint8_t* smem_b = smem + kStages * a_slot_bytes;         // [kStages][BN x 64] bf16 (BN * 128 B is a multiple of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (a_slot_bytes + b_tile_bytes));
  uint64_t* full_bar = bars;                       // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;         // [kMaxStages]
  uint64_t* tmem_full = bars + 2 * kMaxStages;     // [2]
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  float* epi_scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [4 warps][32][33]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (args.staged) tma_prefetch_desc(&tmap_out);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, static_cast<uint32_t>(args.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 0) DBG_T(0);
  // Programmatic dependent launch: this CTA may have started while the previous kernel of the stream (the one that
  // produces the activations) is still running.  Everything above touched only kernel parameters and shared memory.
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t pol_a = l2_policy_evict_first(), pol_b = l2_policy_evict_last();
    WorkItem w;
    // Weight prefetch: A (a bound weight in the weight-streaming orientation) depends on no earlier kernel, so the first
    // kStages k-blocks of it are already on their way into the ring while the producer of the activations finishes.
    int prefetched = 0;
    if (args.prefetch_a) {
      for (int it = 0; prefetched < kStages && get_work<kStreamK>(args, it, w); ++it) {
        const int m0 = w.m_blk * a_tiles * kBM;
        for (int kb = w.kb0; kb < w.kb1 && prefetched < kStages; ++kb, ++prefetched) {
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[prefetched], stage_tx_bytes);   // B's bytes are issued after pdl_wait()
            tma_load_2d_hint(smem_a + prefetched * a_slot_bytes, &tmap_a, &full_bar[prefetched], kb * kBK, m0, pol_a);
          }
          __syncwarp();
        }
      }
    }
    pdl_wait();
    int n_iter = 0;
    for (int it = 0; get_work<kStreamK>(args, it, w); ++it) {
      const int m0 = w.m_blk * a_tiles * kBM, n0 = w.n_blk * BN;
      for (int kb = w.kb0; kb < w.kb1; ++kb, ++n_iter) {
        if (n_iter < prefetched) {
          // slot's first use: A is in flight already, the barrier already expects B's bytes
          if (elect_one()) tma_load_2d_hint(smem_b + stage * b_tile_bytes, &tmap_b, &full_bar[stage], kb * kBK, n0, pol_b);
        } else {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
            if (args.stream_a) {
              tma_load_2d_hint(smem_a + stage * a_slot_bytes, &tmap_a, &full_bar[stage], kb * kBK, m0, pol_a);
              tma_load_2d_hint(smem_b + stage * b_tile_bytes, &tmap_b, &full_bar[stage], kb * kBK, n0, pol_b);
            } else {
              tma_load_2d(smem_a + stage * a_slot_bytes, &tmap_a, &full_bar[stage], kb * kBK, m0);
              tma_load_2d(smem_b + stage * b_tile_bytes, &tmap_b, &full_bar[stage], kb * kBK, n0);
            }
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
    DBG_T(1);
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp loops, one lane issues)
    const uint32_t idesc = umma_idesc_bf16(kBM, static_cast<uint32_t>(BN));
    const uint64_t a_desc0 = umma_desc_k_sw128(smem_u32(smem_a));
    const uint64_t b_desc0 = umma_desc_k_sw128(smem_u32(smem_b));
    const uint64_t a_step = static_cast<uint64_t>(a_slot_bytes >> 4), b_step = static_cast<uint64_t>(b_tile_bytes >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    WorkItem w;
    for (int it = 0; get_work<kStreamK>(args, it, w); ++it) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * acc_cols;
      for (int kb = w.kb0; kb < w.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = a_desc0 + a_step * stage;
          const uint64_t b_desc = b_desc0 + b_step * stage;
          const uint32_t accum = kb > w.kb0 ? 1u : 0u;
          for (int at = 0; at < a_tiles; ++at) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // advance 16 elements = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
              umma_bf16(d_tmem + at * args.sub_stride, a_desc + at * (kATileBytes >> 4) + 2 * k, b_desc + 2 * k, idesc,
                        (accum | (k > 0)) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);  // smem slot is free once these MMAs have read it
          if (kb == w.kb1 - 1) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == args.acc_stages) { acc = 0; acc_phase ^= 1; }
    }
    DBG_T(2);
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    pdl_wait();                    // residual / partial / flag reads and every store come after the earlier kernels
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int m_local = quarter * 32 + lane;
    const int n_chunks = (BN + 31) >> 5;
    int acc = 0;
    uint32_t acc_phase = 0;
    WorkItem w;
    for (int it = 0; get_work<kStreamK>(args, it, w); ++it) {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (warp == 2 && it == 0) DBG_T(3);
      if (kStreamK && w.kind == 2) {
        // head piece of a stream-K tile: the CTAs after this one computed the rest of the k-range early in their
        // own ranges; wait for their flags (they never wait on anything before publishing, and all CTAs are
        // co-resident, so this cannot deadlock)
        for (int f = 1; f <= w.followers; ++f) {
          const unsigned int* fl = args.flags + blockIdx.x + f;
          unsigned int spins = 0;
          while (ld_acquire(fl) == 0u) {
            if (++spins > (1u << 26)) mbar_timeout(nullptr, 0xdead);
          }
        }
        if (warp == 2) DBG_T(4);
      }
      const int ks = w.ks;
      if (args.staged && !(kStreamK && w.kind == 1)) {
        // ---- staged epilogue (weight-streaming orientation, a_tiles == 1): thread = feature row, registers = tokens.
        // Each chunk's 32 tokens x 128 features land in shared memory as [token][feature] (every st.shared of a warp
        // writes 32 consecutive features of one token: conflict-free, one instruction per value with an immediate
        // offset) and leave through one TMA store per chunk, which also clips the ragged token / feature edges.
        // Measured before this path: ~1050 cycles per chunk for the register transpose + 16-byte global stores with one
        // epilogue warp per scheduler (every instruction's latency exposed), ~4900 for a ragged chunk.
        const uint32_t stg = smem_u32(epi_scratch);
        const uint32_t taddr = tmem_base + acc * acc_cols + (static_cast<uint32_t>(quarter * 32) << 16);
        const int cpg = args.staged;                       // chunks per staging group
        const int feat0 = w.m_blk * kBM;                   // first weight row of the tile
        for (int c0 = 0; c0 < n_chunks; c0 += cpg) {
          // the staging buffer may still be read by the previous group's stores
          if (threadIdx.x == 64) bulk_wait_read_all();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int c1 = min(c0 + cpg, n_chunks);
          for (int c = c0; c < c1; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
            if (c == n_chunks - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            }
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (kStreamK && w.kind == 2) {
              for (int f = 1; f <= w.followers; ++f) {
                const float4* src = reinterpret_cast<const float4*>(args.ws) +
                                    (static_cast<long long>(blockIdx.x + f) * 8 + c) * (8 * 128) + m_local;
                float4 p[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) p[q] = __ldcg(src + q * 128);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  v[q * 4] += p[q].x; v[q * 4 + 1] += p[q].y; v[q * 4 + 2] += p[q].z; v[q * 4 + 3] += p[q].w;
                }
              }
            }
            const int lc = c - c0;
            if (args.swiglu) {
              // lanes 0-15 hold gate, lanes 16-31 up of act columns quarter*16 + (lane & 15); lower lanes finish the even
              // tokens, upper lanes the odd ones (one shuffle per token pair); staging row = 64 act columns = 128 B
              const bool lower = lane < 16;
              const uint32_t base = stg + lc * (32 * 128) + (quarter * 16 + (lane & 15)) * 2 + (lower ? 0 : 128);
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float got = __shfl_xor_sync(0xffffffffu, lower ? v[j + 1] : v[j], 16);
                const float g = lower ? v[j] : got;
                const float u = lower ? got : v[j + 1];
                const __nv_bfloat16 a = __float2bfloat16(__fdividef(g, 1.f + __expf(-g)) * u);
                st_shared_u16(base + j * 128, *reinterpret_cast<const uint16_t*>(&a));
              }
            } else if (args.mode == RVL_GEMM_OUT_BF16) {
              const uint32_t base = stg + lc * (32 * 256) + m_local * 2;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const __nv_bfloat16 a = __float2bfloat16(v[j]);
                st_shared_u16(base + j * 256, *reinterpret_cast<const uint16_t*>(&a));
              }
            } else {
              const uint32_t base = stg + lc * (32 * 512) + m_local * 4;
#pragma unroll
              for (int j = 0; j < 32; ++j) st_shared_f32(base + j * 512, v[j]);
            }
          }
          fence_proxy_async();                             // generic-proxy writes -> visible to the TMA engine
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (threadIdx.x == 64) {
            const int chunk_bytes = args.swiglu ? 32 * 128 : (args.mode == RVL_GEMM_OUT_BF16 ? 32 * 256 : 32 * 512);
            for (int c = c0; c < c1; ++c) {
              const int tok0 = w.n_blk * BN + c * 32;
              if (tok0 < args.N && c * 32 < BN)
                tma_store_3d(&tmap_out, epi_scratch + (c - c0) * (chunk_bytes / 4), args.swiglu ? (feat0 >> 1) : feat0, tok0, ks);
            }
            bulk_commit_group();
          }
        }
        if (kStreamK && w.kind == 2) {
          // every epilogue thread has read the followers' partials (the bar.sync above): clear their flags for the next launch
          if (threadIdx.x == 64)
            for (int f = 1; f <= w.followers; ++f) args.flags[blockIdx.x + f] = 0u;
        }
        if (++acc == args.acc_stages) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      for (int at = 0; at < a_tiles; ++at) {
        const int m = (w.m_blk * a_tiles + at) * kBM + m_local;  // A-row owned by this thread
        const uint32_t taddr = tmem_base + acc * acc_cols + at * args.sub_stride + (static_cast<uint32_t>(quarter * 32) << 16);
        const bool m_ok = m < args.M;
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (!kStreamK && warp == 2 && c == 0) DBG_T(4);
          if (!kStreamK && warp == 2 && c == 1) DBG_T(6);
          if (at == a_tiles - 1 && c == n_chunks - 1) {
            // all TMEM reads of this accumulator are done: hand it back to the MMA warp early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          const int n0 = w.n_blk * BN + c * 32;
          if (n0 >= args.N || c * 32 >= BN) continue;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          const int nvalid = min(min(32, args.N - n0), BN - c * 32);
          // stream-K partial tiles live in the workspace in "register order": [cta][a tile][chunk][q][thread][4 floats],
          // so every warp-level 16-byte access is one contiguous 512-byte run
          if (kStreamK && w.kind == 1) {
            float4* dst = reinterpret_cast<float4*>(args.ws) +
                          ((static_cast<long long>(blockIdx.x) * a_tiles + at) * 8 + c) * (8 * 128) + m_local;
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q * 128] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            continue;
          }
          if (kStreamK && w.kind == 2) {
            for (int f = 1; f <= w.followers; ++f) {
              const float4* src = reinterpret_cast<const float4*>(args.ws) +
                                  ((static_cast<long long>(blockIdx.x + f) * a_tiles + at) * 8 + c) * (8 * 128) + m_local;
              float4 p[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) p[q] = __ldcg(src + q * 128);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                v[q * 4] += p[q].x; v[q * 4 + 1] += p[q].y; v[q * 4 + 2] += p[q].z; v[q * 4 + 3] += p[q].w;
              }
            }
          }
          store_chunk(args, v, m, m_ok, n0, nvalid, ks, (w.m_blk * a_tiles + at) * kBM + quarter * 32,
                      epi_scratch + (warp - 2) * (32 * 33), lane);
        }
      }
      if (kStreamK && w.kind == 1) {
        // publish the partial: every epilogue thread fences its own stores, then one thread raises the flag
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) st_release(args.flags + blockIdx.x, 1u);
        if (warp == 2) DBG_T(6);
      }
      if (kStreamK && w.kind == 2) {
        // all four epilogue warps are done with the followers' partials: clear their flags for the next launch
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64)
          for (int f = 1; f <= w.followers; ++f) args.flags[blockIdx.x + f] = 0u;
      }
      if (++acc == args.acc_stages) { acc = 0; acc_phase ^= 1; }
    }
    if (args.staged && threadIdx.x == 64) bulk_wait_all();   // the staging buffer must outlive the stores that read it
    if (warp == 2) DBG_T(5);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) DBG_T(7);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(args.tmem_cols));
  }
}


// ------------------------------------------------------------------------------------------- CTA-pair variant
// Token-major GEMMs with many rows (prefill): two CTAs of a cluster share one 256 x 256 output tile through
// tcgen05.mma.cta_group::2.  Each CTA loads its own 128 A rows and HALF of the B tile (128 of the 256 rows), so a
// k-block costs 32 KB of TMA traffic per SM instead of 48 KB - the single-CTA tile is bound by the ~55-72 B/clk one
// SM's TMA path delivers, not by the tensor pipe.  The even CTA issues the MMAs for both; accumulators (128 rows x
// 256 columns per CTA, two stages) sit in each CTA's own TMEM and each CTA's epilogue warps store its 128 rows.
// Barriers: full[s] lives in the leader and counts the bytes of both CTAs' loads; empty[s] and tmem_full[a] exist in
// both CTAs and are signalled by multicast tcgen05.commit; tmem_empty[a] lives in the leader and is arrived on by the
// 8 epilogue warps of the pair.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const GemmArgs args) {
  constexpr int BN = 256, kHalfN = 128;
  constexpr int kStageBytesP = kATileBytes + kHalfN * kBK * 2;   // 32 KB
  const int kStages = args.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kATileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytesP);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  float* epi_scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);     // leader's instance is the one used: one arrive.expect_tx for the bytes of both CTAs
      mbar_init(&empty_bar[i], 1);    // multicast commit from the leader's MMA warp
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);    // multicast commit
      mbar_init(&tmem_empty[i], 8);   // leader's instance: 4 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer's barriers must exist before anything is signalled remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_trigger();                      // the set-up above overlapped the previous kernel's tail (programmatic dependent launch)
  pdl_wait();

  const int n_clusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;
  const int total_tiles = args.tiles_m * args.tiles_n;     // tiles_m counts 256-row pair tiles

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      int m_blk, n_blk;
      tile_coords(tile, args.tiles_m, args.tiles_n, args.group_m, m_blk, n_blk);
      const int m0 = (m_blk * 2 + rank) * kBM, n0 = n_blk * BN + rank * kHalfN;
      for (int kb = 0; kb < args.k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytesP);
          tma_load_2d_pair(smem_a + stage * kATileBytes, &tmap_a, full_leader, kb * kBK, m0);
          tma_load_2d_pair(smem_b + stage * (kHalfN * kBK * 2), &tmap_b, full_leader, kb * kBK, n0);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------------------------------------------------- MMA issuer (leader CTA only)
      const uint32_t idesc = umma_idesc_bf16(256, BN);
      const uint64_t a_desc0 = umma_desc_k_sw128(smem_u32(smem_a));
      const uint64_t b_desc0 = umma_desc_k_sw128(smem_u32(smem_b));
      const uint64_t a_step = static_cast<uint64_t>(kATileBytes >> 4), b_step = static_cast<uint64_t>((kHalfN * kBK * 2) >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < args.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a_desc = a_desc0 + a_step * stage;
            const uint64_t b_desc = b_desc0 + b_step * stage;
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage], 3);                          // frees the slot in both CTAs
            if (kb == args.k_blocks - 1) umma_commit_pair(&tmem_full[acc], 3);  // accumulators ready in both CTAs
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (both CTAs, own 128 rows)
    const int quarter = warp & 3;
    const int m_local = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      int m_blk, n_blk;
      tile_coords(tile, args.tiles_m, args.tiles_n, args.group_m, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m = (m_blk * 2 + rank) * kBM + m_local;
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      const bool m_ok = m < args.M;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (c == BN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
        }
        const int n0 = n_blk * BN + c * 32;
        if (n0 >= args.N) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        store_chunk(args, v, m, m_ok, n0, min(32, args.N - n0), 0, (m_blk * 2 + rank) * kBM + quarter * 32,
                    epi_scratch + (warp - 2) * (32 * 33), lane);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // nobody frees TMEM or exits while the peer may still signal / be read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------- CTA-pair, weight streaming
// Weight-streaming orientation with MANY tokens (decode at 65 - 256 sequences): the activation tile (tokens x 64, 24 KB at
// 180 tokens) is re-fetched from L2 for every 16 KB weight tile, and one SM ingests at most ~75 B/clk from L2 + HBM together
// (tools/probes/tma_probe2.cu) - at the 1.1 - 1.3 GHz the power-capped chip runs inside the sweep that is 40 KB per ~530
// cycles = ~0.45 us per k-block and SM, i.e. ~4.4 TB/s of weights with every SM pulling: ingest-bound, not HBM-bound.  Here
// two CTAs of a cluster share the activation tile through tcgen05.mma.cta_group::2: each loads its own 128 weight rows
// and HALF of the token rows (28 KB per k-block), the even CTA issues M = 256 MMAs for both, each CTA keeps the accumulator
// of its own 128 weight rows in its TMEM and runs the staged TMA-store epilogue for them.  Work units are (256-row weight
// tile, k-split) pairs dealt round-robin to the clusters; split-k partials go to the caller's partial buffers.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_stream_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const __grid_constant__ CUtensorMap tmap_out, const GemmArgs args) {
  const int kStages = args.stages;
  const int BN = args.bn;                       // tokens (MMA N), multiple of 16
  const int half_rows = BN >> 1;
  const int b_half_bytes = half_rows * kBK * 2;
  const int stage_bytes = kATileBytes + b_half_bytes;
  const int acc_cols = args.sub_stride;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kATileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kATileBytes + ((b_half_bytes + 1023) & ~1023)));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  uint8_t* stage_out = reinterpret_cast<uint8_t*>(bars) + 512;
  const int b_slot_bytes = (b_half_bytes + 1023) & ~1023;        // every B slot starts on a 1024-byte swizzle boundary

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, static_cast<uint32_t>(args.tmem_cols));
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_trigger();

  const int n_clusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;
  const int tiles = args.tiles_m;                               // 256-row weight tiles
  const int n_units = tiles * args.split_k;
  const int kb_per_split = (args.k_blocks + args.split_k - 1) / args.split_k;
  // it-th work item of this cluster: stream-K (k-blocks of all tiles dealt evenly to the clusters, partial tiles fixed up
  // through the workspace exactly as in gemm_bf16_tcgen05_kernel<1, true>, per CTA of the pair) or (tile, k-split) units
  auto next_item = [&](int it, WorkItem& w) -> bool {
    if (args.stream_k) return get_work<true>(args, it, w, cluster_id);
    const int u = cluster_id + it * n_clusters;
    if (u >= n_units) return false;
    w.ks = u / tiles;
    w.m_blk = u - w.ks * tiles;
    w.n_blk = 0;
    w.kb0 = w.ks * kb_per_split;
    w.kb1 = min(args.k_blocks, w.kb0 + kb_per_split);
    w.kind = 0;
    w.followers = 0;
    return true;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    const uint64_t pol_a = l2_policy_evict_first(), pol_b = l2_policy_evict_last();
    int stage = 0;
    uint32_t phase = 0;
    // weight tiles of the first k-blocks go out before the kernel that produces the activations is known to be finished
    int pre = 0;
    if (args.prefetch_a) {
      WorkItem w;
      for (int it = 0; pre < kStages && next_item(it, w); ++it) {
        for (int kb = w.kb0; kb < w.kb1 && pre < kStages; ++kb, ++pre) {
          if (elect_one()) {
            const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[pre]), 0);
            if (leader) mbar_arrive_expect_tx(&full_bar[pre], 2 * stage_bytes);
            tma_load_2d_pair_hint(smem_a + pre * kATileBytes, &tmap_a, full_leader, kb * kBK, (w.m_blk * 2 + static_cast<int>(rank)) * kBM, pol_a);
          }
          __syncwarp();
        }
      }
    }
    pdl_wait();
    int n_iter = 0;
    WorkItem w;
    for (int it = 0; next_item(it, w); ++it) {
      const int m0 = (w.m_blk * 2 + static_cast<int>(rank)) * kBM;
      for (int kb = w.kb0; kb < w.kb1; ++kb, ++n_iter) {
        const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
        if (n_iter < pre) {
          if (elect_one()) tma_load_2d_pair_hint(smem_b + stage * b_slot_bytes, &tmap_b, full_leader, kb * kBK, static_cast<int>(rank) * half_rows, pol_b);
        } else {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_bytes);
            tma_load_2d_pair_hint(smem_a + stage * kATileBytes, &tmap_a, full_leader, kb * kBK, m0, pol_a);
            tma_load_2d_pair_hint(smem_b + stage * b_slot_bytes, &tmap_b, full_leader, kb * kBK, static_cast<int>(rank) * half_rows, pol_b);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------------------------------------------------- MMA issuer (leader CTA only)
      const uint32_t idesc = umma_idesc_bf16(256, static_cast<uint32_t>(BN));
      const uint64_t a_desc0 = umma_desc_k_sw128(smem_u32(smem_a));
      const uint64_t b_desc0 = umma_desc_k_sw128(smem_u32(smem_b));
      const uint64_t a_step = static_cast<uint64_t>(kATileBytes >> 4), b_step = static_cast<uint64_t>(b_slot_bytes >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      WorkItem w;
      for (int it = 0; next_item(it, w); ++it) {
        const int kb0 = w.kb0, kb1 = w.kb1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_cols;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a_desc = a_desc0 + a_step * stage;
            const uint64_t b_desc = b_desc0 + b_step * stage;
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage], 3);
            if (kb == kb1 - 1) umma_commit_pair(&tmem_full[acc], 3);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == args.acc_stages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (both CTAs: own 128 weight rows), staged TMA stores
    pdl_wait();
    const int quarter = warp & 3;
    const int m_local = quarter * 32 + lane;
    const int n_chunks = (BN + 31) >> 5;
    const uint32_t stg = smem_u32(stage_out);
    const int cpg = args.staged;
    int acc = 0;
    uint32_t acc_phase = 0;
    WorkItem w;
    for (int it = 0; next_item(it, w); ++it) {
      const int ks = w.ks, tile = w.m_blk;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * acc_cols + (static_cast<uint32_t>(quarter * 32) << 16);
      const int feat0 = (tile * 2 + static_cast<int>(rank)) * kBM;
      if (w.kind == 1) {
        // tail piece of a stream-K tile: publish this CTA's fp32 partial (register order) and raise its flag
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c == n_chunks - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
          }
          float4* dst = reinterpret_cast<float4*>(args.ws) + (static_cast<long long>(blockIdx.x) * 8 + c) * (8 * 128) + m_local;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            dst[q * 128] = make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]),
                                       __uint_as_float(r[q * 4 + 3]));
        }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) st_release(args.flags + blockIdx.x, 1u);
        if (++acc == args.acc_stages) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      if (w.kind == 2) {
        // head piece: the same-rank CTAs of the following clusters hold the rest of the k-range
        for (int f = 1; f <= w.followers; ++f) {
          const unsigned int* fl = args.flags + blockIdx.x + 2 * f;
          unsigned int spins = 0;
          while (ld_acquire(fl) == 0u) {
            if (++spins > (1u << 26)) mbar_timeout(nullptr, 0xdead);
          }
        }
      }
      for (int c0 = 0; c0 < n_chunks; c0 += cpg) {
        if (threadIdx.x == 64) bulk_wait_read_all();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int c1 = min(c0 + cpg, n_chunks);
        for (int c = c0; c < c1; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c == n_chunks - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
          }
          for (int f = 1; f <= w.followers; ++f) {               // stream-K head: add the followers' partials in a fixed order
            const float4* src = reinterpret_cast<const float4*>(args.ws) +
                                (static_cast<long long>(blockIdx.x + 2 * f) * 8 + c) * (8 * 128) + m_local;
            float4 pp[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) pp[q] = __ldcg(src + q * 128);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              r[q * 4] = __float_as_uint(__uint_as_float(r[q * 4]) + pp[q].x);
              r[q * 4 + 1] = __float_as_uint(__uint_as_float(r[q * 4 + 1]) + pp[q].y);
              r[q * 4 + 2] = __float_as_uint(__uint_as_float(r[q * 4 + 2]) + pp[q].z);
              r[q * 4 + 3] = __float_as_uint(__uint_as_float(r[q * 4 + 3]) + pp[q].w);
            }
          }
          const int lc = c - c0;
          if (args.swiglu) {
            const bool lower = lane < 16;
            const uint32_t base = stg + lc * (32 * 128) + (quarter * 16 + (lane & 15)) * 2 + (lower ? 0 : 128);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float vj = __uint_as_float(r[j]), vj1 = __uint_as_float(r[j + 1]);
              const float got = __shfl_xor_sync(0xffffffffu, lower ? vj1 : vj, 16);
              const float g = lower ? vj : got;
              const float uu = lower ? got : vj1;
              const __nv_bfloat16 a = __float2bfloat16(__fdividef(g, 1.f + __expf(-g)) * uu);
              st_shared_u16(base + j * 128, *reinterpret_cast<const uint16_t*>(&a));
            }
          } else if (args.mode == RVL_GEMM_OUT_BF16) {
            const uint32_t base = stg + lc * (32 * 256) + m_local * 2;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const __nv_bfloat16 a = __float2bfloat16(__uint_as_float(r[j]));
              st_shared_u16(base + j * 256, *reinterpret_cast<const uint16_t*>(&a));
            }
          } else {
            const uint32_t base = stg + lc * (32 * 512) + m_local * 4;
#pragma unroll
            for (int j = 0; j < 32; ++j) st_shared_f32(base + j * 512, __uint_as_float(r[j]));
          }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
          const int chunk_bytes = args.swiglu ? 32 * 128 : (args.mode == RVL_GEMM_OUT_BF16 ? 32 * 256 : 32 * 512);
          for (int c = c0; c < c1; ++c) {
            const int tok0 = c * 32;
            if (tok0 < args.N) tma_store_3d(&tmap_out, stage_out + (c - c0) * chunk_bytes, args.swiglu ? (feat0 >> 1) : feat0, tok0, ks);
          }
          bulk_commit_group();
        }
      }
      if (w.kind == 2 && threadIdx.x == 64)                      // partials consumed (the bar.sync above): flags back to zero
        for (int f = 1; f <= w.followers; ++f) args.flags[blockIdx.x + 2 * f] = 0u;
      if (++acc == args.acc_stages) { acc = 0; acc_phase ^= 1; }
    }
    if (threadIdx.x == 64) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, static_cast<uint32_t>(args.tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128-byte swizzle, OOB -> zeros.
// `ld` (elements): row pitch when the matrix is a column slice of a wider one (0: the rows are dense, pitch = cols)
static int make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int box_rows, std::string* err, int64_t ld = 0) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { *err = "cuTensorMapEncodeTiled entry point not available"; return RVL_ERR_CUDA; }
  if (ld <= 0) ld = cols;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16) {
    *err = "GEMM operand must be 16-byte aligned with a row pitch that is a multiple of 16 bytes";
    return RVL_ERR_INVALID;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box_rows=%d", (int)r,
             (long long)rows, (long long)cols, box_rows);
    *err = buf;
    return RVL_ERR_CUDA;
  }
  return RVL_OK;
}

// Output tensor map for the staged epilogue: [splits][tokens][features] (features contiguous), box = 32 tokens x box_feat
// features, no swizzle (the staging buffer is a dense [token][feature] tile).
static int make_tmap_out(CUtensorMap* tm, const void* base, int64_t n_feat, int64_t n_tok, int64_t n_split, int64_t ld_elems,
                         int64_t split_stride_elems, int box_feat, bool f32, std::string* err) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { *err = "cuTensorMapEncodeTiled entry point not available"; return RVL_ERR_CUDA; }
  const int es = f32 ? 4 : 2;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(n_feat), static_cast<cuuint64_t>(n_tok), static_cast<cuuint64_t>(n_split)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(ld_elems) * es,
                        static_cast<cuuint64_t>(n_split > 1 ? split_stride_elems : ld_elems * n_tok) * es};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_feat), 32u, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim,
                   gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(out) failed (%d) feat=%lld tok=%lld ld=%lld", (int)r, (long long)n_feat,
             (long long)n_tok, (long long)ld_elems);
    *err = buf;
    return RVL_ERR_CUDA;
  }
  return RVL_OK;
}

int make_tmap_bf16_2d(::CUtensorMap_st* tm, const void* base, int64_t rows, int64_t cols, int box_rows, std::string* err, int64_t ld) {
  return make_tmap(tm, base, rows, cols, box_rows, err, ld);
}

static int pow2_at_least(int x) {
  int c = 32;
  while (c < x) c <<= 1;
  return c;
}

// out[tokens, features] = act(X[tokens,K] . W[features,K]^T + bias)
int gemm_bf16(const GemmCall& c, int num_sms, cudaStream_t st, std::string* err) {
  if (c.M <= 0 || c.N <= 0 || c.K <= 0) { *err = "gemm: empty problem"; return RVL_ERR_INVALID; }
  if (c.K % 8 || c.N % 8) { *err = "gemm: K and N must be multiples of 8"; return RVL_ERR_INVALID; }
  const bool partials = c.split_stride > 0;
  if (c.split_k > 1 && !partials && c.out_mode != RVL_GEMM_ADD_F32) { *err = "gemm: split_k needs RVL_GEMM_ADD_F32 or a partial buffer"; return RVL_ERR_INVALID; }
  if (partials && c.out_mode != RVL_GEMM_OUT_F32) { *err = "gemm: split-k partials are fp32 (RVL_GEMM_OUT_F32)"; return RVL_ERR_INVALID; }
  const bool swap = (c.flags & RVL_GEMM_FLAG_SWAP) != 0;
  const Tuning& tn = tuning();                            // RVL_* experiment hooks, read once per process
  const bool env_dbg = tn.plan_debug != 0;
  GemmArgs a{};
  a.K = static_cast<int>(c.K);
  a.k_blocks = static_cast<int>((c.K + kBK - 1) / kBK);
  a.group_m = 16;
  a.ldc = c.ldc;
  a.out = c.out;
  a.bias = reinterpret_cast<const __nv_bfloat16*>(c.bias);
  a.rowmap = c.rowmap;
  a.mode = c.out_mode;
  a.relu = (c.flags & RVL_GEMM_FLAG_RELU) ? 1 : 0;
  a.transposed = swap ? 1 : 0;
  a.swiglu = (c.flags & RVL_GEMM_FLAG_SWIGLU) ? 1 : 0;
  if (a.swiglu && (c.N % 32 || c.out_mode != RVL_GEMM_OUT_BF16 || c.bias || c.rowmap || (c.flags & RVL_GEMM_FLAG_RELU) || c.split_k > 1 || partials)) {
    *err = "gemm: RVL_GEMM_FLAG_SWIGLU needs N % 32 == 0, bf16 output, no bias / ReLU / rowmap / split-k";
    return RVL_ERR_INVALID;
  }
  const void* pa = swap ? c.W : c.A;   // operand that supplies the 128-row MMA dimension
  const void* pb = swap ? c.A : c.W;   // operand that supplies MMA N
  a.M = static_cast<int>(swap ? c.N : c.M);
  a.N = static_cast<int>(swap ? c.M : c.N);
  a.bn = a.N >= 256 ? 256 : ((a.N + 15) / 16) * 16;
  a.stream_a = swap ? 1 : 0;
  a.prefetch_a = (swap && (c.flags & RVL_GEMM_FLAG_W_CONST)) ? 1 : 0;
  a.sub_stride = ((a.bn + 31) / 32) * 32;
  int sk = c.split_k < 1 ? 1 : c.split_k;
  // two A tiles per CTA tile (256 x 256) for big token-major GEMMs: the B tile is fetched once per two A tiles.
  // Weight streaming keeps 128-row tiles: more, smaller units balance better and the partial tiles stay small.
  // (measured on B200: 1-4 % faster in isolation, no gain inside the power-capped sweep, so it stays opt-in)
  a.a_tiles = 1;
  if (tn.a_tiles == 2 && !swap && a.M >= 4 * kBM && a.N >= 256) a.a_tiles = 2;
  a.tiles_m = (a.M + a.a_tiles * kBM - 1) / (a.a_tiles * kBM);
  a.tiles_n = (a.N + a.bn - 1) / a.bn;
  // stream-K for the weight-streaming orientation: needs the workspace, a single n-tile and no explicit split
  a.stream_k = 0;
  // Worth it when plain tiles would leave a ragged last wave (e.g. gate|up: 172 tiles on 148 SMs, measured 58 -> 48 us);
  // for less than one wave of tiles the fix-up traffic eats the gain (qkv: 34.3 vs 35.0 us).
  const int waves = (a.tiles_m + num_sms - 1) / num_sms;
  const bool ragged = a.tiles_m > num_sms && a.tiles_m < 0.7 * waves * num_sms;
  const bool force_sk = tn.stream_k == 2 || (c.flags & RVL_GEMM_FLAG_STREAMK);
  // many tokens: the CTA-pair weight-streaming kernel takes the GEMM (tile mode / split-k partials), no stream-K
  const bool spair_candidate = swap && a.N > 64 && a.M >= 256 && (num_sms % 2 == 0) && !c.bias && !c.rowmap && !a.relu &&
                               c.out_mode != RVL_GEMM_ADD_F32 && !force_sk && tn.spair != 0;
  if (!spair_candidate && swap && c.stream_ws && c.stream_flags && sk == 1 && a.tiles_n == 1 && !a.rowmap && !partials && a.a_tiles == 1 &&
      (ragged || force_sk)) {
    const long long total = static_cast<long long>(a.tiles_m) * a.k_blocks;
    const int ctas = static_cast<int>(total < num_sms ? total : num_sms);
    int upc = static_cast<int>((total + ctas - 1) / ctas);
    if (upc < 8 && total >= 8) upc = 8;                                   // keep pipelines worth starting
    const size_t need = static_cast<size_t>(num_sms) * a.a_tiles * 8 * 8 * 128 * 16;
    if (need <= c.stream_ws_bytes && tn.stream_k != 0) {
      a.stream_k = 1;
      a.units_per_cta = upc;
      a.ws = c.stream_ws;
      a.flags = c.stream_flags;
      a.ws_ld = a.sub_stride;
    }
  }
  a.split_k = a.stream_k ? 1 : (sk > a.k_blocks ? a.k_blocks : sk);
  while (a.split_k > 1 && ((a.k_blocks + a.split_k - 1) / a.split_k) * (a.split_k - 1) >= a.k_blocks) --a.split_k;  // no empty split
  a.split_stride = partials ? c.split_stride : 0;
  a.atomic = (a.split_k > 1 && !partials) ? 1 : 0;
  if (c.split_used) *c.split_used = a.split_k;
  // TMEM: two accumulator stages when they fit in 512 columns
  a.acc_stages = (2 * a.a_tiles * a.sub_stride <= 512) ? 2 : 1;
  a.tmem_cols = pow2_at_least(a.acc_stages * a.a_tiles * a.sub_stride);
  const int stage_bytes = a.a_tiles * kATileBytes + a.bn * kBK * 2;
  // staged TMA-store epilogue for the weight-streaming orientation (plain bf16 / fp32 / SwiGLU outputs)
  const bool out_f32 = c.out_mode != RVL_GEMM_OUT_BF16;
  const int out_es = out_f32 ? 4 : 2;
  a.staged = 0;
  if (swap && a.a_tiles == 1 && !c.bias && !c.rowmap && !a.relu && !a.atomic && c.out_mode != RVL_GEMM_ADD_F32 &&
      (reinterpret_cast<uintptr_t>(c.out) & 15) == 0 && (c.ldc * out_es) % 16 == 0 && (a.split_stride * out_es) % 16 == 0 &&
      tn.staged != 0)
    a.staged = a.swiglu ? 12 : (out_f32 ? 3 : 6);
  a.stages = (a.staged ? kSmemBudgetStaged : kSmemBudget) / stage_bytes;
  if (a.stages > kMaxStages) a.stages = kMaxStages;
  if (a.stages < 2) { *err = "gemm: tile does not fit in shared memory"; return RVL_ERR_INVALID; }
  if (env_dbg)
    fprintf(stderr, "rvl gemm: A rows=%d B rows=%d K=%d bn=%d a_tiles=%d stages=%d acc_stages=%d tmem=%d split_k=%d stream_k=%d upc=%d\n",
            a.M, a.N, a.K, a.bn, a.a_tiles, a.stages, a.acc_stages, a.tmem_cols, a.split_k, a.stream_k, a.units_per_cta);
  // CTA-pair kernel (tcgen05 cta_group::2): token-major GEMMs with >= one wave of 256 x 256 tiles
  const bool pair_ok = !swap && a.split_k == 1 && !a.stream_k && a.N >= 256 && a.M >= 1024 && (num_sms % 2 == 0);
  if (pair_ok && tn.pair != 0) {   // RVL_PAIR=0 falls back to the single-CTA kernel
    a.a_tiles = 1;
    a.bn = 256;
    a.tiles_m = (a.M + 255) / 256;
    a.tiles_n = (a.N + 255) / 256;
    {
      // m-tiles per n-sweep (L2 reuse of the A slab against the weight streaming through).  Measured interleaved on B200 at
      // 33120 tokens (tools/prefill_gemm_ab.py, groups 4 / 6 / 8 / 12 / 16 / 32, round 2): qkv 2.56 / 2.56 / 2.60 / 2.65 / 2.63 /
      // 2.54 ms, gate|up 4.54 / 4.56 / 4.52 / 4.51 / 4.45 / 4.36, o 0.974 / 0.973 / 0.973 / 0.985 / 0.991 / 0.991, down 2.49 / 2.56 /
      // 2.60 / 2.64 / 2.69 / 2.54 - everything within 3 %: 32 for the wide outputs, 8 for the small o projection, 4 for the
      // long-K down projection (a 256-row A tile is 5.6 MB there).  The differences come from DRAM traffic (power), not from
      // the tensor pipe, which is 97 - 99.7 % active either way.
      long long gm = (c.N * c.K < (32ll << 20)) ? 8 : (c.K > 2 * c.N ? 4 : 32);
      if (tn.group_m > 0) gm = tn.group_m;
      a.group_m = static_cast<int>(gm < 2 ? 2 : (gm > 64 ? 64 : gm));
    }
    a.stages = kSmemBudget / (kATileBytes + 128 * kBK * 2);
    if (a.stages > kMaxStages) a.stages = kMaxStages;
    CUtensorMap ta2, tb2;
    int rc2 = make_tmap(&ta2, pa, a.M, c.K, kBM, err);
    if (rc2) return rc2;
    rc2 = make_tmap(&tb2, pb, a.N, c.K, 128, err);
    if (rc2) return rc2;
    static bool pair_attr = false;
    if (!pair_attr) {
      if (cudaFuncSetAttribute(gemm_bf16_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
        *err = "cudaFuncSetAttribute(pair kernel) failed"; return RVL_ERR_CUDA;
      }
      pair_attr = true;
    }
    const int smem2 = a.stages * (kATileBytes + 128 * kBK * 2) + 1024 + 512 + kEpiScratchBytes;
    const int total2 = a.tiles_m * a.tiles_n;
    int grid2 = 2 * (total2 < num_sms / 2 ? total2 : num_sms / 2);
    if (env_dbg) fprintf(stderr, "rvl gemm pair: M=%d N=%d K=%d stages=%d grid=%d\n", a.M, a.N, a.K, a.stages, grid2);
    cudaError_t e2 = launch_gemm_k(gemm_bf16_pair_kernel, dim3(grid2), dim3(kGemmThreads), smem2, st, ta2, tb2, a);
    if (e2 == cudaSuccess) e2 = cudaGetLastError();
    if (e2 != cudaSuccess) { *err = std::string("gemm pair launch: ") + cudaGetErrorString(e2); return RVL_ERR_CUDA; }
    return RVL_OK;
  }
  // CTA-pair weight streaming (see gemm_stream_pair_kernel): many tokens, staged outputs, no stream-K
  if (swap && a.staged && spair_candidate && a.tiles_n == 1 && a.a_tiles == 1 && (num_sms % 2 == 0) && a.M >= 256 && tn.spair != 0) {
    CUtensorMap pa_map, pb_map, po_map;
    a.tiles_m = (a.M + 255) / 256;
    // a ragged last wave of 256-row tiles (gate|up: 86 tiles on 74 clusters): deal the k-blocks out evenly (stream-K).
    // Measured at 180 tokens: gate|up 39.7 us against 44.6 us in tile mode (42.1 us on the single-CTA stream-K kernel); for
    // less than one wave (qkv: 48 tiles, 26.0 vs 30.2 us) or nearly full waves (lm_head: 125 tiles, 49.9 vs 51.6 us) tile mode wins.
    const int clusters = num_sms / 2;
    const int pair_waves = (a.tiles_m + clusters - 1) / clusters;
    const bool pair_ragged = a.tiles_m > clusters && a.tiles_m < 0.7 * pair_waves * clusters;
    a.stream_k = 0;
    if (c.stream_ws && c.stream_flags && sk == 1 && !partials && tn.spair_streamk != 0 &&
        static_cast<size_t>(num_sms) * 8 * 8 * 128 * 16 <= c.stream_ws_bytes && (pair_ragged || tn.spair_streamk == 2)) {
      const long long total = static_cast<long long>(a.tiles_m) * a.k_blocks;
      const int cl = num_sms / 2;
      int upc = static_cast<int>((total + cl - 1) / cl);
      if (upc < 8 && total >= 8) upc = 8;
      a.stream_k = 1;
      a.units_per_cta = upc;
      a.ws = c.stream_ws;
      a.flags = c.stream_flags;
      a.ws_ld = a.sub_stride;
    }
    const int half_bytes = ((a.bn / 2) * kBK * 2 + 1023) & ~1023;
    a.stages = kSmemBudgetStaged / (kATileBytes + half_bytes);
    if (a.stages > kMaxStages) a.stages = kMaxStages;
    a.acc_stages = 2;
    a.tmem_cols = 512;
    int stage_out_bytes = kStageOutBytes;
    // Measured and dropped (round 2, tools/decode_ab.py): a half-size CTA (three 28 KB stages, 16 KB staging buffer, one 256-column
    // accumulator) so that the NEXT weight-streaming GEMM's CTAs become resident beside this kernel's under programmatic
    // dependent launch and fill their ring while this one drains - 8.16 against 7.12 ms per 7B decode step at B = 180 (7.71 vs
    // 6.85 inside CUDA graphs): three stages do not cover the HBM latency, and that costs more than the hidden prologue saves.
    int rcp = make_tmap(&pa_map, pa, a.M, c.K, kBM, err);
    if (rcp) return rcp;
    rcp = make_tmap(&pb_map, pb, a.N, c.K, a.bn / 2, err);
    if (rcp) return rcp;
    rcp = make_tmap_out(&po_map, c.out, a.swiglu ? a.M / 2 : a.M, a.N, a.split_stride > 0 ? a.split_k : 1, c.ldc, a.split_stride,
                        a.swiglu ? 64 : 128, out_f32, err);
    if (rcp) return rcp;
    static bool sp_attr = false;
    if (!sp_attr) {
      if (cudaFuncSetAttribute(gemm_stream_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
        *err = "cudaFuncSetAttribute(stream pair kernel) failed"; return RVL_ERR_CUDA;
      }
      sp_attr = true;
    }
    const int smem_sp = a.stages * (kATileBytes + half_bytes) + 1024 + 512 + stage_out_bytes;
    const int units = a.stream_k ? static_cast<int>((static_cast<long long>(a.tiles_m) * a.k_blocks + a.units_per_cta - 1) / a.units_per_cta)
                                 : a.tiles_m * a.split_k;
    const int grid_sp = 2 * (units < num_sms / 2 ? units : num_sms / 2);
    if (env_dbg) fprintf(stderr, "rvl gemm stream-pair: rows=%d tokens=%d K=%d bn=%d stages=%d split_k=%d grid=%d\n", a.M, a.N, a.K, a.bn, a.stages, a.split_k, grid_sp);
    cudaError_t esp = launch_gemm_k(gemm_stream_pair_kernel, dim3(grid_sp), dim3(kGemmThreads), smem_sp, st, pa_map, pb_map, po_map, a);
    if (esp == cudaSuccess) esp = cudaGetLastError();
    if (esp != cudaSuccess) { *err = std::string("gemm stream-pair launch: ") + cudaGetErrorString(esp); return RVL_ERR_CUDA; }
    return RVL_OK;
  }
  CUtensorMap ta, tb, tout;
  int rc = make_tmap(&ta, pa, a.M, c.K, a.a_tiles * kBM, err);
  if (rc) return rc;
  rc = make_tmap(&tb, pb, a.N, c.K, a.bn, err);
  if (rc) return rc;
  tout = ta;
  if (a.staged) {
    // features = weight rows (a.M; halved by SwiGLU), tokens = a.N
    rc = make_tmap_out(&tout, c.out, a.swiglu ? a.M / 2 : a.M, a.N, a.split_stride > 0 ? a.split_k : 1, c.ldc, a.split_stride,
                       a.swiglu ? 64 : 128, out_f32, err);
    if (rc) return rc;
  }
  const int smem = a.stages * stage_bytes + 1024 + 512 + (a.staged ? kStageOutBytes : kEpiScratchBytes);
  int grid;
  if (a.stream_k) {
    const long long total = static_cast<long long>(a.tiles_m) * a.k_blocks;
    grid = static_cast<int>((total + a.units_per_cta - 1) / a.units_per_cta);
  } else {
    const int total = a.tiles_m * a.tiles_n * a.split_k;
    grid = total < num_sms ? total : num_sms;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaError_t e2 = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaError_t e3 = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { *err = "cudaFuncSetAttribute(max dynamic smem) failed"; return RVL_ERR_CUDA; }
    attr_set = true;
  }
  cudaError_t e;
  if (a.stream_k) e = launch_gemm_k(gemm_bf16_tcgen05_kernel<1, true>, dim3(grid), dim3(kGemmThreads), smem, st, ta, tb, tout, a);
  else if (a.a_tiles == 2) e = launch_gemm_k(gemm_bf16_tcgen05_kernel<2, false>, dim3(grid), dim3(kGemmThreads), smem, st, ta, tb, tout, a);
  else e = launch_gemm_k(gemm_bf16_tcgen05_kernel<1, false>, dim3(grid), dim3(kGemmThreads), smem, st, ta, tb, tout, a);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("gemm launch: ") + cudaGetErrorString(e); return RVL_ERR_CUDA; }
  return RVL_OK;
}

// tools/ only: switch the per-CTA timestamps on/off and read them back
void gemm_debug_enable(int on) { cudaMemcpyToSymbol(g_gemm_dbg_on, &on, sizeof(int)); }
void gemm_debug_read(unsigned long long* out, int n) { cudaMemcpyFromSymbol(out, g_gemm_dbg, sizeof(unsigned long long) * n); }

}  // namespace rvl

extern "C" void rvl_debug_gemm_timestamps(int enable, unsigned long long* out, int n) {
  if (out && n > 0) rvl::gemm_debug_read(out, n);
  rvl::gemm_debug_enable(enable);
}
