// Causal varlen prefill attention on the 5th-gen tensor cores (head_dim 128, no GQA).
//
// Replaces the eager `matmul -> softmax(fp32) -> matmul` attention of transformers' Llama that the reference reaches from
// revisionllm/model/vtimellm_llama.py:79-90 (and, in round 1 of this repo, an mma.sync flash kernel: attention.cu).
//
//   S = Q K^T  : tcgen05.mma, M = 128 query rows x N = 64 keys x K = 128 dims, Q and K tiles K-major in shared memory
//                (TMA, 128-byte swizzle), S in TMEM (two 64-column buffers: QK^T of tile j + 1 overlaps the softmax of tile j)
//   softmax    : four warps, thread = query row = TMEM lane; tcgen05.ld of the row's 64 scores, causal / length mask, online
//                max and sum in fp32 registers (base 2, the scale rides in the FFMA in front of ex2), P rounded to bf16 and
//                written to shared memory in the K-major 128-byte-swizzled layout the next MMA reads
//   O += P V   : tcgen05.mma, M = 128 x N = 128 dims x K = 64 keys, V straight from its TMA tile as an MN-major operand
//                (keys are rows of the tile), O accumulates in TMEM (128 columns); when a row's running maximum grows by more
//                than 2^8 the row of O is rescaled in TMEM by its own thread (tcgen05.ld / st), otherwise the old maximum is
//                kept (P <= 256 is exact enough in bf16 and the final division by the running sum is unaffected)
//   epilogue   : O row / running sum -> bf16 -> global
// Roles per CTA (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2 - 5 softmax / correction / epilogue.
// CTAs are persistent over (sequence, head, 128-row query tile) work items; 256 TMEM columns and ~113 KB of shared memory
// per CTA, so two CTAs share an SM and one's softmax overlaps the other's loads and MMAs.
// A warp whose 32 rows lie past the sequence end, or entirely above the tile's first key (causal), skips the tile's
// exponentials and stores zeros for P.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rvl_internal.h"
#include "rvl_ptx.cuh"

namespace rvl {

namespace {

constexpr int kHD = 128;          // head_dim
constexpr int kQTile = 128;       // query rows per work item (UMMA M)
constexpr int kKTile = 64;        // keys per tile (UMMA N of QK^T, K of PV)
constexpr int kAttnThreads = 192;
constexpr int kQBytes = kQTile * kHD * 2;          // 32 KB: two [128 x 64] halves
constexpr int kKVBytes = kKTile * kHD * 2;         // 16 KB: two [64 x 64] halves
constexpr int kPBytes = kQTile * kKTile * 2;       // 16 KB: [128 x 64] bf16
constexpr int kKVStages = 2;
constexpr int kAttnSmem = 1024 + kQBytes + kKVStages * 2 * kKVBytes + kPBytes + 256;   // 115,968 B: two CTAs per SM
constexpr uint32_t kTmemCols = 256;                // S0 [0, 64) | S1 [64, 128) | O [128, 256)

struct AttnArgs {
  const int32_t* cu_seqlens;
  __nv_bfloat16* out;
  int n_seq, n_heads, n_qt;       // n_qt = query tiles per sequence (from max_seqlen)
  int only_last;                  // only the query tile that holds the last position of each sequence
  float scale_log2;               // log2(e) / sqrt(head_dim)
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MN-major B operand (V tile: rows = keys = the MMA's K dimension, 128 bytes of 64 dims per row, 128-byte swizzle; the two
// 64-dim halves of the tile are `half_bytes` apart).  Canonical layout ((8, n), (8, k)) : ((1, LBO), (8, SBO)) in 16-byte
// units: 8 keys of one 8-row swizzle atom are 128 B apart, SBO = 1024 B between 8-key groups, LBO = distance between the
// 64-dim halves.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t half_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((half_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor, D fp32, A / B bf16, A K-major, B K-major (b_mn = 0) or MN-major (b_mn = 1)
__host__ __device__ constexpr uint32_t attn_idesc(uint32_t M, uint32_t N, uint32_t b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// work item -> (sequence, head, query tile); false when the item has nothing to do (tile past the end of the sequence, or
// not the last tile under only_last).  Every role evaluates it identically.
struct Item {
  int seq, head, qt, s0, L, q0, n_tiles;
};
__device__ __forceinline__ bool decode_item(const AttnArgs& a, int item, Item& it) {
  it.qt = item % a.n_qt;
  const int r = item / a.n_qt;
  it.head = r % a.n_heads;
  it.seq = r / a.n_heads;
  it.s0 = a.cu_seqlens[it.seq];
  it.L = a.cu_seqlens[it.seq + 1] - it.s0;
  it.q0 = it.qt * kQTile;
  if (it.q0 >= it.L) return false;
  if (a.only_last && it.q0 + kQTile < it.L) return false;
  const int last_key = min(it.L, it.q0 + kQTile);         // causal: keys < q0 + 128
  it.n_tiles = (last_key + kKTile - 1) / kKTile;
  return true;
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                        // [2 halves][128 rows][128 B]
  uint8_t* sK = sQ + kQBytes;                                // [stage][2 halves][64 rows][128 B]
  uint8_t* sV = sK + kKVStages * kKVBytes;                   // same
  uint8_t* sP = sV + kKVStages * kKVBytes;                   // [128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
  uint64_t* q_full = bars;               // TMA -> MMA: the item's Q tile landed
  uint64_t* q_empty = bars + 1;          // MMA -> TMA: every QK^T of the item has read Q
  uint64_t* kv_full = bars + 2;          // [2]
  uint64_t* kv_empty = bars + 4;         // [2]  PV of the tile done: K and V slot free
  uint64_t* s_full = bars + 6;           // [2]  QK^T done: scores in TMEM
  uint64_t* s_empty = bars + 8;          // [2]  softmax has read the scores (4 warps)
  uint64_t* p_full = bars + 10;          // softmax wrote P (and rescaled O) (4 warps)
  uint64_t* o_done = bars + 11;          // PV done: O updated, P free
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = args.n_heads * kHD;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_trigger();
  pdl_wait();
  const int n_items = args.n_seq * args.n_heads * args.n_qt;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t n_q = 0, n_kv = 0;                    // items / tiles issued so far (barrier phases)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      Item it;
      if (!decode_item(args, item, it)) continue;
      if (n_q > 0) mbar_wait(q_empty, (n_q - 1) & 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, kQBytes);
        const int row = it.s0 + it.q0, col = it.head * kHD;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int g = 0; g < 2; ++g)
            tma_load_2d(sQ + h * (kQBytes / 2) + g * (64 * 128), &tmap_qkv, q_full, col + h * 64, row + g * 64);
      }
      __syncwarp();
      ++n_q;
      for (int j = 0; j < it.n_tiles; ++j, ++n_kv) {
        const int st = n_kv & 1;
        if (n_kv >= kKVStages) mbar_wait(&kv_empty[st], ((n_kv >> 1) - 1) & 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[st], 2 * kKVBytes);
          const int row = it.s0 + j * kKTile;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tma_load_2d(sK + st * kKVBytes + h * (kKVBytes / 2), &tmap_qkv, &kv_full[st], H + it.head * kHD + h * 64, row);
            tma_load_2d(sV + st * kKVBytes + h * (kKVBytes / 2), &tmap_qkv, &kv_full[st], 2 * H + it.head * kHD + h * 64, row);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_qk = attn_idesc(kQTile, kKTile, 0);
    constexpr uint32_t idesc_pv = attn_idesc(kQTile, kHD, 1);
    const uint64_t q_desc = umma_desc_k_sw128(smem_u32(sQ));
    const uint64_t p_desc = umma_desc_k_sw128(smem_u32(sP));
    uint32_t n_q = 0, n_t = 0;                     // items / tiles so far
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      Item it;
      if (!decode_item(args, item, it)) continue;
      mbar_wait(q_full, n_q & 1);
      tc_fence_after();
      // S(j) = Q K_j^T into S buffer (tile counter & 1); issued one tile ahead of the PV that consumes P(j)
      auto issue_qk = [&](uint32_t t) {
        const int st = t & 1;
        mbar_wait(&kv_full[st], (t >> 1) & 1);
        if (t >= 2) mbar_wait(&s_empty[st], ((t >> 1) - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t k_desc = umma_desc_k_sw128(smem_u32(sK + st * kKVBytes));
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k) {
            // 16 dims = 32 B inside the 128-byte swizzle row: + 2 in the (addr >> 4) field; dims 64 .. 127 live in the second half
            const uint64_t qa = q_desc + ((k >> 2) * ((kQBytes / 2) >> 4)) + 2 * (k & 3);
            const uint64_t kb = k_desc + ((k >> 2) * ((kKVBytes / 2) >> 4)) + 2 * (k & 3);
            umma_bf16(tmem_base + st * kKTile, qa, kb, idesc_qk, k > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[st]);
        }
        __syncwarp();
      };
      issue_qk(n_t);
      for (int j = 0; j < it.n_tiles; ++j) {
        const uint32_t t = n_t + j;
        if (j + 1 < it.n_tiles) issue_qk(t + 1);
        else if (elect_one()) umma_commit(q_empty);            // the item's last QK^T: Q may be overwritten once it completes
        __syncwarp();
        mbar_wait(p_full, t & 1);
        tc_fence_after();
        if (elect_one()) {
          const int st = t & 1;
          const uint64_t v_desc = umma_desc_mn_sw128(smem_u32(sV + st * kKVBytes), kKVBytes / 2);
#pragma unroll
          for (int k = 0; k < kKTile / 16; ++k)     // 16 keys = two 8-row groups = 2048 B of the V tile; 32 B of a P row
            umma_bf16(tmem_base + 2 * kKTile, p_desc + 2 * k, v_desc + k * (2048 >> 4), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(&kv_empty[st]);
          umma_commit(o_done);
        }
        __syncwarp();
      }
      n_t += it.n_tiles;
      ++n_q;
    }
  } else {
    // ------------------------------------------------------------------ softmax, correction, epilogue (thread = query row)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                       // row inside the query tile = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t p_row = smem_u32(sP) + r * 128;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    uint32_t n_t = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      Item it;
      if (!decode_item(args, item, it)) continue;
      const int q_pos = it.q0 + r;                           // position of this thread's query inside the sequence
      const int warp_first = it.q0 + quarter * 32;           // first query position of the warp
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < it.n_tiles; ++j) {
        const uint32_t t = n_t + j;
        const int st = t & 1;
        const int k0 = j * kKTile;
        mbar_wait(&s_full[st], (t >> 1) & 1);
        tc_fence_after();
        // the warp has nothing to exponentiate when all its rows lie past the sequence or above every key of the tile
        const bool skip = warp_first >= it.L || k0 > warp_first + 31;
        uint32_t pk[32];                                     // the row's 64 probabilities as bf16 pairs
        float corr = 1.f;
        if (!skip) {
          uint32_t sr[32];
          float p[64];
          tmem_ld_32x32(tmem_base + lane_addr + st * kKTile, sr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) p[i] = __uint_as_float(sr[i]);
          tmem_ld_32x32(tmem_base + lane_addr + st * kKTile + 32, sr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) p[32 + i] = __uint_as_float(sr[i]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[st]);
          if (k0 + kKTile - 1 > warp_first || k0 + kKTile > it.L) {       // diagonal or tail tile
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (k0 + i > q_pos || k0 + i >= it.L) p[i] = -INFINITY;
          }
          float mx = p[0];
#pragma unroll
          for (int i = 1; i < 64; ++i) mx = fmaxf(mx, p[i]);
          const float m_tile = mx * args.scale_log2;          // -inf when the row has no key in this tile (rows past the end)
          float m_use = m_run;
          if (m_tile > m_run + 8.f || m_run == -INFINITY) {   // lazy rescale: keep the old maximum while P stays <= 2^8
            m_use = fmaxf(m_run, m_tile);
            corr = (m_run == -INFINITY) ? 1.f : ex2f(m_run - m_use);   // first tile: O is overwritten, nothing to rescale
          }
          const float m_sub = (m_use == -INFINITY) ? 0.f : m_use;
          float rs = 0.f;
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            const float a = ex2f(fmaf(p[i], args.scale_log2, -m_sub));
            const float b = ex2f(fmaf(p[i + 1], args.scale_log2, -m_sub));
            rs += a + b;
            pk[i >> 1] = pack_bf16x2(a, b);
          }
          l_run = l_run * corr + rs;
          m_run = m_use;
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[st]);
#pragma unroll
          for (int i = 0; i < 32; ++i) pk[i] = 0u;
        }
        // PV of the previous tile must be complete before its P is overwritten and before O is rescaled
        if (j > 0) {
          mbar_wait(o_done, (t - 1) & 1);
          tc_fence_after();
        }
        if (j > 0 && __any_sync(0xffffffffu, corr != 1.f)) {
#pragma unroll 1
          for (int c = 0; c < kHD / 32; ++c) {
            uint32_t orow[32];
            tmem_ld_32x32(tmem_base + lane_addr + 2 * kKTile + c * 32, orow);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * corr);
            tmem_st_32x32(tmem_base + lane_addr + 2 * kKTile + c * 32, orow);
          }
          tmem_st_wait();
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)                           // 16-byte chunk c of the row goes to chunk (c ^ (row & 7))
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + ((static_cast<uint32_t>(c) ^ sw) << 4)), "r"(pk[4 * c]),
                       "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                       : "memory");
        fence_proxy_async();                                  // generic-proxy writes of P -> visible to the tensor core's reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // ---- epilogue: O row / l -> bf16 -> out
      const uint32_t t_last = n_t + it.n_tiles - 1;
      mbar_wait(o_done, t_last & 1);
      tc_fence_after();
      const bool live = q_pos < it.L;
      const float inv = live ? 1.f / l_run : 0.f;
      __nv_bfloat16* dst = args.out + (static_cast<long long>(it.s0) + q_pos) * H + it.head * kHD;
#pragma unroll 1
      for (int c = 0; c < kHD / 32; ++c) {
        uint32_t orow[32];
        tmem_ld_32x32(tmem_base + lane_addr + 2 * kKTile + c * 32, orow);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(orow[8 * q]) * inv, __uint_as_float(orow[8 * q + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(orow[8 * q + 2]) * inv, __uint_as_float(orow[8 * q + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(orow[8 * q + 4]) * inv, __uint_as_float(orow[8 * q + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(orow[8 * q + 6]) * inv, __uint_as_float(orow[8 * q + 7]) * inv);
            reinterpret_cast<uint4*>(dst + c * 32)[q] = o;
          }
        }
      }
      // the next item's first PV overwrites O: all four warps must be done reading it.  Their next p_full arrival is what lets
      // that PV start, and it comes after these loads (tcgen05.wait::ld above + the fence before the arrive).
      n_t += it.n_tiles;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

// false: the tensor-core kernel does not take this call (the caller falls back to the mma.sync kernel)
bool launch_attn_prefill_tc(const void* qkv, void* out, const int32_t* cu_seqlens, int n_seq, int64_t total_tokens, int max_seqlen,
                            int n_heads, int num_sms, cudaStream_t st, int only_last) {
  if (n_seq <= 0 || max_seqlen <= 0 || total_tokens <= 0) return true;
  const int H = n_heads * kHD;
  CUtensorMap tm;
  std::string err;
  if (make_tmap_bf16_2d(&tm, qkv, total_tokens, 3LL * H, 64, &err) != RVL_OK) return false;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem) != cudaSuccess) return false;
    attr = true;
  }
  AttnArgs a;
  a.cu_seqlens = cu_seqlens;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.n_seq = n_seq;
  a.n_heads = n_heads;
  a.n_qt = (max_seqlen + kQTile - 1) / kQTile;
  a.only_last = only_last;
  a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(kHD));
  const long long items = static_cast<long long>(n_seq) * n_heads * a.n_qt;
  const int grid = static_cast<int>(items < 2LL * num_sms ? items : 2LL * num_sms);
  return launch_k(attn_prefill_tc_kernel, dim3(grid), dim3(kAttnThreads), kAttnSmem, st, tm, a) == cudaSuccess;
}

}  // namespace rvl
