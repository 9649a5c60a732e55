// Attention on the 5th-gen tensor cores, two instantiations of one kernel:
//   * causal varlen prefill, head_dim 128, no GQA: replaces the eager `matmul -> softmax(fp32) -> matmul` attention of
//     transformers' Llama that the reference reaches from revisionllm/model/vtimellm_llama.py:79-90 (and, in round 1 of this
//     repo, an mma.sync flash kernel: attention.cu).  A sequence may name an external context (`seq_pos0` / `seq_ctx_row` of
//     rvl_prefill: the prompt prefix shared by the batch, projected once): its first keys are then rows of another part of
//     the packed stream, and the K / V tiles are loaded as two 32-key boxes each so that a tile may straddle the two ranges.
//   * non-causal, head_dim 96, fixed Tq / Tk, optional key-padding mask and shared key / value sequences: the
//     nn.MultiheadAttention core of the stage-2 ClipEncoder (revisionllm/model/adapter/transformer.py:216-217, :288-289).
//     The tiles keep the 128-dim geometry: the second 64-column box of a head holds its dims 64 .. 95 and 32 columns of the
//     neighbouring head (or zeros past the matrix); QK^T stops after six 16-dim steps, so they are never read, and the 32
//     surplus columns of O are never stored.
//
// Work item = (sequence, head, PAIR of 128-row query tiles).  The two query tiles are two independent "slots" of the CTA
// that share every K / V tile (the later tile needs a superset of the keys of the earlier one):
//   S = Q K^T  : tcgen05.mma, M = 128 query rows x N = 64 keys x K = 128 dims, Q and K tiles K-major in shared memory
//                (TMA, 128-byte swizzle), S in TMEM (two 64-column buffers per slot: QK^T of tile j + 1 overlaps the softmax
//                of tile j)
//   softmax    : four warps per slot, thread = query row = TMEM lane; tcgen05.ld of the row's 64 scores, causal / length mask,
//                online max and sum in fp32 registers (base 2, the scale rides in the FFMA in front of ex2), P rounded to
//                bf16 and written to shared memory in the K-major 128-byte-swizzled layout the next MMA reads
//   O += P V   : tcgen05.mma, M = 128 x N = 128 dims x K = 64 keys, V straight from its TMA tile as an MN-major operand
//                (keys are rows of the tile), O accumulates in TMEM (128 columns per slot); when a row's running maximum
//                grows by more than 2^8 the row of O is rescaled in TMEM by its own thread (tcgen05.ld / st), otherwise the
//                old maximum is kept (P <= 256 is exact enough in bf16, the final division by the running sum is unaffected)
//   epilogue   : O row / running sum -> bf16 -> the warp's own rows of the P buffer -> coalesced 16-byte global stores
// Roles (one CTA per SM, persistent over the items): a TMA producer warp (runs up to four K / V tiles and one item's Q tiles
// ahead), an MMA issuer warp (polls: whichever of the four possible MMAs - QK^T / PV of either slot - has its inputs ready
// goes out), four softmax / correction / epilogue warps per slot.
// At L = 184 the kernel is bound by HBM, not by the tensor pipe: a layer reads the 814 MB qkv stream once and writes 271 MB
// (46 FLOP per byte); what the structure buys is enough loads in flight (224 KB of shared memory per SM) to cover the
// ~1.5 us a tile takes from issue to arrival.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rvl_internal.h"
#include "rvl_ptx.cuh"

namespace rvl {

// tools/attn_timeline.py: globaltimer stamps of CTA 0 - [role][event index][kind]; role 0 producer (kind 0 = K/V tile issued),
// role 1 MMA (0 / 1 = QK issued for slot 0 / 1, 2 / 3 = PV issued), role 2 first softmax warp of slot 0 and role 3 of slot 1
// (0 = scores ready, 1 = exponentials done, 2 = previous PV seen done, 3 = P published); events are indexed by the CTA's
// running K/V tile counter (roles 0, 1) or the slot's running tile counter (roles 2, 3)
constexpr int kDbgEvents = 256;
__device__ unsigned long long g_attn_dbg[4][kDbgEvents][4];
__device__ int g_attn_dbg_on = 0;
__device__ __forceinline__ void dbg_stamp(int role, uint32_t idx, int kind) {
  if (g_attn_dbg_on && blockIdx.x == 0 && idx < kDbgEvents && (threadIdx.x & 31) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_attn_dbg[role][idx][kind] = t;
  }
}

namespace {

constexpr int kHD = 128;          // dims per tile row (head_dim 128; head_dim 96 uses the first 96)
constexpr int kQTile = 128;       // query rows per slot (UMMA M)
constexpr int kKTile = 64;        // keys per tile (UMMA N of QK^T, K of PV)
constexpr int kSlots = 2;
// warps 0 - 3: softmax of slot 0, 4 - 7: slot 1 (TMEM lane quarter = warp % 4), 8: TMA producer, 9: unused, 10 / 11: MMA issuer
// of slot 0 / 1.  The two polling warps sit on schedulers 2 and 3 on purpose: in the common item (L = 184: the later query tile
// holds rows 128 - 183) the softmax warps that do the work are quarters 0 and 1, i.e. schedulers 0 and 1.
constexpr int kAttnThreads = 12 * 32;
constexpr int kProducerWarp = 8, kMmaWarp = 11;
constexpr int kQBytes = kQTile * kHD * 2;          // 32 KB: two [128 x 64] halves
constexpr int kKVBytes = kKTile * kHD * 2;         // 16 KB: two [64 x 64] halves (K; the same again for V)
constexpr int kPBytes = kQTile * kKTile * 2;       // 16 KB: [128 x 64] bf16
constexpr int kKVStages = 4;
constexpr int kAttnSmem = 1024 + kSlots * (kQBytes + kPBytes) + kKVStages * 2 * kKVBytes + 256;   // 230,656 B: one CTA per SM
constexpr uint32_t kTmemCols = 512;                // per slot: S0 [0, 64) | S1 [64, 128) | O [128, 256)

struct AttnArgs {
  const int32_t* cu_seqlens;      // causal: rows [cu[i], cu[i + 1]) of the packed stream are sequence i
  const int32_t* seq_pos0;        // causal, optional: the sequence's own rows start at this position ...
  const int32_t* seq_ctx_row;     // ... and its positions [0, pos0) are the rows [ctx_row, ctx_row + pos0) (multiple of 32)
  const int32_t* kv_seq_idx;      // non-causal, optional: key / value sequence of query sequence i
  const float* key_mask;          // non-causal, optional: fp32 [n_kv_seq, Tk], 0 = padded key
  __nv_bfloat16* out;
  long long out_stride;           // elements between output rows
  int n_seq, n_heads, n_pairs;    // n_pairs = pairs of query tiles per sequence (from the longest sequence)
  int Tq, Tk;                     // non-causal: rows per query / key sequence
  int q_col0, k_col0, v_col0;     // column of head 0 in the q / k / v tensor maps
  int kv_box32;                   // K / V tensor maps have 32-row boxes (external context: a tile = two boxes)
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

// work item -> (sequence, head, pair of query tiles); false when the item has nothing to do.  Every role evaluates it
// identically.  nt[s] = K / V tiles slot s consumes (0: the slot sits this item out); the item loads max(nt) tiles.
// Query rows are counted from the sequence's own first row; keys by POSITION (context first: P positions, then the own rows).
struct Item {
  int seq, head, kv_seq;
  int q_row0, kv_row0, ctx_row0;  // first own row of the queries / keys; first row of the external context
  int P, Lq, Lk;                  // context positions, query rows, key positions (causal: Lk = P + Lq)
  int q0[kSlots], nt[kSlots], n_tiles;
};
// `flip`: which slot takes the earlier query tile alternates from one item of the CTA to the next - the slot that ran the short
// chain (earlier tile: fewer keys) released its Q buffer early, so the NEXT item's long chain can have its Q loaded while the
// current item is still computing.
template <bool kCausal>
__device__ __forceinline__ bool decode_item(const AttnArgs& a, int item, int flip, Item& it) {
  // the pair index is the SLOW index and runs backwards: every CTA of the grid-stride loop gets the same mix of long (late
  // query tiles: more keys) and short items, the long ones first
  const int per_pair = a.n_seq * a.n_heads;
  const int pp = a.n_pairs - 1 - item / per_pair;
  const int r = item - (item / per_pair) * per_pair;
  it.head = r % a.n_heads;
  it.seq = r / a.n_heads;
  if (kCausal) {
    const int s0 = __ldg(a.cu_seqlens + it.seq);
    it.Lq = __ldg(a.cu_seqlens + it.seq + 1) - s0;
    it.q_row0 = it.kv_row0 = s0;
    it.P = a.seq_pos0 ? __ldg(a.seq_pos0 + it.seq) : 0;
    it.ctx_row0 = a.seq_ctx_row ? __ldg(a.seq_ctx_row + it.seq) : 0;
    it.Lk = it.P + it.Lq;
    it.kv_seq = it.seq;
  } else {
    it.kv_seq = a.kv_seq_idx ? __ldg(a.kv_seq_idx + it.seq) : it.seq;
    it.q_row0 = it.seq * a.Tq;
    it.kv_row0 = it.kv_seq * a.Tk;
    it.ctx_row0 = 0;
    it.P = 0;
    it.Lq = a.Tq;
    it.Lk = a.Tk;
  }
  it.n_tiles = 0;
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    const int q0 = (pp * kSlots + (s ^ flip)) * kQTile;
    it.q0[s] = q0;
    const bool on = q0 < it.Lq && !(a.only_last && q0 + kQTile < it.Lq);
    const int keys = kCausal ? it.P + min(it.Lq, q0 + kQTile) : it.Lk;         // causal: key positions < P + q0 + 128
    it.nt[s] = on ? (keys + kKTile - 1) / kKTile : 0;
    it.n_tiles = max(it.n_tiles, it.nt[s]);
  }
  return it.n_tiles > 0;
}

template <int HD, bool kCausal>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
               const __grid_constant__ CUtensorMap tmap_v, const AttnArgs args) {
  static_assert(HD == 128 || HD == 96, "head_dim 128 (Llama) or 96 (ClipEncoder)");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                        // [slot][2 halves][128 rows][128 B]
  uint8_t* sP = sQ + kSlots * kQBytes;                       // [slot][128 rows][128 B]
  uint8_t* sK = sP + kSlots * kPBytes;                       // [stage][2 halves][64 rows][128 B]
  uint8_t* sV = sK + kKVStages * kKVBytes;                   // same
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKVStages * kKVBytes);
  uint64_t* kv_full = bars;              // [4]  TMA -> MMA
  uint64_t* kv_empty = bars + 4;         // [4]  both slots' PV of the tile done (count 2)
  uint64_t* q_full = bars + 8;           // [slot]
  uint64_t* q_empty = bars + 10;         // [slot] every QK^T of the item has read the slot's Q
  uint64_t* s_full = bars + 12;          // [slot][2]  QK^T done: scores in TMEM
  uint64_t* s_empty = bars + 16;         // [slot][2]  softmax has read the scores (4 warps)
  uint64_t* p_full = bars + 20;          // [slot] softmax wrote P (and rescaled O) (4 warps)
  uint64_t* o_done = bars + 22;          // [slot] PV done: O updated, P free
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int i = 0; i < kKVStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);
    }
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&o_done[s], 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[s * 2 + i], 1);
        mbar_init(&s_empty[s * 2 + i], 4);
      }
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_trigger();
  pdl_wait();
  const int n_items = args.n_seq * args.n_heads * args.n_pairs;

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t n_kv = 0, n_it = 0;                   // K/V tiles issued so far; items so far
    uint32_t n_q[kSlots] = {0, 0};                 // items each slot took part in so far
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      Item it;
      if (!decode_item<kCausal>(args, item, n_it & 1, it)) continue;
      ++n_it;
      const int s_long_dbg = it.nt[1] > it.nt[0] ? 1 : 0;
      auto load_q = [&](int s) {
        if (it.nt[s] == 0) return;
        if (n_q[s] > 0) mbar_wait(&q_empty[s], (n_q[s] - 1) & 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[s], kQBytes);
          const int row = it.q_row0 + it.q0[s], col = args.q_col0 + it.head * HD;
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int g = 0; g < 2; ++g)
              tma_load_2d(sQ + s * kQBytes + h * (kQBytes / 2) + g * (64 * 128), &tmap_q, &q_full[s], col + h * 64, row + g * 64);
        }
        __syncwarp();
        dbg_stamp(0, 128 + n_it, s == s_long_dbg ? 1 : 2);      // Q issued (index 128 + item): kind 1 long chain, 2 short chain
        ++n_q[s];
      };
      // in the order the buffers come free: the Q buffer of the slot with the long chain (it ran the short one last time),
      // two K/V tiles, the other Q buffer, the remaining tiles
      const int s_long = it.nt[1] > it.nt[0] ? 1 : 0;
      load_q(s_long);
      for (int j = 0; j < it.n_tiles; ++j, ++n_kv) {
        if (j == 2) load_q(s_long ^ 1);
        const int st = n_kv % kKVStages;
        const uint32_t use = n_kv / kKVStages;
        if (use > 0) mbar_wait(&kv_empty[st], (use - 1) & 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[st], 2 * kKVBytes);
          const int kc = args.k_col0 + it.head * HD, vc = args.v_col0 + it.head * HD;
          if (kCausal && args.kv_box32) {
            // two 32-key boxes per tile and 64-dim half: positions below P are rows of the context, the others own rows
            if (it.P & 31) {
              printf("rvl attention: seq_pos0 = %d is not a multiple of 32\n", it.P);
              __trap();
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int pos = j * kKTile + g * 32;
              const int row = pos < it.P ? it.ctx_row0 + pos : it.kv_row0 + pos - it.P;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                tma_load_2d(sK + st * kKVBytes + h * (kKVBytes / 2) + g * (32 * 128), &tmap_k, &kv_full[st], kc + h * 64, row);
                tma_load_2d(sV + st * kKVBytes + h * (kKVBytes / 2) + g * (32 * 128), &tmap_v, &kv_full[st], vc + h * 64, row);
              }
            }
          } else {
            const int row = it.kv_row0 + j * kKTile;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              tma_load_2d(sK + st * kKVBytes + h * (kKVBytes / 2), &tmap_k, &kv_full[st], kc + h * 64, row);
              tma_load_2d(sV + st * kKVBytes + h * (kKVBytes / 2), &tmap_v, &kv_full[st], vc + h * 64, row);
            }
          }
        }
        __syncwarp();
        dbg_stamp(0, n_kv, 0);
      }
      if (it.n_tiles <= 2) load_q(s_long ^ 1);
    }
  } else if (warp == kMmaWarp || warp == kMmaWarp - 1) {
    // ------------------------------------------------------------------ MMA issuer of one slot (two warps, one per slot)
    constexpr uint32_t idesc_qk = attn_idesc(kQTile, kKTile, 0);
    constexpr uint32_t idesc_pv = attn_idesc(kQTile, kHD, 1);
    const int s = warp - (kMmaWarp - 1);           // the slot this warp serves
    uint32_t n_kv = 0, idle = 0, n_it = 0;         // K/V tiles of finished items; items so far
    uint32_t n_q = 0, n_ts = 0;                    // the slot's items / tiles so far
    const uint64_t q_desc = umma_desc_k_sw128(smem_u32(sQ + s * kQBytes));
    const uint64_t p_desc = umma_desc_k_sw128(smem_u32(sP + s * kPBytes));
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      Item it;
      if (!decode_item<kCausal>(args, item, n_it & 1, it)) continue;
      ++n_it;
      if (s == 0) dbg_stamp(1, 128 + n_it, 0);                  // the MMA warp starts polling for this item
      const int n_my = it.nt[s], n_other = it.nt[s ^ 1];
      int qk_next = 0, pv_next = 0;
      // QK^T of a tile goes out as soon as its K tile has landed, the slot's Q is there and the softmax warps have drained the
      // S buffer it writes (two tiles back); PV as soon as its P is published.  The conditions are POLLED, four lanes testing
      // the four barriers at once (non-blocking test_wait): with a fixed issue order the timeline showed PV(t) waiting ~0.9 us
      // behind the K/V load of tile t + 1, and with one warp polling both slots lane by lane ~0.5 us between two issues.
      while (pv_next < n_my) {
        const uint32_t seq_q = n_kv + qk_next, ts_q = n_ts + qk_next;
        const int st_q = seq_q % kKVStages, sb_q = ts_q & 1;
        const uint32_t ts_p = n_ts + pv_next;
        bool ok = true;
        if (lane == 0) ok = qk_next > 0 || mbar_test_wait(&q_full[s], n_q & 1);
        else if (lane == 1) ok = qk_next < n_my && mbar_test_wait(&kv_full[st_q], (seq_q / kKVStages) & 1);
        else if (lane == 2) ok = ts_q < 2 || mbar_test_wait(&s_empty[s * 2 + sb_q], ((ts_q >> 1) - 1) & 1);
        else if (lane == 3) ok = pv_next < qk_next && mbar_test_wait(&p_full[s], ts_p & 1);
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        const bool qk_ready = qk_next < n_my && (m & 7u) == 7u;
        const bool pv_ready = (m & 8u) != 0;
        if (pv_ready) {                                          // PV first: it is what the softmax warps are waiting for
          const uint32_t seq = n_kv + pv_next;
          const int st = seq % kKVStages;
          tc_fence_after();
          if (elect_one()) {
            const uint64_t v_desc = umma_desc_mn_sw128(smem_u32(sV + st * kKVBytes), kKVBytes / 2);
#pragma unroll
            for (int k = 0; k < kKTile / 16; ++k)     // 16 keys = two 8-row groups = 2048 B of the V tile; 32 B of a P row
              umma_bf16(tmem_base + s * 256 + 2 * kKTile, p_desc + 2 * k, v_desc + k * (2048 >> 4), idesc_pv, (pv_next > 0 || k > 0) ? 1u : 0u);
            umma_commit(&kv_empty[st]);
            // kv_empty counts two arrivals (one per slot): a tile the other slot does not use is released on its behalf by
            // the same MMAs' completion - never by a plain arrive, which could land in the phase of the stage's previous tile
            if (pv_next >= n_other) umma_commit(&kv_empty[st]);
            umma_commit(&o_done[s]);
          }
          __syncwarp();
          dbg_stamp(1, seq, 2 + s);
          ++pv_next;
        }
        if (qk_ready) {
          tc_fence_after();
          if (elect_one()) {
            const uint64_t k_desc = umma_desc_k_sw128(smem_u32(sK + st_q * kKVBytes));
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {
              // 16 dims = 32 B inside the 128-byte swizzle row: + 2 in the (addr >> 4) field; dims 64 .. 127: second half
              const uint64_t qa = q_desc + ((k >> 2) * ((kQBytes / 2) >> 4)) + 2 * (k & 3);
              const uint64_t kb = k_desc + ((k >> 2) * ((kKVBytes / 2) >> 4)) + 2 * (k & 3);
              umma_bf16(tmem_base + s * 256 + sb_q * kKTile, qa, kb, idesc_qk, k > 0 ? 1u : 0u);
            }
            umma_commit(&s_full[s * 2 + sb_q]);
            if (qk_next + 1 == n_my) umma_commit(&q_empty[s]);    // the item's last QK^T of the slot: its Q may be overwritten
          }
          __syncwarp();
          dbg_stamp(1, seq_q, s);
          ++qk_next;
        }
        if (pv_ready || qk_ready) {
          idle = 0;
        } else {
          __nanosleep(20);                                        // leave the issue slots of this scheduler to the softmax warps
          if (++idle > (1u << 24)) mbar_timeout(nullptr, 0xa77);
        }
      }
      n_kv += it.n_tiles;
      n_ts += n_my;
      if (n_my) ++n_q;
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ softmax, correction, epilogue (thread = query row)
    const int slot = warp >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                       // row inside the query tile = TMEM lane
    const uint32_t lane_addr = (static_cast<uint32_t>(quarter * 32) << 16) + slot * 256;
    const uint32_t p_base = smem_u32(sP + slot * kPBytes);
    const uint32_t p_row = p_base + r * 128;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const bool stamp = (warp & 3) == 0;
    uint64_t* my_s_full = s_full + slot * 2;
    uint64_t* my_s_empty = s_empty + slot * 2;
    uint32_t n_t = 0, n_it = 0;                              // the slot's tiles so far; the CTA's items so far
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      Item it;
      if (!decode_item<kCausal>(args, item, n_it & 1, it)) continue;
      ++n_it;
      const int n_my = it.nt[slot];
      if (n_my == 0) continue;
      const int q0 = it.q0[slot];
      const int q_idx = q0 + r;                              // this thread's query row inside the sequence
      const int warp_first = q0 + quarter * 32;              // first query row of the warp
      const int pos_first = it.P + warp_first;               // its position
      const int lim_row = kCausal ? it.P + min(q_idx, it.Lq - 1) : it.Lk - 1;   // last key position this row may see
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_my; ++j) {
        const uint32_t t = n_t + j;
        const int sb = t & 1;
        const int k0 = j * kKTile;
        mbar_wait(&my_s_full[sb], (t >> 1) & 1);
        tc_fence_after();
        if (stamp) dbg_stamp(2 + slot, t, 0);
        // the warp has nothing to exponentiate when all its rows lie past the sequence or above every key of the tile
        const bool skip = warp_first >= it.Lq || (kCausal && k0 > pos_first + 31);
        uint32_t pk[32];                                     // the row's 64 probabilities as bf16 pairs
        float corr = 1.f;
        if (!skip) {
          uint32_t sr0[32], sr1[32];
          float p[64];
          tmem_ld_32x32(tmem_base + lane_addr + sb * kKTile, sr0);
          tmem_ld_32x32(tmem_base + lane_addr + sb * kKTile + 32, sr1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) { p[i] = __uint_as_float(sr0[i]); p[32 + i] = __uint_as_float(sr1[i]); }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&my_s_empty[sb]);
          // keys k0 + i with i > lim are masked for this row (causal, sequence end); groups of 16 keys that are masked for EVERY
          // row of the warp (i > lim_w, warp-uniform) are left out of the maximum and of the exponentials altogether - on a
          // diagonal tile that is up to half of the work of a unit (MUFU) that the two slots' warps of a scheduler share
          const int lim = lim_row - k0;
          const int lim_w = kCausal ? min(min(pos_first + 31, it.Lk - 1) - k0, kKTile - 1) : min(it.Lk - 1 - k0, kKTile - 1);
          if (lim_w < kKTile - 1 || (kCausal && k0 + kKTile - 1 > pos_first)) {
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (i > lim) p[i] = -INFINITY;
          }
          if (!kCausal && args.key_mask) {                    // key padding: one bit per key of the tile, the same for every row
            const float* mrow = args.key_mask + static_cast<long long>(it.kv_seq) * it.Lk + k0;
            const uint32_t lo = __ballot_sync(0xffffffffu, k0 + lane < it.Lk && __ldg(mrow + lane) != 0.f);
            const uint32_t hi = __ballot_sync(0xffffffffu, k0 + 32 + lane < it.Lk && __ldg(mrow + 32 + lane) != 0.f);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (!((lo >> i) & 1u)) p[i] = -INFINITY;
              if (!((hi >> i) & 1u)) p[32 + i] = -INFINITY;
            }
          }
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains (few warps per scheduler: latency is exposed)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (16 * g <= lim_w) {
#pragma unroll
              for (int i = 16 * g; i < 16 * g + 16; i += 4) {
                mx4[0] = fmaxf(mx4[0], p[i]); mx4[1] = fmaxf(mx4[1], p[i + 1]);
                mx4[2] = fmaxf(mx4[2], p[i + 2]); mx4[3] = fmaxf(mx4[3], p[i + 3]);
              }
            }
          }
          const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
          const float m_tile = mx * args.scale_log2;          // -inf when the row has no key in this tile (rows past the end)
          float m_use = m_run;
          if (m_tile > m_run + 8.f || m_run == -INFINITY) {   // lazy rescale: keep the old maximum while P stays <= 2^8
            m_use = fmaxf(m_run, m_tile);
            corr = (m_run == -INFINITY) ? 1.f : ex2f(m_run - m_use);   // first tile: O is overwritten, nothing to rescale
          }
          const float m_sub = (m_use == -INFINITY) ? 0.f : m_use;
          float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (16 * g <= lim_w) {
#pragma unroll
              for (int i = 16 * g; i < 16 * g + 16; i += 2) {
                const float a = ex2f(fmaf(p[i], args.scale_log2, -m_sub));
                const float b = ex2f(fmaf(p[i + 1], args.scale_log2, -m_sub));
                rs4[(i >> 1) & 3] += a + b;
                pk[i >> 1] = pack_bf16x2(a, b);
              }
            } else {
#pragma unroll
              for (int i = 8 * g; i < 8 * g + 8; ++i) pk[i] = 0u;
            }
          }
          l_run = l_run * corr + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
          m_run = m_use;
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&my_s_empty[sb]);
#pragma unroll
          for (int i = 0; i < 32; ++i) pk[i] = 0u;
        }
        if (stamp) dbg_stamp(2 + slot, t, 1);
        // PV of the previous tile must be complete before its P is overwritten and before O is rescaled
        if (j > 0) {
          mbar_wait(&o_done[slot], (t - 1) & 1);
          tc_fence_after();
        }
        if (stamp) dbg_stamp(2 + slot, t, 2);
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
        if (lane == 0) mbar_arrive(&p_full[slot]);
        if (stamp) dbg_stamp(2 + slot, t, 3);
      }
      // ---- epilogue: O row / l -> bf16 -> out
      const uint32_t t_last = n_t + n_my - 1;
      mbar_wait(&o_done[slot], t_last & 1);
      tc_fence_after();
      // Each thread owns one 256-byte output row: stored directly, a warp-level 16-byte store touches 32 different lines.
      // Instead the warp's 32 rows go through ITS OWN 4 KB of the slot's P buffer (free: the last PV has completed), 64 dims
      // at a time, and leave as 4 rows x 128 B per warp-level store.
      const bool live = q_idx < it.Lq;
      const float inv = live ? 1.f / l_run : 0.f;
      const int rows_valid = min(32, it.Lq - warp_first);                 // rows of this warp inside the sequence (<= 0: none)
      __nv_bfloat16* dst_warp = args.out + (static_cast<long long>(it.q_row0) + warp_first) * args.out_stride + it.head * HD;
      const uint32_t p_warp = p_base + quarter * 32 * 128;
      if (rows_valid > 0) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t o0[32], o1[32];
          tmem_ld_32x32(tmem_base + lane_addr + 2 * kKTile + half * 64, o0);
          tmem_ld_32x32(tmem_base + lane_addr + 2 * kKTile + half * 64 + 32, o1);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t* src = c < 4 ? o0 + 8 * c : o1 + 8 * (c - 4);
            const uint32_t w0 = pack_bf16x2(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
            const uint32_t w1 = pack_bf16x2(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
            const uint32_t w2 = pack_bf16x2(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
            const uint32_t w3 = pack_bf16x2(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + ((static_cast<uint32_t>(c) ^ sw) << 4)), "r"(w0), "r"(w1),
                         "r"(w2), "r"(w3)
                         : "memory");
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3), cc = lane & 7;             // row inside the warp's block, 16-byte chunk of the half row
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(p_warp + rr * 128 + ((static_cast<uint32_t>(cc) ^ static_cast<uint32_t>(rr & 7)) << 4)));
            if (rr < rows_valid && half * 64 + cc * 8 < HD)
              *reinterpret_cast<uint4*>(dst_warp + static_cast<long long>(rr) * args.out_stride + half * 64 + cc * 8) = v;
          }
          __syncwarp();
        }
      }
      // The next item's first PV of this slot overwrites O: all four warps must be done reading it.  Their next p_full arrival
      // is what lets that PV start, and it comes after these loads (tcgen05.wait::ld above + the fence before the arrive).
      n_t += n_my;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

template <int HD, bool kCausal>
static bool launch_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnArgs& a, int num_sms, cudaStream_t st) {
  static bool attr = false;                 // one flag per instantiation
  if (!attr) {
    if (cudaFuncSetAttribute(attn_tc_kernel<HD, kCausal>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem) != cudaSuccess) return false;
    attr = true;
  }
  const long long items = static_cast<long long>(a.n_seq) * a.n_heads * a.n_pairs;
  const int grid = static_cast<int>(items < num_sms ? items : num_sms);
  return launch_k(attn_tc_kernel<HD, kCausal>, dim3(grid), dim3(kAttnThreads), kAttnSmem, st, tq, tk, tv, a) == cudaSuccess;
}

// false: the tensor-core kernel does not take this call (the caller falls back to the mma.sync kernel)
bool launch_attn_prefill_tc(const void* qkv, void* out, const int32_t* cu_seqlens, int n_seq, int64_t total_tokens, int max_seqlen,
                            int n_heads, int num_sms, cudaStream_t st, int only_last, const int32_t* seq_pos0,
                            const int32_t* seq_ctx_row) {
  if (n_seq <= 0 || max_seqlen <= 0 || total_tokens <= 0) return true;
  if ((seq_pos0 != nullptr) != (seq_ctx_row != nullptr)) return false;
  const int H = n_heads * kHD;
  CUtensorMap tm, tm_kv;
  std::string err;
  if (make_tmap_bf16_2d(&tm, qkv, total_tokens, 3LL * H, 64, &err) != RVL_OK) return false;
  tm_kv = tm;
  if (seq_pos0 && make_tmap_bf16_2d(&tm_kv, qkv, total_tokens, 3LL * H, 32, &err) != RVL_OK) return false;
  AttnArgs a{};
  a.cu_seqlens = cu_seqlens;
  a.seq_pos0 = seq_pos0;
  a.seq_ctx_row = seq_ctx_row;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.out_stride = H;
  a.n_seq = n_seq;
  a.n_heads = n_heads;
  a.n_pairs = (max_seqlen + kSlots * kQTile - 1) / (kSlots * kQTile);      // max_seqlen counts positions: an upper bound of the own rows
  a.q_col0 = 0;
  a.k_col0 = H;
  a.v_col0 = 2 * H;
  a.kv_box32 = seq_pos0 ? 1 : 0;
  a.only_last = only_last;
  a.scale_log2 = 1.4426950408889634f / sqrtf(128.f);
  return launch_tc<128, true>(tm, tm_kv, tm_kv, a, num_sms, st);
}

// nn.MultiheadAttention core of the ClipEncoder (head_dim 96, non-causal); false: not taken (operands not 16-byte aligned)
bool launch_mha96_tc(const void* q, long long q_stride, const void* k, long long k_stride, const void* v, long long v_stride, void* out,
                     long long out_stride, int n_seq, int n_kv_seq, int n_heads, int Tq, int Tk, const int32_t* kv_seq_idx,
                     const float* key_mask, int num_sms, cudaStream_t st) {
  if (n_seq <= 0 || Tq <= 0 || Tk <= 0) return true;
  if (num_sms <= 0 || n_kv_seq <= 0 || out_stride % 8 || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
  const long long D = 96LL * n_heads;
  if (q_stride < D || k_stride < D || v_stride < D) return false;
  CUtensorMap tq, tk, tv;
  std::string err;
  // the column extent stops at the last head: what a head's second 64-column box reads past it comes back as zeros
  if (make_tmap_bf16_2d(&tq, q, static_cast<int64_t>(n_seq) * Tq, D, 64, &err, q_stride) != RVL_OK) return false;
  if (make_tmap_bf16_2d(&tk, k, static_cast<int64_t>(n_kv_seq) * Tk, D, 64, &err, k_stride) != RVL_OK) return false;
  if (make_tmap_bf16_2d(&tv, v, static_cast<int64_t>(n_kv_seq) * Tk, D, 64, &err, v_stride) != RVL_OK) return false;
  AttnArgs a{};
  a.kv_seq_idx = kv_seq_idx;
  a.key_mask = key_mask;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.out_stride = out_stride;
  a.n_seq = n_seq;
  a.n_heads = n_heads;
  a.n_pairs = (Tq + kSlots * kQTile - 1) / (kSlots * kQTile);
  a.Tq = Tq;
  a.Tk = Tk;
  a.scale_log2 = 1.4426950408889634f / sqrtf(96.f);
  return launch_tc<96, false>(tq, tk, tv, a, num_sms, st);
}

}  // namespace rvl

extern "C" void rvl_debug_attn_timestamps(int enable, unsigned long long* out, int n) {
  if (out && n > 0) cudaMemcpyFromSymbol(out, rvl::g_attn_dbg, sizeof(unsigned long long) * n);
  cudaMemcpyToSymbol(rvl::g_attn_dbg_on, &enable, sizeof(int));
}
