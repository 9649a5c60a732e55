// bf16 GEMM on the 5th-gen tensor cores:  D[m][n] = sum_k A[m][k] * B[n][k]   (both operands K-major).
//
// Replaces every nn.Linear on the ReVisionLLM scoring path (cuBLAS through torch in the reference):
// q/k/v/o, gate/up/down, lm_head (transformers Llama reached from
// revisionllm/model/vtimellm_llama.py:79-90), mm_projector (revisionllm/model/vtimellm_arch.py:42,125)
// and the ClipEncoder linears (revisionllm/model/adapter/transformer.py).
//
// Design (one CTA per SM, persistent over output tiles, warp-specialised):
//   warp 0 / one lane : TMA producer  - cp.async.bulk.tensor 2D loads of a 128 x 64 A tile and a
//                                      BN x 64 B tile per pipeline stage (128-byte swizzle)
//   warp 1 / one lane : MMA issuer    - tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16 x 4 per stage,
//                                      fp32 accumulators in TMEM, double buffered (2 x BN columns)
//   warps 2..5        : epilogue      - tcgen05.ld 32x32b.x32 -> registers -> bias / ReLU / residual /
//                                      row scatter -> global
//   smem ring full/empty mbarriers (TMA <-> MMA) and tmem full/empty mbarriers (MMA <-> epilogue).
// The A operand always supplies the 128-row MMA dimension.  For token-major outputs with many tokens
// the activations are A and the weights B; for small token counts (decode, lm_head on last rows) the
// host swaps the roles so the weight streams through the 128-row slot once and the tokens sit in N
// (epilogue then stores transposed).
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
constexpr int kGroupM = 16;  // tile rasterisation: 16 m-tiles share each n-tile sweep (L2 reuse)

struct GemmArgs {
  int M, N, K;          // A rows, B rows, reduction
  int tiles_m, tiles_n;
  int k_blocks;         // ceil(K / 64)
  int split_k;
  int bm;               // A rows per tile (TMA box rows, <= 128; the MMA always spans 128 smem rows)
  int bn;               // B rows per tile = MMA N (multiple of 16, <= 256)
  int stages;           // smem ring depth
  int tmem_cols;        // power of two >= 2 * bn
  int stream_a;         // A is a weight streamed once from HBM: L2 evict-first for A, evict-last for B
  int prefetch;         // stream_a: k-blocks of A prefetched into L2 ahead of the TMA loads
  long long split_stride;  // > 0: split-k partial s goes to out + s * split_stride (plain stores, no atomics)
  long long ldc;
  void* out;
  const __nv_bfloat16* bias;  // indexed by the feature dimension
  const int* rowmap;          // indexed by the token dimension (optional)
  int mode;                   // RVL_GEMM_OUT_*
  int relu;
  int transposed;             // 0: tokens = A rows (m), features = B rows (n); 1: the other way round
  int atomic;                 // split-k partial sums: atomicAdd into fp32 out
};

constexpr int kMaxStages = 12;
constexpr int kATileBytes = kBM * kBK * 2;           // 16 KB slot; TMA fills bm rows of it
constexpr int kSmemBudget = 227 * 1024 - 1024 /*align slack*/ - 512 /*barriers*/;

__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int& m_blk, int& n_blk) {
  const int group = kGroupM * tiles_n;
  const int gid = t / group;
  const int first_m = gid * kGroupM;
  const int gsz = min(tiles_m - first_m, kGroupM);
  const int r = t - gid * group;
  m_blk = first_m + r % gsz;
  n_blk = r / gsz;
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmArgs args) {
  const int kStages = args.stages;
  const int BN = args.bn;
  const int b_tile_bytes = BN * kBK * 2;
  const uint32_t stage_tx_bytes = static_cast<uint32_t>((args.bm + BN) * kBK * 2);
  const int acc_stride = args.tmem_cols >> 1;              // TMEM columns between the two accumulator stages
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                  // [kStages][128 x 64] bf16
  uint8_t* smem_b = smem + kStages * kATileBytes;          // [kStages][BN x 64] bf16 (BN * 128 B is a multiple of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kATileBytes + b_tile_bytes));
  uint64_t* full_bar = bars;                       // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;         // [kMaxStages]
  uint64_t* tmem_full = bars + 2 * kMaxStages;     // [2]
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
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

  const int tiles_mn = args.tiles_m * args.tiles_n;
  const int total_tiles = tiles_mn * args.split_k;
  const int kb_per_split = (args.k_blocks + args.split_k - 1) / args.split_k;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp runs the loop (warp-convergent control flow keeps descriptors/addresses in uniform
    // registers); one elected lane issues the TMA instructions.
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t pol_a = l2_policy_evict_first(), pol_b = l2_policy_evict_last();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ks = tile / tiles_mn;
      int m_blk, n_blk;
      tile_coords(tile - ks * tiles_mn, args.tiles_m, args.tiles_n, m_blk, n_blk);
      const int kb0 = ks * kb_per_split;
      const int kb1 = min(args.k_blocks, kb0 + kb_per_split);
      const int m0 = m_blk * args.bm, n0 = n_blk * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
          if (args.stream_a) {
            tma_load_2d_hint(smem_a + stage * kATileBytes, &tmap_a, &full_bar[stage], kb * kBK, m0, pol_a);
            tma_load_2d_hint(smem_b + stage * b_tile_bytes, &tmap_b, &full_bar[stage], kb * kBK, n0, pol_b);
          } else {
            tma_load_2d(smem_a + stage * kATileBytes, &tmap_a, &full_bar[stage], kb * kBK, m0);
            tma_load_2d(smem_b + stage * b_tile_bytes, &tmap_b, &full_bar[stage], kb * kBK, n0);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp loops, one lane issues)
    const uint32_t idesc = umma_idesc_bf16(kBM, static_cast<uint32_t>(BN));
    const uint64_t a_desc0 = umma_desc_k_sw128(smem_u32(smem_a));
    const uint64_t b_desc0 = umma_desc_k_sw128(smem_u32(smem_b));
    const uint64_t a_step = static_cast<uint64_t>(kATileBytes >> 4), b_step = static_cast<uint64_t>(b_tile_bytes >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ks = tile / tiles_mn;
      const int kb0 = ks * kb_per_split;
      const int kb1 = min(args.k_blocks, kb0 + kb_per_split);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * acc_stride;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = a_desc0 + a_step * stage;
          const uint64_t b_desc = b_desc0 + b_step * stage;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 elements = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot is free once these MMAs have read it
          if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ks = tile / tiles_mn;
      int m_blk, n_blk;
      tile_coords(tile - ks * tiles_mn, args.tiles_m, args.tiles_n, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m_local = quarter * 32 + lane;
      const int m = m_blk * args.bm + m_local;  // A-row owned by this thread
      const uint32_t taddr = tmem_base + acc * acc_stride + (static_cast<uint32_t>(quarter * 32) << 16);
      const bool m_ok = m_local < args.bm && m < args.M;
      const int n_chunks = (BN + 31) >> 5;
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (c == n_chunks - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        const int n0 = n_blk * BN + c * 32;
        if (n0 >= args.N || c * 32 >= BN) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const int nvalid = min(min(32, args.N - n0), BN - c * 32);
        if (!args.transposed) {
          // thread = token row m, 32 consecutive features n0..n0+31
          if (!m_ok) continue;
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
          if (!m_ok) continue;
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
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(args.tmem_cols));
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
static int make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int box_rows, std::string* err) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) { *err = "cuTensorMapEncodeTiled entry point not available"; return RVL_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (cols * 2) % 16) {
    *err = "GEMM operand must be 16-byte aligned with a row pitch that is a multiple of 16 bytes";
    return RVL_ERR_INVALID;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(cols) * 2};
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

// Smallest power of two >= 2 * bn (two accumulator stages), at least 32 columns.
static int tmem_cols_for(int bn) {
  int c = 32;
  while (c < 2 * bn) c <<= 1;
  return c;
}

static int launch(const CUtensorMap& ta, const CUtensorMap& tb, GemmArgs& a, int num_sms, cudaStream_t st, std::string* err) {
  const int stage_bytes = kATileBytes + a.bn * kBK * 2;
  a.stages = kSmemBudget / stage_bytes;
  if (a.stages > kMaxStages) a.stages = kMaxStages;
  if (a.stages < 2) { *err = "gemm: tile does not fit in shared memory"; return RVL_ERR_INVALID; }
  a.tmem_cols = tmem_cols_for(a.bn);

  const int smem = a.stages * stage_bytes + 1024 + 512;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return RVL_ERR_CUDA; }
    attr_smem = 227 * 1024;
  }
  const int total = a.tiles_m * a.tiles_n * a.split_k;
  const int grid = total < num_sms ? total : num_sms;
  gemm_bf16_tcgen05_kernel<<<grid, kGemmThreads, smem, st>>>(ta, tb, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("gemm launch: ") + cudaGetErrorString(e); return RVL_ERR_CUDA; }
  return RVL_OK;
}

// Weight-streaming plan (few tokens): rows of W per tile (bm) and split-k chosen so that the tiles fill the SMs in
// whole waves.  Measured on B200 (tools/decode_gemm_sweep.py): one CTA streams ~38 GB/s from HBM whatever the
// pipeline depth, so the only way to approach the 6.5 TB/s peak is to keep all 148 SMs streaming equal shares.
static void plan_stream(int64_t features, int64_t K, int bn, int num_sms, bool allow_split, int fixed_sk, int* bm_out,
                        int* sk_out) {
  const double sm_bw = 38e9, fixed = 3.0e-6, clk = 1.8e9;
  const int k_blocks = static_cast<int>((K + kBK - 1) / kBK);
  double best = 1e30;
  int best_bm = 128, best_sk = 1;
  const int sks[6] = {1, 2, 3, 4, 6, 8};
  for (int bm = 128; bm >= 64; bm -= 8) {
    for (int si = 0; si < (fixed_sk > 0 ? 1 : (allow_split ? 6 : 1)); ++si) {
      const int sk = fixed_sk > 0 ? fixed_sk : sks[si];
      if (fixed_sk <= 0 && sk > 1 && k_blocks / sk < 8) continue;
      const int tiles = static_cast<int>((features + bm - 1) / bm) * sk;
      const int waves = (tiles + num_sms - 1) / num_sms;
      const double kb = static_cast<double>((k_blocks + sk - 1) / sk);
      const double t_mem = kb * (bm + 0.25 * bn) * kBK * 2 / sm_bw;          // activations come from L2: cheaper
      const double t_mma = kb * 4.0 * (bn * 0.5) / clk;                         // 128 x bn x 16 MMA ~ bn/2 cycles
      double t = waves * ((t_mem > t_mma ? t_mem : t_mma) + fixed);
      t += sk > 1 ? 1.0e-6 * sk : 0.0;                                           // partial-sum traffic
      if (t < best) { best = t; best_bm = bm; best_sk = sk; }
    }
  }
  *bm_out = best_bm;
  *sk_out = best_sk;
}

// out[tokens, features] = act(X[tokens,K] . W[features,K]^T + bias)
int gemm_bf16(const GemmCall& c, int num_sms, cudaStream_t st, std::string* err) {
  if (c.M <= 0 || c.N <= 0 || c.K <= 0) { *err = "gemm: empty problem"; return RVL_ERR_INVALID; }
  if (c.K % 8 || c.N % 8) { *err = "gemm: K and N must be multiples of 8"; return RVL_ERR_INVALID; }
  const bool partials = c.split_stride > 0;
  if (c.split_k > 1 && !partials && c.out_mode != RVL_GEMM_ADD_F32) { *err = "gemm: split_k needs RVL_GEMM_ADD_F32 or a partial buffer"; return RVL_ERR_INVALID; }
  if (partials && c.out_mode != RVL_GEMM_OUT_F32) { *err = "gemm: split-k partials are fp32 (RVL_GEMM_OUT_F32)"; return RVL_ERR_INVALID; }
  const bool swap = (c.flags & RVL_GEMM_FLAG_SWAP) != 0;
  GemmArgs a{};
  a.K = static_cast<int>(c.K);
  a.k_blocks = static_cast<int>((c.K + kBK - 1) / kBK);
  a.ldc = c.ldc;
  a.out = c.out;
  a.bias = reinterpret_cast<const __nv_bfloat16*>(c.bias);
  a.rowmap = c.rowmap;
  a.mode = c.out_mode;
  a.relu = (c.flags & RVL_GEMM_FLAG_RELU) ? 1 : 0;
  a.transposed = swap ? 1 : 0;
  const void* pa = swap ? c.W : c.A;   // operand that supplies the (up to) 128-row MMA dimension
  const void* pb = swap ? c.A : c.W;   // operand that supplies MMA N
  a.M = static_cast<int>(swap ? c.N : c.M);
  a.N = static_cast<int>(swap ? c.M : c.N);
  a.bm = kBM;
  int sk = c.split_k < 1 ? 1 : c.split_k;
  if (swap) {
    a.bn = a.N >= 256 ? 256 : ((a.N + 15) / 16) * 16;
    a.stream_a = 1;
    const bool may_split = c.auto_plan && (partials || c.out_mode == RVL_GEMM_ADD_F32);
    plan_stream(a.M, c.K, a.bn, num_sms, may_split, c.auto_plan ? 0 : sk, &a.bm, &sk);
    if (c.max_split > 0 && sk > c.max_split) sk = c.max_split;
    // experiment hooks (tools/decode_gemm_sweep.py): RVL_BM / RVL_SK override the plan, RVL_PLAN_DEBUG prints it
    static const char* env_bm = getenv("RVL_BM");
    static const char* env_sk = getenv("RVL_SK");
    static const char* env_dbg = getenv("RVL_PLAN_DEBUG");
    static const char* env_pf = getenv("RVL_PREFETCH");
    a.prefetch = env_pf ? atoi(env_pf) : 16;
    if (env_bm && atoi(env_bm) >= 8) a.bm = atoi(env_bm);
    if (env_sk && atoi(env_sk) >= 1 && (partials || c.out_mode == RVL_GEMM_ADD_F32)) sk = atoi(env_sk);
    if (env_dbg) fprintf(stderr, "rvl plan: features=%d tokens=%d K=%lld bn=%d bm=%d sk=%d\n", a.M, a.N, (long long)c.K, a.bn, a.bm, sk);
  } else {
    a.bn = a.N >= 256 ? 256 : ((a.N + 15) / 16) * 16;
    a.stream_a = 0;
  }
  a.split_k = sk > a.k_blocks ? a.k_blocks : sk;
  while (a.split_k > 1 && ((a.k_blocks + a.split_k - 1) / a.split_k) * (a.split_k - 1) >= a.k_blocks) --a.split_k;  // no empty split
  a.split_stride = partials ? c.split_stride : 0;
  a.atomic = (a.split_k > 1 && !partials) ? 1 : 0;
  a.tiles_m = (a.M + a.bm - 1) / a.bm;
  a.tiles_n = (a.N + a.bn - 1) / a.bn;
  if (c.split_used) *c.split_used = a.split_k;
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, pa, a.M, c.K, a.bm, err);
  if (rc) return rc;
  rc = make_tmap(&tb, pb, a.N, c.K, a.bn, err);
  if (rc) return rc;
  return launch(ta, tb, a, num_sms, st, err);
}

}  // namespace rvl
