// Decode-step layer chain in ONE persistent kernel per attention boundary:
//
//     [o_proj -> residual] -> RMSNorm -> [gate|up + SwiGLU] -> [down -> residual] -> RMSNorm -> [qkv of the next layer | lm_head]
//
// Replaces, for the t > 0 steps of generation (transformers LlamaDecoderLayer reached from
// revisionllm/model/vtimellm_llama.py:79-90 with past_key_values), six kernel launches per layer (four weight-streaming
// GEMMs and two RMSNorms) by one.  Why: at ~180 tokens each of those GEMMs streams 33 - 180 MB of weights in 15 - 40 us,
// and ~7 us of launch / prologue / teardown plus a 2 - 4 us non-overlapped epilogue per launch is what kept the decode
// step at half of HBM speed.  Here
//   * one CTA per SM stays resident across the phases; phases are separated by a grid-wide barrier (one atomic counter);
//   * the TMA producer never stops: while the CTA waits at a barrier it already streams the NEXT phase's weight tiles into
//     the shared-memory ring (weights depend on nothing), and for the first phase it does so before griddepcontrol.wait,
//     i.e. while the attention kernel that precedes this launch is still running;
//   * every GEMM phase is stream-K over all SMs (k-blocks of all 128-row weight tiles dealt evenly, fp32 partial tiles
//     through a workspace, fixed summation order) - the same decomposition as gemm_bf16_tcgen05_kernel<1, true>;
//   * outputs leave through the staged TMA-store epilogue; the residual projections use the TMA reduce-add
//     (cp.reduce.async.bulk.tensor ... .add) straight into the fp32 residual stream, so no split-k partial buffers and no
//     reduction work in RMSNorm remain (each element is added exactly once: deterministic);
//   * RMSNorm phases run on the four epilogue warps of every CTA (rows dealt round-robin).
// Orientation as in the weight-streaming GEMM: weight rows fill the 128-row MMA slot, the (<= 256) tokens sit in MMA N.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "rvl_internal.h"
#include "rvl_ptx.cuh"

namespace rvl {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kThreads = 192;            // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue / RMSNorm
constexpr int kMaxStages = 12;
constexpr int kATileBytes = kBM * kBK * 2;
constexpr int kStageOutBytes = 48 * 1024;
constexpr int kSmemBudget = 227 * 1024 - 1024 - 512 - kStageOutBytes;

struct Phase {
  int kind;               // 0: GEMM, 1: RMSNorm
  // GEMM: out[token][feature] = sum_k x[token][k] * W[feature][k]
  int features;           // weight rows
  int tiles_m, k_blocks, units_per_cta;
  int split_k;            // > 0: plain split-k, CTA b computes k-range (b % split_k) of tile (b / split_k) and stores its fp32
                          //      partial to slot (b % split_k) of the output (RVL_FUSED_OUT_F32 with ldc/slot stride in the map)
  int out_kind;           // RVL_FUSED_OUT_*
  // RMSNorm: x[row] += sum_s partials[s][row] (written back), y[row] = x[row] * rsqrt(mean(x^2) + eps) * w
  float* x;
  const float* partials;
  int n_partials;
  long long partial_stride;
  const __nv_bfloat16* w;
  __nv_bfloat16* y;
  int dim;
  float eps;
};

struct Args {
  int n_phases;
  int n_tokens, bn, stages, sub_stride, tmem_cols;
  float* ws;              // stream-K partial tiles [cta][8 chunks][8][128] float4
  unsigned int* flags;    // stream-K per-CTA epoch flags
  unsigned int epoch0;    // phase p publishes epoch0 + p
  unsigned int* gbar;     // grid barrier counter (monotonic)
  unsigned int gbar_base; // its value when this launch starts
  Phase ph[kFusedMaxPhases];
};

struct Maps {
  CUtensorMap a[kFusedMaxPhases], b[kFusedMaxPhases], out[kFusedMaxPhases];
};

struct Work {
  int m_blk, kb0, kb1, kind, followers;   // kind 0: whole k-range; 1: tail piece (publish partial); 2: head piece (+ followers)
};

__device__ __forceinline__ bool get_work(const Phase& p, int idx, Work& w) {
  if (p.split_k > 0) {
    if (idx > 0 || static_cast<int>(blockIdx.x) >= p.tiles_m * p.split_k) return false;
    const int per = (p.k_blocks + p.split_k - 1) / p.split_k;
    w.m_blk = blockIdx.x / p.split_k;
    w.kb0 = (blockIdx.x % p.split_k) * per;
    w.kb1 = min(p.k_blocks, w.kb0 + per);
    w.kind = 0;
    w.followers = 0;
    return w.kb0 < w.kb1;
  }
  const long long total = static_cast<long long>(p.tiles_m) * p.k_blocks;
  const long long u0 = static_cast<long long>(blockIdx.x) * p.units_per_cta;
  const long long u1 = min(total, u0 + p.units_per_cta);
  if (u0 >= u1) return false;
  const int tile = static_cast<int>(u0 / p.k_blocks) + idx;
  const long long t_begin = static_cast<long long>(tile) * p.k_blocks;
  if (t_begin >= u1) return false;
  w.m_blk = tile;
  w.kb0 = static_cast<int>(max(u0, t_begin) - t_begin);
  w.kb1 = static_cast<int>(min(u1, t_begin + p.k_blocks) - t_begin);
  w.followers = 0;
  if (w.kb0 > 0) {
    w.kind = 1;
  } else if (w.kb1 < p.k_blocks) {
    w.kind = 2;
    const long long t_end = t_begin + p.k_blocks;
    int f = 0;
    while ((static_cast<long long>(blockIdx.x) + f + 1) * p.units_per_cta < t_end) ++f;
    w.followers = f;
  } else {
    w.kind = 0;
  }
  return true;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded like every other wait of the library: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void grid_wait(const unsigned int* bar, unsigned int target) {
  unsigned int spins = 0;
  while (static_cast<int>(ld_acquire_u32(bar) - target) < 0) {
    __nanosleep(40);
    if (++spins > (1u << 24)) {
      printf("rvl: grid barrier timeout block %d target %u value %u\n", blockIdx.x, target, ld_acquire_u32(bar));
      __trap();
    }
  }
}

// tools/fused_timeline.py: per-CTA globaltimer stamps [cta][phase][4] = {producer released (activations available), first
// accumulator ready, epilogue done, arrival at the phase barrier}
__device__ unsigned long long g_fused_dbg[160 * kFusedMaxPhases * 4];
__device__ int g_fused_dbg_on = 0;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FDBG(p, slot) do { if (g_fused_dbg_on) g_fused_dbg[(blockIdx.x * kFusedMaxPhases + (p)) * 4 + (slot)] = gtime(); } while (0)

__device__ __forceinline__ void tma_reduce_add_3d(const void* tmap, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) decode_fused_kernel(const __grid_constant__ Maps maps, const Args args) {
  const int kStages = args.stages;
  const int BN = args.bn;
  const int b_tile_bytes = BN * kBK * 2;
  const uint32_t stage_tx_bytes = static_cast<uint32_t>(kATileBytes + b_tile_bytes);
  const int acc_cols = args.sub_stride;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kATileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kATileBytes + b_tile_bytes));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  float* red = reinterpret_cast<float*>(bars + 2 * kMaxStages + 6);             // [4] RMSNorm partial sums
  uint8_t* stage_out = reinterpret_cast<uint8_t*>(bars) + 512;                  // 48 KB staging of output chunks

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < args.n_phases; ++p)
      if (args.ph[p].kind == 0) {
        tma_prefetch_desc(&maps.a[p]);
        tma_prefetch_desc(&maps.b[p]);
        tma_prefetch_desc(&maps.out[p]);
      }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
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
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t pol_a = l2_policy_evict_first(), pol_b = l2_policy_evict_last();
    Work w;
    for (int p = 0; p < args.n_phases; ++p) {
      const Phase& ph = args.ph[p];
      if (ph.kind != 0) continue;
      // pass 1: the first k-blocks' WEIGHT tiles, as many as slots come free - before the previous phase (or, for the first
      // phase, the previous kernel) is known to be complete
      int pre = 0;
      {
        int s = stage;
        uint32_t sp = phase;
        for (int it = 0; pre < kStages && get_work(ph, it, w); ++it) {
          for (int kb = w.kb0; kb < w.kb1 && pre < kStages; ++kb, ++pre) {
            mbar_wait(&empty_bar[s], sp ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&full_bar[s], stage_tx_bytes);      // the activation tile's bytes follow in pass 2
              tma_load_2d_hint(smem_a + s * kATileBytes, &maps.a[p], &full_bar[s], kb * kBK, w.m_blk * kBM, pol_a);
            }
            __syncwarp();
            if (++s == kStages) { s = 0; sp ^= 1; }
          }
        }
      }
      // the activations of this phase exist once every CTA has finished the previous phase
      if (p == 0) pdl_wait();
      else {
        if (lane == 0) grid_wait(args.gbar, args.gbar_base + static_cast<unsigned int>(p) * gridDim.x);   // one poller per CTA
        __syncwarp();
      }
      fence_proxy_async_all();      // other CTAs' generic stores (RMSNorm output) are about to be read by TMA
      if (lane == 0) FDBG(p, 0);
      int n_iter = 0;
      for (int it = 0; get_work(ph, it, w); ++it) {
        for (int kb = w.kb0; kb < w.kb1; ++kb, ++n_iter) {
          if (n_iter < pre) {
            if (elect_one()) tma_load_2d_hint(smem_b + stage * b_tile_bytes, &maps.b[p], &full_bar[stage], kb * kBK, 0, pol_b);
          } else {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
              tma_load_2d_hint(smem_a + stage * kATileBytes, &maps.a[p], &full_bar[stage], kb * kBK, w.m_blk * kBM, pol_a);
              tma_load_2d_hint(smem_b + stage * b_tile_bytes, &maps.b[p], &full_bar[stage], kb * kBK, 0, pol_b);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc_bf16(kBM, static_cast<uint32_t>(BN));
    const uint64_t a_desc0 = umma_desc_k_sw128(smem_u32(smem_a));
    const uint64_t b_desc0 = umma_desc_k_sw128(smem_u32(smem_b));
    const uint64_t a_step = static_cast<uint64_t>(kATileBytes >> 4), b_step = static_cast<uint64_t>(b_tile_bytes >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    Work w;
    for (int p = 0; p < args.n_phases; ++p) {
      const Phase& ph = args.ph[p];
      if (ph.kind != 0) continue;
      for (int it = 0; get_work(ph, it, w); ++it) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_cols;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a_desc = a_desc0 + a_step * stage;
            const uint64_t b_desc = b_desc0 + b_step * stage;
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > w.kb0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (kb == w.kb1 - 1) umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue + RMSNorm (warps 2..5, 128 threads)
    pdl_wait();
    const int quarter = warp & 3;
    const int m_local = quarter * 32 + lane;
    const int et = threadIdx.x - 64;                 // 0..127
    const int n_chunks = (BN + 31) >> 5;
    const uint32_t stg = smem_u32(stage_out);
    int acc = 0;
    uint32_t acc_phase = 0;
    Work w;
    for (int p = 0; p < args.n_phases; ++p) {
      const Phase& ph = args.ph[p];
      if (ph.kind == 1) {
        // ---- RMSNorm of the residual rows (fp32, read past L1: other SMs' reduce-adds produced them)
        if (p > 0) {
          if (threadIdx.x == 64) grid_wait(args.gbar, args.gbar_base + static_cast<unsigned int>(p) * gridDim.x);
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        const int nvec = ph.dim >> 2;
        for (int row = blockIdx.x; row < args.n_tokens; row += gridDim.x) {
          float4* xr = reinterpret_cast<float4*>(ph.x + static_cast<long long>(row) * ph.dim);
          float4 c[16];
          float ss = 0.f;
          // every load of the row (and of its partial sums) is issued before the first add: one L2 round trip per buffer
          // instead of one per float4 (stores to x in between would otherwise fence the loads behind them)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int idx = et + i * 128;
            if (idx < nvec) c[i] = __ldcg(xr + idx);
          }
          for (int s2 = 0; s2 < ph.n_partials; ++s2) {
            const float4* pr = reinterpret_cast<const float4*>(ph.partials + s2 * ph.partial_stride + static_cast<long long>(row) * ph.dim);
            float4 q[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int idx = et + i * 128;
              if (idx < nvec) q[i] = __ldcg(pr + idx);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int idx = et + i * 128;
              if (idx < nvec) { c[i].x += q[i].x; c[i].y += q[i].y; c[i].z += q[i].z; c[i].w += q[i].w; }
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int idx = et + i * 128;
            if (idx < nvec) {
              if (ph.n_partials > 0) xr[idx] = c[i];
              ss += c[i].x * c[i].x + c[i].y * c[i].y + c[i].z * c[i].z + c[i].w * c[i].w;
            }
          }
          ss = warp_sum(ss);
          asm volatile("bar.sync 1, 128;" ::: "memory");          // red[] free again
          if (lane == 0) red[quarter] = ss;
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const float inv = rsqrtf((red[0] + red[1] + red[2] + red[3]) / static_cast<float>(ph.dim) + ph.eps);
          uint2* yr = reinterpret_cast<uint2*>(ph.y + static_cast<long long>(row) * ph.dim);
          const uint2* wr = reinterpret_cast<const uint2*>(ph.w);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int idx = et + i * 128;
            if (idx < nvec) {
              const uint2 wv = __ldg(wr + idx);
              uint2 o;
              o.x = pack_bf16x2(bf16_lo(wv.x) * (c[i].x * inv), bf16_hi(wv.x) * (c[i].y * inv));
              o.y = pack_bf16x2(bf16_lo(wv.y) * (c[i].z * inv), bf16_hi(wv.y) * (c[i].w * inv));
              yr[idx] = o;
            }
          }
        }
        __threadfence();
        fence_proxy_async_all();    // these generic stores feed other CTAs' TMA loads in the next phase
      } else {
        // ---- GEMM phase: staged epilogue of every work item of this CTA
        const unsigned int epoch = args.epoch0 + static_cast<unsigned int>(p);
        for (int it = 0; get_work(ph, it, w); ++it) {
          mbar_wait(&tmem_full[acc], acc_phase);
          tc_fence_after();
          if (threadIdx.x == 64 && it == 0) FDBG(p, 1);
          const uint32_t taddr = tmem_base + acc * acc_cols + (static_cast<uint32_t>(quarter * 32) << 16);
          if (w.kind == 1) {
            // tail piece of a tile: publish the fp32 partial in register order, then the flag
            for (int c = 0; c < n_chunks; ++c) {
              uint32_t r[32];
              tmem_ld_32x32(taddr + c * 32, r);
              tmem_ld_wait();
              if (c == n_chunks - 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
              }
              float4* dst = reinterpret_cast<float4*>(args.ws) + (static_cast<long long>(blockIdx.x) * 8 + c) * (8 * 128) + m_local;
#pragma unroll
              for (int q = 0; q < 8; ++q)
                dst[q * 128] = make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]),
                                           __uint_as_float(r[q * 4 + 3]));
            }
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (threadIdx.x == 64) st_release_u32(args.flags + blockIdx.x, epoch);
          } else {
            if (w.kind == 2) {
              for (int f = 1; f <= w.followers; ++f) {
                const unsigned int* fl = args.flags + blockIdx.x + f;
                unsigned int spins = 0;
                while (ld_acquire_u32(fl) != epoch) {
                  if (++spins > (1u << 26)) mbar_timeout(nullptr, 0xdead);
                }
              }
            }
            const int cpg = ph.out_kind == RVL_FUSED_OUT_SWIGLU ? 12 : (ph.out_kind == RVL_FUSED_OUT_BF16 ? 6 : 3);
            const int feat0 = w.m_blk * kBM;
            for (int c0 = 0; c0 < n_chunks; c0 += cpg) {
              if (threadIdx.x == 64) bulk_wait_read_all();          // the staging buffer may still feed earlier stores
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
                for (int f = 1; f <= w.followers; ++f) {             // fixed order: deterministic sums
                  const float4* src = reinterpret_cast<const float4*>(args.ws) +
                                      (static_cast<long long>(blockIdx.x + f) * 8 + c) * (8 * 128) + m_local;
                  float4 pp[8];
#pragma unroll
                  for (int q = 0; q < 8; ++q) pp[q] = __ldcg(src + q * 128);
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    v[q * 4] += pp[q].x; v[q * 4 + 1] += pp[q].y; v[q * 4 + 2] += pp[q].z; v[q * 4 + 3] += pp[q].w;
                  }
                }
                const int lc = c - c0;
                if (ph.out_kind == RVL_FUSED_OUT_SWIGLU) {
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
                } else if (ph.out_kind == RVL_FUSED_OUT_BF16) {
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
              fence_proxy_async();
              asm volatile("bar.sync 1, 128;" ::: "memory");
              if (threadIdx.x == 64) {
                const int chunk_bytes = ph.out_kind == RVL_FUSED_OUT_SWIGLU ? 32 * 128 : (ph.out_kind == RVL_FUSED_OUT_BF16 ? 32 * 256 : 32 * 512);
                for (int c = c0; c < c1; ++c) {
                  const int tok0 = c * 32;
                  if (tok0 >= args.n_tokens) continue;
                  const void* src = stage_out + (c - c0) * chunk_bytes;
                  const int slot = ph.split_k > 0 ? static_cast<int>(blockIdx.x % ph.split_k) : 0;
                  if (ph.out_kind == RVL_FUSED_OUT_ADD_F32) tma_reduce_add_3d(&maps.out[p], src, feat0, tok0, 0);
                  else tma_store_3d(&maps.out[p], src, ph.out_kind == RVL_FUSED_OUT_SWIGLU ? (feat0 >> 1) : feat0, tok0, slot);
                }
                bulk_commit_group();
              }
            }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
      // ---- this CTA is done with phase p: its stores are complete and visible before the arrival is
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        FDBG(p, 2);
        bulk_wait_all();
        fence_proxy_async_all();
        __threadfence();
        // one counter serves all phases: a CTA without work in phase p must not arrive for it before phase p - 1 is complete,
        // or its arrival would be counted towards the earlier barrier
        if (p > 0) grid_wait(args.gbar, args.gbar_base + static_cast<unsigned int>(p) * gridDim.x);
        atomicAdd(args.gbar, 1u);
        FDBG(p, 3);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(args.tmem_cols));
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// K-major operand [rows, K] bf16, box [box_rows x 64], 128-byte swizzle
bool map_operand(CUtensorMap* tm, const void* base, int64_t rows, int64_t K, int box_rows) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(K) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// output [tokens][features] (features contiguous), box = 32 tokens x box_feat features, dense
bool map_output(CUtensorMap* tm, const void* base, int64_t n_feat, int64_t n_tok, int64_t ld, int box_feat, bool f32, int n_split = 1) {
  const int es = f32 ? 4 : 2;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(n_feat), static_cast<cuuint64_t>(n_tok), static_cast<cuuint64_t>(n_split)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(ld) * es, static_cast<cuuint64_t>(ld) * es * static_cast<cuuint64_t>(n_tok)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_feat), 32u, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  return encode_fn()(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

int decode_fused(const FusedCall& c, int num_sms, cudaStream_t st, std::string* err) {
  if (!encode_fn()) { *err = "cuTensorMapEncodeTiled entry point not available"; return RVL_ERR_CUDA; }
  if (c.n_phases < 1 || c.n_phases > kFusedMaxPhases || c.n_tokens < 1 || c.n_tokens > 256) {
    *err = "decode_fused: need 1..6 phases and 1..256 tokens";
    return RVL_ERR_INVALID;
  }
  Args a{};
  static Maps maps;     // host staging (copied into the kernel's parameter space at launch)
  a.n_phases = c.n_phases;
  a.n_tokens = c.n_tokens;
  a.bn = ((c.n_tokens + 15) / 16) * 16;
  a.sub_stride = ((a.bn + 31) / 32) * 32;
  int cols = 32;
  while (cols < 2 * a.sub_stride) cols <<= 1;
  a.tmem_cols = cols;
  const int stage_bytes = kATileBytes + a.bn * kBK * 2;
  a.stages = kSmemBudget / stage_bytes;
  if (a.stages > kMaxStages) a.stages = kMaxStages;
  a.ws = c.stream_ws;
  a.flags = c.stream_flags;
  a.epoch0 = c.epoch0;
  a.gbar = c.grid_barrier;
  a.gbar_base = c.grid_barrier_base;
  const int grid = num_sms;
  if (static_cast<size_t>(grid) * 8 * 8 * 128 * 16 > c.stream_ws_bytes) { *err = "decode_fused: stream-K workspace too small"; return RVL_ERR_INVALID; }
  for (int p = 0; p < c.n_phases; ++p) {
    const FusedPhase& f = c.ph[p];
    Phase& ph = a.ph[p];
    ph.kind = f.kind;
    if (f.kind == 1) {
      if (f.dim % 4 || f.dim > 8192) { *err = "decode_fused: RMSNorm dim must be a multiple of 4 and <= 8192"; return RVL_ERR_INVALID; }
      ph.x = f.x; ph.w = reinterpret_cast<const __nv_bfloat16*>(f.norm_w); ph.y = reinterpret_cast<__nv_bfloat16*>(f.y);
      ph.dim = f.dim; ph.eps = f.eps;
      ph.partials = f.partials; ph.n_partials = f.n_partials; ph.partial_stride = static_cast<long long>(c.n_tokens) * f.dim;
      if (f.n_partials < 0 || f.n_partials > 4) { *err = "decode_fused: at most 4 partial buffers"; return RVL_ERR_INVALID; }
      continue;
    }
    if (f.K % 8 || f.features % 8 || (f.out_kind == RVL_FUSED_OUT_SWIGLU && f.features % 32)) { *err = "decode_fused: bad GEMM shape"; return RVL_ERR_INVALID; }
    ph.features = f.features;
    ph.tiles_m = (f.features + kBM - 1) / kBM;
    ph.k_blocks = (f.K + kBK - 1) / kBK;
    const long long total = static_cast<long long>(ph.tiles_m) * ph.k_blocks;
    ph.units_per_cta = static_cast<int>((total + grid - 1) / grid);
    ph.out_kind = f.out_kind;
    ph.split_k = f.split_k;
    if (f.split_k > 0 && (f.out_kind != RVL_FUSED_OUT_F32 || ph.tiles_m * f.split_k > grid || f.split_k > 4)) {
      *err = "decode_fused: split-k phases write fp32 partials and need tiles * split_k <= SMs, split_k <= 4";
      return RVL_ERR_INVALID;
    }
    const bool f32 = f.out_kind == RVL_FUSED_OUT_ADD_F32 || f.out_kind == RVL_FUSED_OUT_F32;
    const bool ok = map_operand(&maps.a[p], f.W, f.features, f.K, kBM) && map_operand(&maps.b[p], f.act, c.n_tokens, f.K, a.bn) &&
                    map_output(&maps.out[p], f.out, f.out_kind == RVL_FUSED_OUT_SWIGLU ? f.features / 2 : f.features, c.n_tokens, f.ldc,
                               f.out_kind == RVL_FUSED_OUT_SWIGLU ? 64 : 128, f32, f.split_k > 0 ? f.split_k : 1);
    if (!ok) { *err = "decode_fused: cuTensorMapEncodeTiled failed"; return RVL_ERR_CUDA; }
  }
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(decode_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      *err = "cudaFuncSetAttribute(decode_fused_kernel) failed";
      return RVL_ERR_CUDA;
    }
    attr = true;
  }
  const int smem = a.stages * stage_bytes + 1024 + 512 + kStageOutBytes;
  cudaError_t e = launch_gemm_k(decode_fused_kernel, dim3(grid), dim3(kThreads), smem, st, maps, a);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("decode_fused launch: ") + cudaGetErrorString(e); return RVL_ERR_CUDA; }
  return RVL_OK;
}

}  // namespace rvl

extern "C" RVL_API void rvl_debug_fused_timestamps(int enable, unsigned long long* out, int n) {
  if (out && n > 0) cudaMemcpyFromSymbol(out, rvl::g_fused_dbg, sizeof(unsigned long long) * n);
  cudaMemcpyToSymbol(rvl::g_fused_dbg_on, &enable, sizeof(int));
}
