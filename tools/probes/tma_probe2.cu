// Micro-benchmark 2: is one SM's TMA rate limited per box (fixed cost), per issuing thread, or by bytes?
//   mode 2D : box [rows x 64] bf16 (rows x 128 B)
//   mode 3D : box [kk x rows x 64] over the matrix viewed as [K/64][rows][64] -> kk k-blocks per instruction
// W issuing warps (one elected lane each), each with its own ring of `stages` slots; data L2-resident or streamed from HBM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tma_probe2 tools/probes/tma_probe2.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../revisionllm_b200/csrc/rvl_ptx.cuh"
using namespace rvl;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, int box_bytes, int rows_per_box, int kk, int stages,
                                                 int iters, int k_blocks, int rows_per_cta, int l2_resident, int is3d,
                                                 unsigned long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int W = blockDim.x >> 5, warp = threadIdx.x >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W * stages * box_bytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < W * stages; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    uint8_t* my = smem + warp * stages * box_bytes;
    uint64_t* mb = bars + warp * stages;
    const int row0 = blockIdx.x * rows_per_cta;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(&mb[s], ((i / stages) - 1) & 1);
      mbar_arrive_expect_tx(&mb[s], box_bytes);
      const int g = i * W + warp;                       // global box index of this CTA
      int kb, r;
      const int kb_groups = k_blocks / kk;
      if (l2_resident) { kb = (g % 4) * kk; r = row0; }
      else { kb = (g % kb_groups) * kk; r = row0 + (g / kb_groups) * rows_per_box; }
      if (is3d) tma_load_3d(my + s * box_bytes, &tm, &mb[s], 0, r, kb);
      else tma_load_2d(my + s * box_bytes, &tm, &mb[s], kb * 64, r);
    }
    for (int i = iters; i < iters + stages; ++i) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(&mb[s], ((i / stages) - 1) & 1);
    }
    cycles[blockIdx.x * 4 + warp] = clock64() - t0;
  }
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int K = 4096, R = 148 * 2048;
  void* d;
  cudaMalloc(&d, size_t(R) * K * 2);
  cudaMemset(d, 1, size_t(R) * K * 2);
  unsigned long long* cyc;
  cudaMalloc(&cyc, 148 * 4 * 8);
  void* p; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  PFN enc = (PFN)p;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Cfg { int is3d, rows, kk, warps, stages, grid; };
  const Cfg cfgs[] = {
      {0, 128, 1, 1, 8, 148}, {0, 128, 1, 1, 8, 96}, {0, 128, 1, 1, 8, 74}, {0, 128, 1, 1, 8, 48}, {0, 128, 1, 1, 12, 96},
      {0, 128, 1, 2, 6, 96}, {0, 128, 1, 2, 6, 48}, {0, 256, 1, 1, 6, 96}, {0, 256, 1, 2, 3, 96}, {0, 256, 1, 2, 3, 48}, {0, 128, 1, 4, 3, 96},
  };
  for (int l2 = 0; l2 < 1; ++l2)
    for (const Cfg& c : cfgs) {
      CUtensorMap tm;
      CUresult rc;
      if (c.is3d) {
        cuuint64_t gd[3] = {64, (cuuint64_t)R, (cuuint64_t)K / 64}; cuuint64_t gs[2] = {(cuuint64_t)K * 2, 128};
        cuuint32_t box[3] = {64, (cuuint32_t)c.rows, (cuuint32_t)c.kk}; cuuint32_t es[3] = {1, 1, 1};
        rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else {
        cuuint64_t gd[2] = {(cuuint64_t)K, (cuuint64_t)R}; cuuint64_t gs[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)c.rows}; cuuint32_t es[2] = {1, 1};
        rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); continue; }
      const int box_bytes = c.rows * 128 * c.kk;
      if (c.warps * c.stages * box_bytes > 200 * 1024) { printf("skip (smem)\n"); continue; }
      const int k_blocks = K / 64;
      // every CTA streams 2048 private rows x K once (HBM) or re-reads 4 boxes (L2)
      const int boxes_total = l2 ? 4096 : (2048 / c.rows) * (k_blocks / c.kk);
      const int iters = boxes_total / c.warps;
      const int smem = c.warps * c.stages * box_bytes + 1024 + 512;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      probe<<<c.grid, 32 * c.warps, smem>>>(tm, box_bytes, c.rows, c.kk, c.stages, iters, k_blocks, 2048, l2, c.is3d, cyc);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      probe<<<c.grid, 32 * c.warps, smem>>>(tm, box_bytes, c.rows, c.kk, c.stages, iters, k_blocks, 2048, l2, c.is3d, cyc);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      unsigned long long h[148 * 4]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < c.grid; ++i) avg += h[i * 4]; avg /= c.grid;
      const double bytes = double(iters) * c.warps * box_bytes;
      printf("%s %s grid=%3d rows=%3d kk=%d warps=%d stages=%d box=%3d KB: %7.1f cyc/box/warp  %6.1f B/clk/SM  chip %7.1f GB/s (%s)\n", l2 ? "L2 " : "HBM",
             c.is3d ? "3D" : "2D", c.grid, c.rows, c.kk, c.warps, c.stages, box_bytes / 1024, avg / iters, bytes / avg, bytes * c.grid / (ms * 1e-3) / 1e9,
             cudaGetErrorString(err));
    }
  return 0;
}
