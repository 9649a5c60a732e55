// Micro-benchmark: how fast can ONE thread per SM stream [rows x 64] bf16 boxes with TMA, as a function of the
// number of loads kept in flight, the box height and whether the data comes from HBM or L2?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tma_probe tools/probes/tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../revisionllm_b200/csrc/rvl_ptx.cuh"
using namespace rvl;

__global__ void __launch_bounds__(32, 1) probe(const __grid_constant__ CUtensorMap tm, int rows_per_box, int stages, int iters,
                                                int k_blocks, int rows_per_cta, int l2_resident, unsigned long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * rows_per_box * 128);
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    const uint32_t bytes = rows_per_box * 128;
    const int row0 = blockIdx.x * rows_per_cta;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(&bars[s], ((i / stages) - 1) & 1);
      mbar_arrive_expect_tx(&bars[s], bytes);
      int kb, r;
      if (l2_resident) { kb = i % 4; r = row0; }                       // same 4 boxes over and over: L2 hits
      else { kb = i % k_blocks; r = row0 + (i / k_blocks) * rows_per_box; }  // stream the CTA's private rows once
      tma_load_2d(smem + s * bytes, &tm, &bars[s], kb * 64, r);
    }
    for (int i = iters; i < iters + stages; ++i) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(&bars[s], ((i / stages) - 1) & 1);
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

// Pure issue cost: N loads into N distinct slots, no waiting in between.
__global__ void __launch_bounds__(32, 1) issue_cost(const __grid_constant__ CUtensorMap tm, int rows_per_box, int n, int with_expect,
                                                     unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + n * rows_per_box * 128);
  if (threadIdx.x == 0) {
    for (int i = 0; i < n; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    const uint32_t bytes = rows_per_box * 128;
    const int row0 = blockIdx.x * 2048;
    long long t0 = clock64();
    if (with_expect == 1) {
      for (int i = 0; i < n; ++i) {
        mbar_arrive_expect_tx(&bars[i], bytes);
        tma_load_2d(smem + i * bytes, &tm, &bars[i], i * 64, row0);
      }
    } else {   // one barrier for all loads
      mbar_arrive_expect_tx(&bars[0], bytes * n);
      for (int i = 0; i < n; ++i) tma_load_2d(smem + i * bytes, &tm, &bars[0], i * 64, row0);
    }
    long long t1 = clock64();
    if (with_expect == 1) { for (int i = 0; i < n; ++i) mbar_wait(&bars[i], 0); } else mbar_wait(&bars[0], 0);
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int K = 4096, R = 148 * 2048;   // 2.4 GB of bf16
  void* d;
  cudaMalloc(&d, size_t(R) * K * 2);
  cudaMemset(d, 1, size_t(R) * K * 2);
  unsigned long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  void* p; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  PFN enc = (PFN)p;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  cudaFuncSetAttribute(issue_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int we = 0; we < 2; ++we)
    for (int rows : {8, 32, 128}) {
      const int n = rows == 128 ? 12 : 40;
      CUtensorMap tm;
      cuuint64_t gd[2] = {(cuuint64_t)K, (cuuint64_t)R}; cuuint64_t gs[1] = {(cuuint64_t)K * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
      enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      unsigned long long* o; cudaMalloc(&o, 148 * 16);
      for (int rep = 0; rep < 2; ++rep) {   // second launch: data in L2
        issue_cost<<<148, 32, n * rows * 128 + 2048>>>(tm, rows, n, we, o);
        cudaDeviceSynchronize();
        unsigned long long h[296]; cudaMemcpy(h, o, sizeof h, cudaMemcpyDeviceToHost);
        double a = 0, b = 0; for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
        printf("issue_cost per_load_barrier=%d rows=%3d n=%d %s: issue %.1f cycles/load, all landed after %.0f cycles\n", we, rows, n,
               rep ? "(L2)" : "(HBM)", a / 148 / n, b / 148);
      }
    }
  for (int l2 = 0; l2 < 2; ++l2)
    for (int rows : {32, 64, 128, 256})
      for (int stages : {1, 2, 4, 6}) {
        if (stages * rows * 128 > 200 * 1024) continue;
        CUtensorMap tm;
        cuuint64_t gd[2] = {(cuuint64_t)K, (cuuint64_t)R}; cuuint64_t gs[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
        enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int k_blocks = K / 64;
        const int iters = l2 ? 2048 : (2048 / rows) * k_blocks;   // DRAM case: each CTA streams 2048 private rows once
        const int smem = stages * rows * 128 + 1024 + 256;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        probe<<<148, 32, smem>>>(tm, rows, stages, iters, k_blocks, 2048, l2, cyc);   // warm
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        probe<<<148, 32, smem>>>(tm, rows, stages, iters, k_blocks, 2048, l2, cyc);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        const double bytes = double(iters) * rows * 128;
        printf("%s rows=%3d stages=%d: %8.1f cycles/box  %6.1f B/clk/SM  chip %7.1f GB/s  (%s)\n", l2 ? "L2 " : "HBM", rows, stages,
               avg / iters, bytes / avg, bytes * 148 / (ms * 1e-3) / 1e9, cudaGetErrorString(err));
      }
  return 0;
}
