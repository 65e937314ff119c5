// Micro-benchmark: sustained tcgen05.mma (kind::f16, M=128, K=16, SS) rate per SM as a function of N, of the A-operand
// start row (swizzle-atom aligned or shifted by whole 128-byte rows) and of the accumulator pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I yolo-inspired-audio-activity-detection_b200/csrc -o gpurun_out/mma_rate tools/micro/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace yad;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int BN, int row_shift, int n_acc, int iters, int b_step, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_ptr, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      uint64_t da[16], db[16];
      uint32_t dt[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t a_addr = a_base + (uint32_t)(((j >> 2) % 4) * 128 + row_shift * (1 + (j >> 2))) * 128u + (j & 3) * 32;
        const uint32_t b_addr = b_base + (uint32_t)(((j >> 2) * b_step) % 2) * (BN * 128) + (j & 3) * 32;
        da[j] = ((uint64_t)desc_hi << 32) | (uint64_t)(((a_addr & 0x3FFFFu) >> 4) | (1u << 16));
        db[j] = ((uint64_t)desc_hi << 32) | (uint64_t)(((b_addr & 0x3FFFFu) >> 4) | (1u << 16));
        dt[j] = (uint32_t)(((j >> 2) % n_acc) * BN);
      }
      t0 = clock64();
#pragma unroll 1
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) umma_bf16(dt[j], da[j], db[j], idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (elect_one() && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_ptr, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  printf("  N shift n_acc b_step   cycles/MMA   ideal   (148 CTAs)\n");
  for (int BN : {16, 32, 64, 128, 256})
    for (int shift : {1})
      for (int n_acc : {1, 2})
        for (int b_step : {1}) {
          if (BN == 256 && n_acc == 2 && false) continue;
          mma_rate_kernel<<<148, 128, 180 * 1024>>>(BN, shift, n_acc, iters, b_step, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long c = 0;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("%4d %5d %5d %6d   %10.1f   %5d   %s\n", BN, shift, n_acc, b_step, (double)c / (iters * 16.0), BN / 2,
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  return 0;
}
