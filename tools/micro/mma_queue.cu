// Micro-benchmark 3: how far can the issuing thread run ahead of the tensor pipe?  G back-to-back tcgen05.mma (kind::f16, M = 128,
// N = 128 or 64, K = 16, SS, precomputed descriptors as in mma_rate.cu), then the thread spins for D cycles (clock64), repeated.
// If the pipe queues q MMAs beyond the one executing, a group costs G * T + max(0, D - q * T) cycles (T = 64 / 48).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I yolo-inspired-audio-activity-detection_b200/csrc -o /tmp/mma_queue tools/micro/mma_queue.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace yad;

template <int G>
__global__ void __launch_bounds__(128, 1) mma_queue_kernel(int BN, int D, int iters, int commit, int nwarp, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2], ring[8];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&ring[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_ptr, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp >= 1 && warp <= nwarp) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      uint64_t da[G], db[G];
      uint32_t dt[G];
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const uint32_t a_addr = a_base + (uint32_t)(((j >> 2) % 4) * 128 + (1 + (j >> 2))) * 128u + (j & 3) * 32;
        const uint32_t b_addr = b_base + (uint32_t)((j >> 2) % 2) * (BN * 128) + (j & 3) * 32;
        da[j] = ((uint64_t)desc_hi << 32) | (uint64_t)(((a_addr & 0x3FFFFu) >> 4) | (1u << 16));
        db[j] = ((uint64_t)desc_hi << 32) | (uint64_t)(((b_addr & 0x3FFFFu) >> 4) | (1u << 16));
        dt[j] = (uint32_t)(((j >> 2) % 2) * BN + (warp - 1) * 256);
      }
      t0 = clock64();
#pragma unroll 1
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < G; ++j) umma_bf16(dt[j], da[j], db[j], idesc, 1u);
        if (commit) umma_commit(&ring[(it & 3) + 4 * (warp - 1)]);
        if (D > 0) {
          const long long ts = clock64();
          while (clock64() - ts < D) {}
        }
      }
      umma_commit(&bar[warp - 1]);
    }
    __syncwarp();
    mbar_wait(&bar[warp - 1], 0);
    t1 = clock64();
    if (elect_one() && blockIdx.x == 0) out[warp - 1] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_ptr, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_queue_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(mma_queue_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(mma_queue_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  printf("   N   G  warps commit     D   cycles/group (per warp)   G*T\n");
  for (int BN : {64, 128})
   for (int nwarp : {1, 2})
    for (int commit : {1})
      for (int D : {0, 64, 128, 192, 256, 384, 512}) {
        const int G = (BN == 64 ? 16 : 8) / nwarp;      // the same MMAs per round of all warps
        if (G == 4)
          mma_queue_kernel<4><<<148, 128, 180 * 1024>>>(BN, D, iters, commit, nwarp, d);
        else if (G == 8)
          mma_queue_kernel<8><<<148, 128, 180 * 1024>>>(BN, D, iters, commit, nwarp, d);
        else
          mma_queue_kernel<16><<<148, 128, 180 * 1024>>>(BN, D, iters, commit, nwarp, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("%4d %3d %5d %7d %5d   %12.1f   %4d   %s\n", BN, G, nwarp, commit, D, (double)c / iters, G * (BN == 64 ? 48 : 64),
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
