// Micro-benchmark 2: what slows a tcgen05.mma stream down inside a warp-specialised kernel?  Sustained cycles per MMA
// (kind::f16, M = 128, K = 16, SS) with, selectable by a bit mask:
//   bit 0: a tcgen05.commit after every group of `grp` MMAs (as a weight-ring release does)
//   bit 1: the whole warp walks the loop, mbarrier try_wait (already complete) + tcgen05.fence + elect.sync per group
//   bit 2: 8 other warps stream tcgen05.ld (32x32b.x32) from the other half of TMEM meanwhile
//   bit 3: those warps also write 16-byte shared-memory chunks (epilogue staging) and read a bias vector
//   bit 4: accumulate flag 0 on the first MMA of every group-of-groups (fresh accumulator), else always accumulate
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I yolo-inspired-audio-activity-detection_b200/csrc -o gpurun_out/mma_rate2 tools/micro/mma_rate2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace yad;

__global__ void __launch_bounds__(320, 1) mma_rate2_kernel(int BN, int mode, int grp, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, ring[8], ready;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&ready, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&ring[i], 1);
    stop = 0;
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
    const int MT = grp > 0 ? grp : 256 / BN; // accumulators per group (default as conv_flat: MT x BN = 256 columns)
    long long t0 = clock64();
    uint32_t slot = 0;
    for (int it = 0; it < iters; ++it) {
      if (mode & 2) {
        mbar_wait(&ready, 1);                // fresh barrier: parity 1 is "complete"
        tc_fence_after();
      }
      if (!(mode & 2) ? lane == 0 : elect_one()) {
        const uint32_t b_lo = ((b_base + (slot & 1) * (BN * 128)) & 0x3FFFFu) >> 4 | (1u << 16);
        const uint32_t a_lo = ((a_base + (slot & 3) * 1024 + 128 * (it % 5)) & 0x3FFFFu) >> 4 | (1u << 16);
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16((uint32_t)(mt * BN), ((uint64_t)desc_hi << 32) | (a_lo + mt * 1024 + 2 * k), ((uint64_t)desc_hi << 32) | (b_lo + 2 * k),
                      idesc, (k > 0 || !(mode & 16) || (it % 9)) ? 1u : 0u);
        }
        if (mode & 1) umma_commit(&ring[slot & 7]);
      }
      if (mode & 2) __syncwarp();
      ++slot;
    }
    if (lane == 0) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    stop = 1;
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  } else if (warp >= 2 && (mode & 12)) {
    const int q = warp & 3;
    uint32_t acc = 0;
    uint8_t* S = smem + 128 * 1024 + (warp - 2) * 4096 + lane * 128;      // own 4 KB per warp: [32 rows][128 B]
    const float* bias = reinterpret_cast<const float*>(smem + 127 * 1024);
    while (!stop) {
      uint32_t v[32];
      tmem_ld32(((uint32_t)(q * 32) << 16) + 256u + ((acc & 7) * 32), v);
      tmem_ld_wait();
      if (mode & 8) {
        const float4* bp = reinterpret_cast<const float4*>(bias);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = bp[j];
          v[4 * j] += __float_as_uint(b4.x);
          v[4 * j + 1] ^= __float_as_uint(b4.y);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(S + ((j ^ (lane & 7)) << 4)) = make_uint4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]);
      }
      acc += v[0] & 1;
      ++acc;
    }
    if (acc == 0xffffffffu) out[1] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_ptr, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4000;
  printf("   N  mode   cycles/MMA   ideal\n");
  for (int grp : {0, 1, 2})
  for (int BN : {16, 64, 128})
    for (int mode : {0, 1, 2, 3, 4, 7, 12, 15, 31}) {
      if (grp > 0 && mode > 3) continue;
      if (grp == 0 && BN == 16) continue;
      mma_rate2_kernel<<<148, 320, 180 * 1024>>>(BN, mode, grp, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0;
      cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const int per_it = (grp > 0 ? grp : 256 / BN) * 4;
      printf("grp %d %4d %5d   %10.1f   %5d   %s\n", grp, BN, mode, (double)c / ((double)iters * per_it), BN == 64 ? 48 : 64,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
