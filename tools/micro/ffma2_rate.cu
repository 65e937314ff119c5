// Micro-benchmark: issue rate of scalar FFMA / FADD against the packed sm_100 forms (fma.rn.f32x2 / add.rn.f32x2)
// and of LDS.64 / LDS.128, per SM sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate.bin
// ffma2_rate.cu ; prints cycles per warp instruction per SMSP for 1, 2 and 4 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int NACC = 16;   // independent accumulators (dependency distance 16 >= pipeline latency)

template <int MODE>
__global__ void rate_kernel(float* out, long long* cyc, float seed) {
  __shared__ __align__(16) float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  float a[NACC];
  float2 a2[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    a[i] = seed + i;
    a2[i] = make_float2(seed + i, seed - i);
  }
  const float m = seed * 1.0001f, c = seed * 0.5f;
  const float2 m2 = make_float2(m, m * 1.01f), c2 = make_float2(c, c * 1.01f);
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) a[i] = fmaf(a[i], m, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) a2[i] = __ffma2_rn(a2[i], m2, c2);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) a[i] = a[i] + m;
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) a2[i] = __fadd2_rn(a2[i], m2);
    } else if (MODE == 4) {   // FFMA with three distinct non-constant registers
#pragma unroll
      for (int i = 0; i < NACC; ++i) a[i] = fmaf(a[i], a[(i + 5) % NACC], a[(i + 9) % NACC]);
    } else if (MODE == 5) {   // LDS.64
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        const float2 v = *reinterpret_cast<const float2*>(sm + ((threadIdx.x * 2 + i * 64 + it * 2) & 4094));
        a[i] += v.x + v.y;
      }
    } else if (MODE == 6) {   // LDS.128
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(sm + ((threadIdx.x * 4 + i * 128 + it * 4) & 4092));
        a[i] += v.x + v.w;
      }
    } else if (MODE == 7) {   // FFMA2 + FFMA mixed 1:1
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        a2[i] = __ffma2_rn(a2[i], m2, c2);
        a[i] = fmaf(a[i], m, c);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += a[i] + a2[i].x + a2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_it) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int wps = 1; wps <= 4; wps *= 2) {
    const int threads = 32 * 4 * wps;
    rate_kernel<MODE><<<148, threads>>>(out, cyc, 1.0f);
    rate_kernel<MODE><<<148, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    // warp instructions issued per SMSP = wps * ITERS * per_it
    printf("%-28s warps/SMSP=%d  cycles per warp-instr per SMSP = %.3f\n", name, wps, avg / ((double)wps * ITERS * per_it));
  }
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("FFMA (reg,reg-const,reg-const)", NACC);
  run<4>("FFMA (3 live regs)", NACC);
  run<1>("FFMA2", NACC);
  run<2>("FADD", NACC);
  run<3>("FADD2", NACC);
  run<7>("FFMA2+FFMA pair", 2 * NACC);
  run<5>("LDS.64 (+2 FADD)", 3 * NACC);
  run<6>("LDS.128 (+2 FADD)", 3 * NACC);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
