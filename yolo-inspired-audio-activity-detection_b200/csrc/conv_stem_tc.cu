// Stem conv1 (2 -> 64 channels, 7x7, stride 2, pad 3; modules/_backbone.py:127,143) on the tcgen05 tensor cores.
//
// K = 98 is too thin and too oddly laid out for TMA (C = 2, overlapping 7-wide windows), so the A operand is
// built by the CTA itself: the fp32 NCHW input patch of one output-row tile is converted to bf16 (channel pairs
// packed in one word), and every (pixel, kh) pair copies its 14 contiguous values (+2 that meet zero weights) as
// one 32-byte K-chunk into the canonical *no-swizzle* K-major UMMA layout (8x16-byte core matrices).
//   GEMM: D[128 pixels, 64] = A[128, 112] * W[64, 112]^T   (K index = kh*16 + kw*2 + c, kw = 7 is zero padding)
// Persistent CTAs, A and the TMEM accumulator double-buffered: the 7 MMAs of tile t run while the CTA stores
// tile t-1 (tcgen05.ld -> bf16 -> 128-byte NHWC pixel rows) and builds tile t+1.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace yad {

constexpr int ST_THREADS = 256;
constexpr int ST_TW = 128;                    // output pixels per tile (one output row segment)
constexpr int ST_K = 112;                     // 7 kh x 16
constexpr int ST_PCOLS = 2 * ST_TW + 6;       // 262 input columns (261 used + the kw = 7 pad column)
constexpr int ST_PPITCH = 264;
constexpr int ST_SBO = (ST_K / 8) * 128;      // 1792 B between 8-row groups
constexpr int ST_A_BYTES = (ST_TW / 8) * ST_SBO;   // 28672
constexpr int ST_B_BYTES = (64 / 8) * ST_SBO;      // 14336

// no-swizzle K-major descriptor: LBO = distance between the two K-adjacent core matrices of one K=16 step,
// SBO = distance between 8-row groups
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__global__ void __launch_bounds__(ST_THREADS, 2)
conv_stem_tc_kernel(const float* __restrict__ x, int B, int H, int W, int Ho, int Wo, int n_wt, int n_tiles,
                    const uint4* __restrict__ w_packed, __nv_bfloat16* __restrict__ out, uint32_t idesc, int swap_lbo_sbo) {
  extern __shared__ __align__(128) uint8_t st_smem[];
  uint8_t* sA = st_smem;                                   // 2 x 28672
  uint8_t* sB = sA + 2 * ST_A_BYTES;                       // 14336
  uint32_t* sP = reinterpret_cast<uint32_t*>(sB + ST_B_BYTES);   // [7][264] bf16x2 patch
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sP + 7 * ST_PPITCH);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < ST_B_BYTES / 16; i += ST_THREADS) reinterpret_cast<uint4*>(sB)[i] = w_packed[i];
  if (tid == 0) {
    mbar_init(&mma_bar[0], 1);
    mbar_init(&mma_bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t lbo = swap_lbo_sbo ? ST_SBO : 128, sbo = swap_lbo_sbo ? 128 : ST_SBO;

  auto epilogue = [&](int tile, int it) {
    const int buf = it & 1;
    mbar_wait(&mma_bar[buf], (uint32_t)((it >> 1) & 1));
    tc_fence_after();
    const int wt = tile % n_wt, ho = (tile / n_wt) % Ho, b = tile / (n_wt * Ho);
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 64 + half * 32), v);
    tmem_ld_wait();
    const int wo = wt * ST_TW + row;
    if (wo < Wo) {
      uint4* op = reinterpret_cast<uint4*>(out + (((int64_t)b * Ho + ho) * Wo + wo) * 64 + half * 32);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[j4 * 8 + e * 2]), __uint_as_float(v[j4 * 8 + e * 2 + 1]));
          w[e] = *reinterpret_cast<uint32_t*>(&h2);
        }
        op[j4] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    tc_fence_before();
  };

  int it = 0, prev_tile = -1;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const int wt = tile % n_wt, ho = (tile / n_wt) % Ho, b = tile / (n_wt * Ho);
    const int wi0 = 2 * wt * ST_TW - 3, hi0 = 2 * ho - 3;
    // (1) input patch: fp32 NCHW -> bf16x2 (c0 | c1 << 16), zero outside the image
    const float* x0 = x + (int64_t)b * 2 * H * W;
    for (int i = tid; i < 7 * ST_PCOLS; i += ST_THREADS) {
      const int r = i / ST_PCOLS, col = i - r * ST_PCOLS;
      const int hi = hi0 + r, wi = wi0 + col;
      float a = 0.0f, c = 0.0f;
      if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
        a = __ldg(x0 + (int64_t)hi * W + wi);
        c = __ldg(x0 + ((int64_t)H + hi) * W + wi);
      }
      __nv_bfloat162 h2 = __floats2bfloat162_rn(a, c);
      sP[r * ST_PPITCH + col] = *reinterpret_cast<uint32_t*>(&h2);
    }
    __syncthreads();
    // (2) A[buf]: row m, K-chunk kh = patch words [kh][2m .. 2m+7]  (two 16-byte core-matrix rows)
    uint8_t* a_buf = sA + buf * ST_A_BYTES;
    for (int i = tid; i < 7 * ST_TW; i += ST_THREADS) {
      const int kh = i >> 7, m = i & (ST_TW - 1);
      const uint2* src = reinterpret_cast<const uint2*>(sP + kh * ST_PPITCH + 2 * m);
      const uint2 p0 = src[0], p1 = src[1], p2 = src[2], p3 = src[3];
      uint8_t* dst = a_buf + (m >> 3) * ST_SBO + (2 * kh) * 128 + (m & 7) * 16;
      *reinterpret_cast<uint4*>(dst) = make_uint4(p0.x, p0.y, p1.x, p1.y);
      *reinterpret_cast<uint4*>(dst + 128) = make_uint4(p2.x, p2.y, p3.x, p3.y);
    }
    fence_proxy_async();     // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    __syncthreads();
    // (3) one thread issues the 7 K=16 MMAs of this tile; completion arrives on mma_bar[buf]
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(a_buf), b_addr = smem_u32(sB);
#pragma unroll
      for (int s = 0; s < ST_K / 16; ++s)
        umma_bf16(tmem_base + (uint32_t)(buf * 64), make_nosw_desc(a_addr + s * 256, lbo, sbo),
                  make_nosw_desc(b_addr + s * 256, lbo, sbo), idesc, s > 0 ? 1u : 0u);
      umma_commit(&mma_bar[buf]);
    }
    // (4) while those run: store the previous tile
    if (prev_tile >= 0) epilogue(prev_tile, it - 1);
    prev_tile = tile;
    __syncthreads();   // TMEM[buf^1] drained and patch / A[buf^1] free before the next iteration overwrites them
  }
  if (prev_tile >= 0) epilogue(prev_tile, it - 1);
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

int init_conv_stem_tc_attrs() {
  cudaError_t e = cudaFuncSetAttribute(conv_stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_stem_tc_kernel) failed: %s", cudaGetErrorString(e));
    return YAD_ERR_CUDA;
  }
  return YAD_OK;
}

}  // namespace yad

extern "C" int yad_conv_stem_tc(const float* x_nchw, int64_t B, int32_t H, int32_t W, const void* weight_packed, void* out_bf16,
                                int32_t flags, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x_nchw && weight_packed && out_bf16, "yad_conv_stem_tc: null pointer");
  YAD_CHECK_ARG(H >= 1 && W >= 1 && B >= 0, "yad_conv_stem_tc: bad B/H/W");
  YAD_CHECK_ARG((reinterpret_cast<uintptr_t>(weight_packed) % 16 == 0) && (reinterpret_cast<uintptr_t>(out_bf16) % 16 == 0),
                "yad_conv_stem_tc: weight/out must be 16-byte aligned");
  if (B == 0) return YAD_OK;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int n_wt = (Wo + ST_TW - 1) / ST_TW;
  const int64_t n_tiles64 = (int64_t)B * Ho * n_wt;
  YAD_CHECK_ARG(n_tiles64 < (1ll << 31), "yad_conv_stem_tc: too many tiles");
  const int n_tiles = (int)n_tiles64;
  const size_t smem = 2 * ST_A_BYTES + ST_B_BYTES + 7 * ST_PPITCH * 4 + 2 * 8 + 16;
  YAD_CUDA(cudaFuncSetAttribute(conv_stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  int grid = 2 * (sm_count() > 0 ? sm_count() : 148);
  if (grid > n_tiles) grid = n_tiles;
  conv_stem_tc_kernel<<<grid, ST_THREADS, smem, (cudaStream_t)stream>>>(
      x_nchw, (int)B, H, W, Ho, Wo, n_wt, n_tiles, reinterpret_cast<const uint4*>(weight_packed),
      reinterpret_cast<__nv_bfloat16*>(out_bf16), idesc, flags & 1);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}
