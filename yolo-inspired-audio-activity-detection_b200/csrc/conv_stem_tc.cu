// Stem conv1 (2 -> 64 channels, 7x7, stride 2, pad 3; modules/_backbone.py:127,143) on the tcgen05 tensor cores.
//
// K = 98 is too thin and too oddly laid out for TMA (C = 2, overlapping 7-wide windows).  Instead of building an
// im2col tile, the CTA keeps the raw input patch in shared memory as channel-interleaved bf16 rows
//     P[row][col] = (x[c0][row][col], x[c1][row][col])            one 4-byte word per input pixel
// and lets the UMMA shared-memory descriptor do the im2col: in the canonical NO-SWIZZLE K-major layout a core
// matrix is 8 rows x 16 B with consecutive rows 16 B apart, the next core matrix along K sits at +LBO and the next
// 8-row group at +SBO.  With LBO = 16 B and SBO = 128 B row r of the operand is simply bytes [16 r, 16 r + 32) of
// the patch row: the 7 (+1 zero-weight) input columns x 2 channels that output pixel wo = 2 r needs for one kh.
// Odd output pixels need the same windows shifted by 8 B, which a descriptor cannot express (16-byte start
// granularity), so the patch is stored twice, the second copy shifted by two columns.
//   GEMM per (output row ho, pixel parity): D[128 pixels, 64] = sum_kh A_kh[128, 16] * W_kh[64, 16]^T
//   (K index inside a kh step = kw * 2 + c, kw = 7 meets zero weights).
// One tile = 8 output rows x 240 output columns of one clip (21 patch rows); 2 CTAs per SM so that one CTA's patch
// fill overlaps the other's MMAs and stores (the kernel is bound by its 503 MB output write at B = 512).
// Output: dense NHWC [B, Ho, Wo, 64], or the space-to-depth flat layout conv2 reads (see yad_b200.h).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace yad {

constexpr int ST_THREADS = 256;
constexpr int ST_HO = 8;                      // output rows per tile
constexpr int ST_SEG = 240;                   // output columns per tile (120 even + 120 odd pixels)
constexpr int ST_PROWS = 2 * ST_HO + 5;       // 21 patch rows
constexpr int ST_PW = 2 * ST_SEG + 8;         // 488 words per patch row (480 + 6 halo + 2)
constexpr int ST_ROWB = ST_PW * 4;            // 1952 B (multiple of 16)
constexpr int ST_COPYB = ST_PROWS * ST_ROWB;  // one copy of the patch
constexpr int ST_K = 112;                     // 7 kh x 16
constexpr int ST_SBO_W = (ST_K / 8) * 128;    // weights: 1792 B between 8-row groups
constexpr int ST_B_BYTES = (64 / 8) * ST_SBO_W;   // 14336
constexpr int ST_TMEM_COLS = 256;             // 4 accumulators of 64 columns: (2 output rows) x (even, odd pixels)

// no-swizzle K-major descriptor: LBO = distance between K-adjacent core matrices, SBO = distance between 8-row groups
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

struct StemParams {
  int32_t B, H, W, Ho, Wo;
  int32_t n_seg, n_hb, n_tiles;
  int32_t s2d;                  // 0: dense NHWC; 1: space-to-depth flat [B][Wp][Hp][256]
  int64_t osb, osh, osw;        // output pixel strides (in pixels of `ld` channels)
  int32_t ld;
  uint32_t idesc;
};

__global__ void __launch_bounds__(ST_THREADS, 2)
conv_stem_tc_kernel(const float* __restrict__ x, const StemParams p, const uint4* __restrict__ w_packed,
                    __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t st_smem[];
  uint8_t* sP = st_smem;                                   // [2 copies][21][1952]
  uint8_t* sB = sP + 2 * ST_COPYB;                         // 14336 (+ slack read by the garbage rows of the last patch row)
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sB + ST_B_BYTES);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < ST_B_BYTES / 16; i += ST_THREADS) reinterpret_cast<uint4*>(sB)[i] = w_packed[i];
  if (tid == 0) {
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, ST_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t p_addr = smem_u32(sP), b_addr = smem_u32(sB);
  const int q = warp & 3, half = warp >> 2;     // epilogue: TMEM lane quadrant, accumulator pair
  const int r = q * 32 + lane;                  // operand row = pixel pair index inside the segment
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int seg = tile % p.n_seg, hb = (tile / p.n_seg) % p.n_hb, b = tile / (p.n_seg * p.n_hb);
    const int ho0 = hb * ST_HO, hi0 = 2 * ho0 - 3, wi0 = 2 * seg * ST_SEG - 3;
    // (1) patch: fp32 NCHW -> bf16x2 words, zero outside the image; copy 1 = copy 0 shifted left by two words
    const float* x0 = x + (int64_t)b * 2 * p.H * p.W;
    uint32_t* P0 = reinterpret_cast<uint32_t*>(sP);
    uint32_t* P1 = reinterpret_cast<uint32_t*>(sP + ST_COPYB);
    // (7 rows x 2 channels = 14 independent loads in flight per thread: the fill is latency-bound otherwise)
    for (int col = tid; col < ST_PW; col += ST_THREADS) {
      const int wi = wi0 + col;
      const bool wok = wi >= 0 && wi < p.W;
#pragma unroll 1
      for (int row0 = 0; row0 < ST_PROWS; row0 += 7) {
        float a[7], c[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const int hi = hi0 + row0 + j;
          const bool ok = wok && hi >= 0 && hi < p.H;
          a[j] = ok ? __ldg(x0 + (int64_t)hi * p.W + wi) : 0.0f;
          c[j] = ok ? __ldg(x0 + ((int64_t)p.H + hi) * p.W + wi) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(a[j], c[j]);
          const uint32_t wv = *reinterpret_cast<uint32_t*>(&h2);
          const int i = (row0 + j) * ST_PW + col;
          P0[i] = wv;
          if (col >= 2) P1[i - 2] = wv;
        }
      }
    }
    fence_proxy_async();     // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    __syncthreads();
    // (2) four batches of (2 output rows) x (even, odd) accumulators: 28 MMAs each, then the epilogue of the batch
    for (int batch = 0; batch < ST_HO / 2; ++batch) {
      if (warp == 0) {
        if (elect_one()) {
          tc_fence_after();
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int hl = batch * 2 + (a >> 1), par = a & 1;
            const uint32_t a0 = p_addr + par * ST_COPYB + (2 * hl) * ST_ROWB;
#pragma unroll
            for (int kh = 0; kh < 7; ++kh)
              umma_bf16(tmem_base + (uint32_t)(a * 64), make_nosw_desc(a0 + kh * ST_ROWB, 16, 128),
                        make_nosw_desc(b_addr + kh * 256, 128, ST_SBO_W), p.idesc, kh > 0 ? 1u : 0u);
          }
          umma_commit(mma_bar);
        }
        __syncwarp();
      }
      mbar_wait(mma_bar, phase);
      phase ^= 1;
      tc_fence_after();
      // epilogue: warp (q, half) stores accumulators 2*half and 2*half+1 (output row hl, even / odd pixels)
      const int hl = batch * 2 + half, ho = ho0 + hl;
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int wo = seg * ST_SEG + 2 * r + par;
        const bool ok = r < ST_SEG / 2 && wo < p.Wo && ho < p.Ho;
        int64_t off;
        if (p.s2d)
          off = ((int64_t)b * p.osb + (int64_t)(ho >> 1) * p.osh + (int64_t)(wo >> 1) * p.osw) * p.ld + ((ho & 1) * 2 + (wo & 1)) * 64;
        else
          off = ((int64_t)b * p.osb + (int64_t)ho * p.osh + (int64_t)wo * p.osw) * p.ld;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((half * 2 + par) * 64 + c0), v);
          tmem_ld_wait();
          if (ok) {
            uint4* op = reinterpret_cast<uint4*>(out + off + c0);
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
        }
      }
      tc_fence_before();
      __syncthreads();       // accumulators drained before the next batch overwrites them / the patch is refilled
    }
  }
  if (warp == 1) tmem_dealloc(tmem_base, ST_TMEM_COLS);
}

static size_t stem_smem_bytes() { return 2 * ST_COPYB + ST_B_BYTES + 256 + 64; }

int init_conv_stem_tc_attrs() {
  cudaError_t e = cudaFuncSetAttribute(conv_stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stem_smem_bytes());
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_stem_tc_kernel) failed: %s", cudaGetErrorString(e));
    return YAD_ERR_CUDA;
  }
  return YAD_OK;
}

}  // namespace yad

extern "C" int yad_conv_stem_tc(const float* x_nchw, int64_t B, int32_t H, int32_t W, const void* weight_packed, void* out_bf16,
                                int32_t s2d_hp, int32_t s2d_wp, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x_nchw && weight_packed && out_bf16, "yad_conv_stem_tc: null pointer");
  YAD_CHECK_ARG(H >= 1 && W >= 1 && B >= 0, "yad_conv_stem_tc: bad B/H/W");
  YAD_CHECK_ARG((reinterpret_cast<uintptr_t>(weight_packed) % 16 == 0) && (reinterpret_cast<uintptr_t>(out_bf16) % 16 == 0),
                "yad_conv_stem_tc: weight/out must be 16-byte aligned");
  if (B == 0) return YAD_OK;
  StemParams p;
  p.B = (int)B;
  p.H = H;
  p.W = W;
  p.Ho = (H - 1) / 2 + 1;
  p.Wo = (W - 1) / 2 + 1;
  p.n_seg = (p.Wo + ST_SEG - 1) / ST_SEG;
  p.n_hb = (p.Ho + ST_HO - 1) / ST_HO;
  const int64_t n_tiles64 = (int64_t)B * p.n_seg * p.n_hb;
  YAD_CHECK_ARG(n_tiles64 < (1ll << 31), "yad_conv_stem_tc: too many tiles");
  p.n_tiles = (int)n_tiles64;
  if (s2d_hp > 0 || s2d_wp > 0) {
    const int H2 = (p.Ho + 1) / 2, W2 = (p.Wo + 1) / 2;
    YAD_CHECK_ARG(s2d_hp >= H2 && s2d_wp >= W2, "yad_conv_stem_tc: space-to-depth pitches (%d,%d) smaller than the image (%d,%d)",
                  s2d_hp, s2d_wp, H2, W2);
    p.s2d = 1;
    p.ld = 256;
    p.osw = s2d_hp;                       // flat layout: f = (b * Wp + w2) * Hp + h2
    p.osh = 1;
    p.osb = (int64_t)s2d_wp * s2d_hp;
  } else {
    p.s2d = 0;
    p.ld = 64;
    p.osw = 1;
    p.osh = p.Wo;
    p.osb = (int64_t)p.Ho * p.Wo;
  }
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  int grid = 2 * (sm_count() > 0 ? sm_count() : 148);
  if (grid > p.n_tiles) grid = p.n_tiles;
  conv_stem_tc_kernel<<<grid, ST_THREADS, stem_smem_bytes(), (cudaStream_t)stream>>>(
      x_nchw, p, reinterpret_cast<const uint4*>(weight_packed), reinterpret_cast<__nv_bfloat16*>(out_bf16));
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}
