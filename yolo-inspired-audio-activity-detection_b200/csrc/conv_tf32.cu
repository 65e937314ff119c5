// TF32 tcgen05 / TMEM convolution kernels for the TRAIN step (fp32 tensors in HBM, tf32 operands, fp32 accumulate in tensor
// memory) - what cuDNN does for the reference's training by default (torch.backends.cudnn.allow_tf32 = True) when
// pipeline/_trainer.py:94-108 runs F.conv2d forward / backward on a GPU.
//
// 1. corr_tf32_kernel: generic tap-list correlation
//        out[b,i,j,n] (+)= bias[n] + sum_t sum_c in[b, i*sh + dh_t, j*sw + dw_t, c] * Wt[n][k_t + c]
//    (implicit GEMM, M = 128 output pixels of a (tb x th x tw) box, N <= 128 channels, K = taps x 32-channel blocks).
//    A operand: 4-D TMA box {32 ch, tw, th, tb} of the NHWC fp32 input at the tap-shifted coordinate of the tap's stride
//    parity class (out-of-range = zero fill = padding), landing as 128 rows x 128 B = the canonical K-major SWIZZLE_128B
//    layout.  The host expresses with it: the forward convolution, the data gradient of a stride-1 convolution (flipped
//    taps, transposed weights) and the data gradient of a stride-2 convolution (one call per output parity class with the
//    sub-filter of that class and output pixel strides of 2).
// 2. wgrad_tf32_kernel: weight gradient  dW[t][ci][co] += sum_pixels X[b, i*sh + dh_t, j*sw + dw_t, ci] * dY[b,i,j,co]
//    (M = 128 input channels, N <= 128 output channels, K = pixels).  NHWC makes BOTH operands MN-major (the channel is
//    the contiguous index, the pixel = K the strided one): the same TMA boxes {32 ch, tw, th, tb} are read by the MMA
//    through MN-major SWIZZLE_128B_BASE32B descriptors - the only MN-major layout the tensor core takes for 32-bit
//    elements, written by TMA's SWIZZLE_128B_ATOM_32B mode (LBO = distance between 32-channel blocks, two 4-pixel K atoms per
//    instruction).  Split over pixel chunks, fp32 `red.add` into dW.
//
// Warp roles (192 threads) as in conv_tc.cu: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <string.h>

namespace yad {

constexpr int TF_THREADS = 192;
constexpr int TF_BM = 128;
constexpr int TF_BK = 32;                        // fp32 elements per K-block (128 B)
constexpr int TF_A_STAGE_BYTES = TF_BM * 128;    // 16 KB
constexpr int TF_MAX_TAPS = 49;
constexpr int TF_MAX_STAGES = 8;

struct TfParams {
  int32_t B, Ho, Wo;
  int32_t tb, th, tw;
  int32_t n_wt, n_ht, n_bt;
  int32_t BN, stages, cin_chunks, n_taps;
  int32_t Cout, ld_out, act, accumulate;
  int32_t ksplit, iters_per_split;   // split-K over the (tap, channel-block) iterations: partial sums are red.add-ed
  uint32_t idesc;
  int8_t tap_map[TF_MAX_TAPS];
  int8_t tap_dw[TF_MAX_TAPS];
  int8_t tap_dh[TF_MAX_TAPS];
  int32_t tap_k[TF_MAX_TAPS];
  int32_t osw, osh, osb;
};

__device__ __forceinline__ uint64_t tf_desc_k_sw128(uint32_t smem_addr) {     // K-major SWIZZLE_128B, SBO = 1024 B
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major operand of 32-bit elements: the only layout the tensor core accepts is SWIZZLE_128B_BASE32B (layout type 1; what
// TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes): rows of 128 B (32 fp32 contiguous along M/N) repeating every `lbo`
// bytes along M/N; 4 K rows per swizzle atom (the 32-byte chunks of a row are XOR-ed with row % 4), K atoms `sbo` bytes apart
__device__ __forceinline__ uint64_t tf_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ------------------------------------------------------------------------------------------------ correlation
__global__ void __launch_bounds__(TF_THREADS, 1)
corr_tf32_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
                 const __grid_constant__ CUtensorMap map_w, const TfParams p, const float* __restrict__ bias,
                 float* __restrict__ out, double* __restrict__ stats) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_stage_bytes = p.BN * 128;
  const int stage_bytes = TF_A_STAGE_BYTES + b_stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + TF_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + TF_MAX_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const int it_lo = blockIdx.z * p.iters_per_split;
  const int it_hi = min(it_lo + p.iters_per_split, p.n_taps * p.cin_chunks);
  const int n_iters = it_hi - it_lo;
  int t = blockIdx.x;
  const int wt = t % p.n_wt;
  t /= p.n_wt;
  const int ht = t % p.n_ht;
  t /= p.n_ht;
  const int bt = t;
  const int w0 = wt * p.tw, h0 = ht * p.th, b0 = bt * p.tb;
  const int n0 = blockIdx.y * p.BN;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < p.BN) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a0);
    prefetch_tmap(&map_w);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)(p.tb * p.th * p.tw * 128 + b_stage_bytes);
      uint32_t s = 0, phase = 0;
      int tap = it_lo / p.cin_chunks, cc = it_lo % p.cin_chunks;
      for (int it = 0; it < n_iters; ++it) {
        const int mi = p.tap_map[tap];
        const CUtensorMap* ma = mi == 0 ? &map_a0 : (mi == 1 ? &map_a1 : (mi == 2 ? &map_a2 : &map_a3));
        mbar_wait(&empty_bar[s], phase ^ 1);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], tx_bytes);
        tma_load_4d(ma, &full_bar[s], sa, cc * TF_BK, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], b0);
        tma_load_2d(&map_w, &full_bar[s], sa + TF_A_STAGE_BYTES, p.tap_k[tap] + cc * TF_BK, n0);
        if (++s == (uint32_t)S) { s = 0; phase ^= 1; }
        if (++cc == p.cin_chunks) { cc = 0; ++tap; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t s = 0, phase = 0;
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(&full_bar[s], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
        const uint64_t da = tf_desc_k_sw128(sa), db = tf_desc_k_sw128(sa + TF_A_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < TF_BK / 8; ++k)      // 8 tf32 = 32 B along K inside the swizzled 128 B row: +2 in the address field
          umma_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), p.idesc, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        if (it == n_iters - 1) umma_commit(tmem_full_bar);
        if (++s == (uint32_t)S) { s = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int rows = p.tb * p.th * p.tw;
    const int wl = r % p.tw;
    const int hl = (r / p.tw) % p.th;
    const int bl = r / (p.tw * p.th);
    const int ow = w0 + wl, oh = h0 + hl, ob = b0 + bl;
    const bool row_ok = (r < rows) && (ow < p.Wo) && (oh < p.Ho) && (ob < p.B);
    const int64_t pix = (int64_t)ob * p.osb + (int64_t)oh * p.osh + (int64_t)ow * p.osw;
    float* orow = out + pix * p.ld_out;
    const bool vec_ok = (p.ld_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const bool add_bias = bias != nullptr && blockIdx.z == 0;
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (stats != nullptr) {
        // per-channel sum / sum of squares of the conv output (the BatchNorm batch moments, modules/_common.py:43-48): lane l
        // ends up with column c0 + l of this warp's 32 rows (32-step transpose-reduce with shuffles), then one fp64 atomic
        float a[32], b2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = row_ok ? __uint_as_float(v[j]) + (bias != nullptr && n0 + c0 + j < p.Cout ? __ldg(bias + n0 + c0 + j) : 0.0f) : 0.0f;
          a[j] = x;
          b2[j] = x * x;
        }
        // butterfly transpose-reduce: after the steps 16, 8, 4, 2, 1 element 0 of lane l is the sum over the warp of column l
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int j = 0; j < off; ++j) {
            const float sa = up ? a[j] : a[j + off], ka = up ? a[j + off] : a[j];
            const float sb = up ? b2[j] : b2[j + off], kb = up ? b2[j + off] : b2[j];
            a[j] = ka + __shfl_xor_sync(0xffffffffu, sa, off);
            b2[j] = kb + __shfl_xor_sync(0xffffffffu, sb, off);
          }
        }
        const float s1 = a[0], s2 = b2[0];
        const int n = n0 + c0 + lane;
        if (n < p.Cout) {
          atomicAdd(stats + n, (double)s1);
          atomicAdd(stats + p.Cout + n, (double)s2);
        }
      }
      if (!row_ok) continue;
      const int nbase = n0 + c0;
      if (nbase >= p.Cout) continue;
      if (p.ksplit > 1) {            // partial sum of this K range (the host zero-filled `out` unless it accumulates)
        for (int j = 0; j < 32; ++j) {
          const int n = nbase + j;
          if (n >= p.Cout) break;
          atomicAdd(orow + n, __uint_as_float(v[j]) + (add_bias ? __ldg(bias + n) : 0.0f));
        }
      } else if (vec_ok && nbase + 32 <= p.Cout) {
        float4* op = reinterpret_cast<float4*>(orow + nbase);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float f[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) f[e] = __uint_as_float(v[j4 * 4 + e]) + (bias != nullptr ? __ldg(bias + nbase + j4 * 4 + e) : 0.0f);
          if (p.accumulate) {
            // out += : a 16-byte reduction executed at the L2 (red.global.add.v4.f32).  Measured: loading the old value and
            // storing the sum from the SM made every partial-line store a DRAM write-back (985 MB written for a 15.7 MB
            // tensor, 26 -> 171 us on the layer1 shape)
            atomicAdd(op + j4, make_float4(f[0], f[1], f[2], f[3]));
          } else {
            op[j4] = make_float4(apply_act(f[0], p.act), apply_act(f[1], p.act), apply_act(f[2], p.act), apply_act(f[3], p.act));
          }
        }
      } else {
        for (int j = 0; j < 32; ++j) {
          const int n = nbase + j;
          if (n >= p.Cout) break;
          const float x = __uint_as_float(v[j]) + (bias != nullptr ? __ldg(bias + n) : 0.0f);
          if (p.accumulate)
            atomicAdd(orow + n, x);
          else
            orow[n] = apply_act(x, p.act);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------ weight gradient
struct WgParams {
  int32_t tb, th, tw, npix;          // pixel box (npix = tb*th*tw, a multiple of 8)
  int32_t n_wt, n_ht, n_bt, n_tiles, tiles_per_chunk;
  int32_t BN, stages;
  int32_t ci_blocks;                 // ceil(Cin / 128)
  int32_t a_blocks;                  // 32-channel blocks of X loaded per tile (<= 4)
  int32_t Cin, Cout;
  uint32_t idesc;
  uint32_t a_lbo;                    // byte distance between the 32-channel blocks of A as the MMA sees them
  int8_t tap_map[TF_MAX_TAPS];
  int8_t tap_dw[TF_MAX_TAPS];
  int8_t tap_dh[TF_MAX_TAPS];
  int8_t tap_dst[TF_MAX_TAPS];       // tap index in dW
};

__global__ void __launch_bounds__(TF_THREADS, 1)
wgrad_tf32_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1,
                  const __grid_constant__ CUtensorMap map_x2, const __grid_constant__ CUtensorMap map_x3,
                  const __grid_constant__ CUtensorMap map_dy, const WgParams p, float* __restrict__ dw) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int blk_bytes = p.npix * 128;                         // one 32-channel block of the pixel box
  const int a_bytes = 4 * blk_bytes, b_bytes = (p.BN / 32) * blk_bytes;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + TF_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + TF_MAX_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const int tap = blockIdx.y / p.ci_blocks, cib = blockIdx.y % p.ci_blocks;
  const int ci0 = cib * 128, n0 = blockIdx.z * p.BN;
  const int t_lo = blockIdx.x * p.tiles_per_chunk;
  const int t_hi = min(t_lo + p.tiles_per_chunk, p.n_tiles);
  const int n_iters = t_hi - t_lo;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < p.BN) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_x0);
    prefetch_tmap(&map_dy);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (n_iters <= 0) {          // (cannot happen with the host's chunking, kept for safety)
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
    return;
  }

  if (warp == 0) {
    if (lane == 0) {
      const int mi = p.tap_map[tap];
      const CUtensorMap* mx = mi == 0 ? &map_x0 : (mi == 1 ? &map_x1 : (mi == 2 ? &map_x2 : &map_x3));
      const int a_blocks = min(p.a_blocks, (p.Cin - ci0 + 31) / 32);
      const int b_blocks = p.BN / 32;
      const uint32_t tx_bytes = (uint32_t)((a_blocks + b_blocks) * blk_bytes);
      uint32_t s = 0, phase = 0;
      for (int tile = t_lo; tile < t_hi; ++tile) {
        int t = tile;
        const int w0 = (t % p.n_wt) * p.tw;
        t /= p.n_wt;
        const int h0 = (t % p.n_ht) * p.th;
        const int b0 = (t / p.n_ht) * p.tb;
        mbar_wait(&empty_bar[s], phase ^ 1);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        mbar_expect_tx(&full_bar[s], tx_bytes);
        for (int a = 0; a < a_blocks; ++a)
          tma_load_4d(mx, &full_bar[s], sa + a * blk_bytes, ci0 + a * 32, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], b0);
        for (int b = 0; b < b_blocks; ++b) tma_load_4d(&map_dy, &full_bar[s], sb + b * blk_bytes, n0 + b * 32, w0, h0, b0);
        if (++s == (uint32_t)S) { s = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t s = 0, phase = 0;
      const int kgroups = p.npix / 8;
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(&full_bar[s], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t sb = sa + (uint32_t)a_bytes;
        for (int k = 0; k < kgroups; ++k) {
          const uint64_t da = tf_desc_mn_sw128(sa + k * 1024, p.a_lbo, 512);
          const uint64_t db = tf_desc_mn_sw128(sb + k * 1024, (uint32_t)blk_bytes, 512);
          umma_tf32(tmem_base, da, db, p.idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
        if (it == n_iters - 1) umma_commit(tmem_full_bar);
        if (++s == (uint32_t)S) { s = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int ci = ci0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    float* drow = dw + ((int64_t)p.tap_dst[tap] * p.Cin + ci) * p.Cout;
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (ci >= p.Cin) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = n0 + c0 + j;
        if (n < p.Cout) atomicAdd(drow + n, __uint_as_float(v[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_map_f32(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(tensor_map_encode_fn());
  if (!fn) {
    set_error("conv_tf32: yad_init() was not called (cuTensorMapEncodeTiled unresolved)");
    return YAD_ERR_ARG;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(f32) failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return YAD_ERR_CUDA;
  }
  return YAD_OK;
}

int init_conv_tf32_attrs() {
  cudaError_t e = cudaFuncSetAttribute(corr_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_tf32) failed: %s", cudaGetErrorString(e));
    return YAD_ERR_CUDA;
  }
  return YAD_OK;
}

static inline int cdiv_(int a, int b) { return (a + b - 1) / b; }

// (tb, th, tw) pixel box with at most `cap` rows (and a multiple of `mult` rows) maximising the useful rows per MMA
static void choose_box(int B, int Ho, int Wo, int cap, int mult, int* tb, int* th, int* tw) {
  double best = -1.0;
  int bb = 1, bh = 1, bw = mult;
  const int wmax = cdiv_(Wo, mult) * mult;
  for (int w = 1; w <= wmax && w <= cap; ++w) {
    for (int h = 1; h <= Ho && h * w <= cap; ++h) {
      for (int b = 1; b <= B && b * h * w <= cap; ++b) {
        if ((b * h * w) % mult) continue;
        const double cover = (double)B * Ho * Wo / ((double)cdiv_(Wo, w) * w * cdiv_(Ho, h) * h * cdiv_(B, b) * b);
        const double eff = cover * (b * h * w) / (double)cap + 1e-6 * w;
        if (eff > best) {
          best = eff;
          bb = b;
          bh = h;
          bw = w;
        }
      }
    }
  }
  *tb = bb;
  *th = bh;
  *tw = bw;
}

// one tensor map per stride parity class of the input (as conv_tc.cu); returns the map index / coordinate shift of a tap
static int build_parity_maps(const float* in, int B, int H, int W, int C, int ld, int sh, int sw, const uint32_t* box,
                             CUtensorMap* maps, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  for (int prh = 0; prh < sh; ++prh) {
    for (int prw = 0; prw < sw; ++prw) {
      const int mi = prh * sw + prw;
      const int Wm = (W - prw + sw - 1) / sw, Hm = (H - prh + sh - 1) / sh;
      if (Wm <= 0 || Hm <= 0) {
        maps[mi] = maps[0];
        continue;
      }
      const float* base = in + ((int64_t)prh * W + prw) * ld;
      const uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wm, (uint64_t)Hm, (uint64_t)B};
      const uint64_t strides[3] = {(uint64_t)sw * ld * 4, (uint64_t)sh * W * ld * 4, (uint64_t)H * W * ld * 4};
      int rc = encode_map_f32(&maps[mi], base, 4, dims, strides, box, swizzle);
      if (rc) return rc;
    }
  }
  for (int mi = sh * sw; mi < 4; ++mi) maps[mi] = maps[0];
  return YAD_OK;
}

}  // namespace yad

extern "C" {

int yad_corr_tf32(const yad_corr_desc* d, const int32_t* tap_dh, const int32_t* tap_dw, const int32_t* tap_k, const float* in,
                  const float* weight, int32_t cout_pad, int64_t k_total, const float* bias, float* out, double* stats,
                  yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(d && tap_dh && tap_dw && tap_k && in && weight && out, "yad_corr_tf32: null pointer");
  YAD_CHECK_ARG(d->Cin % 32 == 0 && d->Cin >= 32 && d->ld_in >= d->Cin && d->ld_in % 4 == 0,
                "yad_corr_tf32: Cin=%d must be a multiple of 32 (zero-pad channels), ld_in=%d a multiple of 4", d->Cin, d->ld_in);
  YAD_CHECK_ARG(cout_pad % 16 == 0 && cout_pad >= d->Cout && d->Cout >= 1 && d->ld_out >= d->Cout, "yad_corr_tf32: bad Cout / cout_pad / ld_out");
  YAD_CHECK_ARG((d->sh == 1 || d->sh == 2) && (d->sw == 1 || d->sw == 2), "yad_corr_tf32: stride (%d,%d) unsupported", d->sh, d->sw);
  YAD_CHECK_ARG(d->n_taps >= 1 && d->n_taps <= TF_MAX_TAPS, "yad_corr_tf32: %d taps (max %d)", d->n_taps, TF_MAX_TAPS);
  YAD_CHECK_ARG(d->B >= 1 && d->H >= 1 && d->W >= 1 && d->Ho >= 1 && d->Wo >= 1, "yad_corr_tf32: empty tensor");
  YAD_CHECK_ARG(!(d->accumulate && d->act != YAD_ACT_NONE), "yad_corr_tf32: accumulate (a reduction at the L2) cannot apply an activation");
  YAD_CHECK_ARG(k_total % 4 == 0 && (reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(weight) % 16 == 0),
                "yad_corr_tf32: pointers must be 16-byte aligned, k_total a multiple of 4");
  TfParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B;
  p.Ho = d->Ho;
  p.Wo = d->Wo;
  choose_box(d->B, d->Ho, d->Wo, 128, 1, &p.tb, &p.th, &p.tw);
  p.n_wt = cdiv_(d->Wo, p.tw);
  p.n_ht = cdiv_(d->Ho, p.th);
  p.n_bt = cdiv_(d->B, p.tb);
  int BN = cout_pad >= 128 ? 128 : cout_pad;
  if (cout_pad % BN != 0) BN = 64;
  if (cout_pad % BN != 0) BN = 32;
  if (cout_pad % BN != 0) BN = 16;
  p.BN = BN;
  p.cin_chunks = d->Cin / TF_BK;
  p.Cout = d->Cout;
  p.ld_out = d->ld_out;
  p.act = d->act;
  p.accumulate = d->accumulate;
  const bool dense = d->out_sw == 0 && d->out_sh == 0 && d->out_sb == 0;
  p.osw = dense ? 1 : d->out_sw;
  p.osh = dense ? d->Wo : d->out_sh;
  p.osb = dense ? d->Ho * d->Wo : d->out_sb;
  // instruction descriptor: D = f32 (1 << 4), A = B = tf32 (2 << 7, 2 << 10), K-major, N >> 3 at bit 17, M >> 4 at bit 24
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TF_BM >> 4) << 24);
  int n_taps = 0;
  for (int t = 0; t < d->n_taps; ++t) {
    const int rh = tap_dh[t], rw = tap_dw[t];
    // skip taps that only ever see padding
    if ((d->Ho - 1) * d->sh + rh < 0 || rh >= d->H || (d->Wo - 1) * d->sw + rw < 0 || rw >= d->W) continue;
    const int prh = ((rh % d->sh) + d->sh) % d->sh, prw = ((rw % d->sw) + d->sw) % d->sw;
    const int ddh = (rh - prh) / d->sh, ddw = (rw - prw) / d->sw;
    YAD_CHECK_ARG(ddh >= -128 && ddh <= 127 && ddw >= -128 && ddw <= 127, "yad_corr_tf32: tap offset out of range");
    YAD_CHECK_ARG(tap_k[t] >= 0 && tap_k[t] + d->Cin <= k_total && tap_k[t] % 4 == 0, "yad_corr_tf32: tap %d weight offset out of range", t);
    p.tap_map[n_taps] = (int8_t)(prh * d->sw + prw);
    p.tap_dh[n_taps] = (int8_t)ddh;
    p.tap_dw[n_taps] = (int8_t)ddw;
    p.tap_k[n_taps] = tap_k[t];
    ++n_taps;
  }
  cudaStream_t st = (cudaStream_t)stream;
  YAD_CHECK_ARG(n_taps > 0, "yad_corr_tf32: no tap touches the input");
  p.n_taps = n_taps;
  // small grids are latency bound: split the K loop over CTAs (partial sums red.add-ed into `out`) until the SMs are covered
  const int n_iters = n_taps * p.cin_chunks, tiles = p.n_wt * p.n_ht * p.n_bt * (cout_pad / BN);
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int ksplit = 1;
  if ((d->accumulate || d->whole_rows) && d->act == YAD_ACT_NONE && tiles < sms && stats == nullptr) {
    ksplit = sms / tiles;
    if (ksplit > n_iters / 2) ksplit = n_iters / 2;
    if (ksplit > 32) ksplit = 32;
    if (ksplit < 1) ksplit = 1;
  }
  p.iters_per_split = cdiv_(n_iters, ksplit);
  ksplit = cdiv_(n_iters, p.iters_per_split);
  p.ksplit = ksplit;
  if (ksplit > 1 && !d->accumulate) {
    YAD_CHECK_ARG(dense, "yad_corr_tf32: whole_rows needs a dense output");
    YAD_CUDA(cudaMemsetAsync(out, 0, (size_t)d->B * d->Ho * d->Wo * d->ld_out * sizeof(float), st));
  }
  const int stage_bytes = TF_A_STAGE_BYTES + BN * 128;
  // two resident CTAs per SM (<= ~100 KB each) when the grid is large, the whole shared memory for one CTA when it is not
  int stages = ((tiles * ksplit > sms ? 100 : 200) * 1024) / stage_bytes;
  if (stages < 3) stages = 3;
  if (stages > TF_MAX_STAGES) stages = TF_MAX_STAGES;
  if (stages > p.iters_per_split) stages = p.iters_per_split;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + (2 * TF_MAX_STAGES + 1) * 8 + 16;
  CUtensorMap maps[4];
  const uint32_t box[4] = {32u, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tb};
  int rc = build_parity_maps(in, d->B, d->H, d->W, d->Cin, d->ld_in, d->sh, d->sw, box, maps);
  if (rc) return rc;
  CUtensorMap map_w;
  {
    const uint64_t dims[2] = {(uint64_t)k_total, (uint64_t)cout_pad};
    const uint64_t strides[1] = {(uint64_t)k_total * 4};
    const uint32_t bx[2] = {32u, (uint32_t)BN};
    rc = encode_map_f32(&map_w, weight, 2, dims, strides, bx);
    if (rc) return rc;
  }
  dim3 grid((unsigned)(p.n_wt * p.n_ht * p.n_bt), (unsigned)(cout_pad / BN), (unsigned)ksplit);
  corr_tf32_kernel<<<grid, TF_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], map_w, p, bias, out, stats);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_wgrad_tf32(const yad_corr_desc* d, const int32_t* tap_dh, const int32_t* tap_dw, const int32_t* tap_dst, const float* x,
                   const float* dy, float* dw, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(d && tap_dh && tap_dw && tap_dst && x && dy && dw, "yad_wgrad_tf32: null pointer");
  YAD_CHECK_ARG(d->ld_in % 32 == 0 && d->ld_in >= d->Cin && d->Cin >= 1, "yad_wgrad_tf32: ld_in=%d must be a multiple of 32 >= Cin", d->ld_in);
  YAD_CHECK_ARG(d->ld_out % 32 == 0 && d->ld_out >= d->Cout && d->Cout >= 1, "yad_wgrad_tf32: ld_out=%d must be a multiple of 32 >= Cout", d->ld_out);
  YAD_CHECK_ARG((d->sh == 1 || d->sh == 2) && (d->sw == 1 || d->sw == 2), "yad_wgrad_tf32: stride (%d,%d) unsupported", d->sh, d->sw);
  YAD_CHECK_ARG(d->n_taps >= 1 && d->n_taps <= TF_MAX_TAPS, "yad_wgrad_tf32: %d taps (max %d)", d->n_taps, TF_MAX_TAPS);
  YAD_CHECK_ARG((reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(dy) % 16 == 0), "yad_wgrad_tf32: unaligned pointer");
  WgParams p;
  memset(&p, 0, sizeof(p));
  choose_box(d->B, d->Ho, d->Wo, 64, 8, &p.tb, &p.th, &p.tw);
  p.npix = p.tb * p.th * p.tw;
  p.n_wt = cdiv_(d->Wo, p.tw);
  p.n_ht = cdiv_(d->Ho, p.th);
  p.n_bt = cdiv_(d->B, p.tb);
  p.n_tiles = p.n_wt * p.n_ht * p.n_bt;
  const int cout32 = cdiv_(d->Cout, 32) * 32, cin32 = cdiv_(d->Cin, 32) * 32;
  int BN = cout32 >= 128 ? 128 : cout32;       // 32, 64, 96 -> 96 is not a power of two: fall back to 32-wide tiles
  if (BN == 96) BN = 32;
  const int n_blocks = cdiv_(cout32, BN);
  p.BN = BN;
  p.ci_blocks = cdiv_(cin32, 128);
  p.a_blocks = cin32 >= 128 ? 4 : cin32 / 32;
  p.Cin = d->Cin;
  p.Cout = d->Cout;
  const int blk_bytes = p.npix * 128;
  // M is always 128: with fewer than four 32-channel blocks the missing ones alias block 0 (their rows are never stored)
  p.a_lbo = p.a_blocks == 4 ? (uint32_t)blk_bytes : 0u;
  if (p.a_blocks == 2 || p.a_blocks == 3) p.a_lbo = (uint32_t)blk_bytes;     // blocks 2/3 read stale (finite or not) rows: discarded
  // D = f32, A = B = tf32, both MN-major (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TF_BM >> 4) << 24);
  int n_taps = 0;
  for (int t = 0; t < d->n_taps; ++t) {
    const int rh = tap_dh[t], rw = tap_dw[t];
    if ((d->Ho - 1) * d->sh + rh < 0 || rh >= d->H || (d->Wo - 1) * d->sw + rw < 0 || rw >= d->W) continue;   // gradient stays 0
    const int prh = ((rh % d->sh) + d->sh) % d->sh, prw = ((rw % d->sw) + d->sw) % d->sw;
    p.tap_map[n_taps] = (int8_t)(prh * d->sw + prw);
    p.tap_dh[n_taps] = (int8_t)((rh - prh) / d->sh);
    p.tap_dw[n_taps] = (int8_t)((rw - prw) / d->sw);
    YAD_CHECK_ARG(tap_dst[t] >= 0 && tap_dst[t] < 128, "yad_wgrad_tf32: bad destination tap");
    p.tap_dst[n_taps] = (int8_t)tap_dst[t];
    ++n_taps;
  }
  if (n_taps == 0) return YAD_OK;
  const int stage_bytes = (4 + BN / 32) * blk_bytes;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > TF_MAX_STAGES) stages = TF_MAX_STAGES;
  YAD_CHECK_ARG(stages >= 2, "yad_wgrad_tf32: pixel box too large for shared memory");
  const int items = n_taps * p.ci_blocks * n_blocks;
  int chunks = cdiv_(2 * (sm_count() > 0 ? sm_count() : 148), items);
  if (chunks > p.n_tiles) chunks = p.n_tiles;
  if (chunks < 1) chunks = 1;
  p.tiles_per_chunk = cdiv_(p.n_tiles, chunks);
  chunks = cdiv_(p.n_tiles, p.tiles_per_chunk);
  if (stages > p.tiles_per_chunk) stages = p.tiles_per_chunk < 2 ? 2 : p.tiles_per_chunk;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + (2 * TF_MAX_STAGES + 1) * 8 + 16;
  CUtensorMap maps[4], map_dy;
  const uint32_t box[4] = {32u, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tb};
  int rc = build_parity_maps(x, d->B, d->H, d->W, d->ld_in, d->ld_in, d->sh, d->sw, box, maps, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  {
    const uint64_t dims[4] = {(uint64_t)d->ld_out, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->ld_out * 4, (uint64_t)d->Wo * d->ld_out * 4, (uint64_t)d->Ho * d->Wo * d->ld_out * 4};
    rc = encode_map_f32(&map_dy, dy, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  }
  YAD_CHECK_ARG(n_taps * p.ci_blocks <= 65535 && n_blocks <= 65535, "yad_wgrad_tf32: grid too large");
  dim3 grid((unsigned)chunks, (unsigned)(n_taps * p.ci_blocks), (unsigned)n_blocks);
  wgrad_tf32_kernel<<<grid, TF_THREADS, smem, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], maps[3], map_dy, p, dw);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

}  // extern "C"
