// tcgen05 / TMEM implicit-GEMM convolution for sm_100a, TMA-fed (bf16 in, fp32 accumulate).
//
// Replaces F.conv2d + folded BatchNorm + activation (+ residual) of the reference's backbone and
// neck (modules/_backbone.py:143-151, torchvision resnet.py:89-105, modules/_common.py:43-48,86-95).
//
// GEMM view:  D[m, n] = sum_k A[m, k] * Wt[n, k]
//   m  = output pixel inside a (tb x th x tw) box of the NHWC output (<= 128 rows per CTA tile)
//   n  = output channel (tile BN <= 256)
//   k  = (tap, cin) ; one K-block = one filter tap x 64 input channels = one 128-byte swizzled row
//
// A operand: 4-D TMA (C, W, H, B) box {64, tw, th, tb} at the tap-shifted coordinate; out-of-range
// coordinates are zero-filled by the TMA unit = the convolution's zero padding.  Stride-2 convs use
// one tensor map per input parity class (base pointer offset, doubled strides) so no element-stride
// feature is needed.  The box lands in shared memory as 128 rows x 128 B with the 128B swizzle, which
// is exactly the canonical K-major SWIZZLE_128B UMMA layout (SBO = 1024 B).
// B operand: 2-D TMA over the packed weight [Cout_pad][taps*Cin] (K-major), box {64, BN}.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
// tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld -> bias/residual/activation -> global).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <string.h>

namespace yad {

constexpr int TC_THREADS = 192;
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                       // bf16 elements per K-block (128 B)
constexpr int TC_A_STAGE_BYTES = TC_BM * 128;   // 16 KB
constexpr int TC_MAX_TAPS = 49;
constexpr int TC_MAX_STAGES = 8;

struct TcParams {
  int32_t B, Ho, Wo;
  int32_t tb, th, tw;            // tile box
  int32_t n_wt, n_ht, n_bt;      // tiles along W, H, B
  int32_t BN, n_ntiles, stages;
  int32_t cin_chunks;            // Cin / 64
  int32_t n_taps;
  int32_t Cout, ld_out, co_off, ld_res, act, out_f32, ld_out2;
  uint32_t idesc;
  // per tap: which parity map, coordinate offsets (in the parity map's units), weight tap index
  int8_t tap_map[TC_MAX_TAPS];
  int8_t tap_dw[TC_MAX_TAPS];
  int8_t tap_dh[TC_MAX_TAPS];
  int8_t tap_widx[TC_MAX_TAPS];
  int32_t cin;                   // K elements per tap in the packed weight
  int32_t osw, osh, osb;         // output pixel strides
  // fused 1x1 downsample branch (torchvision BasicBlock.downsample, resnet.py:99-100): the strided 1x1 convolution of the same
  // input reads exactly the pixels of the main convolution's centre tap, so it runs as ds_chunks extra K-blocks (tap entry
  // n_taps, weights map_w2) into a second accumulator at TMEM column BN; 0 = no second branch
  int32_t ds_chunks, ds_cin, act2;
};

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
//   start address >> 4 | LBO(16 B units, ignored for swizzled K-major) << 16 | SBO = 1024 B (8 rows x 128 B) << 32
//   | version 1 << 46 | layout SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ---------------------------------------------------------------- kernel
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
               const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w2, const TcParams p,
               const float* __restrict__ bias, const __nv_bfloat16* __restrict__ residual, void* __restrict__ out,
               float* __restrict__ out2, const float* __restrict__ bias_ds, __nv_bfloat16* __restrict__ out_ds) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB][B BN*128] | barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_stage_bytes = p.BN * 128;
  const int stage_bytes = TC_A_STAGE_BYTES + b_stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + TC_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + TC_MAX_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const int n_main = p.n_taps * p.cin_chunks;
  const int n_iters = n_main + p.ds_chunks;

  // tile coordinates
  int t = blockIdx.x;
  const int wt = t % p.n_wt;
  t /= p.n_wt;
  const int ht = t % p.n_ht;
  t /= p.n_ht;
  const int bt = t;
  const int w0 = wt * p.tw, h0 = ht * p.th, b0 = bt * p.tb;
  const int n0 = blockIdx.y * p.BN;

  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < (p.ds_chunks ? 2 * p.BN : p.BN)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a0);
    prefetch_tmap(&map_w);
    if (p.ds_chunks) prefetch_tmap(&map_w2);
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
  // programmatic dependent launch: everything above overlapped the previous kernel's tail; its results are visible from here
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)(p.tb * p.th * p.tw * 128 + b_stage_bytes);
      uint32_t s = 0, phase = 0;     // ring (slot, phase): no integer division on the single-thread path
      for (int tap = 0; tap < p.n_taps + (p.ds_chunks ? 1 : 0); ++tap) {
        const int mi = p.tap_map[tap];
        const CUtensorMap* ma = mi == 0 ? &map_a0 : (mi == 1 ? &map_a1 : (mi == 2 ? &map_a2 : &map_a3));
        const CUtensorMap* mw = tap < p.n_taps ? &map_w : &map_w2;
        const int cw = w0 + p.tap_dw[tap], ch = h0 + p.tap_dh[tap];
        const int kbase = p.tap_widx[tap] * p.cin;
        for (int cc = 0; cc < p.cin_chunks; ++cc) {
          mbar_wait(&empty_bar[s], phase ^ 1);      // free on the first lap
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          uint8_t* sb = sa + TC_A_STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], tx_bytes);
          tma_load_4d(ma, &full_bar[s], sa, cc * TC_BK, cw, ch, b0);
          tma_load_2d(mw, &full_bar[s], sb, kbase + cc * TC_BK, n0);
          if (++s == (uint32_t)S) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      uint32_t s = 0, phase = 0;
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(&full_bar[s], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t sb = sa + TC_A_STAGE_BYTES;
        const uint64_t da = make_sw128_desc(sa), db = make_sw128_desc(sb);
        const bool is_ds = it >= n_main;                     // K-blocks of the fused downsample branch: second accumulator
        const uint32_t acc = tmem_base + (is_ds ? (uint32_t)p.BN : 0u);
        const bool first = it == 0 || it == n_main;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the swizzled 128 B row: +2 in the 16-byte address field
          umma_bf16(acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), p.idesc, (!first || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);                       // frees the smem stage when these MMAs retire
        if (it == n_iters - 1) umma_commit(tmem_full_bar);  // accumulator complete
        if (++s == (uint32_t)S) { s = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: 4 warps, one TMEM lane quadrant each =====================
    const int q = warp & 3;                 // TMEM lanes [32q, 32q+32) are accessible to this warp
    const int r = q * 32 + lane;            // tile row == TMEM lane
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int rows = p.tb * p.th * p.tw;
    const int wl = r % p.tw;
    const int hl = (r / p.tw) % p.th;
    const int bl = r / (p.tw * p.th);
    const int ow = w0 + wl, oh = h0 + hl, ob = b0 + bl;
    const bool row_ok = (r < rows) && (ow < p.Wo) && (oh < p.Ho) && (ob < p.B);
    const int64_t pix = (int64_t)ob * p.osb + (int64_t)oh * p.osh + (int64_t)ow * p.osw;
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (row_ok) {
        const int nbase = n0 + c0;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (nbase + 32 <= p.Cout) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __ldg(bias + nbase + j);
          if (residual != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(residual + pix * p.ld_res + nbase);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint4 u = __ldg(rp + j4);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                f[j4 * 8 + e * 2 + 0] += __uint_as_float(w[e] << 16);
                f[j4 * 8 + e * 2 + 1] += __uint_as_float(w[e] & 0xffff0000u);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
          if (out2 != nullptr) {
            float4* o2 = reinterpret_cast<float4*>(out2 + pix * p.ld_out2 + nbase);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) o2[j4] = make_float4(f[j4 * 4], f[j4 * 4 + 1], f[j4 * 4 + 2], f[j4 * 4 + 3]);
          }
          if (p.out_f32) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + pix * p.ld_out + p.co_off + nbase);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) op[j4] = make_float4(f[j4 * 4], f[j4 * 4 + 1], f[j4 * 4 + 2], f[j4 * 4 + 3]);
          } else {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + pix * p.ld_out + p.co_off + nbase);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(f[j4 * 8 + e * 2], f[j4 * 8 + e * 2 + 1]);
                w[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
              op[j4] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        } else {
          // ragged channel tail (Cout not a multiple of 32): scalar path
          for (int j = 0; j < 32; ++j) {
            const int n = nbase + j;
            if (n >= p.Cout) break;
            float x = f[j] + __ldg(bias + n);
            if (residual != nullptr) x += __bfloat162float(residual[pix * p.ld_res + n]);
            x = apply_act(x, p.act);
            if (out2 != nullptr) out2[pix * p.ld_out2 + n] = x;
            if (p.out_f32)
              reinterpret_cast<float*>(out)[pix * p.ld_out + p.co_off + n] = x;
            else
              reinterpret_cast<__nv_bfloat16*>(out)[pix * p.ld_out + p.co_off + n] = __float2bfloat16_rn(x);
          }
        }
      }
    }
    if (p.ds_chunks) {        // downsample branch: bias, activation act2, bf16, same pixel strides and row pitch as the main output
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.BN + c0), v);
        tmem_ld_wait();
        const int nbase = n0 + c0;
        if (row_ok && nbase + 32 <= p.Cout) {
          uint4* op = reinterpret_cast<uint4*>(out_ds + pix * p.ld_out + nbase);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = j4 * 8 + e * 2;
              const float x0 = apply_act(__uint_as_float(v[j]) + __ldg(bias_ds + nbase + j), p.act2);
              const float x1 = apply_act(__uint_as_float(v[j + 1]) + __ldg(bias_ds + nbase + j + 1), p.act2);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
              w[e] = *reinterpret_cast<uint32_t*>(&h2);
            }
            op[j4] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_map_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(tensor_map_encode_fn());
  if (!fn) {
    set_error("conv_tc: yad_init() was not called (cuTensorMapEncodeTiled unresolved)");
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
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return YAD_ERR_CUDA;
  }
  return YAD_OK;
}

int init_conv_tc_attrs() {
  cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_tc_kernel) failed: %s", cudaGetErrorString(e));
    return YAD_ERR_CUDA;
  }
  return YAD_OK;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// choose the (tb, th, tw) output box with tb*th*tw <= 128 maximising useful rows per MMA
static void choose_tile(int B, int Ho, int Wo, int* tb, int* th, int* tw) {
  double best = -1.0;
  int bb = 1, bh = 1, bw = 1;
  for (int w = 1; w <= Wo && w <= 128; ++w) {
    for (int h = 1; h <= Ho && h * w <= 128; ++h) {
      for (int b = 1; b <= B && b * h * w <= 128; ++b) {
        const double cover = (double)B * Ho * Wo / ((double)cdiv(Wo, w) * w * cdiv(Ho, h) * h * cdiv(B, b) * b);
        const double eff = cover * (b * h * w) / 128.0 + 1e-6 * w;  // tie-break: wider rows (longer TMA bursts)
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

}  // namespace yad

static int conv_tc_impl(const yad_conv_desc* d, const void* in, const void* weight, int32_t cout_pad,
                        const float* bias, const void* residual, void* out, int32_t out_dtype,
                        float* out2_f32, int32_t ld_out2, const void* weight_ds, const float* bias_ds, int32_t act_ds, void* out_ds,
                        yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(d && in && weight && bias && out, "yad_conv_tc: null pointer");
  YAD_CHECK_ARG(d->Cin % 64 == 0 && d->Cin >= 64, "yad_conv_tc: Cin=%d must be a multiple of 64 (zero-pad channels)", d->Cin);
  YAD_CHECK_ARG(d->ld_in % 8 == 0 && d->ld_in >= d->Cin, "yad_conv_tc: ld_in=%d must be >= Cin and a multiple of 8", d->ld_in);
  YAD_CHECK_ARG(cout_pad % 16 == 0 && cout_pad >= d->Cout, "yad_conv_tc: cout_pad=%d must be a multiple of 16 >= Cout", cout_pad);
  YAD_CHECK_ARG((d->sh == 1 || d->sh == 2) && (d->sw == 1 || d->sw == 2), "yad_conv_tc: stride (%d,%d) unsupported", d->sh, d->sw);
  YAD_CHECK_ARG(d->kh * d->kw <= TC_MAX_TAPS && d->kh >= 1 && d->kw >= 1, "yad_conv_tc: kernel %dx%d unsupported", d->kh, d->kw);
  YAD_CHECK_ARG(out_dtype == YAD_BF16 || out_dtype == YAD_F32, "yad_conv_tc: bad out dtype");
  YAD_CHECK_ARG(d->ld_out >= d->co_off + d->Cout, "yad_conv_tc: ld_out too small");
  YAD_CHECK_ARG((d->ld_out % 8 == 0) && (d->co_off % 8 == 0), "yad_conv_tc: ld_out/co_off must be multiples of 8 (16 B stores)");
  YAD_CHECK_ARG(residual == nullptr || (d->ld_res % 8 == 0 && d->ld_res >= d->Cout), "yad_conv_tc: bad ld_res");
  YAD_CHECK_ARG(out2_f32 == nullptr || (ld_out2 % 4 == 0 && ld_out2 >= d->Cout), "yad_conv_tc: bad ld_out2");
  YAD_CHECK_ARG((reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(weight) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(out) % 16 == 0),
                "yad_conv_tc: pointers must be 16-byte aligned");
  const int Ho = (d->H + 2 * d->ph - d->kh) / d->sh + 1;
  const int Wo = (d->W + 2 * d->pw - d->kw) / d->sw + 1;
  YAD_CHECK_ARG(Ho >= 1 && Wo >= 1 && d->B >= 1, "yad_conv_tc: empty output");
  const bool in_dense = d->in_sw == 0 && d->in_sh == 0 && d->in_sb == 0;
  const bool out_dense = d->out_sw == 0 && d->out_sh == 0 && d->out_sb == 0;
  const int64_t isw = in_dense ? 1 : d->in_sw, ish = in_dense ? d->W : d->in_sh, isb = in_dense ? (int64_t)d->H * d->W : d->in_sb;
  YAD_CHECK_ARG(isw >= 1 && ish >= isw && isb >= ish, "yad_conv_tc: input pixel strides must satisfy 1 <= in_sw <= in_sh <= in_sb");

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B;
  p.Ho = Ho;
  p.Wo = Wo;
  choose_tile(d->B, Ho, Wo, &p.tb, &p.th, &p.tw);
  p.n_wt = cdiv(Wo, p.tw);
  p.n_ht = cdiv(Ho, p.th);
  p.n_bt = cdiv(d->B, p.tb);
  int BN = cout_pad >= 128 ? 128 : cout_pad;
  if (cout_pad % BN != 0) BN = 64;
  if (cout_pad % BN != 0) BN = 32;
  if (cout_pad % BN != 0) BN = 16;
  p.BN = BN;
  p.n_ntiles = cout_pad / BN;
  p.cin_chunks = d->Cin / 64;
  p.cin = d->Cin;
  p.Cout = d->Cout;
  p.ld_out = d->ld_out;
  p.co_off = d->co_off;
  p.ld_res = d->ld_res;
  p.act = d->act;
  p.out_f32 = (out_dtype == YAD_F32);
  p.ld_out2 = ld_out2;
  p.osw = out_dense ? 1 : d->out_sw;
  p.osh = out_dense ? Wo : d->out_sh;
  p.osb = out_dense ? Ho * Wo : d->out_sb;
  // instruction descriptor: D=f32 (1<<4), A=bf16 (1<<7), B=bf16 (1<<10), K-major A and B, N>>3 at bit 17, M>>4 at bit 24
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  // taps: skip those whose input rows / columns are entirely padding for every output position
  int n_taps = 0;
  for (int kh = 0; kh < d->kh; ++kh) {
    const int lo = kh - d->ph, hi = (Ho - 1) * d->sh + kh - d->ph;
    if (hi < 0 || lo >= d->H) continue;
    for (int kw = 0; kw < d->kw; ++kw) {
      const int lo2 = kw - d->pw, hi2 = (Wo - 1) * d->sw + kw - d->pw;
      if (hi2 < 0 || lo2 >= d->W) continue;
      const int rh = kh - d->ph, rw = kw - d->pw;
      const int prh = ((rh % d->sh) + d->sh) % d->sh, prw = ((rw % d->sw) + d->sw) % d->sw;
      p.tap_map[n_taps] = (int8_t)(prh * d->sw + prw);
      p.tap_dh[n_taps] = (int8_t)((rh - prh) / d->sh);
      p.tap_dw[n_taps] = (int8_t)((rw - prw) / d->sw);
      p.tap_widx[n_taps] = (int8_t)(kh * d->kw + kw);
      ++n_taps;
    }
  }
  YAD_CHECK_ARG(n_taps >= 1, "yad_conv_tc: no tap touches the input");
  p.n_taps = n_taps;
  if (weight_ds != nullptr) {
    // the 1x1 stride-(sh, sw) pad-0 convolution of the same input = the centre tap (kh, kw) = (ph, pw) of the main convolution
    YAD_CHECK_ARG(bias_ds && out_ds && d->ph < d->kh && d->pw < d->kw && n_taps < TC_MAX_TAPS && cout_pad % 32 == 0 && d->Cout == cout_pad &&
                      out_dtype == YAD_BF16 && d->co_off == 0 && reinterpret_cast<uintptr_t>(weight_ds) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(out_ds) % 16 == 0,
                  "yad_conv_tc_dual: needs a centre tap, Cout == cout_pad (a multiple of 32), bf16 output at co_off 0, aligned pointers");
    int centre = -1;
    for (int t = 0; t < n_taps; ++t)
      if (p.tap_widx[t] == d->ph * d->kw + d->pw) centre = t;
    YAD_CHECK_ARG(centre >= 0, "yad_conv_tc_dual: the centre tap is not part of the convolution");
    p.tap_map[n_taps] = p.tap_map[centre];
    p.tap_dh[n_taps] = p.tap_dh[centre];
    p.tap_dw[n_taps] = p.tap_dw[centre];
    p.tap_widx[n_taps] = 0;
    p.ds_chunks = p.cin_chunks;
    p.ds_cin = d->Cin;
    p.act2 = act_ds;
  }

  // pipeline depth: aim for two resident CTAs per SM (<= ~100 KB each) but at least 3 stages
  const int stage_bytes = TC_A_STAGE_BYTES + BN * 128;
  int stages = (100 * 1024) / stage_bytes;
  if (stages < 3) stages = 3;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages > n_taps * p.cin_chunks) stages = n_taps * p.cin_chunks;
  if (stages < 1) stages = 1;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 /*align slack*/ + (2 * TC_MAX_STAGES + 1) * 8 + 16;

  // tensor maps: one per input parity class
  CUtensorMap maps[4];
  const int n_maps = d->sh * d->sw;
  const uint32_t box[4] = {64u, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tb};
  for (int prh = 0; prh < d->sh; ++prh) {
    for (int prw = 0; prw < d->sw; ++prw) {
      const int mi = prh * d->sw + prw;
      const int Wm = (d->W - prw + d->sw - 1) / d->sw, Hm = (d->H - prh + d->sh - 1) / d->sh;
      const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(in) + ((int64_t)prh * ish + prw * isw) * d->ld_in;
      if (Wm <= 0 || Hm <= 0) {  // parity class with no element (e.g. H == 1, odd rows): alias map 0, never referenced
        maps[mi] = maps[0];
        continue;
      }
      const uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)Wm, (uint64_t)Hm, (uint64_t)d->B};
      const uint64_t strides[3] = {(uint64_t)d->sw * isw * d->ld_in * 2, (uint64_t)d->sh * ish * d->ld_in * 2,
                                   (uint64_t)isb * d->ld_in * 2};
      int rc = encode_map_bf16(&maps[mi], base, 4, dims, strides, box);
      if (rc) return rc;
    }
  }
  for (int mi = n_maps; mi < 4; ++mi) maps[mi] = maps[0];
  // every referenced parity map must be non-empty
  CUtensorMap map_w;
  {
    const uint64_t dims[2] = {(uint64_t)d->kh * d->kw * d->Cin, (uint64_t)cout_pad};
    const uint64_t strides[1] = {(uint64_t)d->kh * d->kw * d->Cin * 2};
    const uint32_t bx[2] = {64u, (uint32_t)BN};
    int rc = encode_map_bf16(&map_w, weight, 2, dims, strides, bx);
    if (rc) return rc;
  }
  CUtensorMap map_w2 = map_w;
  if (weight_ds != nullptr) {
    const uint64_t dims[2] = {(uint64_t)d->Cin, (uint64_t)cout_pad};
    const uint64_t strides[1] = {(uint64_t)d->Cin * 2};
    const uint32_t bx[2] = {64u, (uint32_t)BN};
    int rc = encode_map_bf16(&map_w2, weight_ds, 2, dims, strides, bx);
    if (rc) return rc;
  }
  dim3 grid((unsigned)(p.n_wt * p.n_ht * p.n_bt), (unsigned)p.n_ntiles);
  YAD_CUDA(launch_pdl(conv_tc_kernel, grid, dim3(TC_THREADS), smem, (cudaStream_t)stream, maps[0], maps[1], maps[2], maps[3], map_w, map_w2, p,
                      bias, reinterpret_cast<const __nv_bfloat16*>(residual), out, out2_f32, bias_ds,
                      reinterpret_cast<__nv_bfloat16*>(out_ds)));
  return YAD_OK;
}

extern "C" int yad_conv_tc(const yad_conv_desc* d, const void* in, const void* weight, int32_t cout_pad,
                           const float* bias, const void* residual, void* out, int32_t out_dtype,
                           float* out2_f32, int32_t ld_out2, yad_stream_t stream) {
  return conv_tc_impl(d, in, weight, cout_pad, bias, residual, out, out_dtype, out2_f32, ld_out2, nullptr, nullptr, 0, nullptr, stream);
}

extern "C" int yad_conv_tc_dual(const yad_conv_desc* d, const void* in, const void* weight, int32_t cout_pad, const float* bias,
                                void* out, const void* weight_ds, const float* bias_ds, int32_t act_ds, void* out_ds,
                                yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(weight_ds && bias_ds && out_ds, "yad_conv_tc_dual: null pointer");
  return conv_tc_impl(d, in, weight, cout_pad, bias, nullptr, out, YAD_BF16, nullptr, 0, weight_ds, bias_ds, act_ds, out_ds, stream);
}
