// Stem conv1 o conv2 as ONE convolution on the tcgen05 tensor cores.
//
// The reference runs conv1 (2 -> 64, 7x7, stride 2, pad 3, no bias) and conv2 (64 -> 64, 7x7, stride 2, pad 3) back to back with
// nothing in between (modules/_backbone.py:143-146: BatchNorm + ReLU come after conv2), i.e. two LINEAR maps.  Their composition
// is a single 19x19 / stride-4 / pad-9 convolution 2 -> 64 with W12 = conv_transpose(W2, W1) (built in fp64 by the host): 722
// taps per output pixel instead of 98 x 4 + 3136, and the 503 MB intermediate tensor (B = 512) is never written or read.
// The only subtlety is the zero padding of the INTERMEDIATE tensor: conv2 taps that fall outside conv1's output are dropped by
// the reference, so output rows 0, 1 and Ho-1 (and columns 0, 1 and the last one or two) need composite weights built from the
// valid conv2 taps only.  Rows: the CTAs are split into four classes (interior rows 2..6 | row 0 | row 1 | row 7), each keeping
// its own weight variant resident in shared memory.  Columns: this kernel uses the interior-column weights everywhere and
// stem_fixup_kernel below recomputes the handful of border columns exactly (CUDA cores, < 2 % of the pixels).
//
// GEMM view per output row ho and 128 consecutive output pixels: D[128, 64] = sum over the kernel rows dh whose input row
// hi = 4 ho - 9 + dh is inside the image of  A_dh[128, 48] * W_dh[64, 48]^T,  K index = dw * 2 + c (38 real + 10 zero-weight).
// As in conv_stem_tc.cu the A operand is NOT an im2col copy: the patch lives in shared memory as channel-interleaved bf16 rows
// (one 4-byte word per input pixel; frontend stage B writes x_spectral a second time in exactly this form, with zero margins, so
// a patch row is ONE 2144-byte bulk copy) and the no-swizzle K-major UMMA descriptor (LBO = 16 B, SBO = 128 B) reads operand row r
// at byte 16 r of the patch row - exactly the stride-4 window of output pixel r.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace yad {

constexpr int SF_EPI_WARPS = 8;                // warps 0..7: epilogue; warp 8: MMA issue
constexpr int SF_THREADS = 32 * (SF_EPI_WARPS + 2);   // + warp 9: patch producer
constexpr int SF_MAXGRP = 12;                  // row groups of a patch (2 * rows of a class + 1 at most)
constexpr int SF_NACC = 5;                     // accumulator slots (output rows) in TMEM: 5 x 64 columns
constexpr int SF_SEG = 128;                    // output pixels (columns) per tile = MMA M
constexpr int SF_KD = 19;                      // composite kernel size
constexpr int SF_KROW = 48;                    // K per kernel row: 19 taps x 2 channels = 38, padded to 3 MMA steps of 16
constexpr int SF_K = SF_KD * SF_KROW;          // 912
constexpr int SF_PW = 4 * (SF_SEG - 1) + SF_KROW / 2 + 4;   // 536 words per patch row (last window ends at word 4*127 + 24)
constexpr int SF_ROWB = SF_PW * 4;             // 2144 B, multiple of 16
constexpr int SF_MAXROWS = 35;                 // input rows of the interior class: 4*2-9 .. 4*6-9+18 -> -1 .. 33
constexpr int SF_SBO_W = (SF_K / 8) * 128;     // weights: bytes between 8-row (cout) groups
constexpr int SF_B_BYTES = 8 * SF_SBO_W;       // 116736
constexpr int SF_TMEM_COLS = 512;              // up to 5 accumulators of 64 columns

__device__ __forceinline__ uint64_t sf_nosw_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

struct StemFusedParams {
  int32_t B, H, W, Ho, Wo, n_seg;
  int32_t Hp, Wp;                // flat output pitches: pixel (b, h, w) at ((b * Wp + w) * Hp + h) * 64
  int64_t xpitch;                // words per row of the padded bf16 input
  int32_t cta_first[5];          // CTA ranges of the 4 row classes: class c owns blocks [cta_first[c], cta_first[c+1])
  int32_t row_first[4], row_cnt[4];
  uint32_t idesc;
  int32_t dbg;                   // timing experiments (YAD_STEM_DBG; results are wrong with any bit set): 1 no stores, 2 no MMAs, 4 no fills
  int32_t skip_lo, skip_hi;      // output columns [0, skip_lo) and [Wo - skip_hi, Wo) are NOT written: the fix-up kernel owns them
                                 // and may then run concurrently (another stream) instead of after this kernel
};

// 1-D bulk copy global -> shared (TMA unit), completion counted on an mbarrier
__device__ __forceinline__ void sf_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(SF_THREADS, 1)
conv_stem_fused_kernel(const uint32_t* __restrict__ xb, const StemFusedParams p, const uint4* __restrict__ w_classes,
                       const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t sf_smem[];
  uint8_t* sB = sf_smem;                                   // this class's weights, core-matrix layout
  uint8_t* sP = sB + SF_B_BYTES;                           // [rows][SF_ROWB] patch
  float* s_bias = reinterpret_cast<float*>(sP + SF_MAXROWS * SF_ROWB + 256);   // +256: slack read by the last row's last windows
  uint64_t* grp_full = reinterpret_cast<uint64_t*>(s_bias + 64);          // [2][SF_MAXGRP] row group g of patch buffer b has landed
  uint64_t* row_done = grp_full + 2 * SF_MAXGRP;                           // [2][SF_NACC] the MMAs of output row j (buffer b) have retired
  uint64_t* acc_full = row_done + 2 * SF_NACC;                             // [SF_NACC] accumulator slot written by the tensor core
  uint64_t* acc_empty = acc_full + SF_NACC;                                // [SF_NACC] ... drained by the 8 epilogue warps
  uint64_t* w_bar = acc_empty + SF_NACC;                                   // the class's weights have landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_bar + 1);
  int32_t* s_grp = reinterpret_cast<int32_t*>(tmem_ptr_smem + 2);          // [SF_MAXGRP][4]: first row, end row, first / last output row using it

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int cls = 0;
  while (cls < 3 && (int)blockIdx.x >= p.cta_first[cls + 1]) ++cls;
  const int cta_in_cls = blockIdx.x - p.cta_first[cls], n_cta_cls = p.cta_first[cls + 1] - p.cta_first[cls];
  const int ho0 = p.row_first[cls], nrow = p.row_cnt[cls];
  // input rows this class touches: [hi_lo, hi_hi)
  const int hi_lo = max(4 * ho0 - 9, 0), hi_hi = min(4 * (ho0 + nrow - 1) - 9 + SF_KD, p.H);
  const int prow = hi_hi - hi_lo;

  const uint4* wsrc = w_classes + (size_t)cls * (SF_B_BYTES / 16);
  if (tid < 64) s_bias[tid] = bias[tid];
  // Row groups of the patch: maximal runs of input rows used by the same set of output rows of the class (interior class: 9 groups
  // of 2 - 4 rows; a border class: one).  Group g is first needed by output row F(g) and free again once row L(g)'s MMAs have
  // retired, so the producer refills the patch of the NEXT tile group by group while this tile's later rows are still being
  // multiplied: only the 3 rows shared by the first and the last output row sit between two tiles (the whole 69 KB patch did).
  int n_grp = 0;
  {
    int lo = hi_lo;
    while (lo < hi_hi) {
      // next boundary: the smallest window start / window end above lo
      int nxt = hi_hi;
      for (int j = 0; j < nrow; ++j) {
        const int a = 4 * (ho0 + j) - 9, e = a + SF_KD;
        if (a > lo && a < nxt) nxt = a;
        if (e > lo && e < nxt) nxt = e;
      }
      int F = nrow, L = -1;
      for (int j = 0; j < nrow; ++j) {
        const int a = 4 * (ho0 + j) - 9;
        if (a <= lo && lo < a + SF_KD) { F = j < F ? j : F; L = j; }
      }
      if (tid == 0) {
        s_grp[4 * n_grp] = lo - hi_lo;
        s_grp[4 * n_grp + 1] = nxt - hi_lo;
        s_grp[4 * n_grp + 2] = F;
        s_grp[4 * n_grp + 3] = L;
      }
      ++n_grp;
      lo = nxt;
    }
  }
  if (tid == 0) {
    for (int i = 0; i < 2 * SF_MAXGRP; ++i) mbar_init(&grp_full[i], 1);
    for (int i = 0; i < 2 * SF_NACC; ++i) mbar_init(&row_done[i], 1);
    for (int i = 0; i < SF_NACC; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], SF_EPI_WARPS);      // one arrival per epilogue warp
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
    // this class's 114 KB of weights: bulk copies (TMA unit) instead of 23 LDG + STS rounds of the whole CTA; launch constants,
    // so they are fetched before the programmatic-dependent-launch wait
    mbar_expect_tx(w_bar, (uint32_t)SF_B_BYTES);
    constexpr uint32_t W_CHUNK = 14592;                                  // SF_B_BYTES / 8
    static_assert(SF_B_BYTES % W_CHUNK == 0 && W_CHUNK % 16 == 0, "weight chunks");
    for (uint32_t o = 0; o < (uint32_t)SF_B_BYTES; o += W_CHUNK)
      sf_bulk_g2s(sB + o, reinterpret_cast<const uint8_t*>(wsrc) + o, W_CHUNK, w_bar);
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, SF_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp == SF_EPI_WARPS) mbar_wait(w_bar, 0);     // only the MMA warp reads the weights (through the async proxy, like the copy)
  pdl_wait();          // programmatic dependent launch: the 117 KB weight load above overlapped the frontend's tail
  pdl_trigger();
  const int n_tiles = p.B * p.n_seg;

  // Round 2: the tile loop is a pipeline over ACCUMULATOR SLOTS (one output row of 128 pixels x 64 channels each, 5 of them in
  // TMEM).  Warp 8 fills the patch and issues the MMAs row by row, committing each row to its slot's barrier; the 8 epilogue
  // warps drain a slot as soon as it is complete and hand it back.  For the interior class (5 rows per tile) the epilogue of rows
  // 0..3 runs under the MMAs of rows 1..4 and the last row's under the next tile's patch copy and first row; for the one-row
  // border classes up to five tiles' accumulators are in flight.  Before (fill -> all MMAs -> all epilogues, one barrier each)
  // a 5-row tile took 15 us for 7 us of MMA.  The patch is single-buffered (75 KB next to 117 KB of weights): the copy of the
  // next tile starts when the last MMA of this tile has retired.
  // patch buffers: the one-row border classes need at most 14 input rows, so two of their patches fit the 35-row patch area and
  // the copy of tile t + 1 runs under the MMAs of tile t; the interior class (35 rows) has one buffer
  const int NP = 2 * prow <= SF_MAXROWS ? 2 : 1;
  const uint32_t buf_bytes = (uint32_t)prow * SF_ROWB;
  if (warp == SF_EPI_WARPS + 1) {
    // ===================================================================== patch producer (one lane)
    // the input lives as channel-interleaved bf16 words with a 9-word zero margin on the left (and zeros on the right), so the
    // window row of a segment is the 16-byte aligned span [4 wo0, 4 wo0 + SF_PW) of the padded row: one bulk copy per input
    // row, no bounds logic.  Rows outside the image are not loaded (their MMAs are skipped).
    if (lane == 0) {
      int it = 0;
      for (int tile = cta_in_cls; tile < n_tiles; tile += n_cta_cls, ++it) {
        const int pb = NP == 2 ? (it & 1) : 0, use = NP == 2 ? (it >> 1) : it;       // use-th fill of buffer pb
        const int sg = tile % p.n_seg, bb = tile / p.n_seg;
        const uint32_t* src = xb + ((int64_t)bb * p.H + hi_lo) * p.xpitch + 4 * sg * SF_SEG;
        uint8_t* dst = sP + (size_t)pb * buf_bytes;
        for (int g = 0; g < n_grp; ++g) {
          const int r0 = s_grp[4 * g], r1 = s_grp[4 * g + 1], L = s_grp[4 * g + 3];
          if (use > 0) mbar_wait(&row_done[pb * SF_NACC + L], (uint32_t)(use - 1) & 1u);   // its last reader of the previous tile is done
          uint64_t* fb = &grp_full[pb * SF_MAXGRP + g];
          mbar_expect_tx(fb, (uint32_t)(r1 - ((p.dbg & 4) ? r1 - 1 : r0)) * SF_ROWB);
          for (int rr = (p.dbg & 4) ? r1 - 1 : r0; rr < r1; ++rr) sf_bulk_g2s(dst + rr * SF_ROWB, src + (int64_t)rr * p.xpitch, SF_ROWB, fb);
        }
      }
    }
  } else if (warp == SF_EPI_WARPS) {
    // ===================================================================== MMA issue (whole warp walks, one lane issues)
    // A lone thread retires an instruction every 5 - 10 cycles and the tensor pipe queues one MMA behind the running one: the
    // descriptors are running sums (134 sixteen-byte units per patch row, 48 per kernel row of the weights), 2 adds per MMA -
    // the first version rebuilt each with shift / mask / or (13 instructions per MMA, 63 - 76 cycles per 48-cycle MMA).
    constexpr uint32_t A_HI = (128u >> 4) | (1u << 14);                   // SBO 128 B, descriptor version 1, no swizzle
    constexpr uint32_t B_HI = ((uint32_t)SF_SBO_W >> 4) | (1u << 14);
    const uint32_t a_lo0 = ((smem_u32(sP) & 0x3FFFFu) >> 4) | ((16u >> 4) << 16);      // LBO 16 B: overlapping 32-byte windows
    const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
    uint32_t slot = 0, sph = 0;
    int it = 0;
    for (int tile = cta_in_cls; tile < n_tiles; tile += n_cta_cls, ++it) {
      const int pb = NP == 2 ? (it & 1) : 0, use = NP == 2 ? (it >> 1) : it;
      const uint32_t a_buf = a_lo0 + (uint32_t)pb * (buf_bytes >> 4);
      int g = 0;
      for (int j = 0; j < nrow; ++j) {
        while (g < n_grp && s_grp[4 * g + 2] <= j) {                       // the row groups this output row is the first to read
          mbar_wait(&grp_full[pb * SF_MAXGRP + g], (uint32_t)use & 1u);
          ++g;
        }
        mbar_wait(&acc_empty[slot], sph ^ 1);      // drained by the epilogue (free on the first lap)
        tc_fence_after();
        if (elect_one()) {
          // one accumulator per output row; kernel rows whose input row is outside the image are skipped
          const int hi_base = 4 * (ho0 + j) - 9;
          const int dh_lo = hi_base < 0 ? -hi_base : 0, dh_hi = p.H - hi_base < SF_KD ? p.H - hi_base : SF_KD;
          uint32_t a_lo = a_buf + (uint32_t)(hi_base + dh_lo - hi_lo) * (SF_ROWB >> 4);
          uint32_t b_lo = b_lo0 + (uint32_t)dh_lo * 48u;                  // 3 K steps x 256 B per kernel row
          const uint32_t d = tmem_base + slot * 64u;
          uint32_t acc = 0u;
          for (int dh = (p.dbg & 2) ? dh_hi : dh_lo; dh < dh_hi; ++dh) {
            umma_bf16(d, ((uint64_t)A_HI << 32) | a_lo, ((uint64_t)B_HI << 32) | b_lo, p.idesc, acc);
            umma_bf16(d, ((uint64_t)A_HI << 32) | (a_lo + 2u), ((uint64_t)B_HI << 32) | (b_lo + 16u), p.idesc, 1u);
            umma_bf16(d, ((uint64_t)A_HI << 32) | (a_lo + 4u), ((uint64_t)B_HI << 32) | (b_lo + 32u), p.idesc, 1u);
            acc = 1u;
            a_lo += SF_ROWB >> 4;
            b_lo += 48u;
          }
          umma_commit(&acc_full[slot]);
          umma_commit(&row_done[pb * SF_NACC + j]);
        }
        __syncwarp();
        if (++slot == SF_NACC) { slot = 0; sph ^= 1; }
      }
    }
  } else {
    // ===================================================================== epilogue: bias + ReLU, bf16, flat halo layout
    // thread = (pixel r, 32-channel half); the rows of a class are adjacent in memory (h fastest)
    const int q = warp & 3, half = warp >> 2;     // TMEM lane quadrant, channel half
    const int r = q * 32 + lane;                  // pixel inside the segment
    uint32_t slot = 0, sph = 0;
    for (int tile = cta_in_cls; tile < n_tiles; tile += n_cta_cls) {
      const int seg = tile % p.n_seg, b = tile / p.n_seg;
      const int wo = seg * SF_SEG + r;
      const bool ok = wo < p.Wo - p.skip_hi && wo >= p.skip_lo && !(p.dbg & 1);
      for (int j = 0; j < nrow; ++j) {
        mbar_wait(&acc_full[slot], sph);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 64 + half * 32), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[slot])) : "memory");
        }
        if (ok) {
          uint4* op = reinterpret_cast<uint4*>(out + (((int64_t)b * p.Wp + wo) * p.Hp + (ho0 + j)) * 64 + half * 32);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = half * 32 + j4 * 8 + e * 2;
              const float f0 = fmaxf(__uint_as_float(v[j4 * 8 + e * 2]) + s_bias[c], 0.0f);
              const float f1 = fmaxf(__uint_as_float(v[j4 * 8 + e * 2 + 1]) + s_bias[c + 1], 0.0f);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
              w[e] = *reinterpret_cast<uint32_t*>(&h2);
            }
            op[j4] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        if (++slot == SF_NACC) { slot = 0; sph ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, SF_TMEM_COLS);
}

// Exact recomputation of the border COLUMNS (conv2 taps outside conv1's output dropped along W): one CTA per (clip, column),
// thread = (output row, output channel); fp32 on the CUDA cores.  w_var [n_var][4 row classes][19][19][2][64] f32.
struct StemFixupParams {
  int32_t B, H, W, Ho, Hp, Wp, n_cols;
  int32_t col[8], var[8];
  int32_t row_class[8];          // row class of output rows 0..7
};

constexpr int FX_CLIPS = 8;      // clips per CTA: every composite weight is fetched once per 8 clips
constexpr int FX_THREADS = 256;  // thread = (output row 0..7, channel quad 0..15, clip half 0..1): 4 channels x 4 clips of accumulators

__global__ void __launch_bounds__(FX_THREADS, 2)
stem_fixup_kernel(const float* __restrict__ x, const StemFixupParams p, const float* __restrict__ w_var,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
  __shared__ float s_x[2][32][SF_KD + 1][FX_CLIPS];        // clip fastest: the 4 clips of a thread are one 16-byte load
  const int ci = blockIdx.x, b0 = blockIdx.y * FX_CLIPS;
  const int nb = min(FX_CLIPS, p.B - b0);
  const int wo = p.col[ci], var = p.var[ci];
  const int tid = threadIdx.x;
  pdl_wait();          // the fused stem kernel wrote approximate border columns: this kernel must overwrite them afterwards
  pdl_trigger();
  for (int i = tid; i < FX_CLIPS * 2 * 32 * SF_KD; i += blockDim.x) {
    const int dw = i % SF_KD, hi = (i / SF_KD) % 32, c = (i / (SF_KD * 32)) % 2, bb = i / (SF_KD * 32 * 2);
    const int wi = 4 * wo - 9 + dw;
    s_x[c][hi][dw][bb] = (bb < nb && hi < p.H && wi >= 0 && wi < p.W) ? x[(((int64_t)(b0 + bb) * 2 + c) * p.H + hi) * p.W + wi] : 0.0f;
  }
  __syncthreads();
  const int q4 = (tid & 15) * 4, half = (tid >> 4) & 1, ho = tid >> 5;
  if (ho >= p.Ho) return;
  const float* wv = w_var + ((size_t)(var * 4 + p.row_class[ho]) * SF_KD * SF_KD * 2) * 64 + q4;
  float acc[4][4];     // [clip][channel]  (scalar FFMA on purpose: this loop is FMA-pipe bound, FFMA2 measured 40 % slower here)
#pragma unroll
  for (int bb = 0; bb < 4; ++bb)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[bb][e] = bias[q4 + e];
  for (int dh = 0; dh < SF_KD; ++dh) {
    const int hi = 4 * ho - 9 + dh;
    if (hi < 0 || hi >= p.H) continue;
    for (int dw = 0; dw < SF_KD; ++dw) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wv + ((dh * SF_KD + dw) * 2 + c) * 64));
        const float4 xv = *reinterpret_cast<const float4*>(&s_x[c][hi][dw][half * 4]);
        const float xs4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          acc[bb][0] = fmaf(xs4[bb], w.x, acc[bb][0]);
          acc[bb][1] = fmaf(xs4[bb], w.y, acc[bb][1]);
          acc[bb][2] = fmaf(xs4[bb], w.z, acc[bb][2]);
          acc[bb][3] = fmaf(xs4[bb], w.w, acc[bb][3]);
        }
      }
    }
  }
#pragma unroll
  for (int bb = 0; bb < 4; ++bb) {
    const int bl = half * 4 + bb;
    if (bl >= nb) continue;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(fmaxf(acc[bb][0], 0.0f), fmaxf(acc[bb][1], 0.0f));
    __nv_bfloat162 h1 = __floats2bfloat162_rn(fmaxf(acc[bb][2], 0.0f), fmaxf(acc[bb][3], 0.0f));
    uint2 o = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    *reinterpret_cast<uint2*>(out + (((int64_t)(b0 + bl) * p.Wp + wo) * p.Hp + ho) * 64 + q4) = o;
  }
}

// The same fix-up reading the channel-interleaved bf16 copy of x_spectral that the tensor-core kernel reads (one 4-byte word per
// pixel, 9-word zero left margin: window column 4 wo - 9 + dw is word 4 wo + dw, never out of range): 20 KB of shared memory
// instead of 40, so that its CTAs fit next to the resident stem CTA (192 KB) and the two kernels overlap when launched on
// different streams (the stem kernel then skips these columns: skip_lo / skip_hi).
__global__ void __launch_bounds__(FX_THREADS, 2)
stem_fixup_bf16_kernel(const uint32_t* __restrict__ xb, int64_t xpitch, const StemFixupParams p, const float* __restrict__ w_var,
                       const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
  __shared__ __align__(16) uint32_t s_xw[32][SF_KD + 1][FX_CLIPS];        // clip fastest: the 4 clips of a thread are one 16-byte load
  const int ci = blockIdx.x, b0 = blockIdx.y * FX_CLIPS;
  const int nb = min(FX_CLIPS, p.B - b0);
  const int wo = p.col[ci], var = p.var[ci];
  const int tid = threadIdx.x;
  pdl_wait();
  pdl_trigger();
  for (int i = tid; i < FX_CLIPS * 32 * SF_KD; i += blockDim.x) {
    const int dw = i % SF_KD, hi = (i / SF_KD) % 32, bb = i / (SF_KD * 32);
    s_xw[hi][dw][bb] = (bb < nb && hi < p.H) ? __ldg(xb + ((int64_t)(b0 + bb) * p.H + hi) * xpitch + 4 * wo + dw) : 0u;
  }
  __syncthreads();
  const int q4 = (tid & 15) * 4, half = (tid >> 4) & 1, ho = tid >> 5;
  if (ho >= p.Ho) return;
  const float* wv = w_var + ((size_t)(var * 4 + p.row_class[ho]) * SF_KD * SF_KD * 2) * 64 + q4;
  float acc[4][4];
#pragma unroll
  for (int bb = 0; bb < 4; ++bb)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[bb][e] = bias[q4 + e];
  for (int dh = 0; dh < SF_KD; ++dh) {
    const int hi = 4 * ho - 9 + dh;
    if (hi < 0 || hi >= p.H) continue;
    for (int dw = 0; dw < SF_KD; ++dw) {
      const uint4 xv = *reinterpret_cast<const uint4*>(&s_xw[hi][dw][half * 4]);
      const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wv + ((dh * SF_KD + dw) * 2 + c) * 64));
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          const float xs = __uint_as_float(c == 0 ? (xw[bb] << 16) : (xw[bb] & 0xffff0000u));
          acc[bb][0] = fmaf(xs, w.x, acc[bb][0]);
          acc[bb][1] = fmaf(xs, w.y, acc[bb][1]);
          acc[bb][2] = fmaf(xs, w.z, acc[bb][2]);
          acc[bb][3] = fmaf(xs, w.w, acc[bb][3]);
        }
      }
    }
  }
#pragma unroll
  for (int bb = 0; bb < 4; ++bb) {
    const int bl = half * 4 + bb;
    if (bl >= nb) continue;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(fmaxf(acc[bb][0], 0.0f), fmaxf(acc[bb][1], 0.0f));
    __nv_bfloat162 h1 = __floats2bfloat162_rn(fmaxf(acc[bb][2], 0.0f), fmaxf(acc[bb][3], 0.0f));
    uint2 o = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    *reinterpret_cast<uint2*>(out + (((int64_t)(b0 + bl) * p.Wp + wo) * p.Hp + ho) * 64 + q4) = o;
  }
}

static size_t stem_fused_smem_bytes() { return SF_B_BYTES + SF_MAXROWS * SF_ROWB + 256 + 64 * 4 + (2 * SF_MAXGRP + 4 * SF_NACC + 1) * 8 + 16 + SF_MAXGRP * 16; }

int init_conv_stem_fused_attrs() {
  static_assert(SF_ROWB % 16 == 0, "bulk copies move multiples of 16 bytes");
  cudaError_t e = cudaFuncSetAttribute(conv_stem_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stem_fused_smem_bytes());
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_stem_fused_kernel) failed: %s", cudaGetErrorString(e));
    return YAD_ERR_CUDA;
  }
  return YAD_OK;
}

}  // namespace yad

static int conv_stem_fused_impl(const void* x_bf16_padded, int64_t x_pitch, int64_t B, int32_t H, int32_t W, const void* w_classes,
                                const float* bias, void* out_flat_bf16, int32_t Hp, int32_t Wp, int32_t n_cta_interior, int32_t skip_lo,
                                int32_t skip_hi, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x_bf16_padded && w_classes && bias && out_flat_bf16, "yad_conv_stem_fused: null pointer");
  YAD_CHECK_ARG(reinterpret_cast<uintptr_t>(x_bf16_padded) % 16 == 0 && x_pitch % 4 == 0, "yad_conv_stem_fused: input rows must be 16-byte aligned");
  YAD_CHECK_ARG(H == 32 && W >= 8 && B >= 0 && B < (1 << 22), "yad_conv_stem_fused: built for H = 32 (n_mels), W >= 8 (got H=%d W=%d)", H, W);
  YAD_CHECK_ARG((reinterpret_cast<uintptr_t>(w_classes) % 16 == 0) && (reinterpret_cast<uintptr_t>(out_flat_bf16) % 16 == 0),
                "yad_conv_stem_fused: weights / out must be 16-byte aligned");
  if (B == 0) return YAD_OK;
  StemFusedParams p;
  p.B = (int)B;
  p.H = H;
  p.W = W;
  p.Ho = 8;
  p.Wo = (((W - 1) / 2 + 1) - 1) / 2 + 1;
  YAD_CHECK_ARG(Hp >= p.Ho && Wp >= p.Wo, "yad_conv_stem_fused: output pitches (%d,%d) smaller than the image (%d,%d)", Hp, Wp, p.Ho, p.Wo);
  p.n_seg = (p.Wo + SF_SEG - 1) / SF_SEG;
  YAD_CHECK_ARG(x_pitch >= 4 * SF_SEG * (p.n_seg - 1) + SF_PW && x_pitch >= 9 + W,
                "yad_conv_stem_fused: x_pitch %lld too small (need >= %d words: 9-word left margin + W + zero tail)", (long long)x_pitch,
                4 * SF_SEG * (p.n_seg - 1) + SF_PW);
  p.xpitch = x_pitch;
  p.Hp = Hp;
  p.Wp = Wp;
  const int rf[4] = {2, 0, 1, 7}, rc[4] = {5, 1, 1, 1};
  for (int i = 0; i < 4; ++i) p.row_first[i] = rf[i], p.row_cnt[i] = rc[i];
  const int nsm = sm_count() > 0 ? sm_count() : 148;
  const int64_t n_tiles = B * p.n_seg;
  int n_int = n_cta_interior > 0 ? n_cta_interior : (nsm * 70) / 100;      // measured balance point of the four classes (tools/stem_sweep.py;
                                                                           // 62 % before the kernel was pipelined over accumulator slots)
  if (n_int > nsm - 3) n_int = nsm - 3;
  if (n_int < 1) n_int = 1;
  int rest = nsm - n_int;
  int n1 = rest / 3, n2 = rest / 3, n3 = rest - 2 * (rest / 3);
  auto cap = [&](int v) { return (int)(v > n_tiles ? n_tiles : (v < 1 ? 1 : v)); };
  n_int = cap(n_int), n1 = cap(n1), n2 = cap(n2), n3 = cap(n3);
  p.cta_first[0] = 0;
  p.cta_first[1] = n_int;
  p.cta_first[2] = n_int + n1;
  p.cta_first[3] = n_int + n1 + n2;
  p.cta_first[4] = n_int + n1 + n2 + n3;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  YAD_CHECK_ARG(skip_lo >= 0 && skip_hi >= 0 && skip_lo + skip_hi <= p.Wo, "yad_conv_stem_fused: bad skipped column counts");
  static const int stem_dbg = [] { const char* e = getenv("YAD_STEM_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = stem_dbg;
  p.skip_lo = skip_lo;
  p.skip_hi = skip_hi;
  YAD_CUDA(launch_pdl(conv_stem_fused_kernel, dim3((unsigned)p.cta_first[4]), dim3(SF_THREADS), stem_fused_smem_bytes(), (cudaStream_t)stream,
                      reinterpret_cast<const uint32_t*>(x_bf16_padded), p, reinterpret_cast<const uint4*>(w_classes), bias,
                      reinterpret_cast<__nv_bfloat16*>(out_flat_bf16)));
  return YAD_OK;
}

extern "C" int yad_conv_stem_fused(const void* x_bf16_padded, int64_t x_pitch, int64_t B, int32_t H, int32_t W, const void* w_classes,
                                   const float* bias, void* out_flat_bf16, int32_t Hp, int32_t Wp, int32_t n_cta_interior,
                                   yad_stream_t stream) {
  return conv_stem_fused_impl(x_bf16_padded, x_pitch, B, H, W, w_classes, bias, out_flat_bf16, Hp, Wp, n_cta_interior, 0, 0, stream);
}

extern "C" int yad_conv_stem_fused_skip(const void* x_bf16_padded, int64_t x_pitch, int64_t B, int32_t H, int32_t W,
                                        const void* w_classes, const float* bias, void* out_flat_bf16, int32_t Hp, int32_t Wp,
                                        int32_t n_cta_interior, int32_t skip_lo, int32_t skip_hi, yad_stream_t stream) {
  return conv_stem_fused_impl(x_bf16_padded, x_pitch, B, H, W, w_classes, bias, out_flat_bf16, Hp, Wp, n_cta_interior, skip_lo, skip_hi,
                              stream);
}

extern "C" int yad_conv_stem_fused_fixup_bf16(const void* x_bf16_padded, int64_t x_pitch, int64_t B, int32_t H, int32_t W,
                                              const float* w_var, const float* bias, const int32_t* cols, const int32_t* col_var,
                                              int32_t n_cols, void* out_flat_bf16, int32_t Hp, int32_t Wp, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x_bf16_padded && w_var && bias && out_flat_bf16 && (n_cols == 0 || (cols && col_var)), "yad_conv_stem_fused_fixup_bf16: null pointer");
  YAD_CHECK_ARG(H == 32 && W >= 8 && B >= 0 && B <= 65535 * (int64_t)FX_CLIPS && n_cols >= 0 && n_cols <= 8,
                "yad_conv_stem_fused_fixup_bf16: bad arguments");
  if (B == 0 || n_cols == 0) return YAD_OK;
  const int Wo = (((W - 1) / 2 + 1) - 1) / 2 + 1;
  StemFixupParams p;
  p.B = (int)B;
  p.H = H;
  p.W = W;
  p.Ho = 8;
  p.Hp = Hp;
  p.Wp = Wp;
  p.n_cols = n_cols;
  for (int i = 0; i < 8; ++i) {
    p.col[i] = i < n_cols ? cols[i] : 0, p.var[i] = i < n_cols ? col_var[i] : 0;
    YAD_CHECK_ARG(i >= n_cols || (cols[i] >= 0 && cols[i] < Wo && 4 * (int64_t)cols[i] + SF_KD <= x_pitch),
                  "yad_conv_stem_fused_fixup_bf16: column %d outside the image / the padded row", i < n_cols ? cols[i] : 0);
  }
  const int rcls[8] = {1, 2, 0, 0, 0, 0, 0, 3};
  for (int i = 0; i < 8; ++i) p.row_class[i] = rcls[i];
  YAD_CUDA(launch_pdl(stem_fixup_bf16_kernel, dim3((unsigned)n_cols, (unsigned)((B + FX_CLIPS - 1) / FX_CLIPS)), dim3(FX_THREADS), 0,
                      (cudaStream_t)stream, reinterpret_cast<const uint32_t*>(x_bf16_padded), x_pitch, p, w_var, bias,
                      reinterpret_cast<__nv_bfloat16*>(out_flat_bf16)));
  return YAD_OK;
}

extern "C" int yad_conv_stem_fused_fixup(const float* x_nchw, int64_t B, int32_t H, int32_t W, const float* w_var, const float* bias,
                                         const int32_t* cols, const int32_t* col_var, int32_t n_cols, void* out_flat_bf16,
                                         int32_t Hp, int32_t Wp, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x_nchw && w_var && bias && out_flat_bf16 && (n_cols == 0 || (cols && col_var)), "yad_conv_stem_fused_fixup: null pointer");
  YAD_CHECK_ARG(H == 32 && W >= 8 && B >= 0 && B <= 65535 && n_cols >= 0 && n_cols <= 8, "yad_conv_stem_fused_fixup: bad arguments");
  if (B == 0 || n_cols == 0) return YAD_OK;
  StemFixupParams p;
  p.B = (int)B;
  p.H = H;
  p.W = W;
  p.Ho = 8;
  p.Hp = Hp;
  p.Wp = Wp;
  p.n_cols = n_cols;
  for (int i = 0; i < 8; ++i) p.col[i] = i < n_cols ? cols[i] : 0, p.var[i] = i < n_cols ? col_var[i] : 0;
  const int rcls[8] = {1, 2, 0, 0, 0, 0, 0, 3};
  for (int i = 0; i < 8; ++i) p.row_class[i] = rcls[i];
  YAD_CUDA(launch_pdl(stem_fixup_kernel, dim3((unsigned)n_cols, (unsigned)((B + FX_CLIPS - 1) / FX_CLIPS)), dim3(FX_THREADS), 0,
                      (cudaStream_t)stream, x_nchw, p, w_var, bias, reinterpret_cast<__nv_bfloat16*>(out_flat_bf16)));
  return YAD_OK;
}
