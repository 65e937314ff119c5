// Fused log-mel / MFCC frontend for sm_100a.
//
// Stage A (frontend_mel_kernel): raw PCM -> mel power.  One 512-thread CTA per run of 8-frame groups of one clip,
// warp-specialised into two roles that overlap through double-buffered shared memory:
//   resample warps (8): 16-byte cp.async staging of the PCM span (one group ahead) -> sparse polyphase resample
//     (only the non-zero taps of torchaudio's Hann-windowed sinc bank; one QUAD of adjacent phases per thread with
//     its 4 x 21 taps resident in registers for the whole kernel) fused with the Hann analysis window -> frame buffer;
//   FFT warps (8): 1000-point real FFT as a 500-point complex FFT = 25 x 20 with both factors done in registers
//     (fft500.cuh; two shared-memory exchanges instead of four radix passes) -> untangle + |X|^2 -> sparse (CSR)
//     mel filterbank -> [B, 32, T] mel power.
//   The 16 kHz signal, the frames and the spectrum never touch HBM; PCM is read exactly once.
// Stage B (frontend_finish_kernel): per clip global reductions that the reference chains
//   (modules/_architecture.py:98-105): dB with a per-clip top_db floor, DCT-II (MFCC), a second dB with
//   its own per-clip floor, per-(clip, channel) mean / unbiased-std standardisation.
//   Streams the 123 KB mel plane of its clip (L2 resident) instead of staging it.
//
// Reference call sites: torchaudio Resample ([ta] functional.py:1405-1431), torch.stft + abs().pow(2)
// ([ta] functional.py:123,144), MelScale matmul ([ta] transforms/_transforms.py:417), amplitude_to_DB
// ([ta] functional.py:390-403), MFCC ([ta] transforms/_transforms.py:709-718), scale_input
// (modules/_architecture.py:182-189).
#include "common.cuh"
#include "fft500.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <string.h>

namespace yad {

constexpr int FE_NFFT = 1000;
constexpr int FE_FR = 8;            // frames per group
constexpr int FE_NMEL = 32;
constexpr int FE_QW = YAD_FE_QW;    // taps per phase over the quad's common window (zero padded)
constexpr int FE_ROLE = 256;        // threads per role
constexpr int FE_THREADS = 2 * FE_ROLE;
constexpr int FE_FR_WORDS = FE_FR * 2 * FFT_X_STRIDE;   // frame buffer (floats): frames in (stride FFT_Z_STRIDE), spectrum out (FFT_X_STRIDE)
constexpr int FE_Y_WORDS = FE_FR * 2 * FFT_Y_STRIDE;    // pass-A -> pass-B exchange buffer
constexpr int FE_P_STRIDE = 516;    // power-spectrum row pitch (= 4 mod 32: 8 frames x 4 adjacent bins hit 32 distinct banks)

// named barriers (0 is __syncthreads)
constexpr int BAR_RS = 1, BAR_FT = 2, BAR_FULL0 = 3, BAR_EMPTY0 = 5;

struct FeParams {
  int64_t B, L, T;
  int32_t P, O, width;
  int32_t HG;        // hops per group = FE_FR * FE_NFFT / P
  int32_t SX;        // staged PCM floats per group
  int32_t n_groups;  // ceil(T / FE_FR), per clip
  int64_t total_groups;   // B * n_groups
  int64_t taper_len;      // elements of the taper window (0 = none)
  int32_t groups_per_cta;
  int32_t fb_nnz_pad;  // mel CSR values, rounded up to a multiple of 4
  int32_t nquad, nslice;
};

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 1-D bulk copy global -> shared through the TMA unit (cp.async.bulk): one instruction for the whole span, completion
// (byte count) signalled on an mbarrier.  Addresses and size are multiples of 16 bytes.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Stage the zero-padded PCM span of one group: element i of the span is xpad[x0 + i] = x[x0 + i] (zero outside [0, L)).
// Tp = float (EPC = 4 elements per 16-byte chunk) or int16_t (EPC = 8; what audio files hold: torchaudio.load(normalize=True)
// turns it into x / 32768).  Fast path (clip base 16 B aligned, L % EPC == 0): the destination is shifted by (x0 mod EPC)
// elements so that global and shared addresses are congruent mod 16 (returns that shift);
//   * span entirely inside the clip: ONE bulk copy issued by thread 0 of the role (bulk = true; wait on the mbarrier);
//   * span touching either end of the clip: 16-byte cp.async per thread with zero fill.
// Called by the FE_ROLE resample threads only (rt = thread index inside the role).
template <typename Tp, int EPC>
__device__ __forceinline__ int fe_stage_async(Tp* s_dst, const Tp* __restrict__ xb, int64_t x0, int SX, int64_t L, bool fast,
                                              int rt, uint64_t* bar, bool& bulk) {
  bulk = false;
  if (fast) {
    const int shift = (int)(((x0 % EPC) + EPC) % EPC);
    const int nchunk = (shift + SX + EPC - 1) / EPC;
    const int64_t g0 = x0 - shift;                      // multiple of EPC
    if (g0 >= 0 && g0 + (int64_t)EPC * nchunk <= L) {
      bulk = true;
      if (rt == 0) {
        mbar_expect_tx(bar, (uint32_t)nchunk * 16u);
        bulk_g2s(s_dst, xb + g0, (uint32_t)nchunk * 16u, bar);
      }
      return shift;
    }
    for (int c = rt; c < nchunk; c += FE_ROLE) {
      const int64_t gi = g0 + EPC * (int64_t)c;
      const bool ok = gi >= 0 && gi + EPC <= L;
      cp_async16(s_dst + EPC * c, xb + (ok ? gi : 0), ok ? 16 : 0);
    }
    cp_async_commit();
    return shift;
  }
  for (int i = rt; i < SX; i += FE_ROLE) {
    const int64_t src = x0 + i;
    s_dst[i] = (src >= 0 && src < L) ? xb[src] : (Tp)0;
  }
  return 0;
}

template <bool I16, bool TAPER>
__global__ void __launch_bounds__(FE_THREADS, 1)
frontend_mel_kernel(const void* __restrict__ pcm, const float* __restrict__ taper, const FeParams p, const float* __restrict__ taps,
                    const int32_t* __restrict__ tap_base, const int32_t* __restrict__ lane_map,
                    const float* __restrict__ window,
                    const float* __restrict__ twiddle, const float* __restrict__ fb_val,
                    const int32_t* __restrict__ fb_bin, const int32_t* __restrict__ fb_start,
                    float* __restrict__ mel) {
  extern __shared__ __align__(16) float fe_smem[];
  const int sxp = (p.SX + 16 + 7) & ~7;
  float* s_x = fe_smem;                                      // [2][sxp] staged PCM spans (16-bit input: two raw
                                                             // int16 spans in the first half, the converted span in the second)
  float* s_fr = s_x + 2 * sxp;                               // [2][FE_FR_WORDS] frames / FFT exchange
  float* s_Y = s_fr + 2 * FE_FR_WORDS;                       // [FE_FR][25][21] complex: pass-A output in pass-B order
  float* s_P = s_Y + FE_Y_WORDS;                             // [FE_FR][FE_P_STRIDE] power spectrum (bins 0..500)
  float2* s_tw = reinterpret_cast<float2*>(s_P + ((FE_FR * FE_P_STRIDE + 3) & ~3));   // exp(-2 pi i k / 1000), 16 B aligned
  float2* s_twA = s_tw + FE_NFFT;                            // [25 regs][20 n2] pass-A twiddles W_500^(n2 k1(r))
  float* s_win = reinterpret_cast<float*>(s_twA + 500);      // [1000] analysis window
  float* s_fbv = s_win + FE_NFFT;                            // mel filterbank values, CSR over bands
  int* s_fbs = reinterpret_cast<int*>(s_fbv + p.fb_nnz_pad); // [33] row starts, then [32] first bin of each band

  const int tid = threadIdx.x;
  const int role = tid >> 8;          // 0: resample, 1: FFT / mel
  pdl_wait();                         // programmatic dependent launch (the PCM may come from a kernel, e.g. the batch builder)
  pdl_trigger();
  const int rt = tid & (FE_ROLE - 1);
  // Persistent CTAs: the B * n_groups frame groups of the batch are one flat sequence (clip-major) that is cut into gridDim.x
  // contiguous runs; group gl belongs to clip gl / n_groups.  The tables above are loaded once per CTA and the double-buffered
  // staging pipeline runs straight across clip boundaries.
  const float* xf = reinterpret_cast<const float*>(pcm);
  const int16_t* x16 = reinterpret_cast<const int16_t*>(pcm);
  int16_t* s_raw = reinterpret_cast<int16_t*>(s_x);          // [2][sxp] int16 (= sxp floats in total)
  float* s_xf = s_x + sxp;
  const bool fast = I16 ? (((p.L & 7) == 0) && ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0))
                        : (((p.L & 3) == 0) && ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0));
  const int64_t g_first = (int64_t)blockIdx.x * p.groups_per_cta;
  if (g_first >= p.total_groups) return;
  const int n_my = (int)min((int64_t)p.groups_per_cta, p.total_groups - g_first);
  // (clip, group inside the clip) of this CTA's first group: the only division of the kernel; every thread then walks its own
  // counters (the int64 divisions per group were 6 % of all executed instructions)
  const int64_t clip_first = g_first / p.n_groups;
  const int g_in_first = (int)(g_first - clip_first * p.n_groups);
  // start sample (in the zero-padded clip) of group g of a clip
  auto group_x0 = [&](int g) { return (int64_t)g * (p.HG * p.O) - p.width; };
  __shared__ __align__(8) uint64_t s_bar[2];    // completion of the bulk copy into staging buffer 0 / 1
  int shift = 0;
  bool bulk = false;
  if (role == 0) {
    if (rt == 0) {                              // the only thread that ever arrives on them; the others wait after __syncthreads
      mbar_init(&s_bar[0], 1);
      mbar_init(&s_bar[1], 1);
      fence_barrier_init();
    }
    const int64_t x0 = group_x0(g_in_first);
    if (I16)
      shift = fe_stage_async<int16_t, 8>(s_raw, x16 + clip_first * p.L, x0, p.SX, p.L, fast, rt, &s_bar[0], bulk);
    else
      shift = fe_stage_async<float, 4>(s_x, xf + clip_first * p.L, x0, p.SX, p.L, fast, rt, &s_bar[0], bulk);
  }

  for (int i = tid; i < FE_NFFT; i += FE_THREADS) {
    s_tw[i] = make_float2(twiddle[2 * i + 1], -twiddle[2 * i]);   // -i W_1000^k: the untangle pass multiplies by it directly
    s_win[i] = window[i];
  }
  for (int i = tid; i < 500; i += FE_THREADS) {
    const int r = i / 20, n2 = i - r * 20;
    const int e = (2 * n2 * passA_k1_of_reg(r)) % 1000;
    s_twA[i] = make_float2(twiddle[2 * e], twiddle[2 * e + 1]);
  }
  // Mel filterbank, re-packed for 8-byte shared-memory loads: band m covers the contiguous bins [bin0, bin0 + len); its
  // padded row starts at the even bin (bin0 & ~1) (a leading zero weight when bin0 is odd), at an even offset, and is
  // zero-padded to a multiple of 4 weights.  The weights carry the 1/4 of the untangle pass (power spectrum kept as 4 |X|^2).
  for (int i = tid; i < p.fb_nnz_pad; i += FE_THREADS) s_fbv[i] = 0.0f;
  for (int i = tid; i < FE_FR * (FE_P_STRIDE - 501); i += FE_THREADS)      // bins past 500: read (times zero) by the padded rows
    s_P[(i / (FE_P_STRIDE - 501)) * FE_P_STRIDE + 501 + i % (FE_P_STRIDE - 501)] = 0.0f;
  if (tid < FE_NMEL) {        // warp 0: one band per lane, exclusive prefix sum of the padded lengths
    const int fb_s0 = fb_start[tid], fb_len = fb_start[tid + 1] - fb_s0;
    const int bin0 = fb_len > 0 ? fb_bin[fb_s0] : 0;
    const int fb_lead = bin0 & 1;
    const int tot = (fb_lead + fb_len + 3) & ~3;
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += v;
    }
    s_fbs[tid] = incl - tot;
    if (tid == FE_NMEL - 1) s_fbs[FE_NMEL] = incl;
    s_fbs[FE_NMEL + 1 + tid] = bin0 - fb_lead;
  }
  __syncthreads();
  for (int m = tid >> 5; m < FE_NMEL; m += FE_THREADS / 32) {   // one warp per band
    const int s0 = fb_start[m], len = fb_start[m + 1] - s0;
    const int lead = len > 0 ? (fb_bin[s0] & 1) : 0;
    for (int i = tid & 31; i < len; i += 32) s_fbv[s_fbs[m] + lead + i] = 0.25f * fb_val[s0 + i];
  }
  __syncthreads();

  if (role == 0) {
    // ===================================================================== resample role
    // One quad of adjacent phases per thread; consecutive lanes take every second quad so that their x windows are
    // ~11 words apart (an odd stride: conflict-free shared-memory reads).
    // With a host-built lane map (frontend_consts.pack_resample_taps) the (quad, hop slice) items are spread over the warps so
    // that the 32 lanes of every load hit 32 distinct banks; without one, consecutive lanes take every second quad (x windows
    // ~11 words apart) which leaves ~1.9 wavefronts per load.
    int quad, sl;
    bool active;
    if (lane_map != nullptr) {
      const int mp = __ldg(lane_map + rt);
      active = mp >= 0;
      quad = mp & 0xffff;
      sl = mp >> 16;
    } else {
      const int half = (p.nquad + 1) >> 1;
      const int qi = rt % p.nquad;
      sl = rt / p.nquad;
      quad = qi < half ? 2 * qi : 2 * (qi - half) + 1;
      active = sl < p.nslice;
    }
    // taps of the phase pairs (4u, 4u+1) and (4u+2, 4u+3) as packed register pairs: one FFMA2 (fma.rn.f32x2 with the staged
    // sample broadcast to both lanes) advances two phases, so a quad costs 21 LDS + 42 FFMA2 instead of 21 LDS + 84 FFMA.
    // Each lane is an ordinary fmaf chain: results are bit-identical to the scalar form.
    float2 tq01[FE_QW], tq23[FE_QW];
    int base = 0;
    if (active) {
      base = tap_base[quad];
#pragma unroll
      for (int j = 0; j < FE_QW; ++j) {
        tq01[j] = make_float2(__ldg(taps + (quad * 4 + 0) * FE_QW + j), __ldg(taps + (quad * 4 + 1) * FE_QW + j));
        tq23[j] = make_float2(__ldg(taps + (quad * 4 + 2) * FE_QW + j), __ldg(taps + (quad * 4 + 3) * FE_QW + j));
      }
    }
    uint32_t bar_phase = 0u;     // bit b: parity of the next completion on s_bar[b]
    int64_t clip_n = clip_first;   // (clip, group in clip) of the group being STAGED (one ahead of the one being resampled)
    int g_n = g_in_first, gcl = g_in_first;
    // output position of this thread's first quad-hop inside the group: frame f0, sample pos0; it advances by nslice * P
    const int o0 = sl * p.P + 4 * quad, o_step = p.nslice * p.P;
    const int f0 = o0 / FE_NFFT, pos0 = o0 - f0 * FE_NFFT;
    for (int gi = 0; gi < n_my; ++gi) {
      const int buf = gi & 1;
      cp_async_wait_all();
      if (bulk) {
        mbar_wait(&s_bar[buf], (bar_phase >> buf) & 1u);
        bar_phase ^= 1u << buf;
      }
      bar_sync(BAR_RS, FE_ROLE);   // span of this group landed for every thread; nobody still reads the other buffer
      if (I16) {                   // int16 -> float (x / 32768, exact), once per sample; the previous group's reads of s_xf are done
        const int16_t* raw = s_raw + buf * sxp;
        for (int i = rt; i < shift + p.SX; i += FE_ROLE) s_xf[i] = (float)raw[i] * (1.0f / 32768.0f);
        bar_sync(BAR_RS, FE_ROLE);
      }
      int shift_next = 0;
      bool bulk_next = false;
      gcl = g_n;                                             // group inside its clip of the group resampled now
      if (++g_n == p.n_groups) { g_n = 0; ++clip_n; }
      if (gi + 1 < n_my) {
        const int64_t x0n = group_x0(g_n);
        if (I16)
          shift_next = fe_stage_async<int16_t, 8>(s_raw + (buf ^ 1) * sxp, x16 + clip_n * p.L, x0n, p.SX, p.L, fast, rt, &s_bar[buf ^ 1],
                                                  bulk_next);
        else
          shift_next = fe_stage_async<float, 4>(s_x + (buf ^ 1) * sxp, xf + clip_n * p.L, x0n, p.SX, p.L, fast, rt, &s_bar[buf ^ 1],
                                                bulk_next);
      }
      if (gi >= 2) bar_sync(BAR_EMPTY0 + buf, FE_THREADS);   // FFT role released this frame buffer
      if (active) {
        const float* sx = (I16 ? s_xf : s_x + buf * sxp) + shift + base;
        float* fr = s_fr + buf * FE_FR_WORDS;
        // taper_input (modules/_architecture.py:87-94): the resampled signal times a clip-long window, before framing
        const float* tpg = TAPER ? taper + (int64_t)gcl * (FE_FR * FE_NFFT) : nullptr;
        int f = f0, pos = pos0;
        // two hops per iteration: four independent FFMA2 chains per thread instead of two (the role has two warps per SM
        // sub-partition; with two chains each the FMA pipe idled on the 4-cycle dependency)
        auto finish = [&](float2 a01, float2 a23, int h) {
          const float4 w = *reinterpret_cast<const float4*>(s_win + pos);
          if (TAPER) {       // (x * taper) * hann, in the reference's order; samples past the window (unused tail frames) get 0
            const int o = h * p.P + 4 * quad;
            const int64_t n = (int64_t)gcl * (FE_FR * FE_NFFT) + o;
            const float4 tp = n + 4 <= p.taper_len ? __ldg(reinterpret_cast<const float4*>(tpg + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
            a01 = __fmul2_rn(a01, make_float2(tp.x, tp.y));
            a23 = __fmul2_rn(a23, make_float2(tp.z, tp.w));
          }
          a01 = __fmul2_rn(a01, make_float2(w.x, w.y));
          a23 = __fmul2_rn(a23, make_float2(w.z, w.w));
          *reinterpret_cast<float4*>(fr + f * (2 * FFT_Z_STRIDE) + pos) = make_float4(a01.x, a01.y, a23.x, a23.y);
          pos += o_step;
          while (pos >= FE_NFFT) { pos -= FE_NFFT; ++f; }
        };
        int h = sl;
        for (; h + p.nslice < p.HG; h += 2 * p.nslice) {
          const float* xs0 = sx + h * p.O;
          const float* xs1 = xs0 + p.nslice * p.O;
          float2 a01 = make_float2(0.0f, 0.0f), a23 = a01, b01 = a01, b23 = a01;
#pragma unroll
          for (int j = 0; j < FE_QW; ++j) {
            const float xv = xs0[j], yv = xs1[j];
            const float2 xx = make_float2(xv, xv), yy = make_float2(yv, yv);
            a01 = __ffma2_rn(tq01[j], xx, a01);
            a23 = __ffma2_rn(tq23[j], xx, a23);
            b01 = __ffma2_rn(tq01[j], yy, b01);
            b23 = __ffma2_rn(tq23[j], yy, b23);
          }
          finish(a01, a23, h);
          finish(b01, b23, h + p.nslice);
        }
        if (h < p.HG) {      // odd hop count of this slice
          const float* xs = sx + h * p.O;
          float2 a01 = make_float2(0.0f, 0.0f), a23 = make_float2(0.0f, 0.0f);
#pragma unroll
          for (int j = 0; j < FE_QW; ++j) {
            const float xv = xs[j];
            const float2 xx = make_float2(xv, xv);
            a01 = __ffma2_rn(tq01[j], xx, a01);
            a23 = __ffma2_rn(tq23[j], xx, a23);
          }
          finish(a01, a23, h);
        }
      }
      __threadfence_block();
      bar_arrive(BAR_FULL0 + buf, FE_THREADS);
      shift = shift_next;
      bulk = bulk_next;
    }
  } else {
    // ===================================================================== FFT / mel role
    // pass-A work items (frame fA, column n2): half-warp h of warps 0..3 takes frame h, n2 = 0..15 (16 consecutive slots for
    // every load / store; both halves read the same twiddles: broadcast); warp 4 takes n2 = 16..19 of all 8 frames.
    // (The plain item = 20 f + n2 order cost 3.2 shared-memory wavefronts per 64-bit access instead of 2.)
    // Warp placement: a role's warp w sits on SM sub-partition w % 4.  Pass A (160 items = 5 warps) runs on warps 3..7 and pass B
    // (200 items = 6.25 warps) on warps 0..6, so that every sub-partition carries 3 of the 12 pass-warps (with both passes on
    // warps 0.. the first sub-partition had 4 and the last 2); the mel pass pairs a short with a long band group per
    // sub-partition (see below).
    constexpr int A_FIRST = FE_ROLE - FE_FR * 20;
    const bool actA = rt >= A_FIRST, actB = rt < FE_FR * 25;
    const int ia = rt - A_FIRST;
    const int fA = ia < 128 ? (ia >> 4) : ((ia - 128) >> 2);
    const int n2 = ia < 128 ? (ia & 15) : 16 + (ia & 3);
    const int fB = rt / 25, k1 = rt - fB * 25;
    int64_t b = clip_first;
    int g = g_in_first;
    for (int gi = 0; gi < n_my; ++gi) {
      const int buf = gi & 1;
      cf32* zf = reinterpret_cast<cf32*>(s_fr + buf * FE_FR_WORDS);
      cf32* yf = reinterpret_cast<cf32*>(s_Y);
      bar_sync(BAR_FULL0 + buf, FE_THREADS);
      // ---- pass A: 25-point DFTs over n1 (stride 20), twiddle W_500^(n2 k1)
      cf32 v[25];
      if (actA) {
        const cf32* src = zf + fA * FFT_Z_STRIDE + n2;
#pragma unroll
        for (int n1 = 0; n1 < 25; ++n1) v[n1] = src[20 * n1];
        dft25(v);
        const cf32* twa = reinterpret_cast<const cf32*>(s_twA) + n2;
#pragma unroll
        for (int r = 1; r < 25; ++r) {
          const cf32 t = twa[r * 20];
          v[r] = cmulc(v[r], t.x, t.y);
        }
      }
      if (actA) {      // Y lives in its own buffer: no hazard with the other threads' frame loads
        cf32* dst = yf + fA * FFT_Y_STRIDE + n2;
#pragma unroll
        for (int r = 0; r < 25; ++r) dst[passA_k1_of_reg(r) * FFT_Y_PITCH] = v[r];
      }
      bar_sync(BAR_FT, FE_ROLE);      // Y complete; every frame load of pass A is done (the spectrum may overwrite the frames)
      // ---- pass B: 20-point DFTs over n2; X[k1 + 25 k2] in natural order
      if (actB) {
        const cf32* src = yf + fB * FFT_Y_STRIDE + k1 * FFT_Y_PITCH;
#pragma unroll
        for (int j = 0; j < 20; ++j) v[j] = src[j];
        dft20(v);
        cf32* dst = zf + fB * FFT_X_STRIDE + k1;
#pragma unroll
        for (int r = 0; r < 20; ++r) dst[25 * passB_k2_of_reg(r)] = v[r];
      }
      bar_sync(BAR_FT, FE_ROLE);
      // ---- real-FFT untangle + power: X[k] = E + W^k O ; X[500-k] = conj(E - W^k O); one thread owns the pair
      //      (warp = frame, lane walks k: no division, consecutive 8-byte accesses; fully unrolled so that the loads of all 8
      //      pairs are in flight together - the pass is a chain of dependent packed operations behind a shared-memory load)
      const int f_u = rt >> 5;
#pragma unroll
      for (int ku = 0; ku < 8; ++ku) {
        const int f = f_u, k_raw = (rt & 31) + 32 * ku;
        const int k = k_raw > 250 ? 250 : k_raw;                    // lanes past the last pair recompute it; only the store is predicated
        const cf32* zb = zf + f * FFT_X_STRIDE;
        const cf32 zk = zb[k];
        const cf32 zq = zb[k == 0 ? 0 : 500 - k];
        // with zn = conj(zq):  2 E = zk + zn,  D = zk - zn,  2 T = D * (-i W^k)  (s_tw holds -i W^k)
        const cf32 E2 = cadd(zk, cmake(zq.x, -zq.y));
        const cf32 D = cadd(zk, cmake(-zq.x, zq.y));
        const float2 w = s_tw[k];
        const cf32 T2 = __ffma2_rn(cmake(-D.y, D.x), cmake(w.y, w.y), __fmul2_rn(D, cmake(w.x, w.x)));   // D * w
        const cf32 Pv = cadd(E2, T2), Qv = csub(E2, T2);            // 2 X[k], 2 conj(X[500-k])
        const cf32 Ps = __fmul2_rn(Pv, Pv), Qs = __fmul2_rn(Qv, Qv);
        float* pp = s_P + f * FE_P_STRIDE;
        if (k_raw <= 250) {
          pp[k] = Ps.x + Ps.y;                                      // 4 |X[k]|^2 (the 1/4 lives in the mel weights: exact)
          pp[500 - k] = Qs.x + Qs.y;
        }
      }
      bar_sync(BAR_FT, FE_ROLE);      // power spectrum complete; the frame buffer is no longer read
      if (gi + 2 < n_my) {
        __threadfence_block();
        bar_arrive(BAR_EMPTY0 + buf, FE_THREADS);
      }
      // ---- sparse mel filterbank (band m covers the contiguous bins [bin0, bin0 + len)); thread = (frame, band), the 4
      //      bands of a warp are adjacent (similar lengths); 4 independent loads in flight per lane.
      //      (Tried and measured slower: a 4-way split of every band across lanes with shuffle reduction - 2.3x the
      //      instructions; deferring the pass to the warps that idle during the next group's pass A - 18 % slower.)
      {
        // band group (4 adjacent bands) of a warp: warps w and w + 4 share a sub-partition and take groups w and 7 - w (the
        // bands widen with frequency: 7 .. 77 bins)
        const int wr = rt >> 5, mg = wr < 4 ? wr : 11 - wr;
        const int f = rt & (FE_FR - 1), m = 4 * mg + ((rt & 31) >> 3);
        const int64_t t = (int64_t)g * FE_FR + f;
        const float2* pp = reinterpret_cast<const float2*>(s_P + f * FE_P_STRIDE + s_fbs[FE_NMEL + 1 + m]);
        const int s0 = s_fbs[m], n4 = (s_fbs[m + 1] - s0) >> 2;
        const float2* fv = reinterpret_cast<const float2*>(s_fbv + s0);
        float2 acc01 = make_float2(0.0f, 0.0f), acc23 = make_float2(0.0f, 0.0f);
#pragma unroll 4
        for (int i = 0; i < n4; ++i) {
          const float2 p01 = pp[2 * i], p23 = pp[2 * i + 1];
          const float2 w01 = fv[2 * i], w23 = fv[2 * i + 1];
          acc01 = __ffma2_rn(p01, w01, acc01);
          acc23 = __ffma2_rn(p23, w23, acc23);
        }
        const float acc0 = acc01.x, acc1 = acc01.y, acc2 = acc23.x, acc3 = acc23.y;
        if (t < p.T) mel[(b * FE_NMEL + m) * p.T + t] = (acc0 + acc1) + (acc2 + acc3);
      }
      if (++g == p.n_groups) { g = 0; ++b; }
    }
  }
}

// ------------------------------------------------------------------------------------ stage B
constexpr int FB_THREADS = 256;

__device__ __forceinline__ float to_db(float x) { return 10.0f * log10f(fmaxf(x, 1e-10f)); }

template <typename Tv, typename Op>
__device__ __forceinline__ Tv block_reduce(Tv v, Tv* scratch, Op op, Tv ident) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  Tv r = ident;
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) r = op(r, scratch[w]);
  return r;
}

__global__ void __launch_bounds__(FB_THREADS, 4)   // 4 CTAs / SM: 512 clips fit one wave
frontend_finish_kernel(const float* __restrict__ mel, int64_t T, const float* __restrict__ dct, float top_db,
                       int standardise, float* __restrict__ xs, float* __restrict__ tap_meldb,
                       float* __restrict__ tap_mfcc, float* __restrict__ tap_mfdb) {
  __shared__ __align__(16) float s_dct[FE_NMEL * FE_NMEL];
  __shared__ float s_redf[32];
  __shared__ double s_redd[32];
  const int64_t b = blockIdx.x;
  const float* mb = mel + b * FE_NMEL * T;
  float* o0 = xs + (b * 2 + 0) * FE_NMEL * T;
  float* o1 = xs + (b * 2 + 1) * FE_NMEL * T;   // holds the raw MFCC plane between passes 2 and 4
  for (int i = threadIdx.x; i < FE_NMEL * FE_NMEL; i += blockDim.x) s_dct[i] = dct[i];
  auto fmax_op = [](float a, float c) { return fmaxf(a, c); };
  auto dadd_op = [](double a, double c) { return a + c; };

  // pass 1: per-clip max of dB(mel) -> first top_db floor.  10 log10(max(., 1e-10)) is monotonic, so the maximum of the
  //         dB plane is the dB of the maximum: no logarithm in this pass (bit-identical result).
  float mx = 0.0f;
  for (int64_t t = threadIdx.x; t < T; t += blockDim.x)
    for (int m = 0; m < FE_NMEL; ++m) mx = fmaxf(mx, mb[m * T + t]);
  const float floor1 = to_db(block_reduce<float>(mx, s_redf, fmax_op, 0.0f)) - top_db;

  // pass 2: clamped dB-mel (parked in o0), MFCC column (DCT-II of the clamped dB-mel column, parked in o1), the MFCC
  //         maximum and the moments of the dB-mel plane.  Every later pass touches only frames this thread wrote.
  float mxf = -INFINITY;
  double s0 = 0.0, q0 = 0.0;
  for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
    float mf[FE_NMEL];
#pragma unroll
    for (int k = 0; k < FE_NMEL; ++k) mf[k] = 0.0f;
    for (int m = 0; m < FE_NMEL; ++m) {
      const float x = fmaxf(to_db(mb[m * T + t]), floor1);
      o0[m * T + t] = x;
      s0 += (double)x;
      q0 += (double)x * (double)x;
      const float4* dr = reinterpret_cast<const float4*>(s_dct + m * FE_NMEL);
#pragma unroll
      for (int k4 = 0; k4 < FE_NMEL / 4; ++k4) {
        const float4 d4 = dr[k4];
        mf[k4 * 4 + 0] = fmaf(x, d4.x, mf[k4 * 4 + 0]);
        mf[k4 * 4 + 1] = fmaf(x, d4.y, mf[k4 * 4 + 1]);
        mf[k4 * 4 + 2] = fmaf(x, d4.z, mf[k4 * 4 + 2]);
        mf[k4 * 4 + 3] = fmaf(x, d4.w, mf[k4 * 4 + 3]);
      }
    }
#pragma unroll
    for (int k = 0; k < FE_NMEL; ++k) {
      o1[k * T + t] = mf[k];
      if (tap_mfcc) tap_mfcc[(b * FE_NMEL + k) * T + t] = mf[k];
      mxf = fmaxf(mxf, mf[k]);
    }
  }
  const float floor2 = to_db(block_reduce<float>(mxf, s_redf, fmax_op, -INFINITY)) - top_db;   // dB of the max = max of the dB
  const double n_el = (double)FE_NMEL * (double)T;
  const double S0 = block_reduce<double>(s0, s_redd, dadd_op, 0.0);
  const double Q0 = block_reduce<double>(q0, s_redd, dadd_op, 0.0);
  const float mu0 = (float)(S0 / n_el);

  // pass 3: clamped dB(MFCC) (replaces the raw MFCC in o1) and its moments (fp64 accumulators: sum and sum of squares)
  double s1 = 0.0, q1 = 0.0;
  for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
    for (int k = 0; k < FE_NMEL; ++k) {
      const float yf = fmaxf(to_db(o1[k * T + t]), floor2);
      o1[k * T + t] = yf;
      const double y = (double)yf;
      s1 += y;
      q1 += y * y;
    }
  }
  float mu1 = 0.0f, sd0 = 1.0f, sd1 = 1.0f;
  if (standardise) {
    const double S1 = block_reduce<double>(s1, s_redd, dadd_op, 0.0);
    const double Q1 = block_reduce<double>(q1, s_redd, dadd_op, 0.0);
    mu1 = (float)(S1 / n_el);
    // unbiased variance = (sum x^2 - n mean^2) / (n - 1), all in fp64
    sd0 = (float)sqrt(fmax(Q0 - S0 * S0 / n_el, 0.0) / (n_el - 1.0));
    sd1 = (float)sqrt(fmax(Q1 - S1 * S1 / n_el, 0.0) / (n_el - 1.0));
  }
  // pass 4: standardise in place (+ optional taps); no transcendental left
  const float den0 = sd0 + 1e-5f, den1 = sd1 + 1e-5f;
  for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
    for (int m = 0; m < FE_NMEL; ++m) {
      const float x = o0[m * T + t];
      if (tap_meldb) tap_meldb[(b * FE_NMEL + m) * T + t] = x;
      if (standardise) o0[m * T + t] = __fdiv_rn(x - mu0, den0);
    }
    for (int k = 0; k < FE_NMEL; ++k) {
      const float y = o1[k * T + t];
      if (tap_mfdb) tap_mfdb[(b * FE_NMEL + k) * T + t] = y;
      if (standardise) o1[k * T + t] = __fdiv_rn(y - mu1, den1);
    }
  }
}

// Stage B, one-pass form for T <= 1024 frames (any clip up to 64 s): thread t owns frame t.  The clamped dB-mel column is parked
// in shared memory ([32][T] fp32, thread-private column: conflict free), the MFCC column stays in registers, so the mel plane is
// read from global memory ONCE and both output planes are written ONCE (369 KB per clip instead of 1.2 MB in four passes).
// Same operation order per element as frontend_finish_kernel (fmaf chains over m, fp64 moment accumulators).
constexpr int FB2_MAXT = 1024;

template <typename Tv, typename Op>
__device__ __forceinline__ Tv block_reduce_n(Tv v, Tv* scratch, Op op, Tv ident) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  Tv r = ident;
  const int nw = (blockDim.x + 31) >> 5;
  for (int w = 0; w < nw; ++w) r = op(r, scratch[w]);
  return r;
}

__global__ void __launch_bounds__(FB2_MAXT, 1)
frontend_finish_v2_kernel(const float* __restrict__ mel, int64_t B, int64_t T, const float* __restrict__ dct, float top_db,
                          int standardise, float* __restrict__ xs, float* __restrict__ tap_meldb, float* __restrict__ tap_mfcc,
                          float* __restrict__ tap_mfdb, uint32_t* __restrict__ xs_bf16, int64_t bf_pitch, int bf_margin) {
  extern __shared__ __align__(16) float fb2_smem[];
  float* s_dct = fb2_smem;                       // [32][32]
  float* s_x = fb2_smem + FE_NMEL * FE_NMEL;     // [32][T]
  __shared__ float s_redf[32];
  __shared__ double s_red4[32 * 4];
  const int t = threadIdx.x;
  const bool act = t < T;
  for (int i = threadIdx.x; i < FE_NMEL * FE_NMEL; i += blockDim.x) s_dct[i] = dct[i];
  pdl_wait();          // programmatic dependent launch: stage A's mel plane is complete and visible from here
  pdl_trigger();
  auto fmax_op = [](float a, float c) { return fmaxf(a, c); };
  auto dadd_op = [](double a, double c) { return a + c; };
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const float* mb = mel + b * FE_NMEL * T;
    float* o0 = xs + (b * 2 + 0) * FE_NMEL * T;
    float* o1 = xs + (b * 2 + 1) * FE_NMEL * T;
    // 1: mel column -> smem, clip maximum (10 log10(max(., 1e-10)) is monotonic: the maximum of the dB plane is the dB of the maximum)
    float mx = 0.0f;
    if (act) {
#pragma unroll 8
      for (int m = 0; m < FE_NMEL; ++m) {
        const float v = __ldg(mb + m * T + t);
        s_x[m * T + t] = v;
        mx = fmaxf(mx, v);
      }
    }
    const float floor1 = to_db(block_reduce_n<float>(mx, s_redf, fmax_op, 0.0f)) - top_db;
    // 2: clamped dB-mel (parked in smem), MFCC column in registers, moments of the dB-mel plane
    float2 mf2[FE_NMEL / 2];       // MFCC column as register pairs: the DCT advances two coefficients per FFMA2
#pragma unroll
    for (int k = 0; k < FE_NMEL / 2; ++k) mf2[k] = make_float2(0.0f, 0.0f);
    double s0 = 0.0, q0 = 0.0;
    float mxf = -INFINITY;
    if (act) {
      for (int m = 0; m < FE_NMEL; ++m) {
        const float x = fmaxf(to_db(s_x[m * T + t]), floor1);
        s_x[m * T + t] = x;
        s0 += (double)x;
        q0 += (double)x * (double)x;
        const float4* dr = reinterpret_cast<const float4*>(s_dct + m * FE_NMEL);
#pragma unroll
        for (int k4 = 0; k4 < FE_NMEL / 4; ++k4) {
          const float4 d4 = dr[k4];
          const float2 xx = make_float2(x, x);
          mf2[k4 * 2 + 0] = __ffma2_rn(xx, make_float2(d4.x, d4.y), mf2[k4 * 2 + 0]);
          mf2[k4 * 2 + 1] = __ffma2_rn(xx, make_float2(d4.z, d4.w), mf2[k4 * 2 + 1]);
        }
      }
    }
    float* mf = reinterpret_cast<float*>(mf2);      // the same registers, scalar view (all indices below are compile-time)
    if (act) {
#pragma unroll
      for (int k = 0; k < FE_NMEL; ++k) {
        if (tap_mfcc) tap_mfcc[(b * FE_NMEL + k) * T + t] = mf[k];
        mxf = fmaxf(mxf, mf[k]);
      }
    }
    const float floor2 = to_db(block_reduce_n<float>(mxf, s_redf, fmax_op, -INFINITY)) - top_db;
    // 3: clamped dB(MFCC) in registers and its moments
    double s1 = 0.0, q1 = 0.0;
    if (act) {
#pragma unroll
      for (int k = 0; k < FE_NMEL; ++k) {
        const float yf = fmaxf(to_db(mf[k]), floor2);
        mf[k] = yf;
        s1 += (double)yf;
        q1 += (double)yf * (double)yf;
      }
    }
    const double n_el = (double)FE_NMEL * (double)T;
    float mu0 = 0.0f, mu1 = 0.0f, sd0 = 1.0f, sd1 = 1.0f;
    if (standardise) {
      // the four moments in ONE block reduction (two barriers instead of eight): per-warp shuffle sums, lane 0 parks the four
      // partial sums, every thread then adds the warps' partials in the same order (bit-identical to four separate reductions)
      double m4[4] = {s0, q0, s1, q1};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m4[i] += __shfl_xor_sync(0xffffffffu, m4[i], o);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) s_red4[(threadIdx.x >> 5) * 4 + i] = m4[i];
      }
      __syncthreads();
      double S0 = 0.0, Q0 = 0.0, S1 = 0.0, Q1 = 0.0;
      const int nw = (blockDim.x + 31) >> 5;
      for (int w = 0; w < nw; ++w) {
        S0 += s_red4[w * 4 + 0];
        Q0 += s_red4[w * 4 + 1];
        S1 += s_red4[w * 4 + 2];
        Q1 += s_red4[w * 4 + 3];
      }
      mu0 = (float)(S0 / n_el);
      mu1 = (float)(S1 / n_el);
      sd0 = (float)sqrt(fmax(Q0 - S0 * S0 / n_el, 0.0) / (n_el - 1.0));
      sd1 = (float)sqrt(fmax(Q1 - S1 * S1 / n_el, 0.0) / (n_el - 1.0));
    }
    // 4: standardise and write both planes once
    const float den0 = sd0 + 1e-5f, den1 = sd1 + 1e-5f;
    if (act) {
#pragma unroll
      for (int m = 0; m < FE_NMEL; ++m) {
        const float x = s_x[m * T + t];
        if (tap_meldb) tap_meldb[(b * FE_NMEL + m) * T + t] = x;
        if (tap_mfdb) tap_mfdb[(b * FE_NMEL + m) * T + t] = mf[m];
        const float v0 = standardise ? __fdiv_rn(x - mu0, den0) : x;
        const float v1 = standardise ? __fdiv_rn(mf[m] - mu1, den1) : mf[m];
        o0[m * T + t] = v0;
        o1[m * T + t] = v1;
        if (xs_bf16 != nullptr) {      // channel-interleaved bf16 copy with zero margins: the fused stem's patch rows (bulk-copied)
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
          xs_bf16[(b * FE_NMEL + m) * bf_pitch + bf_margin + t] = *reinterpret_cast<uint32_t*>(&h2);
        }
      }
    }
    __syncthreads();     // s_x is reused by the next clip
  }
}

// Stage B, cluster form (round 2, last session): the one-pass kernel above keeps ONE 1024-thread CTA per SM (123 KB of parked
// dB-mel columns), so every block-wide reduction and every phase change (HBM-latency-bound load, issue-bound dB / DCT, store) idles
// the whole SM.  Here a clip is split over a CLUSTER of two 512-thread CTAs (frames [0, T/2) and [T/2, T)); the two maxima and the
// four moments are combined through distributed shared memory (each CTA parks its partial, one cluster barrier, both read both
// partials in rank order: identical bits in both CTAs and for every batch position).  Two CTAs of DIFFERENT clips share an SM
// (61 KB + 64 registers x 512 threads each), so one clip's load / reduction phases run under the other's arithmetic.
// Same operation order per element as frontend_finish_v2_kernel; the fp64 moments are summed per half, then half 0 + half 1.
// (ACC = float was a timing experiment: no difference, the fp64 pipe is not what limits this kernel.)
constexpr int FB3_THREADS = 512;

__device__ __forceinline__ double ld_cluster_f64(const double* p, uint32_t rank) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(mapa_u32(p, rank)) : "memory");
  return v;
}

template <typename ACC>
__global__ void __launch_bounds__(FB3_THREADS, 2)
frontend_finish_v3_kernel(const float* __restrict__ mel, int64_t B, int64_t T, const float* __restrict__ dct, float top_db,
                          int standardise, float* __restrict__ xs, float* __restrict__ tap_meldb, float* __restrict__ tap_mfcc,
                          float* __restrict__ tap_mfdb, uint32_t* __restrict__ xs_bf16, int64_t bf_pitch, int bf_margin, int CS) {
  extern __shared__ __align__(16) float fb3_smem[];
  float* s_dct = fb3_smem;                       // [32][32]
  float* s_x = fb3_smem + FE_NMEL * FE_NMEL;     // [32][TH]
  __shared__ float s_redf[32];
  __shared__ double s_red4[32 * 4];
  __shared__ __align__(8) double s_part[8];      // this CTA's partials for the cluster: 0 max of mel, 1 max of MFCC, 2..5 moments
  const uint32_t rank = cluster_ctarank();
  const int TH = (int)((T + CS - 1) / CS);       // frames per CTA (CS = CTAs per cluster = per clip)
  const int lt = threadIdx.x;
  const int64_t t = (int64_t)rank * TH + lt;
  const bool act = lt < TH && t < T;
  for (int i = threadIdx.x; i < FE_NMEL * FE_NMEL; i += blockDim.x) s_dct[i] = dct[i];
  pdl_wait();          // programmatic dependent launch: stage A's mel plane is complete and visible from here
  pdl_trigger();
  auto fmax_op = [](float a, float c) { return fmaxf(a, c); };
  const int64_t n_clusters = gridDim.x / CS;
  for (int64_t b = blockIdx.x / CS; b < B; b += n_clusters) {
    const float* mb = mel + b * FE_NMEL * T;
    float* o0 = xs + (b * 2 + 0) * FE_NMEL * T;
    float* o1 = xs + (b * 2 + 1) * FE_NMEL * T;
    // 1: mel column -> smem, clip maximum (10 log10(max(., 1e-10)) is monotonic: the maximum of the dB plane is the dB of the maximum)
    float mx = 0.0f;
    if (act) {
#pragma unroll 8
      for (int m = 0; m < FE_NMEL; ++m) {
        const float v = __ldg(mb + m * T + t);
        s_x[m * TH + lt] = v;
        mx = fmaxf(mx, v);
      }
    }
    mx = block_reduce_n<float>(mx, s_redf, fmax_op, 0.0f);
    if (threadIdx.x == 0) s_part[0] = (double)mx;
    cluster_sync_all();
    float gm = 0.0f;
    for (int rk = 0; rk < CS; ++rk) gm = fmaxf(gm, (float)ld_cluster_f64(&s_part[0], rk));
    const float floor1 = to_db(gm) - top_db;
    // 2: clamped dB-mel (parked in smem), MFCC column in registers, moments of the dB-mel plane
    float2 mf2[FE_NMEL / 2];       // MFCC column as register pairs: the DCT advances two coefficients per FFMA2
#pragma unroll
    for (int k = 0; k < FE_NMEL / 2; ++k) mf2[k] = make_float2(0.0f, 0.0f);
    ACC s0 = 0, q0 = 0;
    float mxf = -INFINITY;
    if (act) {
      for (int m = 0; m < FE_NMEL; ++m) {
        const float x = fmaxf(to_db(s_x[m * TH + lt]), floor1);
        s_x[m * TH + lt] = x;
        s0 += (ACC)x;
        q0 += (ACC)x * (ACC)x;
        const float4* dr = reinterpret_cast<const float4*>(s_dct + m * FE_NMEL);
#pragma unroll
        for (int k4 = 0; k4 < FE_NMEL / 4; ++k4) {
          const float4 d4 = dr[k4];
          const float2 xx = make_float2(x, x);
          mf2[k4 * 2 + 0] = __ffma2_rn(xx, make_float2(d4.x, d4.y), mf2[k4 * 2 + 0]);
          mf2[k4 * 2 + 1] = __ffma2_rn(xx, make_float2(d4.z, d4.w), mf2[k4 * 2 + 1]);
        }
      }
    }
    float* mf = reinterpret_cast<float*>(mf2);      // the same registers, scalar view (all indices below are compile-time)
    if (act) {
#pragma unroll
      for (int k = 0; k < FE_NMEL; ++k) {
        if (tap_mfcc) tap_mfcc[(b * FE_NMEL + k) * T + t] = mf[k];
        mxf = fmaxf(mxf, mf[k]);
      }
    }
    mxf = block_reduce_n<float>(mxf, s_redf, fmax_op, -INFINITY);
    if (threadIdx.x == 0) s_part[1] = (double)mxf;
    cluster_sync_all();
    float gf = -INFINITY;
    for (int rk = 0; rk < CS; ++rk) gf = fmaxf(gf, (float)ld_cluster_f64(&s_part[1], rk));
    const float floor2 = to_db(gf) - top_db;
    // 3: clamped dB(MFCC) in registers and its moments
    ACC s1 = 0, q1 = 0;
    if (act) {
#pragma unroll
      for (int k = 0; k < FE_NMEL; ++k) {
        const float yf = fmaxf(to_db(mf[k]), floor2);
        mf[k] = yf;
        s1 += (ACC)yf;
        q1 += (ACC)yf * (ACC)yf;
      }
    }
    const double n_el = (double)FE_NMEL * (double)T;
    float mu0 = 0.0f, mu1 = 0.0f, sd0 = 1.0f, sd1 = 1.0f;
    if (standardise) {
      double m4[4] = {(double)s0, (double)q0, (double)s1, (double)q1};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m4[i] += __shfl_xor_sync(0xffffffffu, m4[i], o);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) s_red4[(threadIdx.x >> 5) * 4 + i] = m4[i];
      }
      __syncthreads();
      if (threadIdx.x < 4) {
        double a = 0.0;
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; ++w) a += s_red4[w * 4 + threadIdx.x];
        s_part[2 + threadIdx.x] = a;
      }
      cluster_sync_all();
      double S0 = 0.0, Q0 = 0.0, S1 = 0.0, Q1 = 0.0;      // rank order: the same bits in every CTA of the cluster
      for (int rk = 0; rk < CS; ++rk) {
        S0 += ld_cluster_f64(&s_part[2], rk);
        Q0 += ld_cluster_f64(&s_part[3], rk);
        S1 += ld_cluster_f64(&s_part[4], rk);
        Q1 += ld_cluster_f64(&s_part[5], rk);
      }
      mu0 = (float)(S0 / n_el);
      mu1 = (float)(S1 / n_el);
      sd0 = (float)sqrt(fmax(Q0 - S0 * S0 / n_el, 0.0) / (n_el - 1.0));
      sd1 = (float)sqrt(fmax(Q1 - S1 * S1 / n_el, 0.0) / (n_el - 1.0));
    }
    // 4: standardise and write both planes once
    const float den0 = sd0 + 1e-5f, den1 = sd1 + 1e-5f;
    if (act) {
#pragma unroll
      for (int m = 0; m < FE_NMEL; ++m) {
        const float x = s_x[m * TH + lt];
        if (tap_meldb) tap_meldb[(b * FE_NMEL + m) * T + t] = x;
        if (tap_mfdb) tap_mfdb[(b * FE_NMEL + m) * T + t] = mf[m];
        const float v0 = standardise ? __fdiv_rn(x - mu0, den0) : x;
        const float v1 = standardise ? __fdiv_rn(mf[m] - mu1, den1) : mf[m];
        o0[m * T + t] = v0;
        o1[m * T + t] = v1;
        if (xs_bf16 != nullptr) {      // channel-interleaved bf16 copy with zero margins: the fused stem's patch rows (bulk-copied)
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
          xs_bf16[(b * FE_NMEL + m) * bf_pitch + bf_margin + t] = *reinterpret_cast<uint32_t*>(&h2);
        }
      }
    }
    __syncthreads();     // s_x is reused by the next clip (s_part: see the ordering argument in DESIGN.md 5.6)
  }
  cluster_sync_all();    // the peer may still be reading this CTA's partials
}

static size_t fe_smem_bytes(int SX, int nnz_pad) {
  const int sxp = (SX + 16 + 7) & ~7;
  return (size_t)(2 * sxp + 2 * FE_FR_WORDS + FE_Y_WORDS + ((FE_FR * FE_P_STRIDE + 3) & ~3) + 2 * FE_NFFT + 1000 + FE_NFFT + nnz_pad +
                  2 * FE_NMEL + 4) * sizeof(float);
}

int init_frontend_attrs() {
  cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(&frontend_mel_kernel<false, false>),
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(reinterpret_cast<const void*>(&frontend_mel_kernel<true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                             226 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(reinterpret_cast<const void*>(&frontend_mel_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                             226 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(reinterpret_cast<const void*>(&frontend_mel_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                             226 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(frontend_finish_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)((FE_NMEL * FE_NMEL + FE_NMEL * FB2_MAXT) * sizeof(float)));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(frontend_finish_v3_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)((FE_NMEL * FE_NMEL + FE_NMEL * (FB2_MAXT / 2)) * sizeof(float)));
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(frontend_mel_kernel) failed: %s", cudaGetErrorString(e));
    return YAD_ERR_CUDA;
  }
  return YAD_OK;
}

}  // namespace yad

extern "C" {

static int frontend_mel_impl(const void* pcm, bool i16, const float* taper, int64_t taper_len, int64_t B, int64_t L, int32_t P, int32_t O,
                             int32_t width,
                             const float* taps, const int32_t* tap_base, const int32_t* lane_map, int32_t window_len,
                             const float* window,
                             const float* twiddle, const float* fb_val, const int32_t* fb_bin,
                             const int32_t* fb_start, int32_t fb_nnz, float* mel, int64_t T, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(pcm && taps && tap_base && window && twiddle && fb_val && fb_bin && fb_start && mel,
                "yad_frontend_mel_power: null pointer");
  YAD_CHECK_ARG(P >= 4 && P % 4 == 0 && P / 4 <= FE_ROLE, "yad_frontend_mel_power: P=%d must be a multiple of 4 and <= %d", P,
                4 * FE_ROLE);
  YAD_CHECK_ARG((FE_FR * FE_NFFT) % P == 0, "yad_frontend_mel_power: %d-sample frame groups must be whole hops of P=%d",
                FE_FR * FE_NFFT, P);
  YAD_CHECK_ARG(O >= 1 && width >= 0 && window_len >= FE_QW, "yad_frontend_mel_power: bad O/width/window_len");
  YAD_CHECK_ARG(B >= 0 && B <= (1 << 24) && L >= 1 && T >= 1, "yad_frontend_mel_power: bad B/L/T");
  YAD_CHECK_ARG(fb_nnz >= 1 && fb_nnz <= 16384, "yad_frontend_mel_power: bad fb_nnz=%d", fb_nnz);
  // frames must exist in the resampled signal: T*1000 <= ceil(P*L/O)
  YAD_CHECK_ARG((int64_t)T * FE_NFFT <= (P * L + O - 1) / O, "yad_frontend_mel_power: T=%lld frames exceed the resampled length",
                (long long)T);
  if (B == 0) return YAD_OK;
  FeParams p;
  p.B = B;
  p.L = L;
  p.T = T;
  p.P = P;
  p.O = O;
  p.width = width;
  p.HG = FE_FR * FE_NFFT / P;
  p.SX = (p.HG - 1) * O + window_len;   // window_len = max(tap_base) + FE_QW
  p.n_groups = (int)((T + FE_FR - 1) / FE_FR);
  p.nquad = P / 4;
  p.nslice = FE_ROLE / p.nquad;
  if (p.nslice > p.HG) p.nslice = p.HG;
  // persistent CTAs (one per SM: the kernel's shared memory allows no more): equal contiguous runs of the flat group sequence
  const int nsm = sm_count() > 0 ? sm_count() : 148;
  p.total_groups = B * (int64_t)p.n_groups;
  const int64_t gpc = (p.total_groups + nsm - 1) / nsm;
  YAD_CHECK_ARG(gpc < (1ll << 30), "yad_frontend_mel_power: batch too large");
  p.groups_per_cta = (int)gpc;
  p.fb_nnz_pad = (fb_nnz + 4 * FE_NMEL + 3) & ~3;   // every band: <= 1 leading + <= 3 trailing zero weights
  const size_t smem = fe_smem_bytes(p.SX, p.fb_nnz_pad);
  YAD_CHECK_ARG(smem <= 226 * 1024, "yad_frontend_mel_power: staging span too large (%zu B of shared memory)", smem);
  dim3 grid((unsigned)((p.total_groups + p.groups_per_cta - 1) / p.groups_per_cta));
  YAD_CHECK_ARG(taper == nullptr || (taper_len >= T * FE_NFFT && reinterpret_cast<uintptr_t>(taper) % 16 == 0),
                "yad_frontend_mel_power: taper window shorter than T*1000 samples or not 16-byte aligned");
  p.taper_len = taper ? taper_len : 0;
#define YAD_FE_LAUNCH(I16_, TAPER_)                                                                                          \
  YAD_CUDA(launch_pdl(frontend_mel_kernel<I16_, TAPER_>, grid, dim3(FE_THREADS), smem, (cudaStream_t)stream, pcm, taper, p, taps,  \
                      tap_base, lane_map, window, twiddle, fb_val, fb_bin, fb_start, mel))
  if (i16) {
    if (taper) YAD_FE_LAUNCH(true, true); else YAD_FE_LAUNCH(true, false);
  } else {
    if (taper) YAD_FE_LAUNCH(false, true); else YAD_FE_LAUNCH(false, false);
  }
#undef YAD_FE_LAUNCH
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_frontend_mel_power(const float* pcm, int64_t B, int64_t L, int32_t P, int32_t O, int32_t width,
                           const float* taps, const int32_t* tap_base, const int32_t* lane_map, int32_t window_len, const float* window,
                           const float* twiddle, const float* fb_val, const int32_t* fb_bin,
                           const int32_t* fb_start, int32_t fb_nnz, float* mel, int64_t T, yad_stream_t stream) {
  return frontend_mel_impl(pcm, false, nullptr, 0, B, L, P, O, width, taps, tap_base, lane_map, window_len, window, twiddle, fb_val, fb_bin, fb_start,
                           fb_nnz, mel, T, stream);
}

int yad_frontend_mel_power_i16(const int16_t* pcm, int64_t B, int64_t L, int32_t P, int32_t O, int32_t width,
                               const float* taps, const int32_t* tap_base, const int32_t* lane_map, int32_t window_len, const float* window,
                               const float* twiddle, const float* fb_val, const int32_t* fb_bin,
                               const int32_t* fb_start, int32_t fb_nnz, float* mel, int64_t T, yad_stream_t stream) {
  return frontend_mel_impl(pcm, true, nullptr, 0, B, L, P, O, width, taps, tap_base, lane_map, window_len, window, twiddle, fb_val, fb_bin, fb_start,
                           fb_nnz, mel, T, stream);
}

int yad_frontend_mel_power_taper(const void* pcm, int32_t pcm_is_i16, const float* taper, int64_t taper_len, int64_t B, int64_t L,
                                 int32_t P, int32_t O, int32_t width, const float* taps, const int32_t* tap_base,
                                 const int32_t* lane_map, int32_t window_len, const float* window, const float* twiddle,
                                 const float* fb_val, const int32_t* fb_bin, const int32_t* fb_start, int32_t fb_nnz, float* mel,
                                 int64_t T, yad_stream_t stream) {
  YAD_CHECK_ARG(taper != nullptr, "yad_frontend_mel_power_taper: null taper window");
  return frontend_mel_impl(pcm, pcm_is_i16 != 0, taper, taper_len, B, L, P, O, width, taps, tap_base, lane_map, window_len, window,
                           twiddle, fb_val, fb_bin, fb_start, fb_nnz, mel, T, stream);
}

static int frontend_finish_impl(const float* mel, int64_t B, int64_t T, const float* dct, float top_db, int32_t standardise,
                                float* x_spectral, float* tap_meldb, float* tap_mfcc, float* tap_mfdb, void* xs_bf16, int64_t bf_pitch,
                                int32_t bf_margin, yad_stream_t stream) {
  YAD_CHECK_ARG(mel && dct && x_spectral && T >= 2 && B >= 0, "yad_frontend_finish: bad arguments");
  YAD_CHECK_ARG(xs_bf16 == nullptr || (T <= yad::FB2_MAXT && bf_margin >= 0 && bf_pitch >= bf_margin + T),
                "yad_frontend_finish_bf16: needs T <= %d and pitch >= margin + T", yad::FB2_MAXT);
  if (B == 0) return YAD_OK;
  if (T <= yad::FB2_MAXT) {     // one-pass form: thread = frame, dB-mel column in shared memory, MFCC column in registers
    const int nsm = yad::sm_count() > 0 ? yad::sm_count() : 148;
    static const bool use_cluster = [] {
      const char* e = getenv("YAD_FE_FINISH_CLUSTER");
      return !(e && e[0] == '0');
    }();
    if (use_cluster && T >= 64) {     // a clip per cluster of two CTAs (frontend_finish_v3_kernel)
      static const int cs_env = [] { const char* e = getenv("YAD_FE_FINISH_CS"); return e ? atoi(e) : 2; }();
      const int CS = (cs_env == 4 && T >= 128) ? 4 : 2;
      const int TH = (int)((T + CS - 1) / CS);
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)(CS * (B < nsm ? B : nsm)));
      cfg.blockDim = dim3((unsigned)((TH + 31) / 32 * 32));
      cfg.dynamicSmemBytes = (size_t)(yad::FE_NMEL * yad::FE_NMEL + yad::FE_NMEL * TH) * sizeof(float);
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute attr[2];
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      int mode = yad::pdl_mode();
      if (mode == 1 && cudaStreamIsCapturing(cfg.stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) mode = 0;
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = mode ? 1 : 0;
      attr[1].id = cudaLaunchAttributeClusterDimension;
      attr[1].val.clusterDim.x = (unsigned)CS;
      attr[1].val.clusterDim.y = 1;
      attr[1].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 2;
      YAD_CUDA(cudaLaunchKernelEx(&cfg, yad::frontend_finish_v3_kernel<double>, mel, B, T, dct, top_db, (int)standardise, x_spectral, tap_meldb,
                                  tap_mfcc, tap_mfdb, reinterpret_cast<uint32_t*>(xs_bf16), bf_pitch, (int)bf_margin, CS));
      return YAD_OK;
    }
    const unsigned grid = (unsigned)(B < nsm ? B : nsm);
    const unsigned threads = (unsigned)((T + 31) / 32 * 32);
    const size_t smem = (size_t)(yad::FE_NMEL * yad::FE_NMEL + yad::FE_NMEL * T) * sizeof(float);
    YAD_CUDA(yad::launch_pdl(yad::frontend_finish_v2_kernel, dim3(grid), dim3(threads), smem, (cudaStream_t)stream, mel, B, T, dct, top_db,
                             (int)standardise, x_spectral, tap_meldb, tap_mfcc, tap_mfdb, reinterpret_cast<uint32_t*>(xs_bf16), bf_pitch,
                             (int)bf_margin));
    return YAD_OK;
  }
  yad::frontend_finish_kernel<<<(unsigned)B, yad::FB_THREADS, 0, (cudaStream_t)stream>>>(
      mel, T, dct, top_db, standardise, x_spectral, tap_meldb, tap_mfcc, tap_mfdb);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_frontend_finish(const float* mel, int64_t B, int64_t T, const float* dct, float top_db, int32_t standardise,
                        float* x_spectral, float* tap_meldb, float* tap_mfcc, float* tap_mfdb, yad_stream_t stream) {
  return frontend_finish_impl(mel, B, T, dct, top_db, standardise, x_spectral, tap_meldb, tap_mfcc, tap_mfdb, nullptr, 0, 0, stream);
}

int yad_frontend_finish_bf16(const float* mel, int64_t B, int64_t T, const float* dct, float top_db, int32_t standardise,
                             float* x_spectral, float* tap_meldb, float* tap_mfcc, float* tap_mfdb, void* xs_bf16, int64_t bf_pitch,
                             int32_t bf_margin, yad_stream_t stream) {
  YAD_CHECK_ARG(xs_bf16 != nullptr, "yad_frontend_finish_bf16: null pointer");
  return frontend_finish_impl(mel, B, T, dct, top_db, standardise, x_spectral, tap_meldb, tap_mfcc, tap_mfdb, xs_bf16, bf_pitch, bf_margin,
                              stream);
}

}  // extern "C"
