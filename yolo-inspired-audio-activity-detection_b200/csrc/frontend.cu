// Fused log-mel / MFCC frontend for sm_100a.
//
// Stage A (frontend_mel_kernel): raw PCM -> mel power, one CTA per run of 8-frame groups of one clip:
//   coalesced HBM reads of the PCM span -> shared memory -> sparse polyphase resample (only the
//   non-zero taps of torchaudio's Hann-windowed sinc bank) fused with the Hann analysis window ->
//   1000-point real FFT as a 500-point complex Stockham FFT (radix 5,5,5,4) entirely in shared memory
//   -> |X|^2 -> sparse (CSR) mel filterbank -> [B, 32, T] mel power.
//   The 16 kHz signal, the frames and the spectrum never touch HBM.
// Stage B (frontend_finish_kernel): per clip global reductions that the reference chains
//   (modules/_architecture.py:98-105): dB with a per-clip top_db floor, DCT-II (MFCC), a second dB with
//   its own per-clip floor, per-(clip, channel) mean / unbiased-std standardisation.
//   Streams the 123 KB mel plane of its clip (L2 resident) five times instead of staging it.
//
// Reference call sites: torchaudio Resample ([ta] functional.py:1405-1431), torch.stft + abs().pow(2)
// ([ta] functional.py:123,144), MelScale matmul ([ta] transforms/_transforms.py:417), amplitude_to_DB
// ([ta] functional.py:390-403), MFCC ([ta] transforms/_transforms.py:709-718), scale_input
// (modules/_architecture.py:182-189).
#include "common.cuh"

namespace yad {

constexpr int FE_NFFT = 1000;
constexpr int FE_FR = 8;          // frames per group
constexpr int FE_NMEL = 32;
constexpr int FE_TPQ = YAD_FE_TPQ; // taps per phase over the pair's common window (zero padded)
constexpr int FE_THREADS = 320;

struct FeParams {
  int64_t B, L, T;
  int32_t P, O, width;
  int32_t HG;        // hops per group = FE_FR * FE_NFFT / P
  int32_t SX;        // staged PCM floats per group
  int32_t n_groups;  // ceil(T / FE_FR)
  int32_t groups_per_cta;
  int32_t fb_nnz_pad;  // mel CSR values, rounded up to a multiple of 4
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// One in-place decimation-in-frequency pass of radix R over FE_FR independent 500-point transforms.
// Block size Nb = R * sub: v[r] = x[blk*Nb + n1 + sub*r]; y = DFT_R(v); x[.. + sub*q] = y[q] * W_Nb^(n1*q).
// Reads and writes hit the same addresses (consecutive threads <-> consecutive n1: conflict-free); after the
// four passes (5,5,5,4) bin k = q1 + 5 q2 + 25 q3 + 125 q4 sits at position 100 q1 + 20 q2 + 4 q3 + q4.
template <int R>
__device__ __forceinline__ void fft_pass_dif(float2* __restrict__ x, const float2* __restrict__ tw, int Nb) {
  constexpr int N = 500, NBF = N / R;      // butterflies per frame
  const int sub = Nb / R;
  const int mul = 1000 / Nb;                // W_Nb = W_1000^mul
  for (int item = threadIdx.x; item < FE_FR * NBF; item += blockDim.x) {
    const int f = item / NBF, j = item - f * NBF;
    const int blk = j / sub, n1 = j - blk * sub;
    float2* px = x + f * N + blk * Nb + n1;
    float2 v[R], y[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = px[r * sub];
    if (R == 5) {
      const float c1 = 0.30901699437494745f, c2 = -0.80901699437494745f;
      const float s1 = 0.95105651629515353f, s2 = 0.58778525229247314f;
      const float2 t1 = make_float2(v[1].x + v[4].x, v[1].y + v[4].y);
      const float2 t2 = make_float2(v[2].x + v[3].x, v[2].y + v[3].y);
      const float2 t3 = make_float2(v[1].x - v[4].x, v[1].y - v[4].y);
      const float2 t4 = make_float2(v[2].x - v[3].x, v[2].y - v[3].y);
      const float2 m1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
      const float2 m2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
      const float2 n1v = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
      const float2 n2v = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
      y[0] = make_float2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
      y[1] = make_float2(m1.x + n1v.y, m1.y - n1v.x);   // m1 - i n1
      y[2] = make_float2(m2.x + n2v.y, m2.y - n2v.x);   // m2 - i n2
      y[3 % R] = make_float2(m2.x - n2v.y, m2.y + n2v.x);   // m2 + i n2
      y[4 % R] = make_float2(m1.x - n1v.y, m1.y + n1v.x);   // m1 + i n1
    } else {
      const float2 a = make_float2(v[0].x + v[2].x, v[0].y + v[2].y);
      const float2 b = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
      const float2 c = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
      const float2 d = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
      y[0] = make_float2(a.x + c.x, a.y + c.y);
      y[1] = make_float2(b.x + d.y, b.y - d.x);          // b - i d
      y[2] = make_float2(a.x - c.x, a.y - c.y);
      y[3] = make_float2(b.x - d.y, b.y + d.x);          // b + i d
    }
    px[0] = y[0];
    if (sub > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) px[q * sub] = cmul(y[q], tw[n1 * q * mul]);
    } else {
#pragma unroll
      for (int q = 1; q < R; ++q) px[q * sub] = y[q];
    }
  }
}

__device__ __forceinline__ int fft_pos(int k) {   // digit-reversed location of bin k after the DIF passes
  const int q1 = k % 5, r1 = k / 5;
  const int q2 = r1 % 5, r2 = r1 / 5;
  const int q3 = r2 % 5, q4 = r2 / 5;
  return 100 * q1 + 20 * q2 + 4 * q3 + q4;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Stage the zero-padded PCM span of group g: element i of the span is xpad[x0 + i] = x[x0 + i] (zero outside [0, L)).
// Fast path (clip base 16 B aligned, L % 4 == 0): 16-byte cp.async with zero fill, destination shifted by
// (x0 mod 4) floats so that global and shared addresses are congruent mod 16.  Returns that shift.
__device__ __forceinline__ int fe_stage_async(float* s_x, const float* __restrict__ xb, int64_t x0, int SX, int64_t L,
                                              bool fast) {
  if (fast) {
    const int shift = (int)(((x0 % 4) + 4) % 4);
    const int nchunk = (shift + SX + 3) >> 2;
    const int64_t g0 = x0 - shift;                      // multiple of 4
    for (int c = threadIdx.x; c < nchunk; c += blockDim.x) {
      const int64_t gi = g0 + 4 * (int64_t)c;
      const bool ok = gi >= 0 && gi + 4 <= L;
      cp_async16(s_x + 4 * c, xb + (ok ? gi : 0), ok ? 16 : 0);
    }
    cp_async_commit();
    return shift;
  }
  for (int i = threadIdx.x; i < SX; i += blockDim.x) {
    const int64_t src = x0 + i;
    s_x[i] = (src >= 0 && src < L) ? __ldg(xb + src) : 0.0f;
  }
  return 0;
}

__global__ void __launch_bounds__(FE_THREADS, 2)
frontend_mel_kernel(const float* __restrict__ pcm, const FeParams p, const float* __restrict__ taps,
                    const int32_t* __restrict__ tap_base, const float* __restrict__ window,
                    const float* __restrict__ twiddle, const float* __restrict__ fb_val,
                    const int32_t* __restrict__ fb_bin, const int32_t* __restrict__ fb_start,
                    float* __restrict__ mel) {
  extern __shared__ __align__(16) float fe_smem[];
  const int sxp = (p.SX + 8 + 3) & ~3;
  float* s_x = fe_smem;                                      // staged PCM span (prefetched one group ahead)
  float* s_fr = s_x + sxp;                                   // FE_FR windowed frames; FFT and power in place
  float2* s_tw = reinterpret_cast<float2*>(s_fr + FE_FR * FE_NFFT);  // exp(-2 pi i k / 1000)
  float* s_win = reinterpret_cast<float*>(s_tw + FE_NFFT);
  float* s_fbv = s_win + FE_NFFT;                            // mel filterbank values, CSR over bands
  int* s_fbs = reinterpret_cast<int*>(s_fbv + p.fb_nnz_pad); // [33] row starts, then [32] first bin of each band
  int* s_pos = s_fbs + 2 * FE_NMEL + 4;                      // [501] digit-reversed position of bin k (500 -> 0)

  const int tid = threadIdx.x;
  const int64_t b = blockIdx.y;
  const float* xb = pcm + b * p.L;
  const bool fast = ((p.L & 3) == 0) && ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0);
  const int g_first = blockIdx.x * p.groups_per_cta;
  if (g_first >= p.n_groups) return;
  int shift = fe_stage_async(s_x, xb, (int64_t)g_first * p.HG * p.O - p.width, p.SX, p.L, fast);

  for (int i = tid; i < FE_NFFT; i += blockDim.x) {
    s_tw[i] = make_float2(twiddle[2 * i], twiddle[2 * i + 1]);
    s_win[i] = window[i];
  }
  const int nnz = fb_start[FE_NMEL];
  for (int i = tid; i < nnz; i += blockDim.x) s_fbv[i] = fb_val[i];
  if (tid <= FE_NMEL) s_fbs[tid] = fb_start[tid];
  if (tid < FE_NMEL) s_fbs[FE_NMEL + 1 + tid] = fb_bin[fb_start[tid]];   // bins of a band are contiguous
  for (int i = tid; i <= 500; i += blockDim.x) s_pos[i] = fft_pos(i % 500);

  // resample role: one pair of adjacent phases per thread, hop slices interleaved across the block
  const int npair = p.P >> 1;
  const int nsl = blockDim.x / npair;
  const int pair = tid % npair, sl = tid / npair;
  const bool rs_active = sl < nsl;
  float t0[FE_TPQ], t1[FE_TPQ];
  int base = 0;
  if (rs_active) {
    base = tap_base[pair];
#pragma unroll
    for (int j = 0; j < FE_TPQ; ++j) {
      t0[j] = taps[(pair * 2 + 0) * FE_TPQ + j];
      t1[j] = taps[(pair * 2 + 1) * FE_TPQ + j];
    }
  }
  float2* zf = reinterpret_cast<float2*>(s_fr);

  for (int gi = 0; gi < p.groups_per_cta; ++gi) {
    const int g = g_first + gi;
    if (g >= p.n_groups) break;
    cp_async_wait_all();
    __syncthreads();  // staged span + tables visible; previous group's mel reads of s_fr are done
    // ---- 1. polyphase resample * Hann window -> frames
    if (rs_active) {
      for (int h = sl; h < p.HG; h += nsl) {
        const float* xs = s_x + shift + h * p.O + base;
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (int j = 0; j < FE_TPQ; ++j) {
          const float xv = xs[j];
          a0 = fmaf(t0[j], xv, a0);
          a1 = fmaf(t1[j], xv, a1);
        }
        const int o = h * p.P + 2 * pair;
        const int pos = o % FE_NFFT;
        *reinterpret_cast<float2*>(s_fr + o) = make_float2(a0 * s_win[pos], a1 * s_win[pos + 1]);
      }
    }
    __syncthreads();
    // ---- 2. prefetch the next group's PCM span (s_x is free now); lands while the FFT runs
    if (gi + 1 < p.groups_per_cta && g + 1 < p.n_groups)
      shift = fe_stage_async(s_x, xb, (int64_t)(g + 1) * p.HG * p.O - p.width, p.SX, p.L, fast);
    // ---- 3. in-place 500-point complex FFT of z[n] = x[2n] + i x[2n+1]
    fft_pass_dif<5>(zf, s_tw, 500);
    __syncthreads();
    fft_pass_dif<5>(zf, s_tw, 100);
    __syncthreads();
    fft_pass_dif<5>(zf, s_tw, 20);
    __syncthreads();
    fft_pass_dif<4>(zf, s_tw, 4);
    __syncthreads();
    // ---- 4. real-FFT untangle + power, in place: X[k] = E + W^k O ; X[500-k] = conj(E - W^k O).
    //         The pair (k, 500-k) is owned by one thread; P[k] overwrites Z[k].x, P[500-k] overwrites Z[500-k].x;
    //         bin 500 (= bin 0's partner) goes to Z[0].y.
    for (int item = tid; item < FE_FR * 251; item += blockDim.x) {
      const int f = item / 251, k = item - f * 251;
      const int pk = s_pos[k], pq = s_pos[500 - k];
      float2* zb = zf + f * 500;
      const float2 zk = zb[pk];
      const float2 zq = zb[pq];
      const float2 zn = make_float2(zq.x, -zq.y);
      const float2 E = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y + zn.y));
      const float2 D = make_float2(zk.x - zn.x, zk.y - zn.y);
      const float2 Od = make_float2(0.5f * D.y, -0.5f * D.x);   // -i/2 * D
      const float2 Tt = cmul(s_tw[k], Od);
      const float pr = E.x + Tt.x, pi = E.y + Tt.y, qr = E.x - Tt.x, qi = E.y - Tt.y;
      const float Pk = pr * pr + pi * pi, Pq = qr * qr + qi * qi;
      if (k == 0) {
        zb[pk] = make_float2(Pk, Pq);
      } else {
        zb[pk].x = Pk;
        zb[pq].x = Pq;     // k == 250: same location, same value
      }
    }
    __syncthreads();
    // ---- 5. sparse mel filterbank (band m covers the contiguous bins [bin0, bin0 + len))
    for (int item = tid; item < FE_FR * FE_NMEL; item += blockDim.x) {
      const int m = item / FE_FR, f = item - m * FE_FR;
      const int64_t t = (int64_t)g * FE_FR + f;
      if (t < p.T) {
        const float2* zb = zf + f * 500;
        const int s0 = s_fbs[m], e0 = s_fbs[m + 1];
        int bin = s_fbs[FE_NMEL + 1 + m];
        float acc = 0.0f;
        for (int i = s0; i < e0; ++i, ++bin) {
          const float pw = (bin == 500) ? zb[0].y : zb[s_pos[bin]].x;
          acc = fmaf(pw, s_fbv[i], acc);
        }
        mel[(b * FE_NMEL + m) * p.T + t] = acc;
      }
    }
  }
  cp_async_wait_all();
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

__global__ void __launch_bounds__(FB_THREADS)
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

  // pass 1: per-clip max of dB(mel)  -> first top_db floor
  float mx = -INFINITY;
  for (int64_t t = threadIdx.x; t < T; t += blockDim.x)
    for (int m = 0; m < FE_NMEL; ++m) mx = fmaxf(mx, to_db(mb[m * T + t]));
  const float floor1 = block_reduce<float>(mx, s_redf, fmax_op, -INFINITY) - top_db;

  // pass 2: MFCC column (DCT-II of the clamped dB-mel column, computed ONCE and parked in o1), its dB maximum,
  //         and the moments of the dB-mel plane.  Every later pass touches only frames this thread wrote.
  float mx2 = -INFINITY;
  double s0 = 0.0, q0 = 0.0;
  for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
    float mf[FE_NMEL];
#pragma unroll
    for (int k = 0; k < FE_NMEL; ++k) mf[k] = 0.0f;
    for (int m = 0; m < FE_NMEL; ++m) {
      const float x = fmaxf(to_db(mb[m * T + t]), floor1);
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
      mx2 = fmaxf(mx2, to_db(mf[k]));
    }
  }
  const float floor2 = block_reduce<float>(mx2, s_redf, fmax_op, -INFINITY) - top_db;
  const double n_el = (double)FE_NMEL * (double)T;
  const double S0 = block_reduce<double>(s0, s_redd, dadd_op, 0.0);
  const double Q0 = block_reduce<double>(q0, s_redd, dadd_op, 0.0);
  const float mu0 = (float)(S0 / n_el);

  // pass 3: moments of the clamped dB(MFCC) plane (fp64 accumulators: sum and sum of squares)
  float mu1 = 0.0f, sd0 = 1.0f, sd1 = 1.0f;
  if (standardise) {
    double s1 = 0.0, q1 = 0.0;
    for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
      for (int k = 0; k < FE_NMEL; ++k) {
        const double y = (double)fmaxf(to_db(o1[k * T + t]), floor2);
        s1 += y;
        q1 += y * y;
      }
    }
    const double S1 = block_reduce<double>(s1, s_redd, dadd_op, 0.0);
    const double Q1 = block_reduce<double>(q1, s_redd, dadd_op, 0.0);
    mu1 = (float)(S1 / n_el);
    // unbiased variance = (sum x^2 - n mean^2) / (n - 1), all in fp64
    sd0 = (float)sqrt(fmax(Q0 - S0 * S0 / n_el, 0.0) / (n_el - 1.0));
    sd1 = (float)sqrt(fmax(Q1 - S1 * S1 / n_el, 0.0) / (n_el - 1.0));
  }
  // pass 4: write x_spectral [B, 2, 32, T] (+ optional taps)
  const float den0 = sd0 + 1e-5f, den1 = sd1 + 1e-5f;
  for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
    for (int m = 0; m < FE_NMEL; ++m) {
      const float x = fmaxf(to_db(mb[m * T + t]), floor1);
      if (tap_meldb) tap_meldb[(b * FE_NMEL + m) * T + t] = x;
      o0[m * T + t] = standardise ? __fdiv_rn(x - mu0, den0) : x;
    }
    for (int k = 0; k < FE_NMEL; ++k) {
      const float y = fmaxf(to_db(o1[k * T + t]), floor2);
      if (tap_mfdb) tap_mfdb[(b * FE_NMEL + k) * T + t] = y;
      o1[k * T + t] = standardise ? __fdiv_rn(y - mu1, den1) : y;
    }
  }
}

static size_t fe_smem_bytes(int SX, int nnz_pad) {
  const int sxp = (SX + 8 + 3) & ~3;
  return (size_t)(sxp + FE_FR * FE_NFFT + 2 * FE_NFFT + FE_NFFT + nnz_pad + 2 * FE_NMEL + 4 + 504) * sizeof(float);
}

int init_frontend_attrs() {
  cudaError_t e = cudaFuncSetAttribute(frontend_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(frontend_mel_kernel) failed: %s", cudaGetErrorString(e));
    return YAD_ERR_CUDA;
  }
  return YAD_OK;
}

}  // namespace yad

extern "C" {

int yad_frontend_mel_power(const float* pcm, int64_t B, int64_t L, int32_t P, int32_t O, int32_t width,
                           const float* taps, const int32_t* tap_base, int32_t window_len, const float* window,
                           const float* twiddle, const float* fb_val, const int32_t* fb_bin,
                           const int32_t* fb_start, int32_t fb_nnz, float* mel, int64_t T, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(pcm && taps && tap_base && window && twiddle && fb_val && fb_bin && fb_start && mel,
                "yad_frontend_mel_power: null pointer");
  YAD_CHECK_ARG(P >= 2 && P % 2 == 0 && P / 2 <= FE_THREADS, "yad_frontend_mel_power: P=%d must be even and <= %d", P,
                2 * FE_THREADS);
  YAD_CHECK_ARG((FE_FR * FE_NFFT) % P == 0, "yad_frontend_mel_power: %d-sample frame groups must be whole hops of P=%d",
                FE_FR * FE_NFFT, P);
  YAD_CHECK_ARG(O >= 1 && width >= 0 && window_len >= FE_TPQ, "yad_frontend_mel_power: bad O/width/window_len");
  YAD_CHECK_ARG(B >= 0 && B <= 65535 && L >= 1 && T >= 1, "yad_frontend_mel_power: bad B/L/T");
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
  p.SX = (p.HG - 1) * O + window_len;   // window_len = max(tap_base) + FE_TPQ
  p.n_groups = (int)((T + FE_FR - 1) / FE_FR);
  p.groups_per_cta = 4;
  p.fb_nnz_pad = (fb_nnz + 3) & ~3;
  const size_t smem = fe_smem_bytes(p.SX, p.fb_nnz_pad);
  YAD_CHECK_ARG(smem <= 200 * 1024, "yad_frontend_mel_power: staging span too large (%zu B)", smem);
  if (smem > 48 * 1024)
    YAD_CUDA(cudaFuncSetAttribute(frontend_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((unsigned)((p.n_groups + p.groups_per_cta - 1) / p.groups_per_cta), (unsigned)B);
  frontend_mel_kernel<<<grid, FE_THREADS, smem, (cudaStream_t)stream>>>(pcm, p, taps, tap_base, window, twiddle, fb_val,
                                                                        fb_bin, fb_start, mel);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_frontend_finish(const float* mel, int64_t B, int64_t T, const float* dct, float top_db, int32_t standardise,
                        float* x_spectral, float* tap_meldb, float* tap_mfcc, float* tap_mfdb, yad_stream_t stream) {
  YAD_CHECK_ARG(mel && dct && x_spectral && T >= 2 && B >= 0, "yad_frontend_finish: bad arguments");
  if (B == 0) return YAD_OK;
  yad::frontend_finish_kernel<<<(unsigned)B, yad::FB_THREADS, 0, (cudaStream_t)stream>>>(
      mel, T, dct, top_db, standardise, x_spectral, tap_meldb, tap_mfcc, tap_mfdb);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

}  // extern "C"
