// 500-point complex FFT building blocks kept in registers (host+device so that tools/fft_selftest.cu can run
// the exact butterfly / index code on the CPU).
//
// 1000-point real FFT of one STFT frame (torch.stft inside torchaudio Spectrogram, [ta] functional.py:123) =
// 500-point complex FFT of z[n] = x[2n] + i x[2n+1] followed by an "untangle" pass.  500 = 25 x 20:
//   pass A: thread (n2 in 0..19) loads z[20 n1 + n2], n1 = 0..24, runs a 25-point DFT (5 x 5) in registers and
//           multiplies by W_500^(n2 k1);
//   pass B: thread (k1 in 0..24) loads the 20 values Y[k1][n2], runs a 20-point DFT (4 x 5) in registers;
//           X[k1 + 25 k2] comes out.
// All butterfly indices are compile-time so the 25 / 20 complex values never leave the register file.
#pragma once
#include <cuda_runtime.h>

namespace yad {

#define YAD_HD __host__ __device__ __forceinline__

// A complex value is a float2 (8-byte aligned: every shared-memory exchange is one LDS.64 / STS.64).  On the device all
// complex arithmetic is written with the packed sm_100 instructions (FADD2 / FMUL2 / FFMA2: one instruction for the real
// and the imaginary lane); their operand modifiers (lane swap .LO_HI, per-lane negation .NP, scalar / immediate broadcast)
// absorb the "multiply by -i" and constant-twiddle patterns, so a complex multiply is 2 instructions instead of 4 and a
// radix-5 butterfly 18 instead of 36.  The host versions (tools/fft_selftest.cu) are the same expressions in scalar form.
typedef float2 cf32;
YAD_HD cf32 cmake(float x, float y) {
  cf32 r;
  r.x = x;
  r.y = y;
  return r;
}
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
YAD_HD cf32 cadd(cf32 a, cf32 b) { return __fadd2_rn(a, b); }
YAD_HD cf32 csub(cf32 a, cf32 b) { return __fadd2_rn(a, cmake(-b.x, -b.y)); }
YAD_HD cf32 cscale(cf32 a, float s) { return __fmul2_rn(a, cmake(s, s)); }
YAD_HD cf32 caxpy(cf32 a, float s, cf32 c) { return __ffma2_rn(a, cmake(s, s), c); }     // a * s + c (s real)
YAD_HD cf32 cmulc(cf32 a, float br, float bi) {                                         // a * (br + i bi)
  return __ffma2_rn(cmake(a.y, a.x), cmake(-bi, bi), __fmul2_rn(a, cmake(br, br)));
}
#else
YAD_HD cf32 cadd(cf32 a, cf32 b) { return cmake(a.x + b.x, a.y + b.y); }
YAD_HD cf32 csub(cf32 a, cf32 b) { return cmake(a.x - b.x, a.y - b.y); }
YAD_HD cf32 cscale(cf32 a, float s) { return cmake(a.x * s, a.y * s); }
YAD_HD cf32 caxpy(cf32 a, float s, cf32 c) { return cmake(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y)); }
YAD_HD cf32 cmulc(cf32 a, float br, float bi) {
  return cmake(fmaf(a.y, -bi, a.x * br), fmaf(a.x, bi, a.y * br));
}
#endif
YAD_HD cf32 csub_i(cf32 a, cf32 b) { return cadd(a, cmake(b.y, -b.x)); }   // a - i b
YAD_HD cf32 cadd_i(cf32 a, cf32 b) { return cadd(a, cmake(-b.y, b.x)); }   // a + i b

// forward 5-point DFT of (v0..v4), in place: v_q <- sum_r v_r exp(-2 pi i r q / 5)
YAD_HD void dft5(cf32& v0, cf32& v1, cf32& v2, cf32& v3, cf32& v4) {
  const float c1 = 0.30901699437494745f, c2 = -0.80901699437494745f;
  const float s1 = 0.95105651629515353f, s2 = 0.58778525229247314f;
  const cf32 t1 = cadd(v1, v4), t2 = cadd(v2, v3), t3 = csub(v1, v4), t4 = csub(v2, v3);
  const cf32 m1 = caxpy(t2, c2, caxpy(t1, c1, v0));
  const cf32 m2 = caxpy(t2, c1, caxpy(t1, c2, v0));
  const cf32 n1 = caxpy(t4, s2, cscale(t3, s1));
  const cf32 n2 = caxpy(t4, -s1, cscale(t3, s2));
  v0 = cadd(cadd(v0, t1), t2);
  v1 = csub_i(m1, n1);
  v2 = csub_i(m2, n2);
  v3 = cadd_i(m2, n2);
  v4 = cadd_i(m1, n1);
}

// forward 4-point DFT, in place
YAD_HD void dft4(cf32& v0, cf32& v1, cf32& v2, cf32& v3) {
  const cf32 a = cadd(v0, v2), b = csub(v0, v2), c = cadd(v1, v3), d = csub(v1, v3);
  v0 = cadd(a, c);
  v1 = csub_i(b, d);
  v2 = csub(a, c);
  v3 = cadd_i(b, d);
}

// cos / sin of 2 pi e / N evaluated by the compiler (constexpr Taylor series in double; the arguments stay below 4.1).
// They must be constant expressions: the earlier __builtin_cos / __builtin_sin form was NOT folded for device code and
// cost ~290 runtime FP64 instructions per thread and FFT pass.
constexpr double ce_sincos(double x, bool want_cos) {
  double term = want_cos ? 1.0 : x, sum = term;
  for (int i = 1; i < 24; ++i) {
    const double d = want_cos ? (double)((2 * i - 1) * (2 * i)) : (double)((2 * i) * (2 * i + 1));
    term = -term * x * x / d;
    sum += term;
  }
  return sum;
}
#define YAD_TW_COS(e, N) ((float)::yad::ce_sincos(6.283185307179586476925286766559 * (double)(e) / (double)(N), true))
#define YAD_TW_SIN(e, N) ((float)::yad::ce_sincos(6.283185307179586476925286766559 * (double)(e) / (double)(N), false))

// 25-point forward DFT.  Input v[n1] (natural order).  Output: bin k = a + 5 b (a, b in 0..4) is left in v[5 a + b].
template <int NV>
YAD_HD void dft25(cf32 (&v)[NV]) {
  static_assert(NV >= 25, "dft25 needs 25 values");
#pragma unroll
  for (int n2 = 0; n2 < 5; ++n2) dft5(v[n2], v[5 + n2], v[10 + n2], v[15 + n2], v[20 + n2]);
  // now v[5 a + n2] = sum_{n1} x[5 n1 + n2] W5^(n1 a); twiddle by W25^(n2 a)
#define YAD_T25(a, n2) { constexpr float c_ = YAD_TW_COS((a) * (n2), 25), s_ = -YAD_TW_SIN((a) * (n2), 25); v[5 * a + n2] = cmulc(v[5 * a + n2], c_, s_); }
  YAD_T25(1, 1) YAD_T25(1, 2) YAD_T25(1, 3) YAD_T25(1, 4)
  YAD_T25(2, 1) YAD_T25(2, 2) YAD_T25(2, 3) YAD_T25(2, 4)
  YAD_T25(3, 1) YAD_T25(3, 2) YAD_T25(3, 3) YAD_T25(3, 4)
  YAD_T25(4, 1) YAD_T25(4, 2) YAD_T25(4, 3) YAD_T25(4, 4)
#undef YAD_T25
#pragma unroll
  for (int a = 0; a < 5; ++a) dft5(v[5 * a], v[5 * a + 1], v[5 * a + 2], v[5 * a + 3], v[5 * a + 4]);
}

// 20-point forward DFT.  Input v[n] natural order (n = 5 n1 + n2, n1 in 0..3, n2 in 0..4).
// Output: bin k = a + 4 b (a in 0..3, b in 0..4) is left in v[5 a + b].
template <int NV>
YAD_HD void dft20(cf32 (&v)[NV]) {
  static_assert(NV >= 20, "dft20 needs 20 values");
#pragma unroll
  for (int n2 = 0; n2 < 5; ++n2) dft4(v[n2], v[5 + n2], v[10 + n2], v[15 + n2]);
  // v[5 a + n2] = sum_{n1} x[5 n1 + n2] W4^(n1 a); twiddle by W20^(n2 a)
#define YAD_T20(a, n2) { constexpr float c_ = YAD_TW_COS((a) * (n2), 20), s_ = -YAD_TW_SIN((a) * (n2), 20); v[5 * a + n2] = cmulc(v[5 * a + n2], c_, s_); }
  YAD_T20(1, 1) YAD_T20(1, 2) YAD_T20(1, 3) YAD_T20(1, 4)
  YAD_T20(2, 1) YAD_T20(2, 2) YAD_T20(2, 3) YAD_T20(2, 4)
  YAD_T20(3, 1) YAD_T20(3, 2) YAD_T20(3, 3) YAD_T20(3, 4)
#undef YAD_T20
#pragma unroll
  for (int a = 0; a < 4; ++a) dft5(v[5 * a], v[5 * a + 1], v[5 * a + 2], v[5 * a + 3], v[5 * a + 4]);
}

// Frame-buffer layouts (units: cf32 = 2 floats).  The three layouts alias the same shared-memory buffer, separated by
// barriers; their frame strides are chosen so that consecutive work items (which straddle frames inside a warp) keep
// walking through distinct banks (see frontend.cu).
constexpr int FFT_NZ = 500;           // complex points per frame
constexpr int FFT_Z_STRIDE = 500;     // natural-order input frames z
constexpr int FFT_X_STRIDE = 505;     // output spectrum (aliases the frame buffer after pass A): pass-B work item i = 25 f + k1
                                      // writes slot 480 f + i + 25 k2, and 480 slots = 0 (mod 32 banks): 16 consecutive
                                      // items always hit 16 distinct bank pairs, also across a frame boundary
constexpr int FFT_Y_PITCH = 21;       // Y[k1][n2] row pitch
constexpr int FFT_Y_STRIDE = 525;     // 25 rows x 21: 1050 words = 26 (mod 32)

// pass A, work item (frame-local n2): register r (= 5 a + b after dft25) holds k1 = a + 5 b
YAD_HD int passA_k1_of_reg(int r) { return (r / 5) + 5 * (r % 5); }
// pass B, register r (= 5 a + b after dft20) holds k2 = a + 4 b
YAD_HD int passB_k2_of_reg(int r) { return (r / 5) + 4 * (r % 5); }

}  // namespace yad
