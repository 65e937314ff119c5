// CUDA-core convolutions (fp32 accumulate):
//  * yad_conv_stem : the 2->64 7x7 stride-2 stem conv1 (modules/_backbone.py:127,143). K = 98 is too
//    thin for a tensor-core tile, and its input is the frontend's NCHW fp32 plane, so it gets its own
//    direct kernel that reads NCHW fp32 and writes NHWC (f32 | bf16).
//  * yad_conv_simt : generic NHWC implicit GEMM for the fp32 parity mode (every conv of the network,
//    any kernel / stride / padding / channel count), 64x64x16 register-tiled.
#include "common.cuh"

namespace yad {

// ------------------------------------------------------------------------------------ stem
constexpr int STEM_CO = 64, STEM_K = 7, STEM_CI = 2, STEM_TW = 32, STEM_THREADS = 128;

template <typename T>
__global__ void __launch_bounds__(STEM_THREADS)
conv_stem_kernel(const float* __restrict__ x, int H, int W, const float* __restrict__ wgt, T* __restrict__ out) {
  // grid: (Ho, B). One CTA = one output row; loops over 32-pixel column tiles.
  __shared__ __align__(16) float s_w[STEM_K * STEM_K * STEM_CI][STEM_CO];          // 25 KB, [tap][co]
  __shared__ float s_x[STEM_CI][STEM_K][2 * STEM_TW + STEM_K - 2 + 1];              // patch, 70 cols
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;   // (H + 2*3 - 7) / 2 + 1
  const int ho = blockIdx.x, b = blockIdx.y;
  for (int i = threadIdx.x; i < STEM_K * STEM_K * STEM_CI * STEM_CO; i += STEM_THREADS) (&s_w[0][0])[i] = wgt[i];
  const int px = threadIdx.x & 31, cg = threadIdx.x >> 5;  // pixel within tile, 16-channel group
  const int PW = 2 * STEM_TW + STEM_K - 2;                 // 69 input columns per tile
  for (int w0 = 0; w0 < Wo; w0 += STEM_TW) {
    __syncthreads();
    for (int i = threadIdx.x; i < STEM_CI * STEM_K * PW; i += STEM_THREADS) {
      const int col = i % PW, r = (i / PW) % STEM_K, c = i / (PW * STEM_K);
      const int hi = 2 * ho + r - 3, wi = 2 * w0 + col - 3;
      float v = 0.0f;
      if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = x[(((int64_t)b * STEM_CI + c) * H + hi) * W + wi];
      s_x[c][r][col] = v;
    }
    __syncthreads();
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
#pragma unroll
    for (int kh = 0; kh < STEM_K; ++kh) {
#pragma unroll
      for (int kw = 0; kw < STEM_K; ++kw) {
#pragma unroll
        for (int c = 0; c < STEM_CI; ++c) {
          const float xv = s_x[c][kh][2 * px + kw];
          const float4* wp = reinterpret_cast<const float4*>(&s_w[(kh * STEM_K + kw) * STEM_CI + c][cg * 16]);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 w4 = wp[j4];
            acc[j4 * 4 + 0] = fmaf(xv, w4.x, acc[j4 * 4 + 0]);
            acc[j4 * 4 + 1] = fmaf(xv, w4.y, acc[j4 * 4 + 1]);
            acc[j4 * 4 + 2] = fmaf(xv, w4.z, acc[j4 * 4 + 2]);
            acc[j4 * 4 + 3] = fmaf(xv, w4.w, acc[j4 * 4 + 3]);
          }
        }
      }
    }
    const int wo = w0 + px;
    if (wo < Wo) {
      T* o = out + (((int64_t)b * Ho + ho) * Wo + wo) * STEM_CO + cg * 16;
#pragma unroll
      for (int j = 0; j < 16; ++j) st_from_float(o + j, acc[j]);
    }
  }
}

// ------------------------------------------------------------------------------------ generic
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

// DGRAD = false: out[b,ho,wo,:] = sum_taps in[b, ho*s + k - p, ...] W[tap]           (d.H, d.W = input size, Ho, Wo = output size)
// DGRAD = true : the data gradient of that conv in gather form.  `in` is dY [B, d.H, d.W, d.Cin] (d.H, d.W, d.Cin hold the
//                FORWARD output size / channels), `out` is dX [B, Ho, Wo, d.Cout] (Ho, Wo, d.Cout = forward input size /
//                channels), wgt is [kh][kw][Cout_fwd][Cin_fwd]; pixel (hi, wi) of dX receives tap (kh, kw) from
//                dY[(hi + p - kh) / s] when that division is exact and in range.
template <typename T, bool DGRAD>
__global__ void __launch_bounds__(SG_THREADS)
conv_simt_kernel(const yad_conv_desc d, int Ho, int Wo, const T* __restrict__ in, const T* __restrict__ wgt, int ld_w,
                 const float* __restrict__ bias, const T* __restrict__ res, T* __restrict__ out) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int64_t M = (int64_t)d.B * Ho * Wo;
  const int64_t m0 = (int64_t)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  // A-load role: row = tid / 4, k-quad = tid % 4
  const int arow = tid >> 2, akq = (tid & 3) * 4;
  const int64_t am = m0 + arow;
  int ab = 0, aho = 0, awo = 0;
  const bool a_ok = am < M;
  if (a_ok) {
    awo = (int)(am % Wo);
    aho = (int)((am / Wo) % Ho);
    ab = (int)(am / ((int64_t)Wo * Ho));
  }
  // B-load role: k = tid / 16, n-quad = (tid % 16) * 4
  const int bk = tid >> 4, bnq = (tid & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int kh = 0; kh < d.kh; ++kh) {
    for (int kw = 0; kw < d.kw; ++kw) {
      int hi, wi;
      bool pix_ok;
      if (DGRAD) {
        const int th = aho + d.ph - kh, tw = awo + d.pw - kw;
        hi = th / d.sh;
        wi = tw / d.sw;
        pix_ok = a_ok && th >= 0 && tw >= 0 && hi * d.sh == th && wi * d.sw == tw && hi < d.H && wi < d.W;
      } else {
        hi = aho * d.sh + kh - d.ph;
        wi = awo * d.sw + kw - d.pw;
        pix_ok = a_ok && hi >= 0 && hi < d.H && wi >= 0 && wi < d.W;
      }
      const T* ap = in + (((int64_t)ab * d.H + hi) * d.W + wi) * d.ld_in;
      const T* wp = wgt + (int64_t)(kh * d.kw + kw) * d.Cin * ld_w;
      for (int c0 = 0; c0 < d.Cin; c0 += SG_BK) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c0 + akq + e;
          As[akq + e][arow] = (pix_ok && c < d.Cin) ? ld_as_float(ap + c) : 0.0f;
        }
        {
          const int c = c0 + bk;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int n = n0 + bnq + e;
            Bs[bk][bnq + e] = (c < d.Cin && n < ld_w) ? ld_as_float(wp + (int64_t)c * ld_w + n) : 0.0f;
          }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
          float a[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= d.Cout) continue;
      float v = acc[i][j] + (bias != nullptr ? bias[n] : 0.0f);
      if (res != nullptr) v += ld_as_float(res + m * d.ld_res + n);
      v = apply_act(v, d.act);
      st_from_float(out + m * d.ld_out + d.co_off + n, v);
    }
  }
}

}  // namespace yad

extern "C" {

int yad_conv_stem(const float* x_nchw, int64_t B, int32_t H, int32_t W, const float* weight, void* out,
                  int32_t out_dtype, yad_stream_t stream) {
  YAD_CHECK_ARG(x_nchw && weight && out, "yad_conv_stem: null pointer");
  YAD_CHECK_ARG(H >= 1 && W >= 1, "yad_conv_stem: bad H=%d W=%d", H, W);
  YAD_CHECK_ARG(out_dtype == YAD_F32 || out_dtype == YAD_BF16, "yad_conv_stem: bad out dtype %d", out_dtype);
  YAD_CHECK_ARG(B <= 65535, "yad_conv_stem: B=%lld exceeds grid.y", (long long)B);
  if (B == 0) return YAD_OK;
  dim3 grid((unsigned)((H - 1) / 2 + 1), (unsigned)B);
  if (out_dtype == YAD_F32)
    yad::conv_stem_kernel<float><<<grid, yad::STEM_THREADS, 0, (cudaStream_t)stream>>>(x_nchw, H, W, weight, (float*)out);
  else
    yad::conv_stem_kernel<__nv_bfloat16>
        <<<grid, yad::STEM_THREADS, 0, (cudaStream_t)stream>>>(x_nchw, H, W, weight, (__nv_bfloat16*)out);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_conv_simt(const yad_conv_desc* d, int32_t dtype, const void* in, const void* weight, int32_t ld_w,
                  const float* bias, const void* residual, void* out, yad_stream_t stream) {
  YAD_CHECK_ARG(d && in && weight && out, "yad_conv_simt: null pointer");
  YAD_CHECK_ARG(dtype == YAD_F32 || dtype == YAD_BF16, "yad_conv_simt: bad dtype %d", dtype);
  YAD_CHECK_ARG(d->Cin >= 1 && d->ld_in >= d->Cin && d->Cout >= 1 && ld_w >= d->Cout && d->ld_out >= d->co_off + d->Cout,
                "yad_conv_simt: inconsistent channel counts");
  YAD_CHECK_ARG(d->sh >= 1 && d->sw >= 1 && d->kh >= 1 && d->kw >= 1, "yad_conv_simt: bad kernel/stride");
  YAD_CHECK_ARG(residual == nullptr || d->ld_res >= d->Cout, "yad_conv_simt: bad ld_res");
  const int Ho = (d->H + 2 * d->ph - d->kh) / d->sh + 1;
  const int Wo = (d->W + 2 * d->pw - d->kw) / d->sw + 1;
  YAD_CHECK_ARG(Ho >= 1 && Wo >= 1, "yad_conv_simt: empty output");
  if (d->B == 0) return YAD_OK;
  const int64_t M = (int64_t)d->B * Ho * Wo;
  dim3 grid((unsigned)((M + yad::SG_BM - 1) / yad::SG_BM), (unsigned)((d->Cout + yad::SG_BN - 1) / yad::SG_BN));
  if (dtype == YAD_F32)
    yad::conv_simt_kernel<float, false><<<grid, yad::SG_THREADS, 0, (cudaStream_t)stream>>>(
        *d, Ho, Wo, (const float*)in, (const float*)weight, ld_w, bias, (const float*)residual, (float*)out);
  else
    yad::conv_simt_kernel<__nv_bfloat16, false><<<grid, yad::SG_THREADS, 0, (cudaStream_t)stream>>>(
        *d, Ho, Wo, (const __nv_bfloat16*)in, (const __nv_bfloat16*)weight, ld_w, bias,
        (const __nv_bfloat16*)residual, (__nv_bfloat16*)out);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_conv_dgrad(const yad_conv_desc* d, const float* dy, const float* weight_t, const float* dx_in, float* dx,
                   yad_stream_t stream) {
  YAD_CHECK_ARG(d && dy && weight_t && dx, "yad_conv_dgrad: null pointer");
  YAD_CHECK_ARG(d->sh >= 1 && d->sw >= 1 && d->kh >= 1 && d->kw >= 1 && d->Cin >= 1 && d->Cout >= 1, "yad_conv_dgrad: bad descriptor");
  const int Ho = (d->H + 2 * d->ph - d->kh) / d->sh + 1;
  const int Wo = (d->W + 2 * d->pw - d->kw) / d->sw + 1;
  YAD_CHECK_ARG(Ho >= 1 && Wo >= 1, "yad_conv_dgrad: empty output");
  YAD_CHECK_ARG(d->ld_in >= d->Cin && d->ld_out >= d->co_off + d->Cout, "yad_conv_dgrad: bad pitches");
  if (d->B == 0) return YAD_OK;
  // the kernel's view: "input" = dY [B, Ho, Wo, Cout] (pitch ld_out, channel slice co_off), "output" = dX [B, H, W, Cin]
  yad_conv_desc k = *d;
  k.H = Ho;
  k.W = Wo;
  k.Cin = d->Cout;
  k.ld_in = d->ld_out;
  k.Cout = d->Cin;
  k.ld_out = d->ld_in;
  k.co_off = 0;
  k.act = YAD_ACT_NONE;
  k.ld_res = dx_in != nullptr ? d->ld_in : 0;
  const int64_t M = (int64_t)d->B * d->H * d->W;
  dim3 grid((unsigned)((M + yad::SG_BM - 1) / yad::SG_BM), (unsigned)((d->Cin + yad::SG_BN - 1) / yad::SG_BN));
  yad::conv_simt_kernel<float, true><<<grid, yad::SG_THREADS, 0, (cudaStream_t)stream>>>(
      k, d->H, d->W, dy + d->co_off, weight_t, d->Cin, nullptr, dx_in, dx);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

}  // extern "C"
