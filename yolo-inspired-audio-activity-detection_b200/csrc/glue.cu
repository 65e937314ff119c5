// Neck glue on NHWC tensors (dtype f32 | bf16, fp32 math): H-mean, bilinear x2 / x0.5 along W,
// cascaded 5-wide max pools.  Every kernel writes into a channel slice of its destination so the
// reference's torch.cat calls (modules/_common.py:183,210,213,257-258) never materialise.
#include "common.cuh"

namespace yad {

// adaptive_avg_pool2d(x, (1, W))  -  modules/_common.py:248-252
template <typename T>
__global__ void hmean_kernel(const T* __restrict__ in, int64_t B, int H, int W, int C, int ld_in, int64_t isw, int64_t ish,
                             int64_t isb, T* __restrict__ out, int ld_out, int co_off) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = B * W * C;
  if (gid >= n) return;
  const int c = (int)(gid % C);
  const int64_t bw = gid / C;
  const int w = (int)(bw % W);
  const int64_t b = bw / W;
  float acc = 0.0f;
  for (int h = 0; h < H; ++h) acc += ld_as_float(in + (b * isb + h * ish + w * isw) * (int64_t)ld_in + c);
  st_from_float(out + (b * W + w) * (int64_t)ld_out + co_off + c, acc / (float)H);
}

// bf16 fast path: 8 channels (16 bytes) per thread
__global__ void hmean_bf16x8_kernel(const __nv_bfloat16* __restrict__ in, int64_t B, int H, int W, int C8, int ld_in,
                                    int64_t isw, int64_t ish, int64_t isb, __nv_bfloat16* __restrict__ out, int ld_out,
                                    int co_off) {
  pdl_wait();
  pdl_trigger();
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = B * W * C8;
  if (gid >= n) return;
  const uint32_t g32 = (uint32_t)gid;            // the host guarantees n < 2^32: 32-bit index arithmetic (64-bit div / mod
  const int c = (int)(g32 % (uint32_t)C8) * 8;   // cost more than the loads of the small stages)
  const uint32_t bw = g32 / (uint32_t)C8;
  const int w = (int)(bw % (uint32_t)W);
  const int64_t b = bw / (uint32_t)W;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  const __nv_bfloat16* p0 = in + (b * isb + w * isw) * (int64_t)ld_in + c;
  for (int h0 = 0; h0 < H; h0 += 8) {          // up to 8 independent 16-byte loads in flight per thread (H = 8, 4, 2)
    uint4 u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      u[j] = h0 + j < H ? __ldg(reinterpret_cast<const uint4*>(p0 + (h0 + j) * ish * (int64_t)ld_in)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t ww[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] += __uint_as_float(ww[e] << 16);
        acc[2 * e + 1] += __uint_as_float(ww[e] & 0xffff0000u);
      }
    }
  }
  const float inv = 1.0f / (float)H;
  uint32_t o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * e] / (float)H, acc[2 * e + 1] / (float)H);
    o[e] = *reinterpret_cast<uint32_t*>(&h2);
  }
  (void)inv;
  *reinterpret_cast<uint4*>(out + (b * W + w) * (int64_t)ld_out + co_off + c) = make_uint4(o[0], o[1], o[2], o[3]);
}

// F.interpolate(mode="bilinear", align_corners=False) along W only - modules/_common.py:173-174,181-182
//   x2 : out[2k] = .75 x[k] + .25 x[max(k-1,0)] ; out[2k+1] = .75 x[k] + .25 x[min(k+1,W-1)]
//   x.5: out[k]  = .5 x[2k] + .5 x[2k+1]
template <typename T>
__global__ void resize_w_kernel(const T* __restrict__ in, int64_t B, int W, int C, int ld_in, int ci_off, int up,
                                T* __restrict__ out, int ld_out, int co_off) {
  const int Wo = up ? 2 * W : W / 2;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = B * Wo * C;
  if (gid >= n) return;
  const int c = (int)(gid % C);
  const int64_t bw = gid / C;
  const int wo = (int)(bw % Wo);
  const int64_t b = bw / Wo;
  const T* row = in + b * W * (int64_t)ld_in + ci_off + c;
  float v;
  if (up) {
    const int k = wo >> 1;
    const int k2 = (wo & 1) ? min(k + 1, W - 1) : max(k - 1, 0);
    // torch: w0 * x[i0] + w1 * x[i1] with i0 < i1 (lambda = .25 / .75)
    const float a = ld_as_float(row + (int64_t)k * ld_in), nb = ld_as_float(row + (int64_t)k2 * ld_in);
    v = (wo & 1) ? (0.75f * a + 0.25f * nb) : (0.25f * nb + 0.75f * a);
  } else {
    v = 0.5f * ld_as_float(row + (int64_t)(2 * wo) * ld_in) + 0.5f * ld_as_float(row + (int64_t)(2 * wo + 1) * ld_in);
  }
  st_from_float(out + (b * Wo + wo) * (int64_t)ld_out + co_off + c, v);
}

// bf16 fast paths of the two kernels above and below: 8 channels (16 bytes) per thread, same arithmetic per element
__device__ __forceinline__ void bf8_unpack(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __uint_as_float(w[e] << 16);
    f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 bf8_pack(const float (&f)[8]) {
  uint32_t o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    o[e] = *reinterpret_cast<uint32_t*>(&h2);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void resize_w_bf16x8_kernel(const __nv_bfloat16* __restrict__ in, int64_t B, int W, int C8, int ld_in, int ci_off, int up,
                                       __nv_bfloat16* __restrict__ out, int ld_out, int co_off) {
  pdl_wait();
  pdl_trigger();
  const int Wo = up ? 2 * W : W / 2;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * Wo * C8) return;
  const uint32_t g32 = (uint32_t)gid;            // 32-bit index arithmetic (the host checks the element count)
  const int c = (int)(g32 % (uint32_t)C8) * 8;
  const uint32_t bw = g32 / (uint32_t)C8;
  const int wo = (int)(bw % (uint32_t)Wo);
  const int64_t b = bw / (uint32_t)Wo;
  const __nv_bfloat16* row = in + b * W * (int64_t)ld_in + ci_off + c;
  float a[8], nb[8], v[8];
  if (up) {
    const int k = wo >> 1;
    const int k2 = (wo & 1) ? min(k + 1, W - 1) : max(k - 1, 0);
    bf8_unpack(__ldg(reinterpret_cast<const uint4*>(row + (int64_t)k * ld_in)), a);
    bf8_unpack(__ldg(reinterpret_cast<const uint4*>(row + (int64_t)k2 * ld_in)), nb);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (wo & 1) ? (0.75f * a[j] + 0.25f * nb[j]) : (0.25f * nb[j] + 0.75f * a[j]);
  } else {
    bf8_unpack(__ldg(reinterpret_cast<const uint4*>(row + (int64_t)(2 * wo) * ld_in)), a);
    bf8_unpack(__ldg(reinterpret_cast<const uint4*>(row + (int64_t)(2 * wo + 1) * ld_in)), nb);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.5f * a[j] + 0.5f * nb[j];
  }
  *reinterpret_cast<uint4*>(out + (b * Wo + wo) * (int64_t)ld_out + co_off + c) = bf8_pack(v);
}

__global__ void sppf_bf16x8_kernel(const __nv_bfloat16* __restrict__ in, int64_t B, int W, int C8, int C, int ld_in, int ci_off,
                                   __nv_bfloat16* __restrict__ out, int ld_out, int co_off) {
  pdl_wait();
  pdl_trigger();
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * W * C8) return;
  const uint32_t g32 = (uint32_t)gid;            // 32-bit index arithmetic (the host checks the element count)
  const int c = (int)(g32 % (uint32_t)C8) * 8;
  const uint32_t bw = g32 / (uint32_t)C8;
  const int w = (int)(bw % (uint32_t)W);
  const int64_t b = bw / (uint32_t)W;
  const __nv_bfloat16* row = in + b * W * (int64_t)ld_in + ci_off + c;
  float m1[8], m2[8], m3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m1[j] = m2[j] = m3[j] = -INFINITY;
#pragma unroll
  for (int d = -6; d <= 6; ++d) {
    const int x = w + d;
    if (x < 0 || x >= W) continue;
    float v[8];
    bf8_unpack(__ldg(reinterpret_cast<const uint4*>(row + (int64_t)x * ld_in)), v);
    const int ad = d < 0 ? -d : d;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (ad <= 2) m1[j] = fmaxf(m1[j], v[j]);
      if (ad <= 4) m2[j] = fmaxf(m2[j], v[j]);
      m3[j] = fmaxf(m3[j], v[j]);
    }
  }
  __nv_bfloat16* o = out + (b * W + w) * (int64_t)ld_out + co_off + c;
  *reinterpret_cast<uint4*>(o) = bf8_pack(m1);
  *reinterpret_cast<uint4*>(o + C) = bf8_pack(m2);
  *reinterpret_cast<uint4*>(o + 2 * C) = bf8_pack(m3);
}

// three cascaded MaxPool2d(k=5, s=1, p=2) at H=1 == running max over windows of 5 / 9 / 13 (-inf padding)
template <typename T>
__global__ void sppf_kernel(const T* __restrict__ in, int64_t B, int W, int C, int ld_in, int ci_off,
                            T* __restrict__ out, int ld_out, int co_off) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = B * W * C;
  if (gid >= n) return;
  const int c = (int)(gid % C);
  const int64_t bw = gid / C;
  const int w = (int)(bw % W);
  const int64_t b = bw / W;
  const T* row = in + b * W * (int64_t)ld_in + ci_off + c;
  float m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  for (int d = -6; d <= 6; ++d) {
    const int x = w + d;
    if (x < 0 || x >= W) continue;
    const float v = ld_as_float(row + (int64_t)x * ld_in);
    const int ad = d < 0 ? -d : d;
    if (ad <= 2) m1 = fmaxf(m1, v);
    if (ad <= 4) m2 = fmaxf(m2, v);
    m3 = fmaxf(m3, v);
  }
  T* o = out + (b * W + w) * (int64_t)ld_out + co_off + c;
  st_from_float(o, m1);
  st_from_float(o + C, m2);
  st_from_float(o + 2 * C, m3);
}

// RepVGG train-form merge (modules/_common.py:90-95): out = lrelu(a + b [+ scale*x + shift]) where a, b are
// the already-activated 3x3 / 1x1 branches and (scale, shift) is the eval-mode identity BatchNorm.
template <typename T>
__global__ void repvgg_merge_kernel(const T* __restrict__ a, const T* __restrict__ bb, const T* __restrict__ x,
                                    const float* __restrict__ scale, const float* __restrict__ shift, int64_t npix,
                                    int C, int ld_ab, int ld_x, T* __restrict__ out, int ld_out, int co_off, int act) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= npix * C) return;
  const int c = (int)(gid % C);
  const int64_t px = gid / C;
  float v = ld_as_float(a + px * ld_ab + c) + ld_as_float(bb + px * ld_ab + c);
  if (x != nullptr) v += fmaf(ld_as_float(x + px * ld_x + c), scale[c], shift[c]);
  st_from_float(out + px * ld_out + co_off + c, apply_act(v, act));
}

// max over the rows h - r .. h + r of one column (clipped to the image: max_pool2d pads with -inf).  Second half of the 2-D
// SPPF pools when the neck keeps its height (backbone: custom): k cascaded 5x5 / stride-1 max pools = a (4k+1) x (4k+1) box
// max = this (r = 2k) applied to the k-fold W cascade that yad_sppf_pools writes (max is separable and the cascades commute).
template <typename T>
__global__ void maxpool_h_kernel(const T* __restrict__ in, int64_t B, int H, int W, int C, int ld_in, int ci_off, int r,
                                 T* __restrict__ out, int ld_out, int co_off) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * H * W * C) return;
  const int c = (int)(gid % C);
  int64_t px = gid / C;
  const int w = (int)(px % W);
  px /= W;
  const int h = (int)(px % H);
  const int64_t b = px / H;
  float m = -INFINITY;
  for (int y = max(h - r, 0); y <= min(h + r, H - 1); ++y)
    m = fmaxf(m, ld_as_float(in + ((b * H + y) * W + w) * (int64_t)ld_in + ci_off + c));
  st_from_float(out + ((b * H + h) * W + w) * (int64_t)ld_out + co_off + c, m);
}

// x_spectral NCHW f32 -> NHWC (dtype) with channel pitch ld_out: input layout of the custom backbone's first conv
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, int64_t B, int C, int64_t HW, T* __restrict__ out, int ld_out) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * HW) return;
  const int64_t b = gid / HW, px = gid - b * HW;
  for (int c = 0; c < C; ++c) st_from_float(out + gid * ld_out + c, in[(b * C + c) * HW + px]);
}


// File-rate resampling in front of the network (inference.py:152-159: torchaudio.transforms.Resample(orig_freq = the file's
// rate, new_freq = the model's sample_rate) with its default Hann-windowed sinc bank, [ta] functional.py:1405-1431):
//   out[b, i * P + p] = sum_k kernel[p][k] * xpad[b, i * O + k],  xpad = x zero-padded by (width, width + O),
// P / O = new / orig rate over their gcd, first ceil(P * L / O) samples.  One thread per output sample; a cold path (runs once
// per file batch, ~KW = 2 * width + O multiply-adds per sample), so no staging: taps and samples come through L1.
template <typename Tin>
__global__ void resample_sinc_kernel(const Tin* __restrict__ x, int64_t B, int64_t L, int O, int P, int width, int KW,
                                     const float* __restrict__ kernel, float* __restrict__ out, int64_t Lout) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Lout) return;
  const int64_t b = idx / Lout, t = idx - b * Lout;
  const int64_t i = t / P;
  const int ph = (int)(t - i * P);
  const int64_t base = i * O - width;
  const Tin* xb = x + b * L;
  const float* kp = kernel + (int64_t)ph * KW;
  const int k_lo = base < 0 ? (int)(-base) : 0;
  const int64_t hi = L - base;
  const int k_hi = hi < KW ? (int)(hi > 0 ? hi : 0) : KW;
  float acc = 0.0f;
  for (int k = k_lo; k < k_hi; ++k) {
    float v;
    if (sizeof(Tin) == 2)
      v = (float)xb[base + k] * (1.0f / 32768.0f);      // 16-bit PCM: what torchaudio.load(normalize=True) hands the reference
    else
      v = (float)xb[base + k];
    acc = fmaf(__ldg(kp + k), v, acc);
  }
  out[idx] = acc;
}

}  // namespace yad

#define YAD_DISPATCH_DTYPE(dtype, KERNEL, ...)                                              \
  if ((dtype) == YAD_F32) {                                                                 \
    using T = float;                                                                        \
    KERNEL<T><<<blocks, threads, 0, (cudaStream_t)stream>>>(__VA_ARGS__);                   \
  } else {                                                                                  \
    using T = __nv_bfloat16;                                                                \
    KERNEL<T><<<blocks, threads, 0, (cudaStream_t)stream>>>(__VA_ARGS__);                   \
  }

extern "C" {

int yad_hmean(const void* in, int32_t dtype, int64_t B, int32_t H, int32_t W, int32_t C, int32_t ld_in, int32_t in_sw,
              int32_t in_sh, int32_t in_sb, void* out, int32_t ld_out, int32_t co_off, yad_stream_t stream) {
  YAD_CHECK_ARG(in && out && H >= 1 && W >= 1 && C >= 1 && ld_in >= C && ld_out >= co_off + C,
                "yad_hmean: bad arguments");
  YAD_CHECK_ARG(dtype == YAD_F32 || dtype == YAD_BF16, "yad_hmean: bad dtype %d", dtype);
  const int64_t n = B * W * C;
  if (n == 0) return YAD_OK;
  const int threads = 256;
  if (dtype == YAD_BF16 && C % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && co_off % 8 == 0 &&
      reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0) {
    const int64_t n8 = B * W * (C / 8);
    YAD_CHECK_ARG(n8 < (1ll << 32), "yad_hmean: tensor too large for the 32-bit index path");
    YAD_CUDA(yad::launch_pdl(yad::hmean_bf16x8_kernel, dim3((unsigned)((n8 + threads - 1) / threads)), dim3(threads), 0, (cudaStream_t)stream,
                             (const __nv_bfloat16*)in, B, H, W, C / 8, ld_in, (int64_t)in_sw, (int64_t)in_sh, (int64_t)in_sb,
                             (__nv_bfloat16*)out, ld_out, co_off));
    return YAD_OK;
  }
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  YAD_DISPATCH_DTYPE(dtype, yad::hmean_kernel, (const T*)in, B, H, W, C, ld_in, (int64_t)in_sw, (int64_t)in_sh, (int64_t)in_sb,
                     (T*)out, ld_out, co_off);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_resize_w(const void* in, int32_t dtype, int64_t B, int32_t W, int32_t C, int32_t ld_in, int32_t ci_off,
                 int32_t up, void* out, int32_t ld_out, int32_t co_off, yad_stream_t stream) {
  YAD_CHECK_ARG(in && out && W >= 1 && C >= 1 && ld_in >= ci_off + C && ld_out >= co_off + C,
                "yad_resize_w: bad arguments");
  YAD_CHECK_ARG(up || (W % 2 == 0), "yad_resize_w: x0.5 needs even W (got %d)", W);
  YAD_CHECK_ARG(dtype == YAD_F32 || dtype == YAD_BF16, "yad_resize_w: bad dtype %d", dtype);
  const int Wo = up ? 2 * W : W / 2;
  const int64_t n = B * Wo * C;
  if (n == 0) return YAD_OK;
  const int threads = 256;
  if (dtype == YAD_BF16 && C % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && ci_off % 8 == 0 && co_off % 8 == 0 &&
      reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 && B * Wo * (int64_t)(C / 8) < (1ll << 32)) {
    const int64_t n8 = B * Wo * (C / 8);
    YAD_CUDA(yad::launch_pdl(yad::resize_w_bf16x8_kernel, dim3((unsigned)((n8 + threads - 1) / threads)), dim3(threads), 0,
                             (cudaStream_t)stream, (const __nv_bfloat16*)in, B, W, C / 8, ld_in, ci_off, up, (__nv_bfloat16*)out, ld_out,
                             co_off));
    return YAD_OK;
  }
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  YAD_DISPATCH_DTYPE(dtype, yad::resize_w_kernel, (const T*)in, B, W, C, ld_in, ci_off, up, (T*)out, ld_out, co_off);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_sppf_pools(const void* in, int32_t dtype, int64_t B, int32_t W, int32_t C, int32_t ld_in, int32_t ci_off,
                   void* out, int32_t ld_out, int32_t co_off, yad_stream_t stream) {
  YAD_CHECK_ARG(in && out && W >= 1 && C >= 1 && ld_in >= ci_off + C && ld_out >= co_off + 3 * C,
                "yad_sppf_pools: bad arguments");
  YAD_CHECK_ARG(dtype == YAD_F32 || dtype == YAD_BF16, "yad_sppf_pools: bad dtype %d", dtype);
  const int64_t n = B * W * C;
  if (n == 0) return YAD_OK;
  const int threads = 256;
  if (dtype == YAD_BF16 && C % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && ci_off % 8 == 0 && co_off % 8 == 0 &&
      reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 && B * W * (int64_t)(C / 8) < (1ll << 32)) {
    const int64_t n8 = B * W * (C / 8);
    YAD_CUDA(yad::launch_pdl(yad::sppf_bf16x8_kernel, dim3((unsigned)((n8 + threads - 1) / threads)), dim3(threads), 0,
                             (cudaStream_t)stream, (const __nv_bfloat16*)in, B, W, C / 8, C, ld_in, ci_off, (__nv_bfloat16*)out, ld_out,
                             co_off));
    return YAD_OK;
  }
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  YAD_DISPATCH_DTYPE(dtype, yad::sppf_kernel, (const T*)in, B, W, C, ld_in, ci_off, (T*)out, ld_out, co_off);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_maxpool_h(const void* in, int32_t dtype, int64_t B, int32_t H, int32_t W, int32_t C, int32_t ld_in, int32_t ci_off,
                  int32_t radius, void* out, int32_t ld_out, int32_t co_off, yad_stream_t stream) {
  YAD_CHECK_ARG(in && out && in != out && H >= 1 && W >= 1 && C >= 1 && radius >= 0 && ld_in >= ci_off + C && ld_out >= co_off + C,
                "yad_maxpool_h: bad arguments (out of place only)");
  YAD_CHECK_ARG(dtype == YAD_F32 || dtype == YAD_BF16, "yad_maxpool_h: bad dtype %d", dtype);
  const int64_t n = B * H * W * C;
  if (n == 0) return YAD_OK;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  YAD_DISPATCH_DTYPE(dtype, yad::maxpool_h_kernel, (const T*)in, B, H, W, C, ld_in, ci_off, radius, (T*)out, ld_out, co_off);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_nchw_to_nhwc(const float* in, int64_t B, int32_t C, int32_t H, int32_t W, void* out, int32_t out_dtype, int32_t ld_out,
                     yad_stream_t stream) {
  YAD_CHECK_ARG(in && out && C >= 1 && H >= 1 && W >= 1 && ld_out >= C, "yad_nchw_to_nhwc: bad arguments");
  YAD_CHECK_ARG(out_dtype == YAD_F32 || out_dtype == YAD_BF16, "yad_nchw_to_nhwc: bad dtype %d", out_dtype);
  const int64_t n = B * (int64_t)H * W;
  if (n == 0) return YAD_OK;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  YAD_DISPATCH_DTYPE(out_dtype, yad::nchw_to_nhwc_kernel, in, B, C, (int64_t)H * W, (T*)out, ld_out);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_repvgg_merge(const void* a, const void* b, const void* x, const float* scale, const float* shift,
                     int32_t dtype, int64_t npix, int32_t C, int32_t ld_ab, int32_t ld_x, void* out, int32_t ld_out,
                     int32_t co_off, int32_t act, yad_stream_t stream) {
  YAD_CHECK_ARG(a && b && out && C >= 1 && ld_ab >= C && ld_out >= co_off + C, "yad_repvgg_merge: bad arguments");
  YAD_CHECK_ARG(x == nullptr || (scale && shift && ld_x >= C), "yad_repvgg_merge: identity branch needs scale/shift/ld_x");
  YAD_CHECK_ARG(dtype == YAD_F32 || dtype == YAD_BF16, "yad_repvgg_merge: bad dtype %d", dtype);
  const int64_t n = npix * C;
  if (n == 0) return YAD_OK;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  YAD_DISPATCH_DTYPE(dtype, yad::repvgg_merge_kernel, (const T*)a, (const T*)b, (const T*)x, scale, shift, npix, C, ld_ab,
                     ld_x, (T*)out, ld_out, co_off, act);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

}  // extern "C"

extern "C" int yad_resample_sinc(const void* x, int32_t x_is_i16, int64_t B, int64_t L, int32_t O, int32_t P, int32_t width,
                                 const float* kernel, float* out, int64_t Lout, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x && kernel && out, "yad_resample_sinc: null pointer");
  YAD_CHECK_ARG(B >= 0 && L >= 1 && O >= 1 && P >= 1 && width >= 0, "yad_resample_sinc: bad sizes");
  YAD_CHECK_ARG(Lout >= 0 && Lout <= ((int64_t)P * L + O - 1) / O, "yad_resample_sinc: Lout=%lld exceeds ceil(P * L / O)", (long long)Lout);
  if (B == 0 || Lout == 0) return YAD_OK;
  const int KW = 2 * width + O;
  const int threads = 256;
  const int64_t n = B * Lout;
  YAD_CHECK_ARG((n + threads - 1) / threads < ((int64_t)1 << 31), "yad_resample_sinc: too many samples");
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  if (x_is_i16)
    resample_sinc_kernel<int16_t><<<blocks, threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const int16_t*>(x), B, L, O, P, width, KW,
                                                                             kernel, out, Lout);
  else
    resample_sinc_kernel<float><<<blocks, threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(x), B, L, O, P, width, KW,
                                                                           kernel, out, Lout);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}
