// Train-mode building blocks of the network (fp32, NHWC rows x channels with explicit pitches), SURVEY section 8 row a17:
// the forward / backward that `TrainerPipeline.__feed` (pipeline/_trainer.py:94-108) drives through autograd.
//   * weight gradient of a convolution (F.conv2d backward w.r.t. weight / bias)
//   * BatchNorm2d in training mode, forward (batch statistics, running-stat update, fused ReLU / LeakyReLU(0.2)) and
//     backward (modules/_common.py:43-48, torchvision resnet.py:89-105)
//   * activation(a + b [+ c]) forward / backward (BasicBlock residual tail, RepVGG branch merge modules/_common.py:90-95)
//   * backward of the neck glue: H-mean, bilinear x2 / x0.5 along W, MaxPool(5,1,2) (modules/_common.py:173-209,248-252)
//   * Dropout (modules/_backbone.py:133,147) with a counter-based mask, backward of the anchor decode
//     (modules/_architecture.py:132-156) including the gradient of the three anchor parameters.
// The data gradient of a convolution is yad_conv_dgrad (conv_simt.cu).  These are correctness-first CUDA-core kernels:
// the inference path is the tuned one; see DESIGN.md for the measured train-step time.
#include "common.cuh"

namespace yad {

constexpr int TN_THREADS = 256;

__device__ __forceinline__ float act_grad(float y, int act) {   // d act / d pre-activation, from the OUTPUT y
  if (act == YAD_ACT_RELU) return y > 0.0f ? 1.0f : 0.0f;
  if (act == YAD_ACT_LRELU02) return y > 0.0f ? 1.0f : 0.2f;
  return 1.0f;
}

// ------------------------------------------------------------------------------------ conv weight gradient
// dW[kh][kw][ci][co] += sum_{b,ho,wo} X[b, ho*s+kh-p, wo*s+kw-p, ci] * dY[b,ho,wo,co] ; 64x64 tile per CTA, split over
// pixel chunks (fp32 atomics into dW).
constexpr int WG_BM = 64, WG_BN = 64, WG_BK = 16, WG_CHUNK = 1024;

__global__ void __launch_bounds__(TN_THREADS)
conv_wgrad_kernel(const yad_conv_desc d, int Ho, int Wo, const float* __restrict__ x, const float* __restrict__ dy,
                  float* __restrict__ dw, int n_ci_tiles) {
  __shared__ float As[WG_BK][WG_BM + 4];   // [pixel][ci]
  __shared__ float Bs[WG_BK][WG_BN + 4];   // [pixel][co]
  const int tap = blockIdx.y / n_ci_tiles, ci0 = (blockIdx.y % n_ci_tiles) * WG_BM;
  const int co0 = blockIdx.z * WG_BN;
  const int kh = tap / d.kw, kw = tap % d.kw;
  const int64_t M = (int64_t)d.B * Ho * Wo;
  const int64_t p0 = (int64_t)blockIdx.x * WG_CHUNK;
  const int64_t p1 = p0 + WG_CHUNK < M ? p0 + WG_CHUNK : M;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 4, lq = (tid & 15) * 4;     // load role: pixel row, channel quad
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int64_t pb = p0; pb < p1; pb += WG_BK) {
    const int64_t pm = pb + lrow;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (pm < p1) {
      const int wo = (int)(pm % Wo), ho = (int)((pm / Wo) % Ho), b = (int)(pm / ((int64_t)Wo * Ho));
      const int hi = ho * d.sh + kh - d.ph, wi = wo * d.sw + kw - d.pw;
      if (hi >= 0 && hi < d.H && wi >= 0 && wi < d.W) {
        const float* xp = x + (((int64_t)b * d.H + hi) * d.W + wi) * d.ld_in;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (ci0 + lq + e < d.Cin) av[e] = xp[ci0 + lq + e];
      }
      const float* yp = dy + pm * d.ld_out + d.co_off;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (co0 + lq + e < d.Cout) bv[e] = yp[co0 + lq + e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[lrow][lq + e] = av[e];
      Bs[lrow][lq + e] = bv[e];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WG_BK; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (ci >= d.Cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < d.Cout && acc[i][j] != 0.0f) atomicAdd(dw + ((int64_t)tap * d.Cin + ci) * d.Cout + co, acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------------------------ column reductions (rows x C)
// out[c] += sum_r f(r, c) in fp64.  MODE 0: (x, x^2) ; MODE 1: (g, g * xhat) with g = dy * act'(y), xhat = (x - mean) * invstd ;
// MODE 2: (x, -) plain column sum
template <int MODE>
__global__ void __launch_bounds__(TN_THREADS)
col_reduce_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ y, int ld_y, const float* __restrict__ dy,
                  int ld_dy, int64_t N, int C, const float* __restrict__ mean, const float* __restrict__ invstd, int act,
                  double* __restrict__ out0, double* __restrict__ out1) {
  __shared__ double s0[8][33], s1[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double a0 = 0.0, a1 = 0.0;
  if (c < C) {
    const float mu = MODE == 1 ? mean[c] : 0.0f, is = MODE == 1 ? invstd[c] : 0.0f;
    for (int64_t r = (int64_t)blockIdx.y * 8 + ry; r < N; r += (int64_t)gridDim.y * 8) {
      if (MODE == 0) {
        const double v = (double)x[r * ld_x + c];
        a0 += v;
        a1 += v * v;
      } else if (MODE == 1) {
        const float g = dy[r * ld_dy + c] * act_grad(y[r * ld_y + c], act);
        a0 += (double)g;
        a1 += (double)g * (double)((x[r * ld_x + c] - mu) * is);
      } else {
        a0 += (double)x[r * ld_x + c];
      }
    }
  }
  s0[ry][cx] = a0;
  s1[ry][cx] = a1;
  __syncthreads();
  if (ry == 0 && c < C) {
    for (int k = 1; k < 8; ++k) {
      a0 += s0[k][cx];
      a1 += s1[k][cx];
    }
    atomicAdd(out0 + c, a0);
    if (MODE != 2) atomicAdd(out1 + c, a1);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, int64_t N, int C, float eps,
                                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = sum[c] / (double)N;
  double var = sumsq[c] / (double)N - mu * mu;
  if (var < 0.0) var = 0.0;
  save_mean[c] = (float)mu;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) {
    const double unb = N > 1 ? var * (double)N / (double)(N - 1) : var;
    running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mu;
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

__global__ void __launch_bounds__(TN_THREADS)
bn_apply_kernel(const float* __restrict__ x, int ld_x, int64_t N, int C, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd, int act,
                float* __restrict__ y, int ld_y) {
  const int64_t n = N * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    const float v = (x[r * ld_x + c] - mean[c]) * invstd[c] * gamma[c] + beta[c];
    y[r * ld_y + c] = apply_act(v, act);
  }
}

// finalize + apply in one launch: every block derives mean / invstd of all C channels from the fp64 sums into shared memory
// (C <= 1024), block 0 also stores them and updates the running statistics
__global__ void __launch_bounds__(TN_THREADS)
bn_apply_fused_kernel(const float* __restrict__ x, int ld_x, int64_t N, int C, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const double* __restrict__ sums, float eps, float momentum,
                      float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ save_mean,
                      float* __restrict__ save_invstd, int act, float* __restrict__ y, int ld_y) {
  __shared__ float s_scale[1024], s_shift[1024];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double mu = sums[c] / (double)N;
    double var = sums[C + c] / (double)N - mu * mu;
    if (var < 0.0) var = 0.0;
    const float mean = (float)mu, invstd = (float)(1.0 / sqrt(var + (double)eps));
    // y = (x - mean) * invstd * gamma + beta, evaluated in the reference's order below
    s_scale[c] = invstd;
    s_shift[c] = mean;
    if (blockIdx.x == 0) {
      save_mean[c] = mean;
      save_invstd[c] = invstd;
      if (running_mean != nullptr) {
        const double unb = N > 1 ? var * (double)N / (double)(N - 1) : var;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unb;
      }
    }
  }
  __syncthreads();
  const int64_t n = N * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    const float v = (x[r * ld_x + c] - s_shift[c]) * s_scale[c] * gamma[c] + beta[c];
    y[r * ld_y + c] = apply_act(v, act);
  }
}

// dx = gamma * invstd * (g - sum_g / N - xhat * sum_gx / N), g = dy * act'(y)
__global__ void __launch_bounds__(TN_THREADS)
bn_bwd_apply_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ y, int ld_y, const float* __restrict__ dy,
                    int ld_dy, int64_t N, int C, const float* __restrict__ gamma, const float* __restrict__ mean,
                    const float* __restrict__ invstd, int act, const double* __restrict__ sum_g, const double* __restrict__ sum_gx,
                    float* __restrict__ dx, int ld_dx, int accumulate, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int64_t n = N * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    const float g = dy[r * ld_dy + c] * act_grad(y[r * ld_y + c], act);
    const float xh = (x[r * ld_x + c] - mean[c]) * invstd[c];
    const float sg = (float)(sum_g[c] / (double)N), sgx = (float)(sum_gx[c] / (double)N);
    const float v = gamma[c] * invstd[c] * (g - sg - xh * sgx);
    dx[r * ld_dx + c] = accumulate ? dx[r * ld_dx + c] + v : v;
    if (r == 0) {
      dgamma[c] += (float)sum_gx[c];
      dbeta[c] += (float)sum_g[c];
    }
  }
}

// ------------------------------------------------------------------------------------ elementwise
// y = act(a + b [+ c])
__global__ void __launch_bounds__(TN_THREADS)
add_act_kernel(const float* __restrict__ a, int ld_a, const float* __restrict__ b, int ld_b, const float* __restrict__ c, int ld_c,
               int64_t N, int C, int act, float* __restrict__ y, int ld_y) {
  const int64_t n = N * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C);
    const int64_t r = i / C;
    float v = a[r * ld_a + ch] + b[r * ld_b + ch];
    if (c != nullptr) v += c[r * ld_c + ch];
    y[r * ld_y + ch] = apply_act(v, act);
  }
}
// g = dy * act'(y) accumulated into up to three gradient tensors
__global__ void __launch_bounds__(TN_THREADS)
add_act_bwd_kernel(const float* __restrict__ y, int ld_y, const float* __restrict__ dy, int ld_dy, int64_t N, int C, int act,
                   float* __restrict__ da, int ld_a, float* __restrict__ db, int ld_b, float* __restrict__ dc, int ld_c) {
  const int64_t n = N * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C);
    const int64_t r = i / C;
    const float g = dy[r * ld_dy + ch] * act_grad(y[r * ld_y + ch], act);
    if (da != nullptr) da[r * ld_a + ch] += g;
    if (db != nullptr) db[r * ld_b + ch] += g;
    if (dc != nullptr) dc[r * ld_c + ch] += g;
  }
}

// dropout: y = x * keep / (1 - p), keep = hash(seed, index) >= p ; the same call with x = dy gives the backward
__device__ __forceinline__ uint32_t mix32(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}
__global__ void __launch_bounds__(TN_THREADS)
dropout_kernel(const float* __restrict__ x, int64_t n, float p, uint64_t seed, const uint64_t* __restrict__ seed_dev, int accumulate,
               float* __restrict__ y) {
  if (seed_dev != nullptr) seed += *seed_dev;      // CUDA-graph replays: the step counter lives on the device
  const float scale = 1.0f / (1.0f - p);
  const uint32_t thr = (uint32_t)((double)p * 4294967296.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = mix32(seed * 0x100000001B3ull + (uint64_t)i) >= thr ? x[i] * scale : 0.0f;
    y[i] = accumulate ? y[i] + v : v;
  }
}

// ------------------------------------------------------------------------------------ glue backward
// forward hmean: out[b,w,c] = mean_h in[b,h,w,c]  ->  din[b,h,w,c] += dout[b,w,c] / H
__global__ void __launch_bounds__(TN_THREADS)
hmean_bwd_kernel(const float* __restrict__ dout, int ld_o, int64_t B, int H, int W, int C, float* __restrict__ din, int ld_i) {
  const int64_t n = B * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t p = i / C;
    const int w = (int)(p % W);
    const int64_t b = p / ((int64_t)W * H);
    din[p * ld_i + c] += dout[(b * W + w) * ld_o + c] / (float)H;
  }
}
// forward resize_w (see glue.cu): up : out[2k] = .75 x[k] + .25 x[max(k-1,0)], out[2k+1] = .75 x[k] + .25 x[min(k+1,W-1)]
//                                  down: out[k] = .5 x[2k] + .5 x[2k+1].   One thread per INPUT element gathers its gradient.
__global__ void __launch_bounds__(TN_THREADS)
resize_w_bwd_kernel(const float* __restrict__ dout, int ld_o, int64_t B, int W, int C, int up, float* __restrict__ din, int ld_i) {
  const int Wo = up ? 2 * W : W / 2;
  const int64_t n = B * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t p = i / C;
    const int k = (int)(p % W);
    const int64_t b = p / W;
    const float* drow = dout + b * Wo * (int64_t)ld_o + c;
    float g = 0.0f;
    if (up) {
      g = 0.75f * (drow[(int64_t)(2 * k) * ld_o] + drow[(int64_t)(2 * k + 1) * ld_o]);
      // x[k] is the "previous" neighbour of out[2(k+1)] and the "next" neighbour of out[2(k-1)+1]; edges clamp onto themselves
      if (k + 1 < W) g += 0.25f * drow[(int64_t)(2 * k + 2) * ld_o];
      else g += 0.25f * drow[(int64_t)(2 * k + 1) * ld_o];
      if (k - 1 >= 0) g += 0.25f * drow[(int64_t)(2 * k - 1) * ld_o];
      else g += 0.25f * drow[(int64_t)(2 * k) * ld_o];
    } else {
      if (k / 2 < Wo) g = 0.5f * drow[(int64_t)(k / 2) * ld_o];
    }
    din[p * ld_i + c] += g;
  }
}
// MaxPool(k=5, s=1, p=2) along W: y[w] = max x[w-2..w+2]; backward routes dy[w] to the FIRST maximum of the window
__global__ void __launch_bounds__(TN_THREADS)
maxpool5_kernel(const float* __restrict__ x, int ld_x, int64_t B, int W, int C, float* __restrict__ y, int ld_y) {
  const int64_t n = B * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t p = i / C;
    const int w = (int)(p % W);
    const int64_t b = p / W;
    float m = -INFINITY;
    for (int j = max(w - 2, 0); j <= min(w + 2, W - 1); ++j) m = fmaxf(m, x[(b * W + j) * ld_x + c]);
    y[p * ld_y + c] = m;
  }
}
__global__ void __launch_bounds__(TN_THREADS)
maxpool5_bwd_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ dy, int ld_dy, int64_t B, int W, int C,
                    float* __restrict__ dx, int ld_dx) {
  const int64_t n = B * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t p = i / C;
    const int w = (int)(p % W);
    const int64_t b = p / W;
    float m = -INFINITY;
    int arg = w;
    for (int j = max(w - 2, 0); j <= min(w + 2, W - 1); ++j) {
      const float v = x[(b * W + j) * ld_x + c];
      if (v > m) {
        m = v;
        arg = j;
      }
    }
    const float g = dy[p * ld_dy + c];
    if (g != 0.0f) atomicAdd(dx + (b * W + arg) * ld_dx + c, g);
  }
}

// ------------------------------------------------------------------------------------ decode backward
// pred = [obj, cls.., centre, width] per (b, g, a);  head channel = a * E + e.
//   centre = clip(((sigmoid(tc) * 2 - 0.5) + g) * stride / scaler, 0, dur) ; width = clip((sigmoid(tw) * 2)^2 * anchor_s, 0, dur)
__global__ void __launch_bounds__(TN_THREADS)
decode_bwd_kernel(const float* __restrict__ head, int ld_h, const float* __restrict__ dpred, int64_t B, int G, int A, int E,
                  const float* __restrict__ anchors_s, float stride_over_scaler, float duration, float* __restrict__ dhead,
                  int ld_dh, float* __restrict__ danchor_s) {
  const int64_t n = B * G * A;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = (int)(i % A);
    const int64_t bg = i / A;
    const int g = (int)(bg % G);
    const float* h = head + bg * ld_h + a * E;
    const float* dp = dpred + i * E;
    float* dh = dhead + bg * ld_dh + a * E;
    for (int e = 0; e < E - 2; ++e) dh[e] = dp[e];
    const float sc = 1.0f / (1.0f + expf(-h[E - 2])), sw = 1.0f / (1.0f + expf(-h[E - 1]));
    const float cen = ((sc * 2.0f - 0.5f) + (float)g) * stride_over_scaler;
    const float wid = (sw * 2.0f) * (sw * 2.0f) * anchors_s[a];
    const float mc = (cen >= 0.0f && cen <= duration) ? 1.0f : 0.0f;   // clip passes the gradient at the bounds
    const float mw = (wid >= 0.0f && wid <= duration) ? 1.0f : 0.0f;
    dh[E - 2] = dp[E - 2] * mc * 2.0f * sc * (1.0f - sc) * stride_over_scaler;
    dh[E - 1] = dp[E - 1] * mw * 8.0f * sw * sw * (1.0f - sw) * anchors_s[a];
    const float da = dp[E - 1] * mw * (sw * 2.0f) * (sw * 2.0f);
    if (da != 0.0f) atomicAdd(danchor_s + a, da);
  }
}

// out (dense, row-major over sizes[4]) (+)= in[sum_k i_k * in_stride_k]: weight re-packing, NCHW <-> NHWC, gradient un-packing
struct Perm4 {
  int32_t n[4];
  int64_t s[4];
};
__global__ void __launch_bounds__(TN_THREADS)
permute4_kernel(const float* __restrict__ in, const Perm4 p, int accumulate, float* __restrict__ out) {
  const int64_t total = (int64_t)p.n[0] * p.n[1] * p.n[2] * p.n[3];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int i3 = (int)(r % p.n[3]);
    r /= p.n[3];
    const int i2 = (int)(r % p.n[2]);
    r /= p.n[2];
    const int i1 = (int)(r % p.n[1]);
    const int i0 = (int)(r / p.n[1]);
    const float v = in[i0 * p.s[0] + i1 * p.s[1] + i2 * p.s[2] + i3 * p.s[3]];
    out[i] = accumulate ? out[i] + v : v;
  }
}
// fp64 accumulator -> fp32 parameter gradient (+=)
__global__ void add_f64_to_f32_kernel(const double* __restrict__ a, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += (float)a[i];
}

// im2col of the stem conv1 (2 -> 64, 7x7, stride 2, pad 3; modules/_backbone.py:127,143): patches[b,ho,wo,k], k = (kh*7 + kw)*C + c
// for k < 49*C, zero above (K padded to a multiple of 32 so that the TF32 tensor-core kernels can treat conv1 as a 1x1 conv)
__global__ void __launch_bounds__(TN_THREADS)
stem_im2col_kernel(const float* __restrict__ x, int64_t B, int C, int H, int W, int Ho, int Wo, int K, float* __restrict__ out) {
  const int K4 = K >> 2;                       // one thread = 4 consecutive k (one 16-byte store)
  const int64_t n = B * Ho * Wo * K4;
  const int kmax = 49 * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k0 = (int)(i % K4) * 4;
    const int64_t p = i / K4;
    const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho);
    const int64_t b = p / ((int64_t)Wo * Ho);
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + e;
      v[e] = 0.0f;
      if (k < kmax) {
        const int c = k % C, t = k / C;
        const int kh = t / 7, kw = t - kh * 7;
        const int hi = 2 * ho + kh - 3, wi = 2 * wo + kw - 3;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) v[e] = __ldg(x + ((b * C + c) * H + hi) * W + wi);
      }
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

static inline unsigned ew_blocks(int64_t n) {
  int64_t b = (n + TN_THREADS - 1) / TN_THREADS;
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace yad

namespace yad {
// All OIHW master weights of the network -> the K-major operands of the TF32 convolutions, ONE launch per step instead of two
// strided-copy kernels per convolution (104 launches, 0.37 ms of a 7.4 ms train step): entry e covers elements
// [start_e, start_{e+1}) of the concatenated parameters; element (o, i, t) goes to wf[(o*KK + t)*cin_pad + i] (forward /
// weight-gradient operand) and wt[(i*KK + t)*coutk + o] (data-gradient operand).  Pad entries are never written (zeroed once).
struct PackEntry {
  const float* src;
  float* wf;
  float* wt;
  int32_t O, I, KK, cin_pad, coutk, pad_;
  int64_t start;
};

__global__ void pack_weights_kernel(const PackEntry* __restrict__ tab, int n, int64_t total) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  int lo = 0, hi = n - 1;
  while (lo < hi) {                      // last entry with start <= gid
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(&tab[mid].start) <= gid) lo = mid; else hi = mid - 1;
  }
  const PackEntry e = tab[lo];
  const uint32_t l = (uint32_t)(gid - e.start);
  const uint32_t t = l % (uint32_t)e.KK, oi = l / (uint32_t)e.KK;
  const uint32_t i = oi % (uint32_t)e.I, o = oi / (uint32_t)e.I;
  const float v = __ldg(e.src + l);
  e.wf[((int64_t)o * e.KK + t) * e.cin_pad + i] = v;
  e.wt[((int64_t)i * e.KK + t) * e.coutk + o] = v;
}
}  // namespace yad

extern "C" {

int yad_pack_weights_tf32(const void* table, int32_t n_entries, int64_t total_elems, yad_stream_t stream) {
  YAD_CHECK_ARG(table && n_entries >= 1 && total_elems >= 1, "yad_pack_weights_tf32: bad arguments");
  static_assert(sizeof(yad::PackEntry) == 56, "host table layout (train_engine.py) assumes 56-byte entries");
  const int threads = 256;
  yad::pack_weights_kernel<<<(unsigned)((total_elems + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const yad::PackEntry*>(table), n_entries, total_elems);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_conv_wgrad(const yad_conv_desc* d, const float* x, const float* dy, float* dw, double* dbias_ws, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(d && x && dy && dw, "yad_conv_wgrad: null pointer");
  const int Ho = (d->H + 2 * d->ph - d->kh) / d->sh + 1;
  const int Wo = (d->W + 2 * d->pw - d->kw) / d->sw + 1;
  YAD_CHECK_ARG(Ho >= 1 && Wo >= 1 && d->Cin >= 1 && d->Cout >= 1, "yad_conv_wgrad: bad descriptor");
  if (d->B == 0) return YAD_OK;
  const int64_t M = (int64_t)d->B * Ho * Wo;
  const int n_ci = (d->Cin + WG_BM - 1) / WG_BM, n_co = (d->Cout + WG_BN - 1) / WG_BN;
  dim3 grid((unsigned)((M + WG_CHUNK - 1) / WG_CHUNK), (unsigned)(d->kh * d->kw * n_ci), (unsigned)n_co);
  YAD_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "yad_conv_wgrad: too many tiles");
  conv_wgrad_kernel<<<grid, TN_THREADS, 0, (cudaStream_t)stream>>>(*d, Ho, Wo, x, dy, dw, n_ci);
  YAD_LAUNCH_CHECK();
  if (dbias_ws != nullptr) {   // column sums of dY (fp64, accumulated): the bias gradient
    dim3 g2((unsigned)((d->Cout + 31) / 32), (unsigned)(M >= 4096 ? 64 : 8));
    col_reduce_kernel<2><<<g2, TN_THREADS, 0, (cudaStream_t)stream>>>(dy + d->co_off, d->ld_out, nullptr, 0, nullptr, 0, M, d->Cout,
                                                                      nullptr, nullptr, 0, dbias_ws, nullptr);
    YAD_LAUNCH_CHECK();
  }
  return YAD_OK;
}

int yad_bn_train_fwd(const float* x, int32_t ld_x, int64_t N, int32_t C, const float* gamma, const float* beta, float eps,
                     float momentum, float* running_mean, float* running_var, int32_t act, float* y, int32_t ld_y,
                     float* save_mean, float* save_invstd, double* ws /* [2*C] */, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x && gamma && beta && y && save_mean && save_invstd && ws && N >= 1 && C >= 1, "yad_bn_train_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  YAD_CUDA(cudaMemsetAsync(ws, 0, 2 * (size_t)C * sizeof(double), st));
  dim3 g((unsigned)((C + 31) / 32), (unsigned)(N >= 4096 ? 64 : 8));
  col_reduce_kernel<0><<<g, TN_THREADS, 0, st>>>(x, ld_x, nullptr, 0, nullptr, 0, N, C, nullptr, nullptr, 0, ws, ws + C);
  YAD_LAUNCH_CHECK();
  bn_finalize_kernel<<<(unsigned)((C + 127) / 128), 128, 0, st>>>(ws, ws + C, N, C, eps, momentum, running_mean, running_var, save_mean,
                                                                 save_invstd);
  YAD_LAUNCH_CHECK();
  bn_apply_kernel<<<ew_blocks(N * C), TN_THREADS, 0, st>>>(x, ld_x, N, C, gamma, beta, save_mean, save_invstd, act, y, ld_y);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_bn_train_apply(const float* x, int32_t ld_x, int64_t N, int32_t C, const float* gamma, const float* beta, float eps,
                       float momentum, float* running_mean, float* running_var, int32_t act, float* y, int32_t ld_y,
                       float* save_mean, float* save_invstd, const double* sums, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x && gamma && beta && y && save_mean && save_invstd && sums && N >= 1 && C >= 1 && C <= 1024,
                "yad_bn_train_apply: bad arguments (C <= 1024)");
  bn_apply_fused_kernel<<<ew_blocks(N * C), TN_THREADS, 0, (cudaStream_t)stream>>>(x, ld_x, N, C, gamma, beta, sums, eps, momentum,
                                                                                  running_mean, running_var, save_mean, save_invstd, act,
                                                                                  y, ld_y);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_bn_train_bwd(const float* x, int32_t ld_x, const float* y, int32_t ld_y, const float* dy, int32_t ld_dy, int64_t N, int32_t C,
                     const float* gamma, const float* save_mean, const float* save_invstd, int32_t act, float* dx, int32_t ld_dx,
                     int32_t accumulate, float* dgamma, float* dbeta, double* ws /* [2*C] */, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x && y && dy && gamma && save_mean && save_invstd && dx && dgamma && dbeta && ws && N >= 1 && C >= 1,
                "yad_bn_train_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  YAD_CUDA(cudaMemsetAsync(ws, 0, 2 * (size_t)C * sizeof(double), st));
  dim3 g((unsigned)((C + 31) / 32), (unsigned)(N >= 4096 ? 64 : 8));
  col_reduce_kernel<1><<<g, TN_THREADS, 0, st>>>(x, ld_x, y, ld_y, dy, ld_dy, N, C, save_mean, save_invstd, act, ws, ws + C);
  YAD_LAUNCH_CHECK();
  bn_bwd_apply_kernel<<<ew_blocks(N * C), TN_THREADS, 0, st>>>(x, ld_x, y, ld_y, dy, ld_dy, N, C, gamma, save_mean, save_invstd, act, ws,
                                                              ws + C, dx, ld_dx, accumulate, dgamma, dbeta);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_add_act(const float* a, int32_t ld_a, const float* b, int32_t ld_b, const float* c, int32_t ld_c, int64_t N, int32_t C,
                int32_t act, float* y, int32_t ld_y, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(a && b && y && N >= 0 && C >= 1, "yad_add_act: bad arguments");
  if (N == 0) return YAD_OK;
  add_act_kernel<<<ew_blocks(N * C), TN_THREADS, 0, (cudaStream_t)stream>>>(a, ld_a, b, ld_b, c, ld_c, N, C, act, y, ld_y);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_add_act_bwd(const float* y, int32_t ld_y, const float* dy, int32_t ld_dy, int64_t N, int32_t C, int32_t act, float* da,
                    int32_t ld_a, float* db, int32_t ld_b, float* dc, int32_t ld_c, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(y && dy && N >= 0 && C >= 1, "yad_add_act_bwd: bad arguments");
  if (N == 0) return YAD_OK;
  add_act_bwd_kernel<<<ew_blocks(N * C), TN_THREADS, 0, (cudaStream_t)stream>>>(y, ld_y, dy, ld_dy, N, C, act, da, ld_a, db, ld_b, dc,
                                                                               ld_c);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_dropout(const float* x, int64_t n, float p, uint64_t seed, int32_t accumulate, float* y, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x && y && n >= 0 && p >= 0.0f && p < 1.0f, "yad_dropout: bad arguments");
  if (n == 0) return YAD_OK;
  dropout_kernel<<<ew_blocks(n), TN_THREADS, 0, (cudaStream_t)stream>>>(x, n, p, seed, nullptr, accumulate, y);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_dropout_dev(const float* x, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev, int32_t accumulate, float* y,
                    yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x && y && seed_dev && n >= 0 && p >= 0.0f && p < 1.0f, "yad_dropout_dev: bad arguments");
  if (n == 0) return YAD_OK;
  dropout_kernel<<<ew_blocks(n), TN_THREADS, 0, (cudaStream_t)stream>>>(x, n, p, seed, seed_dev, accumulate, y);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_hmean_bwd(const float* dout, int32_t ld_o, int64_t B, int32_t H, int32_t W, int32_t C, float* din, int32_t ld_i,
                  yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(dout && din && H >= 1 && W >= 1 && C >= 1, "yad_hmean_bwd: bad arguments");
  if (B == 0) return YAD_OK;
  hmean_bwd_kernel<<<ew_blocks(B * H * W * C), TN_THREADS, 0, (cudaStream_t)stream>>>(dout, ld_o, B, H, W, C, din, ld_i);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_resize_w_bwd(const float* dout, int32_t ld_o, int64_t B, int32_t W, int32_t C, int32_t up, float* din, int32_t ld_i,
                     yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(dout && din && W >= 1 && C >= 1, "yad_resize_w_bwd: bad arguments");
  if (B == 0) return YAD_OK;
  resize_w_bwd_kernel<<<ew_blocks(B * W * C), TN_THREADS, 0, (cudaStream_t)stream>>>(dout, ld_o, B, W, C, up, din, ld_i);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_maxpool5_w(const float* x, int32_t ld_x, int64_t B, int32_t W, int32_t C, float* y, int32_t ld_y, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x && y && W >= 1 && C >= 1, "yad_maxpool5_w: bad arguments");
  if (B == 0) return YAD_OK;
  maxpool5_kernel<<<ew_blocks(B * W * C), TN_THREADS, 0, (cudaStream_t)stream>>>(x, ld_x, B, W, C, y, ld_y);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_maxpool5_w_bwd(const float* x, int32_t ld_x, const float* dy, int32_t ld_dy, int64_t B, int32_t W, int32_t C, float* dx,
                       int32_t ld_dx, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x && dy && dx && W >= 1 && C >= 1, "yad_maxpool5_w_bwd: bad arguments");
  if (B == 0) return YAD_OK;
  maxpool5_bwd_kernel<<<ew_blocks(B * W * C), TN_THREADS, 0, (cudaStream_t)stream>>>(x, ld_x, dy, ld_dy, B, W, C, dx, ld_dx);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_decode_bwd(const float* head, int32_t ld_h, const float* dpred, int64_t B, int32_t G, int32_t A, int32_t nc,
                   const float* anchors_s, float stride_over_scaler, float duration, float* dhead, int32_t ld_dh, float* danchor_s,
                   yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(head && dpred && anchors_s && dhead && danchor_s && G >= 1 && A >= 1 && nc >= 1, "yad_decode_bwd: bad arguments");
  if (B == 0) return YAD_OK;
  decode_bwd_kernel<<<ew_blocks(B * G * A), TN_THREADS, 0, (cudaStream_t)stream>>>(head, ld_h, dpred, B, G, A, 3 + nc, anchors_s,
                                                                                  stride_over_scaler, duration, dhead, ld_dh, danchor_s);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_permute4(const float* in, const int64_t* in_strides, float* out, const int32_t* sizes, int32_t accumulate,
                 yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(in && in_strides && out && sizes, "yad_permute4: null pointer");
  Perm4 p;
  int64_t total = 1;
  for (int k = 0; k < 4; ++k) {
    YAD_CHECK_ARG(sizes[k] >= 1, "yad_permute4: bad size");
    p.n[k] = sizes[k];
    p.s[k] = in_strides[k];
    total *= sizes[k];
  }
  permute4_kernel<<<ew_blocks(total), TN_THREADS, 0, (cudaStream_t)stream>>>(in, p, accumulate, out);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_stem_im2col(const float* x_nchw, int64_t B, int32_t C, int32_t H, int32_t W, int32_t K, float* patches, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x_nchw && patches && C >= 1 && H >= 1 && W >= 1 && K >= 49 * C && K % 4 == 0, "yad_stem_im2col: bad arguments");
  if (B == 0) return YAD_OK;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  stem_im2col_kernel<<<ew_blocks(B * Ho * Wo * (K / 4)), TN_THREADS, 0, (cudaStream_t)stream>>>(x_nchw, B, C, H, W, Ho, Wo, K, patches);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_colsum_f64(const float* x, int32_t ld, int64_t N, int32_t C, double* ws, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(x && ws && N >= 0 && C >= 1 && ld >= C, "yad_colsum_f64: bad arguments");
  if (N == 0) return YAD_OK;
  dim3 g((unsigned)((C + 31) / 32), (unsigned)(N >= 4096 ? 64 : 8));
  col_reduce_kernel<2><<<g, TN_THREADS, 0, (cudaStream_t)stream>>>(x, ld, nullptr, 0, nullptr, 0, N, C, nullptr, nullptr, 0, ws, nullptr);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_add_f64_to_f32(const double* a, int32_t n, float* out, yad_stream_t stream) {
  YAD_CHECK_ARG(a && out && n >= 0, "yad_add_f64_to_f32: bad arguments");
  if (n == 0) return YAD_OK;
  yad::add_f64_to_f32_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, n, out);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

}  // extern "C"
