// Training-side kernels of the hot path (config 5):
//  * anchor matching  - dataset.py:286-365 (AudioDataset.build_target_by_scale): integer-exact.
//  * fused Adam + EMA - torch.optim.Adam with L2 weight decay (train.py:83-90, config.yaml:75-80)
//    and EMAParamsSmoothener.update (smoothener/_ema.py:20-26) over one flat fp32 arena, one launch
//    instead of the reference's ~10 launches per parameter tensor.
#include "common.cuh"
#include <math.h>

namespace yad {

constexpr int BT_THREADS = 1024;

// block-wide exclusive scan of a 0/1 flag; returns this thread's exclusive prefix and the block total
__device__ __forceinline__ int block_excl_scan(int flag, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int excl_in_warp = __popc(bal & ((1u << lane) - 1u));
  __syncthreads();
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  int off = 0, tot = 0;
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) {
    const int c = s_warp[w];
    if (w < warp) off += c;
    tot += c;
  }
  *total = tot;
  return off + excl_in_warp;
}

__global__ void __launch_bounds__(BT_THREADS)
build_targets_kernel(const float* __restrict__ targets, int T, const float* __restrict__ anchors, int A, int G,
                     float anchor_t, float duration, float edge_t, int64_t* __restrict__ batch_idx,
                     int64_t* __restrict__ grid_idx, int64_t* __restrict__ anchor_idx, int64_t* __restrict__ classes,
                     float* __restrict__ cw, int32_t* __restrict__ n_out) {
  __shared__ int s_warp[32];
  const int n = A * T;
  const float Gf = (float)G;

  auto emit = [&](int row, int j, int a, float off) {
    const float c = targets[j * 4 + 2], w = targets[j * 4 + 3];
    batch_idx[row] = (int64_t)targets[j * 4 + 0];
    classes[row] = (int64_t)targets[j * 4 + 1];
    anchor_idx[row] = a;
    cw[row * 2 + 0] = c;
    cw[row * 2 + 1] = w;
    const float gc = __fadd_rn(__fmul_rn(__fdiv_rn(c, duration), Gf), off);
    long long gi = (long long)gc;  // .long(): truncation toward zero
    gi = gi < 0 ? 0 : (gi > G - 1 ? G - 1 : gi);
    grid_idx[row] = gi;
  };
  auto is_base = [&](int i) -> bool {
    const int a = i / T, j = i - a * T;
    const float r = __fdiv_rn(targets[j * 4 + 3], anchors[a]);
    const float m = fmaxf(r, __fdiv_rn(1.0f, r));
    return m < anchor_t;   // NaN compares false, like torch
  };
  auto gc_of = [&](int j) -> float { return __fmul_rn(__fdiv_rn(targets[j * 4 + 2], duration), Gf); };

  // pass 0: base matches (anchor-major).  passes 1 / 2: left / right neighbour copies, in base order.
  int out_base = 0;
  for (int pass = 0; pass < 3; ++pass) {
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
      const int i = i0 + threadIdx.x;
      int flag = 0, a = 0, j = 0;
      float off = 0.0f;
      if (i < n && is_base(i)) {
        a = i / T;
        j = i - a * T;
        const float gc = gc_of(j);
        if (pass == 0) {
          flag = 1;
        } else if (pass == 1) {
          flag = (fmodf(gc, 1.0f) < edge_t) && (gc > 1.0f);
          off = -edge_t;
        } else {
          const float gi = __fsub_rn(Gf, gc);
          flag = (fmodf(gi, 1.0f) < edge_t) && (gi > 1.0f);
          off = edge_t;
        }
      }
      int total;
      const int pos = block_excl_scan(flag, s_warp, &total);
      if (flag) emit(out_base + pos, j, a, off);
      out_base += total;
    }
  }
  if (threadIdx.x == 0) n_out[0] = out_base;
}

__global__ void adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, float* __restrict__ ema, int64_t n, float step_size, float beta1,
                                float beta2, float eps, float wd, float inv_bc2_sqrt, float ema_m) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i]);                       // grad = grad + weight_decay * param
    const float mi = fmaf(1.0f - beta1, gi - m[i], m[i]);      // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(1.0f - beta2, gi * gi, beta2 * v[i]);
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    const float pn = pi - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
    p[i] = pn;
    if (ema != nullptr) ema[i] = fmaf(pn, ema_m, ema[i] * (1.0f - ema_m));
  }
}


// ------------------------------------------------------------------------------------ batch builder
// AudioDataset.__getitem__ tail + collate_fn (dataset.py:132-155,276-283) on the device: ragged clips (C_b channels x n_b
// samples each, packed back to back, fp32 or 16-bit PCM) -> [B, 1, L] fp32, channel mean (audio_tensor.mean(dim=0)) and zero
// padding up to L.  One grid row per clip.
template <typename T>
__global__ void __launch_bounds__(256)
collate_clips_kernel(const T* __restrict__ packed, const int64_t* __restrict__ offset, const int32_t* __restrict__ n_samples,
                     const int32_t* __restrict__ n_channels, int64_t L, float scale, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int64_t n = n_samples[b];
  const int c = n_channels[b];
  const T* src = packed + offset[b];
  float* dst = out + (int64_t)b * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.0f;
    if (i < n) {
      if (c == 1) {
        v = (float)src[i] * scale;
      } else {                       // torch's mean: sum in channel order, then divide
        float acc = 0.0f;
        for (int k = 0; k < c; ++k) acc += (float)src[(int64_t)k * n + i] * scale;
        v = acc / (float)c;
      }
    }
    dst[i] = v;
  }
}

}  // namespace yad

extern "C" {

int yad_build_targets(const float* targets, int32_t T, const float* anchors, int32_t A, int32_t G, float anchor_t,
                      float duration, float edge_t, int64_t* batch_idx, int64_t* grid_idx, int64_t* anchor_idx,
                      int64_t* classes, float* cw, int32_t* n_out, yad_stream_t stream) {
  YAD_CHECK_ARG(targets && anchors && batch_idx && grid_idx && anchor_idx && classes && cw && n_out,
                "yad_build_targets: null pointer");
  YAD_CHECK_ARG(T >= 0 && A >= 1 && G >= 1, "yad_build_targets: bad T/A/G");
  yad::build_targets_kernel<<<1, yad::BT_THREADS, 0, (cudaStream_t)stream>>>(
      targets, T, anchors, A, G, anchor_t, duration, edge_t, batch_idx, grid_idx, anchor_idx, classes, cw, n_out);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_adam_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int64_t n,
                      float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                      float ema_momentum, yad_stream_t stream) {
  YAD_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "yad_adam_ema_step: bad arguments");
  if (n == 0) return YAD_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  const int threads = 256;
  int64_t blocks = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)(yad::sm_count() > 0 ? yad::sm_count() : 148) * 16;
  if (blocks > cap) blocks = cap;
  yad::adam_ema_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, ema, n, step_size, beta1, beta2, eps, weight_decay, inv_bc2_sqrt, ema_momentum);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

int yad_collate_clips(const void* packed, int32_t dtype_i16, const int64_t* offset, const int32_t* n_samples,
                      const int32_t* n_channels, int64_t B, int64_t L, float* out, yad_stream_t stream) {
  YAD_CHECK_ARG(packed && offset && n_samples && n_channels && out && B >= 0 && B <= 65535 && L >= 1, "yad_collate_clips: bad arguments");
  if (B == 0) return YAD_OK;
  const int threads = 256;
  int64_t bx = (L + threads - 1) / threads;
  if (bx > 512) bx = 512;
  dim3 grid((unsigned)bx, (unsigned)B);
  if (dtype_i16)
    yad::collate_clips_kernel<int16_t><<<grid, threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const int16_t*>(packed), offset,
                                                                                 n_samples, n_channels, L, 1.0f / 32768.0f, out);
  else
    yad::collate_clips_kernel<float><<<grid, threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(packed), offset,
                                                                               n_samples, n_channels, L, 1.0f, out);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}

}  // extern "C"
