// YOLO-style detection loss of one scale: forward value, metrics sums and the gradient with respect to the decoded
// predictions.  Replaces AudioDetectionLoss.loss_fn + compute_ciou (modules/_loss.py:115-228).  Default train_config
// (multi_label BCE class loss with label smoothing, BCEWithLogits objectness); the other two branches of the constructor
// (:74-81) are selected by yad_loss_scale_ex: cls_mode 1 = CrossEntropyLoss(weight=class_weights) over the class-valid
// matches (sum_m w[c_m] (logsumexp(x_m) - x_m[c_m]) / sum_m w[c_m]), focal_gamma > 0 = FocalLoss(with_logits=True) objectness
// (mean of alpha (1 - exp(-bce))^gamma bce, :9-37).
//   ciou_m   = CIoU(pred[b,g,a, -2:], target cw_m)          1-D segments dressed as boxes of height 10 (:193-228)
//   box      = mean_m (1 - ciou_m)
//   t_conf   = zeros[B,G,A]; t_conf[b,g,a] = ciou_m          duplicates: the LAST match wins (index_put_ on CPU; Q12)
//   conf     = mean over all B*G*A of BCEWithLogits(pred[...,0], t_conf)
//   cls      = mean over (matches with class != ignore_index) x nc of BCEWithLogits(pred[..., 1:1+nc], smoothed one-hot)
// The caller combines the three scales (conf weights 4/2/1, box_w / conf_w / class_w) and replaces NaN terms by 0
// (handle_nan, :178).  Three launches per scale: matches (ciou, ownership of cells), cells (objectness + its
// gradient, zeroing of the other gradient columns), matches again (box / class terms and their gradients).
// All sums go to fp64 accumulators so the result does not depend on the atomic order beyond 1e-16.
#include "common.cuh"

namespace yad {

constexpr int LS_THREADS = 256;
// accumulator slots (double)
enum { ACC_BOX = 0, ACC_CIOU, ACC_CONF, ACC_CLS, ACC_POS, ACC_NEG, ACC_NNEG, ACC_NVALID, ACC_WSUM, ACC_N };

struct CiouOut {
  float ciou;       // clipped at 0
  float d_pc, d_pw; // d ciou / d (pred centre, pred width), 0 where the clip is active
};

// forward + analytic gradient of compute_ciou (a = v / ((1 + e) - iou) + v is detached, as in the reference)
__device__ __forceinline__ CiouOut ciou_fwd_bwd(float pc, float pw, float tc, float tw) {
  const float e = 1e-8f, hh = 10.0f;
  const float px1 = pc - pw * 0.5f, px2 = pc + pw * 0.5f, tx1 = tc - tw * 0.5f, tx2 = tc + tw * 0.5f;
  const float mn2 = fminf(px2, tx2), mx1 = fmaxf(px1, tx1);
  const float iw_raw = mn2 - mx1;
  const float iw = fmaxf(iw_raw, 0.0f);
  const float inter = iw * hh;
  const float uni = pw * hh + tw * hh - inter;
  const float iou = inter / (uni + e);
  const float mx2 = fmaxf(px2, tx2), mn1 = fminf(px1, tx1);
  const float cwid = mx2 - mn1;
  const float c2 = cwid * cwid + hh * hh + e;
  const float k = 4.0f / (3.14159265358979323846f * 3.14159265358979323846f);
  const float dat = atanf(tw / hh) - atanf(pw / hh);
  const float v = k * dat * dat;
  const float dc = pc - tc;
  const float rho2 = dc * dc;
  const float a = v / ((1.0f + e) - iou) + v;
  const float raw = iou - (rho2 / c2 + a * v);
  CiouOut o;
  o.ciou = fmaxf(raw, 0.0f);
  // selectors of min / max (torch.minimum / maximum split the gradient evenly on ties)
  const float s_mn2 = px2 < tx2 ? 1.0f : (px2 == tx2 ? 0.5f : 0.0f);   // d min(px2,tx2) / d px2
  const float s_mx1 = px1 > tx1 ? 1.0f : (px1 == tx1 ? 0.5f : 0.0f);   // d max(px1,tx1) / d px1
  const float s_mx2 = px2 > tx2 ? 1.0f : (px2 == tx2 ? 0.5f : 0.0f);
  const float s_mn1 = px1 < tx1 ? 1.0f : (px1 == tx1 ? 0.5f : 0.0f);
  const float clipm = iw_raw >= 0.0f ? 1.0f : 0.0f;                    // clip(min=0) passes the gradient at equality
  const float diw_pc = clipm * (s_mn2 - s_mx1), diw_pw = clipm * 0.5f * (s_mn2 + s_mx1);
  const float dint_pc = hh * diw_pc, dint_pw = hh * diw_pw;
  const float duni_pc = -dint_pc, duni_pw = hh - dint_pw;
  const float den = (uni + e) * (uni + e);
  const float diou_pc = (dint_pc * (uni + e) - inter * duni_pc) / den;
  const float diou_pw = (dint_pw * (uni + e) - inter * duni_pw) / den;
  const float dcw_pc = s_mx2 - s_mn1, dcw_pw = 0.5f * (s_mx2 + s_mn1);
  const float dc2_pc = 2.0f * cwid * dcw_pc, dc2_pw = 2.0f * cwid * dcw_pw;
  const float dr_pc = (2.0f * dc * c2 - rho2 * dc2_pc) / (c2 * c2);
  const float dr_pw = (-rho2 * dc2_pw) / (c2 * c2);
  const float r = pw / hh;
  const float dv_pw = k * 2.0f * dat * (-1.0f / (1.0f + r * r)) / hh;
  const float live = raw >= 0.0f ? 1.0f : 0.0f;
  o.d_pc = live * (diou_pc - dr_pc);
  o.d_pw = live * (diou_pw - (dr_pw + a * dv_pw));
  return o;
}

__device__ __forceinline__ float bce_logits(float x, float t) { return fmaxf(x, 0.0f) - x * t + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void block_add(double v, double* dst, double* scratch) {
  v = warp_sum_d(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += scratch[w];
    if (s != 0.0) atomicAdd(dst, s);
  }
}

// pass 1 over matches: ciou, cell ownership (largest match index wins = "last writer"), count of class-valid matches
__global__ void __launch_bounds__(LS_THREADS)
loss_match_kernel(const float* __restrict__ pred, int G, int A, int E, const int64_t* __restrict__ bi,
                  const int64_t* __restrict__ gi, const int64_t* __restrict__ ai, const int64_t* __restrict__ cl,
                  const float* __restrict__ cw, int M, int64_t ignore_index, int nc, const float* __restrict__ class_weights,
                  int32_t* __restrict__ owner, float* __restrict__ ciou_ws, double* __restrict__ acc) {
  __shared__ double scratch[LS_THREADS / 32];
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  double nvalid = 0.0, wsum = 0.0;
  if (m < M) {
    const int64_t cell = (bi[m] * G + gi[m]) * A + ai[m];
    const float* p = pred + cell * E;
    const CiouOut o = ciou_fwd_bwd(p[E - 2], p[E - 1], cw[2 * m], cw[2 * m + 1]);
    ciou_ws[m] = o.ciou;
    atomicMax(owner + cell, m);
    nvalid = cl[m] != ignore_index ? 1.0 : 0.0;
    if (cl[m] != ignore_index) wsum = (class_weights != nullptr && cl[m] >= 0 && cl[m] < nc) ? (double)class_weights[cl[m]] : 1.0;
  }
  block_add(nvalid, acc + ACC_NVALID, scratch);
  block_add(wsum, acc + ACC_WSUM, scratch);       // CrossEntropyLoss(weight) divides by the sum of the targets' weights
}

// pass 2 over cells: objectness BCE vs t_conf, its gradient; zero the other gradient columns
__global__ void __launch_bounds__(LS_THREADS)
loss_cell_kernel(const float* __restrict__ pred, int64_t N, int E, const int32_t* __restrict__ owner,
                 const float* __restrict__ ciou_ws, float conf_scale /* conf_w * scale weight */, float focal_alpha,
                 float focal_gamma /* <= 0: plain BCE */, float* __restrict__ grad, double* __restrict__ acc) {
  __shared__ double scratch[LS_THREADS / 32];
  double s_conf = 0.0, s_neg = 0.0, n_neg = 0.0;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < N; c += (int64_t)gridDim.x * blockDim.x) {
    const int o = owner[c];
    const float t = o >= 0 ? ciou_ws[o] : 0.0f;
    const float x = pred[c * E];
    const float sg = sigmoidf_(x);
    const float bce = bce_logits(x, t);
    float dl = 1.0f;                          // d loss_cell / d bce
    if (focal_gamma > 0.0f) {                 // FocalLoss: f = alpha u^gamma bce with u = 1 - exp(-bce)
      const float pt = expf(-bce), u = 1.0f - pt;
      const float ug = u > 0.0f ? powf(u, focal_gamma) : 0.0f;
      s_conf += (double)(focal_alpha * ug * bce);
      dl = u > 0.0f ? focal_alpha * (focal_gamma * (ug / u) * pt * bce + ug) : 0.0f;
    } else {
      s_conf += (double)bce;
    }
    if (t == 0.0f) {
      s_neg += (double)sg;
      n_neg += 1.0;
    }
    float* g = grad + c * E;
    g[0] = conf_scale * dl * (sg - t) / (float)N;
    for (int j = 1; j < E; ++j) g[j] = 0.0f;
  }
  block_add(s_conf, acc + ACC_CONF, scratch);
  block_add(s_neg, acc + ACC_NEG, scratch);
  block_add(n_neg, acc + ACC_NNEG, scratch);
}

// pass 3 over matches: box and class terms + gradients (scatter-add: duplicates accumulate, as index backward does);
// confusion counts for the accuracy / precision / recall / f1 metrics
__global__ void __launch_bounds__(LS_THREADS)
loss_match_grad_kernel(const float* __restrict__ pred, int G, int A, int E, int nc, const int64_t* __restrict__ bi,
                       const int64_t* __restrict__ gi, const int64_t* __restrict__ ai, const int64_t* __restrict__ cl,
                       const float* __restrict__ cw, int M, int64_t ignore_index, float box_w, float class_w, float cn, int cls_mode,
                       const float* __restrict__ class_weights, float* __restrict__ grad,
                       int32_t* __restrict__ confusion /* [nc][nc] target x predicted */, double* __restrict__ acc) {
  __shared__ double scratch[LS_THREADS / 32];
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  double s_box = 0.0, s_ciou = 0.0, s_cls = 0.0, s_pos = 0.0;
  if (m < M) {
    const int64_t cell = (bi[m] * G + gi[m]) * A + ai[m];
    const float* p = pred + cell * E;
    float* g = grad + cell * E;
    const CiouOut o = ciou_fwd_bwd(p[E - 2], p[E - 1], cw[2 * m], cw[2 * m + 1]);
    s_box = 1.0 - (double)o.ciou;
    s_ciou = (double)o.ciou;
    s_pos = (double)sigmoidf_(p[0]);
    const float gb = -box_w / (float)M;
    if (o.d_pc != 0.0f) atomicAdd(g + E - 2, gb * o.d_pc);
    if (o.d_pw != 0.0f) atomicAdd(g + E - 1, gb * o.d_pw);
    const int64_t c = cl[m];
    if (c != ignore_index) {
      const double nvalid = acc[ACC_NVALID];          // complete: written by the first pass (previous launch)
      int best = 0;
      float bestv = p[1];
      for (int j = 1; j < nc; ++j) {
        if (p[1 + j] > bestv) {
          bestv = p[1 + j];
          best = j;
        }
      }
      if (cls_mode == 0) {                             // BCEWithLogits against the smoothed one-hot row, mean over valid x nc
        const float gs = class_w / (float)(nvalid * nc);
        for (int j = 0; j < nc; ++j) {
          const float x = p[1 + j];
          const float t = (j == (int)c) ? 1.0f - cn : cn;
          s_cls += (double)bce_logits(x, t);
          atomicAdd(g + 1 + j, gs * (sigmoidf_(x) - t));
        }
      } else if (c >= 0 && c < nc) {                   // CrossEntropyLoss: w_c (logsumexp(x) - x_c) / sum of the targets' weights
        const float wc = class_weights != nullptr ? class_weights[c] : 1.0f;
        float se = 0.0f;
        for (int j = 0; j < nc; ++j) se += expf(p[1 + j] - bestv);
        const float lse = bestv + logf(se);
        s_cls += (double)(wc * (lse - p[1 + c]));
        const float gs = class_w * wc / (float)acc[ACC_WSUM];
        for (int j = 0; j < nc; ++j) atomicAdd(g + 1 + j, gs * (expf(p[1 + j] - lse) - (j == (int)c ? 1.0f : 0.0f)));
      }
      if (c >= 0 && c < nc) atomicAdd(confusion + (int)c * nc + best, 1);
    }
  }
  block_add(s_box, acc + ACC_BOX, scratch);
  block_add(s_ciou, acc + ACC_CIOU, scratch);
  block_add(s_cls, acc + ACC_CLS, scratch);
  block_add(s_pos, acc + ACC_POS, scratch);
}

}  // namespace yad

extern "C" int yad_loss_scale_ex(const float* pred, int64_t B, int32_t G, int32_t A, int32_t nc, const int64_t* bi, const int64_t* gi,
                                 const int64_t* ai, const int64_t* cl, const float* cw, int32_t M, float box_w, float conf_scale,
                                 float class_w, float label_smoothing, int64_t ignore_index, int32_t cls_mode,
                                 const float* class_weights, float focal_alpha, float focal_gamma, int32_t* owner_ws, float* ciou_ws,
                                 int32_t* confusion, double* acc, float* grad, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(cls_mode == 0 || cls_mode == 1, "yad_loss_scale_ex: cls_mode must be 0 (multi-label BCE) or 1 (cross entropy)");
  YAD_CHECK_ARG(class_weights == nullptr || cls_mode == 1, "yad_loss_scale_ex: class weights belong to the cross-entropy mode");
  YAD_CHECK_ARG(pred && owner_ws && acc && grad && confusion, "yad_loss_scale: null pointer");
  YAD_CHECK_ARG(B >= 1 && G >= 1 && A >= 1 && nc >= 1 && M >= 0, "yad_loss_scale: bad shape");
  YAD_CHECK_ARG(M == 0 || (bi && gi && ai && cl && cw && ciou_ws), "yad_loss_scale: null match arrays");
  const int E = 3 + nc;
  const int64_t N = B * G * A;
  YAD_CHECK_ARG(N < ((int64_t)1 << 31), "yad_loss_scale: too many cells");
  cudaStream_t st = (cudaStream_t)stream;
  YAD_CUDA(cudaMemsetAsync(owner_ws, 0xFF, (size_t)N * sizeof(int32_t), st));       // -1 = unowned
  YAD_CUDA(cudaMemsetAsync(acc, 0, ACC_N * sizeof(double), st));
  YAD_CUDA(cudaMemsetAsync(confusion, 0, (size_t)nc * nc * sizeof(int32_t), st));
  const unsigned mb = (unsigned)((M + LS_THREADS - 1) / LS_THREADS);
  if (M > 0) {
    loss_match_kernel<<<mb, LS_THREADS, 0, st>>>(pred, G, A, E, bi, gi, ai, cl, cw, M, ignore_index, nc, class_weights, owner_ws, ciou_ws, acc);
    YAD_LAUNCH_CHECK();
  }
  int64_t cb = (N + LS_THREADS - 1) / LS_THREADS;
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  if (cb > cap) cb = cap;
  loss_cell_kernel<<<(unsigned)cb, LS_THREADS, 0, st>>>(pred, N, E, owner_ws, ciou_ws, conf_scale, focal_alpha, focal_gamma, grad, acc);
  YAD_LAUNCH_CHECK();
  if (M > 0) {
    loss_match_grad_kernel<<<mb, LS_THREADS, 0, st>>>(pred, G, A, E, nc, bi, gi, ai, cl, cw, M, ignore_index, box_w, class_w,
                                                     0.5f * label_smoothing, cls_mode, class_weights, grad, confusion, acc);
    YAD_LAUNCH_CHECK();
  }
  return YAD_OK;
}

extern "C" int yad_loss_scale(const float* pred, int64_t B, int32_t G, int32_t A, int32_t nc, const int64_t* bi, const int64_t* gi,
                              const int64_t* ai, const int64_t* cl, const float* cw, int32_t M, float box_w, float conf_scale,
                              float class_w, float label_smoothing, int64_t ignore_index, int32_t* owner_ws, float* ciou_ws,
                              int32_t* confusion, double* acc, float* grad, yad_stream_t stream) {
  return yad_loss_scale_ex(pred, B, G, A, nc, bi, gi, ai, cl, cw, M, box_w, conf_scale, class_w, label_smoothing, ignore_index, 0, nullptr,
                           0.0f, 0.0f, owner_ws, ciou_ws, confusion, acc, grad, stream);
}
