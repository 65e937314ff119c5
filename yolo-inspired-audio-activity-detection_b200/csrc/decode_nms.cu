// Anchor decode + 1-D segment NMS + post-processing (warp-level primitives, one CTA per clip).
// Reference call sites: modules/_architecture.py:113-156 (get_scale_pred), inference.py:42-110
// (process_model_outputs), torchvision ops/boxes.py:48-120 -> torchvision::nms.
//
// All IoU / box arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) so
// that keep-sets are bit-exact against the fp32 CPU implementation given identical inputs.
#include "common.cuh"

namespace yad {

__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// ------------------------------------------------------------------------------------ decode
struct DecodeParams {
  const void* head[4];
  int32_t G[4], ld[4], stride[4], row0[4];
  float anchors[4 * 8];
  const float* anchors_dev;     // optional [n_scales * A] on the device (train mode: the anchors are parameters)
  int32_t n_scales, A, nc, rows_total;
  float center_scaler, duration;
};

template <typename T>
__global__ void decode_kernel(DecodeParams p, int64_t B, float* __restrict__ preds) {
  pdl_wait();
  pdl_trigger();
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * p.rows_total) return;
  const int64_t b = gid / p.rows_total;
  const int row = (int)(gid % p.rows_total);
  int s = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < p.n_scales && row >= p.row0[i]) s = i;
  const int local = row - p.row0[s];
  const int g = local / p.A, a = local % p.A;
  const int E = 3 + p.nc;
  const T* h = reinterpret_cast<const T*>(p.head[s]) + ((int64_t)b * p.G[s] + g) * p.ld[s] + a * E;
  float* o = preds + gid * E;
  for (int j = 0; j < 1 + p.nc; ++j) o[j] = ld_as_float(h + j);
  const float tc = ld_as_float(h + E - 2), tw = ld_as_float(h + E - 1);
  // centers = ((sigmoid*2 - 0.5) + g) * stride / center_scaler      (_architecture.py:146-147)
  float c = __fadd_rn(__fsub_rn(__fmul_rn(sigmoid_f(tc), 2.0f), 0.5f), (float)g);
  c = __fdiv_rn(__fmul_rn(c, (float)p.stride[s]), p.center_scaler);
  // widths = (sigmoid*2)^2 * anchor                                  (_architecture.py:150)
  float w2 = __fmul_rn(sigmoid_f(tw), 2.0f);
  float w = __fmul_rn(__fmul_rn(w2, w2), p.anchors_dev != nullptr ? p.anchors_dev[s * p.A + a] : p.anchors[s * 8 + a]);
  c = fminf(fmaxf(c, 0.0f), p.duration);
  w = fminf(fmaxf(w, 0.0f), p.duration);
  o[E - 2] = c;
  o[E - 1] = w;
}

// ------------------------------------------------------------------------------------ NMS
constexpr int NMS_THREADS = 256;
constexpr int NMS_MAXP = 1024;

__device__ __forceinline__ uint32_t sortable_bits(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// in-place ascending bitonic sort of n_pad (power of two) 64-bit keys in shared memory
__device__ void bitonic_sort_u64(unsigned long long* keys, int n_pad) {
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// The same sort with the keys in REGISTERS (E = n / 256 per thread, key i = tid + 256 m in thread tid): exchange distances
// j >= 256 stay inside a thread, j < 32 are warp shuffles, only j = 32 / 64 / 128 go through shared memory (two buffers in turn:
// one barrier per such stage).  12 block barriers and ~4 k shared-memory wavefronts for n = 1024 instead of 55 and ~14 k: the
// in-place version was half of nms_kernel's time (profiles/r02_ncu_full_nms_b512.txt).  n = 256, 512 or 1024; blockDim.x = 256.
template <int E>
__device__ __forceinline__ void bitonic_sort_u64_regs(unsigned long long* keys, unsigned long long* alt) {
  constexpr int n = 256 * E;
  const int tid = threadIdx.x;
  unsigned long long r[E];
#pragma unroll
  for (int m = 0; m < E; ++m) r[m] = keys[tid + 256 * m];
  unsigned long long* buf[2] = {keys, alt};
  int cur = 0;
#pragma unroll 1
  for (int k = 2; k <= n; k <<= 1) {
#pragma unroll 1
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 256) {
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const int dm = j >> 8;                    // 1 or 2
          if (dm < E && (m & dm) == 0 && (m | dm) < E) {
            const int i = tid + 256 * m;
            const bool up = (i & k) == 0;
            // partner m | dm: selected with compile-time indices
            unsigned long long a = r[m], b = (dm == 1) ? r[(m | 1) < E ? (m | 1) : m] : r[(m | 2) < E ? (m | 2) : m];
            const bool sw = (a > b) == up;
            if (sw) {
              r[m] = b;
              if (dm == 1) r[(m | 1) < E ? (m | 1) : m] = a; else r[(m | 2) < E ? (m | 2) : m] = a;
            }
          }
        }
      } else if (j < 32) {
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const int i = tid + 256 * m;
          const unsigned long long b = __shfl_xor_sync(0xffffffffu, r[m], j);
          const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
          r[m] = keep_min ? (r[m] < b ? r[m] : b) : (r[m] > b ? r[m] : b);
        }
      } else {
        unsigned long long* bf = buf[cur];
#pragma unroll
        for (int m = 0; m < E; ++m) bf[tid + 256 * m] = r[m];
        __syncthreads();
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const int i = tid + 256 * m;
          const unsigned long long b = bf[(tid ^ j) + 256 * m];
          const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
          r[m] = keep_min ? (r[m] < b ? r[m] : b) : (r[m] > b ? r[m] : b);
        }
        cur ^= 1;
      }
    }
  }
  __syncthreads();           // the last readers of either buffer are done
#pragma unroll
  for (int m = 0; m < E; ++m) keys[tid + 256 * m] = r[m];
  __syncthreads();
}

__device__ void bitonic_sort_u64_any(unsigned long long* keys, unsigned long long* alt, int n_pad) {
  if (blockDim.x == 256 && n_pad == 1024)
    bitonic_sort_u64_regs<4>(keys, alt);
  else if (blockDim.x == 256 && n_pad == 512)
    bitonic_sort_u64_regs<2>(keys, alt);
  else if (blockDim.x == 256 && n_pad == 256)
    bitonic_sort_u64_regs<1>(keys, alt);
  else
    bitonic_sort_u64(keys, n_pad);
}

__global__ void __launch_bounds__(NMS_THREADS)
nms_kernel(const float* __restrict__ preds, int P, int nc, double iou_thr, float conf_thr, float duration,
           float box_h, int return_start_end, int32_t* __restrict__ keep_out, int32_t* __restrict__ n_keep_out,
           float* __restrict__ conf_out, float* __restrict__ boxes_out, float* __restrict__ seg_rows,
           int32_t* __restrict__ n_seg_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x;
  const int E = 3 + nc;
  int n_pad = 32;
  while (n_pad < P) n_pad <<= 1;

  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);        // [n_pad]
  float* s_conf = reinterpret_cast<float*>(keys + n_pad);                             // [P] original order
  float* sx1 = s_conf + n_pad;                                                        // [P] sorted order
  float* sx2 = sx1 + n_pad;
  float* sarea = sx2 + n_pad;
  int32_t* s_order = reinterpret_cast<int32_t*>(sarea + n_pad);                       // [P]
  int32_t* s_keep = s_order + n_pad;                                                  // [P] kept sorted positions
  __shared__ uint32_t s_alive[NMS_MAXP / 32];
  __shared__ int s_nseg;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float* pb = preds + (int64_t)b * P * E;
  const float y2 = fminf(fmaxf(box_h, 0.0f), duration);  // coords.clip(0, sample_duration) also hits y2

  // 1. boxes + confidence (inference.py:55-64)
  // Without a keep list to report (keep_out == NULL: process_model_outputs) only boxes with conf > conf_thr can reach the output, and
  // a box below the threshold can only suppress boxes scored even lower (the scan runs in descending score order; inference.py:75-88
  // filters after the NMS): such boxes are dropped BEFORE the sort.  The survivors keep their relative (score desc, index asc)
  // order, so the greedy scan visits exactly the boxes it would have visited and the segments are identical - but the bitonic
  // sort (half of this kernel's time at P = 630 -> 1024 keys) shrinks to the next power of two above the survivor count.
  const bool prefilter = keep_out == nullptr;
  __shared__ int s_nvalid;
  if (threadIdx.x == 0) s_nvalid = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float* r = pb + (int64_t)i * E;
    const float c = r[E - 2], w = r[E - 1];
    const float hw = __fdiv_rn(w, 2.0f);
    const float x1 = fminf(fmaxf(__fsub_rn(c, hw), 0.0f), duration);
    const float x2 = fminf(fmaxf(__fadd_rn(c, hw), 0.0f), duration);
    float mx = r[1];
    for (int j = 1; j < nc; ++j) mx = fmaxf(mx, r[1 + j]);
    float sum = 0.0f;
    for (int j = 0; j < nc; ++j) sum = __fadd_rn(sum, expf(__fsub_rn(r[1 + j], mx)));
    // softmax max element = exp(0) * (1/sum); confidence = class_score * objectness
    const float cf = __fmul_rn(__fdiv_rn(1.0f, sum), sigmoid_f(r[0]));
    s_conf[i] = cf;
    if (conf_out) conf_out[(int64_t)b * P + i] = cf;
    if (boxes_out) {
      boxes_out[((int64_t)b * P + i) * 2 + 0] = x1;
      boxes_out[((int64_t)b * P + i) * 2 + 1] = x2;
    }
    // descending score, ascending index  ->  ascending key
    const unsigned long long key = ((unsigned long long)(~sortable_bits(cf)) << 32) | (unsigned)i;
    if (!prefilter)
      keys[i] = key;
    else if (cf > conf_thr)
      keys[atomicAdd(&s_nvalid, 1)] = key;        // any slot: the sort orders them
  }
  __syncthreads();
  const int Pn = prefilter ? s_nvalid : P;         // boxes that take part in the NMS
  const int W = (Pn + 31) >> 5;                     // alive-mask words
  int n_sort = 32;
  while (n_sort < Pn) n_sort <<= 1;
  for (int i = Pn + threadIdx.x; i < n_sort; i += blockDim.x) keys[i] = ~0ull;
  for (int w = threadIdx.x; w < NMS_MAXP / 32; w += blockDim.x) {
    const int lo = w << 5;
    s_alive[w] = (lo + 32 <= Pn) ? 0xffffffffu : (lo >= Pn ? 0u : ((1u << (Pn - lo)) - 1u));
  }
  __syncthreads();
  bitonic_sort_u64_any(keys, reinterpret_cast<unsigned long long*>(sx1), n_sort);     // sx1 / sx2 (8 n_pad bytes) are free until step 2

  // 2. gather boxes in sorted order
  for (int i = threadIdx.x; i < Pn; i += blockDim.x) {
    const int o = (int)(keys[i] & 0xffffffffu);
    s_order[i] = o;
    const float* r = pb + (int64_t)o * E;
    const float c = r[E - 2], w = r[E - 1];
    const float hw = __fdiv_rn(w, 2.0f);
    const float x1 = fminf(fmaxf(__fsub_rn(c, hw), 0.0f), duration);
    const float x2 = fminf(fmaxf(__fadd_rn(c, hw), 0.0f), duration);
    sx1[i] = x1;
    sx2[i] = x2;
    sarea[i] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, 0.0f));
  }
  __syncthreads();

  // 3. greedy scan (torchvision nms_kernel_impl): pick the next alive box i in score order, then the whole CTA
  //    evaluates row i on the fly - lane k of a warp tests box j = 32*w + k, so shared-memory reads are
  //    conflict-free and only the rows of KEPT boxes are ever computed (no P x P matrix).
  const float hh = fmaxf(0.0f, __fsub_rn(y2, 0.0f));
  int nk = 0, cur = 0;
  while (true) {
    // every warp finds the same next alive index >= cur
    uint32_t wrd = (lane < W) ? s_alive[lane] : 0u;
    const int cw = cur >> 5;
    if (lane < cw) wrd = 0u;
    if (lane == cw) wrd &= ~((1u << (cur & 31)) - 1u);
    const unsigned nz = __ballot_sync(0xffffffffu, wrd != 0u);
    if (nz == 0u) break;
    const int fw = __ffs(nz) - 1;
    const uint32_t fword = __shfl_sync(0xffffffffu, wrd, fw);
    const int i = (fw << 5) + (__ffs(fword) - 1);
    // Without a keep list to report, the scan may stop at the first surviving box that fails the confidence filter: boxes
    // are visited in descending score order and a box can only be suppressed by a higher-scored one, so nothing after this
    // point can change which boxes with conf > conf_thr survive (inference.py:75-88 filters after the NMS).
    if (keep_out == nullptr && !(s_conf[s_order[i]] > conf_thr)) break;
    if (threadIdx.x == 0) s_keep[nk] = i;
    ++nk;
    const float x1i = sx1[i], x2i = sx2[i], ai = sarea[i];
    __syncthreads();   // all warps have read s_alive before anyone clears bits
    for (int w = (i >> 5) + warp; w < W; w += nwarps) {
      const int j = (w << 5) + lane;
      bool sup = false;
      if (j > i && j < Pn) {
        const float ww = fmaxf(0.0f, __fsub_rn(fminf(x2i, sx2[j]), fmaxf(x1i, sx1[j])));
        const float inter = __fmul_rn(ww, hh);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, sarea[j]), inter));
        sup = (double)ovr > iou_thr;   // NaN (two zero-area boxes) compares false: never suppresses
      }
      const unsigned bits = __ballot_sync(0xffffffffu, sup);
      if (lane == 0 && bits) s_alive[w] &= ~bits;
    }
    __syncthreads();
    cur = i + 1;
  }
  if (keep_out != nullptr) {
    for (int i = threadIdx.x; i < P; i += blockDim.x)
      keep_out[(int64_t)b * P + i] = (i < nk) ? s_order[s_keep[i]] : -1;
    if (threadIdx.x == 0) n_keep_out[b] = nk;
  }
  if (seg_rows == nullptr) return;

  // 4. confidence filter + per-clip sort by centre (inference.py:85-99)
  if (threadIdx.x == 0) s_nseg = 0;
  int n_pad2 = 32;
  while (n_pad2 < nk) n_pad2 <<= 1;
  __syncthreads();
  for (int i = threadIdx.x; i < n_pad2; i += blockDim.x) keys[i] = ~0ull;
  __syncthreads();
  for (int i = threadIdx.x; i < nk; i += blockDim.x) {
    const int o = s_order[s_keep[i]];
    if (s_conf[o] > conf_thr) {
      const int slot = atomicAdd(&s_nseg, 1);
      const float c = pb[(int64_t)o * E + E - 2];
      // ascending centre; ties broken by keep rank (the reference's argsort is not declared stable)
      keys[slot] = ((unsigned long long)sortable_bits(c) << 32) | (unsigned)i;
    }
  }
  __syncthreads();
  const int ns = s_nseg;
  bitonic_sort_u64(keys, n_pad2);
  for (int i = threadIdx.x; i < ns; i += blockDim.x) {
    const int o = s_order[s_keep[(int)(keys[i] & 0xffffffffu)]];
    const float* r = pb + (int64_t)o * E;
    float* out = seg_rows + ((int64_t)b * P + i) * 5;
    int label = 0;
    float best = r[1];
    for (int j = 1; j < nc; ++j)
      if (r[1 + j] > best) {
        best = r[1 + j];
        label = j;
      }
    float c = r[E - 2], w = r[E - 1];
    if (return_start_end) {  // inference.py:102-106: start = c - w/2 ; end = start + w ; clip
      const float st = __fsub_rn(c, __fdiv_rn(w, 2.0f));
      const float en = __fadd_rn(st, w);
      c = fminf(fmaxf(st, 0.0f), duration);
      w = fminf(fmaxf(en, 0.0f), duration);
    }
    out[0] = s_conf[o];
    out[1] = r[0];
    out[2] = (float)label;
    out[3] = c;
    out[4] = w;
  }
  if (threadIdx.x == 0) n_seg_out[b] = ns;
}

// single-CTA exclusive scan over per-clip counts + gather into the reference's return layout
__global__ void compact_kernel(const float* __restrict__ seg_rows, const int32_t* __restrict__ n_seg, int64_t B,
                               int P, float* __restrict__ segments, int64_t* __restrict__ batch_idxs,
                               int64_t* __restrict__ total) {
  __shared__ long long s_base;
  __shared__ int s_part[32];
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int64_t b0 = 0; b0 < B; b0 += blockDim.x) {
    const int64_t b = b0 + threadIdx.x;
    const int n = (b < B) ? n_seg[b] : 0;
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_part[w];
    int chunk_total = 0;
    for (int w = 0; w < nw; ++w) chunk_total += s_part[w];
    const long long base = s_base + woff + incl - n;
    for (int i = 0; i < n; ++i) {
      const float* src = seg_rows + (b * P + i) * 5;
      float* dst = segments + (base + i) * 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) dst[k] = src[k];
      batch_idxs[base + i] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = s_base;
}

}  // namespace yad

extern "C" {

static int decode_impl(const void* const* heads, const int32_t* G, const int32_t* ld, const int32_t* stride,
               int32_t n_scales, int32_t dtype, const float* anchors, const float* anchors_dev, int32_t A, int32_t nc,
               float center_scaler, float duration, int64_t B, float* preds, yad_stream_t stream) {
  YAD_CHECK_ARG(n_scales >= 1 && n_scales <= 4, "yad_decode: n_scales=%d not in [1,4]", n_scales);
  YAD_CHECK_ARG(A >= 1 && A <= 8, "yad_decode: A=%d not in [1,8]", A);
  YAD_CHECK_ARG(nc >= 1 && B >= 0 && preds, "yad_decode: bad nc/B/preds");
  YAD_CHECK_ARG(dtype == YAD_F32 || dtype == YAD_BF16, "yad_decode: bad dtype %d", dtype);
  if (B == 0) return YAD_OK;
  yad::DecodeParams p;
  int rows = 0;
  for (int s = 0; s < 4; ++s) {
    p.head[s] = nullptr;
    p.G[s] = p.ld[s] = p.stride[s] = p.row0[s] = 0;
  }
  for (int s = 0; s < n_scales; ++s) {
    YAD_CHECK_ARG(heads[s] && G[s] > 0 && ld[s] >= A * (3 + nc), "yad_decode: scale %d: bad head/G/ld", s);
    p.head[s] = heads[s];
    p.G[s] = G[s];
    p.ld[s] = ld[s];
    p.stride[s] = stride[s];
    p.row0[s] = rows;
    rows += G[s] * A;
    for (int a = 0; a < A; ++a) p.anchors[s * 8 + a] = anchors != nullptr ? anchors[s * A + a] : 0.0f;
  }
  p.anchors_dev = anchors_dev;
  p.n_scales = n_scales;
  p.A = A;
  p.nc = nc;
  p.rows_total = rows;
  p.center_scaler = center_scaler;
  p.duration = duration;
  const int64_t n = B * rows;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  if (dtype == YAD_F32)
    YAD_CUDA(yad::launch_pdl(yad::decode_kernel<float>, dim3(blocks), dim3(threads), 0, (cudaStream_t)stream, p, B, preds));
  else
    YAD_CUDA(yad::launch_pdl(yad::decode_kernel<__nv_bfloat16>, dim3(blocks), dim3(threads), 0, (cudaStream_t)stream, p, B, preds));
  return YAD_OK;
}

int yad_decode(const void* const* heads, const int32_t* G, const int32_t* ld, const int32_t* stride,
               int32_t n_scales, int32_t dtype, const float* anchors, int32_t A, int32_t nc,
               float center_scaler, float duration, int64_t B, float* preds, yad_stream_t stream) {
  YAD_CHECK_ARG(anchors != nullptr, "yad_decode: null anchors");
  return decode_impl(heads, G, ld, stride, n_scales, dtype, anchors, nullptr, A, nc, center_scaler, duration, B, preds, stream);
}

int yad_decode_dev(const void* const* heads, const int32_t* G, const int32_t* ld, const int32_t* stride,
                   int32_t n_scales, int32_t dtype, const float* anchors_dev, int32_t A, int32_t nc,
                   float center_scaler, float duration, int64_t B, float* preds, yad_stream_t stream) {
  YAD_CHECK_ARG(anchors_dev != nullptr, "yad_decode_dev: null anchors");
  return decode_impl(heads, G, ld, stride, n_scales, dtype, nullptr, anchors_dev, A, nc, center_scaler, duration, B, preds, stream);
}

int yad_nms(const float* preds, int64_t B, int32_t P, int32_t nc, double iou_thr, float conf_thr,
            float duration, float box_h, int32_t return_start_end, int32_t* keep, int32_t* n_keep,
            float* conf, float* boxes, float* seg_rows, int32_t* n_seg, yad_stream_t stream) {
  YAD_CHECK_ARG(preds && ((keep == nullptr) == (n_keep == nullptr)) && (keep != nullptr || seg_rows != nullptr),
                "yad_nms: null preds, or keep / n_keep not both given, or nothing to compute");
  YAD_CHECK_ARG(P >= 1 && P <= yad::NMS_MAXP, "yad_nms: P=%d not in [1,%d]", P, yad::NMS_MAXP);
  YAD_CHECK_ARG(nc >= 1 && nc <= 64, "yad_nms: nc=%d not in [1,64]", nc);
  YAD_CHECK_ARG((seg_rows == nullptr) == (n_seg == nullptr), "yad_nms: seg_rows and n_seg go together");
  if (B == 0) return YAD_OK;
  int n_pad = 32;
  while (n_pad < P) n_pad <<= 1;
  const int W = (P + 31) / 32;
  const size_t smem = (size_t)n_pad * 8 + (size_t)n_pad * 4 * 6;
  (void)W;
  if (smem > 48 * 1024)
    YAD_CUDA(cudaFuncSetAttribute(yad::nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  YAD_CUDA(yad::launch_pdl(yad::nms_kernel, dim3((unsigned)B), dim3(yad::NMS_THREADS), smem, (cudaStream_t)stream, preds, (int)P, (int)nc,
                           iou_thr, conf_thr, duration, box_h, (int)return_start_end, keep, n_keep, conf, boxes, seg_rows, n_seg));
  return YAD_OK;
}

int yad_compact_segments(const float* seg_rows, const int32_t* n_seg, int64_t B, int32_t P, float* segments,
                         int64_t* batch_idxs, int64_t* total, yad_stream_t stream) {
  YAD_CHECK_ARG(seg_rows && n_seg && segments && batch_idxs && total, "yad_compact_segments: null pointer");
  YAD_CUDA(yad::launch_pdl(yad::compact_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, seg_rows, n_seg, B, (int)P, segments,
                           batch_idxs, total));
  return YAD_OK;
}

}  // extern "C"
