// Anchor clustering: Lloyd's k-means on the 1-D segment durations, the arithmetic of
// sklearn.cluster.KMeans(algorithm="lloyd") that the reference's compute_anchors.py:72-86 runs (sklearn 1.9:
// _kmeans_single_lloyd / _k_means_lloyd.pyx / _k_means_common.pyx), in fp64 like sklearn on float64 input.
//
// One CTA iterates to convergence on the device (the data set is a few thousand scalars: one launch, no host round trips):
//   E-step   label_i = argmin_j (c_j^2 - 2 x_i c_j), first minimum wins ties (the |x|^2 term is dropped, as in sklearn);
//   M-step   per-cluster sums / counts, reduced in a fixed order (thread-strided partials -> warp shuffles -> warps in
//            order): deterministic, run-to-run bit-identical; empty clusters are re-seeded with the points farthest from
//            their centres (_relocate_empty_clusters_dense);
//   stop     labels unchanged (strict convergence) or sum_j |c_j' - c_j|^2 <= tol (tolerance already scaled by the data
//            variance on the host); without strict convergence the labels are recomputed for the final centres.
#include "common.cuh"

namespace yad {

constexpr int KM_THREADS = 1024;
constexpr int KM_MAXK = 16;

__device__ __forceinline__ int km_label(double x, const double* c, const double* csq, int k) {
  int best = 0;
  double bd = __dsub_rn(csq[0], 2.0 * __dmul_rn(x, c[0]));
  for (int j = 1; j < k; ++j) {
    const double d = __dsub_rn(csq[j], 2.0 * __dmul_rn(x, c[j]));
    if (d < bd) {
      bd = d;
      best = j;
    }
  }
  return best;
}

__global__ void __launch_bounds__(KM_THREADS, 1)
kmeans1d_lloyd_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ centers, int k, int max_iter, double tol,
                      int32_t* __restrict__ labels, int32_t* __restrict__ n_iter_out, double* __restrict__ inertia_out) {
  __shared__ double s_c[KM_MAXK], s_csq[KM_MAXK], s_new[KM_MAXK];
  __shared__ double s_psum[KM_THREADS / 32][KM_MAXK];
  __shared__ double s_pcnt[KM_THREADS / 32][KM_MAXK];
  __shared__ double s_rv[KM_THREADS / 32];
  __shared__ long long s_ri[KM_THREADS / 32];
  __shared__ long long s_chosen[KM_MAXK];
  __shared__ int s_changed, s_stop, s_strict, s_nempty;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < k) {
    s_c[tid] = centers[tid];
    s_csq[tid] = __dmul_rn(centers[tid], centers[tid]);
  }
  if (tid == 0) s_strict = 0;
  for (int64_t i = tid; i < n; i += KM_THREADS) labels[i] = -1;
  __syncthreads();

  int it = 0;
  for (; it < max_iter; ++it) {
    if (tid == 0) s_changed = 0;
    __syncthreads();
    // ---- E-step + thread-local partial sums
    double acc[KM_MAXK], cnt[KM_MAXK];
#pragma unroll
    for (int j = 0; j < KM_MAXK; ++j) acc[j] = 0.0, cnt[j] = 0.0;
    bool changed = false;
    for (int64_t i = tid; i < n; i += KM_THREADS) {
      const double xi = x[i];
      const int l = km_label(xi, s_c, s_csq, k);
      changed |= (l != labels[i]);
      labels[i] = l;
#pragma unroll
      for (int j = 0; j < KM_MAXK; ++j) {
        acc[j] += (l == j) ? xi : 0.0;
        cnt[j] += (l == j) ? 1.0 : 0.0;
      }
    }
    if (changed) s_changed = 1;
#pragma unroll
    for (int j = 0; j < KM_MAXK; ++j) {
      const double a = warp_sum_d(acc[j]), c = warp_sum_d(cnt[j]);
      if (lane == 0) {
        s_psum[warp][j] = a;
        s_pcnt[warp][j] = c;
      }
    }
    __syncthreads();
    if (tid < KM_MAXK) {
      double a = 0.0, c = 0.0;
      for (int w = 0; w < KM_THREADS / 32; ++w) {
        a += s_psum[w][tid];
        c += s_pcnt[w][tid];
      }
      s_psum[0][tid] = a;      // totals live in row 0 from here on
      s_pcnt[0][tid] = c;
    }
    if (tid == 0) s_nempty = 0;
    __syncthreads();
    if (tid == 0) {
      int ne = 0;
      for (int j = 0; j < k; ++j) ne += (s_pcnt[0][j] == 0.0);
      s_nempty = ne;
    }
    __syncthreads();
    // ---- empty clusters: the points farthest from their centre become the new centres (largest first, lowest index on ties)
    const int n_empty = s_nempty;
    for (int e = 0; e < n_empty; ++e) {
      double bv = -1.0;
      long long bi = -1;
      for (int64_t i = tid; i < n; i += KM_THREADS) {
        bool taken = false;
        for (int q = 0; q < e; ++q) taken |= (s_chosen[q] == i);
        if (taken) continue;
        const double d0 = x[i] - s_c[labels[i]];
        const double d = __dmul_rn(d0, d0);
        if (d > bv) {
          bv = d;
          bi = i;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        s_rv[warp] = bv;
        s_ri[warp] = bi;
      }
      __syncthreads();
      if (tid == 0) {
        double v = s_rv[0];
        long long ix = s_ri[0];
        for (int w = 1; w < KM_THREADS / 32; ++w)
          if (s_rv[w] > v || (s_rv[w] == v && s_ri[w] >= 0 && (ix < 0 || s_ri[w] < ix))) {
            v = s_rv[w];
            ix = s_ri[w];
          }
        s_chosen[e] = ix;
        int je = 0;                                // lowest-numbered cluster that is still empty (sklearn's order)
        for (int j = 0; j < k; ++j)
          if (s_pcnt[0][j] == 0.0) {
            je = j;
            break;
          }
        if (ix >= 0) {
          const int old = labels[ix];
          s_psum[0][old] -= x[ix];
          s_pcnt[0][old] -= 1.0;
          s_psum[0][je] = x[ix];
          s_pcnt[0][je] = 1.0;
        }
      }
      __syncthreads();
    }
    // ---- M-step, centre shift, stopping rule (thread 0; k <= 16)
    if (tid == 0) {
      double tot = 0.0;
      for (int j = 0; j < k; ++j) {
        const double nw = s_pcnt[0][j] > 0.0 ? s_psum[0][j] / s_pcnt[0][j] : s_c[j];
        const double d = nw - s_c[j];
        const double sh = sqrt(__dmul_rn(d, d));   // sklearn: per-centre euclidean shift, squared again below
        tot += __dmul_rn(sh, sh);
        s_new[j] = nw;
      }
      for (int j = 0; j < k; ++j) {
        s_c[j] = s_new[j];
        s_csq[j] = __dmul_rn(s_new[j], s_new[j]);
      }
      s_stop = 0;
      if (!s_changed) {
        s_strict = 1;
        s_stop = 1;
      } else if (tot <= tol) {
        s_stop = 1;
      }
    }
    __syncthreads();
    if (s_stop) {
      ++it;
      break;
    }
  }
  // ---- labels for the final centres (unless they are already), inertia
  const bool strict = s_strict != 0;
  double part = 0.0;
  for (int64_t i = tid; i < n; i += KM_THREADS) {
    const double xi = x[i];
    int l = labels[i];
    if (!strict) {
      l = km_label(xi, s_c, s_csq, k);
      labels[i] = l;
    }
    const double d = xi - s_c[l];
    part += __dmul_rn(d, d);
  }
  part = warp_sum_d(part);
  __syncthreads();
  if (lane == 0) s_rv[warp] = part;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int w = 0; w < KM_THREADS / 32; ++w) tot += s_rv[w];
    *inertia_out = tot;
    *n_iter_out = it < max_iter ? it : max_iter;
    for (int j = 0; j < k; ++j) centers[j] = s_c[j];
  }
}

}  // namespace yad

extern "C" int yad_kmeans1d_lloyd(const double* x, int64_t n, double* centers, int32_t k, int32_t max_iter, double tol_abs,
                                  int32_t* labels, int32_t* n_iter, double* inertia, yad_stream_t stream) {
  YAD_CHECK_ARG(x && centers && labels && n_iter && inertia, "yad_kmeans1d_lloyd: null pointer");
  YAD_CHECK_ARG(k >= 1 && k <= yad::KM_MAXK && n >= k && max_iter >= 1 && tol_abs >= 0.0,
                "yad_kmeans1d_lloyd: need 1 <= k <= %d, n >= k, max_iter >= 1 (k=%d, n=%lld)", yad::KM_MAXK, k, (long long)n);
  yad::kmeans1d_lloyd_kernel<<<1, yad::KM_THREADS, 0, (cudaStream_t)stream>>>(x, n, centers, k, max_iter, tol_abs, labels, n_iter,
                                                                            inertia);
  YAD_LAUNCH_CHECK();
  return YAD_OK;
}
