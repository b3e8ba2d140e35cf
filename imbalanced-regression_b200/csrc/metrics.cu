// Evaluation-side reductions of the DIR path: overall and many/median/low-shot MSE, L1 and
// geometric-mean error of a prediction vector, in one pass on the device.
//
//   shot_metrics(preds, labels, train_labels, many_shot_thr=100, low_shot_thr=20)   <- agedb-dir/train.py:338-391
//   validate(): overall MSE / L1 / G-Mean                                           <- agedb-dir/train.py:286-335
//
// The reference loops over np.unique(labels) on the host and, per label value l, counts the training samples with
// int(train_label) == l; a test sample therefore belongs to the "many" group when its label's training count is
// > many_thr, to "low" when it is < low_thr (a label value absent from training, or not integer valued, has count
// 0) and to "median" otherwise.  Here: one exact int64 histogram of int(train_label), then one pass over the test
// samples accumulating, per group, (count, sum d^2, sum |d|, sum log|d|) in fp64.
#include "common.cuh"

namespace dirb200 {

// hist[int(label)]++ for 0 <= int(label) < nbins (no clamping, unlike the LDS histogram of datasets.py:60-63)
__global__ void int_label_hist_kernel(const float* __restrict__ labels, int64_t n, int nbins,
                                      unsigned long long* __restrict__ hist) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = labels[i];
    if (!(v == v)) continue;
    const long long b = (long long)v;            // C truncation == numpy astype(int)
    if (b >= 0 && b < nbins) atomicAdd(&hist[b], 1ull);
  }
}

// out[g][k], g = 0 overall, 1 many, 2 median, 3 low; k = 0 count, 1 sum d^2, 2 sum |d|, 3 sum log|d|
__global__ void __launch_bounds__(256)
shot_metrics_kernel(const float* __restrict__ preds, const float* __restrict__ labels, int64_t n,
                    const unsigned long long* __restrict__ train_hist, int nbins, long long many_thr,
                    long long low_thr, double* __restrict__ out) {
  double acc[3][4];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[g][k] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float l = labels[i];
    const float df = preds[i] - l;               // float32 difference, as the reference's float32 arrays
    const double d = (double)df;
    long long cnt = 0;                           // training samples whose int(label) equals this label value
    if (l >= 0.f && l < (float)nbins && l == floorf(l)) cnt = (long long)train_hist[(int)l];
    const int g = cnt > many_thr ? 0 : (cnt < low_thr ? 2 : 1);
    const double a = fabs(d);
    const double v[4] = {1.0, d * d, a, log(a)};  // log(0) = -inf -> G-Mean 0, as scipy.stats.gmean
#pragma unroll
    for (int gg = 0; gg < 3; ++gg)
      if (gg == g) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[gg][k] += v[k];
      }
  }
  __shared__ double sh[8][12];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double s = warp_sum(acc[g][k]);
      if (lane == 0) sh[warp][g * 4 + k] = s;
    }
  __syncthreads();
  if (threadIdx.x < 12) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w][threadIdx.x];
    // an empty group contributes exactly 0 (and no -inf from log): skip the atomic when nothing was counted
    const double c = [&] {
      double cc = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) cc += sh[w][(threadIdx.x / 4) * 4];
      return cc;
    }();
    if (c > 0.0) {
      atomicAdd(&out[4 + threadIdx.x], t);       // groups 1..3
      atomicAdd(&out[threadIdx.x & 3], t);       // overall
    }
  }
}

}  // namespace dirb200

using namespace dirb200;

extern "C" {

int dirb200_int_label_histogram(const float* labels, int64_t n, int nbins, int64_t* hist, void* stream) {
  DIRB_CHECK_ARG(n >= 0 && nbins > 0 && hist && (labels || n == 0), "int_label_histogram: bad arguments");
  if (n == 0) return DIRB200_OK;
  int64_t g = (n + 255) / 256;
  if (g > 4 * num_sms()) g = 4 * num_sms();
  int_label_hist_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>(labels, n, nbins,
                                                                  reinterpret_cast<unsigned long long*>(hist));
  DIRB_LAUNCHED();
  return DIRB200_OK;
}

int dirb200_shot_metrics(const float* preds, const float* labels, int64_t n, const int64_t* train_hist, int nbins,
                         int many_shot_thr, int low_shot_thr, double* out16, void* stream) {
  DIRB_CHECK_ARG(n >= 0 && nbins > 0 && train_hist && out16 && ((preds && labels) || n == 0),
                 "shot_metrics: bad arguments");
  DIRB_CHECK_ARG(many_shot_thr >= low_shot_thr, "shot_metrics: many_shot_thr must be >= low_shot_thr");
  cudaStream_t st = as_stream(stream);
  DIRB_CUDA(cudaMemsetAsync(out16, 0, 16 * sizeof(double), st));
  if (n == 0) return DIRB200_OK;
  int64_t g = (n + 255) / 256;
  if (g > 2 * num_sms()) g = 2 * num_sms();
  shot_metrics_kernel<<<(unsigned)g, 256, 0, st>>>(preds, labels, n, reinterpret_cast<const unsigned long long*>(train_hist),
                                                 nbins, many_shot_thr, low_shot_thr, out16);
  DIRB_LAUNCHED();
  return DIRB200_OK;
}

}  // extern "C"
