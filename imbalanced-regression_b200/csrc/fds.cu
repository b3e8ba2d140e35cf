// FDS (feature distribution smoothing) kernels -- stage (ii) of the hot path.
//
//   label flags / row binning      <- agedb-dir/fds.py:91-99,120-143 (unique-label loop masks)
//   segmented fp64 accumulation    <- agedb-dir/fds.py:100-102      (mean / var per label bin)
//   finalize + running EMA         <- agedb-dir/fds.py:104-111
//   bin-axis smoothing stencil     <- agedb-dir/fds.py:58-67
//   calibrate forward / backward   <- agedb-dir/fds.py:115-144, agedb-dir/utils.py:97-107
//
// All of it is HBM-bound fp32 streaming work: features are read exactly once
// with 128-bit coalesced loads; per-bin statistics are accumulated as
// (count, sum x, sum x^2) in fp64 so that partial results from row chunks,
// batches and ranks merge by plain addition (all-reduce friendly) without the
// cancellation problems of fp32 sum-of-squares.
#include "common.cuh"

namespace dirb200 {

// ------------------------------------------------------------------ binning
__global__ void label_flags_kernel(const float* __restrict__ labels, int64_t n, float lo, float hi,
                                   int32_t* __restrict__ flags) {
  bool has_lo = false, has_hi = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = labels[i];
    has_lo |= (v == lo);
    has_hi |= (v == hi);
  }
  unsigned blo = __ballot_sync(0xffffffffu, has_lo), bhi = __ballot_sync(0xffffffffu, has_hi);
  if ((threadIdx.x & 31) == 0) {
    if (blo) atomicOr(&flags[0], 1);
    if (bhi) atomicOr(&flags[1], 1);
  }
}

__device__ __forceinline__ int bin_of(float v, float lo, float hi, int nb, bool has_lo, bool has_hi) {
  if (v < lo) return has_lo ? 0 : -1;
  if (v > hi) return has_hi ? nb - 1 : -1;
  if (!(v >= lo)) return -1;  // NaN
  return (int)(v - lo);       // int(label - bucket_start), fds.py:104
}

__global__ void bin_rows_kernel(const float* __restrict__ labels, int64_t n, float lo, float hi, int nb,
                                const int32_t* __restrict__ flags, int32_t* __restrict__ bins) {
  const bool has_lo = flags[0] != 0, has_hi = flags[1] != 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bins[i] = bin_of(labels[i], lo, hi, nb, has_lo, has_hi);
}

// ------------------------------------------------- counting sort of the rows
__global__ void bin_hist_kernel(const int32_t* __restrict__ bins, int64_t n, int nb, int32_t* __restrict__ cnt) {
  extern __shared__ int32_t sh[];
  for (int i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int b = bins[i];
    if (b >= 0 && b < nb) atomicAdd(&sh[b], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x)
    if (sh[i]) atomicAdd(&cnt[i], sh[i]);
}

// single block: exclusive scan of cnt -> offsets[nb+1]; counts64 += cnt
__global__ void bin_scan_kernel(const int32_t* __restrict__ cnt, int nb, int32_t* __restrict__ offsets,
                                int64_t* __restrict__ counts64) {
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = (i < nb) ? cnt[i] : 0;
    // block-wide inclusive scan (blockDim == 1024 max; simple Hillis-Steele in smem)
    __shared__ int32_t tmp[1024];
    tmp[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < blockDim.x; o <<= 1) {
      int t = (threadIdx.x >= o) ? tmp[threadIdx.x - o] : 0;
      __syncthreads();
      tmp[threadIdx.x] += t;
      __syncthreads();
    }
    int incl = tmp[threadIdx.x];
    if (i < nb) {
      offsets[i] = carry + incl - v;
      if (v) counts64[i] += v;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[nb] = carry;
}

__global__ void bin_scatter_kernel(const int32_t* __restrict__ bins, int64_t n, int nb,
                                   const int32_t* __restrict__ offsets, int32_t* __restrict__ cursor,
                                   int32_t* __restrict__ perm, int32_t* __restrict__ sbin) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int b = bins[i];
    if (b < 0 || b >= nb) continue;
    int pos = offsets[b] + atomicAdd(&cursor[b], 1);
    perm[pos] = (int32_t)i;
    sbin[pos] = b;
  }
}

// ------------------------------------------------------ segmented accumulate
// grid.x = row chunks of R sorted positions, grid.y = column slices of
// blockDim.x*VEC columns.  Each thread keeps fp64 (sum, sumsq) for its VEC
// columns and flushes them with fp64 atomics whenever the (sorted) bin
// changes and at the end of its chunk.
template <int VEC>
__device__ __forceinline__ void flush_acc(double* __restrict__ sums, double* __restrict__ sumsq, int bin, int d,
                                          int col, double (&s)[VEC], double (&q)[VEC]) {
  double* ps = sums + (size_t)bin * d + col;
  double* pq = sumsq + (size_t)bin * d + col;
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    atomicAdd(ps + k, s[k]);
    atomicAdd(pq + k, q[k]);
    s[k] = 0.0;
    q[k] = 0.0;
  }
}

template <int VEC, int UNROLL>
__global__ void __launch_bounds__(256)
fds_accumulate_kernel(const float* __restrict__ feat, const int32_t* __restrict__ perm,
                      const int32_t* __restrict__ sbin, const int32_t* __restrict__ offsets, int nb, int d,
                      int rows_per_chunk, double* __restrict__ sums, double* __restrict__ sumsq) {
  const int total = offsets[nb];
  const int start = blockIdx.x * rows_per_chunk;
  if (start >= total) return;
  const int end = min(total, start + rows_per_chunk);
  const int col = (blockIdx.y * blockDim.x + threadIdx.x) * VEC;
  if (col >= d) return;

  double s[VEC], q[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) s[k] = q[k] = 0.0;
  int cur = sbin[start];

  for (int i = start; i < end; i += UNROLL) {
    float v[UNROLL][VEC];
    int b[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int p = i + u;
      if (p < end) {
        const size_t row = (size_t)perm[p];
        b[u] = sbin[p];
        if constexpr (VEC == 4) {
          float4 t = ldg_stream(reinterpret_cast<const float4*>(feat + row * d + col));
          v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w;
        } else {
#pragma unroll
          for (int k = 0; k < VEC; ++k) v[u][k] = __ldg(feat + row * d + col + k);
        }
      } else {
        b[u] = -1;
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (b[u] < 0) continue;
      if (b[u] != cur) {
        flush_acc<VEC>(sums, sumsq, cur, d, col, s, q);
        cur = b[u];
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const double x = (double)v[u][k];
        s[k] += x;
        q[k] = fma(x, x, q[k]);
      }
    }
  }
  flush_acc<VEC>(sums, sumsq, cur, d, col, s, q);
}

// ------------------------------------------------------------------ finalize
__global__ void fds_finalize_kernel(const double* __restrict__ sums, const double* __restrict__ sumsq,
                                    const int64_t* __restrict__ counts, int nb, int d,
                                    float* __restrict__ running_mean, float* __restrict__ running_var,
                                    const float* __restrict__ tracked, double momentum, int first_update) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)nb * d) return;
  const int b = (int)(idx / d);
  const int64_t n = counts[b];
  if (n <= 0) return;
  const double dn = (double)n;
  const double sx = sums[idx], sq = sumsq[idx];
  const double mean = sx / dn;
  double var = 0.0;
  if (n > 1) {
    var = (sq - sx * mean) / (dn - 1.0);
    if (var < 0.0) var = 0.0;
  }
  double factor;
  if (first_update) {
    factor = 0.0;
  } else if (momentum >= 0.0) {
    factor = momentum;
  } else {
    const float tracked_new = tracked[b] + (float)n;   // float32 buffer += n   (fds.py:104)
    factor = 1.0 - dn / (double)tracked_new;           // fds.py:105-106
  }
  const float a = (float)(1.0 - factor), f = (float)factor;
  // (1 - factor) * cur + factor * running, two roundings each like the eager fp32 ops
  running_mean[idx] = __fadd_rn(__fmul_rn(a, (float)mean), __fmul_rn(f, running_mean[idx]));
  running_var[idx] = __fadd_rn(__fmul_rn(a, (float)var), __fmul_rn(f, running_var[idx]));
}

__global__ void fds_bump_tracked_kernel(const int64_t* __restrict__ counts, int nb, float* __restrict__ tracked) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb && counts[b] > 0) tracked[b] += (float)counts[b];
}

// ----------------------------------------------------------- bin-axis stencil
struct Window {
  float w[33];
};

__global__ void fds_smooth_tables_kernel(const float* __restrict__ src, int nb, int d, Window win, int ks,
                                         float* __restrict__ dst) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)nb * d) return;
  const int b = (int)(idx / d), c = (int)(idx % d);
  const int h = (ks - 1) / 2;
  float acc = 0.f;
  for (int j = 0; j < ks; ++j) {
    int k = b + j - h;
    if (k < 0) k = -k;                       // reflect (edge sample not repeated)
    if (k >= nb) k = 2 * (nb - 1) - k;
    acc = __fadd_rn(acc, __fmul_rn(win.w[j], src[(size_t)k * d + c]));
  }
  dst[idx] = acc;
}

// ------------------------------------------------------------------ calibrate
// One CTA per row.  FUSED_FLAGS: the CTA scans the batch's labels itself for
// the two edge flags (b is small on the per-step path), otherwise it reads
// precomputed flags.
template <bool FUSED_FLAGS>
__global__ void __launch_bounds__(256)
fds_calibrate_fwd_kernel(float* __restrict__ x, const float* __restrict__ labels, int64_t nrows, int d, float lo,
                         float hi, int nb, const int32_t* __restrict__ flags, const float* __restrict__ m1,
                         const float* __restrict__ v1, const float* __restrict__ m2, const float* __restrict__ v2,
                         float clip_min, float clip_max, int32_t* __restrict__ rowbin) {
  __shared__ int sh_flags[2];
  __shared__ float sh_red[8];
  const int64_t row = blockIdx.x;
  bool has_lo, has_hi;
  if (FUSED_FLAGS) {
    if (threadIdx.x < 2) sh_flags[threadIdx.x] = 0;
    __syncthreads();
    bool l = false, h = false;
    for (int64_t i = threadIdx.x; i < nrows; i += blockDim.x) {
      float v = labels[i];
      l |= (v == lo);
      h |= (v == hi);
    }
    if (l) sh_flags[0] = 1;
    if (h) sh_flags[1] = 1;
    __syncthreads();
    has_lo = sh_flags[0] != 0;
    has_hi = sh_flags[1] != 0;
  } else {
    has_lo = flags[0] != 0;
    has_hi = flags[1] != 0;
  }
  const int bin = bin_of(labels[row], lo, hi, nb, has_lo, has_hi);
  if (bin < 0) {
    if (threadIdx.x == 0) rowbin[row] = -1;
    return;
  }
  const float* pv1 = v1 + (size_t)bin * d;
  // torch.sum(v1) < 1e-10  -> identity   (utils.py:98-99)
  float part = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) part += pv1[c];
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) sh_red[threadIdx.x >> 5] = part;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) tot += sh_red[w];
  if (tot < 1e-10f) {
    if (threadIdx.x == 0) rowbin[row] = -1;
    return;
  }
  if (threadIdx.x == 0) rowbin[row] = bin;
  const float* pm1 = m1 + (size_t)bin * d;
  const float* pm2 = m2 + (size_t)bin * d;
  const float* pv2 = v2 + (size_t)bin * d;
  float* px = x + (size_t)row * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float a = pv1[c];
    if (a == 0.f) continue;                                   // utils.py:100-104
    float fac = __fdiv_rn(pv2[c], a);
    fac = fminf(fmaxf(fac, clip_min), clip_max);              // torch.clamp
    // (x - m1) * sqrt(factor) + m2, separate roundings (no FMA) like eager torch
    px[c] = __fadd_rn(__fmul_rn(__fsub_rn(px[c], pm1[c]), __fsqrt_rn(fac)), pm2[c]);
  }
}

__global__ void fds_calibrate_bwd_kernel(const float* __restrict__ gout, const int32_t* __restrict__ rowbin,
                                         int64_t nrows, int d, const float* __restrict__ v1,
                                         const float* __restrict__ v2, float clip_min, float clip_max,
                                         float* __restrict__ gin) {
  const int64_t row = blockIdx.x;
  const int bin = rowbin[row];
  const float* pg = gout + (size_t)row * d;
  float* po = gin + (size_t)row * d;
  if (bin < 0) {
    if (pg != po)
      for (int c = threadIdx.x; c < d; c += blockDim.x) po[c] = pg[c];
    return;
  }
  const float* pv1 = v1 + (size_t)bin * d;
  const float* pv2 = v2 + (size_t)bin * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float a = pv1[c];
    float s = 1.f;
    if (a != 0.f) s = __fsqrt_rn(fminf(fmaxf(__fdiv_rn(pv2[c], a), clip_min), clip_max));
    po[c] = pg[c] * s;
  }
}

static inline int grid_for(int64_t n, int block, int cap) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace dirb200

using namespace dirb200;

extern "C" {

int dirb200_fds_label_flags(const float* labels, int64_t n, int bucket_num, int bucket_start, int bin_rule,
                            int32_t* flags, void* stream) {
  DIRB_CHECK_ARG(bin_rule == DIRB200_BIN_AGE, "fds_label_flags: unknown bin_rule %d", bin_rule);
  DIRB_CHECK_ARG(n >= 0 && flags && (labels || n == 0), "fds_label_flags: bad arguments");
  DIRB_CHECK_ARG(bucket_num > bucket_start, "fds_label_flags: bucket_num must exceed bucket_start");
  if (n == 0) return DIRB200_OK;
  label_flags_kernel<<<grid_for(n, 256, 4 * num_sms()), 256, 0, as_stream(stream)>>>(
      labels, n, (float)bucket_start, (float)(bucket_num - 1), flags);
  DIRB_LAUNCHED();
  return DIRB200_OK;
}

int dirb200_fds_bin_rows(const float* labels, int64_t n, int bucket_num, int bucket_start, int bin_rule,
                         const int32_t* flags, int32_t* bins_out, void* stream) {
  DIRB_CHECK_ARG(bin_rule == DIRB200_BIN_AGE, "fds_bin_rows: unknown bin_rule %d", bin_rule);
  DIRB_CHECK_ARG(n >= 0 && flags && (n == 0 || (labels && bins_out)), "fds_bin_rows: bad arguments");
  DIRB_CHECK_ARG(bucket_num > bucket_start, "fds_bin_rows: bucket_num must exceed bucket_start");
  if (n == 0) return DIRB200_OK;
  bin_rows_kernel<<<grid_for(n, 256, 4 * num_sms()), 256, 0, as_stream(stream)>>>(
      labels, n, (float)bucket_start, (float)(bucket_num - 1), bucket_num - bucket_start, flags, bins_out);
  DIRB_LAUNCHED();
  return DIRB200_OK;
}

size_t dirb200_fds_accumulate_workspace_bytes(int64_t n, int nb) {
  if (n < 0 || nb <= 0) return 0;
  // cnt[nb] | cursor[nb] | offsets[nb+1] | perm[n] | sbin[n]     (int32)
  return sizeof(int32_t) * ((size_t)3 * nb + 1 + 2 * (size_t)n) + 64;
}

int dirb200_fds_accumulate(const float* features, const int32_t* bins, int64_t n, int d, int nb, double* sums,
                           double* sumsq, int64_t* counts, void* workspace, size_t workspace_bytes,
                           void* stream) {
  DIRB_CHECK_ARG(n >= 0 && d > 0 && nb > 0 && nb <= 8192, "fds_accumulate: bad sizes n=%lld d=%d nb=%d",
                 (long long)n, d, nb);
  DIRB_CHECK_ARG(n < (int64_t)1 << 31, "fds_accumulate: n too large for one call; stream it in batches");
  DIRB_CHECK_ARG(sums && sumsq && counts, "fds_accumulate: null accumulator");
  if (n == 0) return DIRB200_OK;
  DIRB_CHECK_ARG(features && bins && workspace, "fds_accumulate: null pointer");
  if (workspace_bytes < dirb200_fds_accumulate_workspace_bytes(n, nb)) {
    set_error("fds_accumulate: workspace too small (%zu < %zu)", workspace_bytes,
              dirb200_fds_accumulate_workspace_bytes(n, nb));
    return DIRB200_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  int32_t* cnt = reinterpret_cast<int32_t*>(workspace);
  int32_t* cursor = cnt + nb;
  int32_t* offsets = cursor + nb;
  int32_t* perm = offsets + nb + 1;
  int32_t* sbin = perm + n;
  DIRB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * 2 * nb, st));
  const int g = grid_for(n, 256, 4 * num_sms());
  bin_hist_kernel<<<g, 256, sizeof(int32_t) * nb, st>>>(bins, n, nb, cnt);
  DIRB_LAUNCHED();
  bin_scan_kernel<<<1, 1024, 0, st>>>(cnt, nb, offsets, counts);
  DIRB_LAUNCHED();
  bin_scatter_kernel<<<g, 256, 0, st>>>(bins, n, nb, offsets, cursor, perm, sbin);
  DIRB_LAUNCHED();
  const int rows_per_chunk = n >= 65536 ? 128 : (n >= 8192 ? 16 : 8);
  const bool vec4 = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(features) & 15) == 0);
  const int cols_per_block = 256 * (vec4 ? 4 : 1);
  dim3 grid((unsigned)((n + rows_per_chunk - 1) / rows_per_chunk), (unsigned)((d + cols_per_block - 1) / cols_per_block));
  if (vec4)
    fds_accumulate_kernel<4, 8><<<grid, 256, 0, st>>>(features, perm, sbin, offsets, nb, d, rows_per_chunk, sums, sumsq);
  else
    fds_accumulate_kernel<1, 4><<<grid, 256, 0, st>>>(features, perm, sbin, offsets, nb, d, rows_per_chunk, sums, sumsq);
  DIRB_LAUNCHED();
  return DIRB200_OK;
}

int dirb200_fds_finalize(const double* sums, const double* sumsq, const int64_t* counts, int nb, int d,
                         float* running_mean, float* running_var, float* num_samples_tracked, double momentum,
                         int first_update, void* stream) {
  DIRB_CHECK_ARG(sums && sumsq && counts && running_mean && running_var && num_samples_tracked,
                 "fds_finalize: null pointer");
  DIRB_CHECK_ARG(nb > 0 && d > 0, "fds_finalize: bad sizes");
  cudaStream_t st = as_stream(stream);
  const int64_t tot = (int64_t)nb * d;
  fds_finalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(sums, sumsq, counts, nb, d, running_mean,
                                                                     running_var, num_samples_tracked, momentum,
                                                                     first_update);
  DIRB_LAUNCHED();
  fds_bump_tracked_kernel<<<(nb + 127) / 128, 128, 0, st>>>(counts, nb, num_samples_tracked);
  DIRB_LAUNCHED();
  return DIRB200_OK;
}

int dirb200_fds_smooth_tables(const float* src, int nb, int d, const float* window_host, int ks, float* dst,
                              void* stream) {
  DIRB_CHECK_ARG(src && dst && window_host, "fds_smooth_tables: null pointer");
  DIRB_CHECK_ARG(src != dst, "fds_smooth_tables: src and dst must not alias");
  DIRB_CHECK_ARG(ks >= 1 && ks <= 33 && (ks & 1), "fds_smooth_tables: ks must be odd and <= 33 (got %d)", ks);
  DIRB_CHECK_ARG(nb > (ks - 1) / 2, "fds_smooth_tables: reflect padding needs half_ks < number of bins");
  Window w;
  for (int i = 0; i < 33; ++i) w.w[i] = i < ks ? window_host[i] : 0.f;
  const int64_t tot = (int64_t)nb * d;
  fds_smooth_tables_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, as_stream(stream)>>>(src, nb, d, w, ks, dst);
  DIRB_LAUNCHED();
  return DIRB200_OK;
}

int dirb200_fds_calibrate_fwd(float* x, const float* labels, int64_t b, int d, int bucket_num, int bucket_start,
                              int bin_rule, const float* m1, const float* v1, const float* m2, const float* v2,
                              float clip_min, float clip_max, int32_t* rowbin_out, int32_t* flags_scratch,
                              void* stream) {
  DIRB_CHECK_ARG(bin_rule == DIRB200_BIN_AGE, "fds_calibrate_fwd: unknown bin_rule %d", bin_rule);
  DIRB_CHECK_ARG(b >= 0 && d > 0 && bucket_num > bucket_start, "fds_calibrate_fwd: bad sizes");
  if (b == 0) return DIRB200_OK;
  DIRB_CHECK_ARG(x && labels && m1 && v1 && m2 && v2 && rowbin_out, "fds_calibrate_fwd: null pointer");
  DIRB_CHECK_ARG(b < ((int64_t)1 << 31), "fds_calibrate_fwd: too many rows for one call");
  const float lo = (float)bucket_start, hi = (float)(bucket_num - 1);
  const int nb = bucket_num - bucket_start;
  cudaStream_t st = as_stream(stream);
  if (b <= 2048) {
    fds_calibrate_fwd_kernel<true><<<(unsigned)b, 256, 0, st>>>(x, labels, b, d, lo, hi, nb, nullptr, m1, v1, m2, v2,
                                                                clip_min, clip_max, rowbin_out);
    DIRB_LAUNCHED();
  } else {
    DIRB_CHECK_ARG(flags_scratch, "fds_calibrate_fwd: flags_scratch (int32[2]) is required when b > 2048");
    DIRB_CUDA(cudaMemsetAsync(flags_scratch, 0, 2 * sizeof(int32_t), st));
    label_flags_kernel<<<grid_for(b, 256, 4 * num_sms()), 256, 0, st>>>(labels, b, lo, hi, flags_scratch);
    DIRB_LAUNCHED();
    fds_calibrate_fwd_kernel<false><<<(unsigned)b, 256, 0, st>>>(x, labels, b, d, lo, hi, nb, flags_scratch, m1, v1,
                                                                 m2, v2, clip_min, clip_max, rowbin_out);
    DIRB_LAUNCHED();
  }
  return DIRB200_OK;
}

int dirb200_fds_calibrate_bwd(const float* grad_out, const int32_t* rowbin, int64_t b, int d, const float* v1,
                              const float* v2, float clip_min, float clip_max, float* grad_in, void* stream) {
  DIRB_CHECK_ARG(b >= 0 && d > 0, "fds_calibrate_bwd: bad sizes");
  if (b == 0) return DIRB200_OK;
  DIRB_CHECK_ARG(grad_out && rowbin && v1 && v2 && grad_in, "fds_calibrate_bwd: null pointer");
  fds_calibrate_bwd_kernel<<<(unsigned)b, 256, 0, as_stream(stream)>>>(grad_out, rowbin, b, d, v1, v2, clip_min,
                                                                        clip_max, grad_in);
  DIRB_LAUNCHED();
  return DIRB200_OK;
}

}  // extern "C"
