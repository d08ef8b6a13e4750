// K6: block-cooperative exact order statistics (np.nanmedian / np.median / np.nanstd semantics).
// Replaces np.nanmedian/np.nanstd at /root/reference/src/lightkurve/lightcurve.py:1003-1005,1035,1050,
// np.median at correctors/regressioncorrector.py:279 and the median/std inside astropy sigma_clip (:269).
// MSB-first 8-bit radix select over order-preserving uint64 keys of fp64 values; 8 passes,
// shared-memory integer histogram; exact (no approximation, no sorting of the payload).
#pragma once
#include "common.cuh"

namespace lkb {

struct SelSmem {
  int hist[256];
  double red[33];
  long long redll[33];
  unsigned long long prefix;
  long long k;
  int digit;
};

__device__ __forceinline__ unsigned long long f64_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_unkey(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// k-th smallest (0-based) among the non-NaN values get(0..n-1).  All threads must call.
// Precondition: 0 <= k < (number of non-NaN values).  Result valid in all threads.
template <class Get>
__device__ unsigned long long block_select_key(Get get, int64_t n, long long k, SelSmem& sm) {
  unsigned long long prefix = 0, mask = 0;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sm.hist[i] = 0;
    __syncthreads();
    for (int64_t i0 = 0; i0 < n; i0 += blockDim.x) {         // warp-uniform trip count
      const int64_t i = i0 + threadIdx.x;
      int digit = -1;
      if (i < n) {
        const double v = get(i);
        if (v == v) {
          const unsigned long long key = f64_key(v);
          if ((key & mask) == prefix) digit = (int)((key >> shift) & 255ull);
        }
      }
      // warp-aggregated update: one atomic per distinct digit per warp (a regular cadence makes every
      // dt identical - a naive per-lane atomicAdd would serialise 32-way on one counter)
      const unsigned peers = __match_any_sync(0xffffffffu, digit);
      if (digit >= 0 && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&sm.hist[digit], __popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // warp 0: find the digit whose cumulative count first exceeds k
      const int lane = threadIdx.x;
      int loc[8], tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { loc[j] = sm.hist[lane * 8 + j]; tot += loc[j]; }
      int incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      const long long excl = (long long)incl - tot;
      if (k >= excl && k < (long long)incl) {
        long long kk = k - excl;
        int d = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (kk >= loc[j]) { kk -= loc[j]; d = j + 1; } else break;
        }
        sm.digit = lane * 8 + d;
        sm.k = kk;
      }
    }
    __syncthreads();
    prefix |= ((unsigned long long)sm.digit) << shift;
    mask |= 255ull << shift;
    k = sm.k;
    __syncthreads();
  }
  return prefix;
}

// np.nanmedian over get(0..n-1) (NaN entries ignored; all-NaN -> NaN).  `count_out` (optional)
// receives the number of non-NaN entries.
template <class Get>
__device__ double block_nanmedian(Get get, int64_t n, SelSmem& sm, long long* count_out = nullptr) {
  long long cnt = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = get(i);
    cnt += (v == v) ? 1 : 0;
  }
  const long long m = block_sum_ll(cnt, sm.redll);
  if (count_out) *count_out = m;
  if (m == 0) return __longlong_as_double(0x7ff8000000000000ll);
  const long long klo = (m - 1) / 2, khi = m / 2;
  const unsigned long long key_lo = block_select_key(get, n, klo, sm);
  const double vlo = f64_unkey(key_lo);
  if (khi == klo) return vlo;
  // the next order statistic: equal to vlo if enough duplicates, else the smallest value above it
  long long le = 0;
  double mn = __longlong_as_double(0x7ff0000000000000ll);   // +inf
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = get(i);
    if (v == v) {
      const unsigned long long key = f64_key(v);
      if (key <= key_lo) le++;
      else mn = fmin(mn, v);
    }
  }
  const long long le_tot = block_sum_ll(le, sm.redll);
  // block min
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm.red[threadIdx.x >> 5] = mn;
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = sm.red[0];
    for (int w = 1; w < (int)((blockDim.x + 31) >> 5); ++w) x = fmin(x, sm.red[w]);
    sm.red[32] = x;
  }
  __syncthreads();
  const double vnext = sm.red[32];
  __syncthreads();
  const double vhi = (le_tot >= khi + 1) ? vlo : vnext;
  return (vlo + vhi) / 2.0;
}

// np.nanstd (ddof = 0): two-pass (mean, then squared deviations), NaN ignored.
template <class Get>
__device__ double block_nanstd(Get get, int64_t n, SelSmem& sm, double* mean_out = nullptr) {
  double s = 0.0;
  long long c = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = get(i);
    if (v == v) { s += v; c++; }
  }
  const double tot = block_sum(s, sm.red);
  const long long m = block_sum_ll(c, sm.redll);
  if (m == 0) return __longlong_as_double(0x7ff8000000000000ll);
  const double mean = tot / (double)m;
  if (mean_out) *mean_out = mean;
  double q = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = get(i);
    if (v == v) { const double d = v - mean; q += d * d; }
  }
  const double qq = block_sum(q, sm.red);
  return sqrt(qq / (double)m);
}

}  // namespace lkb
