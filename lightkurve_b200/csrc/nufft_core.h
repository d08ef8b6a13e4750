// Arithmetic of the NUFFT Lomb-Scargle path (ls_nufft.cu) as __host__ __device__ functions: every kernel of that
// path is "one thread = one call of a function in this header" with no shared memory and no intra-CTA
// communication, so the SAME code is compiled for the CPU by tests/nufft_host_harness.cpp and checked against
// numpy's FFT and the fp64 oracle without a GPU (tests/test_nufft_core.py).
//
// Algorithm (type-1 NUFFT, "exponential of semicircle" kernel of Barnett, Magland & af Klinteberg 2019 - the
// method behind lightkurve's optional ls_method="fastnifty", /root/reference/pyproject.toml:48,
// src/lightkurve/periodogram.py:917-946; design study with the accuracy numbers: tools/nufft_ls_model.py):
//
//   sum_n a_n exp(2 pi i (k0 + k) df t_n),  k = 0 .. F-1,   t_n >= 0 sorted,   df * t_max <= 1
//
//   1. cadence n sits at fine-grid coordinate x_n = df t_n M + shift (cells, M = 2^p >= 4 (k0 + F)); it feeds the
//      w cells i0_n .. i0_n + w - 1 (i0_n = ceil(x_n - w/2)) with weight phi((cell - x_n) / (w/2)),
//      phi(z) = exp(beta (sqrt(1 - z^2) - 1)), beta = 2.30 w.  GATHER form: one thread per (cell, pair of light
//      curves) sums the cadences that reach its cell (a contiguous range, because i0 is sorted) - no atomics.
//   2. two light curves share one complex transform (real part / imaginary part); length-M FFT with the +i sign
//      as a sequence of out-of-place Stockham passes of radix 16 / 8 / 4 / 2, one thread per butterfly.
//   3. mode kk = k0 + k of light curve a is (Z[kk] + conj Z[M - kk]) / 2, of light curve b
//      (Z[kk] - conj Z[M - kk]) / 2i; both are multiplied by exp(-2 pi i kk shift / M) / phihat(kk).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LKB_HD __host__ __device__ __forceinline__
#else
#define LKB_HD inline
#if !defined(LKB_CUDA_EMU)          // (tests/native/cuda_emu.h brings CUDA's own vector types)
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
struct double2 { double x, y; };
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
#endif
#endif

namespace lkb {
namespace nufft {

struct Cad {          // per cadence: leftmost fine-grid cell of its kernel support, and (that cell - x_n)
  int32_t i0;
  float d0;           // in [-w/2, -w/2 + 1)
};

LKB_HD int grid_shift(int w) { return (w + 1) / 2 + 1; }            // keeps every i0 >= 1
LKB_HD int64_t table_len(int64_t M, int w) { return M + 2 * (int64_t)w + 4; }   // entries of first_ge

// fine-grid size: power of two >= 2 sigma_min (highest mode index + 1); sigma_min = 2 is the classical upsampling
// factor, 1.25 the low-upsampling choice (smaller transform, wider kernel - es_width / es_beta below)
LKB_HD int fine_grid_log2(int64_t kmax_plus_1, double sigma_min = 2.0) {
  int p = 4;
  while ((double)((int64_t)1 << p) < 2.0 * sigma_min * (double)kmax_plus_1) ++p;
  return p;
}
// upsampling factor a grid of 2^p cells actually gives for modes up to kmax (capped: nothing is gained past 4)
LKB_HD double grid_sigma(int p, int64_t kmax_plus_1) {
  const double s = (double)((int64_t)1 << p) / (2.0 * (double)kmax_plus_1);
  return s > 4.0 ? 4.0 : s;
}
// "exponential of semicircle" kernel phi(z) = exp(beta (sqrt(1 - z^2) - 1)) at upsampling factor sigma (Barnett,
// Magland & af Klinteberg 2019, the finufft parameter rule): beta = 0.976 pi (1 - 1 / (2 sigma)) w  (2.30 w at
// sigma = 2), and the even width that puts the aliasing error near 1e-9 of sum |y|.
LKB_HD double es_beta(int w, double sigma) { return 0.976 * 3.14159265358979323846 * (1.0 - 0.5 / sigma) * (double)w; }
LKB_HD int es_width(double sigma) {
  int w = (int)ceil(20.7 / (3.14159265358979323846 * sqrt(1.0 - 1.0 / sigma)));
  w += w & 1;
  return w < 10 ? 10 : (w > 16 ? 16 : w);
}

LKB_HD Cad cad_entry(double t_rel, double df, int64_t M, int w) {
  const double x = df * t_rel * (double)M + (double)grid_shift(w);
  const double i0 = ceil(x - 0.5 * (double)w);
  Cad c;
  c.i0 = (int32_t)i0;
  c.d0 = (float)(i0 - x);
  return c;
}

LKB_HD float es_eval(float z, float beta) {
  const float q = 1.0f - z * z;
  return q > 0.0f ? expf(beta * (sqrtf(q) - 1.0f)) : 0.0f;
}

// first_ge[c] = number of cadences with i0 < c  (lower bound in the sorted i0), c in [0, table_len)
LKB_HD int32_t first_ge_entry(int64_t c, const Cad* cad, int64_t N) {
  int64_t lo = 0, hi = N;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)cad[mid].i0 < c) lo = mid + 1; else hi = mid;
  }
  return (int32_t)lo;
}

// The two light curves of a pair share one fp32 transform, so its rounding errors scale with the LARGER of the
// two: each is brought to max |y| in [0.5, 1) by an (exact) power of two before it is spread, and scaled back after
// the unpacking.  Returns that power of two (1 for an all-zero light curve).
LKB_HD float pow2_scale(float absmax) {
  if (!(absmax > 0.0f) || !(absmax < 3.0e38f)) return 1.0f;
  int e;
  frexpf(absmax, &e);
  return ldexpf(1.0f, -e);
}

// value of fine-grid cell m for the pair (s0 y0, s1 y1) of light curves (y1 may be NULL: imaginary part 0)
LKB_HD float2 spread_cell(int64_t m, const int32_t* first_ge, const Cad* cad, const float* y0, const float* y1,
                          float s0, float s1, int w, float beta, int64_t M) {
  float2 acc = make_float2(0.0f, 0.0f);
  const float inv_half = 2.0f / (float)w;
  const int64_t L = table_len(M, w);
  for (int wrap = 0; wrap < 2; ++wrap) {           // wrap = 1: cadences whose support runs past cell M - 1
    const int64_t mm = m + (int64_t)wrap * M;
    if (mm + 1 >= L) break;
    int64_t lo_c = mm - w + 1;
    if (lo_c < 0) lo_c = 0;
    const int32_t a = first_ge[lo_c], b = first_ge[mm + 1];
    for (int32_t n = a; n < b; ++n) {
      const Cad c = cad[n];
      const float ph = es_eval((c.d0 + (float)(mm - (int64_t)c.i0)) * inv_half, beta);
      acc.x += ph * (s0 * y0[n]);
      if (y1) acc.y += ph * (s1 * y1[n]);
    }
  }
  return acc;
}

// Ragged batches have one cadence table per light curve and no per-cell lookup table: the cadence range that
// reaches cell m is found by two binary searches in the light curve's own (sorted) table.  Returns the cell value
// of ONE light curve (y == NULL: unit strengths, for the window terms).
LKB_HD float spread_cell_search(int64_t m, const Cad* cad, int64_t n, const float* y, float scale, int w, float beta,
                                int64_t M) {
  float acc = 0.0f;
  const float inv_half = 2.0f / (float)w;
  const int64_t L = table_len(M, w);
  for (int wrap = 0; wrap < 2; ++wrap) {
    if (wrap && m > 2 * (int64_t)w + 2) break;
    const int64_t mm = m + (int64_t)wrap * M;
    if (mm + 1 >= L) break;
    int64_t lo_c = mm - w + 1;
    if (lo_c < 0) lo_c = 0;
    const int32_t a = first_ge_entry(lo_c, cad, n), b = first_ge_entry(mm + 1, cad, n);
    for (int32_t i = a; i < b; ++i) {
      const Cad c = cad[i];
      const float ph = es_eval((c.d0 + (float)(mm - (int64_t)c.i0)) * inv_half, beta);
      acc += y ? ph * (scale * y[i]) : ph;
    }
  }
  return acc;
}

// The same with the kernel weight evaluated in FP64 from the time stamp itself (x = dfM t + shift, exactly as cad_entry
// places the cadence) and rounded once.  The fp32 form above carries ~1e-6 relative weight errors (fp32 offsets, expf):
// a perturbation of the kernel SHAPE the deconvolution does not undo, which leaks a strong line from above the
// frequency grid into the band at ~1e-8 of its amplitude - 7x the rounding noise of the transform itself (emulated
// ragged light curve, flux excursion 600x its in-band peak: rms error 0.085 of the tolerance against 0.012).
LKB_HD float es_eval_acc(double z, double beta) {
  const double q = 1.0 - z * z;
  return q > 0.0 ? (float)exp(beta * (sqrt(q) - 1.0)) : 0.0f;
}
LKB_HD float spread_cell_search_acc(int64_t m, const Cad* cad, const double* t, int64_t n, const float* y, int w,
                                    double beta, double dfM, int64_t M) {
  float acc = 0.0f;
  const double inv_half = 2.0 / (double)w, shift = (double)grid_shift(w);
  const int64_t L = table_len(M, w);
  for (int wrap = 0; wrap < 2; ++wrap) {
    if (wrap && m > 2 * (int64_t)w + 2) break;
    const int64_t mm = m + (int64_t)wrap * M;
    if (mm + 1 >= L) break;
    int64_t lo_c = mm - w + 1;
    if (lo_c < 0) lo_c = 0;
    const int32_t a = first_ge_entry(lo_c, cad, n), b = first_ge_entry(mm + 1, cad, n);
    for (int32_t i = a; i < b; ++i) {
      const double x = dfM * t[i] + shift;
      const float ph = es_eval_acc(((double)mm - x) * inv_half, beta);
      acc += y ? ph * y[i] : ph;
    }
  }
  return acc;
}

// ---- FFT ------------------------------------------------------------------------------------------
LKB_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

LKB_HD double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// complex value type of a transform: float2 (the batch path) or double2 (the escalation pass of nufft_v2.cuh)
template <class CT> struct CplxOf;
template <> struct CplxOf<float2> {
  typedef float real;
  static LKB_HD float2 mk(float a, float b) { return make_float2(a, b); }
};
template <> struct CplxOf<double2> {
  typedef double real;
  static LKB_HD double2 mk(double a, double b) { return make_double2(a, b); }
};

// exp(+2 pi i q / 16), q = 0 .. 7
template <class CT>
LKB_HD CT w16_t(int q) {
  typedef typename CplxOf<CT>::real RT;
  switch (q) {
    case 0: return CplxOf<CT>::mk((RT)1.0, (RT)0.0);
    case 1: return CplxOf<CT>::mk((RT)0.923879532511286756128183189397, (RT)0.382683432365089771728459984030);
    case 2: return CplxOf<CT>::mk((RT)0.707106781186547524400844362105, (RT)0.707106781186547524400844362105);
    case 3: return CplxOf<CT>::mk((RT)0.382683432365089771728459984030, (RT)0.923879532511286756128183189397);
    case 4: return CplxOf<CT>::mk((RT)0.0, (RT)1.0);
    case 5: return CplxOf<CT>::mk((RT)-0.382683432365089771728459984030, (RT)0.923879532511286756128183189397);
    case 6: return CplxOf<CT>::mk((RT)-0.707106781186547524400844362105, (RT)0.707106781186547524400844362105);
    default: return CplxOf<CT>::mk((RT)-0.923879532511286756128183189397, (RT)0.382683432365089771728459984030);
  }
}
LKB_HD float2 w16(int q) { return w16_t<float2>(q); }

// in-place DFT of R points with the +i sign, natural order in and out (decimation in time, recursive halves)
template <int R, class CT = float2>
struct SmallDft {
  static LKB_HD void run(CT* u) {
    CT e[R / 2], o[R / 2];
#pragma unroll
    for (int i = 0; i < R / 2; ++i) { e[i] = u[2 * i]; o[i] = u[2 * i + 1]; }
    SmallDft<R / 2, CT>::run(e);
    SmallDft<R / 2, CT>::run(o);
#pragma unroll
    for (int q = 0; q < R / 2; ++q) {
      const CT tw = cmul(o[q], w16_t<CT>(q * (16 / R)));
      u[q] = CplxOf<CT>::mk(e[q].x + tw.x, e[q].y + tw.y);
      u[q + R / 2] = CplxOf<CT>::mk(e[q].x - tw.x, e[q].y - tw.y);
    }
  }
};
template <class CT>
struct SmallDft<1, CT> {
  static LKB_HD void run(CT*) {}
};

LKB_HD float2 unit_phase(int64_t num, int64_t den) {     // exp(+2 pi i num / den), den a power of two, num < 2^24
  float s, c;
#if defined(__CUDA_ARCH__)
  sincospif(2.0f * (float)num / (float)den, &s, &c);
#else
  const double a = 6.283185307179586476925 * (double)num / (double)den;
  s = (float)sin(a);
  c = (float)cos(a);
#endif
  return make_float2(c, s);
}

// tw[r] = w1^r for r = 1 .. R-1 from ONE accurate evaluation (w1) by a product tree of depth <= 4 (squarings for the
// powers of two, then one product each) - ~4 FMA-pipe operations per twiddle instead of a sincospif.
template <int R>
LKB_HD void twiddle_powers(float2 w1, float2* tw) {
  tw[1] = w1;
#pragma unroll
  for (int r = 2; r < R; ++r) {
    const int hi = (r >= 8) ? 8 : (r >= 4) ? 4 : 2;          // largest power of two <= r (R <= 16)
    tw[r] = (r == hi) ? cmul(tw[r / 2], tw[r / 2]) : cmul(tw[hi], tw[r - hi]);
  }
}

// One butterfly of an out-of-place Stockham pass of radix R over a length-M transform.
//   i  in [0, M/R): butterfly index;  Ns = product of the radices of the earlier passes (1 for the first).
// After passes whose radices multiply to M the output is the DFT in natural order.
// CHAIN = false: every twiddle exp(2 pi i r k / (Ns R)) by its own sincospif (most accurate);
// CHAIN = true : one sincospif per butterfly + twiddle_powers (fewer instructions; measured on the CPU harness:
//                rms transform error at M = 2^19: 1.7e-7 -> 4.3e-7 of the rms output).
// Shared-memory arrays are addressed through skew(): one float2 of padding per 16 keeps the stride-R stores of the
// first passes off a single bank (without it thread i writes word 2 (R i + r): one bank for the whole warp).
LKB_HD int64_t skew(int64_t a) { return a + (a >> 4); }
LKB_HD int skew(int a) { return a + (a >> 4); }             // (32-bit index math inside the shared-memory tiles)

template <int R, bool CHAIN = false, bool SKEW = false>
LKB_HD void fft_pass_butterfly(const float2* x, float2* y, int64_t i, int64_t Ns, int64_t M) {
  const int64_t T = M / R;
  const int64_t k = i & (Ns - 1);
  float2 u[R];
#pragma unroll
  for (int r = 0; r < R; ++r) u[r] = x[SKEW ? skew(i + (int64_t)r * T) : i + (int64_t)r * T];
  if (Ns > 1) {
    if (CHAIN) {
      float2 tw[R];
      twiddle_powers<R>(unit_phase(k, Ns * R), tw);
#pragma unroll
      for (int r = 1; r < R; ++r) u[r] = cmul(u[r], tw[r]);
    } else {
#pragma unroll
      for (int r = 1; r < R; ++r) u[r] = cmul(u[r], unit_phase((int64_t)r * k, Ns * R));
    }
  }
  SmallDft<R>::run(u);
  const int64_t j = (i - k) * R + k;
#pragma unroll
  for (int r = 0; r < R; ++r) y[SKEW ? skew(j + (int64_t)r * Ns) : j + (int64_t)r * Ns] = u[r];
}

// ---- four-step transform (two global sweeps instead of one per radix pass) ---------------------------------------
// M = A * Bc, A = 2^pa.  Input index n = n1 Bc + n2, output index k = k1 + A k2:
//   X[k1 + A k2] = sum_n2 W_Bc^(n2 k2) * W_M^(n2 k1) * [ sum_n1 x[n1 Bc + n2] W_A^(n1 k1) ]
// step 1 ("columns"): length-A transforms over n1 for every n2, times W_M^(n2 k1), stored at [k1][n2];
// step 2 ("rows"):    length-Bc transforms over n2 for every k1, stored at [k1][k2]  (in place).
// Mode k therefore lives at address fourstep_index(k).
LKB_HD int fourstep_pa(int p) { return (p + 1) / 2; }
LKB_HD int64_t fourstep_index(int64_t k, int pa, int64_t Bc) { return (k & (((int64_t)1 << pa) - 1)) * Bc + (k >> pa); }
// float2 elements of one skewed shared-memory line of n elements
LKB_HD int64_t smem_line(int64_t n) { return skew(n) + 1; }

// radix of pass `idx` (0-based) for a length-2^p transform: sixteens first, the remainder last.  0 = done.
LKB_HD int fft_pass_radix(int p, int idx) {
  const int n16 = p / 4, rem = p % 4;
  if (idx < n16) return 16;
  if (idx == n16 && rem) return 1 << rem;
  return 0;
}

// ---- "v2" four-step transform: pruned, tiled, table twiddles (ls_nufft.cu nufft2_* kernels) ---------------------
// M = A * Bc with Bc = 2^V2_PB fixed and A = 2^pa, n = n1 Bc + n2, k = k1 + A k2 as above.  One CTA = V2_THREADS
// threads working in place on V2_TILE = 16 * V2_THREADS points of shared memory: every radix-16 pass is exactly one
// butterfly per thread (read 16 points, barrier, twiddle + DFT + write, barrier).
//   * the spreading writes only the rows n1 < n1max that cadences can reach (df * baseline of the fine grid:
//     20 % at lightkurve's default oversampling), in the layout the column kernel reads contiguously:
//     G[pair][c][n1][j], c = n2 / tc, j = n2 % tc, tc = V2_TILE / A columns per CTA;
//   * the column kernel (length-A transforms) writes T[pair][c][k1][j] - again one contiguous block per CTA;
//   * the row kernel takes R = V2_TILE / (2 Bc) rows k1 = 1 + g R .. (g + 1) R together with their mirror rows
//     A - k1 (mode k pairs with mode M - k when two real series share one complex transform), transforms them
//     (length Bc) and finishes: unpack, deconvolve, epilogue -> power.  R consecutive k1 at one k2 are R consecutive
//     frequency bins, so the power rows are written in R * 4-byte runs.  The transform itself never goes back to
//     global memory.  The CTA that would hold row A / 2 twice takes row 0 (which pairs with itself) instead.
// Twiddles come from tables built once per call in fp64 (no sincospif in the butterflies): per pass a table
// [r][k] = exp(2 pi i r k / (Ns R)) (k fastest: consecutive lanes read consecutive entries), and for the inter-step
// factor exp(2 pi i n2 k1 / M) a two-level table (hi * lo).
constexpr int V2_THREADS = 512;
constexpr int V2_TILE = 16 * V2_THREADS;        // 8192 points = 64 KB of float2
constexpr int V2_PB = 9;                        // Bc = 512
constexpr int V2_P_MIN = V2_PB + 4, V2_P_MAX = V2_PB + 13;      // 16 <= A <= 8192

// entries of the pass tables of a length-2^q transform (the first pass has no twiddles and no table)
LKB_HD int v2_pass_table_len(int q) {
  int len = 0, Ns = 1;
  for (int idx = 0;; ++idx) {
    const int R = fft_pass_radix(q, idx);
    if (R == 0) break;
    if (idx > 0) len += R * Ns;
    Ns *= R;
  }
  return len;
}
// entry e of those tables (e in [0, v2_pass_table_len(q))): exp(2 pi i r k / (Ns R)) as (num, den)
LKB_HD void v2_pass_table_entry(int q, int e, int64_t* num, int64_t* den) {
  int Ns = 1;
  for (int idx = 0;; ++idx) {
    const int R = fft_pass_radix(q, idx);
    if (R == 0) break;
    if (idx > 0) {
      if (e < R * Ns) { *num = (int64_t)(e / Ns) * (e % Ns); *den = (int64_t)Ns * R; return; }
      e -= R * Ns;
    }
    Ns *= R;
  }
  *num = 0; *den = 1;
}
LKB_HD int v2_log2_lo(int p) { return (p + 1) / 2; }          // the inter-step table splits q = hi 2^pl + lo

// ---- unpacking --------------------------------------------------------------------------------------
// Gauss-Legendre nodes and weights on [-1, 1] (Newton iteration on P_n; n <= 64)
LKB_HD void gauss_legendre(int n, double* x, double* w) {
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double z = cos(3.14159265358979323846 * ((double)i + 0.75) / ((double)n + 0.5));
    double pp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) {
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j + 1.0) * z * p2 - (double)j * p3) / ((double)j + 1.0);
      }
      pp = (double)n * (z * p1 - p2) / (z * z - 1.0);
      const double dz = p1 / pp;
      z -= dz;
      if (fabs(dz) < 1e-15) break;
    }
    x[i] = -z;
    x[n - 1 - i] = z;
    w[i] = w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
}

// Fourier coefficient of the kernel at mode kk on an M-cell grid, times the grid-shift phase, INVERTED:
// the factor the raw transform value of mode kk is multiplied with.  glx/glw: Gauss-Legendre nodes/weights on
// [-1, 1] (nq of them; 48 are ample for w <= 16).
LKB_HD void deconv_factor(int64_t kk, int64_t M, int w, double beta, const double* glx, const double* glw, int nq,
                          double* re, double* im) {
  const double half = 0.5 * (double)w;                         // kernel half-width in cells
  const double omega = 6.283185307179586476925 * (double)kk / (double)M;   // radians per cell
  double acc = 0.0;
  for (int q = 0; q < nq; ++q)
    acc += glw[q] * exp(beta * (sqrt(1.0 - glx[q] * glx[q]) - 1.0)) * cos(omega * half * glx[q]);
  const double phihat = half * acc;                            // sum over cells ~ integral in cells
  const double ang = -omega * (double)grid_shift(w);
  *re = cos(ang) / phihat;
  *im = sin(ang) / phihat;
}

// (C + i S) of the two light curves of a pair at mode kk from the packed transform Z (length M); inv0 / inv1 undo
// the pow2_scale factors of the two light curves
// pa > 0: Z is in the four-step layout (fourstep_index), pa = 0: natural order
LKB_HD void unpack_pair(const float2* Z, int64_t kk, int64_t M, float2 dec, float inv0, float inv1, float2* a,
                        float2* b, int pa = 0) {
  const int64_t km = (kk == 0) ? 0 : M - kk;
  const float2 g1 = Z[pa ? fourstep_index(kk, pa, M >> pa) : kk];
  const float2 g2 = Z[pa ? fourstep_index(km, pa, M >> pa) : km];
  const float2 ra = make_float2(0.5f * (g1.x + g2.x), 0.5f * (g1.y - g2.y));      // (g1 + conj g2) / 2
  const float2 rb = make_float2(0.5f * (g1.y + g2.y), 0.5f * (g2.x - g1.x));      // (g1 - conj g2) / 2i
  const float2 da = cmul(ra, dec), db = cmul(rb, dec);
  *a = make_float2(da.x * inv0, da.y * inv0);
  *b = make_float2(db.x * inv1, db.y * inv1);
}

}  // namespace nufft
}  // namespace lkb
