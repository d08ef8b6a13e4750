// liblkb200 - shared-grid Lomb-Scargle by a type-1 NUFFT (LKB_LS_ALGO_NUFFT; OPT-IN in round 1: written and
// verified on the CPU through tests/native/nufft_host_harness.cpp after the round's GPU budget was spent -
// `auto` never selects it until it has been measured on hardware).
//
// Why: the contraction kernels (ls_tc.cu, ls.cu) do 4 N F flops per light curve; on a REGULAR frequency grid
// the same trig sums are the Fourier coefficients of the (non-uniformly sampled) light curve, which a
// spreading step + one FFT give in N w + 2.5 M log2 M flops (config 2: 5e7 instead of 2.6e10 per light curve),
// within 0.02-0.16 of the parity tolerance in fp32 (tools/nufft_ls_model.py) - more accurate than the
// split-fp16 tensor path (0.47).  The whole batch becomes an HBM sweep: fine grids [B/2, M] complex64.
// This is the algorithm behind the reference's optional ls_method="fastnifty" (nifty-ls / finufft,
// /root/reference/pyproject.toml:48, src/lightkurve/periodogram.py:917-946).
//
// Kernels (all "one thread = one function of nufft_core.h"):
//   nufft_cad_kernel        per cadence: leftmost cell + offset of its kernel support on the M-cell grid
//   nufft_first_ge_kernel   per cell: first cadence whose support starts at or after it (binary search)
//   nufft_spread_kernel     per (cell, pair of light curves): gather of the cadences reaching the cell
//   nufft_fft_pass_kernel   per butterfly: out-of-place Stockham pass of radix 16/8/4/2
//   nufft_deconv_kernel     per mode: 1 / phihat and the grid-shift phase
//   nufft_rot_kernel        per frequency: window terms (tau rotation, 1/CC', 1/SS') from the transform of
//                           a_n = 1 on a grid twice as fine (modes kk and 2 kk)
//   nufft_lowrows_kernel    per (low frequency, light curve): direct fp32/fp64 sums with the cos-1 design matrix
//                           for f * baseline <= LS_LOWF_CYCLES (the rows whose sums cancel)
//   nufft_finish_kernel     per (frequency, pair): unpack the two light curves, deconvolve, epilogue -> power
// HBM traffic at config 2 (B = 1024, M = 2^19): spread 2.1 GB written, 5 passes x 4.3 GB, finish ~1 GB read +
// 0.4 GB written  ~ 25 GB  ~ 4 ms at the measured 6.6 TB/s.  Next steps once measured: fuse the spreading into
// the first pass (80 % of its input is zero for oversample 5), shared-memory passes (2 instead of 5 sweeps).
#include <algorithm>

#include "common.cuh"
#include "ls_common.cuh"
#include "nufft_core.h"

namespace lkb {

using nufft::Cad;

namespace {

struct GlNodes {
  double x[32], w[32];
};

__global__ void nufft_cad_kernel(const double* __restrict__ t, int64_t N, double df, int64_t M, int w,
                                 Cad* __restrict__ cad, int* __restrict__ unsorted) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  cad[n] = nufft::cad_entry(t[n], df, M, w);
  if (n > 0 && t[n] < t[n - 1]) *unsorted = 1;
  if (t[n] < 0.0) *unsorted = 1;
}

__global__ void nufft_first_ge_kernel(const Cad* __restrict__ cad, int64_t N, int64_t L, int32_t* __restrict__ first_ge) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < L) first_ge[c] = nufft::first_ge_entry(c, cad, N);
}

// Z[pair][m] = sum over cadences of phi * (y[2 pair][n] + i y[2 pair + 1][n])
__global__ void __launch_bounds__(256)
nufft_spread_kernel(const int32_t* __restrict__ first_ge, const Cad* __restrict__ cad, const float* __restrict__ y,
                    int64_t ystride, const float* __restrict__ absmax, int B, int npairs, int w, float beta, int log2M,
                    float2* __restrict__ Z) {
  const int64_t M = (int64_t)1 << log2M;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)npairs << log2M) return;
  const int64_t pair = gid >> log2M, m = gid & (M - 1);
  const float* y0 = y + (2 * pair) * ystride;
  const float* y1 = (2 * pair + 1 < B) ? y0 + ystride : nullptr;
  const float s0 = absmax ? nufft::pow2_scale(absmax[2 * pair]) : 1.0f;
  const float s1 = (absmax && y1) ? nufft::pow2_scale(absmax[2 * pair + 1]) : 1.0f;
  Z[gid] = nufft::spread_cell(m, first_ge, cad, y0, y1, s0, s1, w, beta, M);
}

template <int R, bool CHAIN>
__global__ void __launch_bounds__(256)
nufft_fft_pass_kernel(const float2* __restrict__ x, float2* __restrict__ y, int64_t Ns, int log2M, int64_t total) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int64_t M = (int64_t)1 << log2M, per = M / R;
  const int64_t pair = gid / per, i = gid - pair * per;
  nufft::fft_pass_butterfly<R, CHAIN>(x + pair * M, y + pair * M, i, Ns, M);
}

template <bool CHAIN>
void launch_pass(int R, unsigned g, const float2* src, float2* dst, int64_t Ns, int p, int64_t total, cudaStream_t st) {
  if (R == 16) LKB_LAUNCH(g, 256, st, nufft_fft_pass_kernel<16, CHAIN>)(src, dst, Ns, p, total);
  else if (R == 8) LKB_LAUNCH(g, 256, st, nufft_fft_pass_kernel<8, CHAIN>)(src, dst, Ns, p, total);
  else if (R == 4) LKB_LAUNCH(g, 256, st, nufft_fft_pass_kernel<4, CHAIN>)(src, dst, Ns, p, total);
  else LKB_LAUNCH(g, 256, st, nufft_fft_pass_kernel<2, CHAIN>)(src, dst, Ns, p, total);
}

// ---- four-step transform with shared-memory sub-transforms (LKB_NUFFT_FFT=smem): two in-place global sweeps ----
// nufft_core.h "four-step transform".  One CTA = a tile of TC columns (step 1) or TR rows (step 2) of one transform
// in two skewed shared-memory buffers; the radix passes are the same butterflies as the global version.
constexpr int FS_THREADS = 256;
// float2 elements per shared-memory buffer (before skew padding): 4096 -> 2 x 34 KB per CTA, 3 CTAs per SM;
// LKB_NUFFT_TILE overrides (power of two, up to 8192)
inline int64_t fs_tile() {
  int64_t v = 4096;
  if (const char* e = getenv("LKB_NUFFT_TILE")) v = atoll(e);
  if (v < 256) v = 256;
  if (v > 8192) v = 8192;
  int64_t p2 = 256;
  while (p2 * 2 <= v) p2 *= 2;
  return p2;
}

template <bool CHAIN>
__device__ __forceinline__ void smem_line_passes(float2*& src, float2*& dst, int lines, int64_t line_stride, int plog2) {
  const int64_t n = (int64_t)1 << plog2;
  int64_t Ns = 1;
  for (int idx = 0;; ++idx) {
    const int R = nufft::fft_pass_radix(plog2, idx);
    if (R == 0) break;
    const int64_t nb = n / R;
    for (int64_t j = threadIdx.x; j < (int64_t)lines * nb; j += blockDim.x) {
      const int64_t c = j / nb, i = j - c * nb;
      const float2* x = src + c * line_stride;
      float2* y = dst + c * line_stride;
      if (R == 16) nufft::fft_pass_butterfly<16, CHAIN, true>(x, y, i, Ns, n);
      else if (R == 8) nufft::fft_pass_butterfly<8, CHAIN, true>(x, y, i, Ns, n);
      else if (R == 4) nufft::fft_pass_butterfly<4, CHAIN, true>(x, y, i, Ns, n);
      else nufft::fft_pass_butterfly<2, CHAIN, true>(x, y, i, Ns, n);
    }
    __syncthreads();
    Ns *= R;
    float2* tmp = src; src = dst; dst = tmp;
  }
}

// step 1: grid (Bc / TC, npairs)
// what the column kernel needs to compute its input cells itself (FUSED: the spreading never touches global memory)
struct SpreadArgs {
  const int32_t* first_ge;
  const Cad* cad;
  const float* y;
  int64_t ystride;
  const float* absmax;
  int B, w;
  float beta;
};

template <bool CHAIN, bool FUSED>
__global__ void __launch_bounds__(FS_THREADS)
nufft_fft_cols_kernel(float2* __restrict__ Z, int p, int pa, int tc, SpreadArgs sp) {
  LKB_DYN_SMEM(float2, smem);
  const int64_t M = (int64_t)1 << p, A = (int64_t)1 << pa, Bc = M >> pa;
  const int64_t stride = nufft::smem_line(A);
  const int64_t pair = blockIdx.y;
  float2* Zp = Z + pair * M;
  const int64_t c0 = (int64_t)blockIdx.x * tc;
  float2 *src = smem, *dst = smem + (int64_t)tc * stride;
  if (FUSED) {
    const float* y0 = sp.y + (2 * pair) * sp.ystride;
    const float* y1 = (2 * pair + 1 < sp.B) ? y0 + sp.ystride : nullptr;
    const float s0 = nufft::pow2_scale(sp.absmax[2 * pair]);
    const float s1 = y1 ? nufft::pow2_scale(sp.absmax[2 * pair + 1]) : 1.0f;
    for (int64_t idx = threadIdx.x; idx < (int64_t)tc * A; idx += blockDim.x) {
      const int64_t c = idx % tc, n1 = idx / tc;
      src[c * stride + nufft::skew(n1)] =
          nufft::spread_cell(n1 * Bc + c0 + c, sp.first_ge, sp.cad, y0, y1, s0, s1, sp.w, sp.beta, M);
    }
  } else {
    for (int64_t idx = threadIdx.x; idx < (int64_t)tc * A; idx += blockDim.x) {
      const int64_t c = idx % tc, n1 = idx / tc;
      src[c * stride + nufft::skew(n1)] = Zp[n1 * Bc + c0 + c];
    }
  }
  __syncthreads();
  smem_line_passes<CHAIN>(src, dst, tc, stride, pa);
  for (int64_t idx = threadIdx.x; idx < (int64_t)tc * A; idx += blockDim.x) {
    const int64_t c = idx % tc, k1 = idx / tc, n2 = c0 + c;
    Zp[k1 * Bc + n2] = nufft::cmul(src[c * stride + nufft::skew(k1)], nufft::unit_phase(n2 * k1, M));
  }
}

// step 2: grid (A / TR, npairs)
template <bool CHAIN>
__global__ void __launch_bounds__(FS_THREADS)
nufft_fft_rows_kernel(float2* __restrict__ Z, int p, int pa, int tr) {
  LKB_DYN_SMEM(float2, smem);
  const int64_t M = (int64_t)1 << p, Bc = M >> pa;
  const int pb = p - pa;
  const int64_t stride = nufft::smem_line(Bc);
  float2* Zp = Z + (int64_t)blockIdx.y * M + (int64_t)blockIdx.x * tr * Bc;
  float2 *src = smem, *dst = smem + (int64_t)tr * stride;
  for (int64_t idx = threadIdx.x; idx < (int64_t)tr * Bc; idx += blockDim.x) {
    const int64_t r = idx / Bc, n2 = idx - r * Bc;
    src[r * stride + nufft::skew(n2)] = Zp[idx];
  }
  __syncthreads();
  smem_line_passes<CHAIN>(src, dst, tr, stride, pb);
  for (int64_t idx = threadIdx.x; idx < (int64_t)tr * Bc; idx += blockDim.x) {
    const int64_t r = idx / Bc, k2 = idx - r * Bc;
    Zp[idx] = src[r * stride + nufft::skew(k2)];
  }
}

// ---- "v2" transform (nufft_core.h): pruned spreading in the column kernel's layout, tiled column transforms with
// table twiddles, row transforms fused with the finish.  Global traffic per pair of light curves at config 2
// (M = 2^19): 0.5 MB flux + 0.84 MB G written + read, 4.2 MB T written + read, 0.8 MB power  = 11.4 MB
// (the five global radix passes: 4.2 MB spread + 5 x 8.4 MB + 1.6 MB unpack reads + 0.8 MB = 48.6 MB).
using nufft::V2_PB;
using nufft::V2_THREADS;
using nufft::V2_TILE;
constexpr int V2_LOG_TILE = 13;
static_assert((1 << V2_LOG_TILE) == V2_TILE, "tile size");

__global__ void nufft2_tables_kernel(int pa, int pb, int p, float2* __restrict__ tw_a, float2* __restrict__ tw_b,
                                     float2* __restrict__ t_hi, float2* __restrict__ t_lo) {
  const int la = nufft::v2_pass_table_len(pa), lb = nufft::v2_pass_table_len(pb), pl = nufft::v2_log2_lo(p);
  const int nlo = 1 << pl, nhi = 1 << (p - pl);
  int e = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  int64_t num = 0, den = 1;
  float2* dst = nullptr;
  if (e < la) { nufft::v2_pass_table_entry(pa, e, &num, &den); dst = tw_a + e; }
  else if ((e -= la) < lb) { nufft::v2_pass_table_entry(pb, e, &num, &den); dst = tw_b + e; }
  else if ((e -= lb) < nhi) { num = e; den = nhi; dst = t_hi + e; }
  else if ((e -= nhi) < nlo) { num = e; den = (int64_t)1 << p; dst = t_lo + e; }
  else return;
  double sn, cs;
  sincospi(2.0 * (double)num / (double)den, &sn, &cs);
  *dst = make_float2((float)cs, (float)sn);
}

// fine-grid cell m of position e of the G layout [c][n1][j]
__device__ __forceinline__ int64_t v2_cell_of(int64_t e, int ptc, int n1max) {
  const int64_t j = e & (((int64_t)1 << ptc) - 1), rest = e >> ptc;
  const int64_t n1 = rest % n1max, c = rest / n1max;
  return (n1 << V2_PB) + (c << ptc) + j;
}

// G[pair][e] for PP pairs of light curves per thread (the kernel weight of a (cell, cadence) is computed once and
// applied to 2 PP light curves)
template <int PP>
__global__ void __launch_bounds__(256)
nufft2_spread_kernel(const int32_t* __restrict__ first_ge, const Cad* __restrict__ cad, const float* __restrict__ y,
                     int64_t ystride, const float* __restrict__ absmax, int B, int npairs, int w, float beta, int p,
                     int ptc, int n1max, float2* __restrict__ G) {
  const int64_t cells = (int64_t)n1max << V2_PB;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cells) return;
  const int64_t M = (int64_t)1 << p, m = v2_cell_of(e, ptc, n1max);
  const int pair0 = (int)blockIdx.y * PP;
  const float* yr[2 * PP];
#pragma unroll
  for (int q = 0; q < 2 * PP; ++q) {
    int b = 2 * pair0 + q;
    if (b > B - 1) b = B - 1;                      // clamped rows are computed and dropped
    yr[q] = y + (int64_t)b * ystride;
  }
  float acc[2 * PP];
#pragma unroll
  for (int q = 0; q < 2 * PP; ++q) acc[q] = 0.0f;
  const float inv_half = 2.0f / (float)w;
  const int64_t L = nufft::table_len(M, w);
  for (int wrap = 0; wrap < 2; ++wrap) {           // wrap = 1: cadences whose support runs past cell M - 1
    const int64_t mm = m + (int64_t)wrap * M;
    if (mm + 1 >= L) break;
    int64_t lo_c = mm - w + 1;
    if (lo_c < 0) lo_c = 0;
    const int32_t a = first_ge[lo_c], b = first_ge[mm + 1];
    for (int32_t n = a; n < b; ++n) {
      const Cad cd = cad[n];
      const float ph = nufft::es_eval((cd.d0 + (float)(mm - (int64_t)cd.i0)) * inv_half, beta);
#pragma unroll
      for (int q = 0; q < 2 * PP; ++q) acc[q] = fmaf(ph, yr[q][n], acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < PP; ++q) {
    const int pair = pair0 + q;
    if (pair >= npairs) break;
    const int b0 = 2 * pair;
    const float s0 = nufft::pow2_scale(absmax[b0]);
    const float s1 = (b0 + 1 < B) ? nufft::pow2_scale(absmax[b0 + 1]) : 0.0f;
    G[(int64_t)pair * cells + e] = make_float2(acc[2 * q] * s0, acc[2 * q + 1] * s1);
  }
}

// one in-place pass of radix R over the lines of a tile (16 points per thread)
template <int R>
__device__ __forceinline__ void v2_pass(float2* buf, int plog, int lstride, int Ns, const float2* __restrict__ tw) {
  constexpr int NB = 16 / R;
  constexpr int LR = (R == 16) ? 4 : (R == 8) ? 3 : (R == 4) ? 2 : 1;
  const int pnb = plog - LR, nb = 1 << pnb;            // butterflies per line
  float2 u[NB][R];
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int b = (int)threadIdx.x + V2_THREADS * q;
    const int line = b >> pnb, i = b & (nb - 1);
    const float2* x = buf + line * lstride;
#pragma unroll
    for (int r = 0; r < R; ++r) u[q][r] = x[nufft::skew(i + r * nb)];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int b = (int)threadIdx.x + V2_THREADS * q;
    const int line = b >> pnb, i = b & (nb - 1), k = i & (Ns - 1);
    float2* x = buf + line * lstride;
    if (Ns > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) u[q][r] = nufft::cmul(u[q][r], tw[r * Ns + k]);
    }
    nufft::SmallDft<R>::run(u[q]);
    const int j = ((i - k) << LR) + k;
#pragma unroll
    for (int r = 0; r < R; ++r) x[nufft::skew(j + r * Ns)] = u[q][r];
  }
  __syncthreads();
}

// all passes of the length-2^plog transforms of the tile's V2_TILE >> plog lines; tw: nufft::v2_pass_table_*
__device__ __forceinline__ void v2_fft_lines(float2* buf, int plog, int lstride, const float2* __restrict__ tw) {
  int Ns = 1;
  for (int idx = 0;; ++idx) {
    const int R = nufft::fft_pass_radix(plog, idx);
    if (R == 0) break;
    if (R == 16) v2_pass<16>(buf, plog, lstride, Ns, tw);
    else if (R == 8) v2_pass<8>(buf, plog, lstride, Ns, tw);
    else if (R == 4) v2_pass<4>(buf, plog, lstride, Ns, tw);
    else v2_pass<2>(buf, plog, lstride, Ns, tw);
    if (idx > 0) tw += R * Ns;
    Ns *= R;
  }
}

// step 1: grid (Bc / tc, npairs): tc columns n2 = c tc + j of one pair, length-A transforms over n1
__global__ void __launch_bounds__(V2_THREADS, 2)
nufft2_cols_kernel(const float2* __restrict__ G, float2* __restrict__ T, int p, int n1max,
                   const float2* __restrict__ tw_a, const float2* __restrict__ t_hi, const float2* __restrict__ t_lo) {
  LKB_DYN_SMEM(float2, buf);
  const int pa = p - V2_PB, ptc = V2_LOG_TILE - pa, tc = 1 << ptc, A = 1 << pa;
  const int lstride = (int)nufft::smem_line(A);
  const int c = (int)blockIdx.x, C = (1 << V2_PB) >> ptc;
  const int64_t pair = blockIdx.y;
  const int nvalid = n1max << ptc;
  const float2* Gp = G + (pair * C + c) * (int64_t)nvalid;
  for (int idx = (int)threadIdx.x; idx < V2_TILE; idx += V2_THREADS) {
    const int n1 = idx >> ptc, j = idx & (tc - 1);
    buf[j * lstride + (int)nufft::skew(n1)] = (idx < nvalid) ? Gp[idx] : make_float2(0.0f, 0.0f);
  }
  __syncthreads();
  v2_fft_lines(buf, pa, lstride, tw_a);
  float2* Tp = T + (pair * C + c) * (int64_t)V2_TILE;
  const int pl = nufft::v2_log2_lo(p);
  const int64_t Mmask = ((int64_t)1 << p) - 1;
  for (int idx = (int)threadIdx.x; idx < V2_TILE; idx += V2_THREADS) {
    const int k1 = idx >> ptc, j = idx & (tc - 1);
    const int64_t q = ((int64_t)((c << ptc) + j) * k1) & Mmask;
    const float2 wq = nufft::cmul(t_hi[q >> pl], t_lo[q & (((int64_t)1 << pl) - 1)]);
    Tp[idx] = nufft::cmul(buf[j * lstride + (int)nufft::skew(k1)], wq);
  }
}

struct V2Finish {
  const float2* dec;
  int64_t k0, F, k_lo;
  const float4* rot;
  const float2* rot2;
  const float* ysum;
  const float* absmax;
  float Nf;
  int normalization;
  float scale;
  int B;
  float* power;
};

// step 2: grid (A / (2 R), npairs): rows k1 = 1 + g R .. (g + 1) R and their mirror rows, length-Bc transforms over n2;
// MODE 1: unpack / deconvolve / epilogue -> power;  MODE 0: the whole transform goes to Zout in the [k1][k2] layout
// (nufft::fourstep_index);  MODE 2: the modes k < nk2_keep * A and their mirrors M - k go to Zout in NATURAL order
// (what the ragged finish kernel reads; R consecutive k1 are R consecutive modes: 64-byte runs).
template <int MODE>
__global__ void __launch_bounds__(V2_THREADS, 2)
nufft2_rows_kernel(const float2* __restrict__ T, int p, const float2* __restrict__ tw_b, V2Finish fa,
                   float2* __restrict__ Zout, int nk2_keep) {
  LKB_DYN_SMEM(float2, buf);
  constexpr int pb = V2_PB, Bc = 1 << pb, pR = V2_LOG_TILE - 1 - pb, R = 1 << pR;
  const int pa = p - pb, A = 1 << pa, ptc = V2_LOG_TILE - pa, tc = 1 << ptc;
  const int lstride = (int)nufft::smem_line(Bc);
  const int g = (int)blockIdx.x;
  const bool last = g == (A >> (pR + 1)) - 1;
  const int64_t pair = blockIdx.y, M = (int64_t)1 << p;
  auto slot_k1 = [&](int s) -> int {
    const int h = s >> pR, r = s & (R - 1);
    if (h == 0) return 1 + g * R + r;
    if (last && r == 0) return 0;                    // instead of a second copy of row A / 2
    return A - (g + 1) * R + r;
  };
  const float2* Tp = T + pair * M;
  for (int e = (int)threadIdx.x; e < V2_TILE; e += V2_THREADS) {
    const int j = e & (tc - 1), r = (e >> ptc) & (R - 1), h = (e >> (ptc + pR)) & 1, c = e >> (ptc + pR + 1);
    const int s = h * R + r;
    buf[s * lstride + (int)nufft::skew((c << ptc) + j)] = Tp[((((int64_t)c << pa) + slot_k1(s)) << ptc) + j];
  }
  __syncthreads();
  v2_fft_lines(buf, pb, lstride, tw_b);
  if (MODE == 0) {
    for (int e = (int)threadIdx.x; e < V2_TILE; e += V2_THREADS) {
      const int s = e >> pb, k2 = e & (Bc - 1);
      Zout[pair * M + ((int64_t)slot_k1(s) << pb) + k2] = buf[s * lstride + (int)nufft::skew(k2)];
    }
    return;
  }
  if (MODE == 2) {
    const int keep = nk2_keep < Bc / 2 ? nk2_keep : Bc / 2;
    for (int item = (int)threadIdx.x; item < 2 * keep * 2 * R; item += V2_THREADS) {
      const int s = item & (2 * R - 1), q = item >> (pR + 1);
      const int k2 = q < keep ? q : Bc - 2 * keep + q;          // [0, keep) and [Bc - keep, Bc)
      Zout[pair * M + (int64_t)slot_k1(s) + ((int64_t)k2 << pa)] = buf[s * lstride + (int)nufft::skew(k2)];
    }
    return;
  }
  int64_t nK2 = ((fa.k0 + fa.F - 1) >> pa) + 1;
  if (nK2 > Bc) nK2 = Bc;
  const int64_t b0 = 2 * pair;
  const bool has1 = b0 + 1 < fa.B;
  const float inv0 = 1.0f / nufft::pow2_scale(fa.absmax[b0]);
  const float inv1 = has1 ? 1.0f / nufft::pow2_scale(fa.absmax[b0 + 1]) : 1.0f;
  const float ys0 = fa.ysum[b0], ys1 = has1 ? fa.ysum[b0 + 1] : 0.0f;
  for (int item = (int)threadIdx.x; item < (int)nK2 * 2 * R; item += V2_THREADS) {
    const int s = item & (2 * R - 1), k2 = item >> (pR + 1);
    const int64_t jj = (int64_t)slot_k1(s) + ((int64_t)k2 << pa) - fa.k0;
    if (jj < fa.k_lo || jj >= fa.F) continue;
    const int h = s >> pR, r = s & (R - 1);
    int ps = (1 - h) * R + (R - 1 - r), pi = Bc - 1 - k2;      // mode M - k: row A - k1, column Bc - 1 - k2
    if (last && h == 0 && r == R - 1) ps = s;                  // row A / 2 mirrors onto itself
    if (last && h == 1 && r == 0) { ps = s; pi = (Bc - k2) & (Bc - 1); }   // row 0: column Bc - k2
    const float2 g1 = buf[s * lstride + (int)nufft::skew(k2)], g2 = buf[ps * lstride + (int)nufft::skew(pi)];
    const float2 ra = make_float2(0.5f * (g1.x + g2.x), 0.5f * (g1.y - g2.y));      // (g1 + conj g2) / 2
    const float2 rb = make_float2(0.5f * (g1.y + g2.y), 0.5f * (g2.x - g1.x));      // (g1 - conj g2) / 2i
    const float2 dc = fa.dec[jj];
    const float2 da = nufft::cmul(ra, dc), db = nufft::cmul(rb, dc);
    const float4 rt = fa.rot[jj];
    const float2 r2 = fa.rot2[jj];
    fa.power[b0 * fa.F + jj] = ls_epilogue_shared(da.x * inv0, da.y * inv0, rt, r2, ys0, fa.Nf, fa.normalization, fa.scale);
    if (has1)
      fa.power[(b0 + 1) * fa.F + jj] = ls_epilogue_shared(db.x * inv1, db.y * inv1, rt, r2, ys1, fa.Nf, fa.normalization, fa.scale);
  }
}

__global__ void nufft_deconv_kernel(int64_t k_first, int64_t count, int64_t M, int w, double beta, GlNodes gl,
                                    float2* __restrict__ dec) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  double re, im;
  nufft::deconv_factor(k_first + k, M, w, beta, gl.x, gl.w, 32, &re, &im);
  dec[k] = make_float2((float)re, (float)im);
}

// window terms of the rows k >= k_lo from the transform Zw (length M2) of unit strengths: mode kk gives
// (C, S) = sum (cos, sin)(2 pi f t), mode 2 kk gives (C2, S2) = sum (cos, sin)(4 pi f t).
// dec2[j] is the deconvolution factor of mode j (j = 0 .. 2 (k0 + F) - 1).
__global__ void nufft_rot_kernel(const float2* __restrict__ Zw, int64_t M2, const float2* __restrict__ dec2, int64_t k0,
                                 int64_t F, int64_t k_lo, double Nd, float4* __restrict__ rot, float2* __restrict__ rot2) {
  const int64_t k = k_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= F) return;
  const int64_t kk = k0 + k;
  float2 a, unused;
  nufft::unpack_pair(Zw, kk, M2, dec2[kk], 1.0f, 1.0f, &a, &unused);
  float2 a2;
  nufft::unpack_pair(Zw, 2 * kk, M2, dec2[2 * kk], 1.0f, 1.0f, &a2, &unused);
  LsSums<double> d;
  d.zero();
  d.c = (double)a.x;
  d.s = (double)a.y;
  d.cc = 0.5 * (Nd + (double)a2.x);       // sum cos^2 = (N + sum cos 2wt) / 2
  d.sc = 0.5 * (double)a2.y;              // sum sin cos = sum sin 2wt / 2
  double ct, st, cc, ss;
  ls_rotation(d, Nd, ct, st, cc, ss);
  const double kf = 1.0 / (2.0 * Nd);
  rot[k] = make_float4((float)ct, (float)st, (float)(kf / cc), (float)(kf / ss));
  rot2[k] = make_float2((float)((d.c * ct + d.s * st) / Nd), (float)((d.s * ct - d.c * st) / Nd));
}

// rows with f * baseline <= LS_LOWF_CYCLES: direct sums, one warp per (row, light curve); the design matrix
// holds cos - 1 (ls_common.cuh) and rot / rot2 of these rows come from the fp64 path of ls_window_kernel.
__global__ void __launch_bounds__(128)
nufft_lowrows_kernel(const double* __restrict__ t, int64_t N, const float* __restrict__ yc, int64_t ystride, int B,
                     const double* __restrict__ freq, int64_t F_low, int64_t F, const float4* __restrict__ rot,
                     const float2* __restrict__ rot2, const float* __restrict__ ysum, int normalization, float scale,
                     float* __restrict__ power) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t job = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (job >= F_low * B) return;
  const int64_t b = job / F_low, k = job - b * F_low;
  const double fr = freq[k];
  const float* y = yc + b * ystride;
  double ch = 0.0, sh = 0.0;
  for (int64_t c0 = 0; c0 < N; c0 += 32 * 64) {
    float pc = 0.f, ps = 0.f;
    const int64_t c1 = min(N, c0 + (int64_t)32 * 64);
    for (int64_t i = c0 + lane; i < c1; i += 32) {
      float s, cm1;
      ls_sincos_cycles_low(fr * t[i], s, cm1);
      const float v = y[i];
      pc = fmaf(v, cm1, pc);
      ps = fmaf(v, s, ps);
    }
    ch += (double)pc;
    sh += (double)ps;
  }
  ch = warp_sum(ch);
  sh = warp_sum(sh);
  if (lane == 0)
    power[b * F + k] = ls_epilogue_shared((float)ch, (float)sh, rot[k], rot2[k], ysum[b], (float)N, normalization,
                                          scale, true);
}

__global__ void __launch_bounds__(256)
nufft_finish_kernel(const float2* __restrict__ Z, int log2M, const float2* __restrict__ dec, int64_t k0, int64_t F,
                    int64_t k_lo, const float4* __restrict__ rot, const float2* __restrict__ rot2,
                    const float* __restrict__ ysum, const float* __restrict__ absmax, float Nf, int normalization,
                    float scale, int B, int npairs, int pa, float* __restrict__ power) {
  const int64_t nk = F - k_lo;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= nk * npairs) return;
  const int64_t pair = gid / nk, k = k_lo + (gid - pair * nk);
  const int64_t M = (int64_t)1 << log2M;
  const int64_t b0 = 2 * pair;
  const float inv0 = 1.0f / nufft::pow2_scale(absmax[b0]);
  const float inv1 = (b0 + 1 < B) ? 1.0f / nufft::pow2_scale(absmax[b0 + 1]) : 1.0f;
  float2 a, b;
  nufft::unpack_pair(Z + pair * M, k0 + k, M, dec[k], inv0, inv1, &a, &b, pa);
  const float4 r = rot[k];
  const float2 r2 = rot2[k];
  power[b0 * F + k] = ls_epilogue_shared(a.x, a.y, r, r2, ysum[b0], Nf, normalization, scale);
  if (b0 + 1 < B) power[(b0 + 1) * F + k] = ls_epilogue_shared(b.x, b.y, r, r2, ysum[b0 + 1], Nf, normalization, scale);
}

// Self-check (LKB_NUFFT_VERIFY=1): 64 warps each pick one (light curve, frequency row >= k_lo) by a hash, recompute
// the two trig sums directly in fp64 and compare them with what the transform delivered; the largest deviation in
// units of 1e-7 * sum |y| goes to *worst (a correct transform stays below ~5, a defect gives >> 100).
__global__ void __launch_bounds__(128)
nufft_verify_kernel(const float2* __restrict__ Z, int log2M, int pa, const float2* __restrict__ dec, int64_t k0, int64_t F,
                    int64_t k_lo, const double* __restrict__ t, int64_t N, const float* __restrict__ yc, int64_t ystride,
                    const float* __restrict__ absmax, const double* __restrict__ freq, int B, float fault,
                    unsigned* __restrict__ worst) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned job = blockIdx.x * (blockDim.x >> 5) + warp;
  unsigned h = job * 2654435761u + 12345u;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  const int64_t b = h % (unsigned)B;
  const int64_t k = k_lo + (int64_t)((h >> 8) % (unsigned)(F - k_lo));
  const float* y = yc + b * ystride;
  const double fr = freq[k];
  double c = 0.0, s = 0.0, l1 = 0.0;
  for (int64_t i = lane; i < N; i += 32) {
    double sn, cs;
    ls_sincos_cycles_f64(fr * t[i], sn, cs);
    const double v = (double)y[i];
    c += v * cs;
    s += v * sn;
    l1 += fabs(v);
  }
  c = warp_sum(c);
  s = warp_sum(s);
  l1 = warp_sum(l1);
  if (lane == 0) {
    const int64_t M = (int64_t)1 << log2M, pair = b >> 1;
    const int64_t b0 = 2 * pair;
    const float inv0 = 1.0f / nufft::pow2_scale(absmax[b0]);
    const float inv1 = (b0 + 1 < B) ? 1.0f / nufft::pow2_scale(absmax[b0 + 1]) : 1.0f;
    float2 ha, hb;
    nufft::unpack_pair(Z + pair * M, k0 + k, M, dec[k], inv0, inv1, &ha, &hb, pa);
    const float2 got = (b & 1) ? hb : ha;
    const double dev = fmax(fabs((double)got.x * (double)fault - c), fabs((double)got.y * (double)fault - s));
    const double units = dev / (1e-7 * fmax(l1, 1e-300));
    atomicMax(worst, (unsigned)fmin(units, 4.0e9));
  }
}

__global__ void nufft_fill_kernel(float* __restrict__ p, int64_t n, float v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

// all FFT passes of `npairs` length-2^p transforms; *result points at the buffer holding the output
int fft_passes(float2* a, float2* b, int p, int npairs, cudaStream_t st, float2** result) {
  const int64_t M = (int64_t)1 << p;
  float2 *src = a, *dst = b;
  int64_t Ns = 1;
  const char* ce = getenv("LKB_NUFFT_TWIDDLE_CHAIN");
  const bool chain = ce && atoi(ce) != 0;
  for (int idx = 0;; ++idx) {
    const int R = nufft::fft_pass_radix(p, idx);
    if (R == 0) break;
    const int64_t total = (int64_t)npairs * (M / R);
    const unsigned g = blocks_for(total, 256);
    // LKB_NUFFT_TWIDDLE_CHAIN=1: one sincospif per butterfly + product tree (fewer instructions, 2.5x the rounding
    // error of the transform; off by default until the passes have been profiled)
    if (chain) launch_pass<true>(R, g, src, dst, Ns, p, total, st);
    else launch_pass<false>(R, g, src, dst, Ns, p, total, st);
    LKB_LAUNCH_CHECK();
    Ns *= R;
    float2* tmp = src; src = dst; dst = tmp;
  }
  *result = src;
  return LKB_OK;
}

// LKB_NUFFT_FFT: "smem" = four-step transform in shared memory, "fused" = the same with the spreading done inside
// the column kernel's load phase (the fine grids are written once, already half transformed)
// "v2" (default where the fine grid allows it, 2^13 .. 2^22 cells) = the pruned, tiled transform with table twiddles
// and the finish fused into the row kernel; "global" = one global sweep per radix pass.
int fft_mode(int p = 0) {
  const char* e = getenv("LKB_NUFFT_FFT");
  if (e && strcmp(e, "smem") == 0) return 1;
  if (e && strcmp(e, "fused") == 0) return 2;
  if (e && strcmp(e, "global") == 0) return 0;
  return (p >= nufft::V2_P_MIN && p <= nufft::V2_P_MAX) ? 3 : 0;
}

// twiddle tables of the v2 transform of 2^p cells in workspace slot `slot`
struct V2Tables {
  const float2 *tw_a, *tw_b, *t_hi, *t_lo;
};
int v2_tables(int p, int slot, cudaStream_t st, V2Tables* out) {
  const int pa = p - V2_PB, pl = nufft::v2_log2_lo(p);
  const int la = nufft::v2_pass_table_len(pa), lb = nufft::v2_pass_table_len(V2_PB), nhi = 1 << (p - pl), nlo = 1 << pl;
  float2* base = nullptr;
  LKB_TRY(ws_get_t<float2>(slot, (size_t)(la + lb + nhi + nlo + 4), &base));
  float2 *tw_a = base, *tw_b = base + la, *t_hi = tw_b + lb, *t_lo = t_hi + nhi;
  LKB_LAUNCH(blocks_for(la + lb + nhi + nlo, 256), 256, st, nufft2_tables_kernel)(pa, V2_PB, p, tw_a, tw_b, t_hi, t_lo);
  LKB_LAUNCH_CHECK();
  out->tw_a = tw_a; out->tw_b = tw_b; out->t_hi = t_hi; out->t_lo = t_lo;
  return LKB_OK;
}
// rows of the [A][Bc] fine grid that cadences can reach when the last one's support starts at cell i0_last
int v2_n1max(int p, int64_t i0_last, int w) {
  const int64_t M = (int64_t)1 << p, A = M >> V2_PB;
  const int64_t reach = i0_last + w + 1;                   // cells [0, reach) (a support running past M wraps to cell 0)
  if (reach >= M) return (int)A;
  const int64_t n = (reach + ((int64_t)1 << V2_PB) - 1) >> V2_PB;
  return (int)(n < 1 ? 1 : (n > A ? A : n));
}
size_t v2_cols_smem(int p) {
  const int pa = p - V2_PB;
  return (size_t)(V2_TILE >> pa) * nufft::smem_line((int64_t)1 << pa) * sizeof(float2);
}
size_t v2_rows_smem() { return (size_t)(V2_TILE >> V2_PB) * nufft::smem_line((int64_t)1 << V2_PB) * sizeof(float2); }

// G (pruned fine grids in the column layout) -> T (column transforms) for `npairs` transforms of 2^p cells
int v2_cols(const float2* G, float2* T, int p, int n1max, int npairs, const V2Tables& tb, cudaStream_t st) {
  const int pa = p - V2_PB, ptc = V2_LOG_TILE - pa;
  const size_t smem = v2_cols_smem(p);
  LKB_CUDA_CHECK(cudaFuncSetAttribute(nufft2_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LKB_LAUNCH_SMEM(dim3((unsigned)((1 << V2_PB) >> ptc), (unsigned)npairs), V2_THREADS, smem, st, nufft2_cols_kernel)(
      G, T, p, n1max, tb.tw_a, tb.t_hi, tb.t_lo);
  LKB_LAUNCH_CHECK();
  return LKB_OK;
}
// T -> power (fa != NULL), -> Zout in the [k1][k2] layout (nk2_keep = 0), or -> Zout in natural order, modes
// k < nk2_keep * A and their mirrors only (nk2_keep > 0)
int v2_rows(const float2* T, int p, int npairs, const V2Tables& tb, const V2Finish* fa, float2* Zout, cudaStream_t st,
            int nk2_keep = 0) {
  const int pa = p - V2_PB, groups = (1 << pa) >> (V2_LOG_TILE - V2_PB);          // A / (2 R)
  const size_t smem = v2_rows_smem();
  LKB_CUDA_CHECK(cudaFuncSetAttribute(nufft2_rows_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LKB_CUDA_CHECK(cudaFuncSetAttribute(nufft2_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LKB_CUDA_CHECK(cudaFuncSetAttribute(nufft2_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid((unsigned)groups, (unsigned)npairs);
  if (fa) LKB_LAUNCH_SMEM(grid, V2_THREADS, smem, st, nufft2_rows_kernel<1>)(T, p, tb.tw_b, *fa, nullptr, 0);
  else if (nk2_keep > 0) LKB_LAUNCH_SMEM(grid, V2_THREADS, smem, st, nufft2_rows_kernel<2>)(T, p, tb.tw_b, V2Finish(), Zout, nk2_keep);
  else LKB_LAUNCH_SMEM(grid, V2_THREADS, smem, st, nufft2_rows_kernel<0>)(T, p, tb.tw_b, V2Finish(), Zout, 0);
  LKB_LAUNCH_CHECK();
  return LKB_OK;
}

// in-place four-step transform of `npairs` length-2^p arrays; result in the [A][Bc] layout (pa returned).
// sp != NULL: the input is not read from Z but spread on the fly from the light curves described by *sp.
template <bool CHAIN>
int fft_fourstep_t(float2* Z, int p, int npairs, cudaStream_t st, int* pa_out, const SpreadArgs* sp) {
  const int pa = nufft::fourstep_pa(p), pb = p - pa;
  const int64_t A = (int64_t)1 << pa, Bc = (int64_t)1 << pb;
  const int tc = (int)std::max<int64_t>(1, std::min<int64_t>(Bc, fs_tile() / A));
  const int tr = (int)std::max<int64_t>(1, std::min<int64_t>(A, fs_tile() / Bc));
  const size_t smem_c = 2 * (size_t)tc * nufft::smem_line(A) * sizeof(float2);
  const size_t smem_r = 2 * (size_t)tr * nufft::smem_line(Bc) * sizeof(float2);
  LKB_CUDA_CHECK(cudaFuncSetAttribute(nufft_fft_cols_kernel<CHAIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
  LKB_CUDA_CHECK(cudaFuncSetAttribute(nufft_fft_cols_kernel<CHAIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
  LKB_CUDA_CHECK(cudaFuncSetAttribute(nufft_fft_rows_kernel<CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
  const dim3 gc((unsigned)(Bc / tc), (unsigned)npairs);
  if (sp) LKB_LAUNCH_SMEM(gc, FS_THREADS, smem_c, st, nufft_fft_cols_kernel<CHAIN, true>)(Z, p, pa, tc, *sp);
  else LKB_LAUNCH_SMEM(gc, FS_THREADS, smem_c, st, nufft_fft_cols_kernel<CHAIN, false>)(Z, p, pa, tc, SpreadArgs());
  LKB_LAUNCH_CHECK();
  LKB_LAUNCH_SMEM(dim3((unsigned)(A / tr), (unsigned)npairs), FS_THREADS, smem_r, st, nufft_fft_rows_kernel<CHAIN>)(Z, p, pa, tr);
  LKB_LAUNCH_CHECK();
  *pa_out = pa;
  return LKB_OK;
}
int fft_fourstep(float2* Z, int p, int npairs, cudaStream_t st, int* pa_out, const SpreadArgs* sp) {
  const char* ce = getenv("LKB_NUFFT_TWIDDLE_CHAIN");
  return (ce && atoi(ce) != 0) ? fft_fourstep_t<true>(Z, p, npairs, st, pa_out, sp)
                               : fft_fourstep_t<false>(Z, p, npairs, st, pa_out, sp);
}

int kernel_width() {
  int w = 8;
  if (const char* e = getenv("LKB_NUFFT_W")) w = atoi(e);
  if (w < 4) w = 4;
  if (w > 12) w = 12;
  return w & ~1;                                  // even widths only
}

}  // namespace

// Regular grid f_k = (k0 + k) df with integer k0 >= 0, df * baseline <= 1, fine grids that fit 2^24 cells.
bool ls_nufft_supported(int64_t F, bool regular, double grid_f0, double grid_df, double t_last) {
  if (!regular || F < 2 || !(grid_df > 0.0) || !(grid_f0 >= 0.0)) return false;
  const double q = grid_f0 / grid_df, k0 = rint(q);
  if (fabs(q - k0) > 1e-9 * fmax(1.0, q) || k0 > 1.0e7) return false;
  if (!(grid_df * t_last <= 1.0 + 1e-9)) return false;      // oversample 1: df * baseline = 1 up to rounding
  const int64_t kmax = (int64_t)k0 + F;
  return nufft::fine_grid_log2(2 * kmax) <= 24;
}

// ---- shared-grid path in two steps: ls_nufft_prepare (tables + window terms, once per call) and ls_nufft_run (spread
// + FFT + finish for a block of light curves; callable per chunk of a pipelined host-mode call, `ws_alt` = 1 selects a
// second set of fine-grid buffers so that two chunks can be in flight on two streams) ----
struct NufftPlan {
  int w, p;
  float beta;
  int64_t k0, M;
  const Cad* cad;
  const int32_t* fge;
  const float2* dec;
  int n1max;            // v2: rows of the [A][Bc] fine grid the cadences reach
  V2Tables tb;          // v2: twiddle tables (valid when fft_mode(p) == 3)
};
static NufftPlan g_plan;

// d_t: times shifted to t[0] = 0 (ascending - checked here).  d_rot / d_rot2 rows [0, F_low) are already filled by
// ls_window_kernel (fp64 path); the rows >= F_low are filled here from one transform of unit strengths.
int ls_nufft_prepare(const double* d_t, int64_t N, int64_t F, double grid_f0, double grid_df, float4* d_rot,
                     float2* d_rot2, int64_t F_low, cudaStream_t st) {
  const int w = kernel_width();
  const float beta = 2.30f * (float)w;
  const int64_t k0 = (int64_t)rint(grid_f0 / grid_df);
  const int p = nufft::fine_grid_log2(k0 + F), p2 = nufft::fine_grid_log2(2 * (k0 + F));
  const int64_t M = (int64_t)1 << p, M2 = (int64_t)1 << p2;
  GlNodes gl;
  nufft::gauss_legendre(32, gl.x, gl.w);

  Cad *cad = nullptr, *cad2 = nullptr;
  int32_t *fge = nullptr, *fge2 = nullptr;
  float2 *dec = nullptr, *dec2 = nullptr, *Zw = nullptr;
  float* ones = nullptr;
  int* flag = nullptr;
  const int64_t L = nufft::table_len(M, w), L2 = nufft::table_len(M2, w);
  LKB_TRY(ws_get_t<Cad>(WS_A, N, &cad));
  LKB_TRY(ws_get_t<int32_t>(WS_B, L, &fge));
  LKB_TRY(ws_get_t<float2>(WS_C, F, &dec));
  LKB_TRY(ws_get_t<Cad>(WS_J, N, &cad2));
  LKB_TRY(ws_get_t<int32_t>(WS_O, L2, &fge2));
  LKB_TRY(ws_get_t<float2>(WS_P, 2 * M2, &Zw));
  LKB_TRY(ws_get_t<float2>(WS_IN3, 2 * (k0 + F), &dec2));
  LKB_TRY(ws_get_t<float>(WS_IN4, N, &ones));
  LKB_TRY(ws_get_t<int>(WS_IN5, 1, &flag));

  // ---- tables of the two fine grids, sortedness check ----
  LKB_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), st));
  LKB_LAUNCH(blocks_for(N, 256), 256, st, nufft_cad_kernel)(d_t, N, grid_df, M, w, cad, flag);
  LKB_LAUNCH_CHECK();
  LKB_LAUNCH(blocks_for(N, 256), 256, st, nufft_cad_kernel)(d_t, N, grid_df, M2, w, cad2, flag);
  LKB_LAUNCH_CHECK();
  int h_flag = 0;
  Cad h_last;
  LKB_CUDA_CHECK(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  LKB_CUDA_CHECK(cudaMemcpyAsync(&h_last, cad + (N - 1), sizeof(Cad), cudaMemcpyDeviceToHost, st));
  LKB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (h_flag) {
    set_error("lkb_ls_power_shared: the NUFFT path needs ascending times");
    return LKB_E_UNSUPPORTED;
  }
  g_plan.n1max = 0;
  if (fft_mode(p) == 3) {
    g_plan.n1max = v2_n1max(p, (int64_t)h_last.i0, w);
    LKB_TRY(v2_tables(p, WS_IN6, st, &g_plan.tb));
  }
  LKB_LAUNCH(blocks_for(L, 256), 256, st, nufft_first_ge_kernel)(cad, N, L, fge);
  LKB_LAUNCH_CHECK();
  LKB_LAUNCH(blocks_for(L2, 256), 256, st, nufft_first_ge_kernel)(cad2, N, L2, fge2);
  LKB_LAUNCH_CHECK();
  LKB_LAUNCH(blocks_for(F, 128), 128, st, nufft_deconv_kernel)(k0, F, M, w, (double)beta, gl, dec);
  LKB_LAUNCH_CHECK();
  LKB_LAUNCH(blocks_for(2 * (k0 + F), 128), 128, st, nufft_deconv_kernel)(0, 2 * (k0 + F), M2, w, (double)beta, gl, dec2);
  LKB_LAUNCH_CHECK();

  // ---- window terms of the rows >= F_low: one transform of unit strengths on the 2x finer grid ----
  if (F_low < F) {
    LKB_LAUNCH(blocks_for(N, 256), 256, st, nufft_fill_kernel)(ones, N, 1.0f);
    LKB_LAUNCH_CHECK();
    LKB_LAUNCH(blocks_for(M2, 256), 256, st, nufft_spread_kernel)(fge2, cad2, ones, 0, nullptr, 1, 1, w, beta, p2, Zw);
    LKB_LAUNCH_CHECK();
    float2* Zw_out = nullptr;
    LKB_TRY(fft_passes(Zw, Zw + M2, p2, 1, st, &Zw_out));
    LKB_LAUNCH(blocks_for(F - F_low, 128), 128, st, nufft_rot_kernel)(Zw_out, M2, dec2, k0, F, F_low, (double)N, d_rot, d_rot2);
    LKB_LAUNCH_CHECK();
  }
  g_plan.w = w;
  g_plan.p = p;
  g_plan.beta = beta;
  g_plan.k0 = k0;
  g_plan.M = M;
  g_plan.cad = cad;
  g_plan.fge = fge;
  g_plan.dec = dec;
  return LKB_OK;
}

// d_yc: centred flux rows [B, ystride] fp32 of THIS block of light curves (with their d_ysumf / d_absmax / d_pow
// rows); d_t / d_freq / d_rot / d_rot2 as in ls_nufft_prepare, which must have run on an earlier point of the stream
// order.  prof: record the library's profiling events around the batch kernels.
int ls_nufft_run(const double* d_t, int64_t N, const float* d_yc, int64_t ystride, const float* d_ysumf,
                 const float* d_absmax, int B, const double* d_freq, int64_t F, const float4* d_rot,
                 const float2* d_rot2, int64_t F_low, int normalization, double norm_scale, float* d_pow,
                 cudaStream_t st, int ws_alt, bool prof) {
  const NufftPlan pl = g_plan;
  const int w = pl.w, p = pl.p;
  const float beta = pl.beta;
  const int64_t k0 = pl.k0, M = pl.M;
  const int npairs = (B + 1) / 2;
  const int mode = fft_mode(p);
  const char* ve = getenv("LKB_NUFFT_VERIFY");
  const bool verify = ve && atoi(ve) != 0 && F_low < F;
  float2 *Za = nullptr, *Zb = nullptr;
  LKB_TRY(ws_get_t<float2>(ws_alt ? WS_OUT4 : WS_H, (size_t)npairs * M, &Za));
  // second buffer: v2 keeps the pruned grids G there (n1max rows of Bc cells per pair; the self-check additionally
  // needs room for a few whole transforms), the other variants a whole second set of fine grids
  const size_t zb_count = (mode == 3) ? std::max((size_t)npairs * ((size_t)pl.n1max << V2_PB), verify ? (size_t)4 * M : (size_t)0)
                                      : (size_t)npairs * M;
  LKB_TRY(ws_get_t<float2>(ws_alt ? WS_OUT5 : WS_I, zb_count, &Zb));
  unsigned* d_worst = nullptr;
  if (verify) LKB_TRY(ws_get_t<unsigned>(ws_alt ? WS_OUT7 : WS_OUT6, 1, &d_worst));

  // ---- the batch: spread, FFT, finish - optionally in groups of light-curve pairs small enough for the fine grids
  // of a group (two buffers) to stay in the 126 MB L2 across the passes (LKB_NUFFT_GROUP_MB, default 0 = one group;
  // to be tuned on hardware: more launches against HBM sweeps turned into L2 sweeps) ----
  int group = npairs;
  if (const char* e = getenv("LKB_NUFFT_GROUP_MB")) {
    const double mb = atof(e);
    if (mb > 0.0) {
      const double per_pair = 2.0 * (double)M * sizeof(float2) / 1048576.0;
      group = (int)fmax(1.0, floor(mb / per_pair));
      if (group > npairs) group = npairs;
    }
  }
  if (prof) prof_begin(st);
  for (int g0 = 0; g0 < npairs; g0 += group) {
    const int np_g = std::min(group, npairs - g0);
    const int B_g = std::min(B - 2 * g0, 2 * np_g);              // light curves in this group
    float2* Za_g = Za + (size_t)g0 * M;
    float2* Zb_g = Zb + (size_t)g0 * M;
    if (mode == 3) {
      // spread (pruned, column layout) -> column transforms -> row transforms + finish
      const int pa = p - V2_PB, ptc = V2_LOG_TILE - pa;
      const size_t cells = (size_t)pl.n1max << V2_PB;
      // with L2-sized groups (LKB_NUFFT_GROUP_MB) every group goes through the SAME buffers, so that they stay in L2
      float2* G_g = (group < npairs) ? Zb : Zb + (size_t)g0 * cells;
      if (group < npairs) Za_g = Za;
      const float* y_g = d_yc + (size_t)2 * g0 * ystride;
      LKB_LAUNCH(dim3(blocks_for((int64_t)cells, 256), (unsigned)((np_g + 3) / 4)), 256, st, nufft2_spread_kernel<4>)(
          pl.fge, pl.cad, y_g, ystride, d_absmax + 2 * g0, B_g, np_g, w, beta, p, ptc, pl.n1max, G_g);
      LKB_LAUNCH_CHECK();
      LKB_TRY(v2_cols(G_g, Za_g, p, pl.n1max, np_g, pl.tb, st));
      if (F_low < F) {
        V2Finish fa;
        fa.dec = pl.dec; fa.k0 = k0; fa.F = F; fa.k_lo = F_low; fa.rot = d_rot; fa.rot2 = d_rot2;
        fa.ysum = d_ysumf + 2 * g0; fa.absmax = d_absmax + 2 * g0; fa.Nf = (float)N; fa.normalization = normalization;
        fa.scale = (float)norm_scale; fa.B = B_g; fa.power = d_pow + (size_t)2 * g0 * F;
        LKB_TRY(v2_rows(Za_g, p, np_g, pl.tb, &fa, nullptr, st));
        if (verify && g0 == 0) {          // self-check: the first pairs' transforms once more, written out this time
          const int np_v = std::min(np_g, 4), B_v = std::min(B_g, 2 * np_v);
          LKB_TRY(v2_rows(Za_g, p, np_v, pl.tb, nullptr, Zb, st));     // G is no longer needed
          const char* fe = getenv("LKB_NUFFT_INJECT_FAULT");
          LKB_CUDA_CHECK(cudaMemsetAsync(d_worst, 0, sizeof(unsigned), st));
          LKB_LAUNCH(16, 128, st, nufft_verify_kernel)(Zb, p, pa, pl.dec, k0, F, F_low, d_t, N, d_yc, ystride, d_absmax,
                                                     d_freq, B_v, fe ? (float)atof(fe) : 1.0f, d_worst);
          LKB_LAUNCH_CHECK();
          unsigned h_worst = 0;
          LKB_CUDA_CHECK(cudaMemcpyAsync(&h_worst, d_worst, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
          LKB_CUDA_CHECK(cudaStreamSynchronize(st));
          if (h_worst > 100u) {
            set_error("NUFFT self-check failed: transform deviates from the direct sums by %u x 1e-7 sum|y|", h_worst);
            return LKB_E_VERIFY;
          }
        }
      }
      continue;
    }
    if (mode != 2) {
      LKB_LAUNCH(blocks_for((int64_t)np_g * M, 256), 256, st, nufft_spread_kernel)(
          pl.fge, pl.cad, d_yc + (size_t)2 * g0 * ystride, ystride, d_absmax + 2 * g0, B_g, np_g, w, beta, p, Za_g);
      LKB_LAUNCH_CHECK();
    }
    float2* Zout = nullptr;
    int pa = 0;                                               // 0: natural order, else the four-step layout
    if (mode == 2) {
      SpreadArgs sp;
      sp.first_ge = pl.fge;
      sp.cad = pl.cad;
      sp.y = d_yc + (size_t)2 * g0 * ystride;
      sp.ystride = ystride;
      sp.absmax = d_absmax + 2 * g0;
      sp.B = B_g;
      sp.w = w;
      sp.beta = beta;
      LKB_TRY(fft_fourstep(Za_g, p, np_g, st, &pa, &sp));
      Zout = Za_g;
    } else if (mode == 1) {
      LKB_TRY(fft_fourstep(Za_g, p, np_g, st, &pa, nullptr));
      Zout = Za_g;
    } else {
      LKB_TRY(fft_passes(Za_g, Zb_g, p, np_g, st, &Zout));
    }
    if (F_low < F) {
      LKB_LAUNCH(blocks_for((F - F_low) * np_g, 256), 256, st, nufft_finish_kernel)(
          Zout, p, pl.dec, k0, F, F_low, d_rot, d_rot2, d_ysumf + 2 * g0, d_absmax + 2 * g0, (float)N, normalization,
          (float)norm_scale, B_g, np_g, pa, d_pow + (size_t)2 * g0 * F);
      LKB_LAUNCH_CHECK();
      if (verify && g0 == 0) {            // self-check on the first group (one small kernel + one read-back)
        const char* fe = getenv("LKB_NUFFT_INJECT_FAULT");       // test hook: pretend the transform is off by x
        LKB_CUDA_CHECK(cudaMemsetAsync(d_worst, 0, sizeof(unsigned), st));
        LKB_LAUNCH(16, 128, st, nufft_verify_kernel)(Zout, p, pa, pl.dec, k0, F, F_low, d_t, N, d_yc, ystride, d_absmax,
                                                   d_freq, B_g, fe ? (float)atof(fe) : 1.0f, d_worst);
        LKB_LAUNCH_CHECK();
        unsigned h_worst = 0;
        LKB_CUDA_CHECK(cudaMemcpyAsync(&h_worst, d_worst, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        LKB_CUDA_CHECK(cudaStreamSynchronize(st));
        if (h_worst > 100u) {
          set_error("NUFFT self-check failed: transform deviates from the direct sums by %u x 1e-7 sum|y|", h_worst);
          return LKB_E_VERIFY;
        }
      }
    }
  }
  if (prof) prof_end(st);
  if (F_low > 0) {
    LKB_LAUNCH(blocks_for(F_low * B, 4), 128, st, nufft_lowrows_kernel)(d_t, N, d_yc, ystride, B, d_freq, F_low, F, d_rot,
                                                                   d_rot2, d_ysumf, normalization, (float)norm_scale,
                                                                   d_pow);
    LKB_LAUNCH_CHECK();
  }
  return LKB_OK;
}



// one-shot form (the whole batch on one stream)
int ls_nufft_launch(const double* d_t, int64_t N, const float* d_yc, int64_t ystride, const float* d_ysumf,
                    const float* d_absmax, int B, const double* d_freq, int64_t F, double grid_f0, double grid_df,
                    float4* d_rot, float2* d_rot2, int64_t F_low, int normalization, double norm_scale, float* d_pow,
                    cudaStream_t st) {
  LKB_TRY(ls_nufft_prepare(d_t, N, F, grid_f0, grid_df, d_rot, d_rot2, F_low, st));
  return ls_nufft_run(d_t, N, d_yc, ystride, d_ysumf, d_absmax, B, d_freq, F, d_rot, d_rot2, F_low, normalization,
                      norm_scale, d_pow, st, 0, true);
}

// =====================================================================================================
// Ragged batches (K1 shapes: every light curve has its own times; one shared regular frequency grid).
// Opt-in through LKB_LS_RAGGED_NUFFT=1 (round 1: CPU-verified arithmetic, CUDA glue not yet run on hardware).
// Same pipeline, with per-light-curve cadence tables in the padded CSR layout of the K1 prologue, cell ranges by
// binary search (nufft::spread_cell_search), and the window terms of each light curve taken from a second
// transform of unit strengths (pairs packed the same way) inside the finish kernel - no rot arrays.
// =====================================================================================================
namespace {

__global__ void nufft_cad_ragged_kernel(const double* __restrict__ t, const int64_t* __restrict__ off,
                                        const int64_t* __restrict__ poff, const double* __restrict__ span, double df,
                                        int64_t M, int64_t M2, int w, Cad* __restrict__ cad, Cad* __restrict__ cad2,
                                        int* __restrict__ bad) {
  const int b = blockIdx.y;
  const int64_t n = off[b + 1] - off[b], po = poff[b];
  if (blockIdx.x == 0 && threadIdx.x == 0 && !(df * span[b] <= 1.0 + 1e-9)) *bad = 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ti = t[po + i];
    cad[po + i] = nufft::cad_entry(ti, df, M, w);
    cad2[po + i] = nufft::cad_entry(ti, df, M2, w);
    if (ti < 0.0 || (i > 0 && ti < t[po + i - 1])) *bad = 1;
  }
}

__global__ void __launch_bounds__(256)
nufft_absmax_ragged_kernel(const float* __restrict__ y, const int64_t* __restrict__ off, const int64_t* __restrict__ poff,
                           float* __restrict__ absmax) {
  __shared__ float s_max[8];
  const int b = blockIdx.x;
  const int64_t n = off[b + 1] - off[b], po = poff[b];
  float mx = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, fabsf(y[po + i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) m = fmaxf(m, s_max[wv]);
    absmax[b] = m;
  }
}

// Z[pair][m] for the flux (y != NULL, scaled) or for unit strengths (y == NULL)
__global__ void __launch_bounds__(256)
nufft_spread_ragged_kernel(const Cad* __restrict__ cad, const float* __restrict__ y, const int64_t* __restrict__ off,
                           const int64_t* __restrict__ poff, const float* __restrict__ absmax, int B, int npairs, int w,
                           float beta, int log2M, float2* __restrict__ Z) {
  const int64_t M = (int64_t)1 << log2M;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)npairs << log2M) return;
  const int64_t pair = gid >> log2M, m = gid & (M - 1);
  const int64_t b0 = 2 * pair, b1 = b0 + 1;
  float2 v = make_float2(0.f, 0.f);
  {
    const int64_t po = poff[b0], n = off[b0 + 1] - off[b0];
    v.x = nufft::spread_cell_search(m, cad + po, n, y ? y + po : nullptr, y ? nufft::pow2_scale(absmax[b0]) : 1.0f, w,
                                    beta, M);
  }
  if (b1 < B) {
    const int64_t po = poff[b1], n = off[b1 + 1] - off[b1];
    v.y = nufft::spread_cell_search(m, cad + po, n, y ? y + po : nullptr, y ? nufft::pow2_scale(absmax[b1]) : 1.0f, w,
                                    beta, M);
  }
  Z[gid] = v;
}

// v2: the same cell values in the column kernel's layout G[pair][c][n1][j], rows n1 < n1max only
__global__ void __launch_bounds__(256)
nufft2_spread_ragged_kernel(const Cad* __restrict__ cad, const float* __restrict__ y, const int64_t* __restrict__ off,
                            const int64_t* __restrict__ poff, const float* __restrict__ absmax, int B, int npairs, int w,
                            float beta, int p, int ptc, int n1max, float2* __restrict__ G) {
  const int64_t cells = (int64_t)n1max << V2_PB;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cells) return;
  const int64_t M = (int64_t)1 << p, m = v2_cell_of(e, ptc, n1max);
  const int64_t pair = blockIdx.y, b0 = 2 * pair, b1 = b0 + 1;
  float2 v = make_float2(0.f, 0.f);
  {
    const int64_t po = poff[b0], n = off[b0 + 1] - off[b0];
    v.x = nufft::spread_cell_search(m, cad + po, n, y ? y + po : nullptr, y ? nufft::pow2_scale(absmax[b0]) : 1.0f, w,
                                    beta, M);
  }
  if (b1 < B) {
    const int64_t po = poff[b1], n = off[b1 + 1] - off[b1];
    v.y = nufft::spread_cell_search(m, cad + po, n, y ? y + po : nullptr, y ? nufft::pow2_scale(absmax[b1]) : 1.0f, w,
                                    beta, M);
  }
  G[pair * cells + e] = v;
}

__device__ __forceinline__ float ragged_power(float2 hs, float2 win1, float2 win2, double Nd, double ysum,
                                              int normalization, double scale) {
  LsSums<double> d;
  d.zero();
  d.ch = (double)hs.x;
  d.sh = (double)hs.y;
  d.c = (double)win1.x;
  d.s = (double)win1.y;
  d.cc = 0.5 * (Nd + (double)win2.x);
  d.sc = 0.5 * (double)win2.y;
  return ls_normalize(ls_power_from_sums(d, Nd, ysum), Nd, normalization, scale);
}

// power[b, k] for the rows that are not "low" for light curve b
__global__ void __launch_bounds__(256)
nufft_finish_ragged_kernel(const float2* __restrict__ Z, int log2M, const float2* __restrict__ Zw, int log2M2,
                           const float2* __restrict__ dec, const float2* __restrict__ dec2, int64_t k0, int64_t F,
                           double f0, double df, const int64_t* __restrict__ off, const double* __restrict__ span,
                           const double* __restrict__ ysum, const float* __restrict__ absmax, int normalization,
                           const double* __restrict__ norm_scale, int B, int npairs, int pa, int pa2,
                           float* __restrict__ power) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= F * npairs) return;
  const int64_t pair = gid / F, k = gid - pair * F;
  const int64_t M = (int64_t)1 << log2M, M2 = (int64_t)1 << log2M2, kk = k0 + k;
  const int64_t b0 = 2 * pair, b1 = b0 + 1;
  const bool has1 = b1 < B;
  const float inv0 = 1.0f / nufft::pow2_scale(absmax[b0]);
  const float inv1 = has1 ? 1.0f / nufft::pow2_scale(absmax[b1]) : 1.0f;
  float2 ha, hb, w1a, w1b, w2a, w2b;
  nufft::unpack_pair(Z + pair * M, kk, M, dec[k], inv0, inv1, &ha, &hb, pa);
  nufft::unpack_pair(Zw + pair * M2, kk, M2, dec2[kk], 1.0f, 1.0f, &w1a, &w1b, pa2);
  nufft::unpack_pair(Zw + pair * M2, 2 * kk, M2, dec2[2 * kk], 1.0f, 1.0f, &w2a, &w2b, pa2);
  const double fr = f0 + (double)k * df;
  if (fr * span[b0] > LS_LOWF_CYCLES) {
    const double Nd = (double)(off[b0 + 1] - off[b0]);
    power[b0 * F + k] = ragged_power(ha, w1a, w2a, Nd, ysum[b0], normalization, norm_scale ? norm_scale[b0] : 1.0);
  }
  if (has1 && fr * span[b1] > LS_LOWF_CYCLES) {
    const double Nd = (double)(off[b1 + 1] - off[b1]);
    power[b1 * F + k] = ragged_power(hb, w1b, w2b, Nd, ysum[b1], normalization, norm_scale ? norm_scale[b1] : 1.0);
  }
}

// rows with f * baseline_b <= LS_LOWF_CYCLES: direct fp64 sums, one warp per (row, light curve)
__global__ void __launch_bounds__(128)
nufft_lowrows_ragged_kernel(const double* __restrict__ t, const float* __restrict__ y, const int64_t* __restrict__ off,
                            const int64_t* __restrict__ poff, const double* __restrict__ span,
                            const double* __restrict__ ysum, double f0, double df, int64_t F_low_max, int64_t F,
                            int normalization, const double* __restrict__ norm_scale, int B, float* __restrict__ power) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t job = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (job >= F_low_max * B) return;
  const int64_t b = job / F_low_max, k = job - b * F_low_max;
  const double fr = f0 + (double)k * df;
  if (k >= F || fr * span[b] > LS_LOWF_CYCLES) return;
  const int64_t n = off[b + 1] - off[b], po = poff[b];
  if (n <= 0) return;
  LsSums<double> d;
  d.zero();
  for (int64_t i = lane; i < n; i += 32) {
    double s, c;
    ls_sincos_cycles_f64(fr * t[po + i], s, c);
    d.add((double)y[po + i], s, c);
  }
  d.warp_reduce();
  if (lane == 0)
    power[b * F + k] = ls_normalize(ls_power_from_sums(d, (double)n, ysum[b]), (double)n, normalization,
                                    norm_scale ? norm_scale[b] : 1.0);
}

}  // namespace

bool ls_nufft_ragged_enabled() {               // `auto` of the ragged entry may use this path (default: yes)
  const char* e = getenv("LKB_LS_RAGGED_NUFFT");
  return !e || atoi(e) != 0;
}

// One shared regular grid f_k = f0 + k df (k0 = f0 / df integer), light curves in the K1 prologue's layout
// (d_t / d_y padded CSR with offsets d_po; d_off the unpadded offsets; d_span, d_ysum per light curve).
// h_span: host copy of d_span.  Returns LKB_E_UNSUPPORTED when a light curve is not eligible (unsorted times,
// df * baseline > 1): the caller then runs the direct kernel.
// The fine grids of all light curves need 24 (M + 2 M2) bytes per pair; batches whose grids exceed
// LKB_NUFFT_RAGGED_MB (default 16384) are processed in groups of pairs through the same buffers.
int ls_nufft_ragged_launch(const double* d_t, const float* d_y, const int64_t* d_off, const int64_t* d_po,
                           const int64_t* h_off, int B, int64_t ptotal, int64_t nmax, const double* d_span,
                           const double* h_span, const double* d_ysum, int64_t F, double f0, double df,
                           int normalization, const double* d_ns, float* d_pow, cudaStream_t st) {
  (void)h_off;
  const double q = f0 / df, k0d = rint(q);
  if (!(df > 0.0) || !(f0 >= 0.0) || fabs(q - k0d) > 1e-9 * fmax(1.0, q) || k0d > 1.0e7) {
    set_error("NUFFT (ragged): the grid is not f_k = (k0 + k) df with integer k0");
    return LKB_E_UNSUPPORTED;
  }
  const int64_t k0 = (int64_t)k0d;
  const int p = nufft::fine_grid_log2(k0 + F), p2 = nufft::fine_grid_log2(2 * (k0 + F));
  if (p2 > 24) { set_error("NUFFT (ragged): fine grid larger than 2^24 cells"); return LKB_E_UNSUPPORTED; }
  for (int b = 0; b < B; ++b) {
    if (!(h_span[b] > 0.0) || !(df * h_span[b] <= 1.0 + 1e-9)) {
      set_error("NUFFT (ragged): a light curve has zero baseline or df * baseline > 1");
      return LKB_E_UNSUPPORTED;
    }
  }
  const int w = kernel_width();
  const float beta = 2.30f * (float)w;
  const int64_t M = (int64_t)1 << p, M2 = (int64_t)1 << p2;
  const int npairs_all = (B + 1) / 2;
  double cap_mb = 16384.0;
  if (const char* e = getenv("LKB_NUFFT_RAGGED_MB")) { const double v = atof(e); if (v > 0.0) cap_mb = v; }
  const double per_pair_mb = (2.0 * (double)M + 2.0 * (double)M2) * sizeof(float2) / 1048576.0;
  int group = (int)fmax(1.0, floor(cap_mb / per_pair_mb));
  if (group > npairs_all) group = npairs_all;
  GlNodes gl;
  nufft::gauss_legendre(32, gl.x, gl.w);

  // v2 transform for both fine grids when they are in range: pruned to the rows the longest light curve reaches
  const bool v2 = fft_mode(p) == 3 && fft_mode(p2) == 3;
  int n1max = 0, n1max2 = 0;
  V2Tables tb, tb2;
  float2* Gbuf = nullptr;
  if (v2) {
    double span_max = 0.0;
    for (int b = 0; b < B; ++b) span_max = fmax(span_max, h_span[b]);
    n1max = v2_n1max(p, (int64_t)nufft::cad_entry(span_max, df, M, w).i0, w);
    n1max2 = v2_n1max(p2, (int64_t)nufft::cad_entry(span_max, df, M2, w).i0, w);
    LKB_TRY(v2_tables(p, WS_IN7, st, &tb));
    LKB_TRY(v2_tables(p2, WS_OUT1, st, &tb2));
    LKB_TRY(ws_get_t<float2>(WS_OUT2, (size_t)group * ((size_t)std::max(n1max, n1max2) << V2_PB), &Gbuf));
  }

  Cad *cad = nullptr, *cad2 = nullptr;
  float2 *dec = nullptr, *dec2 = nullptr, *Za = nullptr, *Zb = nullptr, *Zw = nullptr;
  float* absmax = nullptr;
  int* flag = nullptr;
  LKB_TRY(ws_get_t<Cad>(WS_K, ptotal + 4, &cad));
  LKB_TRY(ws_get_t<Cad>(WS_L, ptotal + 4, &cad2));
  LKB_TRY(ws_get_t<float>(WS_M, B, &absmax));
  LKB_TRY(ws_get_t<float2>(WS_N, (size_t)group * M, &Za));
  LKB_TRY(ws_get_t<float2>(WS_O, (size_t)group * M, &Zb));
  LKB_TRY(ws_get_t<float2>(WS_P, (size_t)2 * group * M2, &Zw));
  LKB_TRY(ws_get_t<float2>(WS_IN4, F, &dec));
  LKB_TRY(ws_get_t<float2>(WS_IN5, 2 * (k0 + F), &dec2));
  LKB_TRY(ws_get_t<int>(WS_IN6, 1, &flag));

  LKB_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), st));
  {
    const unsigned gx = (unsigned)std::min<int64_t>(64, (nmax + 255) / 256);
    LKB_LAUNCH(dim3(gx ? gx : 1, (unsigned)B), 256, st, nufft_cad_ragged_kernel)(d_t, d_off, d_po, d_span, df, M, M2, w, cad,
                                                                          cad2, flag);
    LKB_LAUNCH_CHECK();
  }
  int h_flag = 0;
  LKB_CUDA_CHECK(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  LKB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (h_flag) { set_error("NUFFT (ragged): a light curve has unsorted times"); return LKB_E_UNSUPPORTED; }
  LKB_LAUNCH(B, 256, st, nufft_absmax_ragged_kernel)(d_y, d_off, d_po, absmax);
  LKB_LAUNCH_CHECK();
  LKB_LAUNCH(blocks_for(F, 128), 128, st, nufft_deconv_kernel)(k0, F, M, w, (double)beta, gl, dec);
  LKB_LAUNCH_CHECK();
  LKB_LAUNCH(blocks_for(2 * (k0 + F), 128), 128, st, nufft_deconv_kernel)(0, 2 * (k0 + F), M2, w, (double)beta, gl, dec2);
  LKB_LAUNCH_CHECK();

  prof_begin(st);
  for (int g0 = 0; g0 < npairs_all; g0 += group) {
    const int npairs = std::min(group, npairs_all - g0);
    const int b0 = 2 * g0, Bg = std::min(B - b0, 2 * npairs);
    const int64_t *off_g = d_off + b0, *po_g = d_po + b0;
    const float* amax_g = absmax + b0;
    // rows k with (f0 + k df) * span_b <= LS_LOWF_CYCLES for at least one light curve of the group
    double span_min = 1e300;
    for (int b = b0; b < b0 + Bg; ++b) span_min = fmin(span_min, h_span[b]);
    const double nlow = floor((LS_LOWF_CYCLES / span_min - f0) / df) + 2.0;
    const int64_t F_low_max = nlow < 0.0 ? 0 : (nlow > (double)F ? F : (int64_t)nlow);
    float2* Zw_out = nullptr;
    float2* Zout = nullptr;
    int pa = 0, pa2 = 0;                       // LKB_NUFFT_FFT=smem|fused: four-step transforms (in place)
    if (v2) {
      // window terms (unit strengths, 2x finer grid: modes kk and 2 kk), then the flux; the row kernels write the
      // needed modes and their mirrors in natural order
      const int nk2w = (int)((2 * (k0 + F)) >> (p2 - V2_PB)) + 1, nk2 = (int)((k0 + F) >> (p - V2_PB)) + 1;
      const size_t cells2 = (size_t)n1max2 << V2_PB, cells = (size_t)n1max << V2_PB;
      LKB_LAUNCH(dim3(blocks_for((int64_t)cells2, 256), (unsigned)npairs), 256, st, nufft2_spread_ragged_kernel)(
          cad2, nullptr, off_g, po_g, amax_g, Bg, npairs, w, beta, p2, V2_LOG_TILE - (p2 - V2_PB), n1max2, Gbuf);
      LKB_LAUNCH_CHECK();
      LKB_TRY(v2_cols(Gbuf, Zw, p2, n1max2, npairs, tb2, st));
      Zw_out = Zw + (size_t)npairs * M2;
      LKB_TRY(v2_rows(Zw, p2, npairs, tb2, nullptr, Zw_out, st, nk2w));
      LKB_LAUNCH(dim3(blocks_for((int64_t)cells, 256), (unsigned)npairs), 256, st, nufft2_spread_ragged_kernel)(
          cad, d_y, off_g, po_g, amax_g, Bg, npairs, w, beta, p, V2_LOG_TILE - (p - V2_PB), n1max, Gbuf);
      LKB_LAUNCH_CHECK();
      LKB_TRY(v2_cols(Gbuf, Za, p, n1max, npairs, tb, st));
      Zout = Zb;
      LKB_TRY(v2_rows(Za, p, npairs, tb, nullptr, Zout, st, nk2));
    } else {
    // window terms: unit strengths on the 2x finer grid
    LKB_LAUNCH(blocks_for((int64_t)npairs * M2, 256), 256, st, nufft_spread_ragged_kernel)(cad2, nullptr, off_g, po_g, amax_g,
                                                                                    Bg, npairs, w, beta, p2, Zw);
    LKB_LAUNCH_CHECK();
    if (fft_mode() != 0) {
      LKB_TRY(fft_fourstep(Zw, p2, npairs, st, &pa2, nullptr));
      Zw_out = Zw;
    } else {
      LKB_TRY(fft_passes(Zw, Zw + (size_t)npairs * M2, p2, npairs, st, &Zw_out));
    }
    // flux
    LKB_LAUNCH(blocks_for((int64_t)npairs * M, 256), 256, st, nufft_spread_ragged_kernel)(cad, d_y, off_g, po_g, amax_g, Bg,
                                                                                   npairs, w, beta, p, Za);
    LKB_LAUNCH_CHECK();
    if (fft_mode() != 0) {
      LKB_TRY(fft_fourstep(Za, p, npairs, st, &pa, nullptr));
      Zout = Za;
    } else {
      LKB_TRY(fft_passes(Za, Zb, p, npairs, st, &Zout));
    }
    }
    LKB_LAUNCH(blocks_for(F * npairs, 256), 256, st, nufft_finish_ragged_kernel)(
        Zout, p, Zw_out, p2, dec, dec2, k0, F, f0, df, off_g, d_span + b0, d_ysum + b0, amax_g, normalization,
        d_ns ? d_ns + b0 : nullptr, Bg, npairs, pa, pa2, d_pow + (size_t)b0 * F);
    LKB_LAUNCH_CHECK();
    if (F_low_max > 0) {
      LKB_LAUNCH(blocks_for(F_low_max * Bg, 4), 128, st, nufft_lowrows_ragged_kernel)(
          d_t, d_y, off_g, po_g, d_span + b0, d_ysum + b0, f0, df, F_low_max, F, normalization,
          d_ns ? d_ns + b0 : nullptr, Bg, d_pow + (size_t)b0 * F);
      LKB_LAUNCH_CHECK();
    }
  }
  prof_end(st);
  return LKB_OK;
}

}  // namespace lkb
