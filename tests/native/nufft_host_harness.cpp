// CPU harness for lightkurve_b200/csrc/nufft_core.h (TEST INFRASTRUCTURE): runs the per-thread functions of the
// NUFFT Lomb-Scargle kernels in plain loops - "thread index" = loop variable - so that tests/test_nufft_core.py can
// check the index arithmetic, the butterflies, the pair packing and the deconvolution against numpy without a GPU.
// Built on the fly by the test with:  g++ -O2 -shared -fPIC -o <tmp>/libnufft_harness.so nufft_host_harness.cpp
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../lightkurve_b200/csrc/nufft_core.h"

using namespace lkb::nufft;

static bool g_chain = false;          // twiddles by product tree (harness_set_chain)

template <int R>
static void run_pass(const float2* x, float2* y, int64_t Ns, int64_t M) {
  if (g_chain) for (int64_t i = 0; i < M / R; ++i) fft_pass_butterfly<R, true>(x, y, i, Ns, M);
  else for (int64_t i = 0; i < M / R; ++i) fft_pass_butterfly<R, false>(x, y, i, Ns, M);
}

// length-2^p transform (+i sign) of `in` (interleaved re, im); result in `out`
static void fft_full(const float2* in, float2* out, int p) {
  const int64_t M = (int64_t)1 << p;
  std::vector<float2> a(in, in + M), b(M);
  float2 *src = a.data(), *dst = b.data();
  int64_t Ns = 1;
  for (int idx = 0;; ++idx) {
    const int R = fft_pass_radix(p, idx);
    if (R == 0) break;
    if (R == 16) run_pass<16>(src, dst, Ns, M);
    else if (R == 8) run_pass<8>(src, dst, Ns, M);
    else if (R == 4) run_pass<4>(src, dst, Ns, M);
    else run_pass<2>(src, dst, Ns, M);
    Ns *= R;
    float2* tmp = src; src = dst; dst = tmp;
  }
  memcpy(out, src, sizeof(float2) * M);
}

extern "C" {

void harness_set_chain(int on) { g_chain = on != 0; }

// grid-size and kernel rules of nufft_core.h (the library's fine_log2 / kernel_width / es_beta call these)
int harness_fine_grid_log2(int64_t kmax_plus_1, double sigma_min) { return fine_grid_log2(kmax_plus_1, sigma_min); }
double harness_grid_sigma(int p, int64_t kmax_plus_1) { return grid_sigma(p, kmax_plus_1); }
int harness_es_width(double sigma) { return es_width(sigma); }
double harness_es_beta(int w, double sigma) { return es_beta(w, sigma); }

int harness_fft(const float* in, int p, float* out) {
  fft_full(reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), p);
  return 0;
}

int harness_gauss_legendre(int n, double* x, double* w) {
  gauss_legendre(n, x, w);
  return 0;
}

// trig sums of one pair of light curves:  C + i S = sum_n y[n] exp(2 pi i (k0 + k) df t_rel[n]),  k < F
// out_* have F entries; y1 / C1 / S1 may be NULL.  Returns log2 of the fine-grid size.
// `search` != 0: the ragged-batch variant (binary searches per cell instead of the first_ge table); then the two
// light curves may have DIFFERENT cadence sets: (t_rel, N, y0) and (t1_rel, N1, y1).
int harness_trig_sums_ex(const double* t_rel, int64_t N, const float* y0, const double* t1_rel, int64_t N1,
                         const float* y1, double df, int64_t k0, int64_t F, int w, int search, float* C0, float* S0,
                         float* C1, float* S1);

int harness_trig_sums(const double* t_rel, int64_t N, const float* y0, const float* y1, double df, int64_t k0,
                      int64_t F, int w, float* C0, float* S0, float* C1, float* S1) {
  return harness_trig_sums_ex(t_rel, N, y0, t_rel, N, y1, df, k0, F, w, 0, C0, S0, C1, S1);
}

int harness_trig_sums_ex(const double* t_rel, int64_t N, const float* y0, const double* t1_rel, int64_t N1,
                         const float* y1, double df, int64_t k0, int64_t F, int w, int search, float* C0, float* S0,
                         float* C1, float* S1) {
  const int p = fine_grid_log2(k0 + F);
  const int64_t M = (int64_t)1 << p;
  const float beta = 2.30f * (float)w;
  std::vector<Cad> cad(N);
  for (int64_t n = 0; n < N; ++n) cad[n] = cad_entry(t_rel[n], df, M, w);
  const int64_t L = table_len(M, w);
  std::vector<int32_t> first_ge(L);
  for (int64_t c = 0; c < L; ++c) first_ge[c] = first_ge_entry(c, cad.data(), N);
  float mx0 = 0.f, mx1 = 0.f;
  for (int64_t n = 0; n < N; ++n) mx0 = fmaxf(mx0, fabsf(y0[n]));
  if (y1) for (int64_t n = 0; n < N1; ++n) mx1 = fmaxf(mx1, fabsf(y1[n]));
  const float s0 = pow2_scale(mx0), s1 = pow2_scale(mx1);
  std::vector<float2> grid(M), spec(M);
  if (search) {
    std::vector<Cad> cad1(N1);
    for (int64_t n = 0; n < N1; ++n) cad1[n] = cad_entry(t1_rel[n], df, M, w);
    for (int64_t m = 0; m < M; ++m)
      grid[m] = make_float2(spread_cell_search(m, cad.data(), N, y0, s0, w, beta, M),
                            y1 ? spread_cell_search(m, cad1.data(), N1, y1, s1, w, beta, M) : 0.0f);
  } else {
    for (int64_t m = 0; m < M; ++m) grid[m] = spread_cell(m, first_ge.data(), cad.data(), y0, y1, s0, s1, w, beta, M);
  }
  fft_full(grid.data(), spec.data(), p);
  double glx[32], glw[32];
  gauss_legendre(32, glx, glw);
  for (int64_t k = 0; k < F; ++k) {
    double re, im;
    deconv_factor(k0 + k, M, w, (double)beta, glx, glw, 32, &re, &im);
    float2 a, b;
    unpack_pair(spec.data(), k0 + k, M, make_float2((float)re, (float)im), 1.0f / s0, 1.0f / s1, &a, &b);
    C0[k] = a.x; S0[k] = a.y;
    if (C1) { C1[k] = b.x; S1[k] = b.y; }
  }
  return p;
}

// Intermediate buffers of the table-based (shared-grid) pipeline in the layouts ls_nufft_launch keeps them in its
// workspace: cad [N] {int32 i0, float d0}, first_ge [M + 2w + 4] int32, grid / spec [M] float2, dec [F] float2.
int harness_stages(const double* t_rel, int64_t N, const float* y0, const float* y1, double df, int64_t k0, int64_t F,
                   int w, int32_t* cad_out, int32_t* first_ge_out, float* spec_out, float* dec_out) {
  const int p = fine_grid_log2(k0 + F);
  const int64_t M = (int64_t)1 << p;
  const float beta = 2.30f * (float)w;
  std::vector<Cad> cad(N);
  for (int64_t n = 0; n < N; ++n) cad[n] = cad_entry(t_rel[n], df, M, w);
  memcpy(cad_out, cad.data(), sizeof(Cad) * N);
  const int64_t L = table_len(M, w);
  std::vector<int32_t> first_ge(L);
  for (int64_t c = 0; c < L; ++c) first_ge[c] = first_ge_entry(c, cad.data(), N);
  memcpy(first_ge_out, first_ge.data(), sizeof(int32_t) * L);
  float mx0 = 0.f, mx1 = 0.f;
  for (int64_t n = 0; n < N; ++n) { mx0 = fmaxf(mx0, fabsf(y0[n])); if (y1) mx1 = fmaxf(mx1, fabsf(y1[n])); }
  std::vector<float2> grid(M), spec(M);
  for (int64_t m = 0; m < M; ++m)
    grid[m] = spread_cell(m, first_ge.data(), cad.data(), y0, y1, pow2_scale(mx0), pow2_scale(mx1), w, beta, M);
  fft_full(grid.data(), spec.data(), p);
  memcpy(spec_out, spec.data(), sizeof(float2) * M);
  double glx[32], glw[32];
  gauss_legendre(32, glx, glw);
  for (int64_t k = 0; k < F; ++k) {
    double re, im;
    deconv_factor(k0 + k, M, w, (double)beta, glx, glw, 32, &re, &im);
    dec_out[2 * k] = (float)re;
    dec_out[2 * k + 1] = (float)im;
  }
  return p;
}

}  // extern "C"
