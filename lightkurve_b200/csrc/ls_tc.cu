// K2c: shared-cadence-grid Lomb-Scargle contraction on the 5th-generation tensor cores.
//
//   Ch[f, b] = sum_n cos(2 pi f t_n) y_b[n],   Sh[f, b] = sum_n sin(2 pi f t_n) y_b[n]
// is a GEMM  D[M = frequencies, N = light curves] = A[M, K = cadences] * Y[N, K]^T  whose A operand
// (the sin/cos design matrix, F x N_cad - 1.3e10 elements at BASELINE config 2) is never
// materialised in HBM: producer warps synthesise each 128 x 32 tile directly into shared
// memory in the UMMA canonical K-major SWIZZLE_64B layout, while the TMA engine streams the
// matching 256 x 32 flux tile.  One elected thread issues tcgen05.mma (kind::f16, M128 N256
// K16, cta_group::1); the cos and sin accumulators (128 lanes x 256 fp32 columns each) fill the
// SM's 512 TMEM columns, and the epilogue reads them back with tcgen05.ld, applies the
// per-frequency tau rotation / CC' SS' terms (ls_window_kernel) and lightkurve's normalisation.
//
// Precision: fp16 operands alone would leave ~1e-3 relative error in weak bins.
// Both operands are split hi + lo (fp16 + fp16 residual, power-of-two pre-scaling so that the
// residuals stay normal numbers) and three products are accumulated per k-step
// (Ah*Yh + Ah*Yl + Al*Yh; Al*Yl ~ 2^-22 is dropped), fp32 accumulation in TMEM.
// Roofline: algorithmic flops = 4 F N B; issued tensor flops = 3x that.
//
// Accumulation (measured on B200, profiles/): tcgen05 adds into the fp32 accumulator with
// TRUNCATION, so a chain of k MMAs shrinks the sum by ~0.35 k 2^-24 (2.5e-4 at 65 000 cadences,
// 12 k chained MMAs - outside the LS tolerance).  The cadence axis is therefore split into
// segments of <= TC_SEG_STAGES pipeline stages: the CTA walks the segments of its tile back to
// back (the smem pipeline never drains), four dedicated epilogue warps copy each finished
// segment's accumulators out of TMEM as an fp32 partial (Ch, Sh) plane while the producers
// already fill the next stages, and ls_tc_finish_kernel sums the planes with round-to-nearest
// CUDA-core adds before the epilogue math.
//
// Warp roles (704 threads): warp 0 = TMA producer (flux tiles), warp 1 = MMA issuer + TMEM owner, warps 2..17 =
// design-matrix generators, warps 18..21 = epilogue (warp_id % 4 covers the four TMEM lane quadrants).
// Kernels in this file: ls_tcg_kernel (shipped: generator warps in two groups that fill alternate stages),
// ls_tc_kernel (its lock-step predecessor, kept for A/B timing and the LKB_TC_DEBUG experiments) and ls_tc2_kernel
// (CTA-pair / cta_group::2 variant, opt-in) - DESIGN.md section 4 K2 has the measurements behind that choice.
#include "common.cuh"
#include "ptx.cuh"
#include "ls_common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace lkb {

constexpr int TC_BM = 128;            // frequencies per CTA (TMEM lanes)
constexpr int TC_BN = 256;            // light curves per CTA (columns per accumulator)
constexpr int TC_BK = 32;             // cadences per pipeline stage (64-byte fp16 rows, SWIZZLE_64B)
constexpr int TC_STAGES = 3;
constexpr int TC_GEN_WARPS = 16;      // 512 generator threads: one 8-cadence chunk of one row per stage each
constexpr int TC_EPI_WARPS = 4;
constexpr int TC_THREADS = (2 + TC_GEN_WARPS + TC_EPI_WARPS) * 32;
constexpr int TC_A_TILE = TC_BM * TC_BK * 2;          // 8 KB
constexpr int TC_Y_TILE = TC_BN * TC_BK * 2;          // 16 KB
constexpr int TC_STAGE_BYTES = 4 * TC_A_TILE + 2 * TC_Y_TILE;   // 64 KB
constexpr float TC_A_SCALE = 256.0f;                  // 2^8: keeps fp16 residuals of cos/sin normal
constexpr int TC_SCRATCH = TC_GEN_WARPS * 8 * 16;     // per-warp broadcast slots for the phase table
constexpr size_t TC_SMEM = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + TC_SCRATCH;
constexpr int TC_SEG_STAGES = 64;     // <= 2048 cadences (384 chained MMAs) per accumulator chain

// UMMA shared-memory descriptor, K-major, SWIZZLE_64B: 8-row atoms of 512 B (SBO), version 1.
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(512 >> 4) << 32;                     // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                              // layout type: SWIZZLE_64B
  return d;
}
// instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=256
constexpr uint32_t TC_IDESC = (1u << 4) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

// flux -> power-of-two scaled fp16 hi/lo planes: yhl[0][b][n] = hi, yhl[1][b][n] = lo.  One block
// per light curve; also returns the sum of the EFFECTIVE values (hi + lo) / scale for the mean term.
__global__ void __launch_bounds__(256)
tc_split_flux_kernel(const float* __restrict__ yc, const float* __restrict__ absmax, int B, int64_t Npad,
                     __half* __restrict__ yhl, float* __restrict__ inv_scale, float* __restrict__ ysum_eff) {
  __shared__ double red[33];
  const int b = blockIdx.x;
  const float am = absmax[b];
  // scale = 2^k with scale*absmax in [2^13, 2^14); all-zero light curves keep scale 1
  int e = 0;
  float sc = 1.0f;
  if (am > 0.f && isfinite(am)) {
    frexpf(am, &e);                       // am = m * 2^e, m in [0.5, 1)
    sc = ldexpf(1.0f, 14 - e);
  }
  double acc = 0.0;
  for (int64_t i = (int64_t)threadIdx.x * 2; i < Npad; i += (int64_t)blockDim.x * 2) {
    const float2 v = *reinterpret_cast<const float2*>(yc + (int64_t)b * Npad + i);
    const float a0 = v.x * sc, a1 = v.y * sc;
    const __half2 h = __floats2half2_rn(a0, a1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
    const float2 lf = __half22float2(l);
    *reinterpret_cast<__half2*>(yhl + (int64_t)b * Npad + i) = h;
    *reinterpret_cast<__half2*>(yhl + ((int64_t)B + b) * Npad + i) = l;
    acc += (double)hf.x + (double)lf.x + (double)hf.y + (double)lf.y;
  }
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) {
    inv_scale[b] = 1.0f / (sc * TC_A_SCALE);
    ysum_eff[b] = (float)(tot / (double)sc);
  }
}

struct TcParams {
  const double* t;          // [Npad] shifted times (padding cadences hold 0)       (irregular grids)
  const ulonglong2* tab;    // [Npad] fixed-point phase table {a_n, b_n}            (regular grids)
  const double* freq;       // [F]
  const float4* rot;        // [F]
  const float2* rot2;       // [F]
  const float* ysum;        // [B] sum of the effective (hi + lo) / scale flux
  const float* inv_scale;   // [B]
  float* power;             // [B, F]
  float* part;              // [nseg, 2, B, F] raw partial (Ch, Sh) accumulators when nseg > 1
  int64_t N, Npad, F;
  int B;
  int normalization;
  float norm_scale;
  int seg_stages, nseg;
  double* wsum;             // [F, 4 chunks, 4] window sums (S, C, CC, SC) accumulated by the generators of
                            // the blockIdx.y == 0 CTAs (nullptr: the separate window kernel is used)
  double lowf_max;          // frequencies <= this are "low rows": design matrix carries cos - 1
  int debug;                // LKB_TC_DEBUG bit mask: timing experiments only (results are wrong when set)
  float low_mul;            // extra output scale of the low rows (1; a variant may generate them at half scale)
  double f0, df;            // regular grid (REGULAR kernels)
};

// split fp32 (c0, c1) and (s0, s1) into fp16 hi / residual words
__device__ __forceinline__ void tc_split2(float c0, float c1, float s0, float s1, uint32_t& ch, uint32_t& cl,
                                          uint32_t& sh, uint32_t& sl) {
  c0 *= TC_A_SCALE; c1 *= TC_A_SCALE; s0 *= TC_A_SCALE; s1 *= TC_A_SCALE;
  const __half2 hc = __floats2half2_rn(c0, c1), hs = __floats2half2_rn(s0, s1);
  const float2 fc = __half22float2(hc), fs = __half22float2(hs);
  const __half2 lc = __floats2half2_rn(c0 - fc.x, c1 - fc.y), ls = __floats2half2_rn(s0 - fs.x, s1 - fs.y);
  ch = *reinterpret_cast<const uint32_t*>(&hc);
  cl = *reinterpret_cast<const uint32_t*>(&lc);
  sh = *reinterpret_cast<const uint32_t*>(&hs);
  sl = *reinterpret_cast<const uint32_t*>(&ls);
}

template <bool REGULAR>
__global__ void __launch_bounds__(TC_THREADS, 1)
ls_tc_kernel(const __grid_constant__ CUtensorMap ymap, const TcParams p) {
  extern __shared__ unsigned char tc_smem_raw[];
  // 1024-byte aligned carve-up (swizzle atoms need their natural alignment)
  const uint32_t raw = ptx::smem_u32(tc_smem_raw);
  unsigned char* smem = tc_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)TC_STAGES * TC_STAGE_BYTES);
  uint64_t* full_y = bars;                         // [STAGES] TMA bytes landed
  uint64_t* full_a = bars + TC_STAGES;             // [STAGES] generator warps done
  uint64_t* empty = bars + 2 * TC_STAGES;          // [STAGES] MMAs of the stage retired
  uint64_t* acc_full = bars + 3 * TC_STAGES;       // a segment's accumulators are complete
  uint64_t* acc_empty = bars + 3 * TC_STAGES + 1;  // the epilogue warps have drained TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 2);
  unsigned char* scratch = reinterpret_cast<unsigned char*>(bars) + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t f0 = (int64_t)blockIdx.x * TC_BM;
  const int b0 = blockIdx.y * TC_BN;
  const int nst = (int)(p.Npad / TC_BK);           // all stages; segments are p.seg_stages long
  const uint32_t gen_ns = (p.debug & 8) ? 0u : ((p.debug & 32) ? 20u : 100u), mma_ns = (p.debug & 16) ? 0u : 40u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      ptx::mbar_init(&full_y[s], 1);
      ptx::mbar_init(&full_a[s], TC_GEN_WARPS);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_empty, TC_EPI_WARPS);
    ptx::mbar_fence_init();
    ptx::prefetch_tensormap(&ymap);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: flux hi/lo tiles =================
    if (lane == 0) {
      for (int it = 0; it < nst; ++it) {
        const int s = it % TC_STAGES;
        if (it >= TC_STAGES) ptx::mbar_wait_sleep(&empty[s], ((it / TC_STAGES) - 1) & 1, 200);
        unsigned char* st = smem + (size_t)s * TC_STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&full_y[s], 2 * TC_Y_TILE);
        ptx::tma_load_2d(st + 4 * TC_A_TILE, &ymap, it * TC_BK, b0, &full_y[s]);
        ptx::tma_load_2d(st + 4 * TC_A_TILE + TC_Y_TILE, &ymap, it * TC_BK, p.B + b0, &full_y[s]);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      for (int it = 0; it < nst; ++it) {
        const int s = it % TC_STAGES;
        const uint32_t ph = (it / TC_STAGES) & 1;
        const int seg = it / p.seg_stages;
        const bool seg_first = (it - seg * p.seg_stages) == 0;
        const bool seg_last = (it + 1 == nst) || ((it + 1) % p.seg_stages == 0);
        ptx::mbar_wait_sleep(&full_y[s], ph, mma_ns);
        ptx::mbar_wait_sleep(&full_a[s], ph, mma_ns);
        if (seg_first && seg > 0) ptx::mbar_wait_sleep(acc_empty, (seg - 1) & 1, 40);   // TMEM drained
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + (size_t)s * TC_STAGE_BYTES);
        const uint32_t a_ch = sa, a_cl = sa + TC_A_TILE, a_sh = sa + 2 * TC_A_TILE, a_sl = sa + 3 * TC_A_TILE;
        const uint32_t y_h = sa + 4 * TC_A_TILE, y_l = y_h + TC_Y_TILE;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint32_t ko = k * 32;                      // 16 fp16 = 32 bytes along K inside the 64 B row
          const uint32_t first = (seg_first && k == 0) ? 0u : 1u;
          const uint64_t dyh = tc_smem_desc(y_h + ko), dyl = tc_smem_desc(y_l + ko);
          // cos accumulator: columns [0, 256)
          ptx::umma_f16_ss(tmem, tc_smem_desc(a_ch + ko), dyh, TC_IDESC, first);
          if (!(p.debug & 1)) {
            ptx::umma_f16_ss(tmem, tc_smem_desc(a_ch + ko), dyl, TC_IDESC, 1u);
            ptx::umma_f16_ss(tmem, tc_smem_desc(a_cl + ko), dyh, TC_IDESC, 1u);
          }
          // sin accumulator: columns [256, 512)
          ptx::umma_f16_ss(tmem + TC_BN, tc_smem_desc(a_sh + ko), dyh, TC_IDESC, first);
          if (!(p.debug & 1)) {
            ptx::umma_f16_ss(tmem + TC_BN, tc_smem_desc(a_sh + ko), dyl, TC_IDESC, 1u);
            ptx::umma_f16_ss(tmem + TC_BN, tc_smem_desc(a_sl + ko), dyh, TC_IDESC, 1u);
          }
        }
        ptx::umma_commit(&empty[s]);          // smem stage reusable once these MMAs retire
        if (seg_last) ptx::umma_commit(acc_full);
      }
    }
  } else if (warp < 2 + TC_GEN_WARPS) {
    // ================= design-matrix generators =================
    // thread -> (frequency row, one 16-byte chunk = 8 cadences of the stage); a warp covers 32 rows of
    // ONE chunk, so its 8 cadences' table entries are prefetched by lanes 0..7 one stage ahead and
    // broadcast through a private shared-memory scratch (no L2 latency on the critical path).
    const int gw = warp - 2;                              // 0..15
    const int row = (gw & 3) * 32 + lane;                 // frequency row inside the tile
    const int chunk = gw >> 2;                            // which 8-cadence chunk of the stage
    const uint32_t row_off = (uint32_t)row * 64u + ((((uint32_t)chunk) ^ (uint32_t)((row >> 1) & 3)) << 4);
    ulonglong2* my_scr = reinterpret_cast<ulonglong2*>(scratch + gw * 128);
    const unsigned long long kfreq = (unsigned long long)(f0 + row);          // global frequency index
    const double fr = REGULAR ? (p.f0 + (double)(f0 + row) * p.df) : ((f0 + row < p.F) ? p.freq[f0 + row] : 0.0);
    const bool low_row = fabs(fr) <= p.lowf_max;
    // window sums of this (row, chunk): fp32 within a segment, flushed to fp64 at segment ends
    const bool do_win = (p.wsum != nullptr) && (blockIdx.y == 0);
    float w_s = 0.f, w_c = 0.f, w_cc = 0.f, w_sc = 0.f;
    double W_s = 0.0, W_c = 0.0, W_cc = 0.0, W_sc = 0.0;
    ulonglong2 nxt = make_ulonglong2(0ull, 0ull);
    auto prefetch = [&](int it) {
      if (lane < 8) {
        const int64_t n = (int64_t)it * TC_BK + chunk * 8 + lane;
        if (REGULAR) nxt = p.tab[n];
        else nxt.x = (unsigned long long)__double_as_longlong(p.t[n]);
      }
    };
    prefetch(0);
    for (int it = 0; it < nst; ++it) {
      const int s = it % TC_STAGES;
      if (lane < 8) my_scr[lane] = nxt;
      __syncwarp();
      if (it + 1 < nst) prefetch(it + 1);
      if (it >= TC_STAGES) ptx::mbar_wait_sleep(&empty[s], ((it / TC_STAGES) - 1) & 1, gen_ns);
      unsigned char* st = smem + (size_t)s * TC_STAGE_BYTES;
      uint32_t ch[4] = {0, 0, 0, 0}, cl[4] = {0, 0, 0, 0}, sh[4] = {0, 0, 0, 0}, sl[4] = {0, 0, 0, 0};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (p.debug & 2) break;
        float s0, c0, s1, c1;
        const ulonglong2 e0 = my_scr[2 * q], e1 = my_scr[2 * q + 1];
        if (REGULAR) {
          if (low_row) {
            ls_sincos_fixed_low(e0.x + kfreq * e0.y, s0, c0);
            ls_sincos_fixed_low(e1.x + kfreq * e1.y, s1, c1);
          } else {
            ls_sincos_fixed(e0.x + kfreq * e0.y, s0, c0);
            ls_sincos_fixed(e1.x + kfreq * e1.y, s1, c1);
          }
        } else {
          if (low_row) {
            ls_sincos_cycles_low(fr * __longlong_as_double((long long)e0.x), s0, c0);
            ls_sincos_cycles_low(fr * __longlong_as_double((long long)e1.x), s1, c1);
          } else {
            ls_sincos_cycles(fr * __longlong_as_double((long long)e0.x), s0, c0);
            ls_sincos_cycles(fr * __longlong_as_double((long long)e1.x), s1, c1);
          }
        }
        if (do_win) {
          const float rc0 = low_row ? 1.0f + c0 : c0, rc1 = low_row ? 1.0f + c1 : c1;   // low rows carry cos - 1
          w_s += s0 + s1;
          w_c += rc0 + rc1;
          w_cc = fmaf(rc0, rc0, fmaf(rc1, rc1, w_cc));
          w_sc = fmaf(s0, rc0, fmaf(s1, rc1, w_sc));
        }
        tc_split2(c0, c1, s0, s1, ch[q], cl[q], sh[q], sl[q]);
      }
      if (do_win && (((it + 1) % p.seg_stages) == 0 || it + 1 == nst)) {
        W_s += (double)w_s; W_c += (double)w_c; W_cc += (double)w_cc; W_sc += (double)w_sc;
        w_s = w_c = w_cc = w_sc = 0.f;
      }
      *reinterpret_cast<uint4*>(st + row_off) = make_uint4(ch[0], ch[1], ch[2], ch[3]);
      *reinterpret_cast<uint4*>(st + TC_A_TILE + row_off) = make_uint4(cl[0], cl[1], cl[2], cl[3]);
      *reinterpret_cast<uint4*>(st + 2 * TC_A_TILE + row_off) = make_uint4(sh[0], sh[1], sh[2], sh[3]);
      *reinterpret_cast<uint4*>(st + 3 * TC_A_TILE + row_off) = make_uint4(sl[0], sl[1], sl[2], sl[3]);
      if (!(p.debug & 4)) ptx::fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full_a[s]);
    }
    if (do_win && f0 + row < p.F) {
      double* o = p.wsum + ((f0 + row) * 4 + chunk) * 4;
      o[0] = W_s; o[1] = W_c; o[2] = W_cc; o[3] = W_sc;
    }
  } else {
    // ================= epilogue warps (TMEM lane quadrant = warp % 4) =================
    const int quad = warp & 3;
    const int64_t f = f0 + quad * 32 + lane;
    const bool f_ok = f < p.F;
    const float4 r = f_ok ? p.rot[f] : make_float4(1.f, 0.f, 0.f, 0.f);
    const float2 r2 = f_ok ? p.rot2[f] : make_float2(0.f, 0.f);
    const bool low_out = f_ok && fabs(p.freq[f]) <= p.lowf_max;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const float Nf = (float)p.N;
    const int64_t plane = (int64_t)p.B * p.F;
    for (int seg = 0; seg < p.nseg; ++seg) {
      ptx::mbar_wait_sleep(acc_full, seg & 1, 500);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < TC_BN; c0 += 16) {
        uint32_t vc[16], vs[16];
        ptx::tmem_ld_32x32b_x16(lane_addr + c0, vc);
        ptx::tmem_ld_32x32b_x16(lane_addr + TC_BN + c0, vs);
        ptx::tmem_ld_wait();
        if (f_ok) {
          if (p.nseg == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int b = b0 + c0 + j;
              if (b < p.B) {
                const float h = p.inv_scale[b];
                p.power[(int64_t)b * p.F + f] = ls_epilogue_shared(__uint_as_float(vc[j]) * h, __uint_as_float(vs[j]) * h,
                                                                  r, r2, p.ysum[b], Nf, p.normalization, p.norm_scale, low_out);
              }
            }
          } else {
            float* pc = p.part + (int64_t)(seg * 2) * plane + (int64_t)(b0 + c0) * p.F + f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (b0 + c0 + j < p.B) {
                pc[(int64_t)j * p.F] = __uint_as_float(vc[j]);
                pc[plane + (int64_t)j * p.F] = __uint_as_float(vs[j]);
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty);
    }
  }
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

// =====================================================================================
// Grouped-generator variant.  Timing experiments on the kernel above (LKB_TC_DEBUG, bench config 2) showed its
// step time is the SUM of three parts - 32.6 ms with neither design-matrix math nor the correction MMAs,
// +22.9 ms for the math, +20.2 ms for the 8 extra MMAs per stage - i.e. the generator warps are the serial
// bottleneck: all 16 of them fill the SAME stage in lock step, so the fixed per-stage costs (barrier wake-up,
// table hand-off, async-proxy fence, arrive) are paid once per stage on the critical path.  Here the generator
// warps form TG_GROUPS groups that fill alternate stages (each thread: one row x TG_GROUPS chunks), so one group's
// fixed costs overlap the other group's math.
// =====================================================================================
constexpr int TG_GROUPS = 2;
constexpr size_t TG_SMEM = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024 + 256 + 2 * TC_SCRATCH;

template <bool REGULAR>
__global__ void __launch_bounds__(TC_THREADS, 1)
ls_tcg_kernel(const __grid_constant__ CUtensorMap ymap, const TcParams p) {
  extern __shared__ unsigned char tc_smem_raw[];
  // 1024-byte aligned carve-up (swizzle atoms need their natural alignment)
  const uint32_t raw = ptx::smem_u32(tc_smem_raw);
  unsigned char* smem = tc_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)TC_STAGES * TC_STAGE_BYTES);
  uint64_t* full_y = bars;                         // [STAGES] TMA bytes landed
  uint64_t* full_a = bars + TC_STAGES;             // [STAGES] generator warps done
  uint64_t* empty = bars + 2 * TC_STAGES;          // [STAGES] MMAs of the stage retired
  uint64_t* acc_full = bars + 3 * TC_STAGES;       // a segment's accumulators are complete
  uint64_t* acc_empty = bars + 3 * TC_STAGES + 1;  // the epilogue warps have drained TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 2);
  unsigned char* scratch = reinterpret_cast<unsigned char*>(bars) + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t f0 = (int64_t)blockIdx.x * TC_BM;
  const int b0 = blockIdx.y * TC_BN;
  const int nst = (int)(p.Npad / TC_BK);           // all stages; segments are p.seg_stages long
  const uint32_t park_ns = (p.debug & 64) ? 0u : 20000u;     // suspend-time hint of the barrier waits

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      ptx::mbar_init(&full_y[s], 1);
      ptx::mbar_init(&full_a[s], TC_GEN_WARPS / TG_GROUPS);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_empty, TC_EPI_WARPS);
    ptx::mbar_fence_init();
    ptx::prefetch_tensormap(&ymap);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: flux hi/lo tiles =================
    if (lane == 0) {
      for (int it = 0; it < nst; ++it) {
        const int s = it % TC_STAGES;
        if (it >= TC_STAGES) ptx::mbar_wait_park(&empty[s], ((it / TC_STAGES) - 1) & 1, park_ns);
        unsigned char* st = smem + (size_t)s * TC_STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&full_y[s], 2 * TC_Y_TILE);
        ptx::tma_load_2d(st + 4 * TC_A_TILE, &ymap, it * TC_BK, b0, &full_y[s]);
        ptx::tma_load_2d(st + 4 * TC_A_TILE + TC_Y_TILE, &ymap, it * TC_BK, p.B + b0, &full_y[s]);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      for (int it = 0; it < nst; ++it) {
        const int s = it % TC_STAGES;
        const uint32_t ph = (it / TC_STAGES) & 1;
        const int seg = it / p.seg_stages;
        const bool seg_first = (it - seg * p.seg_stages) == 0;
        const bool seg_last = (it + 1 == nst) || ((it + 1) % p.seg_stages == 0);
        ptx::mbar_wait_park(&full_y[s], ph, park_ns);
        ptx::mbar_wait_park(&full_a[s], ph, park_ns);
        if (seg_first && seg > 0) ptx::mbar_wait_park(acc_empty, (seg - 1) & 1, park_ns);   // TMEM drained
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + (size_t)s * TC_STAGE_BYTES);
        const uint32_t a_ch = sa, a_cl = sa + TC_A_TILE, a_sh = sa + 2 * TC_A_TILE, a_sl = sa + 3 * TC_A_TILE;
        const uint32_t y_h = sa + 4 * TC_A_TILE, y_l = y_h + TC_Y_TILE;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint32_t ko = k * 32;                      // 16 fp16 = 32 bytes along K inside the 64 B row
          const uint32_t first = (seg_first && k == 0) ? 0u : 1u;
          const uint64_t dyh = tc_smem_desc(y_h + ko), dyl = tc_smem_desc(y_l + ko);
          // cos accumulator: columns [0, 256)
          ptx::umma_f16_ss(tmem, tc_smem_desc(a_ch + ko), dyh, TC_IDESC, first);
          ptx::umma_f16_ss(tmem, tc_smem_desc(a_ch + ko), dyl, TC_IDESC, 1u);
          ptx::umma_f16_ss(tmem, tc_smem_desc(a_cl + ko), dyh, TC_IDESC, 1u);
          // sin accumulator: columns [256, 512)
          ptx::umma_f16_ss(tmem + TC_BN, tc_smem_desc(a_sh + ko), dyh, TC_IDESC, first);
          ptx::umma_f16_ss(tmem + TC_BN, tc_smem_desc(a_sh + ko), dyl, TC_IDESC, 1u);
          ptx::umma_f16_ss(tmem + TC_BN, tc_smem_desc(a_sl + ko), dyh, TC_IDESC, 1u);
        }
        ptx::umma_commit(&empty[s]);          // smem stage reusable once these MMAs retire
        if (seg_last) ptx::umma_commit(acc_full);
      }
    }
  } else if (warp < 2 + TC_GEN_WARPS) {
    // ================= design-matrix generators =================
    // thread -> (frequency row, one 16-byte chunk = 8 cadences of the stage); a warp covers 32 rows of
    // ONE chunk, so its 8 cadences' table entries are prefetched by lanes 0..7 one stage ahead and
    // broadcast through a private shared-memory scratch (no L2 latency on the critical path).
    const int gw = warp - 2;                              // 0..15
    const int group = gw / (TC_GEN_WARPS / TG_GROUPS);    // which stages this warp fills: it = group (mod TG_GROUPS)
    const int gw8 = gw % (TC_GEN_WARPS / TG_GROUPS);
    const int row = (gw8 & 3) * 32 + lane;                // frequency row inside the tile
    const int cb = (gw8 >> 2) * TG_GROUPS;                // first of this thread's TG_GROUPS 8-cadence chunks
    ulonglong2* my_scr = reinterpret_cast<ulonglong2*>(scratch + gw * 256);
    const unsigned long long kfreq = (unsigned long long)(f0 + row);          // global frequency index
    const double fr = REGULAR ? (p.f0 + (double)(f0 + row) * p.df) : ((f0 + row < p.F) ? p.freq[f0 + row] : 0.0);
    const bool low_row = fabs(fr) <= p.lowf_max;
    ulonglong2 nxt = make_ulonglong2(0ull, 0ull);
    auto prefetch = [&](int it) {
      if (lane < 8 * TG_GROUPS) {
        const int64_t n = (int64_t)it * TC_BK + cb * 8 + lane;
        if (REGULAR) nxt = p.tab[n];
        else nxt.x = (unsigned long long)__double_as_longlong(p.t[n]);
      }
    };
    if (group < nst) prefetch(group);
    for (int it = group; it < nst; it += TG_GROUPS) {
      const int s = it % TC_STAGES;
      if (lane < 8 * TG_GROUPS) my_scr[lane] = nxt;
      __syncwarp();
      if (it + TG_GROUPS < nst) prefetch(it + TG_GROUPS);
      if (it >= TC_STAGES) ptx::mbar_wait_park(&empty[s], ((it / TC_STAGES) - 1) & 1, park_ns);
      unsigned char* st = smem + (size_t)s * TC_STAGE_BYTES;
#pragma unroll
      for (int cc = 0; cc < TG_GROUPS; ++cc) {
        const int chunk = cb + cc;
        const uint32_t row_off = (uint32_t)row * 64u + ((((uint32_t)chunk) ^ (uint32_t)((row >> 1) & 3)) << 4);
        uint32_t ch[4], cl[4], sh[4], sl[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float s0, c0, s1, c1;
          const ulonglong2 e0 = my_scr[cc * 8 + 2 * q], e1 = my_scr[cc * 8 + 2 * q + 1];
          if (REGULAR) {
            if (low_row) {
              ls_sincos_fixed_low(e0.x + kfreq * e0.y, s0, c0);
              ls_sincos_fixed_low(e1.x + kfreq * e1.y, s1, c1);
            } else {
              ls_sincos_fixed(e0.x + kfreq * e0.y, s0, c0);
              ls_sincos_fixed(e1.x + kfreq * e1.y, s1, c1);
            }
          } else {
            if (low_row) {
              ls_sincos_cycles_low(fr * __longlong_as_double((long long)e0.x), s0, c0);
              ls_sincos_cycles_low(fr * __longlong_as_double((long long)e1.x), s1, c1);
            } else {
              ls_sincos_cycles(fr * __longlong_as_double((long long)e0.x), s0, c0);
              ls_sincos_cycles(fr * __longlong_as_double((long long)e1.x), s1, c1);
            }
          }
          tc_split2(c0, c1, s0, s1, ch[q], cl[q], sh[q], sl[q]);
        }
        *reinterpret_cast<uint4*>(st + row_off) = make_uint4(ch[0], ch[1], ch[2], ch[3]);
        *reinterpret_cast<uint4*>(st + TC_A_TILE + row_off) = make_uint4(cl[0], cl[1], cl[2], cl[3]);
        *reinterpret_cast<uint4*>(st + 2 * TC_A_TILE + row_off) = make_uint4(sh[0], sh[1], sh[2], sh[3]);
        *reinterpret_cast<uint4*>(st + 3 * TC_A_TILE + row_off) = make_uint4(sl[0], sl[1], sl[2], sl[3]);
      }
      ptx::fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full_a[s]);
    }
  } else {
    // ================= epilogue warps (TMEM lane quadrant = warp % 4) =================
    const int quad = warp & 3;
    const int64_t f = f0 + quad * 32 + lane;
    const bool f_ok = f < p.F;
    const float4 r = f_ok ? p.rot[f] : make_float4(1.f, 0.f, 0.f, 0.f);
    const float2 r2 = f_ok ? p.rot2[f] : make_float2(0.f, 0.f);
    const bool low_out = f_ok && fabs(p.freq[f]) <= p.lowf_max;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const float Nf = (float)p.N;
    const int64_t plane = (int64_t)p.B * p.F;
    for (int seg = 0; seg < p.nseg; ++seg) {
      ptx::mbar_wait_park(acc_full, seg & 1, park_ns);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < TC_BN; c0 += 16) {
        uint32_t vc[16], vs[16];
        ptx::tmem_ld_32x32b_x16(lane_addr + c0, vc);
        ptx::tmem_ld_32x32b_x16(lane_addr + TC_BN + c0, vs);
        ptx::tmem_ld_wait();
        if (f_ok) {
          if (p.nseg == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int b = b0 + c0 + j;
              if (b < p.B) {
                const float h = p.inv_scale[b];
                p.power[(int64_t)b * p.F + f] = ls_epilogue_shared(__uint_as_float(vc[j]) * h, __uint_as_float(vs[j]) * h,
                                                                  r, r2, p.ysum[b], Nf, p.normalization, p.norm_scale, low_out);
              }
            }
          } else {
            float* pc = p.part + (int64_t)(seg * 2) * plane + (int64_t)(b0 + c0) * p.F + f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (b0 + c0 + j < p.B) {
                pc[(int64_t)j * p.F] = __uint_as_float(vc[j]);
                pc[plane + (int64_t)j * p.F] = __uint_as_float(vs[j]);
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty);
    }
  }
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

// =====================================================================================
// CTA-pair variant (cta_group::2).  Measured motivation: with one CTA per tile the M128 N256 K16 SS MMAs read
// 12 KB of operands per 128 clk (96 B/clk) while the generators + TMA write another 64 KB per 1536-clk stage
// (42 B/clk) - more than the 128 B/clk of one SM's shared memory, so the tensor pipe idles ~30 %.  A pair of
// CTAs (two SMs of a TPC) computes a 256-frequency x 256-light-curve tile with M = 256 MMAs issued by the
// leader: each SM supplies its own 128 design-matrix rows and only HALF of the flux tile, i.e. 8 KB per MMA
// per SM and a 48 KB stage (4 stages fit).  Barriers: flux bytes of both CTAs and the "rows ready" arrivals of
// both CTAs' generator warps are counted on the leader's mbarriers (remote arrives / cta_group::2 TMA);
// tcgen05.commit multicasts "stage free" and "accumulators complete" to both CTAs.
// =====================================================================================
constexpr int T2_STAGES = 4;
constexpr int T2_BNH = TC_BN / 2;                                  // light curves staged per CTA
constexpr int T2_Y_TILE = T2_BNH * TC_BK * 2;                      // 8 KB
constexpr int T2_STAGE_BYTES = 4 * TC_A_TILE + 2 * T2_Y_TILE;      // 48 KB
constexpr size_t T2_SMEM = (size_t)T2_STAGES * T2_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + TC_SCRATCH;
constexpr uint32_t T2_IDESC = (1u << 4) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);

template <bool REGULAR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
ls_tc2_kernel(const __grid_constant__ CUtensorMap ymap, const TcParams p) {
  extern __shared__ unsigned char tc_smem_raw[];
  const uint32_t raw = ptx::smem_u32(tc_smem_raw);
  unsigned char* smem = tc_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)T2_STAGES * T2_STAGE_BYTES);
  uint64_t* full_y = bars;                         // [STAGES] (leader) flux bytes of BOTH CTAs landed
  uint64_t* full_a = bars + T2_STAGES;             // [STAGES] (leader) generator warps of BOTH CTAs done
  uint64_t* empty = bars + 2 * T2_STAGES;          // [STAGES] (each CTA) MMAs of the stage retired
  uint64_t* acc_full = bars + 3 * T2_STAGES;       // (each CTA) a segment's accumulators are complete
  uint64_t* acc_empty = bars + 3 * T2_STAGES + 1;  // (leader) the epilogue warps of BOTH CTAs drained TMEM
  uint64_t* loc_a = bars + 3 * T2_STAGES + 2;      // [STAGES] (peer) its own generator warps done -> forwarded
  uint64_t* loc_acc = bars + 4 * T2_STAGES + 2;    // (peer) its own epilogue warps done -> forwarded
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * T2_STAGES + 3);
  unsigned char* scratch = reinterpret_cast<unsigned char*>(bars) + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int64_t f0 = (int64_t)blockIdx.x * TC_BM;           // this CTA's 128 frequency rows
  const int b_tile = blockIdx.y * TC_BN;                    // the pair's 256 light curves
  const int b_half = b_tile + (int)rank * T2_BNH;           // the half this CTA stages
  const int nst = (int)(p.Npad / TC_BK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < T2_STAGES; ++s) {
      ptx::mbar_init(&full_y[s], 1);
      ptx::mbar_init(&full_a[s], TC_GEN_WARPS + 1);      // own generator warps + the peer's forwarder
      ptx::mbar_init(&empty[s], 1);
      ptx::mbar_init(&loc_a[s], TC_GEN_WARPS);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_empty, TC_EPI_WARPS + 1);
    ptx::mbar_init(loc_acc, TC_EPI_WARPS);
    ptx::mbar_fence_init();
    ptx::prefetch_tensormap(&ymap);
  }
  if (warp == 1) {
    ptx::tmem_alloc_2cta(tmem_slot, 512);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                     // both CTAs' barriers are initialised before any remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: this CTA's half of the flux hi/lo tiles =================
    if (lane == 0) {
      for (int it = 0; it < nst; ++it) {
        const int s = it % T2_STAGES;
        if (it >= T2_STAGES) ptx::mbar_wait_sleep(&empty[s], ((it / T2_STAGES) - 1) & 1, 200);
        unsigned char* st = smem + (size_t)s * T2_STAGE_BYTES;
        if (leader) ptx::mbar_arrive_expect_tx(&full_y[s], 4 * T2_Y_TILE);        // 2 planes x 2 CTAs
        const uint32_t bar = ptx::leader_addr(&full_y[s]);
        ptx::tma_load_2d_2sm(st + 4 * TC_A_TILE, &ymap, it * TC_BK, b_half, bar);
        ptx::tma_load_2d_2sm(st + 4 * TC_A_TILE + T2_Y_TILE, &ymap, it * TC_BK, p.B + b_half, bar);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (leader && lane == 0) {
      for (int it = 0; it < nst; ++it) {
        const int s = it % T2_STAGES;
        const uint32_t ph = (it / T2_STAGES) & 1;
        const int seg = it / p.seg_stages;
        const bool seg_first = (it - seg * p.seg_stages) == 0;
        const bool seg_last = (it + 1 == nst) || ((it + 1) % p.seg_stages == 0);
        ptx::mbar_wait_sleep(&full_y[s], ph, 40);
        ptx::mbar_wait_cluster_sleep(&full_a[s], ph, 40);
        if (seg_first && seg > 0) ptx::mbar_wait_cluster_sleep(acc_empty, (seg - 1) & 1, 40);   // TMEM drained
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + (size_t)s * T2_STAGE_BYTES);
        const uint32_t a_ch = sa, a_cl = sa + TC_A_TILE, a_sh = sa + 2 * TC_A_TILE, a_sl = sa + 3 * TC_A_TILE;
        const uint32_t y_h = sa + 4 * TC_A_TILE, y_l = y_h + T2_Y_TILE;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint32_t ko = k * 32;
          const uint32_t first = (seg_first && k == 0) ? 0u : 1u;
          const uint64_t dyh = tc_smem_desc(y_h + ko), dyl = tc_smem_desc(y_l + ko);
          ptx::umma_f16_ss_2cta(tmem, tc_smem_desc(a_ch + ko), dyh, T2_IDESC, first);
          ptx::umma_f16_ss_2cta(tmem, tc_smem_desc(a_ch + ko), dyl, T2_IDESC, 1u);
          ptx::umma_f16_ss_2cta(tmem, tc_smem_desc(a_cl + ko), dyh, T2_IDESC, 1u);
          ptx::umma_f16_ss_2cta(tmem + TC_BN, tc_smem_desc(a_sh + ko), dyh, T2_IDESC, first);
          ptx::umma_f16_ss_2cta(tmem + TC_BN, tc_smem_desc(a_sh + ko), dyl, T2_IDESC, 1u);
          ptx::umma_f16_ss_2cta(tmem + TC_BN, tc_smem_desc(a_sl + ko), dyh, T2_IDESC, 1u);
        }
        ptx::umma_commit_2cta(&empty[s]);          // both CTAs' stage s reusable once these MMAs retire
        if (seg_last) ptx::umma_commit_2cta(acc_full);
      }
    } else if (!leader && lane == 0) {
      // ---- peer CTA: forward "my rows are ready" / "my TMEM is drained" to the leader's barriers.  A remote
      // release-arrive costs a MEMBAR.ALL.GPU (measured: 12 % of all stall samples when every generator warp
      // did it); here it is paid once per stage by an otherwise idle thread with no memory traffic of its own.
      const uint32_t full_a_leader = ptx::leader_addr(&full_a[0]);
      const uint32_t acc_empty_leader = ptx::leader_addr(acc_empty);
      for (int it = 0; it < nst; ++it) {
        const int s = it % T2_STAGES;
        ptx::mbar_wait_sleep(&loc_a[s], (it / T2_STAGES) & 1, 40);
        ptx::mbar_arrive_cluster(full_a_leader + 8u * (uint32_t)s);
        const bool seg_last = (it + 1 == nst) || ((it + 1) % p.seg_stages == 0);
        if (seg_last) {
          ptx::mbar_wait_sleep(loc_acc, (it / p.seg_stages) & 1, 200);
          ptx::mbar_arrive_cluster(acc_empty_leader);
        }
      }
    }
  } else if (warp < 2 + TC_GEN_WARPS) {
    // ================= design-matrix generators (as in ls_tc_kernel; "rows ready" goes to the leader) =========
    const int gw = warp - 2;
    const int row = (gw & 3) * 32 + lane;
    const int chunk = gw >> 2;
    const uint32_t row_off = (uint32_t)row * 64u + ((((uint32_t)chunk) ^ (uint32_t)((row >> 1) & 3)) << 4);
    ulonglong2* my_scr = reinterpret_cast<ulonglong2*>(scratch + gw * 128);
    const unsigned long long kfreq = (unsigned long long)(f0 + row);
    const double fr = REGULAR ? (p.f0 + (double)(f0 + row) * p.df) : ((f0 + row < p.F) ? p.freq[f0 + row] : 0.0);
    const bool low_row = fabs(fr) <= p.lowf_max;
    ulonglong2 nxt = make_ulonglong2(0ull, 0ull);
    auto prefetch = [&](int it) {
      if (lane < 8) {
        const int64_t n = (int64_t)it * TC_BK + chunk * 8 + lane;
        if (REGULAR) nxt = p.tab[n];
        else nxt.x = (unsigned long long)__double_as_longlong(p.t[n]);
      }
    };
    prefetch(0);
    for (int it = 0; it < nst; ++it) {
      const int s = it % T2_STAGES;
      if (lane < 8) my_scr[lane] = nxt;
      __syncwarp();
      if (it + 1 < nst) prefetch(it + 1);
      if (it >= T2_STAGES) ptx::mbar_wait_sleep(&empty[s], ((it / T2_STAGES) - 1) & 1, 100);
      unsigned char* st = smem + (size_t)s * T2_STAGE_BYTES;
      uint32_t ch[4], cl[4], sh[4], sl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float s0, c0, s1, c1;
        const ulonglong2 e0 = my_scr[2 * q], e1 = my_scr[2 * q + 1];
        if (REGULAR) {
          if (low_row) {
            ls_sincos_fixed_low(e0.x + kfreq * e0.y, s0, c0);
            ls_sincos_fixed_low(e1.x + kfreq * e1.y, s1, c1);
          } else {
            ls_sincos_fixed(e0.x + kfreq * e0.y, s0, c0);
            ls_sincos_fixed(e1.x + kfreq * e1.y, s1, c1);
          }
        } else {
          if (low_row) {
            ls_sincos_cycles_low(fr * __longlong_as_double((long long)e0.x), s0, c0);
            ls_sincos_cycles_low(fr * __longlong_as_double((long long)e1.x), s1, c1);
          } else {
            ls_sincos_cycles(fr * __longlong_as_double((long long)e0.x), s0, c0);
            ls_sincos_cycles(fr * __longlong_as_double((long long)e1.x), s1, c1);
          }
        }
        tc_split2(c0, c1, s0, s1, ch[q], cl[q], sh[q], sl[q]);
      }
      *reinterpret_cast<uint4*>(st + row_off) = make_uint4(ch[0], ch[1], ch[2], ch[3]);
      *reinterpret_cast<uint4*>(st + TC_A_TILE + row_off) = make_uint4(cl[0], cl[1], cl[2], cl[3]);
      *reinterpret_cast<uint4*>(st + 2 * TC_A_TILE + row_off) = make_uint4(sh[0], sh[1], sh[2], sh[3]);
      *reinterpret_cast<uint4*>(st + 3 * TC_A_TILE + row_off) = make_uint4(sl[0], sl[1], sl[2], sl[3]);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(leader ? &full_a[s] : &loc_a[s]);
    }
  } else {
    // ================= epilogue warps (own TMEM: this CTA's 128 frequency rows x the pair's 256 light curves) ====
    const int quad = warp & 3;
    const int64_t f = f0 + quad * 32 + lane;
    const bool f_ok = f < p.F;
    const float4 r = f_ok ? p.rot[f] : make_float4(1.f, 0.f, 0.f, 0.f);
    const float2 r2 = f_ok ? p.rot2[f] : make_float2(0.f, 0.f);
    const bool low_out = f_ok && fabs(p.freq[f]) <= p.lowf_max;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const float Nf = (float)p.N;
    const int64_t plane = (int64_t)p.B * p.F;
    for (int seg = 0; seg < p.nseg; ++seg) {
      ptx::mbar_wait_sleep(acc_full, seg & 1, 500);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < TC_BN; c0 += 16) {
        uint32_t vc[16], vs[16];
        ptx::tmem_ld_32x32b_x16(lane_addr + c0, vc);
        ptx::tmem_ld_32x32b_x16(lane_addr + TC_BN + c0, vs);
        ptx::tmem_ld_wait();
        if (f_ok) {
          if (p.nseg == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int b = b_tile + c0 + j;
              if (b < p.B) {
                const float h = p.inv_scale[b];
                p.power[(int64_t)b * p.F + f] = ls_epilogue_shared(__uint_as_float(vc[j]) * h, __uint_as_float(vs[j]) * h,
                                                                  r, r2, p.ysum[b], Nf, p.normalization, p.norm_scale, low_out);
              }
            }
          } else {
            float* pc = p.part + (int64_t)(seg * 2) * plane + (int64_t)(b_tile + c0) * p.F + f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (b_tile + c0 + j < p.B) {
                pc[(int64_t)j * p.F] = __uint_as_float(vc[j]);
                pc[plane + (int64_t)j * p.F] = __uint_as_float(vs[j]);
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(leader ? acc_empty : loc_acc);
    }
  }
  __syncthreads();
  ptx::cluster_sync();                     // nobody exits (or frees TMEM) while the peer may still touch this CTA
  if (warp == 1) ptx::tmem_dealloc_2cta(tmem, 512);
}

// rot / rot2 from the window sums accumulated inside ls_tc_kernel (regular grids).  The padding
// cadences (phase 0: cos = 1, sin = 0) are removed analytically.  Low-frequency rows are NOT
// written here: ls_window_kernel's full-fp64 path owns them.
__global__ void ls_tc_rot_kernel(const double* __restrict__ wsum, int64_t F, int64_t N, int64_t Npad,
                                 const double* __restrict__ freq, double lowf_max, float4* __restrict__ rot,
                                 float2* __restrict__ rot2) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F || fabs(freq[f]) <= lowf_max) return;
  LsSums<double> d;
  d.zero();
  for (int c = 0; c < 4; ++c) {
    const double* o = wsum + (f * 4 + c) * 4;
    d.s += o[0]; d.c += o[1]; d.cc += o[2]; d.sc += o[3];
  }
  const double pad = (double)(Npad - N);
  d.c -= pad;
  d.cc -= pad;
  double ct, st, cc, ss;
  ls_rotation(d, (double)N, ct, st, cc, ss);
  const double k = 1.0 / (2.0 * (double)N);
  rot[f] = make_float4((float)ct, (float)st, (float)(k / cc), (float)(k / ss));
  rot2[f] = make_float2((float)((d.c * ct + d.s * st) / (double)N), (float)((d.s * ct - d.c * st) / (double)N));
}

// sum the split-K partials (round-to-nearest fp32 adds), undo the operand scaling, apply the epilogue
__global__ void __launch_bounds__(256)
ls_tc_finish_kernel(const float* __restrict__ part, int nseg, int B, int64_t F, const float4* __restrict__ rot,
                    const float2* __restrict__ rot2, const float* __restrict__ ysum,
                    const float* __restrict__ inv_scale, const double* __restrict__ freq, double lowf_max, float N,
                    int normalization, float norm_scale, float* __restrict__ power, float low_mul) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (f >= F) return;
  const int64_t plane = (int64_t)B * F;
  const float* pc = part + (int64_t)b * F + f;
  float ch = 0.f, sh = 0.f;
  int s = 0;
  for (; s + 8 <= nseg; s += 8) {        // 16 independent loads in flight per thread
    float c[8], d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c[j] = __ldcs(pc + (int64_t)(2 * (s + j)) * plane);
      d[j] = __ldcs(pc + (int64_t)(2 * (s + j) + 1) * plane);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { ch += c[j]; sh += d[j]; }
  }
  for (; s < nseg; ++s) {
    ch += pc[(int64_t)(2 * s) * plane];
    sh += pc[(int64_t)(2 * s + 1) * plane];
  }
  const bool low = fabs(freq[f]) <= lowf_max;
  const float h = inv_scale[b] * (low ? low_mul : 1.0f);
  power[(int64_t)b * F + f] = ls_epilogue_shared(ch * h, sh * h, rot[f], rot2[f], ysum[b], N, normalization, norm_scale, low);
}

// ---- host -------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

bool ls_tc_supported(int B, int64_t N, int64_t F) {
  // worthwhile only when a flux tile is reasonably full; any shape is functionally fine
  return B >= 64 && N >= 256 && F >= 128;
}

bool ls_tc_window_in_kernel(int64_t Npad, bool regular) {
  // the generators can accumulate the window sums themselves when the epilogue is deferred to the
  // finish kernel (more than one segment) and the phases come from the fixed-point table
  int seg_cap = TC_SEG_STAGES;
  if (const char* e = getenv("LKB_TC_SEG_STAGES")) { const int v = atoi(e); if (v > 0) seg_cap = v; }
  // Measured on B200 (bench c2): folding the sums in costs the tc kernel more (65 -> 71 ms: the
  // generator warps are already the co-bottleneck) than the separate 4.4 ms window kernel, so this
  // is opt-in.
  const bool pair_off = !(getenv("LKB_TC_2CTA") != nullptr && atoi(getenv("LKB_TC_2CTA")) != 0);  // 1-CTA kernel only
  return pair_off && regular && (Npad / TC_BK) > seg_cap && getenv("LKB_TC_WINDOW_IN_KERNEL") != nullptr;
}

int ls_tc_launch(const double* d_t, const ulonglong2* d_tab, int64_t N, int64_t Npad, const float* d_yc,
                 const float* d_absmax, int B, const double* d_freq, int64_t F, float4* d_rot,
                 float2* d_rot2, bool window_in_kernel, double lowf_max, double grid_f0, double grid_df,
                 int normalization, double norm_scale, float* d_pow, cudaStream_t st, cudaEvent_t rot_ready, int ws_alt) {
  // ws_alt = 1: second set of workspace slots, so that two calls may be in flight on two streams
  const int S_YHL = ws_alt ? WS_OUT4 : WS_H, S_INV = ws_alt ? WS_OUT5 : WS_I, S_YSUM = ws_alt ? WS_OUT6 : WS_O,
            S_PART = ws_alt ? WS_OUT7 : WS_J;
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return LKB_E_CUDA; }
  __half* d_yhl = nullptr;
  float *d_inv = nullptr, *d_ysum = nullptr;
  LKB_TRY(ws_get_t<__half>(S_YHL, (size_t)2 * B * Npad, &d_yhl));
  LKB_TRY(ws_get_t<float>(S_INV, B, &d_inv));
  LKB_TRY(ws_get_t<float>(S_YSUM, B, &d_ysum));
  tc_split_flux_kernel<<<B, 256, 0, st>>>(d_yc, d_absmax, B, Npad, d_yhl, d_inv, d_ysum);
  LKB_LAUNCH_CHECK();

  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)Npad, (cuuint64_t)(2 * (int64_t)B)};
  const cuuint64_t strides[1] = {(cuuint64_t)Npad * sizeof(__half)};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BN};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_yhl, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)cr); return LKB_E_CUDA; }

  // CTA-pair kernel (LKB_TC_2CTA=1): half-height flux boxes, one per CTA of the pair.  Measured on the bench
  // (config 2, same box, back to back): 72.6 ms at 1702 MHz vs 71.2 ms at 1590 MHz for the one-CTA kernel - both
  // sit on the board's power cap (the pair variant relieves shared-memory bandwidth, the clocks drop until the
  // energy per step is the same), so the simpler one-CTA kernel stays the default.
  const bool use_pair = getenv("LKB_TC_2CTA") != nullptr && atoi(getenv("LKB_TC_2CTA")) != 0;
  CUtensorMap map2;
  if (use_pair) {
    const cuuint32_t box2[2] = {(cuuint32_t)TC_BK, (cuuint32_t)T2_BNH};
    cr = enc(&map2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_yhl, dims, strides, box2, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)cr); return LKB_E_CUDA; }
  }
  static bool attr_set = false;
  if (!attr_set) {
    LKB_CUDA_CHECK(cudaFuncSetAttribute(ls_tcg_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM));
    LKB_CUDA_CHECK(cudaFuncSetAttribute(ls_tcg_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM));
    LKB_CUDA_CHECK(cudaFuncSetAttribute(ls_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    LKB_CUDA_CHECK(cudaFuncSetAttribute(ls_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    LKB_CUDA_CHECK(cudaFuncSetAttribute(ls_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM));
    LKB_CUDA_CHECK(cudaFuncSetAttribute(ls_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM));
    attr_set = true;
  }
  const bool regular = d_tab != nullptr;
  const int nst_total = (int)(Npad / TC_BK);
  int seg_cap = TC_SEG_STAGES;
  if (const char* e = getenv("LKB_TC_SEG_STAGES")) { const int v = atoi(e); if (v > 0) seg_cap = v; }
  const int nseg0 = (nst_total + seg_cap - 1) / seg_cap;
  const int seg_stages = (nst_total + nseg0 - 1) / nseg0;     // balanced segments
  const int nseg = (nst_total + seg_stages - 1) / seg_stages; // every segment non-empty
  float* d_part = nullptr;
  if (nseg > 1) LKB_TRY(ws_get_t<float>(S_PART, (size_t)nseg * 2 * B * F, &d_part));

  TcParams p;
  p.t = d_t; p.tab = d_tab; p.freq = d_freq; p.rot = d_rot; p.rot2 = d_rot2; p.ysum = d_ysum; p.inv_scale = d_inv; p.power = d_pow; p.part = d_part;
  p.N = N; p.Npad = Npad; p.F = F; p.B = B; p.normalization = normalization; p.norm_scale = (float)norm_scale;
  p.seg_stages = seg_stages; p.nseg = nseg;
  p.lowf_max = lowf_max; p.f0 = grid_f0; p.df = grid_df;
  p.low_mul = 1.0f;
  p.debug = getenv("LKB_TC_DEBUG") ? atoi(getenv("LKB_TC_DEBUG")) : 0;
  p.wsum = nullptr;
  if (window_in_kernel && nseg > 1 && regular) LKB_TRY(ws_get_t<double>(WS_P, (size_t)F * 16, &p.wsum));
  dim3 grid((unsigned)((F + TC_BM - 1) / TC_BM), (unsigned)((B + TC_BN - 1) / TC_BN));
  if (nseg == 1) LKB_CUDA_CHECK(cudaStreamWaitEvent(st, rot_ready, 0));   // direct epilogue needs rot
  prof_begin(st);
  if (!use_pair && !(getenv("LKB_TC_GEN_GROUPS") != nullptr && atoi(getenv("LKB_TC_GEN_GROUPS")) == 1) &&
             !window_in_kernel && p.debug == 0) {
    // default: generator warps in two groups that fill alternate stages (LKB_TC_GEN_GROUPS=1: lock-step kernel)
    if (regular) ls_tcg_kernel<true><<<grid, TC_THREADS, TG_SMEM, st>>>(map, p);
    else ls_tcg_kernel<false><<<grid, TC_THREADS, TG_SMEM, st>>>(map, p);
  } else if (use_pair) {
    dim3 grid2(2u * (unsigned)((F + 2 * TC_BM - 1) / (2 * TC_BM)), grid.y);      // pairs of 128-row CTAs
    if (regular) ls_tc2_kernel<true><<<grid2, TC_THREADS, T2_SMEM, st>>>(map2, p);
    else ls_tc2_kernel<false><<<grid2, TC_THREADS, T2_SMEM, st>>>(map2, p);
  } else {
    if (regular) ls_tc_kernel<true><<<grid, TC_THREADS, TC_SMEM, st>>>(map, p);
    else ls_tc_kernel<false><<<grid, TC_THREADS, TC_SMEM, st>>>(map, p);
  }
  prof_end(st);
  LKB_LAUNCH_CHECK();
  if (p.wsum) {
    ls_tc_rot_kernel<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(p.wsum, F, N, Npad, d_freq, lowf_max, d_rot, d_rot2);
    LKB_LAUNCH_CHECK();
  }
  if (nseg > 1) {
    LKB_CUDA_CHECK(cudaStreamWaitEvent(st, rot_ready, 0));
    ls_tc_finish_kernel<<<dim3((unsigned)((F + 255) / 256), (unsigned)B), 256, 0, st>>>(
        d_part, nseg, B, F, d_rot, d_rot2, d_ysum, d_inv, d_freq, lowf_max, (float)N, normalization, (float)norm_scale,
        d_pow, p.low_mul);
    LKB_LAUNCH_CHECK();
  }
  return LKB_OK;
}

}  // namespace lkb
