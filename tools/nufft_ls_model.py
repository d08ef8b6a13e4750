"""Design study for round 2 (CPU, numpy): Lomb-Scargle trig sums by a type-1 NUFFT instead of the O(N*F)
contraction - the algorithm behind lightkurve's optional ``ls_method="fastnifty"`` (nifty-ls / finufft,
/root/reference/pyproject.toml:48, periodogram.py:917-946) - with the accuracy it reaches against the
exact sums, in the precision a GPU kernel would use.

    S(f_k) = sum_n y_n exp(2 pi i f_k t_n),  f_k = f0 + k df,  k = 0 .. F-1
           = sum_n a_n exp(i k x_n),  a_n = y_n exp(2 pi i f0 (t_n - t0)),  x_n = 2 pi df (t_n - t0)

  1. spread the strengths a_n onto a fine periodic grid of M = 2^ceil(log2(sigma * 2F)) cells with the
     "exponential of semicircle" kernel phi(z) = exp(beta (sqrt(1 - z^2) - 1)), |z| <= 1, of width w cells
     (Barnett, Magland & af Klinteberg 2019); for a SHARED cadence grid the (cell, cadence) weights are the same
     for every light curve of the batch - the spreading is a sparse [M x N] times dense [N x B] product;
  2. one length-M FFT per light curve;
  3. divide mode k by the kernel's Fourier coefficient, undo the t0 shift, feed (Ch, Sh) to the same
     floating-mean epilogue the contraction kernels use (window terms come from the same transform of a_n = 1
     at f and 2f, once per cadence grid).

Work per light curve: N*w + 2.5 M log2 M flops instead of 4 N F - for BASELINE config 2 (N = 65 000,
F = 1e5, M = 2^19) that is 5e7 instead of 2.6e10, and the whole batch becomes an HBM sweep (the fine grids of
1024 light curves are 2 GB in fp32).

Usage:  python tools/nufft_ls_model.py [w] [float32|float64]
prints the worst tolerance excess  |P - P_exact| / (1e-5 max(P) + 1e-4 P)  (the parity bound of
tests/test_gpu_engine.py) for a noise-only and a strong-signal light curve on the config-2 cadence grid.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def es_kernel(z, beta):
    out = np.zeros_like(z)
    m = np.abs(z) < 1
    out[m] = np.exp(beta * (np.sqrt(1 - z[m] ** 2) - 1))
    return out


def es_kernel_ft(k, M, w, beta, nquad=400):
    """Fourier coefficients of the periodised kernel on the M-cell grid at integer modes k (Gauss-Legendre)."""
    z, wq = np.polynomial.legendre.leggauss(nquad)
    phi = np.exp(beta * (np.sqrt(1 - z ** 2) - 1))
    # phi lives on |x| <= w/2 cells = (w/2) * 2 pi / M in angle; integral of phi(x) exp(-i k x) dx
    half = 0.5 * w * 2 * np.pi / M
    return half * (np.cos(np.outer(k, z) * half) @ (wq * phi))


def trig_sums_nufft(t, a, f0, df, F, w=8, sigma=2.0, dtype=np.float64):
    """(C, S)[k] = sum_n a_n (cos, sin)(2 pi f_k t_n) for real strengths a; returns float64 arrays."""
    cdt = np.complex64 if dtype == np.float32 else np.complex128
    t0 = t.min()
    M = 1 << int(np.ceil(np.log2(sigma * 2 * F)))
    beta = 2.30 * w
    x = (df * (t - t0)) % 1.0 * M                        # fine-grid coordinate in cells (fp64 phase, as in K1/K2)
    strengths = (a * np.exp(2j * np.pi * f0 * (t - t0))).astype(cdt)
    grid = np.zeros(M, dtype=cdt)
    i0 = np.ceil(x - 0.5 * w).astype(np.int64)           # leftmost cell within the kernel support
    for j in range(w):
        cell = i0 + j
        z = ((cell - x) / (0.5 * w)).astype(dtype)
        wt = es_kernel(z, dtype(beta)).astype(dtype)
        np.add.at(grid, cell % M, strengths * wt)
    spec = np.fft.ifft(grid.astype(cdt)) * M             # sum_m g_m exp(+2 pi i k m / M)
    spec = spec.astype(cdt)[:F]
    k = np.arange(F)
    phihat = es_kernel_ft(k, M, w, beta) * M / (2 * np.pi)
    c = spec.astype(np.complex128) / phihat
    c = c * np.exp(2j * np.pi * t0 * (f0 + df * k))
    return c.real, c.imag


def ls_power_from_sums(N, Ch, Sh, C, S, C2, S2):
    """The floating-mean epilogue (astropy slow_impl / ls_common.cuh), y already centred, w = 1/N."""
    Ch, Sh, C, S, C2, S2 = (v / N for v in (Ch, Sh, C, S, C2, S2))
    tan_num = S2 - 2 * S * C
    tan_den = C2 - (C * C - S * S)
    tau2 = np.arctan2(tan_num, tan_den)
    ct, st = np.cos(0.5 * tau2), np.sin(0.5 * tau2)
    c2t, s2t = np.cos(tau2), np.sin(tau2)
    YC = Ch * ct + Sh * st
    YS = Sh * ct - Ch * st
    Ct, St = C * ct + S * st, S * ct - C * st
    CC = 0.5 * (1 + C2 * c2t + S2 * s2t) - Ct * Ct
    SS = 0.5 * (1 - C2 * c2t - S2 * s2t) - St * St
    return 0.5 * N * (YC * YC / CC + YS * YS / SS)


def ls_psd_nufft(t, y, f0, df, F, w=8, dtype=np.float64):
    y = y - y.mean()
    ones = np.ones_like(t)
    Ch, Sh = trig_sums_nufft(t, y, f0, df, F, w, dtype=dtype)
    C, S = trig_sums_nufft(t, ones, f0, df, F, w, dtype=np.float64)          # once per cadence grid: keep fp64
    C2, S2 = trig_sums_nufft(t, ones, 2 * f0, 2 * df, F, w, dtype=np.float64)
    return ls_power_from_sums(len(t), Ch, Sh, C, S, C2, S2)


def main():
    import bench
    from oracle import ls as ols
    w = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dtype = np.float32 if (len(sys.argv) > 2 and sys.argv[2] == "float32") else np.float64
    t, Y, freq = bench.make_workload_sample("c2", 1002, n_lc=4)
    rng = np.random.default_rng(7)
    N, F = len(t), len(freq)
    f0, df = float(freq[0]), float(freq[1] - freq[0])
    cases = {"noise only": (1 + 3e-4 * rng.standard_normal(N)).astype(np.float32).astype(np.float64),
             "3 sinusoids + noise (bench LC 0)": Y[0].astype(np.float64),
             "strong line": 1 + 1e-2 * np.sin(2 * np.pi * 3.3217 * t) + 1e-4 * rng.standard_normal(N)}
    sel = np.unique(np.concatenate([np.arange(0, F, 97), np.arange(0, 400), np.arange(F - 400, F)]))
    for name, y in cases.items():
        t0 = time.perf_counter()
        p = ls_psd_nufft(t, y, f0, df, F, w=w, dtype=dtype)
        dt = time.perf_counter() - t0
        amp = np.sqrt(np.maximum(p, 0)) * np.sqrt(4.0 / N)
        ref_sel = np.sqrt(ols.ls_slow_psd(t, y, freq[sel])) * np.sqrt(4.0 / N)
        # max(P) of the exact spectrum: the peak is inside `sel` only by luck, so take it from the NUFFT result
        pmax = amp.max()
        excess = np.abs(amp[sel] - ref_sel) / (1e-5 * pmax + 1e-4 * ref_sel)
        print("%-34s w=%d %s  worst tolerance excess %.3f (median %.4f) over %d bins   [%.1f s]"
              % (name, w, np.dtype(dtype).name, excess.max(), np.median(excess), len(sel), dt))


if __name__ == "__main__":
    main()
