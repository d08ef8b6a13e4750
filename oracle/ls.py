"""Oracle: Lomb-Scargle as lightkurve calls astropy (TEST INFRASTRUCTURE ONLY).

Reference call site: /root/reference/src/lightkurve/periodogram.py:961-975
    LS = LombScargle(time, flux, nterms=1, normalization="psd")   # dy=None,
    power = LS.power(frequency, method=ls_method)                 # fit_mean, center_data = True
    psd:       power *= 2 / (N * oversample_factor * fs)          # :969-973
    amplitude: power = sqrt(power) * sqrt(4 / N)                  # :974-975

astropy is not present in /root/reference nor installable here; the two
functions below restate astropy >= 5.0
``timeseries/periodograms/lombscargle/implementations/slow_impl.py`` and
``fast_impl.py`` + ``utils.py`` (trig_sum, extirpolate, bitceil).
PARITY UNPINNED for power values (see oracle/__init__.py); cross-checked in
tests against scipy.signal.lombscargle(floating_mean=True).
"""
import math

import numpy as np


def center(y):
    """astropy: w = dy**-2 (dy=1) ; w /= w.sum() ; y = y - dot(w, y)."""
    y = np.asarray(y, dtype=np.float64)
    w = np.ones_like(y)
    w /= w.sum()
    return y - np.dot(w, y), w


def ls_slow_psd(t, y, freq, chunk=256):
    """Exact floating-mean GLS (Zechmeister & Kuerster 2009), astropy
    ``lombscargle_slow`` with fit_mean=center_data=True, dy=1,
    normalization="psd"  ->  P = 0.5 * N * (YC^2/CC + YS^2/SS).
    """
    t = np.asarray(t, dtype=np.float64)
    freq = np.asarray(freq, dtype=np.float64)
    y, w = center(y)
    N = len(t)
    out = np.empty(len(freq), dtype=np.float64)
    Y = np.dot(w, y)
    wy = w * y
    for i0 in range(0, len(freq), chunk):
        f = freq[i0:i0 + chunk]
        omega_t = 2.0 * np.pi * f[:, None] * t[None, :]
        s = np.sin(omega_t)
        c = np.cos(omega_t)
        S2 = 2.0 * (s * c) @ w
        C2 = 2.0 * (0.5 - s * s) @ w
        S = s @ w
        C = c @ w
        S2 = S2 - 2.0 * S * C
        C2 = C2 - (C * C - S * S)
        tau = 0.5 * np.arctan2(S2, C2)
        ott = omega_t - tau[:, None]
        st = np.sin(ott)
        ct = np.cos(ott)
        Ctau = ct @ w
        Stau = st @ w
        YC = ct @ wy - Y * Ctau
        YS = st @ wy - Y * Stau
        CC = (ct * ct) @ w - Ctau * Ctau
        SS = (st * st) @ w - Stau * Stau
        with np.errstate(divide="ignore", invalid="ignore"):
            p = YC * YC / CC + YS * YS / SS
        out[i0:i0 + chunk] = p * (0.5 * N)
    return out


def bitceil(n):
    """Smallest power of two >= n (astropy utils.bitceil)."""
    n = int(n)
    return 1 << max(0, (n - 1).bit_length())


def extirpolate(x, y, N, M=4):
    """Press & Rybicki Lagrange spreading (astropy utils.extirpolate)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y).ravel()
    result = np.zeros(N, dtype=y.dtype)
    integers = x % 1 == 0
    np.add.at(result, x[integers].astype(int), y[integers])
    x, y = x[~integers], y[~integers]
    ilo = np.clip((x - M // 2).astype(int), 0, N - M)
    numerator = y * np.prod(x - ilo - np.arange(M)[:, None], 0)
    denominator = math.factorial(M - 1)
    for j in range(M):
        if j > 0:
            denominator *= j / (j - M)
        ind = ilo + (M - 1 - j)
        np.add.at(result, ind, numerator / (denominator * (x - ind)))
    return result


def trig_sum(t, h, df, N, f0=0.0, freq_factor=1, oversampling=5, use_fft=True, Mfft=4):
    """S_k = sum_j h_j sin(2 pi f_k t_j), C_k likewise (astropy utils.trig_sum)."""
    df = df * freq_factor
    f0 = f0 * freq_factor
    t = np.asarray(t, dtype=np.float64)
    h = np.asarray(h)
    if use_fft:
        t0 = t.min()
        Nfft = bitceil(N * oversampling)
        if f0 > 0:
            h = h * np.exp(2j * np.pi * f0 * (t - t0))
        tnorm = ((t - t0) * Nfft * df) % Nfft
        grid = extirpolate(tnorm, h, Nfft, Mfft)
        fftgrid = np.fft.ifft(grid)[:N]
        if t0 != 0:
            f = f0 + df * np.arange(N)
            fftgrid = fftgrid * np.exp(2j * np.pi * t0 * f)
        C = Nfft * fftgrid.real
        S = Nfft * fftgrid.imag
    else:
        f = f0 + df * np.arange(N)
        C = np.dot(h, np.cos(2 * np.pi * f * t[:, None]))
        S = np.dot(h, np.sin(2 * np.pi * f * t[:, None]))
    return S, C


def ls_fast_psd(t, y, f0, df, Nf, use_fft=True):
    """astropy ``lombscargle_fast`` (lightkurve's DEFAULT ls_method="fast",
    periodogram.py:650), regular grid f0 + df*arange(Nf), psd normalisation."""
    t = np.asarray(t, dtype=np.float64)
    y, w = center(y)
    N = len(t)
    kw = dict(f0=f0, df=df, use_fft=use_fft, N=Nf)
    Sh, Ch = trig_sum(t, w * y, **kw)
    S2, C2 = trig_sum(t, w, freq_factor=2, **kw)
    S, C = trig_sum(t, w, **kw)
    with np.errstate(divide="ignore", invalid="ignore"):
        tan_2omega_tau = (S2 - 2 * S * C) / (C2 - (C * C - S * S))
        S2w = tan_2omega_tau / np.sqrt(1 + tan_2omega_tau * tan_2omega_tau)
        C2w = 1 / np.sqrt(1 + tan_2omega_tau * tan_2omega_tau)
        Cw = np.sqrt(0.5) * np.sqrt(1 + C2w)
        Sw = np.sqrt(0.5) * np.sign(S2w) * np.sqrt(1 - C2w)
        YC = Ch * Cw + Sh * Sw
        YS = Sh * Cw - Ch * Sw
        CC = 0.5 * (1 + C2 * C2w + S2 * S2w)
        SS = 0.5 * (1 - C2 * C2w - S2 * S2w)
        CC -= (C * Cw + S * Sw) ** 2
        SS -= (S * Cw - C * Sw) ** 2
        power = YC * YC / CC + YS * YS / SS
    return power * (0.5 * N)


def design_matrix(t, frequency, bias=True, nterms=1):
    """astropy implementations.mle.design_matrix with dy = 1: [1, sin(w t), cos(w t), sin(2 w t), ...]."""
    t = np.asarray(t, dtype=np.float64)
    cols = [np.ones(len(t))] if bias else []
    for i in range(1, nterms + 1):
        cols.append(np.sin(2 * np.pi * i * frequency * t))
        cols.append(np.cos(2 * np.pi * i * frequency * t))
    return np.transpose(np.vstack(cols))


def ls_chi2_psd(t, y, freq, nterms=1):
    """astropy ``lombscargle_chi2`` (fit_mean=center_data=True, dy=1, normalization="psd"):
    P = 0.5 * XTy^T (XTX)^-1 XTy per frequency - what lightkurve gets for ls_method in
    {"chi2", "fastchi2"} with nterms >= 1 (periodogram.py:948-964)."""
    t = np.asarray(t, dtype=np.float64)
    yw, _ = center(y)
    out = np.empty(len(freq))
    for k, f in enumerate(np.asarray(freq, dtype=np.float64)):
        X = design_matrix(t, f, True, nterms)
        XTX = X.T @ X
        XTy = X.T @ yw
        try:
            out[k] = 0.5 * (XTy @ np.linalg.solve(XTX, XTy))
        except np.linalg.LinAlgError:
            out[k] = np.nan
    return out


def ls_model(t, y, frequency, t_fit, nterms=1):
    """astropy ``LombScargle.model`` / mle.periodic_fit (fit_mean=center_data=True, dy=1) as called at
    periodogram.py:1010.  Times are taken relative to t[0] like astropy does for Time inputs."""
    t = np.asarray(t, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    t0 = t[0]
    y_mean = y.mean()
    X = design_matrix(t - t0, frequency, True, nterms)
    theta = np.linalg.solve(X.T @ X, X.T @ (y - y_mean))
    return y_mean + design_matrix(np.asarray(t_fit, dtype=np.float64) - t0, frequency, True, nterms) @ theta


def is_regular(frequency):
    """astropy implementations.main._is_regular (periodogram.py:933)."""
    frequency = np.asarray(frequency)
    if frequency.ndim != 1:
        return False
    if len(frequency) == 1:
        return True
    d = np.diff(frequency)
    return bool(np.allclose(d[0], d))


def default_frequency_grid(t, oversample_factor=5.0, nyquist_factor=1,
                           minimum_frequency=None, maximum_frequency=None):
    """lightkurve grid (periodogram.py:850-911), freq_unit = 1/day."""
    t = np.asarray(t, dtype=np.float64)
    nyquist = 0.5 * (1.0 / np.median(np.diff(t)))
    fs = (1.0 / (t[-1] - t[0])) / oversample_factor
    fmin = fs if minimum_frequency is None else minimum_frequency
    fmax = nyquist * nyquist_factor if maximum_frequency is None else maximum_frequency
    return np.arange(fmin, fmax, fs), fs, nyquist


def lk_normalize(power_psd, n_time, normalization="amplitude", oversample_factor=5.0, fs=None):
    """lightkurve rescale, periodogram.py:969-975."""
    if normalization == "psd":
        return power_psd * (2.0 / (n_time * oversample_factor * fs))
    return np.sqrt(power_psd) * np.sqrt(4.0 / n_time)


def lombscargle(t, y, frequency=None, normalization="amplitude", ls_method="fast",
                oversample_factor=None, nterms=1, **grid_kw):
    """End-to-end restatement of LombScarglePeriodogram.from_lightcurve numerics
    for finite inputs (NaNs must be dropped by the caller, periodogram.py:785-790).
    Returns (frequency, power, ls_method_used)."""
    t = np.asarray(t, dtype=np.float64)
    if oversample_factor is None:
        oversample_factor = 5.0 if normalization == "amplitude" else 1.0
    grid, fs, _ = default_frequency_grid(t, oversample_factor, **grid_kw)
    if frequency is None:
        frequency = grid
    frequency = np.asarray(frequency, dtype=np.float64)
    if not is_regular(frequency) and ls_method in ("fast", "fastchi2"):
        ls_method = {"fast": "slow", "fastchi2": "chi2"}[ls_method]
    if ls_method in ("chi2", "fastchi2"):
        p = ls_chi2_psd(t - t[0], y, frequency, nterms)      # (fastchi2 = FFT approximation of the same sums)
    elif ls_method == "fast":
        f0 = frequency[0]
        df = frequency[1] - frequency[0]
        p = ls_fast_psd(t, y, f0, df, len(frequency))
    else:
        p = ls_slow_psd(t, y, frequency)
    return frequency, lk_normalize(p, len(t), normalization, oversample_factor, fs), ls_method
