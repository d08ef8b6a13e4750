"""CPU tests that PIN THE ORACLE (no GPU): the reference's own behavioural / known-answer tests
for the hot path ported onto oracle/, the committed golden vectors, and
independent cross-checks (scipy.signal.lombscargle floating_mean=True; numpy vs C BLS)."""
import os

import numpy as np
import pytest
import scipy.signal

from oracle import bls as obls
from oracle import detrend as odet
from oracle import ls as ols

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------ Lomb-Scargle
def test_ls_slow_matches_scipy_floating_mean():
    """Independent third-party implementation of the same estimator (Zechmeister & Kuerster)."""
    rng = np.random.default_rng(1)
    t = np.sort(rng.uniform(0, 100, 500))
    y = 1 + 0.1 * rng.normal(size=500) + 0.3 * np.sin(2 * np.pi * t / 7.3)
    f = np.linspace(0.01, 2, 300)
    p = ols.ls_slow_psd(t, y, f)
    q = scipy.signal.lombscargle(t, y, 2 * np.pi * f, floating_mean=True, normalize=False)
    np.testing.assert_allclose(p, q, rtol=1e-10)


def test_ls_fast_without_fft_equals_slow_and_fft_is_close():
    rng = np.random.default_rng(2)
    t = np.sort(rng.uniform(0, 60, 400))
    y = 1 + 0.05 * rng.normal(size=400) + 0.2 * np.sin(2 * np.pi * t / 3.1)
    f0, df, nf = 0.02, 0.01, 250
    f = f0 + df * np.arange(nf)
    slow = ols.ls_slow_psd(t, y, f)
    np.testing.assert_allclose(ols.ls_fast_psd(t, y, f0, df, nf, use_fft=False), slow, rtol=1e-9)
    fast = ols.ls_fast_psd(t, y, f0, df, nf)
    assert np.max(np.abs(fast - slow)) < 5e-3 * slow.max()
    assert np.argmax(fast) == np.argmax(slow)


def test_reference_periodogram_can_find_periods():
    """/root/reference/tests/test_periodogram.py:102-114"""
    rng = np.random.default_rng(1001)
    t = np.arange(1000.0)
    y = rng.normal(1, 0.1, 1000) + np.sin((t / t.max()) * 20 * np.pi)
    y /= np.median(y)
    f, p, method = ols.lombscargle(t, y)
    assert len(f) == 2497 and method == "fast"
    assert np.isclose(1 / f[np.nanargmax(p)], 100, rtol=1e-3)


def test_reference_masked_flux_nans_gives_exact_zeros():
    """/root/reference/tests/test_periodogram.py:445-457 (after NaN removal: t=[1,3,4], flux=1)"""
    f, p, _ = ols.lombscargle([1.0, 3.0, 4.0], [1.0, 1.0, 1.0])
    assert not np.isnan(p).all()
    assert (p == 0).all()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_reference_ls_method_basics_beta_lyr(dtype):
    """/root/reference/tests/test_periodogram.py:460-488: fast / nterms=1 finds pi to 1 decimal"""
    t = np.arange(0, 30, 0.1)
    f = np.array(np.sin(t * 2 + np.pi / 2) + np.sin(t) + 1, dtype=dtype)
    f = f / np.median(f)
    fr, p, _ = ols.lombscargle(t + 2457000, f)
    np.testing.assert_almost_equal(1 / fr[np.nanargmax(p)], np.pi, decimal=1)


def test_reference_uneven_grid_switches_to_slow():
    """/root/reference/tests/test_periodogram.py:491-515"""
    t = np.arange(0, 30, 0.1)
    f = np.sin(t * 2 + np.pi / 2) + np.sin(t) + 1
    f /= np.median(f)
    freq = 1 / np.arange(1, 10, 0.01)
    fr, p, method = ols.lombscargle(t, f, frequency=freq)
    assert method == "slow"
    np.testing.assert_almost_equal(1 / fr[np.nanargmax(p)], np.pi, decimal=1)


def test_amplitude_normalisation_returns_the_sine_amplitude():
    t = np.arange(0, 100, 0.02)
    y = 1 + 0.0123 * np.sin(2 * np.pi * 0.71 * t)
    f, p, _ = ols.lombscargle(t, y, ls_method="slow", minimum_frequency=0.6, maximum_frequency=0.8)
    assert abs(p.max() - 0.0123) < 2e-4


# ------------------------------------------------------------------ BLS
def _transit_lc(rng):
    time = np.arange(0, 20, 0.02)
    flux = np.ones_like(time)
    flux[np.abs((time - 0.5 + 0.5 * 2.0) % 2.0 - 0.5 * 2.0) < 0.5 * 0.1] = 1.0 - 0.2
    flux += 0.01 * rng.normal(size=len(time))
    return time, flux


def test_reference_bls_period_recovery():
    """/root/reference/tests/test_periodogram.py:331-361"""
    time, flux = _transit_lc(np.random.default_rng(3))
    r = obls.boxleastsquares(time, flux)
    np.testing.assert_almost_equal(r["period"][np.argmax(r["power"])], 2.0, decimal=2)
    keep = np.ones(len(time), bool)
    keep[10] = False                                     # "sneaky NaN" is dropped by remove_nans
    r = obls.boxleastsquares(time[keep], flux[keep], dy=np.full(keep.sum(), np.nan))
    np.testing.assert_almost_equal(r["period"][np.argmax(r["power"])], 2.0, decimal=2)


def test_reference_bls_explicit_periods_and_errors():
    """/root/reference/tests/test_periodogram.py:434-442"""
    r = obls.boxleastsquares([1.0, 2.0, 3.0], [4.0, 5.0, 6.0], period=[1, 2, 3, 4, 5], impl="numpy")
    assert np.array_equal(r["period"], [1, 2, 3, 4, 5])
    with pytest.raises(ValueError, match="maximum transit duration"):
        obls.bls_power_c(np.arange(10.0), np.ones(10), None, [0.1], [0.2])


def test_bls_c_and_numpy_restatements_agree():
    rng = np.random.default_rng(4)
    time, flux = _transit_lc(rng)
    dy = 0.01 * rng.uniform(0.8, 1.2, len(time))
    dur = np.array([0.05, 0.1, 0.2])
    per = obls.autoperiod(time, dur, 0.5, 6.0, frequency_factor=30)
    for objective in ("likelihood", "snr"):
        a = obls.bls_power_c(time, flux, dy, per, dur, objective=objective, return_bins=True)
        b = obls.bls_power_numpy(time, flux, dy, per, dur, objective=objective, return_bins=True)
        assert np.array_equal(a["bins"], b["bins"])
        for k in obls.RESULT_FIELDS:
            np.testing.assert_allclose(a[k], b[k], rtol=1e-10, atol=1e-13, err_msg=k)
    t = np.sort(rng.uniform(0, 27, 5000))
    assert np.array_equal(obls.bin_index_c(t - t.min(), 0.0, 1.2345, 0.005),
                          obls.bin_index(t - t.min(), 0.0, 1.2345, 0.005))


def test_bls_autoperiod_matches_lightkurve_size_guard():
    """periodogram.py:1138-1158: npoints formula vs the grid autoperiod actually builds"""
    time = np.linspace(0, 10, 200)
    dur = obls.DEFAULT_DURATIONS
    lo, hi = obls.lk_default_period_bounds(time, dur)
    per = obls.autoperiod(time, dur, lo, hi, frequency_factor=10)
    df = 10 * np.min(dur) / (time.max() - time.min()) ** 2
    assert abs(len(per) - int((1 / lo - 1 / hi) / df)) <= 2
    assert np.all(np.diff(per) > 0) and np.isclose(per[0], lo) and abs(1 / per[-1] - 1 / hi) <= df


# ------------------------------------------------------------------ flatten / sigma_clip / regression
def test_reference_flatten_known_answers():
    """/root/reference/tests/test_lightcurve.py:1297-1317, 1344-1360, 1284-1294"""
    t = np.arange(6.0)
    f = np.array([10, 20, 30, 40, 50, 60])
    flat, _, trend = odet.flatten(t, f, window_length=3, polyorder=1)
    np.testing.assert_allclose(flat, 1.0)
    flat, _, trend = odet.flatten(t, f, window_length=7, polyorder=1)        # median fallback
    np.testing.assert_allclose(trend, np.median(f))
    odet.flatten(t, f, window_length=3, polyorder=5)                          # clamp
    odet.flatten(t, f, window_length=3, polyorder=1, break_tolerance=None)
    x = np.arange(2000.0)
    y = np.sin(np.arange(2000) / 200) / 100 + 1
    y[250] -= 0.01
    flat, _, _ = odet.flatten(x, y, window_length=25, niters=2, sigma=3)
    assert np.isclose(flat, 1, rtol=1e-5).sum() == 1999
    flat, flat_err, _ = odet.flatten([1.0, 2, 3, 4, 5], [np.nan, 1.1, 1.2, np.nan, 1.4],
                                     [1.0, np.nan, 1.2, 1.3, np.nan], window_length=3)
    assert len(flat) == 5 and np.isfinite(flat).sum() == 3 and np.isfinite(flat_err).sum() == 3


def test_sigma_clip_mask_semantics():
    rng = np.random.default_rng(5)
    x = rng.normal(size=1000)
    x[[3, 77]] = [9.0, -11.0]
    x[5] = np.nan
    m = odet.sigma_clip_mask(x, sigma=5)
    assert m[3] and m[77] and m[5] and m.sum() == 3
    # iterative: a second, smaller outlier only falls after the first is removed
    y = np.concatenate([np.zeros(50) + 1e-3 * rng.normal(size=50), [100.0, 0.05]])
    assert odet.sigma_clip_mask(y, sigma=5)[-1]


def test_reference_regression_known_answers():
    """/root/reference/tests/correctors/test_regressioncorrector.py:13-83"""
    X = np.array([[1.0, 1.0], [1.0, 2.0]])
    y = np.array([5.0, 10.0])
    np.testing.assert_almost_equal(odet.regress(X, y)["coefficients"], [0, 5])
    np.testing.assert_almost_equal(odet.regress(X, y, flux_err=np.array([0.1, 0.1]))["coefficients"], [0, 5])
    r = odet.regress(X, y, prior_mu=np.array([99.0, 99.0]), prior_sigma=np.array([1e-9, 1e-9]))
    np.testing.assert_almost_equal(r["coefficients"], [99, 99])
    r = odet.regress(X, y, prior_mu=np.array([99.0, 99.0]), prior_sigma=np.array([1e9, 1e9]))
    np.testing.assert_almost_equal(r["coefficients"], [0, 5])
    # sinusoid + noise is removed exactly (:51-83): corrected_lc.normalize().flux == 1
    size = 100
    time = np.linspace(1, 100, size)
    true_flux = np.ones(size)
    noise = np.sin(time / 5)
    X2 = np.vstack([noise, np.ones(size)]).T
    for kw in (dict(flux_err=0.1 * np.ones(size)),
               dict(flux_err=0.1 * np.ones(size), prior_mu=np.array([0.1, 0.1]), prior_sigma=np.array([1e6, 1e6])),
               dict()):
        r = odet.regress(X2, true_flux + noise, **kw)
        np.testing.assert_almost_equal(r["corrected"] / np.median(r["corrected"]), true_flux)


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name", ["ls_c1", "bls_small", "flatten_small", "regress_small"])
def test_golden_vectors_reproduce(name):
    """tests/golden/*.npz were written by tests/golden/make_golden.py (same oracle, pinned in time:
    any later edit of the oracle that changes numbers is caught here)."""
    path = os.path.join(GOLDEN, name + ".npz")
    g = np.load(path)
    if name == "ls_c1":
        f, p, _ = ols.lombscargle(g["t"], g["y"], ls_method="slow")
        np.testing.assert_allclose(f, g["frequency"], rtol=1e-14)
        np.testing.assert_allclose(p, g["power_slow"], rtol=1e-9, atol=1e-14)
        _, pf, _ = ols.lombscargle(g["t"], g["y"], ls_method="fast")
        np.testing.assert_allclose(pf, g["power_fast"], rtol=1e-7, atol=1e-12)
    elif name == "bls_small":
        r = obls.bls_power_c(g["t"], g["y"], g["dy"], g["period"], g["durations"], return_bins=True)
        assert np.array_equal(r["bins"], g["bins"])
        for k in obls.RESULT_FIELDS:
            np.testing.assert_allclose(r[k], g[k], rtol=1e-12, atol=1e-15, err_msg=k)
    elif name == "flatten_small":
        flat, _, trend = odet.flatten(g["t"], g["f"], window_length=int(g["window_length"]), niters=3)
        np.testing.assert_allclose(trend, g["trend"], rtol=1e-12)
    else:
        r = odet.regress(g["X"], g["y"], g["fe"], sigma=5, niters=5)
        np.testing.assert_allclose(r["coefficients"], g["coefficients"], rtol=1e-9)
        assert np.array_equal(r["outlier_mask"], g["outlier_mask"])
