"""Seismology front end on the host (no GPU): the reference's own tests of the estimators, which work on a
synthetic signal-to-noise spectrum (/root/reference/tests/seismology/test_butler.py:28-135,180-240 and
test_stellar_estimators.py), ported to lightkurve_b200.  `Seismology.from_lightcurve` (GPU periodogram + GPU
log-median background) is covered by tests/test_gpu_shim.py."""
import numpy as np
import pytest
from scipy.signal import unit_impulse as deltafn

from lightkurve_b200 import units as u
from lightkurve_b200.periodogram import Periodogram, SNRPeriodogram
from lightkurve_b200.seismology import (DELTANU_SOL, G_SOL, NUMAX_SOL, TEFF_SOL, Seismology, estimate_logg,
                                        estimate_mass, estimate_radius)
from lightkurve_b200.utils import LightkurveWarning

cM, cR, clogg = (1.30, 0.09), (9.91, 0.24), (2.559, 0.009)
ceteff, cenumax, cedeltanu = 80, 0.75, 0.012
cteff, cnumax, cdeltanu = 4531, 46.12, 4.934


def generate_test_spectrum():
    """A simple solar-like oscillator spectrum (test_butler.py:28-46)."""
    f = np.arange(0, 4000.0, 0.4)
    p = np.ones(len(f))
    nmx = 2500.0
    fs = f.max() / len(f)
    s = 0.25 * nmx / 2.335
    p *= 10 * np.exp(-0.5 * (f - nmx) ** 2 / s ** 2)
    m = np.zeros(len(f))
    lo, hi = int(np.floor(0.5 * nmx / fs)), int(np.floor(1.5 * nmx / fs))
    deltanu_true = 0.294 * nmx ** 0.772
    for modeloc in np.arange(lo, hi, deltanu_true / 2, dtype=int):
        m += deltafn(len(f), modeloc)
    p *= m
    p += 1
    return f, p, nmx, deltanu_true


def _snr(f, p, unit=u.microhertz):
    return SNRPeriodogram(u.Quantity(f, unit), u.Quantity(p, None))


def test_estimate_numax_basics():
    f, p, true_numax, _ = generate_test_spectrum()
    snr = _snr(f, p)
    numax = snr.to_seismology().estimate_numax()
    assert np.isclose(true_numax, numax.value, atol=0.1 * true_numax)
    assert numax.unit == u.microhertz
    assert "numax" in repr(numax) and "ACF2D" in repr(numax)
    rsnr = snr[(snr.frequency.value > 1600) & (snr.frequency.value < 3200)]
    numax = rsnr.to_seismology().estimate_numax()
    assert np.isclose(true_numax, numax.value, atol=0.1 * true_numax)
    fday = u.Quantity(f, u.microhertz).to(1 / u.day)
    numax = SNRPeriodogram(fday, u.Quantity(p, None)).to_seismology().estimate_numax()
    nmxday = u.Quantity(true_numax, u.microhertz).to(1 / u.day)
    assert np.isclose(nmxday.value, numax.value, atol=0.1 * nmxday.value)
    assert numax.unit == 1 / u.day
    f2 = f + np.random.default_rng(0).uniform(size=len(f))
    with pytest.raises(ValueError, match="uniformly spaced"):
        _snr(f2, p).to_seismology().estimate_numax()


def test_estimate_numax_kwargs():
    f, p, true_numax, _ = generate_test_spectrum()
    std = 0.25 * true_numax / 2.335
    butler = _snr(f, p).to_seismology()
    numaxs = np.linspace(true_numax - 2 * std, true_numax + 2 * std, 500)
    assert np.isclose(butler.estimate_numax(numaxs=numaxs).value, true_numax, atol=0.1 * true_numax)
    for bad in (np.linspace(-5, 5.0), np.linspace(1.0, 5000.0)):
        with pytest.raises(ValueError):
            butler.estimate_numax(numaxs=bad)
    assert np.isclose(butler.estimate_numax(window_width=200.0).value, true_numax, atol=0.1 * true_numax)
    ww = u.Quantity(200.0, u.microhertz).to(1 / u.day)
    assert np.isclose(butler.estimate_numax(window_width=ww).value, true_numax, atol=0.1 * true_numax)
    for bad in (-5, 1e6, 0.001):
        with pytest.raises(ValueError):
            butler.estimate_numax(window_width=bad)
        with pytest.raises(ValueError):
            butler.estimate_numax(spacing=bad)
    assert np.isclose(butler.estimate_numax(spacing=15.0).value, true_numax, atol=0.1 * true_numax)
    sp = u.Quantity(15.0, u.microhertz).to(1 / u.day)
    assert np.isclose(butler.estimate_numax(spacing=sp).value, true_numax, atol=0.1 * true_numax)
    daynumaxs = u.Quantity(numaxs, u.microhertz).to(1 / u.day)
    numax = butler.estimate_numax(numaxs=daynumaxs)
    assert np.isclose(numax.value, true_numax, atol=0.1 * true_numax) and numax.unit == u.microhertz
    assert set(numax.diagnostics) == {"numaxs", "acf2d", "window_width", "metric", "metric_smooth"}


def test_estimate_deltanu_basics_and_kwargs():
    f, p, _, true_deltanu = generate_test_spectrum()
    snr = _snr(f, p)
    butler = snr.to_seismology()
    with pytest.raises(AttributeError, match="estimate_numax"):
        butler.estimate_deltanu()
    numax = butler.estimate_numax()
    deltanu = butler.estimate_deltanu()
    assert np.isclose(true_deltanu, deltanu.value, atol=0.25 * true_deltanu)
    assert deltanu.unit == u.microhertz
    rsnr = snr[(snr.frequency.value > 1600) & (snr.frequency.value < 3200)]
    b2 = rsnr.to_seismology()
    b2.estimate_numax()
    assert np.isclose(true_deltanu, b2.estimate_deltanu().value, atol=0.25 * true_deltanu)
    fday = u.Quantity(f, u.microhertz).to(1 / u.day)
    b3 = SNRPeriodogram(fday, u.Quantity(p, None)).to_seismology()
    b3.estimate_numax()
    dday = u.Quantity(true_deltanu, u.microhertz).to(1 / u.day)
    assert np.isclose(dday.value, b3.estimate_deltanu().value, atol=0.25 * dday.value)
    f2 = f + np.random.default_rng(0).uniform(size=len(f))
    with pytest.raises(ValueError, match="uniformly spaced"):
        _snr(f2, p).to_seismology().estimate_deltanu(numax=100)
    # kwargs
    assert np.isclose(butler.estimate_deltanu(numax=numax).value, true_deltanu, atol=0.25 * true_deltanu)
    for bad in (-5.0, 5000):
        with pytest.raises(ValueError):
            butler.estimate_deltanu(numax=bad)
    daynumax = u.Quantity(numax.value, u.microhertz).to(1 / u.day)
    d = butler.estimate_deltanu(numax=daynumax)
    assert np.isclose(d.value, true_deltanu, atol=0.25 * true_deltanu) and d.unit == u.microhertz


def test_constants():
    assert (NUMAX_SOL.n, NUMAX_SOL.s) == (3090.0, 30.0)
    assert (DELTANU_SOL.n, DELTANU_SOL.s) == (135.1, 0.1)
    assert (TEFF_SOL.n, TEFF_SOL.s) == (5772.0, 0.8)
    assert np.isclose(G_SOL.value, 27420, rtol=1e-4) and G_SOL.unit == u.cm / u.second ** 2


def _check(q, ref):
    assert np.isclose(q.value, ref[0], atol=ref[1])
    assert np.isclose(q.error.value, ref[1], atol=0.1)


def test_scaling_relations_known_answers():
    """test_stellar_estimators.py:16-230 (values of a red giant; errors by linear propagation)."""
    R = estimate_radius(cnumax, cdeltanu, cteff)
    assert R.unit == u.solRad and np.isclose(R.value, cR[0], rtol=cR[1])
    R = estimate_radius(u.Quantity(cnumax, u.microhertz).to(1 / u.day), u.Quantity(cdeltanu, u.microhertz).to(u.hertz),
                        u.Quantity(cteff, u.Kelvin))
    assert np.isclose(R.value, cR[0], rtol=cR[1])
    R = estimate_radius(cnumax, cdeltanu, cteff, cenumax, cedeltanu, ceteff)
    assert R.error.unit == u.solRad
    _check(R, cR)
    _check(estimate_radius(cnumax, cdeltanu, cteff, u.Quantity(cenumax, u.microhertz).to(1 / u.day), cedeltanu, ceteff), cR)
    M = estimate_mass(cnumax, cdeltanu, cteff)
    assert M.unit == u.solMass and np.isclose(M.value, cM[0], rtol=cM[1])
    _check(estimate_mass(cnumax, cdeltanu, cteff, cenumax, cedeltanu, ceteff), cM)
    logg = estimate_logg(cnumax, cteff)
    assert logg.unit == u.dex and np.isclose(logg.value, clogg[0], rtol=clogg[1])
    logg = estimate_logg(cnumax, cteff, cenumax, ceteff)
    assert logg.error.unit == u.dex
    _check(logg, clogg)
    _check(estimate_logg(u.Quantity(cnumax, u.microhertz).to(1 / u.day), cteff, cenumax, u.Quantity(ceteff, u.Kelvin)), clogg)
    # without ALL errors given, only the solar constants contribute
    assert estimate_radius(cnumax, cdeltanu, cteff, cenumax).error.value < 0.15


def test_stellar_estimator_calls_through_seismology():
    """test_butler.py:288-314."""
    f, p, _, _ = generate_test_spectrum()
    butler = _snr(f, p).to_seismology()
    butler.estimate_numax()
    butler.estimate_deltanu()
    with pytest.raises(ValueError, match="effective temperature"):
        butler.estimate_radius()
    mass = butler.estimate_mass(cteff)
    rad = butler.estimate_radius(cteff)
    logg = butler.estimate_logg(cteff)
    assert np.isclose(mass.value, estimate_mass(butler.numax, butler.deltanu, cteff).value)
    assert np.isclose(rad.value, estimate_radius(butler.numax, butler.deltanu, cteff).value)
    assert np.isclose(logg.value, estimate_logg(butler.numax, cteff).value)
    butler.periodogram.meta["TEFF"] = cteff
    assert np.isclose(butler.estimate_radius().value, rad.value)
    for name in ("numax", "deltanu", "mass", "radius", "logg"):
        assert name in repr(butler)


def test_seismology_warns_without_background_correction():
    f, p, _, _ = generate_test_spectrum()
    with pytest.warns(LightkurveWarning, match="background-corrected"):
        Seismology(Periodogram(u.Quantity(f[1:], u.microhertz), u.Quantity(p[1:], None)))


def test_fill_gaps_host_logic():
    """lightcurve.py:1329-1427 (no cadence-number column): a cadence every median step wherever the spacing exceeds
    1.2 steps; original samples untouched, flux_err interpolated, filler ~ N(mean, std).  Without a GPU the CDPP
    estimate is unavailable and the reference's own fallback (nanstd of the flux) is taken."""
    import lightkurve_b200 as lk
    rng = np.random.default_rng(2)
    t = np.arange(400) * 0.02
    keep = np.ones(400, bool)
    keep[100:140] = False
    keep[300] = False
    flux = 1 + 1e-3 * rng.normal(size=400)
    ferr = np.linspace(1e-3, 2e-3, 400)
    lc = lk.LightCurve(time=t[keep], flux=flux[keep], flux_err=ferr[keep], label="x")
    np.random.seed(3)
    filled = lc.fill_gaps()
    assert len(filled) == 400 and filled.meta["LABEL"] == "x"
    np.testing.assert_allclose(filled.time.value, t, atol=1e-9)
    np.testing.assert_array_equal(filled.flux.value[keep], flux[keep])
    np.testing.assert_allclose(filled.flux_err.value, ferr, rtol=1e-9)            # linear in time: interpolation exact
    filler = filled.flux.value[~keep]
    assert abs(filler.mean() - 1) < 1e-3 and 3e-4 < filler.std() < 3e-3
    with pytest.raises(NotImplementedError):
        lc.fill_gaps(method="spline")
    both = lc.append(lc)
    assert len(both) == 2 * len(lc) and both.flux.unit == lc.flux.unit
