"""Parity at BASELINE.json's FULL sizes (config 2, 3, 4 shapes) - the oracle cannot cover these exhaustively in
seconds, so: spot checks of random (light curve, bin/period) entries against the oracle, plus size-independent
properties of the estimators (scaling, batch-permutation invariance, mean-shift invariance, idempotence).
Runs in well under a minute on one B200."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import bls as obls, detrend as odet, ls as ols  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from lightkurve_b200 import engine as eng
    eng.init(0)
    return eng


def test_config2_full_size_lombscargle(engine):
    """1024 light curves x 65 000 cadences x 1e5 frequencies (bench.py workload): the whole tensor-core call."""
    from bench import make_workload
    t, Y, freq = make_workload("c2", 1002)
    B, N = Y.shape
    F = len(freq)
    assert (B, N, F) == (1024, 65000, 100000)
    out = engine.ls_power_shared(t, Y, freq, "amplitude")
    assert out.shape == (B, F) and np.isfinite(out).all()
    rng = np.random.default_rng(7)
    # (a) spot check against the exact fp64 sums: 4 light curves x 150 random bins (+ the lowest and highest bins)
    for b in rng.choice(B, 4, replace=False):
        bins = np.unique(np.concatenate([rng.choice(F, 150, replace=False), [0, 1, 2, F - 1]]))
        ref = np.sqrt(ols.ls_slow_psd(t, Y[b].astype(np.float64), freq[bins])) * np.sqrt(4.0 / N)
        pmax = float(out[b].max())
        tol = 1e-5 * max(pmax, ref.max()) + 1e-4 * ref                     # amplitude spectrum: same form as the psd bound
        assert np.all(np.abs(out[b][bins] - ref) <= 2 * tol), (b, np.max(np.abs(out[b][bins] - ref) / tol))
    # (c) batch-permutation invariance: every light curve's spectrum is independent of its neighbours (bitwise)
    perm = rng.permutation(B)
    out_p = engine.ls_power_shared(t, np.ascontiguousarray(Y[perm]), freq, "amplitude")
    assert np.array_equal(out_p, out[perm])
    # (d) the amplitude spectrum is linear in the flux about its mean: y -> 1 + 4 (y - 1) is EXACT in fp32 for these
    # fluxes (|y - 1| << 1), so every bin must scale by 4 up to the kernel's own tolerance
    sub = rng.choice(B, 256, replace=False)
    Y4 = (np.float32(1.0) + np.float32(4.0) * (Y[sub] - np.float32(1.0))).astype(np.float32)
    assert np.array_equal(Y4.astype(np.float64), 1.0 + 4.0 * (Y[sub].astype(np.float64) - 1.0))
    o1 = engine.ls_power_shared(t, np.ascontiguousarray(Y[sub]), freq[:20000], "amplitude")
    o4 = engine.ls_power_shared(t, Y4, freq[:20000], "amplitude")
    big = o1 > 2e-2 * o1.max(axis=1, keepdims=True)
    np.testing.assert_allclose(o4[big], 4.0 * o1[big], rtol=5e-3)


def test_config3_shape_bls(engine):
    """TESS 2-min shape: 20 000 cadences x 50 000 periods x 10 durations (6 of the 256 light curves)."""
    rng = np.random.default_rng(1003)
    N, P, B = 20000, 50000, 6
    t = 1325 + np.arange(N + 720) / 720.0
    t = np.concatenate([t[: N // 2], t[N // 2 + 720:]])[:N]
    ys, dys = [], []
    for b in range(B):
        y = 1 + 5e-4 * rng.normal(size=N)
        per, dep, dur = rng.uniform(1, 8), 10 ** rng.uniform(-3.3, -2), rng.uniform(0.05, 0.3)
        y[np.abs((t - t[0] - 0.3 * per + 0.5 * per) % per - 0.5 * per) < 0.5 * dur] -= dep
        ys.append(y)
        dys.append(np.full(N, 5e-4))
    duration = np.linspace(0.05, 0.33, 10)
    period = 1.0 / np.linspace(1 / 0.3314, 1 / 9.26, P)
    res = engine.bls_power([t] * B, ys, dys, period, duration, return_bins=True)
    for b in (0, B - 1):
        sel = np.sort(rng.choice(P, 4000, replace=False))
        ref = obls.bls_power_c(t, ys[b], dys[b], period[sel], duration, return_bins=True)
        same = np.all(res["bins"][b][sel] == ref["bins"], axis=1)
        for i in np.flatnonzero(~same):                   # only mathematical ties of bls.c itself may differ
            n_, d_ = res["bins"][b][sel][i]
            o = obls.objective_at(t, ys[b], dys[b], period[sel][i], duration, int(n_), int(d_))
            assert abs(o - ref["power"][i]) <= 1e-10 * abs(ref["power"][i])
        assert (~same).mean() < 0.02
        for k in ("power", "depth", "depth_snr"):
            np.testing.assert_allclose(res[k][b][sel], ref[k], rtol=1e-9, atol=1e-12)
    # mean-shift invariance: BLS subtracts the median itself
    res2 = engine.bls_power([t], [ys[0] + 0.25], [dys[0]], period[::25], duration)
    np.testing.assert_allclose(res2["power"][0], res["power"][0][::25], rtol=1e-6)


def test_config4_shape_flatten_and_regression(engine):
    """Kepler shape: 65 000 cadences, window 401; K = 151 regressors (64 of the 4096 light curves)."""
    rng = np.random.default_rng(1004)
    N, B, K = 65000, 64, 151
    t = 131.5 + np.arange(N) * 0.0204336
    X = np.cumsum(rng.normal(size=(N, K - 1)), axis=0)
    X, _ = np.linalg.qr(X - X.mean(0))                       # orthonormalised random-walk "CBVs" 
    X = np.hstack([X * np.sqrt(N), np.ones((N, 1))])
    W = rng.normal(size=(B, K)) * 1e-3
    Y = 1 + W @ X.T + 3e-4 * rng.normal(size=(B, N))
    out_idx = rng.choice(N, (B, 200))
    for b in range(B):
        Y[b, out_idx[b]] += 8 * 3e-4
    fe = 3e-4 * rng.uniform(0.8, 1.2, (B, N))
    r = engine.regress(X, Y, fe, None, None, None, sigma=5, niters=5)
    assert (r["status"] == 0).all()
    for b in (0, B - 1):
        ref = odet.regress(X, Y[b], fe[b], None, None, None, sigma=5, niters=5)
        assert np.array_equal(r["outlier_mask"][b], ref["outlier_mask"])
        np.testing.assert_allclose(r["coefficients"][b], ref["coefficients"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(r["model"][b], ref["model"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(r["coefficients"][:, :-1], W[:, :-1], atol=2e-5)   # the injected coefficients come back
    # flatten of the corrected light curves, window_length = 401
    corrected = [Y[b] - r["model"][b] + 1e-3 * np.sin(2 * np.pi * (t - t[0]) / 7.0) for b in range(8)]
    flat, flat_err, trend = engine.flatten([t] * 8, corrected, [fe[b] for b in range(8)], None, window_length=401,
                                           polyorder=2, break_tolerance=5, niters=3, sigma=3)
    for b in (0, 7):
        rf, re_, rt = odet.flatten(t, corrected[b], fe[b], window_length=401, polyorder=2, break_tolerance=5, niters=3,
                                   sigma=3)
        np.testing.assert_allclose(trend[b], rt, rtol=1e-9)
        np.testing.assert_allclose(flat[b], rf, rtol=1e-9)
    # flattening a flattened light curve changes it by less than the noise that is left (measured ratio 0.39)
    flat2, _, _ = engine.flatten([t], [flat[0]], [fe[0]], None, window_length=401, polyorder=2, break_tolerance=5,
                                 niters=3, sigma=3)
    assert np.nanstd(flat2[0] - flat[0]) < 0.7 * np.nanstd(flat[0] - 1.0)
