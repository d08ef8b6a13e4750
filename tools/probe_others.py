"""GPU probe: throughput of the other hot-path kernels on BASELINE-config-shaped inputs
(K1 ragged LS / config 5, K3 BLS / config 3, K4 flatten + K5 regression / config 4), with the CPU
oracle timed on a small sample beside them.  Prints one line per kernel; results go to profiles/."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lightkurve_b200 import engine  # noqa: E402
from oracle import bls as obls, detrend as odet, ls as ols  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0      # fraction of the full config batch
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["c1", "k1", "bls", "flatten", "regress", "pg"]
engine.init(0)
rng = np.random.default_rng(1003)


def timed(fn, reps=2):
    fn()
    engine.profile_enable(True)
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    wall = (time.perf_counter() - t0) / reps
    k = engine.profile_read()
    engine.profile_enable(False)
    return out, wall, float(np.mean(k)) * 1e-3


if "k1" in which:      # config 5 shape: ragged, irregular sampling, common grid of F = 20000 up to 50 / d
    B = max(8, int(16384 / 8 * scale))            # one GPU's share of config 5
    F = 20000
    times, fluxes = [], []
    for _ in range(B):
        grid = 1325 + np.arange(int(27.8 * 720)) / 720.0
        n = min(len(grid), int(round(10 ** rng.uniform(np.log10(2000), np.log10(20000)))))
        keep = np.sort(rng.choice(len(grid), n, replace=False))
        t = grid[keep] + rng.uniform(-20, 20, n) / 86400.0
        times.append(t)
        fluxes.append((1 + 1e-3 * np.sin(2 * np.pi * 3.1 * t) + 3e-4 * rng.normal(size=n)).astype(np.float32))
    freq = np.linspace(50.0 / F, 50.0, F)
    out, wall, ker = timed(lambda: engine.ls_power_ragged(times, fluxes, freq, "amplitude"), reps=1)
    units = F * sum(len(t) for t in times)
    t0 = time.perf_counter(); ols.ls_fast_psd(times[0], fluxes[0].astype(np.float64), freq[0], freq[1] - freq[0], F); cpu = time.perf_counter() - t0
    ref = np.sqrt(ols.ls_slow_psd(times[0], fluxes[0].astype(np.float64), freq[:500])) * np.sqrt(4.0 / len(times[0]))
    err = np.max(np.abs(out[0][:500] - ref) / (1e-5 * ref.max() + 1e-4 * ref))
    print("K1 ragged LS : B=%d F=%d units=%.3e  kernel %.1f ms  %.3e bin*cad/s (wall %.1f ms incl. H2D/D2H)  "
          "cpu fast 1 LC %.3f s -> %.3e equiv units/s  tol-excess %.2f" % (B, F, units, ker * 1e3, units / ker, wall * 1e3,
                                                                      cpu, F * len(times[0]) / cpu, err))

if "bls" in which:     # config 3: 256 TESS LCs x 20000 cadences x 50000 periods x 10 durations
    B = max(2, int(256 * scale))
    N, P = 20000, 50000
    t = 1325 + np.arange(N + 720) / 720.0
    t = np.concatenate([t[: N // 2], t[N // 2 + 720:]])[:N]
    times = [t] * B
    fluxes, errs = [], []
    for b in range(B):
        y = 1 + 5e-4 * rng.normal(size=N)
        per0, dep, dur0 = rng.uniform(1, 8), 10 ** rng.uniform(np.log10(5e-4), -2), rng.uniform(0.05, 0.3)
        if b % 4 != 3:
            y[np.abs((t - t[0] - 0.7 + 0.5 * per0) % per0 - 0.5 * per0) < 0.5 * dur0] -= dep
        fluxes.append(y)
        errs.append(np.full(N, 5e-4))
    duration = np.linspace(0.05, 0.33, 10)
    period = 1.0 / np.linspace(1 / 0.3314, 1 / 9.26, P)
    res, wall, ker = timed(lambda: engine.bls_power(times, fluxes, errs, period, duration), reps=1)
    t0 = time.perf_counter(); ref = obls.bls_power_c(t, fluxes[0], errs[0], period[::50], duration); cpu = time.perf_counter() - t0
    ok = np.allclose(res["power"][0][::50], ref["power"], rtol=1e-9)
    print("K3 BLS       : B=%d N=%d P=%d D=10  kernel %.1f ms  %.3e (LC,period)/s  eff. %.0f GB/s of 24N+56 B  "
          "(wall %.1f ms)  cpu C/OpenMP %.3e periods/s  parity %s" % (B, N, P, ker * 1e3, B * P / ker,
          B * P * (24 * N + 56) / ker / 1e9, wall * 1e3, len(ref["power"]) / cpu, ok))

if "flatten" in which or "regress" in which:      # config 4
    B = max(4, int(4096 * scale))
    N, K = 65000, 151
    keep = np.sort(rng.choice(71500, N, replace=False))
    tt = 131.5 + keep * 0.0204336
    X = np.cumsum(rng.normal(size=(N, K - 1)), axis=0)
    X, _ = np.linalg.qr(X)
    X = np.hstack([X * np.sqrt(N), np.ones((N, 1))])
    W = rng.normal(size=(B, K)) * 1e-3
    slow = np.cumsum(rng.normal(size=N)) * 1e-5
    Y = 1 + W @ X.T + slow[None, :] + 3e-4 * rng.normal(size=(B, N))
    for b in range(B):
        o = rng.choice(N, N // 300, replace=False)
        Y[b, o] += 8 * 3e-4
    FE = 3e-4 * rng.uniform(0.8, 1.2, (B, N))
    if "flatten" in which:
        (flat, fe, tr), wall, ker = timed(lambda: engine.flatten([tt] * B, list(Y), list(FE), None, window_length=401),
                                          reps=1)
        t0 = time.perf_counter(); r = odet.flatten(tt, Y[0], FE[0], window_length=401); cpu = time.perf_counter() - t0
        ok = np.allclose(tr[0], r[2], rtol=1e-9)
        print("K4 flatten   : B=%d N=%d w=401 niters=3  kernel %.1f ms  %.1f LC/s  eff. %.0f GB/s of 63N B  (wall %.1f ms)  "
              "cpu scipy 1 core %.1f LC/s  parity %s" % (B, N, ker * 1e3, B / ker, B * 63 * N / ker / 1e9, wall * 1e3,
                                                       1 / cpu, ok))
    if "regress" in which:
        rr, wall, ker = timed(lambda: engine.regress(X, Y, FE, None, np.zeros(K), np.full(K, np.inf)), reps=1)
        t0 = time.perf_counter(); r = odet.regress(X, Y[0], FE[0], None, np.zeros(K), np.full(K, np.inf)); cpu = time.perf_counter() - t0
        ok = np.allclose(rr["coefficients"][0], r["coefficients"], rtol=1e-7, atol=1e-10) and \
            np.array_equal(rr["outlier_mask"][0], r["outlier_mask"])
        print("K5 regress   : B=%d N=%d K=%d niters=5  Gram kernel %.1f ms (%.1f TFLOP/s fp64 of N*K^2 flop)  "
              "total wall %.1f ms  %.1f LC/s  cpu numpy %.1f LC/s  parity %s" % (
                  B, N, K, ker * 1e3, B * 2.0 * N * K * K / 2 / ker / 1e12, wall * 1e3, B / wall, 1 / cpu, ok))


if "pg" in which:      # the step after config 2: log-median background of B periodograms of F = 1e5 bins
    from oracle import pg as opg
    B = max(4, int(1024 * scale))
    F = 100000
    freq = (np.arange(F) + 1) * (13.6 / F)
    power = rng.chisquare(2, size=(B, F)) * (1 + 3.0 / (1 + freq))
    out, wall, ker = timed(lambda: engine.pg_logmedian(freq, power, 0.01), reps=1)
    t0 = time.perf_counter(); ref = opg.smooth_logmedian(freq, power[0], 0.01); cpu = time.perf_counter() - t0
    ok = np.allclose(out[0], ref, rtol=1e-13, equal_nan=True)
    nwin = len(engine.logmedian_windows(freq, 0.01)[0])
    print("K6 logmedian : B=%d F=%d windows=%d  median kernel %.2f ms  %.1f periodograms/s (wall %.1f ms incl. H2D/D2H)  "
          "cpu numpy %.2f s/periodogram  parity %s" % (B, F, nwin, ker * 1e3, B / ker, wall * 1e3, cpu, ok))


if "c1" in which:      # config 1: one synthetic sinusoid light curve, 1000 cadences, default to_periodogram()
    import lightkurve_b200 as lk
    rng1 = np.random.default_rng(1001)
    t = np.arange(1000.0)
    y = rng1.normal(1, 0.1, 1000) + np.sin(t / t.max() * 20 * np.pi)
    lc = lk.LightCurve(time=t, flux=y / np.median(y))
    pg = lc.to_periodogram()
    t0 = time.perf_counter()
    for _ in range(200):
        pg = lc.to_periodogram()
    dt_api = (time.perf_counter() - t0) / 200
    fr = np.asarray(pg.frequency.value)
    yy = np.asarray(lc.flux.value)
    t0 = time.perf_counter()
    for _ in range(200):
        out = engine.ls_power_ragged([t], [yy], fr, "amplitude")
    dt_eng = (time.perf_counter() - t0) / 200
    t0 = time.perf_counter(); _, pref, _ = ols.lombscargle(t, yy); cpu = time.perf_counter() - t0
    ok = np.nanargmax(pref) == np.nanargmax(pg.power.value)
    print("C1 plumbing  : N=1000 F=%d  LightCurve.to_periodogram() %.0f us per call (engine call alone %.0f us)  "
          "cpu oracle fast %.0f us  period %.2f d  argmax parity %s" % (len(fr), dt_api * 1e6, dt_eng * 1e6, cpu * 1e6,
                                                                   float(pg.period_at_max_power.value), ok))
