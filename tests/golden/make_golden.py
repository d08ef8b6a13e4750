"""Writes the golden fixtures under tests/golden/ from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  astropy/lightkurve cannot be imported in the build
container, so these are ORACLE outputs frozen in time (they pin the oracle against later edits and
give the GPU tests fixed inputs/outputs); scipy/numpy - the reference's real arithmetic for
flatten / regression - are called directly.  Record of versions is stored in each file."""
import os
import sys

import numpy as np
import scipy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import bls as obls, detrend as odet, ls as ols  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
VERS = "numpy %s scipy %s" % (np.__version__, scipy.__version__)


def main():
    # BASELINE config 1: 1000-cadence sinusoid, default lightkurve grid (F = 2497)
    rng = np.random.default_rng(1001)
    t = np.arange(1000.0)
    y = 1 + rng.normal(0, 0.1, 1000) + np.sin(20 * np.pi * t / 999)
    y /= np.median(y)
    f, p_slow, _ = ols.lombscargle(t, y, ls_method="slow")
    _, p_fast, _ = ols.lombscargle(t, y, ls_method="fast")
    np.savez_compressed(os.path.join(HERE, "ls_c1.npz"), t=t, y=y, frequency=f, power_slow=p_slow,
                        power_fast=p_fast, versions=VERS)

    rng = np.random.default_rng(1003)
    tt = 1325 + np.arange(3000) / 720.0 * 6
    tt = tt[(tt < 1335) | (tt > 1336)]
    fl = 1 + 5e-4 * rng.normal(size=len(tt))
    fl[np.abs((tt - 1325.7 + 1.6) % 3.2 - 1.6) < 0.06] -= 4e-3
    dy = np.full(len(tt), 5e-4)
    dur = np.linspace(0.05, 0.33, 10)
    per = obls.autoperiod(tt, dur, 0.4, 8.0, frequency_factor=40)
    r = obls.bls_power_c(tt, fl, dy, per, dur, return_bins=True)
    np.savez_compressed(os.path.join(HERE, "bls_small.npz"), t=tt, y=fl, dy=dy, period=per, durations=dur,
                        bins=r["bins"], versions=VERS, **{k: r[k] for k in obls.RESULT_FIELDS})

    rng = np.random.default_rng(1004)
    tf = 131.5 + np.arange(4000) * 0.0204336
    tf = np.delete(tf, np.r_[700:760, 2500:2512])
    ff = 1 + 0.01 * np.sin(tf / 2.0) + 3e-4 * rng.normal(size=len(tf))
    ff[rng.choice(len(tf), 12, replace=False)] += 3e-3
    _, _, trend = odet.flatten(tf, ff, window_length=401, niters=3)
    np.savez_compressed(os.path.join(HERE, "flatten_small.npz"), t=tf, f=ff, window_length=401, trend=trend,
                        versions=VERS)

    rng = np.random.default_rng(1005)
    N, K = 1500, 24
    X = np.cumsum(rng.normal(size=(N, K - 1)), axis=0)
    X, _ = np.linalg.qr(X)
    X = np.hstack([X * np.sqrt(N), np.ones((N, 1))])
    yv = 1 + X @ (1e-3 * rng.normal(size=K)) + 3e-4 * rng.normal(size=N)
    yv[rng.choice(N, 5, replace=False)] += 3e-3
    fe = 3e-4 * rng.uniform(0.8, 1.2, N)
    rr = odet.regress(X, yv, fe, sigma=5, niters=5)
    np.savez_compressed(os.path.join(HERE, "regress_small.npz"), X=X, y=yv, fe=fe, coefficients=rr["coefficients"],
                        outlier_mask=rr["outlier_mask"], model=rr["model"], versions=VERS)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
