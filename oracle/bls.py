"""Oracle: Box Least Squares as lightkurve calls astropy (TEST INFRASTRUCTURE ONLY).

Reference call site: /root/reference/src/lightkurve/periodogram.py:1095-1169
    bls = BoxLeastSquares(lc.time, lc.flux, dy)
    period = bls.autoperiod(duration, minimum_period, maximum_period, frequency_factor)
    result = bls.power(period, duration)   # objective="likelihood", method="fast", oversample=10

Restates astropy >= 5.0 ``timeseries/periodograms/bls/core.py`` (autoperiod,
power pre-processing) and ``bls.c`` (numpy loop here, plain C in bls_c.c).
PARITY UNPINNED for values (see oracle/__init__.py).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DEFAULT_DURATIONS = [0.05, 0.10, 0.15, 0.20, 0.25, 0.33]  # periodogram.py:1102

RESULT_FIELDS = ("power", "depth", "depth_err", "duration", "transit_time",
                 "depth_snr", "log_likelihood")


def autoperiod(t, duration, minimum_period=None, maximum_period=None,
               minimum_n_transit=3, frequency_factor=1.0):
    """astropy BoxLeastSquares.autoperiod (core.py)."""
    t = np.asarray(t, dtype=np.float64)
    duration = np.atleast_1d(np.asarray(duration, dtype=np.float64))
    baseline = t.max() - t.min()
    min_duration = duration.min()
    df = frequency_factor * min_duration / baseline ** 2
    if minimum_period is None:
        minimum_period = 2.0 * duration.max()
    if maximum_period is None:
        if minimum_n_transit <= 1:
            raise ValueError("minimum number of transits must be greater than 1")
        maximum_period = baseline / (minimum_n_transit - 1)
    if maximum_period < minimum_period:
        minimum_period, maximum_period = maximum_period, minimum_period
    if minimum_period <= 0.0:
        raise ValueError("minimum period must be positive")
    minimum_frequency = 1.0 / maximum_period
    maximum_frequency = 1.0 / minimum_period
    nf = 1 + int(np.round((maximum_frequency - minimum_frequency) / df))
    return 1.0 / (maximum_frequency - df * np.arange(nf))


def lk_default_period_bounds(t, duration):
    """lightkurve defaults, periodogram.py:1114-1128."""
    t = np.asarray(t, dtype=np.float64)
    dt = np.median(np.diff(t))
    minimum_period = np.max([dt * 4, np.max(duration) + dt])
    maximum_period = (np.max(t) - np.min(t)) / 3.0
    return minimum_period, maximum_period


def _prepare(t, y, dy):
    t = np.ascontiguousarray(t, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if dy is None:
        ivar = np.ones_like(y)
    else:
        ivar = 1.0 / np.ascontiguousarray(np.broadcast_to(dy, y.shape), dtype=np.float64) ** 2
    t_ref = np.min(t)
    return t - t_ref, y - np.median(y), np.ascontiguousarray(ivar), t_ref


def _validate(period, duration):
    period = np.atleast_1d(np.asarray(period, dtype=np.float64))
    duration = np.atleast_1d(np.asarray(duration, dtype=np.float64))
    if period.ndim != 1 or period.size == 0:
        raise ValueError("period must be 1-dimensional")
    if np.min(period) <= np.max(duration):
        raise ValueError("The maximum transit duration must be shorter than the minimum period")
    return np.ascontiguousarray(period), np.ascontiguousarray(duration)


def bin_index(t_rel, min_t, period, bin_duration):
    """The index rule that must be bit-exact on the GPU (bls.c)."""
    return (np.abs(np.fmod(t_rel - min_t, period)) / bin_duration).astype(np.int32) + 1


def bls_power_numpy(t, y, dy, period, duration, oversample=10, objective="likelihood",
                    return_bins=False):
    """Pure-numpy restatement (slow; small cases only)."""
    period, duration = _validate(period, duration)
    trel, yc, ivar, t_ref = _prepare(t, y, dy)
    use_snr = objective == "snr"
    P = len(period)
    out = {k: np.zeros(P) for k in RESULT_FIELDS}
    bins = np.full((P, 2), -1, dtype=np.int32)
    bin_duration = duration.min() / float(oversample)
    min_t = trel.min()
    wy = yc * ivar
    sum_y = 0.0
    sum_ivar = 0.0
    for n in range(len(trel)):  # same left-to-right order as the C loop
        sum_y += wy[n]
        sum_ivar += ivar[n]
    eps = np.finfo(np.float64).eps
    for p in range(P):
        per = period[p]
        n_bins = int(np.ceil(per / bin_duration)) + oversample
        ind = bin_index(trel, min_t, per, bin_duration)
        mean_y = np.zeros(n_bins + 1)
        mean_ivar = np.zeros(n_bins + 1)
        np.add.at(mean_y, ind, wy)
        np.add.at(mean_ivar, ind, ivar)
        lo = n_bins - oversample
        mean_y[lo:lo + oversample] = mean_y[1:oversample + 1]
        mean_ivar[lo:lo + oversample] = mean_ivar[1:oversample + 1]
        mean_y = np.cumsum(mean_y)
        mean_ivar = np.cumsum(mean_ivar)
        best = -np.inf
        for d in duration:
            dur = int(np.floor(d / bin_duration + 0.5)) if d / bin_duration >= 0 else int(np.ceil(d / bin_duration - 0.5))
            n_max = n_bins - dur
            if n_max < 0:
                continue
            y_in = mean_y[dur:dur + n_max + 1] - mean_y[:n_max + 1]
            ivar_in = mean_ivar[dur:dur + n_max + 1] - mean_ivar[:n_max + 1]
            y_out = sum_y - y_in
            ivar_out = sum_ivar - ivar_in
            ok = ~((ivar_in < eps) | (ivar_out < eps))
            with np.errstate(divide="ignore", invalid="ignore"):
                yi = y_in / ivar_in
                yo = y_out / ivar_out
                depth = yo - yi
                depth_err = np.sqrt(1.0 / ivar_in + 1.0 / ivar_out)
                snr = depth / depth_err
                ll = 0.5 * ivar_in * (yo - yi) * (yo - yi)
            obj = snr if use_snr else ll
            ok &= yo >= yi
            if not ok.any():
                continue
            objm = np.where(ok, obj, -np.inf)
            n = int(np.argmax(objm))  # first maximum
            if objm[n] > best:
                best = objm[n]
                out["power"][p] = objm[n]
                out["depth"][p] = depth[n]
                out["depth_err"][p] = depth_err[n]
                out["depth_snr"][p] = snr[n]
                out["log_likelihood"][p] = ll[n]
                out["duration"][p] = dur * bin_duration
                out["transit_time"][p] = np.fmod(n * bin_duration + 0.5 * (dur * bin_duration) + min_t, per)
                bins[p] = (n, dur)
        if best == -np.inf:
            out["power"][p] = -np.inf
    out["transit_time"] = out["transit_time"] + t_ref
    out["period"] = period
    if return_bins:
        out["bins"] = bins
    return out


def objective_at(t, y, dy, period, duration, n, dur, oversample=10, objective="likelihood"):
    """Objective of ONE box (start bin n, dur bins) at ONE period, in the oracle's arithmetic.
    Used by the parity tests to recognise mathematically tied boxes (e.g. a box at phase ~0 and
    its duplicate in the wrap-padded bins), which bls.c itself separates only by rounding noise."""
    trel, yc, ivar, _ = _prepare(t, y, dy)
    duration = np.atleast_1d(np.asarray(duration, dtype=np.float64))
    bd = duration.min() / float(oversample)
    n_bins = int(np.ceil(period / bd)) + oversample
    ind = bin_index(trel, trel.min(), period, bd)
    my = np.zeros(n_bins + 1)
    mi = np.zeros(n_bins + 1)
    np.add.at(my, ind, yc * ivar)
    np.add.at(mi, ind, ivar)
    lo = n_bins - oversample
    my[lo:lo + oversample] = my[1:oversample + 1]
    mi[lo:lo + oversample] = mi[1:oversample + 1]
    my = np.cumsum(my)
    mi = np.cumsum(mi)
    y_in = my[n + dur] - my[n]
    i_in = mi[n + dur] - mi[n]
    y_out = np.sum(yc * ivar) - y_in
    i_out = np.sum(ivar) - i_in
    yi, yo = y_in / i_in, y_out / i_out
    if objective == "snr":
        return (yo - yi) / np.sqrt(1.0 / i_in + 1.0 / i_out)
    return 0.5 * i_in * (yo - yi) ** 2


def _load():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "_build", "liboracle_bls.so")
    if not os.path.exists(path):
        build()
    lib = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    ip = ctypes.POINTER(ctypes.c_int)
    lib.oracle_bls_fast.restype = ctypes.c_int
    lib.oracle_bls_fast.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_int, dp, ctypes.c_int, dp,
                                    ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp, dp, dp, ip]
    lib.oracle_bls_bin_index.restype = None
    lib.oracle_bls_bin_index.argtypes = [ctypes.c_int, dp, ctypes.c_double, ctypes.c_double,
                                         ctypes.c_double, ip]
    _LIB = lib
    return lib


def build():
    """Compile bls_c.c -> oracle/_build/liboracle_bls.so (gcc, OpenMP, strict IEEE)."""
    import subprocess
    out = os.path.join(_HERE, "_build")
    os.makedirs(out, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-std=c99",
                           "-ffp-contract=off", os.path.join(_HERE, "bls_c.c"),
                           "-o", os.path.join(out, "liboracle_bls.so"), "-lm"])


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def bls_power_c(t, y, dy, period, duration, oversample=10, objective="likelihood",
                return_bins=False):
    """C restatement (OpenMP over periods, like astropy's bls.c)."""
    lib = _load()
    period, duration = _validate(period, duration)
    trel, yc, ivar, t_ref = _prepare(t, y, dy)
    P = len(period)
    res = [np.zeros(P) for _ in range(7)]
    bins = np.full((P, 2), -1, dtype=np.int32)
    rc = lib.oracle_bls_fast(len(trel), _dp(trel), _dp(yc), _dp(ivar), P, _dp(period),
                             len(duration), _dp(duration), int(oversample),
                             1 if objective == "snr" else 0,
                             _dp(res[0]), _dp(res[1]), _dp(res[2]), _dp(res[3]), _dp(res[4]),
                             _dp(res[5]), _dp(res[6]),
                             bins.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    if rc == 1:
        raise ValueError("Invalid period grid")
    if rc == 2:
        raise ValueError("The maximum transit duration must be shorter than the minimum period")
    if rc != 0:
        raise RuntimeError("oracle_bls_fast failed: %d" % rc)
    out = dict(power=res[0], depth=res[1], depth_err=res[2], duration=res[3],
               transit_time=res[4] + t_ref, depth_snr=res[5], log_likelihood=res[6], period=period)
    if return_bins:
        out["bins"] = bins
    return out


def bin_index_c(t_rel, min_t, period, bin_duration):
    lib = _load()
    t_rel = np.ascontiguousarray(t_rel, dtype=np.float64)
    out = np.zeros(len(t_rel), dtype=np.int32)
    lib.oracle_bls_bin_index(len(t_rel), _dp(t_rel), float(min_t), float(period),
                             float(bin_duration), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return out


def boxleastsquares(t, y, dy=None, duration=None, period=None, minimum_period=None,
                    maximum_period=None, frequency_factor=10, oversample=10,
                    objective="likelihood", impl="c"):
    """End-to-end numerics of BoxLeastSquaresPeriodogram.from_lightcurve for
    finite inputs (periodogram.py:1095-1169)."""
    t = np.asarray(t, dtype=np.float64)
    if duration is None:
        duration = DEFAULT_DURATIONS
    if dy is not None and not np.isfinite(dy).all():
        dy = None
    if period is None:
        lo, hi = lk_default_period_bounds(t, duration)
        minimum_period = lo if minimum_period is None else minimum_period
        maximum_period = hi if maximum_period is None else maximum_period
        period = autoperiod(t, duration, minimum_period, maximum_period,
                            frequency_factor=frequency_factor)
    f = bls_power_c if impl == "c" else bls_power_numpy
    return f(t, y, dy, period, duration, oversample, objective)
