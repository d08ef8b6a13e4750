"""TEST INFRASTRUCTURE: oracle-backed stand-ins for the batched engine entry points
(`lightkurve_b200.engine.*`), used by tests/test_shim_on_oracle.py to run the BODIES of the shim-level GPU tests
(tests/test_gpu_shim.py) on a machine without a GPU.  What that checks: the whole Python layer (argument
handling, grids, units, normalisation bookkeeping, result objects) and the expectations written in those tests,
against the fp64 oracle - so that on the GPU box only the kernels can make them fail.  Nothing in the product
imports this module."""
import numpy as np

from oracle import bls as obls, detrend as odet, ls as ols, pg as opg


def _norm(p, n, normalization, scale):
    if normalization == "amplitude":
        return np.sqrt(np.maximum(p, 0)) * np.sqrt(4.0 / n)
    if normalization == "psd":
        return p * scale
    return p


def _scales(norm_scale, B):
    return [None] * B if norm_scale is None else list(np.broadcast_to(norm_scale, (B,)))


def ls_power_ragged(times, fluxes, frequency, normalization="amplitude", norm_scale=None, algo="auto"):
    B = len(times)
    per_lc = isinstance(frequency, (list, tuple))
    sc = _scales(norm_scale, B)
    out = []
    for b in range(B):
        f = np.asarray(frequency[b] if per_lc else frequency, dtype=np.float64)
        y = np.asarray(fluxes[b], dtype=np.float64)
        if len(y) == 0 or np.all(y == y[0]):
            out.append(np.zeros(len(f), np.float32))
            continue
        with np.errstate(all="ignore"):
            p = ols.ls_slow_psd(times[b], y, f)
        out.append(_norm(p, len(y), normalization, sc[b]).astype(np.float32))
    return out if per_lc else np.stack(out)


def ls_power_shared(t, Y, frequency, normalization="amplitude", norm_scale=None, algo="auto", out=None):
    res = ls_power_ragged([t] * len(Y), list(Y), frequency, normalization,
                          None if norm_scale is None else [norm_scale] * len(Y))
    if out is not None:
        out[...] = res
        return out
    return res


def ls_power_chi2(times, fluxes, frequency, nterms=1, normalization="amplitude", norm_scale=None, return_theta=False):
    B = len(times)
    per_lc = isinstance(frequency, (list, tuple))
    sc = _scales(norm_scale, B)
    out, thetas = [], []
    for b in range(B):
        f = np.asarray(frequency[b] if per_lc else frequency, dtype=np.float64)
        t = np.asarray(times[b], dtype=np.float64)
        y = np.asarray(fluxes[b], dtype=np.float64)
        with np.errstate(all="ignore"):
            p = ols.ls_chi2_psd(t, y, f, nterms)
        out.append(_norm(p, len(y), normalization, sc[b]).astype(np.float32))
        if return_theta:
            th = np.empty((len(f), 2 * nterms + 1))
            for k, fk in enumerate(f):
                X = ols.design_matrix(t - t[0], fk, True, nterms)
                th[k] = np.linalg.solve(X.T @ X, X.T @ (y - y.mean()))
            thetas.append(th)
    if not per_lc:
        out = np.stack(out)
        thetas = np.stack(thetas) if return_theta else None
    return (out, thetas) if return_theta else out


def bls_power(times, fluxes, flux_errs, period, duration, oversample=10, objective="likelihood", return_bins=False):
    B = len(times)
    rs = [obls.bls_power_c(times[b], fluxes[b], None if flux_errs is None else flux_errs[b], period, duration,
                           oversample=oversample, objective=objective, return_bins=return_bins) for b in range(B)]
    res = {k: np.stack([r[k] for r in rs]) for k in obls.RESULT_FIELDS}
    res["period"] = np.ascontiguousarray(np.atleast_1d(period), dtype=np.float64)
    if return_bins:
        res["bins"] = np.stack([r["bins"] for r in rs])
    return res


def flatten(times, fluxes, flux_errs=None, masks=None, window_length=101, polyorder=2, break_tolerance=5,
            niters=3, sigma=3):
    outs = ([], [], [])
    for b in range(len(times)):
        r = odet.flatten(times[b], np.asarray(fluxes[b], dtype=np.float64),
                         None if flux_errs is None else flux_errs[b], window_length=window_length,
                         polyorder=polyorder, break_tolerance=break_tolerance, niters=niters, sigma=sigma,
                         mask=None if masks is None else masks[b])
        for o, v in zip(outs, r):
            o.append(np.asarray(v, dtype=np.float64))
    return outs


def regress(X, Y, flux_err=None, cadence_mask=None, prior_mu=None, prior_sigma=None, sigma=5, niters=5,
            return_cov=False):
    X = np.asarray(X, dtype=np.float64)
    Y = np.atleast_2d(np.asarray(Y, dtype=np.float64))
    B, N = Y.shape
    K = X.shape[-1]
    fe = None if flux_err is None else np.broadcast_to(flux_err, Y.shape)
    cm = None if cadence_mask is None else np.broadcast_to(np.asarray(cadence_mask, dtype=bool), Y.shape)
    out = dict(coefficients=np.full((B, K), np.nan), model=np.full((B, N), np.nan),
               outlier_mask=np.zeros((B, N), bool), status=np.zeros(B, np.int32))
    if return_cov:
        out["covariance"] = np.full((B, K, K), np.nan)
    for b in range(B):
        Xb = X[b] if X.ndim == 3 else X
        try:
            r = odet.regress(Xb, Y[b], None if fe is None else fe[b], None if cm is None else cm[b], prior_mu,
                             prior_sigma, sigma=sigma, niters=niters)
        except np.linalg.LinAlgError:
            out["status"][b] = -4
            continue
        out["coefficients"][b], out["model"][b], out["outlier_mask"][b] = r["coefficients"], r["model"], r["outlier_mask"]
        if return_cov:
            use = (np.ones(N, bool) if cm is None else cm[b]) & ~r["outlier_mask"]
            e = np.ones(use.sum()) if fe is None or not np.any(np.isfinite(fe[b])) else fe[b][use]
            A = Xb[use].T @ (Xb[use] / e[:, None] ** 2)
            if prior_sigma is not None:
                A = A + np.diag(1.0 / np.asarray(prior_sigma, dtype=np.float64) ** 2)
            out["covariance"][b] = np.linalg.inv(A)
    return out


def nanmedian_std(arrays):
    with np.errstate(all="ignore"):
        return (np.array([np.nanmedian(a) if len(a) else np.nan for a in arrays]),
                np.array([np.nanstd(a) if len(a) else np.nan for a in arrays]))


def pg_logmedian(frequency, power, filter_width):
    p = np.asarray(power, dtype=np.float64)
    with np.errstate(all="ignore"):
        r = np.stack([opg.smooth_logmedian(frequency, row, filter_width) for row in np.atleast_2d(p)])
    return r[0] if p.ndim == 1 else r


FAKES = dict(ls_power_ragged=ls_power_ragged, ls_power_shared=ls_power_shared, ls_power_chi2=ls_power_chi2,
             bls_power=bls_power, flatten=flatten, regress=regress, nanmedian_std=nanmedian_std,
             pg_logmedian=pg_logmedian, init=lambda device=0: None)
