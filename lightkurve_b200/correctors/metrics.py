"""Over-fitting metric of a systematics correction: /root/reference/src/lightkurve/correctors/metrics.py:23-123
(``overfit_metric_lombscargle``).  Same definition and the same use of numpy's
global random stream; the difference is where the arithmetic runs: the original / corrected periodograms are
computed once (the reference recomputes the identical pair in every iteration) and the ``n_samples`` white-noise
periodograms, which share one cadence grid, go through the batched GPU call in one launch.
``underfit_metric_neighbors`` needs MAST neighbour searches and is out of scope.
"""
import numpy as np

from ..collections import LightCurveCollection
from ..lightcurve import LightCurve

__all__ = ["overfit_metric_lombscargle"]


def overfit_metric_lombscargle(original_lc, corrected_lc, n_samples=10):
    """Change in broad-band Lomb-Scargle power introduced by a correction, mapped to [0, 1] (0 bad, 1 good);
    0.5 means the introduced noise has the power level of the light curve's uncertainties."""
    orig_lc = original_lc.copy()
    orig_lc = orig_lc.remove_nans().normalize()
    orig_lc -= 1.0
    corrected_lc = corrected_lc.copy()
    corrected_lc = corrected_lc.remove_nans().normalize()
    corrected_lc -= 1.0
    if len(corrected_lc) == 0:
        return 1.0

    pg_orig = orig_lc.to_periodogram()
    pg_corrected = corrected_lc.to_periodogram(frequency=pg_orig.frequency)
    pg_change = np.asarray(pg_corrected.power.value) - np.asarray(pg_orig.power.value)
    pg_change = pg_change[~np.isnan(pg_change)]
    n_positive = len(np.nonzero(pg_change > 0.0)[0])

    n_cad = len(orig_lc)
    mean_err = np.nanmean(np.asarray(corrected_lc.flux_err.value))
    # the reference draws randn(n, 1) once per iteration, in order: same stream
    noise = [(np.random.randn(n_cad, 1) * mean_err).T[0] for _ in range(n_samples)]
    noise_lcs = LightCurveCollection([LightCurve(time=orig_lc.time, flux=w, flux_err=np.zeros(n_cad)) for w in noise])
    noise_pgs = noise_lcs.to_periodogram() if n_samples > 0 else []

    metric_per_iter = []
    for pg_noise in noise_pgs:
        mean_noise_power = np.nanmean(np.asarray(pg_noise.power.value))
        if n_positive == 0:
            metric_per_iter.append(0.0)
        else:
            denominator = n_positive * mean_noise_power
            metric_per_iter.append(np.inf if denominator == 0 else np.sum(pg_change[pg_change > 0.0]) / denominator)
    with np.errstate(over="ignore"):
        metric = np.mean(metric_per_iter)
        return float(2.0 / (1 + np.exp(np.max([metric, 0.0]))))
