"""Oracle: periodogram post-processing (TEST INFRASTRUCTURE ONLY).

Restates /root/reference/src/lightkurve/periodogram.py:260-284 (Periodogram.smooth, method="logmedian")
verbatim in numpy.  Pinned by the reference's own properties (tests/test_periodogram.py:177-248: mean of the
smoothed white-noise spectrum within 5 % of the mean power; flattened white noise has mean ~1).
"""
import numpy as np


def smooth_logmedian(frequency, power, filter_width=0.1):
    frequency = np.asarray(frequency, dtype=np.float64)
    power = np.asarray(power, dtype=np.float64)
    count = np.zeros(len(frequency), dtype=int)
    bkg = np.zeros_like(frequency)
    x0 = np.log10(frequency[0])
    corr_factor = (8.0 / 9.0) ** 3
    while x0 < np.log10(frequency[-1]):
        m = np.abs(np.log10(frequency) - x0) < filter_width
        if len(bkg[m] > 0):
            with np.errstate(all="ignore"):
                bkg[m] += np.nanmedian(power[m]) / corr_factor
            count[m] += 1
        x0 += 0.5 * filter_width
    with np.errstate(all="ignore"):
        bkg /= count
    return bkg
