"""CPU oracle for the lightkurve periodogram-and-detrending hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  ``lightkurve_b200`` never imports it: the
product path is CUDA-only and fails loudly when the extension is missing.

What it restates (fp64, numpy/scipy + one plain-C file):

* ``ls``       - astropy ``LombScargle(...).power`` as lightkurve calls it
                 (``/root/reference/src/lightkurve/periodogram.py:961-975``):
                 the exact ``slow`` sums and the default ``fast``
                 (extirpolation + FFT) approximation, plus lightkurve's own
                 amplitude/psd rescale.
* ``bls``      - astropy ``BoxLeastSquares.autoperiod/.power(method="fast")``
                 as called at ``periodogram.py:1161-1169`` (numpy and C
                 restatements of ``bls.c``; C one is OpenMP over periods like
                 the original).
* ``detrend``  - ``LightCurve.flatten`` (``lightcurve.py:996-1070``) on top of
                 the REAL ``scipy.signal.savgol_filter`` / ``interp1d``;
                 astropy ``sigma_clip`` defaults; ``RegressionCorrector``
                 ``_fit_coefficients``/``correct``
                 (``correctors/regressioncorrector.py:127-189,244-279``) on top
                 of the REAL ``numpy.linalg.solve``.

PARITY STATUS
-------------
astropy is not installable in the build container and is not vendored in
/root/reference, so the LS and BLS restatements follow astropy's published
algorithms from memory of upstream (``lombscargle/implementations/{slow,fast}_impl.py``,
``bls/bls.c``) and are anchored on the reference's own behavioural tests
(``tests/test_periodogram.py``) which ``tests/test_oracle_*.py`` port.
The reference holds NO golden power array for LS or BLS, so for
LS/BLS POWER VALUES the header must say: **parity unpinned** beyond (a) those
behavioural pins, (b) an independent cross-check of the LS math against
``scipy.signal.lombscargle(floating_mean=True)`` (a third-party implementation
of the same Zechmeister & Kuerster estimator present in this image), and (c)
agreement between the independent numpy and C BLS restatements.
flatten / regression are pinned: they call the very scipy/numpy primitives the
reference calls and reproduce the reference's known-answer tests.
"""
