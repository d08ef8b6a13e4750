"""`RegressionCorrector`: /root/reference/src/lightkurve/correctors/regressioncorrector.py.

``__init__`` validation (:88-117) and the result bookkeeping of ``correct`` (:300-342) are mirrored
on the host; the numerics - the ``niters`` x {weighted normal equations with Gaussian priors, LU
solve, model, sigma_clip} loop of :244-279 - are ONE C-ABI call (``lkb_regress``) into CUDA
kernels.  ``RegressionCorrector.correct_batch`` is the new collection-level entry (many light
curves sharing one design matrix, BASELINE config 4) with the same per-light-curve semantics.
"""
import logging
import warnings

import numpy as np

from ..lightcurve import LightCurve
from ..units import Quantity
from .designmatrix import (DesignMatrix, DesignMatrixCollection, SparseDesignMatrix,
                           SparseDesignMatrixCollection)

log = logging.getLogger(__name__)

__all__ = ["RegressionCorrector"]


class RegressionCorrector:
    """Remove noise using linear regression against a DesignMatrix."""

    def __init__(self, lc):
        fe = np.asarray(lc.flux_err.value, dtype=np.float64)
        if np.any([~np.isfinite(np.asarray(lc.time.value)), ~np.isfinite(np.asarray(lc.flux.value))]):
            raise ValueError(
                "Input light curve has NaNs in time or flux. "
                "Please remove NaNs before correction "
                "(e.g. using `lc = lc.remove_nans()`)."
            )
        if np.any(~np.isfinite(fe)) and not np.all(~np.isfinite(fe)):
            raise ValueError(
                "Input light curve has NaNs in `flux_err`. "
                "Please remove NaNs before correction "
                "(e.g. using `lc = lc.remove_nans()`)."
            )
        if np.any(fe[np.isfinite(fe)] <= 0):
            raise ValueError(
                "Input light curve contains flux uncertainties "
                "smaller than or equal to zero. Please remove "
                "these (e.g. using `lc = lc[lc.flux_err > 0]`)."
            )
        self.lc = lc
        self.design_matrix_collection = None
        self.coefficients = None
        self.corrected_lc = None
        self.model_lc = None
        self.diagnostic_lightcurves = None

    def __repr__(self):
        return "RegressionCorrector (ID: {})".format(self.lc.targetid)

    @property
    def dmc(self):
        """Shorthand for self.design_matrix_collection."""
        return self.design_matrix_collection

    @property
    def original_lc(self):
        return self.lc

    @staticmethod
    def _dense_X(dmc):
        """The collection's matrix as the C-contiguous float64 array ``lkb_regress`` takes.  The GPU Gram
        kernel is dense, so a sparse collection is densified here, once per call."""
        X = dmc.X
        if hasattr(X, "toarray"):
            X = X.toarray()
        return np.ascontiguousarray(X, dtype=np.float64)

    @staticmethod
    def _as_collection(design_matrix_collection):
        if not isinstance(design_matrix_collection, DesignMatrixCollection):
            if isinstance(design_matrix_collection, SparseDesignMatrix):        # regressioncorrector.py:225-232
                design_matrix_collection = SparseDesignMatrixCollection([design_matrix_collection])
            elif isinstance(design_matrix_collection, DesignMatrix):
                design_matrix_collection = DesignMatrixCollection([design_matrix_collection])
            else:
                raise TypeError("design_matrix_collection must be a DesignMatrix or DesignMatrixCollection")
        return design_matrix_collection

    def correct(self, design_matrix_collection, cadence_mask=None, sigma=5, niters=5, propagate_errors=False):
        """Find the best fit correction for the light curve (regressioncorrector.py:191-309)."""
        from .. import engine
        dmc = self._as_collection(design_matrix_collection)
        dmc.validate()
        self.design_matrix_collection = dmc
        n = len(self.lc.time)
        if cadence_mask is None:
            self.cadence_mask = np.ones(n, bool)
        else:
            self.cadence_mask = np.asarray(cadence_mask, dtype=bool)
        X = self._dense_X(dmc)
        if X.shape[0] != n:
            raise ValueError("design matrix has {} rows but the light curve has {} cadences".format(X.shape[0], n))
        fe = np.asarray(self.lc.flux_err.value, dtype=np.float64)
        fe_arg = None if np.all(~np.isfinite(fe)) else fe[None, :]          # :157-160
        res = engine.regress(X, np.asarray(self.lc.flux.value, dtype=np.float64)[None, :], fe_arg,
                             self.cadence_mask[None, :], np.asarray(dmc.prior_mu, dtype=np.float64),
                             np.asarray(dmc.prior_sigma, dtype=np.float64), sigma=sigma, niters=niters,
                             return_cov=bool(propagate_errors))
        if res["status"][0] != 0:
            raise np.linalg.LinAlgError("Singular matrix")
        self.outlier_mask = res["outlier_mask"][0]
        self.coefficients = res["coefficients"][0]
        model_err = None
        if propagate_errors:
            # (X^T W X + prior)^-1 comes from the GPU (np.linalg.inv at :185); the 100 multivariate-normal
            # draws use numpy's global RNG exactly like the reference (:280-297) so that a seeded run draws
            # the same stream - RNG bookkeeping, not hot-path arithmetic.
            self.coefficients_err = res["covariance"][0]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                samples = np.asarray([X.dot(np.random.multivariate_normal(self.coefficients, self.coefficients_err))
                                      for idx in range(100)]).T
            model_err = np.abs(np.percentile(samples, [16, 84], axis=1)
                               - np.median(samples, axis=1)[:, None].T).mean(axis=0)
        else:
            self.coefficients_err = np.zeros(len(self.coefficients)) * np.nan
        self._finish(res["model"][0], model_err)
        return self.corrected_lc

    def _finish(self, model_flux, model_err=None):
        unit = self.lc.flux.unit
        if model_err is None:
            model_err = np.zeros(len(model_flux))
        self.model_lc = LightCurve(time=self.lc.time, flux=Quantity(model_flux, unit),
                                   flux_err=Quantity(model_err, unit))
        self.corrected_lc = self.lc.copy()
        self.corrected_lc.flux = self.lc.flux - self.model_lc.flux
        self.corrected_lc.flux_err = (self.lc.flux_err ** 2 + self.model_lc.flux_err ** 2) ** 0.5
        self.diagnostic_lightcurves = self._create_diagnostic_lightcurves()

    def _create_diagnostic_lightcurves(self):
        """Model light curve of every sub design matrix (regressioncorrector.py:311-342)."""
        if self.coefficients is None:
            raise ValueError("you need to call `correct()` first")
        lcs = {}
        idx = 0
        for submatrix in self.dmc:
            k = submatrix.shape[1]
            firstcol, lastcol = idx, idx + k
            idx = lastcol
            model_flux = np.asarray(submatrix.X.dot(self.coefficients[firstcol:lastcol]), dtype=np.float64)
            lcs[submatrix.name] = LightCurve(time=self.lc.time, flux=Quantity(model_flux, self.lc.flux.unit),
                                             flux_err=Quantity(np.zeros(len(model_flux)), self.lc.flux.unit),
                                             label=submatrix.name)
        return lcs

    # ---- collection-level entry (new API; == per-LC loop) -----------------------------------
    @staticmethod
    def correct_batch(lightcurves, design_matrix_collection, cadence_mask=None, sigma=5, niters=5):
        """Correct many light curves that share one design matrix in a single device call.
        Returns the list of RegressionCorrector objects (each with corrected_lc, model_lc,
        coefficients, outlier_mask set) - identical to calling ``correct`` on each."""
        from .. import engine
        correctors = [RegressionCorrector(lc) for lc in lightcurves]
        dmc = RegressionCorrector._as_collection(design_matrix_collection)
        dmc.validate()
        X = RegressionCorrector._dense_X(dmc)
        n = X.shape[0]
        Y = np.stack([np.asarray(lc.flux.value, dtype=np.float64) for lc in lightcurves])
        if Y.shape[1] != n:
            raise ValueError("all light curves must have as many cadences as the design matrix has rows")
        FE = np.stack([np.asarray(lc.flux_err.value, dtype=np.float64) for lc in lightcurves])
        allnan = np.all(~np.isfinite(FE), axis=1)
        if allnan.all():
            fe_arg = None
        else:
            fe_arg = np.where(allnan[:, None], 1.0, FE)
        cm = None if cadence_mask is None else np.broadcast_to(np.asarray(cadence_mask, dtype=bool), Y.shape)
        res = engine.regress(X, Y, fe_arg, cm, np.asarray(dmc.prior_mu, dtype=np.float64),
                             np.asarray(dmc.prior_sigma, dtype=np.float64), sigma=sigma, niters=niters)
        for b, rc in enumerate(correctors):
            if res["status"][b] != 0:
                raise np.linalg.LinAlgError("Singular matrix (light curve %d)" % b)
            rc.design_matrix_collection = dmc
            rc.cadence_mask = np.ones(n, bool) if cm is None else np.asarray(cm[b])
            rc.outlier_mask = res["outlier_mask"][b]
            rc.coefficients = res["coefficients"][b]
            rc.coefficients_err = np.zeros(X.shape[1]) * np.nan
            rc._finish(res["model"][b])
        return correctors
