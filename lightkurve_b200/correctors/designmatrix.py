"""`DesignMatrix` / `DesignMatrixCollection`: the parts of
/root/reference/src/lightkurve/correctors/designmatrix.py that RegressionCorrector.correct uses
(X, prior_mu, prior_sigma, validate, append_constant, split, standardize, pca, collect, collection hstack);
plotting, sparse matrices and the spline builders are out of scope.
"""
import warnings
from copy import deepcopy

import numpy as np
import pandas as pd

from ..utils import LightkurveWarning

__all__ = ["DesignMatrix", "DesignMatrixCollection"]


class DesignMatrix:
    """A matrix of column vectors for use in linear regression (designmatrix.py:28-384)."""

    def __init__(self, df, columns=None, name="unnamed_matrix", prior_mu=None, prior_sigma=None):
        if not isinstance(df, pd.DataFrame):
            df = pd.DataFrame(df)
        self.df = df
        if columns is not None:
            df.columns = columns
        self.columns = list(df.columns)
        self.name = name
        prior_mu = getattr(prior_mu, "value", prior_mu)
        if prior_mu is None:
            prior_mu = np.zeros(len(df.T))
        self.prior_mu = np.atleast_1d(prior_mu)
        prior_sigma = getattr(prior_sigma, "value", prior_sigma)
        if prior_sigma is None:
            prior_sigma = np.ones(len(df.T)) * np.inf
        self.prior_sigma = np.atleast_1d(prior_sigma)

    @property
    def X(self):
        """Design matrix "X" to be used in RegressionCorrector objects."""
        return self.df.values

    @property
    def values(self):
        return self.df.values

    @property
    def shape(self):
        return self.X.shape

    @property
    def rank(self):
        """Matrix rank computed using numpy.linalg.matrix_rank (setup-time, host)."""
        return np.linalg.matrix_rank(self.values)

    def copy(self):
        return deepcopy(self)

    def split(self, row_indices, inplace=False):
        """Split each regressor into several columns at ``row_indices`` (designmatrix.py:160-215)."""
        if isinstance(row_indices, int):
            row_indices = [row_indices]
        if (len(row_indices) == 0) or (row_indices == [0]) or (row_indices is None):
            return self
        dm = self if inplace else self.copy()
        x = np.arange(len(dm.df))
        dfs = []
        boundaries = np.append(np.append(0, row_indices), len(dm.df))
        for idx, a, b in zip(range(len(boundaries) - 1), boundaries[:-1], boundaries[1:]):
            new_columns = dict(("{}".format(val), "{}".format(val) + " {}".format(idx + 1))
                               for val in list(dm.df.columns))
            dfs.append(dm.df.copy().rename(columns=new_columns))
            dfs[-1].loc[~np.isin(x, np.arange(a, b)), :] = 0
        dm.df = pd.concat(dfs, axis=1)
        dm.columns = list(dm.df.columns)
        dm.prior_mu = np.hstack([dm.prior_mu for idx in range(len(dfs))])
        dm.prior_sigma = np.hstack([dm.prior_sigma for idx in range(len(dfs))])
        return dm

    def standardize(self, inplace=False):
        """Median-subtract and sigma-divide every non-constant column (designmatrix.py:216-250); zeros are
        treated as missing and stay zero."""
        ar = np.asarray(np.copy(self.df), dtype=float)
        ar[ar == 0] = np.nan
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            is_const = np.nanstd(ar, axis=0) == 0
            median = np.atleast_2d(np.nanmedian(ar, axis=0)[~is_const])
            std = np.atleast_2d(np.nanstd(ar, axis=0)[~is_const])
            ar[:, ~is_const] = (ar[:, ~is_const] - median) / std
        new_df = pd.DataFrame(ar, columns=self.columns).fillna(0)
        dm = self if inplace else self.copy()
        dm.df = new_df
        return dm

    def pca(self, nterms=6, n_iter=10):
        """Principal components of the (column-centred) matrix as a new DesignMatrix (designmatrix.py:252-282).
        The reference uses the randomised `fbpca.pca`; this is the exact thin SVD (`n_iter` is accepted and ignored),
        so the components agree up to sign and the randomised solver's error."""
        if nterms > self.shape[1]:
            nterms = self.shape[1]
        a = np.asarray(self.values, dtype=float)
        u_, _, _ = np.linalg.svd(a - a.mean(axis=0), full_matrices=False)
        return DesignMatrix(u_[:, :nterms], name=self.name)

    def append_constant(self, prior_mu=0, prior_sigma=np.inf, inplace=False):
        """Append a column of ones named "offset" (designmatrix.py:284-304)."""
        dm = self if inplace else self.copy()
        extra_df = pd.DataFrame(np.atleast_2d(np.ones(self.shape[0])).T, columns=["offset"])
        dm.df = pd.concat([self.df, extra_df], axis=1)
        dm.columns = list(dm.df.columns)
        dm.prior_mu = np.append(self.prior_mu, prior_mu)
        dm.prior_sigma = np.append(self.prior_sigma, prior_sigma)
        return dm

    def _validate(self, rank=True):
        if rank:
            if self.rank < (0.5 * self.shape[1]):
                warnings.warn(
                    "The design matrix has low rank ({}) compared to the "
                    "number of columns ({}), which suggests that the "
                    "matrix contains duplicate or correlated columns. "
                    "This may prevent the regression from succeeding. "
                    "Consider reducing the dimensionality by calling the "
                    "`pca()` method.".format(self.rank, self.shape[1]),
                    LightkurveWarning,
                )
        if self.prior_mu is not None:
            if len(self.prior_mu) != self.shape[1]:
                raise ValueError("`prior_mu` must have shape {}" "".format(self.shape[1]))
        if self.prior_sigma is not None:
            if len(self.prior_sigma) != self.shape[1]:
                raise ValueError("`prior_sigma` must have shape {}" "".format(self.shape[1]))
            if np.any(np.asarray(self.prior_sigma) <= 0):
                raise ValueError("`prior_sigma` values cannot be smaller than " "or equal to zero")

    def validate(self, rank=True):
        """Emits LightkurveWarning if the matrix has low rank; checks prior shapes (designmatrix.py:306-349)."""
        self._validate()

    def __getitem__(self, key):
        return self.df[key].values

    def collect(self, matrix):
        """Join two design matrices into a collection (designmatrix.py:382-384)."""
        return DesignMatrixCollection([self, matrix])

    def __repr__(self):
        return "{} DesignMatrix {}".format(self.name, self.shape)


class DesignMatrixCollection:
    """Object which stores multiple design matrices (designmatrix.py:387-553)."""

    def __init__(self, matrices):
        self.matrices = matrices
        self.X = np.hstack(tuple(m.X for m in self.matrices))
        self._child_class = DesignMatrix
        self.validate()

    @property
    def values(self):
        return np.hstack(tuple(m.values for m in self.matrices))

    @property
    def prior_mu(self):
        return np.hstack([m.prior_mu for m in self])

    @property
    def prior_sigma(self):
        return np.hstack([m.prior_sigma for m in self])

    @property
    def columns(self):
        return np.hstack([d.columns for d in self])

    def split(self, row_indices):
        return self.__class__([d.split(row_indices) for d in self])

    def standardize(self):
        return self.__class__([d.standardize() for d in self])

    def __getitem__(self, key):
        try:
            return self.matrices[key]
        except Exception:
            arg = np.argwhere([m.name == key for m in self.matrices])
            return self.matrices[arg[0][0]]

    def __iter__(self):
        return iter(self.matrices)

    def __len__(self):
        return len(self.matrices)

    def validate(self):
        [d.validate() for d in self]

    def __repr__(self):
        return "DesignMatrixCollection:\n" + "".join(["\t{}\n".format(i.__repr__()) for i in self])

    def to_designmatrix(self, name=None):
        if name is None:
            name = self.matrices[0].name
        return self._child_class(self.X, columns=self.columns, prior_mu=self.prior_mu,
                                 prior_sigma=self.prior_sigma, name=name)
