"""`DesignMatrix` / `DesignMatrixCollection` and their `scipy.sparse` twins: the parts of
/root/reference/src/lightkurve/correctors/designmatrix.py that feed RegressionCorrector.correct
(X, prior_mu, prior_sigma, validate, append_constant, split, standardize, pca, collect, collection hstack,
to_sparse / to_dense, the B-spline matrix builders).  Plotting is out of scope.

All of this is O(N K) set-up on the host, as in the reference.  The regression itself runs on the GPU
with a DENSE design matrix (FP64 tensor-core Gram kernel): a `SparseDesignMatrix` keeps
the reference's storage and semantics on the host and is densified once when it is handed to
``lkb_regress``.
"""
import warnings
from copy import deepcopy

import numpy as np
import pandas as pd
from scipy import sparse

from ..utils import LightkurveWarning

__all__ = ["DesignMatrix", "DesignMatrixCollection", "SparseDesignMatrix", "SparseDesignMatrixCollection",
           "create_spline_matrix", "create_sparse_spline_matrix"]


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

    def to_sparse(self):
        """The same matrix as a `SparseDesignMatrix` (CSR storage; designmatrix.py:367-380)."""
        return SparseDesignMatrix(sparse.csr_matrix(self.values), name=self.name, columns=self.columns,
                                  prior_mu=self.prior_mu, prior_sigma=self.prior_sigma)

    def collect(self, matrix):
        """Join two design matrices into a collection (designmatrix.py:382-384)."""
        return DesignMatrixCollection([self, matrix])

    def __repr__(self):
        return "{} DesignMatrix {}".format(self.name, self.shape)


class DesignMatrixCollection:
    """Object which stores multiple design matrices (designmatrix.py:387-553)."""

    def __init__(self, matrices):
        if np.any([sparse.issparse(m.X) for m in matrices]):        # designmatrix.py:408-425
            warnings.warn("Some matrices are `SparseDesignMatrix` objects. "
                          "Sparse matrices will be converted to dense matrices.", LightkurveWarning)
            matrices = [m.copy().to_dense() if isinstance(m, SparseDesignMatrix) else m for m in matrices]
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


class SparseDesignMatrix(DesignMatrix):
    """`DesignMatrix` stored as a `scipy.sparse` matrix (designmatrix.py:556-791)."""

    def __init__(self, X, columns=None, name="unnamed_matrix", prior_mu=None, prior_sigma=None):
        if not sparse.issparse(X):
            raise ValueError("Must pass a `scipy.sparse` matrix (e.g. `scipy.sparse.csr_matrix`)")
        ncol = X.shape[1]
        self.columns = np.arange(ncol) if columns is None else columns
        self.name = name
        self._X = X
        self.prior_mu = np.zeros(ncol) if prior_mu is None else prior_mu
        self.prior_sigma = np.ones(ncol) * np.inf if prior_sigma is None else prior_sigma
        self._child_class = SparseDesignMatrix
        self.validate()

    @property
    def X(self):
        return self._X

    @property
    def values(self):
        """Dense 2-D copy of the matrix."""
        return self._X.toarray()

    def validate(self, rank=False):
        """The rank check needs the dense matrix, so it is off unless asked for (designmatrix.py:615-619)."""
        self._validate(rank=rank)

    def split(self, row_indices, inplace=False):
        """One copy of every column per row block, zero outside its block; blocks end at `row_indices`
        (0 and n_rows are ignored); columns that come out empty are dropped (designmatrix.py:621-681)."""
        if not hasattr(row_indices, "__iter__"):
            row_indices = [row_indices]
        nrows = self.shape[0]
        cuts = [int(r) for r in row_indices if r != 0 and r != nrows]
        if len(cuts) == 0:
            return self
        dm = self if inplace else self.copy()
        src = sparse.csr_matrix(dm._X)
        rows = np.arange(nrows)
        blocks = []
        for block_rows in np.array_split(rows, cuts):
            keep = sparse.diags(np.isin(rows, block_rows).astype(src.dtype))
            blocks.append(keep @ src)
        stacked = sparse.hstack(blocks, format="csr")
        nonempty = np.asarray(stacked.sum(axis=0)).ravel() != 0
        nblocks = len(cuts) + 1
        if dm.columns is not None:
            dm.columns = ["{}_{}".format(c, b) for b in range(nblocks) for c in dm.columns]
        dm._X = stacked[:, nonempty].tolil()
        dm.prior_mu = np.tile(np.asarray(self.prior_mu), nblocks)[nonempty]
        dm.prior_sigma = np.tile(np.asarray(self.prior_sigma), nblocks)[nonempty]
        return dm

    def standardize(self, inplace=False):
        """z-scores of the STORED (non-zero) entries of every column (sample standard deviation, ddof = 1);
        zeros stay zero and constant columns are left alone (designmatrix.py:683-726)."""
        dm = self if inplace else self.copy()
        csc = sparse.csc_matrix(dm._X, dtype=np.float64)
        csc.eliminate_zeros()
        out = csc.copy()
        with np.errstate(all="ignore"):
            for j in range(csc.shape[1]):
                lo, hi = csc.indptr[j], csc.indptr[j + 1]
                v = csc.data[lo:hi]
                if len(v) == 0:
                    continue
                mean = v.sum() / len(v)
                std = (np.sum((v - mean) ** 2) * (1.0 / (len(v) - 1))) ** 0.5 if len(v) > 1 else np.nan
                if std == 0:
                    mean, std = 0.0, 1.0
                out.data[lo:hi] = (v - mean) * (1.0 / std)
        dm._X = out.tocsr()
        return dm

    def pca(self, nterms=6, **kwargs):
        return super().pca(nterms, **kwargs).to_sparse()

    def append_constant(self, prior_mu=0, prior_sigma=np.inf, inplace=False):
        """Append a column of ones (designmatrix.py:746-762; the reference does not extend `columns` either)."""
        dm = self if inplace else self.copy()
        ones = sparse.csr_matrix(np.ones((dm.shape[0], 1)))
        dm._X = sparse.hstack([dm._X, ones], format="lil")
        dm.prior_mu = np.append(dm.prior_mu, prior_mu)
        dm.prior_sigma = np.append(dm.prior_sigma, prior_sigma)
        return dm

    def __getitem__(self, key):
        loc = np.where(np.asarray(self.columns) == key)[0]
        if len(loc) == 0:
            raise ValueError("No such column as `{}`.".format(key))
        return sparse.csr_matrix(self._X)[:, loc].toarray()

    def __repr__(self):
        return "{} SparseDesignMatrix {}".format(self.name, self.shape)

    def collect(self, matrix):
        return SparseDesignMatrixCollection([self, matrix])

    def to_sparse(self):
        return self

    def to_dense(self):
        """The same matrix as a dense `DesignMatrix` (designmatrix.py:777-791)."""
        return DesignMatrix(self.values, name=self.name, columns=self.columns, prior_mu=self.prior_mu,
                            prior_sigma=self.prior_sigma)


class SparseDesignMatrixCollection(DesignMatrixCollection):
    """A set of sparse design matrices, stacked column-wise as CSR (designmatrix.py:793-848)."""

    def __init__(self, matrices):
        if not np.all([sparse.issparse(m.X) for m in matrices]):
            warnings.warn("Not all matrices are `SparseDesignMatrix` objects. "
                          "Dense matrices will be converted to sparse matrices.", LightkurveWarning)
            matrices = [m if isinstance(m, SparseDesignMatrix) else m.copy().to_sparse() for m in matrices]
        self.matrices = matrices
        self.X = sparse.hstack([m.X for m in self.matrices], format="csr")
        self._child_class = SparseDesignMatrix
        self.validate()

    def __repr__(self):
        return "SparseDesignMatrixCollection:\n" + "".join(["\t{}\n".format(i.__repr__()) for i in self])


# ---- B-spline design matrices -------------------------------------------------------------------
def _bspline_basis_table(x, degree, knots):
    """All B-spline basis functions of `degree` on the knot vector `knots` at the points `x`, by the
    Cox-de Boor recurrence evaluated bottom-up: row i of the result is B_{i,degree}(x).  Conventions of the
    reference's recursive `_spline_basis_vector` (designmatrix.py:853-893): degree-0 pieces are 1 on the CLOSED
    interval [knots[i], knots[i+1]], and a term whose knot span is zero contributes nothing."""
    x = np.asarray(x, dtype=np.float64)
    knots = np.asarray(knots, dtype=np.float64)
    table = ((x[None, :] >= knots[:-1, None]) & (x[None, :] <= knots[1:, None])).astype(np.float64)
    for k in range(1, degree + 1):
        nrow = len(knots) - k - 1
        nxt = np.zeros((nrow, len(x)))
        for i in range(nrow):
            da = knots[i + k] - knots[i]
            db = knots[i + k + 1] - knots[i + 1]
            alpha1 = (x - knots[i]) / da if da != 0 else np.zeros(len(x))
            alpha2 = (knots[i + k + 1] - x) / db if db != 0 else np.zeros(len(x))
            nxt[i] = table[i] * alpha1 + table[i + 1] * alpha2
        table = nxt
    return table


def create_sparse_spline_matrix(x, n_knots=20, knots=None, degree=3, name="spline"):
    """B-spline basis of `x` as a `SparseDesignMatrix` (designmatrix.py:896-949).  Without `knots`, interior knots
    sit midway between the samples that end each of `n_knots - degree` equal-count chunks of sorted `x`.  The
    clamped knot vector repeats min(x) and max(x) `degree + 1` times; basis functions that vanish on all of `x`
    are dropped."""
    x = np.asarray(x, np.float64)
    if not isinstance(n_knots, int):
        raise ValueError("`n_knots` must be an integer.")
    if n_knots - degree <= 0:
        raise ValueError("n_knots must be greater than degree.")
    if (knots is None) and (n_knots is not None):
        order = np.argsort(x)
        ends = np.asarray([chunk[-1] for chunk in np.array_split(order, n_knots - degree)[:-1]])
        knots = [np.mean([x[k], x[k + 1]]) for k in ends]
    elif (knots is None) and (n_knots is None):
        raise ValueError("Pass either `n_knots` or `knots`.")
    inner = np.unique(np.append(np.append(x.min(), knots), x.max()))
    # the reference starts its recursion at knot index -1 with `degree` copies of min(x); that is the clamped
    # knot vector with `degree + 1` copies (the extra span has zero width and contributes nothing)
    clamped = np.concatenate([[x.min()] * degree, inner, [x.max()] * degree])
    table = _bspline_basis_table(x, degree, clamped)
    rows = [sparse.csr_matrix(row) for row in table if row.sum() != 0]
    return SparseDesignMatrix(sparse.vstack(rows, format="csr").T, name=name)


def create_spline_matrix(x, n_knots=20, knots=None, degree=3, name="spline", include_intercept=True):
    """B-spline basis of `x` as a dense `DesignMatrix` (designmatrix.py:952-1000).  The reference builds it with
    ``patsy.dmatrix("bs(x, df|knots, degree, include_intercept) - 1")``; patsy is absent here, so this restates
    patsy's `bs`: interior knots at equally spaced quantiles of `x` (numpy's default linear percentile) unless
    given, boundary knots min/max repeated `degree + 1` times, basis by `scipy.interpolate.splev`, the first
    basis function dropped when `include_intercept` is False."""
    from scipy.interpolate import splev
    x = np.asarray(x, dtype=np.float64)
    order = degree + 1
    if knots is not None:
        inner = np.asarray(knots, dtype=np.float64)
    else:
        n_inner = n_knots - order + (0 if include_intercept else 1)
        if n_inner < 0:
            raise ValueError("df={} is too small for degree={} and include_intercept={}; must be >= {}".format(
                n_knots, degree, include_intercept, order - (0 if include_intercept else 1)))
        quantiles = np.linspace(0, 1, n_inner + 2)[1:-1]
        inner = np.asarray([np.percentile(x, 100 * q) for q in quantiles])
    all_knots = np.sort(np.concatenate([[x.min(), x.max()] * order, inner]))
    n_bases = len(all_knots) - order
    basis = np.empty((len(x), n_bases))
    for i in range(n_bases):
        coefs = np.zeros(n_bases)
        coefs[i] = 1
        basis[:, i] = splev(x, (all_knots, coefs, degree))
    if not include_intercept:
        basis = basis[:, 1:]
    df = pd.DataFrame(basis, columns=["knot{}".format(idx + 1) for idx in range(basis.shape[1])])
    return DesignMatrix(df, name=name)
