"""Host-side (no GPU) behaviour the reference's own tests pin on the hot path's containers:
/root/reference/tests/test_periodogram.py:364-431 (error messages),
/root/reference/tests/correctors/test_designmatrix.py:12-141 (DesignMatrix / DesignMatrixCollection),
/root/reference/tests/correctors/test_regressioncorrector.py:86-118 (input validation of RegressionCorrector),
/root/reference/tests/test_lightcurve.py:242-392 (fold, cycle numbering, odd/even masks),
/root/reference/tests/correctors/test_sparsedesignmatrix.py:22-186 (SparseDesignMatrix, collections, splines;
the B-spline basis also against outputs of the reference's own recursion, tests/golden/spline_basis.npz)."""
import os
import warnings

import numpy as np
import pandas as pd
import pytest
from numpy.testing import assert_array_equal

import lightkurve_b200 as lk
from lightkurve_b200 import units as u
from scipy import sparse

from lightkurve_b200.correctors import (DesignMatrix, DesignMatrixCollection, RegressionCorrector,
                                        SparseDesignMatrix, SparseDesignMatrixCollection,
                                        create_sparse_spline_matrix, create_spline_matrix)
from lightkurve_b200.periodogram import Periodogram
from lightkurve_b200.utils import LightkurveWarning


def test_periodogram_error_messages():
    lc = lk.LightCurve(time=np.arange(1000), flux=np.random.normal(1, 0.1, 1000), flux_err=np.zeros(1000) + 0.1)
    with pytest.raises(ValueError):
        lc.to_periodogram(maximum_frequency=0.1, minimum_period=10)
    with pytest.raises(ValueError) as err:
        lc.to_periodogram(maximum_frequency=0.1, minimum_frequency=10)
    assert err.value.args[0] == "minimum_frequency cannot be larger than maximum_frequency"
    with pytest.raises(ValueError) as err:
        lc.to_periodogram(maximum_period=0.1, minimum_period=10)
    assert err.value.args[0] == "minimum_period cannot be larger than maximum_period"
    with pytest.raises(ValueError):
        lc.to_periodogram(frequency=np.arange(10), period=np.arange(10))
    cases = [
        (lambda: Periodogram([0], [1]), "frequency must be an `astropy.units.Quantity` object."),
        (lambda: Periodogram([0] * u.Hz, [1]), "power must be an `astropy.units.Quantity` object."),
        (lambda: Periodogram([0] * u.Hz, [1] * u.K), "frequency and power must have a length greater than 1."),
        (lambda: Periodogram([0, 1, 2, 3] * u.Hz, [1, 1] * u.K), "frequency and power must have the same length."),
        (lambda: Periodogram([0, 1, 2] * u.K, [1, 1, 1] * u.K), "Frequency must be in units of 1/time."),
        (lambda: Periodogram([0, 1, 2] * u.Hz, [1, 1, 1] * u.K).bin(binsize=-2),
         "binsize must be larger than or equal to 1"),
    ]
    for fn, msg in cases:
        with pytest.raises(ValueError) as err:
            fn()
        assert err.value.args[0] == msg
    for fn in (lambda: Periodogram([0, 1, 2] * u.Hz, [1, 1, 1] * u.K).bin(method="not-implemented"),
               lambda: Periodogram([0, 1, 2] * u.Hz, [1, 1, 1] * u.K).smooth(method="not-implemented")):
        with pytest.raises(ValueError) as err:
            fn()
        assert "method 'not-implemented' is not supported" in err.value.args[0]


def test_designmatrix_basics():
    size, name = 10, "testmatrix"
    df = pd.DataFrame({"vector1": np.ones(size), "vector2": np.zeros(size), "vector3": np.ones(size)})
    dm = DesignMatrix(df, name=name)
    assert dm.columns == ["vector1", "vector2", "vector3"]
    assert dm.name == name
    assert dm.shape == (size, 3)
    assert (dm["vector1"] == df["vector1"]).all()
    assert dm.append_constant().shape == (size, 4)
    assert dm.pca(nterms=2).shape == (size, 2)
    assert dm.split([10]).shape == (size, 6)
    dm.__repr__()
    dm = DesignMatrix(df, name=name)
    dm.append_constant(inplace=True)
    assert dm.shape == (size, 4)
    dm = DesignMatrix(df, name=name)
    dm.split([10], inplace=True)
    assert dm.shape == (size, 6)


def test_designmatrix_from_numpy_and_dict():
    dm = DesignMatrix(np.ones((10, 2)))
    assert dm.columns == [0, 1]
    assert dm.name == "unnamed_matrix"
    assert (dm[0] == np.ones(10)).all()
    dm = DesignMatrix({"centroid_col": np.ones(10), "centroid_row": np.ones(10)}, name="motion_systematics")
    assert dm.shape == (10, 2)
    assert (dm["centroid_col"] == np.ones(10)).all()


def test_split():
    dm = DesignMatrix({"a": np.linspace(0, 9, 10), "b": np.linspace(100, 109, 10)})
    assert dm.shape == (10, 2)
    assert dm.split(2).shape == (10, 4)
    assert dm.split([2, 8]).shape == (10, 6)
    assert (dm.split([2, 8]).values[2:, 0:2] == 0).all()
    assert (dm.split([2, 8]).values[:8, 4:] == 0).all()
    assert len(set(dm.split(2).columns)) == 4


def test_standardize_and_pca():
    dm = DesignMatrix({"const": np.ones(10)})
    assert (dm.standardize()["const"] == dm["const"]).all()
    np.random.seed(3)
    dm = DesignMatrix({"normal": np.random.normal(loc=5, scale=3, size=100)})
    assert np.round(np.median(dm.standardize()["normal"]), 3) == 0
    assert np.round(np.std(dm.standardize()["normal"]), 1) == 1
    dm.standardize(inplace=True)
    dm = DesignMatrix({"a": np.random.normal(10, 20, 10), "b": np.random.normal(40, 10, 10),
                       "c": np.random.normal(60, 5, 10)})
    for nterms in [1, 2, 3]:
        pc = dm.pca(nterms=nterms)
        assert pc.shape == (10, nterms)
        np.testing.assert_allclose(pc.values.T @ pc.values, np.eye(nterms), atol=1e-12)    # orthonormal components


def test_collection_basics():
    dm1 = DesignMatrix(np.ones((5, 1)), columns=["col1"], name="matrix1")
    dm2 = DesignMatrix(np.zeros((5, 2)), columns=["col2", "col3"], name="matrix2")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", LightkurveWarning)              # (an all-zero matrix has rank 0)
        dmc = DesignMatrixCollection([dm1, dm2])
        assert_array_equal(dmc["matrix1"].values, dm1.values)
        assert_array_equal(dmc["matrix2"].values, dm2.values)
        assert_array_equal(dmc.values, np.hstack((dm1.values, dm2.values)))
        dmc.__repr__()
        dmc = dm1.collect(dm2)
        assert_array_equal(dmc["matrix1"].values, dm1.values)
        assert_array_equal(dmc.values, np.hstack((dm1.values, dm2.values)))
        assert isinstance(dmc.to_designmatrix(), DesignMatrix)


def test_designmatrix_rank_warning():
    dm = DesignMatrix({"a": [1, 2, 3]})
    assert dm.rank == 1
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        dm.validate(rank=True)                                          # good rank: no warning
    with pytest.warns(LightkurveWarning, match="rank"):
        dm = DesignMatrix({"a": [1, 2, 3], "b": [1, 1, 1], "c": [1, 1, 1], "d": [1, 1, 1], "e": [3, 4, 5]})
        assert dm.rank == 2
        dm.validate(rank=True)


def test_regressioncorrector_input_validation():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", LightkurveWarning)
        bad = [lk.LightCurve(flux=[5, 10], flux_err=[np.nan, 1]), lk.LightCurve(flux=[np.nan, 10], flux_err=[1, 1])]
    for lc in bad:
        with pytest.raises(ValueError):
            RegressionCorrector(lc)
    RegressionCorrector(lk.LightCurve(flux=[5, 10], flux_err=[np.nan, np.nan]))     # all-NaN errors are allowed
    for fe in ([1, 0], [1, -10]):                                                   # regression test for #668
        with pytest.raises(ValueError):
            RegressionCorrector(lk.LightCurve(flux=[5, 10], flux_err=fe))


def test_lightcurve_fold():
    lc = lk.LightCurve(time=np.linspace(0, 10, 100), flux=np.zeros(100) + 1, targetid=999, label="mystar",
                       meta={"CCD": 2})
    fold = lc.fold(period=1)
    np.testing.assert_almost_equal(fold.phase.value[0], -0.5, 2)
    np.testing.assert_almost_equal(np.min(fold.phase.value), -0.5, 2)
    np.testing.assert_almost_equal(np.max(fold.phase.value), 0.5, 2)
    assert np.min(fold.cycle) == 0                       # reference issue #1397: fold() without epoch_time
    assert np.max(fold.cycle) == 10
    assert fold.targetid == lc.targetid and fold.label == lc.label
    assert set(lc.meta).issubset(set(fold.meta)) and lc.meta["CCD"] == fold.meta["CCD"]
    assert_array_equal(np.sort(fold.time_original.value), lc.time.value)
    fold = lc.fold(period=1, epoch_time=-0.1)
    np.testing.assert_almost_equal(fold.phase.value[0], -0.5, 2)
    np.testing.assert_almost_equal(np.max(fold.phase.value), 0.5, 2)
    assert np.min(fold.cycle) == 0 and np.max(fold.cycle) == 10
    kep = lk.LightCurve(time=lk.units.Time(np.linspace(0, 10, 100), format="bkjd"), flux=np.ones(100))
    with pytest.warns(LightkurveWarning, match="appears to be given in JD"):
        kep.fold(10, 2456600)
    fold = lc.fold(period=1.5, normalize_phase=False)
    np.testing.assert_almost_equal(np.max(fold.phase.value) - np.min(fold.phase.value), 1.5, 1)
    fold = lc.fold(period=1.5, normalize_phase=True)
    assert fold.time.unit == u.dimensionless_unscaled
    np.testing.assert_almost_equal(np.max(fold.phase.value) - np.min(fold.phase.value), 1, 1)
    assert len(fold) == 100
    fold_copy = fold.copy()
    assert_array_equal(fold.time.value, fold_copy.time.value)
    assert fold is not fold_copy and fold.flux is not fold_copy.flux
    lc.fold(period=1 * u.day, epoch_time=5 * u.day)      # reference issue #520: quantities accepted


@pytest.mark.parametrize("normalize_phase", [False, True])
def test_lightcurve_fold_odd_even_masks(normalize_phase):
    epoch_time, period = 3, 4
    lc = lk.LightCurve(time=np.linspace(0, 10, 100), flux=np.zeros(100), targetid=999, label="mystar", meta={"CCD": 2})
    lc.flux = u.Quantity(np.sin((period * 0.75 + lc.time.value - epoch_time) * 2 * np.pi / period))
    fold = lc.fold(period=period, epoch_time=epoch_time, epoch_phase=0.5, normalize_phase=normalize_phase)
    odd, even = fold.odd_mask, fold.even_mask
    assert len(odd) == len(fold.time) and np.all(odd == ~even)
    wrapped = lc.fold(period=period, epoch_time=epoch_time, epoch_phase=0.5, normalize_phase=normalize_phase,
                      wrap_phase=0.25)
    np.testing.assert_almost_equal(wrapped.phase.value[-1], 0.25 if normalize_phase else 0.25, decimal=1)
    t = fold.time_original.value
    # cycle 0: [0, 1), 1: [1, 5), 2: [5, 9), 3: [9, 10]
    expected_cycle = np.where(t < 1, 0, np.where(t < 5, 1, np.where(t < 9, 2, 3)))
    assert_array_equal(fold.cycle, expected_cycle)
    assert_array_equal(even, (t < 1) | ((t >= 5) & (t < 9)))


# ---- sparse design matrices (reference: tests/correctors/test_sparsedesignmatrix.py) ----------------
def test_sparse_designmatrix_basics():
    size, name = 10, "testmatrix"
    X = sparse.csr_matrix(np.vstack([np.ones(size), np.arange(size), np.arange(size) ** 2]).T)
    cols = ["vector1", "vector2", "vector3"]
    dm = SparseDesignMatrix(X, name=name, columns=cols)
    assert dm.columns == cols and dm.name == name and dm.shape == (size, 3)
    assert dm.append_constant().shape == (size, 4)
    assert dm.pca(nterms=2).shape == (size, 2)
    assert isinstance(dm.pca(nterms=2), SparseDesignMatrix)
    assert dm.split([5]).shape == (size, 6)
    assert "SparseDesignMatrix" in repr(dm)
    dm.append_constant(inplace=True)
    assert dm.shape == (size, 4)
    dm = SparseDesignMatrix(X, name=name, columns=cols)
    dm.split([5], inplace=True)
    assert dm.shape == (size, 6)
    with pytest.raises(ValueError, match="scipy.sparse"):
        SparseDesignMatrix(np.ones((3, 2)))
    with pytest.raises(ValueError, match="No such column"):
        dm["nope"]


def test_sparse_split():
    X = sparse.csr_matrix(np.vstack([np.linspace(0, 9, 10), np.linspace(100, 109, 10)]).T)
    dm = SparseDesignMatrix(X, columns=["a", "b"])
    assert dm.shape == (10, 2)
    assert dm.split(2).shape == (10, 4)
    assert dm.split([2, 8]).shape == (10, 6)
    assert (dm.split([2, 8]).values[2:, 0:2] == 0).all()
    assert (dm.split([2, 8]).values[:8, 4:] == 0).all()
    assert len(set(dm.split(4).columns)) == 4
    # same values as the dense split (dense keeps the all-zero copy of column "a" in block 0, sparse drops it)
    dense = DesignMatrix({"a": np.linspace(1, 10, 10), "b": np.linspace(100, 109, 10)})
    assert_array_equal(dense.to_sparse().split([2, 8]).values, dense.split([2, 8]).values)
    assert dm.split([0]) is dm and dm.split([10]) is dm                 # nothing to split at the ends
    sp = dm.split([2, 8])
    assert len(sp.prior_mu) == len(sp.prior_sigma) == sp.shape[1]


def test_sparse_standardize():
    dm = SparseDesignMatrix(sparse.csr_matrix(np.ones((10, 1))), columns=["const"])
    assert (dm.standardize()["const"] == dm["const"]).all()             # zero spread: unchanged
    rng = np.random.default_rng(5)
    v = rng.normal(loc=5, scale=3, size=100)
    dm = SparseDesignMatrix(sparse.csr_matrix(v[:, None]), columns=["normal"])
    z = dm.standardize()["normal"].ravel()
    np.testing.assert_allclose(z, (v - v.mean()) / v.std(ddof=1), rtol=1e-12)
    # zeros are "not stored": they stay zero and do not enter the statistics
    w = v.copy()
    w[::4] = 0
    z = SparseDesignMatrix(sparse.csr_matrix(w[:, None]), columns=["x"]).standardize()["x"].ravel()
    nz = w != 0
    assert (z[~nz] == 0).all()
    np.testing.assert_allclose(z[nz], (w[nz] - w[nz].mean()) / w[nz].std(ddof=1), rtol=1e-12)
    dm.standardize(inplace=True)


def test_sparse_collection_basics():
    size = 5
    dm1 = DesignMatrix(np.ones((size, 1)), columns=["col1"], name="matrix1").to_sparse()
    dm2 = DesignMatrix(np.zeros((size, 2)), columns=["col2", "col3"], name="matrix2").to_sparse()
    dmc = SparseDesignMatrixCollection([dm1, dm2])
    assert_array_equal(dmc["matrix1"].values, dm1.values)
    assert_array_equal(dmc["matrix2"].values, dm2.values)
    assert_array_equal(dmc.values, np.hstack((dm1.values, dm2.values)))
    assert sparse.issparse(dmc.X) and dmc.X.shape == (size, 3)
    assert "SparseDesignMatrixCollection" in repr(dmc)
    dmc = dm1.collect(dm2)
    assert isinstance(dmc, SparseDesignMatrixCollection)
    assert_array_equal(dmc.values, np.hstack((dm1.values, dm2.values)))
    dm1d = DesignMatrix(np.ones((size, 1)), columns=["col1"], name="matrix1")
    with pytest.warns(LightkurveWarning, match="Sparse matrices will be converted to dense matrices."):
        dmc = DesignMatrixCollection([dm1d, dm2])
    assert not np.any([sparse.issparse(d.X) for d in dmc])
    with pytest.warns(LightkurveWarning, match="Dense matrices will be converted to sparse matrices."):
        dmc = SparseDesignMatrixCollection([dm1d, dm2])
    assert np.all([sparse.issparse(d.X) for d in dmc])
    assert isinstance(dmc.to_designmatrix(), SparseDesignMatrix)


def test_sparse_designmatrix_rank():
    dm = DesignMatrix({"a": [1, 2, 3]}).to_sparse()
    assert dm.rank == 1
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        dm.validate(rank=True)
    dm = DesignMatrix({"a": [1, 2, 3], "b": [1, 1, 1], "c": [1, 1, 1], "d": [1, 1, 1], "e": [3, 4, 5]}).to_sparse()
    assert dm.rank == 2
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        dm.validate()                                                   # sparse: rank check off by default
    with pytest.warns(LightkurveWarning, match="rank"):
        dm.validate(rank=True)


def test_splines_dense_equals_sparse():
    x = np.linspace(0, 1, 100)
    dense = create_spline_matrix(x, knots=[0.1, 0.3, 0.6, 0.9], degree=2)
    sp = create_sparse_spline_matrix(x, knots=[0.1, 0.3, 0.6, 0.9], degree=2)
    assert np.allclose(dense.values, sp.values)
    assert isinstance(dense, DesignMatrix) and isinstance(sp, SparseDesignMatrix)
    # defaults: 20 cubic basis functions that sum to one everywhere (partition of unity)
    for m in (create_spline_matrix(x), create_sparse_spline_matrix(x)):
        assert m.shape == (100, 20)
        np.testing.assert_allclose(m.values.sum(axis=1), 1.0, atol=1e-14)
    assert list(create_spline_matrix(x).columns) == ["knot%d" % (i + 1) for i in range(20)]
    assert create_spline_matrix(x, n_knots=8, include_intercept=False).shape == (100, 8)
    with pytest.raises(ValueError, match="integer"):
        create_sparse_spline_matrix(x, n_knots=5.0)
    with pytest.raises(ValueError, match="greater than degree"):
        create_sparse_spline_matrix(x, n_knots=3, degree=3)


def test_sparse_spline_matches_reference_recursion():
    """Bit-exact against matrices produced by the reference's own `_spline_basis_vector` recursion
    (tests/golden/make_spline_golden.py)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "spline_basis.npz"))
    for i in range(int(g["ncases"])):
        n_knots, degree = (int(v) for v in g["cfg%d" % i])
        m = create_sparse_spline_matrix(g["x%d" % i], n_knots=n_knots, degree=degree)
        assert_array_equal(m.values, g["m%d" % i])


def test_regression_oracle_with_sparse_matrix_is_the_dense_problem():
    """`RegressionCorrector` hands a sparse collection to the GPU as its dense equivalent: the host-side
    conversion must reproduce the matrix exactly (the GPU parity tests then cover both)."""
    x = np.linspace(0, 10, 200)
    dmc = SparseDesignMatrixCollection([create_sparse_spline_matrix(x, n_knots=10),
                                        DesignMatrix(np.ones((200, 1)), name="offset").to_sparse()])
    X = RegressionCorrector._dense_X(dmc)
    assert isinstance(X, np.ndarray) and X.flags.c_contiguous and X.dtype == np.float64
    assert_array_equal(X, dmc.values)
    assert isinstance(RegressionCorrector._as_collection(create_sparse_spline_matrix(x)), SparseDesignMatrixCollection)


# ---- LightCurve.bin (reference: tests/test_lightcurve.py:707-767, 1543-1549, 2272-2276) -------------------
def test_bin():
    lc = lk.LightCurve(time=np.arange(10), flux=2 * np.ones(10), flux_err=2 ** 0.5 * np.ones(10))
    binned_lc = lc.bin(binsize=2)
    np.testing.assert_allclose(binned_lc.flux.value, 2 * np.ones(5))
    np.testing.assert_allclose(binned_lc.flux_err.value, np.sqrt(((2 ** 0.5) ** 2 + (2 ** 0.5) ** 2) / 2) * np.ones(5))
    assert len(binned_lc.time) == 5
    with pytest.raises(TypeError):
        lc.bin(method="doesnotexist")
    lc = lk.LightCurve(time=np.arange(10), flux=2 * np.ones(10))
    np.testing.assert_allclose(lc.bin(binsize=2).flux_err.value, np.zeros(5))        # no errors: std of the bin
    lc = lk.LightCurve(time=np.arange(2000), flux=np.random.normal(loc=42, scale=0.01, size=2000))
    assert np.round(lc.bin(2000).flux_err.value[0], 2) == 0.01                       # regression test for #500
    lk.LightCurve(flux=[0, 0, 0]).bin(bins=2)                                        # #1162
    lk.LightCurve(time=np.arange(50), flux=np.ones(50)).bin(binsize=15)              # #705
    with pytest.raises(ValueError, match="Only one of"):
        lc.bin(bins=2, binsize=2)
    with pytest.raises(ValueError, match="conflicts"):
        lc.bin(bins=2, time_bin_size=1)
    with pytest.raises(TypeError, match="integer"):
        lc.bin(bins=2.5)


def test_bin_semantics():
    lc = lk.LightCurve(time=np.arange(10), flux=np.arange(10.0), flux_err=np.ones(10), label="x")
    lc.meta["SECTOR"] = 99
    b = lc.bin(time_bin_size=5)
    assert b.meta == lc.meta                                                         # #1040
    np.testing.assert_allclose(b.time.value, [2.5, 7.5])                             # bin centres
    np.testing.assert_allclose(b.flux.value, [2.0, 7.0])                             # [0, 5) and [5, 10)
    b = lc.bin(time_bin_size=3, n_bins=5)                                            # the fifth bin [12, 15) is empty
    np.testing.assert_allclose(b.flux.value[:4], [1.0, 4.0, 7.0, 9.0])
    assert np.isnan(b.flux.value[4]) and np.isnan(b.flux_err.value[4])
    b = lc.bin(time_bin_size=u.Quantity(48, u.hour), time_bin_start=1.0, aggregate_func=np.nanmedian)
    np.testing.assert_allclose(b.flux.value, [1.5, 3.5, 5.5, 8.0])                   # 4 bins; t = 9 closes the last one
    shuffled = lc[np.array([3, 1, 2, 0, 9, 8, 7, 6, 5, 4])]
    np.testing.assert_allclose(shuffled.bin(time_bin_size=5).flux.value, [2.0, 7.0])
    np.testing.assert_allclose(lc.bin(bins=np.array([0, 4, 9])).flux.value, [1.5, 6.5])  # cadence-index edges
    folded = lk.LightCurve(time=np.arange(2000), flux=np.random.normal(loc=42, scale=0.01, size=2000)).fold(period=100)
    assert np.round(folded.bin(time_bin_size=100).flux_err.value[0], 2) == 0.01       # #927
