"""`CotrendingBasisVectors` on the host (no GPU): /root/reference/tests/correctors/test_cbvcorrector.py:40-196 ported
(the plotting calls dropped; astropy Tables replaced by plain mappings)."""
import warnings

import numpy as np
import pytest

import lightkurve_b200 as lk
from lightkurve_b200.correctors import CBVCorrector, CotrendingBasisVectors
from lightkurve_b200.units import Time
from lightkurve_b200.utils import LightkurveWarning


def test_constructor_and_designmatrix():
    data = {"CADENCENO": [1, 2, 3], "GAP": [False, True, False], "VECTOR_1": [2.0, 3.0, 4.0], "VECTOR_3": [3.0, 4.0, 5.0]}
    cbvs = CotrendingBasisVectors(data=data, time=Time([443.51090033, 443.53133457, 443.55176891]))
    assert cbvs.cbv_indices == [1, 3]
    assert np.all(cbvs.time.value == [443.51090033, 443.53133457, 443.55176891])
    cbvs = CotrendingBasisVectors(data={"VECTOR_3": [2.0, 3.0, 4.0], "VECTOR_12": [3.0, 4.0, 5.0]},
                                  time=[443.51090033, 443.53133457, 443.55176891])
    assert cbvs.cbv_indices == [3, 12]
    assert np.all(cbvs.gap_indicators == [False, False, False])       # GAP / CADENCENO are auto-initialised
    assert np.all(cbvs.cadenceno == [0, 1, 2])
    data = {"CADENCENO": [1, 2, 3], "GAP": [False, True, False], "VECTOR_1": [1.0, 2.0, 3.0],
            "VECTOR_2": [4.0, 5.0, 6.0], "VECTOR_3": [7.0, 8.0, 9.0]}
    cbvs = CotrendingBasisVectors(data, [1569.44053967, 1569.44192856, 1569.44331746])
    dm = cbvs.to_designmatrix(cbv_indices=[1, 3, 5], name="test cbv set")   # index 5 does not exist: ignored
    assert dm.shape == (3, 2) and dm.name == "test cbv set"
    assert np.all(dm["VECTOR_1"] == np.array([1.0, 2.0, 3.0]))
    assert np.all(dm["VECTOR_3"] == np.array([7.0, 8.0, 9.0]))
    with pytest.raises(KeyError):
        dm["VECTOR_2"]
    with pytest.raises(ValueError):
        cbvs.to_designmatrix(cbv_indices=[0, 1, 2])                   # 1-based indexing
    with pytest.raises(ValueError):
        cbvs.to_designmatrix("Doh!")
    assert cbvs.to_designmatrix().shape == (3, 3)


def test_align():
    sample_lc = lk.LightCurve(time=[1, 2, 3, 4, 6, 7], flux=[1, 2, 3, 4, 6, 7], flux_err=[0.1] * 6,
                              cadenceno=[1, 2, 3, 4, 6, 7])
    data = {"CADENCENO": [1, 2, 3, 5, 6], "GAP": [False, True, False, False, False], "VECTOR_1": [1.0, 2.0, 3.0, 5.0, 6.0]}
    cbvs = CotrendingBasisVectors(data, [1569.43915078, 1569.44053967, 1569.44192856, 1569.44470635, 1569.44609524])
    cbvs = cbvs.align(sample_lc)                                       # trims cadence 5, inserts NaNs at 4 and 7
    assert np.all(sample_lc.cadenceno == cbvs.cadenceno)
    assert len(cbvs.cadenceno) == 6
    assert np.all(cbvs.gap_indicators[[1, 3, 5]])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", LightkurveWarning)
        dm = cbvs.to_designmatrix(cbv_indices=[1])
    assert np.all(dm["VECTOR_1"][[0, 1, 2, 4]] == [1.0, 2.0, 3.0, 6.0])
    assert np.all(np.isnan(dm["VECTOR_1"][[3, 5]]))
    with pytest.raises(Exception, match="cadence numbers"):
        cbvs.align(lk.LightCurve(time=[1, 2], flux=[1, 2]))


def test_interpolate():
    n_lc, n_cbv = 20, 10
    x_lc = np.linspace(0.0, 2 * np.pi, num=n_lc)
    sample_lc = lk.LightCurve(time=x_lc, flux=np.sin(x_lc), flux_err=np.full(n_lc, 0.1), cadenceno=np.arange(n_lc))
    x = np.linspace(0.0, 2 * np.pi, num=n_cbv)
    data = {"CADENCENO": np.arange(n_cbv), "GAP": np.full(n_cbv, False), "VECTOR_1": np.cos(x),
            "VECTOR_2": np.sin(x + np.pi * 0.125)}
    it = CotrendingBasisVectors(data, x).interpolate(sample_lc, extrapolate=False)
    assert np.all(it.time.value == sample_lc.time.value)
    assert np.abs(it["VECTOR_1"].value - np.cos(x_lc)).max() < 0.05    # PCHIP through 10 samples of a cosine
    x = np.linspace(0.0, 1.5 * np.pi, num=n_cbv)
    data.update({"VECTOR_1": np.cos(x), "VECTOR_2": np.sin(x + np.pi * 0.125)})
    cbvs = CotrendingBasisVectors(data, x)
    outside = np.nonzero(sample_lc.time.value > 1.5 * np.pi)[0]
    assert np.all(cbvs.interpolate(sample_lc, extrapolate=False)["VECTOR_1"].value[outside] == 0.0)
    assert np.all(cbvs.interpolate(sample_lc, extrapolate=True)["VECTOR_1"].value[outside] != 0.0)


def test_cbvcorrector_constructor_without_gpu():
    """Constructor behaviour of test_CBVCorrector (:339-354) that needs no kernel."""
    from lightkurve_b200 import units as u
    lc = lk.LightCurve(time=[1, 2, 3, 4, 5], flux=[1, 2, np.nan, 4, 5], flux_err=[0.1] * 5, cadenceno=[1, 2, 3, 4, 5],
                       flux_unit=u.electron / u.second)
    c = CBVCorrector(lc, do_not_load_cbvs=True)
    assert len(c.lc.flux) == 4                                         # the NaN is gone
    np.testing.assert_allclose(np.nanmedian(c.lc.flux.value), np.nanmedian(lc.flux.value))
    assert "no CBVs" in repr(c)
    with pytest.raises(NotImplementedError, match="MAST"):
        CBVCorrector(lc)                                               # loading CBV files is out of scope
    with pytest.raises(AssertionError, match="e-/s"):
        CBVCorrector(lk.LightCurve(time=[1, 2], flux=[1, 2], flux_err=[0.1, 0.1]), do_not_load_cbvs=True)
    with pytest.raises(Exception, match="interpolate_cbvs must be True"):
        CBVCorrector(lc, extrapolate_cbvs=True, do_not_load_cbvs=True)
