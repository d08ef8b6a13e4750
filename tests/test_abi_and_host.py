"""CPU tests (no GPU): the C-ABI library loads and exports every symbol of include/lkb200.h, fails
loudly without a device (no CPU fallback), and the host-side logic of the shim (grids, validation,
error strings, units, containers, sharding) mirrors the reference."""
import logging
import os
import re
import warnings

import numpy as np
import pytest
import scipy.signal

import lightkurve_b200 as lk
from lightkurve_b200 import _lib, engine, units as u
from lightkurve_b200.periodogram import Periodogram, LombScarglePeriodogram, BoxLeastSquaresPeriodogram
from lightkurve_b200.correctors import DesignMatrix, DesignMatrixCollection, RegressionCorrector
from oracle import ls as ols, bls as obls

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "lkb200.h")).read()
    declared = set(re.findall(r"\b(lkb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), "liblkb200.so lacks %s" % name
    assert declared == set(_lib.SIGNATURES), "ctypes signatures out of sync with the header"
    assert lib.lkb_version() >= 1


def test_no_cpu_fallback_without_device():
    if engine.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.EngineError, match="no CPU fallback"):
        engine.ls_power_ragged([np.arange(10.0)], [np.ones(10)], np.linspace(0.1, 1, 5))
    with pytest.raises(_lib.EngineError):
        engine.flatten([np.arange(10.0)], [np.ones(10)])
    with pytest.raises(_lib.EngineError):
        lk.LightCurve(time=np.arange(10), flux=np.ones(10)).to_periodogram()


def test_product_never_imports_the_oracle():
    import subprocess
    import sys
    code = ("import sys; import lightkurve_b200, lightkurve_b200.dist, lightkurve_b200.engine; "
            "bad=[m for m in sys.modules if m.split('.')[0]=='oracle']; assert not bad, bad")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lightkurve_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


@pytest.mark.parametrize("w,p", [(5, 2), (101, 2), (401, 2), (51, 3), (7, 1), (3, 1), (25, 4)])
def test_savgol_tables_match_scipy(w, p):
    lib = _lib.load()
    c = np.zeros(w)
    e = np.zeros(w * (w // 2))
    assert lib.lkb_savgol_tables(w, p, _lib.ptr(c), _lib.ptr(e)) == 0
    np.testing.assert_allclose(c, scipy.signal.savgol_coeffs(w, p)[::-1], rtol=0, atol=1e-12)
    # edges: savgol_filter(mode="interp") of a random window
    rng = np.random.default_rng(w)
    x = rng.normal(size=w)
    ref = scipy.signal.savgol_filter(x, w, p)              # length == window: both edges from the same fit
    half = w // 2
    edge = e.reshape(w, half)
    np.testing.assert_allclose(x @ edge, ref[:half], rtol=0, atol=1e-10)
    np.testing.assert_allclose(x[::-1] @ edge, ref[::-1][:half], rtol=0, atol=1e-10)


# ------------------------------------------------------------------ units / containers
def test_units_and_quantity():
    q = u.Quantity([1.0, 2.0, 3.0], "electron/s")
    assert (q ** 2 / u.microhertz).unit == (u.electron / u.s) ** 2 / u.microhertz
    assert np.sqrt(q * q).unit == u.electron / u.s
    f = u.Quantity([1.0, 2.0], 1 / u.day)
    np.testing.assert_allclose(f.to(u.microhertz).value, np.array([1.0, 2.0]) / 86400 * 1e6)
    assert (1.0 / f).unit == u.day and f.unit == 1 / u.day
    assert not u.dimensionless_unscaled and u.ppm
    with pytest.raises(u.UnitConversionError):
        q.to(u.day)


def test_lightcurve_container_semantics():
    lc = lk.LightCurve(time=np.arange(10), flux=np.arange(10.0) + 1)
    assert len(lc) == 10 and np.isnan(lc.flux_err.value).all() and lc.time.format == "jd"
    assert len(lc[2:5]) == 3 and lc[np.arange(10) % 2 == 0].flux.value[1] == 3
    lc2 = lk.LightCurve(flux=[1.0, np.nan, 3.0])
    assert np.array_equal(lc2.time.value, [0, 1, 2]) and len(lc2.remove_nans()) == 2
    lc[4:6] = np.nan
    assert np.isnan(lc.flux.value[4:6]).all() and len(lc.remove_nans()) == 8
    with pytest.raises(ValueError, match="is not supported"):
        lc.to_periodogram(method="not-a-method")
    coll = lk.LightCurveCollection([lc, lc2])
    assert len(coll) == 2 and coll[1] is lc2 and len(coll[[True, False]]) == 1


# ------------------------------------------------------------------ Periodogram / LS host logic
def test_periodogram_error_messages():
    """/root/reference/tests/test_periodogram.py:364-431"""
    K = u.K
    with pytest.raises(ValueError) as err:
        Periodogram([0], [1])
    assert err.value.args[0] == "frequency must be an `astropy.units.Quantity` object."
    with pytest.raises(ValueError) as err:
        Periodogram([0] * u.Hz, [1])
    assert err.value.args[0] == "power must be an `astropy.units.Quantity` object."
    with pytest.raises(ValueError) as err:
        Periodogram([0] * u.Hz, [1] * K)
    assert err.value.args[0] == "frequency and power must have a length greater than 1."
    with pytest.raises(ValueError) as err:
        Periodogram([0, 1, 2, 3] * u.Hz, [1, 1] * K)
    assert err.value.args[0] == "frequency and power must have the same length."
    with pytest.raises(ValueError) as err:
        Periodogram([0, 1, 2] * K, [1, 1, 1] * K)
    assert err.value.args[0] == "Frequency must be in units of 1/time."
    with pytest.raises(ValueError) as err:
        Periodogram([0, 1, 2] * u.Hz, [1, 1, 1] * K).bin(binsize=-2)
    assert err.value.args[0] == "binsize must be larger than or equal to 1"
    with pytest.raises(ValueError) as err:
        Periodogram([0, 1, 2] * u.Hz, [1, 1, 1] * K).bin(method="not-implemented")
    assert "method 'not-implemented' is not supported" in err.value.args[0]
    with pytest.raises(ValueError) as err:
        Periodogram([0, 1, 2] * u.Hz, [1, 1, 1] * K).smooth(method="not-implemented")
    assert "method 'not-implemented' is not supported" in err.value.args[0]


def test_ls_prepare_grid_defaults_and_conflicts(caplog):
    rng = np.random.default_rng(0)
    lc = lk.LightCurve(time=np.arange(1000), flux=rng.normal(1, 0.1, 1000), flux_err=np.zeros(1000) + 0.1)
    p = LombScarglePeriodogram._prepare(lc)
    grid, fs, nyq = ols.default_frequency_grid(np.arange(1000.0))
    assert len(p["frequency"]) == 2497 and p["frequency"].unit == 1 / u.day
    np.testing.assert_allclose(p["frequency"].value, grid, rtol=1e-15)
    assert p["ls_method"] == "fast" and p["default_view"] == "frequency" and p["oversample_factor"] == 5.0
    p = LombScarglePeriodogram._prepare(lc, normalization="psd")
    assert p["frequency"].unit == u.microhertz and p["oversample_factor"] == 1.0
    np.testing.assert_allclose(p["frequency"].to(1 / u.day).value[0], 1 / 999.0)
    # user grids passed through exactly (tests/test_periodogram.py:147-161)
    fr = np.arange(0.1, 1, 0.01)
    p = LombScarglePeriodogram._prepare(lc, frequency=fr)
    assert np.isclose(np.sum(fr - p["frequency"].value), 0, rtol=1e-14)
    with caplog.at_level(logging.WARNING):
        p = LombScarglePeriodogram._prepare(lc, period=np.arange(1.0, 10, 0.1))
    assert p["ls_method"] == "slow" and p["default_view"] == "period"
    assert "not evenly sampled in frequency" in caplog.text
    with pytest.raises(ValueError):
        LombScarglePeriodogram._prepare(lc, maximum_frequency=0.1, minimum_period=10)
    with pytest.raises(ValueError) as err:
        LombScarglePeriodogram._prepare(lc, maximum_frequency=0.1, minimum_frequency=10)
    assert err.value.args[0] == "minimum_frequency cannot be larger than maximum_frequency"
    with pytest.raises(ValueError) as err:
        LombScarglePeriodogram._prepare(lc, maximum_period=0.1, minimum_period=10)
    assert err.value.args[0] == "minimum_period cannot be larger than maximum_period"
    with pytest.raises(ValueError):
        LombScarglePeriodogram._prepare(lc, frequency=np.arange(10), period=np.arange(10))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        p = LombScarglePeriodogram._prepare(lc, nterms=2)                 # fast + nterms>1 -> warn, nterms=1
    assert p["nterms"] == 1 and any(issubclass(x.category, lk.LightkurveWarning) for x in w)
    # NaN removal happens before the grid is built (tests/test_periodogram.py:38-40)
    lc[400:500] = np.nan
    assert len(LombScarglePeriodogram._prepare(lc)["time"]) == 900


def test_bls_prepare_defaults_and_errors(caplog):
    rng = np.random.default_rng(1)
    lc = lk.LightCurve(time=np.linspace(0, 10, 200), flux=rng.normal(100, 0.1, 200), flux_err=np.zeros(200) + 0.1)
    p = BoxLeastSquaresPeriodogram._prepare(lc)
    lo, hi = obls.lk_default_period_bounds(np.linspace(0, 10, 200), obls.DEFAULT_DURATIONS)
    ref = obls.autoperiod(np.linspace(0, 10, 200), obls.DEFAULT_DURATIONS, lo, hi, frequency_factor=10)
    np.testing.assert_allclose(p["period"], ref, rtol=1e-15)
    assert p["dy"] is not None and p["objective"] == "likelihood" and p["oversample"] == 10
    with pytest.raises(ValueError, match="Periodogram is too large to evaluate"):
        BoxLeastSquaresPeriodogram._prepare(lc, frequency_factor=0.00001)
    with pytest.raises(ValueError, match="period"):
        BoxLeastSquaresPeriodogram._prepare(lk.LightCurve(time=[1, 2, 3], flux=[4, 5, 6]),
                                            period=[1, 2, 3, np.nan, 4])
    p = BoxLeastSquaresPeriodogram._prepare(lk.LightCurve(time=[1, 2, 3], flux=[4, 5, 6]), period=[1, 2, 3, 4, 5])
    assert np.array_equal(p["period"], [1, 2, 3, 4, 5]) and p["dy"] is None
    with pytest.raises(ValueError, match="maximum transit duration"):
        BoxLeastSquaresPeriodogram._prepare(lc, period=[0.2, 1.0], duration=0.3)


# ------------------------------------------------------------------ correctors host logic
def test_designmatrix_and_regressioncorrector_validation():
    """/root/reference/tests/correctors/test_designmatrix.py + test_regressioncorrector.py:86-118"""
    size = 10
    dm = DesignMatrix({"vector1": np.ones(size), "vector2": np.arange(size)}, name="matrix")
    assert dm.shape == (size, 2) and dm.columns == ["vector1", "vector2"]
    assert (dm.prior_mu == 0).all() and np.isinf(dm.prior_sigma).all()
    dm2 = dm.append_constant()
    assert dm2.shape == (size, 3) and dm2.columns[-1] == "offset" and len(dm2.prior_sigma) == 3
    assert dm.split([5]).shape == (size, 4)
    dmc = DesignMatrixCollection([dm, DesignMatrix(np.arange(size) ** 2.0, name="sq")])
    assert dmc.X.shape == (size, 3) and len(dmc.prior_mu) == 3 and dmc["sq"].name == "sq"
    assert repr(dm) == "matrix DesignMatrix (10, 2)"
    with pytest.warns(lk.LightkurveWarning, match="low rank"):
        DesignMatrix(np.ones((size, 4))).validate()
    with pytest.raises(ValueError, match="prior_sigma"):
        DesignMatrix(np.ones((size, 2)), prior_sigma=[1, 0]).validate()
    lc = lk.LightCurve(flux=[5, 10])
    lc.flux.view(np.ndarray)[1] = np.nan
    with pytest.raises(ValueError, match="NaNs in time or flux"):
        RegressionCorrector(lc)
    with pytest.raises(ValueError, match="NaNs in `flux_err`"):
        RegressionCorrector(lk.LightCurve(flux=[5.0, 10.0], flux_err=[1.0, np.nan]))
    with pytest.raises(ValueError, match="smaller than or equal to zero"):
        RegressionCorrector(lk.LightCurve(flux=[5.0, 10.0], flux_err=[1.0, 0.0]))
    RegressionCorrector(lk.LightCurve(flux=[5.0, 10.0]))                     # all-NaN flux_err is fine


def test_shard_by_length_balances_and_partitions():
    from lightkurve_b200.dist import shard_by_length
    rng = np.random.default_rng(3)
    lens = np.round(10 ** rng.uniform(np.log10(2000), np.log10(20000), 1000)).astype(int)
    shards = shard_by_length(lens, 8)
    allidx = np.sort(np.concatenate(shards))
    assert np.array_equal(allidx, np.arange(1000))
    work = np.array([lens[s].sum() for s in shards], dtype=float)
    assert work.max() / work.min() < 1.05


def test_fold_matches_astropy_timeseries_fold_semantics():
    """reference tests/test_lightcurve.py:1589-1606 (fold v2 API) + the phase formula of astropy TimeSeries.fold."""
    import lightkurve_b200 as lk
    from lightkurve_b200 import units as u
    lc = lk.LightCurve(time=np.linspace(0, 10, 100), flux=np.zeros(100) + 1)
    fld = lc.fold(period=1)
    fld2 = lc.fold(period=1 * u.day)
    np.testing.assert_array_equal(fld.phase.value, fld2.phase.value)
    assert isinstance(fld, lk.FoldedLightCurve) and (np.diff(fld.phase.value) >= 0).all()
    assert fld.phase.value.min() >= -0.5 and fld.phase.value.max() < 0.5
    fn = lc.fold(period=2.5, epoch_time=1.0, normalize_phase=True)
    assert fn.phase.unit == u.dimensionless_unscaled and fn.phase.value.min() >= -0.5 and fn.phase.value.max() < 0.5
    t = np.asarray(lc.time.value)
    want = np.sort(((t - 1.0) + 1.25) % 2.5 - 1.25) / 2.5
    np.testing.assert_allclose(fn.phase.value, want, rtol=0, atol=1e-15)
    np.testing.assert_array_equal(np.sort(fn.time_original.value), t)
    fw = lc.fold(period=2.0, epoch_time=0.0, wrap_phase=2.0)           # phases in [0, P)
    assert fw.phase.value.min() >= 0 and fw.phase.value.max() < 2.0
    with pytest.raises(ValueError):
        lc.fold(period=2.0, wrap_phase=3.0)
    assert fw.meta["PERIOD"].value == 2.0 and fw.cycle.max() == 5


def test_logmedian_windows_match_reference_loop():
    """Host-side window builder of Periodogram.smooth(method="logmedian") (periodogram.py:267-277): identical bin sets."""
    from lightkurve_b200.engine import logmedian_windows
    rng = np.random.default_rng(0)
    for f in ((np.arange(5000) + 1) * 0.0137, np.sort(rng.uniform(0.01, 300, 3000)), np.logspace(-2, 2, 777)):
        for fw in (0.01, 0.1, 0.033):
            lo, hi = logmedian_windows(f, fw)
            logf = np.log10(f)
            x0, w = logf[0], 0
            while x0 < logf[-1]:
                idx = np.flatnonzero(np.abs(logf - x0) < fw)
                if len(idx):
                    assert (idx[0], idx[-1] + 1, len(idx)) == (lo[w], hi[w], hi[w] - lo[w])
                else:
                    assert lo[w] == hi[w]
                x0 += 0.5 * fw
                w += 1
            assert w == len(lo)


def test_nccl_entry_points_without_a_communicator():
    """The exchange step of the C ABI (include/lkb200.h, multi-GPU section): NCCL is bound at run time, ids are
    128 bytes, and the collective refuses to run before lkb_nccl_init - all checkable without a GPU."""
    from lightkurve_b200 import _lib as L, engine
    assert engine.nccl_rank_world() == (-1, 0)
    with pytest.raises(ValueError, match="no communicator"):
        L.check(L.load().lkb_allgather_f32(None, 4, None, None))
    with pytest.raises(ValueError, match="no communicator"):
        engine.allgather_f32(None)
    with pytest.raises(ValueError, match="128 bytes"):
        engine.nccl_init(0, 1, b"short")
    if engine.nccl_version() > 0:                                   # an NCCL library is present in this image
        a, b = engine.nccl_unique_id(), engine.nccl_unique_id()
        assert len(a) == len(b) == 128 and a != b
    engine.nccl_shutdown()                                          # no communicator: a no-op


def test_diagnostic_workspace_readback_needs_an_initialised_engine():
    """lkb_ws_read (diagnostic entry used by tools/nufft_gpu_check.py) refuses to run without a bound device and
    validates its arguments; the slot names of the Python wrapper follow enum Slot of csrc/common.cuh."""
    from lightkurve_b200 import engine
    assert engine.WS_SLOTS["A"] == 0 and engine.WS_SLOTS["P"] == 15 and engine.WS_SLOTS["IN0"] == 16
    assert engine.WS_SLOTS["OUT0"] == 24 and engine.WS_SLOTS["X0"] == 32 and engine.WS_SLOTS["Y7"] == 47
    assert len(engine.WS_SLOTS) == 48
    with pytest.raises(ValueError, match="not initialised"):
        engine.ws_read("A", 4, np.float32)
