"""The bodies of the shim-level GPU tests (tests/test_gpu_shim.py - the reference's own tests of the hot path,
ported) run on the CPU with the engine entry points replaced by the fp64 oracle (tests/_oracle_engine.py).

Purpose: the Python layer between the user and the C ABI - argument handling, frequency / period grids, units,
normalisations, sigma-clip bookkeeping, result objects - and the EXPECTATIONS of those tests are verified here
without a GPU, so on the B200 box only the kernels themselves can make them fail.  This is test infrastructure:
the product never imports the oracle (tests/test_abi_and_host.py::test_product_never_imports_the_oracle) and has no
CPU fallback."""
import importlib
import inspect

import pytest

import _oracle_engine

SKIP = {
    # asserts on device-only behaviour (kernel selection / batching equalities computed twice on the device)
}


def _shim_tests():
    mod = importlib.import_module("test_gpu_shim")
    out = []
    for name, fn in inspect.getmembers(mod, inspect.isfunction):
        if name.startswith("test_") and fn.__module__ == mod.__name__ and name not in SKIP:
            out.append((name, fn))
    return mod, out


_MOD, _TESTS = _shim_tests()


def _cases():
    import itertools
    for name, fn in _TESTS:
        marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
        axes = []
        for m in marks:                                         # stacked parametrize marks: cartesian product
            argnames = [a.strip() for a in m.args[0].split(",")]
            rows = []
            for values in m.args[1]:
                values = tuple(values) if isinstance(values, (tuple, list)) and len(argnames) > 1 else (values,)
                rows.append(dict(zip(argnames, values)))
            axes.append(rows)
        for combo in itertools.product(*axes):
            params = {}
            for d in combo:
                params.update(d)
            ident = name if not params else "%s[%s]" % (name, "-".join(str(v) for v in params.values()))
            yield pytest.param(fn, params, id=ident)


@pytest.fixture
def oracle_engine(monkeypatch):
    from lightkurve_b200 import engine
    for name, fake in _oracle_engine.FAKES.items():
        monkeypatch.setattr(engine, name, fake)
    yield engine


@pytest.mark.parametrize("fn,params", list(_cases()))
def test_gpu_shim_test_body_on_the_oracle(fn, params, oracle_engine, caplog, monkeypatch):
    kwargs = dict(params)
    sig = inspect.signature(fn).parameters
    if "caplog" in sig:
        kwargs["caplog"] = caplog
    if "monkeypatch" in sig:
        kwargs["monkeypatch"] = monkeypatch
    if "engine" in sig:
        kwargs["engine"] = oracle_engine
    import numpy as np
    from conftest import seed_for
    np.random.seed(seed_for(fn.__name__))            # the data the GPU run of this test will see
    fn(**kwargs)
