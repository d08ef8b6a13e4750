"""GPU test of the opt-in NUFFT Lomb-Scargle path (algo="nufft", lightkurve_b200/csrc/ls_nufft.cu).

Status: the arithmetic of every kernel is verified on the CPU (tests/test_nufft_core.py runs the same
`__host__ __device__` functions through a g++ harness) and the whole translation unit runs on a CUDA-on-CPU layer
against the oracle (tests/test_nufft_emulated.py).  Written after round 1's GPU budget was spent (then marked xfail);
it has passed on hardware in every run of round 2, so the marks are gone.  The work runs in a child process so that a
device fault cannot poison the CUDA context of the rest of the suite.  The file name sorts last on purpose."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _excess(got, ref):
    return float(np.max(np.abs(got - ref) / (1e-5 * ref.max() + 1e-4 * ref)))


def _worker(out_q):
    sys.path.insert(0, ROOT)
    from lightkurve_b200 import engine
    from oracle import ls as ols
    engine.init(0)
    os.environ["LKB_NUFFT_VERIFY"] = "1"           # the path's built-in self-check (direct fp64 spot sums) stays on
    rng = np.random.default_rng(21)
    res = {}
    # (1) default lightkurve grid (f0 = df = 1 / (5 T)), odd batch, amplitudes spanning 3 decades
    idx = np.flatnonzero(rng.uniform(size=4000) > 0.1)[:3000]
    t = 131.5 + idx * 0.0204336
    N = len(t)
    df = 1.0 / (5.0 * (t[-1] - t[0]))
    freq = df * (1 + np.arange(2000))
    Y = np.stack([1 + a * np.sin(2 * np.pi * f * t) + s * rng.normal(size=N)
                  for a, f, s in [(1e-2, 3.1, 1e-4), (0, 1, 5e-5), (1e-4, 7.7, 3e-4), (3e-3, 0.9, 1e-3), (0, 1, 1e-3)]])
    got = np.asarray(engine.ls_power_shared(t, Y.astype(np.float32), freq, "amplitude", algo="nufft"), dtype=np.float64)
    worst = 0.0
    for b in range(len(Y)):
        y = Y[b].astype(np.float32).astype(np.float64)
        ref = np.sqrt(ols.ls_slow_psd(t, y, freq)) * np.sqrt(4.0 / N)
        worst = max(worst, _excess(got[b], ref))
    res["default grid, amplitude"] = worst
    # (2) psd normalisation, oversample 1 (df * baseline = 1: the last cadence wraps around the fine grid), f64 flux
    df1 = 1.0 / (t[-1] - t[0])
    freq1 = df1 * (1 + np.arange(700))
    scale = 2.0 / (N * 1.0 * df1)
    got = np.asarray(engine.ls_power_shared(t, Y[:2], freq1, "psd", norm_scale=scale, algo="nufft"), dtype=np.float64)
    worst = 0.0
    for b in range(2):
        ref = ols.ls_slow_psd(t, Y[b], freq1) * scale
        worst = max(worst, float(np.max(np.abs(got[b] - ref) / (2e-5 * ref.max() + 2e-4 * ref))))
    res["oversample 1, psd"] = worst
    # (3) against the CUDA-core contraction kernel on a larger batch (sizes the oracle does not reach)
    B, N3, F3 = 64, 20000, 30000
    t3 = np.sort(rng.uniform(0, 90.0, N3))
    df3 = 1.0 / (5.0 * (t3[-1] - t3[0]))
    f3 = df3 * (1 + np.arange(F3))
    Y3 = (1 + 1e-3 * np.sin(2 * np.pi * 2.2 * t3)[None, :] + 3e-4 * rng.normal(size=(B, N3))).astype(np.float32)
    a = np.asarray(engine.ls_power_shared(t3, Y3, f3, "amplitude", algo="nufft"), dtype=np.float64)
    s = np.asarray(engine.ls_power_shared(t3, Y3, f3, "amplitude", algo="simt"), dtype=np.float64)
    res["vs simt kernel"] = max(_excess(a[b], s[b]) for b in range(B))
    # (5) host-mode call with more than 256 light curves: the chunk-pipelined loop (two streams, two buffer sets)
    B5, N5, F5 = 300, 1500, 1000
    t5 = np.sort(rng.uniform(0, 40.0, N5))
    f5 = (1.0 / (5.0 * (t5[-1] - t5[0]))) * (1 + np.arange(F5))
    Y5 = (1 + 2e-3 * np.sin(2 * np.pi * 1.3 * t5)[None, :] + 10 ** rng.uniform(-4, -3, (B5, 1)) *
          rng.normal(size=(B5, N5))).astype(np.float32)
    a5 = np.asarray(engine.ls_power_shared(t5, Y5, f5, "amplitude", algo="nufft"), dtype=np.float64)
    s5 = np.asarray(engine.ls_power_shared(t5, Y5, f5, "amplitude", algo="simt"), dtype=np.float64)
    res["pipelined vs simt kernel"] = max(_excess(a5[b], s5[b]) for b in range(B5))
    # (4) shapes the path must refuse
    try:
        engine.ls_power_shared(t, Y[:2].astype(np.float32), np.sort(rng.uniform(0.1, 5, 100)), "amplitude", algo="nufft")
        res["irregular grid refused"] = False
    except Exception:
        res["irregular grid refused"] = True
    out_q.put(res)


def _worker_ragged(out_q):
    """Ragged batches: LKB_LS_RAGGED_NUFFT=1 routes lkb_ls_power (one shared regular grid) through the NUFFT kernels."""
    sys.path.insert(0, ROOT)
    from lightkurve_b200 import engine
    from oracle import ls as ols
    engine.init(0)
    rng = np.random.default_rng(33)
    times, fluxes = [], []
    for i in range(7):                                           # odd count: the last pair is half empty
        n = int(rng.integers(300, 4000))
        t = 1325.0 + np.sort(rng.uniform(0, 27.4 * rng.uniform(0.5, 1.0), n))
        times.append(t)
        a = [1e-2, 0.0, 1e-4, 3e-3, 0.0, 1e-3, 2e-5][i]
        fluxes.append((1 + a * np.sin(2 * np.pi * 2.2 * t) + 10 ** rng.uniform(-4.3, -3) * rng.normal(size=n)).astype(
            np.float32 if i % 2 else np.float64))
    F = 3000
    freq = np.linspace(12.0 / F, 12.0, F)                        # f0 = df: k0 = 1; df * baseline <= 0.11
    os.environ["LKB_LS_RAGGED_NUFFT"] = "1"
    got = np.asarray(engine.ls_power_ragged(times, [f.astype(np.float64) for f in fluxes], freq, "amplitude"),
                     dtype=np.float64)
    os.environ["LKB_LS_RAGGED_NUFFT"] = "0"
    direct = np.asarray(engine.ls_power_ragged(times, [f.astype(np.float64) for f in fluxes], freq, "amplitude"),
                        dtype=np.float64)
    res = {"vs direct kernel": max(_excess(got[b], direct[b]) for b in range(len(times)))}
    worst = 0.0
    for b in (0, 1, 6):
        y = fluxes[b].astype(np.float64)
        ref = np.sqrt(ols.ls_slow_psd(times[b], y, freq)) * np.sqrt(4.0 / len(y))
        worst = max(worst, _excess(got[b], ref))
    res["vs oracle"] = worst
    res["differs from direct"] = bool(np.any(got != direct))    # proves the NUFFT kernels actually ran
    # a batch with an unsorted light curve falls back to the direct kernel (same numbers as without the switch)
    os.environ["LKB_LS_RAGGED_NUFFT"] = "1"
    t_bad = times[0][::-1].copy()
    fb = engine.ls_power_ragged([t_bad, times[1]], [fluxes[0][::-1].astype(np.float64), fluxes[1].astype(np.float64)],
                                freq, "amplitude")
    res["fallback ok"] = _excess(np.asarray(fb[0], dtype=np.float64), direct[0]) < 1.0
    out_q.put(res)


def _run_child(target):
    import queue
    import time
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    p = ctx.Process(target=target, args=(out_q,), daemon=True)
    p.start()
    res, t0 = None, time.time()
    try:
        while res is None:
            try:
                res = out_q.get(timeout=1.0)
            except queue.Empty:
                assert p.is_alive() or p.exitcode == 0, "worker died with exit code %r" % p.exitcode
                assert time.time() - t0 < 300, "timed out"
    finally:
        if p.is_alive():
            p.join(timeout=30)
        if p.is_alive():
            p.kill()
    return res


def test_ragged_nufft_path_matches_direct_kernel_and_oracle():
    res = _run_child(_worker_ragged)
    print("ragged NUFFT path, worst tolerance excess per case:", res)
    assert res["differs from direct"] is True
    assert res["vs direct kernel"] < 1.0
    assert res["vs oracle"] < 1.0
    assert res["fallback ok"] is True


def test_nufft_path_matches_oracle_and_simt_kernel():
    import queue
    import time
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(out_q,), daemon=True)
    p.start()
    res, t0 = None, time.time()
    try:
        while res is None:
            try:
                res = out_q.get(timeout=1.0)
            except queue.Empty:
                assert p.is_alive() or p.exitcode == 0, "worker died with exit code %r" % p.exitcode
                assert time.time() - t0 < 300, "timed out"
    finally:
        if p.is_alive():
            p.join(timeout=30)
        if p.is_alive():
            p.kill()
    print("NUFFT path, worst tolerance excess per case:", res)
    assert res["irregular grid refused"] is True
    assert res["default grid, amplitude"] < 1.0
    assert res["oversample 1, psd"] < 1.0
    assert res["vs simt kernel"] < 1.0
    assert res["pipelined vs simt kernel"] < 1.0
