"""CPU verification of the NUFFT Lomb-Scargle path (lightkurve_b200/csrc/ls_nufft.cu): every kernel of that path
is one call of a `__host__ __device__` function of csrc/nufft_core.h per thread, so the SAME code is compiled
here for the host (tests/native/nufft_host_harness.cpp, g++) and checked against numpy's FFT, direct trig sums
and the fp64 oracle - index arithmetic of the Stockham passes, butterflies, twiddles, gather spreading with
wrap-around, pair packing, deconvolution.  The CUDA-only glue (launch shapes, epilogue, low rows) is covered by
the opt-in GPU test in tests/test_gpu_zz_nufft.py."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import ls as ols

HERE = os.path.dirname(os.path.abspath(__file__))
c_vp, c_i64 = ctypes.c_void_p, ctypes.c_int64


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("nufft") / "libnufft_harness.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out,
                           os.path.join(HERE, "native", "nufft_host_harness.cpp")])
    lib = ctypes.CDLL(out)
    lib.harness_fft.argtypes = [c_vp, ctypes.c_int, c_vp]
    lib.harness_gauss_legendre.argtypes = [ctypes.c_int, c_vp, c_vp]
    lib.harness_trig_sums.argtypes = [c_vp, c_i64, c_vp, c_vp, ctypes.c_double, c_i64, c_i64, ctypes.c_int,
                                      c_vp, c_vp, c_vp, c_vp]
    lib.harness_set_chain.argtypes = [ctypes.c_int]
    lib.harness_stages.argtypes = [c_vp, c_i64, c_vp, c_vp, ctypes.c_double, c_i64, c_i64, ctypes.c_int,
                                   c_vp, c_vp, c_vp, c_vp]
    lib.harness_trig_sums_ex.argtypes = [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, ctypes.c_double, c_i64, c_i64,
                                         ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp, c_vp]
    return lib


def trig_sums(lib, t, y0, y1, df, k0, F, w=8):
    t = np.ascontiguousarray(t, dtype=np.float64)
    y0 = np.ascontiguousarray(y0, dtype=np.float32)
    out = [np.zeros(F, np.float32) for _ in range(4)]
    if y1 is None:
        lib.harness_trig_sums(t.ctypes.data, len(t), y0.ctypes.data, None, df, k0, F, w, out[0].ctypes.data,
                              out[1].ctypes.data, None, None)
    else:
        y1 = np.ascontiguousarray(y1, dtype=np.float32)
        lib.harness_trig_sums(t.ctypes.data, len(t), y0.ctypes.data, y1.ctypes.data, df, k0, F, w,
                              *[o.ctypes.data for o in out])
    return [o.astype(np.float64) for o in out]


def test_gauss_legendre_nodes(harness):
    x, w = np.zeros(32), np.zeros(32)
    harness.harness_gauss_legendre(32, x.ctypes.data, w.ctypes.data)
    rx, rw = np.polynomial.legendre.leggauss(32)
    np.testing.assert_allclose(x, rx, atol=1e-15)
    np.testing.assert_allclose(w, rw, atol=1e-14)


@pytest.mark.parametrize("p", range(4, 14))                 # every radix mix: 16^a x {-, 2, 4, 8}
def test_stockham_passes_equal_numpy_fft(harness, p):
    rng = np.random.default_rng(p)
    M = 1 << p
    z = (rng.normal(size=M) + 1j * rng.normal(size=M)).astype(np.complex64)
    out = np.zeros(M, np.complex64)
    harness.harness_fft(z.ctypes.data, p, out.ctypes.data)
    ref = np.fft.ifft(z.astype(np.complex128)) * M            # +i sign
    assert np.abs(out - ref).max() <= 5e-7 * np.abs(ref).max()


@pytest.mark.parametrize("N,F,k0,oversample,w,tol", [(500, 300, 1, 5, 8, 2e-7), (2000, 1000, 0, 1.0, 8, 3e-7),
                                                     (3000, 5000, 3, 2, 6, 6e-7), (777, 64, 1, 5, 8, 2e-7),
                                                     (1500, 700, 1, 1.0, 10, 2e-7)])
def test_trig_sums_equal_direct_sums(harness, N, F, k0, oversample, w, tol):
    """oversample = 1 puts the last cadence on the wrap-around of the fine grid (df * baseline = 1)."""
    rng = np.random.default_rng(N)
    t = np.sort(rng.uniform(0, 27.0, N))
    t -= t[0]
    df = 1.0 / (oversample * t[-1])
    y0 = rng.normal(size=N).astype(np.float32)
    y1 = (np.sin(2 * np.pi * 1.3 * t) + 0.1 * rng.normal(size=N)).astype(np.float32)
    C0, S0, C1, S1 = trig_sums(harness, t, y0, y1, df, k0, F, w)
    ph = 2 * np.pi * np.outer((k0 + np.arange(F)) * df, t)
    for y, C, S in ((y0, C0, S0), (y1, C1, S1)):
        yd = y.astype(np.float64)
        err = max(np.abs(C - np.cos(ph) @ yd).max(), np.abs(S - np.sin(ph) @ yd).max())
        assert err <= tol * np.abs(yd).sum()
    # single light curve (imaginary part empty) gives the same numbers for the first of the pair
    C0s, S0s, _, _ = trig_sums(harness, t, y0, None, df, k0, F, w)
    np.testing.assert_allclose(C0s, C0, atol=2e-7 * np.abs(y0).sum())
    np.testing.assert_allclose(S0s, S0, atol=2e-7 * np.abs(y0).sum())


def _power_from_sums(N, ch, sh, C, S, C2, S2, ysum):
    """ls_rotation + ls_power_from_sums of csrc/ls_common.cuh, with the window sums the way nufft_rot_kernel
    builds them: sum cos^2 = (N + C2) / 2, sum sin cos = S2 / 2."""
    Sb, Cb, CCb, SCb = S / N, C / N, 0.5 * (N + C2) / N, 0.5 * S2 / N
    SSb = 1.0 - CCb
    s2 = 2.0 * SCb - 2.0 * Sb * Cb
    c2 = (2.0 * CCb - 1.0) - (Cb * Cb - Sb * Sb)
    ta = 0.5 * np.arctan2(s2, c2)
    ct, st = np.cos(ta), np.sin(ta)
    ctau, stau = Cb * ct + Sb * st, Sb * ct - Cb * st
    ccp = CCb * ct * ct + 2 * SCb * ct * st + SSb * st * st - ctau * ctau
    ssp = SSb * ct * ct - 2 * SCb * ct * st + CCb * st * st - stau * stau
    Y = ysum / N
    yc = (ch * ct + sh * st) / N - Y * ctau
    ys = (sh * ct - ch * st) / N - Y * stau
    return 0.5 * N * (yc * yc / ccp + ys * ys / ssp)


@pytest.mark.parametrize("w,limit", [(8, 0.1), (6, 0.5)])
def test_full_pipeline_meets_the_parity_tolerance(harness, w, limit):
    """Kepler-like shared grid with gaps, lightkurve's default frequency grid (f0 = df = 1 / (5 T)): amplitude
    spectra of a noise-only and a signal light curve within the GPU parity bound
    |P - P_slow64| <= 1e-5 max(P) + 1e-4 P  (rows above the low-frequency cut, as in nufft_finish_kernel)."""
    rng = np.random.default_rng(3)
    idx = np.flatnonzero(rng.uniform(size=5200) > 0.12)[:4000]
    t = idx * 0.0204336
    N = len(t)
    df = 1.0 / (5.0 * t[-1])
    F, k0 = 6000, 1
    freq = (k0 + np.arange(F)) * df
    noise = (1 + 3e-4 * rng.normal(size=N)).astype(np.float32)
    signal = (1 + 5e-3 * np.sin(2 * np.pi * 7.3 * t) + 2e-4 * rng.normal(size=N)).astype(np.float32)
    ys = [y.astype(np.float64) - y.astype(np.float64).mean() for y in (noise, signal)]
    yc = [y.astype(np.float32) for y in ys]
    C0, S0, C1, S1 = trig_sums(harness, t, yc[0], yc[1], df, k0, F, w)
    Cw, Sw, _, _ = trig_sums(harness, t, np.ones(N, np.float32), None, df, 0, 2 * (k0 + F), w)
    kk = k0 + np.arange(F)
    low = freq * t[-1] <= 2.0
    for y64, y32, ch, sh in ((ys[0], yc[0], C0, S0), (ys[1], yc[1], C1, S1)):
        p = _power_from_sums(float(N), ch, sh, Cw[kk], Sw[kk], Cw[2 * kk], Sw[2 * kk], float(y32.astype(np.float64).sum()))
        amp = np.sqrt(np.maximum(p, 0)) * np.sqrt(4.0 / N)
        ref = np.sqrt(ols.ls_slow_psd(t, y64 + 1.0, freq)) * np.sqrt(4.0 / N)
        excess = np.abs(amp - ref) / (1e-5 * ref.max() + 1e-4 * ref)
        assert excess[~low].max() < limit


def test_pair_members_of_very_different_amplitude(harness):
    """Two light curves share one fp32 transform; the power-of-two pre-scaling keeps the small one's error
    relative to ITS OWN size (without it the error below is ~1000x larger)."""
    rng = np.random.default_rng(8)
    N, F = 3000, 2000
    t = np.sort(rng.uniform(0, 40.0, N))
    t -= t[0]
    df = 1.0 / (5.0 * t[-1])
    small = (1e-5 * rng.normal(size=N)).astype(np.float32)
    big = (1e-2 * np.sin(2 * np.pi * 2.0 * t)).astype(np.float32)
    C0, S0, C1, S1 = trig_sums(harness, t, small, big, df, 1, F)
    ph = 2 * np.pi * np.outer((1 + np.arange(F)) * df, t)
    for y, C, S in ((small, C0, S0), (big, C1, S1)):
        yd = y.astype(np.float64)
        err = max(np.abs(C - np.cos(ph) @ yd).max(), np.abs(S - np.sin(ph) @ yd).max())
        assert err <= 2e-7 * np.abs(yd).sum()


def test_ragged_pair_with_different_cadence_sets(harness):
    """Ragged batches (K1): the two light curves of a pair have their own times; cell ranges come from binary
    searches in each light curve's sorted cadence table (`spread_cell_search`)."""
    rng = np.random.default_rng(12)
    ta = np.sort(rng.uniform(0, 27.4, 2500))
    ta -= ta[0]
    tb = np.sort(rng.uniform(0, 20.0, 900))
    tb -= tb[0]
    F, k0, w = 1500, 1, 8
    df = 1.0 / (5.0 * 27.4)
    ya = rng.normal(size=len(ta)).astype(np.float32)
    yb = (3e-3 * np.sin(2 * np.pi * 0.7 * tb)).astype(np.float32)
    out = [np.zeros(F, np.float32) for _ in range(4)]
    harness.harness_trig_sums_ex(ta.ctypes.data, len(ta), ya.ctypes.data, tb.ctypes.data, len(tb), yb.ctypes.data,
                                 df, k0, F, w, 1, *[o.ctypes.data for o in out])
    f = (k0 + np.arange(F)) * df
    for t, y, C, S in ((ta, ya, out[0], out[1]), (tb, yb, out[2], out[3])):
        ph = 2 * np.pi * np.outer(f, t)
        yd = y.astype(np.float64)
        err = max(np.abs(C - np.cos(ph) @ yd).max(), np.abs(S - np.sin(ph) @ yd).max())
        assert err <= 2e-7 * np.abs(yd).sum()
    # the search variant and the table variant are the same arithmetic
    ref = trig_sums(harness, ta, ya, None, df, k0, F, w)
    out2 = [np.zeros(F, np.float32) for _ in range(2)]
    harness.harness_trig_sums_ex(ta.ctypes.data, len(ta), ya.ctypes.data, ta.ctypes.data, len(ta), None, df, k0, F, w,
                                 1, out2[0].ctypes.data, out2[1].ctypes.data, None, None)
    np.testing.assert_array_equal(out2[0], ref[0].astype(np.float32))
    np.testing.assert_array_equal(out2[1], ref[1].astype(np.float32))


def test_tiny_and_degenerate_inputs(harness):
    """1-40 cadences, duplicate times, k0 = 0 (mode 0), 2-17 bins, df * baseline = 1 and 1/3."""
    rng = np.random.default_rng(4)
    worst = 0.0
    for N in (1, 2, 3, 5, 9, 40):
        for F in (2, 3, 17):
            for k0 in (0, 1, 5):
                for oversample in (1.0, 3.0):
                    t = np.sort(rng.uniform(0, 10, N))
                    t -= t[0]
                    if N > 2:
                        t[1] = t[2]
                    df = 1.0 / (oversample * max(t[-1], 1.0))
                    y = rng.normal(size=N).astype(np.float32)
                    C, S, _, _ = trig_sums(harness, t, y, None, df, k0, F)
                    ph = 2 * np.pi * np.outer((k0 + np.arange(F)) * df, t)
                    err = max(np.abs(C - np.cos(ph) @ y).max(), np.abs(S - np.sin(ph) @ y).max())
                    worst = max(worst, err / np.abs(y).sum())
    assert worst < 2e-6


def test_gpu_bringup_tool_accepts_the_harness_buffers(harness):
    """tools/nufft_gpu_check.py compares the GPU's intermediate buffers with reference formulas written in numpy;
    fed with the buffers of the CPU harness (same layouts) every stage must pass - and a corrupted buffer must not."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("nufft_gpu_check", os.path.join(os.path.dirname(HERE), "tools",
                                                                               "nufft_gpu_check.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    rng = np.random.default_rng(2)
    N, F, k0, w = 700, 400, 1, 8
    t = np.sort(rng.uniform(0, 30, N))
    t -= t[0]
    df = 1.0 / (5.0 * t[-1])
    y0 = (3e-3 * rng.normal(size=N)).astype(np.float32)
    y1 = rng.normal(size=N).astype(np.float32)
    p = 4
    while (1 << p) < 4 * (k0 + F):
        p += 1
    M = 1 << p
    cad = np.zeros(2 * N, np.int32)
    fge = np.zeros(M + 2 * w + 4, np.int32)
    spec_buf = np.zeros((M, 2), np.float32)
    dec = np.zeros((F, 2), np.float32)
    assert harness.harness_stages(t.ctypes.data, N, y0.ctypes.data, y1.ctypes.data, df, k0, F, w, cad.ctypes.data,
                                  fge.ctypes.data, spec_buf.ctypes.data, dec.ctypes.data) == p
    C0, S0, _, _ = trig_sums(harness, t, y0, y1, df, k0, F, w)
    stages = tool.check_stages(cad, fge, dec, spec_buf, t, y0, df, k0, F, w, (C0, S0))
    assert [ok for _, ok, _ in stages] == [True, True, True], stages
    bad = spec_buf.copy()
    bad[k0 + 5] *= 1.01
    stages = tool.check_stages(cad, fge, dec, bad, t, y0, df, k0, F, w, (C0, S0))
    assert [ok for _, ok, _ in stages] == [True, True, False]


def test_twiddle_product_tree_variant(harness):
    """LKB_NUFFT_TWIDDLE_CHAIN=1 (one sincospif per butterfly, the other twiddles by a product tree): the transform
    error grows ~2.5x and the whole pipeline stays far inside the parity tolerance."""
    rng = np.random.default_rng(31)
    try:
        harness.harness_set_chain(1)
        for p in (9, 14):
            M = 1 << p
            z = (rng.normal(size=M) + 1j * rng.normal(size=M)).astype(np.complex64)
            out = np.zeros(M, np.complex64)
            harness.harness_fft(z.ctypes.data, p, out.ctypes.data)
            ref = np.fft.ifft(z.astype(np.complex128)) * M
            assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()
        N, F, k0 = 3000, 4000, 1
        t = np.sort(rng.uniform(0, 60.0, N))
        t -= t[0]
        df = 1.0 / (5.0 * t[-1])
        y = (3e-4 * rng.normal(size=N)).astype(np.float32)
        C, S, _, _ = trig_sums(harness, t, y, None, df, k0, F)
        ph = 2 * np.pi * np.outer((k0 + np.arange(F)) * df, t)
        yd = y.astype(np.float64)
        err = max(np.abs(C - np.cos(ph) @ yd).max(), np.abs(S - np.sin(ph) @ yd).max())
        assert err <= 3e-7 * np.abs(yd).sum()
    finally:
        harness.harness_set_chain(0)


def test_grid_size_and_kernel_rules(harness):
    """The fine-grid / kernel-parameter rules (nufft_core.h) the library uses, and the copy of the grid-size rule in
    bench.py's byte model: smallest power of two with an upsampling factor >= sigma_min over the highest mode; width
    10 and beta = 2.30 w at sigma = 2 (the validated setting), wider kernels below."""
    import ctypes
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    import bench
    harness.harness_fine_grid_log2.argtypes = [ctypes.c_int64, ctypes.c_double]
    harness.harness_grid_sigma.argtypes = [ctypes.c_int, ctypes.c_int64]
    harness.harness_grid_sigma.restype = ctypes.c_double
    harness.harness_es_width.argtypes = [ctypes.c_double]
    harness.harness_es_beta.argtypes = [ctypes.c_int, ctypes.c_double]
    harness.harness_es_beta.restype = ctypes.c_double
    os.environ.pop("LKB_NUFFT_SIGMA", None)
    for kmax in (17, 261, 3501, 7002, 14001, 100001, 100002, 131072, 131073, 5000000):
        p = harness.harness_fine_grid_log2(kmax, 2.0)
        assert (1 << p) >= 4 * kmax and ((1 << (p - 1)) < 4 * kmax or p == 4)
        assert p == bench._nufft_fine_log2(kmax)                       # the byte model sizes the same grid
        assert 2.0 <= harness.harness_grid_sigma(p, kmax) <= 4.0
        p_low = harness.harness_fine_grid_log2(kmax, 1.25)
        assert p_low in (p, p - 1) and (1 << p_low) >= 2.5 * kmax
    assert harness.harness_fine_grid_log2(100001, 2.0) == 19 and harness.harness_fine_grid_log2(100001, 1.25) == 18
    assert harness.harness_es_width(2.0) == 10 and harness.harness_es_width(1.31) == 14 and harness.harness_es_width(1.25) == 16
    assert abs(harness.harness_es_beta(10, 2.0) - 23.0) < 0.01        # 2.30 w
    assert harness.harness_es_beta(14, 1.31) < harness.harness_es_beta(14, 2.0)
