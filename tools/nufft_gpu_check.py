"""Round-2 bring-up aid for the NUFFT Lomb-Scargle path: runs a small shared-grid call with algo="nufft" on the
GPU, reads the intermediate buffers back (lkb_ws_read) and compares them STAGE BY STAGE with the CPU harness
(tests/native/nufft_host_harness.cpp, the same __host__ __device__ code compiled with g++) and with the fp64
oracle, so that a defect in the CUDA glue is localised to one kernel in one run.

    python tools/nufft_gpu_check.py            # on the GPU box: prints one line per stage and a verdict

Stages: cadence table (slot A) -> first_ge table (B) -> deconvolution factors (C) -> spread grid / FFT output
(H, I: whichever holds the last pass) -> rotation terms (F) -> power vs oracle.
Workspace slots as used by ls_nufft_launch in lightkurve_b200/csrc/ls_nufft.cu."""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_harness():
    out = os.path.join(tempfile.mkdtemp(), "libnufft_harness.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "native", "nufft_host_harness.cpp")])
    lib = ctypes.CDLL(out)
    c_vp, c_i64 = ctypes.c_void_p, ctypes.c_int64
    lib.harness_trig_sums.argtypes = [c_vp, c_i64, c_vp, c_vp, ctypes.c_double, c_i64, c_i64, ctypes.c_int,
                                      c_vp, c_vp, c_vp, c_vp]
    lib.harness_fft.argtypes = [c_vp, ctypes.c_int, c_vp]
    return lib


def check_stages(cad, fge, dec, Z0, trel, yc0, df, k0, F, w, sums0):
    """Compare the intermediate buffers of one call (cad [2N] int32 view of {i0, d0}, first_ge, dec [F, 2],
    Z0 [M, 2] = final transform of pair 0) with what they must be; `sums0` = (C, S) of light curve 0 from the CPU
    harness.  Returns [(stage, ok, detail)]."""
    N = len(trel)
    p = 4
    while (1 << p) < 4 * (k0 + F):
        p += 1
    M = 1 << p
    shift = (w + 1) // 2 + 1
    out = []
    i0 = np.asarray(cad[0::2], dtype=np.int64)
    d0 = np.asarray(cad[1::2]).view(np.float32)
    x = df * trel * M + shift
    i0_ref = np.ceil(x - 0.5 * w).astype(np.int64)
    out.append(("cadence table", np.array_equal(i0, i0_ref) and np.allclose(d0, i0_ref - x, atol=1e-6),
                "max |d0 err| %.2e" % np.abs(d0 - (i0_ref - x)).max()))
    L = M + 2 * w + 4
    out.append(("first_ge table", np.array_equal(fge[:L], np.searchsorted(i0_ref, np.arange(L), side="left")), ""))
    z = Z0[:, 0].astype(np.float64) + 1j * Z0[:, 1].astype(np.float64)
    kk = k0 + np.arange(F)
    g1, g2 = z[kk], z[(M - kk) % M]
    mx = float(np.abs(yc0).max())
    s0 = 1.0 if mx == 0 else 2.0 ** -(np.frexp(np.float32(mx))[1])          # nufft::pow2_scale
    a = 0.5 * (g1 + np.conj(g2)) * (dec[:, 0] + 1j * dec[:, 1]) / s0
    ref = sums0[0] + 1j * sums0[1]
    err = np.abs(a - ref).max() / max(np.abs(yc0).sum(), 1e-30)
    out.append(("spread + FFT + deconvolution", err < 3e-6, "pair 0, LC 0: max err / |y|_1 = %.2e" % err))
    return out


def main():
    from lightkurve_b200 import engine
    from oracle import ls as ols
    engine.init(0)
    lib = build_harness()
    rng = np.random.default_rng(2)
    w = 8
    idx = np.flatnonzero(rng.uniform(size=2600) > 0.1)[:2000]
    t = 131.5 + idx * 0.0204336
    N = len(t)
    trel = t - t[0]
    df = 1.0 / (5.0 * trel[-1])
    F, k0 = 1200, 1
    freq = df * (k0 + np.arange(F))
    Y = np.stack([1 + 1e-2 * np.sin(2 * np.pi * 3.1 * t) + 1e-4 * rng.normal(size=N),
                  1 + 3e-4 * rng.normal(size=N), 1 + 1e-3 * rng.normal(size=N)]).astype(np.float32)
    B = len(Y)
    got = np.asarray(engine.ls_power_shared(t, Y, freq, "amplitude", algo="nufft"), dtype=np.float64)
    p = 4
    while (1 << p) < 4 * (k0 + F):
        p += 1
    M = 1 << p
    yc = (Y.astype(np.float64) - Y.astype(np.float64).mean(axis=1, keepdims=True)).astype(np.float32)
    C0, S0, C1, S1 = (np.zeros(F, np.float32) for _ in range(4))
    lib.harness_trig_sums(trel.ctypes.data, N, yc[0].ctypes.data, yc[1].ctypes.data, df, k0, F, w, C0.ctypes.data,
                          S0.ctypes.data, C1.ctypes.data, S1.ctypes.data)
    npass = (p + 3) // 4
    cad = engine.ws_read("A", 2 * N, np.int32)
    fge = engine.ws_read("B", M + 2 * w + 4, np.int32)
    dec = engine.ws_read("C", 2 * F, np.float32).reshape(F, 2)
    Z = engine.ws_read("H" if npass % 2 == 0 else "I", 2 * M * ((B + 1) // 2), np.float32).reshape(-1, M, 2)
    stages = check_stages(cad, fge, dec, Z[0], trel, yc[0], df, k0, F, w, (C0.astype(np.float64), S0.astype(np.float64)))
    worst = 0.0
    for b in range(B):
        ref = np.sqrt(ols.ls_slow_psd(t, Y[b].astype(np.float64), freq)) * np.sqrt(4.0 / N)
        worst = max(worst, float(np.max(np.abs(got[b] - ref) / (1e-5 * ref.max() + 1e-4 * ref))))
    stages.append(("power vs fp64 oracle", worst < 1.0, "worst tolerance excess %.3f" % worst))
    for stage, ok, detail in stages:
        print("%-30s %s  %s" % (stage, "ok  " if ok else "FAIL", detail))
    ok_all = all(ok for _, ok, _ in stages)
    print("VERDICT:", "NUFFT path verified on this GPU" if ok_all else "see the first FAIL above")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
