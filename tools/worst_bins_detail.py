#!/usr/bin/env python
"""Where exactly are the Lomb-Scargle kernel families furthest from the fp64 oracle at full config-2 size?
Prints the 40 worst (light curve, bin) pairs with every family's value; GPU box only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from lightkurve_b200 import engine
    from bench import make_workload
    import test_gpu_fullsize as tf
    engine.init(0)
    t, Y, freq = make_workload("c2", 1002)
    algos = ["nufft", "simt", "tcgen05"]
    outs = {a: engine.ls_power_shared(t, Y, freq, "amplitude", algo=a) for a in algos}
    B, F = outs["nufft"].shape
    T = t[-1] - t[0]
    pmax = outs["nufft"].max(axis=1)
    tol = 1e-5 * pmax[:, None] + 1e-4 * outs["nufft"]
    d = np.abs(outs["simt"] - outs["nufft"]) / tol
    flat = np.argpartition(d.ravel(), -3000)[-3000:]
    bb, kk = np.unravel_index(flat, (B, F))
    ref = np.empty(len(flat))
    for b in np.unique(bb):
        sel = bb == b
        ref[sel] = tf._oracle_amplitude(t, Y[b], freq[kk[sel]])
    tol_ref = 1e-5 * np.maximum(pmax[bb], ref) + 1e-4 * ref
    ex = {a: np.abs(outs[a][bb, kk] - ref) / tol_ref for a in algos}
    # candidates for NUFFT misses as well: where it differs most from tcgen05, and the bins just above the low rows
    d2 = np.abs(outs["tcgen05"] - outs["nufft"]) / tol
    flat2 = np.argpartition(d2.ravel(), -1500)[-1500:]
    b2, k2 = np.unravel_index(flat2, (B, F))
    lowb = np.repeat(np.arange(B), 8)
    lowk = np.tile(np.arange(8, 16), B)
    bb2, kk2 = np.concatenate([b2, lowb]), np.concatenate([k2, lowk])
    ref2 = np.empty(len(bb2))
    for b in np.unique(bb2):
        sel = bb2 == b
        ref2[sel] = tf._oracle_amplitude(t, Y[b], freq[kk2[sel]])
    tol2 = 1e-5 * np.maximum(pmax[bb2], ref2) + 1e-4 * ref2
    exn = np.abs(outs["nufft"][bb2, kk2] - ref2) / tol2
    print("NUFFT: its 25 worst pairs among %d candidates (b, k, f*T, oracle, nufft, pmax_b, excess, |k - k_peak|, std(y))" % len(bb2))
    kpk0 = outs["nufft"].argmax(axis=1)
    for i in np.argsort(-exn)[:25]:
        b, k = int(bb2[i]), int(kk2[i])
        print("%5d %7d %9.2f %12.5e %12.5e %10.3e %6.2f %7d %10.3e" % (b, k, freq[k] * T, ref2[i], outs["nufft"][b, k], pmax[b],
                                                                  exn[i], abs(k - int(kpk0[b])), Y[b].astype(np.float64).std()))
    order = np.argsort(-np.maximum(ex["nufft"], ex["simt"]))[:40]
    print("baseline T = %.3f d, df*T = %.3f; rows with f*T <= 2 are 'low rows'" % (T, freq[0] * T))
    print("%5s %7s %8s %12s %12s %12s %12s %10s %10s | excess nufft simt tc | partner pmax" %
          ("b", "k", "f*T", "oracle", "nufft", "simt", "tcgen05", "pmax_b", "tol"))
    for i in order:
        b, k = int(bb[i]), int(kk[i])
        print("%5d %7d %8.2f %12.5e %12.5e %12.5e %12.5e %10.3e %10.3e | %6.2f %6.2f %6.2f | %.3e" % (
            b, k, freq[k] * T, ref[i], outs["nufft"][b, k], outs["simt"][b, k], outs["tcgen05"][b, k], pmax[b], tol_ref[i],
            ex["nufft"][i], ex["simt"][i], ex["tcgen05"][i], pmax[b ^ 1]))
    low = kk < 12
    print("pairs in low rows (k < 12): %d of %d; worst excess outside the low rows: nufft %.2f simt %.2f tcgen05 %.2f" % (
        int(low.sum()), len(kk), ex["nufft"][~low].max(), ex["simt"][~low].max(), ex["tcgen05"][~low].max()))
    # histogram of the worst bins' positions relative to the light curve's own peak
    kpk = outs["nufft"].argmax(axis=1)
    print("distance |k - k_peak| of the 40 worst:", [int(abs(int(kk[i]) - int(kpk[bb[i]]))) for i in order])
    # permutation sensitivity (pair partners): same light curves, partner = a quiet one vs a loud one
    quiet = np.argsort(pmax)[:2]
    loud = np.argsort(pmax)[-2:]
    probe = np.ascontiguousarray(np.stack([Y[quiet[0]], Y[quiet[1]], Y[quiet[0]], Y[loud[1]], Y[loud[0]], Y[loud[1]]]))
    o = engine.ls_power_shared(t, probe, freq, "amplitude", algo="nufft")
    tq = 1e-5 * o[0].max() + 1e-4 * o[0]
    dq = np.abs(o[2] - o[0]) / tq
    print("quiet LC (pmax %.3e) paired with a quiet / a loud (pmax %.3e) partner: worst change %.2f tol at k = %d (f*T %.2f)"
          % (pmax[quiet[0]], pmax[loud[1]], dq.max(), int(dq.argmax()), freq[int(dq.argmax())] * T))
    refq = tf._oracle_amplitude(t, Y[quiet[0]], freq[:2000])
    for name, row in (("quiet partner", o[0]), ("loud partner", o[2])):
        e = np.abs(row[:2000] - refq) / (1e-5 * max(row.max(), refq.max()) + 1e-4 * refq)
        print("  %s: worst excess vs oracle over the first 2000 bins %.2f at k = %d" % (name, e.max(), int(e.argmax())))


if __name__ == "__main__":
    main()
