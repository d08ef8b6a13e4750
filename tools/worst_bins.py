#!/usr/bin/env python
"""Which Lomb-Scargle kernel family is how far from the fp64 oracle at the FULL config-2 size?

Runs nufft / tcgen05 (several split-K segment lengths) / simt on 1024 x 65 000 x 1e5, takes the (light curve, bin)
pairs where they disagree most plus random pairs, evaluates the oracle there and prints the worst excess over the
tolerance per family.  GPU box only:  python tools/worst_bins.py [--quick]"""
import json
import os
import sys
import time

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
    if "--quick" in sys.argv:
        Y = Y[:128]
    algos = ["nufft", "tcgen05", "tcgen05:seg32", "tcgen05:seg16", "simt"]
    env = {"tcgen05:seg32": {"LKB_TC_SEG_STAGES": "32"}, "tcgen05:seg16": {"LKB_TC_SEG_STAGES": "16"}}
    t0 = time.time()
    worst, (bb, kk), excess = tf.worst_bin_excess(engine, t, Y, freq, algos, env=env)
    out = {"worst_excess": worst, "pairs": int(len(bb)), "seconds": time.time() - t0,
           "note": "tol = stated tolerance 1e-5 max(P) + 1e-4 P; tol_data = floor relative to sqrt(2) std(y) as well",
           "p99_excess": {a: float(np.quantile(e, 0.99)) for a, e in excess.items()},
           "median_excess": {a: float(np.median(e)) for a, e in excess.items()}}
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "worst_bins.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
