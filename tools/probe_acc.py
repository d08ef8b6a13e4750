"""GPU probe: accuracy of the tcgen05 path vs accumulation-chain length (LKB_TC_SEG_STAGES)."""
import os
import sys
import numpy as np
sys.path.insert(0, ".")
from lightkurve_b200 import engine
from oracle import ls as ols
engine.init(0)
rng = np.random.default_rng(23)
B, N, F = 70, 65000, 200
keep = np.sort(rng.choice(71500, N, replace=False))
t = 131.5 + keep * 0.0204336
Y = (1 + 10 ** rng.uniform(-4, -2, (B, 1)) * np.sin(2 * np.pi * rng.uniform(0.05, 13, (B, 1)) * t[None, :])
     + 10 ** rng.uniform(-4.3, -3, (B, 1)) * rng.normal(size=(B, N))).astype(np.float32)
freq = np.sort(rng.uniform(0.01, 13.6, F))
refs = {b: np.sqrt(ols.ls_slow_psd(t, Y[b].astype(np.float64), freq)) * np.sqrt(4.0 / N) for b in (0, 33, 69)}
sim = engine.ls_power_shared(t, Y, freq, "amplitude", algo="simt")
modes = [(int(a), 0) for a in os.environ.get("PROBE_SEGS", "100000,512,256,128,64,32").split(",")]
for seg, fp8 in modes:
    os.environ["LKB_TC_SEG_STAGES"] = str(seg)
    os.environ["LKB_TC_FP8LO"] = str(fp8)
    out = engine.ls_power_shared(t, Y, freq, "amplitude", algo="tcgen05")
    row = []
    for b, ref in refs.items():
        e = out[b] - ref
        row.append("lc%d mean rel %.2e max|rel| %.2e excess %.2f" % (b, np.mean(e / ref), np.max(np.abs(e / ref)),
                   np.max(np.abs(e) / (1e-5 * ref.max() + 1e-4 * ref))))
    print("seg_stages %6d%s: %s" % (seg, " fp8lo" if fp8 else "      ", " | ".join(row)))
os.environ["LKB_TC_FP8LO"] = "0"
row = []
for b, ref in refs.items():
    e = sim[b] - ref
    row.append("lc%d mean rel %.2e max|rel| %.2e excess %.2f" % (b, np.mean(e / ref), np.max(np.abs(e / ref)),
               np.max(np.abs(e) / (1e-5 * ref.max() + 1e-4 * ref))))
print("simt             : %s" % " | ".join(row))
