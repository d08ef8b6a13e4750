"""GPU probe: shared-grid Lomb-Scargle, SIMT vs tcgen05 - agreement, accuracy vs the oracle, timing."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from lightkurve_b200 import engine  # noqa: E402
from oracle import ls as ols  # noqa: E402

B, N, F = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (1024, 8192, 4096))]
algos = sys.argv[4].split(",") if len(sys.argv) > 4 else ["simt", "tcgen05"]
engine.init(0)
rng = np.random.default_rng(7)
keep = np.sort(rng.choice(int(N * 1.1), N, replace=False))
t = 131.5 + keep * 0.0204336
amp = 10 ** rng.uniform(-4, -2, (B, 1))
Y = (1 + amp * np.sin(2 * np.pi * rng.uniform(0.05, 20, (B, 1)) * t[None, :] + rng.uniform(0, 6, (B, 1)))
     + 10 ** rng.uniform(-4.3, -3, (B, 1)) * rng.normal(size=(B, N))).astype(np.float32)
freq = (1 + np.arange(F)) / (5 * (t[-1] - t[0]))
dt, dY, df = torch.tensor(t, device="cuda"), torch.tensor(Y, device="cuda"), torch.tensor(freq, device="cuda")
res = {}
for algo in algos:
    out = engine.ls_power_shared(dt, dY, df, "amplitude", algo=algo)
    torch.cuda.synchronize()
    engine.profile_enable(True)
    t0 = time.time()
    for _ in range(3):
        out = engine.ls_power_shared(dt, dY, df, "amplitude", algo=algo)
    torch.cuda.synchronize()
    wall = (time.time() - t0) / 3
    kms = engine.profile_read()
    engine.profile_enable(False)
    res[algo] = out.cpu().numpy()
    units = F * N * B
    print("%-8s wall %.2f ms  kernel %.2f ms  %.3e bin*cad/s  tensor-equiv %.1f TFLOP/s (algorithmic 4FNB)" % (
        algo, wall * 1e3, kms.mean(), units / (kms.mean() * 1e-3), 4 * units / (kms.mean() * 1e-3) / 1e12))
for b in (0, B // 2, B - 1):
    ref = np.sqrt(ols.ls_slow_psd(t, Y[b].astype(np.float64), freq[:256])) * np.sqrt(4.0 / N)
    for algo in algos:
        e = np.abs(res[algo][b, :256] - ref)
        print("lc %d %-8s max|err|/max(P) %.2e   max rel err %.2e   tol-excess %.3f" % (
            b, algo, e.max() / ref.max(), (e / ref).max(), (e / (1e-5 * ref.max() + 1e-4 * ref)).max()))
if len(algos) == 2:
    d = np.abs(res[algos[0]] - res[algos[1]])
    print("simt vs tc: max abs diff / max %.2e" % (d.max() / res[algos[0]].max()))
