#!/usr/bin/env bash
# Device sanitizers on the GPU box: memcheck over smoke() (every kernel family once) and
# racecheck over the shared-memory-heavy kernels (NUFFT v2 tiles, flatten v2).  Summaries -> gpurun_out/sanitize/.
set -u
O=gpurun_out/sanitize
mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
export PYTHONUNBUFFERED=1
echo "=== memcheck: smoke() ==="
timeout 1500 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/memcheck_smoke.log 2>&1; echo "rc=$?"
grep -E "ERROR SUMMARY|smoke ok|Invalid|========= Error" $O/memcheck_smoke.log | head -10
cat > $O/_small.py <<'PY'
import numpy as np, os, sys
sys.path.insert(0, ".")
os.environ["LKB_NUFFT_ESCALATE"] = "50"       # (the test light curve's flux excursion is 219 x its in-band peak)
from lightkurve_b200 import engine
engine.init(0)
rng = np.random.default_rng(3)
N, B, F = 3000, 9, 3600
keep = np.ones(3300, bool)
for c in (200, 700, 1300, 1900, 2500, 3000):
    keep[c:c + 50] = False                                   # a regular cadence with six gaps
t = 100.0 + 0.02 * np.flatnonzero(keep)
freq = (1 + np.arange(F)) / (5.0 * (t[-1] - t[0]))
Y = (1 + 1e-3 * np.sin(2 * np.pi * 1.1 * t)[None, :] + 3e-4 * rng.normal(size=(B, N))).astype(np.float32)
Y[0] = (1 + 1e-2 * np.sin(2 * np.pi * 18.3 * t) + 1e-5 * rng.normal(size=N)).astype(np.float32)   # loud line above the grid
engine.ls_power_shared(t, Y, freq, "amplitude")
assert engine.ls_last_algo() == "nufft" and engine.ls_last_escalated() == 1, engine.ls_last_escalated()
engine.ls_power_ragged([t[:2000], t[:2500], t], [Y[0, :2000], Y[1, :2500], Y[2]], freq, "amplitude", algo="nufft")
tt = np.arange(0, 120, 0.02)
f = 1 + 0.01 * np.sin(tt / 3.0) + 1e-3 * rng.normal(size=len(tt))
engine.flatten([tt, tt[:3000]], [f, f[:3000]], None, None, window_length=101)
if os.environ.get("LKB_SANITIZE_SKIP_TC"):
    print("small ok")       # racecheck does not model the mbarrier / tcgen05.commit ordering of the TMA + tcgen05 pipelines
    sys.exit(0)             # (it reports every stage re-use as a write-after-write hazard): those kernels run under memcheck only
Nr, Kr, Br = 4096, 24, 64                                    # tcgen05 Gram + DMMA right-hand sides + factor reuse
X = np.hstack([rng.normal(size=(Nr, Kr - 1)), np.ones((Nr, 1))])
Yr = 1 + (rng.normal(size=(Br, Kr)) * 1e-3) @ X.T + 3e-4 * rng.normal(size=(Br, Nr))
engine.regress(X, Yr, np.full((Br, Nr), 3e-4), niters=2, sigma=5)
print("small ok")
PY
echo "=== racecheck: NUFFT v2 + flatten v2 ==="
LKB_SANITIZE_SKIP_TC=1 timeout 1500 $CS --tool racecheck --error-exitcode 9 --print-limit 20 python $O/_small.py > $O/racecheck_small.log 2>&1; echo "rc=$?"
grep -E "RACECHECK SUMMARY|small ok|hazard|========= Error" $O/racecheck_small.log | head -10
echo "=== memcheck: NUFFT v2 + flatten v2 ==="
timeout 1500 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python $O/_small.py > $O/memcheck_small.log 2>&1; echo "rc=$?"
grep -E "ERROR SUMMARY|small ok|Invalid|========= Error" $O/memcheck_small.log | head -10
echo "=== done ==="
