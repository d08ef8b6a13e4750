#!/usr/bin/env bash
# Round-2 GPU run 23: regression leg variance check (two fresh processes, and two calls inside one process)
set -u
O=gpurun_out/r2_run23
mkdir -p $O
export PYTHONUNBUFFERED=1
for i in 1 2; do
timeout 900 python bench.py --steps 5 --warmup 3 --legs regress --no-cpu-baseline > $O/bench_regress$i.json 2> $O/bench_regress$i.err
python - $O/bench_regress$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    v = d["secondary"]["regress"]
    print("regress: device ms %.1f e2e ms %.1f gram ms %.1f" % (v["ms_per_step"], v["e2e"]["ms_per_step"], v["roofline"]["kernel_ms"]))
except Exception as e:
    print("no bench line:", e)
PY
done
timeout 900 python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, ".")
from lightkurve_b200 import engine
import bench
engine.init(0)
N, K, B = 65000, 151, 4096
tt, X, Y, FE = bench.make_c4_workload(1004, B, N, K)
for i in range(3):
    engine.profile_enable(True)
    t0 = time.perf_counter()
    engine.regress(X, Y, FE, None, np.zeros(K), np.full(K, np.inf), sigma=5, niters=5)
    wall = time.perf_counter() - t0
    kms = engine.profile_read()
    engine.profile_enable(False)
    print("call %d: gram %.1f ms, device %.1f ms, wall %.1f ms" % (i, kms[0], sum(kms), 1e3 * wall))
PY
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
echo "=== done ==="
