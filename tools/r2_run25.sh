#!/usr/bin/env bash
# Round-2 GPU run 25: ragged NUFFT with FP64-evaluated kernel weights - cost (A/B) and the ragged parity tests
set -u
O=gpurun_out/r2_run25
mkdir -p $O
export PYTHONUNBUFFERED=1
for v in "" "LKB_NUFFT_RAGGED_W32=1"; do
env $v timeout 900 python bench.py --steps 10 --warmup 3 --legs ls_ragged --no-cpu-baseline > $O/bench_ragged.json 2> $O/bench_ragged.err
python - $O/bench_ragged.json "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    v = d["secondary"]["ls_ragged"]
    print("[%s] ls_ragged: ms %.2f value %.4g e2e ms %.2f" % (sys.argv[2], v["ms_per_step"], v["value"], v["e2e"]["ms_per_step"]))
except Exception as e:
    print("no bench line:", e)
PY
done
timeout 1200 python -m pytest tests -m gpu -q -rxXs -s -k "ragged or config5 or collection or nufft" > $O/pytest.log 2>&1; echo "rc=$?"
grep -E "worst-bin excess|passed|failed" $O/pytest.log | head
echo "=== done ==="
