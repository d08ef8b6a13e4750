#!/usr/bin/env bash
# Round-2 GPU run 14/15: flatten with fused statistics, cut candidates, batched loads, carried dt bracket; sanitizers
set -u
O=gpurun_out/r2_run15
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. flatten tests + leg ==="
timeout 900 python -m pytest tests -m gpu -q -rxXs -k "flatten or config4" > $O/pytest_sel.log 2>&1; echo "rc=$?"
tail -3 $O/pytest_sel.log
timeout 1200 python bench.py --steps 10 --warmup 3 --legs flatten --no-cpu-baseline > $O/bench_flatten.json 2> $O/bench_flatten.err; echo "rc=$?"
python - $O/bench_flatten.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    v = d["secondary"]["flatten"]
    print("flatten: value %.4g LC/s ms %.2f e2e ms %.1f frac %.3f" % (v["value"], v["ms_per_step"], v["e2e"]["ms_per_step"], v["roofline"]["frac"]))
    print("LS: ms/step %.3f e2e %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("no bench line:", e)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:flatten2 -c 1 -o $O/r02_flatten2_f python tools/probe_others.py 0.03 flatten > $O/ncu_flatten.log 2>&1
echo "=== 2. sanitizers ==="
bash tools/sanitize_gpu.sh 2>&1 | tail -14
echo "=== done ==="
