#!/usr/bin/env bash
# Round-2 GPU run 8: flatten v2 with broadcast selects / retuned sampling median; NUFFT kernel width 10/12 worst-bin detail
set -u
O=gpurun_out/r2_run8
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. flatten tests + leg ==="
timeout 900 python -m pytest tests -m gpu -q -rxXs -k "flatten or config4" > $O/pytest_sel.log 2>&1; echo "rc=$?"
tail -4 $O/pytest_sel.log
timeout 1200 python bench.py --steps 10 --warmup 3 --legs flatten > $O/bench_flatten.json 2> $O/bench_flatten.err; echo "rc=$?"
python - $O/bench_flatten.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    v = d["secondary"]["flatten"]
    print("flatten: value %.4g LC/s ms %.2f e2e ms %.1f frac %.3f parity %s" % (v["value"], v["ms_per_step"], v["e2e"]["ms_per_step"], v["roofline"]["frac"], v.get("parity_on_sample")))
except Exception as e:
    print("no bench line:", e)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:flatten2 -c 1 -o $O/r02_flatten2_d python tools/probe_others.py 0.03 flatten > $O/ncu_flatten.log 2>&1
echo "=== 2. worst-bin detail at kernel width 10 and 12 ==="
LKB_NUFFT_W=10 timeout 900 python tools/worst_bins_detail.py > $O/worst_bins_detail_w10.log 2>&1; echo "rc=$?"; head -30 $O/worst_bins_detail_w10.log
LKB_NUFFT_W=12 timeout 900 python tools/worst_bins.py > $O/worst_bins_w12.log 2>&1; echo "rc=$?"; grep -A3 '"nufft"' $O/worst_bins_w12.log | head -8
echo "=== done ==="
