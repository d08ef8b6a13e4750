#!/usr/bin/env bash
# Round-2 GPU run 9: kernel width 10 by default + precision escalation (double-precision pass for flagged light curves)
set -u
O=gpurun_out/r2_run9
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. worst bins (default settings) ==="
timeout 900 python tools/worst_bins.py > $O/worst_bins.log 2>&1; echo "rc=$?"; grep -A3 '"nufft"' $O/worst_bins.log | head -8
echo "=== 2. bench headline, with and without escalation ==="
for e in 250 0; do
LKB_NUFFT_ESCALATE=$e timeout 400 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $O/bench_esc$e.json 2> $O/bench_esc$e.err
python - $O/bench_esc$e.json $e <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("escalate=%s: ms/step %.3f kernel_ms %.3f e2e ms %.2f frac %.3f escalated %s" % (sys.argv[2], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["config"].get("escalated_per_step")))
except Exception as e:
    print("no bench line:", e)
PY
done
echo "=== 3. full GPU suite ==="
timeout 1500 python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu.log 2>&1; echo "rc=$?"
tail -12 $O/pytest_gpu.log
echo "=== 4. launch list of one step ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02_c2_escalation.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $O/ncu_bench.log 2>&1
python - $O/launches_r02_c2_escalation.csv <<'PY'
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d["Metric Name"] == "gpu__time_duration.sum":
            k = d["Kernel Name"][:60]
            v = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(d["Metric Unit"], 1e-6)
            a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, ms) in agg.items(): print("%-62s x%-4d %.3f ms" % (k, n, ms))
PY
echo "=== done ==="
