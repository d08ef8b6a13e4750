#!/usr/bin/env bash
# Round-2 GPU run 16: regression with the one-pass-per-round sigma clip and two-rows-per-warp LU
set -u
O=gpurun_out/r2_run16
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. regression tests + leg ==="
timeout 900 python -m pytest tests -m gpu -q -rxXs -k "regress or config4 or corrector or nanmedian or select" > $O/pytest_sel.log 2>&1; echo "rc=$?"
tail -4 $O/pytest_sel.log
timeout 1200 python bench.py --steps 10 --warmup 3 --legs regress > $O/bench_regress.json 2> $O/bench_regress.err; echo "rc=$?"
python - $O/bench_regress.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    v = d["secondary"]["regress"]
    print("regress: value %.4g LC/s device ms %.1f e2e ms %.1f gram ms %.1f parity %s" % (v["value"], v["ms_per_step"], v["e2e"]["ms_per_step"], v["roofline"]["kernel_ms"], v.get("parity_on_sample")))
except Exception as e:
    print("no bench line:", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02_regress_d.csv python tools/probe_others.py 0.125 regress > $O/ncu_regress.log 2>&1
python - $O/launches_r02_regress_d.csv <<'PY'
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d["Metric Name"] == "gpu__time_duration.sum":
            k = d["Kernel Name"][:40]
            v = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(d["Metric Unit"], 1e-6)
            a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, ms) in agg.items(): print("%-42s x%-4d %.3f ms" % (k, n, ms))
PY
echo "=== done ==="
