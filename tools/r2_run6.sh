#!/usr/bin/env bash
# Round-2 GPU run 6: fp64 NUFFT weights, refined tcgen05 regression, flatten v2 with unordered cuts.
set -u
O=gpurun_out/r2_run6
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. worst bins (fp64 weights) ==="
timeout 900 python tools/worst_bins.py > $O/worst_bins.log 2>&1; echo "rc=$?"; grep -A3 '"nufft"' $O/worst_bins.log | head -12
echo "=== 2. full GPU suite ==="
timeout 1800 python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu.log 2>&1; echo "rc=$?"
tail -12 $O/pytest_gpu.log; grep -E "^E  |Error" $O/pytest_gpu.log | head -20
echo "=== 3. bench with all legs + launch list of the regression leg ==="
timeout 1500 python bench.py --steps 10 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"
python - $O/bench_full.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("LS c2: ms/step %.3f kernel_ms %.3f e2e %.3f ms frac %.3f value %.4g" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["value"]))
    for k, v in d["secondary"].items():
        if "error" in v: print(k, "ERROR", v["error"]); continue
        print(k, "value %.4g %s  ms %.3f  e2e ms %.3f  roofline %.4g %s frac %.3f  kernel_ms %s cpu %s  parity %s" % (v["value"], v["unit"], v["ms_per_step"], v["e2e"]["ms_per_step"], v["roofline"]["achieved"], v["roofline"]["unit"], v["roofline"]["frac"], v["roofline"].get("kernel_ms"), v.get("cpu_baseline", {}).get("value"), v.get("parity_on_sample")))
except Exception as e:
    print("no bench line:", e)
PY
tail -3 $O/bench_full.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02_regress.csv python tools/probe_others.py 0.125 regress > $O/ncu_regress.log 2>&1
python - $O/launches_r02_regress.csv <<'PY'
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
echo "=== 4. ncu flatten ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flatten2_kernel -c 1 -o $O/r02_flatten2_c python tools/probe_others.py 0.03 flatten > $O/ncu_flatten.log 2>&1
tail -2 $O/ncu_flatten.log
ls -la $O
echo "=== done ==="
