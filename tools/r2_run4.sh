#!/usr/bin/env bash
# Round-2 GPU run 4: K4 v2 (streaming flatten) on hardware, low-row split-K kernel, worst-bin detail for the NUFFT family.
set -u
O=gpurun_out/r2_run4
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. flatten + LS tests ==="
timeout 1500 python -m pytest tests -m gpu -q -rxXs -k "flatten or config4 or ls_shared or nufft or plan_cache or detrend or iterative" > $O/pytest_sel.log 2>&1; echo "rc=$?"
tail -15 $O/pytest_sel.log
echo "=== 2. bench: headline + flatten / ragged legs ==="
timeout 900 python bench.py --steps 10 --warmup 3 --legs flatten,ls_ragged > $O/bench_legs.json 2> $O/bench_legs.err; echo "rc=$?"
python - $O/bench_legs.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("LS c2: ms/step %.3f kernel_ms %.3f e2e %.3f ms frac %.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
    for k, v in d["secondary"].items():
        if "error" in v: print(k, "ERROR", v["error"]); continue
        print(k, "value %.4g %s  ms %.3f  e2e ms %.3f  roofline frac %.3f  cpu %s  parity %s" % (v["value"], v["unit"], v["ms_per_step"], v["e2e"]["ms_per_step"], v["roofline"]["frac"], v.get("cpu_baseline", {}).get("value"), v.get("parity_on_sample")))
except Exception as e:
    print("no bench line:", e)
PY
tail -3 $O/bench_legs.err
echo "=== 3. worst-bin detail ==="
timeout 900 python tools/worst_bins_detail.py > $O/worst_bins_detail.log 2>&1; echo "rc=$?"
head -32 $O/worst_bins_detail.log
echo "=== 4. ncu: flatten v2 (1/32 of config 4) and the NUFFT kernels ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flatten2_kernel -c 1 -o $O/r02_flatten2 python tools/probe_others.py 0.03 flatten > $O/ncu_flatten.log 2>&1
tail -2 $O/ncu_flatten.log
B="python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nufft2_(spread|cols|rows|lowrows)_kernel" -c 4 -o $O/r02_nufft_v2r2 $B > $O/ncu_v2.log 2>&1
tail -2 $O/ncu_v2.log
echo "=== 5. BLS: warp instructions per (light curve, period) over a whole period grid (issue-slot roofline) ==="
timeout 900 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:bls_search_kernel --csv --log-file $O/bls_inst_r02.csv python tools/probe_others.py 0.125 bls > $O/ncu_bls.log 2>&1
tail -2 $O/ncu_bls.log
python - $O/bls_inst_r02.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; inst = 0.0; ns = 0.0; n = 0
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        if d["Metric Name"] == "smsp__inst_executed.sum": inst += v; n += 1
        if d["Metric Name"] == "gpu__time_duration.sum": ns += v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(d["Metric Unit"], 1)
pairs = 2 * 32 * 50000          # probe_others runs the search twice (warm-up + timed)
print("bls_search launches %d  warp-instructions %.4e  per (LC, period) %.1f  kernel time %.2f ms" % (n, inst, inst / pairs, ns / 1e6))
PY
ls -la $O
echo "=== done ==="
