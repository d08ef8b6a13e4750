#!/usr/bin/env bash
# Round-2 GPU run 10: sync-free precision escalation; regression leg with DMMA right-hand sides and factor reuse
set -u
O=gpurun_out/r2_run10
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. regression tests + leg ==="
timeout 900 python -m pytest tests -m gpu -q -rxXs -k "regress or config4 or corrector" > $O/pytest_sel.log 2>&1; echo "rc=$?"
tail -4 $O/pytest_sel.log
for v in "" "LKB_REGRESS_RHS_SIMT=1" "LKB_REGRESS_REFACTOR=1"; do
env $v timeout 1200 python bench.py --steps 10 --warmup 3 --legs regress > $O/bench_regress.json 2> $O/bench_regress.err; echo "rc=$? [$v]"
python - $O/bench_regress.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    v = d["secondary"]["regress"]
    print("regress: value %.4g LC/s e2e ms %.1f kernel_ms %.1f parity %s" % (v["value"], v["e2e"]["ms_per_step"], v["roofline"]["kernel_ms"], v.get("parity_on_sample")))
except Exception as e:
    print("no bench line:", e)
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02_regress_c.csv python tools/probe_others.py 0.125 regress > $O/ncu_regress.log 2>&1
python - $O/launches_r02_regress_c.csv <<'PY'
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
echo "=== 2. LS: worst bins + headline ==="
timeout 900 python tools/worst_bins.py > $O/worst_bins.log 2>&1; echo "rc=$?"; grep -A3 '"nufft"' $O/worst_bins.log | head -4
for e in 250 0; do
LKB_NUFFT_ESCALATE=$e timeout 400 python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu-baseline > $O/bench_esc$e.json 2> $O/bench_esc$e.err
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
tail -8 $O/pytest_gpu.log
echo "=== 4. sanitizers (NUFFT with escalation) ==="
bash tools/sanitize_gpu.sh 2>&1 | tail -12
echo "=== done ==="
