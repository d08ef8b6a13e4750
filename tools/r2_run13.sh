#!/usr/bin/env bash
# Round-2 GPU run 13: rows kernel with the light curve as the fast block index; sanitizers with a forced escalation
set -u
O=gpurun_out/r2_run13
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. headline ==="
timeout 600 python bench.py --no-secondary --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python - $O/bench.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("steps %d ms/step %.3f kernel_ms %.3f e2e ms %.2f frac %.3f escalated %s" % (d["steps"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["config"].get("escalated_per_step")))
except Exception as e:
    print("no bench line:", e)
PY
echo "=== 2. LS GPU tests ==="
timeout 1500 python -m pytest tests -m gpu -q -rxXs -k "ls or nufft or lomb or config2 or shared or ragged or periodogram" > $O/pytest_ls.log 2>&1; echo "rc=$?"
tail -4 $O/pytest_ls.log
echo "=== 3. sanitizers ==="
bash tools/sanitize_gpu.sh 2>&1 | tail -14
echo "=== 4. launch list ==="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "rc=$?"
python - $O/launches.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); hdr = None
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d["Metric Name"] == "gpu__time_duration.sum" and "1024" in d.get("Grid Size", "") and ("cols_kernel" in d["Kernel Name"] or "rows_kernel" in d["Kernel Name"]):
            print(d["Kernel Name"][:50], d["Grid Size"], d["Metric Value"], d["Metric Unit"])
PY
echo "=== done ==="
