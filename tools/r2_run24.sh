#!/usr/bin/env bash
# Round-2 GPU run 24: the final default bench line and reference arm (record for profiles/)
set -u
O=gpurun_out/r2_run24
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?"
python - $O/bench_full.json $O/bench_reference.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print("LS c2: ms/step %.3f value %.4g e2e ms %.2f (%.4g) frac %.3f escalated %s launches %s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["config"].get("escalated_per_step"), d.get("gpu_launches")))
print("cpu_baseline", d.get("cpu_baseline"))
for k, v in d["secondary"].items():
    print("%-10s value %.4g ms %.2f e2e ms %.1f frac %.3f parity %s" % (k, v["value"], v["ms_per_step"], v["e2e"]["ms_per_step"], v["roofline"]["frac"], v.get("parity_on_sample")))
print("reference arm: %.4g on %s cores; e2e / reference = %.0f" % (r["value"], r["cpu_baseline"]["cores"], d["e2e"]["value"] / r["value"]))
PY
echo "=== done ==="
