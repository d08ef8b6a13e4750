#!/usr/bin/env bash
# Round-2 GPU run 17: final single-GPU record - GPU suite, default bench line, reference arm, launch list
set -u
O=gpurun_out/r2_run17
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. full GPU suite ==="
timeout 1500 python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu.log 2>&1; echo "rc=$?"
tail -7 $O/pytest_gpu.log
echo "=== 2. smoke ==="
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== 3. default bench ==="
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"
python - $O/bench_full.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("LS c2: steps %d ms/step %.3f value %.4g e2e ms %.2f (%.4g) frac %.3f escalated %s launches %s clocks %s" % (d["steps"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["config"].get("escalated_per_step"), d.get("gpu_launches"), d.get("clocks")))
    print("cpu_baseline", d.get("cpu_baseline"))
    for k, v in d["secondary"].items():
        print("%-10s value %.4g ms %.2f e2e ms %.1f frac %.3f cpu %.4g parity %s" % (k, v["value"], v["ms_per_step"], v["e2e"]["ms_per_step"], v["roofline"]["frac"], v.get("cpu_baseline", {}).get("value", float("nan")), v.get("parity_on_sample")))
except Exception as e:
    print("no bench line:", e)
PY
echo "=== 4. reference arm ==="
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?"
tail -c 500 $O/bench_reference.json
echo "=== 5. launch list ==="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02_bench_c2_final.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "rc=$?"
echo "=== done ==="
