#!/usr/bin/env bash
# Round-2 GPU run 21: one escalation round per batch; LS tests + headline + launch list
set -u
O=gpurun_out/r2_run21
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -rxXs -k "ls or nufft or lomb or config2 or config5 or shared or ragged or periodogram" > $O/pytest_ls.log 2>&1; echo "rc=$?"
tail -3 $O/pytest_ls.log
timeout 600 python bench.py --no-secondary --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python - $O/bench.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("ms/step %.3f value %.4g e2e ms %.3f (%.4g) frac %.3f launches/step %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"] / d["steps"]))
except Exception as e:
    print("no bench line:", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02_bench_c2_final.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "rc=$?"
echo "=== done ==="
