#!/usr/bin/env bash
# Round-2 GPU run 19: full GPU suite after the fixed low-row slices; headline + launch list
set -u
O=gpurun_out/r2_run19
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. full GPU suite ==="
timeout 1500 python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu.log 2>&1; echo "rc=$?"
tail -7 $O/pytest_gpu.log
echo "=== 2. headline ==="
timeout 600 python bench.py --no-secondary --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python - $O/bench.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("ms/step %.3f value %.4g e2e ms %.3f (%.4g) frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print("no bench line:", e)
PY
echo "=== done ==="
