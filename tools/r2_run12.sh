#!/usr/bin/env bash
# Round-2 GPU run 12: after the register-spill fix of the transform kernels - headline, LS tests, sanitizers, ncu captures
set -u
O=gpurun_out/r2_run12
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. headline ==="
for e in 250 0; do
LKB_NUFFT_ESCALATE=$e timeout 600 python bench.py --no-secondary --no-cpu-baseline > $O/bench_esc$e.json 2> $O/bench_esc$e.err
python - $O/bench_esc$e.json $e <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("escalate=%s: steps %d ms/step %.3f kernel_ms %.3f e2e ms %.2f frac %.3f escalated %s clocks %s" % (sys.argv[2], d["steps"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["config"].get("escalated_per_step"), d.get("clocks")))
except Exception as e:
    print("no bench line:", e)
PY
done
echo "=== 2. full GPU suite ==="
timeout 1500 python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu.log 2>&1; echo "rc=$?"
tail -6 $O/pytest_gpu.log
echo "=== 3. sanitizers ==="
bash tools/sanitize_gpu.sh 2>&1 | tail -14
echo "=== 4. ncu --set full: transform kernels (float2 batch launch, then the double2 escalation launch) ==="
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"nufft2_cols_kernel|nufft2_rows_kernel|nufft2_cols_list_kernel|nufft2_rows_list_kernel" -c 4 -o $O/r02_nufft_final python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $O/ncu_nufft.log 2>&1; echo "rc=$?"
echo "=== 5. launch list of the bench command ==="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02_bench_final.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "rc=$?"
echo "=== 6. full default bench ==="
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"
python - $O/bench_full.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("LS c2: ms/step %.3f value %.4g e2e ms %.2f frac %.3f escalated %s launches %s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["config"].get("escalated_per_step"), d.get("gpu_launches")))
    for k, v in d["secondary"].items():
        print("%-10s value %.4g ms %.2f e2e ms %.1f frac %.3f" % (k, v["value"], v["ms_per_step"], v["e2e"]["ms_per_step"], v["roofline"]["frac"]))
except Exception as e:
    print("no bench line:", e)
PY
echo "=== done ==="
