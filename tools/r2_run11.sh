#!/usr/bin/env bash
# Round-2 GPU run 11: final single-GPU evidence - full bench line, sanitizers, ncu captures of the round's new kernels
set -u
O=gpurun_out/r2_run11
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. full default bench (all legs) ==="
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"
python - $O/bench_full.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("LS c2: ms/step %.3f value %.4g e2e ms %.2f frac %.3f escalated %s launches %s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["config"].get("escalated_per_step"), d.get("gpu_launches")))
    for k, v in d["secondary"].items():
        print("%-10s value %.4g ms %.2f e2e ms %.1f frac %.3f cpu %.4g parity %s" % (k, v["value"], v["ms_per_step"], v["e2e"]["ms_per_step"], v["roofline"]["frac"], v.get("cpu_baseline", {}).get("value", float("nan")), v.get("parity_on_sample")))
    print("clocks", d.get("clocks"))
except Exception as e:
    print("no bench line:", e)
PY
echo "=== 2. reference arm ==="
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?"
tail -c 600 $O/bench_reference.json
echo "=== 3. sanitizers ==="
bash tools/sanitize_gpu.sh 2>&1 | tail -14
echo "=== 4. ncu --set full: double-precision rows, DMMA right-hand side, flag/low-row kernels ==="
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"nufft2_rows_kernel.*double2|nufft2_cols_kernel.*double2|nufft2_lowacc" -c 3 -o $O/r02_nufft_escalation python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $O/ncu_esc.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"rt_rhs_mma|rg_resolve" -c 2 -o $O/r02_regress_rhs_mma python tools/probe_others.py 0.125 regress > $O/ncu_rhs.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"nufft2_rows_kernel.*float2|nufft2_cols_kernel.*float2" -c 2 -o $O/r02_nufft_final python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $O/ncu_nufft.log 2>&1; echo "rc=$?"
echo "=== 5. launch list of the default bench command ==="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02_bench_default.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "rc=$?"
ls -la $O | head -30
echo "=== done ==="
