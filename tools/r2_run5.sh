#!/usr/bin/env bash
# Round-2 GPU run 5: tcgen05 Gram (K5t), flatten with the two-pass median, NUFFT kernel width 10, full suite.
set -u
O=gpurun_out/r2_run5
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. regression + flatten tests ==="
timeout 1200 python -m pytest tests -m gpu -q -rxXs -k "regress or flatten or config4 or corrector" > $O/pytest_sel.log 2>&1; echo "rc=$?"
tail -15 $O/pytest_sel.log
echo "=== 2. bench: flatten + regress legs ==="
timeout 1200 python bench.py --steps 10 --warmup 3 --legs flatten,regress > $O/bench_legs.json 2> $O/bench_legs.err; echo "rc=$?"
python - $O/bench_legs.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("LS c2: ms/step %.3f kernel_ms %.3f e2e %.3f ms frac %.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
    for k, v in d["secondary"].items():
        if "error" in v: print(k, "ERROR", v["error"]); continue
        print(k, "value %.4g %s  ms %.3f  e2e ms %.3f  roofline %.4g %s frac %.3f  kernel_ms %.2f cpu %s  parity %s" % (v["value"], v["unit"], v["ms_per_step"], v["e2e"]["ms_per_step"], v["roofline"]["achieved"], v["roofline"]["unit"], v["roofline"]["frac"], v["roofline"]["kernel_ms"], v.get("cpu_baseline", {}).get("value"), v.get("parity_on_sample")))
except Exception as e:
    print("no bench line:", e)
PY
tail -3 $O/bench_legs.err
echo "=== 3. NUFFT kernel width 10: worst bins + timing ==="
LKB_NUFFT_W=10 timeout 900 python tools/worst_bins.py > $O/worst_bins_w10.log 2>&1; echo "rc=$?"; grep -A3 '"nufft"' $O/worst_bins_w10.log | head -8
LKB_NUFFT_W=10 timeout 400 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $O/bench_w10.json 2> $O/bench_w10.err
LKB_NUFFT_W=12 timeout 900 python tools/worst_bins.py > $O/worst_bins_w12.log 2>&1; grep -A3 '"nufft"' $O/worst_bins_w12.log | head -4
python - $O/bench_w10.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("W=10: ms/step %.3f kernel_ms %.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"]))
except Exception as e:
    print("no bench line:", e)
PY
echo "=== 4. full GPU suite ==="
timeout 1800 python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu.log 2>&1; echo "rc=$?"
tail -12 $O/pytest_gpu.log
echo "=== 5. ncu: tcgen05 Gram + flatten2 ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rt_gram_kernel" -c 1 -o $O/r02_rt_gram python tools/probe_others.py 0.125 regress > $O/ncu_rt.log 2>&1
tail -2 $O/ncu_rt.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flatten2_kernel -c 1 -o $O/r02_flatten2_b python tools/probe_others.py 0.03 flatten > $O/ncu_flatten.log 2>&1
tail -2 $O/ncu_flatten.log
ls -la $O
echo "=== done ==="
