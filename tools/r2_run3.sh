#!/usr/bin/env bash
# Round-2 GPU run 3: NUFFT v2 real-mode (one transform per light curve), plan cache, new bench legs.
set -u
O=gpurun_out/r2_run3
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. GPU test suite ==="
timeout 1800 python -m pytest tests -m gpu -q -rxXs > $O/pytest_gpu.log 2>&1; echo "rc=$?"
tail -25 $O/pytest_gpu.log
echo "=== 2. worst-bin sweep ==="
timeout 900 python tools/worst_bins.py > $O/worst_bins.log 2>&1; echo "rc=$?"; cat $O/worst_bins.log | tail -40
echo "=== 3. bench headline only, then with all legs ==="
timeout 400 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $O/bench_v2r.json 2> $O/bench_v2r.err; echo "rc=$?"
python - $O/bench_v2r.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("ms/step %.3f  kernel_ms %.3f  e2e ms %.3f  launches %d family %s roofline frac %.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["config"]["kernel_family"], d["roofline"]["frac"]))
except Exception as e:
    print("no bench line:", e)
PY
timeout 1500 python bench.py --steps 10 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"
tail -c 6000 $O/bench_full.json; tail -5 $O/bench_full.err
echo "=== 4. c5 on one GPU (16384 ragged light curves) ==="
timeout 900 python bench.py --workload c5 --steps 3 --warmup 2 > $O/bench_c5_n1.json 2> $O/bench_c5_n1.err; echo "rc=$?"
tail -c 2500 $O/bench_c5_n1.json; tail -5 $O/bench_c5_n1.err
echo "=== 5. ncu: launch list + full captures of the v2 kernels ==="
B="python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_r02_bench_c2_nufft_v2r.csv $B > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nufft2_(spread|cols|rows|lowrows)_kernel" -c 4 -o $O/r02_nufft_v2r $B > $O/ncu_v2.log 2>&1
tail -3 $O/ncu_v2.log
ls -la $O
echo "=== done ==="
