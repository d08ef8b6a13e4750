#!/usr/bin/env bash
# Round-2 GPU run 2: NUFFT v2 on hardware (tests, bench variants, ncu), worst-bin detail.
set -u
O=gpurun_out/r2_run2
mkdir -p $O
export PYTHONUNBUFFERED=1

echo "=== 1. NUFFT GPU tests (v2 default) ==="
timeout 900 python -m pytest tests/test_gpu_zz_nufft.py tests/test_gpu_engine.py -m gpu -q -rxXs -k "nufft or ls_" > $O/pytest_nufft.log 2>&1; echo "rc=$?"
tail -12 $O/pytest_nufft.log

echo "=== 2. bench: v2 (default) and variants ==="
for cfg in "v2" "v2 LKB_NUFFT_GROUP_MB=48" "v2 LKB_NUFFT_GROUP_MB=96" "global LKB_NUFFT_FFT=global"; do
  set -- $cfg
  tag=$1; shift
  tagf=$(echo "$cfg" | tr ' =' '__')
  env "$@" timeout 400 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $O/bench_$tagf.json 2> $O/bench_$tagf.err || echo "bench failed"
  python - $O/bench_$tagf.json "$cfg" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-40s ms/step %.3f  kernel_ms %.3f  e2e ms %.3f  launches %d family %s" % (sys.argv[2], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["config"]["kernel_family"]))
except Exception as e:
    print("no bench line:", e)
PY
done

echo "=== 3. worst-bin detail ==="
timeout 900 python tools/worst_bins_detail.py > $O/worst_bins_detail.log 2>&1; echo "rc=$?"
cat $O/worst_bins_detail.log

echo "=== 4. ragged probe (config-5 share x 1/4): v2 vs global passes vs direct ==="
timeout 300 python tools/probe_others.py 0.25 k1 2>&1 | tail -1
LKB_NUFFT_FFT=global timeout 300 python tools/probe_others.py 0.25 k1 2>&1 | tail -1

echo "=== 5. ncu: launch list + full captures of the v2 kernels ==="
B="python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_r02_bench_c2_nufft_v2.csv $B > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nufft2_(spread|cols|rows)_kernel" -c 3 -o $O/r02_nufft_v2 $B > $O/ncu_v2.log 2>&1
tail -3 $O/ncu_v2.log
ls -la $O
echo "=== done ==="
